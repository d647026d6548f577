"""Drop-in for /root/reference/oisatgmi/interpolator.py -- same names, same
arguments, same return records, GPU execution.

    interpolator(interpolator_type, grid_size, sat_data, ctm_models_coordinate, flag_thresh=0.75)
    _upscaler(X, Y, Z, ctm_models_coordinate, grid_size, threshold, tri=None, error=False)

What runs where
  host   geometry plan (oisatgmi_b200/plan.py): Qhull triangulation + scipy's
         directed walk, as the reference does (interpolator.py:153, 13-15), but
         ONCE per granule instead of once per field, and only for mesh nodes
         that K0 keeps;
  GPU    K0 proximity predicate (interpolator.py:145-150,16), quality mask
         (:126-128) and K2: every field and every level of the granule through
         the stencil in a single launch (:162-283).
There is no CPU fallback: without liboisat.so / a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib, plan as _plan
from .config import kind_of, satellite_amf, satellite_opt

__all__ = ["interpolator", "_upscaler"]


class _FieldSpec:
    __slots__ = ("name", "array", "nlev", "error")

    def __init__(self, name, array, nlev=1, error=False):
        self.name, self.array, self.nlev, self.error = name, array, nlev, error


def _flat_levels(arr, nlev, n_px):
    """(nlev, ...) reader array -> contiguous (nlev, n_px) in its native float dtype."""
    a = _dev.native_float(arr)
    return np.ascontiguousarray(a).reshape(nlev, n_px)


def apply_plan(gp: "_plan.GranulePlan", specs, good_dev, n_px, n_out):
    """Run K2 for a list of _FieldSpec; returns (device tensor [rows, n_out]
    pre-filled with NaN, {name: (row0, nlev)})."""
    L = _lib.lib()
    rows = sum(s.nlev for s in specs)
    out = _dev.full((rows, n_out), float("nan"))
    cells_d, vert_d, w_d = gp.dev()
    nwin = gp.nwin
    keep_alive = []
    layout = {}
    row0 = 0
    # the C-ABI takes up to 16 field descriptors per call
    for start in range(0, len(specs), 16):
        chunk = specs[start:start + 16]
        farr = (_lib.Field * len(chunk))()
        chunk_row0 = row0
        for i, s in enumerate(chunk):
            if hasattr(s.array, "data_ptr"):     # already on the device ([nlev * n_px], native dtype)
                d = s.array
            else:
                d = _dev.to_device(_flat_levels(s.array, s.nlev, n_px))
            keep_alive.append(d)
            farr[i].data = d.data_ptr()
            farr[i].dtype = _dev.dtype_code(d)
            farr[i].op = _lib.OP_SQUARE_NATIVE if s.error else _lib.OP_NONE
            farr[i].post = _lib.POST_SQRT if s.error else _lib.POST_NONE
            farr[i].nlev = s.nlev
            farr[i].lev_stride = n_px
            farr[i].box_weight = (1.0 / (nwin * nwin)) if s.error else (1.0 / nwin)
            layout[s.name] = (row0, s.nlev)
            row0 += s.nlev
        out_view = out[chunk_row0:]
        _lib.check(L.oisat_interp_apply(vert_d.data_ptr(), w_d.data_ptr(), nwin, gp.n_cells,
                                        _dev.ptr(good_dev), farr, len(chunk),
                                        out_view.data_ptr(), n_out, cells_d.data_ptr(),
                                        _dev.stream()))
    return out, layout, keep_alive


def quality_mask(qflag, thresh):
    """good[p] = quality_flag[p] > thresh on the device (interpolator.py:126-128)."""
    L = _lib.lib()
    q = np.ascontiguousarray(_dev.native_float(np.asarray(qflag)).squeeze()).ravel()
    qd = _dev.to_device(q)
    good = _dev.empty((q.size,), "uint8")
    _lib.check(L.oisat_quality_mask(qd.data_ptr(), _dev.dtype_code(q), q.size, float(thresh),
                                    good.data_ptr(), _dev.stream()))
    return good


def field_specs(sat_data):
    """The fields the reference grids, in its order (interpolator.py:162-283)."""
    kind = kind_of(sat_data)
    specs = [_FieldSpec("vcd", sat_data.vcd)]
    if kind == "amf":
        specs.append(_FieldSpec("amf", sat_data.amf))
    if np.size(sat_data.tropopause) != 1:
        specs.append(_FieldSpec("tropopause", sat_data.tropopause))
    specs.append(_FieldSpec("uncertainty", sat_data.uncertainty, error=True))
    nlev = np.shape(sat_data.pressure_mid)[0]
    if kind == "amf":
        if np.size(sat_data.scattering_weights) != 1:
            specs.append(_FieldSpec("scattering_weights", sat_data.scattering_weights, nlev))
            specs.append(_FieldSpec("pressure_mid", sat_data.pressure_mid, nlev))
    else:
        for name in ("aprior_column", "surface_pressure", "apriori_surface"):
            src = getattr(sat_data, name)
            # size-1 placeholders (GOSAT, filler_gosat.py:198-200) are uninitialised
            # memory in the reference; the result is unused garbage there and a
            # size-1 NaN here
            if np.size(src) != 1 and np.asarray(src).any():
                specs.append(_FieldSpec(name, src))
        specs.append(_FieldSpec("x_col", sat_data.x_col))
        if sat_data.sensor == "MOPITT":
            specs.append(_FieldSpec("averaging_kernels", sat_data.averaging_kernels, nlev + 1))
        elif sat_data.sensor == "GOSAT":
            specs.append(_FieldSpec("averaging_kernels", sat_data.averaging_kernels, nlev))
            specs.append(_FieldSpec("pressure_weight", sat_data.pressure_weight, nlev))
        specs.append(_FieldSpec("pressure_mid", sat_data.pressure_mid, nlev))
        specs.append(_FieldSpec("apriori_profile", sat_data.apriori_profile, nlev))
    return specs


def interpolator(interpolator_type: int, grid_size: float, sat_data, ctm_models_coordinate: dict,
                 flag_thresh=0.75):
    """Grid one L2/L3 granule onto the model grid (or onto the working mesh when
    the model is finer than `grid_size`).  Mirrors interpolator.py:100-291:
    returns a new satellite_amf / satellite_opt, or None when the pixel centres
    cannot be triangulated or nothing of the granule falls on the grid."""
    if interpolator_type not in (1, 2, 4):
        # type 3 (RBFInterpolator, interpolator.py:21-27) is not used by any reader
        #  and not built; anything else raises there too (:34-36)
        raise Exception("other type of interpolation methods has not been implemented yet")
    _dev.require_cuda()
    kind = kind_of(sat_data)
    gpl = _plan.grid_plan(ctm_models_coordinate, grid_size)
    lat = np.asarray(sat_data.latitude_center)
    lon = np.asarray(sat_data.longitude_center)
    n_px = lat.size
    if interpolator_type == 1:
        gp = _plan.granule_plan(lon, lat, gpl, radius=grid_size * 2.0)
    else:
        # nearest pixel: NearestNDInterpolator over the Delaunay points (type 2; the
        # granule is skipped when the triangulation fails, :151-155) or the KD-tree
        # query itself (type 4) -- the same answer
        if interpolator_type == 2 and not _plan.triangulable(lon, lat):
            return None
        gp = _plan.nearest_plan(lon, lat, gpl, radius=grid_size * 2.0)
    if gp is None:
        return None
    n_out = int(np.prod(gpl.out_shape))
    if gp.n_cells == 0:
        return None  # every node is NaN -> interpolator.py:165-167
    good = quality_mask(sat_data.quality_flag, flag_thresh)
    specs = field_specs(sat_data)
    out, layout, _keep = apply_plan(gp, specs, good, n_px, n_out)
    host = _dev.to_host(out)

    def take(name):
        if name not in layout:
            return None
        r0, nl = layout[name]
        blk = host[r0:r0 + nl].reshape((nl,) + tuple(gpl.out_shape))
        return blk[0] if (nl == 1 and name not in _LEVELLED) else blk

    vcd = take("vcd")
    if np.isnan(vcd).all():
        return None  # interpolator.py:165-167
    if gpl.upscale:
        up_x, up_y, needed = gpl.ctm_lon, gpl.ctm_lat, False
    else:
        up_x, up_y = gpl.mesh()
        needed = True
    trop = take("tropopause")
    if trop is None:
        trop = np.empty((1))
    unc = take("uncertainty")
    nlev = np.shape(sat_data.pressure_mid)[0]
    if kind == "amf":
        sw = take("scattering_weights")
        if sw is None:
            sw = np.empty((1))
            pmid = np.zeros((nlev,) + tuple(gpl.out_shape))
        else:
            pmid = take("pressure_mid")
        res = satellite_amf(vcd, take("amf"), sat_data.time, trop, up_y, up_x, [], [], unc, [],
                            pmid, sw, needed, [], [], [], [])
    else:
        def opt(name):
            v = take(name)
            return np.full((1,), np.nan) if v is None else v
        pw = take("pressure_weight")
        res = satellite_opt(vcd, sat_data.time, [], trop, up_y, up_x, [], [], unc, [],
                            take("pressure_mid"), take("averaging_kernels"), needed, [], [], [],
                            opt("aprior_column"), take("apriori_profile"),
                            opt("surface_pressure"), opt("apriori_surface"), take("x_col"),
                            np.empty((1)) if pw is None else pw, sat_data.sensor)
    return res


_LEVELLED = {"scattering_weights", "pressure_mid", "averaging_kernels", "pressure_weight",
             "apriori_profile"}


def _upscaler(X, Y, Z, ctm_models_coordinate: dict, grid_size: float, threshold: float, tri=None,
              error=False):
    """Box mean + nearest-node sampling of a gridded field onto another grid,
    or pass-through (interpolator.py:48-97).  Imported by the vertical-operator
    modules exactly as in the reference; `tri` is accepted and ignored, as there."""
    _dev.require_cuda()
    L = _lib.lib()
    coords = ctm_models_coordinate
    dlon, dlat = _plan.grid_spacing(coords)
    if not ((dlon >= grid_size) or (dlat >= grid_size)):
        return X, Y, Z, True
    ky, kx = _plan.box_extent(dlon, dlat, grid_size)
    d, idx = _plan.nearest_node_table(X, Y, coords["Longitude"], coords["Latitude"])
    ok = ~(d > threshold * 2.0)
    Zc = np.asarray(Z)
    if Zc.dtype not in (np.float32, np.float64):
        Zc = Zc.astype(np.float64)
    H, W = Zc.shape
    src = _dev.to_device(Zc)
    nn = _dev.to_device(idx.astype(np.int32))
    okd = _dev.to_device(ok.astype(np.uint8))
    out = _dev.empty((idx.size,))
    norm = (kx * ky) ** 2 if error else (kx * ky)
    _lib.check(L.oisat_grid_resample(src.data_ptr(), None, _lib.SRC_VALUE, _dev.dtype_code(Zc),
                                     1, H, W, ky, kx,
                                     1.0 / norm, nn.data_ptr(), okd.data_ptr(), idx.size,
                                     out.data_ptr(), idx.size, _dev.stream()))
    res = _dev.to_host(out).reshape(np.shape(coords["Latitude"]))
    return coords["Longitude"], coords["Latitude"], res, False
