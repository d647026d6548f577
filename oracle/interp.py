"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's L2 -> grid stage.

Follows /root/reference/oisatgmi/interpolator.py (and the GOSAT gap filler,
filler_gosat.py) step by step with the same numpy/scipy calls the reference
makes, so that it can be pinned against the live reference in the build
container (tests/test_oracle_vs_reference.py) and then travel to the GPU box,
where /root/reference does not exist.  Parity status: PINNED (reference run
here; fixtures in tests/golden/ made by oracle/make_golden.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module; the product package never does.
"""
from __future__ import annotations

import numpy as np
from scipy import signal
from scipy.interpolate import (LinearNDInterpolator, NearestNDInterpolator,
                               RBFInterpolator)
from scipy.spatial import Delaunay, cKDTree

from oisatgmi_b200.config import kind_of, satellite_amf, satellite_opt


def _spacing(coords):
    """CTM spacing from the first two columns/rows (interpolator.py:119-120)."""
    lon, lat = coords["Longitude"], coords["Latitude"]
    return np.abs(lon[0, 0] - lon[0, 1]), np.abs(lat[0, 0] - lat[1, 0])


def _query_points(X, Y):
    # what scipy's _ndim_coords_from_arrays returns for a 2-tuple of meshgrids
    return np.stack(np.broadcast_arrays(X, Y), axis=-1).astype(np.float64)


def regrid(handle, Z, X, Y, kind, dists, threshold, reach=2.0):
    """One field onto the (X, Y) mesh.  interpolator.py:10-37 (reach = 2, i.e.
    NaN beyond 2*threshold from the nearest pixel); filler_gosat.py:11-32 uses
    reach = 1 and has no KD-tree mode."""
    zflat = np.asarray(Z).flatten()
    if kind == 1:
        out = LinearNDInterpolator(handle, zflat, fill_value=np.nan)((X, Y))
    elif kind == 2:
        out = NearestNDInterpolator(handle, zflat)((X, Y))
    elif kind == 3:
        q = np.stack([X.ravel(), Y.ravel()], -1)
        out = RBFInterpolator(handle, zflat, neighbors=5)(q).reshape(np.shape(X))
    elif kind == 4:
        _, idx = handle.query(np.column_stack((X.ravel(), Y.ravel())))
        out = zflat[idx].reshape(X.shape)
    else:
        raise Exception("other type of interpolation methods has not been implemented yet")
    out[dists > threshold * reach] = np.nan
    return out


def upscale(X, Y, Z, coords, grid_size, threshold, tri=None, error=False):
    """Box mean over floor(ctm spacing / grid_size) fine points and nearest
    sample at the CTM centres, or pass-through when the model is finer than the
    working grid (interpolator.py:48-97).  `error=True` divides by (kx*ky)^2:
    variance of a window mean (interpolator.py:44-46, 72-73)."""
    dlon, dlat = _spacing(coords)
    if not ((dlon >= grid_size) or (dlat >= grid_size)):
        return X, Y, Z, True
    kx = np.floor(dlon / grid_size)
    ky = np.floor(dlat / grid_size)
    kx = 1 if kx == 0 else kx
    ky = 1 if ky == 0 else ky
    norm = (kx * ky) ** 2 if error else (kx * ky)
    box = np.ones((int(ky), int(kx))) / norm
    Z = signal.convolve2d(Z, box, boundary="symm", mode="same")
    pts = np.column_stack((X.flatten(), Y.flatten())).astype(np.float64)
    tree = cKDTree(pts)
    dists, _ = tree.query(_query_points(coords["Longitude"], coords["Latitude"]))
    Z = regrid(tree, Z, coords["Longitude"], coords["Latitude"], 4, dists, threshold)
    return coords["Longitude"], coords["Latitude"], Z, False


def interpolator(interpolator_type, grid_size, sat_data, ctm_models_coordinate,
                 flag_thresh=0.75):
    """Restates interpolator.py:100-291."""
    coords = ctm_models_coordinate
    dlon, dlat = _spacing(coords)
    threshold_ctm = np.sqrt(dlon ** 2 + dlat ** 2)
    mask = np.multiply(sat_data.quality_flag > flag_thresh, 1.0).squeeze()
    mask[mask != 1.0] = np.nan
    pts = np.zeros((np.size(sat_data.latitude_center), 2))
    pts[:, 0] = sat_data.longitude_center.flatten()
    pts[:, 1] = sat_data.latitude_center.flatten()
    lat_all = coords["Latitude"].flatten()
    lon_all = coords["Longitude"].flatten()
    lon_axis = np.arange(lon_all.min(), lon_all.max() + grid_size, grid_size)
    lat_axis = np.arange(lat_all.min(), lat_all.max() + grid_size, grid_size)
    X, Y = np.meshgrid(lon_axis, lat_axis)
    tree = cKDTree(pts)
    dists, _ = tree.query(_query_points(X, Y))
    if interpolator_type < 3:
        try:
            handle = Delaunay(pts)
        except Exception:
            return None
    elif interpolator_type == 3:
        handle = pts
    else:
        handle = tree

    state = {}

    def grid(z, error=False):
        fine = regrid(handle, z, X, Y, interpolator_type, dists, grid_size)
        gx, gy, out, needed = upscale(X, Y, fine, coords, grid_size, threshold_ctm, error=error)
        state.update(x=gx, y=gy, needed=needed)
        return out

    def grid_levels(arr3, nlev):
        first = grid(arr3[0].squeeze() * mask)
        out = np.zeros((nlev,) + first.shape)
        out[0] = first
        for z in range(1, nlev):
            out[z] = grid(arr3[z].squeeze() * mask)
        return out

    vcd = grid(sat_data.vcd * mask)
    up_x, up_y, needed = state["x"], state["y"], state["needed"]
    if np.isnan(np.nanmean(vcd.flatten())):
        return None
    kind = kind_of(sat_data)
    if kind == "amf":
        amf = grid(sat_data.amf * mask)
    if np.size(sat_data.tropopause) != 1:
        tropopause = grid(sat_data.tropopause * mask)
    else:
        tropopause = np.empty((1))
    # sigma is squared in the INPUT dtype (float16 for the OMI/TROPOMI readers)
    uncertainty = np.sqrt(grid(sat_data.uncertainty ** 2 * mask, error=True))

    if kind == "amf":
        nlev = np.shape(sat_data.pressure_mid)[0]
        if np.size(sat_data.scattering_weights) != 1:
            sw = grid_levels(sat_data.scattering_weights, nlev)
            pmid = grid_levels(sat_data.pressure_mid, nlev)
        else:
            sw = np.empty((1))
            pmid = np.zeros((nlev,) + np.shape(up_x))
        return satellite_amf(vcd, amf, sat_data.time, tropopause, up_y, up_x, [], [],
                             uncertainty, [], pmid, sw, needed, [], [], [], [])

    # optimal-estimation products (MOPITT, GOSAT): interpolator.py:216-287
    nlev = np.shape(sat_data.pressure_mid)[0]
    extras = {}
    for name in ("aprior_column", "surface_pressure", "apriori_surface"):
        src = getattr(sat_data, name)
        # the reference tests `.any()`; for GOSAT these are uninitialised
        # size-1 arrays (filler_gosat.py:198-200) and the result is garbage
        # that nothing downstream reads -> "don't care", kept as size-1 NaN.
        if np.size(src) != 1 and src.any():
            extras[name] = grid(src * mask)
        else:
            extras[name] = np.full((1,), np.nan)
    x_col = grid(sat_data.x_col * mask)
    if sat_data.sensor == "MOPITT":
        aks = grid_levels(sat_data.averaging_kernels, nlev + 1)
        pw = np.empty((1))
    elif sat_data.sensor == "GOSAT":
        aks = grid_levels(sat_data.averaging_kernels, nlev)
        pw = grid_levels(sat_data.pressure_weight, nlev)
    pmid = grid_levels(sat_data.pressure_mid, nlev)
    ap_prof = grid_levels(sat_data.apriori_profile, nlev)
    return satellite_opt(vcd, sat_data.time, [], tropopause, up_y, up_x, [], [],
                         uncertainty, [], pmid, aks, needed, [], [], [],
                         extras["aprior_column"], ap_prof, extras["surface_pressure"],
                         extras["apriori_surface"], x_col, pw, sat_data.sensor)


def filler_gosatxch4(grid_size, sat_data, flag_thresh=0.75):
    """Sparse soundings -> global image (filler_gosat.py:87-201).  The filler's
    fake 0.1 degree float16 'model' grid (filler_gosat.py:124-129) has a
    float16 spacing of 0.125 < grid_size, so its _upscaler always takes the
    pass-through branch: the output lives on the float16 fine mesh."""
    mask = np.multiply(sat_data.quality_flag > flag_thresh, 1.0).squeeze()
    nearest_mask = mask  # alias, as in filler_gosat.py:102
    mask[mask != 1.0] = np.nan
    pts = np.zeros((np.size(sat_data.latitude_center), 2))
    pts[:, 0] = sat_data.longitude_center.flatten()
    pts[:, 1] = sat_data.latitude_center.flatten()
    try:
        tri = Delaunay(pts)
    except Exception:
        return None
    lon_axis = np.arange(-180.0, 180.0 + grid_size, grid_size)
    lat_axis = np.arange(-90.0, 90.0 + grid_size, grid_size)
    X, Y = np.meshgrid(lon_axis.astype("float16"), lat_axis.astype("float16"))
    fx, fy = np.meshgrid(np.arange(-180.0, 181.0, 0.1).astype("float16"),
                         np.arange(-90.0, 91.0, 0.1).astype("float16"))
    fake = {"Latitude": fy, "Longitude": fx}
    dlon, dlat = _spacing(fake)
    assert not ((dlon >= grid_size) or (dlat >= grid_size))  # pass-through branch
    tree = cKDTree(pts)
    dists, _ = tree.query(_query_points(X, Y))

    def grid(z, kind=1):
        return regrid(tri, z, X, Y, kind, dists, grid_size, reach=1.0)

    vcd = grid(sat_data.x_col * mask)
    xch4 = grid(sat_data.x_col * mask)
    qflag = grid(nearest_mask, kind=2)
    uncertainty = np.sqrt(grid(sat_data.uncertainty ** 2 * mask))
    nlev = np.shape(sat_data.pressure_mid)[0]

    def levels(arr):
        out = np.zeros((nlev,) + X.shape)
        for z in range(nlev):
            out[z] = grid(arr[z, :].squeeze() * mask)
        return out

    aks = levels(sat_data.averaging_kernels)
    pmid = levels(sat_data.pressure_mid)
    ap_prof = levels(sat_data.apriori_profile)
    pw = levels(sat_data.pressure_weight)
    return satellite_opt(vcd, sat_data.time, [], np.empty((1)), Y, X, [], [], uncertainty,
                         qflag, pmid, aks, [], [], [], [], np.empty((1)), ap_prof,
                         np.empty((1)), np.empty((1)), xch4, pw, "GOSAT")
