"""Drop-in for /root/reference/oisatgmi/interpolator_ssmis.py:
`interpolator_ssmis(interpolator_type, grid_size, sat_data, ctm_models_coordinate)`.

The SSMIS path grids the monthly water-vapour map in two Delaunay-linear steps
(interpolator_ssmis.py:96-168): map -> working mesh (axes rounded to float16, :137; NaN beyond
`grid_size` of the nearest map point, not 2 x, :21-22,150), a box mean on the mesh, then mesh ->
model cell centres with the module's own `_upscaler` (:43-94; NaN beyond the model's cell
diagonal).  No quality mask; the uncertainty goes through the (kx*ky)^2 kernel unsquared and no
square root follows (:152-155).  Both steps are geometry plans here (the inputs are lattices, so
the plans are built by Qhull + scipy's walk, exact, and cached by geometry: every month of a
record shares them); K2 applies them, K6 is the box mean.  Only interpolator type 1 is built --
the reader uses no other (reader.py:1302-1303).
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, plan as _plan
from .config import satellite_ssmis
from .interpolator import _FieldSpec, apply_plan

__all__ = ["interpolator_ssmis"]

_fake = None


def _fake_coords():
    """A 0.1 degree float16 lattice: its float16 spacing (0.125) is below any working mesh this
    module meets, so GridPlan keeps the output on the mesh handed to it (no box window)."""
    global _fake
    if _fake is None:
        fx, fy = np.meshgrid(np.arange(-180.0, 181.0, 0.1).astype("float16"),
                             np.arange(-90.0, 91.0, 0.1).astype("float16"))
        _fake = {"Latitude": fy, "Longitude": fx}
    return _fake


def interpolator_ssmis(interpolator_type: int, grid_size: float, sat_data, ctm_models_coordinate: dict):
    if interpolator_type != 1:
        raise Exception("other type of interpolation methods has not been implemented yet")
    _dev.require_cuda()
    L = _lib.lib()
    clat = np.asarray(ctm_models_coordinate["Latitude"])
    clon = np.asarray(ctm_models_coordinate["Longitude"])
    dlon, dlat = _plan.grid_spacing(ctm_models_coordinate)
    threshold_ctm = np.sqrt(dlon ** 2 + dlat ** 2)
    lon_axis = np.arange(np.min(clon), np.max(clon) + grid_size, grid_size).astype("float16")
    lat_axis = np.arange(np.min(clat), np.max(clat) + grid_size, grid_size).astype("float16")
    gpl = _plan.grid_plan(_fake_coords(), grid_size, mesh=(lon_axis, lat_axis))
    if gpl.upscale:
        raise NotImplementedError("grid_size <= 0.125 degree is outside this module's design")
    lat, lon = np.asarray(sat_data.latitude_center), np.asarray(sat_data.longitude_center)
    gp = _plan.granule_plan(lon, lat, gpl, radius=grid_size)          # map -> mesh, reach 1
    if gp is None:
        return None
    H, W = gpl.out_shape
    n_mesh = H * W
    specs = [_FieldSpec("vcd", sat_data.vcd), _FieldSpec("uncertainty", sat_data.uncertainty)]
    mesh, layout, _keep = apply_plan(gp, specs, None, lat.size, n_mesh)    # (2, n_mesh) float64
    X, Y = np.meshgrid(lon_axis, lat_axis)
    if not ((dlon >= grid_size) or (dlat >= grid_size)):
        host = _dev.to_host(mesh)
        return satellite_ssmis(host[0].reshape(H, W), host[1].reshape(H, W), sat_data.time, Y, X,
                               True, [], "SSMIS")
    # box mean on the mesh (convolve2d, boundary='symm', mode='same'), :60-73
    ky, kx = _plan.box_extent(dlon, dlat, grid_size)
    ident = _dev.to_device(np.arange(n_mesh, dtype=np.int32))
    ones = _dev.full((n_mesh,), 1, "uint8")
    smooth = _dev.empty((2, n_mesh))
    for row, norm in ((0, kx * ky), (1, (kx * ky) ** 2)):
        _lib.check(L.oisat_grid_resample(mesh[row].data_ptr(), None, _lib.SRC_VALUE, _lib.F64, 1,
                                         H, W, ky, kx, 1.0 / norm, ident.data_ptr(), ones.data_ptr(),
                                         n_mesh, smooth[row].data_ptr(), n_mesh, _dev.stream()))
    # mesh -> model cell centres, Delaunay-linear on the mesh lattice, NaN beyond the cell diagonal
    if not (np.array_equal(clon, np.broadcast_to(clon[0:1, :], clon.shape))
            and np.array_equal(clat, np.broadcast_to(clat[:, 0:1], clat.shape))):
        raise _lib.OisatError("interpolator_ssmis needs a rectilinear model grid")
    gpl2 = _plan.grid_plan(_fake_coords(), 1.0, mesh=(clon[0, :], clat[:, 0]))
    gp2 = _plan.granule_plan(X, Y, gpl2, radius=threshold_ctm)
    if gp2 is None:
        return None
    n_out = clat.size
    out, _, _keep2 = apply_plan(gp2, [_FieldSpec("vcd", smooth[0]), _FieldSpec("uncertainty", smooth[1])],
                                None, n_mesh, n_out)
    host = _dev.to_host(out)
    return satellite_ssmis(host[0].reshape(clat.shape), host[1].reshape(clat.shape), sat_data.time,
                           clat, clon, False, [], "SSMIS")
