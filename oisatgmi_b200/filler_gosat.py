"""Drop-in for /root/reference/oisatgmi/filler_gosat.py:
`filler_gosatxch4(grid_size, sat_data, flag_thresh=0.75)`.

Sparse GOSAT soundings -> a global `grid_size` degree image by Delaunay-linear
interpolation, NaN further than `grid_size` (NOT doubled, filler_gosat.py:17)
from the nearest sounding.  The filler's private `_upscaler` is handed a fake
0.1 degree float16 "model" mesh whose float16 spacing (0.125) is always below
`grid_size`, so it is a pass-through (filler_gosat.py:55,82-85, 122-130): the
result lives on the float16-rounded working mesh.  Same plan machinery as
`interpolator`, reach 1 instead of 2, plus one nearest-neighbour field (the
quality mask, filler_gosat.py:153-155).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import NearestNDInterpolator

from . import _dev, plan as _plan
from .config import satellite_opt
from .interpolator import _FieldSpec, apply_plan, quality_mask

__all__ = ["filler_gosatxch4"]


def filler_gosatxch4(grid_size: float, sat_data, flag_thresh=0.75):
    _dev.require_cuda()
    lon_axis = np.arange(-180.0, 180.0 + grid_size, grid_size).astype("float16")
    lat_axis = np.arange(-90.0, 90.0 + grid_size, grid_size).astype("float16")
    fx, fy = np.meshgrid(np.arange(-180.0, 181.0, 0.1).astype("float16"),
                         np.arange(-90.0, 91.0, 0.1).astype("float16"))
    fake = {"Latitude": fy, "Longitude": fx}
    gpl = _plan.grid_plan(fake, grid_size, mesh=(lon_axis, lat_axis))
    if gpl.upscale:
        raise NotImplementedError("grid_size <= 0.125 degree is outside the filler's design")
    lat = np.asarray(sat_data.latitude_center)
    lon = np.asarray(sat_data.longitude_center)
    n_px = lat.size
    gp = _plan.granule_plan(lon, lat, gpl, radius=grid_size)
    if gp is None:
        return None
    shape = gpl.out_shape
    n_out = shape[0] * shape[1]
    good = quality_mask(sat_data.quality_flag, flag_thresh)
    nlev = np.shape(sat_data.pressure_mid)[0]
    specs = [_FieldSpec("x_col", sat_data.x_col),
             _FieldSpec("uncertainty", sat_data.uncertainty, error=True),
             _FieldSpec("averaging_kernels", sat_data.averaging_kernels, nlev),
             _FieldSpec("pressure_mid", sat_data.pressure_mid, nlev),
             _FieldSpec("apriori_profile", sat_data.apriori_profile, nlev),
             _FieldSpec("pressure_weight", sat_data.pressure_weight, nlev)]
    out, layout, _keep = apply_plan(gp, specs, good, n_px, n_out)
    host = _dev.to_host(out)

    def take(name, levelled=False):
        r0, nl = layout[name]
        blk = host[r0:r0 + nl].reshape((nl,) + tuple(shape))
        return blk if levelled else blk[0].copy()

    X, Y = np.meshgrid(lon_axis, lat_axis)
    # quality flag: nearest-neighbour field of the 1/NaN mask, NaN beyond reach
    # (filler_gosat.py:100-103,153-155) -- a host lookup on 65k nodes
    mask = np.multiply(np.asarray(sat_data.quality_flag) > flag_thresh, 1.0).squeeze()
    mask[mask != 1.0] = np.nan
    pts = np.column_stack((lon.flatten().astype(np.float64), lat.flatten().astype(np.float64)))
    qflag = NearestNDInterpolator(pts, mask.flatten())((X, Y))
    qflag = np.where(gp.keep.reshape(shape), qflag, np.nan)
    xch4 = take("x_col")
    return satellite_opt(xch4.copy(), sat_data.time, [], np.empty((1)), Y, X, [], [],
                         take("uncertainty"), qflag, take("pressure_mid", True),
                         take("averaging_kernels", True), [], [], [], [], np.empty((1)),
                         take("apriori_profile", True), np.empty((1)), np.empty((1)), xch4,
                         take("pressure_weight", True), 'GOSAT')
