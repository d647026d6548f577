"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SSMIS water-vapour path
(SURVEY.md section 8f-4):

  ssmis_wv           /root/reference/oisatgmi/reader.py:1277-1297 (ssmis_reader_wv up to its
                     interpolator call; file access answered from a dictionary)
  interpolator_ssmis /root/reference/oisatgmi/interpolator_ssmis.py:96-168, with ITS OWN
                     _interpolosis (:12-33, reach = grid_size, not doubled) and _upscaler
                     (:43-94: box mean on the working mesh, then a SECOND Delaunay-linear
                     interpolation from the mesh to the model cell centres -- not the nearest
                     node sampling of interpolator.py)
  pwv_calculator     /root/reference/oisatgmi/pwv_cal.py:7-101 (model precipitable water)

Same numpy / scipy calls as the reference, in the same order.  Quirks kept: the working mesh
axes are rounded to float16 (:137), no quality mask, the uncertainty is NOT squared before the
(kx*ky)^2 kernel and no square root follows (:152-155), `tri` of pwv_calculator is built and
never used by the `_upscaler` it is handed to (pwv_cal.py:31-34,85-86).

Parity status: PINNED -- tests/test_oracle_vs_reference.py runs the unmodified functions on the
same seeded case and compares bit for bit; fixture tests/golden/ssmis_pwv.npz.  Only tests/,
smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import datetime

import numpy as np
from scipy import signal
from scipy.interpolate import LinearNDInterpolator, NearestNDInterpolator, RBFInterpolator
from scipy.spatial import Delaunay, cKDTree

from oisatgmi_b200.config import satellite_ssmis
from oracle import interp as _interp


def ssmis_wv(v, yyyymm):
    """reader.py:1280-1297; `v`: 'latitude', 'longitude' (1-D axes, 0..360) and
    'atmosphere_water_vapor_content' (scaled bytes of the monthly map); `yyyymm` is what the
    reader parses from the file name (:1280-1283)."""
    time = datetime.datetime(int(yyyymm[0:4]), int(yyyymm[4:6]), 1)
    lat = v["latitude"].astype("float32")
    lon = v["longitude"].astype("float32")
    lon[lon > 180.0] = lon[lon > 180.0] - 360.0
    lon, lat = np.meshgrid(lon, lat)
    pwv = np.array(v["atmosphere_water_vapor_content"]).astype("float32")
    pwv[pwv > 250.0] = np.nan
    pwv = pwv * 0.3
    pwv[np.where((pwv >= 75.0) | (np.isinf(pwv)))] = np.nan
    return satellite_ssmis(pwv, pwv * 0.05, time, lat, lon, False, [], "SSMI")


def _regrid(handle, Z, X, Y, kind, dists, threshold):
    """interpolator_ssmis.py:12-33."""
    z = np.asarray(Z).flatten()
    if kind == 1:
        out = LinearNDInterpolator(handle, z, fill_value=np.nan)((X, Y))
        out[dists > threshold] = np.nan
    elif kind == 2:
        out = NearestNDInterpolator(handle, z)((X, Y))
        out[dists > threshold] = np.nan
    elif kind == 3:
        q = np.stack([X.ravel(), Y.ravel()], -1)
        out = RBFInterpolator(handle, z, neighbors=5)(q).reshape(np.shape(X))
        out[dists > threshold * 3.0] = np.nan
    else:
        raise Exception("other type of interpolation methods has not been implemented yet")
    return out


def _upscale(X, Y, Z, coords, grid_size, threshold, error=False):
    """interpolator_ssmis.py:43-94."""
    clat, clon = coords["Latitude"], coords["Longitude"]
    dlon = np.abs(clon[0, 0] - clon[0, 1])
    dlat = np.abs(clat[0, 0] - clat[1, 0])
    if not ((dlon >= grid_size) or (dlat >= grid_size)):
        return X, Y, Z, True
    kx = np.floor(dlon / grid_size)
    ky = np.floor(dlat / grid_size)
    kx = 1 if kx == 0 else kx
    ky = 1 if ky == 0 else ky
    norm = (ky * kx) ** 2 if error else (ky * kx)
    Z = signal.convolve2d(Z, np.ones((int(ky), int(kx))) / norm, boundary="symm", mode="same")
    pts = np.zeros((np.size(X), 2))
    pts[:, 0] = X.flatten()
    pts[:, 1] = Y.flatten()
    tri = Delaunay(pts)
    dists, _ = cKDTree(pts).query(_interp._query_points(clon, clat))
    return clon, clat, _regrid(tri, Z, clon, clat, 1, dists, threshold), False


def interpolator_ssmis(interpolator_type, grid_size, sat_data, ctm_models_coordinate):
    """interpolator_ssmis.py:96-168."""
    clat, clon = ctm_models_coordinate["Latitude"], ctm_models_coordinate["Longitude"]
    dlon = np.abs(clon[0, 0] - clon[0, 1])
    dlat = np.abs(clat[0, 0] - clat[1, 0])
    threshold_ctm = np.sqrt(dlon ** 2 + dlat ** 2)
    pts = np.zeros((np.size(sat_data.latitude_center), 2))
    pts[:, 0] = sat_data.longitude_center.flatten()
    pts[:, 1] = sat_data.latitude_center.flatten()
    try:
        tri = Delaunay(pts)
    except Exception:
        return None
    lon_grid = np.arange(np.min(clon.flatten()), np.max(clon.flatten()) + grid_size, grid_size)
    lat_grid = np.arange(np.min(clat.flatten()), np.max(clat.flatten()) + grid_size, grid_size)
    X, Y = np.meshgrid(lon_grid.astype("float16"), lat_grid.astype("float16"))
    grid = np.zeros((2,) + np.shape(X))
    grid[0], grid[1] = X, Y
    dists, _ = cKDTree(pts).query(_interp._query_points(grid[0], grid[1]))
    handle = pts if interpolator_type == 3 else tri
    up_x, up_y, vcd, needed = _upscale(
        X, Y, _regrid(handle, sat_data.vcd, X, Y, interpolator_type, dists, grid_size),
        ctm_models_coordinate, grid_size, threshold_ctm)
    _, _, err, _ = _upscale(
        X, Y, _regrid(handle, sat_data.uncertainty, X, Y, interpolator_type, dists, grid_size),
        ctm_models_coordinate, grid_size, threshold_ctm, error=True)
    return satellite_ssmis(vcd, err, sat_data.time, up_y, up_x, needed, [], "SSMIS")


def pwv_calculator(ctm_data, sat_data):
    """pwv_cal.py:7-101: precipitable water of the model, kg m-2 / 1000, on the satellite's
    grid cells that hold a retrieval."""
    stamps = []
    for c in ctm_data:
        for t in c.time:
            stamps.append(t.year * 10000 + t.month * 100 + t.day + t.hour / 24.0
                          + t.minute / 60.0 / 24.0 + t.second / 3600.0 / 24.0)
    stamps = np.array(stamps)
    g = 9.80665
    for k, gr in enumerate(sat_data):
        if gr is None:
            continue
        t = gr.time
        day = 0
        if ctm_data[0].averaged == False:  # noqa: E712
            day = int(np.floor(np.argmin(np.abs(t.year * 10000 + t.month * 100 + t.day - stamps))))
        c = ctm_data[day]
        if c.ctmtype in ("ECCOH", "FREE"):
            dp = c.delta_p[:, :, :].squeeze()
            prof = c.gas_profile[:, :, :].squeeze()
        elif c.ctmtype == "GMI":
            prof = np.nanmean(c.gas_profile[:, :, :, :], axis=0).squeeze()
            dp = np.nanmean(c.delta_p[:, :, :, :], axis=0).squeeze()
        pc = dp * prof / g / 10000.0
        if gr.ctm_upscaled_needed == True:  # noqa: E712
            new = np.zeros((np.shape(dp)[0],) + np.shape(gr.longitude_center)) * np.nan
            sat = {"Longitude": gr.longitude_center, "Latitude": gr.latitude_center}
            ds_lon = np.abs(sat["Longitude"][0, 0] - sat["Longitude"][0, 1])
            ds_lat = np.abs(sat["Latitude"][0, 0] - sat["Latitude"][1, 0])
            thr = np.sqrt(ds_lon ** 2 + ds_lat ** 2)
            clon, clat = ctm_data[0].longitude, ctm_data[0].latitude
            gs = np.sqrt(np.abs(clon[0, 0] - clon[0, 1]) ** 2 + np.abs(clat[0, 0] - clat[1, 0]) ** 2)
            for z in range(np.shape(prof)[0]):
                _, _, new[z], _ = _interp.upscale(clon, clat, pc[z], sat, gs, thr)
            pc = new
        pwv = np.nansum(pc / 1000.0, axis=0).squeeze()
        pwv[np.isnan(gr.vcd)] = np.nan
        pwv[np.isinf(gr.vcd)] = np.nan
        sat_data[k].ctm_vcd = pwv
    return sat_data
