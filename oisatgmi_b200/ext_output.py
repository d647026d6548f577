"""GEOS ExtData form of the monthly scaling factors: what the reference's
tools/convert2EXT.py:32-78 derives from the diagnostics `driver.write_to_nc` wrote
 -- `SF(time, lat, lon)` float64 with 1-D `lat` / `lon` axes and one
time stamp per month -- plus the all-ones months it fabricates before the record starts
(:80-124).  The data side lives here (`ext_fields`, `ones_fields`); the NetCDF container is
file I/O and is written only where `netCDF4` is importable (`write_ext`).

`ext_fields` takes the float32 variables of `driver.oisatgmi.output_fields()` /
`MonthPipeline.output_fields()` (K8), so a month can go from the device to its ExtData file
without the round trip through the diagnostics file.
"""
from __future__ import annotations

import datetime

import numpy as np

GLOBAL_ATTRIBUTES = {
    "Source": "OI-SAT-GMI tool (https://doi.org/10.5281/zenodo.7757427)",
    "Version": "0.0.7",
    "Institution": "NASA GSFC Code 614",
    "Contact": "Amir Souri (a.souri@nasa.gov or ahsouri@gmail.com)",
}


def _common(lat, lon, year, month):
    lat, lon = np.asarray(lat), np.asarray(lon)
    t0 = datetime.datetime(int(year), int(month), 1)
    return {
        "time": np.array([0.0]),                                       # hours since t0
        "time_units": "hours since " + t0.strftime("%Y-%m-%d %H:%M:%S"),
        "lat": lat[:, 0].squeeze().astype(np.float64),                  # convert2EXT.py:65
        "lon": lon[0, :].squeeze().astype(np.float64),                  # :66
    }


def ext_fields(fields, yyyymm):
    """`fields`: mapping with 'lat', 'lon' (2-D) and 'scaling_factor' as stored by
    write_to_nc (float32); `yyyymm`: the month, 'YYYYMM' (the reference parses it from the
    diagnostics file name, convert2EXT.py:36-38)."""
    out = _common(fields["lat"], fields["lon"], yyyymm[0:4], yyyymm[4:6])
    sf = np.asarray(fields["scaling_factor"])
    out["SF"] = sf.astype(np.float64).reshape((1,) + sf.shape)         # :67, an 'f8' variable
    return out


def ones_fields(lat, lon, year, month):
    """A month outside the assimilated record: SF == 1 everywhere (convert2EXT.py:80-124)."""
    out = _common(lat, lon, year, month)
    out["SF"] = np.ones((1,) + np.shape(lat), dtype=np.float64)
    return out


def write_ext(path, ext):
    """One ExtData file from `ext_fields` / `ones_fields` (needs the netCDF4 package)."""
    import time
    try:
        from netCDF4 import Dataset
    except Exception as exc:
        raise RuntimeError("write_ext needs the netCDF4 package; ext_fields() returns the "
                           "variables it would store") from exc
    nc = Dataset(path, "w", format="NETCDF4")
    nc.createDimension("time", 1)
    nc.createDimension("lat", ext["lat"].size)
    nc.createDimension("lon", ext["lon"].size)
    t = nc.createVariable("time", "f8", ("time",))
    t.long_name = "time"
    t.units = ext["time_units"]
    la = nc.createVariable("lat", "f8", ("lat",))
    la.units, la.long_name = "degrees_north", "latitude"
    lo = nc.createVariable("lon", "f8", ("lon",))
    lo.units, lo.long_name = "degrees_east", "longitude"
    sf = nc.createVariable("SF", "f8", ("time", "lat", "lon",))
    sf.units = "fraction"
    t[:], la[:], lo[:] = ext["time"], ext["lat"], ext["lon"]
    sf[:, :, :] = ext["SF"]
    for k, v in GLOBAL_ATTRIBUTES.items():
        setattr(nc, k, v)
    nc.creation_time = time.strftime("%Y-%m-%d %H:%M:%S", time.localtime())
    nc.close()
