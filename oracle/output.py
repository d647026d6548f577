"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the data side of
driver.write_to_nc (/root/reference/oisatgmi/driver.py:156-227, SURVEY.md section
8f-3): the float32 variables the reference stores per month, the emission scaling
factor among them (:203-206).  The NetCDF container itself is file I/O, not restated.

Parity status: PINNED -- tests/test_oracle_vs_reference.py runs the unmodified
`oisatgmi.write_to_nc` against a recording stand-in for netCDF4.Dataset and compares
every stored array bit for bit.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module; the product package never does.
"""
from __future__ import annotations

import numpy as np

NAMES = ("sat_averaged_vcd", "ctm_averaged_vcd_prior", "ctm_averaged_vcd_posterior",
         "sat_averaged_error", "ak_OI", "error_OI", "scaling_factor", "aux1", "aux2")


def output_fields(obj):
    """`obj`: anything with the attributes `driver.oisatgmi` holds after average() and
    oi().  A float64 array assigned to a float32 NetCDF variable is cast like astype."""
    f32 = lambda a: np.asarray(a).astype(np.float32)  # noqa: E731
    with np.errstate(all="ignore"):
        scaling = obj.ctm_averaged_vcd_corrected / obj.ctm_averaged_vcd
    scaling[np.where((np.isnan(scaling)) | (np.isinf(scaling)) | (scaling == 0.0))] = 1.0
    first = next(s for s in obj.reader_obj.sat_data if s is not None)
    return {
        "sat_averaged_vcd": f32(obj.sat_averaged_vcd),
        "ctm_averaged_vcd_prior": f32(obj.ctm_averaged_vcd),
        "ctm_averaged_vcd_posterior": f32(obj.ctm_averaged_vcd_corrected),
        "sat_averaged_error": f32(obj.sat_averaged_error),
        "ak_OI": f32(obj.ak_OI), "error_OI": f32(obj.error_OI),
        "scaling_factor": f32(scaling), "aux1": f32(obj.aux1), "aux2": f32(obj.aux2),
        "lon": f32(first.longitude_center), "lat": f32(first.latitude_center),
        "time": obj.avg_time.strftime("%Y-%m-%d %H:%M:%S"),
    }


def ext_fields(lat, lon, scaling_factor, year, month):
    """tools/convert2EXT.py:32-78 (data side): the variables of one GEOS ExtData file from the
    `lat`, `lon`, `scaling_factor` arrays of a diagnostics file.  PINNED: the unmodified script
    is run against recording stand-ins for netCDF4.Dataset (tests/test_oracle_vs_reference.py)."""
    import datetime
    t0 = datetime.datetime(int(year), int(month), 1) + datetime.timedelta(seconds=int(0.0))
    sf = np.zeros((1,) + np.shape(lat), dtype=np.float64)
    sf[:, :, :] = scaling_factor
    return {"time": np.array([0.0]), "time_units": "hours since " + t0.strftime("%Y-%m-%d %H:%M:%S"),
            "lat": np.array(lat[:, 0].squeeze(), dtype=np.float64),
            "lon": np.array(lon[0, :].squeeze(), dtype=np.float64), "SF": sf}
