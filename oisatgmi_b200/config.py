"""In-memory wire format between the hot-path stages.

The reference passes four positional dataclasses between every layer
(/root/reference/oisatgmi/config.py:6-73) and constructs them positionally
(e.g. interpolator.py:285-290), so the field ORDER is part of the boundary.
The records are generated here from plain field tables; the tables are the
single source of truth for the order and are also used by the device
hand-off code to know which members are per-pixel arrays.
"""
from __future__ import annotations

import dataclasses as _dc
import datetime as _dt
import typing as _t

import numpy as _np

_A = _np.ndarray

# (name, type) in the reference's positional order
_SATELLITE_AMF = (
    ("vcd", _A), ("amf", _A), ("time", _dt.datetime), ("tropopause", _A),
    ("latitude_center", _A), ("longitude_center", _A),
    ("latitude_corner", _A), ("longitude_corner", _A),
    ("uncertainty", _A), ("quality_flag", _A),
    ("pressure_mid", _A), ("scattering_weights", _A),
    ("ctm_upscaled_needed", bool), ("ctm_vcd", _A),
    ("ctm_time_at_sat", _dt.datetime), ("old_amf", _A), ("new_amf", _A),
)

_SATELLITE_OPT = (
    ("vcd", _A), ("time", _dt.datetime), ("profile", _A), ("tropopause", _A),
    ("latitude_center", _A), ("longitude_center", _A),
    ("latitude_corner", _A), ("longitude_corner", _A),
    ("uncertainty", _A), ("quality_flag", _A),
    ("pressure_mid", _A), ("averaging_kernels", _A),
    ("ctm_upscaled_needed", bool), ("ctm_vcd", _A), ("ctm_xcol", _A),
    ("ctm_time_at_sat", _dt.datetime),
    ("aprior_column", _A), ("apriori_profile", _A), ("surface_pressure", _A),
    ("apriori_surface", _A), ("x_col", _A), ("pressure_weight", _A),
    ("sensor", str),
)

_SATELLITE_SSMIS = (
    ("vcd", _A), ("uncertainty", _A), ("time", _dt.datetime),
    ("latitude_center", _A), ("longitude_center", _A),
    ("ctm_upscaled_needed", bool), ("ctm_vcd", _A), ("sensor", str),
)

# NB: "tempeature_mid" is the reference's own spelling (config.py:70) and is
# part of the interface.
_CTM_MODEL = (
    ("latitude", _A), ("longitude", _A), ("time", list),
    ("gas_profile", _A), ("pressure_mid", _A), ("tempeature_mid", _A),
    ("delta_p", _A), ("ctmtype", str), ("averaged", bool),
)


def _record(name: str, table) -> type:
    cls = _dc.make_dataclass(name, [(n, t) for n, t in table])
    cls.__module__ = __name__
    cls.__doc__ = "%s(%s)" % (name, ", ".join(n for n, _ in table))
    return cls


satellite_amf = _record("satellite_amf", _SATELLITE_AMF)
satellite_opt = _record("satellite_opt", _SATELLITE_OPT)
satellite_ssmis = _record("satellite_ssmis", _SATELLITE_SSMIS)
ctm_model = _record("ctm_model", _CTM_MODEL)


def field_values(obj) -> list:
    """Positional field values of a record (no deep copy)."""
    return [getattr(obj, f.name) for f in _dc.fields(obj)]


def convert(obj, target_cls):
    """Rebuild `obj` as `target_cls` (same positional layout); used by the test
    harness to hand identical inputs to the reference's own dataclasses."""
    return target_cls(*field_values(obj))


def kind_of(obj) -> _t.Optional[str]:
    """'amf' / 'opt' / 'ssmis' by duck-typing on the field set, so that records
    created by the reference's own config module are accepted as well."""
    names = {f.name for f in _dc.fields(obj)} if _dc.is_dataclass(obj) else set()
    if "scattering_weights" in names:
        return "amf"
    if "averaging_kernels" in names:
        return "opt"
    if "sensor" in names and "vcd" in names:
        return "ssmis"
    return None
