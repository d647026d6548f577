"""Drop-in for /root/reference/oisatgmi/pwv_cal.py: `pwv_calculator(ctm_data, sat_data)`.

Model precipitable water on the grid of every gridded SSMIS map (pwv_cal.py:7-101): the layer
partial columns delta_p * q / g / 10000 (float32, K `oisat_pwv_partial`), resampled to the
satellite mesh first when the model is the finer grid (`interpolator._upscaler` per layer in the
reference, K6 here), summed over the layers and masked where the map has no value
(`oisat_pwv_column`).  Every non-None granule gets `.ctm_vcd`; the list is returned.  The
reference also triangulates the whole model grid per call and hands the result to an
`_upscaler` that ignores it (:31-34,85-86); that is not reproduced.
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, _vertical as _v
from .ak_conv_mopitt import model_fields

__all__ = ["pwv_calculator"]


def pwv_calculator(ctm_data: list, sat_data: list):
    _dev.require_cuda()
    L = _lib.lib()
    stamps, _ = _v.ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        _, day = _v.closest_day(ctm_data, stamps, g.time)
        _pmid, prof_d, dp_d = model_fields(ctm_data, day)
        if prof_d.dtype != _dev.torch().float32 or dp_d.dtype != _dev.torch().float32:
            raise _lib.OisatError("model fields must be float32 as delivered by the readers")
        n_lev, n_cell = int(dp_d.shape[0]), int(dp_d.shape[1])
        pc = _dev.empty((n_lev, n_cell), "float32")
        _lib.check(L.oisat_pwv_partial(dp_d.data_ptr(), prof_d.data_ptr(), n_lev * n_cell,
                                       pc.data_ptr(), _dev.stream()))
        upscaled = g.ctm_upscaled_needed == True  # noqa: E712
        if upscaled:
            (pc,) = _v.resample_to_sat([(pc, None, _lib.SRC_VALUE)], ctm_data, g)   # float64
        shape = np.shape(g.vcd)
        vcd_d = _dev.to_device(np.ascontiguousarray(np.asarray(g.vcd, dtype=np.float64)).ravel())
        out = _dev.empty((vcd_d.numel(),))
        _lib.check(L.oisat_pwv_column(pc.data_ptr(), _dev.dtype_code(pc), n_lev, vcd_d.numel(),
                                      vcd_d.data_ptr(), out.data_ptr(), _dev.stream()))
        pwv = _dev.to_host(out).reshape(shape)
        g.ctm_vcd = pwv if upscaled else pwv.astype(np.float32)     # np.nansum keeps the dtype
    return sat_data
