"""Drop-in for /root/reference/oisatgmi/ak_conv_gosat.py:
`ak_conv_gosat(ctm_data, sat_data)`.

Contract as in the reference (ak_conv_gosat.py:8-147): every non-None granule
gets `.ctm_xcol` (ppbv), an all-NaN `.ctm_vcd` (by design, :138) and
`.ctm_time_at_sat`; the loop is keyed on `x_col`, not `vcd` (:118-121).
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, _vertical as _v
from .ak_conv_mopitt import sat_grid_fields

__all__ = ["ak_conv_gosat"]


def ak_conv_gosat(ctm_data: list, sat_data: list):
    _dev.require_cuda()
    L = _lib.lib()
    stamps, _ = _v.ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        k, day = _v.closest_day(ctm_data, stamps, g.time)
        pmid_d, prof_d, _third, mode = sat_grid_fields(ctm_data, g, day)
        n_ctm = pmid_d.shape[0]
        shape = np.shape(g.x_col)
        xcol = np.asarray(g.x_col, dtype=np.float64)
        valid = np.flatnonzero(~np.isnan(xcol).ravel())
        n = valid.size
        nlev = np.shape(g.pressure_mid)[0]
        cidx = _dev.to_device(valid.astype(np.int32))
        x_d = _dev.to_device(_v.compact(xcol, valid))
        psat_d = _dev.to_device(_v.compact(g.pressure_mid, valid, nlev))
        ak_d = _dev.to_device(_v.compact(g.averaging_kernels, valid, nlev))
        approf_d = _dev.to_device(_v.compact(g.apriori_profile, valid, nlev))
        pw_d = _dev.to_device(_v.compact(g.pressure_weight, valid, nlev))
        out_d = _dev.empty((n,))
        _lib.check(L.oisat_vertical_gosat(
            n, None, cidx.data_ptr(), x_d.data_ptr(), psat_d.data_ptr(), ak_d.data_ptr(),
            approf_d.data_ptr(), pw_d.data_ptr(), nlev, n, pmid_d.data_ptr(), prof_d.data_ptr(),
            mode, n_ctm, pmid_d.shape[1], out_d.data_ptr(), _dev.stream()))
        g.ctm_vcd = np.zeros_like(np.asarray(g.vcd, dtype=np.float64)) * np.nan
        g.ctm_xcol = _v.scatter(shape, valid, _dev.to_host(out_d))
        g.ctm_time_at_sat = stamps[k]
    return sat_data
