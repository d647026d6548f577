"""Drop-in for /root/reference/oisatgmi/ak_conv_mopitt.py:
`ak_conv_mopitt(ctm_data, sat_data)`.

Contract as in the reference (ak_conv_mopitt.py:8-149): every non-None granule
gets `.ctm_vcd`, `.ctm_xcol` (ppmv) and `.ctm_time_at_sat`; the list is
returned.  The model profile is interpolated in log pressure to the nine
MOPITT levels and folded with the ten averaging-kernel rows (row 0 = surface)
by one K3 launch per granule; when the model is finer than the 1 degree L3
mesh its fields are first resampled with K6 (ak_conv_mopitt.py:79-110).  The
reference also triangulates all 207,936 model points per call and never uses
the result (ak_conv_mopitt.py:31-34); that is not reproduced.
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, _vertical as _v

__all__ = ["ak_conv_mopitt"]


def model_fields(ctm_data, day):
    """(p_mid, profile, delta_p) device slabs: ECCOH/FREE files are 3-D, GMI
    files are averaged over their time axis first (ak_conv_mopitt.py:60-77)."""
    c = ctm_data[day]
    if c.ctmtype in ("ECCOH", "FREE"):
        get = lambda a: _v.ctm_slot_device(a, None)  # noqa: E731
    elif c.ctmtype == "GMI":
        get = _v.ctm_time_mean_device
    else:
        raise _lib.OisatError("unsupported model type %r" % (c.ctmtype,))
    return get(c.pressure_mid), get(c.gas_profile), get(c.delta_p)


def sat_grid_fields(ctm_data, g, day):
    """Model columns on the satellite grid: (p_mid, profile, air or delta_p, mode)."""
    pmid_d, prof_d, dp_d = model_fields(ctm_data, day)
    if g.ctm_upscaled_needed == True:  # noqa: E712
        if pmid_d.dtype != _dev.torch().float32:
            raise _lib.OisatError("model fields must be float32 as delivered by the readers")
        pmid_s, prof_s, air_s = _v.resample_to_sat(
            [(pmid_d, None, _lib.SRC_VALUE), (prof_d, None, _lib.SRC_VALUE),
             (dp_d, None, _lib.SRC_AIR_COLUMN)], ctm_data, g)
        return pmid_s, prof_s, air_s, 1
    return pmid_d, prof_d, dp_d, 0


def ak_conv_mopitt(ctm_data: list, sat_data: list):
    _dev.require_cuda()
    L = _lib.lib()
    stamps, _ = _v.ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        k, day = _v.closest_day(ctm_data, stamps, g.time)
        pmid_d, prof_d, third_d, mode = sat_grid_fields(ctm_data, g, day)
        n_ctm = pmid_d.shape[0]
        shape = np.shape(g.vcd)
        vcd = np.asarray(g.vcd, dtype=np.float64)
        valid = np.flatnonzero(~np.isnan(vcd).ravel())
        n = valid.size
        nlev = np.shape(g.pressure_mid)[0]
        cidx = _dev.to_device(valid.astype(np.int32))
        vcd_d = _dev.to_device(_v.compact(vcd, valid))
        apcol_d = _dev.to_device(_v.compact(g.aprior_column, valid))
        apsfc_d = _dev.to_device(_v.compact(g.apriori_surface, valid))
        psat_d = _dev.to_device(_v.compact(g.pressure_mid, valid, nlev))
        ak_d = _dev.to_device(_v.compact(g.averaging_kernels, valid, nlev + 1))
        approf_d = _dev.to_device(_v.compact(g.apriori_profile, valid, nlev))
        col_d = _dev.empty((n,))
        xcol_d = _dev.empty((n,))
        _lib.check(L.oisat_vertical_mopitt(
            n, None, cidx.data_ptr(), vcd_d.data_ptr(), apcol_d.data_ptr(), apsfc_d.data_ptr(),
            psat_d.data_ptr(), ak_d.data_ptr(), approf_d.data_ptr(), nlev, n, pmid_d.data_ptr(),
            prof_d.data_ptr(), third_d.data_ptr(), mode, n_ctm, pmid_d.shape[1],
            col_d.data_ptr(), xcol_d.data_ptr(), _dev.stream()))
        g.ctm_vcd = _v.scatter(shape, valid, _dev.to_host(col_d))
        g.ctm_xcol = _v.scatter(shape, valid, _dev.to_host(xcol_d))
        g.ctm_time_at_sat = stamps[k]
    return sat_data
