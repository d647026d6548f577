"""Drop-in for /root/reference/oisatgmi/amf_recal.py: `amf_recal(ctm_data, sat_data)`.

Same contract as the reference (amf_recal.py:121-185): the granules of
`sat_data` are updated in place -- `.vcd` (re-scaled with the new air-mass
factor), `.ctm_vcd`, `.old_amf`, `.new_amf`, `.ctm_time_at_sat` -- `None`
granules are skipped and the same list is returned.  The per-cell Python loop
with one scipy interp1d object per cell (amf_recal.py:93-119) is replaced by
one K3 launch per granule (one warp per valid cell).
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, _vertical as _v

__all__ = ["amf_recal"]


def _ctm_fields(ctm_data, day, hour):
    """Device slabs [nlev][ny*nx] of the matched slot (amf_recal.py:39-49)."""
    c = ctm_data[day]
    slot = None if ctm_data[0].ctmtype == "FREE" else hour
    return (_v.ctm_slot_device(c.pressure_mid, slot), _v.ctm_slot_device(c.gas_profile, slot),
            _v.ctm_slot_device(c.delta_p, slot))


def amf_recal(ctm_data: list, sat_data: list):
    _dev.require_cuda()
    L = _lib.lib()
    stamps, fracs = _v.ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        k, day, hour = _v.closest_slot(ctm_data, stamps, fracs, g.time)
        pmid_d, prof_d, dp_d = _ctm_fields(ctm_data, day, hour)
        n_ctm = pmid_d.shape[0]
        mode = 0 if pmid_d.dtype == _dev.torch().float32 else 1
        if g.ctm_upscaled_needed == True:  # noqa: E712
            pmid_d, pc_d = _v.resample_to_sat(
                [(pmid_d, None, _lib.SRC_VALUE), (dp_d, prof_d, _lib.SRC_PARTIAL_COLUMN)],
                ctm_data, g)
            prof_d, dp_d, mode = pc_d, None, 1
        elif mode == 1:
            raise _lib.OisatError("model fields must be float32 as delivered by the readers")
        shape = np.shape(g.vcd)
        vcd = np.asarray(g.vcd, dtype=np.float64)
        valid = np.flatnonzero(~np.isnan(vcd).ravel())
        has_trop = np.size(g.tropopause) != 1
        n = valid.size
        cidx = _dev.to_device(valid.astype(np.int32))
        vcd_d = _dev.to_device(_v.compact(vcd, valid))
        trop_d = _dev.to_device(_v.compact(g.tropopause, valid)) if has_trop else None
        ctm_vcd_d = _dev.empty((n,))
        if np.size(g.scattering_weights) == 1:
            # no scattering weights: amf_recal.py:160-171
            _lib.check(L.oisat_vertical_column(
                n, None, cidx.data_ptr(), vcd_d.data_ptr(), _dev.ptr(trop_d), pmid_d.data_ptr(),
                prof_d.data_ptr(), _dev.ptr(dp_d), mode, n_ctm, pmid_d.shape[1],
                ctm_vcd_d.data_ptr(), _dev.stream()))
            col = _v.scatter(shape, valid, _dev.to_host(ctm_vcd_d))
            g.ctm_vcd = col.astype(np.float32) if mode == 0 else col
            g.ctm_time_at_sat = stamps[k]
            g.old_amf = np.empty((1))
            g.new_amf = np.empty((1))
            continue
        nlev = np.shape(g.pressure_mid)[0]
        amf_d = _dev.to_device(_v.compact(g.amf, valid))
        psat_d = _dev.to_device(_v.compact(g.pressure_mid, valid, nlev))
        sw_d = _dev.to_device(_v.compact(g.scattering_weights, valid, nlev))
        new_amf_d = _dev.empty((n,))
        vcd_out_d = _dev.empty((n,))
        _lib.check(L.oisat_vertical_amf(
            n, None, cidx.data_ptr(), vcd_d.data_ptr(), amf_d.data_ptr(), _dev.ptr(trop_d),
            psat_d.data_ptr(), sw_d.data_ptr(), nlev, n, pmid_d.data_ptr(), prof_d.data_ptr(),
            _dev.ptr(dp_d), mode, n_ctm, pmid_d.shape[1], new_amf_d.data_ptr(),
            ctm_vcd_d.data_ptr(), vcd_out_d.data_ptr(), _dev.stream()))
        g.old_amf = getattr(g, "amf", None)
        g.new_amf = _v.scatter(shape, valid, _dev.to_host(new_amf_d))
        g.vcd = _v.scatter(shape, valid, _dev.to_host(vcd_out_d))
        g.ctm_vcd = _v.scatter(shape, valid, _dev.to_host(ctm_vcd_d))
        g.ctm_time_at_sat = stamps[k]
    return sat_data
