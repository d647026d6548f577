"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-cell vertical operators.

  amf_recal       /root/reference/oisatgmi/amf_recal.py:121-185
  ak_conv_mopitt  /root/reference/oisatgmi/ak_conv_mopitt.py:8-149
  ak_conv_gosat   /root/reference/oisatgmi/ak_conv_gosat.py:8-147

Same numpy/scipy calls as the reference (per-cell `interp1d` in log pressure,
`nansum`), including its dtype behaviour: model fields are float32 and stay
float32 through the partial-column arithmetic.  Parity status:
PINNED against the live reference (tests/test_oracle_vs_reference.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import interpolate

from oracle.interp import upscale

G0 = 9.80665
M_AIR = 28.97e-3
N_AVO = 6.02214076e23


def _stamp(t):
    """YYYYMMDD.fraction-of-day (amf_recal.py:7-16)."""
    return (t.year * 10000 + t.month * 100 + t.day + t.hour / 24.0
            + t.minute / 60.0 / 24.0 + t.second / 3600.0 / 24.0)


def _day_fraction(t):
    return t.hour / 24.0 + t.minute / 60.0 / 24.0 + t.second / 3600.0 / 24.0


def _ctm_clock(ctm_data):
    stamps, fracs = [], []
    for c in ctm_data:
        stamps.extend(_stamp(t) for t in c.time)
        fracs.extend(_day_fraction(t) for t in c.time)
    return np.array(stamps), np.array(fracs)


def partial_column(delta_p, profile):
    # amf_recal.py:51-56 -- left-to-right, float32 stays float32 (NEP 50)
    return delta_p * profile / G0 / M_AIR * N_AVO * 1e-4 * 1e-15 * 100.0 * 1e-9


def air_column(delta_p):
    # ak_conv_mopitt.py:68
    return delta_p / G0 / M_AIR * N_AVO * 1e-4 * 1e-15 * 100.0


def _resample_to_sat(fields, ctm_data, granule):
    """CTM -> satellite grid for the model-finer-than-grid branch
    (amf_recal.py:58-83, ak_conv_mopitt.py:79-110): level by level through the
    interpolator's `_upscaler` with the roles of the two grids swapped."""
    sat = {"Longitude": granule.longitude_center, "Latitude": granule.latitude_center}
    dlon_s = np.abs(sat["Longitude"][0, 0] - sat["Longitude"][0, 1])
    dlat_s = np.abs(sat["Latitude"][0, 0] - sat["Latitude"][1, 0])
    thr = np.sqrt(dlon_s ** 2 + dlat_s ** 2)
    clon, clat = ctm_data[0].longitude, ctm_data[0].latitude
    gs = np.sqrt(np.abs(clon[0, 0] - clon[0, 1]) ** 2 + np.abs(clat[0, 0] - clat[1, 0]) ** 2)
    nlev = fields[0].shape[0]
    outs = [np.full((nlev,) + np.shape(sat["Longitude"]), np.nan) for _ in fields]
    for z in range(nlev):
        for o, f in zip(outs, fields):
            _, _, o[z], _ = upscale(clon, clat, f[z], sat, gs, thr)
    return outs


def amf_recal(ctm_data, sat_data):
    stamps, fracs = _ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        if not ctm_data[0].averaged:
            k = int(np.argmin(np.abs(_stamp(g.time) - stamps)))
            day, hour = int(np.floor(k / 8.0)), int(k % 8)
        else:
            k = int(np.argmin(np.abs(_day_fraction(g.time) - fracs)))
            day, hour = 0, k
        c = ctm_data[day]
        if ctm_data[0].ctmtype == "FREE":
            pmid, prof, dp = c.pressure_mid.squeeze(), c.gas_profile.squeeze(), c.delta_p.squeeze()
        else:
            pmid, prof, dp = (c.pressure_mid[hour].squeeze(), c.gas_profile[hour].squeeze(),
                              c.delta_p[hour].squeeze())
        pc = partial_column(dp, prof)
        if g.ctm_upscaled_needed == True:  # noqa: E712 (reference spelling)
            nlev = pmid.shape[0]
            per_level_pc = np.stack([partial_column(dp[z], prof[z]) for z in range(nlev)])
            pmid, pc = _resample_to_sat([pmid, per_level_pc], ctm_data, g)
        has_trop = np.size(g.tropopause) != 1
        if np.size(g.scattering_weights) == 1:
            # no scattering weights (e.g. O3): amf_recal.py:160-171
            if has_trop:
                for z in range(pc.shape[0]):
                    pc[z][pmid[z] < g.tropopause] = np.nan
            col = np.nansum(pc, axis=0)
            col[np.isnan(g.vcd)] = np.nan
            g.ctm_vcd = col
            g.ctm_time_at_sat = stamps[k]
            g.old_amf = np.empty((1))
            g.new_amf = np.empty((1))
            continue
        new_amf = np.full_like(g.vcd, np.nan)
        col = np.full_like(g.vcd, np.nan)
        for i in range(g.vcd.shape[0]):
            for j in range(g.vcd.shape[1]):
                if np.isnan(g.vcd[i, j]):
                    continue
                pc_ij = pc[:, i, j]          # a view: masking writes through (amf_recal.py:101,114)
                p_ij = pmid[:, i, j]
                f = interpolate.interp1d(np.log(g.pressure_mid[:, i, j].squeeze()),
                                         g.scattering_weights[:, i, j].squeeze(),
                                         fill_value="extrapolate")
                sw = f(np.log(p_ij))
                sw[np.isinf(sw)] = 0.0
                if has_trop:
                    below = p_ij < g.tropopause[i, j]
                    sw[below] = np.nan
                    pc_ij[below] = np.nan
                scd = np.nansum(sw * pc_ij)
                col[i, j] = np.nansum(pc_ij)
                new_amf[i, j] = scd / col[i, j] if col[i, j] != 0 else np.nan
        g.old_amf = getattr(g, "amf", None)
        new_amf[np.isnan(g.vcd)] = np.nan
        g.new_amf = new_amf
        g.vcd = (g.amf * g.vcd) / new_amf
        col[np.isnan(g.vcd)] = np.nan
        col[np.isinf(g.vcd)] = np.nan
        g.ctm_vcd = col
        g.ctm_time_at_sat = stamps[k]
    return sat_data


def _ak_fields(ctm_data, g, stamps):
    """Common head of the two AK operators (ak_conv_mopitt.py:41-114)."""
    if ctm_data[0].averaged == False:  # noqa: E712
        day_stamp = g.time.year * 10000 + g.time.month * 100 + g.time.day
        k = int(np.argmin(np.abs(day_stamp - stamps)))
        day = int(np.floor(k))
    else:
        k, day = 0, 0
    c = ctm_data[day]
    if c.ctmtype in ("ECCOH", "FREE"):
        pmid, prof, dp = c.pressure_mid.squeeze(), c.gas_profile.squeeze(), c.delta_p.squeeze()
    elif c.ctmtype == "GMI":
        pmid = np.nanmean(c.pressure_mid, axis=0).squeeze()
        prof = np.nanmean(c.gas_profile, axis=0).squeeze()
        dp = np.nanmean(c.delta_p, axis=0).squeeze()
    pc = partial_column(dp, prof)
    air = air_column(dp)
    if g.ctm_upscaled_needed == True:  # noqa: E712
        pmid, prof, pc, air = _resample_to_sat([pmid, prof, pc, air], ctm_data, g)
    return k, pmid, prof, pc, air


def ak_conv_mopitt(ctm_data, sat_data):
    stamps, _ = _ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        k, pmid, prof, pc, air = _ak_fields(ctm_data, g, stamps)
        col = np.zeros_like(g.vcd) * np.nan
        xcol = np.zeros_like(g.vcd) * np.nan
        for i in range(g.vcd.shape[0]):
            for j in range(g.vcd.shape[1]):
                if np.isnan(g.vcd[i, j]):
                    continue
                x = prof[:, i, j].squeeze()
                f = interpolate.interp1d(np.log(pmid[:, i, j].squeeze()), x,
                                         fill_value=np.nan, bounds_error=False)
                xi = f(np.log(g.pressure_mid[:, i, j].squeeze()))
                prof_part = g.aprior_column[i, j] + np.nansum(
                    g.averaging_kernels[1::, i, j].squeeze()
                    * (np.log10(xi) - np.log10(g.apriori_profile[:, i, j].squeeze())))
                sfc_part = g.averaging_kernels[0, i, j].squeeze() * (
                    np.log10(x[0]) - np.log10(g.apriori_surface[i, j]))
                col[i, j] = prof_part + sfc_part
                xcol[i, j] = 1e6 * col[i, j] / np.nansum(air[:, i, j].squeeze())
        col[np.isnan(g.vcd)] = np.nan
        col[np.isinf(g.vcd)] = np.nan
        g.ctm_vcd = col
        g.ctm_xcol = xcol
        g.ctm_time_at_sat = stamps[k]
    return sat_data


def ak_conv_gosat(ctm_data, sat_data):
    stamps, _ = _ctm_clock(ctm_data)
    for g in sat_data:
        if g is None:
            continue
        k, pmid, prof, pc, air = _ak_fields(ctm_data, g, stamps)
        col = np.zeros_like(g.vcd) * np.nan
        xcol = np.zeros_like(g.vcd) * np.nan
        for i in range(g.x_col.shape[0]):
            for j in range(g.x_col.shape[1]):
                if np.isnan(g.x_col[i, j]):
                    continue
                f = interpolate.interp1d(np.log(pmid[:, i, j].squeeze()),
                                         prof[:, i, j].squeeze(), fill_value="extrapolate")
                xi = f(np.log(g.pressure_mid[:, i, j].squeeze()))
                ap = g.apriori_profile[:, i, j].squeeze()
                t = ap + (xi - ap) * g.averaging_kernels[:, i, j].squeeze()
                t = t * g.pressure_weight[:, i, j].squeeze()
                t[t <= 0] = np.nan
                xcol[i, j] = np.nansum(t)
        g.ctm_vcd = col  # all NaN by design (ak_conv_gosat.py:138)
        xcol[np.isinf(g.x_col)] = np.nan
        xcol[np.isnan(g.x_col)] = np.nan
        g.ctm_xcol = xcol
        g.ctm_time_at_sat = stamps[k]
    return sat_data
