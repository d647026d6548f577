"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the temporal averaging stage.

Follows /root/reference/oisatgmi/averaging.py:11-120 including its quirks
(SURVEY.md appendix D): the sat VCD array starts as zeros and the other four as
NaN (:53-63); the reduction block sits outside the month loop (:97-108), so only
the LAST month of the range is stored; the error is sqrt(sum(sigma^2)/n^2) with
n = number of finite sigma^2 (:11-24).  Parity status: PINNED against the live
reference.  Only tests/, smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import datetime

import numpy as np

from oisatgmi_b200.config import kind_of


def error_averager(var_stack):
    """(G, ny, nx) variances -> sqrt(sum_finite / n_finite^2); averaging.py:11-24.
    The reference compacts each cell's finite values and calls np.sum on the
    1-D result (pairwise summation) -- kept cell by cell for that reason."""
    _, ny, nx = np.shape(var_stack)
    out = np.zeros((ny, nx)) * np.nan
    for i in range(ny):
        for j in range(nx):
            col = np.array(var_stack[:, i, j], dtype=np.float64)
            col[np.isinf(col)] = np.nan
            keep = col[~np.isnan(col)]
            out[i, j] = np.sum(keep) / (np.size(keep) ** 2)
    return np.sqrt(out)


def averaging(startdate, enddate, reader_obj):
    d0 = datetime.date(int(startdate[0:4]), int(startdate[5:7]), int(startdate[8:10]))
    d1 = datetime.date(int(enddate[0:4]), int(enddate[5:7]), int(enddate[8:10]))
    days = [d0 + datetime.timedelta(n) for n in range(int((d1 - d0).days))]
    months = np.array([d.month for d in days])
    years = np.array([d.year for d in days])
    first = next(g for g in reader_obj.sat_data if g is not None)
    ny, nx = np.shape(first.latitude_center)[0], np.shape(first.latitude_center)[1]
    nm = len(range(np.min(months), np.max(months) + 1))
    nyr = len(range(np.min(years), np.max(years) + 1))
    sat_vcd = np.zeros((ny, nx, nm, nyr))
    sat_err = np.zeros_like(sat_vcd) * np.nan
    ctm_vcd = np.zeros_like(sat_vcd) * np.nan
    aux1 = np.zeros_like(sat_vcd) * np.nan
    aux2 = np.zeros_like(sat_vcd) * np.nan
    for year in range(np.min(years), np.max(years) + 1):
        for month in range(np.min(months), np.max(months) + 1):
            sel = [g for g in reader_obj.sat_data
                   if g is not None and g.time.year == year and g.time.month == month]
            times = [g.time for g in sel]
            s_vcd = np.array([g.vcd for g in sel])
            s_vcd[np.isinf(s_vcd)] = np.nan
            s_err = np.array([g.uncertainty for g in sel])
            s_ctm = np.array([g.ctm_vcd for g in sel])
            a1, a2 = [], []
            for g in sel:
                k = kind_of(g)
                if k == "amf":
                    a1.append(g.new_amf)
                    a2.append(g.old_amf)
                elif k == "opt":
                    a1.append(g.x_col)
                    a2.append(g.ctm_xcol)
                else:
                    a1.append(np.nan * g.vcd)
                    a2.append(np.nan * g.vcd)
            a1, a2 = np.array(a1), np.array(a2)
        # NB: dedented on purpose -- averaging.py:97-108 runs once per YEAR with
        # the loop variables left over from the last month iteration.
        mi, yi = month - min(months), year - min(years)
        if np.size(s_vcd) != 0:
            sat_vcd[:, :, mi, yi] = np.squeeze(np.nanmean(s_vcd, axis=0))
            sat_err[:, :, mi, yi] = error_averager(s_err ** 2)
            ctm_vcd[:, :, mi, yi] = np.squeeze(np.nanmean(s_ctm, axis=0))
        if np.size(a1) != 0:
            aux1[:, :, mi, yi] = np.squeeze(np.nanmean(a1, axis=0))
            aux2[:, :, mi, yi] = np.squeeze(np.nanmean(a2, axis=0))
    stamps = [t.timestamp() for t in times]
    mean_time = datetime.datetime.fromtimestamp(sum(stamps) / len(stamps))
    return (sat_vcd.squeeze(), sat_err.squeeze(), ctm_vcd.squeeze(), aux1.squeeze(),
            aux2.squeeze(), mean_time)
