"""Drop-in for /root/reference/oisatgmi/averaging.py:
`averaging(startdate, enddate, reader_obj) -> (sat_vcd, sat_err, ctm_vcd, aux1, aux2, avg_datetime)`.

The reference stacks every granule grid of the month and reduces with
np.nanmean / a Python triple loop (averaging.py:64-108, 11-24).  Here each
granule is added, in list order, to a [10][n_cell] device block of running sums
and exact counts (K4); the means and sqrt(sum sigma^2 / n^2) are taken once at
the end.  Quirks kept on purpose: only the LAST month of a
multi-month range is stored (the reduction sits outside the month loop,
averaging.py:97-108), the satellite VCD slab starts as zeros and the other four
as NaN (:53-63), and the mean time stamp goes through local-time
`datetime.fromtimestamp` (:116-118).
"""
from __future__ import annotations

import datetime

import numpy as np

from . import _dev, _lib
from .config import kind_of

__all__ = ["averaging", "MonthAccumulator"]


class MonthAccumulator:
    """[10][n_cell] float64 running sums (rows 0-4) and counts (rows 5-9) of
    (sat vcd, sigma^2, model vcd, aux1, aux2) on the device."""

    def __init__(self, n_cell):
        self.n_cell = int(n_cell)
        self.acc = _dev.zeros((10, self.n_cell))

    def add(self, vcd, sigma, ctm_vcd, aux1, aux2, sigma_is_variance=False):
        """Arguments: device float64 tensors of n_cell elements, or None."""
        L = _lib.lib()
        fn = L.oisat_accum_add_variance if sigma_is_variance else L.oisat_accum_add
        _lib.check(fn(self.acc.data_ptr(), self.n_cell, _dev.ptr(vcd), _dev.ptr(sigma),
                      _dev.ptr(ctm_vcd), _dev.ptr(aux1), _dev.ptr(aux2), _dev.stream()))

    def finalize(self):
        L = _lib.lib()
        outs = [_dev.empty((self.n_cell,)) for _ in range(5)]
        _lib.check(L.oisat_accum_finalize(self.acc.data_ptr(), self.n_cell,
                                          *[o.data_ptr() for o in outs], _dev.stream()))
        return outs


def _grid_or_none(a, n_cell):
    """Host array -> device float64 vector; size-1 placeholders (O3 granules,
    amf_recal.py:169-170) broadcast like numpy would in np.array(list)."""
    if a is None or isinstance(a, list):
        return None
    a = np.asarray(a, dtype=np.float64)
    if a.size == 1:
        a = np.full(n_cell, a.reshape(-1)[0])
    return _dev.to_device(a.ravel())


def averaging(startdate: str, enddate: str, reader_obj):
    _dev.require_cuda()
    d0 = datetime.date(int(startdate[0:4]), int(startdate[5:7]), int(startdate[8:10]))
    d1 = datetime.date(int(enddate[0:4]), int(enddate[5:7]), int(enddate[8:10]))
    days = [d0 + datetime.timedelta(n) for n in range(int((d1 - d0).days))]
    months = np.array([d.month for d in days])
    years = np.array([d.year for d in days])
    first = next(g for g in reader_obj.sat_data if g is not None)
    ny, nx = np.shape(first.latitude_center)[0], np.shape(first.latitude_center)[1]
    n_cell = ny * nx
    nm = len(range(np.min(months), np.max(months) + 1))
    nyr = len(range(np.min(years), np.max(years) + 1))
    sat_vcd = np.zeros((ny, nx, nm, nyr))
    sat_err = np.zeros_like(sat_vcd) * np.nan
    ctm_vcd = np.zeros_like(sat_vcd) * np.nan
    aux1 = np.zeros_like(sat_vcd) * np.nan
    aux2 = np.zeros_like(sat_vcd) * np.nan
    times = []
    for year in range(np.min(years), np.max(years) + 1):
        # averaging.py:97-108 is dedented out of the month loop: only the last
        # month of the range is reduced, once per year
        month = int(np.max(months))
        sel = [g for g in reader_obj.sat_data
               if g is not None and g.time.year == year and g.time.month == month]
        times = [g.time for g in sel]
        if not sel:
            continue
        acc = MonthAccumulator(n_cell)
        for g in sel:
            k = kind_of(g)
            if k == "amf":
                a1, a2 = g.new_amf, g.old_amf
            elif k == "opt":
                a1, a2 = g.x_col, g.ctm_xcol
            else:
                a1 = a2 = None
            unc = np.asarray(g.uncertainty)
            narrow = unc.dtype in (np.float16, np.float32)
            if narrow:
                # averaging.py:101 squares the stacked uncertainties in THEIR dtype
                with np.errstate(over="ignore"):
                    unc = unc ** 2
            acc.add(_grid_or_none(g.vcd, n_cell), _grid_or_none(unc, n_cell),
                    _grid_or_none(g.ctm_vcd, n_cell), _grid_or_none(a1, n_cell),
                    _grid_or_none(a2, n_cell), sigma_is_variance=narrow)
        outs = [_dev.to_host(o).reshape(ny, nx) for o in acc.finalize()]
        mi, yi = month - min(months), year - min(years)
        sat_vcd[:, :, mi, yi], sat_err[:, :, mi, yi], ctm_vcd[:, :, mi, yi] = outs[0], outs[1], outs[2]
        aux1[:, :, mi, yi], aux2[:, :, mi, yi] = outs[3], outs[4]
    stamps = [t.timestamp() for t in times]
    mean_time = datetime.datetime.fromtimestamp(sum(stamps) / len(stamps))
    return (sat_vcd.squeeze(), sat_err.squeeze(), ctm_vcd.squeeze(), aux1.squeeze(),
            aux2.squeeze(), mean_time)
