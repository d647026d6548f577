"""Drop-in for /root/reference/oisatgmi/optimal_interpolation.py:
`OI(Xa, Y, Sa, So, regularization_on=True) -> (Xb, AK, increment, error)`.

Element-wise (diagonal) optimal interpolation.  As in the reference
(optimal_interpolation.py:14) negative observations are clipped IN PLACE in the
caller's `Y`.  The 99 regularisation factors are swept in one GPU pass whose
sums follow numpy's pairwise order (K5, csrc/k5_oi.cu), the knee is picked on
the host (kneedle.py) and the chosen factor is applied by a second launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib
from .kneedle import knee_index

__all__ = ["OI"]


def regularisation_factors(on=True):
    return np.arange(0.1, 10, 0.1) if on == True else np.array([1.0])  # noqa: E712


def sweep_device(Sa_d, So_d, factors):
    """nanmean(AK_r) for every factor; returns float64 host array."""
    L = _lib.lib()
    n = Sa_d.numel()
    nf = len(factors)
    sums = _dev.empty((nf,))
    cnts = _dev.empty((nf,))
    wbytes = int(L.oisat_oi_sweep_workspace(n, nf))
    work = _dev.empty(((wbytes + 7) // 8,))
    fac = (C.c_double * nf)(*[float(f) for f in factors])
    _lib.check(L.oisat_oi_sweep(Sa_d.data_ptr(), So_d.data_ptr(), n, fac, nf, sums.data_ptr(),
                                cnts.data_ptr(), work.data_ptr(), wbytes, _dev.stream()))
    with np.errstate(invalid="ignore", divide="ignore"):
        return _dev.to_host(sums) / _dev.to_host(cnts)


_KNEED = None


def kneed_available() -> bool:
    """The reference's own third-party dependency decides when it is importable
    (kneedle.knee_index); otherwise the restated Kneedle runs, on the device."""
    global _KNEED
    if _KNEED is None:
        _KNEED = _kneed_importable()
    return _KNEED


def _kneed_importable() -> bool:
    try:
        import kneed  # noqa: F401
        return True
    except ImportError:
        return False


def sweep_knee_apply_device(xa_d, y_d, Sa_d, So_d, factors):
    """Sweep, knee and update queued back to back with no host round trip: returns
    (xb, ak, inc, err, pick, factor, means) as device tensors (pick: int32 scalar)."""
    L = _lib.lib()
    n = Sa_d.numel()
    nf = len(factors)
    sums, cnts, means = _dev.empty((nf,)), _dev.empty((nf,)), _dev.empty((nf,))
    wbytes = int(L.oisat_oi_sweep_workspace(n, nf))
    work = _dev.empty(((wbytes + 7) // 8,))
    fac = (C.c_double * nf)(*[float(f) for f in factors])
    s = _dev.stream()
    _lib.check(L.oisat_oi_sweep(Sa_d.data_ptr(), So_d.data_ptr(), n, fac, nf, sums.data_ptr(),
                                cnts.data_ptr(), work.data_ptr(), wbytes, s))
    pick = _dev.empty((1,), "int32")
    factor = _dev.empty((1,))
    _lib.check(L.oisat_oi_knee(fac, nf, sums.data_ptr(), cnts.data_ptr(), pick.data_ptr(),
                               factor.data_ptr(), means.data_ptr(), s))
    outs = [_dev.empty((n,)) for _ in range(4)]
    _lib.check(L.oisat_oi_apply_dev(xa_d.data_ptr(), y_d.data_ptr(), Sa_d.data_ptr(), So_d.data_ptr(),
                                    n, factor.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                                    outs[2].data_ptr(), outs[3].data_ptr(), s))
    return outs + [pick, factor, means]


def apply_device(xa_d, y_d, Sa_d, So_d, factor):
    L = _lib.lib()
    n = xa_d.numel()
    outs = [_dev.empty((n,)) for _ in range(4)]
    _lib.check(L.oisat_oi_apply(xa_d.data_ptr(), y_d.data_ptr(), Sa_d.data_ptr(), So_d.data_ptr(),
                                n, float(factor), outs[0].data_ptr(), outs[1].data_ptr(),
                                outs[2].data_ptr(), outs[3].data_ptr(), _dev.stream()))
    return outs


def OI(Xa: np.ndarray, Y: np.ndarray, Sa: np.ndarray, So: np.ndarray, regularization_on=True):
    _dev.require_cuda()
    Y[Y < 0] = 0.0  # in place, like the reference: the caller's array is clipped too
    shape = np.shape(Xa)
    dev = [_dev.to_device(np.asarray(a, dtype=np.float64).ravel()) for a in (Xa, Y, Sa, So)]
    factors = regularisation_factors(regularization_on)
    pick = 0
    if regularization_on == True:  # noqa: E712
        means = sweep_device(dev[2], dev[3], factors)
        pick = knee_index(factors, means)
    xb, ak, inc, err = apply_device(dev[0], dev[1], dev[2], dev[3], float(factors[pick]))
    back = lambda t: _dev.to_host(t).reshape(shape)  # noqa: E731
    return back(xb), back(ak), back(inc), back(err)
