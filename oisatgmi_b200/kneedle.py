"""Host-side knee selection for the OI regularisation sweep.

The reference delegates to the third-party package `kneed`
(`KneeLocator(x, y, direction='increasing').knee`,
/root/reference/oisatgmi/optimal_interpolation.py:37-39; pinned kneed==0.8.3 in
requirements.txt:9).  `kneed` is used when it is importable; otherwise the
Kneedle algorithm (Satopaa et al., 2011) is evaluated here for the one
configuration the reference uses: concave, increasing, S = 1, first knee.
Ninety-nine float64 values -- this is control flow, not a kernel.
"""
from __future__ import annotations

import numpy as np


def _rel_extrema(y, cmp):
    """Indices i where cmp(y[i], y[i-1]) and cmp(y[i], y[i+1]) with the ends
    clipped (scipy.signal.argrelextrema(order=1, mode='clip'))."""
    n = len(y)
    idx = np.arange(n)
    left = y[np.clip(idx - 1, 0, n - 1)]
    right = y[np.clip(idx + 1, 0, n - 1)]
    return np.flatnonzero(cmp(y, left) & cmp(y, right))


def knee_increasing_concave(x, y, S: float = 1.0):
    """The knee abscissa (an element of x) or None."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xn = (x - x.min()) / (x.max() - x.min())
    yn = (y - y.min()) / (y.max() - y.min())
    yd = yn - xn
    maxima = _rel_extrema(yd, np.greater_equal)
    minima = _rel_extrema(yd, np.less_equal)
    if maxima.size == 0:
        return None
    tmx = yd[maxima] - S * np.abs(np.diff(xn).mean())
    is_max = np.zeros(len(x), bool)
    is_max[maxima] = True
    is_min = np.zeros(len(x), bool)
    is_min[minima] = True
    threshold = None
    threshold_index = None
    seen = 0
    for i in range(int(maxima[0]), len(x)):
        if xn[i] == 1.0:
            break
        if is_max[i]:
            threshold = tmx[seen]
            threshold_index = i
            seen += 1
        if is_min[i]:
            threshold = 0.0
        if yd[i + 1] < threshold:
            return x[threshold_index]
    return None


def knee_index(x, y) -> int:
    """Index into x of the knee, 0 when none is found -- the reference's
    fallback (optimal_interpolation.py:39-41)."""
    x = np.asarray(x, dtype=np.float64)
    try:
        from kneed import KneeLocator  # the reference's own dependency, when present
        knee = KneeLocator(x, np.asarray(y), direction="increasing").knee
    except ImportError:
        knee = knee_increasing_concave(x, y)
    hit = np.argwhere(x == knee) if knee is not None else np.empty((0, 1))
    return int(hit[0][0]) if np.size(hit) != 0 else 0
