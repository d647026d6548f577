"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimal-interpolation stage.

  OI            /root/reference/oisatgmi/optimal_interpolation.py:6-52
  bias_correct  /root/reference/oisatgmi/driver.py:65-106
  oi (Sa, So)   /root/reference/oisatgmi/driver.py:108-114

Strictly element-wise (diagonal covariances): K = Sa r / (Sa r + So) for 99
regularisation factors r = 0.1 ... 9.9, the factor picked at the knee of
r -> nanmean(AK).  The knee itself comes from oracle/kneedle.py (PARITY
UNPINNED: third-party kneed, see that file); everything else here is PINNED
against the live reference, and `regularization_on=False` is knee-free.
Only tests/, smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import numpy as np

from oracle.kneedle import KneeLocator

BIAS = {("TROPOMI", "NO2"): (0.32, 0.66), ("TROPOMI", "HCHO"): (0.90, 0.59),
        ("OMI", "NO2"): (0.32, 0.63), ("OMI", "HCHO"): (0.821, 0.79)}


def OI(Xa, Y, Sa, So, regularization_on=True):
    Y[Y < 0] = 0.0  # in place: the caller's averaged satellite field is clipped too
    factors = list(np.arange(0.1, 10, 0.1)) if regularization_on == True else [1.0]  # noqa: E712
    gains, post_var, aks, ak_means = [], [], [], []
    for r in factors:
        r = float(r)
        K = Sa * r * (Sa * r + So) ** (-1)
        Sb = (np.ones_like(K) - K) * Sa * r
        AK = np.ones_like(Sb) - Sb / (Sa * r)
        gains.append(K)
        post_var.append(Sb)
        aks.append(AK)
        ak_means.append(np.nanmean(AK.flatten()))
    pick = 0
    if regularization_on == True:  # noqa: E712
        knee = KneeLocator(np.array(factors), np.array(ak_means), direction="increasing").knee
        hit = np.argwhere(np.array(factors) == knee)
        pick = int(hit[0][0]) if np.size(hit) != 0 else 0
    inc = gains[pick] * (Y - Xa)
    return Xa + inc, aks[pick], inc, np.sqrt(post_var[pick]), pick, np.array(ak_means)


def bias_correct(sat_type, gasname, sat_vcd):
    if (sat_type, gasname) in BIAS:
        a, b = BIAS[(sat_type, gasname)]
        return (sat_vcd - a) / b
    return sat_vcd


def oi_from_means(sensor, ctm_vcd, sat_vcd, sat_err, aux1, aux2, error_ctm=50.0):
    if sensor != "GOSAT":
        return OI(ctm_vcd, sat_vcd, (ctm_vcd * error_ctm / 100.0) ** 2, sat_err ** 2)
    return OI(aux2, aux1, (aux2 * error_ctm / 100.0) ** 2, sat_err ** 2)
