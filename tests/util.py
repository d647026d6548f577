import numpy as np

# tolerance stated by BASELINE.json north_star for float64 analysis fields
RTOL_FP64 = 1e-6


def assert_same_mask(a, b, name=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (name, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "%s: NaN masks differ at %d cells" % (name, int((na != nb).sum()))


def assert_field(a, b, name="", rtol=RTOL_FP64, atol=0.0, scale=None):
    """Bit-exact NaN mask, values within rtol (relative to the reference value).
    `scale` (array) widens the yardstick to max(|b|, |scale|): used for the OI
    increment K*(y - x_b), a difference of two analysis-sized numbers whose
    relative error is otherwise unbounded by cancellation."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert_same_mask(a, b, name)
    inf_a, inf_b = np.isinf(a), np.isinf(b)
    assert np.array_equal(inf_a, inf_b) and np.array_equal(a[inf_a], b[inf_b]), name + ": inf differ"
    f = np.isfinite(a) & np.isfinite(b)
    if f.any():
        err = np.abs(a[f] - b[f])
        ref = np.abs(b[f])
        if scale is not None:
            ref = np.maximum(ref, np.abs(np.broadcast_to(np.asarray(scale, np.float64), b.shape)[f]))
        lim = rtol * ref + atol
        worst = float(np.max(err / np.maximum(ref, 1e-300)))
        assert np.all(err <= lim), "%s: max rel err %.3e > %.1e" % (name, worst, rtol)


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    f = np.isfinite(a) & np.isfinite(b)
    if not f.any():
        return 0.0
    return float(np.max(np.abs(a[f] - b[f]) / np.maximum(np.abs(b[f]), 1e-300)))
