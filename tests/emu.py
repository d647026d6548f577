"""numpy emulation of what the K0/K2 kernels compute, used ONLY by the CPU test
suite to check the host-side plan composition without a GPU.  It is not a
fallback: nothing in oisatgmi_b200/ imports it."""
import numpy as np
from scipy.spatial import cKDTree


def distmask(lon, lat, gplan, radius):
    pts = np.column_stack((np.asarray(lon, np.float64).ravel(), np.asarray(lat, np.float64).ravel()))
    X, Y = gplan.mesh()
    d, _ = cKDTree(pts).query(np.stack([X, Y], -1))
    return ~(d > radius)


def apply_stencil(gp, values, good, error=False):
    """values: (n_px,) float64 (already squared in native dtype for error fields)."""
    z = np.where(good, values, np.nan)
    nwin = gp.nwin
    acc = np.zeros(gp.n_cells)
    bw = 1.0 / (nwin * nwin) if error else 1.0 / nwin
    for k in range(nwin):
        fine = np.zeros(gp.n_cells)
        for j in range(3):
            fine = fine + gp.w[3 * k + j] * z[gp.vert[3 * k + j]]
        acc = acc + fine * bw
    out = np.full(int(np.prod(gp.gplan.out_shape)), np.nan)
    out[gp.cells] = np.sqrt(acc) if error else acc
    return out.reshape(gp.gplan.out_shape)
