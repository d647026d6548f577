"""Host-side logic that needs no GPU: geometry-plan composition, the knee
selector, time matching, pair/tile tables, and the C-ABI surface."""
import os
import re
import ctypes

import numpy as np
import pytest

import cases
import emu
from oisatgmi_b200 import _lib, kneedle, plan
import synth
from util import assert_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------- C-ABI surface
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "oisat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oisat_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 20
    h = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    assert set(names) == set(_lib.PROTOTYPES), sorted(set(names) ^ set(_lib.PROTOTYPES))
    L = _lib.lib()   # loading and the calls below need no GPU
    assert L.oisat_abi_version() == 2
    assert L.oisat_pack_record_halfs(47, 0) == 96
    assert L.oisat_pack_record_halfs(35, 1) == 80
    assert L.oisat_oi_sweep_workspace(207936, 99) > 0


def test_bad_arguments_are_reported_not_crashed():
    L = _lib.lib()
    rc = L.oisat_distmask(None, None, _lib.F32, 10, None, 4, None, 4, 0.5, None, None)
    assert rc == -1 and b"null pointer" in L.oisat_last_error()
    with pytest.raises(_lib.OisatError):
        _lib.check(rc)


def test_no_cuda_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oisatgmi_b200 import optimal_interpolation
    a = np.ones((4, 4))
    with pytest.raises(_lib.OisatError):
        optimal_interpolation.OI(a.copy(), a.copy(), a.copy(), a.copy())


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Field) == 40
    # oisat_fused_args: see include/oisat.h; natural alignment, no packing
    assert ctypes.sizeof(_lib.FusedArgs) % 8 == 0
    assert _lib.FusedArgs.box_weight_err.offset == _lib.FusedArgs.box_weight.offset + 8


# ------------------------------------------------------------------ geometry plan
@pytest.mark.parametrize("name", ["omi_no2", "tropomi_no2"])
def test_plan_reproduces_oracle_gridding(name):
    from oracle import interp as ointerp
    c = cases.amf_case(name)
    g = c["granules"][0]
    o = ointerp.interpolator(1, c["grid_size"], cases.clone(g), c["coords"],
                             flag_thresh=c["flag_thresh"])
    gpl = plan.grid_plan(c["coords"], c["grid_size"])
    keep = emu.distmask(g.longitude_center, g.latitude_center, gpl, 2 * c["grid_size"])
    gp = plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 2 * c["grid_size"],
                           keep=keep, cache=False)
    good = (g.quality_flag > c["flag_thresh"]).ravel()
    f64 = lambda a: np.asarray(a).astype(np.float64).ravel()  # noqa: E731
    assert_field(emu.apply_stencil(gp, f64(g.vcd), good), o.vcd, "vcd", rtol=1e-14)
    assert_field(emu.apply_stencil(gp, f64(g.uncertainty ** 2), good, error=True), o.uncertainty,
                 "uncertainty", rtol=1e-14)
    assert_field(emu.apply_stencil(gp, f64(g.scattering_weights[3]), good),
                 o.scattering_weights[3], "sw[3]", rtol=1e-14)
    # restricted point location (only nodes K0 keeps) gives the same plan as the full walk
    old = plan.FULL_WALK_MAX_NODES
    try:
        plan.FULL_WALK_MAX_NODES = 0
        gp2 = plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 2 * c["grid_size"],
                                keep=keep, cache=False)
    finally:
        plan.FULL_WALK_MAX_NODES = old
    assert np.array_equal(gp.cells, gp2.cells) and np.array_equal(gp.vert, gp2.vert)
    assert np.array_equal(gp.w, gp2.w)


def test_box_window_matches_convolve2d():
    from scipy import signal
    rng = np.random.default_rng(0)
    Z = rng.standard_normal((9, 11))
    for ky, kx in [(2, 2), (5, 6), (1, 3), (1, 1)]:
        ref = signal.convolve2d(Z, np.ones((ky, kx)) / (ky * kx), boundary="symm", mode="same")
        node = np.arange(Z.size)
        win = plan.box_window(9, 11, node, ky, kx)
        got = Z.ravel()[win].mean(axis=1).reshape(Z.shape)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)


def test_degenerate_granule_returns_none():
    gpl = plan.grid_plan(cases.coords(), 0.25)
    lon = np.linspace(-100, -90, 12).astype(np.float32)   # collinear: Qhull raises
    lat = np.full(12, 40.0, np.float32)
    keep = emu.distmask(lon, lat, gpl, 0.5)
    assert plan.granule_plan(lon, lat, gpl, 0.5, keep=keep, cache=False) is None


# ------------------------------------------------------------------------- knee
def test_kneedle_matches_oracle_restatement():
    from oracle.kneedle import KneeLocator
    rng = np.random.default_rng(5)
    x = np.arange(0.1, 10, 0.1)
    for trial in range(200):
        a = rng.uniform(0.05, 5.0)
        y = x / (x + a) * rng.uniform(0.2, 1.0) + rng.uniform(0, 0.2)
        if trial % 3 == 0:
            y = y + rng.normal(0, 1e-3, y.shape)      # wiggles: several local maxima
        want = KneeLocator(x, y, direction="increasing").knee
        got = kneedle.knee_increasing_concave(x, y)
        assert (want is None and got is None) or want == got
        idx = kneedle.knee_index(x, y)
        assert idx == (0 if want is None else int(np.argwhere(x == want)[0][0]))
    assert kneedle.knee_increasing_concave(x, np.full_like(x, np.nan)) is None


# ----------------------------------------------------------------- time matching
def test_time_matching_follows_reference_rules():
    import datetime as dt
    from oisatgmi_b200 import _vertical as v
    ctm = cases.ctm()
    stamps, fracs = v.ctm_clock(ctm)
    k, day, hour = v.closest_slot(ctm, stamps, fracs, dt.datetime(2005, 6, 9, 13, 40))
    assert (day, hour) == (0, 4)      # slots at 01:30, 04:30, ... -> 13:30 is index 4
    ctm[0].averaged = False
    k, day, hour = v.closest_slot(ctm, stamps, fracs, dt.datetime(2005, 6, 1, 23, 0))
    assert (day, hour) == (0, 7)
    assert v.closest_day(ctm, stamps, dt.datetime(2005, 6, 20))[1] == 7  # day stamp vs slot stamps


# --------------------------------------------------------- native triangulation
def _tri_set(t):
    return set(map(tuple, np.sort(np.asarray(t), axis=1)))


def test_native_delaunay_equals_qhull_on_general_position_inputs():
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(12)
    cases_xy = []
    for n in (3, 4, 7, 50, 3000):
        cases_xy.append((rng.uniform(-50, 50, n), rng.uniform(-20, 20, n)))
    # swath pieces: the real use (float32 coordinates, date-line wrap included)
    for geo in (synth.regional_geo(cases.REGION), dict(node_lon_deg=172.0, u0_deg=10, u1_deg=25)):
        lat, lon = synth.swath_geolocation(180, 60, rng=rng, **geo)
        cases_xy.append((lon.ravel().astype(np.float64), lat.ravel().astype(np.float64)))
    for x, y in cases_xy:
        tri, ties = plan.native_delaunay(x, y)
        ref = Delaunay(np.column_stack((x, y)))
        assert ties == 0
        assert _tri_set(tri) == _tri_set(ref.simplices), len(x)
    # a tight cluster (1e-3) next to a wide one (5): circumcircle margins fall below
    # Qhull's tolerance, which is relative to the global coordinate range -- Qhull
    # then returns a non-Delaunay diagonal.  The native builder must flag the input
    # (so the product takes Qhull's answer, like the reference) ...
    x = np.concatenate([rng.normal(0, 1e-3, 300), rng.normal(40, 5, 300)])
    y = np.concatenate([rng.normal(0, 1e-3, 300), rng.normal(-10, 5, 300)])
    tri, ties = plan.native_delaunay(x, y)
    assert ties > 0


def test_native_delaunay_reports_ties_and_degenerate_input():
    lon, lat = np.meshgrid(np.arange(10.0), np.arange(8.0))
    tri, ties = plan.native_delaunay(lon.ravel(), lat.ravel())      # co-circular quads everywhere
    assert tri is not None and len(tri) == 2 * 9 * 7 and ties > 0    # valid, but not unique
    tri, ties = plan.native_delaunay(np.arange(12.0), np.full(12, 3.0))   # collinear
    assert tri is None
    tri, ties = plan.native_delaunay(np.array([0.0, 1.0]), np.array([0.0, 1.0]))
    assert tri is None
    # exact duplicates are not vertices (Qhull lists them as coplanar points)
    x = np.array([0.0, 1.0, 0.0, 1.0, 0.3, 0.3])
    y = np.array([0.0, 0.0, 1.0, 1.1, 0.4, 0.4])
    tri, _ = plan.native_delaunay(x, y)
    assert len({int(v) for v in tri.ravel()} & {4, 5}) == 1


def test_native_delaunay_exact_predicates_on_nearly_degenerate_input():
    """Lattice perturbed by ~1 ulp: the floating-point filter cannot decide, the
    exact stage must, and the result must still be THE Delaunay triangulation
    (checked with exact rational arithmetic on every interior edge)."""
    from fractions import Fraction
    rng = np.random.default_rng(4)
    gx, gy = np.meshgrid(np.arange(12.0), np.arange(9.0))
    x = (gx.ravel() + rng.integers(-2, 3, gx.size) * 2.0 ** -50)
    y = (gy.ravel() + rng.integers(-2, 3, gy.size) * 2.0 ** -50)
    tri, ties = plan.native_delaunay(x, y)
    assert tri is not None
    F = lambda v: Fraction(float(v))  # noqa: E731

    def incircle(a, b, c, d):
        rows = []
        for p in (a, b, c):
            dx, dy = F(x[p]) - F(x[d]), F(y[p]) - F(y[d])
            rows.append((dx, dy, dx * dx + dy * dy))
        (a0, a1, a2), (b0, b1, b2), (c0, c1, c2) = rows
        return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0)

    def orient(a, b, c):
        return (F(x[a]) - F(x[c])) * (F(y[b]) - F(y[c])) - (F(y[a]) - F(y[c])) * (F(x[b]) - F(x[c]))

    edges = {}
    for t in tri:
        for k in range(3):
            e = tuple(sorted((int(t[k]), int(t[(k + 1) % 3]))))
            edges.setdefault(e, []).append(int(t[(k + 2) % 3]))
    violations = 0
    for (a, b), opp in edges.items():
        if len(opp) != 2:
            continue
        c, d = opp
        if orient(a, b, c) < 0:
            a, b = b, a
        if incircle(a, b, c, d) > 0:
            violations += 1
    assert violations == 0


def _valid_triangulation(x, y, tri):
    """All triangles non-degenerate and their areas add up to the hull's."""
    from scipy.spatial import ConvexHull
    a, b, c = (tri[:, k] for k in range(3))
    area2 = (x[b] - x[a]) * (y[c] - y[a]) - (y[b] - y[a]) * (x[c] - x[a])
    hull = ConvexHull(np.column_stack((x, y)))
    return bool(np.all(area2 != 0)) and np.isclose(np.abs(area2).sum() / 2, hull.volume, rtol=1e-9)


def test_lattice_delaunay_equals_general_builder():
    """oisat_h_delaunay_swath (2-D lon/lat: coarse-to-fine lattice insertion) must give
    the triangle set of the general builder / Qhull on every kind of swath."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(5)
    swaths = []
    for geo in (synth.regional_geo(cases.REGION),                       # mid-latitude piece
                dict(node_lon_deg=172.0, u0_deg=10, u1_deg=25),         # date-line crossing
                dict(node_lon_deg=30.0, u0_deg=60, u1_deg=120)):        # over the pole: lat turns round
        lat, lon = synth.swath_geolocation(150, 40, rng=rng, **geo)
        swaths.append((lon.astype(np.float64), lat.astype(np.float64)))
    lat, lon = swaths[0][1], swaths[0][0]
    swaths.append((lon[:, ::-1].copy(), lat[:, ::-1].copy()))           # mirrored handedness
    swaths.append((lon.T.copy(), lat.T.copy()))                         # long axis second
    swaths.append((lon[:2], lat[:2]))                                   # two scan lines
    swaths.append((lon[:, :2].copy(), lat[:, :2].copy()))               # two pixels per line
    gx, gy = np.meshgrid(np.arange(23.0), np.arange(17.0))              # jittered regular lattice
    swaths.append((gx + rng.uniform(-0.3, 0.3, gx.shape), gy + rng.uniform(-0.3, 0.3, gy.shape)))
    for lon, lat in swaths:
        tri, ties, path = plan.native_delaunay_path(lon, lat)
        assert path == 1 and ties == 0, (lon.shape, path, ties)
        ref, ties_ref = plan.native_delaunay(lon.ravel(), lat.ravel())
        assert ties_ref == 0
        assert _tri_set(tri) == _tri_set(ref), lon.shape
        assert _tri_set(tri) == _tri_set(Delaunay(np.column_stack((lon.ravel(), lat.ravel()))).simplices)


def test_lattice_delaunay_degenerate_lattices():
    # exactly regular lattice: points land ON edges and hull edges while inserting
    # (2 -> 4 splits, also against ghost triangles); valid, but ties everywhere
    gx, gy = np.meshgrid(np.arange(13.0), np.arange(9.0))
    tri, ties, path = plan.native_delaunay_path(gx, gy)
    assert path == 1 and ties > 0 and len(tri) == 2 * 12 * 8
    assert _valid_triangulation(gx.ravel(), gy.ravel(), tri)
    # rows exactly collinear, columns sheared: still on-edge insertions, no co-circular quads
    sx = gx + 0.375 * gy + 0.015625 * gx * gx        # dyadic: exact in floating point
    tri, ties, path = plan.native_delaunay_path(sx, gy)
    assert path == 1 and _valid_triangulation(sx.ravel(), gy.ravel(), tri)
    ref, _ = plan.native_delaunay(sx.ravel(), gy.ravel())
    if ties == 0:
        assert _tri_set(tri) == _tri_set(ref)
    # repeated points are not vertices
    dx, dy = gx + 0.2 * np.sin(gy), gy + 0.1 * np.cos(3 * gx)
    dx[4, 5], dy[4, 5] = dx[4, 4], dy[4, 4]
    tri, ties, path = plan.native_delaunay_path(dx, dy)
    assert path == 1 and len({4 * 13 + 4, 4 * 13 + 5} & {int(v) for v in tri.ravel()}) == 1
    assert _valid_triangulation(dx.ravel(), dy.ravel(), tri)
    # all points on one line: no triangle at all
    tri, ties, path = plan.native_delaunay_path(gx, 2.0 * gx + 1.0)
    assert tri is None
    # non-finite coordinates: scipy.spatial.Delaunay raises and the reference skips the
    # granule (interpolator.py:152-155), so no triangulation is reported either way
    for bad in (np.nan, np.inf):
        nx = dx.copy()
        nx[3, 7] = bad
        assert plan.native_delaunay_path(nx, dy)[0] is None
        assert plan.native_delaunay(nx.ravel(), dy.ravel())[0] is None


def test_lattice_delaunay_randomised_against_qhull():
    """Seeded sweep over lattice shapes (2 x 2 up to 40 x 23), smooth warps, shears, folds
    (a lattice that doubles back on itself, like a swath over the pole), float32-rounded
    coordinates and strong anisotropy: whenever the builder reports no tie its triangle set
    is Qhull's; in every case the triangulation is valid."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(60):
        rows, cols = int(rng.integers(2, 41)), int(rng.integers(2, 24))
        i, j = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64),
                           indexing="ij")
        ax, ay = rng.uniform(0.05, 3.0, 2)                      # anisotropic spacing
        x = ax * j + rng.uniform(-1.5, 1.5) * i
        y = ay * i + rng.uniform(-0.5, 0.5) * j
        amp = rng.uniform(0.0, 0.45) * min(ax, ay)
        x = x + amp * np.sin(0.9 * i + rng.uniform(0, 6)) * np.cos(0.7 * j)
        y = y + amp * np.cos(0.8 * j + rng.uniform(0, 6))
        if trial % 5 == 0:                                      # fold: rows turn back
            y = np.abs(y - 0.6 * y.max()) + 0.013 * i
        x = x + rng.uniform(-1e-3, 1e-3, x.shape)                # general position
        y = y + rng.uniform(-1e-3, 1e-3, y.shape)
        if trial % 3 == 0:
            x, y = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
        tri, ties, path = plan.native_delaunay_path(x, y)
        assert path == 1 and tri is not None, (trial, rows, cols)
        assert _valid_triangulation(x.ravel(), y.ravel(), tri), (trial, rows, cols)
        if ties == 0:
            ref = Delaunay(np.column_stack((x.ravel(), y.ravel()))).simplices
            if _tri_set(tri) != _tri_set(ref):
                # Qhull may merge/flip near-degenerate facets; the general exact builder is
                # the arbiter then
                gen, gen_ties = plan.native_delaunay(x.ravel(), y.ravel())
                assert gen_ties == 0 and _tri_set(tri) == _tri_set(gen), (trial, rows, cols)
            checked += 1
    assert checked >= 40


def _seed_whole(x, y, flags=0):
    """oisat_h_delaunay_seed on a 2-D lattice: (tri, half, info) or None when it declines."""
    import ctypes as C
    L = _lib.lib()
    xs, ys = np.ascontiguousarray(x.ravel(), dtype=np.float64), np.ascontiguousarray(y.ravel(), dtype=np.float64)
    n = xs.size
    tri, half = np.empty((2 * n, 3), np.int32), np.empty((2 * n, 3), np.int32)
    info, ties = np.zeros(5, np.int64), C.c_int64(0)
    nt = L.oisat_h_delaunay_seed(xs.ctypes.data, ys.ctypes.data, x.shape[0], x.shape[1], tri.ctypes.data,
                                 2 * n, half.ctypes.data, C.byref(ties), flags, info.ctypes.data)
    assert nt >= 0
    return (tri[:nt].copy(), half[:nt].copy(), info) if nt else None


def _flip_rounds(x, y, tri, half, max_rounds=4096):
    L = _lib.lib()
    xs, ys = np.ascontiguousarray(x.ravel(), dtype=np.float64), np.ascontiguousarray(y.ravel(), dtype=np.float64)
    res = np.zeros(4, np.int64)
    assert L.oisat_h_flip_rounds(xs.ctypes.data, ys.ctypes.data, tri.ctypes.data, half.ctypes.data,
                                 len(tri), max_rounds, res.ctypes.data) == 0
    return res


def _assemble_numpy(parts):
    """What oisat_seed_assemble does (k12_flip.cu), thread by thread, in numpy."""
    rows, cols, sg, base = parts["rows"], parts["cols"], parts["sigma"], 6 * parts["n_quads"]
    qc = cols - 1
    nt = parts["n_tri"]
    tri, half = np.full(3 * nt, -1, np.int32), np.full(3 * nt, -1, np.int32)
    q = parts["qtri"].reshape(rows - 1, qc)

    def slot(t, side):
        t0, t1 = 3 * t, 3 * t + 3
        return ((t0, t0 + 2, t1, t1 + 1) if sg > 0 else (t0 + 2, t0, t1 + 2, t1 + 1))[side]
    for r in range(rows - 1):
        for c in range(qc):
            t = q[r, c]
            if t < 0:
                continue
            a = r * cols + c
            b, cc, d = a + 1, a + cols, a + cols + 1
            t0, t1 = 3 * t, 3 * t + 3
            if sg > 0:
                tri[t0:t0 + 6] = (a, b, cc, b, d, cc)
                half[t0 + 1], half[t1 + 2] = t1 + 2, t0 + 1
            else:
                tri[t0:t0 + 6] = (a, cc, b, b, cc, d)
                half[t0 + 1], half[t1] = t1, t0 + 1
            if r > 0 and q[r - 1, c] >= 0:
                half[slot(t, 0)] = slot(q[r - 1, c], 3)
            if c > 0 and q[r, c - 1] >= 0:
                half[slot(t, 1)] = slot(q[r, c - 1], 2)
            if c + 1 < qc and q[r, c + 1] >= 0:
                half[slot(t, 2)] = slot(q[r, c + 1], 1)
            if r + 2 < rows and q[r + 1, c] >= 0:
                half[slot(t, 3)] = slot(q[r + 1, c], 0)
    for j in range(3 * parts["n_outside"]):
        tri[base + j] = parts["otri"][j]
        g = parts["ohalf"][j]
        half[base + j] = g
        if 0 <= g < base:
            half[g] = base + j
    return tri.reshape(-1, 3), half.reshape(-1, 3)


def _seed_swaths():
    rng = np.random.default_rng(5)
    swaths = []
    for geo in (synth.regional_geo(cases.REGION),                       # mid-latitude piece
                dict(node_lon_deg=172.0, u0_deg=10, u1_deg=25),         # date-line crossing
                dict(node_lon_deg=179.0, u0_deg=-40, u1_deg=40),        # running along the date line
                dict(node_lon_deg=30.0, u0_deg=60, u1_deg=120)):        # over the pole: lat turns round
        lat, lon = synth.swath_geolocation(150, 40, rng=rng, **geo)
        swaths.append((lon.astype(np.float64), lat.astype(np.float64)))
    lon, lat = swaths[0]
    swaths.append((lon[:, ::-1].copy(), lat[:, ::-1].copy()))           # mirrored handedness
    swaths.append((lon.T.copy(), lat.T.copy()))                         # long axis second
    swaths.append((lon.astype(np.float32).astype(np.float64), lat.astype(np.float32).astype(np.float64)))
    gx, gy = np.meshgrid(np.arange(23.0), np.arange(17.0))              # jittered regular lattice
    swaths.append((gx + rng.uniform(-0.3, 0.3, gx.shape), gy + rng.uniform(-0.3, 0.3, gy.shape)))
    return swaths


def test_lattice_seed_plus_flip_rounds_is_the_delaunay_triangulation():
    """K12's construction without a device: the seed (oisat_h_delaunay_seed: lattice quads +
    exact triangulation of the seam) is a valid triangulation of the hull with consistent twin
    pointers; the rounds of independent flips (oisat_h_flip_rounds, the device's per-edge code
    run serially) end with no undecided edge and give the incremental builder's triangle set;
    serial Lawson flips (flags bit 0) give the same; the PARTS the device receives assemble
    (numpy model of oisat_seed_assemble) to the very arrays of the host assembly."""
    seeded = 0
    for lon, lat in _seed_swaths():
        ref, ties, path = plan.native_delaunay_path(lon, lat)
        got = _seed_whole(lon, lat)
        if got is None:
            continue
        seeded += 1
        tri, half, info = got
        assert len(tri) == len(ref) and _valid_triangulation(lon.ravel(), lat.ravel(), tri)
        h, t = half.ravel(), tri.ravel()
        inner = np.flatnonzero(h >= 0)
        nxt = lambda e: e - e % 3 + (e + 1) % 3   # noqa: E731
        assert np.array_equal(h[h[inner]], inner)
        assert np.array_equal(t[inner], t[nxt(h[inner])]) and np.array_equal(t[nxt(inner)], t[h[inner]])
        parts = plan.native_seed_parts(lon, lat)
        assert parts is not None and parts["n_tri"] == len(tri)
        atri, ahalf = _assemble_numpy(parts)
        assert np.array_equal(atri, tri) and np.array_equal(ahalf, half)
        res = _flip_rounds(lon, lat, tri, half)
        assert res[2] == 0 and res[3] == 0 and res[0] < 400, res
        assert _tri_set(tri) == _tri_set(ref), lon.shape
        h = half.ravel()
        inner = np.flatnonzero(h >= 0)
        assert np.array_equal(h[h[inner]], inner)
        law = _seed_whole(lon, lat, flags=1)
        assert _tri_set(law[0]) == _tri_set(ref) and abs(law[2][3] - res[1]) <= 0.01 * res[1]
    assert seeded >= 6


def test_lattice_seed_declines_or_is_exact_on_random_lattices():
    """The randomised lattice sweep of the incremental builder's test (warps, shears, folds,
    float32 coordinates): the seed either declines (folds, repeated points) or, with the
    replayed rounds, reproduces the incremental builder's triangulation; an exactly regular
    lattice (co-circular quads everywhere) is reported as undecidable, never as Delaunay."""
    rng = np.random.default_rng(2024)
    done = 0
    for trial in range(40):
        rows, cols = int(rng.integers(2, 41)), int(rng.integers(2, 24))
        i, j = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
        ax, ay = rng.uniform(0.05, 3.0, 2)
        x = ax * j + rng.uniform(-1.5, 1.5) * i
        y = ay * i + rng.uniform(-0.5, 0.5) * j
        amp = rng.uniform(0.0, 0.45) * min(ax, ay)
        x = x + amp * np.sin(0.9 * i + rng.uniform(0, 6)) * np.cos(0.7 * j)
        y = y + amp * np.cos(0.8 * j + rng.uniform(0, 6))
        if trial % 5 == 0:
            y = np.abs(y - 0.6 * y.max()) + 0.013 * i
        x = x + rng.uniform(-1e-3, 1e-3, x.shape)
        y = y + rng.uniform(-1e-3, 1e-3, y.shape)
        got = _seed_whole(x, y)
        if got is None:
            continue
        tri, half, info = got
        assert _valid_triangulation(x.ravel(), y.ravel(), tri), trial
        res = _flip_rounds(x, y, tri, half)
        ref, ties, path = plan.native_delaunay_path(x, y)
        if ties == 0 and res[3] == 0:
            assert res[2] == 0 and _tri_set(tri) == _tri_set(ref), trial
            done += 1
    assert done >= 15
    gx, gy = np.meshgrid(np.arange(13.0), np.arange(9.0))
    got = _seed_whole(gx, gy)
    if got is not None:
        res = _flip_rounds(gx, gy, got[0], got[1])
        assert res[3] > 0 and res[2] == 0
    # a cut-off run says so
    lon, lat = _seed_swaths()[0]
    tri, half, _ = _seed_whole(lon, lat)
    assert _flip_rounds(lon, lat, tri, half, max_rounds=2)[2] > 0


def test_lattice_seed_never_loses_a_point():
    """A pixel that does not sit where the lattice puts it (moved into somebody else's quad: its
    own four quads are folded and dropped) used to vanish from the seed -- all the triangles of
    the seam triangulation around it are labelled "inside" and replaced by lattice quads that do
    not know it.  The construction must decline such lattices (or get them right), and a
    randomised sweep over swath pieces, warped / folded / jittered lattices and displaced or
    repeated pixels must never give a triangulation other than the incremental builder's."""
    rng = np.random.default_rng(12)
    gx, gy = np.meshgrid(np.arange(19.0), np.arange(13.0))
    lon, lat = gx + rng.uniform(-0.3, 0.3, gx.shape), gy + rng.uniform(-0.3, 0.3, gy.shape)
    lon[7, 12] = lon[0, 0] + 0.4                       # row 7, but at the left edge of the map
    got = _seed_whole(lon, lat)
    ref, ties, path = plan.native_delaunay_path(lon, lat)
    assert ties == 0 and len(ref) == 2 * lon.size - 2 - (len(ref) * 3 - 2 * _interior_edges(ref))
    if got is not None:
        res = _flip_rounds(lon, lat, got[0], got[1])
        assert res[3] == 0 and _tri_set(got[0]) == _tri_set(ref)
    checked = declined = 0
    for trial in range(400):
        kind = trial % 4
        if kind == 0:
            lat, lon = synth.swath_geolocation(int(rng.integers(2, 120)), int(rng.integers(2, 40)),
                                               node_lon_deg=float(rng.uniform(-180, 180)),
                                               u0_deg=float(rng.uniform(-85, 60)), u1_deg=float(rng.uniform(61, 140)),
                                               rng=rng)
        elif kind == 1:
            u0 = float(rng.uniform(-80, 70))
            lat, lon = synth.swath_geolocation(int(rng.integers(2, 100)), int(rng.integers(2, 40)),
                                               node_lon_deg=float(rng.uniform(160, 200)), u0_deg=u0,
                                               u1_deg=u0 + float(rng.uniform(1, 30)), rng=rng)
        elif kind == 2:
            rows, cols = int(rng.integers(2, 40)), int(rng.integers(2, 30))
            i, j = np.meshgrid(np.arange(rows, dtype=float), np.arange(cols, dtype=float), indexing="ij")
            lon = (rng.uniform(0.05, 3) * j + rng.uniform(-1.5, 1.5) * i
                   + rng.uniform(0, 0.4) * np.sin(0.9 * i) * np.cos(0.7 * j) + rng.uniform(-1e-3, 1e-3, i.shape))
            lat = rng.uniform(0.05, 3) * i + rng.uniform(-0.5, 0.5) * j + rng.uniform(-1e-3, 1e-3, i.shape)
            if rng.random() < 0.3:
                lat = np.abs(lat - 0.6 * lat.max()) + 0.013 * i
        else:
            rows, cols = int(rng.integers(2, 25)), int(rng.integers(2, 25))
            gx, gy = np.meshgrid(np.arange(cols, dtype=float), np.arange(rows, dtype=float))
            lon = gx + rng.uniform(-0.3, 0.3, gx.shape)
            lat = gy + (rng.uniform(-0.3, 0.3, gy.shape) if rng.random() < 0.5 else 0.0)
            if rng.random() < 0.5:                     # a displaced (sometimes repeated) pixel
                r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
                lon[r, c] = lon[0, 0] + (0.0 if rng.random() < 0.3 else 0.37)
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
        if trial % 3 == 0:
            lon, lat = lon.astype(np.float32).astype(np.float64), lat.astype(np.float32).astype(np.float64)
        got = _seed_whole(lon, lat)
        assert (got is None) == (plan.native_seed_parts(lon, lat) is None)
        if got is None:
            declined += 1
            continue
        tri, half, _ = got
        assert _valid_triangulation(lon.ravel(), lat.ravel(), tri), trial
        res = _flip_rounds(lon, lat, tri, half)
        ref, ties, path = plan.native_delaunay_path(lon, lat)
        if res[3] == 0 and ties == 0 and ref is not None:
            assert res[2] == 0 and _tri_set(tri) == _tri_set(ref), (trial, kind, lon.shape)
            checked += 1
    assert checked >= 150 and declined >= 20


def _interior_edges(tri):
    e = np.sort(np.concatenate((tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]])), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    return int((cnt == 2).sum())


def _near_ties_numpy(x, y, tri, half):
    """count_near_ties (csrc/delaunay.cpp) restated with numpy from (tri, half)."""
    x = np.asarray(x, np.float64).ravel()
    y = np.asarray(y, np.float64).ravel()
    t = tri.ravel()
    h = half.ravel()
    a = np.flatnonzero(h > np.arange(h.size))
    b = h[a]
    a0, b0 = a - a % 3, b - b % 3
    p0, pr, pl, p1 = t[a0 + (a + 2) % 3], t[a], t[a0 + (a + 1) % 3], t[b0 + (b + 2) % 3]
    adx, ady, bdx, bdy, cdx, cdy = x[p0] - x[p1], y[p0] - y[p1], x[pl] - x[p1], y[pl] - y[p1], \
        x[pr] - x[p1], y[pr] - y[p1]
    det = (adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) + \
        (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) + \
        (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady)
    area2 = np.abs((x[p0] - x[pr]) * (y[pl] - y[pr]) - (y[p0] - y[pr]) * (x[pl] - x[pr]))
    m = max(np.abs(x).max(), np.abs(y).max())
    return int((np.abs(det) <= 2e-14 * max(m * m, 1e-300) * area2).sum())


def _adjacency_cases():
    rng = np.random.default_rng(9)
    lat, lon = synth.swath_geolocation(120, 30, rng=rng, **synth.regional_geo(cases.REGION))
    gx, gy = np.meshgrid(np.arange(13.0), np.arange(9.0))
    return [("swath", lon.astype(np.float64), lat.astype(np.float64)),
            ("regular", gx, gy),
            ("half-regular", gx + np.where(gy > 4, 0.17 * np.sin(gx), 0.0), gy)]


def test_delaunay_adjacency_output_is_the_twin_map_and_carries_the_tie_report():
    """oisat_h_delaunay_swath_adj: same triangles as oisat_h_delaunay_swath, `half` pairs
    every interior edge with its reversed twin, and hull ties + the near-tie scan over
    (tri, half) equal the complete report of the one-call form."""
    for name, lon, lat in _adjacency_cases():
        tri, half, hull_ties, maxabs = plan.native_delaunay_adj(lon, lat)
        ref, ties_ref, path = plan.native_delaunay_path(lon, lat)
        assert path == 1 and half is not None, name
        assert np.array_equal(tri, ref), name
        t, h = tri.ravel(), half.ravel()
        inner = np.flatnonzero(h >= 0)
        assert np.array_equal(h[h[inner]], inner), name
        nxt = lambda e: e - e % 3 + (e + 1) % 3   # noqa: E731
        assert np.array_equal(t[inner], t[nxt(h[inner])]) and np.array_equal(t[nxt(inner)], t[h[inner]])
        assert (h < 0).sum() >= 3, name
        assert maxabs == max(np.abs(lon).max(), np.abs(lat).max())
        assert hull_ties + _near_ties_numpy(lon, lat, tri, half) == ties_ref, name
    assert ties_ref > 0   # the last case does have ties


def test_granule_plans_routes_ties_and_failures(monkeypatch):
    """Control flow of plan.granule_plans without a GPU (device steps stubbed): a granule with
    exact hull ties goes to builder v0, one whose device scan finds an affected near-tie too,
    an untriangulable one stays None, the rest take the two-phase v1 path; the thread pool is
    the persistent one."""
    import types
    calls = []

    class FakeDev:
        @staticmethod
        def to_device(a, *k, **kw):
            return a

        @staticmethod
        def to_host(a):
            return np.asarray(a)

        @staticmethod
        def device():
            return types.SimpleNamespace(index=0)

    def fake_adj(lon, lat, pinned=False, device_index=None):
        kind = int(lon[0, 0])
        if kind == 3:
            return None, None, 0, 0.0
        return np.zeros((1, 3), np.int32), np.zeros((1, 3), np.int32), (2 if kind == 1 else 0), 1.0

    monkeypatch.setattr(plan, "_dev", FakeDev)
    monkeypatch.setattr(plan, "distance_mask", lambda lo, la, g, r: np.ones(4, np.uint8))
    monkeypatch.setattr(plan, "native_delaunay_adj", fake_adj)
    monkeypatch.setattr(plan, "_plan_v1_enqueue", lambda tri, ll, g, keep, half, m: dict(kind=int(np.ravel(ll[0])[0])))
    monkeypatch.setenv("OISAT_DELAUNAY", "host")
    monkeypatch.setattr(plan, "_kept_cells", lambda st: None)
    monkeypatch.setattr(plan, "_plan_v1_finish", lambda st, g, cells=None: None if st["kind"] == 2 else ("v1", st["kind"]))
    monkeypatch.setattr(plan, "_plan_v0", lambda lon, lat, g, keep: calls.append(int(lon[0, 0])) or ("v0", int(lon[0, 0])))
    monkeypatch.setenv("OISAT_PLAN", "auto")
    gplan = types.SimpleNamespace(upscale=True)
    kinds = [0, 1, 2, 3, 0, 0]
    lons = [np.full((2, 2), float(k)) for k in kinds]
    out = plan.granule_plans(lons, lons, gplan, 0.5, workers=3)
    assert out == [("v1", 0), ("v0", 1), ("v0", 2), None, ("v1", 0), ("v1", 0)]
    assert sorted(calls) == [1, 2]
    assert plan._plan_pool(3) is plan._plan_pool(3)


def test_cache_keys_see_every_change_of_a_large_array():
    """plan._digest keys the cached grid plans; arrays of a model grid's size are fingerprinted
    at memory speed (word sum + dot product with fixed odd weights) instead of hashed: one
    changed bit, two swapped elements, another dtype or shape must still change the key, equal
    content must not."""
    rng = np.random.default_rng(3)
    a = rng.normal(size=(361, 576))
    b = rng.normal(size=(361, 576)).astype(np.float32)
    d = plan._digest(a, b)
    assert d == plan._digest(a.copy(), b.copy())
    a1 = a.copy()
    a1[200, 300] = np.nextafter(a1[200, 300], np.inf)
    a2 = a.copy()
    a2[5, 6], a2[5, 7] = a[5, 7], a[5, 6]
    b1 = b.copy()
    b1[-1, -1] += np.float32(1.0)              # the odd tail word of a float32 array
    for changed in ((a1, b), (a2, b), (a, b1), (a.reshape(576, 361), b), (a.astype(np.float32), b), (b, a)):
        assert plan._digest(*changed) != d
    small = np.arange(12.0)
    assert plan._digest(small) == plan._digest(small.copy()) != plan._digest(small[::-1])


def test_ext_fields_match_the_oracle_restatement():
    """oisatgmi_b200.ext_output (GEOS ExtData form of the scaling factors, tools/convert2EXT.py)
    against oracle.output.ext_fields, which is pinned to the unmodified script where the
    reference is present."""
    from oisatgmi_b200 import ext_output
    from oracle import output as ooutput
    rng = np.random.default_rng(9)
    lat, lon = np.meshgrid(np.arange(-90.0, 90.5, 0.5), np.arange(-180.0, 180.0, 0.625), indexing="ij")
    f = {"lat": lat.astype(np.float32), "lon": lon.astype(np.float32),
         "scaling_factor": rng.uniform(0.2, 3.0, lat.shape).astype(np.float32)}
    got = ext_output.ext_fields(f, "201912")
    want = ooutput.ext_fields(f["lat"], f["lon"], f["scaling_factor"], 2019, 12)
    assert got["time_units"] == want["time_units"] == "hours since 2019-12-01 00:00:00"
    for k in ("time", "lat", "lon", "SF"):
        assert got[k].dtype == np.float64 and np.array_equal(got[k], want[k]), k
    ones = ext_output.ones_fields(f["lat"], f["lon"], 1990, 1)
    assert ones["SF"].shape == (1,) + lat.shape and np.all(ones["SF"] == 1.0)
