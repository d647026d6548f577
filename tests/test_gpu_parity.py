"""Parity tests proper: the CUDA path, called through the reference-shaped
drop-in functions (which go through the C-ABI), against the oracle on the same
seeded inputs and against the reference's golden fixtures.

Bar (BASELINE.json north_star): bit-exact NaN masks, indices and counts;
float64 fields within 1e-6 relative (util.RTOL_FP64).  Most fields land at
~1e-15; the bar is only approached where the reference itself rounds through
float32 (amf_recal.py:51-56,108,116)."""
import numpy as np
import pytest

import cases
import chains
from util import RTOL_FP64, assert_field, max_rel

pytestmark = pytest.mark.gpu


def _compare(store, want, tight=()):
    keys = [k for k in want if k != "input_sha256"]
    assert set(keys) == set(store), sorted(set(keys) ^ set(store))
    worst = {}
    for k in keys:
        rtol = 1e-12 if k.startswith(tight) else RTOL_FP64
        scale = None
        if k.endswith(("increment_OI", ".inc")):
            # the prior of the OI: the model column, or aux2 for GOSAT (driver.py:113-114),
            # whose ctm_vcd is NaN by design (ak_conv_gosat.py:138)
            scale = want["avg.ctm_vcd"]
            if np.isnan(scale).all():
                scale = want["avg.aux2"]
        assert_field(store[k], want[k], k, rtol=rtol, scale=scale)
        worst[k.split(".")[0]] = max(worst.get(k.split(".")[0], 0.0), max_rel(store[k], want[k]))
    return worst


@pytest.mark.parametrize("name", list(cases.CASES))
def test_amf_chain_vs_oracle_and_golden(name, golden):
    got, _ = chains.amf_chain(chains.cuda_impl(), name)
    want, _ = chains.amf_chain(chains.oracle_impl(), name)
    # gridding is float64 in a fixed order: far tighter than the bar
    worst = _compare(got, want, tight=("interp",))
    print(name, "max rel err vs oracle:", worst)
    _compare(got, golden(name), tight=("interp",))


def test_mopitt_chain_vs_oracle_and_golden(golden):
    got, _ = chains.mopitt_chain(chains.cuda_impl())
    want, _ = chains.mopitt_chain(chains.oracle_impl())
    print("mopitt", _compare(got, want, tight=("interp",)))
    _compare(got, golden("mopitt_co"), tight=("interp",))


def test_gosat_chain_vs_oracle_and_golden(golden):
    got, _ = chains.gosat_chain(chains.cuda_impl())
    want, _ = chains.gosat_chain(chains.oracle_impl())
    print("gosat", _compare(got, want, tight=("interp", "fill")))
    _compare(got, golden("gosat_xch4"), tight=("interp", "fill"))


def test_o3_chain_vs_oracle_and_golden(golden):
    """Granules without scattering weights: oisat_vertical_column (amf_recal.py:160-171), the
    Dobson-unit conversion of average() and OI on the result, through the oisatgmi class."""
    got, _ = chains.o3_chain(chains.cuda_impl())
    want, _ = chains.o3_chain(chains.oracle_impl())
    print("o3", _compare(got, want, tight=("interp",)))
    _compare(got, golden("omi_o3"), tight=("interp",))


@pytest.mark.parametrize("fine", [False, True])
def test_ssmis_chain_vs_oracle_and_golden(fine, golden):
    """SSMIS water vapour: reader front-end (bit for bit), interpolator_ssmis (two Delaunay-linear
    steps with the box mean between them), pwv_calculator (model precipitable water, on the
    model grid or resampled to the mesh), then the month through the oisatgmi class."""
    got, _ = chains.ssmis_chain(chains.cuda_impl(), fine)
    want, _ = chains.ssmis_chain(chains.oracle_impl(), fine)
    for k in ("read.vcd", "read.uncertainty", "read.latitude_center", "read.longitude_center"):
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k], equal_nan=True), k
    assert got["pwv0.ctm_vcd"].dtype == want["pwv0.ctm_vcd"].dtype
    print("ssmis", fine, _compare(got, want, tight=("interp", "read")))
    _compare(got, golden("ssmis_pwv_fine" if fine else "ssmis_pwv"), tight=("interp", "read"))


def test_distance_predicate_is_bit_exact():
    """K0 against scipy's cKDTree distances, including nodes that sit within one
    ulp of the threshold."""
    import emu
    from oisatgmi_b200 import _dev, plan
    c = cases.amf_case("omi_no2")
    gpl = plan.grid_plan(c["coords"], 0.25)
    g = c["granules"][0]
    # move some pixels so that their distance to a node is exactly / almost the radius
    lon = np.array(g.longitude_center, dtype=np.float64).ravel()
    lat = np.array(g.latitude_center, dtype=np.float64).ravel()
    X, Y = gpl.mesh()
    for i, (dx, dy) in enumerate([(0.5, 0.0), (0.3, 0.4), (np.nextafter(0.5, 1), 0.0),
                                  (np.nextafter(0.5, 0), 0.0), (0.0, -0.5)]):
        lon[i], lat[i] = X[5 + i, 7] + dx, Y[5 + i, 7] + dy
    want = emu.distmask(lon, lat, gpl, 0.5)
    got = _dev.to_host(plan.distance_mask(_dev.to_device(lon), _dev.to_device(lat), gpl, 0.5))
    assert np.array_equal(got.astype(bool), want.ravel())
    lon32, lat32 = lon.astype(np.float32), lat.astype(np.float32)
    want = emu.distmask(lon32, lat32, gpl, 0.5)
    got = _dev.to_host(plan.distance_mask(_dev.to_device(lon32), _dev.to_device(lat32), gpl, 0.5))
    assert np.array_equal(got.astype(bool), want.ravel())


def test_nearest_pixel_matches_kdtree_query():
    """oisat_nearest_pixel against cKDTree.query: the same pixel index for every node
    inside the radius, INT32_MAX outside; both output layouts of the stencil."""
    from scipy.spatial import cKDTree
    from oisatgmi_b200 import _dev, plan
    c = cases.amf_case("omi_kdtree")
    radius = 2 * c["grid_size"]
    for region_coords in (c["coords"], None):
        if region_coords is None:   # a model finer than the working mesh: output on the mesh
            lat = np.arange(30.0, 50.0, 0.1)
            lon = np.arange(-105.0, -75.0, 0.1)
            X, Y = np.meshgrid(lon, lat)
            region_coords = {"Latitude": Y, "Longitude": X}
        gpl = plan.grid_plan(region_coords, c["grid_size"])
        g = c["granules"][0]
        lon = np.asarray(g.longitude_center)
        lat = np.asarray(g.latitude_center)
        gp = plan.nearest_plan(lon, lat, gpl, radius)
        X, Y = gpl.mesh()
        pts = np.column_stack((lon.ravel().astype(np.float64), lat.ravel().astype(np.float64)))
        d, idx = cKDTree(pts).query(np.column_stack((X.ravel(), Y.ravel())))
        inside = ~(d > radius)
        if gpl.upscale:
            valid = gpl.nn_ok & inside[gpl.window].all(axis=1)
            cells = np.flatnonzero(valid)
            want = idx[gpl.window[cells]]                     # (n, nwin)
        else:
            cells = np.flatnonzero(inside)
            want = idx[cells][:, None]
        assert np.array_equal(gp.cells, cells)
        vert = gp.vert.reshape(gp.nwin, 3, -1)                # stencil-major (3*nwin, n)
        assert np.array_equal(vert[:, 0, :].T, want)
        assert np.array_equal(vert[:, 1, :], vert[:, 0, :]) and np.array_equal(vert[:, 2, :], vert[:, 0, :])
        w = gp.w.reshape(gp.nwin, 3, -1)
        assert np.all(w[:, 0, :] == 1.0) and np.all(w[:, 1:, :] == 0.0)


def test_oi_knee_index_and_means_are_bit_exact():
    """The 99 nanmean(AK_r) follow numpy's pairwise order -> identical floats ->
    identical discrete knee decision."""
    from oisatgmi_b200 import _dev, optimal_interpolation as oi
    from oracle import oi as ooi
    rng = np.random.default_rng(3)
    for shape in [(41, 49), (361, 576), (7,), (129,), (1000,)]:
        xa = np.abs(rng.standard_normal(shape)) * 3 + 0.1
        y = xa * (1 + 0.3 * rng.standard_normal(shape))
        sig = np.abs(rng.standard_normal(shape)) + 0.05
        hole = rng.uniform(size=shape) < 0.3
        xa[hole] = np.nan
        Sa, So = (xa * 0.5) ** 2, sig ** 2
        want = ooi.OI(xa.copy(), y.copy(), Sa, So)
        means = oi.sweep_device(_dev.to_device(Sa.ravel()), _dev.to_device(So.ravel()),
                                oi.regularisation_factors(True))
        assert np.array_equal(means, want[5]), shape
        yy = y.copy()
        got = oi.OI(xa.copy(), yy, Sa, So)
        assert np.array_equal(yy, np.where(y < 0, 0.0, y))     # clipped in place
        for a, b, n in zip(got, want[:4], ["xb", "ak", "inc", "err"]):
            assert_field(a, b, n, rtol=1e-15)


def test_accumulator_matches_numpy_nanmean_bit_for_bit():
    from oisatgmi_b200 import _dev
    from oisatgmi_b200.averaging import MonthAccumulator
    from oracle.averaging import error_averager
    rng = np.random.default_rng(8)
    G, n = 17, 3000
    stack = rng.standard_normal((5, G, n)) * 10 ** rng.uniform(-2, 2, (5, G, n))
    stack[rng.uniform(size=stack.shape) < 0.4] = np.nan
    stack[0, 3, :50] = np.inf                  # inf -> NaN for the satellite column only
    stack[1] = np.abs(stack[1])
    acc = MonthAccumulator(n)
    for g in range(G):
        acc.add(*[_dev.to_device(stack[q, g]) for q in range(5)])
    got = [_dev.to_host(o) for o in acc.finalize()]
    v = stack[0].copy()
    v[np.isinf(v)] = np.nan
    with np.errstate(all="ignore"):
        assert np.array_equal(got[0], np.nanmean(v, axis=0), equal_nan=True)
        for q in (2, 3, 4):
            assert np.array_equal(got[q], np.nanmean(stack[q], axis=0), equal_nan=True)
        want_err = error_averager((stack[1] ** 2)[:, None, :])[0]
    counts = _dev.to_host(acc.acc[5:])
    assert np.array_equal(counts[1], np.sum(np.isfinite(stack[1] ** 2), axis=0))   # exact counts
    assert_field(got[1], want_err, "sat_err", rtol=1e-14)   # pairwise vs sequential sum (A.6)


def test_averaging_squares_narrow_uncertainties_in_their_own_dtype():
    """averaging.py:101 squares the stacked uncertainty grids in THEIR dtype: a float32 (or
    float16) grid is squared in float32 (float16), not in float64.  The drop-in against the
    oracle on grids whose uncertainty is float32."""
    from oisatgmi_b200 import averaging as cavg
    from oracle import averaging as oavg
    _, grids = chains.amf_chain(chains.oracle_impl(), "omi_no2", stop_after="amf")
    for g in grids:
        g.uncertainty = np.asarray(g.uncertainty).astype(np.float32)
    want = oavg.averaging("2005-06-01", "2005-07-01", cases.reader_ns(cases.clone(grids)))
    got = cavg.averaging("2005-06-01", "2005-07-01", cases.reader_ns(cases.clone(grids)))
    for k, name in enumerate(["sat_vcd", "sat_err", "ctm_vcd", "aux1", "aux2"]):
        assert_field(got[k], want[k], name, rtol=1e-14)
    # squaring in float64 instead would be off by ~1e-8 relative: make sure that is not what runs
    wide = [cases.clone(g) for g in grids]
    for g in wide:
        g.uncertainty = np.asarray(g.uncertainty).astype(np.float64)
    other = oavg.averaging("2005-06-01", "2005-07-01", cases.reader_ns(wide))
    f = np.isfinite(want[1])
    assert np.max(np.abs(other[1][f] - want[1][f]) / want[1][f]) > 1e-10


def test_constant_field_and_unit_weights_properties():
    """Known answers: a constant field grids to the same constant; SW == 1 gives
    AMF == 1 up to the float32 rounding of the reference's column sum
    (amf_recal.py:116: the denominator is a float32 nansum, the numerator a
    float64 one); So -> inf returns the prior."""
    from oisatgmi_b200 import amf_recal, interpolator, optimal_interpolation as oi
    c = cases.amf_case("omi_no2")
    g = cases.clone(c["granules"][0])
    g.vcd[...] = np.float16(2.5)
    g.scattering_weights[...] = np.float16(1.0)
    r = interpolator.interpolator(1, 0.25, g, c["coords"], flag_thresh=0.0)
    f = np.isfinite(r.vcd)
    assert f.sum() > 200 and np.max(np.abs(r.vcd[f] - 2.5)) < 1e-14
    assert np.max(np.abs(r.scattering_weights[:, f] - 1.0)) < 1e-14
    out = amf_recal.amf_recal(c["ctm"], [r])[0]
    f = np.isfinite(out.new_amf)
    assert f.sum() > 200 and np.max(np.abs(out.new_amf[f] - 1.0)) < 5e-7
    xa = np.full((5, 5), 3.0)
    res = oi.OI(xa.copy(), xa * 2, (xa * 0.5) ** 2, np.full((5, 5), np.inf), regularization_on=False)
    assert np.array_equal(res[0], xa)


def test_deterministic_run_to_run():
    a, _ = chains.amf_chain(chains.cuda_impl(), "omi_hcho", stop_after="amf")
    b, _ = chains.amf_chain(chains.cuda_impl(), "omi_hcho", stop_after="amf")
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_empty_and_all_masked_granules():
    from oisatgmi_b200 import interpolator
    c = cases.amf_case("omi_no2")
    g = cases.clone(c["granules"][0])
    g.quality_flag[...] = -100.0                      # everything masked -> all NaN -> None
    assert interpolator.interpolator(1, 0.25, g, c["coords"], flag_thresh=0.0) is None
    g = cases.clone(c["granules"][0])
    g.longitude_center = g.longitude_center + np.float32(150.0)   # off the regional grid
    assert interpolator.interpolator(1, 0.25, g, c["coords"], flag_thresh=0.0) is None
    with pytest.raises(Exception):
        interpolator.interpolator(5, 0.25, c["granules"][0], c["coords"])


@pytest.mark.parametrize("delaunay", ["device", "host"])
@pytest.mark.parametrize("name", ["omi_no2", "tropomi_no2"])
def test_gpu_plan_builder_equals_host_plan(name, delaunay, monkeypatch):
    """Builder v1 (native Delaunay -- finished on the device by K12, or built on the host --
    + K1 point location on the GPU) against builder v0 (Qhull + scipy's walk, what the
    reference itself runs): same kept cells, same triangle per window node, weights to
    rounding."""
    from oisatgmi_b200 import plan
    monkeypatch.setenv("OISAT_DELAUNAY", delaunay)
    c = cases.amf_case(name)
    gpl = plan.grid_plan(c["coords"], c["grid_size"])
    for g in c["granules"]:
        monkeypatch.setenv("OISAT_PLAN", "v0")
        p0 = plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 2 * c["grid_size"],
                               cache=False)
        monkeypatch.setenv("OISAT_PLAN", "auto")
        p1 = plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 2 * c["grid_size"],
                               cache=False)
        assert p0.builder == "v0" and p1.builder == ("v1d" if delaunay == "device" else "v1")
        assert np.array_equal(p0.cells, p1.cells)                       # bit-exact masks
        S, n = p0.vert.shape
        # same triangle per node (vertex order inside a triangle is free)
        v0 = np.sort(p0.vert.reshape(S // 3, 3, n), axis=1)
        v1 = np.sort(p1.vert.reshape(S // 3, 3, n), axis=1)
        assert np.array_equal(v0, v1)
        o0 = np.argsort(p0.vert.reshape(S // 3, 3, n), axis=1)
        o1 = np.argsort(p1.vert.reshape(S // 3, 3, n), axis=1)
        w0 = np.take_along_axis(p0.w.reshape(S // 3, 3, n), o0, axis=1)
        w1 = np.take_along_axis(p1.w.reshape(S // 3, 3, n), o1, axis=1)
        assert np.max(np.abs(w0 - w1)) < 1e-12


@pytest.mark.parametrize("product", cases.READER_PRODUCTS)
def test_reader_front_end_is_bit_exact(product, golden):
    """K7 (reader.py:707-983 between the file reads and the interpolator call) against the
    oracle and against the fixture the reference's own reader functions produced: same
    dtypes, same bits -- float16 roundings, flag decoding, -0.0 and NaN patterns included."""
    from oisatgmi_b200 import reader_frontend
    from oracle import reader as oreader
    got = chains.reader_chain(reader_frontend, product)
    chains.same_bits(got, chains.reader_chain(oreader, product))
    chains.same_bits(got, golden("reader_" + product))


def test_output_fields_are_bit_exact():
    """K8 (the data side of driver.write_to_nc, driver.py:156-227) against the oracle:
    float32 casts and the cleaned scaling factor, bit for bit, on the CUDA chain's month."""
    from oisatgmi_b200 import driver
    from oracle import output as ooutput
    obj = chains.month_object(chains.cuda_impl())
    d = driver.oisatgmi()
    d.__dict__.update(obj.__dict__)
    got = d.output_fields()
    want = ooutput.output_fields(obj)
    assert set(got) == set(want)
    for k, a in want.items():
        if k == "time":
            assert got[k] == a
        else:
            assert got[k].dtype == np.float32 and np.array_equal(got[k], a, equal_nan=True), k
    s = got["scaling_factor"]
    assert np.all(s[0, :4] == 1.0)          # x/0, x/NaN, x/inf and 0/x all become 1


def test_near_tie_scan_on_the_device_equals_the_host_scan():
    """oisat_near_ties (device) == count_near_ties inside oisat_h_delaunay_swath (host) on
    a swath without ties, an exactly regular lattice (ties everywhere) and a half-regular
    one; float32 and float64 coordinates."""
    import ctypes as C  # noqa: F401
    from oisatgmi_b200 import _dev, _lib, plan
    from test_host_logic import _adjacency_cases
    L = _lib.lib()
    for name, lon, lat in _adjacency_cases():
        for dt in (np.float64, np.float32):
            lo, la = lon.astype(dt), lat.astype(dt)
            tri, half, hull_ties, maxabs = plan.native_delaunay_adj(lo, la)
            _, total, path = plan.native_delaunay_path(lo, la)
            assert path == 1 and half is not None
            d_tri, d_half = _dev.to_device(tri), _dev.to_device(half)
            d_lo, d_la = _dev.to_device(lo.ravel()), _dev.to_device(la.ravel())
            out = _dev.empty((1,), "int64")
            flag = _dev.zeros((tri.shape[0],), "uint8")
            _lib.check(L.oisat_near_ties(d_tri.data_ptr(), d_half.data_ptr(), tri.shape[0],
                                         d_lo.data_ptr(), d_la.data_ptr(), _dev.dtype_code(d_lo),
                                         float(maxabs), out.data_ptr(), flag.data_ptr(), _dev.stream()))
            assert hull_ties + int(out.item()) == total, (name, dt)
            n_flag = int(flag.sum().item())
            assert (n_flag == 0) == (int(out.item()) == 0) and n_flag <= 2 * int(out.item())
            # nodes "located" in flagged triangles are counted; INT32_MAX = not located
            node_tri = _dev.full((tri.shape[0] + 5,), 2 ** 31 - 1, "int32")
            node_tri[:tri.shape[0]] = _dev.torch().arange(tri.shape[0], dtype=_dev.torch().int32,
                                                          device=node_tri.device)
            cnt = _dev.empty((1,), "int64")
            _lib.check(L.oisat_flagged_nodes(node_tri.data_ptr(), node_tri.numel(), flag.data_ptr(),
                                             cnt.data_ptr(), _dev.stream()))
            assert int(cnt.item()) == n_flag


def test_device_triangulation_equals_the_host_builders(monkeypatch):
    """K12 (oisat_seed_assemble + oisat_flip_delaunay): the assembled seed is the host
    assembly bit for bit; after the rounds, `tri` / `half` are bit-identical to the serial
    replay of the same rounds on the host (the set of flips of a round does not depend on
    the order the threads ran in), no edge is left undecided, and the triangle set is the
    incremental builder's; float32 and float64 coordinates; with the grid-wide rounds only,
    with block 0's tail from the start, and the default mix."""
    from oisatgmi_b200 import _dev, _lib, plan
    from test_host_logic import _seed_swaths, _seed_whole, _flip_rounds, _tri_set
    L = _lib.lib()
    t = _dev.torch()
    done = 0
    meshes = []
    for k, (lon, lat) in enumerate(_seed_swaths()):
        for dt in (np.float64, np.float32):
            lo, la = lon.astype(dt), lat.astype(dt)
            parts = plan.native_seed_parts(lo, la)
            if parts is None:
                continue
            seed = _seed_whole(lo.astype(np.float64), la.astype(np.float64))
            d_lo, d_la = _dev.to_device(lo.ravel()), _dev.to_device(la.ravel())
            # the seed alone
            nt = parts["n_tri"]
            d_tri, d_half = _dev.empty((nt, 3), "int32"), _dev.empty((nt, 3), "int32")
            q, ot, oh = (_dev.to_device(np.ascontiguousarray(parts[n])) for n in ("qtri", "otri", "ohalf"))
            _lib.check(L.oisat_seed_assemble(q.data_ptr(), parts["rows"], parts["cols"], parts["sigma"],
                                             parts["n_quads"], ot.data_ptr(), oh.data_ptr(),
                                             parts["n_outside"], d_tri.data_ptr(), d_half.data_ptr(),
                                             _dev.stream()))
            assert np.array_equal(d_tri.cpu().numpy(), seed[0]) and np.array_equal(d_half.cpu().numpy(), seed[1])
            h_tri, h_half = seed[0].copy(), seed[1].copy()
            want = _flip_rounds(lo.astype(np.float64), la.astype(np.float64), h_tri, h_half)
            ref, ties, path = plan.native_delaunay_path(lo, la)
            for tail in ("0", "1000000000", None):
                if tail is None:
                    monkeypatch.delenv("OISAT_FLIP_TAIL", raising=False)
                else:
                    monkeypatch.setenv("OISAT_FLIP_TAIL", tail)
                tri, half, res, _keep = plan.device_triangulation(parts, (d_lo, d_la))
                t.cuda.synchronize()
                res = res.cpu().numpy()
                assert list(res) == list(want), (k, dt, tail, res, want)
                meshes.append((parts, tri, half, (d_lo, d_la), h_tri, h_half)) if tail is None else None
                assert res[2] == 0 and res[3] == 0
                assert np.array_equal(tri.cpu().numpy(), h_tri) and np.array_equal(half.cpu().numpy(), h_half)
            assert _tri_set(tri.cpu().numpy()) == _tri_set(ref)
            done += 1
    assert done >= 12
    # the same meshes as ONE batch (per coordinate dtype): every mesh ends as it did alone
    for code in (np.float64, np.float32):
        group = [m for m in meshes if m[3][0].cpu().numpy().dtype == code]
        fresh = [plan.seed_assemble_device(m[0]) for m in group]
        result, _work = plan.flip_batch_device([(f[0], f[1], m[3]) for f, m in zip(fresh, group)])
        t.cuda.synchronize()
        result = result.cpu().numpy()
        assert (result[:, 2:] == 0).all() and len(set(result[:, 0])) == 1 and result[0, 0] < 400
        for f, m in zip(fresh, group):
            assert np.array_equal(f[0].cpu().numpy(), m[4]) and np.array_equal(f[1].cpu().numpy(), m[5])


def test_plans_from_the_device_triangulation_equal_the_host_ones(monkeypatch):
    """granule_plans with OISAT_DELAUNAY=device (K12) against OISAT_DELAUNAY=host (incremental
    builder): same kept cells; the same stencil per cell up to the rotation of a triangle's
    vertices (weights to 1e-12); builder tag 'v1d'.  Full OMI-sized granules, one of them
    crossing the date line."""
    from oisatgmi_b200 import plan
    import synth
    coords = synth.ctm_coordinates()
    gplan = plan.grid_plan(coords, 0.25)
    lons, lats = [], []
    for node in (10.0, 175.0, -120.0):
        lat, lon = synth.swath_geolocation(1644, 60, node_lon_deg=node)[:2]
        lons.append(lon)
        lats.append(lat)
    out = {}
    for mode in ("device", "host"):
        monkeypatch.setenv("OISAT_DELAUNAY", mode)
        plan.clear_caches()
        gplan = plan.grid_plan(coords, 0.25)
        out[mode] = plan.granule_plans(lons, lats, gplan, 2 * 0.25)
    for a, b in zip(out["device"], out["host"]):
        assert a is not None and b is not None
        assert a.builder == "v1d" and b.builder == "v1", (a.builder, b.builder)
        assert a.flips > 1000 and a.flip_rounds < 400
        assert np.array_equal(a.cells, b.cells)
        # host accessors are stencil-major (3 * nwin, n_cells), entry 3 k + j = vertex j of window node k
        va, wa = a.vert.T.reshape(len(a.cells), -1, 3), a.w.T.reshape(len(a.cells), -1, 3)
        vb, wb = b.vert.T.reshape(len(b.cells), -1, 3), b.w.T.reshape(len(b.cells), -1, 3)
        ia, ib = np.argsort(va, axis=2, kind="stable"), np.argsort(vb, axis=2, kind="stable")
        va, wa = np.take_along_axis(va, ia, 2), np.take_along_axis(wa, ia, 2)
        vb, wb = np.take_along_axis(vb, ib, 2), np.take_along_axis(wb, ib, 2)
        same = (va == vb).all(axis=2)
        # a mesh node exactly on an edge may sit in either triangle: its weight for the
        # vertex that differs is 0 within rounding
        assert same.mean() > 0.9999
        np.testing.assert_allclose(wa[same], wb[same], rtol=0, atol=1e-12)


def test_knee_on_the_device_equals_the_host_kneedle():
    """oisat_oi_knee (one thread on the device) against kneedle.knee_index on saturating
    curves with and without wiggles (several local maxima of the difference curve), curves
    without a knee, NaNs; and the pipeline gives the same index either way."""
    import ctypes as C
    from oisatgmi_b200 import _dev, _lib, kneedle
    L = _lib.lib()
    rng = np.random.default_rng(11)
    x = np.arange(0.1, 10, 0.1)
    fac = (C.c_double * len(x))(*[float(v) for v in x])
    curves = []
    for trial in range(300):
        a = rng.uniform(0.05, 5.0)
        y = x / (x + a) * rng.uniform(0.2, 1.0) + rng.uniform(0, 0.2)
        if trial % 3 == 0:
            y = y + rng.normal(0, 1e-3, y.shape)
        if trial % 7 == 0:
            y = y + rng.normal(0, 3e-2, y.shape)
        curves.append(y)
    curves += [np.full_like(x, np.nan), np.linspace(0, 1, len(x)), x ** 2, np.sqrt(x), -x,
               np.where(np.arange(len(x)) == 40, np.nan, np.sqrt(x))]
    picks = set()
    for y in curves:
        cnt = rng.integers(1, 50, size=len(x)).astype(np.float64)
        sums = y * cnt
        with np.errstate(invalid="ignore"):
            means_host = sums / cnt
        want = kneedle.knee_index(x, means_host)
        pick, factor, means = _dev.empty((1,), "int32"), _dev.empty((1,)), _dev.empty((len(x),))
        d_sums, d_cnt = _dev.to_device(sums), _dev.to_device(cnt)
        _lib.check(L.oisat_oi_knee(fac, len(x), d_sums.data_ptr(), d_cnt.data_ptr(), pick.data_ptr(),
                                   factor.data_ptr(), means.data_ptr(), _dev.stream()))
        assert int(pick.item()) == want
        assert float(factor.item()) == x[want]
        assert np.array_equal(_dev.to_host(means), means_host, equal_nan=True)
        picks.add(want)
    assert len(picks) > 10 and 0 in picks


def test_pipeline_knee_on_device_equals_knee_on_host(monkeypatch):
    from test_gpu_fused import run_pipeline
    _, dev = run_pipeline("omi_hcho")
    monkeypatch.setenv("OISAT_KNEE", "host")
    _, host = run_pipeline("omi_hcho")
    assert dev["knee_index"] == host["knee_index"] and dev["factor"] == host["factor"]
    assert np.array_equal(np.asarray(dev["ak_means"]), np.asarray(host["ak_means"]), equal_nan=True)
    for k in ("ctm_averaged_vcd_corrected", "ak_OI", "increment_OI", "error_OI"):
        assert np.array_equal(dev[k], host[k], equal_nan=True), k
