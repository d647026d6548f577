"""The fused month pipeline (pack -> gather+AMF -> ordered accumulation -> OI)
against the oracle chain interpolator -> amf_recal -> averaging -> bias -> OI and
against the stage-by-stage CUDA drop-ins."""
import numpy as np
import pytest

import cases
import chains
from util import RTOL_FP64, assert_field, max_rel

pytestmark = pytest.mark.gpu

KEYS = {"avg.sat_vcd": None, "avg.sat_err": "sat_averaged_error", "avg.ctm_vcd": "ctm_averaged_vcd",
        "avg.aux1": "aux1", "avg.aux2": "aux2", "oi.y": "sat_averaged_vcd",
        "oi.ctm_averaged_vcd_corrected": "ctm_averaged_vcd_corrected", "oi.ak_OI": "ak_OI",
        "oi.increment_OI": "increment_OI", "oi.error_OI": "error_OI"}


def same_bits_where_defined(a, b):
    """NaN at the same places, identical bits everywhere else.  (The NaNs of a pair with a
    masked pixel are written by oisat_pair_alive in the tile form and produced by arithmetic in
    the other forms: their payload bits are not part of the contract.)"""
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and \
        np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


def run_pipeline(name):
    from oisatgmi_b200.pipeline import MonthPipeline
    c = cases.amf_case(name)
    pipe = MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor=c["sensor"],
                         gas=c["gas"], error_ctm=50.0, interpolator_type=c["kind"])
    for g in c["granules"]:
        assert pipe.add_granule(cases.clone(g))
    res = pipe.results_to_host(pipe.run())
    return pipe, res


@pytest.mark.parametrize("name", [n for n in cases.CASES if n not in cases.FINE_MODEL_SPACING])
def test_fused_pipeline_matches_oracle_chain(name, golden):
    pipe, res = run_pipeline(name)
    want, _ = chains.amf_chain(chains.oracle_impl(), name)
    gold = golden(name)
    worst = 0.0
    for key, attr in KEYS.items():
        if attr is None:
            continue   # the un-corrected satellite mean is overwritten by bias + clip, like driver.py
        scale = want["avg.ctm_vcd"] if "increment" in key else None
        assert_field(res[attr], want[key], key, rtol=RTOL_FP64, scale=scale)
        assert_field(res[attr], gold[key], key + "(golden)", rtol=RTOL_FP64, scale=scale)
        worst = max(worst, max_rel(res[attr], want[key]))
    print(name, "fused vs oracle: worst rel err %.2e, knee index %d" % (worst, res["knee_index"]))
    # discrete outputs: the knee index equals the one the oracle picks
    from oracle import oi as ooi
    xa, y = want["avg.ctm_vcd"], want["oi.y"]
    pick = ooi.OI(np.array(xa), np.array(y), (np.array(xa) * 0.5) ** 2,
                  np.array(want["avg.sat_err"]) ** 2)[4]
    assert res["knee_index"] == pick


def test_fused_pipeline_equals_stagewise_cuda_chain():
    """Same device functions on both paths -> identical monthly means."""
    pipe, res = run_pipeline("omi_no2")
    got, _ = chains.amf_chain(chains.cuda_impl(), "omi_no2")
    for key, attr in KEYS.items():
        if attr is None:
            continue
        assert_field(res[attr], got[key], key, rtol=1e-13,
                     scale=got["avg.ctm_vcd"] if "increment" in key else None)


@pytest.mark.parametrize("name", ["omi_hcho", "tropomi_no2"])
def test_day_batch_equals_granule_by_granule(name):
    """MonthPipeline.add_day (reader arrays in ONE page-locked block per granule, the bulk on a
    copy stream, the day's triangulations finished on the device as a batch) gives the bits of
    add_granule called once per record (pageable arrays, one plan at a time); also with pageable
    arrays handed to add_day; with the host triangulation the fields agree to rounding."""
    import os
    from oisatgmi_b200.pipeline import MonthPipeline, _HostBlock
    c = cases.amf_case(name)

    def pipe():
        return MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor=c["sensor"],
                             gas=c["gas"], error_ctm=50.0, interpolator_type=c["kind"])
    _, want = run_pipeline(name)
    recs = [cases.clone(g) for g in c["granules"]]
    hosts = [MonthPipeline.host_arrays(g, pin=True) for g in recs]
    assert all(isinstance(h, _HostBlock) and h.block.is_pinned() and h.head > 0 for h in hosts)
    for k, a in (("vcd", recs[0].vcd), ("sw", recs[0].scattering_weights)):
        assert np.array_equal(hosts[0][k].numpy(), np.asarray(a).reshape(-1), equal_nan=True)
    variants = [dict(hosts=hosts), dict(hosts=None, pin=False)]
    for kw in variants:
        p = pipe()
        assert p.add_day(recs, **kw) == len(recs)
        got = p.results_to_host(p.run())
        for k, a in want.items():
            if isinstance(a, np.ndarray) and a.dtype.kind == "f":
                assert same_bits_where_defined(a, np.asarray(got[k])), (k, sorted(kw))
        assert got["knee_index"] == want["knee_index"]
    old = os.environ.get("OISAT_DELAUNAY")
    os.environ["OISAT_DELAUNAY"] = "host"
    try:
        p = pipe()
        assert p.add_day(recs, hosts=hosts) == len(recs)
        got = p.results_to_host(p.run())
    finally:
        if old is None:
            del os.environ["OISAT_DELAUNAY"]
        else:
            os.environ["OISAT_DELAUNAY"] = old
    # another triangulation builder numbers and rotates the triangles differently: the three
    # products of a stencil triple are then added in another order (masks stay bit-exact)
    for k, a in want.items():
        if isinstance(a, np.ndarray) and a.dtype.kind == "f" and k != "increment_OI":
            assert_field(np.asarray(got[k]), a, k, rtol=1e-9)


def test_fused_counts_are_exact():
    """Per-cell granule counts (rows 5-9 of the accumulator) against the oracle's
    gridded granules: integers, bit-exact."""
    from oisatgmi_b200 import _dev
    pipe, _ = run_pipeline("omi_hcho")
    _, grids = chains.amf_chain(chains.oracle_impl(), "omi_hcho", stop_after="amf")
    counts = _dev.to_host(pipe._buf["acc"][5:]).reshape((5,) + tuple(pipe.gplan.out_shape))
    want = [sum(np.isfinite(np.where(np.isinf(g.vcd), np.nan, g.vcd)).astype(int) for g in grids),
            sum(np.isfinite(g.uncertainty ** 2).astype(int) for g in grids),
            sum((~np.isnan(g.ctm_vcd)).astype(int) for g in grids),
            sum((~np.isnan(g.new_amf)).astype(int) for g in grids),
            sum((~np.isnan(g.old_amf)).astype(int) for g in grids)]
    for q in range(5):
        assert np.array_equal(counts[q], want[q]), q


def test_pack_roundtrip():
    """Every reader value must be recoverable from the packed records (and
    sigma^2 must be the float16 square)."""
    from oisatgmi_b200 import _dev, _lib
    L = _lib.lib()
    g = cases.amf_case("omi_no2")["granules"][0]
    nlev, n_px = g.pressure_mid.shape[0], g.vcd.size
    R = int(L.oisat_pack_record_halfs(nlev, 1))
    d = lambda a: _dev.to_device(np.ascontiguousarray(a).reshape(-1))  # noqa: E731
    rec = _dev.empty((n_px, R), "float16")
    ins = [d(g.scattering_weights), d(g.pressure_mid), d(g.vcd), d(g.uncertainty), d(g.tropopause)]
    _lib.check(L.oisat_pack_granule(ins[0].data_ptr(), ins[1].data_ptr(), nlev, ins[2].data_ptr(),
                                    ins[3].data_ptr(), ins[4].data_ptr(), n_px, rec.data_ptr(),
                                    _dev.stream()))
    rec = _dev.to_host(rec)
    nchunk = R // 8
    rows = np.concatenate([g.scattering_weights.reshape(nlev, -1), g.pressure_mid.reshape(nlev, -1),
                           g.vcd.reshape(1, -1), (g.uncertainty ** 2).reshape(1, -1),
                           g.tropopause.reshape(1, -1)])
    for r in range(rows.shape[0]):
        slot = (r % nchunk) * 8 + r // nchunk
        assert np.array_equal(rec[:, slot], rows[r], equal_nan=True), r


def test_pipeline_output_fields_match_float32_casts():
    """MonthPipeline.output_fields (K8) == what driver.write_to_nc would store from the
    pipeline's float64 results (driver.py:180-222)."""
    from oisatgmi_b200.pipeline import MonthPipeline
    c = cases.amf_case("omi_hcho")
    pipe = MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor=c["sensor"],
                         gas=c["gas"], error_ctm=50.0)
    for g in c["granules"]:
        assert pipe.add_granule(cases.clone(g))
    dev = pipe.run()
    f32 = pipe.output_fields(dev)
    f64 = pipe.results_to_host(dev)
    pairs = {"sat_averaged_vcd": "sat_averaged_vcd", "ctm_averaged_vcd_prior": "ctm_averaged_vcd",
             "ctm_averaged_vcd_posterior": "ctm_averaged_vcd_corrected",
             "sat_averaged_error": "sat_averaged_error", "ak_OI": "ak_OI", "error_OI": "error_OI",
             "aux1": "aux1", "aux2": "aux2"}
    for name, src in pairs.items():
        assert np.array_equal(f32[name], f64[src].astype(np.float32), equal_nan=True), name
    with np.errstate(all="ignore"):
        s = f64["ctm_averaged_vcd_corrected"] / f64["ctm_averaged_vcd"]
    s[np.isnan(s) | np.isinf(s) | (s == 0.0)] = 1.0
    assert np.array_equal(f32["scaling_factor"], s.astype(np.float32))


@pytest.mark.parametrize("name", ["omi_hcho", "tropomi_no2", "tropomi_nearest"])
def test_no_kernel_writes_outside_its_buffers(name, monkeypatch):
    """Every output buffer of the month pipeline between canary zones (OISAT_GUARD=1):
    records, masked AMF, row buffer, staged values, accumulator, derived model fields.
    The pool offers no compute-sanitizer, so this is the out-of-bounds-write check --
    ragged pixel counts, partial tiles and the last block of every kernel included."""
    monkeypatch.setenv("OISAT_GUARD", "1")
    pipe, res = run_pipeline(name)
    assert pipe.check_guards()
    monkeypatch.setenv("OISAT_FUSED", "single")
    pipe2, res2 = run_pipeline(name)
    assert pipe2.check_guards()
    for k in ("sat_averaged_vcd", "aux1", "ctm_averaged_vcd"):
        assert_field(res2[k], res[k], k, rtol=1e-12)


@pytest.mark.parametrize("name", ["omi_hcho", "omi_no2", "tropomi_no2", "tropomi_nearest",
                                  "omi_no2_kinked"])
def test_tile_form_is_bit_identical_to_split_form(name, monkeypatch):
    """oisat_fused_amf_tile (one launch, gridded columns in shared memory, bisection) and
    oisat_fused_amf_split (two launches, row buffer, merge walk) evaluate the same
    arithmetic in the same order: every staged value, hence every monthly field, is equal
    bit for bit -- NaN patterns included."""
    monkeypatch.setenv("OISAT_GUARD", "1")
    monkeypatch.setenv("OISAT_FUSED", "tile")
    pipe_t, res_t = run_pipeline(name)
    assert pipe_t.fused_form == "tile" and pipe_t.check_guards()
    staged_t = pipe_t._buf["staged"].cpu().numpy()
    monkeypatch.setenv("OISAT_FUSED", "split")
    pipe_s, res_s = run_pipeline(name)
    assert pipe_s.fused_form == "split" and pipe_s.check_guards()
    staged_s = pipe_s._buf["staged"].cpu().numpy()
    assert staged_t.shape == staged_s.shape and staged_t.size > 0
    assert same_bits_where_defined(staged_t, staged_s)
    n_alive = int(pipe_t._buf["n_alive"].item())
    assert 0 < n_alive < staged_t.shape[1]            # the case does have masked pixels
    for k in res_s:
        if isinstance(res_s[k], np.ndarray):
            assert np.array_equal(res_t[k], res_s[k], equal_nan=True), k


@pytest.mark.parametrize("name", ["omi_hcho", "omi_no2", "tropomi_no2", "omi_no2_kinked"])
def test_tile_form_generic_build_equals_specialised_build(name, monkeypatch):
    """The BASELINE products run a build of the tile kernel with their level counts and
    stencil size as compile-time constants; OISAT_TILE_GENERIC=1 forces the run-time build
    every other shape takes.  Same staged values bit for bit."""
    monkeypatch.setenv("OISAT_FUSED", "tile")
    pipe_a, _ = run_pipeline(name)
    a = pipe_a._buf["staged"].cpu().numpy()
    monkeypatch.setenv("OISAT_TILE_GENERIC", "1")
    pipe_b, _ = run_pipeline(name)
    b = pipe_b._buf["staged"].cpu().numpy()
    assert a.size > 0 and same_bits_where_defined(a, b)


@pytest.mark.parametrize("name", ["omi_hcho", "omi_no2", "tropomi_no2", "omi_no2_kinked"])
def test_tile_form_packed_lanes_equal_half_warp_lanes(name, monkeypatch):
    """OISAT_TILE_PACKED=1 deals the (pair, chunk) work of the gather to consecutive threads
    (no idle lanes for records of fewer than 16 chunks); the half-warp-per-pair layout is the
    other build.  Same staged values bit for bit."""
    monkeypatch.setenv("OISAT_FUSED", "tile")
    monkeypatch.setenv("OISAT_TILE_PACKED", "0")
    pipe_a, _ = run_pipeline(name)
    a = pipe_a._buf["staged"].cpu().numpy()
    monkeypatch.setenv("OISAT_TILE_PACKED", "1")
    monkeypatch.setenv("OISAT_GUARD", "1")
    pipe_b, _ = run_pipeline(name)
    assert pipe_b.check_guards()
    b = pipe_b._buf["staged"].cpu().numpy()
    assert a.size > 0 and same_bits_where_defined(a, b)


def test_live_pair_list_is_exactly_the_unmasked_pairs():
    """oisat_pair_alive against numpy: a pair is live iff none of its 3*nwin stencil pixels is
    masked (interpolator.py:126-128); the list holds every live pair once."""
    from oisatgmi_b200 import _dev
    pipe, _ = run_pipeline("omi_no2")
    host, dev = pipe.build_tables()
    vert = _dev.to_host(dev["vert"]).astype(np.int64)            # (n_pairs, S)
    bad = np.concatenate([~(np.asarray(g.host["qflag"].numpy(), dtype=np.float64) > pipe.flag_thresh)
                          for g in pipe.granules])
    px0 = host["px0"][host["gran"]]
    want = ~bad[vert + px0[:, None]].any(axis=1)
    n = int(pipe._buf["n_alive"].item())
    got = np.sort(_dev.to_host(pipe._buf["alive_pairs"])[:n])
    assert np.array_equal(got, np.flatnonzero(want))
    staged = _dev.to_host(pipe._buf["staged"])
    assert np.all(np.isnan(staged[:, ~want]))


def test_month_without_granules_still_finishes():
    """A rank whose share of the month is empty must not raise before the all-reduce (the other
    ranks would wait for it forever): it contributes a zero accumulator block and the month
    finalises to empty fields."""
    from oisatgmi_b200.pipeline import MonthPipeline
    c = cases.amf_case("omi_no2")
    pipe = MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor="OMI", gas="NO2")
    res = pipe.results_to_host(pipe.run())
    for k in ("ctm_averaged_vcd", "sat_averaged_vcd", "sat_averaged_error",
              "ctm_averaged_vcd_corrected"):
        assert res[k].shape == tuple(pipe.gplan.out_shape) and np.all(np.isnan(res[k])), k


OPT_KEYS = {"avg.sat_err": "sat_averaged_error", "avg.ctm_vcd": "ctm_averaged_vcd",
            "oi.ctm_averaged_vcd_corrected": "ctm_averaged_vcd_corrected", "oi.ak_OI": "ak_OI",
            "oi.increment_OI": "increment_OI", "oi.error_OI": "error_OI"}


def test_opt_month_pipeline_on_a_model_coarser_than_the_lattice():
    """MOPITT against a 2 x 2.5 degree model: the gridded granules land on the model grid
    (box mean + nearest node, interpolator.py:64-93) and the AK convolution reads the model
    columns as they are -- OptMonthPipeline's other branch, against the oracle's month and
    against the stage-wise drop-ins."""
    from oisatgmi_b200.opt_pipeline import OptMonthPipeline
    c = cases.mopitt_case(coarse=True)
    pipe = OptMonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], "MOPITT")
    assert pipe.gplan.upscale
    for g in c["granules"]:
        assert pipe.add_granule(cases.clone(g))
    res = pipe.results_to_host(pipe.run())
    want, _ = chains.mopitt_chain(chains.oracle_impl(), coarse=True)
    got, _ = chains.mopitt_chain(chains.cuda_impl(), coarse=True)
    for key, attr in (("avg.sat_err", "sat_averaged_error"), ("avg.ctm_vcd", "ctm_averaged_vcd"),
                      ("avg.aux1", "aux1"), ("avg.aux2", "aux2"), ("oi.y", "sat_averaged_vcd"),
                      ("oi.ctm_averaged_vcd_corrected", "ctm_averaged_vcd_corrected"),
                      ("oi.ak_OI", "ak_OI"), ("oi.error_OI", "error_OI")):
        assert_field(res[attr], want[key], key, rtol=RTOL_FP64)
        assert_field(got[key], want[key], key + " (drop-ins)", rtol=RTOL_FP64)
    assert np.isfinite(res["ctm_averaged_vcd_corrected"]).sum() > 50


@pytest.mark.parametrize("sensor", ["MOPITT", "GOSAT"])
def test_opt_month_pipeline_matches_oracle_month(sensor, golden):
    """The device-resident satellite_opt month (gap filling, gridding, model resampling, AK
    convolution, accumulation, OI -- OI on aux2/aux1 for GOSAT) against the oracle's month
    through the oisatgmi class and against the reference's golden month."""
    from oisatgmi_b200.opt_pipeline import OptMonthPipeline
    c = cases.mopitt_case() if sensor == "MOPITT" else cases.gosat_case()
    pipe = OptMonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor)
    for g in c["granules"]:
        assert pipe.add_granule(cases.clone(g))
    res = pipe.results_to_host(pipe.run())
    chain = chains.mopitt_chain if sensor == "MOPITT" else chains.gosat_chain
    want, _ = chain(chains.oracle_impl())
    gold = golden("mopitt_co" if sensor == "MOPITT" else "gosat_xch4")
    keys = dict(OPT_KEYS)
    # OI clips its Y in place: the satellite mean, or aux1 for GOSAT (driver.py:113-114)
    keys["oi.y"] = "aux1" if sensor == "GOSAT" else "sat_averaged_vcd"
    if sensor == "GOSAT":
        keys["avg.sat_vcd"] = "sat_averaged_vcd"
    else:
        keys["avg.aux1"] = "aux1"
    keys["avg.aux2"] = "aux2"
    prior = want["avg.aux2"] if sensor == "GOSAT" else want["avg.ctm_vcd"]
    for key, attr in keys.items():
        scale = prior if "increment" in key else None
        assert_field(res[attr], want[key], key, rtol=RTOL_FP64, scale=scale)
        assert_field(res[attr], gold[key], key + "(golden)", rtol=RTOL_FP64, scale=scale)
    assert np.isfinite(res["ctm_averaged_vcd_corrected"]).sum() > 100
    # the second run REPLAYS the recorded launches of the granule loop: same bits; and so does a
    # pipeline that never records (OISAT_OPT_REPLAY=0)
    again = pipe.results_to_host(pipe.run())
    assert pipe._program is not None and len(pipe._program[0].calls) >= 4 * len(c["granules"])
    import os
    os.environ["OISAT_OPT_REPLAY"] = "0"
    try:
        plain_pipe = OptMonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor)
        for g in c["granules"]:
            assert plain_pipe.add_granule(cases.clone(g))
        plain = plain_pipe.results_to_host(plain_pipe.run())
        assert plain_pipe._program is None
    finally:
        del os.environ["OISAT_OPT_REPLAY"]
    for k, v in res.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, again[k], equal_nan=True), k
            assert np.array_equal(v, plain[k], equal_nan=True), k
    from oracle import oi as ooi
    pick = ooi.oi_from_means(sensor, np.array(want["avg.ctm_vcd"]), np.array(want["avg.sat_vcd"]),
                             np.array(want["avg.sat_err"]), np.array(want["avg.aux1"]),
                             np.array(want["avg.aux2"]))[4]
    assert res["knee_index"] == pick


@pytest.mark.parametrize("generic", ["0", "1"])
@pytest.mark.parametrize("name", ["omi_hcho", "omi_no2", "tropomi_no2", "omi_no2_kinked",
                                  "tropomi_nearest"])
def test_warp_specialised_form_equals_one_tile_per_block_form(name, generic, monkeypatch):
    """The persistent producer / consumer kernel (default) and the one-tile-per-block kernel
    (OISAT_TILE_WS=0) run the same arithmetic in the same order: identical staged bits, for the
    builds with compile-time geometry and for the run-time build."""
    monkeypatch.setenv("OISAT_FUSED", "tile")
    monkeypatch.setenv("OISAT_TILE_GENERIC", generic)
    monkeypatch.setenv("OISAT_TILE_WS", "0")
    pipe_a, _ = run_pipeline(name)
    a = pipe_a._buf["staged"].cpu().numpy()
    monkeypatch.setenv("OISAT_TILE_WS", "1")
    monkeypatch.setenv("OISAT_GUARD", "1")
    pipe_b, _ = run_pipeline(name)
    assert pipe_b.check_guards()
    b = pipe_b._buf["staged"].cpu().numpy()
    assert a.size > 0 and same_bits_where_defined(a, b)
    # and a second run of the persistent kernel gives the same bits (no race between the roles)
    pipe_c, _ = run_pipeline(name)
    assert same_bits_where_defined(b, pipe_c._buf["staged"].cpu().numpy())


@pytest.mark.parametrize("name", ["omi_hcho", "omi_no2", "tropomi_no2", "omi_no2_kinked"])
def test_bulk_copy_gather_equals_lane_copy_gather(name, monkeypatch):
    """Records fetched by one bulk asynchronous copy per (pair, entry) (TMA engine, the default
    of the packed builds) or by 16-byte cp.async per lane (OISAT_TILE_BULK=0): same bytes in
    shared memory, same arithmetic, identical staged bits (tropomi_no2: six sweeps, so the
    mbarrier goes through six phases)."""
    monkeypatch.setenv("OISAT_FUSED", "tile")
    monkeypatch.setenv("OISAT_TILE_BULK", "0")
    pipe_a, _ = run_pipeline(name)
    a = pipe_a._buf["staged"].cpu().numpy()
    monkeypatch.setenv("OISAT_TILE_BULK", "1")
    monkeypatch.setenv("OISAT_GUARD", "1")
    pipe_b, _ = run_pipeline(name)
    assert pipe_b.check_guards()
    b = pipe_b._buf["staged"].cpu().numpy()
    assert a.size > 0 and same_bits_where_defined(a, b)
