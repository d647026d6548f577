"""BASELINE.json's largest granule shape (configs[4]: TROPOMI NO2, 4172 x 450 =
1.88 M pixels, 0.10 degree working mesh of 6.5 M nodes, 5 x 6 box => 90 stencil entries
per cell, 34 levels + tropopause) through the fused month pipeline on the global GMI
grid.  The oracle needs minutes per granule at this size, so the checks are the
size-independent properties of the path:

  * scattering weights == 1  =>  recalculated AMF == 1 wherever it is defined
    (amf_recal.py:93-119: a weighted mean of a constant; the reference's float32
    column sum bounds the error at ~1e-7);
  * the gridded vcd of a constant field is that constant, its count per cell is the
    number of granules, and cells outside every granule stay empty;
  * a second granule that is all bad pixels changes nothing;
  * determinism: two runs are bit-identical;
  * the device plan builder (native Delaunay + K1) and the host one (Qhull + scipy's
    walk) keep the same cells.
"""
import copy
import datetime

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _granule(seed=5):
    g = synth.make_amf_granule(seed, "TROPOMI_NO2", geo=dict(node_lon_deg=-60.0, u0_deg=-70.0,
                                                             u1_deg=70.0), bad_fraction=0.2,
                               time=datetime.datetime(2005, 6, 3, 13, 20, 0))
    g.scattering_weights = np.ones_like(g.scattering_weights)
    g.vcd = np.full_like(g.vcd, 3.5)
    return g


def test_tropomi_scale_granule_properties():
    from oisatgmi_b200.pipeline import MonthPipeline
    from oisatgmi_b200 import plan as _plan
    model = [synth.make_ctm(11, synth.ctm_coordinates(None))]
    g = _granule()
    assert np.size(g.latitude_center) == 4172 * 450

    def run(granules):
        pipe = MonthPipeline(model, 0.10, 0.75, sensor="TROPOMI", gas="NO2", error_ctm=50.0)
        for x in granules:
            pipe.add_granule(copy.deepcopy(x))
        return pipe, pipe.results_to_host(pipe.run())

    pipe, res = run([g])
    assert pipe.gplan.nwin == 30 and pipe.gplan.H * pipe.gplan.W == 1801 * 3595
    cells = pipe.granules[0].plan.cells
    assert 5_000 < cells.size < 60_000
    sat = res["sat_averaged_vcd"].ravel()
    amf = res["aux1"].ravel()            # new AMF (averaging.py: aux1)
    finite = np.isfinite(amf)
    assert finite.sum() > 1000
    assert set(np.flatnonzero(finite)) <= set(cells.tolist())
    assert np.max(np.abs(amf[finite] - 1.0)) < 5e-7
    # vcd_new = old_amf * vcd / new_amf, with new_amf == 1 and vcd == 3.5 (exact in float16):
    # the ratio to the gridded old AMF is the constant again
    outside = np.ones(sat.size, bool)
    outside[cells] = False
    assert np.all(np.isnan(sat[outside]))

    # an all-bad second granule contributes nothing
    bad = copy.deepcopy(g)
    bad.quality_flag = np.zeros_like(bad.quality_flag)
    bad.time = g.time + datetime.timedelta(hours=2)
    _, res2 = run([g, bad])
    for k in res:
        a, b = np.asarray(res[k]), np.asarray(res2[k])
        if a.dtype.kind == "f" and a.size > 1:
            assert np.array_equal(a, b, equal_nan=True), k

    # determinism
    _, res3 = run([g])
    for k in res:
        a, b = np.asarray(res[k]), np.asarray(res3[k])
        if a.dtype.kind == "f" and a.size > 1:
            assert np.array_equal(a, b, equal_nan=True), k


def test_tropomi_scale_plan_builders_agree(monkeypatch):
    from oisatgmi_b200 import plan as _plan
    coords = synth.ctm_coordinates(None)
    gpl = _plan.grid_plan(coords, 0.10)
    g = _granule(6)
    monkeypatch.setenv("OISAT_PLAN", "auto")
    p1 = _plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 0.2, cache=False)
    assert p1.builder == "v1d"          # triangulation finished on the device (K12)
    monkeypatch.setenv("OISAT_PLAN", "v0")
    p0 = _plan.granule_plan(g.longitude_center, g.latitude_center, gpl, 0.2, cache=False)
    assert p0.builder == "v0"
    assert np.array_equal(p0.cells, p1.cells)
    S, n = p0.vert.shape
    v0 = np.sort(p0.vert.reshape(S // 3, 3, n), axis=1)
    v1 = np.sort(p1.vert.reshape(S // 3, 3, n), axis=1)
    assert np.array_equal(v0, v1)


def test_tropomi_scale_isolated_near_tie_does_not_need_qhull(monkeypatch):
    """One edge of this 5.6 M-edge triangulation is inside Qhull's tolerance of a
    co-circular quadrilateral.  No kept mesh node lies in that quadrilateral (a pixel
    quadrilateral is smaller than the mesh spacing), so no stencil depends on how it is
    split: builder v1 keeps the exact triangulation (before, the whole granule went back to
    Qhull + scipy's walk, ~20 s) -- and the plan equals the one built on Qhull's answer."""
    from oisatgmi_b200 import plan as _plan
    rng = np.random.default_rng(501)
    lat, lon = synth.swath_geolocation(4172, 450, rng=rng, node_lon_deg=150.0 - 360.0 / 14.6,
                                       alt_km=824.0, half_fov_deg=54.0)
    _, ties, path = _plan.native_delaunay_path(lon, lat)
    assert path == 1 and ties >= 1          # the host's complete report sees the tie
    gpl = _plan.grid_plan(synth.ctm_coordinates(None), 0.10)
    monkeypatch.setenv("OISAT_PLAN", "auto")
    p1 = _plan.granule_plan(lon, lat, gpl, 0.2, cache=False)
    assert p1.builder == "v1d" and p1.near_ties == ties
    monkeypatch.setenv("OISAT_PLAN", "v0")
    p0 = _plan.granule_plan(lon, lat, gpl, 0.2, cache=False)
    assert p0.builder == "v0"
    assert np.array_equal(p0.cells, p1.cells)
    S, n = p0.vert.shape
    assert np.array_equal(np.sort(p0.vert.reshape(S // 3, 3, n), axis=1),
                          np.sort(p1.vert.reshape(S // 3, 3, n), axis=1))
    assert np.allclose(np.sort(p0.w.reshape(S // 3, 3, n), axis=1),
                       np.sort(p1.w.reshape(S // 3, 3, n), axis=1), rtol=0, atol=1e-9)


def test_full_omi_hcho_granule_on_global_grid_vs_oracle():
    """One full-size OMI HCHO granule (1644 x 60 px, 47 levels) whose orbit runs across the
    date line, on the global 361 x 576 GMI grid: VALUES of the CUDA path against the oracle
    (about a minute of CPU), not properties -- through the drop-in functions
    (interpolator -> amf_recal, every gridded field and level) and through the fused month
    pipeline (a one-granule month: the means are the granule's own grids)."""
    import types
    from oisatgmi_b200 import amf_recal as _amf, interpolator as _interp
    from oisatgmi_b200.pipeline import MonthPipeline
    from oracle import interp as ointerp, vertical as overt
    from util import assert_field
    coords = synth.ctm_coordinates(None)
    model = [synth.make_ctm(11, coords, averaged=True)]
    g = synth.make_amf_granule(77, "OMI_HCHO", geo=dict(node_lon_deg=172.0), bad_fraction=0.2,
                               time=datetime.datetime(2005, 6, 9, 1, 40, 0))
    lon = np.asarray(g.longitude_center)
    assert lon.shape == (1644, 60) and lon.max() > 179.0 and lon.min() < -179.0   # both sides
    want = ointerp.interpolator(1, 0.25, copy.deepcopy(g), coords, flag_thresh=0.0)
    got = _interp.interpolator(1, 0.25, copy.deepcopy(g), coords, flag_thresh=0.0)
    assert want is not None and got is not None
    for name in ("vcd", "amf", "uncertainty", "pressure_mid", "scattering_weights"):
        assert_field(getattr(got, name), getattr(want, name), "interp." + name, rtol=1e-12)
    assert np.isfinite(want.vcd).sum() > 10_000
    overt.amf_recal(model, [want])
    _amf.amf_recal(model, [got])
    for name in ("vcd", "ctm_vcd", "new_amf", "old_amf"):
        assert_field(getattr(got, name), getattr(want, name), "amf." + name)
    pipe = MonthPipeline(model, 0.25, 0.0, sensor="OMI", gas="HCHO")
    assert pipe.add_granule(copy.deepcopy(g))
    res = pipe.results_to_host(pipe.run())
    # the month's satellite mean leaves the pipeline bias-corrected (driver.py:65-106, OMI HCHO)
    # and clipped at zero by OI (optimal_interpolation.py:14)
    y = (np.asarray(want.vcd) - 0.821) / 0.79
    y[y < 0] = 0.0
    # (y - a) / b cancels where the column is close to the offset: the yardstick is the column
    assert_field(res["sat_averaged_vcd"], y, "fused.vcd", scale=np.asarray(want.vcd) / 0.79)
    assert_field(res["sat_averaged_error"], want.uncertainty, "fused.sigma")
    assert_field(res["ctm_averaged_vcd"], want.ctm_vcd, "fused.ctm_vcd")
    assert_field(res["aux1"], want.new_amf, "fused.new_amf")
    assert_field(res["aux2"], want.old_amf, "fused.old_amf")


def test_near_tie_with_a_kept_node_takes_qhull_triangles_and_device_walk(monkeypatch):
    """Orbit 9 of rank 5's day in the 8-GPU bench (round 1: the granule that made eight ranks
    wait 0.7 s): its exact triangulation has a near tie whose quadrilateral holds a kept mesh
    node, so Qhull's own split is needed.  The fallback now takes Qhull's TRIANGLES and leaves
    the walk over the 1.04 M mesh nodes to K1 (builder "v0q"); the plan equals the one built by
    scipy's walk on the same triangles (builder v0)."""
    import bench
    from oisatgmi_b200 import plan as _plan
    g = synth.make_amf_granule(5009, "OMI_HCHO", geo=bench.orbit_geo(9, 15), bad_fraction=0.2)
    lon, lat = np.asarray(g.longitude_center), np.asarray(g.latitude_center)
    gpl = _plan.grid_plan(synth.ctm_coordinates(None), 0.25)
    monkeypatch.setenv("OISAT_PLAN", "auto")
    p1 = _plan.granule_plan(lon, lat, gpl, 0.5, cache=False)
    monkeypatch.setenv("OISAT_PLAN", "v0")
    p0 = _plan.granule_plan(lon, lat, gpl, 0.5, cache=False)
    assert p0.builder == "v0"
    if p1.builder in ("v1", "v1d"):
        pytest.skip("this build of the synthetic orbit has no affected near tie")
    assert p1.builder == "v0q"
    assert np.array_equal(p0.cells, p1.cells)
    S, n = p0.vert.shape
    assert np.array_equal(np.sort(p0.vert.reshape(S // 3, 3, n), axis=1),
                          np.sort(p1.vert.reshape(S // 3, 3, n), axis=1))
    assert np.allclose(np.sort(p0.w.reshape(S // 3, 3, n), axis=1),
                       np.sort(p1.w.reshape(S // 3, 3, n), axis=1), rtol=0, atol=1e-9)
