#!/usr/bin/env python
"""Measurement of the widened rows at BASELINE sizes, the CPU
restatement timed beside them.  One JSON line per row on stdout:

    python tests/bench_rows.py > profiles/rNN_rows.jsonl

(It lives under tests/ because it runs the oracle next to the kernels; only tests/,
smoke() and bench.py's CPU legs may do that.)

  f-1  reader front-end (K7), one TROPOMI NO2 granule (4172 x 450 px, 34 levels)
  f-2  nearest-neighbour plan (K0-style scatter + plan fill), one OMI granule
  f-3  output fields (K8), one month on the GMI grid

Device times are CUDA events around the kernels with inputs resident in HBM (median of
5 after 2 warm-ups); `achieved` uses the algorithmic bytes stated per row; the CPU
number is oracle/ (numpy/scipy, the reference's algorithm) on one host core.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from oisatgmi_b200 import _dev, _lib, plan as _plan, reader_frontend as rf  # noqa: E402
import synth  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def gpu_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out))


def tropomi_vars(nt=4172, nxt=450, seed=1):
    import cases
    small = cases.reader_vars("tropomi_no2", seed=seed, nt=64, nxt=16)
    rng = np.random.default_rng(seed)
    shape = (nt, nxt)
    v = dict(small)
    for k, a in small.items():
        a = np.asarray(a)
        if a.shape[:2] == (64, 16):
            v[k] = np.resize(a, shape + a.shape[2:]).astype(a.dtype)
    v["delta_time"] = (np.arange(nt) * 840).astype(np.int64)[:, None]
    v["surface_pressure"] = rng.uniform(60000.0, 103000.0, shape).astype(np.float32)
    return v


def row_reader():
    from oracle import reader as oreader
    v = tropomi_vars()
    n_px = 4172 * 450
    L = 34
    dev = {k: rf._up(a) for k, a in v.items() if np.ndim(a) >= 2 and np.shape(a)[:2] == (4172, 450)}
    lib = _lib.lib()
    f16 = lambda *s: _dev.empty(s, "float16")  # noqa: E731
    out_sw, out_p, out_v, out_u, out_q, out_t = f16(L, n_px), f16(L, n_px), f16(n_px), f16(n_px), \
        f16(n_px), f16(n_px)
    tm5_a = np.concatenate(((np.asarray(v["tm5_constant_a"]) / 100.0)[:, 0], 0), axis=None)
    tm5_b = np.concatenate((np.asarray(v["tm5_constant_b"])[:, 0], 0), axis=None)
    a_d, b_d = _dev.to_device(tm5_a.astype(np.float64)), _dev.to_device(tm5_b.astype(np.float64))
    import ctypes as C
    fac = (C.c_double * 3)(6.02214, 1e19, 1e-15)
    s = _dev.stream()

    def run():
        c = _lib.check
        col = dev["nitrogendioxide_tropospheric_column"]
        c(lib.oisat_reader_scale_f16(col.data_ptr(), 2, n_px, fac, 3, out_v.data_ptr(), s))
        c(lib.oisat_reader_scale_f16(dev["nitrogendioxide_tropospheric_column_precision"].data_ptr(), 2,
                                     n_px, fac, 3, out_u.data_ptr(), s))
        c(lib.oisat_reader_scale_f16(dev["qa_value"].data_ptr(), 2, n_px, fac, 0, out_q.data_ptr(), s))
        c(lib.oisat_reader_pmid(3, a_d.data_ptr(), b_d.data_ptr(), dev["surface_pressure"].data_ptr(), 2,
                                100.0, L, n_px, out_p.data_ptr(), s))
        c(lib.oisat_reader_weights(dev["averaging_kernel"].data_ptr(), 2, 1, L, n_px,
                                   dev["air_mass_factor_total"].data_ptr(), 2, out_sw.data_ptr(), s))
        c(lib.oisat_reader_tropopause(dev["tm5_tropopause_layer_index"].data_ptr(), out_p.data_ptr(), L,
                                      n_px, out_t.data_ptr(), s))

    ms = gpu_ms(run)
    # bytes: file variables read once (float32) + reader-dtype arrays written once (float16)
    b = n_px * (4 * (3 + 1 + 1 + L + 1) + 2 * (3 + 2 * L + 1))
    t0 = time.perf_counter()
    oreader.tropomi_no2(v, True)
    cpu_s = time.perf_counter() - t0
    return dict(row="8f-1 reader front-end (K7)", workload="TROPOMI NO2 granule 4172x450 px, 34 levels",
                unit="px/s", value=n_px / (ms * 1e-3), kernel_ms=ms, algorithmic_bytes=b,
                achieved_GBps=b / (ms * 1e-3) / 1e9, peak_GBps=peak(),
                frac=b / (ms * 1e-3) / 1e9 / peak(), cpu_px_per_s=n_px / cpu_s, cpu_s=cpu_s,
                cpu_kind="port (oracle/reader.py, numpy, 1 core)")


def row_nearest():
    from scipy.spatial import cKDTree
    coords = synth.ctm_coordinates()
    gpl = _plan.grid_plan(coords, 0.25)
    g = synth.make_amf_granule(3, "OMI_HCHO", geo=dict(node_lon_deg=-60.0))
    lon, lat = np.asarray(g.longitude_center), np.asarray(g.latitude_center)
    lonlat = (_dev.to_device(_plan.coord_array(lon)), _dev.to_device(_plan.coord_array(lat)))
    gpl.dev_tables()
    holder = {}

    def run():
        holder["p"] = _plan.nearest_plan(lon, lat, gpl, 0.5, lonlat_dev=lonlat)

    ms = gpu_ms(run)          # includes the one small D2H of the kept-cell flags
    X, Y = gpl.mesh()
    t0 = time.perf_counter()
    pts = np.column_stack((lon.ravel().astype(np.float64), lat.ravel().astype(np.float64)))
    cKDTree(pts).query(np.column_stack((X.ravel(), Y.ravel())))
    cpu_s = time.perf_counter() - t0
    n_px = lon.size
    return dict(row="8f-2 nearest-neighbour plan (types 2/4)",
                workload="OMI granule 98,640 px onto the 721x1439 mesh, GMI grid, 2x2 box",
                unit="px/s", value=n_px / (ms * 1e-3), plan_ms=ms, kept_cells=int(holder["p"].n_cells),
                cpu_px_per_s=n_px / cpu_s, cpu_s=cpu_s,
                cpu_kind="the reference's own step: cKDTree.query of all 1.04 M mesh nodes "
                         "(interpolator.py:144-150), 1 core; latency-bound, no roofline quoted")


def row_output():
    from oracle import output as ooutput
    import datetime
    import types
    n = 361 * 576
    rng = np.random.default_rng(5)
    arrs = [rng.uniform(0.1, 9.0, n) for _ in range(8)]
    arrs[1][::97] = 0.0
    arrs[2][::89] = np.nan
    dev = [_dev.to_device(a) for a in arrs]
    out = _dev.empty((9, n), "float32")
    lib = _lib.lib()

    def run():
        _lib.check(lib.oisat_output_fields(n, *[d.data_ptr() for d in dev], out.data_ptr(),
                                           _dev.stream()))

    ms = gpu_ms(run)
    b = n * (8 * 8 + 9 * 4)
    sh = (361, 576)
    obj = types.SimpleNamespace(
        sat_averaged_vcd=arrs[0].reshape(sh), ctm_averaged_vcd=arrs[1].reshape(sh),
        ctm_averaged_vcd_corrected=arrs[2].reshape(sh), sat_averaged_error=arrs[3].reshape(sh),
        ak_OI=arrs[4].reshape(sh), error_OI=arrs[5].reshape(sh), aux1=arrs[6].reshape(sh),
        aux2=arrs[7].reshape(sh), avg_time=datetime.datetime(2005, 6, 15),
        reader_obj=types.SimpleNamespace(sat_data=[types.SimpleNamespace(
            longitude_center=np.zeros(sh), latitude_center=np.zeros(sh))]))
    t0 = time.perf_counter()
    ooutput.output_fields(obj)
    cpu_s = time.perf_counter() - t0
    return dict(row="8f-3 output fields (K8)", workload="one month on the 361x576 GMI grid",
                unit="cells/s", value=n / (ms * 1e-3), kernel_ms=ms, algorithmic_bytes=b,
                achieved_GBps=b / (ms * 1e-3) / 1e9, peak_GBps=peak(),
                frac=b / (ms * 1e-3) / 1e9 / peak(),
                note="20 MB working set: launch-latency bound (one 5 us launch), not HBM bound",
                cpu_cells_per_s=n / cpu_s, cpu_s=cpu_s, cpu_kind="port (oracle/output.py, numpy, 1 core)")


def row_reader_mopitt():
    """8f-1 (round 2): MOPITT CO L3 front-end at the product's size (360 x 180 x 9 levels, 10 AK
    rows): oisat_reader_clean / _mopitt_xcol through reader_frontend.mopitt_co (uploads of the
    file variables and the device -> host copy of the record included: it is the drop-in call)."""
    import cases
    from oracle import reader as oreader
    small = cases.reader_vars("mopitt_co")
    rng = np.random.default_rng(2)
    nlon, nlat = 360, 180
    v = {}
    for k, a in small.items():
        a = np.asarray(a)
        if a.ndim >= 2 and a.shape[:2] == (72, 36):
            v[k] = np.resize(a, (nlon, nlat) + a.shape[2:]).astype(a.dtype)
        else:
            v[k] = a
    v["Latitude"] = np.linspace(-89.5, 89.5, nlat).astype(np.float32)
    v["Longitude"] = np.linspace(-179.5, 179.5, nlon).astype(np.float32)
    n_px = nlon * nlat
    t0 = time.perf_counter()
    for _ in range(3):
        rf.mopitt_co(v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        rf.mopitt_co(v)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    oreader.mopitt_co({k: np.array(a) for k, a in v.items()})
    cpu_s = time.perf_counter() - t0
    return dict(row="8f-1 reader front-end, MOPITT CO (K7 oisat_reader_clean)",
                workload="MOP03 daily L3, 360x180 px, 9 levels + 10 AK rows",
                unit="px/s", value=n_px / (ms * 1e-3), call_ms=ms,
                note="whole drop-in call from host variables to host record (H2D + 9 launches + D2H): "
                     "a 12 MB working set, host- and copy-bound",
                cpu_px_per_s=n_px / cpu_s, cpu_s=cpu_s, cpu_kind="port (oracle/reader.py, numpy, 1 core)")


def row_pwv():
    """8f-4 (round 2): model precipitable water on the global GMI grid (72 layers x 207,936
    cells): oisat_pwv_partial + oisat_pwv_column against the numpy statement (pwv_cal.py:62-93)."""
    n_lev, n = 72, 361 * 576
    rng = np.random.default_rng(3)
    dp = rng.uniform(1.0, 30.0, (n_lev, n)).astype(np.float32)
    q = rng.uniform(1e-6, 2e-2, (n_lev, n)).astype(np.float32)
    vcd = rng.uniform(1.0, 60.0, n)
    vcd[::7] = np.nan
    dpd, qd, vd = _dev.to_device(dp), _dev.to_device(q), _dev.to_device(vcd)
    pc = _dev.empty((n_lev, n), "float32")
    out = _dev.empty((n,))
    lib = _lib.lib()

    def run():
        _lib.check(lib.oisat_pwv_partial(dpd.data_ptr(), qd.data_ptr(), n_lev * n, pc.data_ptr(),
                                         _dev.stream()))
        _lib.check(lib.oisat_pwv_column(pc.data_ptr(), _lib.F32, n_lev, n, vd.data_ptr(),
                                        out.data_ptr(), _dev.stream()))

    ms = gpu_ms(run)
    t0 = time.perf_counter()
    want = np.nansum((dp * q / 9.80665 / 10000.0) / 1000.0, axis=0)
    want[np.isnan(vcd)] = np.nan
    cpu_s = time.perf_counter() - t0
    got = _dev.to_host(out)
    assert np.array_equal(got.astype(np.float32), want, equal_nan=True)     # bit for bit
    b = n_lev * n * (4 + 4 + 4 + 4) + n * 16
    return dict(row="8f-4 model precipitable water (oisat_pwv_partial + oisat_pwv_column)",
                workload="72 layers x 361x576 GMI cells", unit="cells/s", value=n / (ms * 1e-3),
                kernel_ms=ms, algorithmic_bytes=b, achieved_GBps=b / (ms * 1e-3) / 1e9,
                peak_GBps=peak(), frac=b / (ms * 1e-3) / 1e9 / peak(), bit_identical=True,
                cpu_cells_per_s=n / cpu_s, cpu_s=cpu_s, cpu_kind="numpy statement of pwv_cal.py:62-93, 1 core")


def row_ssmis():
    """8f-4 (round 2): one monthly SSMIS map (1440 x 720, 0.25 degree) onto the global GMI grid
    through interpolator_ssmis: the first call builds the two lattice plans (Qhull + scipy's
    walk on the host, exact for lattices; cached by geometry, every month of a record shares
    them), later calls are the GPU work alone.  Oracle = the reference's algorithm (two Delaunay
    triangulations and LinearNDInterpolator evaluations per field, interpolator_ssmis.py)."""
    import types
    from oisatgmi_b200 import interpolator_ssmis as issmis
    from oracle import ssmis as ossmis
    rng = np.random.default_rng(1)
    lat = np.arange(-89.875, 90.0, 0.25)
    lon = np.arange(0.125, 360.0, 0.25)
    wv = np.clip(60 + 90 * (synth._smooth2d(rng, (lat.size, lon.size), scale=3.0) + 0.5), 0, 249)
    wv[~synth._coherent_mask(rng, wv.shape, 0.3)] = 255
    v = {"latitude": lat, "longitude": lon, "atmosphere_water_vapor_content": wv.astype(np.uint8)}
    rec = rf.ssmis_wv({k: np.array(a) for k, a in v.items()}, "200506")
    coords = synth.ctm_coordinates()
    t0 = time.perf_counter()
    got = issmis.interpolator_ssmis(1, 0.25, rec, coords)
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t0
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        got = issmis.interpolator_ssmis(1, 0.25, rec, coords)
    torch.cuda.synchronize()
    later_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    want = ossmis.interpolator_ssmis(1, 0.25, ossmis.ssmis_wv({k: np.array(a) for k, a in v.items()},
                                                              "200506"), coords)
    cpu_s = time.perf_counter() - t0
    same_mask = bool(np.array_equal(np.isnan(got.vcd), np.isnan(want.vcd)))
    f = np.isfinite(want.vcd)
    rel = float(np.max(np.abs(got.vcd[f] - want.vcd[f]) / np.maximum(np.abs(want.vcd[f]), 1e-300)))
    n_px = lat.size * lon.size
    return dict(row="8f-4 interpolator_ssmis (two plan steps + box mean)",
                workload="SSMIS monthly map 1440x720 onto the 361x576 GMI grid, 0.25 degree mesh",
                unit="px/s", value=n_px / (later_ms * 1e-3), call_ms_plans_cached=later_ms,
                first_call_s_building_plans=first_s, same_nan_mask=same_mask, max_rel_err_vcd=rel,
                cpu_px_per_s=n_px / cpu_s, cpu_s=cpu_s, cpu_kind="port (oracle/ssmis.py, scipy, 1 core)")


def main():
    _dev.require_cuda()
    rows = (row_reader, row_nearest, row_output, row_reader_mopitt, row_pwv, row_ssmis)
    want = sys.argv[1:]
    for fn in rows:
        if want and fn.__name__ not in want:
            continue
        print(json.dumps(fn()), flush=True)


if __name__ == "__main__":
    main()
