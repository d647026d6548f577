#!/usr/bin/env python
"""The other single-GPU configurations of BASELINE.json through the fused month pipeline at
full size, one JSON line each (bench.py itself stays on configs[1], the configuration the
metric is quoted on):

    python tests/bench_configs.py [tropomi_no2] [omi_no2] [mopitt_co] [gosat_xch4] > profiles/rNN_configs.jsonl
    python -m torch.distributed.run --nproc-per-node 8 tests/bench_configs.py gosat_xch4

  tropomi_no2  configs[4], one GPU's share: TROPOMI-scale NO2, 4172 x 450 px granules
               (1.88 M px, 34 levels, 0.10 degree mesh, 5 x 6 box => 90-entry stencils),
               14 orbits = one day = 26 M px
  omi_no2      configs[0]: OMI NO2 (35 levels + tropopause mask), one month of 435 granules
  mopitt_co    configs[2]: MOPITT CO L3, 30 daily lattices, AK convolution + OI (OptMonthPipeline)
  gosat_xch4   configs[3]: GOSAT XCH4, 30 days of soundings, gap filling + AK convolution + OI on
               aux2/aux1 (OptMonthPipeline); also under torchrun, days dealt to the ranks

A "step" is what bench.py times: pack, fused gather + AMF, ordered accumulation, OI, with
reader arrays and geometry plans resident in HBM; CUDA events on the launching stream, 3
warm-ups, 5 timed steps; inputs (>= 4 GB) are larger than L2.  Bytes per pixel as in
SURVEY.md section 8d (reader dtypes): P * B_px + pairs * (864 + 160) + 14 * 8 * n_cell.
(Lives under tests/ like bench_rows.py; it does not run the oracle.)
"""
import copy
import datetime
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]   # tests/: the synthetic generators (synth.py)

import torch  # noqa: E402

import bench  # noqa: E402
from oisatgmi_b200 import _lib, plan as _plan  # noqa: E402
import synth  # noqa: E402
from oisatgmi_b200.pipeline import MonthPipeline  # noqa: E402

CONFIGS = {
    "tropomi_no2": dict(product="TROPOMI_NO2", sensor="TROPOMI", gas="NO2", grid_size=0.10, thresh=0.75,
                        orbits=14, days=1, bytes_px=4 + 4 + 8 + 8 + 2 + 2 + 2 + 34 * 2 + 34 * 2,
                        label="configs[4] (one GPU's share): TROPOMI-scale NO2, one day = 14 orbits "
                              "x 4172 x 450 px x 34 levels, 0.10 degree mesh, 90-entry stencils"),
    "omi_no2": dict(product="OMI_NO2", sensor="OMI", gas="NO2", grid_size=0.25, thresh=0.0,
                    orbits=15, days=29, bytes_px=4 + 4 + 8 + 8 + 2 + 2 + 2 + 35 * 2 + 35 * 2,
                    label="configs[0]: OMI NO2 one-month OI, 435 granules x 98,640 px x 35 levels "
                          "+ tropopause mask"),
}


def run(name):
    c = CONFIGS[name]
    model = bench.make_model()
    t0 = time.perf_counter()
    day = []
    for i in range(c["orbits"]):
        t = datetime.datetime(2005, 6, 1) + datetime.timedelta(seconds=1800 + i * 5933)
        day.append(synth.make_amf_granule(500 + i, c["product"], geo=bench.orbit_geo(i, c["orbits"]),
                                          bad_fraction=0.2, time=t))
    gen_s = time.perf_counter() - t0
    pipe = MonthPipeline(model, c["grid_size"], c["thresh"], sensor=c["sensor"], gas=c["gas"],
                         error_ctm=50.0)
    hosts = [MonthPipeline.host_arrays(g, pin=False) for g in day]
    lons = [np.asarray(g.longitude_center) for g in day]
    lats = [np.asarray(g.latitude_center) for g in day]
    t0 = time.perf_counter()
    plans = _plan.granule_plans(lons, lats, pipe.gplan, c["grid_size"] * 2.0)
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    for d in range(c["days"]):
        for i, g in enumerate(day):
            gg = copy.copy(g)
            gg.time = g.time + datetime.timedelta(days=d)
            pipe.add_granule(gg, plan=plans[i], host=hosts[i])
    pipe.upload_ctm()
    host_t, _ = pipe.build_tables()
    pipe.allocate()
    torch.cuda.synchronize()
    n_px, n_pairs = pipe.n_pixels(), host_t["n_pairs"]
    for _ in range(3):
        pipe.run()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    marks_all = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 5
    for _ in range(steps):
        marks = []
        res = pipe.run(marks)
        marks_all.append(marks)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    phase = {}
    for marks in marks_all:
        for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
            phase.setdefault(n1, []).append(a.elapsed_time(b))
    phase = {k: float(np.mean(v)) for k, v in phase.items()}
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    total_bytes = n_px * c["bytes_px"] + n_pairs * (864 + 160) + 14 * 8 * pipe.n_cell
    value = n_px / (ms * 1e-3)
    return {"config": name, "workload": c["label"], "metric": "L2 pixels/sec through interp+AMF+grid+OI",
            "value": value, "unit": "px/s", "ms_per_step": ms, "steps": steps, "warmup": 3,
            "pixels": int(n_px), "pairs": int(n_pairs), "granules": len(pipe.granules),
            "stencil_entries": 3 * pipe.gplan.nwin, "fused_form": pipe.fused_form,
            "input_GB": pipe.input_bytes() / 1e9, "phase_ms": phase,
            "bytes_per_px": total_bytes / n_px,
            "roofline": {"bound": "hbm", "achieved": total_bytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": total_bytes / (ms * 1e-3) / 1e9 / peak,
                         "what": "whole step, SURVEY 8d algorithmic bytes"},
            "plan_build_s_one_day": plan_s, "synth_s": gen_s, "knee_index": int(res["knee_index"]),
            "gpu_launches_per_step": int((_lib.launch_count() - launches0) / steps)}


# ------------------------------------------------------------------ satellite_opt months
OPT_CONFIGS = {
    "mopitt_co": dict(sensor="MOPITT", days=30, gas_scale=40.0,
                      # reader dtypes per L3 pixel (reader.py:1130-1213): vcd f16, sigma f32, quality
                      # f16, x_col f32, 10 AK rows f16, 9 pressure levels f16, 9 a-priori levels f32,
                      # a-priori column f16, a-priori surface f32, lon/lat f32
                      bytes_px=2 + 4 + 2 + 4 + 10 * 2 + 9 * 2 + 9 * 4 + 2 + 4 + 8,
                      label="configs[2]: MOPITT CO L3 one-month AK convolution + OI, 30 daily "
                            "360 x 180 lattices x 9 levels (10 AK rows), daily 361x576x72 model "
                            "resampled to the 1 degree mesh"),
    "gosat_xch4": dict(sensor="GOSAT", days=30, gas_scale=600.0, soundings=3000,
                       # per sounding (reader.py:1216-1275): xch4, sigma, quality f64, 4 x 20 levels
                       # f64 (AK, pressure, a priori, pressure weight), lon/lat f32
                       bytes_px=8 + 8 + 8 + 4 * 20 * 8 + 8,
                       label="configs[3]: GOSAT XCH4 one-month gap filling + AK convolution + OI "
                             "(on aux2/aux1), 30 days x 3000 soundings x 20 levels, daily "
                             "361x576x72 model resampled to the 1 degree mesh"),
}


def _opt_month(c):
    """Synthetic month: `days` granules and a daily 3-D model (one set of fields shared by the
    days -- distinct time stamps, so the per-day resampling is done for every day)."""
    coords = synth.ctm_coordinates()
    base = synth.make_ctm(21, coords, ctmtype="ECCOH", averaged=False, gas_scale=c["gas_scale"],
                          date=datetime.datetime(2005, 6, 1))
    model = []
    for d in range(c["days"]):
        m = copy.copy(base)
        m.time = [datetime.datetime(2005, 6, 1) + datetime.timedelta(days=d)]
        model.append(m)
    grans = []
    for d in range(c["days"]):
        t = datetime.datetime(2005, 6, 1 + d, 12)
        if c["sensor"] == "MOPITT":
            grans.append(synth.make_mopitt_granule(300 + d % 6, time=t))
        else:
            grans.append(synth.make_gosat_soundings(400 + d % 6, n=c["soundings"], time=t))
    return model, grans


def run_opt(name):
    """One JSON line for a satellite_opt month (configs[2] / [3]); under torchrun the days are
    dealt to the ranks (sharding.assign) and one all-reduce merges the accumulators."""
    import torch.distributed as dist
    from oisatgmi_b200 import sharding
    from oisatgmi_b200.opt_pipeline import OptMonthPipeline
    c = OPT_CONFIGS[name]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    pg = None
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        pg = dist.group.WORLD
    model, grans = _opt_month(c)
    mine = sharding.assign([g.time for g in grans], rank, world)
    pipe = OptMonthPipeline(model, 1.0, 0.0, c["sensor"], process_group=pg)
    t0 = time.perf_counter()
    for i in mine:
        pipe.add_granule(grans[i])
    torch.cuda.synchronize()
    add_s = time.perf_counter() - t0            # uploads + geometry plans (lattice plan cached)
    for _ in range(3):
        res = pipe.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    steps = 5
    marks_all = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        marks = []
        res = pipe.run(marks)
        marks_all.append(marks)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_px = pipe.n_pixels()
    if world > 1:
        tv = torch.tensor([ms, float(n_px)], device="cuda", dtype=torch.float64)
        tmax = tv.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tv)
        ms, n_px = float(tmax[0].item()), int(tv[1].item())
    phase = {}
    for marks in marks_all:
        for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
            phase.setdefault(n1, []).append(a.elapsed_time(b))
    phase = {k: float(np.mean(v)) for k, v in phase.items()}
    # end to end: a fresh pipeline from host records every step (uploads, plans, kernels, D2H)
    e2e_t = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p2 = OptMonthPipeline(model, 1.0, 0.0, c["sensor"], process_group=pg)
        for i in mine:
            p2.add_granule(grans[i])
        out = p2.results_to_host(p2.run())
        torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t[1:]))
    if world > 1:
        tv = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        e2e_s = float(tv.item())
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n_cell = pipe.n_cell
    # SURVEY 8d style: reader bytes of every pixel + per granule the model column (3 x 72 f32)
    # and the accumulator traffic (10 x f64 read + write) of every mesh cell + the OI pass
    n_gran = c["days"]
    total_bytes = n_px * c["bytes_px"] + n_gran * n_cell * (864 + 160) + 14 * 8 * n_cell
    value = n_px / (ms * 1e-3)
    cpu = None
    if rank == 0:
        cpu = _opt_cpu_baseline(c, model, grans[0])
    line = {"config": name, "workload": c["label"], "metric": "L2 pixels/sec through interp+AK+grid+OI",
            "value": value, "unit": "px/s", "n_gpus": world, "ms_per_step": ms, "steps": steps, "warmup": 3,
            "pixels": int(n_px), "granules": n_gran, "mesh_cells": int(n_cell),
            "input_MB": pipe.input_bytes() / 1e6, "phase_ms": phase, "bytes_per_px": total_bytes / n_px,
            "roofline": {"bound": "hbm", "achieved": total_bytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": total_bytes / (ms * 1e-3) / 1e9 / peak,
                         "what": "whole step, SURVEY 8d style algorithmic bytes; the month is %d "
                                 "small granules, the step is launch-latency bound" % n_gran},
            "e2e": {"value": n_px / e2e_s, "unit": "px/s", "s_per_step": e2e_s,
                    "what": "fresh pipeline from host records: uploads, geometry plans (cached "
                            "lattice plan; per-day Delaunay of the soundings for GOSAT), kernels, "
                            "D2H of the monthly fields"},
            "add_granules_s": add_s, "cpu_baseline": cpu, "knee_index": int(res["knee_index"]),
            "gpu_launches_per_step": int((_lib.launch_count() - launches0) / (steps + 3))}
    if world > 1:
        dist.barrier()
    return line if rank == 0 else None


def _opt_cpu_baseline(c, model, granule):
    """The oracle port of the reference on ONE granule of the month, one host core: gap filling
    (GOSAT) + interpolator + ak_conv_*, measured."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import interp as ointerp, vertical as overt
    coords = {"Latitude": model[0].latitude, "Longitude": model[0].longitude}
    t0 = time.perf_counter()
    g = copy.deepcopy(granule)
    if c["sensor"] == "GOSAT":
        g = ointerp.filler_gosatxch4(1.0, g, flag_thresh=0.0)
    grid = ointerp.interpolator(1, 1.0, g, coords, flag_thresh=0.0)
    (overt.ak_conv_mopitt if c["sensor"] == "MOPITT" else overt.ak_conv_gosat)(model[:1], [grid])
    wall = time.perf_counter() - t0
    n_px = int(np.size(granule.latitude_center))
    return {"value": n_px / wall, "unit": "px/s", "cores": 1, "kind": "port",
            "sample": "one %s granule (%d px) through oracle/ (gap filling for GOSAT, interpolator, "
                      "ak_conv): %.1f s of CPU wall time, measured" % (c["sensor"], n_px, wall)}


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CONFIGS)):
        line = run_opt(name) if name in OPT_CONFIGS else run(name)
        if line is not None:
            print(json.dumps(line), flush=True)
        torch.cuda.empty_cache()
