#!/usr/bin/env python
"""The other single-GPU configurations of BASELINE.json through the fused month pipeline at
full size, one JSON line each (bench.py itself stays on configs[1], the configuration the
metric is quoted on):

    python tests/bench_configs.py [tropomi_no2] [omi_no2] > profiles/r01_configs.jsonl

  tropomi_no2  configs[4], one GPU's share: TROPOMI-scale NO2, 4172 x 450 px granules
               (1.88 M px, 34 levels, 0.10 degree mesh, 5 x 6 box => 90-entry stencils),
               14 orbits = one day = 26 M px
  omi_no2      configs[0]: OMI NO2 (35 levels + tropopause mask), one month of 435 granules

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
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oisatgmi_b200 import _lib, plan as _plan, synth  # noqa: E402
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


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CONFIGS)):
        print(json.dumps(run(name)), flush=True)
        torch.cuda.empty_cache()
