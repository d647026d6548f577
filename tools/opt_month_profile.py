#!/usr/bin/env python
"""One MOPITT (or GOSAT) month through OptMonthPipeline: two runs (record, replay).  Meant to
run under `ncu --metrics gpu__time_duration.sum` for a per-kernel launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch  # noqa: E402

import bench_configs  # noqa: E402
from oisatgmi_b200.opt_pipeline import OptMonthPipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mopitt_co"
c = dict(bench_configs.OPT_CONFIGS[name])
c["days"] = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model, grans = bench_configs._opt_month(c)
pipe = OptMonthPipeline(model, 1.0, 0.0, c["sensor"])
for g in grans:
    pipe.add_granule(g)
for _ in range(2):
    pipe.run()
torch.cuda.synchronize()
print("done", pipe.n_pixels())
