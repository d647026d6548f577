"""cProfile of the host side of one e2e step (bench.py's `e2e`): where the wall time of
`add_day` (uploads + plans) and `allocate` (tables + buffers) goes.

    python tools/e2e_profile.py [n_orbits]
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from oisatgmi_b200.pipeline import MonthPipeline  # noqa: E402


def main():
    n_orb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    model = bench.make_model()
    day = bench.make_day(0, n_orb)
    hosts = [MonthPipeline.host_arrays(g, pin=True) for g in day]

    def new_pipe():
        return MonthPipeline(model, bench.GRID_SIZE, bench.FLAG_THRESH, sensor="OMI", gas="HCHO",
                             error_ctm=50.0)
    ctm_dev = None
    for it in range(3):
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        p = new_pipe()
        if ctm_dev is not None:
            p._ctm_dev, p._ctm_slots = ctm_dev
        pra = cProfile.Profile()
        if os.environ.get("OISAT_PLAN_TRACE") != "1":
            pra.enable()
        p.add_day(day, hosts=hosts)
        ta = time.perf_counter()
        torch.cuda.synchronize()
        pra.disable()
        t1 = time.perf_counter()
        pr.enable()
        p.allocate()
        torch.cuda.synchronize()
        pr.disable()
        t2 = time.perf_counter()
        if it == 2:
            print("add_day returned after %.1f ms, drained after %.1f ms" % ((ta - t0) * 1e3, (t1 - t0) * 1e3))
            if os.environ.get("OISAT_PLAN_TRACE") != "1":
                pstats.Stats(pra).sort_stats("cumulative").print_stats(25)
        out = p.results_to_host(p.run())
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ctm_dev = (p._ctm_dev, list(p._ctm_slots))
        print("iter %d: add_day %.1f ms, allocate %.1f ms, run+d2h %.1f ms" %
              (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
        if it == 2:
            pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
            pr2 = cProfile.Profile()
            p = new_pipe()
            p._ctm_dev, p._ctm_slots = ctm_dev
            p.add_day(day, hosts=hosts)
            p.allocate()
            torch.cuda.synchronize()
            pr2.enable()
            out = p.results_to_host(p.run())
            torch.cuda.synchronize()
            pr2.disable()
            pstats.Stats(pr2).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
