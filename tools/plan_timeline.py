"""Timeline of one day batch of geometry plans (what bounds bench.py's `e2e`): when each
native triangulation finishes on the thread pool and how long the device part takes.

    python tools/plan_timeline.py [n_orbits] [workers]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from oisatgmi_b200 import _dev, plan as _plan  # noqa: E402
from oisatgmi_b200.pipeline import MonthPipeline  # noqa: E402


def main():
    n_orb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else min(n_orb, os.cpu_count())
    from concurrent.futures import ThreadPoolExecutor, as_completed
    model = bench.make_model()
    day = bench.make_day(0, n_orb)
    pipe = MonthPipeline(model, bench.GRID_SIZE, bench.FLAG_THRESH, sensor="OMI", gas="HCHO",
                         error_ctm=50.0)
    gplan = pipe.gplan
    lons = [np.asarray(g.longitude_center) for g in day]
    lats = [np.asarray(g.latitude_center) for g in day]
    radius = bench.GRID_SIZE * 2.0
    _plan.granule_plans(lons, lats, gplan, radius)
    torch.cuda.synchronize()
    for fn in (_plan.native_delaunay, _plan.native_delaunay_adj):
        best = min(_t(lambda: fn(lons[0], lats[0])) for _ in range(5))
        print("%s serial: %.1f ms" % (fn.__name__, best * 1e3))
        for w in (4, 8, 15):
            def run():
                with ThreadPoolExecutor(w) as ex:
                    list(ex.map(lambda i: fn(lons[i], lats[i]), range(n_orb)))
            print("  %d threads, %d granules: %.1f ms" % (w, n_orb, min(_t(run) for _ in range(3)) * 1e3))
    lonlat = [(_dev.to_device(_plan.coord_array(lons[i])), _dev.to_device(_plan.coord_array(lats[i])))
              for i in range(n_orb)]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        keeps = [_plan.distance_mask(lo, la, gplan, radius) for lo, la in lonlat]
        ev, pending = [], []
        with ThreadPoolExecutor(workers) as ex:
            futures = {ex.submit(_plan.native_delaunay_adj, lons[i], lats[i], True): i for i in range(n_orb)}
            for fut in as_completed(futures):
                i = futures[fut]
                tri, half, ties, maxabs = fut.result()
                t1 = time.perf_counter()
                pending.append((i, _plan._plan_v1_enqueue(tri, lonlat[i], gplan, keeps[i], half, maxabs)))
                t2 = time.perf_counter()
                ev.append((i, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
        t3 = time.perf_counter()
        for i, st in pending:
            _plan._plan_v1_finish(st, gplan)
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        print("rep %d total %.1f ms (pool %.1f, finish %.1f, drain %.1f); (granule, ready at ms, enqueue ms):"
              % (rep, (t5 - t0) * 1e3, (t3 - t0) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3),
              " ".join("%d:%.0f+%.1f" % e for e in ev), flush=True)
        t0 = time.perf_counter()
        _plan.granule_plans(lons, lats, gplan, radius, lonlat_dev=lonlat)
        torch.cuda.synchronize()
        print("granule_plans: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)


def _t(f):
    t0 = time.perf_counter()
    f()
    return time.perf_counter() - t0


if __name__ == "__main__":
    main()
