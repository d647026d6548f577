"""Thread scaling of the native swath Delaunay (host only, no GPU needed).

    python tools/delaunay_scaling.py [n_orbits]
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from oisatgmi_b200 import plan as _plan  # noqa: E402


def main():
    n_orb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    day = bench.make_day(0, n_orb)
    lons = [np.asarray(g.longitude_center) for g in day]
    lats = [np.asarray(g.latitude_center) for g in day]
    _plan.native_delaunay(lons[0], lats[0])
    os.system("lscpu | egrep 'Model name|^CPU\\(s\\)|Thread|Core|Socket|MHz' 1>&2")
    for workers in (1, 2, 4, 8, 12, 15, 16):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda i: _plan.native_delaunay(lons[i], lats[i]), range(n_orb)))
        dt = time.perf_counter() - t0
        print("workers %2d: %.1f ms wall, %.1f ms per granule-thread" %
              (workers, dt * 1e3, dt * 1e3 * min(workers, n_orb) / n_orb), flush=True)


if __name__ == "__main__":
    main()
