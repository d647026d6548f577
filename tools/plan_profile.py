"""Wall-clock breakdown of the geometry-plan build for one day batch (the part of
`e2e` that dominates): K0, native Delaunay (serial and threaded), device part.

    python tools/plan_profile.py [n_orbits]
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
    model = bench.make_model()
    day = bench.make_day(0, n_orb)
    pipe = MonthPipeline(model, bench.GRID_SIZE, bench.FLAG_THRESH, sensor="OMI", gas="HCHO",
                         error_ctm=50.0)
    gplan = pipe.gplan
    lons = [np.asarray(g.longitude_center) for g in day]
    lats = [np.asarray(g.latitude_center) for g in day]
    radius = bench.GRID_SIZE * 2.0
    sync = torch.cuda.synchronize
    _plan.granule_plans(lons, lats, gplan, radius)   # warm-up (module load, tables)
    sync()
    for rep in range(2):
        t = {}
        t0 = time.perf_counter()
        lonlat = [(_dev.to_device(_plan.coord_array(lons[i])), _dev.to_device(_plan.coord_array(lats[i])))
                  for i in range(n_orb)]
        sync(); t["h2d_coords"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        keeps = [_plan.distance_mask(lo, la, gplan, radius) for lo, la in lonlat]
        sync(); t["k0"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        one = _plan.native_delaunay(lons[0], lats[0])
        t["delaunay_one_serial"] = time.perf_counter() - t0
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        with ThreadPoolExecutor(min(n_orb, os.cpu_count())) as ex:
            tris = list(ex.map(lambda i: _plan.native_delaunay(lons[i], lats[i]), range(n_orb)))
        t["delaunay_threaded"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        out = [_plan._plan_v1_device(tris[i][0], lonlat[i], gplan, keeps[i]) for i in range(n_orb)]
        sync(); t["device_part"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        _plan.granule_plans(lons, lats, gplan, radius)
        sync(); t["granule_plans_total"] = time.perf_counter() - t0
        print({k: round(v * 1e3, 2) for k, v in t.items()}, "n_tri", one[0].shape, "ties",
              [x[1] for x in tris][:4], flush=True)


if __name__ == "__main__":
    main()


def device_breakdown():
    """Per-call wall time (with a sync after each) of the device part for one granule."""
    from oisatgmi_b200 import _lib
    model = bench.make_model()
    day = bench.make_day(0, 2)
    pipe = MonthPipeline(model, bench.GRID_SIZE, bench.FLAG_THRESH, sensor="OMI", gas="HCHO",
                         error_ctm=50.0)
    gplan = pipe.gplan
    lon = np.asarray(day[1].longitude_center)
    lat = np.asarray(day[1].latitude_center)
    radius = bench.GRID_SIZE * 2.0
    sync = torch.cuda.synchronize
    L = _lib.lib()
    lo, la = _dev.to_device(_plan.coord_array(lon)), _dev.to_device(_plan.coord_array(lat))
    keep = _plan.distance_mask(lo, la, gplan, radius)
    tri_host, ties = _plan.native_delaunay(lon, lat)
    _plan._plan_v1_device(tri_host, (lo, la), gplan, keep)
    sync()
    for rep in range(2):
        t = {}

        def lap(name, t0):
            sync()
            t[name] = round((time.perf_counter() - t0) * 1e3, 3)

        xs, ys = gplan.dev_axes()
        window, nn_ok = gplan.dev_tables()
        t0 = time.perf_counter(); tri = _dev.to_device(tri_host); lap("h2d_tri", t0)
        t0 = time.perf_counter()
        node_tri = _dev.full((gplan.H * gplan.W,), 2 ** 31 - 1, "int32")
        work = _dev.empty((2 * tri.shape[0] + 2,), "int32")
        lap("alloc_fill", t0)
        code = _dev.dtype_code(lo)
        s = _dev.stream()
        t0 = time.perf_counter()
        _lib.check(L.oisat_locate(tri.data_ptr(), tri.shape[0], lo.data_ptr(), la.data_ptr(), code,
                                  xs.data_ptr(), gplan.W, ys.data_ptr(), gplan.H, keep.data_ptr(),
                                  node_tri.data_ptr(), work.data_ptr(), s))
        lap("locate", t0)
        n_cell = int(np.prod(gplan.out_shape))
        ok = _dev.empty((n_cell,), "uint8")
        t0 = time.perf_counter()
        _lib.check(L.oisat_plan_cells(window.data_ptr(), gplan.nwin, nn_ok.data_ptr(), n_cell,
                                      node_tri.data_ptr(), ok.data_ptr(), s))
        lap("plan_cells", t0)
        t0 = time.perf_counter(); okh = _dev.to_host(ok); lap("d2h_ok", t0)
        t0 = time.perf_counter(); cells = np.flatnonzero(okh); lap("flatnonzero", t0)
        n = cells.size
        S = 3 * gplan.nwin
        t0 = time.perf_counter()
        vert = _dev.empty((n, S), "int32")
        w = _dev.empty((n, S))
        cells_d = _dev.to_device(cells.astype(np.int32))
        lap("alloc_h2d_cells", t0)
        t0 = time.perf_counter()
        _lib.check(L.oisat_plan_fill(cells_d.data_ptr(), n, window.data_ptr(), gplan.nwin,
                                     node_tri.data_ptr(), tri.data_ptr(), lo.data_ptr(),
                                     la.data_ptr(), code, xs.data_ptr(), gplan.W, ys.data_ptr(), 1,
                                     vert.data_ptr(), w.data_ptr(), s))
        lap("plan_fill", t0)
        print("device part:", t, "n_cells", n, "nwin", gplan.nwin, "mesh", gplan.H, gplan.W, flush=True)


if __name__ == "__main__" and len(sys.argv) > 2:
    device_breakdown()
