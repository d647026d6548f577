"""K12 timing: host seed parts and the device part (assemble + flip rounds + check) of the
triangulation of one granule, for several tail thresholds (OISAT_FLIP_TAIL).

    python tools/flip_profile.py [omi|tropomi] > profiles/rNN_flip_profile.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from oisatgmi_b200 import _dev, plan  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "omi"
    nt, nx = (1644, 60) if kind == "omi" else (4172, 450)
    t = _dev.torch()
    rng = np.random.default_rng(3)
    for node in (10.0, 176.0):
        lat, lon = synth.swath_geolocation(nt, nx, node_lon_deg=node, rng=rng)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            parts = plan.native_seed_parts(lon, lat, pinned=True)
            best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(3):
            plan.native_delaunay_adj(lon, lat, pinned=True)
        host_ms = (time.perf_counter() - t0) / 3 * 1e3
        print("%s node %.0f: %d px, %d triangles (%d outside the lattice); host seed parts %.2f ms, "
              "incremental host builder %.2f ms" % (kind, node, lon.size, parts["n_tri"],
                                                    parts["n_outside"], best * 1e3, host_ms))
        d_lo, d_la = _dev.to_device(lon.ravel()), _dev.to_device(lat.ravel())
        for tail in ("0", "512", "2048", "8192", "32768", "1000000000"):
            os.environ["OISAT_FLIP_TAIL"] = tail
            for _ in range(2):
                plan.device_triangulation(parts, (d_lo, d_la))
            t.cuda.synchronize()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                tri, half, res, keep = plan.device_triangulation(parts, (d_lo, d_la))
            e1.record()
            t.cuda.synchronize()
            r = res.cpu().numpy()
            print("  tail %-10s device %.3f ms per granule; rounds %d flips %d bad %d undecided %d"
                  % (tail, e0.elapsed_time(e1) / reps, r[0], r[1], r[2], r[3]))


def batch(kind="omi", n=15):
    """The granules of a day through the rounds together (what plan.granule_plans does)."""
    nt, nx = (1644, 60) if kind == "omi" else (4172, 450)
    t = _dev.torch()
    rng = np.random.default_rng(4)
    parts, coords = [], []
    for k in range(n):
        lat, lon = synth.swath_geolocation(nt, nx, node_lon_deg=-170.0 + 24.0 * k, rng=rng)
        parts.append(plan.native_seed_parts(lon, lat, pinned=True))
        coords.append((_dev.to_device(lon.ravel()), _dev.to_device(lat.ravel())))
    for tail, cut in (("0", None), ("512", None), ("4096", None), ("512", 1), ("512", 2), ("512", 4),
                      ("512", 8), ("512", 16), ("512", 32), ("512", 64)):
        os.environ["OISAT_FLIP_TAIL"] = tail
        os.environ.pop("OISAT_FLIP_MAX_ROUNDS", None)
        if cut is not None:
            os.environ["OISAT_FLIP_MAX_ROUNDS"] = str(cut)
        times = []
        for rep in range(4):
            meshes = [plan.seed_assemble_device(p) for p in parts]
            t.cuda.synchronize()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            result, work = plan.flip_batch_device([(m[0], m[1], c) for m, c in zip(meshes, coords)])
            e1.record()
            t.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        r = result.cpu().numpy()
        print("batch of %d %s granules, tail %-6s%s: %.3f ms (min of 3 after warm-up); rounds %d flips %d "
              "bad %d undecided %d" % (n, kind, tail, "" if cut is None else " cut after %d rounds" % cut,
                                       min(times[1:]), r[0, 0], r[0, 1], r[:, 2].sum(), r[:, 3].sum()))
    os.environ.pop("OISAT_FLIP_MAX_ROUNDS", None)


if __name__ == "__main__":
    main()
    batch(sys.argv[1] if len(sys.argv) > 1 else "omi", 15 if (len(sys.argv) < 2 or sys.argv[1] == "omi") else 14)
