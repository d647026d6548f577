#!/usr/bin/env python
"""Regenerate profiles/<round>_* from the captures a gpurun call brought back:

    python tools/make_profiles.py <launches.csv> <full.ncu-rep> <bench_n1.json> [round, default r02]

  launches.csv    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv
                  of `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline`
  full.ncu-rep    ncu --set full --clock-control none --import-source on of
                  `python bench.py --steps 1 --warmup 1 --days 8 --no-e2e --no-cpu-baseline`
  bench_n1.json   the line of a plain `python bench.py` run (no profiler)
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(HERE, "profiles")
STEP = ["ctm_prepare_kernel", "pack_batch_kernel", "pair_alive_kernel", "fused_tile_kernel", "fused_ws_kernel", "gather_rows_kernel", "vertical_rows_kernel",
        "accum_pairs_kernel", "accum_finalize_kernel", "oi_prepare_kernel", "oi_sweep_leaf_kernel",
        "oi_sweep_combine_kernel", "oi_apply_kernel"]


def short(k):
    return k.replace("void ", "").replace("oisat::", "")


def main():
    launches, rep, bench_path = sys.argv[1:4]
    RND = sys.argv[4] if len(sys.argv) > 4 else "r02"
    shutil.copy(launches, os.path.join(PROF, RND + "_launches.csv"))
    if os.path.abspath(bench_path) != os.path.join(PROF, RND + "_bench_n1.json"):
        shutil.copy(bench_path, os.path.join(PROF, RND + "_bench_n1.json"))
    bench = json.load(open(bench_path))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    in_step = {k: a for k, a in agg.items() if any(s in k for s in STEP)}
    n_steps = max(a[0] for a in in_step.values())
    tot = sum(a[1] for a in in_step.values())
    out = ["# " + RND + " launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` of",
           "# `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` (435 OMI HCHO granules, "
           "42.9 M px, 7.36 M pairs per step;",
           "# %d steps captured = 1 warm-up + 2 timed; the 15 plan-building launches per kernel belong to "
           "set-up, not to the step)." % n_steps,
           "# Per-launch times are cold-cache and serialised: the SHARE of the step is what must agree "
           "with bench.py's CUDA-event phases.", "",
           "| kernel | launches | total ms | ms per step | share of step kernels |", "|---|---|---|---|---|"]
    for k, a in sorted(in_step.items(), key=lambda x: -x[1][1]):
        out.append("| `%s` | %d | %.3f | %.3f | %.1f %% |" % (short(k), a[0], a[1], a[1] / n_steps,
                                                            100 * a[1] / tot))
    out.append("| **sum** | | %.3f | **%.3f** | |" % (tot, tot / n_steps))
    out += ["", "Set-up (geometry plans of the 15 distinct orbits, once per run; inside `e2e`, not inside "
            "`value`):", "", "| kernel | launches | total ms | ms per granule |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if k in in_step or a[0] % 15 or a[0] == 0:
            continue
        out.append("| `%s` | %d | %.3f | %.3f |" % (short(k), a[0], a[1], a[1] / a[0]))
    ph = bench["roofline"]["phase_ms"]
    s_ph = sum(ph.values())
    pack = sum(a[1] for k, a in in_step.items() if "pack" in k)
    fused = sum(a[1] for k, a in in_step.items() if "rows_kernel" in k or "fused_tile" in k)
    out += ["", "bench.py CUDA-event phases of the same workload without ncu (`profiles/" + RND + "_bench_n1.json`): "
            + ", ".join("%s %.2f ms" % kv for kv in ph.items()) + " (step %.2f ms)." % bench["ms_per_step"],
            "Shares agree: pack %.0f %% (events) vs %.0f %% (ncu), fused step %.0f %% vs %.0f %%."
            % (100 * ph["pack"] / s_ph, 100 * pack / tot, 100 * ph["fused"] / s_ph, 100 * fused / tot)]
    open(os.path.join(PROF, RND + "_launches_summary.md"), "w").write("\n".join(out) + "\n")

    header = ("# " + RND + " (final code of the round) ncu --set full --clock-control none --import-source on; "
              "cold-cache, serialised replays -- not bench numbers\n# command: python bench.py --steps 1 "
              "--warmup 1 --days 8 --no-e2e --no-cpu-baseline   (120 OMI HCHO granules, 11.8 M px, 2.03 M "
              "pairs per launch)")
    md = subprocess.run([sys.executable, os.path.join(HERE, "tools", "ncu_summary.py"), rep, header],
                        capture_output=True, text=True).stdout
    open(os.path.join(PROF, RND + "_ncu_summary.md"), "w").write(md)

    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hh, un = rr[0], rr[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    kernels = {}
    for d in rr[2:]:
        def val(k):
            return float(d[hh.index(k)]) * scale[un[hh.index(k)]]
        kernels[short(d[hh.index("Kernel Name")].split("(")[0])] = {
            "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
            "duration_ms": float(d[hh.index("gpu__time_duration.sum")])}
    json.dump({"command": "python bench.py --steps 1 --warmup 1 --days 8 --no-e2e --no-cpu-baseline",
               "n_px": 120 * 98640, "kernels": kernels,
               "source": "profiles/" + RND + "_ncu_summary.md (ncu --set full, one launch each)"},
              open(os.path.join(PROF, RND + "_traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, RND + "_launches_summary.md")).read())


if __name__ == "__main__":
    main()
