#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel.

ncu's CSV source page is SASS-only; this joins it with `nvdisasm --print-line-info`
of the same cubin (built with -lineinfo) by instruction offset.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> <mangled-substring> [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, want, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(here, "oisatgmi_b200", "csrc", "liboisat.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
lines = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)],
                         capture_output=True, text=True).stdout
    cur, loc = None, None
    for ln in out.splitlines():
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            cur = m.group(1) if mangled in m.group(1) else None
            loc = None
            continue
        if cur is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines[int(m.group(1), 16)] = (loc, m.group(2).strip())
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                        text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
sec, hdr, body = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        if body:
            break
        sec = r[1] if want in r[1] else None
        hdr = None
        continue
    if sec is None:
        continue
    if hdr is None and r and r[0] == "Address":
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
base = int(body[0][ci["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in body:
    off = int(r[ci["Address"]], 16) - base
    loc = lines.get(off, (None, ""))[0]
    n, s = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
    agg[loc][0] += n
    agg[loc][1] += s
    tot_i += n
    tot_s += s
print("kernel:", sec, "instructions", tot_i, "samples", tot_s)
src_cache = {}
for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if loc:
        path = os.path.join(here, "oisatgmi_b200", "csrc", loc[0])
        if path not in src_cache and os.path.exists(path):
            src_cache[path] = open(path).read().splitlines()
        if path in src_cache and loc[1] - 1 < len(src_cache[path]):
            text = src_cache[path][loc[1] - 1].strip()[:70]
    print("%5.1f%% inst %5.1f%% smp  %-22s %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1),
                                               "%s:%d" % loc if loc else "?", text))
