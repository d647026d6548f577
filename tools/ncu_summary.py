#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one table per captured kernel.

    python tools/ncu_summary.py <report.ncu-rep> "<header line>" > profiles/<name>.md
"""
import csv
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("threads active per warp instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("regs/thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("dynamic smem/block", "launch__shared_mem_per_block_dynamic"),
    ("blocks/SM limit: registers", "launch__occupancy_limit_registers"),
    ("blocks/SM limit: shared memory", "launch__occupancy_limit_shared_mem"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("LSU data pipe % (wavefronts)", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("stall samples: long scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard"),
    ("stall samples: short scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard"),
    ("stall samples: wait (fixed latency)", "smsp__pcsamp_warps_issue_stalled_wait"),
    ("stall samples: barrier", "smsp__pcsamp_warps_issue_stalled_barrier"),
    ("stall samples: branch resolving", "smsp__pcsamp_warps_issue_stalled_branch_resolving"),
    ("stall samples: selected (issuing)", "smsp__pcsamp_warps_issue_stalled_selected"),
    ("samples", "smsp__pcsamp_sample_count"),
]


def main():
    rep = sys.argv[1]
    print(sys.argv[2] if len(sys.argv) > 2 else "# ncu summary of %s" % rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("\n## %s" % d[idx["Kernel Name"]].split("(")[0])
        print("| metric | value | unit |\n|---|---|---|")
        for label, key in METRICS:
            if key in idx:
                print("| %s (`%s`) | %s | %s |" % (label, key, d[idx[key]], units[idx[key]]))


if __name__ == "__main__":
    main()
