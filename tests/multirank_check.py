"""Launched by torchrun on N GPUs: every rank processes its share of the granules
(dealt by day), one NCCL all-reduce merges the accumulators, and every rank must
end with the single-GPU result.  Prints one line per rank; exit code != 0 on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
from oisatgmi_b200 import sharding  # noqa: E402
from oisatgmi_b200.pipeline import MonthPipeline  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = cases.amf_case("omi_no2")
    times = [g.time for g in c["granules"]]

    def run(idx, pg):
        p = MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor="OMI", gas="NO2",
                          process_group=pg)
        for i in idx:
            assert p.add_granule(cases.clone(c["granules"][i]))
        return p, p.results_to_host(p.run())

    mine = sharding.assign(times, rank, world)
    # more ranks than days leaves some ranks without a granule: they must still join the
    # all-reduce with a zero block (MonthPipeline.run) and end with the same month
    p, res = run(mine, dist.group.WORLD)
    _, ref = run(range(len(times)), None)
    ok = res["knee_index"] == ref["knee_index"]
    for k in ("ctm_averaged_vcd", "sat_averaged_vcd", "sat_averaged_error", "aux1", "aux2",
              "ctm_averaged_vcd_corrected", "ak_OI", "error_OI"):
        a, b = res[k], ref[k]
        same_mask = np.array_equal(np.isnan(a), np.isnan(b))
        f = np.isfinite(a) & np.isfinite(b)
        rel = float(np.max(np.abs(a[f] - b[f]) / np.maximum(np.abs(b[f]), 1e-300))) if f.any() else 0
        ok = ok and same_mask and rel < 1e-12
        print("rank %d %-28s mask %s rel %.2e" % (rank, k, same_mask, rel), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d owns %s: %s" % (rank, mine, "OK" if ok else "MISMATCH"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
