"""world_size-2 run of the N>1 decomposition on CPU (gloo): granules dealt by
day, local [10][n_cell] accumulators, one all-reduce, replicated finalisation.
Gridded granules come from the oracle here (there is no GPU in this suite);
what is under test is the sharding / merge logic of oisatgmi_b200/sharding.py
that bench.py and MonthPipeline use with NCCL."""
import datetime
import os
import tempfile

import numpy as np
import pytest

import cases
import chains


def _accumulate(grids):
    """numpy statement of K4 (csrc/k4_accum.cu): rows 0-4 sums, 5-9 counts."""
    n = grids[0].vcd.size
    acc = np.zeros((10, n))
    for g in grids:
        vals = [np.where(np.isinf(g.vcd), np.nan, g.vcd), g.uncertainty ** 2, g.ctm_vcd, g.new_amf,
                g.old_amf]
        vals[1] = np.where(np.isinf(vals[1]), np.nan, vals[1])
        for q, v in enumerate(vals):
            v = np.asarray(v, dtype=np.float64).ravel()
            ok = ~np.isnan(v)
            acc[q, ok] += v[ok]
            acc[5 + q, ok] += 1.0
    return acc


def _worker(rank, world, path, port):
    import torch
    import torch.distributed as dist
    from oisatgmi_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.load(path, allow_pickle=True)
    grids = list(data["grids"])
    mine = sharding.assign([g.time for g in grids], rank, world)
    acc = torch.from_numpy(_accumulate([grids[i] for i in mine]) if mine
                           else np.zeros((10, grids[0].vcd.size)))
    sharding.merge_accumulators(acc)
    if rank == 0:
        np.save(path + ".merged.npy", acc.numpy())
        np.save(path + ".owned.npy", np.array(mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank():
    import torch.multiprocessing as mp
    from oisatgmi_b200 import sharding
    _, grids = chains.amf_chain(chains.oracle_impl(), "omi_no2", stop_after="amf")
    # three granules on three different days -> ranks own {day0, day2} and {day1}
    times = [g.time for g in grids]
    assert sharding.assign(times, 0, 2) == [0, 2] and sharding.assign(times, 1, 2) == [1]
    assert sharding.assign(times, 0, 1) == [0, 1, 2]
    same_day = [datetime.datetime(2005, 6, 3, h) for h in (1, 5, 9)]
    assert sharding.assign(same_day, 0, 2) == [0, 1, 2] and sharding.assign(same_day, 1, 2) == []
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "grids.npz")
        with open(path, "wb") as f:
            np.savez(f, grids=np.array(grids, dtype=object))
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, path, port), nprocs=2, join=True)
        merged = np.load(path + ".merged.npy")
        owned = np.load(path + ".owned.npy")
    assert list(owned) == [0, 2]
    single = _accumulate(grids)
    assert np.array_equal(merged[5:], single[5:])            # counts: exact integers
    np.testing.assert_allclose(merged[:5], single[:5], rtol=1e-12, atol=0)
    # and the merged block finalises to the oracle's monthly means
    from oracle import averaging as oavg
    want = oavg.averaging("2005-06-01", "2005-07-01", cases.reader_ns(grids))
    with np.errstate(all="ignore"):
        np.testing.assert_allclose((merged[0] / merged[5]).reshape(want[0].shape), want[0],
                                   rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(np.sqrt(merged[1] / merged[6] ** 2).reshape(want[1].shape),
                                   want[1], rtol=1e-12, equal_nan=True)
