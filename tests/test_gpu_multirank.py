"""N-rank equality of the sharded month: run with
`gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... tests/multirank_check.py`;
under plain pytest on one GPU the two ranks are emulated as two pipelines whose
accumulator blocks are summed, which exercises the same merge arithmetic."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def test_sharded_accumulators_equal_single_pipeline():
    from oisatgmi_b200 import _dev, sharding
    from oisatgmi_b200.pipeline import MonthPipeline
    c = cases.amf_case("omi_no2")

    def pipe_for(idx):
        p = MonthPipeline(c["ctm"], c["grid_size"], c["flag_thresh"], sensor="OMI", gas="NO2")
        for i in idx:
            assert p.add_granule(cases.clone(c["granules"][i]))
        p.allocate()
        p.run_prepare(); p.run_pack(); p.run_alive(); p.run_fused(); p.run_accumulate()
        return p

    times = [g.time for g in c["granules"]]
    full = pipe_for(range(len(times)))
    parts = [pipe_for(sharding.assign(times, r, 2)) for r in range(2)]
    merged = parts[0]._buf["acc"] + parts[1]._buf["acc"]
    a, b = _dev.to_host(merged), _dev.to_host(full._buf["acc"])
    assert np.array_equal(a[5:], b[5:])                       # counts exact
    np.testing.assert_allclose(a[:5], b[:5], rtol=1e-12, atol=0)
    # replicated OI on the merged block picks the same knee and the same analysis
    parts[0]._buf["acc"].copy_(merged)
    r0 = parts[0].results_to_host(parts[0].run_oi())
    r1 = full.results_to_host(full.run_oi())
    assert r0["knee_index"] == r1["knee_index"]
    np.testing.assert_allclose(r0["ctm_averaged_vcd_corrected"], r1["ctm_averaged_vcd_corrected"],
                               rtol=1e-12, equal_nan=True)
