"""The oracle restatement against the golden fixtures the UNMODIFIED reference
produced (oracle/make_golden.py).  This is what pins the oracle on machines
where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

import cases
import chains
from util import assert_field

# fields that pass through numpy's float32 log (amf_recal.py:108) can move in the
# last float32 ulp between CPUs with different SIMD dispatch; everything else is
# float64 arithmetic in a fixed order
LOOSE = ("amf", "ak", "pwv", "avg.", "oi")


def _compare(store, gold):
    assert set(k for k in gold if k != "input_sha256") == set(store), \
        sorted(set(gold) ^ set(store))
    for k, v in store.items():
        rtol = 1e-6 if k.startswith(LOOSE) else 1e-12
        assert_field(v, gold[k], k, rtol=rtol)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_amf_chain_matches_reference_fixture(name, golden):
    store, _ = chains.amf_chain(chains.oracle_impl(), name)
    _compare(store, golden(name))


def test_mopitt_chain_matches_reference_fixture(golden):
    store, _ = chains.mopitt_chain(chains.oracle_impl())
    _compare(store, golden("mopitt_co"))


def test_gosat_chain_matches_reference_fixture(golden):
    store, _ = chains.gosat_chain(chains.oracle_impl())
    _compare(store, golden("gosat_xch4"))


def test_o3_chain_matches_reference_fixture(golden):
    store, _ = chains.o3_chain(chains.oracle_impl())
    _compare(store, golden("omi_o3"))


@pytest.mark.parametrize("fine", [False, True])
def test_ssmis_chain_matches_reference_fixture(fine, golden):
    store, _ = chains.ssmis_chain(chains.oracle_impl(), fine)
    _compare(store, golden("ssmis_pwv_fine" if fine else "ssmis_pwv"))


def test_inputs_did_not_drift(golden):
    import hashlib
    from oisatgmi_b200 import config
    for name in cases.CASES:
        c = cases.amf_case(name)
        h = hashlib.sha256()
        for g in c["granules"]:
            for v in config.field_values(g):
                if isinstance(v, np.ndarray):
                    h.update(np.ascontiguousarray(v).tobytes())
        assert h.hexdigest() == str(golden(name)["input_sha256"]), name


@pytest.mark.parametrize("product", cases.READER_PRODUCTS)
def test_reader_front_end_matches_reference_fixture(product, golden):
    from oracle import reader as oreader
    chains.same_bits(chains.reader_chain(oreader, product), golden("reader_" + product))
