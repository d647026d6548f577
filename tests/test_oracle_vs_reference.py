"""Pins the oracle restatement against the LIVE, unmodified reference
(/root/reference through oracle/ref_shim.py).  Runs only in the build container;
on the GPU box the golden fixtures (tests/test_golden_oracle.py) take over."""
import contextlib
import io
import types

import numpy as np
import pytest

import cases
import chains
from oisatgmi_b200 import config
from oracle import ref_shim
from util import assert_field

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason="reference tree not present on this machine")


reference_impl = chains.reference_impl


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        assert_field(a[k], b[k], k, rtol=0.0)      # bit-identical


@pytest.mark.parametrize("name", list(cases.CASES))
def test_amf_chain_is_bit_identical_to_reference(name):
    _same(chains.amf_chain(chains.oracle_impl(), name)[0], chains.amf_chain(reference_impl(), name)[0])


def test_mopitt_chain_is_bit_identical_to_reference():
    _same(chains.mopitt_chain(chains.oracle_impl())[0], chains.mopitt_chain(reference_impl())[0])


def test_gosat_chain_is_bit_identical_to_reference():
    _same(chains.gosat_chain(chains.oracle_impl())[0], chains.gosat_chain(reference_impl())[0])


def test_o3_chain_is_bit_identical_to_reference():
    _same(chains.o3_chain(chains.oracle_impl())[0], chains.o3_chain(reference_impl())[0])


def test_bias_table_matches_driver():
    ref = ref_shim.load_reference()
    from oracle import oi as ooi
    d = ref.driver.oisatgmi()
    for (sensor, gas), (a, b) in ooi.BIAS.items():
        d.sat_averaged_vcd = np.array([[2.0, 5.0]])
        with contextlib.redirect_stdout(io.StringIO()):
            d.bias_correct(sensor, gas)
        assert np.array_equal(d.sat_averaged_vcd, (np.array([[2.0, 5.0]]) - a) / b)


@pytest.mark.parametrize("product", cases.READER_PRODUCTS)
def test_reader_front_end_is_bit_identical_to_reference(product):
    """The unmodified reader functions of the reference (reader.py:707-983, 1130-1275), their
    file access answered from a dictionary (oracle/make_golden.py:reader_chain), against the
    restatement: every field, dtype and NaN bit for bit."""
    from oracle import make_golden, reader as oreader
    want = make_golden.reader_chain(ref_shim.load_reference(), product)
    chains.same_bits(chains.reader_chain(oreader, product), want)


def test_output_fields_are_bit_identical_to_write_to_nc(tmp_path):
    """driver.write_to_nc (driver.py:156-227), unmodified, writing into a recording
    stand-in for netCDF4.Dataset: every stored array equals oracle.output's."""
    import sys
    from oracle import output as ooutput
    ref_shim.load_reference()
    drv = sys.modules["oisatgmi.driver"]
    stored = {}

    class Var:
        def __init__(self, name, kind):
            self.name, self.kind = name, kind

        def __setitem__(self, key, value):
            a = np.asarray(value)
            stored[self.name] = a.astype(np.float32) if self.kind == "f" else a

    class RecordingDataset:
        def __init__(self, path, mode):
            pass

        def createDimension(self, *a):
            pass

        def createVariable(self, name, kind, dims):
            return Var(name, kind)

        def close(self):
            pass

    obj = chains.month_object(chains.oracle_impl())
    ref_obj = drv.oisatgmi()
    ref_obj.__dict__.update(obj.__dict__)
    saved = drv.Dataset
    drv.Dataset = RecordingDataset
    try:
        ref_obj.write_to_nc("month", output_folder=str(tmp_path))
    finally:
        drv.Dataset = saved
    got = ooutput.output_fields(obj)
    assert set(stored) == set(got)
    for k, a in stored.items():
        if k == "time":
            assert b"".join(a.tolist()).decode() == got[k]
        else:
            assert a.dtype == got[k].dtype == np.float32
            assert np.array_equal(a, got[k], equal_nan=True), k
