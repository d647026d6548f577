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


READER_CALLS = [("omi_no2", (True,)), ("omi_no2", (False,)), ("omi_hcho", ()),
                ("tropomi_no2", (True,)), ("tropomi_no2", (False,))]
READER_FIELDS = ("vcd", "amf", "tropopause", "latitude_center", "longitude_center", "uncertainty",
                 "quality_flag", "pressure_mid", "scattering_weights")


def reference_reader(product, v, args):
    """The unmodified reader function of the reference, with its only file access
    (`_read_group_nc`, reader.py:51-67) answered from the dictionary `v`."""
    import sys
    ref_shim.load_reference()
    rd = sys.modules["oisatgmi.reader"]
    fn = {"omi_no2": rd.omi_reader_no2, "omi_hcho": rd.omi_reader_hcho,
          "tropomi_no2": rd.tropomi_reader_no2}[product]
    saved = rd._read_group_nc
    rd._read_group_nc = lambda fname, group, var: np.squeeze(np.array(v[var]))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if product == "omi_hcho":
                return fn("dir/granule.nc", None, True)
            return fn("dir/granule.nc", args[0], None, True)
    finally:
        rd._read_group_nc = saved


@pytest.mark.parametrize("product,args", READER_CALLS)
def test_reader_front_end_is_bit_identical_to_reference(product, args):
    from oracle import reader as oreader
    v = cases.reader_vars(product)
    want = reference_reader(product, v, args)
    got = getattr(oreader, product)(v, *args)
    assert want is not None and got.time == want.time
    for n in READER_FIELDS:
        a, b = np.asarray(getattr(got, n)), np.asarray(getattr(want, n))
        if b.size > 1:
            assert a.dtype == b.dtype and a.shape == b.shape, n
            assert np.array_equal(a, b, equal_nan=True), n


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
