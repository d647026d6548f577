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


def test_ext_fields_are_bit_identical_to_convert2EXT(tmp_path, monkeypatch):
    """tools/convert2EXT.py, unmodified, run as the script it is against recording stand-ins
    for netCDF4.Dataset (read side: one diagnostics file held in memory; write side: every
    assignment recorded): the variables of the ExtData files equal oracle.output.ext_fields
    and oisatgmi_b200.ext_output (a host-only module), for the real month and for the all-ones
    months the script fabricates."""
    import runpy
    import sys
    import types
    from oisatgmi_b200 import ext_output
    from oracle import output as ooutput
    rng = np.random.default_rng(5)
    lat, lon = np.meshgrid(np.arange(-90.0, 90.5, 10.0), np.arange(-180.0, 180.0, 20.0), indexing="ij")
    diag = {"lat": lat.astype(np.float32), "lon": lon.astype(np.float32),
            "scaling_factor": rng.uniform(0.5, 2.0, lat.shape).astype(np.float32)}
    diag_dir, ext_dir = tmp_path / "diag", tmp_path / "ext"
    diag_dir.mkdir()
    (diag_dir / "HCHO_200506.nc").write_bytes(b"")
    written = {}

    class Var:
        def __init__(self, store, name):
            self.store, self.name = store, name

        def __setitem__(self, key, value):
            self.store[self.name] = np.array(np.broadcast_to(
                np.asarray(value, dtype=np.float64), self.store["_shape_" + self.name]))

    class FakeDataset:
        def __init__(self, path, mode, format=None):
            self.mode = mode
            if mode == "r":
                self.variables = diag
            else:
                self.store = written.setdefault(path.split("/")[-1], {})
                self.dims = {}

        def createDimension(self, name, size):
            self.dims[name] = size
            return size

        def createVariable(self, name, kind, dims):
            assert kind == "f8"
            self.store["_shape_" + name] = tuple(self.dims[d] for d in dims)
            v = Var(self.store, name)
            object.__setattr__(self, "_last", v)
            return _Attr(v, self.store)

        def close(self):
            pass

    class _Attr:
        """Variable handle that records attributes (units) as well as data."""
        def __init__(self, var, store):
            object.__setattr__(self, "_var", var)
            object.__setattr__(self, "_store", store)

        def __setitem__(self, key, value):
            self._var[key] = value

        def __setattr__(self, name, value):
            self._store[self._var.name + "." + name] = value

    monkeypatch.setitem(sys.modules, "netCDF4", types.SimpleNamespace(Dataset=FakeDataset))
    monkeypatch.setattr(sys, "argv", ["convert2EXT.py", str(diag_dir), str(ext_dir)])
    with contextlib.redirect_stdout(io.StringIO()):
        runpy.run_path(ref_shim.REFERENCE_ROOT + "/tools/convert2EXT.py", run_name="__main__")
    assert "HCHO_200506.nc" in written and "HCHO_199001.nc" in written

    def check(store, want):
        for k in ("time", "lat", "lon", "SF"):
            assert store[k].dtype == want[k].dtype == np.float64 and store[k].shape == want[k].shape, k
            assert np.array_equal(store[k], want[k]), k
        assert store["time.units"] == want["time_units"]

    real = written["HCHO_200506.nc"]
    check(real, ooutput.ext_fields(diag["lat"], diag["lon"], diag["scaling_factor"], 2005, 6))
    check(real, ext_output.ext_fields(diag, "200506"))
    fake = written["HCHO_199001.nc"]
    check(fake, ext_output.ones_fields(diag["lat"], diag["lon"], 1990, 1))
    assert np.all(fake["SF"] == 1.0)


@pytest.mark.parametrize("fine", [False, True])
def test_ssmis_chain_is_bit_identical_to_reference(fine):
    _same(chains.ssmis_chain(chains.oracle_impl(), fine)[0],
          chains.ssmis_chain(reference_impl(), fine)[0])


def test_drop_in_signatures_match_the_reference():
    """The switch INTEGRATION.md section 1 describes rebinds the reference's imports to the
    drop-in modules: every replaced callable must take the same parameters (names, order,
    defaults) as the one it replaces, and the `oisatgmi` class the same methods."""
    import importlib
    import inspect
    ref_shim.load_reference()
    pairs = [("interpolator", "interpolator"), ("interpolator", "_upscaler"),
             ("filler_gosat", "filler_gosatxch4"), ("amf_recal", "amf_recal"),
             ("ak_conv_mopitt", "ak_conv_mopitt"), ("ak_conv_gosat", "ak_conv_gosat"),
             ("averaging", "averaging"), ("optimal_interpolation", "OI"),
             ("interpolator_ssmis", "interpolator_ssmis"), ("pwv_cal", "pwv_calculator")]
    for mod, name in pairs:
        ref = getattr(importlib.import_module("oisatgmi." + mod), name)
        ours = getattr(importlib.import_module("oisatgmi_b200." + mod), name)
        a, b = inspect.signature(ref), inspect.signature(ours)
        assert [(p.name, p.default) for p in a.parameters.values()] == \
               [(p.name, p.default) for p in b.parameters.values()], (mod, name, str(a), str(b))
    ref_cls = importlib.import_module("oisatgmi.driver").oisatgmi
    our_cls = importlib.import_module("oisatgmi_b200.driver").oisatgmi
    for meth in ("read_data", "recal_amf", "cal_pwv", "conv_ak", "average", "bias_correct", "oi",
                 "write_to_nc"):
        a, b = inspect.signature(getattr(ref_cls, meth)), inspect.signature(getattr(our_cls, meth))
        assert [(p.name, p.default) for p in a.parameters.values()] == \
               [(p.name, p.default) for p in b.parameters.values()], (meth, str(a), str(b))


def test_mopitt_chain_on_a_coarse_model_is_bit_identical_to_reference():
    """The other branch of _upscaler / ak_conv_mopitt: a model coarser than the L3 lattice."""
    _same(chains.mopitt_chain(chains.oracle_impl(), coarse=True)[0],
          chains.mopitt_chain(reference_impl(), coarse=True)[0])
