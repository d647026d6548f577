"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference package.

`/root/reference` exists only in the build container (never on the GPU box), so
everything here is used solely by `oracle/make_golden.py` (fixture generation)
and by `-m "not gpu"` tests that pin the oracle restatement against the live
reference.  Nothing under `oisatgmi_b200/` may import this module.

The reference does not import out of the box in this image:
  1. `oisatgmi/__init__.py:1` pulls in driver -> reader/report, which need
     netCDF4, h5py, fpdf, matplotlib, basemap (absent) -> empty stub modules;
  2. `oisatgmi/interpolator.py:7` imports a private scipy symbol from its
     pre-1.14 location -> re-attached from `scipy.interpolate._interpnd`;
  3. `oisatgmi/optimal_interpolation.py:3` imports third-party `kneed`
     (pinned kneed==0.8.3, requirements.txt:9; not installed, not vendored)
     -> served by the oracle's own Kneedle restatement (oracle/kneedle.py).
No file under /root/reference is modified or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OISAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "oisatgmi"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Return a namespace with the reference's hot-path callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    import scipy.interpolate._interpnd as _p
    import scipy.interpolate.interpnd as _pub
    if not hasattr(_pub, "_ndim_coords_from_arrays"):
        _pub._ndim_coords_from_arrays = _p._ndim_coords_from_arrays

    class _Missing:  # placeholder for I/O classes the hot path never touches
        def __init__(self, *a, **k):
            raise RuntimeError("I/O dependency stubbed out in the oracle harness")

    _stub("netCDF4", Dataset=_Missing)
    _stub("h5py", File=_Missing)
    _stub("fpdf", FPDF=_Missing)
    mpl = _stub("matplotlib", colormaps={})
    mpl.__path__ = []  # behave like a package
    _stub("matplotlib.pyplot")
    _stub("matplotlib.ticker", FormatStrFormatter=_Missing)
    tk = _stub("mpl_toolkits")
    tk.__path__ = []
    _stub("mpl_toolkits.basemap", Basemap=_Missing)

    if "kneed" not in sys.modules:
        try:
            importlib.import_module("kneed")
        except Exception:
            from oracle import kneedle as _kn
            _stub("kneed", KneeLocator=_kn.KneeLocator)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.config = importlib.import_module("oisatgmi.config")
    ns.interpolator_mod = importlib.import_module("oisatgmi.interpolator")
    ns.interpolator = ns.interpolator_mod.interpolator
    ns._upscaler = ns.interpolator_mod._upscaler
    ns.filler_gosatxch4 = importlib.import_module("oisatgmi.filler_gosat").filler_gosatxch4
    ns.amf_recal = importlib.import_module("oisatgmi.amf_recal").amf_recal
    ns.ak_conv_mopitt = importlib.import_module("oisatgmi.ak_conv_mopitt").ak_conv_mopitt
    ns.ak_conv_gosat = importlib.import_module("oisatgmi.ak_conv_gosat").ak_conv_gosat
    av = importlib.import_module("oisatgmi.averaging")
    ns.averaging = av.averaging
    ns.error_averager = av.error_averager
    ns.OI = importlib.import_module("oisatgmi.optimal_interpolation").OI
    ns.driver = importlib.import_module("oisatgmi.driver")
    ns.interpolator_ssmis = importlib.import_module("oisatgmi.interpolator_ssmis").interpolator_ssmis
    ns.pwv_calculator = importlib.import_module("oisatgmi.pwv_cal").pwv_calculator
    _loaded = ns
    return ns
