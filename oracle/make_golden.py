"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the
UNMODIFIED reference (/root/reference, through oracle/ref_shim.py) on the seeded
cases of tests/cases.py.  Run in the build container only:

    python -m oracle.make_golden

The reference cannot travel to the GPU box; these fixtures (and this script)
do.  Each file holds the reference's outputs plus a checksum of the generated
inputs, so a test can tell "inputs drifted" from "outputs differ".
3-D fields are stored for the levels in cases.LEVEL_SUBSET to keep fixtures small, except
in the cases of cases.ALL_LEVELS, which keep every level.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oisatgmi_b200 import config  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def checksum(granules) -> str:
    h = hashlib.sha256()
    for g in granules:
        for v in config.field_values(g):
            if isinstance(v, np.ndarray):
                h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _with_checksum(store, granules):
    out = {"input_sha256": np.array(checksum(granules))}
    out.update(store)
    return out


def amf_chain(ref, name):
    """tests/chains.py drives the reference's own functions (chains.reference_impl): the
    fixtures hold exactly the keys the tests compare."""
    import chains
    return _with_checksum(chains.amf_chain(chains.reference_impl(), name)[0],
                          cases.amf_case(name)["granules"])


def mopitt_chain(ref):
    import chains
    return _with_checksum(chains.mopitt_chain(chains.reference_impl())[0],
                          cases.mopitt_case()["granules"])


def gosat_chain(ref):
    import chains
    return _with_checksum(chains.gosat_chain(chains.reference_impl())[0],
                          cases.gosat_case()["granules"])


def ssmis_chain(ref, fine):
    import chains
    store = chains.ssmis_chain(chains.reference_impl(), fine)[0]
    v = cases.ssmis_case(fine)["vars"]
    h = hashlib.sha256()
    for k in sorted(v):
        h.update(np.ascontiguousarray(v[k]).tobytes())
    out = {"input_sha256": np.array(h.hexdigest())}
    out.update(store)
    return out


def o3_chain(ref):
    import chains
    return _with_checksum(chains.o3_chain(chains.reference_impl())[0],
                          cases.o3_case()["granules"])


def reader_chain(ref, product):
    """Outputs of the reference's own reader functions (reader.py:707-983, 1130-1275) with their
    file access (`_read_group_nc`, `_read_nc`, the MOPITT file attributes) answered from
    cases.reader_vars().  GOSAT: the record is taken where the reader hands it to the gap
    filler (reader.py:1266), which has its own fixtures."""
    import chains
    rd = sys.modules["oisatgmi.reader"]
    fn = {"omi_no2": rd.omi_reader_no2, "omi_hcho": rd.omi_reader_hcho,
          "tropomi_no2": rd.tropomi_reader_no2, "mopitt_co": rd.mopitt_reader_co,
          "gosat_xch4": rd.gosat_reader_xch4}[product]
    v = cases.reader_vars(product)
    saved = {n: getattr(rd, n) for n in ("_read_group_nc", "_read_nc", "_get_nc_attr_group_mopitt",
                                         "filler_gosatxch4")}
    rd._read_group_nc = lambda fname, group, var: np.squeeze(np.array(v[var]))
    rd._read_nc = lambda fname, var: np.squeeze(np.array(v[var]))
    rd._get_nc_attr_group_mopitt = lambda fname: {"StartTime": v["StartTime"], "StopTime": v["StopTime"]}
    rd.filler_gosatxch4 = lambda grid_size, sat, flag_thresh=0.75: sat
    store = {}
    try:
        for prod, args in chains.READER_CALLS:
            if prod != product:
                continue
            if product in ("omi_hcho", "mopitt_co", "gosat_xch4"):
                r = quiet(fn, "dir/granule.nc", None, True)
            else:
                r = quiet(fn, "dir/granule.nc", args[0], None, True)
            chains.reader_record(store, "trop%d" % int(args[0]) if args else "all", r)
    finally:
        for n, f in saved.items():
            setattr(rd, n, f)
    return store


def main():
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    jobs = {name: (lambda n=name: amf_chain(ref, n)) for name in cases.CASES}
    jobs["mopitt_co"] = lambda: mopitt_chain(ref)
    jobs["gosat_xch4"] = lambda: gosat_chain(ref)
    jobs["omi_o3"] = lambda: o3_chain(ref)
    jobs["ssmis_pwv"] = lambda: ssmis_chain(ref, False)
    jobs["ssmis_pwv_fine"] = lambda: ssmis_chain(ref, True)
    for product in cases.READER_PRODUCTS:
        jobs["reader_" + product] = lambda p=product: reader_chain(ref, p)
    only = sys.argv[1:]
    for name, job in jobs.items():
        if only and name not in only:
            continue
        store = job()
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **store)
        print("%-12s %3d arrays  %7.1f kB" % (name, len(store), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
