"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the
UNMODIFIED reference (/root/reference, through oracle/ref_shim.py) on the seeded
cases of tests/cases.py.  Run in the build container only:

    python -m oracle.make_golden

The reference cannot travel to the GPU box; these fixtures (and this script)
do.  Each file holds the reference's outputs plus a checksum of the generated
inputs, so a test can tell "inputs drifted" from "outputs differ".
3-D fields are stored for the levels in cases.LEVEL_SUBSET to keep fixtures small.
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


def put(store, prefix, obj, names):
    for n in names:
        v = getattr(obj, n)
        if isinstance(v, np.ndarray) and v.size > 1:
            store["%s.%s" % (prefix, n)] = cases.subset_levels(v)


def amf_chain(ref, name):
    c = cases.amf_case(name)
    store = {"input_sha256": np.array(checksum(c["granules"]))}
    rcls = ref.config.satellite_amf
    grids = []
    for i, g in enumerate(c["granules"]):
        r = quiet(ref.interpolator, c["kind"], c["grid_size"], config.convert(cases.clone(g), rcls),
                  c["coords"], flag_thresh=c["flag_thresh"])
        assert r is not None
        put(store, "interp%d" % i, r, ["vcd", "amf", "tropopause", "uncertainty", "pressure_mid",
                                        "scattering_weights"])
        grids.append(r)
    rctm = [config.convert(m, ref.config.ctm_model) for m in c["ctm"]]
    grids = quiet(ref.amf_recal, rctm, grids)
    for i, r in enumerate(grids):
        put(store, "amf%d" % i, r, ["vcd", "ctm_vcd", "new_amf", "old_amf"])
    avg = quiet(ref.averaging, "2005-06-01", "2005-07-01", cases.reader_ns(grids))
    for n, v in zip(["sat_vcd", "sat_err", "ctm_vcd", "aux1", "aux2"], avg[:5]):
        store["avg." + n] = v
    d = ref.driver.oisatgmi()
    d.sat_averaged_vcd, d.sat_averaged_error, d.ctm_averaged_vcd, d.aux1, d.aux2 = \
        [np.array(v) for v in avg[:5]]
    quiet(d.bias_correct, c["sensor"], c["gas"])
    quiet(d.oi, c["sensor"], 50.0)
    store["oi.y"] = d.sat_averaged_vcd
    for n in ["ctm_averaged_vcd_corrected", "ak_OI", "increment_OI", "error_OI"]:
        store["oi." + n] = getattr(d, n)
    # knee-free variant pins everything in OI except the third-party knee
    y2 = np.array(avg[0])
    r2 = quiet(ref.OI, np.array(avg[2]), y2, (np.array(avg[2]) * 0.5) ** 2, np.array(avg[1]) ** 2,
               regularization_on=False)
    for n, v in zip(["xb", "ak", "inc", "err"], r2):
        store["oi_noreg." + n] = v
    return store


def mopitt_chain(ref):
    c = cases.mopitt_case()
    store = {"input_sha256": np.array(checksum(c["granules"]))}
    grids = []
    for i, g in enumerate(c["granules"]):
        r = quiet(ref.interpolator, 1, c["grid_size"],
                  config.convert(cases.clone(g), ref.config.satellite_opt), c["coords"],
                  flag_thresh=c["flag_thresh"])
        assert r is not None and r.ctm_upscaled_needed
        put(store, "interp%d" % i, r, ["vcd", "uncertainty", "x_col", "aprior_column",
                                        "surface_pressure", "apriori_surface", "pressure_mid",
                                        "averaging_kernels", "apriori_profile"])
        grids.append(r)
    rctm = [config.convert(m, ref.config.ctm_model) for m in c["ctm"]]
    grids = quiet(ref.ak_conv_mopitt, rctm, grids)
    for i, r in enumerate(grids):
        put(store, "ak%d" % i, r, ["ctm_vcd", "ctm_xcol"])
    return store


def gosat_chain(ref):
    c = cases.gosat_case()
    store = {"input_sha256": np.array(checksum(c["granules"]))}
    grids = []
    for i, g in enumerate(c["granules"]):
        f = quiet(ref.filler_gosatxch4, 1.0, config.convert(cases.clone(g), ref.config.satellite_opt),
                  flag_thresh=0.0)
        assert f is not None
        put(store, "fill%d" % i, f, ["vcd", "x_col", "uncertainty", "quality_flag", "pressure_mid",
                                      "averaging_kernels", "apriori_profile", "pressure_weight"])
        r = quiet(ref.interpolator, 1, c["grid_size"], f, c["coords"], flag_thresh=0.0)
        assert r is not None and r.ctm_upscaled_needed
        put(store, "interp%d" % i, r, ["vcd", "uncertainty", "x_col", "pressure_mid",
                                        "averaging_kernels", "apriori_profile", "pressure_weight"])
        grids.append(r)
    rctm = [config.convert(m, ref.config.ctm_model) for m in c["ctm"]]
    grids = quiet(ref.ak_conv_gosat, rctm, grids)
    for i, r in enumerate(grids):
        put(store, "ak%d" % i, r, ["ctm_xcol"])
    return store


READER_CALLS = [("omi_no2", (True,)), ("omi_no2", (False,)), ("omi_hcho", ()),
                ("tropomi_no2", (True,)), ("tropomi_no2", (False,))]
READER_FIELDS = ("vcd", "amf", "tropopause", "latitude_center", "longitude_center", "uncertainty",
                 "quality_flag", "pressure_mid", "scattering_weights")


def reader_chain(ref, product):
    """Outputs of the reference's own reader functions (reader.py:707-983) with their file
    access (`_read_group_nc`) answered from cases.reader_vars()."""
    rd = sys.modules["oisatgmi.reader"]
    fn = {"omi_no2": rd.omi_reader_no2, "omi_hcho": rd.omi_reader_hcho,
          "tropomi_no2": rd.tropomi_reader_no2}[product]
    v = cases.reader_vars(product)
    saved = rd._read_group_nc
    rd._read_group_nc = lambda fname, group, var: np.squeeze(np.array(v[var]))
    store = {}
    try:
        for prod, args in READER_CALLS:
            if prod != product:
                continue
            r = quiet(fn, "dir/granule.nc", None, True) if product == "omi_hcho" else \
                quiet(fn, "dir/granule.nc", args[0], None, True)
            tag = "trop%d" % int(args[0]) if args else "all"
            store[tag + ".time"] = np.array(r.time.isoformat())
            for n in READER_FIELDS:
                a = np.asarray(getattr(r, n))
                if a.size > 1:
                    store["%s.%s" % (tag, n)] = a
    finally:
        rd._read_group_nc = saved
    return store


def main():
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    jobs = {name: (lambda n=name: amf_chain(ref, n)) for name in cases.CASES}
    jobs["mopitt_co"] = lambda: mopitt_chain(ref)
    jobs["gosat_xch4"] = lambda: gosat_chain(ref)
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
