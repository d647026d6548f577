"""Run the same chain of stages through a given implementation (oracle or the
CUDA drop-ins) on the seeded cases; returns a flat {name: array} store with the
same keys oracle/make_golden.py writes."""
import numpy as np

import cases


def put(store, prefix, obj, names):
    for n in names:
        v = getattr(obj, n)
        if isinstance(v, np.ndarray) and v.size > 1:
            store["%s.%s" % (prefix, n)] = cases.subset_levels(v)


def amf_chain(impl, name, stop_after=None):
    """impl: namespace with interpolator, amf_recal, averaging, OI, bias (dict)."""
    c = cases.amf_case(name)
    store = {}
    grids = []
    for i, g in enumerate(c["granules"]):
        r = impl.interpolator(c["kind"], c["grid_size"], cases.clone(g), c["coords"],
                              flag_thresh=c["flag_thresh"])
        assert r is not None
        put(store, "interp%d" % i, r, ["vcd", "amf", "tropopause", "uncertainty", "pressure_mid",
                                        "scattering_weights"])
        grids.append(r)
    if stop_after == "interp":
        return store, grids
    grids = impl.amf_recal(c["ctm"], grids)
    for i, r in enumerate(grids):
        put(store, "amf%d" % i, r, ["vcd", "ctm_vcd", "new_amf", "old_amf"])
    if stop_after == "amf":
        return store, grids
    avg = impl.averaging("2005-06-01", "2005-07-01", cases.reader_ns(grids))
    for n, v in zip(["sat_vcd", "sat_err", "ctm_vcd", "aux1", "aux2"], avg[:5]):
        store["avg." + n] = v
    a, b = impl.bias.get((c["sensor"], c["gas"]), (None, None))
    y = np.array(avg[0]) if a is None else (np.array(avg[0]) - a) / b
    xa = np.array(avg[2])
    res = impl.OI(xa, y, (xa * 50.0 / 100.0) ** 2, np.array(avg[1]) ** 2)
    store["oi.y"] = y
    for n, v in zip(["ctm_averaged_vcd_corrected", "ak_OI", "increment_OI", "error_OI"], res[:4]):
        store["oi." + n] = v
    y2 = np.array(avg[0])
    r2 = impl.OI(np.array(avg[2]), y2, (np.array(avg[2]) * 0.5) ** 2, np.array(avg[1]) ** 2,
                 regularization_on=False)
    for n, v in zip(["xb", "ak", "inc", "err"], r2[:4]):
        store["oi_noreg." + n] = v
    return store, grids


def mopitt_chain(impl):
    c = cases.mopitt_case()
    store = {}
    grids = []
    for i, g in enumerate(c["granules"]):
        r = impl.interpolator(1, c["grid_size"], cases.clone(g), c["coords"],
                              flag_thresh=c["flag_thresh"])
        assert r is not None and r.ctm_upscaled_needed
        put(store, "interp%d" % i, r, ["vcd", "uncertainty", "x_col", "aprior_column",
                                        "surface_pressure", "apriori_surface", "pressure_mid",
                                        "averaging_kernels", "apriori_profile"])
        grids.append(r)
    grids = impl.ak_conv_mopitt(c["ctm"], grids)
    for i, r in enumerate(grids):
        put(store, "ak%d" % i, r, ["ctm_vcd", "ctm_xcol"])
    return store, grids


def gosat_chain(impl):
    c = cases.gosat_case()
    store = {}
    grids = []
    for i, g in enumerate(c["granules"]):
        f = impl.filler_gosatxch4(1.0, cases.clone(g), flag_thresh=0.0)
        assert f is not None
        put(store, "fill%d" % i, f, ["vcd", "x_col", "uncertainty", "quality_flag", "pressure_mid",
                                      "averaging_kernels", "apriori_profile", "pressure_weight"])
        r = impl.interpolator(1, c["grid_size"], f, c["coords"], flag_thresh=0.0)
        assert r is not None and r.ctm_upscaled_needed
        put(store, "interp%d" % i, r, ["vcd", "uncertainty", "x_col", "pressure_mid",
                                        "averaging_kernels", "apriori_profile", "pressure_weight"])
        grids.append(r)
    grids = impl.ak_conv_gosat(c["ctm"], grids)
    for i, r in enumerate(grids):
        put(store, "ak%d" % i, r, ["ctm_xcol"])
    return store, grids


def oracle_impl():
    import types
    from oracle import averaging as oavg, interp as ointerp, oi as ooi, vertical as overt
    return types.SimpleNamespace(
        interpolator=ointerp.interpolator, filler_gosatxch4=ointerp.filler_gosatxch4,
        amf_recal=overt.amf_recal, ak_conv_mopitt=overt.ak_conv_mopitt,
        ak_conv_gosat=overt.ak_conv_gosat, averaging=oavg.averaging, OI=ooi.OI, bias=ooi.BIAS)


def cuda_impl():
    import types
    from oisatgmi_b200 import (ak_conv_gosat, ak_conv_mopitt, amf_recal, averaging, driver,
                               filler_gosat, interpolator, optimal_interpolation)
    return types.SimpleNamespace(
        interpolator=interpolator.interpolator, filler_gosatxch4=filler_gosat.filler_gosatxch4,
        amf_recal=amf_recal.amf_recal, ak_conv_mopitt=ak_conv_mopitt.ak_conv_mopitt,
        ak_conv_gosat=ak_conv_gosat.ak_conv_gosat, averaging=averaging.averaging,
        OI=optimal_interpolation.OI, bias=driver.BIAS_CORRECTION)


READER_CALLS = [("omi_no2", (True,)), ("omi_no2", (False,)), ("omi_hcho", ()),
                ("tropomi_no2", (True,)), ("tropomi_no2", (False,))]
READER_FIELDS = ("vcd", "amf", "tropopause", "latitude_center", "longitude_center", "uncertainty",
                 "quality_flag", "pressure_mid", "scattering_weights")


def reader_chain(module, product):
    """Reader front-end of one product through `module` (oracle.reader or
    oisatgmi_b200.reader_frontend): same keys as oracle/make_golden.py's reader_chain.
    Arrays keep their dtypes: the comparison is bit for bit."""
    v = cases.reader_vars(product)
    store = {}
    for prod, args in READER_CALLS:
        if prod != product:
            continue
        r = getattr(module, product)(v, *args)
        tag = "trop%d" % int(args[0]) if args else "all"
        store[tag + ".time"] = np.array(r.time.isoformat())
        for n in READER_FIELDS:
            a = np.asarray(getattr(r, n))
            if a.size > 1:
                store["%s.%s" % (tag, n)] = a
    return store


def same_bits(store, gold):
    assert set(store) == set(gold), sorted(set(store) ^ set(gold))
    for k, a in store.items():
        b = gold[k]
        assert a.dtype == b.dtype and a.shape == b.shape, (k, a.dtype, b.dtype, a.shape, b.shape)
        assert np.array_equal(a, b, equal_nan=(a.dtype.kind == "f")), k


def month_object(impl, name="omi_no2"):
    """Attributes `driver.oisatgmi` holds after average() -> bias_correct() -> oi(), from
    the amf chain of `impl` (used for the write_to_nc data side, driver.py:156-227)."""
    import datetime
    import types
    store, grids = amf_chain(impl, name)
    o = types.SimpleNamespace(
        sat_averaged_vcd=store["oi.y"], sat_averaged_error=store["avg.sat_err"],
        ctm_averaged_vcd=store["avg.ctm_vcd"].copy(), aux1=store["avg.aux1"], aux2=store["avg.aux2"],
        ctm_averaged_vcd_corrected=store["oi.ctm_averaged_vcd_corrected"], ak_OI=store["oi.ak_OI"],
        increment_OI=store["oi.increment_OI"], error_OI=store["oi.error_OI"],
        avg_time=datetime.datetime(2005, 6, 16, 12, 0, 0),
        reader_obj=types.SimpleNamespace(sat_data=[None] + list(grids)))
    # the cases the clean-up of the scaling factor exists for (driver.py:204-206)
    o.ctm_averaged_vcd[0, :4] = [0.0, np.nan, np.inf, 1.0]
    o.ctm_averaged_vcd_corrected = o.ctm_averaged_vcd_corrected.copy()
    o.ctm_averaged_vcd_corrected[0, :4] = [1.0, 2.0, 3.0, 0.0]
    return o
