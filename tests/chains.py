"""Run the same chain of stages through a given implementation (oracle or the
CUDA drop-ins) on the seeded cases; returns a flat {name: array} store with the
same keys oracle/make_golden.py writes."""
import numpy as np

import cases


def put(store, prefix, obj, names, all_levels=False):
    for n in names:
        v = getattr(obj, n)
        if isinstance(v, np.ndarray) and v.size > 1:
            store["%s.%s" % (prefix, n)] = v if all_levels else cases.subset_levels(v)


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
                                        "scattering_weights"], all_levels=name in cases.ALL_LEVELS)
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


def month_tail(impl, store, sensor, gas, ctm, grids, per_granule, skip_aux=False):
    """The rest of the month exactly as run/job.py:61-84 drives it, through the `oisatgmi`
    class of the implementation (driver.py:36-114): conv_ak / recal_amf -> average ->
    bias_correct -> oi.  `per_granule`: names stored after the vertical stage.
    skip_aux: granules without scattering weights carry np.empty(1) air-mass factors
    (amf_recal.py:169-170), so aux1 / aux2 are uninitialised memory."""
    d = impl.driver(ctm, grids)
    if sensor in ("MOPITT", "GOSAT"):
        d.conv_ak(sensor)
    elif sensor == "SSMIS":
        d.cal_pwv()
    else:
        d.recal_amf()
    grids = d.reader_obj.sat_data
    tag = {"MOPITT": "ak", "GOSAT": "ak", "SSMIS": "pwv"}.get(sensor, "amf")
    for i, r in enumerate(grids):
        put(store, "%s%d" % (tag, i), r, per_granule)
    d.average("2005-06-01", "2005-07-01", gasname=gas)
    names = ["sat_vcd", "sat_err", "ctm_vcd"] + ([] if skip_aux else ["aux1", "aux2"])
    for n, a in zip(names, ["sat_averaged_vcd", "sat_averaged_error", "ctm_averaged_vcd",
                            "aux1", "aux2"]):
        store["avg." + n] = np.array(getattr(d, a))
    d.bias_correct(sensor, gas)
    d.oi(sensor, error_ctm=50.0)
    # OI clips its Y argument in place (optimal_interpolation.py:14): the satellite mean for
    # every sensor but GOSAT, whose Y is aux1 (driver.py:113-114)
    store["oi.y"] = np.array(d.aux1 if sensor == "GOSAT" else d.sat_averaged_vcd)
    for n in ["ctm_averaged_vcd_corrected", "ak_OI", "increment_OI", "error_OI"]:
        store["oi." + n] = np.array(getattr(d, n))
    return grids


def mopitt_chain(impl, coarse=False):
    c = cases.mopitt_case(coarse)
    store = {}
    grids = []
    for i, g in enumerate(c["granules"]):
        r = impl.interpolator(1, c["grid_size"], cases.clone(g), c["coords"],
                              flag_thresh=c["flag_thresh"])
        assert r is not None and bool(r.ctm_upscaled_needed) == (not coarse)
        put(store, "interp%d" % i, r, ["vcd", "uncertainty", "x_col", "aprior_column",
                                        "surface_pressure", "apriori_surface", "pressure_mid",
                                        "averaging_kernels", "apriori_profile"])
        grids.append(r)
    grids = month_tail(impl, store, "MOPITT", "CO", c["ctm"], grids, ["ctm_vcd", "ctm_xcol"])
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
    grids = month_tail(impl, store, "GOSAT", "CH4", c["ctm"], grids, ["ctm_xcol"])
    return store, grids


def o3_chain(impl):
    """OMI total O3: no scattering weights, so amf_recal only grabs the model column
    (amf_recal.py:160-171) and average() converts it to Dobson units (driver.py:62-63)."""
    c = cases.o3_case()
    store = {}
    grids = []
    for i, g in enumerate(c["granules"]):
        r = impl.interpolator(1, c["grid_size"], cases.clone(g), c["coords"],
                              flag_thresh=c["flag_thresh"])
        assert r is not None
        put(store, "interp%d" % i, r, ["vcd", "amf", "uncertainty"])
        assert np.size(r.scattering_weights) == 1 and np.size(r.tropopause) == 1
        grids.append(r)
    grids = month_tail(impl, store, "OMI", "O3", c["ctm"], grids, ["vcd", "ctm_vcd"],
                       skip_aux=True)
    return store, grids


def ssmis_chain(impl, fine=False):
    """SSMIS water vapour (run/job.py:67-68): reader front-end -> interpolator_ssmis ->
    pwv_calculator -> average -> oi, the month through the oisatgmi class."""
    c = cases.ssmis_case(fine)
    store = {}
    r = impl.ssmis_wv({k: np.array(v) for k, v in c["vars"].items()}, c["yyyymm"])
    put(store, "read", r, ["vcd", "uncertainty", "latitude_center", "longitude_center"])
    g = impl.interpolator_ssmis(1, c["grid_size"], r, c["coords"])
    assert g is not None and bool(g.ctm_upscaled_needed) == bool(fine)
    put(store, "interp", g, ["vcd", "uncertainty", "latitude_center", "longitude_center"])
    month_tail(impl, store, "SSMIS", "H2O", c["ctm"], [g], ["ctm_vcd"], skip_aux=True)
    return store, [g]


def _attached(d, ctm, grids):
    d.reader_obj = cases.reader_ns(grids, ctm)
    return d


def oracle_impl():
    import types
    from oracle import (averaging as oavg, driver as odriver, interp as ointerp, oi as ooi,
                        ssmis as ossmis, vertical as overt)
    return types.SimpleNamespace(
        interpolator=ointerp.interpolator, filler_gosatxch4=ointerp.filler_gosatxch4,
        amf_recal=overt.amf_recal, ak_conv_mopitt=overt.ak_conv_mopitt,
        ak_conv_gosat=overt.ak_conv_gosat, averaging=oavg.averaging, OI=ooi.OI, bias=ooi.BIAS,
        ssmis_wv=ossmis.ssmis_wv, interpolator_ssmis=ossmis.interpolator_ssmis,
        driver=lambda ctm, grids: _attached(odriver.oisatgmi(), ctm, grids))


def reference_impl():
    """The LIVE unmodified reference (build container only, through oracle/ref_shim.py):
    what oracle/make_golden.py records and tests/test_oracle_vs_reference.py compares with."""
    import contextlib
    import io
    import types
    from oisatgmi_b200 import config
    from oracle import ref_shim
    ref = ref_shim.load_reference()

    def quiet(fn):
        def run(*a, **k):
            with contextlib.redirect_stdout(io.StringIO()):
                return fn(*a, **k)
        return run

    def conv(g):
        cls = ref.config.satellite_amf if config.kind_of(g) == "amf" else ref.config.satellite_opt
        return g if isinstance(g, cls) else config.convert(g, cls)

    def ctm(models):
        return [m if isinstance(m, ref.config.ctm_model) else config.convert(m, ref.config.ctm_model)
                for m in models]

    class QuietDriver(ref.driver.oisatgmi):
        pass

    for meth in ("recal_amf", "conv_ak", "cal_pwv", "average", "bias_correct", "oi"):
        setattr(QuietDriver, meth, quiet(getattr(ref.driver.oisatgmi, meth)))

    return types.SimpleNamespace(
        interpolator=quiet(lambda k, gs, g, c, flag_thresh=0.75: ref.interpolator(
            k, gs, conv(g), c, flag_thresh=flag_thresh)),
        filler_gosatxch4=quiet(lambda gs, g, flag_thresh=0.75: ref.filler_gosatxch4(
            gs, conv(g), flag_thresh=flag_thresh)),
        amf_recal=quiet(lambda m, s: ref.amf_recal(ctm(m), s)),
        ak_conv_mopitt=quiet(lambda m, s: ref.ak_conv_mopitt(ctm(m), s)),
        ak_conv_gosat=quiet(lambda m, s: ref.ak_conv_gosat(ctm(m), s)),
        averaging=quiet(ref.averaging), OI=quiet(ref.OI),
        bias={("TROPOMI", "NO2"): (0.32, 0.66), ("TROPOMI", "HCHO"): (0.90, 0.59),
              ("OMI", "NO2"): (0.32, 0.63), ("OMI", "HCHO"): (0.821, 0.79)},
        ssmis_wv=_reference_ssmis_reader(ref, quiet),
        interpolator_ssmis=quiet(lambda k, gs, g, c: ref.interpolator_ssmis(
            k, gs, g if isinstance(g, ref.config.satellite_ssmis)
            else config.convert(g, ref.config.satellite_ssmis), c)),
        driver=lambda models, grids: _attached(QuietDriver(), ctm(models), grids))


def _reference_ssmis_reader(ref, quiet):
    """ssmis_reader_wv (reader.py:1277-1305), unmodified, with its file access (`_read_nc`,
    `_read_ssmi`) answered from the dictionary and the month in a file name it can parse."""
    import sys

    def run(v, yyyymm):
        rd = sys.modules["oisatgmi.reader"]
        saved = rd._read_nc, rd._read_ssmi
        rd._read_nc = lambda fname, var: np.squeeze(np.array(v[var]))
        rd._read_ssmi = lambda fname, var: np.squeeze(np.array(v[var]))
        try:
            return quiet(rd.ssmis_reader_wv)("dir/f17_%sv7.nc" % yyyymm, None)
        finally:
            rd._read_nc, rd._read_ssmi = saved
    return run


def cuda_impl():
    import types
    from oisatgmi_b200 import (ak_conv_gosat, ak_conv_mopitt, amf_recal, averaging, driver,
                               filler_gosat, interpolator, interpolator_ssmis,
                               optimal_interpolation, reader_frontend)
    return types.SimpleNamespace(
        interpolator=interpolator.interpolator, filler_gosatxch4=filler_gosat.filler_gosatxch4,
        amf_recal=amf_recal.amf_recal, ak_conv_mopitt=ak_conv_mopitt.ak_conv_mopitt,
        ak_conv_gosat=ak_conv_gosat.ak_conv_gosat, averaging=averaging.averaging,
        OI=optimal_interpolation.OI, bias=driver.BIAS_CORRECTION,
        ssmis_wv=reader_frontend.ssmis_wv, interpolator_ssmis=interpolator_ssmis.interpolator_ssmis,
        driver=lambda ctm, grids: _attached(driver.oisatgmi(), ctm, grids))


READER_CALLS = [("omi_no2", (True,)), ("omi_no2", (False,)), ("omi_hcho", ()),
                ("tropomi_no2", (True,)), ("tropomi_no2", (False,)),
                ("mopitt_co", ()), ("gosat_xch4", ())]
READER_FIELDS = ("vcd", "amf", "tropopause", "latitude_center", "longitude_center", "uncertainty",
                 "quality_flag", "pressure_mid", "scattering_weights")
READER_FIELDS_OPT = ("vcd", "latitude_center", "longitude_center", "uncertainty", "quality_flag",
                     "pressure_mid", "averaging_kernels", "aprior_column", "apriori_profile",
                     "surface_pressure", "apriori_surface", "x_col", "pressure_weight")


def reader_record(store, tag, r):
    store[tag + ".time"] = np.array(r.time.isoformat())
    names = READER_FIELDS if hasattr(r, "scattering_weights") else READER_FIELDS_OPT
    for n in names:
        a = np.asarray(getattr(r, n))
        if a.size > 1:
            store["%s.%s" % (tag, n)] = a


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
        reader_record(store, "trop%d" % int(args[0]) if args else "all", r)
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
