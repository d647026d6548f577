"""Seeded parity cases shared by oracle/make_golden.py (which runs the LIVE
reference on them in the build container) and by the test-suite (which runs the
oracle restatement and the CUDA path on the same inputs).  Small on purpose: the
oracle finishes each case in seconds."""
import copy
import datetime
import types

import numpy as np

import synth

REGION = (30.0, 50.0, -105.0, -75.0)        # 41 x 49 model cells
REGION_AK = (10.0, 60.0, -130.0, -60.0)     # MOPITT / GOSAT: 1 degree lattice needs room
LEVEL_SUBSET = (0, 5, -1)                    # 3-D fields are stored for these levels only ...
ALL_LEVELS = ("omi_hcho",)                   # ... except in these cases: every level is compared


def coords(region=REGION):
    return synth.ctm_coordinates(region)


def ctm(region=REGION, ctmtype="GMI", averaged=True, seed=11, gas_scale=1.0):
    return [synth.make_ctm(seed, coords(region), ctmtype=ctmtype, averaged=averaged,
                           gas_scale=gas_scale)]


def amf_granules(product, seeds, nt, nxt, region=REGION, bad_fraction=0.25):
    geo0 = synth.regional_geo(region)
    out = []
    for i, s in enumerate(seeds):
        geo = dict(geo0)
        geo["node_lon_deg"] = geo0["node_lon_deg"] + 4.0 * (i - (len(seeds) - 1) / 2.0)
        t = datetime.datetime(2005, 6, 3 + 2 * i, 13 + i % 3, 10 * i % 60, 7)
        out.append(synth.make_amf_granule(s, product, nt=nt, nxt=nxt, geo=geo,
                                          bad_fraction=bad_fraction, time=t))
    return out


CASES = {
    # name: (product, grid_size, flag_thresh, seeds, nt, nxt[, interpolator type])
    "omi_no2": ("OMI_NO2", 0.25, 0.0, (3, 4, 5), 260, 60),
    "omi_hcho": ("OMI_HCHO", 0.25, 0.0, (7, 8), 260, 60),
    "tropomi_no2": ("TROPOMI_NO2", 0.10, 0.75, (9,), 620, 140),
    # the nearest-neighbour gridding modes (interpolator.py:17-20 type 2: TROPOMI HCHO,
    # reader.py:698-700; :28-33 type 4: TEMPO, reader.py:528-530) on the same kind of data
    "tropomi_nearest": ("TROPOMI_NO2", 0.10, 0.5, (12,), 420, 120, 2),
    "omi_kdtree": ("OMI_HCHO", 0.25, 0.0, (13, 14), 200, 60, 4),
    # a model FINER than the working mesh (HiGMI / CMAQ / FREE, amf_recal.py:58-83): the
    # gridded granule stays on the mesh (ctm_upscaled_needed) and amf_recal brings the model's
    # pressure and partial column to it with _upscaler, level by level
    "omi_no2_fine_model": ("OMI_NO2", 0.25, 0.0, (15, 16), 200, 60, 1),
    # satellite pressure profiles that are NOT strictly monotone, which scipy's interp1d
    # (assume_sorted=False, amf_recal.py:103-107) argsorts: exact ties and level swaps over
    # whole blocks of scan lines (so the GRIDDED profile has them too), NaN pressures and NaN
    # scattering weights at scattered pixels -- the slow path of the vertical kernels
    "omi_no2_kinked": ("OMI_NO2", 0.25, 0.0, (17, 18), 260, 60, 1),
}
KINKED = ("omi_no2_kinked",)


def kink_profiles(granules):
    """Deterministic damage to the pressure / scattering-weight profiles (in place)."""
    for gi, g in enumerate(granules):
        rng = np.random.default_rng(100 + gi)
        p = g.pressure_mid
        sw = g.scattering_weights
        L, nt, nxt = p.shape
        k0, k1 = 5 + gi, 12 + 2 * gi
        p[k0 + 1, 40:70, :] = p[k0, 40:70, :]                       # ties
        a = p[k1, 110:140, :].copy()                                  # a swap two levels apart
        p[k1, 110:140, :] = p[k1 + 2, 110:140, :]
        p[k1 + 2, 110:140, :] = a
        p[L - 1, 180:200, :] = p[L - 2, 180:200, :] * np.float16(1.5)   # top level out of order
        for arr, frac in ((p, 0.004), (sw, 0.004)):                   # scattered NaNs
            idx = rng.choice(nt * nxt, size=max(1, int(frac * nt * nxt)), replace=False)
            lev = rng.integers(0, L, size=idx.size)
            arr[lev, idx // nxt, idx % nxt] = np.nan
    return granules


REGION_FINE = (36.0, 46.0, -98.0, -84.0)
FINE_MODEL_SPACING = {"omi_no2_fine_model": 0.125}   # degrees, both directions


def amf_case(name):
    product, gs, thr, seeds, nt, nxt = CASES[name][:6]
    kind = CASES[name][6] if len(CASES[name]) > 6 else 1
    if name in FINE_MODEL_SPACING:
        d = FINE_MODEL_SPACING[name]
        c = synth.ctm_coordinates(REGION_FINE, dlat=d, dlon=d)
        model = [synth.make_ctm(17, c, averaged=True)]
        return dict(product=product, grid_size=gs, flag_thresh=thr, kind=kind,
                    granules=amf_granules(product, seeds, nt, nxt, region=REGION_FINE), coords=c,
                    ctm=model, sensor=product.split("_")[0], gas=product.split("_")[1])
    granules = amf_granules(product, seeds, nt, nxt)
    if name in KINKED:
        kink_profiles(granules)
    return dict(product=product, grid_size=gs, flag_thresh=thr, kind=kind,
                granules=granules, coords=coords(), ctm=ctm(),
                sensor=product.split("_")[0], gas=product.split("_")[1])


def mopitt_case(coarse=False):
    """`coarse`: a 2 x 2.5 degree model, COARSER than the 1 degree L3 lattice -- the gridded
    granule then lands on the model grid itself (interpolator.py:64-93) and ak_conv_mopitt uses
    the model fields as they are (no resampling, ak_conv_mopitt.py:79)."""
    c = synth.ctm_coordinates(REGION_AK, dlat=2.0, dlon=2.5) if coarse else coords(REGION_AK)
    model = [synth.make_ctm(21, c, ctmtype="ECCOH", averaged=False, gas_scale=40.0,
                            date=datetime.datetime(2005, 6, 1))]
    grans = [synth.make_mopitt_granule(31 + i, region=REGION_AK,
                                       time=datetime.datetime(2005, 6, 4 + i, 12)) for i in range(2)]
    return dict(granules=grans, coords=c, ctm=model, grid_size=1.0, flag_thresh=0.0)


def gosat_case():
    c = coords(REGION_AK)
    model = [synth.make_ctm(22, c, ctmtype="ECCOH", averaged=False, gas_scale=600.0,
                            date=datetime.datetime(2005, 6, 1))]
    grans = [synth.make_gosat_soundings(41 + i, n=1500, region=(15.0, 55.0, -125.0, -65.0),
                                        time=datetime.datetime(2005, 6, 6 + i, 3)) for i in range(2)]
    return dict(granules=grans, coords=c, ctm=model, grid_size=1.0, flag_thresh=0.0)


def o3_case():
    """OMI total O3 as omi_reader_o3 hands it over (reader.py:1037-1042): no scattering
    weights, no tropopause, `amf` = the column itself, `pressure_mid` an empty list."""
    granules = amf_granules("OMI_O3", (23, 24), 220, 60)
    for g in granules:
        g.amf = g.vcd
        g.pressure_mid = []
    return dict(granules=granules, coords=coords(), ctm=ctm(gas_scale=30.0), grid_size=0.25,
                flag_thresh=0.0, sensor="OMI", gas="O3")


def ssmis_vars(seed=51, region=REGION):
    """Variables of one monthly SSMIS map as ssmis_reader_wv reads them (reader.py:1285-1292):
    0.25 degree axes (longitude 0..360) and the scaled-byte water-vapour map with its flag
    values above 250 (land, ice, no data)."""
    rng = np.random.default_rng(seed)
    la0, la1, lo0, lo1 = region
    lat = np.arange(la0 - 2.0, la1 + 2.0, 0.25) + 0.125
    lon = (np.arange(lo0 - 2.0, lo1 + 2.0, 0.25) + 0.125) % 360.0
    shape = (lat.size, lon.size)
    wv = 60.0 + 90.0 * (synth._smooth2d(rng, shape, scale=3.0) + 0.5)
    wv = np.clip(wv, 0.0, 249.0)
    land = synth._coherent_mask(rng, shape, 0.3)
    wv[~land] = rng.choice([251.0, 252.0, 253.0, 254.0, 255.0], size=int((~land).sum()))
    wv[rng.uniform(size=shape) < 0.01] = 250.0          # 75.0 after scaling: rejected (>= 75)
    return {"latitude": lat.astype(np.float64), "longitude": lon.astype(np.float64),
            "atmosphere_water_vapor_content": wv.astype(np.uint8)}


def ssmis_case(fine=False):
    """SSMIS precipitable water against a water-vapour model field: the model coarser than the
    0.25 degree working mesh (the reference's production setting) or finer (`fine`: the
    interpolator leaves the map on the mesh and pwv_calculator resamples the model)."""
    if fine:
        c = synth.ctm_coordinates(REGION_FINE, dlat=0.125, dlon=0.125)
        region = REGION_FINE
    else:
        c, region = coords(), REGION
    model = [synth.make_ctm(61, c, ctmtype="ECCOH", averaged=False, gas_scale=2.0e3,
                            date=datetime.datetime(2005, 6, 1))]
    return dict(vars=ssmis_vars(region=region), yyyymm="200506", coords=c, ctm=model, grid_size=0.25)


def reader_ns(sat_data, ctm_data=None):
    return types.SimpleNamespace(sat_data=sat_data, ctm_data=ctm_data)


def clone(x):
    return copy.deepcopy(x)


def subset_levels(a):
    a = np.asarray(a)
    if a.ndim == 3:
        return a[list(LEVEL_SUBSET)]
    return a


# ---------------------------------------------------------------- reader front-end
READER_PRODUCTS = ("omi_no2", "omi_hcho", "tropomi_no2", "mopitt_co", "gosat_xch4")


def reader_vars(product, seed=3, nt=90, nxt=24):
    """File variables of one level-2 granule as `_read_group_nc` (reader.py:51-67) hands
    them to the readers: names, shapes and dtypes of the real products, synthetic values
    with the awkward ones included (fill values, negative / huge / NaN scattering weights,
    cloud fractions exactly on the float16 image of the threshold, every flag pattern)."""
    rng = np.random.default_rng(seed)
    geo = synth.regional_geo(REGION)
    lat, lon = synth.swath_geolocation(nt, nxt, rng=rng, **geo)
    shape = (nt, nxt)

    def f32(a):
        return np.asarray(a, dtype=np.float32)

    def weights(L, pixel_major):
        sw = rng.uniform(0.05, 3.0, shape + (L,)) if pixel_major else rng.uniform(0.05, 3.0, (L,) + shape)
        bad = rng.uniform(size=sw.shape)
        sw[bad < 0.01] = np.nan
        sw[(bad >= 0.01) & (bad < 0.02)] = -0.3
        sw[(bad >= 0.02) & (bad < 0.03)] = 250.0
        sw[(bad >= 0.03) & (bad < 0.035)] = np.inf
        sw[(bad >= 0.035) & (bad < 0.04)] = 1e6       # overflows float16 -> inf
        return f32(sw)

    cloud = rng.uniform(0.0, 1.0, shape)
    cloud[0, :4] = [0.3, np.float16(0.3), 0.4, np.float16(0.4)]   # threshold images
    cloud[1, :2] = [0.2998046875, 0.39990234375]
    if product == "omi_no2":
        v = {"Time": np.full(nt, 3.91e8) + np.arange(nt) * 2.0,
             "Latitude": lat.astype(np.float32), "Longitude": lon.astype(np.float32),
             "CloudFraction": f32(cloud),
             "TerrainReflectivity": f32(rng.uniform(0.0, 0.4, shape)),
             "VcdQualityFlags": rng.integers(0, 16, shape).astype(np.int16),
             "ScatteringWeightPressure": f32(synth.OMI_NO2_PRESSURES),
             "ScatteringWeight": weights(35, True),
             "TropopausePressure": f32(rng.uniform(90.0, 320.0, shape))}
        for sfx in ("", "Trop"):
            col = np.exp(rng.normal(35.0, 1.0, shape))
            col[rng.uniform(size=shape) < 0.02] = -1.2676506e30        # the product's fill value
            v["ColumnAmountNO2" + sfx] = f32(col)
            v["ColumnAmountNO2" + sfx + "Std"] = f32(np.exp(rng.normal(33.5, 0.5, shape)))
            v["Amf" + sfx] = f32(rng.uniform(0.4, 2.5, shape))
        return v
    if product == "omi_hcho":
        return {"time": np.full(nt, 3.92e8) + np.arange(nt) * 2.0,
                "latitude": lat.astype(np.float32), "longitude": lon.astype(np.float32),
                "column_amount": np.exp(rng.normal(36.0, 1.0, shape)),           # float64 in the file
                "column_uncertainty": np.exp(rng.normal(35.0, 0.5, shape)),
                "amf": f32(rng.uniform(0.4, 2.5, shape)),
                "cloud_fraction": f32(cloud),
                "main_data_quality_flag": rng.integers(0, 3, shape).astype(np.int8),
                "surface_pressure": f32(rng.uniform(600.0, 1030.0, shape)),
                "scattering_weights": weights(47, False)}
    if product == "tropomi_no2":
        a, b = synth._hybrid_table(35)
        v = {"time": np.float64(1.7e8), "delta_time": (np.arange(nt) * 840).astype(np.int64)[:, None]
             * np.ones((1, 1), np.int64),
             "latitude": lat.astype(np.float32), "longitude": lon.astype(np.float32),
             "air_mass_factor_total": f32(rng.uniform(0.8, 3.0, shape)),
             "air_mass_factor_troposphere": f32(rng.uniform(0.4, 2.5, shape)),
             "qa_value": f32(rng.integers(0, 101, shape) / 100.0),
             "tm5_constant_a": f32(np.stack([a[:-1], a[1:]], axis=1) * 100.0),
             "tm5_constant_b": f32(np.stack([b[:-1], b[1:]], axis=1)),
             "surface_pressure": f32(rng.uniform(60000.0, 103000.0, shape)),
             "averaging_kernel": weights(34, True),
             "tm5_tropopause_layer_index": rng.integers(-1, 36, shape).astype(np.int32)}
        for name in ("nitrogendioxide_total_column", "nitrogendioxide_tropospheric_column"):
            v[name] = f32(np.exp(rng.normal(-9.5, 1.0, shape)))               # mol m-2
            v[name + "_precision"] = f32(np.exp(rng.normal(-11.0, 0.5, shape)))
        return v
    if product == "mopitt_co":
        # MOP03 daily L3 'Data Fields' (reader.py:1143-1203): (lon, lat[, level]) float32 grids
        # with the product's -9999 fill value, 9 fixed pressure levels
        nlon, nlat = 72, 36
        shp = (nlon, nlat)

        def fill(a, frac=0.3):
            a = f32(a)
            a[rng.uniform(size=a.shape) < frac] = -9999.0
            return a
        return {"StartTime": 3.93e8, "StopTime": 3.93e8 + 86399.0,
                "Latitude": f32(np.linspace(-87.5, 87.5, nlat)),
                "Longitude": f32(np.linspace(-177.5, 177.5, nlon)),
                "RetrievedCOTotalColumnDay": fill(np.exp(rng.normal(42.0, 0.3, shp))),
                "DryAirColumnDay": f32(2.1e25 * (1.0 + 0.02 * rng.normal(size=shp))),
                "APrioriCOMixingRatioProfileDay": fill(rng.uniform(40.0, 160.0, shp + (9,)), 0.1),
                "APrioriCOSurfaceMixingRatioDay": fill(rng.uniform(60.0, 200.0, shp), 0.1),
                "SurfacePressureDay": f32(rng.uniform(600.0, 1030.0, shp)),
                "APrioriCOTotalColumnDay": fill(np.exp(rng.normal(42.0, 0.2, shp)), 0.1),
                "RetrievedCOTotalColumnMeanUncertaintyDay": f32(np.exp(rng.normal(39.0, 0.4, shp))),
                "Pressure": f32([900., 800., 700., 600., 500., 400., 300., 200., 100.]),
                "TotalColumnAveragingKernelDay": fill(np.exp(rng.normal(37.0, 0.5, shp + (10,))), 0.1)}
    if product == "gosat_xch4":
        # GOSAT proxy XCH4 level 2 (reader.py:1228-1262): per-sounding vectors, (N, 20) profiles
        n, L = 700, 20
        frac = (np.arange(L) + 0.5) / L
        ps = rng.uniform(600.0, 1013.0, n)

        def holes(a, frac_bad=0.05, val=-999.0):
            a = np.array(a, dtype=np.float32)
            a[rng.uniform(size=a.shape) < frac_bad] = val
            return a
        return {"time": 1.118e9 + np.arange(n) * 4.0,
                "latitude": f32(rng.uniform(-60.0, 75.0, n)), "longitude": f32(rng.uniform(-170.0, 170.0, n)),
                "xch4": holes(rng.normal(1800.0, 20.0, n)),
                "ch4_profile_apriori": holes(1850.0 - 300.0 * frac[None] ** 2 + rng.normal(0, 5.0, (n, L))),
                "xch4_quality_flag": (rng.uniform(size=n) < 0.2).astype(np.int8),
                "xch4_uncertainty": f32(rng.uniform(5.0, 15.0, n)),
                "pressure_levels": holes(ps[:, None] * (1.0 - frac[None]) + 0.1, 0.02, 0.0),
                "xch4_averaging_kernel": holes(rng.uniform(0.3, 1.2, (n, L)), 0.03, -1.0),
                "pressure_weight": holes(np.full((n, L), 1.0 / L), 0.03, 0.0)}
    raise KeyError(product)
