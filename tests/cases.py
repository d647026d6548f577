"""Seeded parity cases shared by oracle/make_golden.py (which runs the LIVE
reference on them in the build container) and by the test-suite (which runs the
oracle restatement and the CUDA path on the same inputs).  Small on purpose: the
oracle finishes each case in seconds."""
import copy
import datetime
import types

import numpy as np

from oisatgmi_b200 import synth

REGION = (30.0, 50.0, -105.0, -75.0)        # 41 x 49 model cells
REGION_AK = (10.0, 60.0, -130.0, -60.0)     # MOPITT / GOSAT: 1 degree lattice needs room
LEVEL_SUBSET = (0, 5, -1)                    # 3-D fields are stored for these levels only


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
}
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
    return dict(product=product, grid_size=gs, flag_thresh=thr, kind=kind,
                granules=amf_granules(product, seeds, nt, nxt), coords=coords(), ctm=ctm(),
                sensor=product.split("_")[0], gas=product.split("_")[1])


def mopitt_case():
    c = coords(REGION_AK)
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


def reader_ns(sat_data, ctm_data=None):
    return types.SimpleNamespace(sat_data=sat_data, ctm_data=ctm_data)


def clone(x):
    return copy.deepcopy(x)


def subset_levels(a):
    a = np.asarray(a)
    if a.ndim == 3:
        return a[list(LEVEL_SUBSET)]
    return a
