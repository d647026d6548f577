"""Reader front-end on the GPU: the conversions the
reference's level-2 readers apply to the file variables before they call
`interpolator` -- reader.py:807-903 `omi_reader_no2`, :906-983 `omi_reader_hcho`,
:707-804 `tropomi_reader_no2`.

File access is not part of this module (netCDF4/h5py are the reference's I/O
layer, out of scope).  Each function takes `v`, a mapping from the variable names
the reference reads to arrays with the file's dtypes and shapes -- exactly what
its `_read_group_nc` (reader.py:51-67) returns -- and produces the `satellite_amf`
record the reference's function builds (same field dtypes, bit for bit), through
liboisat's K7 kernels.  With `device=True` the big arrays stay on the GPU (torch
tensors), ready for `pipeline.MonthPipeline`.

There is no CPU fallback: without liboisat.so / a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C
import datetime

import numpy as np

from . import _dev, _lib
from .config import satellite_amf

# hybrid coefficients of the OMI HCHO product's 47 layers (reader.py:954-957)
OMI_HCHO_A = (0., 0.04804826, 6.593752, 13.1348, 19.61311, 26.09201, 32.57081, 38.98201, 45.33901,
              51.69611, 58.05321, 64.36264, 70.62198, 78.83422, 89.09992, 99.36521, 109.1817,
              118.9586, 128.6959, 142.91, 156.26, 169.609, 181.619, 193.097, 203.259, 212.15,
              218.776, 223.898, 224.363, 216.865, 201.192, 176.93, 150.393, 127.837, 108.663,
              92.36572, 78.51231, 56.38791, 40.17541, 28.36781, 19.7916, 9.292942, 4.076571,
              1.65079, 0.6167791, 0.211349, 0.06600001, 0.01)
OMI_HCHO_B = (1., 0.984952, 0.963406, 0.941865, 0.920387, 0.898908, 0.877429, 0.856018, 0.8346609,
              0.8133039, 0.7919469, 0.7706375, 0.7493782, 0.721166, 0.6858999, 0.6506349, 0.6158184,
              0.5810415, 0.5463042, 0.4945902, 0.4437402, 0.3928911, 0.3433811, 0.2944031,
              0.2467411, 0.2003501, 0.1562241, 0.1136021, 0.06372006, 0.02801004, 0.006960025,
              8.175413e-09, 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.)


def _up(a):
    """Host array -> device tensor in a dtype the kernels take (floats as they are,
    every integer type as int32: the value, not the width, is what the readers use)."""
    a = np.asarray(a)
    if a.dtype.kind in "iub":
        a = a.astype(np.int32)
    elif a.dtype not in (np.float16, np.float32, np.float64):
        a = a.astype(np.float64)
    return _dev.to_device(np.ascontiguousarray(a))


def _scale(src, factors, shape):
    L = _lib.lib()
    d = _up(src)
    out = _dev.empty((d.numel(),), "float16")
    f = (C.c_double * max(len(factors), 1))(*factors)
    _lib.check(L.oisat_reader_scale_f16(d.data_ptr(), _dev.dtype_code(d), d.numel(), f, len(factors),
                                        out.data_ptr(), _dev.stream()))
    return out.reshape(tuple(shape))


def _quality(mode, flags, cloud, terrain, shape):
    L = _lib.lib()
    f, c = _up(flags), _up(cloud)
    t = _up(terrain) if terrain is not None else None
    out = _dev.empty((f.numel(),))
    _lib.check(L.oisat_reader_quality(mode, f.data_ptr(), _dev.dtype_code(f), c.data_ptr(),
                                      _dev.dtype_code(c), _dev.ptr(t),
                                      _dev.dtype_code(t) if t is not None else 0, f.numel(),
                                      out.data_ptr(), _dev.stream()))
    return out.reshape(tuple(shape))


def _weights(src, pixel_major, n_lev, shape, scale=None):
    L = _lib.lib()
    d = _up(src)
    n_px = int(np.prod(shape))
    s = None
    if scale is not None:
        s = np.asarray(scale)
        s = _up(s if s.dtype in (np.float32, np.float64) else s.astype(np.float64))
    out = _dev.empty((n_lev, n_px), "float16")
    _lib.check(L.oisat_reader_weights(d.data_ptr(), _dev.dtype_code(d), int(pixel_major), n_lev, n_px,
                                      _dev.ptr(s), _dev.dtype_code(s) if s is not None else 0,
                                      out.data_ptr(), _dev.stream()))
    return out.reshape((n_lev,) + tuple(shape))


def _pmid(mode, a, b, ps, ps_div, n_lev, shape):
    L = _lib.lib()
    n_px = int(np.prod(shape))
    a_d = _dev.to_device(np.ascontiguousarray(a, dtype=np.float64))
    b_d = _dev.to_device(np.ascontiguousarray(b, dtype=np.float64)) if b is not None else None
    p = _up(ps) if ps is not None else None
    out = _dev.empty((n_lev, n_px), "float16")
    _lib.check(L.oisat_reader_pmid(mode, a_d.data_ptr(), _dev.ptr(b_d), _dev.ptr(p),
                                   _dev.dtype_code(p) if p is not None else 0, float(ps_div), n_lev,
                                   n_px, out.data_ptr(), _dev.stream()))
    return out.reshape((n_lev,) + tuple(shape))


def _finish(fields, device):
    return fields if device else [(_dev.to_host(f) if hasattr(f, "data_ptr") else f) for f in fields]


def _epoch(seconds, year):
    return datetime.datetime(year, 1, 1) + datetime.timedelta(seconds=int(seconds))


def _record(vcd, amf, time, trop, lat, lon, unc, qf, p_mid, sw):
    return satellite_amf(vcd, amf, time, trop, lat, lon, [], [], unc, qf, p_mid, sw, [], [], [], [],
                         [])


def omi_no2(v, trop, read_ak=True, device=False):
    """reader.py:820-896."""
    _dev.require_cuda()
    time = _epoch(np.squeeze(np.nanmean(v["Time"])), 1993)
    lat = np.asarray(v["Latitude"]).astype("float32")
    lon = np.asarray(v["Longitude"]).astype("float32")
    shape = lat.shape
    sfx = "Trop" if trop else ""
    vcd = _scale(v["ColumnAmountNO2" + sfx], (1e-15,), shape)
    unc = _scale(v["ColumnAmountNO2" + sfx + "Std"], (1e-15,), shape)
    qf = _quality(0, v["VcdQualityFlags"], v["CloudFraction"], v["TerrainReflectivity"], shape)
    levels = np.asarray(v["ScatteringWeightPressure"]).astype("float16").astype(np.float64)
    p_mid = _pmid(0, levels, None, None, 0.0, 35, shape)
    sw = _weights(v["ScatteringWeight"], True, 35, shape) if read_ak else np.empty((1))
    tropopause = _scale(v["TropopausePressure"], (), shape) if trop else np.empty((1))
    vcd, unc, qf, p_mid, sw, tropopause = _finish([vcd, unc, qf, p_mid, sw, tropopause], device)
    return _record(vcd, v["Amf" + sfx], time, tropopause, lat, lon, unc, qf, p_mid, sw)


def omi_hcho(v, read_ak=True, device=False):
    """reader.py:920-974."""
    _dev.require_cuda()
    time = _epoch(np.squeeze(np.nanmean(v["time"])), 1993)
    lat = np.asarray(v["latitude"]).astype("float32")
    lon = np.asarray(v["longitude"]).astype("float32")
    shape = lat.shape
    vcd = _scale(v["column_amount"], (1e-15,), shape)
    unc = _scale(v["column_uncertainty"], (1e-15,), shape)
    qf = _quality(1, v["main_data_quality_flag"], v["cloud_fraction"], None, shape)
    n_lev = len(OMI_HCHO_A) - 1
    p_mid = _pmid(1, OMI_HCHO_A, OMI_HCHO_B, v["surface_pressure"], 0.0, n_lev, shape)
    sw = _weights(v["scattering_weights"], False, n_lev, shape) if read_ak else np.empty((1))
    vcd, unc, qf, p_mid, sw = _finish([vcd, unc, qf, p_mid, sw], device)
    return _record(vcd, v["amf"], time, np.empty((1)), lat, lon, unc, qf, p_mid, sw)


def tropomi_no2(v, trop, read_ak=True, device=False):
    """reader.py:721-797."""
    _dev.require_cuda()
    L = _lib.lib()
    t = v["time"] + np.nanmean(np.array(v["delta_time"]), axis=0) / 1000.0
    time = _epoch(np.squeeze(t), 2010)
    lat = np.asarray(v["latitude"]).astype("float32")
    lon = np.asarray(v["longitude"]).astype("float32")
    shape = lat.shape
    amf_total = v["air_mass_factor_total"]
    if not trop:
        col, amf = "nitrogendioxide_total_column", amf_total
    else:
        col, amf = "nitrogendioxide_tropospheric_column", v["air_mass_factor_troposphere"]
    factors = (6.02214, 1e19, 1e-15)      # mol m-2 -> 1e15 molecules cm-2 (:752-754)
    vcd = _scale(v[col], factors, shape)
    unc = _scale(v[col + "_precision"], factors, shape)
    qf = _scale(v["qa_value"], (), shape)
    # 35 interface coefficients: tiny host arrays, numpy's own promotion (:757-760)
    tm5_a = np.concatenate(((np.asarray(v["tm5_constant_a"]) / 100.0)[:, 0], 0), axis=None)
    tm5_b = np.concatenate((np.asarray(v["tm5_constant_b"])[:, 0], 0), axis=None)
    # float32 coefficients keep the whole expression in float32 (numpy float32 scalars)
    mode = 3 if (tm5_a.dtype == np.float32 and tm5_b.dtype == np.float32) else 2
    p_mid = _pmid(mode, tm5_a, tm5_b, np.asarray(v["surface_pressure"]).astype("float32"), 100.0,
                  34, shape)
    sw = _weights(v["averaging_kernel"], True, 34, shape, scale=amf_total) if read_ak \
        else np.empty((1))
    if trop:
        layer = _up(v["tm5_tropopause_layer_index"])
        tropopause = _dev.empty((layer.numel(),), "float16")
        _lib.check(L.oisat_reader_tropopause(layer.data_ptr(), p_mid.data_ptr(), 34, layer.numel(),
                                             tropopause.data_ptr(), _dev.stream()))
        tropopause = tropopause.reshape(tuple(shape))
    else:
        tropopause = np.empty((1))
    vcd, unc, qf, p_mid, sw, tropopause = _finish([vcd, unc, qf, p_mid, sw, tropopause], device)
    return _record(vcd, amf, time, tropopause, lat, lon, unc, qf, p_mid, sw)


# ------------------------------------------------------------------ satellite_opt products
def _clean(src, shape_px, n_lev=1, pixel_major=False, pre=0, factors=(), out="float32", post=0):
    """oisat_reader_clean on one file variable: returns a device tensor shaped
    (n_lev,) + shape_px (or shape_px when n_lev == 1)."""
    L = _lib.lib()
    a = np.asarray(src)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    d = _dev.to_device(np.ascontiguousarray(a))
    n_px = int(np.prod(shape_px))
    o = _dev.empty((n_lev * n_px,), out)
    f = (C.c_double * max(len(factors), 1))(*factors)
    code = {"float16": _lib.F16, "float32": _lib.F32, "float64": _lib.F64}[out]
    _lib.check(L.oisat_reader_clean(d.data_ptr(), _dev.dtype_code(d), n_px, n_lev, int(pixel_major),
                                    pre, f, len(factors), code, post, o.data_ptr(), _dev.stream()))
    return o.reshape(((n_lev,) if n_lev > 1 else ()) + tuple(shape_px))


def _native(a):
    a = np.asarray(a)
    return "float64" if a.dtype not in (np.float32,) else "float32"


def mopitt_co(v, read_ak=True, device=False):
    """reader.py:1143-1203 (`mopitt_reader_co` up to its interpolator call).  `v`: the MOP03
    'Data Fields' variables plus StartTime / StopTime of the file attributes."""
    from .config import satellite_opt
    _dev.require_cuda()
    L = _lib.lib()
    time = _epoch(0.5 * (v["StartTime"] + v["StopTime"]), 1993)
    lat = np.asarray(v["Latitude"]).astype("float32")
    lon = np.asarray(v["Longitude"]).astype("float32")
    lon, lat = np.meshgrid(lon, lat)
    lon, lat = np.transpose(lon), np.transpose(lat)
    shape = lat.shape
    vcd = _clean(v["RetrievedCOTotalColumnDay"], shape, pre=3, factors=(1e-15,), out="float16")
    dry = _up(np.asarray(v["DryAirColumnDay"]).astype(np.float32))
    x_col = _dev.empty((vcd.numel(),), "float32")
    _lib.check(L.oisat_reader_mopitt_xcol(vcd.data_ptr(), dry.data_ptr(), vcd.numel(),
                                          x_col.data_ptr(), _dev.stream()))
    x_col = x_col.reshape(tuple(shape))
    ap_prof = _clean(v["APrioriCOMixingRatioProfileDay"], shape, 9, True, pre=1,
                     out=_native(v["APrioriCOMixingRatioProfileDay"]))
    ap_sfc = _clean(v["APrioriCOSurfaceMixingRatioDay"], shape, pre=1,
                    out=_native(v["APrioriCOSurfaceMixingRatioDay"]))
    ap_col = _clean(v["APrioriCOTotalColumnDay"], shape, factors=(1e-15,), out="float16", post=1)
    unc = _clean(v["RetrievedCOTotalColumnMeanUncertaintyDay"], shape, factors=(1e-15,), out="float32")
    levels = np.asarray(v["Pressure"]).astype("float16").astype(np.float64)
    p_mid = _pmid(0, levels, None, None, 0.0, 9, shape)
    if read_ak:
        aks = _clean(v["TotalColumnAveragingKernelDay"], shape, 10, True, factors=(1e-15,), out="float16")
    else:
        aks = np.empty((1))
    vcd, x_col, ap_prof, ap_sfc, ap_col, unc, p_mid, aks = _finish(
        [vcd, x_col, ap_prof, ap_sfc, ap_col, unc, p_mid, aks], device)
    quality = np.ones(shape, dtype=np.float16)
    return satellite_opt(vcd, time, [], np.empty((1)), lat, lon, [], [], unc, quality, p_mid, aks, [],
                         [], [], [], ap_col, ap_prof, v["SurfacePressureDay"], ap_sfc, x_col, [],
                         "MOPITT")


def gosat_xch4(v, read_ak=True, device=False):
    """reader.py:1228-1264 (`gosat_reader_xch4` up to the gap filler)."""
    from .config import satellite_opt
    _dev.require_cuda()
    time = datetime.datetime(1970, 1, 1) + datetime.timedelta(
        seconds=int(np.squeeze(np.nanmean(v["time"]))))
    lat = np.asarray(v["latitude"]).astype("float32")
    lon = np.asarray(v["longitude"]).astype("float32")
    n = (lat.size,)
    nlev = int(np.shape(v["pressure_levels"])[1])
    xch4 = _clean(v["xch4"], n, pre=3, out=_native(v["xch4"]))
    ap = _clean(v["ch4_profile_apriori"], n, nlev, True, pre=1, out=_native(v["ch4_profile_apriori"]))
    p_mid = _clean(v["pressure_levels"], n, nlev, True, pre=1, out=_native(v["pressure_levels"]))
    if read_ak:
        aks = _clean(v["xch4_averaging_kernel"], n, nlev, True, pre=1,
                     out=_native(v["xch4_averaging_kernel"]))
        pw = _clean(v["pressure_weight"], n, nlev, True, pre=1, out=_native(v["pressure_weight"]))
    else:
        aks, pw = np.empty((1)), np.empty((1))
    xch4, ap, p_mid, aks, pw = _finish([xch4, ap, p_mid, aks, pw], device)
    return satellite_opt(xch4, time, [], np.empty((1)), lat, lon, [], [], v["xch4_uncertainty"],
                         1 - v["xch4_quality_flag"], p_mid, aks, [], [], [], [], np.empty((1)), ap,
                         np.empty((1)), np.empty((1)), xch4, pw, "GOSAT")


def ssmis_wv(v, yyyymm, device=False):
    """reader.py:1280-1297 (`ssmis_reader_wv` up to its interpolator call): `v` holds the
    'latitude' / 'longitude' axes and the scaled 'atmosphere_water_vapor_content' map; `yyyymm`
    is the month the reader parses from the file name."""
    from .config import satellite_ssmis
    _dev.require_cuda()
    time = datetime.datetime(int(yyyymm[0:4]), int(yyyymm[4:6]), 1)
    lat = np.asarray(v["latitude"]).astype("float32")
    lon = np.asarray(v["longitude"]).astype("float32")
    lon[lon > 180.0] = lon[lon > 180.0] - 360.0
    lon, lat = np.meshgrid(lon, lat)
    raw = np.asarray(v["atmosphere_water_vapor_content"])
    src = _dev.to_device(np.ascontiguousarray(raw)) if raw.dtype == np.uint8 else _up(raw)
    pwv = _dev.empty((src.numel(),), "float32")
    unc = _dev.empty((src.numel(),), "float32")
    _lib.check(_lib.lib().oisat_reader_ssmis(src.data_ptr(), _dev.dtype_code(src), src.numel(),
                                             pwv.data_ptr(), unc.data_ptr(), _dev.stream()))
    pwv, unc = pwv.reshape(lat.shape), unc.reshape(lat.shape)
    pwv, unc = _finish([pwv, unc], device)
    return satellite_ssmis(pwv, unc, time, lat, lon, False, [], "SSMI")
