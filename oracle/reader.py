"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reader front-end: what the
reference's level-2 readers do to the file variables before they call
`interpolator`: unit scaling and float16 quantisation,
quality flags, scattering-weight clean-up, mid-level pressures, tropopause.

Follows /root/reference/oisatgmi/reader.py: omi_reader_no2 (:807-903),
omi_reader_hcho (:906-983), tropomi_reader_no2 (:707-804), mopitt_reader_co (:1130-1213),
gosat_reader_xch4 (:1216-1275, up to the gap filler).  File access is not
part of it: `v` maps the variable names those functions read to arrays with the
file's dtypes and shapes (what `_read_group_nc`, :51-67, returns: np.squeeze of the
variable).  Every expression keeps numpy's promotion rules (NEP 50: Python
scalars adopt the array dtype, numpy float64 scalars do not), which decide where
float16 / float32 rounding happens.

Parity status: PINNED -- tests/test_oracle_vs_reference.py runs the unmodified
reader functions with `_read_group_nc` replaced by a dictionary lookup and
compares bit for bit; fixtures in tests/golden/reader_*.npz.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module; the product package never does.
"""
from __future__ import annotations

import datetime

import numpy as np

from oisatgmi_b200.config import satellite_amf

# hybrid coefficients of the OMI HCHO product's 47 layers (reader.py:954-957)
OMI_HCHO_A = np.array([
    0., 0.04804826, 6.593752, 13.1348, 19.61311, 26.09201, 32.57081, 38.98201, 45.33901, 51.69611,
    58.05321, 64.36264, 70.62198, 78.83422, 89.09992, 99.36521, 109.1817, 118.9586, 128.6959,
    142.91, 156.26, 169.609, 181.619, 193.097, 203.259, 212.15, 218.776, 223.898, 224.363, 216.865,
    201.192, 176.93, 150.393, 127.837, 108.663, 92.36572, 78.51231, 56.38791, 40.17541, 28.36781,
    19.7916, 9.292942, 4.076571, 1.65079, 0.6167791, 0.211349, 0.06600001, 0.01])
OMI_HCHO_B = np.array([
    1., 0.984952, 0.963406, 0.941865, 0.920387, 0.898908, 0.877429, 0.856018, 0.8346609, 0.8133039,
    0.7919469, 0.7706375, 0.7493782, 0.721166, 0.6858999, 0.6506349, 0.6158184, 0.5810415,
    0.5463042, 0.4945902, 0.4437402, 0.3928911, 0.3433811, 0.2944031, 0.2467411, 0.2003501,
    0.1562241, 0.1136021, 0.06372006, 0.02801004, 0.006960025, 8.175413e-09, 0., 0., 0., 0., 0.,
    0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.])


def _epoch(seconds, year):
    return datetime.datetime(year, 1, 1) + datetime.timedelta(seconds=int(seconds))


def _clean(sw):
    """reader.py:887-888 (and :777-778, :967-968): bad weights become 0, in place."""
    sw[np.where((np.isnan(sw)) | (np.isinf(sw)) | (sw > 100.0) | (sw < 0.0))] = 0.0
    return sw


def omi_no2(v, trop, read_ak=True):
    """reader.py:820-896."""
    time = _epoch(np.squeeze(np.nanmean(v["Time"])), 1993)
    lat = v["Latitude"].astype("float32")
    lon = v["Longitude"].astype("float32")
    sfx = "Trop" if trop else ""
    vcd = v["ColumnAmountNO2" + sfx]
    amf = v["Amf" + sfx]
    unc = v["ColumnAmountNO2" + sfx + "Std"]
    vcd = (vcd * 1e-15).astype("float16")
    unc = (unc * 1e-15).astype("float16")
    cloud_ok = np.multiply(v["CloudFraction"].astype("float16") < 0.3, 1.0).squeeze()
    terrain_ok = np.multiply(v["TerrainReflectivity"].astype("float16") < 0.2, 1.0).squeeze()
    raw = v["VcdQualityFlags"].astype("float16")
    # bit 0 clear, or bit 0 set with bit 1 clear -> usable (:861-869); the reference decides
    # from the last two characters of the binary text of int(flag)
    bits = np.abs(raw.astype(np.int64)) & 3
    qf = np.ones_like(raw) * -100.0
    qf[bits != 3] = 1.0
    qf = qf * cloud_ok * terrain_ok
    ps = v["ScatteringWeightPressure"].astype("float16")
    p_mid = np.zeros((35,) + vcd.shape).astype("float16")
    for z in range(35):
        p_mid[z, :, :] = ps[z]
    if read_ak:
        sw = _clean(v["ScatteringWeight"].astype("float16").transpose((2, 0, 1)))
    else:
        sw = np.empty((1))
    tropopause = v["TropopausePressure"].astype("float16") if trop else np.empty((1))
    return satellite_amf(vcd, amf, time, tropopause, lat, lon, [], [], unc, qf, p_mid, sw, [], [],
                         [], [], [])


def omi_hcho(v, read_ak=True):
    """reader.py:920-974."""
    time = _epoch(np.squeeze(np.nanmean(v["time"])), 1993)
    lat = v["latitude"].astype("float32")
    lon = v["longitude"].astype("float32")
    vcd = (v["column_amount"] * 1e-15).astype("float16")
    unc = (v["column_uncertainty"] * 1e-15).astype("float16")
    amf = v["amf"]
    cloud_ok = np.multiply(v["cloud_fraction"].astype("float16") < 0.4, 1.0).squeeze()
    qf = np.multiply(v["main_data_quality_flag"].astype("float16") == 0.0, 1.0).squeeze()
    qf = qf * cloud_ok
    ps = v["surface_pressure"].astype("float16")
    a0, b0 = OMI_HCHO_A, OMI_HCHO_B
    p_mid = np.zeros((a0.size - 1,) + vcd.shape).astype("float16")
    for z in range(a0.size - 1):
        p_mid[z, :, :] = 0.5 * ((a0[z] + b0[z] * ps) + (a0[z + 1] + b0[z + 1] * ps))
    sw = _clean(v["scattering_weights"].astype("float16")) if read_ak else np.empty((1))
    return satellite_amf(vcd, amf, time, np.empty((1)), lat, lon, [], [], unc, qf, p_mid, sw, [],
                         [], [], [], [])


def tropomi_no2(v, trop, read_ak=True):
    """reader.py:721-797."""
    t = v["time"] + np.nanmean(np.array(v["delta_time"]), axis=0) / 1000.0
    time = _epoch(np.squeeze(t), 2010)
    lat = v["latitude"].astype("float32")
    lon = v["longitude"].astype("float32")
    amf_total = v["air_mass_factor_total"]
    if not trop:
        vcd = v["nitrogendioxide_total_column"]
        amf = amf_total
        unc = v["nitrogendioxide_total_column_precision"]
    else:
        vcd = v["nitrogendioxide_tropospheric_column"]
        amf = v["air_mass_factor_troposphere"]
        unc = v["nitrogendioxide_tropospheric_column_precision"]
    vcd = (vcd * 6.02214 * 1e19 * 1e-15).astype("float16")
    unc = (unc * 6.02214 * 1e19 * 1e-15).astype("float16")
    qf = v["qa_value"].astype("float16")
    tm5_a = v["tm5_constant_a"] / 100.0
    tm5_a = np.concatenate((tm5_a[:, 0], 0), axis=None)
    tm5_b = np.concatenate((v["tm5_constant_b"][:, 0], 0), axis=None)
    ps = v["surface_pressure"].astype("float32") / 100.0
    p_mid = np.zeros((34,) + vcd.shape).astype("float16")
    if read_ak:
        sw = np.zeros((34,) + vcd.shape).astype("float16")
        aks = v["averaging_kernel"].astype("float16")
    else:
        sw = np.empty((1))
    for z in range(34):
        p_mid[z, :, :] = 0.5 * (tm5_a[z] + tm5_b[z] * ps[:, :] + tm5_a[z + 1] + tm5_b[z + 1] * ps[:, :])
        if read_ak:
            sw[z, :, :] = aks[:, :, z] * amf_total
    if read_ak:
        sw = _clean(sw)
    if trop:
        layer = v["tm5_tropopause_layer_index"]
        tropopause = np.zeros_like(layer).astype("float16")
        inside = (layer > 0) & (layer < 34)
        ii, jj = np.nonzero(inside)
        tropopause[ii, jj] = p_mid[layer[ii, jj], ii, jj]
        tropopause[~inside] = np.nan
    else:
        tropopause = np.empty((1))
    return satellite_amf(vcd, amf, time, tropopause, lat, lon, [], [], unc, qf, p_mid, sw, [], [],
                         [], [], [])


# ------------------------------------------------------------------ satellite_opt products
def mopitt_co(v, read_ak=True):
    """reader.py:1143-1203 (mopitt_reader_co before its interpolator call).  `v` holds the
    MOP03 'Data Fields' variables plus StartTime / StopTime of the file attributes
    (_get_nc_attr_group_mopitt, :47-56)."""
    from oisatgmi_b200.config import satellite_opt
    t = 0.5 * (v["StartTime"] + v["StopTime"])
    time = _epoch(t, 1993)
    lat = v["Latitude"].astype("float32")
    lon = v["Longitude"].astype("float32")
    lon, lat = np.meshgrid(lon, lat)
    lon, lat = np.transpose(lon), np.transpose(lat)
    vcd = np.array(v["RetrievedCOTotalColumnDay"])
    vcd[np.where((vcd <= 0) | (np.isinf(vcd)))] = np.nan
    vcd = (vcd * 1e-15).astype("float16")
    dry = v["DryAirColumnDay"]
    with np.errstate(all="ignore"):
        x_col = (1e6 * vcd / (dry * 1e-15)).astype("float32")     # 1e6 * float16 overflows: kept
    ap_prof = np.array(v["APrioriCOMixingRatioProfileDay"]).transpose((2, 0, 1))
    ap_prof[ap_prof <= 0] = np.nan
    ap_sfc = np.array(v["APrioriCOSurfaceMixingRatioDay"])
    p_sfc = v["SurfacePressureDay"]
    ap_sfc[ap_sfc <= 0] = np.nan
    ap_col = (v["APrioriCOTotalColumnDay"] * 1e-15).astype("float16")
    ap_col[ap_col <= 0] = np.nan
    unc = (v["RetrievedCOTotalColumnMeanUncertaintyDay"] * 1e-15).astype("float32")
    ps = v["Pressure"].astype("float16")
    p_mid = np.zeros((9, np.shape(vcd)[0], np.shape(vcd)[1])).astype("float16")
    if read_ak == True:  # noqa: E712
        aks = v["TotalColumnAveragingKernelDay"] * 1e-15
        aks = aks.transpose((2, 0, 1)).astype("float16")
    else:
        aks = np.empty((1))
    for z in range(0, 9):
        p_mid[z, :, :] = ps[z]
    return satellite_opt(vcd, time, [], np.empty((1)), lat, lon, [], [], unc, np.ones_like(vcd),
                         p_mid, aks, [], [], [], [], ap_col, ap_prof, p_sfc, ap_sfc, x_col, [],
                         "MOPITT")


def gosat_xch4(v, read_ak=True):
    """reader.py:1228-1264 (gosat_reader_xch4 before the gap filler)."""
    from oisatgmi_b200.config import satellite_opt
    time = datetime.datetime(1970, 1, 1) + datetime.timedelta(
        seconds=int(np.squeeze(np.nanmean(v["time"]))))
    lat = v["latitude"].astype("float32")
    lon = v["longitude"].astype("float32")
    xch4 = np.array(v["xch4"])
    xch4[np.where((xch4 <= 0) | (np.isinf(xch4)))] = np.nan
    ap = np.array(v["ch4_profile_apriori"]).transpose()
    ap[ap <= 0] = np.nan
    qflag = v["xch4_quality_flag"]
    unc = v["xch4_uncertainty"]
    p_mid = np.array(v["pressure_levels"])
    p_mid[p_mid <= 0] = np.nan
    if read_ak == True:  # noqa: E712
        aks = np.array(v["xch4_averaging_kernel"]).transpose()
        pw = np.array(v["pressure_weight"]).transpose()
        aks[aks <= 0] = np.nan
        pw[pw <= 0] = np.nan
    else:
        aks, pw = np.empty((1)), np.empty((1))
    p_mid = np.transpose(p_mid)
    return satellite_opt(xch4, time, [], np.empty((1)), lat, lon, [], [], unc, 1 - qflag, p_mid,
                         aks, [], [], [], [], np.empty((1)), ap, np.empty((1)), np.empty((1)), xch4,
                         pw, "GOSAT")
