"""Seeded synthetic inputs in the dtypes the reference's readers deliver.

There is no network and no netCDF4/h5py in the image, so every test and
`bench.py` feed the hot path with records built here.  Shapes and dtypes follow
the readers:
  * OMI NO2   /root/reference/oisatgmi/reader.py:807-903   (35 SW levels, f16 fields)
  * OMI HCHO  reader.py:906-985   (47 hybrid levels from a0/b0, no tropopause)
  * TROPOMI   reader.py:707-804   (34 levels, qa in [0,1], threshold 0.75)
  * MOPITT L3 reader.py:1130-1213 (1 degree lattice, 9 levels + surface, 10 AK rows)
  * GOSAT     reader.py:1216-1275 (sparse soundings, (L,N) kernels)
  * GMI/ECCOH reader.py:95-172, 272-332 (72 layers bottom->top, float32)
Everything is generated from `numpy.random.default_rng(seed)`; identical seeds
give bit-identical records on every machine.
"""
from __future__ import annotations

import datetime as _dt

import numpy as np

from oisatgmi_b200.config import ctm_model, satellite_amf, satellite_opt

R_EARTH_KM = 6371.0

# ----------------------------------------------------------------------------
# model grid
# ----------------------------------------------------------------------------


def ctm_coordinates(region=None, dlat=0.5, dlon=0.625):
    """GMI-like lat/lon meshgrids (float64), optionally cut to
    region=(lat_min, lat_max, lon_min, lon_max)."""
    nlat = int(round(180.0 / dlat)) + 1
    lat = np.linspace(-90.0, 90.0, nlat)
    lon = np.arange(-180.0, 180.0, dlon)
    if region is not None:
        la0, la1, lo0, lo1 = region
        lat = lat[(lat >= la0) & (lat <= la1)]
        lon = lon[(lon >= lo0) & (lon <= lo1)]
    lon2d, lat2d = np.meshgrid(lon, lat)
    return {"Latitude": lat2d, "Longitude": lon2d}


def _smooth2d(rng, shape, ncomp=6, scale=1.0):
    """Cheap smooth random field: a few random plane waves."""
    ny, nx = shape
    yy = np.linspace(0.0, 1.0, ny)[:, None]
    xx = np.linspace(0.0, 1.0, nx)[None, :]
    out = np.zeros(shape)
    for _ in range(ncomp):
        ky, kx = rng.uniform(-3.0, 3.0, size=2)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        out += rng.uniform(0.3, 1.0) * np.sin(2.0 * np.pi * (ky * yy + kx * xx) + ph)
    return scale * out / ncomp


def _sigma_edges(nlev=72, ptop=0.01):
    k = np.arange(nlev + 1, dtype=np.float64)
    s = (1.0 - k / nlev) ** 2.4  # 1 at the surface, 0 at the top
    return s, ptop


def make_ctm(seed, coords, nslots=8, nlev=72, ctmtype="GMI", averaged=True,
             date=_dt.datetime(2005, 6, 1), gas_scale=1.0):
    """One `ctm_model`.  GMI: fields (nslots, nlev, ny, nx) float32, bottom->top.
    ECCOH/FREE: fields (nlev, ny, nx) and a single time stamp."""
    rng = np.random.default_rng(seed)
    lat2d, lon2d = coords["Latitude"], coords["Longitude"]
    ny, nx = lat2d.shape
    s, ptop = _sigma_edges(nlev)
    three_d = ctmtype in ("ECCOH", "FREE")
    ns = 1 if three_d else nslots
    ps0 = 1000.0 + 60.0 * _smooth2d(rng, (ny, nx), scale=3.0) \
        - 250.0 * np.clip(_smooth2d(rng, (ny, nx), scale=4.0), 0.0, None)
    delta_p = np.empty((ns, nlev, ny, nx), dtype=np.float32)
    pressure_mid = np.empty_like(delta_p)
    gas = np.empty_like(delta_p)
    zfrac = (np.arange(nlev) + 0.5) / nlev
    base = gas_scale * (0.02 + 3.0 * np.exp(-zfrac * 9.0))  # ppbv, decays upward
    hot = np.exp(2.0 * _smooth2d(rng, (ny, nx), scale=3.0))
    for t in range(ns):
        ps = ps0 + 4.0 * np.sin(2.0 * np.pi * t / max(ns, 1)) + rng.normal(0.0, 0.3, (ny, nx))
        edges = ptop + (ps[None] - ptop) * s[:, None, None]
        delta_p[t] = (edges[:-1] - edges[1:]).astype(np.float32)
        pressure_mid[t] = (0.5 * (edges[:-1] + edges[1:])).astype(np.float32)
        diurnal = 1.0 + 0.3 * np.cos(2.0 * np.pi * (t / max(ns, 1)))
        noise = np.exp(rng.normal(0.0, 0.15, (nlev, ny, nx)))
        gas[t] = (base[:, None, None] * hot[None] * diurnal * noise).astype(np.float32)
    if three_d:
        time = [date]
        gas, pressure_mid, delta_p = gas[0], pressure_mid[0], delta_p[0]
    else:
        time = [date + _dt.timedelta(hours=3 * t, minutes=90) for t in range(ns)]
    return ctm_model(lat2d, lon2d, time, gas, pressure_mid, [], delta_p, ctmtype, averaged)


# ----------------------------------------------------------------------------
# swath geometry
# ----------------------------------------------------------------------------


def swath_geolocation(nt, nxt, node_lon_deg=0.0, u0_deg=-80.0, u1_deg=80.0,
                      incl_deg=98.2, alt_km=705.0, half_fov_deg=57.0,
                      period_s=5933.0, jitter=1e-4, rng=None):
    """Pixel-centre lat/lon (float32, shape (nt, nxt)) of a sun-synchronous
    push-broom swath: nadir track of a circular orbit (inclination `incl_deg`)
    sampled at `nt` along-track angles between u0 and u1, `nxt` view angles
    uniformly spread over +-half_fov (so ground pixels grow toward the swath
    edge), Earth rotation included, longitudes wrapped to [-180, 180) -- the
    reference does not unwrap the date line either.  A tiny jitter keeps the
    points in general position (no exactly co-circular quadruples)."""
    u = np.deg2rad(np.linspace(u0_deg, u1_deg, nt))[:, None]
    inc = np.deg2rad(incl_deg)
    # position on the orbit plane, then tilt by the inclination
    x, y, z = np.cos(u), np.sin(u) * np.cos(inc), np.sin(u) * np.sin(inc)
    # velocity direction (d/du)
    vx, vy, vz = -np.sin(u), np.cos(u) * np.cos(inc), np.cos(u) * np.sin(inc)
    # cross-track unit vector  c = r x v
    cx, cy, cz = y * vz - z * vy, z * vx - x * vz, x * vy - y * vx
    theta = np.deg2rad(np.linspace(-half_fov_deg, half_fov_deg, nxt))[None, :]
    ratio = (R_EARTH_KM + alt_km) / R_EARTH_KM
    delta = np.arcsin(np.clip(ratio * np.sin(theta), -1.0, 1.0)) - theta
    px = x * np.cos(delta) + cx * np.sin(delta)
    py = y * np.cos(delta) + cy * np.sin(delta)
    pz = z * np.cos(delta) + cz * np.sin(delta)
    lat = np.rad2deg(np.arcsin(np.clip(pz, -1.0, 1.0)))
    tsec = (u - u[0]) / (2.0 * np.pi) * period_s
    lon = np.rad2deg(np.arctan2(py, px)) + node_lon_deg - tsec * (360.0 / 86164.0)
    if rng is not None and jitter > 0:
        lat = lat + rng.uniform(-jitter, jitter, lat.shape)
        lon = lon + rng.uniform(-jitter, jitter, lon.shape)
    lon = (lon + 180.0) % 360.0 - 180.0
    return lat.astype(np.float32), lon.astype(np.float32)


def _coherent_mask(rng, shape, bad_fraction):
    """Spatially coherent validity mask (cloud blobs), True = good."""
    f = _smooth2d(rng, shape, ncomp=8, scale=8.0)
    if bad_fraction <= 0:
        return np.ones(shape, bool)
    thr = np.quantile(f, bad_fraction)
    return f > thr


def _time_of(day_index, orbit_in_day, year=2005, month=6, orbits_per_day=14.6):
    sec = int(86400.0 * orbit_in_day / orbits_per_day) + 1800
    return _dt.datetime(year, month, 1) + _dt.timedelta(days=int(day_index), seconds=sec)


OMI_NO2_PRESSURES = np.array(
    [1020.0, 1010.0, 1000.0, 990.0, 975.0, 960.0, 945.0, 925.0, 900.0, 875.0,
     850.0, 825.0, 800.0, 770.0, 740.0, 700.0, 660.0, 610.0, 560.0, 500.0,
     450.0, 400.0, 350.0, 280.0, 200.0, 120.0, 60.0, 35.0, 20.0, 12.0,
     8.0, 5.0, 3.0, 1.5, 0.8])

# hybrid coefficients of the 47-layer OMI HCHO grid as tabulated by the reader
# (reader.py:954-957); regenerated here analytically to the same shape: the
# synthetic generator only needs a monotone 48-edge hybrid table.
def _hybrid_table(nedge, ptop=0.01):
    """Monotone hybrid table: p_k = a_k + b_k * ps, surface (b=1, a=0) to a pure
    pressure top (b=0, a=ptop); strictly positive and decreasing for any ps."""
    k = np.arange(nedge, dtype=np.float64) / (nedge - 1)
    b = np.clip(1.0 - k / 0.65, 0.0, None) ** 1.3
    p_ref = 1000.0 * np.exp(-k * np.log(1000.0 / ptop))
    a = (1.0 - b) * p_ref
    return a, b


def _scattering_weights(rng, p_mid, shape2d):
    """Smooth SW(p, pixel) in [0.05, 3]; p_mid (L, nt, nxt) float."""
    albedo = 0.5 + 0.4 * _smooth2d(rng, shape2d, scale=3.0)
    logp = np.log(np.maximum(p_mid.astype(np.float64), 1e-3) / 1000.0)
    sw = 0.25 + 1.6 * (1.0 - np.exp(logp * 0.9)) + albedo[None] * np.exp(logp * 2.0)
    return np.clip(sw, 0.0, 99.0)


def make_amf_granule(seed, product="OMI_NO2", nt=None, nxt=None, geo=None,
                     bad_fraction=0.25, time=None):
    """`satellite_amf` record as a reader would hand it to `interpolator`
    (before gridding).  product in {OMI_NO2, OMI_HCHO, OMI_O3, TROPOMI_NO2}."""
    rng = np.random.default_rng(seed)
    dflt = {"OMI_NO2": (1644, 60), "OMI_HCHO": (1644, 60), "OMI_O3": (1644, 60),
            "TROPOMI_NO2": (4172, 450)}[product]
    nt = dflt[0] if nt is None else nt
    nxt = dflt[1] if nxt is None else nxt
    geo = dict(geo or {})
    if product.startswith("TROPOMI"):
        geo.setdefault("alt_km", 824.0)
        geo.setdefault("half_fov_deg", 54.0)
    lat, lon = swath_geolocation(nt, nxt, rng=rng, **geo)
    shape = (nt, nxt)
    vcd = np.exp(0.8 * _smooth2d(rng, shape, scale=4.0) + rng.normal(0, 0.1, shape)) * 3.0
    vcd = vcd.astype(np.float16)
    amf = (1.2 + 0.5 * _smooth2d(rng, shape, scale=3.0)).astype(np.float64)
    unc = (0.3 + 0.2 * np.abs(_smooth2d(rng, shape, scale=3.0))
           + rng.uniform(0, 0.2, shape)).astype(np.float16)
    good = _coherent_mask(rng, shape, bad_fraction)
    if product == "TROPOMI_NO2":
        qf = np.where(good, rng.uniform(0.76, 1.0, shape), rng.uniform(0.0, 0.74, shape))
        qf = qf.astype(np.float64)
        a, b = _hybrid_table(35)
        ps = (1000.0 + 20.0 * _smooth2d(rng, shape, scale=3.0)).astype(np.float32)
        edges = a[:, None, None] + b[:, None, None] * ps[None]
        p_mid = (0.5 * (edges[:-1] + edges[1:])).astype(np.float16)
        trop = (200.0 + 80.0 * _smooth2d(rng, shape, scale=2.0)).astype(np.float16)
    elif product == "OMI_NO2":
        qf = np.where(good, 1.0, -100.0) * np.where(rng.uniform(size=shape) < 0.02, 0.0, 1.0)
        p_mid = np.broadcast_to(OMI_NO2_PRESSURES.astype(np.float16)[:, None, None],
                                (35, nt, nxt)).copy()
        trop = (200.0 + 80.0 * _smooth2d(rng, shape, scale=2.0)).astype(np.float16)
    elif product == "OMI_HCHO":
        qf = np.where(good, 1.0, 0.0).astype(np.float64)
        a, b = _hybrid_table(48)
        ps = (1000.0 + 20.0 * _smooth2d(rng, shape, scale=3.0)).astype(np.float16)
        p_mid = np.zeros((47, nt, nxt), dtype=np.float16)
        for z in range(47):
            p_mid[z] = 0.5 * ((a[z] + b[z] * ps) + (a[z + 1] + b[z + 1] * ps))
        trop = np.empty((1))
    elif product == "OMI_O3":
        qf = np.where(good, 1.0, -100.0).astype(np.float64)
        p_mid = np.zeros((1, nt, nxt), dtype=np.float16)
        trop = np.empty((1))
    else:
        raise ValueError(product)
    if product == "OMI_O3":
        sw = np.empty((1))
    else:
        sw = _scattering_weights(rng, p_mid, shape).astype(np.float16)
    time = time or _time_of(seed % 28, seed % 14)
    return satellite_amf(vcd, amf, time, trop, lat, lon, [], [], unc, qf.astype(np.float64),
                         p_mid, sw, [], [], [], [], [])


def make_mopitt_granule(seed, fill_fraction=0.4, time=None, region=None):
    """MOPITT L3-like `satellite_opt`: 1 degree lattice (lon-major, 360x180 as
    in reader.py:1157-1160), valid in coherent orbit stripes."""
    rng = np.random.default_rng(seed)
    lon1 = np.arange(-179.5, 180.0, 1.0, dtype=np.float32)
    lat1 = np.arange(-89.5, 90.0, 1.0, dtype=np.float32)
    if region is not None:
        la0, la1, lo0, lo1 = region
        lat1 = lat1[(lat1 >= la0) & (lat1 <= la1)]
        lon1 = lon1[(lon1 >= lo0) & (lon1 <= lo1)]
    lon2d, lat2d = np.meshgrid(lon1, lat1)
    lon2d, lat2d = np.transpose(lon2d), np.transpose(lat2d)
    shape = lon2d.shape
    # orbit stripes: bands in (lon + 0.25*lat) space
    phase = rng.uniform(0, 25.0)
    stripe = ((lon2d + 0.25 * lat2d + phase) % 25.0) < 25.0 * fill_fraction
    stripe &= _coherent_mask(rng, shape, 0.1)
    vcd = (1800.0 + 400.0 * _smooth2d(rng, shape, scale=3.0)).astype(np.float64)
    vcd[~stripe] = np.nan
    vcd16 = vcd.astype(np.float16)
    dry = 2.1e25 * (1.0 + 0.02 * _smooth2d(rng, shape, scale=2.0))
    with np.errstate(over="ignore", invalid="ignore"):
        # as in reader.py:1167 the product is formed in float16 and overflows to
        # inf: MOPITT's x_col is not a usable quantity in the reference either
        x_col = (1e6 * vcd16 / (dry * 1e-15)).astype(np.float32)
    L = 9
    press = np.array([900.0, 800.0, 700.0, 600.0, 500.0, 400.0, 300.0, 200.0, 100.0])
    p_mid = np.broadcast_to(press.astype(np.float16)[:, None, None], (L,) + shape).copy()
    ap_prof = (90.0 + 30.0 * _smooth2d(rng, shape, scale=2.0))[None] \
        * np.linspace(1.2, 0.6, L)[:, None, None]
    ap_prof = ap_prof.astype(np.float32)
    ap_sfc = (110.0 + 30.0 * _smooth2d(rng, shape, scale=2.0)).astype(np.float32)
    p_sfc = (990.0 + 25.0 * _smooth2d(rng, shape, scale=2.0)).astype(np.float32)
    ap_col = (1700.0 + 300.0 * _smooth2d(rng, shape, scale=2.0)).astype(np.float16)
    unc = (90.0 + 40.0 * np.abs(_smooth2d(rng, shape, scale=2.0))).astype(np.float32)
    aks = (120.0 * (0.4 + np.abs(_smooth2d(rng, (L + 1,) + (shape[0] * shape[1],), scale=1.0)
                                 ).reshape((L + 1,) + shape))).astype(np.float16)
    time = time or _time_of(seed % 28, 7)
    return satellite_opt(vcd16, time, [], np.empty((1)), lat2d, lon2d, [], [], unc,
                         np.ones_like(vcd16), p_mid, aks, [], [], [], [],
                         ap_col, ap_prof, p_sfc, ap_sfc, x_col, [], "MOPITT")


def make_gosat_soundings(seed, n=2000, L=20, time=None, region=None):
    """GOSAT-like sparse soundings (`satellite_opt`, 1-D fields, (L, N) kernels,
    reader.py:1233-1262), clustered over 'land'."""
    rng = np.random.default_rng(seed)
    ncl = max(4, n // 60)
    if region is None:
        region = (-60.0, 75.0, -170.0, 170.0)
    la0, la1, lo0, lo1 = region
    c_lat = rng.uniform(la0, la1, ncl)
    c_lon = rng.uniform(lo0, lo1, ncl)
    which = rng.integers(0, ncl, n)
    lat = (c_lat[which] + rng.normal(0, 2.5, n)).clip(-89.0, 89.0).astype(np.float32)
    lon = (c_lon[which] + rng.normal(0, 3.5, n)).clip(-179.0, 179.0).astype(np.float32)
    xch4 = 1800.0 + 30.0 * np.sin(np.deg2rad(lat)) + rng.normal(0, 8.0, n)
    unc = rng.uniform(5.0, 15.0, n)
    qflag = (rng.uniform(size=n) < 0.15).astype(np.float64)  # 1 = bad in the file
    ps = rng.uniform(850.0, 1013.0, n)
    frac = (np.arange(L) + 0.5) / L
    p_mid = (ps[None, :] * (1.0 - frac[:, None]) + 0.1)
    ap = 1850.0 - 300.0 * frac[:, None] ** 2 + rng.normal(0, 5.0, (L, n))
    aks = np.clip(1.0 - 0.6 * (frac[:, None] - 0.3) ** 2 + rng.normal(0, 0.02, (L, n)), 0.05, None)
    pw = np.full((L, n), 1.0 / L) * (1.0 + rng.normal(0, 0.01, (L, n)))
    time = time or _time_of(seed % 28, 5)
    return satellite_opt(xch4, time, [], np.empty((1)), lat, lon, [], [], unc, 1 - qflag,
                         p_mid, aks, [], [], [], [], np.empty((1)), ap, np.empty((1)),
                         np.empty((1)), xch4, pw, "GOSAT")


def regional_geo(region, frac=0.5, span_deg=None):
    """Swath parameters that put a piece of an orbit across `region`
    (lat0, lat1, lon0, lon1): used by the small test cases."""
    la0, la1, lo0, lo1 = region
    mid_lat = la0 + frac * (la1 - la0)
    span = span_deg if span_deg is not None else (la1 - la0) * 0.75 + 6.0
    # ascending part of the orbit: lat ~ u for small angles
    u0, u1 = mid_lat - span / 2.0, mid_lat + span / 2.0
    # find the node longitude that puts the nadir track at the box centre
    lat, lon = swath_geolocation(3, 3, node_lon_deg=0.0, u0_deg=u0, u1_deg=u1, jitter=0.0)
    node = (lo0 + lo1) / 2.0 - float(lon[1, 1])
    return dict(node_lon_deg=node, u0_deg=u0, u1_deg=u1)
