"""Geometry plans: everything `interpolator` does that depends only on WHERE the
pixels are, not on what they measured.

The reference re-derives the same geometry for every one of the ~74 fields of a
granule (interpolator.py:162-283: LinearNDInterpolator construction + point
location, a KD-tree over the 1.04 M fine nodes rebuilt per call, a box
convolution).  All of that is linear in the pixel values, so it collapses to one
stencil per output cell:

    out[c] = sum_{f in window(nn[c])} box_weight * sum_{k<3} w[f,k] * z[vert[f,k]]

NaN if any node of the window is outside the convex hull / beyond the distance
mask, or any vertex value is NaN.  Two plan levels:

  * GridPlan     constant per (model grid, grid_size): fine mesh axes, box window,
                 nearest-node table.  The nearest-node table has EXACT distance ties
                 on half of the GMI longitude columns and scipy's KD-tree breaks them
                 in traversal order, so it is produced by the
                 same cKDTree call the reference makes (interpolator.py:82-88), once.
  * GranulePlan  per pixel geometry: Delaunay triangulation (Qhull, like the
                 reference: interpolator.py:153), point location with the same
                 directed walk (`Delaunay.find_simplex`) and barycentric weights
                 from `Delaunay.transform` -- which reproduces LinearNDInterpolator's
                 NaN mask exactly (appendix E) -- composed with the GridPlan.
                 The distance mask comes from the GPU (K0, oisat_distmask).
                 Plans are cached by a hash of the pixel coordinates, so
                 constant-geometry products (MOPITT L3, GOSAT second pass) pay
                 for the triangulation once.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np
from scipy.spatial import Delaunay, cKDTree

from . import _dev, _lib

# meshes up to this many nodes are located in one in-order find_simplex call,
# exactly like LinearNDInterpolator walks them (matters only for degenerate,
# lattice-like inputs where ties depend on the walk's start simplex); larger
# meshes (swaths, general position) only locate the nodes that survive K0.
FULL_WALK_MAX_NODES = 400_000


def _digest(*arrays) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class _LRU(OrderedDict):
    def __init__(self, cap):
        super().__init__()
        self.cap = cap

    def get_or(self, key, make):
        if key in self:
            self.move_to_end(key)
            return self[key]
        val = make()
        self[key] = val
        while len(self) > self.cap:
            self.popitem(last=False)
        return val


_grid_plans = _LRU(8)
_granule_plans = _LRU(16)
_nn_tables = _LRU(8)


def grid_spacing(coords):
    lon, lat = coords["Longitude"], coords["Latitude"]
    return np.abs(lon[0, 0] - lon[0, 1]), np.abs(lat[0, 0] - lat[1, 0])


def _reflect(i, n):
    """scipy.signal.convolve2d(boundary='symm') index reflection."""
    i = np.asarray(i).copy()
    for _ in range(4):
        i = np.where(i < 0, -i - 1, i)
        i = np.where(i >= n, 2 * n - 1 - i, i)
    return i


def box_extent(dlon, dlat, grid_size):
    """(ky, kx) of interpolator.py:66-71."""
    kx = np.floor(dlon / grid_size)
    ky = np.floor(dlat / grid_size)
    return int(1 if ky == 0 else ky), int(1 if kx == 0 else kx)


def nearest_node_table(X, Y, tlon, tlat):
    """`cKDTree(mesh nodes).query(target points)` exactly as interpolator.py:78-88
    (tie-breaking is scipy's; the table is geometry-only and cached)."""
    key = _digest(X, Y, tlon, tlat)

    def make():
        pts = np.zeros((np.size(X), 2))
        pts[:, 0] = np.asarray(X).flatten()
        pts[:, 1] = np.asarray(Y).flatten()
        tree = cKDTree(pts)
        q = np.stack(np.broadcast_arrays(tlon, tlat), axis=-1).astype(np.float64)
        d, idx = tree.query(q)
        return d.ravel(), idx.ravel().astype(np.int64)

    return _nn_tables.get_or(key, make)


def box_window(H, W, node, ky, kx):
    """Fine-mesh flat indices of the (ky, kx) window that convolve2d('same',
    'symm') averages into `node` (offsets
    range(-(k//2), k-(k//2)), reflected at the edges).  Shape (n, ky*kx)."""
    r0, c0 = node // W, node % W
    rr = _reflect(r0[:, None] + np.arange(-(ky // 2), ky - ky // 2)[None, :], H)
    cc = _reflect(c0[:, None] + np.arange(-(kx // 2), kx - kx // 2)[None, :], W)
    return (rr[:, :, None] * W + cc[:, None, :]).reshape(len(node), ky * kx)


class GridPlan:
    """Constant part of the plan for one (model grid, grid_size)."""

    def __init__(self, coords, grid_size, mesh=None):
        lat, lon = np.asarray(coords["Latitude"]), np.asarray(coords["Longitude"])
        self.grid_size = float(grid_size)
        self.ctm_shape = lat.shape
        self.ctm_lat, self.ctm_lon = lat, lon
        dlon, dlat = grid_spacing(coords)
        self.threshold_ctm = np.sqrt(dlon ** 2 + dlat ** 2)
        if mesh is None:
            # interpolator.py:136-143
            self.lon_axis = np.arange(lon.min(), lon.max() + grid_size, grid_size)
            self.lat_axis = np.arange(lat.min(), lat.max() + grid_size, grid_size)
        else:
            self.lon_axis, self.lat_axis = (np.asarray(m, dtype=np.float64) for m in mesh)
        self.W, self.H = len(self.lon_axis), len(self.lat_axis)
        self.upscale = bool((dlon >= grid_size) or (dlat >= grid_size))  # interpolator.py:64
        if self.upscale:
            self.ky, self.kx = box_extent(dlon, dlat, grid_size)
            X, Y = np.meshgrid(self.lon_axis, self.lat_axis)
            d, idx = nearest_node_table(X, Y, lon, lat)
            self.nn = idx
            self.nn_ok = ~(d > self.threshold_ctm * 2.0)  # interpolator.py:33 via :90-91
            self.window = box_window(self.H, self.W, idx, self.ky, self.kx)
            self.out_shape = lat.shape
        else:
            self.ky = self.kx = 1
            self.out_shape = (self.H, self.W)
        self._dev_axes = None

    @property
    def nwin(self):
        return self.ky * self.kx

    def mesh(self):
        return np.meshgrid(self.lon_axis, self.lat_axis)

    def dev_axes(self):
        if self._dev_axes is None:
            self._dev_axes = (_dev.to_device(self.lon_axis), _dev.to_device(self.lat_axis))
        return self._dev_axes


def grid_plan(coords, grid_size, mesh=None) -> GridPlan:
    key = (float(grid_size), _digest(coords["Latitude"], coords["Longitude"]),
           None if mesh is None else _digest(*mesh))
    gp = _grid_plans.get_or(key, lambda: GridPlan(coords, grid_size, mesh))
    gp.key = key
    return gp


def distance_mask(lon_dev, lat_dev, gplan: GridPlan, radius: float) -> np.ndarray:
    """K0 on the GPU: keep[f] = not (distance to nearest pixel > radius)."""
    L = _lib.lib()
    xs, ys = gplan.dev_axes()
    keep = _dev.zeros((gplan.H * gplan.W,), "uint8")
    _lib.check(L.oisat_distmask(lon_dev.data_ptr(), lat_dev.data_ptr(), _dev.dtype_code(lon_dev),
                                lon_dev.numel(), xs.data_ptr(), gplan.W, ys.data_ptr(), gplan.H,
                                float(radius), keep.data_ptr(), _dev.stream()))
    return keep


class GranulePlan:
    """Per-granule stencil: `cells` (flat output index), `vert`/`w` of shape
    (3*nwin, n_cells)."""

    def __init__(self, gplan, cells, vert, w, keep=None):
        self.gplan = gplan
        self.keep = keep  # bool over mesh nodes: K0 predicate
        self.cells = cells
        self.vert = vert
        self.w = w
        self.n_cells = int(cells.size)
        self.nwin = gplan.nwin
        self._dev = None

    def dev(self):
        if self._dev is None:
            self._dev = (_dev.to_device(self.cells.astype(np.int32)),
                         _dev.to_device(self.vert), _dev.to_device(self.w))
        return self._dev


def coord_array(a):
    """Pixel coordinates for K0: float32/float64 as delivered, anything else
    (e.g. the filler's float16 mesh) widened exactly to float64."""
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a).ravel()


def triangulate(lon, lat):
    pts = np.zeros((np.size(lat), 2))
    pts[:, 0] = np.asarray(lon).flatten()
    pts[:, 1] = np.asarray(lat).flatten()
    return Delaunay(pts)  # Qhull; raises on degenerate input (caller returns None)


def locate(tri, qx, qy):
    """Containing simplex and barycentric weights of each query point, evaluated
    the way scipy's LinearNDInterpolator does (qhull._barycentric_coordinates)."""
    q = np.column_stack((qx, qy))
    s = tri.find_simplex(q)
    inside = s >= 0
    T = tri.transform[s[inside]]
    d0 = qx[inside] - T[:, 2, 0]
    d1 = qy[inside] - T[:, 2, 1]
    c0 = T[:, 0, 0] * d0 + T[:, 0, 1] * d1
    c1 = T[:, 1, 0] * d0 + T[:, 1, 1] * d1
    c2 = (1.0 - c0) - c1
    wts = np.full((len(s), 3), np.nan)
    wts[inside, 0], wts[inside, 1], wts[inside, 2] = c0, c1, c2
    verts = np.zeros((len(s), 3), dtype=np.int32)
    verts[inside] = tri.simplices[s[inside]]
    return s, verts, wts


def granule_plan(lon, lat, gplan: GridPlan, radius: float, lonlat_dev=None, cache=True,
                 keep=None):
    """Build (or fetch) the stencil of one granule.  Returns None when Qhull
    cannot triangulate the pixel centres (interpolator.py:152-155).  `keep`
    (host bool array over the mesh nodes) replaces the K0 launch; it exists so
    that the host-side composition can be unit-tested without a GPU."""
    lon = np.asarray(lon)
    lat = np.asarray(lat)
    key = (_digest(lon, lat), gplan.key, float(radius))
    if cache and key in _granule_plans:
        _granule_plans.move_to_end(key)
        return _granule_plans[key]
    keep_dev = None
    if keep is None:
        if lonlat_dev is None:
            lonlat_dev = (_dev.to_device(coord_array(lon)), _dev.to_device(coord_array(lat)))
        keep_dev = distance_mask(lonlat_dev[0], lonlat_dev[1], gplan, radius)  # async
    try:
        tri = triangulate(lon, lat)  # Qhull runs while K0 is in flight
    except Exception:
        return None
    if keep_dev is not None:
        keep = _dev.to_host(keep_dev)
    keep = np.asarray(keep).astype(bool).ravel()
    n_nodes = gplan.H * gplan.W
    if n_nodes <= FULL_WALK_MAX_NODES:
        cand = np.arange(n_nodes)
    else:
        cand = np.flatnonzero(keep)
    qx = gplan.lon_axis[cand % gplan.W]
    qy = gplan.lat_axis[cand // gplan.W]
    s, verts, wts = locate(tri, qx, qy)
    valid = np.zeros(n_nodes, dtype=bool)
    valid[cand] = (s >= 0)
    valid &= keep
    pos = np.zeros(n_nodes, dtype=np.int64)
    pos[cand] = np.arange(len(cand))
    if gplan.upscale:
        ok = gplan.nn_ok & valid[gplan.window].all(axis=1)
        cells = np.flatnonzero(ok)
        nodes = gplan.window[cells]              # (n, nwin)
    else:
        cells = np.flatnonzero(valid)
        nodes = cells[:, None]
    sel = pos[nodes]                              # (n, nwin)
    S = 3 * gplan.nwin
    vert = np.ascontiguousarray(verts[sel].reshape(len(cells), S).T)   # (3*nwin, n)
    w = np.ascontiguousarray(wts[sel].reshape(len(cells), S).T)
    plan = GranulePlan(gplan, cells, vert.astype(np.int32), w, keep)
    if cache:
        _granule_plans[key] = plan
        while len(_granule_plans) > _granule_plans.cap:
            _granule_plans.popitem(last=False)
    return plan


def clear_caches():
    _grid_plans.clear()
    _granule_plans.clear()
    _nn_tables.clear()


# ---------------------------------------------------------------------------
# many granules at once: K0 on the GPU for all of them, then the host part
# (Qhull + walk) on every core.  Worker processes are forked AFTER the K0 masks
# are on the host and never touch CUDA.
# ---------------------------------------------------------------------------
_pool_job = None


def _pool_worker(i):
    lons, lats, keeps, gplan, radius = _pool_job
    p = granule_plan(lons[i], lats[i], gplan, radius, keep=keeps[i], cache=False)
    if p is None:
        return None
    return p.cells, p.vert, p.w


def granule_plans(lons, lats, gplan: GridPlan, radius: float, workers=None, lonlat_dev=None):
    """Plans for a batch of granules (list of lon/lat arrays)."""
    import multiprocessing as mp
    import os
    global _pool_job
    n = len(lons)
    keeps_dev = []
    for i in range(n):
        if lonlat_dev is not None:
            lo, la = lonlat_dev[i]
        else:
            lo, la = _dev.to_device(coord_array(lons[i])), _dev.to_device(coord_array(lats[i]))
        keeps_dev.append(distance_mask(lo, la, gplan, radius))
    keeps = [_dev.to_host(k).astype(bool) for k in keeps_dev]
    workers = min(n, workers or os.cpu_count() or 1)
    _pool_job = (lons, lats, keeps, gplan, radius)
    try:
        if workers <= 1:
            raw = [_pool_worker(i) for i in range(n)]
        else:
            with mp.get_context("fork").Pool(workers) as pool:
                raw = pool.map(_pool_worker, range(n), chunksize=1)
    finally:
        _pool_job = None
    out = []
    for r, k in zip(raw, keeps):
        out.append(None if r is None else GranulePlan(gplan, r[0], r[1], r[2], k))
    return out
