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


_FP_WEIGHTS = {}


def _fingerprint(a: np.ndarray) -> bytes:
    """Content fingerprint of a large array at memory speed: the 64-bit words, their wrapping
    sum and their wrapping dot product with fixed odd pseudo-random weights (any changed,
    swapped or moved word changes it).  blake2b over the 3.3 MB of a model grid's coordinates
    cost 4 ms per pipeline object; this costs 0.3 ms."""
    b = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    n8 = b.size // 8
    words = b[:8 * n8].view(np.uint64)
    w = _FP_WEIGHTS.get(n8)
    if w is None:
        rng = np.random.default_rng(0x0153A7)
        w = _FP_WEIGHTS[n8] = rng.integers(0, 2 ** 63, size=n8, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        while len(_FP_WEIGHTS) > 8:
            _FP_WEIGHTS.pop(next(iter(_FP_WEIGHTS)))
    with np.errstate(over="ignore"):
        s0 = np.add.reduce(words, dtype=np.uint64)
        s1 = np.dot(words, w)
    return b"%d:%d:" % (int(s0), int(s1)) + b[8 * n8:].tobytes()


def _digest(*arrays) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(_fingerprint(a) if a.nbytes >= (1 << 18) else a.tobytes())
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

    def dev_tables(self):
        """(window (n_cell, nwin) int32, nn_ok (n_cell) uint8) on the device."""
        if getattr(self, "_dev_tables", None) is None:
            self._dev_tables = (_dev.to_device(self.window.astype(np.int32)),
                                _dev.to_device(self.nn_ok.astype(np.uint8)))
        return self._dev_tables


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
    """Per-granule stencil: `cells` (flat output index of every kept cell, host
    int64) plus, for each kept cell, the 3*nwin (vertex, weight) entries.  The
    entries live either on the host (built by scipy, v0: `vert`/`w`,
    stencil-major (3*nwin, n_cells)) or on the device (built by K1, v1: pair-major
    (n_cells, 3*nwin)); the accessors convert lazily."""

    def __init__(self, gplan, cells, vert=None, w=None, keep=None, dev_pairs=None, builder="v0"):
        self.gplan = gplan
        self.keep = keep  # bool over mesh nodes (K0 predicate); host array or None
        self.cells = cells
        self.n_cells = int(cells.size)
        self.nwin = gplan.nwin
        self.builder = builder
        self._host = None if vert is None else (vert, w)
        self._pairs = dev_pairs      # (vert (n, S) int32, w (n, S) float64) device tensors
        self._stencil = None         # (cells int32, vert (S, n), w (S, n)) device tensors

    @property
    def vert(self):
        return self._host_arrays()[0]

    @property
    def w(self):
        return self._host_arrays()[1]

    def _host_arrays(self):
        if self._host is None:
            v, w = self._pairs
            self._host = (np.ascontiguousarray(_dev.to_host(v).T),
                          np.ascontiguousarray(_dev.to_host(w).T))
        return self._host

    def dev_pairs(self):
        """Pair-major device tensors for the fused kernel."""
        if self._pairs is None:
            v, w = self._host
            self._pairs = (_dev.to_device(np.ascontiguousarray(v.T)),
                           _dev.to_device(np.ascontiguousarray(w.T)))
        return self._pairs

    def dev(self):
        """(cells, vert, w) stencil-major device tensors for K2."""
        if self._stencil is None:
            cells = _dev.to_device(self.cells.astype(np.int32))
            if self._host is not None:
                self._stencil = (cells, _dev.to_device(self._host[0]), _dev.to_device(self._host[1]))
            else:
                v, w = self._pairs
                self._stencil = (cells, v.t().contiguous(), w.t().contiguous())
        return self._stencil


def coord_array(a):
    """Pixel coordinates for K0/K1: float32/float64 as delivered, anything else
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


def native_delaunay(lon, lat):
    """oisat_h_delaunay (csrc/delaunay.cpp): (triangles (n_tri, 3) int32, n_ties),
    or (None, 0) when no triangle exists.  The call releases the GIL."""
    tri, ties, _ = native_delaunay_path(lon, lat)
    return tri, ties


def native_delaunay_path(lon, lat):
    """As native_delaunay, plus which builder ran: 1 = structured-swath fast path
    (2-D lon/lat whose lattice quads are all convex), 0 = general sweep-hull.
    OISAT_DELAUNAY=general forces the latter."""
    import ctypes as C
    import os
    L = _lib.lib()
    shape = np.shape(lon)
    x = np.ascontiguousarray(np.asarray(lon, dtype=np.float64).ravel())
    y = np.ascontiguousarray(np.asarray(lat, dtype=np.float64).ravel())
    n = x.size
    if n < 3:
        return None, 0, 0
    tri = np.empty((2 * n, 3), dtype=np.int32)
    ties = C.c_int64(0)
    path = C.c_int32(0)
    if len(shape) == 2 and min(shape) >= 2 and os.environ.get("OISAT_DELAUNAY") != "general":
        nt = L.oisat_h_delaunay_swath(x.ctypes.data, y.ctypes.data, shape[0], shape[1],
                                      tri.ctypes.data, 2 * n, C.byref(ties), C.byref(path))
    else:
        nt = L.oisat_h_delaunay(x.ctypes.data, y.ctypes.data, n, tri.ctypes.data, 2 * n,
                                C.byref(ties))
    if nt <= 0:
        return None, 0, int(path.value)
    return tri[:nt], int(ties.value), int(path.value)


def native_delaunay_adj(lon, lat, pinned=False, device_index=None):
    """Triangulation for builder v1 with the near-tie scan left to the device:
    (tri, half, n_ties, max_abs_coord).  `half` (twin half-edge of every triangle edge) is
    None when the general builder ran -- then n_ties is already the complete report."""
    import ctypes as C
    import os
    L = _lib.lib()
    shape = np.shape(lon)
    if not (len(shape) == 2 and min(shape) >= 2 and os.environ.get("OISAT_DELAUNAY") != "general"
            and os.environ.get("OISAT_NEAR_TIES") != "host"):
        tri, ties = native_delaunay(lon, lat)
        return tri, None, ties, 0.0
    x = np.ascontiguousarray(np.asarray(lon, dtype=np.float64).ravel())
    y = np.ascontiguousarray(np.asarray(lat, dtype=np.float64).ravel())
    n = x.size
    if pinned:   # page-locked outputs: their upload does not block the calling thread
        t = _dev.torch()
        # (a pool thread's current device is 0 whatever the rank's device is: pin through the
        # caller's device so that a rank does not open a context on somebody else's GPU)
        with t.cuda.device(device_index if device_index is not None else t.cuda.current_device()):
            tri = t.empty((2 * n, 3), dtype=t.int32, pin_memory=True).numpy()
            half = t.empty((2 * n, 3), dtype=t.int32, pin_memory=True).numpy()
    else:
        tri = np.empty((2 * n, 3), dtype=np.int32)
        half = np.empty((2 * n, 3), dtype=np.int32)
    ties = C.c_int64(0)
    path = C.c_int32(0)
    nt = L.oisat_h_delaunay_swath_adj(x.ctypes.data, y.ctypes.data, shape[0], shape[1],
                                      tri.ctypes.data, 2 * n, half.ctypes.data, C.byref(ties),
                                      C.byref(path))
    if nt <= 0:
        return None, None, 0, 0.0
    if not path.value:
        return tri[:nt], None, int(ties.value), 0.0
    maxabs = float(max(np.abs(x).max(), np.abs(y).max()))
    return tri[:nt], half[:nt], int(ties.value), maxabs


def _seed_mode():
    """OISAT_DELAUNAY=device (default): structured swaths are triangulated by K12 -- host seed
    parts + flips on the device; host: the incremental swath builder; general: the sweep-hull."""
    import os
    return os.environ.get("OISAT_DELAUNAY", "device")


def native_seed_parts(lon, lat, pinned=False, device_index=None):
    """Host share of the device triangulation (oisat_h_delaunay_seed_parts): a dict with the
    quad table and the triangles outside the lattice, or None when the construction declines
    (folded lattices, ties on the hull of the seam: the incremental builder serves)."""
    L = _lib.lib()
    shape = np.shape(lon)
    if not (len(shape) == 2 and min(shape) >= 2):
        return None
    x = np.ascontiguousarray(np.asarray(lon, dtype=np.float64).ravel())
    y = np.ascontiguousarray(np.asarray(lat, dtype=np.float64).ravel())
    n = x.size
    nq = (shape[0] - 1) * (shape[1] - 1)
    cap = 2 * n // 3 + 16                       # triangles outside the lattice (seam <= n / 3)
    if pinned:
        t = _dev.torch()
        with t.cuda.device(device_index if device_index is not None else t.cuda.current_device()):
            buf = t.empty((nq + 6 * cap,), dtype=t.int32, pin_memory=True).numpy()
    else:
        buf = np.empty((nq + 6 * cap,), dtype=np.int32)
    qtri, otri, ohalf = buf[:nq], buf[nq:nq + 3 * cap], buf[nq + 3 * cap:]
    info = np.zeros(5, dtype=np.int64)
    maxabs = np.zeros(1, dtype=np.float64)
    nt = L.oisat_h_delaunay_seed_parts(x.ctypes.data, y.ctypes.data, shape[0], shape[1],
                                       qtri.ctypes.data, otri.ctypes.data, ohalf.ctypes.data, cap,
                                       info.ctypes.data, maxabs.ctypes.data)
    if nt <= 0:
        return None
    n_out = int(info[2])
    return dict(pinned=bool(pinned), qtri=qtri, otri=otri[:3 * n_out], ohalf=ohalf[:3 * n_out], n_quads=int(info[0]),
                n_outside=n_out, sigma=int(info[3]), n_tri=int(nt), rows=int(shape[0]),
                cols=int(shape[1]), maxabs=float(maxabs[0]))


def host_triangulation(lon, lat, pinned=False, device_index=None):
    """What a pool thread does for one granule: ("seed", parts) when K12 will finish the
    triangulation on the device, else ("tri", (tri, half, ties, maxabs)) from the host builders."""
    if _seed_mode() == "device":
        parts = native_seed_parts(lon, lat, pinned, device_index)
        if parts is not None:
            return "seed", parts
    return "tri", native_delaunay_adj(lon, lat, pinned, device_index)


def seed_assemble_device(parts):
    """Queues the upload of the seed parts and oisat_seed_assemble: (tri, half, keep-alive)."""
    L = _lib.lib()
    t = _dev.torch()
    dev = _dev.device()
    nt = parts["n_tri"]
    tri = _dev.empty((nt, 3), "int32")
    half = _dev.empty((nt, 3), "int32")
    if parts.get("pinned"):
        # page-locked parts are read by the kernel where they are (unified addressing: 0.5 MB
        # over the bus, coalesced).  As copies they would queue on the copy engine behind the
        # day's 365 MB of reader arrays and hold the flips back until those are through
        # (measured: the first batch of flips ended at 8.4 ms instead of ~6).
        qtri = otri = ohalf = None
        ptrs = (parts["qtri"].ctypes.data, parts["otri"].ctypes.data, parts["ohalf"].ctypes.data)
    else:
        qtri = t.from_numpy(parts["qtri"]).to(dev, non_blocking=True)
        otri = t.from_numpy(parts["otri"]).to(dev, non_blocking=True)
        ohalf = t.from_numpy(parts["ohalf"]).to(dev, non_blocking=True)
        ptrs = (qtri.data_ptr(), otri.data_ptr(), ohalf.data_ptr())
    _lib.check(L.oisat_seed_assemble(ptrs[0], parts["rows"], parts["cols"], parts["sigma"],
                                     parts["n_quads"], ptrs[1], ptrs[2],
                                     parts["n_outside"], tri.data_ptr(), half.data_ptr(), _dev.stream()))
    return tri, half, (qtri, otri, ohalf, parts)


def flip_batch_device(meshes):
    """Queues oisat_flip_delaunay_batch for [(tri, half, (lon_dev, lat_dev)), ...] -- the
    granules of a day go through the rounds of flips together.  Returns (result (n, 4) int64
    on the device: rounds and flips of the batch, then per mesh the edges left non-Delaunay
    and the edges the filter cannot decide, keep-alive).  Coordinates of one dtype."""
    L = _lib.lib()
    n = len(meshes)
    items = (_lib.FlipItem * n)()
    code = _dev.dtype_code(meshes[0][2][0])
    for k, (tri, half, (lo, la)) in enumerate(meshes):
        if _dev.dtype_code(lo) != code:
            raise _lib.OisatError("flip batch: coordinates of mixed dtypes")
        items[k] = _lib.FlipItem(tri.data_ptr(), half.data_ptr(), tri.shape[0], lo.data_ptr(), la.data_ptr())
    per_chunk = max(sum(m[0].shape[0] for m in meshes[c:c + 32]) for c in range(0, n, 32))
    work = _dev.empty((int(L.oisat_flip_workspace_bytes(per_chunk)),), "uint8")
    result = _dev.empty((n, 4), "int64")
    import ctypes as C
    _lib.check(L.oisat_flip_delaunay_batch(C.cast(items, C.c_void_p), n, code, work.data_ptr(),
                                           result.data_ptr(), _dev.stream()))
    return result, work


def device_triangulation(parts, lonlat_dev):
    """K12 for one granule: seed parts -> (tri, half, result[4], keep-alive) on the device;
    result[2:4] must be zero (read it after the stream has run) for the triangulation to be
    the Delaunay one."""
    tri, half, keep = seed_assemble_device(parts)
    result, work = flip_batch_device([(tri, half, lonlat_dev)])
    return tri, half, result[0], (keep, work)


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


def _plan_mode():
    import os
    return os.environ.get("OISAT_PLAN", "auto")


def _plan_v0(lon, lat, gplan, keep):
    """Host plan: Qhull + scipy's directed walk (exact for every input class)."""
    try:
        tri = triangulate(lon, lat)
    except Exception:
        return None
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
    return GranulePlan(gplan, cells, vert.astype(np.int32), w, keep, builder="v0")


def _plan_v0_device(lon, lat, lonlat_dev, gplan, keep_dev):
    """Qhull's triangulation with the device's point location and stencil fill: the fallback
    for a swath whose exact Delaunay triangulation has a NEAR tie inside a quadrilateral that
    holds a kept mesh node (Qhull merges such a quadrilateral into one facet and splits it its
    own way, so its answer is needed) -- but only its triangles are: scipy's sequential walk over
    the 1.04 M mesh nodes (0.3 s of the 0.7 s this fallback used to cost, and what made one
    granule in 120 stall an 8-rank step) is K1's job as in builder v1.  Lattice inputs, whose
    mesh nodes sit ON triangle edges everywhere, keep the host walk (_plan_v0): there the
    choice among the triangles sharing an edge decides which NaN vertices a node sees."""
    try:
        tri = triangulate(lon, lat)
    except Exception:
        return None
    simp = np.ascontiguousarray(tri.simplices.astype(np.int32))
    plan = _plan_v1_finish(_plan_v1_enqueue(simp, lonlat_dev, gplan, keep_dev), gplan)
    plan.builder = "v0q"
    return plan


def _plan_v1_enqueue(tri_host, lonlat_dev, gplan, keep_dev, half_host=None, maxabs=0.0, seed=None,
                     mesh=None):
    """First half of the device part of a v1 plan, queued without waiting for anything:
    upload of the triangulation (or, with `seed`, its construction on the device: K12; or,
    with `mesh` = (tri, half, maxabs, flip_host_row, keep-alive), a triangulation K12 has
    already been queued for as part of a batch),
    near-tie scan (with `half_host`, native_delaunay_adj), point location (K1), per-cell
    validity, and the copy of the per-cell flags (n_cell bytes) + tie count to pinned host
    memory.  Returns the state _plan_v1_finish needs."""
    L = _lib.lib()
    t = _dev.torch()
    dev = _dev.device()
    lo, la = lonlat_dev
    xs, ys = gplan.dev_axes()
    window, nn_ok = gplan.dev_tables()
    s = _dev.stream()
    ties_dev = tri_flag = flip_dev = flip_host = half = seed_keep = None
    if mesh is not None:
        tri, half, maxabs, flip_host, seed_keep = mesh
    elif seed is not None:
        tri, half, flip_dev, seed_keep = device_triangulation(seed, lonlat_dev)
        maxabs = seed["maxabs"]
    else:
        tri = t.from_numpy(tri_host).to(dev, non_blocking=True)    # asynchronous when pinned
        if half_host is not None:
            half = t.from_numpy(half_host).to(dev, non_blocking=True)
    if half is not None:
        ties_dev = _dev.empty((2,), "int64")     # [near ties, kept nodes inside tied quadrilaterals]
        tri_flag = _dev.zeros((tri.shape[0],), "uint8")
        _lib.check(L.oisat_near_ties(tri.data_ptr(), half.data_ptr(), tri.shape[0], lo.data_ptr(),
                                     la.data_ptr(), _dev.dtype_code(lo), float(maxabs),
                                     ties_dev.data_ptr(), tri_flag.data_ptr(), s))
    node_tri = _dev.full((gplan.H * gplan.W,), 2 ** 31 - 1, "int32")
    code = _dev.dtype_code(lo)
    work = _dev.empty((2 * tri.shape[0] + 2,), "int32")
    _lib.check(L.oisat_locate(tri.data_ptr(), tri.shape[0], lo.data_ptr(), la.data_ptr(), code,
                              xs.data_ptr(), gplan.W, ys.data_ptr(), gplan.H, keep_dev.data_ptr(),
                              node_tri.data_ptr(), work.data_ptr(), s))
    if tri_flag is not None:
        _lib.check(L.oisat_flagged_nodes(node_tri.data_ptr(), gplan.H * gplan.W, tri_flag.data_ptr(),
                                         ties_dev[1:].data_ptr(), s))
    n_cell = int(np.prod(gplan.out_shape))
    ok = _dev.empty((n_cell,), "uint8")
    _lib.check(L.oisat_plan_cells(window.data_ptr(), gplan.nwin, nn_ok.data_ptr(), n_cell,
                                  node_tri.data_ptr(), ok.data_ptr(), s))
    ok_host = t.empty((n_cell,), dtype=t.uint8, pin_memory=True)
    ok_host.copy_(ok, non_blocking=True)
    ties_host = None
    if ties_dev is not None:
        ties_host = t.empty((2,), dtype=t.int64, pin_memory=True)
        ties_host.copy_(ties_dev, non_blocking=True)
    if flip_dev is not None:
        flip_host = t.empty((4,), dtype=t.int64, pin_memory=True)
        flip_host.copy_(flip_dev, non_blocking=True)
    done = t.cuda.Event()
    done.record()
    return dict(tri=tri, node_tri=node_tri, ok_host=ok_host, ties_host=ties_host, done=done,
                flip_host=flip_host, lonlat=lonlat_dev, code=code,
                keep=(tri_host, half_host, half, work, ok, ties_dev, tri_flag, seed, seed_keep))


def _kept_cells(st):
    """Waits for the granule's flags and lists the kept cells (host)."""
    st["done"].synchronize()
    return np.flatnonzero(st["ok_host"].numpy().view(np.bool_))   # flags are 0 / 1


def _plan_v1_finish(st, gplan, cells=None):
    """Second half: waits for ITS flags only, lists the kept cells on the host and queues
    the stencil fill.  None when the near-tie scan found a tie (the caller takes v0)."""
    L = _lib.lib()
    st["done"].synchronize()
    if st["flip_host"] is not None:
        bad, unsure = int(st["flip_host"][2]), int(st["flip_host"][3])
        if bad or unsure:        # an edge the filter cannot decide: the exact host builder's job
            return "host"
    near_ties = affected = 0
    if st["ties_host"] is not None:
        near_ties, affected = int(st["ties_host"][0]), int(st["ties_host"][1])
    if affected != 0 and _plan_mode() != "v1":
        # a kept mesh node lies in a quadrilateral that is not uniquely split (or is inside
        # Qhull's tolerance): Qhull's own answer is needed.  Ties whose quadrilaterals hold no
        # kept node (the usual case for an isolated near-tie: a pixel quadrilateral is smaller
        # than the mesh spacing) cannot change any stencil and are accepted.
        return None
    if cells is None:
        cells = _kept_cells(st)
    lo, la = st["lonlat"]
    xs, ys = gplan.dev_axes()
    window, _ = gplan.dev_tables()
    n = cells.size
    S = 3 * gplan.nwin
    vert = _dev.empty((n, S), "int32")
    w = _dev.empty((n, S))
    cells_d = None
    if n:
        # the kernel reads the list from page-locked host memory (unified addressing): as a copy
        # it would wait on the copy engine behind the day's reader arrays
        t = _dev.torch()
        cells_d = t.empty((n,), dtype=t.int32, pin_memory=True)
        cells_d.numpy()[...] = cells
        _lib.check(L.oisat_plan_fill(cells_d.data_ptr(), n, window.data_ptr(), gplan.nwin,
                                     st["node_tri"].data_ptr(), st["tri"].data_ptr(), lo.data_ptr(),
                                     la.data_ptr(), st["code"], xs.data_ptr(), gplan.W, ys.data_ptr(), 1,
                                     vert.data_ptr(), w.data_ptr(), _dev.stream()))
    plan = GranulePlan(gplan, cells, keep=None, dev_pairs=(vert, w),
                       builder="v1d" if st["flip_host"] is not None else "v1")
    plan.near_ties = near_ties
    plan._cells_pinned = cells_d            # the buffer must outlive the launch that reads it
    if st["flip_host"] is not None:
        plan.flip_rounds, plan.flips = int(st["flip_host"][0]), int(st["flip_host"][1])
    return plan


def _plan_v1_device(tri_host, lonlat_dev, gplan, keep_dev, half_host=None, maxabs=0.0):
    """Device part of the v1 plan for one granule (enqueue + finish)."""
    return _plan_v1_finish(_plan_v1_enqueue(tri_host, lonlat_dev, gplan, keep_dev, half_host, maxabs),
                           gplan)


def triangulable(lon, lat) -> bool:
    """Would scipy.spatial.Delaunay accept these points?  Interpolator type 2 builds
    the triangulation although NearestNDInterpolator only uses its points, and skips
    the granule when that fails (interpolator.py:151-155): non-finite coordinates,
    fewer than three points, or all points on one line."""
    x = np.asarray(lon, dtype=np.float64).ravel()
    y = np.asarray(lat, dtype=np.float64).ravel()
    if x.size < 3 or not (np.isfinite(x).all() and np.isfinite(y).all()):
        return False
    d = np.flatnonzero((x != x[0]) | (y != y[0]))
    if d.size == 0:
        return False
    j = d[0]
    cross = (x - x[0]) * (y[j] - y[0]) - (y - y[0]) * (x[j] - x[0])
    return bool(np.any(cross != 0.0))


def nearest_plan(lon, lat, gplan: GridPlan, radius: float, lonlat_dev=None):
    """Stencil of the nearest-neighbour gridding modes (interpolator.py:17-20,28-33):
    every mesh node within `radius` of a pixel takes the value of its nearest pixel
    (K0's scatter with an atomic minimum), then the same box window / nearest-node
    composition as the linear mode.  Built entirely on the device."""
    L = _lib.lib()
    if lonlat_dev is None:
        lonlat_dev = (_dev.to_device(coord_array(lon)), _dev.to_device(coord_array(lat)))
    lo, la = lonlat_dev
    xs, ys = gplan.dev_axes()
    n_nodes = gplan.H * gplan.W
    s = _dev.stream()
    work = _dev.empty((n_nodes,), "int64")
    node_px = _dev.empty((n_nodes,), "int32")
    _lib.check(L.oisat_nearest_pixel(lo.data_ptr(), la.data_ptr(), _dev.dtype_code(lo), lo.numel(),
                                     xs.data_ptr(), gplan.W, ys.data_ptr(), gplan.H, float(radius),
                                     work.data_ptr(), node_px.data_ptr(), s))
    if gplan.upscale:
        window, nn_ok = gplan.dev_tables()
        n_cell = int(np.prod(gplan.out_shape))
        ok = _dev.empty((n_cell,), "uint8")
        _lib.check(L.oisat_plan_cells(window.data_ptr(), gplan.nwin, nn_ok.data_ptr(), n_cell,
                                      node_px.data_ptr(), ok.data_ptr(), s))
        cells = np.flatnonzero(_dev.to_host(ok))
        wptr = window.data_ptr()
    else:
        cells = np.flatnonzero(_dev.to_host(node_px) != 2 ** 31 - 1)
        wptr = None
    n = cells.size
    S = 3 * gplan.nwin
    vert = _dev.empty((n, S), "int32")
    w = _dev.empty((n, S))
    if n:
        cells_d = _dev.to_device(cells.astype(np.int32))
        _lib.check(L.oisat_plan_fill_nearest(cells_d.data_ptr(), n, wptr, gplan.nwin,
                                             node_px.data_ptr(), 1, vert.data_ptr(), w.data_ptr(), s))
    return GranulePlan(gplan, cells, keep=None, dev_pairs=(vert, w), builder="nearest")


def granule_plan(lon, lat, gplan: GridPlan, radius: float, lonlat_dev=None, cache=True,
                 keep=None):
    """Build (or fetch) the stencil of one granule.  Returns None when the pixel
    centres cannot be triangulated (interpolator.py:152-155).

    Builder v1 (default for model-grid targets): native Delaunay on the host,
    point location and stencil assembly on the GPU.  It is used when the native
    triangulation met no exact tie (collinear / co-circular input), i.e. when the
    Delaunay triangulation is unique and therefore the one Qhull returns; lattice
    inputs (MOPITT L3, GOSAT second pass) tie everywhere and go through builder v0
    (Qhull + scipy's own walk), cached by geometry.  OISAT_PLAN=v0 forces v0.
    `keep` (host bool array over the mesh nodes) replaces the K0 launch and forces
    v0; it exists so that the host-side composition can be unit-tested without a GPU."""
    lon = np.asarray(lon)
    lat = np.asarray(lat)
    key = (_digest(lon, lat), gplan.key, float(radius))
    if cache and key in _granule_plans:
        _granule_plans.move_to_end(key)
        return _granule_plans[key]
    plan = None
    if keep is not None:
        plan = _plan_v0(lon, lat, gplan, np.asarray(keep).astype(bool).ravel())
    else:
        if lonlat_dev is None:
            lonlat_dev = (_dev.to_device(coord_array(lon)), _dev.to_device(coord_array(lat)))
        keep_dev = distance_mask(lonlat_dev[0], lonlat_dev[1], gplan, radius)  # async
        use_v1 = gplan.upscale and _plan_mode() != "v0"
        if use_v1:
            kind, got = host_triangulation(lon, lat)      # runs while K0 is in flight
            near_tie = False                               # a v1 plan was refused for a near tie
            if kind == "seed":
                plan = _plan_v1_finish(_plan_v1_enqueue(None, lonlat_dev, gplan, keep_dev, seed=got), gplan)
                if isinstance(plan, str):                  # K12 met an undecidable edge
                    plan = None
                    kind, got = "tri", native_delaunay_adj(lon, lat)
                else:
                    near_tie = plan is None
            if kind == "tri":
                tri, half, ties, maxabs = got
                if tri is None:
                    return None
                if ties == 0 or _plan_mode() == "v1":
                    plan = _plan_v1_device(tri, lonlat_dev, gplan, keep_dev, half, maxabs)
                    near_tie = plan is None
            if near_tie and _plan_mode() != "v0walk":
                plan = _plan_v0_device(lon, lat, lonlat_dev, gplan, keep_dev)
        if plan is None:
            plan = _plan_v0(lon, lat, gplan, _dev.to_host(keep_dev).astype(bool))
    if plan is not None and cache:
        _granule_plans[key] = plan
        while len(_granule_plans) > _granule_plans.cap:
            _granule_plans.popitem(last=False)
    return plan


def clear_caches():
    _grid_plans.clear()
    _granule_plans.clear()
    _nn_tables.clear()


_POOLS = {}


def _plan_pool(workers):
    """The triangulation threads live as long as the process: a thread that has built one
    granule keeps its malloc arena warm, so the next day's 5 MB work arrays are not mapped and
    page-faulted again (measured on the build host: 8 granules on 8 threads 56-67 ms with a
    pool per call, 45-50 ms with this one)."""
    from concurrent.futures import ThreadPoolExecutor
    ex = _POOLS.get(workers)
    if ex is None:
        ex = _POOLS[workers] = ThreadPoolExecutor(workers, thread_name_prefix="oisat-plan")
    return ex


def _plan_workers(n, workers=None):
    import os
    if workers is None:
        # one process per GPU: the ranks of a node share its cores
        local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        workers = max(1, (os.cpu_count() or 1) // local)
    return max(1, min(n, workers))


def submit_triangulations(lons, lats, gplan: GridPlan, workers=None):
    """Puts the host share of the plans of a batch on the thread pool and returns at once
    ({future: index}, or None when this grid does not take builder v1): the caller can queue
    uploads and K0 while the pool works, then hand the futures to granule_plans."""
    if not gplan.upscale or _plan_mode() == "v0":
        return None
    n = len(lons)
    ex = _plan_pool(_plan_workers(n, workers))
    dev_index = _dev.device().index
    return {ex.submit(host_triangulation, lons[i], lats[i], True, dev_index): i for i in range(n)}


def granule_plans(lons, lats, gplan: GridPlan, radius: float, workers=None, lonlat_dev=None,
                  futures=None):
    """Plans for a batch of granules (lists of lon/lat arrays), as an event loop: K0 for all of
    them; the host shares of the triangulations on a thread pool (the C calls release the GIL;
    `futures`: already submitted, submit_triangulations); the seeds that have arrived are flipped
    on the device as batches (K12), each granule's device part follows its batch, and granules
    whose device part has run get their second half while the host waits for slower seeds."""
    import os
    n = len(lons)
    import time as _time
    t_start = _time.perf_counter()
    if futures is None:
        futures = submit_triangulations(lons, lats, gplan, workers)
    if lonlat_dev is None:
        lonlat_dev = [(_dev.to_device(coord_array(lons[i])), _dev.to_device(coord_array(lats[i])))
                      for i in range(n)]
    keeps = [distance_mask(lo, la, gplan, radius) for lo, la in lonlat_dev]
    if futures is None:
        return [granule_plan(lons[i], lats[i], gplan, radius, lonlat_dev=lonlat_dev[i], cache=False)
                for i in range(n)]
    out = [None] * n
    pending = []
    seeded = []
    # The first half of a granule's device part (uploads, near-tie scan, point location,
    # flags to pinned memory) is queued as soon as ITS triangulation is done, while the
    # others are still on the pool -- and without waiting for the GPU: with a
    # synchronisation per granule the 15 device parts of a day (2 ms each) ran one after
    # the other behind the slowest triangulation.
    trace = [] if os.environ.get("OISAT_PLAN_TRACE") == "1" else None
    marks = []

    def mark(label):
        if trace is not None:
            e = _dev.torch().cuda.Event(enable_timing=True)
            e.record()
            marks.append((label, (_time.perf_counter() - t_start) * 1e3, e))

    mark("K0 queued")

    def flush():
        # The seeds that have arrived go through the rounds of flips together (one launch: a
        # round costs its two barriers whatever it holds), then each granule's device part.
        if not seeded:
            return
        t = _dev.torch()
        groups = {}
        for k, sd in enumerate(seeded):
            groups.setdefault(_dev.dtype_code(lonlat_dev[sd[0]][0]), []).append(k)
        for ks in groups.values():
            result, work = flip_batch_device([(seeded[k][2], seeded[k][3], lonlat_dev[seeded[k][0]]) for k in ks])
            flip_host = t.empty((len(ks), 4), dtype=t.int64, pin_memory=True)
            flip_host.copy_(result, non_blocking=True)
            mark("flips of %d" % len(ks))
            for row, k in enumerate(ks):
                i, parts, tri, half, keep = seeded[k]
                mesh = (tri, half, parts["maxabs"], flip_host[row], (keep, work, result, parts))
                pending.append((i, _plan_v1_enqueue(None, lonlat_dev[i], gplan, keeps[i], mesh=mesh)))
        mark("K1 of the batch")
        del seeded[:]

    def handle(fut):
        i = futures[fut]
        kind, got = fut.result()
        t_a = _time.perf_counter()
        if kind == "seed":       # K12 finishes the triangulation on the device
            seeded.append((i, got) + seed_assemble_device(got))
        else:
            tri, half, ties, maxabs = got
            if tri is None:
                return
            if ties == 0 or _plan_mode() == "v1":
                pending.append((i, _plan_v1_enqueue(tri, lonlat_dev[i], gplan, keeps[i], half, maxabs)))
            else:
                out[i] = _plan_v0(lons[i], lats[i], gplan, _dev.to_host(keeps[i]).astype(bool))
                return
        if trace is not None:
            trace.append("%d:%.0f+%.1f" % (i, (t_a - t_start) * 1e3, (_time.perf_counter() - t_a) * 1e3))

    def finalize(i, st):
        # second half: kept cells on the host, stencil fill queued
        out[i] = _plan_v1_finish(st, gplan)
        if isinstance(out[i], str):      # K12 met an edge its filter cannot decide: exact builder
            tri, half, ties, maxabs = native_delaunay_adj(lons[i], lats[i])
            out[i] = None
            if tri is None:
                return
            if ties == 0 or _plan_mode() == "v1":
                out[i] = _plan_v1_device(tri, lonlat_dev[i], gplan, keeps[i], half, maxabs)
            else:
                out[i] = _plan_v0(lons[i], lats[i], gplan, _dev.to_host(keeps[i]).astype(bool))
                return
        if out[i] is None and _plan_mode() != "v0walk":      # near tie: Qhull's triangles, K1's walk
            out[i] = _plan_v0_device(lons[i], lats[i], lonlat_dev[i], gplan, keeps[i])
        if out[i] is None:
            out[i] = _plan_v0(lons[i], lats[i], gplan, _dev.to_host(keeps[i]).astype(bool))

    def finish_ready():
        # granules whose device part has run are finished while the host would otherwise wait
        # for the slower seeds of the day
        for k, (i, st) in enumerate(pending):
            ev = st.get("done")
            if ev is not None and ev.query():
                del pending[k]
                finalize(i, st)
                return True
        return False

    from concurrent.futures import wait, FIRST_COMPLETED
    not_done = set(futures)
    while not_done:
        ready = [f for f in not_done if f.done()]
        if ready:
            for f in ready:
                not_done.discard(f)
                handle(f)
            continue
        # nothing has arrived: do not let the device idle behind the slowest seed of the day (a
        # date-line crosser takes twice as long on the host) -- half a day's worth of seeds is a
        # batch of its own -- and use the wait for the finished granules' second halves
        if len(seeded) >= max(4, n // 2):
            flush()
        elif not finish_ready():
            wait(not_done, timeout=0.0005, return_when=FIRST_COMPLETED)
    flush()
    t_pool = _time.perf_counter()
    for i, st in list(pending):
        finalize(i, st)
    del pending[:]
    if trace is not None:
        import sys
        mark("end")
        _dev.torch().cuda.synchronize()
        print("granule_plans device marks (label: host ms / device ms): " +
              "; ".join("%s: %.1f / %.1f" % (lab, th, marks[0][2].elapsed_time(e) + marks[0][1])
                        for lab, th, e in marks), file=sys.stderr, flush=True)
        print("granule_plans trace: pool %.1f ms, finish %.1f ms; (granule:ready+enqueue ms) %s" %
              ((t_pool - t_start) * 1e3, (_time.perf_counter() - t_pool) * 1e3, " ".join(trace)),
              file=sys.stderr, flush=True)
    return out
