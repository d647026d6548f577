"""Fused month pipeline: the whole hot path of one (year, month) bin for float16
`satellite_amf` products (OMI NO2 / HCHO, TROPOMI NO2) without ever
materialising a gridded granule.

    reader records (host) --upload--> HBM
    K0  oisat_distmask        \\  geometry plan per granule (native Delaunay on the host,
    plan.granule_plan(s)      /   K1 point location on the GPU), concatenated into pair tables
    oisat_pack_batch          quality mask + pixel-major float16 records, one launch per month
    oisat_fused_amf_split     gather-interpolate, then AMF recalculation per (granule, cell)
    oisat_accum_pairs         ordered segmented reduction -> [10][n_cell] sums / counts
    (torch.distributed all_reduce of the accumulator block when sharded, section 8e)
    oisat_accum_finalize, oisat_oi_prepare, oisat_oi_sweep, oisat_oi_knee, oisat_oi_apply_dev
    (knee on the host, through the reference's own `kneed`, when that package is importable)

It computes exactly what `interpolator` -> `amf_recal` -> `averaging` ->
`bias_correct` -> `oi` compute through the drop-in modules (same device
functions), and is what `bench.py` times.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import _dev, _lib, _vertical as _v, plan as _plan
from .driver import BIAS_CORRECTION
from .kneedle import knee_index
from .optimal_interpolation import (apply_device, kneed_available, regularisation_factors,
                                    sweep_device, sweep_knee_apply_device)


class _DeviceScalar:
    """A one-element device tensor that reads as a host number when asked (int(), float(),
    ==): the month step itself never waits for it."""

    def __init__(self, t, kind):
        self.t, self.kind = t, kind

    def value(self):
        return self.kind(self.t.item())

    def __int__(self):
        return int(self.value())

    def __index__(self):
        return int(self.value())

    def __float__(self):
        return float(self.value())

    def __eq__(self, other):
        return self.value() == other

    def __repr__(self):
        return repr(self.value())


class _DeviceVector:
    def __init__(self, t):
        self.t = t

    def value(self):
        return _dev.to_host(self.t)

    def __array__(self, dtype=None, copy=None):
        a = self.value()
        return a if dtype is None else a.astype(dtype)


class _HostBlock(dict):
    """host_arrays(pin=True): field -> view into one page-locked block (`block`; `offsets` of the
    views; the first `head` bytes hold lon and lat, which the plan builders need first)."""

    def __init__(self):
        super().__init__()
        self.block, self.offsets, self.head = None, {}, 0


class _Granule:
    __slots__ = ("n_px", "nlev", "has_trop", "dev", "plan", "slot", "time", "host", "dev_block")


def finalize_and_oi(acc, n_cell, sensor, gas, error_ctm):
    """[10][n_cell] accumulator block -> monthly means (averaging.py:97-108), bias correction
    (driver.py:65-106), the 99-factor OI sweep, the knee and the update (driver.py:108-114,
    optimal_interpolation.py:6-52).  GOSAT runs OI on (aux2, aux1) = (mean model XCH4, mean
    satellite XCH4) instead of the columns (driver.py:113-114).  Returns device tensors."""
    L = _lib.lib()
    n = int(n_cell)
    means = [_dev.empty((n,)) for _ in range(5)]
    _lib.check(L.oisat_accum_finalize(acc.data_ptr(), n, *[m.data_ptr() for m in means],
                                      _dev.stream()))
    sat_vcd, sat_err, ctm_vcd, aux1, aux2 = means
    a, b = BIAS_CORRECTION.get((sensor, gas), (0.0, 1.0))
    xa, y = (aux2, aux1) if sensor == "GOSAT" else (ctm_vcd, sat_vcd)
    Sa, So = _dev.empty((n,)), _dev.empty((n,))
    _lib.check(L.oisat_oi_prepare(xa.data_ptr(), y.data_ptr(), sat_err.data_ptr(), n,
                                  a, b, float(error_ctm), Sa.data_ptr(), So.data_ptr(),
                                  _dev.stream()))
    factors = regularisation_factors(True)
    if kneed_available() or os.environ.get("OISAT_KNEE") == "host":
        ak_means = sweep_device(Sa, So, factors)   # one small D2H: 99 sums + 99 counts
        pick = knee_index(factors, ak_means)       # the reference's own kneed when importable
        xb, ak, inc, err = apply_device(xa, y, Sa, So, float(factors[pick]))
        extra = dict(knee_index=pick, ak_means=ak_means, factor=float(factors[pick]),
                     knee_source="kneed" if kneed_available() else "restated-host")
    else:
        # sweep -> knee -> update on the device, no host round trip inside the step
        xb, ak, inc, err, pick, factor, ak_means = sweep_knee_apply_device(xa, y, Sa, So, factors)
        extra = dict(knee_index=_DeviceScalar(pick, int), ak_means=_DeviceVector(ak_means),
                     factor=_DeviceScalar(factor, float), knee_source="restated-device")
    return dict(sat_averaged_vcd=sat_vcd, sat_averaged_error=sat_err, ctm_averaged_vcd=ctm_vcd,
                aux1=aux1, aux2=aux2, ctm_averaged_vcd_corrected=xb, ak_OI=ak,
                increment_OI=inc, error_OI=err, **extra)


_COPY_STREAMS = {}


def _copy_stream():
    """One side stream per device for host -> device copies that should not queue in front
    of kernels (add_day)."""
    t = _dev.torch()
    d = t.cuda.current_device()
    if d not in _COPY_STREAMS:
        _COPY_STREAMS[d] = t.cuda.Stream(device=d)
    return _COPY_STREAMS[d]


class MonthPipeline:
    def __init__(self, ctm_data, grid_size, flag_thresh, sensor="OMI", gas="NO2", error_ctm=50.0,
                 process_group=None, interpolator_type=1):
        _dev.require_cuda()
        if interpolator_type not in (1, 2, 4):  # interpolator.py:10-36; 3 (RBF) is not built
            raise Exception("other type of interpolation methods has not been implemented yet")
        self.interpolator_type = int(interpolator_type)
        self.ctm_data = ctm_data
        self.coords = {"Latitude": ctm_data[0].latitude, "Longitude": ctm_data[0].longitude}
        self.grid_size = float(grid_size)
        self.flag_thresh = float(flag_thresh)
        self.sensor, self.gas, self.error_ctm = sensor, gas, float(error_ctm)
        self.pg = process_group
        self.gplan = _plan.grid_plan(self.coords, grid_size)
        if not self.gplan.upscale:
            raise _lib.OisatError("the fused pipeline needs a model grid coarser than grid_size")
        self.n_cell = int(np.prod(self.gplan.out_shape))
        self.granules = []
        self._stamps, self._fracs = _v.ctm_clock(ctm_data)
        self._ctm_dev = None
        self._ctm_slots = []
        self._tables = None
        self._buf = None
        self.timings = {}

    # ------------------------------------------------------------------ inputs
    def upload_ctm(self):
        """Model fields as [n_used_slots][n_lev][n_cell] float32 device tensors.  Only the
        time slots the granules of THIS pipeline (this rank's share) refer to are uploaded
        and `gran_slot` indexes the compact block: a non-averaged month (8 slots a day, 44 GB
        for the three fields) never has to be resident as a whole (amf_recal.py:39-49 picks
        one slot per granule)."""
        used = sorted({g.slot for g in self.granules}) or [0]
        if self._ctm_dev is not None and set(used) <= set(self._ctm_slots):
            return self._ctm_dev
        t = _dev.torch()
        free = self.ctm_data[0].ctmtype == "FREE"
        per = 1 if free else int(np.asarray(self.ctm_data[0].pressure_mid).shape[0])

        def slab(name, slot):
            c = self.ctm_data[slot // per] if not free else self.ctm_data[slot]
            a = np.asarray(getattr(c, name))
            if a.dtype != np.float32:
                raise _lib.OisatError("model fields must be float32 as delivered by the readers")
            if a.ndim == 4:
                a = a[slot % per]
            return _dev.to_device(np.ascontiguousarray(a).reshape(1, a.shape[0], -1))

        def stack(name):
            parts = [slab(name, s) for s in used]
            return parts[0] if len(parts) == 1 else t.cat(parts, dim=0)

        self._ctm_dev = (stack("pressure_mid"), stack("gas_profile"), stack("delta_p"))
        self._ctm_slots = list(used)
        self._tables = None            # gran_slot indexes the compact block
        self._buf = None
        return self._ctm_dev

    def share_ctm(self, other):
        """Use the model block `other` already uploaded (one upload per month)."""
        self._ctm_dev, self._ctm_slots = other._ctm_dev, list(other._ctm_slots)

    def _slot_of(self, t_sat):
        k, day, hour = _v.closest_slot(self.ctm_data, self._stamps, self._fracs, t_sat)
        if self.ctm_data[0].ctmtype == "FREE":
            return day
        per = np.asarray(self.ctm_data[0].pressure_mid).shape[0]
        return day * per + hour

    FIELDS = ("lon", "lat", "vcd", "sigma", "amf", "qflag", "pmid", "sw", "trop")

    @staticmethod
    def host_arrays(sat, pin=False):
        """Reader record -> flat host tensors in the reader's own dtypes (pinned on
        request, so that uploads are asynchronous DMA copies)."""
        t = _dev.torch()
        for name in ("vcd", "uncertainty", "pressure_mid", "scattering_weights"):
            if np.asarray(getattr(sat, name)).dtype != np.float16:
                raise _lib.OisatError("fused path expects float16 %s (reader dtypes)" % name)
        has_trop = np.size(sat.tropopause) != 1
        src = {
            "lon": _plan.coord_array(sat.longitude_center),
            "lat": _plan.coord_array(sat.latitude_center),
            "vcd": sat.vcd, "sigma": sat.uncertainty,
            "amf": _dev.native_float(np.asarray(sat.amf)),
            "qflag": _dev.native_float(np.asarray(sat.quality_flag).squeeze()),
            "pmid": sat.pressure_mid, "sw": sat.scattering_weights,
            "trop": sat.tropopause if has_trop else None,
        }
        if not pin:
            return {k: (None if a is None else t.from_numpy(np.ascontiguousarray(a).reshape(-1)))
                    for k, a in src.items()}
        # page-locked: ONE block per granule with the arrays as views into it (256-byte
        # aligned, lon / lat first), so that add_day moves a granule with two copies
        # instead of nine (the ~100 small copy calls of a day were 2 ms of its host time)
        out = _HostBlock()
        arrays = {k: np.ascontiguousarray(a).reshape(-1) for k, a in src.items() if a is not None}
        off = 0
        for k, a in arrays.items():
            out.offsets[k] = off
            off = (off + a.nbytes + 255) // 256 * 256
            if k == "lat":
                out.head = off
        out.block = t.empty((off,), dtype=t.uint8, pin_memory=True)
        for k in src:
            if src[k] is None:
                out[k] = None
                continue
            a = arrays[k]
            view = out.block[out.offsets[k]:out.offsets[k] + a.nbytes].view(t.from_numpy(a[:0]).dtype)
            view.numpy()[...] = a
            out[k] = view
        return out

    def add_granule(self, sat, plan=None, host=None, pin=False):
        """Upload one reader record (before gridding) and attach its geometry plan
        (built here unless given).  Returns False when the granule is skipped like
        the reference would (Qhull failure or nothing on the grid,
        interpolator.py:152-155,165-167)."""
        g = _Granule()
        g.host = host if host is not None else self.host_arrays(sat, pin=pin)
        g.n_px = int(np.size(sat.latitude_center))
        g.nlev = np.shape(sat.pressure_mid)[0]
        g.has_trop = g.host["trop"] is not None
        dev = _dev.device()
        g.dev = {k: (None if h is None else h.to(dev, non_blocking=True))
                 for k, h in g.host.items()}
        if plan is None:
            lon, lat = np.asarray(sat.longitude_center), np.asarray(sat.latitude_center)
            if self.interpolator_type == 1:
                plan = _plan.granule_plan(lon, lat, self.gplan, radius=self.grid_size * 2.0,
                                          lonlat_dev=(g.dev["lon"], g.dev["lat"]), cache=False)
            elif self.interpolator_type == 4 or _plan.triangulable(lon, lat):
                plan = _plan.nearest_plan(lon, lat, self.gplan, radius=self.grid_size * 2.0,
                                          lonlat_dev=(g.dev["lon"], g.dev["lat"]))
        g.plan = plan
        g.slot = self._slot_of(sat.time)
        g.time = sat.time
        if g.plan is None or g.plan.n_cells == 0:
            return False
        self.granules.append(g)
        self._tables = None
        self._buf = None
        return True

    def add_day(self, sats, hosts=None, pin=False):
        """A batch of reader records (one day of orbits): every array is queued for
        upload first (asynchronous DMA when the host tensors are pinned), then the
        geometry plans of the whole batch are built -- triangulations on the host
        thread pool, the device part granule by granule -- while the copies are in
        flight.  Returns the number of granules kept."""
        import os
        import time as _time
        trace = os.environ.get("OISAT_PLAN_TRACE") == "1"
        t_0 = _time.perf_counter()
        dev = _dev.device()
        t = _dev.torch()
        lons = [np.asarray(s.longitude_center) for s in sats]
        lats = [np.asarray(s.latitude_center) for s in sats]
        # the host share of the plans starts first: it needs nothing from the device
        futures = (_plan.submit_triangulations(lons, lats, self.gplan)
                   if self.interpolator_type == 1 else None)
        staged = []
        for i, sat in enumerate(sats):
            g = _Granule()
            g.host = hosts[i] if hosts is not None else self.host_arrays(sat, pin=pin)
            g.n_px = int(np.size(sat.latitude_center))
            g.nlev = np.shape(sat.pressure_mid)[0]
            g.has_trop = g.host["trop"] is not None
            g.dev = {}
            if isinstance(g.host, _HostBlock):
                # one device block per granule, the fields are views; lon / lat (its head) now
                g.dev_block = t.empty(g.host.block.shape, dtype=t.uint8, device=dev)
                g.dev_block[:g.host.head].copy_(g.host.block[:g.host.head], non_blocking=True)
                for k, h in g.host.items():
                    g.dev[k] = None if h is None else (
                        g.dev_block[g.host.offsets[k]:g.host.offsets[k] + h.numel() * h.element_size()]
                        .view(h.dtype))
            else:
                for k in ("lon", "lat"):          # the plan builders need these first
                    g.dev[k] = g.host[k].to(dev, non_blocking=True)
            staged.append(g)
        # the bulk of the reader arrays travels on a copy stream of its own: the plan kernels
        # queued below need only lon / lat and would otherwise sit behind 24 MB per granule
        main = t.cuda.current_stream()
        blocks = all(isinstance(g.host, _HostBlock) for g in staged)
        pinned = blocks or all(h is None or h.is_pinned() for g in staged for h in g.host.values())
        side = _copy_stream() if pinned else None
        for g in staged:
            for k, h in g.host.items():
                if k not in g.dev:
                    g.dev[k] = None if h is None else (
                        t.empty(h.shape, dtype=h.dtype, device=dev) if side is not None
                        else h.to(dev, non_blocking=True))
        copied = None
        if side is not None:
            side.wait_stream(main)          # the destinations were allocated on `main`
            with t.cuda.stream(side):
                for g in staged:
                    if isinstance(g.host, _HostBlock):
                        g.dev_block[g.host.head:].copy_(g.host.block[g.host.head:], non_blocking=True)
                        continue
                    for k, h in g.host.items():
                        if k not in ("lon", "lat") and h is not None:
                            g.dev[k].copy_(h, non_blocking=True)
                copied = t.cuda.Event()
                copied.record()
        lonlat = [(g.dev["lon"], g.dev["lat"]) for g in staged]
        radius = self.grid_size * 2.0
        t_1 = _time.perf_counter()
        if self.interpolator_type == 1:
            plans = _plan.granule_plans(lons, lats, self.gplan, radius, lonlat_dev=lonlat,
                                        futures=futures)
        else:
            plans = [(_plan.nearest_plan(lons[i], lats[i], self.gplan, radius, lonlat_dev=lonlat[i])
                      if self.interpolator_type == 4 or _plan.triangulable(lons[i], lats[i])
                      else None) for i in range(len(sats))]
        if trace:
            import sys
            t_2 = _time.perf_counter()
            done = copied.query() if copied is not None else True
            if copied is not None:
                copied.synchronize()
            print("add_day trace: staging %.1f ms, plans %.1f ms; bulk copies %s when the plans were "
                  "(waited %.1f ms more)" % ((t_1 - t_0) * 1e3, (t_2 - t_1) * 1e3,
                                            "done" if done else "NOT done", (_time.perf_counter() - t_2) * 1e3),
                  file=sys.stderr, flush=True)
        if copied is not None:
            main.wait_event(copied)          # whatever is queued from here on sees the arrays
        kept = 0
        for g, sat, plan in zip(staged, sats, plans):
            g.plan = plan
            g.slot = self._slot_of(sat.time)
            g.time = sat.time
            if plan is None or plan.n_cells == 0:
                continue
            self.granules.append(g)
            kept += 1
        self._tables = None
        self._buf = None
        return kept

    def refresh_inputs(self):
        """Host -> device copy of every granule's reader arrays (the per-step
        upload of the end-to-end measurement).  Returns the bytes copied."""
        n = 0
        for g in self.granules:
            for k, h in g.host.items():
                if h is not None:
                    g.dev[k].copy_(h, non_blocking=True)
                    n += h.numel() * h.element_size()
        return n

    def input_bytes(self):
        """Bytes of reader-typed pixel data resident on the device."""
        tot = 0
        for g in self.granules:
            tot += sum(t.numel() * t.element_size() for t in g.dev.values() if t is not None)
        return tot

    # ------------------------------------------------------------ pair tables
    def build_tables(self):
        """Concatenate the granule plans into the tile / pair / segment tables the
        fused kernels consume (all geometry, no pixel values)."""
        if self._tables is not None:
            return self._tables
        G = self.granules
        if not G:
            raise _lib.OisatError("no granule on the grid")
        self.upload_ctm()
        slot_index = {s: i for i, s in enumerate(self._ctm_slots)}
        nlev = {g.nlev for g in G}
        trop = {g.has_trop for g in G}
        if len(nlev) != 1 or len(trop) != 1:
            raise _lib.OisatError("granules of one month must share the level layout")
        S = 3 * self.gplan.nwin
        cells = np.concatenate([g.plan.cells for g in G]).astype(np.int64)
        gran = np.concatenate([np.full(g.plan.n_cells, i, np.int32) for i, g in enumerate(G)])
        # stencil entries stay on the device (pair-major); only `cells` is host data
        t = _dev.torch()
        pairs = [g.plan.dev_pairs() for g in G]
        vert = t.cat([p[0] for p in pairs]) if len(pairs) > 1 else pairs[0][0]   # (n_pairs, S)
        w = t.cat([p[1] for p in pairs]) if len(pairs) > 1 else pairs[0][1]
        n_pairs = cells.size
        px0 = np.concatenate(([0], np.cumsum([g.n_px for g in G])[:-1])).astype(np.int64)
        host = dict(n_pairs=n_pairs, n_tiles=0, S=S, px0=px0,
                    total_px=int(sum(g.n_px for g in G)), cells=cells, gran=gran)
        d = _dev.to_device
        pair_cell = d(cells.astype(np.int32))
        # segments: pairs of each model cell in granule order, built on the device
        # (oisat_segment_tables) -- the host only concatenated the cell lists
        seg_start = _dev.empty((self.n_cell + 1,), "int64")
        seg_pair = _dev.empty((n_pairs,), "int64")
        seg_work = _dev.empty((self.n_cell,), "int32")
        _lib.check(_lib.lib().oisat_segment_tables(pair_cell.data_ptr(), n_pairs, self.n_cell,
                                                   seg_start.data_ptr(), seg_pair.data_ptr(),
                                                   seg_work.data_ptr(), _dev.stream()))
        dev = dict(vert=vert, w=w, seg_start=seg_start, seg_pair=seg_pair, gran_px0=d(px0),
                   pair_gran=d(gran), pair_cell=pair_cell,
                   gran_slot=d(np.array([slot_index[g.slot] for g in G], np.int32)))
        self._tables = (host, dev)
        return self._tables

    def _tile_tables(self):
        """Tile table of the single-kernel form (oisat_fused_amf): runs of pairs of one
        granule inside one 32-aligned cell segment.  The tile and split forms do not read it."""
        host, dev = self.build_tables()
        if "tile_pair0" in dev:
            return
        cells, gran = host["cells"], host["gran"]
        tid = gran.astype(np.int64) * ((self.n_cell + 31) // 32 + 1) + cells // 32
        starts = np.concatenate(([0], np.flatnonzero(np.diff(tid)) + 1))
        bits = (np.uint32(1) << (cells % 32).astype(np.uint32))
        tile_mask = np.bitwise_or.reduceat(bits, starts).astype(np.uint32)
        d = _dev.to_device
        dev.update(tile_pair0=d(starts.astype(np.int64)), tile_gran=d(gran[starts].astype(np.int32)),
                   tile_cell0=d(((cells[starts] // 32) * 32).astype(np.int32)),
                   tile_mask=d(tile_mask.view(np.int32)))
        host["n_tiles"] = len(starts)

    def plan_bytes(self):
        _, dev = self.build_tables()
        return sum(t.numel() * t.element_size() for t in dev.values())

    # ---------------------------------------------------------------- kernels
    def allocate(self):
        host, _ = self.build_tables()
        L = _lib.lib()
        g0 = self.granules[0]
        R = int(L.oisat_pack_record_halfs(g0.nlev, int(g0.has_trop)))
        t = _dev.torch()
        pm, pr, dp = self.upload_ctm()
        # device table of the batch pack launch
        items = (_lib.PackItem * len(self.granules))()
        block0 = 0
        blocks_of = []
        for i, (g, p0) in enumerate(zip(self.granules, host["px0"])):
            d = g.dev
            items[i].sw, items[i].p_mid = d["sw"].data_ptr(), d["pmid"].data_ptr()
            items[i].vcd, items[i].sigma = d["vcd"].data_ptr(), d["sigma"].data_ptr()
            items[i].trop = _dev.ptr(d["trop"])
            items[i].qflag = d["qflag"].data_ptr()
            items[i].amf = d["amf"].data_ptr()
            items[i].n_px, items[i].px0, items[i].block0 = g.n_px, int(p0), block0
            blocks_of.append(int(L.oisat_pack_blocks(g.n_px)))
            block0 += blocks_of[-1]
        raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
        # OISAT_GUARD=1 (tests): every output buffer sits between two canary zones that
        # check_guards() inspects after a run -- the out-of-bounds-write check of this path
        self._guards = []
        guard = os.environ.get("OISAT_GUARD") == "1"

        def out(shape, dtype="float64", zero=False):
            if not guard:
                return (_dev.zeros if zero else _dev.empty)(shape, dtype)
            n = int(np.prod(shape))
            pad = 4096
            big = _dev.full((n + 2 * pad,), 77, dtype)
            self._guards.append((big, pad, n))
            view = big[pad:pad + n]
            if zero:
                view.zero_()
            return view.reshape(shape)

        self._buf = dict(
            records=out((host["total_px"], R), "float16"),
            amf_masked=out((host["total_px"],)),
            px_bad=out((host["total_px"],), "uint8"),
            alive_pairs=out((host["n_pairs"],), "int32"), n_alive=out((1,), "int64"),
            staged=out((5, host["n_pairs"])),
            acc=out((10, self.n_cell), zero=True),
            ctm_logp=out(tuple(pm.shape), "float32"), ctm_pcol=out(tuple(pm.shape), "float32"),
            rows=(out(((host["n_pairs"] + 15) // 16,
                       int(L.oisat_rows_per_pair(g0.nlev, int(g0.has_trop))), 16))
                  if self.split else None),
            pack_items=_dev.to_device(raw), pack_blocks=block0,
            pack_block_item=_dev.to_device(np.repeat(np.arange(len(blocks_of), dtype=np.int32),
                                                     blocks_of)),
        )
        return self._buf

    def check_guards(self):
        """True when no kernel wrote outside its output buffers (needs OISAT_GUARD=1)."""
        _dev.torch().cuda.synchronize()
        for big, pad, n in self._guards:
            if not bool((big[:pad] == 77).all()) or not bool((big[pad + n:] == 77).all()):
                return False
        return bool(self._guards)

    def run_prepare(self):
        """Derived model fields (float32 log pressure, partial column)."""
        L = _lib.lib()
        pm, pr, dp = self.upload_ctm()
        buf = self._buf
        _lib.check(L.oisat_ctm_prepare(pm.data_ptr(), pr.data_ptr(), dp.data_ptr(), pm.numel(),
                                       buf["ctm_logp"].data_ptr(), buf["ctm_pcol"].data_ptr(),
                                       _dev.stream()))

    def run_pack(self):
        """Quality mask + pixel-major packing of every granule, one launch."""
        L = _lib.lib()
        buf = self._buf
        g0 = self.granules[0]
        _lib.check(L.oisat_pack_batch_masked(
            buf["pack_items"].data_ptr(), len(self.granules), buf["pack_blocks"],
            buf["pack_block_item"].data_ptr(), g0.nlev, int(g0.has_trop),
            _dev.dtype_code(g0.dev["qflag"]), self.flag_thresh, _dev.dtype_code(g0.dev["amf"]),
            buf["records"].data_ptr(), buf["amf_masked"].data_ptr(), buf["px_bad"].data_ptr(),
            int(self.fused_form == "tile"), _dev.stream()))

    def fused_args(self):
        host, dev = self.build_tables()
        buf = self._buf
        pm, pr, dp = self.upload_ctm()
        g0 = self.granules[0]
        nwin = self.gplan.nwin
        a = _lib.FusedArgs()
        if self.fused_form == "single":
            self._tile_tables()
            a.n_tiles = host["n_tiles"]
            a.tile_granule = dev["tile_gran"].data_ptr()
            a.tile_cell0 = dev["tile_cell0"].data_ptr()
            a.tile_pair0 = dev["tile_pair0"].data_ptr()
            a.tile_mask = dev["tile_mask"].data_ptr()
        a.n_pairs = host["n_pairs"]
        a.nwin = nwin
        a.vert = dev["vert"].data_ptr()
        a.w = dev["w"].data_ptr()
        a.box_weight = 1.0 / nwin
        a.box_weight_err = 1.0 / (nwin * nwin)
        a.n_granules = len(self.granules)
        a.gran_record0 = dev["gran_px0"].data_ptr()
        a.gran_px0 = dev["gran_px0"].data_ptr()
        a.gran_slot = dev["gran_slot"].data_ptr()
        a.n_records = host["total_px"]
        a.records = buf["records"].data_ptr()
        a.amf_masked = buf["amf_masked"].data_ptr()
        a.n_sat_lev = g0.nlev
        a.has_trop = int(g0.has_trop)
        a.ctm_pmid = pm.data_ptr()
        a.ctm_logp, a.ctm_pcol = buf["ctm_logp"].data_ptr(), buf["ctm_pcol"].data_ptr()
        a.n_ctm_lev = pm.shape[1]
        a.n_cell = self.n_cell
        a.staged = buf["staged"].data_ptr()
        a.pair_granule = dev["pair_gran"].data_ptr()
        a.pair_cell = dev["pair_cell"].data_ptr()
        if "pair_rec0" not in dev:     # per-pair copies of the per-granule facts (tile form)
            if int(pm.shape[0]) * pm.shape[1] * self.n_cell >= 2 ** 31:
                raise _lib.OisatError("model fields too large for 32-bit element offsets")
            dev["pair_rec0"] = _dev.empty((host["n_pairs"],), "int64")
            dev["pair_ctm_off"] = _dev.empty((host["n_pairs"],), "int32")
            _lib.check(_lib.lib().oisat_pair_tables(
                host["n_pairs"], dev["pair_gran"].data_ptr(), dev["pair_cell"].data_ptr(),
                dev["gran_px0"].data_ptr(), dev["gran_slot"].data_ptr(), int(pm.shape[1]),
                self.n_cell, int(pm.shape[0]), dev["pair_rec0"].data_ptr(),
                dev["pair_ctm_off"].data_ptr(), _dev.stream()))
        a.pair_record0 = dev["pair_rec0"].data_ptr()
        a.pair_ctm_off = dev["pair_ctm_off"].data_ptr()
        if self.fused_form == "tile":
            a.alive_pairs = buf["alive_pairs"].data_ptr()
            a.n_alive = buf["n_alive"].data_ptr()
        return a

    @property
    def fused_form(self):
        """Which form of the fused step runs: "tile" (default: one launch, 16 pairs per
        block, gridded columns kept in shared memory), "split" (two launches joined by a
        row buffer in HBM) -- both need records of fewer than 16 chunks -- or "single"
        (half warp per pair end to end, any record width).  OISAT_FUSED overrides."""
        import os
        g0 = self.granules[0]
        halfs = int(_lib.lib().oisat_pack_record_halfs(g0.nlev, int(g0.has_trop)))
        want = os.environ.get("OISAT_FUSED", "tile")
        if want not in ("tile", "split", "single"):
            raise _lib.OisatError("OISAT_FUSED must be tile, split or single")
        if halfs // 8 >= 16:
            return "single"
        if want == "tile" and g0.nlev > 62:
            return "split"
        return want

    @property
    def split(self):
        return self.fused_form == "split"

    def run_alive(self):
        """Tile form only: pairs with a masked stencil pixel are NaN in every field
        (interpolator.py:126-128); they get their NaNs here and the fused kernel runs over the
        compact list of the others."""
        if self.fused_form != "tile":
            return
        a = self.fused_args()
        _lib.check(_lib.lib().oisat_pair_alive(
            a.n_pairs, a.nwin, a.vert, a.w, a.pair_record0, a.pair_granule, a.gran_px0,
            self._buf["px_bad"].data_ptr(), a.amf_masked, a.box_weight, a.staged,
            a.alive_pairs, a.n_alive, _dev.stream()))

    def run_fused(self):
        L = _lib.lib()
        a = self.fused_args()
        form = self.fused_form
        if form == "tile":
            _lib.check(L.oisat_fused_amf_tile(C.byref(a), _dev.stream()))
        elif form == "split":
            _lib.check(L.oisat_fused_amf_split(C.byref(a), self._buf["rows"].data_ptr(),
                                               _dev.stream()))
        else:
            _lib.check(L.oisat_fused_amf(C.byref(a), _dev.stream()))

    def run_accumulate(self):
        L = _lib.lib()
        host, dev = self.build_tables()
        buf = self._buf
        buf["acc"].zero_()
        _lib.check(L.oisat_accum_pairs(buf["acc"].data_ptr(), self.n_cell,
                                       dev["seg_start"].data_ptr(), dev["seg_pair"].data_ptr(),
                                       buf["staged"].data_ptr(), host["n_pairs"], _dev.stream()))
        if self.pg is not None:
            from .sharding import merge_accumulators
            merge_accumulators(buf["acc"], self.pg)   # the path's only collective (NCCL)

    def run_oi(self):
        """Means, bias correction, OI sweep, knee, apply.  Returns device tensors."""
        return finalize_and_oi(self._buf["acc"], self.n_cell, self.sensor, self.gas, self.error_ctm)

    def run(self, marks=None):
        """pack -> fused -> accumulate -> OI on the current stream.  `marks`
        (a list) receives (phase, cuda event) pairs recorded on that stream."""
        if not self.granules:
            # a rank whose share of the month is empty (fewer days than ranks, or every granule
            # skipped): nothing to grid, but it still owns a zero accumulator block and MUST
            # join the all-reduce, or the other ranks wait for it forever
            self._buf = dict(acc=_dev.zeros((10, self.n_cell)))
            if self.pg is not None:
                from .sharding import merge_accumulators
                merge_accumulators(self._buf["acc"], self.pg)
            return self.run_oi()
        if getattr(self, "_buf", None) is None:
            self.allocate()

        def mark(name):
            if marks is not None:
                e = _dev.torch().cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        self.run_prepare()
        mark("prepare")
        self.run_pack()
        mark("pack")
        self.run_alive()
        mark("alive")
        self.run_fused()
        mark("fused")
        self.run_accumulate()
        mark("accumulate")
        res = self.run_oi()
        mark("oi")
        return res

    def results_to_host(self, res):
        """Device results -> numpy.  The fields are copied into page-locked buffers, all copies
        in flight at once and ONE wait (a `.cpu()` per field goes through pageable staging and
        waits each time: 2 ms for the nine fields of a month, against 0.4 ms)."""
        t = _dev.torch()
        shape = self.gplan.out_shape
        out, late = {}, []
        for k, v in res.items():
            if isinstance(v, (_DeviceScalar, _DeviceVector)):
                late.append((k, v))
            elif hasattr(v, "data_ptr"):
                h = t.empty(v.shape, dtype=v.dtype, pin_memory=True)
                h.copy_(v, non_blocking=True)
                out[k] = h
            else:
                out[k] = v
        t.cuda.current_stream().synchronize()
        for k, v in list(out.items()):
            if hasattr(v, "data_ptr"):
                out[k] = v.numpy().reshape(shape)
        for k, v in late:
            out[k] = v.value()
        return out

    def output_fields(self, res):
        """The nine float32 variables driver.write_to_nc stores (driver.py:180-222), from the
        device results of run(): one K8 launch and one float32 device -> host copy."""
        from .driver import oisatgmi
        n = self.n_cell
        out = _dev.empty((9, n), "float32")
        keys = ("sat_averaged_vcd", "ctm_averaged_vcd", "ctm_averaged_vcd_corrected",
                "sat_averaged_error", "ak_OI", "error_OI", "aux1", "aux2")
        _lib.check(_lib.lib().oisat_output_fields(n, *[res[k].data_ptr() for k in keys],
                                                  out.data_ptr(), _dev.stream()))
        host = _dev.to_host(out)
        shape = self.gplan.out_shape
        return {name: host[k].reshape(shape) for k, name in enumerate(oisatgmi.OUTPUT_NAMES)}

    def n_pixels(self):
        return int(sum(g.n_px for g in self.granules))
