"""Fused month pipeline: the whole hot path of one (year, month) bin for float16
`satellite_amf` products (OMI NO2 / HCHO, TROPOMI NO2) without ever
materialising a gridded granule.

    reader records (host) --upload--> HBM
    K0  oisat_distmask        \\  geometry plan per granule (host Qhull + walk,
    plan.granule_plan         /   oisatgmi_b200/plan.py), concatenated into pair tables
    oisat_quality_mask, oisat_pack_granule      pixel-major float16 records
    oisat_fused_amf           gather-interpolate + AMF recalculation per (granule, cell)
    oisat_accum_pairs         ordered segmented reduction -> [10][n_cell] sums / counts
    (torch.distributed all_reduce of the accumulator block when sharded, section 8e)
    oisat_accum_finalize, oisat_oi_prepare, oisat_oi_sweep, knee (host), oisat_oi_apply

It computes exactly what `interpolator` -> `amf_recal` -> `averaging` ->
`bias_correct` -> `oi` compute through the drop-in modules (same device
functions), and is what `bench.py` times.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib, _vertical as _v, plan as _plan
from .driver import BIAS_CORRECTION
from .kneedle import knee_index
from .optimal_interpolation import apply_device, regularisation_factors, sweep_device


class _Granule:
    __slots__ = ("n_px", "nlev", "has_trop", "dev", "plan", "slot", "time", "host")


class MonthPipeline:
    def __init__(self, ctm_data, grid_size, flag_thresh, sensor="OMI", gas="NO2", error_ctm=50.0,
                 process_group=None):
        _dev.require_cuda()
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
        self._tables = None
        self.timings = {}

    # ------------------------------------------------------------------ inputs
    def upload_ctm(self):
        """Model fields as [n_slots][n_lev][n_cell] float32 device tensors."""
        if self._ctm_dev is not None:
            return self._ctm_dev
        t = _dev.torch()

        def stack(name):
            parts = []
            for c in self.ctm_data:
                a = np.asarray(getattr(c, name))
                if a.dtype != np.float32:
                    raise _lib.OisatError("model fields must be float32 as delivered by the readers")
                if a.ndim == 3:
                    a = a[None]
                parts.append(_dev.to_device(a.reshape(a.shape[0], a.shape[1], -1)))
            return parts[0] if len(parts) == 1 else t.cat(parts, dim=0)

        self._ctm_dev = (stack("pressure_mid"), stack("gas_profile"), stack("delta_p"))
        return self._ctm_dev

    def _slot_of(self, t_sat):
        k, day, hour = _v.closest_slot(self.ctm_data, self._stamps, self._fracs, t_sat)
        if self.ctm_data[0].ctmtype == "FREE":
            return day
        per = np.asarray(self.ctm_data[0].pressure_mid).shape[0]
        return day * per + hour

    def add_granule(self, sat, pin=False):
        """Upload one reader record (before gridding) and build its geometry plan.
        Returns False when the granule is skipped like the reference would
        (Qhull failure or nothing on the grid, interpolator.py:152-155,165-167)."""
        for name in ("vcd", "uncertainty", "pressure_mid", "scattering_weights"):
            if np.asarray(getattr(sat, name)).dtype != np.float16:
                raise _lib.OisatError("fused path expects float16 %s (reader dtypes)" % name)
        g = _Granule()
        lat = np.asarray(sat.latitude_center)
        lon = np.asarray(sat.longitude_center)
        g.n_px = lat.size
        g.nlev = np.shape(sat.pressure_mid)[0]
        g.has_trop = np.size(sat.tropopause) != 1
        up = lambda a: _dev.to_device(np.ascontiguousarray(a).reshape(-1), pin=pin)  # noqa: E731
        g.dev = {
            "lon": up(_plan.coord_array(lon)), "lat": up(_plan.coord_array(lat)),
            "vcd": up(sat.vcd), "sigma": up(sat.uncertainty),
            "amf": up(_dev.native_float(np.asarray(sat.amf))),
            "qflag": up(_dev.native_float(np.asarray(sat.quality_flag).squeeze())),
            "pmid": up(sat.pressure_mid), "sw": up(sat.scattering_weights),
            "trop": up(sat.tropopause) if g.has_trop else None,
        }
        g.plan = _plan.granule_plan(lon, lat, self.gplan, radius=self.grid_size * 2.0,
                                    lonlat_dev=(g.dev["lon"], g.dev["lat"]), cache=False)
        g.slot = self._slot_of(sat.time)
        g.time = sat.time
        if g.plan is None or g.plan.n_cells == 0:
            return False
        self.granules.append(g)
        self._tables = None
        return True

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
        nlev = {g.nlev for g in G}
        trop = {g.has_trop for g in G}
        if len(nlev) != 1 or len(trop) != 1:
            raise _lib.OisatError("granules of one month must share the level layout")
        S = 3 * self.gplan.nwin
        cells = np.concatenate([g.plan.cells for g in G]).astype(np.int64)
        gran = np.concatenate([np.full(g.plan.n_cells, i, np.int32) for i, g in enumerate(G)])
        vert = np.concatenate([g.plan.vert.T for g in G]).astype(np.int32)   # (n_pairs, S)
        w = np.concatenate([g.plan.w.T for g in G])
        n_pairs = cells.size
        # tiles: runs of pairs of one granule inside one 32-aligned cell segment
        tid = gran.astype(np.int64) * ((self.n_cell + 31) // 32 + 1) + cells // 32
        starts = np.concatenate(([0], np.flatnonzero(np.diff(tid)) + 1))
        tile_pair0 = starts.astype(np.int64)
        tile_gran = gran[starts].astype(np.int32)
        tile_cell0 = ((cells[starts] // 32) * 32).astype(np.int32)
        bits = (np.uint32(1) << (cells % 32).astype(np.uint32))
        tile_mask = np.bitwise_or.reduceat(bits, starts).astype(np.uint32)
        # segments: pairs of each model cell in granule order (stable sort by cell)
        order = np.argsort(cells, kind="stable").astype(np.int64)
        seg_start = np.zeros(self.n_cell + 1, np.int64)
        np.cumsum(np.bincount(cells, minlength=self.n_cell), out=seg_start[1:])
        px0 = np.concatenate(([0], np.cumsum([g.n_px for g in G])[:-1])).astype(np.int64)
        host = dict(n_pairs=n_pairs, n_tiles=len(starts), S=S, px0=px0,
                    total_px=int(sum(g.n_px for g in G)))
        d = _dev.to_device
        dev = dict(vert=d(np.ascontiguousarray(vert)), w=d(np.ascontiguousarray(w)),
                   tile_pair0=d(tile_pair0), tile_gran=d(tile_gran), tile_cell0=d(tile_cell0),
                   tile_mask=d(tile_mask.view(np.int32)), seg_start=d(seg_start),
                   seg_pair=d(order), gran_px0=d(px0),
                   gran_slot=d(np.array([g.slot for g in G], np.int32)))
        self._tables = (host, dev)
        return self._tables

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
        self._buf = dict(
            records=_dev.empty((host["total_px"], R), "float16"),
            good=_dev.empty((host["total_px"],), "uint8"),
            amf=t.cat([g.dev["amf"] for g in self.granules]),
            staged=_dev.empty((5, host["n_pairs"])),
            acc=_dev.zeros((10, self.n_cell)),
        )
        return self._buf

    def run_pack(self):
        """Quality mask + pixel-major packing of every granule."""
        L = _lib.lib()
        host, _ = self.build_tables()
        buf = self._buf
        s = _dev.stream()
        R = buf["records"].shape[1]
        for g, p0 in zip(self.granules, host["px0"]):
            d = g.dev
            _lib.check(L.oisat_quality_mask(d["qflag"].data_ptr(), _dev.dtype_code(d["qflag"]),
                                            g.n_px, self.flag_thresh,
                                            buf["good"].data_ptr() + int(p0), s))
            _lib.check(L.oisat_pack_granule(d["sw"].data_ptr(), d["pmid"].data_ptr(), g.nlev,
                                            d["vcd"].data_ptr(), d["sigma"].data_ptr(),
                                            _dev.ptr(d["trop"]), g.n_px,
                                            buf["records"].data_ptr() + int(p0) * R * 2, s))

    def fused_args(self):
        host, dev = self.build_tables()
        buf = self._buf
        pm, pr, dp = self.upload_ctm()
        g0 = self.granules[0]
        nwin = self.gplan.nwin
        a = _lib.FusedArgs()
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
        a.records = buf["records"].data_ptr()
        a.good = buf["good"].data_ptr()
        a.amf = buf["amf"].data_ptr()
        a.amf_dtype = _dev.dtype_code(buf["amf"])
        a.n_sat_lev = g0.nlev
        a.has_trop = int(g0.has_trop)
        a.ctm_pmid, a.ctm_prof, a.ctm_dp = pm.data_ptr(), pr.data_ptr(), dp.data_ptr()
        a.n_ctm_lev = pm.shape[1]
        a.n_cell = self.n_cell
        a.staged = buf["staged"].data_ptr()
        return a

    def run_fused(self):
        L = _lib.lib()
        a = self.fused_args()
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
            import torch.distributed as dist
            dist.all_reduce(buf["acc"], op=dist.ReduceOp.SUM, group=self.pg)

    def run_oi(self):
        """Means, bias correction, OI sweep, knee, apply.  Returns device tensors."""
        L = _lib.lib()
        buf = self._buf
        n = self.n_cell
        means = [_dev.empty((n,)) for _ in range(5)]
        _lib.check(L.oisat_accum_finalize(buf["acc"].data_ptr(), n,
                                          *[m.data_ptr() for m in means], _dev.stream()))
        sat_vcd, sat_err, ctm_vcd, aux1, aux2 = means
        a, b = BIAS_CORRECTION.get((self.sensor, self.gas), (0.0, 1.0))
        Sa, So = _dev.empty((n,)), _dev.empty((n,))
        _lib.check(L.oisat_oi_prepare(ctm_vcd.data_ptr(), sat_vcd.data_ptr(), sat_err.data_ptr(), n,
                                      a, b, self.error_ctm, Sa.data_ptr(), So.data_ptr(),
                                      _dev.stream()))
        factors = regularisation_factors(True)
        ak_means = sweep_device(Sa, So, factors)   # one small D2H: 99 sums + 99 counts
        pick = knee_index(factors, ak_means)
        xb, ak, inc, err = apply_device(ctm_vcd, sat_vcd, Sa, So, float(factors[pick]))
        return dict(sat_averaged_vcd=sat_vcd, sat_averaged_error=sat_err, ctm_averaged_vcd=ctm_vcd,
                    aux1=aux1, aux2=aux2, ctm_averaged_vcd_corrected=xb, ak_OI=ak,
                    increment_OI=inc, error_OI=err, knee_index=pick, ak_means=ak_means,
                    factor=float(factors[pick]))

    def run(self):
        """pack -> fused -> accumulate -> OI on the current stream."""
        if getattr(self, "_buf", None) is None:
            self.allocate()
        self.run_pack()
        self.run_fused()
        self.run_accumulate()
        return self.run_oi()

    def results_to_host(self, res):
        shape = self.gplan.out_shape
        out = {}
        for k, v in res.items():
            out[k] = _dev.to_host(v).reshape(shape) if hasattr(v, "data_ptr") else v
        return out

    def n_pixels(self):
        return int(sum(g.n_px for g in self.granules))
