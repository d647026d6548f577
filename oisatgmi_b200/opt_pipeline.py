"""Fused month pipeline for `satellite_opt` products -- MOPITT CO (BASELINE configs[2]) and
GOSAT XCH4 (configs[3]) -- with every intermediate resident on the device:

    reader records (host) --upload once--> HBM
    [GOSAT] gap filling        filler_gosat.py:87-201   K0/K1 plan of the soundings + K2,
                                                        nearest-sounding quality field
    gridding                   interpolator.py:100-291  K2 on the (cached) lattice plan
    model -> satellite mesh    ak_conv_mopitt.py:79-110 K6, once per model day
    AK convolution             ak_conv_mopitt.py:118-138 / ak_conv_gosat.py:118-141   K3
    temporal accumulation      averaging.py:64-108      K4, granules in list order
    (one all-reduce of the accumulator block when the month is sharded over ranks)
    means, OI                  driver.py:108-114, optimal_interpolation.py:6-52       K5
                               (GOSAT: OI on aux2 / aux1)

It computes what `interpolator` -> `ak_conv_*` -> `averaging` -> `bias_correct` -> `oi` compute
through the drop-in modules (same device functions) without their host round trips: the
drop-ins bring every gridded field back to numpy per granule, as the reference's signatures
demand; here only the nine monthly fields leave the device.  A granule whose gridded column is
entirely NaN is skipped by the reference (interpolator.py:165-167); here it simply adds nothing.
"""
from __future__ import annotations

import numpy as np

from . import _dev, _lib, _vertical as _v, plan as _plan
from .ak_conv_mopitt import model_fields
from .interpolator import _FieldSpec, apply_plan
from .pipeline import finalize_and_oi

__all__ = ["OptMonthPipeline"]


class _Granule:
    __slots__ = ("time", "n_px", "nlev", "dev", "lon", "lat", "plan", "fill_plan", "near_plan")


def _dev_flat(a):
    a = _dev.native_float(np.asarray(a))
    return _dev.to_device(np.ascontiguousarray(a).reshape(-1))


class OptMonthPipeline:
    FILL_GRID = 1.0      # reader.py:1266: filler_gosatxch4(1.0, ...)

    def __init__(self, ctm_data, grid_size, flag_thresh, sensor, gas=None, error_ctm=50.0,
                 process_group=None):
        _dev.require_cuda()
        if sensor not in ("MOPITT", "GOSAT"):
            raise _lib.OisatError("OptMonthPipeline serves MOPITT and GOSAT")
        self.ctm_data = ctm_data
        self.coords = {"Latitude": ctm_data[0].latitude, "Longitude": ctm_data[0].longitude}
        self.grid_size, self.flag_thresh = float(grid_size), float(flag_thresh)
        self.sensor = sensor
        self.gas = gas or {"MOPITT": "CO", "GOSAT": "CH4"}[sensor]
        self.error_ctm = float(error_ctm)
        self.pg = process_group
        self.gplan = _plan.grid_plan(self.coords, grid_size)
        self.n_cell = int(np.prod(self.gplan.out_shape))
        self.granules = []
        self._stamps, _ = _v.ctm_clock(ctm_data)
        self._model_on_mesh = {}       # model day -> (p_mid, profile, third, mode) on the device
        self._fill = None              # GOSAT: (grid plan of the filler mesh, X, Y)
        self._resample = None          # K6 geometry (model grid -> output mesh)
        self._acc = None
        self._program = None           # recorded launches of the granule loop (see run)
        self._program_keep = None
        self.n_skipped = 0

    # ------------------------------------------------------------------ inputs
    def _filler_mesh(self):
        if self._fill is None:
            gs = self.FILL_GRID
            lon_axis = np.arange(-180.0, 180.0 + gs, gs).astype("float16")
            lat_axis = np.arange(-90.0, 90.0 + gs, gs).astype("float16")
            fx, fy = np.meshgrid(np.arange(-180.0, 181.0, 0.1).astype("float16"),
                                 np.arange(-90.0, 91.0, 0.1).astype("float16"))
            gpl = _plan.grid_plan({"Latitude": fy, "Longitude": fx}, gs, mesh=(lon_axis, lat_axis))
            X, Y = np.meshgrid(lon_axis, lat_axis)
            self._fill = (gpl, X, Y)
        return self._fill

    def add_granule(self, sat):
        """Upload one reader record (MOPITT L3 lattice / GOSAT soundings, before gridding) and
        attach its geometry plan(s).  False = skipped like the reference would (Qhull failure
        or nothing on the grid, interpolator.py:152-155, filler_gosat.py:139-142)."""
        g = _Granule()
        g.time = sat.time
        g.nlev = int(np.shape(sat.pressure_mid)[0])
        lon, lat = np.asarray(sat.longitude_center), np.asarray(sat.latitude_center)
        g.n_px = int(lat.size)
        names = ["vcd", "uncertainty", "quality_flag", "x_col", "averaging_kernels", "pressure_mid",
                 "apriori_profile"]
        names += ["aprior_column", "apriori_surface"] if self.sensor == "MOPITT" else ["pressure_weight"]
        g.dev = {n: _dev_flat(np.squeeze(getattr(sat, n)) if n == "quality_flag" else getattr(sat, n))
                 for n in names}
        radius = self.grid_size * 2.0
        if self.sensor == "GOSAT":
            gpl_f, X, Y = self._filler_mesh()
            g.fill_plan = _plan.granule_plan(lon, lat, gpl_f, radius=self.FILL_GRID)
            if g.fill_plan is None or g.fill_plan.n_cells == 0:
                self.n_skipped += 1
                return False
            g.near_plan = _plan.nearest_plan(lon, lat, gpl_f, radius=self.FILL_GRID)
            g.plan = _plan.granule_plan(X, Y, self.gplan, radius=radius)      # lattice: cached
        else:
            g.fill_plan = g.near_plan = None
            g.plan = _plan.granule_plan(lon, lat, self.gplan, radius=radius)  # L3 lattice: cached
        if g.plan is None or g.plan.n_cells == 0:
            self.n_skipped += 1
            return False
        self.granules.append(g)
        self._program = self._program_keep = None
        return True

    def n_pixels(self):
        return int(sum(g.n_px for g in self.granules))

    def input_bytes(self):
        return int(sum(t.numel() * t.element_size() for g in self.granules for t in g.dev.values()))

    # ------------------------------------------------------------------ stages
    def _good(self, qflag_dev, thresh):
        good = _dev.empty((qflag_dev.numel(),), "uint8")
        _lib.check(_lib.lib().oisat_quality_mask(qflag_dev.data_ptr(), _dev.dtype_code(qflag_dev),
                                                 qflag_dev.numel(), float(thresh), good.data_ptr(),
                                                 _dev.stream()))
        return good

    def _fill_gaps(self, g):
        """filler_gosat.py:87-201 on the device: the soundings -> float64 rows on the 1 degree
        filler mesh.  Returns ({name: device rows}, quality field)."""
        gpl_f, _, _ = self._filler_mesh()
        n_mesh = int(np.prod(gpl_f.out_shape))
        d, L = g.dev, g.nlev
        good = self._good(d["quality_flag"], self.flag_thresh)
        specs = [_FieldSpec("x_col", d["x_col"]),
                 _FieldSpec("uncertainty", d["uncertainty"], error=True),
                 _FieldSpec("averaging_kernels", d["averaging_kernels"], L),
                 _FieldSpec("pressure_mid", d["pressure_mid"], L),
                 _FieldSpec("apriori_profile", d["apriori_profile"], L),
                 _FieldSpec("pressure_weight", d["pressure_weight"], L)]
        out, layout, _keep = apply_plan(g.fill_plan, specs, good, g.n_px, n_mesh)
        # quality field: the 1/NaN mask at the nearest sounding, NaN beyond reach
        # (filler_gosat.py:100-103,153-155)
        ones = _dev.full((g.n_px,), 1.0)
        q, _, _keep2 = apply_plan(g.near_plan, [_FieldSpec("q", ones)], good, g.n_px, n_mesh)
        rows = {name: out[r0:r0 + nl].reshape(-1) for name, (r0, nl) in layout.items()}
        return rows, q.reshape(-1), n_mesh

    def _grid(self, g):
        """interpolator.py:100-291 on the device: float64 rows on the output mesh."""
        L = g.nlev
        if self.sensor == "GOSAT":
            rows, qflag, n_px = self._fill_gaps(g)
            good = self._good(qflag, 0.0)      # reader.py:1271: flag_thresh=0.0 on the filled image
            # vcd and x_col are the same image (filler_gosat.py:193,199): gridded once
            specs = [_FieldSpec("x_col", rows["x_col"]),
                     _FieldSpec("uncertainty", rows["uncertainty"], error=True),
                     _FieldSpec("averaging_kernels", rows["averaging_kernels"], L),
                     _FieldSpec("pressure_weight", rows["pressure_weight"], L),
                     _FieldSpec("pressure_mid", rows["pressure_mid"], L),
                     _FieldSpec("apriori_profile", rows["apriori_profile"], L)]
        else:
            d, n_px = g.dev, g.n_px
            good = self._good(d["quality_flag"], self.flag_thresh)
            specs = [_FieldSpec("vcd", d["vcd"]),
                     _FieldSpec("uncertainty", d["uncertainty"], error=True),
                     _FieldSpec("aprior_column", d["aprior_column"]),
                     _FieldSpec("apriori_surface", d["apriori_surface"]),
                     _FieldSpec("x_col", d["x_col"]),
                     _FieldSpec("averaging_kernels", d["averaging_kernels"], L + 1),
                     _FieldSpec("pressure_mid", d["pressure_mid"], L),
                     _FieldSpec("apriori_profile", d["apriori_profile"], L)]
        out, layout, keep = apply_plan(g.plan, specs, good, n_px, self.n_cell)
        return out, layout, keep

    def _model(self, day):
        """Model fields of the matched day on the output mesh (K6 when the model is finer than
        the mesh, ak_conv_mopitt.py:79-110): resampled once per day, not once per granule."""
        hit = self._model_on_mesh.get(day)
        if hit is not None:
            return hit
        pmid_d, prof_d, dp_d = model_fields(self.ctm_data, day)
        if not self.gplan.upscale:
            if pmid_d.dtype != _dev.torch().float32:
                raise _lib.OisatError("model fields must be float32 as delivered by the readers")
            if self._resample is None:       # geometry of the resampling: once per pipeline
                X, Y = self.gplan.mesh()
                self._resample = _v.resample_geometry(self.ctm_data, X, Y)
            pm, pr, third = _v.resample_with(
                self._resample, [(pmid_d, None, _lib.SRC_VALUE), (prof_d, None, _lib.SRC_VALUE),
                                 (dp_d, None, _lib.SRC_AIR_COLUMN)])
            hit = (pm, pr, third, 1)
        else:
            hit = (pmid_d, prof_d, dp_d, 0)
        self._model_on_mesh[day] = hit
        return hit

    def _convolve(self, g, out, layout):
        """K3 over every cell of the mesh (cells without a retrieval come out NaN)."""
        L = _lib.lib()
        _, day = _v.closest_day(self.ctm_data, self._stamps, g.time)
        pmid, prof, third, mode = self._model(day)
        n = self.n_cell

        def row(name):
            return out[layout[name][0]].data_ptr()

        xcol = _dev.empty((n,))
        if self.sensor == "MOPITT":
            col = _dev.empty((n,))
            _lib.check(L.oisat_vertical_mopitt(
                n, None, None, row("vcd"), row("aprior_column"), row("apriori_surface"),
                row("pressure_mid"), row("averaging_kernels"), row("apriori_profile"), g.nlev, n,
                pmid.data_ptr(), prof.data_ptr(), third.data_ptr(), mode, pmid.shape[0],
                pmid.shape[1], col.data_ptr(), xcol.data_ptr(), _dev.stream()))
            return out[layout["vcd"][0]], col, xcol
        _lib.check(L.oisat_vertical_gosat(
            n, None, None, row("x_col"), row("pressure_mid"), row("averaging_kernels"),
            row("apriori_profile"), row("pressure_weight"), g.nlev, n, pmid.data_ptr(),
            prof.data_ptr(), mode, pmid.shape[0], pmid.shape[1], xcol.data_ptr(), _dev.stream()))
        return out[layout["x_col"][0]], None, xcol       # ctm_vcd is NaN by design (:138)

    # --------------------------------------------------------------------- run
    def _granule_steps(self, acc):
        """Every launch of the granule loop, in order (gap filling, gridding, model resampling
        once per model day, AK convolution, accumulation)."""
        L = _lib.lib()
        self._model_on_mesh = {}     # the model is resampled inside the step, once per model day
        for g in self.granules:
            out, layout, _keep = self._grid(g)
            vcd, col, xcol = self._convolve(g, out, layout)
            _lib.check(L.oisat_accum_add(
                acc.data_ptr(), self.n_cell, vcd.data_ptr(), out[layout["uncertainty"][0]].data_ptr(),
                _dev.ptr(col), out[layout["x_col"][0]].data_ptr(), xcol.data_ptr(), _dev.stream()))

    def run(self, marks=None):
        """One month step.  The first run executes the granule loop through the Python layers
        and RECORDS its launches (`_lib.recording`): inputs are resident, plans cached and
        every buffer of the loop is kept alive, so the arguments of those ~8 launches per
        granule never change; later runs replay them call by call -- the months of these
        sensors are 30 small granules, and assembling a launch in Python (descriptor structs,
        allocations, look-ups) costs more than the kernel it starts.  OISAT_OPT_REPLAY=0 runs
        the loop afresh every time; add_granule() drops the recording."""
        import os

        def mark(name):
            if marks is not None:
                e = _dev.torch().cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        stream = _dev.stream()
        replay = os.environ.get("OISAT_OPT_REPLAY", "1") != "0"
        if replay and self._program is not None and self._program[2] == stream:
            rec, acc, _ = self._program
            acc.zero_()
            rec.replay()
        else:
            acc = _dev.zeros((10, self.n_cell))
            if replay:
                keep = [acc]
                with _dev.keep_allocations(keep), _lib.recording() as rec:
                    self._granule_steps(acc)
                self._program = (rec, acc, stream)
                self._program_keep = keep
            else:
                self._granule_steps(acc)
        self._acc = acc
        mark("granules")
        merged = acc
        if self.pg is not None:
            from .sharding import merge_accumulators
            merged = acc.clone()          # the recording's block keeps this rank's own sums
            merge_accumulators(merged, self.pg)
        res = finalize_and_oi(merged, self.n_cell, self.sensor, self.gas, self.error_ctm)
        mark("oi")
        return res

    def results_to_host(self, res):
        shape = tuple(self.gplan.out_shape)
        out = {}
        for k, v in res.items():
            if hasattr(v, "value"):
                out[k] = v.value()
            else:
                out[k] = _dev.to_host(v).reshape(shape) if hasattr(v, "data_ptr") else v
        return out
