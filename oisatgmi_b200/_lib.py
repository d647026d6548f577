"""ctypes binding of liboisat.so (declared in include/oisat.h).

The product path has no CPU fallback: if the shared library is missing the
first call raises, loudly, with the build command.  Loading the library does
not touch CUDA (static cudart initialises lazily), so importing this module in
joblib worker processes or on a CPU-only box is safe.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liboisat.so")

F16, F32, F64, U8, I32 = 1, 2, 3, 4, 5
OP_NONE, OP_SQUARE_NATIVE = 0, 1
POST_NONE, POST_SQRT = 0, 1
SRC_VALUE, SRC_PARTIAL_COLUMN, SRC_AIR_COLUMN = 0, 1, 2

vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double


class Field(C.Structure):
    """struct oisat_field"""
    _fields_ = [("data", vp), ("dtype", i32), ("op", i32), ("post", i32), ("nlev", i32),
                ("lev_stride", i64), ("box_weight", f64)]


class FusedArgs(C.Structure):
    """struct oisat_fused_args"""
    _fields_ = [
        ("n_tiles", i64), ("tile_granule", vp), ("tile_cell0", vp), ("tile_pair0", vp),
        ("tile_mask", vp),
        ("n_pairs", i64), ("nwin", i32), ("vert", vp), ("w", vp), ("box_weight", f64),
        ("box_weight_err", f64),
        ("n_granules", i32), ("gran_record0", vp), ("gran_px0", vp), ("gran_slot", vp),
        ("n_records", i64), ("records", vp), ("amf_masked", vp), ("n_sat_lev", i32),
        ("has_trop", i32),
        ("ctm_pmid", vp), ("ctm_logp", vp), ("ctm_pcol", vp), ("n_ctm_lev", i32), ("n_cell", i64),
        ("staged", vp), ("pair_granule", vp), ("pair_cell", vp),
        ("pair_record0", vp), ("pair_ctm_off", vp), ("alive_pairs", vp), ("n_alive", vp),
    ]


class FlipItem(C.Structure):
    """struct oisat_flip_item"""
    _fields_ = [("tri", vp), ("half", vp), ("n_tri", i64), ("px", vp), ("py", vp)]


class PackItem(C.Structure):
    """struct oisat_pack_item"""
    _fields_ = [("sw", vp), ("p_mid", vp), ("vcd", vp), ("sigma", vp), ("trop", vp), ("qflag", vp),
                ("amf", vp), ("n_px", i64), ("px0", i64), ("block0", i64)]


# name -> (restype, argtypes); mirrors include/oisat.h one to one
PROTOTYPES = {
    "oisat_last_error": (C.c_char_p, []),
    "oisat_abi_version": (C.c_int, []),
    "oisat_launch_count": (i64, []),
    "oisat_distmask": (C.c_int, [vp, vp, i32, i64, vp, i64, vp, i64, f64, vp, vp]),
    "oisat_h_delaunay": (i64, [vp, vp, i64, vp, i64, C.POINTER(i64)]),
    "oisat_h_delaunay_swath": (i64, [vp, vp, i64, i64, vp, i64, C.POINTER(i64), C.POINTER(i32)]),
    "oisat_h_delaunay_swath_adj": (i64, [vp, vp, i64, i64, vp, i64, vp, C.POINTER(i64), C.POINTER(i32)]),
    "oisat_h_delaunay_seed_parts": (i64, [vp, vp, i64, i64, vp, vp, vp, i64, vp, vp]),
    "oisat_h_delaunay_seed": (i64, [vp, vp, i64, i64, vp, i64, vp, C.POINTER(i64), i32, vp]),
    "oisat_h_flip_rounds": (C.c_int, [vp, vp, vp, vp, i64, i64, vp]),
    "oisat_seed_assemble": (C.c_int, [vp, i64, i64, i32, i64, vp, vp, i64, vp, vp, vp]),
    "oisat_flip_workspace_bytes": (i64, [i64]),
    "oisat_flip_delaunay": (C.c_int, [vp, vp, i64, vp, vp, i32, vp, vp, vp]),
    "oisat_flip_delaunay_batch": (C.c_int, [vp, i32, i32, vp, vp, vp]),
    "oisat_near_ties": (C.c_int, [vp, vp, i64, vp, vp, i32, f64, vp, vp, vp]),
    "oisat_flagged_nodes": (C.c_int, [vp, i64, vp, vp, vp]),
    "oisat_locate": (C.c_int, [vp, i64, vp, vp, i32, vp, i64, vp, i64, vp, vp, vp, vp]),
    "oisat_plan_cells": (C.c_int, [vp, i32, vp, i64, vp, vp, vp]),
    "oisat_plan_fill": (C.c_int, [vp, i64, vp, i32, vp, vp, vp, vp, i32, vp, i64, vp, i32, vp, vp,
                                  vp]),
    "oisat_nearest_pixel": (C.c_int, [vp, vp, i32, i64, vp, i64, vp, i64, f64, vp, vp, vp]),
    "oisat_plan_fill_nearest": (C.c_int, [vp, i64, vp, i32, vp, i32, vp, vp, vp]),
    "oisat_reader_scale_f16": (C.c_int, [vp, i32, i64, C.POINTER(f64), i32, vp, vp]),
    "oisat_reader_quality": (C.c_int, [i32, vp, i32, vp, i32, vp, i32, i64, vp, vp]),
    "oisat_reader_weights": (C.c_int, [vp, i32, i32, i32, i64, vp, i32, vp, vp]),
    "oisat_reader_pmid": (C.c_int, [i32, vp, vp, vp, i32, f64, i32, i64, vp, vp]),
    "oisat_reader_tropopause": (C.c_int, [vp, vp, i32, i64, vp, vp]),
    "oisat_reader_clean": (C.c_int, [vp, i32, i64, i32, i32, i32, C.POINTER(f64), i32, i32, i32, vp, vp]),
    "oisat_reader_mopitt_xcol": (C.c_int, [vp, vp, i64, vp, vp]),
    "oisat_quality_mask": (C.c_int, [vp, i32, i64, f64, vp, vp]),
    "oisat_interp_apply": (C.c_int, [vp, vp, i32, i64, vp, C.POINTER(Field), i32, vp, i64, vp, vp]),
    "oisat_vertical_amf": (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, i32,
                                     i32, i64, vp, vp, vp, vp]),
    "oisat_vertical_column": (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, vp]),
    "oisat_vertical_mopitt": (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp,
                                        i32, i32, i64, vp, vp, vp]),
    "oisat_vertical_gosat": (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, i32, i32,
                                       i64, vp, vp]),
    "oisat_grid_resample": (C.c_int, [vp, vp, i32, i32, i32, i64, i64, i32, i32, f64, vp, vp, i64,
                                      vp, i64, vp]),
    "oisat_accum_add": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "oisat_accum_add_variance": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "oisat_accum_finalize": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "oisat_oi_prepare": (C.c_int, [vp, vp, vp, i64, f64, f64, f64, vp, vp, vp]),
    "oisat_oi_sweep_workspace": (i64, [i64, i32]),
    "oisat_oi_sweep": (C.c_int, [vp, vp, i64, C.POINTER(f64), i32, vp, vp, vp, i64, vp]),
    "oisat_oi_apply": (C.c_int, [vp, vp, vp, vp, i64, f64, vp, vp, vp, vp, vp]),
    "oisat_oi_knee": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp]),
    "oisat_oi_apply_dev": (C.c_int, [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]),
    "oisat_output_fields": (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "oisat_pack_record_halfs": (i64, [i32, i32]),
    "oisat_pack_granule": (C.c_int, [vp, vp, i32, vp, vp, vp, i64, vp, vp]),
    "oisat_pack_blocks": (i64, [i64]),
    "oisat_pack_batch": (C.c_int, [vp, i32, i64, i32, i32, i32, f64, i32, vp, vp, vp]),
    "oisat_pack_batch_indexed": (C.c_int, [vp, i32, i64, vp, i32, i32, i32, f64, i32, vp, vp, vp]),
    "oisat_pack_batch_masked": (C.c_int, [vp, i32, i64, vp, i32, i32, i32, f64, i32, vp, vp, vp, i32, vp]),
    "oisat_pair_alive": (C.c_int, [i64, i32, vp, vp, vp, vp, vp, vp, vp, f64, vp, vp, vp, vp]),
    "oisat_ctm_prepare": (C.c_int, [vp, vp, vp, i64, vp, vp, vp]),
    "oisat_fused_amf": (C.c_int, [C.POINTER(FusedArgs), vp]),
    "oisat_rows_per_pair": (i64, [i32, i32]),
    "oisat_fused_amf_split": (C.c_int, [C.POINTER(FusedArgs), vp, vp]),
    "oisat_fused_amf_tile": (C.c_int, [C.POINTER(FusedArgs), vp]),
    "oisat_segment_tables": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
    "oisat_pair_tables": (C.c_int, [i64, vp, vp, vp, vp, i32, i64, i64, vp, vp, vp]),
    "oisat_reader_ssmis": (C.c_int, [vp, i32, i64, vp, vp, vp]),
    "oisat_pwv_partial": (C.c_int, [vp, vp, i64, vp, vp]),
    "oisat_pwv_column": (C.c_int, [vp, i32, i32, i64, vp, vp, vp]),
    "oisat_accum_pairs": (C.c_int, [vp, i64, vp, vp, vp, i64, vp]),
}

_lock = threading.Lock()
_lib = None


class OisatError(RuntimeError):
    pass


class Recorder:
    """Stands in for the library while it is installed (`with recording() as rec`): every
    kernel-launching entry point called on this thread is executed AND noted as (function,
    arguments).  A month step whose arguments do not change from one run to the next -- resident
    inputs, cached plans, buffers kept alive by `_dev.keep_allocations` -- can then be replayed
    call by call (`replay`) without the Python work that assembled it (descriptor structs,
    allocations, look-ups: ~30 us per launch, more than the small kernels themselves take)."""

    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith("oisat_") or fn.restype is not C.c_int or name == "oisat_abi_version":
            return fn

        def call(*args):
            self.calls.append((fn, args))
            return fn(*args)
        return call

    def replay(self):
        for fn, args in self.calls:
            rc = fn(*args)
            if rc != 0:
                check(rc)


_recorder = None
_recorder_thread = None


class recording:
    def __enter__(self):
        global _recorder, _recorder_thread
        if _recorder is not None:
            raise OisatError("a recording is already in progress")
        rec = Recorder(lib())
        _recorder, _recorder_thread = rec, threading.get_ident()
        return rec

    def __exit__(self, *exc):
        global _recorder, _recorder_thread
        _recorder = _recorder_thread = None
        return False


def lib():
    """The loaded library (raises if it has not been built)."""
    global _lib
    if _recorder is not None and threading.get_ident() == _recorder_thread:
        return _recorder
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OisatError(
                "liboisat.so is missing (%s): the CUDA extension is the product path and there "
                "is no CPU fallback.  Build it with `python -m oisatgmi_b200.csrc.build`."
                % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if h.oisat_abi_version() != 2:
            raise OisatError("liboisat ABI mismatch")
        _lib = h
        return _lib


def check(rc: int):
    if rc != 0:
        raise OisatError("liboisat error %d: %s" % (rc, lib().oisat_last_error().decode()))


def launch_count() -> int:
    return int(lib().oisat_launch_count())
