"""ctypes binding of the C-ABI shared library ``csrc/libspb200.so`` (include/spb200.h).

The product path has no CPU fallback: if the library is missing or a call fails, an
exception is raised (``NativeLibraryError`` / ``RuntimeError``).  Loading the library does not
create a CUDA context.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPB200_LIB") or os.path.join(_HERE, "csrc", "libspb200.so")   # env: tuning variants only

TILE = 128
PAD = 4
PACK_WORDS = 4 + 5 * TILE
PAIR_NOUT = 16
GN_PAIR_NOUT = 48
GN_SEG_NOUT = 10
GN_NA = 36
LM_NSTATE = 8
ADAM_PAIR = 24
ADAM_SEG = 2
WIN_OPT_POSE, WIN_OPT_AFF, WIN_OPT_SEEDS = 1, 2, 4
WIN_ADAM_FRAME = 16
WIN_NSTATE = 8
MAX_INLINE_PAIRS = 16


class NativeLibraryError(RuntimeError):
    pass


class SpbGeom(C.Structure):
    _fields_ = [("uv", C.c_void_p), ("logd", C.c_void_p), ("tiles", C.c_void_p),
                ("seg_tile", C.c_void_p), ("seg_lkp", C.c_void_p), ("K", C.c_void_p),
                ("n_pts", C.c_int32), ("n_pad", C.c_int32), ("n_seg", C.c_int32),
                ("n_tiles", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class SpbPair(C.Structure):
    _fields_ = [("trg_rgba", C.c_void_p), ("src_rgb", C.c_void_p), ("tile_pack", C.c_void_p),
                ("K_trg", C.c_void_p),
                ("pose", C.c_void_p), ("k", C.c_void_p), ("aff_src", C.c_void_p),
                ("aff_trg", C.c_void_p), ("geom", C.c_int32), ("Hl", C.c_int32),
                ("Wl", C.c_int32), ("tau", C.c_float)]


class SpbStats(C.Structure):
    _fields_ = [("src_pts", C.c_void_p), ("moved_pts", C.c_void_p), ("trg_px", C.c_void_p),
                ("residual_raw", C.c_void_p), ("trg_ok", C.c_void_p), ("full_mask", C.c_void_p)]


class SpbFrameJob(C.Structure):
    _fields_ = [("src_u8", C.c_void_p), ("trg_u8", C.c_void_p), ("src_planar", C.c_void_p), ("src_rgb", C.c_void_p),
                ("pack", C.c_void_p), ("trg_rgba", C.c_void_p), ("geom", C.c_int32), ("Hl", C.c_int32),
                ("Wl", C.c_int32), ("pad_", C.c_int32)]


class SpbWindow(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("n_frames", C.c_int32), ("n_edges", C.c_int32), ("seg_total", C.c_int32),
                ("win_frame_off", C.c_void_p), ("win_edge_off", C.c_void_p), ("edge_src", C.c_void_p),
                ("edge_trg", C.c_void_p), ("edge_w", C.c_void_p), ("edge_seg_off", C.c_void_p),
                ("frame_seg_off", C.c_void_p), ("frame_seg_cnt", C.c_void_p), ("frame_flags", C.c_void_p),
                ("frame_T", C.c_void_p), ("frame_aff", C.c_void_p), ("k", C.c_void_p), ("edge_pose", C.c_void_p),
                ("adam_frame", C.c_void_p), ("adam_seg", C.c_void_p), ("win_state", C.c_void_p),
                ("edge_tw", C.c_void_p)]


_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_PROTOS = {
    "spb_compact_count": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "spb_compact_scan": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spb_tile_table": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "spb_compact_count_depth": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "spb_compact_fill_depth": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spb_compact_fill": (_i, [_vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spb_pack_rgba": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "spb_sample_source": (_i, [C.POINTER(SpbGeom), _vp, _i, _i, _vp, _vp]),
    "spb_build_tile_pack": (_i, [C.POINTER(SpbGeom), _vp, _vp, _vp]),
    "spb_image_tt": (_i, [_vp, _i, _i, _vp, _vp]),
    "spb_ingest_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "spb_cost_grad": (_i, [C.POINTER(SpbGeom), C.POINTER(SpbPair), _i, _vp, _vp, _vp, _vp, _vp,
                           C.POINTER(SpbStats), _vp]),
    "spb_cost_grad_points": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(SpbPair), _vp, _vp, _vp]),
    "spb_workspace_floats": (_i64, [C.POINTER(SpbGeom), _i, _i]),
    "spb_workspace_floats_points": (_i64, [_i]),
    "spb_gn_ctas": (_i, [_i, _i]),
    "spb_gn_work_stride": (_i64, [_i, _i]),
    "spb_gn_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _f, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "spb_gn_iterate": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                            _vp, _vp]),
    "spb_grad_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "spb_adam_update": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _d, _vp]),
    "spb_adam_iterate": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d,
                              _d, _d, _d, _vp, _vp, _vp]),
    "spb_window_poses": (_i, [C.POINTER(SpbWindow), _vp]),
    "spb_window_update": (_i, [C.POINTER(SpbWindow), _vp, _vp, _d, _d, _d, _d, _d, _d, _d, _vp]),
    "spb_window_iterate": (_i, [_vp, _vp, C.POINTER(SpbWindow), _i, _i, _vp, _i64, _vp, _vp, _d, _d, _d, _d, _d, _d, _d,
                                _vp, _vp, _vp]),
    "spb_lm_saved_floats": (_i, [_i, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "spb_lm_update": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spb_dense_depths": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "spb_depth_splat": (_i, [C.POINTER(SpbGeom), _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "spb_depth_splat_points": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "spb_depth_avg_dense": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "spb_depth_avg_compact": (_i, [C.POINTER(SpbGeom), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spb_fill_nearest": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "spb_pyr_down": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "spb_lift_points": (_i, [C.POINTER(SpbGeom), _vp, _vp, _vp, _vp, _vp]),
    "spb_segment_reinit": (_i, [C.POINTER(SpbGeom), _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "spb_tile_points": (_i, []),
    "spb_version": (_i, []),
}

EXPORTS = tuple(_PROTOS)
_lib = None


def lib():
    """The loaded library; raises NativeLibraryError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or super_primitive_b200/csrc/build.sh).  There is no CPU fallback.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _PROTOS.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        global TILE, PACK_WORDS
        TILE = int(handle.spb_tile_points())      # the tile size is a build-time constant of the library
        PACK_WORDS = 4 + 5 * TILE
    return _lib


def check(status, what):
    if status != 0:
        if status > 0:
            raise RuntimeError(f"{what}: CUDA error {status}")
        raise RuntimeError(f"{what}: invalid arguments (code {status})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
