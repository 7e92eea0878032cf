"""VOID depth-completion tail (BASELINE config 4).

    render_depth_avg          depth_completion/segment_based_completion.py:21-27 (dense drop-in, in place)
    render_segments_avg       lines 48-54 fused: unproject_kf_to_depths -> mask -> keep visible segments -> average,
                              computed from the compact geometry without the (N,H,W) tensor
"""
from __future__ import annotations

import torch

from . import _native as nat
from .geometry import _f32c, _stream, geometry_of


def render_depth_avg(depths):
    """(N,H,W) stacked per-segment depths (anything < 1e-6 = absent) -> (average (H,W), invalid (H,W) bool).
    Like the reference, entries < 1e-6 of ``depths`` are set to 0 in place."""
    if not depths.is_cuda:
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    if depths.dtype != torch.float32 or not depths.is_contiguous():
        raise AssertionError("render_depth_avg expects a contiguous float32 (N,H,W) tensor (it is modified in place)")
    N, H, W = depths.shape
    out = torch.empty((H, W), dtype=torch.float32, device=depths.device)
    invalid = torch.empty((H, W), dtype=torch.uint8, device=depths.device)
    nat.check(nat.lib().spb_depth_avg_dense(depths.data_ptr(), N, H, W, out.data_ptr(), invalid.data_ptr(), _stream()),
              "spb_depth_avg_dense")
    return out, invalid.bool()


def render_segments_avg(kf, keypoint_logdepth, visible_seg=None):
    """Average depth render of a keyframe's (visible) segments with seeds ``keypoint_logdepth``."""
    with torch.no_grad():
        geom = geometry_of(kf)
        k_c = _f32c(keypoint_logdepth)
        dev = k_c.device
        vis = None
        if visible_seg is not None:
            vis = (visible_seg if visible_seg.dtype == torch.bool else visible_seg != 0).contiguous().view(torch.uint8)
        HW = geom.H * geom.W
        acc = torch.empty(HW, dtype=torch.int64, device=dev)        # 32.32 fixed-point sums (order-independent atomics)
        cnt = torch.empty(HW, dtype=torch.int32, device=dev)
        out = torch.empty((geom.H, geom.W), dtype=torch.float32, device=dev)
        invalid = torch.empty((geom.H, geom.W), dtype=torch.uint8, device=dev)
        nat.check(nat.lib().spb_depth_avg_compact(geom.cref, k_c.data_ptr(), nat.ptr(vis), acc.data_ptr(), cnt.data_ptr(),
                                                  out.data_ptr(), invalid.data_ptr(), _stream()),
                  "spb_depth_avg_compact")
    return out, invalid.bool()
