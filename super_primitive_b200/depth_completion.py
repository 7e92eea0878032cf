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
    return out, invalid.view(torch.bool)


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
    return out, invalid.view(torch.bool)


def complete_batch(kfs, sparse_depths, mode='median', fill_holes=False):
    """The reference's per-frame `DepthCompletion.infer_depth` tail (depth_completion/segment_based_completion.py:44-54:
    `segment_based_depth_reinit(partial_depth, kf, 'median', return_info=True)` -> `unproject_kf_to_depths` -> mask ->
    drop the unseeded segments -> `render_depth_avg`) for a BATCH of independent frames with two host syncs for the whole
    batch instead of two per frame: all mask compactions are queued and their point counts read back together
    (`geometry.geometries_of`), then every frame's re-initialisation and average render are queued, and the
    'no segment saw a depth' condition (the reference fails on `torch.median` of an empty tensor) is checked once at the end.

    kfs: keyframes; sparse_depths: (H,W) tensors (0 = no measurement; like the reference, entries < 1e-6 are clamped to
    1e-6 in place).  Returns a list of (depth (H,W), invalid (H,W) bool, k (N,), visible (N,) bool); with ``fill_holes``
    every tuple carries a fifth entry, the map with its invalid pixels filled from the nearest valid one
    (`fill_in_tools.fill_depth`, depth_completion/fill_in_tools.py:5-7)."""
    from .geometry import geometries_of
    if len(kfs) != len(sparse_depths):
        raise AssertionError("one sparse depth map per keyframe expected")
    lib = nat.lib()
    out = []
    with torch.no_grad():
        geoms = geometries_of(kfs)
        nvis_all = torch.empty(len(kfs), dtype=torch.int32, device=geoms[0].uv.device) if geoms else None
        for i, (kf, geom, est0) in enumerate(zip(kfs, geoms, sparse_depths)):
            dev = geom.uv.device
            if tuple(est0.shape) != (geom.H, geom.W):
                raise AssertionError("estimated_depth must have the keyframe's geometry size")
            est = est0
            if est.dtype != torch.float32 or not est.is_contiguous() or est.device != dev:
                est = est.to(device=dev, dtype=torch.float32).contiguous()
            N, HW = geom.N, geom.H * geom.W
            seg_val = torch.empty(N, dtype=torch.float32, device=dev)
            visible = torch.empty(N, dtype=torch.uint8, device=dev)
            k = torch.empty(N, dtype=torch.float32, device=dev)
            nat.check(lib.spb_segment_reinit(geom.cref, est.data_ptr(), 1 if mode == 'median' else 0, seg_val.data_ptr(),
                                             visible.data_ptr(), k.data_ptr(), nvis_all[i:i + 1].data_ptr(), _stream()),
                      "spb_segment_reinit")
            if est0.is_floating_point():
                est0.clamp_(min=1e-6)                      # = masked_fill_(est0 < 1e-6, 1e-6): NaN stays NaN
            acc = torch.empty(HW, dtype=torch.int64, device=dev)
            cnt = torch.empty(HW, dtype=torch.int32, device=dev)
            depth = torch.empty((geom.H, geom.W), dtype=torch.float32, device=dev)
            invalid = torch.empty((geom.H, geom.W), dtype=torch.uint8, device=dev)
            nat.check(lib.spb_depth_avg_compact(geom.cref, k.data_ptr(), visible.data_ptr(), acc.data_ptr(), cnt.data_ptr(),
                                                depth.data_ptr(), invalid.data_ptr(), _stream()), "spb_depth_avg_compact")
            if fill_holes:
                from .fill_in_tools import fill_depth
                out.append((depth, invalid.view(torch.bool), k, visible.view(torch.bool), fill_depth(depth, invalid.view(torch.bool))))
            else:
                out.append((depth, invalid.view(torch.bool), k, visible.view(torch.bool)))
        if geoms and int(nvis_all.min().item()) == 0:
            raise IndexError("complete_batch: a frame has no segment with a valid depth estimate")
    return out
