"""Drop-in replacement for the reference's ``core.depth_render``.

    estimate_depth_kf_native   core/depth_render.py:7-21 (+ core/ops.py:59-96 estimate_depth_diff)

Renders a keyframe's segments into the view ``pose`` as an (H,W) depth map, 0 = empty.  With
``mean=False`` duplicates resolve deterministically to the LAST point in (segment,row,col) order --
the CPU semantics of ``scatter_``; the reference's own GPU path is non-deterministic there.
"""
from __future__ import annotations

import torch

from . import _native as nat
from .geometry import _f32c, _stream, geometry_of


def estimate_depth_kf_native(kf, kf_logdepth, pose=None, mean=False):
    with torch.no_grad():
        geom = geometry_of(kf)
        if not bool(torch.isfinite(kf_logdepth).all()):
            raise AssertionError("kf_logdepth is not finite")
        k_c = _f32c(kf_logdepth)
        dev = k_c.device
        H, W = geom.H, geom.W
        keys = torch.empty(H * W, dtype=torch.int64, device=dev)
        acc = torch.empty(H * W, dtype=torch.int64, device=dev) if mean else None   # 32.32 fixed-point sums
        out = torch.empty((H, W), dtype=torch.float32, device=dev)
        pose_c = None if pose is None else _f32c(pose)
        nat.check(nat.lib().spb_depth_splat(geom.cref, k_c.data_ptr(), nat.ptr(pose_c), 1 if mean else 0,
                                            keys.data_ptr(), nat.ptr(acc), out.data_ptr(), _stream()),
                  "spb_depth_splat")
    return out
