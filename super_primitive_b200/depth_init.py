"""Drop-in replacement for the reference's ``odometery.depth_init.segment_based_depth_reinit``
(odometery/depth_init.py:10-67) -- SURVEY.md section 8(f) rank 1, the step that follows
``estimate_depth_kf_native`` on every new keyframe and the core of VOID depth completion.

For every segment: k_b = (mean | lower median) over the segment's pixels with a valid estimate
(>= 1e-6) of ``log(est_depth) - logdepth_perseg``, plus the log-depth at the keypoint; segments
without any valid pixel get the lower median of the visible ones.  The reference runs a Python loop
of ``torch.median`` over masked dense tensors; here one CTA per segment radix-selects over the compact
point list (csrc/spb_reinit.cu).

Faithful side effects: like the reference, grad mode is switched off and left off, and invalid entries
of a tensor ``estimated_depth`` are clamped to 1e-6 in place.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .geometry import _stream, geometry_of


def segment_based_depth_reinit(estimated_depth, kf, mode='mean', return_info=False):
    assert mode == 'mean' or mode == 'median'
    torch.set_grad_enabled(False)
    geom = geometry_of(kf)
    device = geom.uv.device
    if isinstance(estimated_depth, np.ndarray):
        estimated_depth = torch.from_numpy(estimated_depth).to(device)
    if tuple(estimated_depth.shape) != (geom.H, geom.W):
        raise AssertionError("estimated_depth must have the keyframe's geometry size")
    est = estimated_depth
    if est.dtype != torch.float32 or not est.is_contiguous() or est.device != device:
        est = est.to(device=device, dtype=torch.float32).contiguous()
    N = geom.N
    seg_val = torch.empty(N, dtype=torch.float32, device=device)
    visible = torch.empty(N, dtype=torch.uint8, device=device)
    out = torch.empty(N, dtype=torch.float32, device=device)
    nvis = torch.empty(1, dtype=torch.int32, device=device)
    nat.check(nat.lib().spb_segment_reinit(geom.cref, est.data_ptr(), 1 if mode == 'median' else 0,
                                           seg_val.data_ptr(), visible.data_ptr(), out.data_ptr(), nvis.data_ptr(),
                                           _stream()), "spb_segment_reinit")
    # reference semantics: estimated_depth[estimated_depth < eps] = eps (in place, after the read above)
    if estimated_depth.is_floating_point():
        estimated_depth.clamp_(min=1e-6)             # = masked_fill_(est < 1e-6, 1e-6): NaN stays NaN
    if int(nvis.item()) == 0:
        # torch.median of an empty tensor: the reference fails here too
        raise IndexError("segment_based_depth_reinit: no segment has a valid depth estimate")
    torch.set_grad_enabled(False)
    if return_info:
        return out, visible.view(torch.bool)
    return out
