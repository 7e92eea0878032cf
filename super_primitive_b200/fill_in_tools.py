"""Hole filling of completed depth maps (reference depth_completion/fill_in_tools.py).

    fill_depth(depth, invalid_mask)          fill_in_tools.py:5-7 -- every invalid pixel takes the depth of the nearest
                                             valid pixel (exact Euclidean feature transform; scipy's tie-breaking),
                                             bit-identical to `depth[tuple(distance_transform_edt(invalid, False, True))]`
    fill_depth_batch(depths, invalid_masks)  the same for a stack of frames in two launches
    fill_single_griddata                     fill_in_tools.py:9-21 -- NOT provided: its first stage is scipy's Delaunay
                                             (Qhull) linear interpolation, whose result on a pixel lattice (all points
                                             co-circular in fours) is decided by Qhull's own tie-breaking

The reference works on numpy arrays (`evaluate_void.py:124-125`); numpy inputs are uploaded, filled on the GPU and
returned as numpy arrays, CUDA tensors stay on the device.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .geometry import _stream


def _fill(depth, invalid, want_indices):
    if depth.dim() != 3 or tuple(invalid.shape) != tuple(depth.shape):
        raise AssertionError("fill_depth expects depth and invalid_mask of the same (H,W) / (F,H,W) shape")
    if not depth.is_cuda or not invalid.is_cuda:
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    F, H, W = (int(v) for v in depth.shape)
    exact = depth.dtype == torch.float32           # other dtypes: the kernel finds the pixels, the values are gathered
    d = depth.contiguous() if exact else depth.to(torch.float32).contiguous()
    m = (invalid if invalid.dtype == torch.bool else invalid != 0).contiguous().view(torch.uint8)
    out = torch.empty_like(d)
    scratch = torch.empty((F, H, W), dtype=torch.int32, device=d.device)
    idx = torch.empty((F, 2, H, W), dtype=torch.int32, device=d.device) if (want_indices or not exact) else None
    nat.check(nat.lib().spb_fill_nearest(d.data_ptr(), m.data_ptr(), F, H, W, scratch.data_ptr(), out.data_ptr(),
                                         nat.ptr(idx), _stream()), "spb_fill_nearest")
    if not exact:
        rows, cols = idx[:, 0].long() % H, idx[:, 1].long()         # row -1 (no valid pixel) wraps like numpy
        out = depth[torch.arange(F, device=d.device)[:, None, None], rows, cols]
    return out, (idx if want_indices else None)


def fill_depth_batch(depths, invalid_masks, return_indices=False, device="cuda"):
    """(F,H,W) depth maps + (F,H,W) bool masks (True = to be filled) -> (F,H,W) filled maps
    [, (F,2,H,W) int32 (row, col) of the pixel each value came from]."""
    as_numpy = isinstance(depths, np.ndarray)
    d = torch.as_tensor(depths)
    m = torch.as_tensor(invalid_masks)
    if as_numpy:
        d, m = d.to(device), m.to(device)
    out, idx = _fill(d, m, return_indices)
    if as_numpy:
        out = out.cpu().numpy()
        idx = idx.cpu().numpy() if idx is not None else None
    return (out, idx) if return_indices else out


def fill_depth(depth, invalid_mask, return_indices=False, device="cuda"):
    """depth_completion/fill_in_tools.py:5-7 for one (H,W) map; numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    if isinstance(depth, np.ndarray):
        res = fill_depth_batch(depth[None], np.asarray(invalid_mask)[None], return_indices, device)
    else:
        res = fill_depth_batch(depth[None], invalid_mask[None], return_indices, device)
    if return_indices:
        return res[0][0], res[1][0]
    return res[0]


def fill_single_griddata(depths_zbuff, pred_invalid_np):
    raise NotImplementedError(
        "fill_single_griddata (depth_completion/fill_in_tools.py:9-21) is not provided: scipy.interpolate.griddata's "
        "Delaunay triangulation of a pixel lattice is decided by Qhull's tie-breaking, so no result is comparable "
        "with the reference's; use fill_depth (its second stage, exact) instead")
