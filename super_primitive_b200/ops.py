"""Small tensor helpers with the names and semantics of the reference's ``core.ops``.

Only the visualisation tools of the reference call these directly (tool/viz.py:71,121); on the
alignment path the same arithmetic runs fused inside the CUDA kernels (csrc/spb_align.cu), so these
are plain tensor expressions kept for API compatibility, differentiable like the reference's.

    transform_points_batch  core/ops.py:5-17      project_points_batch  core/ops.py:19-40
    project_points          core/ops.py:42-43     transform_points      core/dense_optim.py:117-122
    estimate_depth_diff     core/ops.py:59-96  (CUDA z-splat, csrc/spb_geom.cu; CUDA tensors only)
"""
from __future__ import annotations

import torch


def transform_points_batch(points_3d, poses):
    R = poses[:, :3, :3]
    t = poses[:, :3, 3]
    if points_3d.dim() == 2:
        moved = torch.einsum('bij,nj->bni', R, points_3d)
    else:
        moved = torch.einsum('bij,bnj->bni', R, points_3d)
    return moved + t[:, None, :]


def project_points_batch(points_3d, K):
    """Pinhole projection with the reference's guarded reciprocal: 1/z where |z| > 1e-6, else 1e-6."""
    eps = 1e-6
    z = points_3d[..., 2]
    big = torch.abs(z) > eps
    zi = torch.where(big, 1.0 / torch.where(big, z, torch.ones_like(z)), torch.full_like(z, eps))
    u = points_3d[..., 0] * K[..., 0, 0][:, None] * zi + K[..., 0, 2][:, None]
    v = points_3d[..., 1] * K[..., 1, 1][:, None] * zi + K[..., 1, 2][:, None]
    return torch.stack([u, v], dim=-1)


def project_points(points_3d, K):
    return project_points_batch(points_3d[None], K[None])[0]


def transform_points(points_3d, pose):
    return torch.matmul(points_3d, pose[:3, :3].T) + pose[:3, 3]


def estimate_depth_diff(points_3d, K, spatial_dim, mean=False):
    """z-splat of a point cloud (already in the target frame) into an ``spatial_dim`` depth image.
    Returns ``(image (1,H,W), valid_depth (P,) bool)`` like the reference.  Duplicates resolve to the last
    point in order (the CPU semantics of ``scatter_``) or to ``sum / (count + 1)`` when ``mean=True``."""
    from . import _native as nat
    from .geometry import _f32c, _stream
    if not points_3d.is_cuda:
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    with torch.no_grad():
        pts = _f32c(points_3d).reshape(-1, 3)
        H, W = int(spatial_dim[0]), int(spatial_dim[1])
        dev = pts.device
        keys = torch.empty(H * W, dtype=torch.int64, device=dev)
        acc = torch.empty(H * W, dtype=torch.int64, device=dev) if mean else None   # 32.32 fixed-point sums
        out = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        valid = torch.empty(pts.shape[0], dtype=torch.uint8, device=dev)
        Kc = _f32c(K)
        nat.check(nat.lib().spb_depth_splat_points(pts.data_ptr(), pts.shape[0], Kc.data_ptr(), H, W, 1 if mean else 0,
                                                   keys.data_ptr(), nat.ptr(acc), out.data_ptr(), valid.data_ptr(),
                                                   _stream()), "spb_depth_splat_points")
    return out, valid.bool()
