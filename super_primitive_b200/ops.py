"""Small tensor helpers with the names and semantics of the reference's ``core.ops``.

Only the visualisation tools of the reference call these directly (tool/viz.py:71,121); on the
alignment path the same arithmetic runs fused inside the CUDA kernels (csrc/spb_align.cu), so these
are plain tensor expressions kept for API compatibility, differentiable like the reference's.

    transform_points_batch  core/ops.py:5-17      project_points_batch  core/ops.py:19-40
    project_points          core/ops.py:42-43     transform_points      core/dense_optim.py:117-122
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
