"""Deterministic synthetic inputs for the dense alignment path.

No dataset, SAM or normal network is available offline, so keyframes are generated:
a smooth sinusoid RGB image (meaningful image gradients), a pinhole camera, N segment
masks (vertical strips, dilated/overlapping strips, or seeded random rectangles to mimic
SAM), a per-segment log-depth ramp and one keypoint per segment.  The same generator
feeds the GPU path, the CPU oracle and the CPU baseline timing so every arm sees
identical inputs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .keyframe import KeyFrame


def pinhole(H, W, dtype=torch.float32):
    f = 0.8 * W
    return torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=dtype)


def sinus_image(H, W, shift=(0.0, 0.0), noise=0.0, seed=0, dtype=torch.float32):
    """I_c(x, y) = 0.5 + 0.5 sin(2 pi (3x + 2y + c/3)) on the unit square, sampled at pixel
    centres displaced by ``shift`` pixels (col, row).  Optional seeded uniform noise."""
    ys = (torch.arange(H, dtype=torch.float64) + shift[1]) / H
    xs = (torch.arange(W, dtype=torch.float64) + shift[0]) / W
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    chans = [0.5 + 0.5 * torch.sin(2 * math.pi * (3 * xx + 2 * yy + c / 3.0)) for c in range(3)]
    img = torch.stack(chans, 0)
    if noise > 0:
        g = torch.Generator().manual_seed(seed)
        img = img + (torch.rand(img.shape, generator=g, dtype=torch.float64) * 2 - 1) * noise
    return img.to(dtype)


def segment_masks(H, W, N, kind="strips", seed=0):
    """Returns (masks (N,H,W) bool, keypoints_rc (N,2) int64 row/col inside each mask)."""
    masks = torch.zeros((N, H, W), dtype=torch.bool)
    kp = torch.zeros((N, 2), dtype=torch.int64)
    if kind in ("strips", "overlap"):
        cols = torch.arange(W)
        ids = (cols * N) // W
        pad = 4 if kind == "overlap" else 0
        for b in range(N):
            own = torch.nonzero(ids == b).flatten()
            lo = max(int(own[0]) - pad, 0)
            hi = min(int(own[-1]) + pad + 1, W)
            masks[b, :, lo:hi] = True
            kp[b, 0] = H // 2
            kp[b, 1] = int(min(max(round((b + 0.5) * W / N), int(own[0])), int(own[-1])))
    elif kind == "rects":
        rng = np.random.RandomState(seed)
        for b in range(N):
            h = int(rng.randint(max(H // 8, 2), max(H // 2, 3)))
            w = int(rng.randint(max(W // 8, 2), max(W // 2, 3)))
            r0 = int(rng.randint(0, H - h + 1))
            c0 = int(rng.randint(0, W - w + 1))
            masks[b, r0:r0 + h, c0:c0 + w] = True
            # punch a hole so masks are not convex (ragged rows)
            if h > 6 and w > 6:
                masks[b, r0 + h // 3:r0 + h // 3 + 2, c0 + w // 3:c0 + w // 3 + 3] = False
            kp[b, 0] = r0 + h // 2 + 2
            kp[b, 1] = c0 + w // 2 + 2
            if not masks[b, kp[b, 0], kp[b, 1]]:
                rr, cc = torch.nonzero(masks[b], as_tuple=True)
                kp[b, 0], kp[b, 1] = rr[len(rr) // 2], cc[len(cc) // 2]
    else:
        raise ValueError(kind)
    return masks, kp


def normalise_rc(rc, dims):
    """(row, col) pixel -> [-1, 1] with the reference's (dims-1) convention and float32
    reciprocal (tool/point_utils.py:31-35)."""
    inv = 1.0 / (torch.as_tensor(dims, dtype=torch.float32) - 1)
    return 2 * rc.to(torch.float32) * inv - 1


def make_keyframe(H, W, N, kind="strips", seed=0, noise=0.0, shift=(0.0, 0.0),
                  supporting=False, dtype=torch.float32):
    img = sinus_image(H, W, shift=shift, noise=noise, seed=seed, dtype=dtype)
    K = pinhole(H, W, dtype)
    if supporting:
        return KeyFrame(img, K)
    masks, kp = segment_masks(H, W, N, kind, seed)
    x = (torch.arange(W, dtype=torch.float64) / W)[None, None, :].expand(N, H, W)
    y = (torch.arange(H, dtype=torch.float64) / H)[None, :, None].expand(N, H, W)
    # per-segment ramp with a different tilt per segment, zero outside the mask
    tilt = (0.05 * torch.cos(torch.arange(N, dtype=torch.float64)))[:, None, None]
    logd = ((0.1 * x + tilt * y) * masks).to(dtype)
    keypoints = normalise_rc(kp, (H, W)).to(dtype)
    return KeyFrame(img, K, logd, keypoints, masks)


def small_pose(tx=0.02, ty=0.0, tz=0.0, rx=0.0, ry=0.0, rz=0.0, dtype=torch.float32):
    """4x4 rigid transform from a translation and XYZ Euler angles (float64 internally)."""
    cx, sx = math.cos(rx), math.sin(rx)
    cy, sy = math.cos(ry), math.sin(ry)
    cz, sz = math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [tx, ty, tz]
    return torch.tensor(T, dtype=dtype)


_BLUR = torch.tensor([[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]]) / 16.0


def image_pyramid(image, start_level, end_level):
    """Coarse-to-fine list of (C,H_l,W_l) images: 3x3 [1 2 1]^2/16 blur with reflect padding
    then 2x decimation, semantics of reference image/gaussian_pyramid.py:53-85."""
    out = []
    x = image[None]
    C = image.shape[0]
    ker = _BLUR.to(image.dtype).to(image.device).repeat(C, 1, 1, 1)
    for i in range(end_level - 1):
        if i >= start_level:
            out.insert(0, x[0])
        x = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect"),
                                       ker, groups=C)[:, :, 0::2, 0::2]
    out.insert(0, x[0])
    return out


def level_intrinsics(K, level):
    """K_img of pyramid level ``level`` (0 = full res): reference
    image/gaussian_pyramid.py:43-51,113-119 (T @ K with T = [[s,0,s],[0,s,s],[0,0,1]])."""
    s = 2.0 ** (-level)
    T = torch.tensor([[s, 0, s], [0, s, s], [0, 0, 1]], dtype=K.dtype, device=K.device)
    return T @ K


def keyframe_pyramid(kf, start_level, end_level):
    """Coarse-to-fine keyframes with ``geo_down=False`` semantics (the only mode the callers
    use, reference image/keyframe.py:77-148): only image / K_img change per level."""
    imgs = image_pyramid(kf.image[:3], start_level, end_level)
    levels = list(range(start_level, end_level))[::-1]
    out = []
    for img, lvl in zip(imgs, levels):
        out.append(KeyFrame(img, kf.K.clone(), kf.logdepth_perseg, kf.keypoints,
                            kf.keypoint_regions, level_intrinsics(kf.K, lvl), kf.id))
    return out


# Named workloads from BASELINE.json `configs`.
CONFIGS = {
    "C1": dict(H=192, W=256, N=8, levels=(0, 1)),
    "C2": dict(H=480, W=640, N=64, levels=(0, 3)),
    "C3": dict(H=224, W=288, N=100, levels=(0, 3)),
    "C4": dict(H=480, W=640, N=100, levels=(0, 1)),
    "C5": dict(H=768, W=1024, N=256, levels=(0, 1)),
}


def two_frame_problem(H, W, N, kind="strips", seed=0, noise=0.0, shift=(2.0, 1.0)):
    """Source keyframe with geometry + a supporting target frame whose image is the same
    scene displaced by ``shift`` pixels; initial log-depth seeds log 2; initial pose ~2 cm translation + a few mrad rotation."""
    src = make_keyframe(H, W, N, kind=kind, seed=seed, noise=noise)
    trg = make_keyframe(H, W, N, shift=shift, noise=noise, seed=seed + 1, supporting=True)
    k0 = torch.full((N,), math.log(2.0), dtype=torch.float32)
    # generic position on purpose: with an axis-aligned integer-pixel displacement every warped point
    # lands exactly on a texel row/column, where the slope of bilinear interpolation is discontinuous
    # and float32 rounding (in the reference too) decides which side is taken
    pose0 = small_pose(0.02, 0.004, -0.003, 0.003, -0.002, 0.0015)
    return src, trg, k0, pose0


def planar_scene_pair(H, W, N, pose_true, z0=2.0, kind="strips", seed=0):
    """A geometrically CONSISTENT two-frame problem: a fronto-parallel textured plane at depth ``z0`` in the source
    frame, rendered exactly (analytic texture, ray/plane intersection) into the source view and into the view
    ``pose_true`` (source -> target).  Aligning the two must recover ``pose_true`` (up to the joint scale of
    translation and depth).  Returns (src keyframe, target frame, true log-depth seeds)."""
    K = pinhole(H, W, torch.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]

    def tex(X, Y):
        chans = []
        for c in range(3):
            ph = 2 * math.pi * c / 3.0
            chans.append(0.5 + 0.3 * torch.sin(2 * math.pi * (1.5 * X + 1.0 * Y) + ph)
                         + 0.15 * torch.sin(2 * math.pi * (-0.9 * X + 2.3 * Y) + 1.7 * ph))
        return torch.stack(chans, 0)

    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    src_img = tex((u - cx) * z0 / fx, (v - cy) * z0 / fy)
    T = pose_true.to(torch.float64)
    R, t = T[:3, :3], T[:3, 3]
    d = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)          # target-frame rays
    Rt_d = d @ R                                                                       # (R^T d) per pixel
    Rt_t = R.T @ t
    lam = (z0 + Rt_t[2]) / Rt_d[..., 2]
    Ps = lam[..., None] * Rt_d - Rt_t                                                  # R^T (lam d - t)
    trg_img = tex(Ps[..., 0], Ps[..., 1])
    masks, kp = segment_masks(H, W, N, kind, seed)
    logd = torch.zeros((N, H, W), dtype=torch.float32)                                 # plane: constant depth
    src = KeyFrame(src_img.float(), K.float(), logd, normalise_rc(kp, (H, W)), masks)
    trg = KeyFrame(trg_img.float(), K.float())
    k_true = torch.full((N,), math.log(z0), dtype=torch.float32)
    return src, trg, k_true


def mapping_window(H, W, N, n_kf=3, n_supp=1, kind="overlap", seed=0, noise=0.01, affine=True, window_full=True,
                   dtype=torch.float32):
    """One mapping window in the reference's shape (odometery/odometery.py:451-479, 576-648, 798-820): ``n_kf``
    keyframes connected to their neighbours (i -> i-1, i+1) plus ``n_supp`` supporting frames per keyframe that serve
    as targets of their own keyframe AND of the next one; the first keyframe's pose and brightness are held, and its
    seeds too when the window is full.  Returns {'frames': [...], 'edges': [...]} (see window.py)."""
    frames, supp_of = [], [[] for _ in range(n_kf)]
    for i in range(n_kf):
        kf = make_keyframe(H, W, N, kind=kind, seed=seed + 7 * i, noise=noise, shift=(1.7 * i, 0.9 * i), dtype=dtype)
        T = small_pose(0.02 * i, -0.004 * i, 0.003 * i, 0.003 * i, -0.002 * i, 0.0015 * i, dtype=dtype)
        frames.append(dict(T=T, image=kf.image, K=kf.K, kf=kf,
                           k=torch.full((N,), math.log(2.0) + 0.01 * i, dtype=dtype),
                           aff=torch.tensor([0.02 * i, -0.01 * i], dtype=dtype) if affine else None,
                           opt_pose=i > 0, opt_aff=affine and i > 0, opt_seeds=(i > 0) or not window_full))
    for i in range(n_kf):
        for j in range(n_supp):
            a = i + (j + 1) / (n_supp + 1.0)
            fr = make_keyframe(H, W, N, seed=seed + 100 + 13 * i + j, noise=noise, shift=(1.7 * a, 0.9 * a),
                               supporting=True, dtype=dtype)
            T = small_pose(0.02 * a, -0.004 * a, 0.003 * a, 0.003 * a, -0.002 * a, 0.0015 * a, dtype=dtype)
            supp_of[i].append(len(frames))
            frames.append(dict(T=T, image=fr.image, K=fr.K, kf=None, k=None,
                               aff=torch.tensor([0.01 * a, 0.005 * a], dtype=dtype) if affine else None,
                               opt_pose=True, opt_aff=affine, opt_seeds=False))
    edges = []
    for s in range(n_kf):
        if s > 0:
            edges.append((s, s - 1))
        if s < n_kf - 1:
            edges.append((s, s + 1))
        for ss in ([s, s - 1] if s > 0 else [s]):
            edges.extend((s, t) for t in supp_of[ss])
    return dict(frames=frames, edges=edges)
