"""ORACLE (test infrastructure -- never imported by the product path).

A CPU restatement, in plain PyTorch, of the reference's dense photometric alignment path,
kept operation-for-operation equivalent to the reference's dense ``(N,H,W)`` formulation so
that (a) its outputs and autograd gradients are the parity target for the CUDA path and
(b) timing it on host cores is a fair stand-in ("port") for the reference's own
PyTorch-CPU path when ``/root/reference`` is not mounted (GPU box).

Pinned against the live reference: ``tests/golden/make_golden.py`` imports
``/root/reference/core`` in the build container, runs both on identical seeded inputs and
freezes the reference's outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this port reproduces them (bit-exact on CPU, same
op order).  The reference itself ships no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to /root/reference).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- coordinates
def to_unit(px, dims):
    """pixel -> [-1,1], float32 reciprocal of (dims-1).  tool/point_utils.py:31-35"""
    scale = 1.0 / (torch.as_tensor(dims, dtype=torch.float32, device=px.device) - 1)
    return 2 * px * scale - 1


def from_unit(xn, dims):
    """[-1,1] -> integer pixel (round half even).  tool/point_utils.py:37-40"""
    d = torch.as_tensor(dims, dtype=torch.float32, device=xn.device)
    return (0.5 * (d - 1) * (xn + 1)).round().long()


def _dims_of(logd):
    return logd.shape[1:] if logd.dim() == 3 else logd.shape


# ----------------------------------------------------------------------------- geometry
def seeded_logdepth(k, keypoints, regions, logd):
    """Shift every segment's log-depth so its keypoint has log-depth k_b, zero outside the
    mask.  core/dense_optim.py:38-80 (dense (N,H,W) add and multiply, two finiteness asserts)."""
    n = keypoints.shape[0]
    assert torch.isfinite(k).all()
    per_seg = logd.dim() == 3
    if per_seg:
        assert logd.shape[0] == n
    rc = from_unit(keypoints, _dims_of(logd))
    r, c = rc[:, 0].long(), rc[:, 1].long()
    at_kp = logd[torch.arange(n, device=r.device), r, c] if per_seg else logd[r, c]
    shift = k - at_kp
    dense = logd if per_seg else logd.unsqueeze(0).expand(n, -1, -1)
    dense = dense + shift[:, None, None]
    dense = dense * regions
    assert torch.all(torch.isfinite(dense))
    return dense


def backproject(uv, z, K):
    """Pinhole unprojection of (u,v) pixel columns with depth z.  core/dense_optim.py:19-35"""
    assert uv.shape[0] == z.shape[0]
    zz = z.reshape(-1)
    X = (uv[:, 0].reshape(-1).float() - K[0, 2]) * zz / K[0, 0]
    Y = (uv[:, 1].reshape(-1).float() - K[1, 2]) * zz / K[1, 1]
    return torch.stack([X, Y, zz], dim=1)


def lift_segments(depth, regions, K, want_coords=False):
    """Mask compaction in (segment,row,col) order + unprojection.  core/dense_optim.py:89-114"""
    count = torch.sum(regions, dim=[1, 2]).sum()
    b, r, c = torch.where(regions)
    z = depth[b, r, c]
    assert len(z) == count
    uv = torch.stack([c, r], dim=1)
    pts = backproject(uv, z, K)
    return (pts, b, uv) if want_coords else (pts, b)


def rigid_batch(pts, poses):
    """R X + t for B poses via einsum.  core/ops.py:5-17"""
    R, t = poses[:, :3, :3], poses[:, :3, 3]
    eq = 'bij, nj -> bni' if pts.dim() == 2 else 'bij, bnj -> bni'
    return torch.einsum(eq, R, pts) + t[:, None, :]


def rigid(pts, pose):
    """core/dense_optim.py:117-122"""
    return torch.matmul(pts, pose[:3, :3].T) + pose[:3, 3]


def pinhole_batch(pts, K):
    """Projection with guarded reciprocal depth (eps 1e-6, masked assignment).
    core/ops.py:19-40"""
    eps = 1e-6
    fx, fy, cx, cy = K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    zi = torch.ones_like(z) * eps
    # the mask is evaluated twice (two compares + two boolean-index passes), exactly as in the
    # reference, so that the CPU timing of this port stays representative
    zi[torch.abs(z) > eps] = 1.0 / z[torch.abs(z) > eps]
    u = x * fx[:, None] * zi + cx[:, None]
    v = y * fy[:, None] * zi + cy[:, None]
    return torch.stack([u, v], dim=-1)


def pinhole(pts, K):
    """core/ops.py:42-43"""
    return pinhole_batch(pts[None], K[None])[0]


# ----------------------------------------------------------------------------- sampling
def bilinear(img, grid):
    """grid_sample(bilinear, zeros, align_corners=True) + |coord|<=0.99 validity.
    core/dense_optim.py:128-140.  img (B,C,H,W), grid (B,P,2) -> (B,C,P), (B,P)"""
    ok = torch.all(torch.abs(grid) <= 0.99, dim=-1)
    out = F.grid_sample(img, grid.unsqueeze(1), mode="bilinear", padding_mode="zeros",
                        align_corners=True)
    return out.squeeze(2), ok


def sample_single(image, pts, K, spatial_dim=None):
    """project -> normalise by the *geometry* dims -> sample the level image; AND with
    z > 1e-7.  core/dense_optim.py:143-162"""
    front = pts[..., 2].detach() > 1e-7
    uv = pinhole(pts, K)
    if spatial_dim is None:
        spatial_dim = image.shape[1:]
    g = to_unit(uv, (spatial_dim[1], spatial_dim[0]))
    vals, ok = bilinear(image[None], g[None])
    return vals, torch.logical_and(ok, front)


def sample_batch(image, pts, K, spatial_dim=None):
    """Batched variant: z > 1e-6, flip to (row,col) for normalisation and back.
    core/dense_optim_batch.py:12-46"""
    front = pts[..., 2].detach() > 1e-6
    uv = pinhole(pts, K) if pts.dim() == 2 else pinhole_batch(pts, K)
    if not torch.isfinite(uv).all():
        assert torch.all(torch.isfinite(uv))
    if spatial_dim is None:
        spatial_dim = image.shape[1:]
    g = to_unit(uv.flip(-1), spatial_dim).flip(-1)
    if image.dim() == 3:
        image = image.unsqueeze(0)
    if g.dim() == 2:
        g = g.unsqueeze(0)
    vals, ok = bilinear(image, g)
    return vals, torch.logical_and(ok, front)


def brightness(trg_px, src_ab, trg_ab):
    """exp(-(a_t-a_s)) rgb + (b_t-b_s) on the first three channels.
    core/dense_optim.py:202-225"""
    rgb, rest = trg_px[:, :3], trg_px[:, 3:]
    if src_ab is None:
        assert trg_ab is None
        return trg_px
    if src_ab.dim() == 1:
        src_ab = src_ab.unsqueeze(0)
    if trg_ab.dim() == 1:
        trg_ab = trg_ab.unsqueeze(0)
    sa, sb = torch.split(src_ab, [1, 1], dim=-1)
    ta, tb = torch.split(trg_ab, [1, 1], dim=-1)
    a = ta[:, None].expand(-1, 3, -1) - sa[:, None].expand(-1, 3, -1)
    b = tb[:, None].expand(-1, 3, -1) - sb[:, None].expand(-1, 3, -1)
    return torch.cat([torch.exp(-a) * rgb + b, rest], dim=1)


def masked_l1(src_px, trg_px, mask, want_raw=False):
    """mean over (3, P) of |(src - trg) * mask| ('colour' mode: first 3 channels; the normal
    term of the reference is dead code).  core/dense_optim.py:228-261,
    core/cost_utils.py:4-19"""
    r = (src_px[:, :3] - trg_px[:, :3]) * mask
    raw = r.detach().clone() if want_raw else None
    return torch.abs(r).mean(dim=[1, 2]), raw


# ----------------------------------------------------------------------------- entry points
def dense_depths(kf, k):
    """core/dense_optim.py:164-174  -> (N,H,W) depth, 1 outside masks"""
    return torch.exp(seeded_logdepth(k, kf.keypoints, kf.keypoint_regions, kf.get_logdepth()))


def lift_keyframe(kf, k):
    """core/dense_optim.py:176-200"""
    dims = kf.geo_spatial_dim()
    depth = dense_depths(kf, k)
    pts, seg = lift_segments(depth, kf.keypoint_regions, kf.K)
    px, ok = sample_single(kf.image, pts, kf.K, spatial_dim=dims)
    return {'src_pixels': px, 'src_valid_mask': ok, 'src_pts': pts, 'segm_ids': seg,
            'spatial_size': dims}


def _keypoint_stats_single(src, trg, k, pose, depth, dims):
    kp_uv = from_unit(src.keypoints, depth.shape[1:]).flip(-1)
    kp3 = rigid(backproject(kp_uv, torch.exp(k), src.K), pose)
    _, ok = sample_single(trg.image, kp3, trg.K, spatial_dim=dims)
    return {'src_in_trg_keypoints': pinhole(kp3, trg.K_img),
            'src_in_trg_keypoints_z': kp3[:, 2],
            'src_in_trg_keypoints_valid_mask': ok}


def cost_single(src, trg, k, pose, cost_config, affine_comp=None):
    """core/dense_optim.py:265-363 (colour mode)"""
    stats_level = cost_config['collect_stats']
    dims = src.geo_spatial_dim()
    depth = dense_depths(src, k)
    pts, seg = lift_segments(depth, src.keypoint_regions, src.K)
    extra = _keypoint_stats_single(src, trg, k, pose, depth, dims) if stats_level > 1 else {}
    assert torch.all(torch.isfinite(pts))
    moved = rigid(pts, pose)
    src_px, src_ok = sample_single(src.image, pts, src.K, spatial_dim=dims)
    assert torch.all(torch.isfinite(moved))
    trg_px, trg_ok = sample_single(trg.image, moved, trg.K, spatial_dim=dims)
    both = trg_ok[:, None].long() * src_ok[:, None].long()
    if affine_comp is not None:
        trg_px = brightness(trg_px, affine_comp[0], affine_comp[1])
    res, raw = masked_l1(src_px, trg_px, both, want_raw=stats_level > 0)
    assert torch.all(torch.isnan(pts) == False)      # noqa: E712  (the reference's guards)
    assert torch.all(torch.isnan(src_px) == False)   # noqa: E712
    assert torch.all(torch.isnan(res) == False)      # noqa: E712
    out = {'residual': res}
    if stats_level > 0:
        out.update({'segm_ids': seg, 'src_pixels': src_px, 'src_in_trg_pixels': trg_px,
                    'src_valid_mask': src_ok, 'trg_valid_mask': trg_ok, 'full_mask': both,
                    'src_pts': pts, 'src_in_trg_pts': moved, 'residual_raw': raw,
                    'median_depth': None})
        out.update(extra)
    return out


def cost_precomputed(pre, trg, pose, cost_config, affine_comp=None):
    """core/dense_optim.py:365-403"""
    moved = rigid(pre['src_pts'], pose)
    trg_px, trg_ok = sample_single(trg.image, moved, trg.K, spatial_dim=pre['spatial_size'])
    both = trg_ok[:, None].long() * pre['src_valid_mask'][:, None].long()
    if affine_comp is not None:
        trg_px = brightness(trg_px, affine_comp[0], affine_comp[1])
    res, _ = masked_l1(pre['src_pixels'], trg_px, both)
    return {'residual': res}


def cost_batch(src, trg_images, trg_Ks, k, poses, cost_config, affine_comp=None):
    """core/dense_optim_batch.py:50-147 (colour mode)"""
    stats_level = cost_config['collect_stats']
    dims = src.geo_spatial_dim()
    depth = dense_depths(src, k)
    pts, seg = lift_segments(depth, src.keypoint_regions, src.K)
    assert torch.all(torch.isfinite(pts))
    moved = rigid_batch(pts, poses)
    extra = {}
    if stats_level > 1:
        kp_uv = from_unit(src.keypoints, depth.shape[1:]).flip(-1)
        kp3 = rigid_batch(backproject(kp_uv, torch.exp(k), src.K), poses)
        _, ok = sample_batch(trg_images, kp3, trg_Ks, spatial_dim=dims)
        extra = {'src_in_trg_keypoints': pinhole_batch(kp3, trg_Ks),
                 'src_in_trg_keypoints_z': kp3[..., 2],
                 'src_in_trg_keypoints_valid_mask': ok}
    src_px, src_ok = sample_single(src.image, pts, src.K, spatial_dim=dims)
    trg_px, trg_ok = sample_batch(trg_images, moved, trg_Ks, spatial_dim=dims)
    both = trg_ok[:, None].long() * src_ok[:, None].long()
    if affine_comp is not None:
        trg_px = brightness(trg_px, affine_comp[0], affine_comp[1])
    res, raw = masked_l1(src_px, trg_px, both, want_raw=stats_level > 0)
    assert torch.all(torch.isnan(pts) == False)      # noqa: E712
    assert torch.all(torch.isnan(src_px) == False)   # noqa: E712
    assert torch.all(torch.isnan(res) == False)      # noqa: E712
    out = {'residual': res}
    if stats_level > 0:
        out.update({'segm_ids': seg, 'src_pixels': src_px, 'src_in_trg_pixels': trg_px,
                    'src_valid_mask': src_ok, 'trg_valid_mask': trg_ok, 'full_mask': both,
                    'src_pts': pts, 'src_in_trg_pts': moved, 'residual_raw': raw,
                    'median_depth': None})
        out.update(extra)
    return out


# ----------------------------------------------------------------------------- depth render
def splat_depth(pts, K, dims, mean=False):
    """z-splat with .long() truncation; last writer wins on CPU (point order) or
    scatter_reduce mean that includes the initial zero.  core/ops.py:59-96"""
    front = pts[..., 2].detach() > 1e-6
    with torch.no_grad():
        rc = pinhole(pts, K).flip(-1).long()
    z = pts[..., 2]
    canvas = torch.zeros((1, *dims), device=rc.device, dtype=torch.float32)
    r, c = rc[..., 0], rc[..., 1]
    keep = front * (r >= 0) * (r < dims[0]) * (c >= 0) * (c < dims[1])
    z, r, c = z[keep], r[keep], c[keep]
    flat = canvas.reshape(-1)
    idx = r * dims[1] + c
    if mean:
        flat.scatter_reduce_(0, idx, z, reduce='mean')
    else:
        flat.scatter_(0, idx, z)
    return flat.reshape((1, *dims)), keep


def render_keyframe_depth(kf, k, pose=None, mean=False):
    """core/depth_render.py:7-21"""
    with torch.no_grad():
        pts = lift_keyframe(kf, k)['src_pts']
        if pose is None:
            pose = torch.eye(4, device=k.device)
        img, _ = splat_depth(rigid(pts, pose), kf.K, kf.geo_spatial_dim(), mean=mean)
    return img[0]


# ----------------------------------------------------------------------------- next row
def segment_median_reinit(est_depth, kf, mode='median'):
    """Per-segment mean / lower-median of (log est_depth - logdepth_perseg) over pixels that
    are in the mask and have a valid estimate; invisible segments get the median of the
    visible ones.  odometery/depth_init.py:10-67.  Does not modify ``est_depth``."""
    assert mode in ('mean', 'median')
    eps = 1e-6
    n = kf.keypoints.shape[0]
    _, H, W = kf.logdepth_perseg.shape
    rc = from_unit(kf.keypoints, (H, W))
    b = torch.arange(n, device=kf.logdepth_perseg.device)
    est = est_depth.clone()
    bad = est < eps
    good = ~bad
    est[bad] = eps
    shifts = torch.log(est)[None] - kf.logdepth_perseg
    valid = kf.keypoint_regions * good[None]
    shifts = shifts * valid
    cnt = valid.sum((1, 2))
    vis = cnt > 0
    out = torch.zeros(n, device=est.device)
    at_kp = kf.logdepth_perseg[b, rc[:, 0], rc[:, 1]]
    if mode == 'mean':
        out[vis] = shifts[vis].sum((1, 2)) / cnt[vis]
    else:
        vals = [torch.median(s[m]) for s, m in zip(shifts[vis], valid[vis])]
        out[vis] = torch.stack(vals, 0)
    out[vis] += at_kp[vis]
    out[~vis] = torch.median(out[vis])
    return out, vis


def average_render(depths):
    """Per-pixel average of the valid entries of stacked depth maps; zeroes entries < 1e-6 in place.
    depth_completion/segment_based_completion.py:21-27"""
    invalid = depths.max(dim=0)[0] < 1e-6
    depths[depths < 1e-6] = 0.0
    n_valid = ((depths > 1e-6).sum(dim=0) + 1e-6)
    return depths.sum(dim=0) / n_valid, invalid


def completion_render(kf, k, visible):
    """depth_completion/segment_based_completion.py:48-54: dense depths, -1 outside masks, drop unseeded
    segments, average."""
    d = dense_depths(kf, k)
    d[kf.keypoint_regions == 0] = -1
    d = d[visible]
    return average_render(d)
