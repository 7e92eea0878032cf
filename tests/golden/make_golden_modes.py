"""Golden fixture for the non-colour residual modes, generated from the LIVE reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_modes.py

`core.dense_optim.photomeric_cost` with mode 'colour_norm' on 6-channel keyframes (RGB + normals) and
`core.dense_optim_batch.photomeric_cost_batch` with mode 'colour_norm_kappa' on 7-channel ones are executed unmodified
on CPU float32; stored: the inputs, the residual (it equals the colour residual -- the reference never assigns its
normal term, core/dense_optim.py:241-261), the autograd gradients and the all-channel per-point statistics
(`src_pixels` with the normals rotated by the detached R, `src_in_trg_pixels`; core/normal_cost.py:11-30).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

from super_primitive_b200 import synthetic as syn          # noqa: E402
from super_primitive_b200.keyframe import KeyFrame          # noqa: E402
import core.dense_optim as ref_do                            # noqa: E402
import core.dense_optim_batch as ref_dob                     # noqa: E402
from image.keyframe import KeyFrame as RefKeyFrame           # noqa: E402


def with_channels(kf, extra, seed):
    g = torch.Generator().manual_seed(seed)
    H, W = kf.image.shape[1:]
    n = torch.randn(3, H, W, generator=g)
    n = n / n.norm(dim=0, keepdim=True)
    chans = [kf.image, n] + ([torch.rand(1, H, W, generator=g)] if extra == 4 else [])
    return KeyFrame(torch.cat(chans, 0), kf.K, kf.logdepth_perseg, kf.keypoints, kf.keypoint_regions, kf.K_img)


def ref_kf(kf):
    return RefKeyFrame(kf.image, kf.K, kf.logdepth_perseg, kf.keypoints, kf.keypoint_regions, K_img=kf.K_img)


def main():
    H, W, N, B = 48, 64, 5, 2
    src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind="rects", seed=13, noise=0.01)
    store = {}
    # ---- single target, 'colour_norm', 6 channels, brightness terms on
    s6, t6 = with_channels(src, 3, 1), with_channels(trg, 3, 2)
    cfg = {'mode': 'colour_norm', 'collect_stats': 1, 'normal_loss': 'lecrec', 'normal_weight': 0.1}
    k = k0.clone().requires_grad_(True)
    pose = syn.small_pose(0.02, -0.01, 0.005, 0.01, -0.02, 0.015).requires_grad_(True)
    a_s, a_t = torch.tensor([0.02, 0.01]), torch.tensor([-0.01, 0.02], requires_grad=True)
    out = ref_do.photomeric_cost(ref_kf(s6), ref_kf(t6), k, pose, cfg, (a_s, a_t))
    out['residual'].mean().backward()
    colour = ref_do.photomeric_cost(ref_kf(src), ref_kf(trg), k0, pose.detach(), {'mode': 'colour', 'collect_stats': 0},
                                    (a_s, a_t.detach()))
    assert torch.equal(out['residual'].detach(), colour['residual'])      # the normal term is dead code upstream
    store.update(s_src_image=s6.image.numpy(), s_trg_image=t6.image.numpy(), s_pose=pose.detach().numpy(),
                 s_aff_src=a_s.numpy(), s_aff_trg=a_t.detach().numpy(), s_residual=out['residual'].detach().numpy(),
                 s_g_k=k.grad.numpy(), s_g_pose=pose.grad.numpy(), s_g_aff_trg=a_t.grad.numpy(),
                 s_src_pixels=out['src_pixels'].detach().numpy(), s_src_in_trg_pixels=out['src_in_trg_pixels'].detach().numpy(),
                 s_residual_raw=out['residual_raw'].numpy())
    # ---- batch of two targets, 'colour_norm_kappa', 7 channels
    s7 = with_channels(src, 4, 3)
    timgs = torch.stack([with_channels(trg, 4, 4).image, with_channels(src, 4, 5).image])
    Ks = torch.stack([trg.K, trg.K])
    poses = torch.stack([syn.small_pose(0.02, 0.0, 0.0, 0.01, 0.0, -0.01), syn.small_pose(-0.01, 0.01, 0.0, 0.0, 0.02, 0.0)])
    poses.requires_grad_(True)
    kb = k0.clone().requires_grad_(True)
    cfgb = {'mode': 'colour_norm_kappa', 'collect_stats': 1, 'normal_loss': 'lecrec', 'normal_weight': 0.1}
    outb = ref_dob.photomeric_cost_batch(ref_kf(s7), timgs, Ks, kb, poses, cfgb)
    outb['residual'].mean().backward()
    store.update(b_src_image=s7.image.numpy(), b_trg_images=timgs.numpy(), b_poses=poses.detach().numpy(),
                 b_residual=outb['residual'].detach().numpy(), b_g_k=kb.grad.numpy(), b_g_poses=poses.grad.numpy(),
                 b_src_pixels=outb['src_pixels'].detach().numpy(),
                 b_src_in_trg_pixels=outb['src_in_trg_pixels'].detach().numpy())
    store.update(K=src.K.numpy(), K_img=src.K_img.numpy(), logdepth=src.logdepth_perseg.numpy(),
                 keypoints=src.keypoints.numpy(), regions=src.keypoint_regions.numpy(), k=k0.numpy())
    np.savez_compressed(os.path.join(HERE, "modes.npz"), **store)
    print("wrote modes.npz", {n: v.shape for n, v in store.items() if n.endswith("pixels")})


if __name__ == "__main__":
    main()
