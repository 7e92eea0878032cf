"""Drop-in replacement for the reference's ``core.dense_optim_batch``: one source keyframe against
B target images with B poses / intrinsics (windowed mapping, odometery/odometery.py:833).

    photomeric_cost_batch   core/dense_optim_batch.py:50-147
    get_pixels_batch        core/dense_optim_batch.py:12-46 -- fused into the kernel
                            (depth threshold 1e-6 instead of the single path's 1e-7)
"""
from __future__ import annotations

import torch

from .dense_optim import (LazyResult, _affine_pair, _check_cfg, _extra_channel_stats, _keypoint_stats, _mode_channels,
                          _PairCost, _point_stats)
from .geometry import _f32c, geometry_of, pack_rgba


def photomeric_cost_batch(src_keyframe, trg_images, trg_Ks, src_keypoint_logdepth, poses, cost_config,
                          affine_comp=None):
    """Returns ``{'residual': (B,)}`` (+ statistics when ``collect_stats > 0``)."""
    collect_stats, check = _check_cfg(cost_config)
    _mode_channels(cost_config['mode'], src_keyframe.image.shape[0])
    geom = geometry_of(src_keyframe)
    level = geom.level_buffers(src_keyframe.image)
    trg_rgba = pack_rgba(trg_images)
    B = trg_rgba.shape[0]
    if poses.shape[0] != B:
        raise AssertionError("one pose per target image expected")
    a_s, a_t = _affine_pair(affine_comp)
    Ks = _f32c(trg_Ks)
    if Ks.dim() == 3 and Ks.shape[0] != B:
        raise AssertionError("one intrinsics matrix per target image expected")
    tau = 1e-6
    residual = _PairCost.apply(src_keypoint_logdepth, poses, a_s, a_t, geom, level, trg_rgba, Ks, tau, check)
    if collect_stats <= 0:
        return {'residual': residual}
    k_c = _f32c(src_keypoint_logdepth).clone()
    poses_c = _f32c(poses).clone()
    as_c = None if a_s is None else _f32c(a_s).reshape(-1).clone()
    at_c = None if a_t is None else _f32c(a_t).reshape(-1, 2).expand(B, 2).contiguous().clone()
    src_image, mode = src_keyframe.image, cost_config['mode']

    def produce():
        with torch.no_grad():
            out = _point_stats(geom, src_image, level, trg_rgba, Ks, poses_c, k_c, as_c, at_c, tau, True)
            if src_image.shape[0] > 3:
                out = _extra_channel_stats(out, geom, src_image, trg_images, Ks, poses_c, mode)
            if collect_stats > 1:
                out.update(_keypoint_stats(geom, k_c, poses_c, Ks, Ks, tau, True))
        return out

    res = LazyResult(residual, produce)
    if cost_config.get('eager_stats', False):
        res._fill()
    return res
