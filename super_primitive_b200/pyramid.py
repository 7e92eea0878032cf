"""Drop-in for the reference's ``image.keyframe.keyframe_pyramid`` (image/keyframe.py:77-148) -- SURVEY.md
section 8(f) rank 3: the coarse-to-fine list of keyframes built once per frame before the alignment loop.

With ``geo_down=False`` (every caller) only the image and ``K_img`` change per level: the RGB channels are
blurred (3x3 [1 2 1]^2/16, reflect padding) and decimated by 2 per level by the CUDA kernel
``spb_pyr_down``; normal channels (image[3:]) are nearest-neighbour decimated like the reference's
``DepthPyramidModule``; geometry tensors are shared, so the compact geometry is built once and reused by
every level.  ``geo_down=True`` additionally decimates log-depth and masks with ``[::2, ::2]`` slicing.

Like the reference, grad mode is switched off at entry and ON at exit (a side effect its callers rely on).
"""
from __future__ import annotations

import torch

from . import _native as nat
from .geometry import _f32c, _stream
from .keyframe import KeyFrame


def _pyr_down(img):
    C, H, W = img.shape
    out = torch.empty((C, (H + 1) // 2, (W + 1) // 2), dtype=torch.float32, device=img.device)
    nat.check(nat.lib().spb_pyr_down(img.data_ptr(), C, H, W, out.data_ptr(), _stream()), "spb_pyr_down")
    return out


def _levels(x, start_level, end_level, step):
    pyr = []
    for i in range(end_level - 1):
        if i >= start_level:
            pyr.insert(0, x)
        x = step(x)
    pyr.insert(0, x)
    return pyr


def level_intrinsics(K, level):
    """image/gaussian_pyramid.py:43-51,113-119: T @ K with T = [[s,0,s],[0,s,s],[0,0,1]], s = 2^-level."""
    s = 2.0 ** (-level)
    T = torch.tensor([[s, 0, s], [0, s, s], [0, 0, 1]], dtype=K.dtype, device=K.device)
    return torch.matmul(T, K)


def keyframe_pyramid(keyframe, start_level, end_level, geo_down=False, drop_normals=False, grayscale=False):
    torch.set_grad_enabled(False)
    if grayscale:
        raise NotImplementedError("grayscale pyramids are not used by any caller of the alignment path")
    image = keyframe.image
    if not image.is_cuda:
        torch.set_grad_enabled(True)
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    rgb = _f32c(image[:3])
    image_pyr = _levels(rgb, start_level, end_level, _pyr_down)
    nearest = lambda t: t[..., 0::2, 0::2]          # noqa: E731  DepthPyramidModule 'nearest_neighbor'
    with_normals = image.shape[0] > 3
    normals_pyr = _levels(image[3:], start_level, end_level, nearest) if with_normals else [None] * len(image_pyr)
    supporting = keyframe.is_supporting() if hasattr(keyframe, "is_supporting") else keyframe.keypoints is None
    if geo_down and not supporting:
        depth_pyr = _levels(keyframe.logdepth_perseg, start_level, end_level, nearest)
        mask_pyr = _levels(keyframe.keypoint_regions, start_level, end_level, nearest)
    else:
        depth_pyr = mask_pyr = [None] * len(image_pyr)
    intr_pyr = [level_intrinsics(keyframe.K, i) for i in range(start_level, end_level)][::-1]
    out = []
    compact = getattr(keyframe, "_spb_geometry", None)
    if compact is not None and not geo_down:
        # handover.CompactKeyFrame: the levels share its compact geometry, no dense tensor is touched
        from .handover import CompactKeyFrame
        for img, intr, norms in zip(image_pyr, intr_pyr, normals_pyr):
            if norms is not None and not drop_normals:
                img = torch.cat([img, norms.to(img.dtype)], dim=0)
            out.append(CompactKeyFrame(img, keyframe.K.clone(), compact, keyframe.keypoints, K_img=intr,
                                       id=getattr(keyframe, "id", None)))
        torch.set_grad_enabled(True)
        return out
    for img, depth, mask, intr, norms in zip(image_pyr, depth_pyr, mask_pyr, intr_pyr, normals_pyr):
        if norms is not None and not drop_normals:
            img = torch.cat([img, norms.to(img.dtype)], dim=0)
        cls = type(keyframe) if type(keyframe).__name__ == "KeyFrame" else KeyFrame
        out.append(cls(img,
                       K=intr if geo_down else keyframe.K.clone(),
                       logdepth_perseg=(depth if geo_down else keyframe.logdepth_perseg),
                       keypoints=keyframe.keypoints,
                       keypoint_regions=((mask.bool() if mask is not None else None) if geo_down
                                         else keyframe.keypoint_regions),
                       K_img=intr, id=getattr(keyframe, "id", None)))
    torch.set_grad_enabled(True)
    return out
