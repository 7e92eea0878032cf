"""Keyframe container accepted by the alignment path.

The alignment functions are duck-typed: anything with the attributes below works,
including the reference's own ``image.keyframe.KeyFrame`` (an ``nn.Module``,
reference ``image/keyframe.py:20-65``).  This light container exists so tests, the
benchmark and multi-GPU shards can build inputs without importing the reference.

Attributes (same meaning as the reference):
    image              (C, H_l, W_l) float32 -- pyramid-level image (C >= 3, RGB first)
    K                  (3, 3) float32        -- intrinsics of the *geometry* grid (H, W)
    K_img              (3, 3) float32        -- intrinsics scaled to the image level
    logdepth_perseg    (N, H, W) float32     -- per-segment log-depth up to a shift, 0 outside the mask
    keypoints          (N, 2) float32        -- (row, col) in [-1, 1], ``(dims-1)`` convention
    keypoint_regions   (N, H, W) bool        -- segment masks (may overlap)
"""
from __future__ import annotations

import torch


class KeyFrame:
    def __init__(self, image, K, logdepth_perseg=None, keypoints=None,
                 keypoint_regions=None, K_img=None, id=None):
        self.image = image
        self.K = K
        self.K_img = K if K_img is None else K_img
        self.id = id
        self.supporting = (logdepth_perseg is None or keypoints is None
                           or keypoint_regions is None)
        self.logdepth_perseg = None
        self.keypoints = None
        self.keypoint_regions = None
        if not self.supporting:
            if keypoints.shape[0] != keypoint_regions.shape[0]:
                raise AssertionError("one keypoint per segment mask expected")
            self.logdepth_perseg = logdepth_perseg
            self.keypoints = keypoints
            self.keypoint_regions = keypoint_regions

    # -- reference-compatible accessors (image/keyframe.py:49-65) ---------------
    def get_logdepth(self):
        return self.logdepth_perseg

    def geo_spatial_dim(self):
        ld = self.logdepth_perseg
        return ld.shape[1:] if ld.dim() == 3 else ld.shape

    def is_supporting(self):
        return self.supporting

    def num_segments(self):
        return self.keypoint_regions.shape[0]

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return KeyFrame(mv(self.image), mv(self.K), mv(self.logdepth_perseg),
                        mv(self.keypoints), mv(self.keypoint_regions),
                        mv(self.K_img), self.id)

    def __repr__(self):
        n = 0 if self.supporting else self.keypoints.shape[0]
        return f"KeyFrame(image={tuple(self.image.shape)}, segments={n})"
