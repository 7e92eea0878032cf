"""Shared helpers for the test-suite: golden fixture loading and comparison utilities."""
from __future__ import annotations

import os

import numpy as np
import torch

from super_primitive_b200.keyframe import KeyFrame

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL_CASES = ["tiny_rects", "tiny_strips", "c1_overlap", "pyr3_rects"]
STATS_CASES = ["tiny_rects", "tiny_strips"]
CFG0 = {'mode': 'colour', 'collect_stats': 0}
CFG2 = {'mode': 'colour', 'collect_stats': 2}


class Golden:
    """One frozen reference case (see tests/golden/make_golden.py)."""

    def __init__(self, name, device="cpu"):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.device = torch.device(device)
        self.N = int(self.z["N"])
        self.B = int(self.z["B"]) if "B" in self.z else 1
        self.H, self.W = int(self.z["H"]), int(self.z["W"])
        self.with_affine = bool(self.z["with_affine"]) if "with_affine" in self.z else False
        self.n_levels = sum(1 for k in self.z.files if k.endswith("_src_image") and k.startswith("L"))

    def t(self, key):
        return torch.from_numpy(self.z[key]).to(self.device)

    def has(self, key):
        return key in self.z.files

    def src(self, level):
        tag = f"L{level}_"
        return KeyFrame(self.t(tag + "src_image"), self.t("src_K"), self.t("src_logdepth"),
                        self.t("src_keypoints"), self.t("src_regions"), K_img=self.t(tag + "src_K_img"))

    def trg(self, level, j=0):
        tag = f"L{level}_"
        return KeyFrame(self.t(tag + "trg_images")[j], self.t("src_K"), K_img=self.t(tag + "src_K_img"))

    def trg_images(self, level):
        return self.t(f"L{level}_trg_images")

    def trg_Ks(self):
        return self.t("src_K")[None].repeat(self.B, 1, 1)

    def k(self):
        return self.t("k")

    def poses(self):
        return self.t("poses")

    def affine(self, j=None):
        if not self.with_affine:
            return None
        a_t = self.t("aff_trg")
        return self.t("aff_src"), (a_t if j is None else a_t[j])


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over the whole array (scale-relative, not element-wise)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: scale-relative error {e:.3e} > {tol:.1e}"


def assert_close_elem(a, b, tol, what="", floor=1e-3):
    """ELEMENT-wise bar: |a - b| <= tol * max(|b|, floor * max|b|) for every entry -- a small per-segment gradient
    cannot hide behind the largest one (the scale-relative bar above lets it)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-30))
    e = np.abs(a - b) / scale
    i = int(np.argmax(e))
    assert e.max() <= tol, (f"{what}: element {np.unravel_index(i, e.shape)} off by {e.max():.3e} of its own "
                            f"magnitude (got {a.flat[i]:.6e}, want {b.flat[i]:.6e}) > {tol:.1e}")


def elem_err(a, b, floor=1e-3):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-30))
    return float((np.abs(a - b) / scale).max())


def to_np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
