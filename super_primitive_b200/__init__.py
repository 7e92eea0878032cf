"""super_primitive_b200 -- B200-native (sm_100a) dense photometric alignment, a drop-in for the hot
path of makezur/super_primitive (``core.dense_optim``, ``core.dense_optim_batch``,
``core.depth_render``).  See DESIGN.md / INTEGRATION.md.
"""
from __future__ import annotations

import sys
import types

__version__ = "0.1.0"

_ALIASES = ("dense_optim", "dense_optim_batch", "depth_render", "ops")


def install_as_core(force=True):
    """Make ``import core.dense_optim`` (etc.) resolve to this package so the reference's callers
    (odometery/, depth_completion/, gui/, tool/viz.py) run unchanged.  Call before importing them.
    The reference's remaining ``core`` modules (cost_utils, normal_cost) are not needed by callers."""
    import importlib
    pkg = sys.modules.get("core")
    if pkg is None or force:
        pkg = types.ModuleType("core")
        pkg.__path__ = []          # namespace-like package
        sys.modules["core"] = pkg
    for name in _ALIASES:
        mod = importlib.import_module(f"{__name__}.{name}")
        sys.modules[f"core.{name}"] = mod
        setattr(pkg, name, mod)
    return pkg


def install_depth_init():
    """Also route ``odometery.depth_init`` (segment_based_depth_reinit) to the CUDA path.  Call before
    ``odometery.odometery`` / ``depth_completion`` are imported; the rest of the reference's ``odometery``
    package stays the reference's."""
    import importlib
    mod = importlib.import_module(f"{__name__}.depth_init")
    sys.modules["odometery.depth_init"] = mod
    return mod


def install_pyramid():
    """Replace ``image.keyframe.keyframe_pyramid`` of the (already importable) reference with the CUDA pyramid;
    callers use ``keyframe.keyframe_pyramid(...)`` through the module attribute, so patching it is enough."""
    import importlib
    ref_mod = importlib.import_module("image.keyframe")
    mod = importlib.import_module(f"{__name__}.pyramid")
    ref_mod.keyframe_pyramid = mod.keyframe_pyramid
    return mod


def install_image_tt():
    """Replace ``tool.etc.image_tt`` (tool/etc.py:37-40; the frame hand-over of frontend/process_frame.py:216,258) of the
    already importable reference with the device-side conversion: the 8-bit frame is uploaded (3 bytes per pixel
    instead of 12) and divided by 255 on the device -- bit-identical output.  Calls that ask for a CPU tensor (the SAM
    pre-resize, frontend/process_frame.py:101) keep running the reference's own function.  ``frontend.process_frame``
    binds the name at import (``from tool.etc import image_tt``), so it is patched too when already imported."""
    import importlib

    import torch
    ref_mod = importlib.import_module("tool.etc")
    mod = importlib.import_module(f"{__name__}.frames")
    reference_image_tt = ref_mod.image_tt

    def image_tt(image, device='cuda'):
        if torch.device(device).type != 'cuda':
            return reference_image_tt(image, device)
        return mod.image_tt(image, device)

    image_tt.__wrapped__ = mod.image_tt
    ref_mod.image_tt = image_tt
    fp = sys.modules.get("frontend.process_frame")
    if fp is not None and hasattr(fp, "image_tt"):
        fp.image_tt = image_tt
    return image_tt
