"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), module aliasing for the reference's callers, the lazy statistics
dictionary, the synthetic generator and the pyramid semantics."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "spb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(spb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from super_primitive_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _declared_functions()
    assert len(names) >= 18
    handle = ctypes.CDLL(_native.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in include/spb200.h but not exported: {missing}"
    # and the Python binding covers the same set
    assert set(names) == set(_native.EXPORTS), set(names) ^ set(_native.EXPORTS)
    assert _native.lib().spb_version() >= 100


def test_struct_layouts_match_header_sizes():
    from super_primitive_b200 import _native
    assert ctypes.sizeof(_native.SpbGeom) == 6 * 8 + 6 * 4
    assert ctypes.sizeof(_native.SpbPair) == 8 * 8 + 4 * 4
    assert ctypes.sizeof(_native.SpbStats) == 6 * 8
    assert ctypes.sizeof(_native.SpbFrameJob) == 6 * 8 + 4 * 4
    assert ctypes.sizeof(_native.SpbWindow) == 4 * 4 + 17 * 8


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    from super_primitive_b200 import _native
    lib = _native.lib()
    assert lib.spb_compact_count(None, 1, 4, 4, None, None) == -1
    assert lib.spb_pack_rgba(None, 0, 1, 4, 4, None, None) == -1
    assert lib.spb_cost_grad(None, None, 1, None, None, None, None, None, None, None) == -1
    assert lib.spb_gn_accumulate(None, None, None, 1, 1, 1e-3, 0, None, 0, None, None, None, None, None) == -1
    assert lib.spb_segment_reinit(None, None, 1, None, None, None, None, None) == -1
    assert lib.spb_image_tt(None, 4, 4, None, None) == -1
    assert lib.spb_ingest_u8(None, None, 1, 1, 1, 1, None) == -1
    assert lib.spb_window_poses(None, None) == -1
    empty = _native.SpbWindow()
    assert lib.spb_window_update(ctypes.byref(empty), None, None, 1e-4, 1e-2, 1e-5, 0.9, 0.999, 1e-8, 0.0, None) == -1


def test_install_as_core_aliases_modules():
    import super_primitive_b200 as spb
    saved = {k: v for k, v in sys.modules.items() if k == "core" or k.startswith("core.")}
    try:
        spb.install_as_core()
        import core.dense_optim as do
        import core.dense_optim_batch as dob
        import core.depth_render as dr
        import core.ops as ops
        for fn in ("photomeric_cost", "photomeric_cost_precomputed", "unproject_kf", "unproject_kf_to_depths",
                   "transform_points", "project_points"):
            assert callable(getattr(do, fn)), fn
        assert callable(dob.photomeric_cost_batch) and callable(dr.estimate_depth_kf_native)
        assert callable(ops.project_points_batch) and callable(ops.transform_points_batch)
    finally:
        for k in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_install_image_tt_patches_the_reference_helper():
    import types
    import super_primitive_b200 as spb
    from super_primitive_b200 import frames
    saved = {k: sys.modules.get(k) for k in ("tool", "tool.etc", "frontend", "frontend.process_frame")}
    try:
        tool, etc = types.ModuleType("tool"), types.ModuleType("tool.etc")
        fe, fp = types.ModuleType("frontend"), types.ModuleType("frontend.process_frame")
        etc.image_tt = fp.image_tt = lambda image, device='cuda': "reference"
        tool.etc, fe.process_frame = etc, fp
        sys.modules.update({"tool": tool, "tool.etc": etc, "frontend": fe, "frontend.process_frame": fp})
        hook = spb.install_image_tt()
        assert etc.image_tt is hook and fp.image_tt is hook and hook.__wrapped__ is frames.image_tt
        assert hook(np.zeros((4, 4, 3), np.uint8), 'cpu') == "reference"     # the SAM pre-resize keeps the reference's path
        import inspect
        assert list(inspect.signature(frames.image_tt).parameters) == ['image', 'device']      # tool/etc.py:37
        with pytest.raises(AssertionError):
            frames.image_tt(np.zeros((4, 4, 3), np.float32))          # 8-bit HWC frames only
        with pytest.raises(RuntimeError):
            frames.image_tt(np.zeros((4, 4, 3), np.uint8), device="cpu")   # no CPU fallback
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_signatures_match_reference_call_surface():
    import inspect
    from super_primitive_b200 import dense_optim as do, dense_optim_batch as dob, depth_render as dr, depth_init as di
    sig = lambda f: list(inspect.signature(f).parameters)   # noqa: E731
    assert sig(do.photomeric_cost) == ['src_keyframe', 'trg_keyframe', 'src_keypoint_logdepth', 'pose', 'cost_config',
                                       'affine_comp']
    assert sig(do.photomeric_cost_precomputed) == ['src_precomputed', 'trg_keyframe', 'pose', 'cost_config',
                                                   'affine_comp']
    assert sig(dob.photomeric_cost_batch) == ['src_keyframe', 'trg_images', 'trg_Ks', 'src_keypoint_logdepth', 'poses',
                                              'cost_config', 'affine_comp']
    assert sig(do.unproject_kf) == ['kf', 'keypoint_logdepth', 'jacobian']
    assert sig(do.unproject_kf_to_depths) == ['kf', 'keypoint_logdepth']
    assert sig(dr.estimate_depth_kf_native) == ['kf', 'kf_logdepth', 'pose', 'mean']
    assert sig(di.segment_based_depth_reinit) == ['estimated_depth', 'kf', 'mode', 'return_info']


def test_cpu_tensors_are_refused():
    from super_primitive_b200 import dense_optim as do, synthetic as syn
    src, trg, k0, pose0 = syn.two_frame_problem(16, 24, 2)
    with pytest.raises(RuntimeError):
        do.photomeric_cost(src, trg, k0, pose0, {'mode': 'colour', 'collect_stats': 0})
    # mode handling mirrors the reference (core/cost_utils.py:4-19, core/dense_optim.py:228-236): outside 'colour' the
    # config must carry normal_loss / normal_weight (KeyError), the image must have the mode's channel count, and
    # 'norm_kappa' (no colour term: the reference's residual is the constant 0.0) is rejected
    with pytest.raises(KeyError):
        do.photomeric_cost(src, trg, k0, pose0, {'mode': 'colour_norm', 'collect_stats': 0})
    full = {'collect_stats': 0, 'normal_loss': 'lecrec', 'normal_weight': 0.1}
    with pytest.raises(AssertionError):
        do.photomeric_cost(src, trg, k0, pose0, dict(full, mode='colour_norm'))          # 3 channels, 6 needed
    with pytest.raises(NotImplementedError):
        do.photomeric_cost(src, trg, k0, pose0, dict(full, mode='norm_kappa'))
    with pytest.raises(ValueError):
        do.photomeric_cost(src, trg, k0, pose0, dict(full, mode='depth'))


def test_lazy_result_materialises_once():
    from super_primitive_b200.dense_optim import LazyResult
    calls = []

    def produce():
        calls.append(1)
        return {'a': 1, 'median_depth': None}

    r = LazyResult(torch.zeros(1), produce)
    assert r['residual'].shape == (1,) and calls == []
    assert 'residual' in r and calls == []
    assert r['a'] == 1 and calls == [1]
    assert set(r.keys()) == {'residual', 'a', 'median_depth'} and len(r) == 3 and calls == [1]
    assert r.get('median_depth', 5) is None
    assert dict(r.items())['a'] == 1 and calls == [1]


def test_synthetic_generator_contract():
    from super_primitive_b200 import synthetic as syn
    for kind in ("strips", "overlap", "rects"):
        kf = syn.make_keyframe(48, 64, 6, kind=kind, seed=2)
        assert kf.image.shape == (3, 48, 64) and kf.keypoint_regions.shape == (6, 48, 64)
        assert kf.keypoint_regions.dtype == torch.bool and kf.logdepth_perseg.shape == (6, 48, 64)
        assert torch.all(kf.logdepth_perseg[~kf.keypoint_regions] == 0)
        rc = (0.5 * (torch.tensor([48., 64.]) - 1) * (kf.keypoints + 1)).round().long()
        assert all(bool(kf.keypoint_regions[b, rc[b, 0], rc[b, 1]]) for b in range(6)), kind
        assert kf.keypoint_regions.flatten(1).any(1).all()
    strips = syn.make_keyframe(48, 64, 8, kind="strips").keypoint_regions
    assert int(strips.sum()) == 48 * 64 and int(strips.sum(0).max()) == 1     # exact partition
    a = syn.make_keyframe(32, 40, 4, kind="rects", seed=9, noise=0.02)
    b = syn.make_keyframe(32, 40, 4, kind="rects", seed=9, noise=0.02)
    assert torch.equal(a.image, b.image) and torch.equal(a.keypoint_regions, b.keypoint_regions)


def test_pyramid_matches_reference_semantics():
    """3x3 [1 2 1]^2/16 blur with reflect padding then [::2, ::2]; geometry untouched (geo_down=False)."""
    from super_primitive_b200 import synthetic as syn
    kf = syn.make_keyframe(32, 48, 3, kind="strips", noise=0.01, seed=1)
    pyr = syn.keyframe_pyramid(kf, 0, 3)
    assert [tuple(p.image.shape[1:]) for p in pyr] == [(8, 12), (16, 24), (32, 48)]
    assert all(p.logdepth_perseg is kf.logdepth_perseg and p.keypoint_regions is kf.keypoint_regions for p in pyr)
    img = kf.image.double().numpy()
    pad = np.pad(img, ((0, 0), (1, 1), (1, 1)), mode="reflect")
    ker = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]) / 16.0
    blur = sum(ker[i, j] * pad[:, i:i + 32, j:j + 48] for i in range(3) for j in range(3))
    assert np.allclose(pyr[1].image.numpy(), blur[:, ::2, ::2], atol=1e-6)
    assert torch.allclose(pyr[0].K_img[0, 0], kf.K[0, 0] * 0.25)


def test_precomputed_dict_drops_its_fast_path_handle_when_edited():
    """`unproject_kf` hands back a dict subclass carrying a private handle on the compact geometry (so that
    `photomeric_cost_precomputed` can run the fused kernel); any edit invalidates it, and it pickles as a plain dict."""
    import copy
    import pickle
    from super_primitive_b200.dense_optim import _Precomputed
    base = {'src_pts': torch.zeros(4, 3), 'src_pixels': torch.zeros(1, 3, 4), 'src_valid_mask': torch.ones(1, 4, dtype=torch.bool),
            'segm_ids': torch.zeros(4, dtype=torch.int64), 'spatial_size': (2, 2)}
    for edit in (lambda d: d.__setitem__('src_pts', d['src_pts'] * 2), lambda d: d.update(extra=1),
                 lambda d: d.pop('segm_ids'), lambda d: d.setdefault('x', 0), lambda d: d.__delitem__('segm_ids')):
        d = _Precomputed(base)
        d._spb = ("geom", "level", "k", d['src_pts'], d['src_pixels'])
        assert d._spb is not None and d['spatial_size'] == (2, 2)          # reads keep the handle
        edit(d)
        assert d._spb is None
    d = _Precomputed(base)
    d._spb = ("geom", "level", "k", d['src_pts'], d['src_pixels'])
    for clone in (pickle.loads(pickle.dumps(d)), copy.deepcopy(d)):
        assert type(clone) is dict and set(clone) == set(base)
    assert dict(d) == base and type(dict(d)) is dict


def test_mode_table_matches_the_reference_split():
    """core/cost_utils.py:4-19: channel counts the modes insist on."""
    from super_primitive_b200.dense_optim import _mode_channels
    _mode_channels('colour', 3)
    _mode_channels('colour', 7)                 # 'colour' slices the first three channels of anything
    _mode_channels('colour_norm', 6)
    _mode_channels('colour_norm_kappa', 7)
    for mode, c in (('colour_norm', 3), ('colour_norm', 7), ('colour_norm_kappa', 6)):
        with pytest.raises(AssertionError):
            _mode_channels(mode, c)
