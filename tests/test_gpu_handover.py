"""GPU: keyframe hand-over from the frontend without dense tensors (super_primitive_b200/handover.py)
against tests/golden/handover.npz -- the reference's own `put_keypoints_back` inside the last lines of
`FrontProcessorNew.process_to_kf` (tests/golden/make_golden_handover.py)."""
import os

import numpy as np
import pytest
import torch

from tests.common import CFG0, assert_close, to_np

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["half", "odd", "same"])
def test_compact_geometry_from_integrated_depth_equals_the_dense_route(name):
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.geometry import CompactGeometry
    from super_primitive_b200.handover import geometry_from_frontend
    z = np.load(os.path.join(HERE, "golden", "handover.npz"))
    H, W = (int(v) for v in z[f"{name}_size"])
    K = syn.pinhole(H, W).cuda()
    g, kp_norm, good = geometry_from_frontend(torch.from_numpy(z[f"{name}_depth"]).cuda(),
                                              torch.from_numpy(z[f"{name}_kps"]).cuda(), K, (H, W))
    assert np.array_equal(to_np(good), z[f"{name}_good"])                       # empty segments dropped like the reference
    assert np.array_equal(to_np(kp_norm), z[f"{name}_keypoints"])               # keypoints snapped to the same pixels
    # the compact arrays equal what the dense route builds from the reference's (masks, logdepth, keypoints)
    dense = CompactGeometry(torch.from_numpy(z[f"{name}_masks"]).cuda(), torch.from_numpy(z[f"{name}_logdepth"]).cuda(),
                            torch.from_numpy(z[f"{name}_keypoints"]).cuda(), K)
    assert (g.N, g.P, g.P_pad, g.n_tiles) == (dense.N, dense.P, dense.P_pad, dense.n_tiles)
    assert torch.equal(g.uv, dense.uv) and torch.equal(g.tiles, dense.tiles) and torch.equal(g.seg_tile, dense.seg_tile)
    assert torch.equal(g.kp_rc, dense.kp_rc)
    np.testing.assert_allclose(to_np(g.logd), to_np(dense.logd), rtol=3e-7, atol=1e-7)      # logf on the device vs torch CPU
    np.testing.assert_allclose(to_np(g.seg_lkp), to_np(dense.seg_lkp), rtol=3e-7, atol=1e-7)


def test_compact_keyframe_runs_the_alignment_path_like_a_dense_keyframe():
    from super_primitive_b200 import dense_optim as do, depth_render, synthetic as syn
    from super_primitive_b200.handover import CompactKeyFrame, keyframe_from_frontend
    from super_primitive_b200.keyframe import KeyFrame
    from super_primitive_b200.pyramid import keyframe_pyramid
    from super_primitive_b200.geometry import geometry_of
    z = np.load(os.path.join(HERE, "golden", "handover.npz"))
    H, W = (int(v) for v in z["half_size"])
    K = syn.pinhole(H, W).cuda()
    img = syn.sinus_image(H, W, noise=0.01, seed=1).cuda()
    trg = KeyFrame(syn.sinus_image(H, W, shift=(1.5, 0.5), noise=0.01, seed=2).cuda(), K)
    ckf = keyframe_from_frontend(img, K, torch.from_numpy(z["half_depth"]).cuda(), torch.from_numpy(z["half_kps"]).cuda(), (H, W))
    assert isinstance(ckf, CompactKeyFrame) and ckf._dense is None and ckf.geo_spatial_dim() == (H, W)
    dkf = KeyFrame(img, K, torch.from_numpy(z["half_logdepth"]).cuda(), torch.from_numpy(z["half_keypoints"]).cuda(),
                   torch.from_numpy(z["half_masks"]).cuda())
    N = ckf.num_segments()
    k0 = torch.full((N,), float(np.log(2.0)), device="cuda")
    pose0 = syn.small_pose(0.02, 0.004, -0.003, 0.003, -0.002, 0.0015).cuda()
    outs = []
    for kf in (ckf, dkf):
        k, pose = k0.clone().requires_grad_(True), pose0.clone().requires_grad_(True)
        out = do.photomeric_cost(kf, trg, k, pose, CFG0)
        out['residual'].mean().backward()
        outs.append((out['residual'].detach(), k.grad, pose.grad))
    for a, b, what in zip(outs[0], outs[1], ("residual", "d/dk", "d/dpose")):
        assert_close(to_np(a), to_np(b), 2e-6, what)
    assert ckf._dense is None                                   # the cost never touched a dense tensor
    # pyramid levels share the compact geometry; depth render and re-lifting work
    levels = keyframe_pyramid(ckf, 0, 2)
    assert all(geometry_of(lv) is ckf._spb_geometry for lv in levels) and ckf._dense is None
    d_c = depth_render.estimate_depth_kf_native(ckf, k0, pose0)
    d_d = depth_render.estimate_depth_kf_native(dkf, k0, pose0)
    assert_close(to_np(d_c), to_np(d_d), 2e-6, "depth render")
    # the dense tensors, when somebody does ask, are the reference's
    assert np.array_equal(to_np(ckf.keypoint_regions), z["half_masks"])
    np.testing.assert_allclose(to_np(ckf.logdepth_perseg), z["half_logdepth"], rtol=3e-7, atol=1e-7)
