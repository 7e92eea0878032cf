"""CPU: the oracle (torch port + float64 closed form) against the frozen reference outputs."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_port as port
from tests.common import CFG0, CFG2, FULL_CASES, STATS_CASES, Golden, assert_close, to_np


def _grads(fn, leaves):
    out = fn()
    out['residual'].mean().backward()
    return out, [None if l is None else l.grad for l in leaves]


@pytest.mark.parametrize("case", FULL_CASES)
def test_port_reproduces_reference(case):
    g = Golden(case)
    for lvl in range(g.n_levels):
        tag = f"L{lvl}_"
        k = g.k().requires_grad_(True)
        pose = g.poses()[0].clone().requires_grad_(True)
        aff = g.affine(0)
        if aff is not None:
            aff = tuple(a.clone().requires_grad_(True) for a in aff)
        out, gr = _grads(lambda: port.cost_single(g.src(lvl), g.trg(lvl), k, pose, CFG0, aff), [k, pose])
        assert_close(to_np(out['residual']), g.z[tag + "single_residual"], 1e-6, "residual")
        assert_close(to_np(gr[0]), g.z[tag + "single_g_k"], 1e-5, "g_k")
        assert_close(to_np(gr[1]), g.z[tag + "single_g_pose"], 1e-5, "g_pose")
        # batch
        k = g.k().requires_grad_(True)
        poses = g.poses().clone().requires_grad_(True)
        aff = g.affine()
        out = port.cost_batch(g.src(lvl), g.trg_images(lvl), g.trg_Ks(), k, poses, CFG0, aff)
        out['residual'].mean().backward()
        assert_close(to_np(out['residual']), g.z[tag + "batch_residual"], 1e-6, "batch residual")
        assert_close(to_np(k.grad), g.z[tag + "batch_g_k"], 1e-5, "batch g_k")
        assert_close(to_np(poses.grad), g.z[tag + "batch_g_poses"], 1e-5, "batch g_poses")


@pytest.mark.parametrize("case", STATS_CASES)
def test_port_stats_and_geometry(case):
    g = Golden(case)
    lvl = g.n_levels - 1
    out = port.cost_single(g.src(lvl), g.trg(lvl), g.k(), g.poses()[0], CFG2, g.affine(0))
    for key in ['segm_ids', 'src_valid_mask', 'trg_valid_mask', 'full_mask']:
        assert np.array_equal(to_np(out[key]), g.z[f"L{lvl}_single_{key}"]), key
    for key in ['src_pts', 'src_in_trg_pts', 'residual_raw', 'src_in_trg_pixels', 'src_in_trg_keypoints']:
        assert_close(to_np(out[key]), g.z[f"L{lvl}_single_{key}"], 1e-6, key)
    with torch.no_grad():
        assert_close(to_np(port.dense_depths(g.src(lvl), g.k())), g.z["dense_depths"], 1e-6, "dense")
        for tag, pose, mean in [("render_id", None, False), ("render_pose", g.poses()[0], False),
                                ("render_mean", g.poses()[0], True)]:
            assert_close(to_np(port.render_keyframe_depth(g.src(lvl), g.k(), pose, mean)), g.z[tag], 1e-6, tag)
        est = g.t("reinit_est_depth")
        for mode in ("median", "mean"):
            kk, vis = port.segment_median_reinit(est, g.src(lvl), mode)
            assert_close(to_np(kk), g.z[f"reinit_{mode}"], 1e-6, mode)
            assert np.array_equal(to_np(vis), g.z["reinit_visible"])


@pytest.mark.parametrize("case", FULL_CASES)
@pytest.mark.parametrize("batch", [False, True])
def test_closed_form_matches_reference(case, batch):
    """float64 analytic cost + gradients vs the reference's float32 autograd."""
    g = Golden(case)
    geo = cf.compact_geometry(g.z["src_regions"], g.z["src_logdepth"], g.z["src_keypoints"])
    for lvl in range(g.n_levels):
        tag = f"L{lvl}_"
        simg = g.z[tag + "src_image"]
        timgs = g.z[tag + "trg_images"]
        js = range(g.B) if batch else [0]
        costs, gk = [], 0.0
        for j in js:
            aff = None
            if g.with_affine:
                aff = (g.z["aff_src"], g.z["aff_trg"][j])
            r = cf.evaluate(geo, simg, timgs[j], g.z["src_K"], g.z["src_K"], g.z["k"], g.z["poses"][j],
                            aff, batch_thresholds=batch)
            costs.append(r["cost"])
            scale = 1.0 / len(list(js))
            gk = gk + r["g_k"] * scale
            ref_gp = g.z[tag + ("batch_g_poses" if batch else "single_g_pose")]
            ref_gp = ref_gp[j] if batch else ref_gp
            assert_close(r["g_pose"] * scale, ref_gp, 2e-4, f"g_pose[{j}]")
            if g.with_affine:
                ref_at = g.z[tag + ("batch_g_aff_trg" if batch else "single_g_aff_trg")]
                ref_at = ref_at[j] if batch else ref_at
                assert_close(r["g_aff_trg"] * scale, ref_at, 2e-4, "g_aff_trg")
        assert_close(np.array(costs), g.z[tag + ("batch_residual" if batch else "single_residual")], 2e-5, "cost")
        assert_close(gk, g.z[tag + ("batch_g_k" if batch else "single_g_k")], 2e-4, "g_k")


def test_closed_form_gn_blocks_are_consistent():
    """J^T W r from the GN blocks must equal the analytic L1 gradient when W = 1/|r| (IRLS),
    up to the 1/(3P) normalisation -- ties the (unpinned) GN extension to the pinned gradient."""
    g = Golden("tiny_rects")
    geo = cf.compact_geometry(g.z["src_regions"], g.z["src_logdepth"], g.z["src_keypoints"])
    lvl = g.n_levels - 1
    r = cf.evaluate(geo, g.z[f"L{lvl}_src_image"], g.z[f"L{lvl}_trg_images"][0], g.z["src_K"], g.z["src_K"],
                    g.z["k"], g.z["poses"][0], None, want_gn=True, irls_eps=1e-12)
    P3 = 3.0 * r["P"]
    assert_close(r["g_d"] / P3, r["g_k"], 1e-9, "depth block")
    # pose block is in the left tangent: g_tau = sum gY ; g_phi = sum Y x gY
    gp = r["g_pose"]
    assert_close(r["g_p"][:3] / P3, gp[:3, 3], 1e-9, "translation block")
    xi, dk = cf.lm_step(r["A"], r["B"], r["D"], r["g_p"], r["g_d"], 1e-3)
    assert np.all(np.isfinite(xi)) and np.all(np.isfinite(dk))


def test_synthetic_pyramid_reproduces_reference_pyramid():
    """synthetic.keyframe_pyramid (used to build every multi-level test input) against the reference's own
    keyframe_pyramid outputs, odd image size."""
    import os
    from tests.common import GOLDEN_DIR
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.keyframe import KeyFrame
    z = np.load(os.path.join(GOLDEN_DIR, "pyramid_odd.npz"))
    kf = KeyFrame(torch.from_numpy(z["image"]), torch.from_numpy(z["K"]))
    for a, b in [(0, 3), (1, 4), (0, 1)]:
        levels = syn.keyframe_pyramid(kf, a, b)
        assert len(levels) == int(z[f"p{a}{b}_n"])
        for i, lv in enumerate(levels):
            assert_close(to_np(lv.image), z[f"p{a}{b}_L{i}_image"], 1e-6, "image")
            assert_close(to_np(lv.K_img), z[f"p{a}{b}_L{i}_K_img"], 1e-7, "K_img")


def test_completion_render_reproduces_reference():
    """oracle.completion_render against the reference's render_depth_avg pipeline (completion.npz)."""
    import os
    from tests.common import GOLDEN_DIR
    from super_primitive_b200.keyframe import KeyFrame
    z = np.load(os.path.join(GOLDEN_DIR, "completion.npz"))
    t = lambda k: torch.from_numpy(z[k])   # noqa: E731
    kf = KeyFrame(t("src_image"), t("src_K"), t("src_logdepth"), t("src_keypoints"), t("src_regions"))
    with torch.no_grad():
        avg, inv = port.completion_render(kf, t("k"), t("visible"))
    assert np.array_equal(to_np(inv), z["invalid"])
    assert_close(to_np(avg), z["avg"], 1e-6, "average render")


def test_frontend_handover_oracle_reproduces_the_reference():
    """oracle/frontend_handover.py against tests/golden/handover.npz: the reference's own `put_keypoints_back` inside
    the last lines of `process_to_kf` (nearest resampling, threshold, keypoint snap, logarithm; empty segments dropped)."""
    import os
    from oracle import frontend_handover as port
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handover.npz"))
    for name in ("half", "odd", "same"):
        H, W = (int(v) for v in z[f"{name}_size"])
        kp, masks, logd, good = port.handover(torch.from_numpy(z[f"{name}_depth"]), torch.from_numpy(z[f"{name}_kps"]), (H, W))
        assert np.array_equal(kp.numpy(), z[f"{name}_keypoints"])
        assert np.array_equal(masks.numpy(), z[f"{name}_masks"])
        assert np.array_equal(logd.numpy(), z[f"{name}_logdepth"])
        assert np.array_equal(good.numpy(), z[f"{name}_good"])
        # the fixture exercises what it should: dropped segments, moved keypoints
        assert good.sum() < len(good)
        moved = np.abs(kp.numpy() - z[f"{name}_kps"][z[f"{name}_good"]]).max(axis=1) > 1e-6
        assert moved.any()


def test_fill_depth_oracle_reproduces_the_reference():
    """oracle/fill_oracle.py against tests/golden/fill_depth.npz: the reference's own `fill_depth`
    (depth_completion/fill_in_tools.py:5-7, scipy's Euclidean feature transform underneath), indices and filled maps bit
    for bit -- including scipy's choice among equidistant pixels and its answer for a map with no valid pixel."""
    import hashlib
    import os
    from oracle import fill_oracle as port
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fill_depth.npz"))
    names = sorted(k[:-len("_invalid")] for k in z.files if k.endswith("_invalid"))
    assert {"lattice", "all_invalid", "one_valid", "all_valid"} <= set(names)
    ties = 0
    for name in names:
        inv, depth = z[f"{name}_invalid"], z[f"{name}_depth"]
        ind = port.nearest_valid_indices(inv)
        assert np.array_equal(ind, z[f"{name}_indices"]), name
        assert np.array_equal(port.fill_depth(depth, inv), z[f"{name}_filled"]), name
        if name == "lattice":                               # the fixture does exercise ties
            vr, vc = np.nonzero(~inv)
            for r, c in ((2, 2), (2, 6), (6, 2)):
                d2 = (vr - r) ** 2 + (vc - c) ** 2
                ties += int((d2 == d2.min()).sum() > 1)
    assert ties == 3
    H, W = (int(v) for v in z["vga_shape"])
    inv = np.unpackbits(z["vga_invalid_bits"])[:H * W].reshape(H, W).astype(bool)
    ind = port.nearest_valid_indices(inv)
    assert hashlib.sha256(np.ascontiguousarray(ind).tobytes()).digest() == z["vga_indices_sha256"].tobytes()
