"""GPU: the other BASELINE.json configurations as parity / property cases.

C3-like  288x224, ~120 SAM-like segments (ragged, overlapping, many partial tiles), 3-level pyramid,
         tracking (precomputed) and mapping (batch) paths
C2/C5    full-size shapes through size-independent properties: the compact-geometry kernel and the
         pre-lifted-points kernel must agree, results are bit-reproducible, batch == stacked singles
"""
import numpy as np
import pytest
import torch

from tests.common import CFG0, assert_close, rel_err, to_np

pytestmark = pytest.mark.gpu


def _leaf(t):
    return t.clone().requires_grad_(True)


def _f64(kf):
    from super_primitive_b200.keyframe import KeyFrame
    c = lambda t: None if t is None else (t.double() if t.is_floating_point() else t)   # noqa: E731
    return KeyFrame(c(kf.image), c(kf.K), c(kf.logdepth_perseg), c(kf.keypoints), kf.keypoint_regions, c(kf.K_img))


def _bar(e_gpu, e_ref, what):
    assert e_gpu <= max(1e-4, 2.0 * e_ref) and e_gpu <= 1e-3, f"{what}: GPU vs float64 {e_gpu:.2e}, reference {e_ref:.2e}"


def test_c3_many_small_segments_tracking_and_mapping():
    from oracle import ref_port as port
    from super_primitive_b200 import dense_optim as do, dense_optim_batch as dob, synthetic as syn
    H, W, N, B = 224, 288, 120, 5
    src0 = syn.make_keyframe(H, W, N, kind="rects", seed=31, noise=0.02)
    trgs0 = [syn.make_keyframe(H, W, N, shift=(1.0 + 0.4 * j, 0.5 - 0.2 * j), noise=0.02, seed=40 + j, supporting=True)
             for j in range(B)]
    k0 = float(np.log(2.0)) + 0.1 * torch.randn(N, generator=torch.Generator().manual_seed(2))
    poses0 = torch.stack([syn.small_pose(0.02 - 0.004 * j, 0.003 * j, -0.002 * j, 0.004, -0.003 * j, 0.002)
                          for j in range(B)])
    aff_s0 = torch.tensor([0.03, -0.01])
    aff_t0 = torch.tensor([[0.01 * j, 0.02 - 0.01 * j] for j in range(B)])
    spyr = syn.keyframe_pyramid(src0, 0, 3)
    tpyrs = [syn.keyframe_pyramid(t, 0, 3) for t in trgs0]
    for lvl in (0, 2):
        s = spyr[lvl]
        # ---- mapping: one source against B targets, affine on
        imgs = torch.stack([tp[lvl].image for tp in tpyrs])
        Ks = torch.stack([t.K for t in trgs0])
        k64, p64 = _leaf(k0.double()), _leaf(poses0.double())
        as64, at64 = _leaf(aff_s0.double()), _leaf(aff_t0.double())
        r64 = port.cost_batch(_f64(s), imgs.double(), Ks.double(), k64, p64, CFG0, (as64, at64))
        r64['residual'].mean().backward()
        k32, p32 = _leaf(k0), _leaf(poses0)
        r32 = port.cost_batch(s, imgs, Ks, k32, p32, CFG0, (aff_s0, aff_t0))
        r32['residual'].mean().backward()
        kg, pg = _leaf(k0.cuda()), _leaf(poses0.cuda())
        asg, atg = _leaf(aff_s0.cuda()), _leaf(aff_t0.cuda())
        out = dob.photomeric_cost_batch(s.to("cuda"), imgs.cuda(), Ks.cuda(), kg, pg, CFG0, (asg, atg))
        out['residual'].mean().backward()
        assert_close(to_np(out['residual']), to_np(r64['residual']), 2e-5, f"batch residual L{lvl}")
        _bar(rel_err(to_np(kg.grad), to_np(k64.grad)), rel_err(to_np(k32.grad), to_np(k64.grad)), "batch g_k")
        _bar(rel_err(to_np(pg.grad), to_np(p64.grad)), rel_err(to_np(p32.grad), to_np(p64.grad)), "batch g_poses")
        assert_close(to_np(asg.grad), to_np(as64.grad), 1e-3, "g_aff_src")
        assert_close(to_np(atg.grad), to_np(at64.grad), 1e-3, "g_aff_trg")
        # ---- tracking: pre-lifted points against one target
        with torch.no_grad():
            pre = do.unproject_kf(s.to("cuda"), k0.cuda())
            pre_ref = port.lift_keyframe(s, k0)
        assert pre['src_pts'].shape == pre_ref['src_pts'].shape
        pose_g, pose_c = _leaf(poses0[0].cuda()), _leaf(poses0[0])
        og = do.photomeric_cost_precomputed(pre, tpyrs[0][lvl].to("cuda"), pose_g, CFG0)
        og['residual'].mean().backward()
        oc = port.cost_precomputed(pre_ref, tpyrs[0][lvl], pose_c, CFG0)
        oc['residual'].mean().backward()
        assert_close(to_np(og['residual']), to_np(oc['residual']), 2e-5, f"tracking residual L{lvl}")
        assert_close(to_np(pose_g.grad), to_np(pose_c.grad), 1e-3, "tracking g_pose")


def test_more_targets_than_one_launch_packs():
    """B = 20 > 16 pairs per inline launch: the host splits the batch; results equal 20 single calls."""
    from super_primitive_b200 import dense_optim as do, dense_optim_batch as dob, synthetic as syn
    H, W, N, B = 64, 96, 6, 20
    src = syn.make_keyframe(H, W, N, kind="overlap", seed=3, noise=0.01).to("cuda")
    imgs = torch.stack([syn.sinus_image(H, W, shift=(0.5 + 0.1 * j, 0.2 * (j % 3)), noise=0.01, seed=j) for j in range(B)]).cuda()
    Ks = src.K[None].repeat(B, 1, 1)
    poses = torch.stack([syn.small_pose(0.01 + 0.001 * j, 0.002, 0.0, 0.001 * j, 0.002, -0.001) for j in range(B)]).cuda()
    k = torch.full((N,), float(np.log(2.0)), device="cuda")
    kb, pb = _leaf(k), _leaf(poses)
    out = dob.photomeric_cost_batch(src, imgs, Ks, kb, pb, CFG0)
    out['residual'].sum().backward()
    from super_primitive_b200.keyframe import KeyFrame
    gk = torch.zeros_like(k)
    for j in range(B):
        kj, pj = _leaf(k), _leaf(poses[j])
        # the batch path uses the 1e-6 depth threshold, the single path 1e-7: identical for these depths
        r = do.photomeric_cost(src, KeyFrame(imgs[j], src.K), kj, pj, CFG0)
        r['residual'].sum().backward()
        assert_close(to_np(out['residual'][j:j + 1]), to_np(r['residual']), 1e-6, f"residual {j}")
        assert_close(to_np(pb.grad[j]), to_np(pj.grad), 1e-5, f"pose grad {j}")
        gk += kj.grad
    assert_close(to_np(kb.grad), to_np(gk), 1e-5, "summed depth gradient")


@pytest.mark.parametrize("H,W,N", [(480, 640, 64), (768, 1024, 256)])
def test_full_size_consistency_properties(H, W, N):
    """BASELINE full sizes (C2, C5): (i) bit-reproducible, (ii) the compact kernel and the pre-lifted-points
    kernel (two independent code paths) agree on cost and pose gradient, (iii) non-trivial numbers."""
    from super_primitive_b200 import dense_optim as do, synthetic as syn
    src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind="overlap", seed=7, noise=0.01)
    src, trg, k0, pose0 = src.to("cuda"), trg.to("cuda"), k0.cuda(), pose0.cuda()
    runs = []
    for _ in range(2):
        k, pose = _leaf(k0), _leaf(pose0)
        out = do.photomeric_cost(src, trg, k, pose, CFG0)
        out['residual'].mean().backward()
        runs.append((out['residual'].detach().clone(), k.grad.clone(), pose.grad.clone()))
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b), "results must be bit-reproducible (deterministic reductions)"
    res, gk, gp = runs[0]
    assert 1e-3 < float(res) < 1.0 and float(gk.abs().max()) > 0 and float(gp.abs().max()) > 0
    with torch.no_grad():
        pre = do.unproject_kf(src, k0)
    P = pre['src_pts'].shape[0]
    assert P == int(src.keypoint_regions.sum())
    pose = _leaf(pose0)
    out2 = do.photomeric_cost_precomputed(dict(pre), trg, pose, CFG0)      # a plain dict: the generic point-list kernel
    out2['residual'].mean().backward()
    assert_close(to_np(out2['residual']), to_np(res), 2e-5, "compact vs points kernel: cost")
    assert_close(to_np(pose.grad), to_np(gp), 2e-4, "compact vs points kernel: pose gradient")
    # the dict as unproject_kf returned it is served by the fused compact kernel at the lifted seeds: same launch
    # as photomeric_cost => bit-identical
    pose3 = _leaf(pose0)
    out3 = do.photomeric_cost_precomputed(pre, trg, pose3, CFG0)
    out3['residual'].mean().backward()
    assert torch.equal(out3['residual'], res) and torch.equal(pose3.grad, gp)
    pre['src_pts'] = pre['src_pts'] * 1.0                                  # an edited dict falls back to the generic kernel
    assert getattr(pre, "_spb", None) is None
    out4 = do.photomeric_cost_precomputed(pre, trg, _leaf(pose0), CFG0)
    assert_close(to_np(out4['residual']), to_np(res), 2e-5, "edited dict: generic kernel")


def test_full_size_gn_converges_c2():
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    src, trg, k0, pose0 = syn.two_frame_problem(480, 640, 64, kind="overlap", seed=1, noise=0.01)
    src, trg = src.to("cuda"), trg.to("cuda")
    batch = AlignmentBatch([make_problem(src, trg.image, trg.K, pose0.cuda(), k0.cuda())])
    batch.gn_accumulate()
    c0 = float(batch.costs()[0])
    batch.run_gn(25)
    torch.cuda.synchronize()
    c1 = float(batch.lm_state[0, 1]) / (3 * float(batch.pts_per_problem[0]))
    assert c1 < 0.5 * c0, (c0, c1)
    assert float(batch.lm_state[0, 3]) >= 3


def test_estimate_depth_diff_points_against_oracle():
    """core/ops.py:59-96 on an arbitrary point cloud (general pose => no pixel-boundary degeneracy)."""
    from oracle import ref_port as port
    from super_primitive_b200 import ops, synthetic as syn
    src = syn.make_keyframe(96, 128, 10, kind="rects", seed=12)
    k = torch.full((10,), float(np.log(2.0)))
    pose = syn.small_pose(0.03, -0.02, 0.01, 0.02, -0.015, 0.01)
    with torch.no_grad():
        pts = port.rigid(port.lift_keyframe(src, k)['src_pts'], pose)
    for mean in (False, True):
        ref_img, ref_valid = port.splat_depth(pts, src.K, (96, 128), mean=mean)
        img, valid = ops.estimate_depth_diff(pts.cuda(), src.K.cuda(), (96, 128), mean=mean)
        assert img.shape == ref_img.shape and valid.shape == ref_valid.shape
        assert (to_np(valid) != to_np(ref_valid)).mean() < 1e-3
        bad = np.abs(to_np(img) - to_np(ref_img)) > 1e-4 * np.maximum(np.abs(to_np(ref_img)), 1e-3)
        assert bad.mean() < 2e-3, f"mean={mean}: {bad.sum()} differing pixels"


def test_keyframe_pyramid_matches_reference():
    """image/keyframe.py:77-148 (geo_down=False): the CUDA blur/decimate against the reference's own pyramid
    frozen in tests/golden/pyramid_odd.npz (odd sizes 45x70)."""
    import os
    from tests.common import GOLDEN_DIR
    from super_primitive_b200.keyframe import KeyFrame
    from super_primitive_b200.pyramid import keyframe_pyramid
    z = np.load(os.path.join(GOLDEN_DIR, "pyramid_odd.npz"))
    kf = KeyFrame(torch.from_numpy(z["image"]).cuda(), torch.from_numpy(z["K"]).cuda())
    for a, b in [(0, 3), (1, 4), (0, 1)]:
        levels = keyframe_pyramid(kf, a, b)
        assert torch.is_grad_enabled()
        assert len(levels) == int(z[f"p{a}{b}_n"])
        for i, lv in enumerate(levels):
            ref = z[f"p{a}{b}_L{i}_image"]
            assert tuple(lv.image.shape) == ref.shape
            assert_close(to_np(lv.image), ref, 1e-6, f"pyramid {a}{b} level {i}")
            assert_close(to_np(lv.K_img), z[f"p{a}{b}_L{i}_K_img"], 1e-7, "K_img")
            assert lv.is_supporting()


def test_depth_completion_average_render():
    """BASELINE config 4 tail: render_depth_avg (dense, in place) and the fused compact variant vs the oracle."""
    from oracle import ref_port as port
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.depth_completion import render_depth_avg, render_segments_avg
    kf = syn.make_keyframe(120, 160, 30, kind="rects", seed=8)
    k = float(np.log(2.0)) + 0.2 * torch.randn(30, generator=torch.Generator().manual_seed(4))
    vis = torch.ones(30, dtype=torch.bool)
    vis[[3, 11, 12]] = False
    with torch.no_grad():
        ref_avg, ref_inv = port.completion_render(kf, k, vis)
        dense = port.dense_depths(kf, k)
        dense[kf.keypoint_regions == 0] = -1
        dense = dense[vis].contiguous()
    d_gpu = dense.clone().cuda()
    avg, inv = render_depth_avg(d_gpu)
    assert np.array_equal(to_np(inv), to_np(ref_inv))
    assert_close(to_np(avg), to_np(ref_avg), 1e-6, "dense average")
    assert float(d_gpu.min()) >= 0.0                      # negatives zeroed in place like the reference
    avg2, inv2 = render_segments_avg(kf.to("cuda"), k.cuda(), vis.cuda())
    assert np.array_equal(to_np(inv2), to_np(ref_inv))
    assert_close(to_np(avg2), to_np(ref_avg), 1e-5, "compact average")
    # overlapping segments meet in a pixel in any order: the fixed-point accumulation makes the render bit-reproducible
    # (30 overlapping rectangles: most pixels receive several contributions)
    for _ in range(3):
        again, _ = render_segments_avg(kf.to("cuda"), k.cuda(), vis.cuda())
        assert torch.equal(again, avg2)


def test_depth_completion_against_reference_golden():
    import os
    from tests.common import GOLDEN_DIR
    from super_primitive_b200.depth_completion import render_segments_avg
    from super_primitive_b200.keyframe import KeyFrame
    z = np.load(os.path.join(GOLDEN_DIR, "completion.npz"))
    t = lambda k: torch.from_numpy(z[k]).cuda()   # noqa: E731
    kf = KeyFrame(t("src_image"), t("src_K"), t("src_logdepth"), t("src_keypoints"), t("src_regions"))
    avg, inv = render_segments_avg(kf, t("k"), t("visible"))
    assert np.array_equal(to_np(inv), z["invalid"])
    assert_close(to_np(avg), z["avg"], 1e-5, "average render vs reference")


def test_edge_cases_empty_segment_many_segments_single_pixel():
    """Edge cases the reference's dense formulation handles implicitly: a segment with an EMPTY mask, single-pixel
    segments, and more segments than the kernel's shared-memory shift cache (N > 512 -> global fallback)."""
    from oracle import ref_port as port
    from super_primitive_b200 import dense_optim as do, synthetic as syn
    from super_primitive_b200.depth_init import segment_based_depth_reinit
    from super_primitive_b200.keyframe import KeyFrame
    H, W, N = 48, 64, 600
    g = torch.Generator().manual_seed(5)
    masks = torch.zeros((N, H, W), dtype=torch.bool)
    kp = torch.zeros((N, 2), dtype=torch.int64)
    for b in range(N):
        r, c = int(torch.randint(2, H - 2, (1,), generator=g)), int(torch.randint(2, W - 2, (1,), generator=g))
        kp[b, 0], kp[b, 1] = r, c
        if b % 7 == 0:
            masks[b, r, c] = True                                  # single-pixel segment
        elif b != 13:                                              # segment 13 stays EMPTY
            masks[b, r - 1:r + 2, c - 2:c + 2] = True
    x = (torch.arange(W, dtype=torch.float64) / W)[None, None, :].expand(N, H, W)
    logd = ((0.1 * x) * masks).float()
    src = KeyFrame(syn.sinus_image(H, W, noise=0.01, seed=1), syn.pinhole(H, W), logd, syn.normalise_rc(kp, (H, W)), masks)
    trg = KeyFrame(syn.sinus_image(H, W, shift=(1.5, 0.5), noise=0.01, seed=2), syn.pinhole(H, W))
    k0 = float(np.log(2.0)) + 0.05 * torch.randn(N, generator=g)
    pose0 = syn.small_pose(0.02, 0.004, -0.003, 0.003, -0.002, 0.0015)
    k, pose = _leaf(k0), _leaf(pose0)
    ref = port.cost_single(src, trg, k, pose, CFG0)
    ref['residual'].mean().backward()
    kg, pg = _leaf(k0.cuda()), _leaf(pose0.cuda())
    out = do.photomeric_cost(src.to("cuda"), trg.to("cuda"), kg, pg, CFG0)
    out['residual'].mean().backward()
    assert_close(to_np(out['residual']), to_np(ref['residual']), 2e-5, "residual")
    assert_close(to_np(kg.grad), to_np(k.grad), 1e-3, "g_k")
    assert float(kg.grad[13]) == 0.0 and float(k.grad[13]) == 0.0    # empty segment: no gradient
    assert_close(to_np(pg.grad), to_np(pose.grad), 1e-3, "g_pose")
    # re-initialisation with an empty segment: it is 'invisible' and takes the median of the visible ones
    est = 1.5 + torch.rand((H, W), generator=g)
    kk_ref, vis_ref = port.segment_median_reinit(est.clone(), src, 'median')
    kk, vis = segment_based_depth_reinit(est.clone().cuda(), src.to("cuda"), 'median', return_info=True)
    torch.set_grad_enabled(True)
    assert np.array_equal(to_np(vis), to_np(vis_ref)) and not bool(vis[13])
    assert_close(to_np(kk), to_np(kk_ref), 1e-5, "reinit with empty segment")


def test_batched_depth_completion_equals_the_per_frame_path():
    """`depth_completion.complete_batch` (one read-back of the point counts for all frames) against
    `segment_based_depth_reinit` + `render_segments_avg` frame by frame: bit-identical maps, seeds and visibility."""
    from super_primitive_b200 import depth_completion as dc, depth_init, synthetic as syn
    from super_primitive_b200.geometry import clear_caches
    kfs, sparse = [], []
    for i in range(5):
        kf = syn.make_keyframe(64 + 16 * (i % 2), 96, 7 + i, kind="rects", seed=20 + i).to("cuda")
        g = torch.Generator().manual_seed(i)
        H, W = kf.keypoint_regions.shape[1:]
        d = torch.zeros(H, W)
        idx = torch.randperm(H * W, generator=g)[:300]
        d.view(-1)[idx] = 1.0 + torch.rand(300, generator=g)
        kfs.append(kf)
        sparse.append(d.cuda())
    want = []
    for kf, sp in zip(kfs, sparse):
        k, vis = depth_init.segment_based_depth_reinit(sp.clone(), kf, 'median', return_info=True)
        depth, inv = dc.render_segments_avg(kf, k, vis)
        want.append((depth, inv, k, vis))
    torch.set_grad_enabled(True)
    clear_caches()
    got = dc.complete_batch(kfs, [s.clone() for s in sparse], 'median')
    assert len(got) == 5
    for (d0, i0, k0, v0), (d1, i1, k1, v1) in zip(want, got):
        assert torch.equal(d0, d1) and torch.equal(i0, i1) and torch.equal(k0, k1) and torch.equal(v0, v1)
    with pytest.raises(IndexError):                      # a frame without any measurement: the reference fails too
        dc.complete_batch(kfs[:2], [sparse[0].clone(), torch.zeros_like(sparse[1])], 'median')
