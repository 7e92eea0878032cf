"""GPU: the Gauss-Newton / LM extension (no reference counterpart -- SURVEY R1) against the float64
closed-form oracle, plus convergence properties of the device-resident LM loop."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from tests.common import Golden, assert_close, to_np

pytestmark = pytest.mark.gpu

# IRLS weights 1/max(|r|, eps) amplify float32 rounding of small residuals (d w / w = d r / r), so the
# float32 blocks are compared with the float64 oracle at 1e-3 (scale-relative), the cost at 2e-5.
GN_TOL = 1e-3


def _tri8(A8):
    out = np.zeros((8, 8))
    q = 0
    for r in range(8):
        for c in range(r, 8):
            out[r, c] = out[c, r] = A8[q]
            q += 1
    return out


def _batch_from_golden(g, lvl, with_affine, js=(0,)):
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    src = g.src(lvl)
    probs = []
    for j in js:
        aff = g.affine(j) if with_affine else None
        probs.append(make_problem(src, g.trg(lvl, j).image, g.trg(lvl, j).K, g.poses()[j], g.k(),
                                  aff_src=None if aff is None else aff[0], aff_trg=None if aff is None else aff[1]))
    return AlignmentBatch(probs, with_affine=with_affine, irls_eps=1e-3)


@pytest.mark.parametrize("case,with_affine", [("tiny_rects", False), ("tiny_rects", True), ("pyr3_rects", False)])
def test_normal_equations_match_closed_form(case, with_affine):
    g = Golden(case, "cuda")
    for lvl in range(g.n_levels):
        batch = _batch_from_golden(g, lvl, with_affine, js=range(g.B))
        batch.gn_accumulate()
        torch.cuda.synchronize()
        geo = cf.compact_geometry(g.z["src_regions"], g.z["src_logdepth"], g.z["src_keypoints"])
        for j in range(g.B):
            aff = (g.z["aff_src"], g.z["aff_trg"][j]) if with_affine else None
            r = cf.evaluate(geo, g.z[f"L{lvl}_src_image"], g.z[f"L{lvl}_trg_images"][j], g.z["src_K"], g.z["src_K"],
                            g.z["k"], g.z["poses"][j], aff, want_gn=True, irls_eps=1e-3, with_affine_cols=with_affine)
            gp = to_np(batch.gn_pair[j]).astype(np.float64)
            gs = to_np(batch.gn_seg[j * g.N:(j + 1) * g.N]).astype(np.float64)
            A = _tri8(gp[:36])
            np_ = 8 if with_affine else 6
            assert_close(A[:np_, :np_], r["A"], GN_TOL, f"A (pair {j}, level {lvl})")
            assert_close(gp[36:36 + np_], r["g_p"], GN_TOL, "g_p")
            assert_close(gs[:, :np_].T, r["B"], GN_TOL, "B")
            assert_close(gs[:, 8], r["D"], GN_TOL, "D")
            assert_close(gs[:, 9], r["g_d"], GN_TOL, "g_d")
            assert_close(gp[44] / (3 * r["P"]), r["cost"], 2e-5, "cost")
            if with_affine:      # the packed 6-column path does not accumulate the (report-only) weighted cost
                assert_close(gp[45], r["wcost"], GN_TOL, "weighted cost")
            if not with_affine:
                assert np.all(A[6:, :] == 0) and np.all(gs[:, 6:8] == 0)


def test_lm_step_matches_closed_form():
    g = Golden("tiny_rects", "cuda")
    lvl = g.n_levels - 1
    batch = _batch_from_golden(g, lvl, False)
    k0 = to_np(batch.k).copy()
    T0 = to_np(batch.poses_matrix()[0]).astype(np.float64)
    batch.gn_step()
    torch.cuda.synchronize()
    geo = cf.compact_geometry(g.z["src_regions"], g.z["src_logdepth"], g.z["src_keypoints"])
    r = cf.evaluate(geo, g.z[f"L{lvl}_src_image"], g.z[f"L{lvl}_trg_images"][0], g.z["src_K"], g.z["src_K"],
                    g.z["k"], g.z["poses"][0], None, want_gn=True, irls_eps=1e-3)
    xi, dk = cf.lm_step(r["A"], r["B"], r["D"], r["g_p"], r["g_d"], 1e-3)
    T1 = cf.se3_exp(xi) @ T0
    assert_close(to_np(batch.k) - k0, dk, 2e-3, "dk")
    assert_close(to_np(batch.poses_matrix()[0]), T1, 1e-5, "pose after one LM step")


def test_lm_loop_decreases_cost_and_recovers_shift():
    """Target = source image displaced by a known pixel shift: the LM loop must reduce the L1 cost
    monotonically over accepted steps and end well below the start."""
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    H, W, N = 96, 128, 8
    probs = []
    for seed in range(3):
        src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind="overlap", seed=seed, noise=0.0)
        src, trg = src.to("cuda"), trg.to("cuda")
        probs.append(make_problem(src, trg.image, trg.K, pose0.cuda(), k0.cuda()))
    batch = AlignmentBatch(probs, irls_eps=1e-3)
    batch.gn_accumulate()
    c0 = to_np(batch.costs()).copy()
    accepted = [c0]
    for _ in range(30):
        batch.gn_step()
        st = to_np(batch.lm_state)
        accepted.append(st[:, 1] / (3 * to_np(batch.pts_per_problem)))
    acc = np.stack(accepted)
    assert np.all(np.diff(acc[1:], axis=0) <= 1e-9), "accepted cost must be non-increasing"
    assert np.all(acc[-1] < 0.5 * c0), f"cost {c0} -> {acc[-1]}"
    assert np.all(np.isfinite(to_np(batch.poses))) and np.all(np.isfinite(to_np(batch.k)))
    st = to_np(batch.lm_state)
    assert np.all(st[:, 3] >= 3), "several steps must have been accepted"


def test_graph_replay_equals_eager():
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    H, W, N = 64, 96, 6

    def build():
        src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind="rects", seed=4, noise=0.01)
        src, trg = src.to("cuda"), trg.to("cuda")
        return AlignmentBatch([make_problem(src, trg.image, trg.K, pose0.cuda(), k0.cuda())])

    a = build()
    a.run_gn(6)
    b = build()
    graph = b.capture_gn(6)          # the warm-up step before capture is rolled back: replay = 6 iterations
    graph.replay()
    torch.cuda.synchronize()
    assert_close(to_np(b.poses), to_np(a.poses), 1e-6, "graph vs eager poses")
    assert_close(to_np(b.k), to_np(a.k), 1e-6, "graph vs eager k")


def test_gn_recovers_pose_of_a_consistent_planar_scene():
    """Geometrically consistent pair (analytic plane rendering): from a perturbed start the device-resident GN/LM
    loop must drive the photometric cost to the interpolation-noise floor and recover the true rotation and the
    true translation up to the joint (translation, depth) scale gauge."""
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    H, W, N = 120, 160, 8
    T_true = syn.small_pose(0.03, -0.02, 0.01, 0.010, -0.015, 0.020)
    src, trg, k_true = syn.planar_scene_pair(H, W, N, T_true, z0=2.0, kind="strips")
    src, trg = src.to("cuda"), trg.to("cuda")
    T0 = torch.eye(4)
    k0 = k_true + 0.05
    batch = AlignmentBatch([make_problem(src, trg.image, trg.K, T0.cuda(), k0.cuda())], irls_eps=1e-3)
    batch.gn_accumulate()
    c0 = float(batch.costs()[0])
    batch.run_gn(40)
    torch.cuda.synchronize()
    c1 = float(batch.lm_state[0, 1]) / (3 * float(batch.pts_per_problem[0]))
    assert c1 < 0.05 * c0, (c0, c1)
    T = to_np(batch.poses_matrix()[0]).astype(np.float64)
    Rerr = T[:3, :3] @ to_np(T_true)[:3, :3].T.astype(np.float64)
    ang = np.arccos(np.clip((np.trace(Rerr) - 1) / 2, -1, 1))
    assert ang < 2e-3, f"rotation error {ang:.2e} rad"
    scale = np.exp(float(batch.k.mean()) - float(k_true.mean()))          # joint scale gauge of (t, depth)
    t_est = T[:3, 3] / scale
    assert np.linalg.norm(t_est - to_np(T_true)[:3, 3]) < 0.15 * np.linalg.norm(to_np(T_true)[:3, 3]), (t_est, scale)


def test_pose_only_lm_holds_the_seeds_and_matches_the_reduced_system():
    """hold_depth: the tracker's parameter set (pose only).  One LM step must equal the solve of the damped pose
    block alone (no Schur elimination), the seeds must not move, and the loop must still reduce the cost."""
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.solver import AlignmentBatch, make_problem
    g = Golden("tiny_rects", "cuda")
    lvl = g.n_levels - 1
    src = g.src(lvl)
    prob = make_problem(src, g.trg(lvl, 0).image, g.trg(lvl, 0).K, g.poses()[0], g.k())
    batch = AlignmentBatch([prob], hold_depth=True)
    k0 = to_np(batch.k).copy()
    T0 = to_np(batch.poses_matrix()[0]).astype(np.float64)
    batch.gn_step()
    torch.cuda.synchronize()
    geo = cf.compact_geometry(g.z["src_regions"], g.z["src_logdepth"], g.z["src_keypoints"])
    r = cf.evaluate(geo, g.z[f"L{lvl}_src_image"], g.z[f"L{lvl}_trg_images"][0], g.z["src_K"], g.z["src_K"],
                    g.z["k"], g.z["poses"][0], None, want_gn=True, irls_eps=1e-3)
    lam = 1e-3
    A = r["A"] + lam * np.diag(np.diag(r["A"]))
    xi = np.linalg.solve(A, -r["g_p"])
    assert np.array_equal(to_np(batch.k), k0), "held seeds moved"
    assert_close(to_np(batch.poses_matrix()[0]), cf.se3_exp(xi) @ T0, 1e-5, "pose after one pose-only LM step")
    # consistent scene, true depth: pose-only tracking must recover the pose
    H, W, N = 120, 160, 8
    T_true = syn.small_pose(0.03, -0.02, 0.01, 0.010, -0.015, 0.020)
    s2, t2, k_true = syn.planar_scene_pair(H, W, N, T_true, z0=2.0, kind="strips")
    s2, t2 = s2.to("cuda"), t2.to("cuda")
    b2 = AlignmentBatch([make_problem(s2, t2.image, t2.K, torch.eye(4).cuda(), k_true.cuda())], hold_depth=True)
    b2.gn_accumulate()
    c0 = float(b2.costs()[0])
    b2.run_gn(40)
    torch.cuda.synchronize()
    c1 = float(b2.lm_state[0, 1]) / (3 * float(b2.pts_per_problem[0]))
    assert c1 < 0.1 * c0, (c0, c1)
    assert torch.equal(b2.k.cpu(), k_true.to(torch.float32))
    T = to_np(b2.poses_matrix()[0]).astype(np.float64)
    Rerr = T[:3, :3] @ to_np(T_true)[:3, :3].T.astype(np.float64)
    ang = np.arccos(np.clip((np.trace(Rerr) - 1) / 2, -1, 1))
    assert ang < 3e-3, f"rotation error {ang:.2e} rad"
    t_true = to_np(T_true)[:3, 3]
    assert np.linalg.norm(T[:3, 3] - t_true) < 0.15 * np.linalg.norm(t_true), T[:3, 3]     # depth known: no gauge
