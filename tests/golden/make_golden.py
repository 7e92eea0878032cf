"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, CPU):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (makezur/super_primitive @ 37e7d76) ships no tests or golden vectors, so the
parity target is frozen here: for each seeded synthetic case the reference's own functions
(`core.dense_optim`, `core.dense_optim_batch`, `core.depth_render`, `odometery.depth_init`)
are executed unmodified on CPU float32 and their outputs + autograd gradients are stored as
``<case>.npz`` together with the exact inputs.  While generating, ``oracle/ref_port.py`` is
run on the same inputs and asserted to reproduce the reference, pinning the port.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)

from super_primitive_b200 import synthetic as syn          # noqa: E402
from oracle import ref_port as port                          # noqa: E402
from tests.se3 import se3_exp_t                              # noqa: E402

import core.dense_optim as ref_do                            # noqa: E402
import core.dense_optim_batch as ref_dob                     # noqa: E402
import core.depth_render as ref_dr                           # noqa: E402
from image.keyframe import KeyFrame as RefKeyFrame           # noqa: E402

# odometery.depth_init imports only torch/numpy + tool/core -> importable
import odometery.depth_init as ref_di                        # noqa: E402


def ref_kf(kf):
    return RefKeyFrame(kf.image, kf.K, kf.logdepth_perseg, kf.keypoints, kf.keypoint_regions,
                       K_img=kf.K_img)


def t2n(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def same(a, b, what):
    a, b = t2n(a), t2n(b)
    if a.dtype == bool or np.issubdtype(a.dtype, np.integer):
        assert np.array_equal(a, b), what
    else:
        assert np.allclose(a, b, rtol=0, atol=0, equal_nan=True), f"port != reference: {what} " \
            f"max|d|={np.abs(a.astype(np.float64) - b.astype(np.float64)).max()}"


def grads_of(fn, leaves):
    for l in leaves:
        if l is not None and l.grad is not None:
            l.grad = None
    out = fn()
    out['residual'].mean().backward()
    return out, [None if l is None else l.grad.clone() for l in leaves]


def inputs_dict(src, prefix="src_"):
    return {prefix + "image": t2n(src.image), prefix + "K": t2n(src.K), prefix + "K_img": t2n(src.K_img),
            prefix + "logdepth": t2n(src.logdepth_perseg), prefix + "keypoints": t2n(src.keypoints),
            prefix + "regions": t2n(src.keypoint_regions)}


def case_full(name, H, W, N, kind, noise, levels, B, with_affine, seed, stats=True):
    torch.manual_seed(seed)
    cfg = {'mode': 'colour', 'collect_stats': 2 if stats else 0}
    src0 = syn.make_keyframe(H, W, N, kind=kind, seed=seed, noise=noise)
    trgs0 = [syn.make_keyframe(H, W, N, shift=(1.5 + 0.5 * j, 0.75 - 0.25 * j), noise=noise,
                               seed=seed + 10 + j, supporting=True) for j in range(B)]
    src_pyr = syn.keyframe_pyramid(src0, *levels)
    trg_pyrs = [syn.keyframe_pyramid(t, *levels) for t in trgs0]
    g = torch.Generator().manual_seed(seed)
    k0 = (torch.log(torch.tensor(2.0)) + 0.2 * (torch.rand(N, generator=g) - 0.5)).float()
    poses0 = torch.stack([syn.small_pose(0.02 + 0.01 * j, -0.01 * j, 0.005 * j,
                                         0.03 - 0.01 * j, 0.01 * j, -0.02 + 0.005 * j) for j in range(B)])
    aff_s0 = torch.tensor([0.05, 0.01]) if with_affine else None
    aff_t0 = torch.tensor([[-0.02 + 0.01 * j, 0.03 - 0.005 * j] for j in range(B)]) if with_affine else None

    store = dict(H=H, W=W, N=N, B=B, levels=np.array(levels), with_affine=with_affine,
                 k=t2n(k0), poses=t2n(poses0))
    store.update(inputs_dict(src0))
    del store["src_image"]          # == finest pyramid level, stored below as L<last>_src_image
    if with_affine:
        store["aff_src"], store["aff_trg"] = t2n(aff_s0), t2n(aff_t0)

    for li, src in enumerate(src_pyr):
        trg_l = [tp[li] for tp in trg_pyrs]
        tag = f"L{li}_"
        store[tag + "src_image"] = t2n(src.image)
        store[tag + "trg_images"] = np.stack([t2n(t.image) for t in trg_l])
        store[tag + "src_K_img"] = t2n(src.K_img)
        # ---------------- single-target cost: reference vs port ----------------
        k = k0.clone().requires_grad_(True)
        pose = poses0[0].clone().requires_grad_(True)
        a_s = aff_s0.clone().requires_grad_(True) if with_affine else None
        a_t = aff_t0[0].clone().requires_grad_(True) if with_affine else None
        aff = (a_s, a_t) if with_affine else None
        rsrc, rtrg = ref_kf(src), ref_kf(trg_l[0])
        out_r, g_r = grads_of(lambda: ref_do.photomeric_cost(rsrc, rtrg, k, pose, cfg, aff),
                              [k, pose, a_s, a_t])
        out_p, g_p = grads_of(lambda: port.cost_single(src, trg_l[0], k, pose, cfg, aff),
                              [k, pose, a_s, a_t])
        for key in out_r:
            if out_r[key] is not None:
                same(out_r[key], out_p[key], f"{name}/{tag}single/{key}")
        for a, b, nm in zip(g_r, g_p, ["k", "pose", "a_s", "a_t"]):
            if a is not None:
                same(a, b, f"{name}/{tag}single/grad_{nm}")
        store[tag + "single_residual"] = t2n(out_r['residual'])
        store[tag + "single_g_k"] = t2n(g_r[0])
        store[tag + "single_g_pose"] = t2n(g_r[1])
        if with_affine:
            store[tag + "single_g_aff_src"] = t2n(g_r[2])
            store[tag + "single_g_aff_trg"] = t2n(g_r[3])
        if stats:
            for key in ['segm_ids', 'src_pixels', 'src_in_trg_pixels', 'src_valid_mask',
                        'trg_valid_mask', 'full_mask', 'src_pts', 'src_in_trg_pts', 'residual_raw',
                        'src_in_trg_keypoints', 'src_in_trg_keypoints_z',
                        'src_in_trg_keypoints_valid_mask']:
                store[tag + "single_" + key] = t2n(out_r[key])

        # ---------------- precomputed (tracking) path ----------------
        with torch.no_grad():
            pre_r = ref_do.unproject_kf(rsrc, k0)
            pre_p = port.lift_keyframe(src, k0)
        for key in ['src_pixels', 'src_valid_mask', 'src_pts', 'segm_ids']:
            same(pre_r[key], pre_p[key], f"{name}/{tag}unproject_kf/{key}")
        pose = poses0[0].clone().requires_grad_(True)
        a_s = aff_s0.clone().requires_grad_(True) if with_affine else None
        a_t = aff_t0[0].clone().requires_grad_(True) if with_affine else None
        aff = (a_s, a_t) if with_affine else None
        out_r, g_r = grads_of(lambda: ref_do.photomeric_cost_precomputed(pre_r, rtrg, pose, cfg, aff),
                              [pose, a_s, a_t])
        out_p, g_p = grads_of(lambda: port.cost_precomputed(pre_p, trg_l[0], pose, cfg, aff),
                              [pose, a_s, a_t])
        same(out_r['residual'], out_p['residual'], f"{name}/{tag}pre/residual")
        for a, b in zip(g_r, g_p):
            if a is not None:
                same(a, b, f"{name}/{tag}pre/grad")
        store[tag + "pre_residual"] = t2n(out_r['residual'])
        store[tag + "pre_g_pose"] = t2n(g_r[0])
        if with_affine:
            store[tag + "pre_g_aff_src"] = t2n(g_r[1])
            store[tag + "pre_g_aff_trg"] = t2n(g_r[2])

        # ---------------- batch (mapping) path ----------------
        k = k0.clone().requires_grad_(True)
        poses = poses0.clone().requires_grad_(True)
        a_s = aff_s0.clone().requires_grad_(True) if with_affine else None
        a_t = aff_t0.clone().requires_grad_(True) if with_affine else None
        aff = (a_s, a_t) if with_affine else None
        imgs = torch.stack([t.image for t in trg_l])
        Ks = torch.stack([t.K for t in trg_l])
        out_r, g_r = grads_of(lambda: ref_dob.photomeric_cost_batch(rsrc, imgs, Ks, k, poses, cfg, aff),
                              [k, poses, a_s, a_t])
        out_p, g_p = grads_of(lambda: port.cost_batch(src, imgs, Ks, k, poses, cfg, aff),
                              [k, poses, a_s, a_t])
        for key in out_r:
            if out_r[key] is not None:
                same(out_r[key], out_p[key], f"{name}/{tag}batch/{key}")
        for a, b in zip(g_r, g_p):
            if a is not None:
                same(a, b, f"{name}/{tag}batch/grad")
        store[tag + "batch_residual"] = t2n(out_r['residual'])
        store[tag + "batch_g_k"] = t2n(g_r[0])
        store[tag + "batch_g_poses"] = t2n(g_r[1])
        if with_affine:
            store[tag + "batch_g_aff_src"] = t2n(g_r[2])
            store[tag + "batch_g_aff_trg"] = t2n(g_r[3])
        if stats:
            for key in ['trg_valid_mask', 'full_mask', 'src_in_trg_pts', 'residual_raw',
                        'src_in_trg_pixels', 'src_in_trg_keypoints', 'src_in_trg_keypoints_z',
                        'src_in_trg_keypoints_valid_mask']:
                store[tag + "batch_" + key] = t2n(out_r[key])

    # ---------------- geometry-only entry points (finest level) ----------------
    src = src_pyr[-1]
    rsrc = ref_kf(src)
    with torch.no_grad():
        dd_r = ref_do.unproject_kf_to_depths(rsrc, k0)
        same(dd_r, port.dense_depths(src, k0), f"{name}/dense_depths")
        if stats:
            store["dense_depths"] = t2n(dd_r)
        else:                        # large: keep a float64 checksum per segment instead
            store["dense_depths_segsum"] = t2n(dd_r.double().sum((1, 2)))
        pre_r = ref_do.unproject_kf(rsrc, k0)
        store["pre_src_pts"] = t2n(pre_r['src_pts'])
        store["pre_src_pixels"] = t2n(pre_r['src_pixels'])
        store["pre_src_valid_mask"] = t2n(pre_r['src_valid_mask'])
        store["pre_segm_ids"] = t2n(pre_r['segm_ids'])
        for tag, pose, mean in [("render_id", None, False), ("render_pose", poses0[0], False),
                                ("render_mean", poses0[0], True)]:
            im_r = ref_dr.estimate_depth_kf_native(rsrc, k0, pose, mean=mean)
            torch.set_grad_enabled(True)
            same(im_r, port.render_keyframe_depth(src, k0, pose, mean=mean), f"{name}/{tag}")
            store[tag] = t2n(im_r)
        # per-segment re-initialisation from a rendered depth (next row, depth_init.py)
        est = ref_dr.estimate_depth_kf_native(rsrc, k0, poses0[0])
        for mode in ("median", "mean"):
            kk_r, vis_r = ref_di.segment_based_depth_reinit(est.clone(), rsrc, mode, return_info=True)
            kk_p, vis_p = port.segment_median_reinit(est.clone(), src, mode)
            same(kk_r, kk_p, f"{name}/reinit_{mode}")
            same(vis_r, vis_p, f"{name}/reinit_vis")
            store[f"reinit_{mode}"] = t2n(kk_r)
            store["reinit_visible"] = t2n(vis_r)
        store["reinit_est_depth"] = t2n(est)
        torch.set_grad_enabled(True)

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k_: v for k_, v in store.items() if v is not None})
    print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB")


def case_adam(name, H, W, N, kind, steps, seed):
    """K-step Adam trajectory (reference LRs: log-depth 1e-3, pose 1e-2,
    odometery/two_frame_sfm.py:117-121) over (k, xi) with T = Exp(xi) T0."""
    cfg = {'mode': 'colour', 'collect_stats': 0}
    src = syn.make_keyframe(H, W, N, kind=kind, seed=seed, noise=0.01)
    trg = syn.make_keyframe(H, W, N, shift=(2.0, 1.0), noise=0.01, seed=seed + 1, supporting=True)
    k0 = torch.full((N,), float(np.log(2.0)))
    T0 = syn.small_pose(0.02, 0.0, 0.0, 0.01, 0.0, 0.0)

    def run(cost_fn, s, t):
        k = torch.nn.Parameter(k0.clone())
        xi = torch.nn.Parameter(torch.zeros(6))
        opt = torch.optim.Adam([{'params': [k], 'lr': 1e-3}, {'params': [xi], 'lr': 1e-2}], lr=1e-3)
        traj_k, traj_xi, losses = [], [], []
        for _ in range(steps):
            pose = se3_exp_t(xi) @ T0
            loss = cost_fn(s, t, k, pose, cfg)['residual'].mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            traj_k.append(k.detach().clone())
            traj_xi.append(xi.detach().clone())
            losses.append(loss.detach().clone())
        return torch.stack(traj_k), torch.stack(traj_xi), torch.stack(losses)

    tk_r, tx_r, l_r = run(ref_do.photomeric_cost, ref_kf(src), ref_kf(trg))
    tk_p, tx_p, l_p = run(port.cost_single, src, trg)
    same(tk_r, tk_p, f"{name}/traj_k")
    same(tx_r, tx_p, f"{name}/traj_xi")
    store = dict(H=H, W=W, N=N, steps=steps, k0=t2n(k0), T0=t2n(T0), traj_k=t2n(tk_r),
                 traj_xi=t2n(tx_r), losses=t2n(l_r), trg_image=t2n(trg.image))
    store.update(inputs_dict(src))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **store)
    print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB; loss {float(l_r[0]):.6f} -> {float(l_r[-1]):.6f}")


def main():
    torch.set_num_threads(8)
    torch.set_grad_enabled(True)
    # ragged overlapping rectangles, affine, 2-level pyramid, full per-point stats
    case_full("tiny_rects", 40, 56, 5, "rects", 0.02, (0, 2), 3, True, seed=3)
    # exact partition, no affine, single level
    case_full("tiny_strips", 32, 48, 4, "strips", 0.0, (0, 1), 2, False, seed=5)
    # BASELINE config 1 shape: 256x192, 8 segments, 1 level (scalars + gradients only)
    case_full("c1_overlap", 192, 256, 8, "overlap", 0.01, (0, 1), 2, True, seed=7, stats=False)
    # 3-level pyramid at a reduced C2-like shape
    case_full("pyr3_rects", 96, 128, 12, "rects", 0.01, (0, 3), 2, True, seed=11, stats=False)
    case_adam("adam_c1", 96, 128, 8, "overlap", 40, seed=13)
    case_pyramid("pyramid_odd", 45, 70, 4, seed=17)
    case_completion("completion", 60, 80, 12, seed=19)




def case_pyramid(name, H, W, N, seed):
    """The reference's own keyframe_pyramid (image/keyframe.py:77-148, geo_down=False) on a synthetic keyframe,
    incl. odd sizes: pins the blur/decimate and K_img semantics used by synthetic.keyframe_pyramid and by the
    CUDA pyramid kernel."""
    from image.keyframe import keyframe_pyramid as ref_pyr
    kf = syn.make_keyframe(H, W, N, kind="rects", seed=seed, noise=0.02)
    store = dict(H=H, W=W, image=t2n(kf.image), K=t2n(kf.K))
    for (a, b) in [(0, 3), (1, 4), (0, 1)]:
        levels = ref_pyr(ref_kf(kf), a, b)
        torch.set_grad_enabled(True)
        mine = syn.keyframe_pyramid(kf, a, b)
        assert len(levels) == len(mine)
        for i, (r, m) in enumerate(zip(levels, mine)):
            same(r.image, m.image, f"{name}/pyr{a}{b}/L{i}/image")
            same(r.K_img, m.K_img, f"{name}/pyr{a}{b}/L{i}/K_img")
            same(r.K, m.K, f"{name}/pyr{a}{b}/L{i}/K")
            assert r.logdepth_perseg is kf.logdepth_perseg or torch.equal(r.logdepth_perseg, kf.logdepth_perseg)
            store[f"p{a}{b}_L{i}_image"] = t2n(r.image)
            store[f"p{a}{b}_L{i}_K_img"] = t2n(r.K_img)
        store[f"p{a}{b}_n"] = len(levels)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **store)
    print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB")


def _reference_function(path, name, namespace):
    """Compile ONE function of a reference module that cannot be imported here (its module imports SAM etc.) and
    return it: the reference's own code runs, nothing is copied."""
    import ast
    tree = ast.parse(open(os.path.join(REF, path)).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    ns = dict(namespace)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REF, path), "exec"), ns)
    return ns[name]


def case_completion(name, H, W, N, seed):
    """VOID depth-completion tail: the reference's render_depth_avg (depth_completion/segment_based_completion.py:21-27)
    applied to unproject_kf_to_depths output with unseeded segments dropped (lines 48-54)."""
    ref_avg = _reference_function("depth_completion/segment_based_completion.py", "render_depth_avg", {"torch": torch})
    kf = syn.make_keyframe(H, W, N, kind="rects", seed=seed)
    g = torch.Generator().manual_seed(seed)
    k = float(np.log(2.0)) + 0.2 * torch.randn(N, generator=g)
    vis = torch.rand(N, generator=g) > 0.2
    with torch.no_grad():
        d = ref_do.unproject_kf_to_depths(ref_kf(kf), k)
        d[kf.keypoint_regions == 0] = -1
        d = d[vis]
        avg_r, inv_r = ref_avg(d.clone())
        avg_p, inv_p = port.completion_render(kf, k, vis)
    same(avg_r, avg_p, f"{name}/avg")
    same(inv_r, inv_p, f"{name}/invalid")
    store = dict(H=H, W=W, N=N, k=t2n(k), visible=t2n(vis), avg=t2n(avg_r), invalid=t2n(inv_r))
    store.update(inputs_dict(kf))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **store)
    print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB")


def case_renorm(name, seed):
    """renormalise_se3 (lie/lie_algebra.py:41-48), the pose clean-up of the mapping loop (odometery/odometery.py:867,
    880), run from the reference's own module on slightly denormalised poses -- incl. rotations near pi so that all
    four quaternion candidates are exercised.  lie/lie_algebra.py imports lietorch (absent, unpinned) at module
    level without using it in these functions: an empty stub module stands in for the import."""
    sys.modules.setdefault("lietorch", types.ModuleType("lietorch"))
    from lie import lie_algebra as ref_la
    from oracle import window_loop as wl
    from oracle.adam_loop import exp_se3
    g = torch.Generator().manual_seed(seed)
    T_in, T_out = [], []
    for i in range(48):
        xi = torch.randn(6, generator=g, dtype=torch.float64) * (0.5 if i % 2 else 3.0)
        T = exp_se3(xi).float()
        T[:3] += 1e-3 * torch.randn(3, 4, generator=g)
        ref = ref_la.renormalise_se3(T.clone())
        same(ref, wl.renormalise(T), f"{name}/{i}")
        T_in.append(t2n(T))
        T_out.append(t2n(ref))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, T_in=np.stack(T_in), T_out=np.stack(T_out))
    print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    if "--pyramid-only" in sys.argv:
        case_pyramid("pyramid_odd", 45, 70, 4, seed=17)
    elif "--completion-only" in sys.argv:
        case_completion("completion", 60, 80, 12, seed=19)
    elif "--renorm-only" in sys.argv:
        case_renorm("renorm", seed=23)
    else:
        main()
        case_renorm("renorm", seed=23)
