"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by the product path).

CPU restatement of the reference's first-order optimisation loop, the oracle of the device-resident Adam
iteration (`spb_adam_iterate`, super_primitive_b200/csrc/spb_adam.cuh):

* optimiser: `torch.optim.Adam` exactly as the reference builds it -- parameter groups with their own learning
  rates, default betas / eps (odometery/two_frame_sfm.py:117-121: seeds 1e-3, poses 1e-2;
  odometery/odometery.py:303-310: tracker increment + affine 5e-3);
* pose bookkeeping of the tracker (odometery/odometery.py:386-403): the optimised variable is a twist `delta`,
  the cost is evaluated at `pose_to_mat(delta) @ T`, after `optim.step()` the increment is folded into the pose and
  `delta` is re-zeroed (`zero_out_lietorch_tensor`, lie/lietorch_utils.py:22-25) while the Adam state persists;
* cost: oracle/ref_port.py (pinned bit-exactly to the live reference by tests/test_oracle_golden.py).

PARITY UNPINNED for the retraction: upstream uses lietorch (`LieGroupParameter.retr().matrix()`), a C++/CUDA
dependency installed from an unpinned git HEAD (install.sh:9-14) and absent here.  Its documented semantics --
`retr(a) = Exp(a) * X` with the twist ordered (translation, rotation) -- are restated with
`torch.linalg.matrix_exp` of the 4x4 twist matrix, whose autograd derivative at delta = 0 is exactly the
left-perturbation Jacobian.
"""
import torch

from oracle import ref_port as port


def twist_matrix(delta):
    """(6,) twist (tau, phi) -> 4x4 element of se(3)."""
    tau, phi = delta[:3], delta[3:]
    z = torch.zeros((), dtype=delta.dtype)
    rows = [torch.stack([z, -phi[2], phi[1], tau[0]]),
            torch.stack([phi[2], z, -phi[0], tau[1]]),
            torch.stack([-phi[1], phi[0], z, tau[2]]),
            torch.stack([z, z, z, z])]
    return torch.stack(rows)


def exp_se3(delta):
    return torch.linalg.matrix_exp(twist_matrix(delta))


def tracker_adam(src, trg, k0, pose0, iters, lr_pose=1e-2, lr_k=1e-3, lr_aff=5e-3, affine=None, opt_affine=False,
                 cost_config=None):
    """Runs `iters` iterations; returns dict(pose, k, aff_trg, costs[iters]) in the dtype of the inputs.
    `affine` = (src (2,), trg (2,)) or None; the target terms are optimised when `opt_affine`.
    `src` / `trg` may be lists of pyramid levels (coarse -> fine) with `iters` a list of the same length: the
    reference's coarse-to-fine schedule (`for pyr_level ... for iter in range(steps[pyr_level])`,
    odometery/odometery.py:376-384) with ONE optimiser across the levels."""
    cfg = cost_config or {'mode': 'colour', 'collect_stats': 0}
    if not isinstance(src, (list, tuple)):
        src, trg, iters = [src], [trg], [iters]
    dt = k0.dtype
    was = torch.is_grad_enabled()
    torch.set_grad_enabled(True)
    try:
        delta = torch.zeros(6, dtype=dt, requires_grad=True)
        k = k0.detach().clone().requires_grad_(True)
        T = pose0.detach().clone()
        groups = [{'params': [k], 'lr': lr_k}, {'params': [delta], 'lr': lr_pose}]
        aff_s = aff_t = None
        if affine is not None:
            aff_s = affine[0].detach().clone()
            aff_t = affine[1].detach().clone().requires_grad_(opt_affine)
            if opt_affine:
                groups.append({'params': [aff_t], 'lr': lr_aff})
        opt = torch.optim.Adam(groups, lr=1e-3)
        costs = []
        for src_l, trg_l, n in zip(src, trg, iters):
            for _ in range(n):
                pose = exp_se3(delta) @ T
                res = port.cost_single(src_l, trg_l, k, pose, cfg, None if affine is None else (aff_s, aff_t))
                loss = torch.mean(res['residual'])
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                costs.append(float(loss.detach()))
                with torch.no_grad():
                    T = exp_se3(delta.detach()) @ T
                    delta.zero_()
        return {'pose': T.detach(), 'k': k.detach(), 'aff_trg': None if aff_t is None else aff_t.detach(),
                'costs': costs}
    finally:
        torch.set_grad_enabled(was)


def sfm_adam(src_levels, trg_levels, k0, T0s, iters_per_level, lr_k=1e-3, lr_pose=1e-2, cost_config=None):
    """The two-frame SfM loop, odometery/two_frame_sfm.py:127-206: one source keyframe against `len(T0s)` supporting
    frames, coarse-to-fine over the pyramid levels (`src_levels[l]`, `trg_levels[j][l]`), `iters_per_level` iterations per
    level (500 in the reference).  Unlike the tracker, the increment is NOT folded: every pose is `Exp(delta_j) @ T0_j` with
    `delta_j` accumulating in ONE Adam (seeds lr 1e-3, increments lr 1e-2, :117-121); the very first iteration evaluates the
    cost but takes no step (`if count > 0`, :201-205); loss = sum_j mean|residual_j|.
    Returns dict(k, deltas [(6,)...], poses [(4,4)...], losses)."""
    cfg = cost_config or {'mode': 'colour', 'collect_stats': 0}
    was = torch.is_grad_enabled()
    torch.set_grad_enabled(True)
    try:
        dt = k0.dtype
        k = k0.detach().clone().requires_grad_(True)
        deltas = [torch.zeros(1, 6, dtype=dt, requires_grad=True) for _ in T0s]
        opt = torch.optim.Adam([{'params': k, 'lr': lr_k}, {'params': deltas, 'lr': lr_pose}], lr=1e-3)
        losses, count = [], 0
        for lvl in range(len(src_levels)):
            for _ in range(iters_per_level):
                per = []
                for j, T0 in enumerate(T0s):
                    pose = exp_se3(deltas[j][0]) @ T0
                    res = port.cost_single(src_levels[lvl], trg_levels[j][lvl], k, pose, cfg)
                    per.append(torch.mean(torch.abs(res['residual'])))
                loss = torch.sum(torch.stack(per))
                if count > 0:
                    loss.backward()
                    opt.step()
                    opt.zero_grad()
                count += 1
                losses.append(float(loss.detach()))
        return {'k': k.detach(), 'deltas': [d.detach()[0] for d in deltas],
                'poses': [(exp_se3(d.detach()[0]) @ T0) for d, T0 in zip(deltas, T0s)], 'losses': losses}
    finally:
        torch.set_grad_enabled(was)
