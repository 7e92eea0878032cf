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
    `affine` = (src (2,), trg (2,)) or None; the target terms are optimised when `opt_affine`."""
    cfg = cost_config or {'mode': 'colour', 'collect_stats': 0}
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
        for _ in range(iters):
            pose = exp_se3(delta) @ T
            res = port.cost_single(src, trg, k, pose, cfg, None if affine is None else (aff_s, aff_t))
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
