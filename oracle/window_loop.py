"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by the product path).

CPU restatement of the reference's windowed mapping loop (odometery/odometery.py:687-915), the oracle of the
device-resident window iteration (`spb_window_iterate`, super_primitive_b200/csrc/spb_window.cu):

* frames of a window carry a camera-to-world pose T_f, a twist increment delta_f held at zero, optional brightness
  terms and, for keyframes, the segments + log-depth seeds;
* per source keyframe s ONE `photomeric_cost_batch` call over its targets at the relative poses
  `Delta_b @ inv(T_b) @ T_s @ inv(Delta_s)` (:793,:817), loss = sum_s mean_b residual (:845-851);
* ONE `torch.optim.Adam` with the reference's parameter groups (:628-638): seeds lr_k, pose increments lr_pose,
  brightness terms lr_aff;
* after `optim.step()`: T_f <- T_f @ inv(Delta_f), `renormalise_se3`, delta re-zeroed with its Adam state kept
  (:861-882); early stop on the relative loss change (:907-915).

Cost: oracle/ref_port.py (pinned bit-exactly to the live reference).  `renormalise` restates
lie/lie_algebra.py:10-48,56-118 (pytorch3d's matrix <-> quaternion conversions) and is pinned to the reference's own
function by tests/golden/renorm.npz (generated with a stub standing in for the absent lietorch import).
PARITY UNPINNED for the twist exponential: upstream uses lietorch (unpinned git HEAD, absent here); restated with
`torch.linalg.matrix_exp` as in oracle/adam_loop.py.
"""
import torch

from oracle import ref_port as port
from oracle.adam_loop import exp_se3


def matrix_to_quaternion(m):
    """lie/lie_algebra.py:56-118 for one 3x3 matrix (real part first)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.reshape(9).unbind()
    q2 = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22])
    q_abs = torch.where(q2 > 0, torch.sqrt(torch.clamp(q2, min=0)), torch.zeros_like(q2))
    cand = torch.stack([
        torch.stack([q_abs[0] ** 2, m21 - m12, m02 - m20, m10 - m01]),
        torch.stack([m21 - m12, q_abs[1] ** 2, m10 + m01, m02 + m20]),
        torch.stack([m02 - m20, m10 + m01, q_abs[2] ** 2, m12 + m21]),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[3] ** 2]),
    ])
    flr = torch.tensor(0.1, dtype=m.dtype)
    cand = cand / (2.0 * torch.maximum(q_abs[:, None], flr))
    return cand[int(torch.argmax(q_abs))]


def quaternion_to_matrix(q):
    """lie/lie_algebra.py:10-38"""
    r, i, j, k = q.unbind()
    two_s = 2.0 / (q * q).sum()
    return torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)]).reshape(3, 3)


def renormalise(T):
    """renormalise_se3, lie/lie_algebra.py:41-48"""
    T = T.clone()
    T[:3, :3] = quaternion_to_matrix(matrix_to_quaternion(T[:3, :3]))
    return T


def group_edges(edges):
    """[(src, trg)] -> [(src, [trg...])] in order of first appearance (the reference iterates its connectivity dict)."""
    order, by = [], {}
    for s, t in edges:
        if s not in by:
            by[s] = []
            order.append(s)
        by[s].append(t)
    return [(s, by[s]) for s in order]


def mapping_adam(frames, edges, iters, lr_pose=1e-4, lr_k=1e-2, lr_aff=1e-5, stop_tol=0.0, cost_config=None):
    """frames: list of dicts {T (4,4), image (3,Hl,Wl), K (3,3), aff (2,)|None, opt_pose, opt_aff,
    kf KeyFrame|None, k (N,)|None, opt_seeds}; edges: [(src, trg)].
    Returns dict(T=[...], k=[...|None], aff=[...|None], losses=[...], steps=int)."""
    cfg = cost_config or {'mode': 'colour', 'collect_stats': 0}
    was = torch.is_grad_enabled()
    torch.set_grad_enabled(True)
    try:
        dt = frames[0]['T'].dtype
        T = [f['T'].detach().clone() for f in frames]
        use_aff = frames[0].get('aff') is not None
        delta = [torch.zeros(6, dtype=dt, requires_grad=bool(f.get('opt_pose'))) for f in frames]
        aff = [None if not use_aff else f['aff'].detach().clone().requires_grad_(bool(f.get('opt_aff'))) for f in frames]
        k = [None if f.get('kf') is None else f['k'].detach().clone().requires_grad_(bool(f.get('opt_seeds')))
             for f in frames]
        groups = [{'params': [x for x in k if x is not None and x.requires_grad], 'lr': lr_k},
                  {'params': [d for d in delta if d.requires_grad], 'lr': lr_pose}]
        if use_aff:
            groups.append({'params': [a for a in aff if a.requires_grad], 'lr': lr_aff})
        groups = [g for g in groups if g['params']]
        opt = torch.optim.Adam(groups, lr=1e-3)
        grouped = group_edges(edges)
        losses, prev, steps = [], float('inf'), 0
        for _ in range(iters):
            per_src = []
            for s, trgs in grouped:
                src_delta = exp_se3(delta[s])
                poses = torch.stack([exp_se3(delta[b]) @ torch.linalg.inv(T[b]) @ T[s] @ torch.linalg.inv(src_delta)
                                     for b in trgs])
                images = torch.stack([frames[b]['image'] for b in trgs])
                Ks = torch.stack([frames[b]['K'] for b in trgs])
                ac = (aff[s], torch.stack([aff[b] for b in trgs])) if use_aff else None
                res = port.cost_batch(frames[s]['kf'], images, Ks, k[s], poses, cfg, ac)
                per_src.append(res['residual'].mean())
            loss = torch.sum(torch.stack(per_src))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            steps += 1
            with torch.no_grad():
                for f in range(len(frames)):
                    T[f] = renormalise(T[f] @ torch.linalg.inv(exp_se3(delta[f].detach())))
                    delta[f].zero_()
            lv = float(loss.detach())
            losses.append(lv)
            if stop_tol > 0 and abs(lv - prev) / prev < stop_tol:
                break
            prev = lv
        return {'T': [t.detach() for t in T], 'k': [None if x is None else x.detach() for x in k],
                'aff': [None if a is None else a.detach() for a in aff], 'losses': losses, 'steps': steps}
    finally:
        torch.set_grad_enabled(was)
