"""Golden fixture for the mapping window from the reference's OWN caller code.

Run in the build container only (needs /root/reference, CPU):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_callers.py

`Odometery.mapping` (odometery/odometery.py:687-967) is executed unmodified -- connectivity, relative-pose products,
`photomeric_cost_batch`, loss, the single Adam with its parameter groups, pose folding + `renormalise_se3`, early stop,
result write-back -- on a synthetic window (3 keyframes, one supporting frame each, brightness terms on) whose state is
placed into an `Odometery` object built without its constructor (the constructor needs the SAM / normal networks).
What cannot run here is stubbed and nothing else:
  * `lietorch` (C++/CUDA, unpinned git HEAD, absent): a stand-in with the documented semantics -- `SE3.Identity`,
    `.matrix()`, `LieGroupParameter` = zero tangent (1,6) whose `.retr()` is `Exp(tangent) * group`, the exponential
    being `torch.linalg.matrix_exp` of the twist matrix (translation first).  This is the PARITY-UNPINNED part.
  * `frontend.process_frame`, `data` (import SAM / geffnet / trimesh at module level; not used by `mapping`).
  * `torch.cuda.synchronize` (no GPU here; `mapping` calls it for its wall-clock print) and the hard-coded
    `cuda:0` of `SE3.Identity(1).to(device)` (the stand-in ignores the device).
The outputs are frozen in mapping_window.npz; oracle/window_loop.py is asserted to reproduce them, which pins the
oracle of `spb_window_iterate` to the reference's bookkeeping.
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

from oracle.adam_loop import exp_se3                         # noqa: E402
from oracle import window_loop as wl                         # noqa: E402
from super_primitive_b200 import synthetic as syn            # noqa: E402

ITERS = 8
SHAPE = dict(H=40, W=56, N=4, n_kf=3, n_supp=1, kind="rects", seed=3, affine=True, window_full=False)


def install_stubs():
    lt = types.ModuleType("lietorch")

    class SE3:
        def __init__(self, mat):
            self.mat = mat                                   # (1,4,4)
            self.shape = (1,)
            # lietorch stores translation + quaternion; zero_out_lietorch_tensor (lie/lietorch_utils.py:21-25) also zeroes
            # the identity elements of the frames that are not optimised -- a zero quaternion still maps to R = I
            self.data = torch.tensor([[0., 0., 0., 0., 0., 0., 1.]])

        @staticmethod
        def Identity(n):
            return SE3(torch.eye(4)[None].clone())

        def to(self, device):
            return self

        def matrix(self):
            return self.mat

    class LieGroupParameter(torch.Tensor):
        @staticmethod
        def __new__(cls, group, requires_grad=True):
            return torch.Tensor._make_subclass(cls, torch.zeros(1, 6), requires_grad)

        def __init__(self, group, requires_grad=True):
            self.group = group

        def retr(self):
            return SE3((exp_se3(self.as_subclass(torch.Tensor)[0]) @ self.group.mat[0])[None])

    lt.SE3, lt.LieGroupParameter = SE3, LieGroupParameter
    sys.modules["lietorch"] = lt
    fe = types.ModuleType("frontend")
    fe.__path__ = []
    sys.modules["frontend"] = fe
    sys.modules["frontend.process_frame"] = types.ModuleType("frontend.process_frame")
    fe.process_frame = sys.modules["frontend.process_frame"]
    sys.modules["data"] = types.ModuleType("data")
    torch.cuda.synchronize = lambda *a, **k: None


class _Queue:
    """Stands in for the GUI queue; keeps the last message (`mapping` pushes all poses at the end, :918-933)."""
    last = None

    def push(self, msg, *a, **k):
        self.last = msg


def build_odometery(window):
    import odometery.odometery as od
    from image.keyframe import KeyFrame as RefKeyFrame
    frames, n_kf = window['frames'], SHAPE['n_kf']
    o = object.__new__(od.Odometery)                         # no constructor: it loads the SAM / normal networks
    o.config = {'aligment': {'cost_params': {'normal_loss': 'cosine', 'normal_weight': 0.0, 'depth_median_weight': 0.0},
                             'median_loss_weight': 0.0, 'mapping': {'supp_every_n': 3}}}
    o.window_size, o.mono_init, o.opt_supporting, o.affine_compensation = 5, False, True, True
    o.initialised, o.mapping_scheduled, o.paused = True, False, False
    o.pose_to_mat = lambda x: (x.retr() if isinstance(x, sys.modules["lietorch"].LieGroupParameter) else x).matrix()[0]
    o.viz_queue = _Queue()
    o.check_if_paused = lambda: None
    o.tracked_poses_to_supp = lambda: None                   # the running supporting frames are set below
    o.global_kf_trajectory, o.global_kf_scale = {}, {}
    o.tracked_frames, o.tracked_poses, o.tracked_timestamps, o.tracked_affines = [], [], [], []
    kfs = frames[:n_kf]
    o.kfs = [RefKeyFrame(f['kf'].image, f['kf'].K, f['kf'].logdepth_perseg, f['kf'].keypoints, f['kf'].keypoint_regions)
             for f in kfs]
    o.kf_poses = [f['T'].clone() for f in kfs]
    o.kf_logdepths = [f['k'].clone() for f in kfs]
    o.kf_affines = [f['aff'].clone() for f in kfs]
    o.kf_timestamps = [10 * i for i in range(n_kf)]
    # supporting frames: frame n_kf + i belongs to keyframe i (n_supp = 1); those of the LAST keyframe are the "running" ones
    o.supp_kfs_class, o.supp_kfs_opt = [[] for _ in range(n_kf)], [[] for _ in range(n_kf)]
    o.curr_supp_kfs, o.curr_supp_kf_poses, o.curr_supp_kf_timestamps, o.curr_supp_kf_affines = [], [], [], []
    for i in range(n_kf):
        f = frames[n_kf + i]
        kf, ts = RefKeyFrame(f['image'], f['K']), 10 * i + 5
        if i == n_kf - 1:
            o.curr_supp_kfs.append(kf)
            o.curr_supp_kf_poses.append(f['T'].clone())
            o.curr_supp_kf_timestamps.append(ts)
            o.curr_supp_kf_affines.append(f['aff'].clone())
        else:
            o.supp_kfs_class[i].append(od.SupportingKF(kf=kf, timestamp=ts))
            o.supp_kfs_opt[i].append(od.ParamsSupportingKF(pose=f['T'].clone(), timestamp=ts, affine=f['aff'].clone()))
    return o


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    install_stubs()
    window = syn.mapping_window(**SHAPE)
    assert SHAPE['n_supp'] == 1
    o = build_odometery(window)
    o.mapping(num_iters=ITERS, mode='map')
    torch.set_grad_enabled(True)
    n_kf = SHAPE['n_kf']
    T = [p.detach() for p in o.kf_poses]
    k = [x.detach() for x in o.kf_logdepths]
    aff = [a.detach() for a in o.kf_affines]
    # supporting frame of keyframe i = frame n_kf + i.  `mapping` writes the optimised supporting-frame parameters back
    # per `self.supp_kfs_opt[src_id]` (:947-957), which is empty for the LAST keyframe (its supporting frames are the
    # "running" ones), so their results never reach the object: the poses are taken from the final GUI message instead
    # (all frames, :918-933) and the brightness terms of that one frame are not observable (stored as NaN).
    viz_poses = o.viz_queue.last[3]
    assert len(viz_poses) == len(window['frames'])
    for i in range(n_kf):
        T.append(viz_poses[n_kf + i].detach())
        if i == n_kf - 1:
            aff.append(torch.full((2,), float('nan')))
        else:
            assert torch.equal(o.supp_kfs_opt[i][0].pose, viz_poses[n_kf + i])
            aff.append(o.supp_kfs_opt[i][0].affine.detach())
    for i in range(n_kf):
        assert torch.equal(viz_poses[i], o.kf_poses[i])
    # the oracle must reproduce the reference's run (same learning rates: 'map' mode, odometery.py:579-586)
    got = wl.mapping_adam(window['frames'], window['edges'], ITERS, lr_pose=1e-4, lr_k=1e-2, lr_aff=1e-5, stop_tol=1e-8)
    worst = 0.0
    for f in range(len(window['frames'])):
        worst = max(worst, float((got['T'][f] - T[f]).abs().max()))
        if not torch.isnan(aff[f]).any():
            worst = max(worst, float((got['aff'][f] - aff[f]).abs().max()))
        if f < n_kf:
            worst = max(worst, float((got['k'][f] - k[f]).abs().max()))
    moved = max(float((T[f] - window['frames'][f]['T']).abs().max()) for f in range(len(T)))
    print(f"oracle vs reference mapping(): max |d| = {worst:.3e} (parameters moved by up to {moved:.3e})")
    assert worst < 2e-6, worst
    store = dict(iters=ITERS, T=np.stack([t.numpy() for t in T]), k=np.stack([x.numpy() for x in k]),
                 aff=np.stack([a.numpy() for a in aff]), lr=np.array([1e-4, 1e-2, 1e-5]), stop_tol=1e-8,
                 **{"shape_" + key: np.array(val) for key, val in SHAPE.items()})
    path = os.path.join(HERE, "mapping_window.npz")
    np.savez_compressed(path, **store)
    print(f"mapping_window: wrote {os.path.getsize(path) / 1e3:.0f} kB")


TRACK = dict(H=48, W=64, N=5, kind="overlap", seed=11, noise=0.01, iters=12, lr=5e-3)


def case_tracker():
    """`Odometery.track_frame` (odometery/odometery.py:323-447) run unmodified: one frame tracked against the latest
    keyframe at one pyramid level -- `unproject_kf` precompute, `photomeric_cost_precomputed` at
    `Delta @ inv(T_frame) @ T_kf`, Adam on the increment (track.lr) and on the frame's brightness terms (5e-3), folding
    `T_frame <- T_frame @ inv(Delta)`, increment re-zeroed, final `renormalise_se3`.  Only `init_supporting_frame` (the
    frontend call that resizes the frame) is replaced.  Pins oracle/adam_loop.tracker_adam, the oracle of
    `spb_adam_iterate`: the oracle optimises the relative pose T <- Exp(delta) T directly, which is the same iteration
    written without the camera-to-world products, so the comparison is to rounding, not bit for bit."""
    import odometery.odometery as od
    from image.keyframe import KeyFrame as RefKeyFrame
    from oracle import adam_loop
    c = TRACK
    src, trg, k0, pose0 = syn.two_frame_problem(c['H'], c['W'], c['N'], kind=c['kind'], seed=c['seed'], noise=c['noise'])
    aff_src, aff_trg = torch.tensor([0.05, 0.01]), torch.tensor([-0.02, 0.03])
    T_kf = syn.small_pose(0.3, -0.1, 0.2, 0.05, -0.04, 0.03)              # camera-to-world of the keyframe
    T_frame = T_kf @ torch.linalg.inv(pose0)                              # so that inv(T_frame) @ T_kf = pose0
    o = object.__new__(od.Odometery)
    o.config = {'aligment': {'cost_params': {'normal_loss': 'cosine', 'normal_weight': 0.0, 'depth_median_weight': 0.0},
                             'track': {'pyramid_min': 0, 'pyramid_max': 1, 'steps': [c['iters']], 'lr': c['lr']}}}
    o.affine_compensation = True
    o.pose_to_mat = lambda x: (x.retr() if isinstance(x, sys.modules["lietorch"].LieGroupParameter) else x).matrix()[0]
    o.viz_queue = _Queue()
    o.kfs = [RefKeyFrame(src.image, src.K, src.logdepth_perseg, src.keypoints, src.keypoint_regions)]
    o.kf_poses, o.kf_logdepths, o.kf_affines, o.kf_timestamps = [T_kf.clone()], [k0.clone()], [aff_src.clone()], ["000000"]
    o.current_track, o.current_aff = T_frame.clone(), aff_trg.clone()
    o.tracked_frames, o.tracked_poses, o.tracked_timestamps, o.tracked_affines = [], [], [], []
    o.global_track_trajectory = {}
    supp = RefKeyFrame(trg.image, trg.K)
    o.init_supporting_frame = lambda frame: (supp, o.current_track, o.current_aff.detach().clone())
    res = o.track_frame(None, "000001")
    torch.set_grad_enabled(True)
    rel = torch.linalg.inv(res['pose']) @ T_kf                            # the relative pose the tracker ended at
    want = adam_loop.tracker_adam(src, trg, k0, pose0, c['iters'], lr_pose=c['lr'], lr_k=0.0, lr_aff=5e-3,
                                  affine=(aff_src, aff_trg), opt_affine=True)
    e_pose = float((want['pose'] - rel).abs().max())
    e_aff = float((want['aff_trg'] - res['affine']).abs().max())
    moved = float((rel - pose0).abs().max())
    print(f"oracle vs reference track_frame(): pose {e_pose:.3e}, brightness {e_aff:.3e} (pose moved by {moved:.3e})")
    assert e_pose < 5e-6 and e_aff < 1e-6 and torch.equal(want['k'], k0)
    store = dict(rel_pose=rel.numpy(), frame_pose=res['pose'].numpy(), kf_pose=T_kf.numpy(), aff_trg=res['affine'].numpy(),
                 aff_src=aff_src.numpy(), aff_trg0=aff_trg.numpy(), **{"cfg_" + key: np.array(val) for key, val in c.items()})
    path = os.path.join(HERE, "tracker.npz")
    np.savez_compressed(path, **store)
    print(f"tracker: wrote {os.path.getsize(path) / 1e3:.0f} kB")


SFM = dict(H=48, W=64, N=8, kind="overlap", seed=21, noise=0.01, levels=2, n_supp=2)


def case_sfm():
    """`SfM.run` (odometery/two_frame_sfm.py:127-215) run unmodified -- BASELINE config 0 in miniature: one source
    keyframe with 8 segments against two supporting frames, 2 pyramid levels x 500 iterations, the reference's Adam
    (seeds 1e-3, poses 1e-2), poses as `LieGroupParameter`s that are never re-zeroed.  `init_keyframes` /
    `init_optimisation` (dataset + SAM frontend + random start) are replaced by fixed synthetic state; the GUI queues
    are inert.  Pins oracle/adam_loop.sfm_adam bit for bit."""
    import queue
    import odometery.two_frame_sfm as sfm_mod
    from image.keyframe import KeyFrame as RefKeyFrame
    from oracle import adam_loop
    lt = sys.modules["lietorch"]
    c = SFM
    src = syn.make_keyframe(c['H'], c['W'], c['N'], kind=c['kind'], seed=c['seed'], noise=c['noise'])
    trgs = [syn.make_keyframe(c['H'], c['W'], c['N'], shift=(2.0 + j, 1.0 - 0.5 * j), noise=c['noise'],
                              seed=c['seed'] + 1 + j, supporting=True) for j in range(c['n_supp'])]
    T0s = [syn.small_pose(0.02 + 0.01 * j, 0.004, -0.003, 0.003, -0.002 * (j + 1), 0.0015) for j in range(c['n_supp'])]
    k0 = torch.log(2.0 + 2.0 * torch.rand(c['N'], generator=torch.Generator().manual_seed(c['seed'])))   # :103-105
    o = object.__new__(sfm_mod.SfM)
    o.config = {'aligment': {'pyramid_min': 0, 'pyramid_max': c['levels'],
                             'cost_params': {'normal_loss': 'cosine', 'normal_weight': 0.0, 'depth_median_weight': 0.0}}}
    o.paused = False
    o.init_keyframes = lambda: None
    o.init_optimisation = lambda: None
    o.kf_queue, o.viz_queue, o.pause_queue = _Queue(), _Queue(), queue.Queue()
    o.waitev = types.SimpleNamespace(wait=lambda: None)
    o.src_keyframe = RefKeyFrame(src.image, src.K, src.logdepth_perseg, src.keypoints, src.keypoint_regions)
    pose_to_mat = lambda x: x.retr().matrix()[0]                                                         # noqa: E731
    o.supp_frames = [(RefKeyFrame(t.image, t.K), lt.LieGroupParameter(lt.SE3(T0[None].clone())), pose_to_mat)
                     for t, T0 in zip(trgs, T0s)]
    o.src_depth_keypoints_opt = torch.nn.Parameter(k0.clone())
    o.instatiate_optimisation()
    o.run()
    torch.set_grad_enabled(True)
    k_ref = o.src_depth_keypoints_opt.detach()
    d_ref = [p.detach().as_subclass(torch.Tensor)[0] for _, p, _ in o.supp_frames]
    src_lv = syn.keyframe_pyramid(src, 0, c['levels'])
    trg_lv = [syn.keyframe_pyramid(t, 0, c['levels']) for t in trgs]
    got = adam_loop.sfm_adam(src_lv, trg_lv, k0, T0s, 500)
    worst = max([float((got['k'] - k_ref).abs().max())] + [float((a - b).abs().max()) for a, b in zip(got['deltas'], d_ref)])
    print(f"oracle vs reference SfM.run(): max |d| = {worst:.3e} (seeds moved by {float((k_ref - k0).abs().max()):.3e}, "
          f"increments up to {max(float(d.abs().max()) for d in d_ref):.3e})")
    assert worst == 0.0, worst
    store = dict(k0=k0.numpy(), k=k_ref.numpy(), deltas=np.stack([d.numpy() for d in d_ref]),
                 T0s=np.stack([T.numpy() for T in T0s]), loss_first=got['losses'][0], loss_last=got['losses'][-1],
                 **{"cfg_" + key: np.array(val) for key, val in c.items()})
    path = os.path.join(HERE, "sfm_run.npz")
    np.savez_compressed(path, **store)
    print(f"sfm_run: wrote {os.path.getsize(path) / 1e3:.0f} kB; loss {got['losses'][0]:.5f} -> {got['losses'][-1]:.5f}")


if __name__ == "__main__":
    main()
    case_tracker()
    case_sfm()
