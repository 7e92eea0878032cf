"""CPU: the mapping-window update (spb_window_update) without a GPU.

The per-item arithmetic of the kernel (super_primitive_b200/csrc/spb_window_math.h) is compiled with g++ into a host
harness (tests/host/window_host.cpp, test infrastructure) and driven with per-edge gradients taken from the pinned
cost port by autograd; the resulting trajectory of poses / seeds / brightness terms must follow
oracle/window_loop.py -- the reference's mapping loop (odometery/odometery.py:687-915) with ONE torch.optim.Adam over
the full autograd graph.  This checks the chain rule through `Delta_b inv(T_b) T_s inv(Delta_s)`, the edge -> frame
gradient sums and loss weights, the Adam arithmetic, the pose folding + quaternion renormalisation and the early
stop.  The GPU test (tests/test_gpu_window.py) then checks the kernel's thread mapping with the real gradient kernel.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from super_primitive_b200 import _native as nat
from super_primitive_b200 import synthetic as syn
from super_primitive_b200.window import window_layout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host", "window_host.cpp")
    out_dir = os.path.join(HERE, "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libwindow_host.so")
    deps = [src, os.path.join(ROOT, "super_primitive_b200", "csrc", "spb_window_math.h"),
            os.path.join(ROOT, "include", "spb200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.window_update_host.argtypes = [C.POINTER(nat.SpbWindow), C.c_void_p, C.c_void_p] + [C.c_double] * 7
    lib.window_poses_host.argtypes = [C.POINTER(nat.SpbWindow)]
    lib.renormalise_host.argtypes = [C.c_void_p]
    lib.se3_exp_host.argtypes = [C.c_void_p, C.c_void_p]
    return lib


class HostWindows:
    """The state arrays of SpbWindow in host memory (numpy), laid out by the product's own window_layout()."""

    def __init__(self, windows):
        self.lay = lay = window_layout(windows)
        self.frames = [f for w in windows for f in w['frames']]
        F, E = lay['n_frames'], lay['n_edges']
        self.use_aff = self.frames[0].get('aff') is not None
        f32 = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))     # noqa: E731
        self.frame_T = np.stack([f32(f['T']).reshape(16) for f in self.frames])
        self.frame_aff = np.stack([f32(f['aff']) for f in self.frames]) if self.use_aff else None
        self.k = np.concatenate([f32(f['k']) for f in self.frames if f.get('kf') is not None])
        self.edge_pose = np.zeros((E, 16), np.float32)
        self.adam_frame = np.zeros((F, nat.WIN_ADAM_FRAME), np.float32)
        self.adam_seg = np.zeros((lay['seg_total'], nat.ADAM_SEG), np.float32)
        self.win_state = np.zeros((lay['n_windows'], nat.WIN_NSTATE), np.float32)
        self.edge_tw = np.zeros((E, 12), np.float32)
        self.out_pair = np.zeros((E, nat.PAIR_NOUT), np.float32)
        self.out_gk = np.zeros(lay['gk_total'], np.float32)
        p = lambda a: a.ctypes.data                                                     # noqa: E731
        self.c = nat.SpbWindow(lay['n_windows'], F, E, lay['seg_total'],
                               *(p(lay[n]) for n in ('win_frame_off', 'win_edge_off', 'edge_src', 'edge_trg', 'edge_w',
                                                     'edge_seg_off', 'frame_seg_off', 'frame_seg_cnt', 'frame_flags')),
                               p(self.frame_T), None if self.frame_aff is None else p(self.frame_aff), p(self.k),
                               p(self.edge_pose), p(self.adam_frame), p(self.adam_seg), p(self.win_state), p(self.edge_tw))

    def edge_gradients(self):
        """Fill out_pair / out_gk from the pinned cost port: autograd with respect to the edge's relative pose, the
        source seeds and the brightness terms (what spb_grad_accumulate produces on the device)."""
        from oracle import ref_port as port
        lay = self.lay
        cfg = {'mode': 'colour', 'collect_stats': 0}
        with torch.enable_grad():
            for e in range(lay['n_edges']):
                s, t = int(lay['edge_src'][e]), int(lay['edge_trg'][e])
                so, n = int(lay['frame_seg_off'][s]), int(lay['frame_seg_cnt'][s])
                pose = torch.from_numpy(self.edge_pose[e].reshape(1, 4, 4).copy()).requires_grad_(True)
                k = torch.from_numpy(self.k[so:so + n].copy()).requires_grad_(True)
                ac, a_t = None, None
                if self.use_aff:
                    a_t = torch.from_numpy(self.frame_aff[t][None].copy()).requires_grad_(True)
                    ac = (torch.from_numpy(self.frame_aff[s].copy()), a_t)
                fs, ft = self.frames[s], self.frames[t]
                res = port.cost_batch(fs['kf'], ft['image'][None], ft['K'][None], k, pose, cfg, ac)
                cost = res['residual'][0]
                cost.backward()
                row = self.out_pair[e]
                row[:] = 0
                row[0] = float(cost.detach())
                g = pose.grad[0].numpy()
                row[1:4] = g[:3, 3]
                row[4:13] = g[:3, :3].reshape(9)
                if a_t is not None:
                    row[13:15] = a_t.grad[0].numpy()
                o = int(lay['edge_seg_off'][e])
                self.out_gk[o:o + n] = k.grad.numpy()


def _f64_frames(window):
    from tests.test_gpu_adam import _f64
    c = lambda t: None if t is None else t.double()      # noqa: E731
    out = []
    for f in window['frames']:
        g = dict(f)
        g.update(T=c(f['T']), image=c(f['image']), K=c(f['K']), aff=c(f['aff']), k=c(f['k']),
                 kf=None if f['kf'] is None else _f64(f['kf']))
        out.append(g)
    return dict(frames=out, edges=window['edges'])


def test_layout_follows_the_reference_connectivity():
    w = syn.mapping_window(24, 32, 3, n_kf=3, n_supp=1)
    # keyframes 0..2, supporting frames 3..5 (one per keyframe); the supporting frame of keyframe s-1 also serves s
    assert w['edges'] == [(0, 1), (0, 3), (1, 0), (1, 2), (1, 4), (1, 3), (2, 1), (2, 5), (2, 4)]
    lay = window_layout([w, w])
    assert lay['n_frames'] == 12 and lay['n_edges'] == 18 and lay['seg_total'] == 18
    assert lay['win_frame_off'].tolist() == [0, 6, 12] and lay['win_edge_off'].tolist() == [0, 9, 18]
    assert lay['edge_src'][9:].tolist() == [s + 6 for s, _ in w['edges']]
    np.testing.assert_allclose(lay['edge_w'][:9], [1 / 2] * 2 + [1 / 4] * 4 + [1 / 3] * 3, rtol=1e-7)
    assert lay['frame_flags'][:6].tolist() == [0, 7, 7, 3, 3, 3]
    assert lay['edge_seg_off'].tolist() == list(range(0, 54, 3))
    with pytest.raises(ValueError):
        window_layout([dict(frames=w['frames'], edges=[(3, 0)])])       # a supporting frame cannot be a source


def test_host_renormalise_and_exp_match_the_oracle(host):
    from oracle import window_loop as wl
    from oracle.adam_loop import exp_se3
    rng = np.random.default_rng(0)
    for i in range(40):
        xi = rng.normal(size=6) * (0.5 if i % 2 else 3.0)              # includes rotations near pi (all branches)
        E = np.zeros(12)
        host.se3_exp_host(xi.ctypes.data, E.ctypes.data)
        want = exp_se3(torch.from_numpy(xi)).numpy()
        np.testing.assert_allclose(E.reshape(3, 4), want[:3], atol=1e-12)
        T = np.eye(4, dtype=np.float32)
        T[:3] = (E.reshape(3, 4) + rng.normal(size=(3, 4)) * 1e-3).astype(np.float32)   # slightly denormalised
        got = T.copy()
        host.renormalise_host(got.ctypes.data)
        ref = wl.renormalise(torch.from_numpy(T)).numpy()
        np.testing.assert_allclose(got, ref, atol=2e-6)
        R = got[:3, :3].astype(np.float64)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-6)


def test_renormalise_port_is_pinned_to_the_reference():
    """tests/golden/renorm.npz holds the reference's own renormalise_se3 outputs (make_golden.py)."""
    from oracle import window_loop as wl
    z = np.load(os.path.join(HERE, "golden", "renorm.npz"))
    for T, want in zip(z["T_in"], z["T_out"]):
        got = wl.renormalise(torch.from_numpy(T)).numpy()
        assert np.array_equal(got, want)


@pytest.mark.parametrize("affine,n_kf,n_supp", [(True, 3, 1), (False, 2, 2)])
def test_host_window_update_follows_the_reference_mapping_loop(host, affine, n_kf, n_supp):
    from oracle import window_loop as wl
    iters = 6
    lrs = dict(lr_pose=1e-3, lr_k=1e-2, lr_aff=1e-3)
    w = syn.mapping_window(40, 56, 4, n_kf=n_kf, n_supp=n_supp, kind="rects", seed=3, affine=affine)
    want64 = wl.mapping_adam(_f64_frames(w)['frames'], w['edges'], iters, **lrs)
    want32 = wl.mapping_adam(w['frames'], w['edges'], iters, **lrs)
    hw = HostWindows([w])
    host.window_poses_host(C.byref(hw.c))
    losses = []
    for _ in range(iters):
        hw.edge_gradients()
        host.window_update_host(C.byref(hw.c), hw.out_pair.ctypes.data, hw.out_gk.ctypes.data, lrs['lr_pose'],
                                lrs['lr_k'], lrs['lr_aff'], 0.9, 0.999, 1e-8, 0.0)
        losses.append(float(hw.win_state[0, 1]))
    err = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))   # noqa: E731
    assert hw.win_state[0, 0] == iters
    np.testing.assert_allclose(losses, want64['losses'], rtol=1e-4)
    lay = hw.lay
    for f in range(lay['n_frames']):
        e_host, e_ref = err(hw.frame_T[f].reshape(4, 4), want64['T'][f]), err(want32['T'][f], want64['T'][f])
        assert e_host <= max(1e-5, 2 * e_ref), f"pose of frame {f}: {e_host:.2e} (float32 oracle {e_ref:.2e})"
        if want64['k'][f] is not None:
            o, n = lay['frame_seg_off'][f], lay['frame_seg_cnt'][f]
            e_host, e_ref = err(hw.k[o:o + n], want64['k'][f]), err(want32['k'][f], want64['k'][f])
            assert e_host <= max(1e-5, 2 * e_ref), f"seeds of frame {f}: {e_host:.2e} (float32 oracle {e_ref:.2e})"
        if affine:
            e_host = err(hw.frame_aff[f], want64['aff'][f])
            e_ref = err(want32['aff'][f], want64['aff'][f])
            assert e_host <= max(1e-5, 2 * e_ref), f"brightness of frame {f}: {e_host:.2e}"
    # held parameters did not move; optimised ones moved by about lr per step (Adam)
    assert np.array_equal(hw.k[:4], w['frames'][0]['k'].numpy())
    assert abs(hw.k[4] - float(w['frames'][1]['k'][0])) > 1e-3
    # relative poses of the next iteration are consistent with the frame poses
    for e in range(lay['n_edges']):
        Ts = hw.frame_T[lay['edge_src'][e]].reshape(4, 4).astype(np.float64)
        Tt = hw.frame_T[lay['edge_trg'][e]].reshape(4, 4).astype(np.float64)
        np.testing.assert_allclose(hw.edge_pose[e].reshape(4, 4), np.linalg.inv(Tt) @ Ts, atol=1e-6)


def test_host_early_stop_freezes_the_window(host):
    from oracle import window_loop as wl
    w = syn.mapping_window(32, 40, 3, n_kf=2, n_supp=0, kind="strips", seed=1, affine=False)
    lrs = dict(lr_pose=1e-4, lr_k=1e-4, lr_aff=0.0)
    tol = 5e-3                                                    # loose on purpose: stops after a few iterations
    want = wl.mapping_adam(w['frames'], w['edges'], 30, stop_tol=tol, **lrs)
    assert 2 <= want['steps'] < 30
    hw = HostWindows([w])
    host.window_poses_host(C.byref(hw.c))
    for _ in range(30):
        hw.edge_gradients()
        host.window_update_host(C.byref(hw.c), hw.out_pair.ctypes.data, hw.out_gk.ctypes.data, lrs['lr_pose'],
                                lrs['lr_k'], lrs['lr_aff'], 0.9, 0.999, 1e-8, tol)
    assert hw.win_state[0, 3] == 1.0 and hw.win_state[0, 0] == want['steps']
    np.testing.assert_allclose(hw.k[3:6], want['k'][1].numpy(), atol=1e-6)


def _golden_window():
    z = np.load(os.path.join(HERE, "golden", "mapping_window.npz"))
    shape = {key[6:]: z[key].item() for key in z.files if key.startswith("shape_")}
    return z, syn.mapping_window(**shape), shape


def test_mapping_oracle_is_pinned_to_the_reference_caller():
    """tests/golden/mapping_window.npz holds what the reference's OWN `Odometery.mapping` (odometery/odometery.py:687-967,
    run unmodified with only lietorch stubbed, tests/golden/make_golden_callers.py) leaves after 8 iterations of a
    3-keyframe window: oracle/window_loop.py must reproduce it bit for bit."""
    from oracle import window_loop as wl
    z, w, shape = _golden_window()
    lr = z["lr"]
    got = wl.mapping_adam(w['frames'], w['edges'], int(z["iters"]), lr_pose=float(lr[0]), lr_k=float(lr[1]),
                          lr_aff=float(lr[2]), stop_tol=float(z["stop_tol"]))
    assert got['steps'] == int(z["iters"])
    for f in range(len(w['frames'])):
        assert np.array_equal(got['T'][f].numpy(), z["T"][f]), f"pose of frame {f}"
        if not np.isnan(z["aff"][f]).any():
            assert np.array_equal(got['aff'][f].numpy(), z["aff"][f]), f"brightness of frame {f}"
        if f < shape['n_kf']:
            assert np.array_equal(got['k'][f].numpy(), z["k"][f]), f"seeds of frame {f}"
    # the fixture is not trivial: every optimised pose moved by ~ lr_pose * iterations, the first keyframe's did not
    assert np.array_equal(z["T"][0], w['frames'][0]['T'].numpy())
    assert 5e-4 < np.abs(z["T"][1] - w['frames'][1]['T'].numpy()).max() < 1e-3
    assert np.abs(z["k"][0] - w['frames'][0]['k'].numpy()).max() > 5e-2      # window not full: first keyframe's seeds move


def test_host_window_update_matches_the_reference_caller_golden(host):
    """The kernel's arithmetic (host build) with port gradients against the reference's own mapping() result."""
    z, w, shape = _golden_window()
    lr = z["lr"]
    hw = HostWindows([w])
    host.window_poses_host(C.byref(hw.c))
    for _ in range(int(z["iters"])):
        hw.edge_gradients()
        host.window_update_host(C.byref(hw.c), hw.out_pair.ctypes.data, hw.out_gk.ctypes.data, float(lr[0]), float(lr[1]),
                                float(lr[2]), 0.9, 0.999, 1e-8, float(z["stop_tol"]))
    assert hw.win_state[0, 0] == int(z["iters"]) and hw.win_state[0, 3] == 0.0
    np.testing.assert_allclose(hw.frame_T.reshape(-1, 4, 4), z["T"], atol=2e-7)
    np.testing.assert_allclose(hw.k.reshape(shape['n_kf'], -1), z["k"], atol=2e-6)
    ok = ~np.isnan(z["aff"]).any(axis=1)
    np.testing.assert_allclose(hw.frame_aff[ok], z["aff"][ok], atol=2e-7)
