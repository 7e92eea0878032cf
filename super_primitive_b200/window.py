"""Device-resident windowed mapping: the reference's coupled optimisation over a window of keyframes and supporting
frames (odometery/odometery.py:687-915) as batched launches.

The reference loops on the host: per iteration and per source keyframe it rebuilds every relative pose with
``pose_to_mat`` / ``torch.linalg.inv`` products (:775-829), calls ``photomeric_cost_batch`` (:833), sums the losses,
runs autograd + ONE ``torch.optim.Adam`` over all poses / seeds / brightness terms (:845-858), folds the pose
increments, renormalises (:861-882) and syncs on ``loss.item()`` for the early stop (:907-915).  Here every edge
(source keyframe -> target frame) of every window is one pair of the batched gradient launch and
``spb_window_iterate`` (include/spb200.h) performs the coupled update on the device: three launches per iteration for
any number of windows, no host synchronisation, CUDA-graph capturable.

Frame / window description (shared with the oracle, oracle/window_loop.py):
    window = {'frames': [frame, ...], 'edges': [(src, trg), ...]}       indices local to the window
    frame  = {'T': (4,4) camera-to-world, 'image': (3,Hl,Wl) level image, 'K': (3,3),
              'aff': (2,) or None, 'opt_pose': bool, 'opt_aff': bool,
              'kf': KeyFrame or None (sources), 'k': (N,) seeds, 'opt_seeds': bool}
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .geometry import CompactGeometry, _f32c, _stream, pack_rgba
from .solver import _struct_array_to_device


def window_layout(windows):
    """Index arrays of SpbWindow for a list of windows (host logic, no device access).

    Frames and edges are numbered window after window; an edge's weight is 1 / (number of targets of its source)
    (loss = sum_src mean_b cost, odometery/odometery.py:845-851); seeds are laid out keyframe after keyframe and the
    per-edge gradient slots edge after edge."""
    win_frame_off, win_edge_off = [0], [0]
    edge_src, edge_trg, edge_w, edge_seg_off = [], [], [], []
    frame_seg_off, frame_seg_cnt, frame_flags = [], [], []
    seg_total = gk_total = 0
    for w in windows:
        frames, edges = w['frames'], w['edges']
        if not frames or not edges:
            raise ValueError("a window needs frames and edges")
        base = win_frame_off[-1]
        n_out = {}
        for s, t in edges:
            if not (0 <= s < len(frames) and 0 <= t < len(frames)) or s == t:
                raise ValueError(f"bad edge ({s}, {t})")
            if frames[s].get('kf') is None:
                raise ValueError(f"edge source {s} is not a keyframe")
            n_out[s] = n_out.get(s, 0) + 1
        for f in frames:
            n = 0 if f.get('kf') is None else int(f['k'].shape[0])
            frame_seg_off.append(seg_total)
            frame_seg_cnt.append(n)
            seg_total += n
            frame_flags.append((nat.WIN_OPT_POSE if f.get('opt_pose') else 0) |
                               (nat.WIN_OPT_AFF if f.get('opt_aff') and f.get('aff') is not None else 0) |
                               (nat.WIN_OPT_SEEDS if f.get('opt_seeds') and n > 0 else 0))
        for s, t in edges:
            edge_src.append(base + s)
            edge_trg.append(base + t)
            edge_w.append(1.0 / n_out[s])
            edge_seg_off.append(gk_total)
            gk_total += frame_seg_cnt[base + s]
        win_frame_off.append(base + len(frames))
        win_edge_off.append(win_edge_off[-1] + len(edges))
    i32 = lambda a: np.asarray(a, dtype=np.int32)      # noqa: E731
    return dict(n_windows=len(windows), n_frames=win_frame_off[-1], n_edges=win_edge_off[-1], seg_total=seg_total,
                gk_total=gk_total, win_frame_off=i32(win_frame_off), win_edge_off=i32(win_edge_off),
                edge_src=i32(edge_src), edge_trg=i32(edge_trg), edge_w=np.asarray(edge_w, dtype=np.float32),
                edge_seg_off=i32(edge_seg_off), frame_seg_off=i32(frame_seg_off), frame_seg_cnt=i32(frame_seg_cnt),
                frame_flags=np.asarray(frame_flags, dtype=np.uint8))


class MappingWindows:
    """Any number of independent mapping windows, optimised together on the device."""

    def __init__(self, windows, tau=1e-6):
        lib = nat.lib()
        lay = window_layout(windows)
        self.layout = lay
        frames = [f for w in windows for f in w['frames']]
        dev = frames[0]['image'].device
        if dev.type != 'cuda':
            raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
        self.device = dev
        F, E = lay['n_frames'], lay['n_edges']
        self.n_windows, self.n_frames, self.n_edges = lay['n_windows'], F, E
        self.use_affine = any(f.get('aff') is not None for f in frames)
        if self.use_affine and not all(f.get('aff') is not None for f in frames):
            raise ValueError("brightness terms must be given for every frame or for none")
        # state
        self.frame_T = torch.stack([_f32c(f['T']).reshape(16) for f in frames]).contiguous()
        self.frame_aff = torch.stack([_f32c(f['aff']).reshape(2) for f in frames]).contiguous() if self.use_affine else None
        ks = [_f32c(f['k']).reshape(-1) for f in frames if f.get('kf') is not None]
        self.k = torch.cat(ks).contiguous()
        self.K = torch.stack([_f32c(f['K']).reshape(9) for f in frames]).contiguous()
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)      # noqa: E731
        self.edge_pose = z(E, 16)
        self.adam_frame = z(F, nat.WIN_ADAM_FRAME)
        self.adam_seg = z(max(lay['seg_total'], 1), nat.ADAM_SEG)
        self.win_state = z(self.n_windows, nat.WIN_NSTATE)
        self.edge_tw = z(E, 12)
        self.out_pair = z(E, nat.PAIR_NOUT)
        self.out_gk = z(max(lay['gk_total'], 1))
        up = lambda a: torch.from_numpy(a).to(dev)                           # noqa: E731
        self._idx = {name: up(lay[name]) for name in ('win_frame_off', 'win_edge_off', 'edge_src', 'edge_trg', 'edge_w',
                                                      'edge_seg_off', 'frame_seg_off', 'frame_seg_cnt', 'frame_flags')}
        # per-frame device buffers: compact geometry + level buffers for sources, packed RGBA for targets
        self.geoms, self._src, self._rgba = [], {}, {}
        gidx = {}
        is_trg = set(int(t) for t in lay['edge_trg'])
        for i, f in enumerate(frames):
            kf = f.get('kf')
            if kf is not None:
                g = f.get('geom') or getattr(kf, "_spb_geometry", None) or \
                    CompactGeometry(kf.keypoint_regions, kf.get_logdepth(), kf.keypoints, kf.K)
                if int(f['k'].shape[0]) != g.N:
                    raise AssertionError("one log-depth seed per segment expected")
                gidx[i] = len(self.geoms)
                self.geoms.append(g)
                self._src[i] = g.level_buffers(kf.image)
            if i in is_trg:
                self._rgba[i] = pack_rgba(f['image'])[0]
        self.max_tiles = max(g.n_tiles for g in self.geoms)
        self._P = np.array([self.geoms[gidx[int(s)]].P for s in lay['edge_src']], dtype=np.int64)
        # descriptors
        garr = (nat.SpbGeom * len(self.geoms))()
        for i, g in enumerate(self.geoms):
            garr[i] = g.c
        parr = (nat.SpbPair * E)()
        for e in range(E):
            s, t = int(lay['edge_src'][e]), int(lay['edge_trg'][e])
            q = parr[e]
            q.trg_rgba = self._rgba[t].data_ptr()
            q.src_rgb = self._src[s][0].data_ptr()
            q.tile_pack = self._src[s][1].data_ptr()
            q.K_trg = self.K[t].data_ptr()
            q.pose = self.edge_pose[e].data_ptr()
            q.k = self.k.data_ptr() + 4 * int(lay['frame_seg_off'][s])
            q.aff_src = self.frame_aff[s].data_ptr() if self.use_affine else None
            q.aff_trg = self.frame_aff[t].data_ptr() if self.use_affine else None
            q.geom = gidx[s]
            q.Hl, q.Wl = self._rgba[t].shape[0], self._rgba[t].shape[1]
            q.tau = float(tau)           # batch path threshold, core/dense_optim_batch.py:15
        self.d_geoms = _struct_array_to_device(garr, dev)
        self.d_pairs = _struct_array_to_device(parr, dev)
        self.ctas = lib.spb_gn_ctas(self.max_tiles, E)
        self.work_stride = int(lib.spb_gn_work_stride(self.max_tiles, E))
        self.work = torch.empty(E * self.work_stride, dtype=torch.float32, device=dev)
        self.c = nat.SpbWindow(self.n_windows, F, E, lay['seg_total'],
                               *(self._idx[n].data_ptr() for n in ('win_frame_off', 'win_edge_off', 'edge_src', 'edge_trg',
                                                                    'edge_w', 'edge_seg_off', 'frame_seg_off',
                                                                    'frame_seg_cnt', 'frame_flags')),
                               self.frame_T.data_ptr(), None if self.frame_aff is None else self.frame_aff.data_ptr(),
                               self.k.data_ptr(), self.edge_pose.data_ptr(), self.adam_frame.data_ptr(),
                               self.adam_seg.data_ptr(), self.win_state.data_ptr(), self.edge_tw.data_ptr())
        nat.check(lib.spb_window_poses(C.byref(self.c), _stream()), "spb_window_poses")
        self.launches = 1

    # ---- iteration ---------------------------------------------------------------------------------
    def step(self, lr_pose=1e-4, lr_k=1e-2, lr_aff=1e-5, betas=(0.9, 0.999), eps=1e-8, stop_tol=0.0, ev=None):
        """One mapping iteration of every window (reference learning rates: odometery/odometery.py:579-586)."""
        e0, e1 = (None, None) if ev is None else (ev[0].cuda_event, ev[1].cuda_event)
        nat.check(nat.lib().spb_window_iterate(self.d_geoms.data_ptr(), self.d_pairs.data_ptr(), C.byref(self.c),
                                               self.max_tiles, 1 if self.use_affine else 0, self.work.data_ptr(),
                                               self.work_stride, self.out_pair.data_ptr(), self.out_gk.data_ptr(),
                                               float(lr_pose), float(lr_k), float(lr_aff), float(betas[0]),
                                               float(betas[1]), float(eps), float(stop_tol), e0, e1, _stream()),
                  "spb_window_iterate")
        self.launches += 3

    def run(self, iters, **kw):
        for _ in range(iters):
            self.step(**kw)

    def capture(self, iters, **kw):
        """CUDA-graph ``iters`` iterations; replaying the graph applies exactly those (the eager warm-up step taken
        before capture -- module load, lazy allocations -- is rolled back).  Returns the graph."""
        state = [t for t in (self.frame_T, self.frame_aff, self.k, self.edge_pose, self.adam_frame, self.adam_seg,
                             self.win_state) if t is not None]
        saved = [t.clone() for t in state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.step(**kw)
        torch.cuda.current_stream().wait_stream(s)
        for t, v in zip(state, saved):
            t.copy_(v)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(iters):
                self.step(**kw)
        return graph

    # ---- results -----------------------------------------------------------------------------------
    def poses(self):
        """(n_frames, 4, 4) camera-to-world poses."""
        return self.frame_T.reshape(self.n_frames, 4, 4)

    def seeds_of(self, frame):
        """log-depth seeds of global frame index ``frame`` (a keyframe)."""
        o, n = int(self.layout['frame_seg_off'][frame]), int(self.layout['frame_seg_cnt'][frame])
        return self.k[o:o + n]

    def losses(self):
        """loss of every window at the parameters of the last evaluated iteration."""
        return self.win_state[:, 1]

    def steps_done(self):
        return self.win_state[:, 0]

    def converged(self):
        return self.win_state[:, 3] != 0

    def result_vectors(self):
        """One flat float32 vector per window -- [frame poses (F*16) | seeds of its keyframes | brightness terms (F*2,
        if any) | loss] -- the payload of the final cross-rank gather (`shard.gather_ragged`)."""
        lay, out = self.layout, []
        for w in range(self.n_windows):
            f0, f1 = int(lay['win_frame_off'][w]), int(lay['win_frame_off'][w + 1])
            k0 = int(lay['frame_seg_off'][f0])
            k1 = int(lay['frame_seg_off'][f1 - 1]) + int(lay['frame_seg_cnt'][f1 - 1])
            parts = [self.frame_T[f0:f1].reshape(-1), self.k[k0:k1]]
            if self.frame_aff is not None:
                parts.append(self.frame_aff[f0:f1].reshape(-1))
            parts.append(self.win_state[w, 1:2])
            out.append(torch.cat(parts))
        return out

    def algorithmic_bytes_per_iter(self):
        """SURVEY.md section 8(d) summed over the edges: 24 P + 12 Hl Wl + outputs per edge."""
        total = 0
        for e in range(self.n_edges):
            t = int(self.layout['edge_trg'][e])
            N = int(self.layout['frame_seg_cnt'][int(self.layout['edge_src'][e])])
            Hl, Wl = self._rgba[t].shape[0], self._rgba[t].shape[1]
            total += 24 * int(self._P[e]) + 12 * Hl * Wl + 4 * (12 + N + 4 + 1)
        return total
