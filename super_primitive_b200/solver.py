"""Batched, device-resident alignment of many independent (source keyframe, target frame) problems.

This is the "many problems per launch" driver of SURVEY.md section 7.1 step 7: every problem's
compact geometry, images, pose, log-depth seeds and affine terms live in HBM; descriptor arrays
(``SpbGeom[]``, ``SpbPair[]``) are uploaded once; one fused kernel launch per iteration covers
all problems (grid.y = problem) and a second small kernel performs the per-problem damped
Schur-complement solve and the SE(3) retraction -- no host synchronisation inside the loop, so a
whole optimisation can be captured in a CUDA graph.

Two iteration kinds:
  * ``gn_step``   IRLS Gauss-Newton / LM on the reference's L1 objective (extension named by
                  BASELINE.json; the reference itself only has Adam + autograd)
  * ``grad_step`` cost + first-order gradient (what the reference's backward() yields), for
                  Adam-parity loops driven from the host
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .geometry import CompactGeometry, _f32c, _stream, pack_rgba


def _struct_array_to_device(arr, device):
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


class AlignmentBatch:
    """n independent two-frame problems.

    problems: list of dicts with keys
        geom      CompactGeometry of the source keyframe (may be shared between problems)
        src_rgb   [3][n_pad] cached source samples at the level   } ``geom.level_buffers(image)``
        pack      [n_tiles][PACK_WORDS] tile-major level buffer   }
        trg_rgba  (Hl,Wl,4) packed target level image
        K_trg     (3,3) target intrinsics
        pose      (4,4) initial source->target transform
        k         (N,) initial log-depth seeds
        aff_src, aff_trg   optional (2,) tensors
        tau       optional front-of-camera threshold (default 1e-7)
        levels    optional list, COARSE -> FINE, of dict(src_rgb, pack, trg_rgba): the image pyramid of the pair
                  (geometry, pose, seeds and brightness terms are shared by the levels, as keyframe_pyramid(geo_down=False)
                  shares them, image/keyframe.py:125-146).  Without it the problem has the one level given by the
                  top-level src_rgb / pack / trg_rgba.  Every problem of a batch must have the same number of levels.
    """

    def __init__(self, problems, with_affine=False, irls_eps=1e-3, lam0=1e-3, hold_depth=False):
        lib = nat.lib()
        self.n = n = len(problems)
        if n < 1:
            raise ValueError("empty batch")
        dev = problems[0]['trg_rgba'].device
        self.device = dev
        self.with_affine = bool(with_affine)
        # hold_depth: the GN/LM loop moves only the pose (+ target affine); the log-depth seeds stay (the reference's
        # tracker, odometery/odometery.py:303-310).  For the Adam loop pass lr_k=0.
        self.hold_depth = bool(hold_depth)
        self.irls_eps = float(irls_eps)
        # unique geometries
        geoms, gidx = [], []
        seen = {}
        for p in problems:
            g = p['geom']
            if id(g) not in seen:
                seen[id(g)] = len(geoms)
                geoms.append(g)
            gidx.append(seen[id(g)])
        self.geoms = geoms
        seg_cnt = np.array([p['geom'].N for p in problems], dtype=np.int32)
        seg_off = np.zeros(n, dtype=np.int32)
        seg_off[1:] = np.cumsum(seg_cnt)[:-1]
        self.seg_total = int(seg_cnt.sum())
        self.seg_cnt_host, self.seg_off_host = seg_cnt, seg_off
        self.max_tiles = max(g.n_tiles for g in geoms)
        self.points_total = int(sum(p['geom'].P for p in problems))
        self._P = [int(p['geom'].P) for p in problems]
        self.pts_per_problem = torch.tensor(self._P, dtype=torch.float32, device=dev)
        # state
        self.poses = torch.stack([_f32c(p['pose']).reshape(16) for p in problems]).contiguous()
        self.k = torch.cat([_f32c(p['k']).reshape(-1) for p in problems]).contiguous()
        self.K_trg = torch.stack([_f32c(p['K_trg']).reshape(9) for p in problems]).contiguous()
        zero2 = torch.zeros(2, dtype=torch.float32, device=dev)
        self.aff_src = torch.stack([_f32c(p.get('aff_src', zero2) if p.get('aff_src') is not None else zero2)
                                    for p in problems]).contiguous()
        self.aff_trg = torch.stack([_f32c(p.get('aff_trg', zero2) if p.get('aff_trg') is not None else zero2)
                                    for p in problems]).contiguous()
        def lv_of(p):
            lv = p.get('levels')
            return list(lv) if lv else [dict(src_rgb=p['src_rgb'], pack=p['pack'], trg_rgba=p['trg_rgba'])]

        plv = [lv_of(p) for p in problems]
        self.n_levels = len(plv[0])
        if any(len(lv) != self.n_levels for lv in plv):
            raise ValueError("every problem of a batch needs the same number of pyramid levels")
        self._keep_lv = [[(lv[l]['src_rgb'], lv[l]['trg_rgba'], lv[l]['pack']) for lv in plv]
                         for l in range(self.n_levels)]
        use_aff = self.with_affine or any(p.get('aff_src') is not None for p in problems)
        self.use_affine = use_aff
        # descriptor arrays: one SpbPair array per pyramid level (only the image-derived pointers differ)
        garr = (nat.SpbGeom * len(geoms))()
        for i, g in enumerate(geoms):
            garr[i] = g.c
        self.d_pairs_lv = []
        for l in range(self.n_levels):
            parr = (nat.SpbPair * n)()
            for i, p in enumerate(problems):
                src_rgb, trg_rgba, pack = self._keep_lv[l][i]
                q = parr[i]
                q.trg_rgba = trg_rgba.data_ptr()
                q.src_rgb = src_rgb.data_ptr()
                q.tile_pack = pack.data_ptr()
                q.K_trg = self.K_trg[i].data_ptr()
                q.pose = self.poses[i].data_ptr()
                q.k = self.k.data_ptr() + 4 * int(seg_off[i])
                q.aff_src = self.aff_src[i].data_ptr() if use_aff else None
                q.aff_trg = self.aff_trg[i].data_ptr() if use_aff else None
                q.geom = gidx[i]
                q.Hl, q.Wl = trg_rgba.shape[0], trg_rgba.shape[1]
                q.tau = float(p.get('tau', 1e-7))
            self.d_pairs_lv.append(_struct_array_to_device(parr, dev))
        self.level = self.n_levels - 1                       # finest
        self.d_pairs = self.d_pairs_lv[self.level]
        self._keep = self._keep_lv[self.level]
        self.d_geoms = _struct_array_to_device(garr, dev)
        self.d_seg_off = torch.from_numpy(seg_off).to(dev)
        self.d_seg_cnt = torch.from_numpy(seg_cnt).to(dev)
        # workspaces / outputs
        self.ctas = lib.spb_gn_ctas(self.max_tiles, n)
        self.work_stride = int(lib.spb_gn_work_stride(self.max_tiles, n))
        self.work = torch.empty(n * self.work_stride, dtype=torch.float32, device=dev)
        self.gn_pair = torch.zeros((n, nat.GN_PAIR_NOUT), dtype=torch.float32, device=dev)
        self.gn_seg = torch.zeros((self.seg_total, nat.GN_SEG_NOUT), dtype=torch.float32, device=dev)
        self.out_pair = torch.zeros((n, nat.PAIR_NOUT), dtype=torch.float32, device=dev)
        self.out_gk = torch.zeros(self.seg_total, dtype=torch.float32, device=dev)
        self.lm_state = torch.zeros((n, nat.LM_NSTATE), dtype=torch.float32, device=dev)
        self.lm_state[:, 0] = lam0
        pf, sf = C.c_int64(), C.c_int64()
        nat.check(lib.spb_lm_saved_floats(n, self.seg_total, C.byref(pf), C.byref(sf)), "spb_lm_saved_floats")
        self.saved_pair = torch.zeros(pf.value, dtype=torch.float32, device=dev)
        self.saved_seg = torch.zeros(sf.value, dtype=torch.float32, device=dev)
        self.launches = 0

    # ---- per-iteration entry points ------------------------------------------------------------
    def gn_accumulate(self, ev=None):
        """``ev`` = (torch.cuda.Event, torch.cuda.Event) recorded around the fused kernel alone."""
        e0, e1 = (None, None) if ev is None else (ev[0].cuda_event, ev[1].cuda_event)
        nat.check(nat.lib().spb_gn_accumulate(self.d_geoms.data_ptr(), self.d_pairs.data_ptr(),
                                              self.d_seg_off.data_ptr(), self.n, self.max_tiles, self.irls_eps,
                                              1 if self.with_affine else (2 if self.use_affine else 0),
                                              self.work.data_ptr(), self.work_stride,
                                              self.gn_pair.data_ptr(), self.gn_seg.data_ptr(), e0, e1, _stream()),
                  "spb_gn_accumulate")
        self.launches += 2

    def lm_update(self):
        nat.check(nat.lib().spb_lm_update(self.gn_pair.data_ptr(), self.gn_seg.data_ptr(), self.d_seg_off.data_ptr(),
                                          self.d_seg_cnt.data_ptr(), self.n, 1 if self.with_affine else 0,
                                          1 if self.hold_depth else 0,
                                          self.poses.data_ptr(), self.k.data_ptr(),
                                          self.aff_trg.data_ptr() if self.with_affine else None,
                                          self.lm_state.data_ptr(), self.saved_pair.data_ptr(),
                                          self.saved_seg.data_ptr(), _stream()), "spb_lm_update")
        self.launches += 1

    def gn_step(self, ev=None):
        """One GN/LM iteration for every problem in two launches: the fused residual+Jacobian+normal-equation
        kernel, then finalize + damped solve + retraction (``spb_gn_iterate``)."""
        e0, e1 = (None, None) if ev is None else (ev[0].cuda_event, ev[1].cuda_event)
        nat.check(nat.lib().spb_gn_iterate(self.d_geoms.data_ptr(), self.d_pairs.data_ptr(), self.d_seg_off.data_ptr(),
                                           self.d_seg_cnt.data_ptr(), self.n, self.max_tiles, self.irls_eps,
                                           1 if self.with_affine else (2 if self.use_affine else 0),
                                           1 if self.hold_depth else 0,
                                           self.work.data_ptr(), self.work_stride, self.gn_pair.data_ptr(),
                                           self.gn_seg.data_ptr(), self.poses.data_ptr(), self.k.data_ptr(),
                                           self.aff_trg.data_ptr() if self.with_affine else None,
                                           self.lm_state.data_ptr(), self.saved_pair.data_ptr(),
                                           self.saved_seg.data_ptr(), e0, e1, _stream()), "spb_gn_iterate")
        self.launches += 2

    def grad_step(self, ev=None):
        """Cost + first-order gradient for every problem (Adam-parity quantities): fills
        ``out_pair`` (n,16) and ``out_gk`` (seg_total,)."""
        e0, e1 = (None, None) if ev is None else (ev[0].cuda_event, ev[1].cuda_event)
        nat.check(nat.lib().spb_grad_accumulate(self.d_geoms.data_ptr(), self.d_pairs.data_ptr(),
                                                self.d_seg_off.data_ptr(), self.n, self.max_tiles,
                                                1 if self.use_affine else 0,
                                                self.work.data_ptr(), self.work_stride, self.out_pair.data_ptr(),
                                                self.out_gk.data_ptr(), e0, e1, _stream()), "spb_grad_accumulate")
        self.launches += 2

    def _adam_buffers(self):
        if getattr(self, "adam_pair", None) is None:
            self.adam_pair = torch.zeros((self.n, nat.ADAM_PAIR), dtype=torch.float32, device=self.device)
            self.adam_seg = torch.zeros((self.seg_total, nat.ADAM_SEG), dtype=torch.float32, device=self.device)

    def adam_step(self, ev=None, lr_pose=1e-2, lr_k=1e-3, lr_aff=5e-3, betas=(0.9, 0.999), eps=1e-8):
        """One first-order iteration for every problem in two launches (``spb_adam_iterate``): the fused
        residual+gradient kernel, then finalize + torch.optim.Adam update of (twist increment, log-depth seeds,
        target affine when ``with_affine``) + retraction ``T <- Exp(delta) T`` -- the reference's optimiser
        (odometery/two_frame_sfm.py:117-121 learning rates; tracker bookkeeping odometery/odometery.py:386-403)
        kept on the device, no host synchronisation."""
        self._adam_buffers()
        e0, e1 = (None, None) if ev is None else (ev[0].cuda_event, ev[1].cuda_event)
        nat.check(nat.lib().spb_adam_iterate(self.d_geoms.data_ptr(), self.d_pairs.data_ptr(), self.d_seg_off.data_ptr(),
                                             self.d_seg_cnt.data_ptr(), self.n, self.max_tiles,
                                             1 if self.with_affine else (2 if self.use_affine else 0),
                                             self.work.data_ptr(), self.work_stride, self.out_pair.data_ptr(),
                                             self.out_gk.data_ptr(), self.poses.data_ptr(), self.k.data_ptr(),
                                             self.aff_trg.data_ptr() if self.with_affine else None,
                                             self.adam_pair.data_ptr(), self.adam_seg.data_ptr(), float(lr_pose),
                                             float(lr_k), float(lr_aff), float(betas[0]), float(betas[1]), float(eps),
                                             e0, e1, _stream()), "spb_adam_iterate")
        self.launches += 2

    def adam_update(self, lr_pose=1e-2, lr_k=1e-3, lr_aff=5e-3, betas=(0.9, 0.999), eps=1e-8):
        """The update alone, from the gradients ``grad_step`` left in ``out_pair`` / ``out_gk``."""
        self._adam_buffers()
        nat.check(nat.lib().spb_adam_update(self.out_pair.data_ptr(), self.out_gk.data_ptr(), self.d_seg_off.data_ptr(),
                                            self.d_seg_cnt.data_ptr(), self.n, 1 if self.with_affine else 0,
                                            self.poses.data_ptr(), self.k.data_ptr(),
                                            self.aff_trg.data_ptr() if self.with_affine else None,
                                            self.adam_pair.data_ptr(), self.adam_seg.data_ptr(), float(lr_pose),
                                            float(lr_k), float(lr_aff), float(betas[0]), float(betas[1]), float(eps),
                                            _stream()), "spb_adam_update")
        self.launches += 1

    # ---- coarse-to-fine schedule ------------------------------------------------------------------
    def set_level(self, level):
        """Select the pyramid level the next iterations run on (0 = coarsest).  Poses, seeds, brightness terms and
        the optimiser state carry over; the LM acceptance baseline is re-armed because the cost of another level is
        not comparable (the damping factor is kept)."""
        if not 0 <= level < self.n_levels:
            raise ValueError(f"level {level} outside 0..{self.n_levels - 1}")
        if level != self.level:
            self.level = level
            self.d_pairs = self.d_pairs_lv[level]
            self._keep = self._keep_lv[level]
            self.lm_state[:, 2] = 0.0

    def _schedule(self, iters):
        """iters: int (all on the current level) or a per-level sequence, coarse -> fine, the reference's
        `for pyr_level ... for i in range(steps)` (odometery/two_frame_sfm.py:150-155, odometery/odometery.py:376-384)."""
        if isinstance(iters, int):
            return [(self.level, iters)]
        iters = list(iters)
        if len(iters) != self.n_levels:
            raise ValueError(f"{len(iters)} iteration counts for {self.n_levels} pyramid levels")
        return [(l, int(n)) for l, n in enumerate(iters) if int(n) > 0]

    def run_adam(self, iters, **kw):
        for level, n in self._schedule(iters):
            self.set_level(level)
            for _ in range(n):
                self.adam_step(**kw)

    def _capture(self, step, iters, kw):
        sched = self._schedule(iters)
        saved = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for level, _ in sched:                     # warm-up outside capture (module load, lazy allocations) ...
                self.set_level(level)
                step(**kw)
        torch.cuda.current_stream().wait_stream(s)
        self._restore(saved)                           # ... which must not count as an iteration
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for level, n in sched:
                self.set_level(level)
                for _ in range(n):
                    step(**kw)
        return graph

    def _state_tensors(self):
        ts = [self.poses, self.k, self.aff_trg, self.lm_state, self.saved_pair, self.saved_seg]
        if getattr(self, "adam_pair", None) is not None:
            ts += [self.adam_pair, self.adam_seg]
        return ts

    def _snapshot(self):
        return (self.level, [t.clone() for t in self._state_tensors()])

    def _restore(self, saved):
        level, vals = saved
        for t, v in zip(self._state_tensors(), vals):
            t.copy_(v)
        self.level = level
        self.d_pairs = self.d_pairs_lv[level]
        self._keep = self._keep_lv[level]

    def capture_adam(self, iters, **kw):
        """CUDA-graph `iters` first-order iterations (int, or per level coarse -> fine); replaying the graph applies
        exactly those iterations (the warm-up step taken before capture is rolled back).  Returns the graph."""
        self._adam_buffers()
        return self._capture(self.adam_step, iters, kw)

    def grad_costs(self):
        """mean |r| per problem at the parameters of the last gradient evaluation (grad_step / adam_step)."""
        return self.out_pair[:, 0]

    def run_gn(self, iters):
        for level, n in self._schedule(iters):
            self.set_level(level)
            for _ in range(n):
                self.gn_step()

    def capture_gn(self, iters):
        """CUDA-graph `iters` GN iterations (int, or per level coarse -> fine; launch-bound small batches); replaying
        applies exactly those iterations.  Returns the graph."""
        return self._capture(self.gn_step, iters, {})

    # ---- results -------------------------------------------------------------------------------
    def costs(self):
        """mean |r| per problem at the last evaluated parameters (GN accumulation)."""
        return self.gn_pair[:, nat.GN_NA + 8] / (3.0 * self.pts_per_problem)

    def poses_matrix(self):
        return self.poses.reshape(self.n, 4, 4)

    def k_of(self, i):
        o = int(self.seg_off_host[i])
        return self.k[o:o + int(self.seg_cnt_host[i])]

    def k_padded(self):
        """(n, N_max) log-depth seeds padded with NaN (for the final cross-rank gather)."""
        nmax = int(self.seg_cnt_host.max())
        out = torch.full((self.n, nmax), float('nan'), dtype=torch.float32, device=self.device)
        for i in range(self.n):
            out[i, :int(self.seg_cnt_host[i])] = self.k_of(i)
        return out

    def algorithmic_bytes_per_iter(self, gn=True):
        """per pair 24 P + 4 C Hl Wl + outputs (C = 3)."""
        total = 0
        for i, (src_rgb, trg, _pack) in enumerate(self._keep):
            P = self._P[i]
            Hl, Wl = trg.shape[0], trg.shape[1]
            N = int(self.seg_cnt_host[i])
            outs = 4 * (12 + N + 4 + 1) + (4 * (21 + 6 * N + N) if gn else 0)
            total += 24 * P + 12 * Hl * Wl + outs
        return total


def make_problem(src_kf, trg_image, trg_K, pose, k, geom=None, aff_src=None, aff_trg=None, tau=1e-7, levels=None):
    """Convenience: build one problem dict from a source keyframe (dense) and a planar target image.
    levels = (start_level, end_level): additionally the image pyramid of both frames, `keyframe_pyramid`'s levels
    (image/keyframe.py:77-148: 3x3 blur + decimation per level, geometry shared), coarse -> fine."""
    if geom is None:
        geom = getattr(src_kf, "_spb_geometry", None)          # handover.CompactKeyFrame
    if geom is None:
        geom = CompactGeometry(src_kf.keypoint_regions, src_kf.get_logdepth(), src_kf.keypoints, src_kf.K)
    out = dict(geom=geom, K_trg=trg_K, pose=pose, k=k, aff_src=aff_src, aff_trg=aff_trg, tau=tau)
    if levels is None:
        src_rgb, pack = geom.level_buffers(src_kf.image)
        out.update(src_rgb=src_rgb, pack=pack, trg_rgba=pack_rgba(trg_image)[0])
        return out
    from .pyramid import _levels, _pyr_down
    start, end = levels
    lv = []
    for s_img, t_img in zip(_levels(_f32c(src_kf.image[:3]), start, end, _pyr_down),
                            _levels(_f32c(trg_image[:3]), start, end, _pyr_down)):
        src_rgb, pack = geom.level_buffers(s_img)
        lv.append(dict(src_rgb=src_rgb, pack=pack, trg_rgba=pack_rgba(t_img)[0], src_image=s_img, trg_image=t_img))
    out.update(levels=lv, src_rgb=lv[-1]['src_rgb'], pack=lv[-1]['pack'], trg_rgba=lv[-1]['trg_rgba'])
    return out
