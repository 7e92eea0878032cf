"""The other BASELINE.json configurations for bench.py (`--workload c2levels | c3 | c4 | c5 | compaction`).

Every workload works on a GLOBAL list of independent units (pairs / tracked frames / depth-completion frames) that is
dealt to the ranks with `shard.shard_indices` (unit u belongs to rank u % world, no data-path collective), timed on the
device as the max over ranks, and whose per-unit results are collected once with `shard.gather_results`.  With more
than one rank, rank 0 afterwards rebuilds the shard of rank 1 from the unit seeds, runs it alone and requires the
gathered results of those units to be BIT-equal: a unit's result does not depend on where it ran.

    c2levels   BASELINE config 2 per pyramid level (160x120 / 320x240 / 640x480 targets, full-resolution geometry)
    c3         TUM-shape tracking (288x224): pose + brightness only, the reference's 300 iterations per frame
               (odometery/odometery.py:365-403) as the device-resident Adam iteration
    c4         VOID depth completion (640x480, 100 segments per frame): per frame mask compaction, per-segment lower
               median re-initialisation from the sparse depth and the average render
               (depth_completion/segment_based_completion.py:30-62)
    c5         stress: 1024x768, 256 segments per pair, GN/LM iteration over as many pairs as asked
    compaction one-time cost of the ordered mask compaction per keyframe (core/dense_optim.py:89-114 is what it replaces)

Synthetic inputs are generated ON the device (the dense (N,H,W) tensors of config 5 are 1 GB per keyframe).
"""
from __future__ import annotations

import json
import math
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
UNIT_ITERS = "iterations/s"


# ------------------------------------------------------------------------------------------------------------------
# context / timing helpers
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args, rank, world, device, dist):
        self.args, self.rank, self.world, self.device, self.dist = args, rank, world, device, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int(self, v):
        t = torch.tensor([v], dtype=torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed_steps(ctx, step, steps, warm, kernel_events=True):
    """`step(ev)` runs one step, recording the (start, stop) event pair `ev` around its dominant kernel when given.
    Returns (total ms for `steps` steps = max over ranks, mean kernel ms on this rank or None)."""
    for _ in range(warm):
        step(None)
    ctx.barrier()
    kev = None
    if kernel_events:
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in kev:
            a.record()
            b.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for i in range(steps):
        step(kev[i] if kev else None)
    e1.record()
    ctx.barrier()
    total = ctx.max_ms(e0.elapsed_time(e1))
    kern = float(np.mean([a.elapsed_time(b) for a, b in kev])) if kev else None
    return total, kern


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs on the device
# ------------------------------------------------------------------------------------------------------------------
def grid_shape(N, H, W):
    """N = gx * gy cells with cells as square as the image allows."""
    best = (N, 1)
    for gy in range(1, N + 1):
        if N % gy == 0:
            gx = N // gy
            if abs(math.log((W / gx) / (H / gy))) < abs(math.log((W / best[0]) / (H / best[1]))):
                best = (gx, gy)
    return best


def device_image(H, W, device, shift=(0.0, 0.0), seed=0, noise=0.01):
    """the generator of super_primitive_b200.synthetic.sinus_image, evaluated in float32 on the device"""
    ys = (torch.arange(H, dtype=torch.float32, device=device) + shift[1]) / H
    xs = (torch.arange(W, dtype=torch.float32, device=device) + shift[0]) / W
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    img = torch.stack([0.5 + 0.5 * torch.sin(2 * math.pi * (3 * xx + 2 * yy + c / 3.0)) for c in range(3)], 0)
    if noise > 0:
        g = torch.Generator(device=device).manual_seed(seed)
        img = img + (torch.rand(img.shape, generator=g, device=device) * 2 - 1) * noise
    return img.clamp_(0, 1)


def device_keyframe(H, W, N, device, seed=0, pad=2, noise=0.01):
    """Source keyframe with N SAM-like blob segments: a gx x gy grid of cells, each dilated by `pad` pixels (masks
    overlap like SAM's), per-segment log-depth ramp with its own tilt, keypoint at the cell centre."""
    from super_primitive_b200.keyframe import KeyFrame
    from super_primitive_b200.synthetic import pinhole
    gx, gy = grid_shape(N, H, W)
    cols = torch.arange(W, device=device)
    rows = torch.arange(H, device=device)
    cx0 = (torch.arange(gx, device=device) * W) // gx
    cx1 = ((torch.arange(gx, device=device) + 1) * W) // gx
    cy0 = (torch.arange(gy, device=device) * H) // gy
    cy1 = ((torch.arange(gy, device=device) + 1) * H) // gy
    in_x = (cols[None, :] >= (cx0[:, None] - pad)) & (cols[None, :] < (cx1[:, None] + pad))       # (gx, W)
    in_y = (rows[None, :] >= (cy0[:, None] - pad)) & (rows[None, :] < (cy1[:, None] + pad))       # (gy, H)
    masks = (in_y[:, None, :, None] & in_x[None, :, None, :]).reshape(N, H, W)
    x = (cols.to(torch.float32) / W)[None, None, :]
    y = (rows.to(torch.float32) / H)[None, :, None]
    tilt = (0.05 * torch.cos(torch.arange(N, dtype=torch.float32, device=device) + seed))[:, None, None]
    logd = (0.1 * x + tilt * y) * masks
    kp_r = ((cy0 + cy1) // 2)[:, None].expand(gy, gx).reshape(N)
    kp_c = ((cx0 + cx1) // 2)[None, :].expand(gy, gx).reshape(N)
    inv = 1.0 / (torch.tensor([H, W], dtype=torch.float32, device=device) - 1)
    keypoints = 2 * torch.stack([kp_r, kp_c], 1).to(torch.float32) * inv - 1
    img = device_image(H, W, device, seed=seed, noise=noise)
    return KeyFrame(img, pinhole(H, W).to(device), logd, keypoints, masks)


def unit_rng(unit):
    return torch.Generator().manual_seed(1000003 * (unit + 1))


def start_pose(unit):
    from super_primitive_b200.synthetic import small_pose
    g = unit_rng(unit)
    r = (torch.rand(6, generator=g) - 0.5)
    return small_pose(0.02 + 0.01 * float(r[0]), 0.004 + 0.004 * float(r[1]), -0.003 + 0.004 * float(r[2]),
                      0.003 + 0.002 * float(r[3]), -0.002 + 0.002 * float(r[4]), 0.0015 + 0.002 * float(r[5]))


# ------------------------------------------------------------------------------------------------------------------
# unit lists
# ------------------------------------------------------------------------------------------------------------------
def build_pair_units(units, H, W, N, device, n_geoms=2, levels=None, with_affine=False, pad=2):
    """Problems for the global unit ids `units` (deterministic in the unit id, so any rank can rebuild any unit).
    `n_geoms` distinct source geometries per build are cycled through by unit id (a geometry is a keyframe's masks;
    building one per pair at config 5 would read 1 GB of dense tensors per pair)."""
    from super_primitive_b200.geometry import CompactGeometry
    from super_primitive_b200.solver import make_problem
    from super_primitive_b200.keyframe import KeyFrame
    geoms = {}
    probs = []
    for u in units:
        gi = u % n_geoms
        if gi not in geoms:
            kf = device_keyframe(H, W, N, device, seed=gi, pad=pad)
            geoms[gi] = (CompactGeometry(kf.keypoint_regions, kf.logdepth_perseg, kf.keypoints, kf.K), kf.K, kf.keypoints)
            del kf
        geom, K, _ = geoms[gi]
        g = unit_rng(u)
        sh = (torch.rand(2, generator=g) * 2.0 + 0.5).tolist()
        src_img = device_image(H, W, device, seed=2 * u + 11)
        trg_img = device_image(H, W, device, shift=(sh[0], sh[1]), seed=2 * u + 12)
        k0 = (math.log(2.0) + (torch.rand(N, generator=g) * 0.1 - 0.05)).to(device)
        src = KeyFrame(src_img, K, None, None, None)
        aff = (torch.tensor([0.01, 0.0], device=device), torch.tensor([0.0, 0.01], device=device)) if with_affine else (None, None)
        p = make_problem(src, trg_img, K, start_pose(u).to(device), k0, geom=geom, aff_src=aff[0], aff_trg=aff[1],
                         levels=levels)
        p['unit'] = u
        probs.append(p)
    return probs


def shard_check(ctx, n_units, run_units, gathered, what):
    """With more than one rank: rank 0 rebuilds rank 1's shard from the unit seeds, runs it alone (same batch size =>
    same launch configuration) and compares bit for bit with what was gathered.  Returns the JSON fragment."""
    from super_primitive_b200.shard import shard_indices
    if ctx.world == 1:
        return {"ranks": 1, "checked": 0, "what": "single rank: nothing to compare"}
    out = None
    if ctx.rank == 0:
        units = shard_indices(n_units, 1, ctx.world)
        local = run_units(units)
        ii = torch.tensor(units, dtype=torch.int64, device=ctx.device)
        eq = all(torch.equal(a[ii].nan_to_num(-7.0), b.nan_to_num(-7.0)) for a, b in zip(gathered, local))
        out = {"ranks": ctx.world, "checked": len(units), "bit_equal": bool(eq),
               "what": f"rank 0 re-ran the {len(units)} units of rank 1 alone; gathered {what} of those units compared "
                       "bit for bit"}
    ctx.barrier()
    return out


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------
def line_base(ctx, metric, unit, value, steps, warm, total_ms, workload, scaling, extra_cfg):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": ctx.world, "steps": steps, "warmup": warm,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload=workload, parallelism=f"shard{ctx.world}",
                                                                 **extra_cfg)}


def run_c2levels(ctx):
    """config 2, every pyramid level: GN/LM iteration on 64 pairs per GPU (the headline's batch), geometry at full
    resolution, target + cached source samples of the level."""
    from super_primitive_b200.solver import AlignmentBatch
    from super_primitive_b200.shard import shard_indices
    a = ctx.args
    H, W, N = 480, 640, 64
    n_units = a.pairs * ctx.world
    probs = build_pair_units(shard_indices(n_units, ctx.rank, ctx.world), H, W, N, ctx.device, n_geoms=4, levels=(0, 3), pad=4)
    batch = AlignmentBatch(probs)
    peak, src = hbm_peak()
    steps, warm = max(1, a.steps), max(3, a.warmup)
    levels = []
    for lv in range(batch.n_levels):
        batch.set_level(lv)
        for mode, fn in (("gn", batch.gn_step), ("grad", batch.adam_step)):
            total, kern = timed_steps(ctx, lambda ev, fn=fn: fn(ev), steps, warm)
            bytes_ = batch.algorithmic_bytes_per_iter(gn=(mode == "gn"))
            Hl, Wl = batch._keep[0][1].shape[:2]
            levels.append({"level": lv, "target": f"{Wl}x{Hl}", "iteration": mode,
                           "value": n_units * steps / (total * 1e-3), "ms_per_step": total / steps, "kernel_ms": kern,
                           "algorithmic_bytes_per_launch": int(bytes_), "frac": bytes_ / (kern * 1e-3) / 1e9 / peak})
    if ctx.rank == 0:
        fin = [l for l in levels if l["level"] == batch.n_levels - 1 and l["iteration"] == "gn"][0]
        line = line_base(ctx, "GN-iters/sec per pyramid level (640x480, 64 primitives)", "GN-iters/s", fin["value"], steps,
                         warm, fin["ms_per_step"] * steps, "C2 per level: blob segments (8x8 grid cells dilated 4 px), "
                         f"P={batch.points_total // batch.n} points/pair, geometry at full resolution", "weak",
                         {"pairs_per_gpu": a.pairs, "l2": "inputs larger than L2 at the finest level only"})
        line["levels"] = levels
        line["roofline"] = {"bound": "hbm", "achieved": fin["frac"] * peak, "peak": peak, "unit": "GB/s", "frac": fin["frac"],
                            "traffic": None, "peak_source": src, "kernel": "k_align_global<GN>", "kernel_ms": fin["kernel_ms"]}
        print(json.dumps(line), flush=True)


def run_c3(ctx):
    """config 3, tracking: every unit is one (keyframe, new frame) problem at the TUM shape; 300 Adam iterations on the
    pose increment + the frame's brightness terms, seeds held (lr_k = 0) -- odometery/odometery.py:300-312,365-403."""
    from super_primitive_b200.solver import AlignmentBatch
    from super_primitive_b200.shard import gather_results, shard_indices
    a = ctx.args
    H, W, N = 224, 288, a.segments or 100
    per_gpu = a.units or 256
    n_units = per_gpu * ctx.world
    iters_per_frame = 300
    kw = dict(lr_pose=5e-3, lr_k=0.0, lr_aff=5e-3)

    def make(units):
        return AlignmentBatch(build_pair_units(units, H, W, N, ctx.device, n_geoms=8, with_affine=True), with_affine=True)

    mine = shard_indices(n_units, ctx.rank, ctx.world)
    batch = make(mine)
    steps, warm = max(1, a.steps), max(3, a.warmup)
    total, kern = timed_steps(ctx, lambda ev: batch.adam_step(ev, **kw), steps, warm)
    # one tracked frame = 300 iterations replayed from a CUDA graph (no host work inside)
    graph = batch.capture_adam(iters_per_frame, **kw)
    ctx.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    graph.replay()
    f1.record()
    ctx.barrier()
    frame_ms = ctx.max_ms(f0.elapsed_time(f1))
    res = gather_results(batch.poses_matrix(), batch.k_padded(), batch.grad_costs(), n_units)

    def rerun(units):
        b = make(units)
        for _ in range(warm + steps):
            b.adam_step(**kw)
        g = b.capture_adam(iters_per_frame, **kw)
        g.replay()
        torch.cuda.synchronize()
        return b.poses_matrix(), b.k_padded(), b.grad_costs()

    chk = shard_check(ctx, n_units, rerun, res, "poses / seeds / costs")
    if ctx.rank == 0:
        peak, src = hbm_peak()
        bytes_ = batch.algorithmic_bytes_per_iter(gn=False)
        P = batch.points_total // batch.n
        line = line_base(ctx, "tracking iterations/sec (288x224, pose + brightness)", UNIT_ITERS,
                         n_units * steps / (total * 1e-3), steps, warm, total,
                         f"C3 tracking: {N} blob segments, P={P} points/frame, finest level, {per_gpu} independent "
                         "(keyframe, frame) problems per GPU", "weak",
                         {"units_per_gpu": per_gpu, "iteration": "Adam on pose increment + brightness, seeds held "
                          "(the reference's tracker)", "working_set_bytes_per_gpu": int(per_gpu * (20 * P + 16 * H * W))})
        line["roofline"] = {"bound": "hbm", "achieved": bytes_ / (kern * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": bytes_ / (kern * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": src,
                            "kernel": "k_align_global<GRAD,affine>", "kernel_ms": kern,
                            "algorithmic_bytes_per_launch": int(bytes_)}
        line["tracked_frames"] = {"iterations_per_frame": iters_per_frame, "ms_per_frame_batch": frame_ms,
                                  "frames_per_s": n_units / (frame_ms * 1e-3),
                                  "what": "300 iterations (config/tum/odom_desk.yaml steps [0,0,300]) replayed from one "
                                          "CUDA graph for every unit of the batch"}
        line["shard_check"] = chk
        line["gpu_launches"] = 2 * steps
        print(json.dumps(line), flush=True)


def run_c5(ctx):
    """config 5, stress: 1024x768, 256 segments per pair; `--units` pairs in total (default 1024), GN/LM iteration."""
    from super_primitive_b200.solver import AlignmentBatch
    from super_primitive_b200.shard import gather_results, shard_indices
    a = ctx.args
    H, W, N = 768, 1024, a.segments or 256
    n_units = a.units or 1024

    def make(units):
        return AlignmentBatch(build_pair_units(units, H, W, N, ctx.device, n_geoms=2))

    mine = shard_indices(n_units, ctx.rank, ctx.world)
    t0 = time.perf_counter()
    batch = make(mine)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    steps, warm = max(1, a.steps), max(3, a.warmup)
    total, kern = timed_steps(ctx, lambda ev: batch.gn_step(ev), steps, warm)
    costs = batch.lm_state[:, 1] / (3.0 * batch.pts_per_problem)
    res = gather_results(batch.poses_matrix(), batch.k_padded(), costs, n_units)

    def rerun(units):
        b = make(units)
        for _ in range(warm + steps):
            b.gn_step()
        torch.cuda.synchronize()
        return b.poses_matrix(), b.k_padded(), b.lm_state[:, 1] / (3.0 * b.pts_per_problem)

    chk = shard_check(ctx, n_units, rerun, res, "poses / seeds / costs")
    if ctx.rank == 0:
        peak, src = hbm_peak()
        bytes_ = batch.algorithmic_bytes_per_iter(gn=True)
        P = batch.points_total // batch.n
        line = line_base(ctx, "GN-iters/sec (1024x768, 256 primitives, stress batch)", "GN-iters/s",
                         n_units * steps / (total * 1e-3), steps, warm, total,
                         f"C5 stress: {n_units} pairs in total, 1024x768, {N} blob segments, P={P} points/pair, finest level",
                         "strong", {"units_total": n_units, "units_this_rank": len(mine),
                                    "working_set_bytes_per_gpu": int(len(mine) * (20 * P + 16 * H * W)),
                                    "geometry": "2 distinct source geometries per rank shared by its pairs (images per pair)",
                                    "build_s": build_s, "l2": "inputs larger than L2 (126 MB), no flush"})
        line["roofline"] = {"bound": "hbm", "achieved": bytes_ / (kern * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": bytes_ / (kern * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": src,
                            "kernel": "k_align_global<GN>", "kernel_ms": kern, "algorithmic_bytes_per_launch": int(bytes_)}
        line["shard_check"] = chk
        line["gpu_launches"] = 2 * steps
        print(json.dumps(line), flush=True)


def sparse_depth(kf_depth, keep, seed):
    """VOID-style sparse depth: `keep` random valid pixels of a dense depth map, 0 elsewhere."""
    H, W = kf_depth.shape
    g = torch.Generator(device=kf_depth.device).manual_seed(seed)
    idx = torch.randperm(H * W, generator=g, device=kf_depth.device)[:keep]
    out = torch.zeros(H * W, dtype=torch.float32, device=kf_depth.device)
    out[idx] = kf_depth.reshape(-1)[idx]
    return out.reshape(H, W)


def run_c4(ctx):
    """config 4, VOID depth completion: `--units` frames in total (default 256), per frame ordered mask compaction of the
    (100,480,640) dense keyframe, per-segment lower-median re-initialisation from ~1500 sparse depths and the average
    render of the seeded segments -- depth_completion/segment_based_completion.py:30-62 through the drop-in surface."""
    from super_primitive_b200 import depth_completion as dc, depth_init, geometry
    from super_primitive_b200.shard import shard_indices
    a = ctx.args
    H, W, N = 480, 640, a.segments or 100
    n_units = a.units or 256
    mine = shard_indices(n_units, ctx.rank, ctx.world)
    frames = []
    for u in mine:
        kf = device_keyframe(H, W, N, ctx.device, seed=u, pad=3, noise=0.0)
        true_depth = 1.5 + 0.5 * torch.sin(torch.linspace(0, 3.0, W, device=ctx.device))[None, :].expand(H, W).contiguous()
        frames.append((kf, sparse_depth(true_depth, 1500, u)))
    torch.cuda.synchronize()
    sums = torch.zeros((len(mine), 2), dtype=torch.float64, device=ctx.device)

    def step(_ev):
        for i, (kf, sp) in enumerate(frames):
            geometry.clear_caches()                       # a new frame: its masks are compacted, never reused
            k, vis = depth_init.segment_based_depth_reinit(sp.clone(), kf, 'median', return_info=True)
            depth, invalid = dc.render_segments_avg(kf, k, vis)
            sums[i, 0] = depth.sum(dtype=torch.float64)
            sums[i, 1] = invalid.sum()
        torch.set_grad_enabled(True)

    steps, warm = max(1, a.steps), max(3, a.warmup)
    total_pf, _ = timed_steps(ctx, step, steps, warm, kernel_events=False)

    # the same frames through the batched entry point: two host syncs per 32 frames instead of two per frame
    sums_b = torch.zeros_like(sums)
    CH = 32

    def step_batched(_ev):
        geometry.clear_caches()
        for c0 in range(0, len(frames), CH):
            chunk = frames[c0:c0 + CH]
            res = dc.complete_batch([kf for kf, _ in chunk], [sp.clone() for _, sp in chunk], 'median')
            for j, (depth, invalid, _k, _vis) in enumerate(res):
                sums_b[c0 + j, 0] = depth.sum(dtype=torch.float64)
                sums_b[c0 + j, 1] = invalid.sum()
            geometry.clear_caches()

    total_b, _ = timed_steps(ctx, step_batched, steps, warm, kernel_events=False)
    batched_equal = bool(torch.equal(sums, sums_b))
    # stage split (untimed pass, device events per stage on one frame)
    kf, sp = frames[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    geometry.clear_caches()
    ev[0].record()
    geom = geometry.geometry_of(kf)
    ev[1].record()
    k, vis = depth_init.segment_based_depth_reinit(sp.clone(), kf, 'median', return_info=True)
    ev[2].record()
    dc.render_segments_avg(kf, k, vis)
    ev[3].record()
    torch.cuda.synchronize()
    torch.set_grad_enabled(True)
    stage = {"compaction_ms": ev[0].elapsed_time(ev[1]), "reinit_ms": ev[1].elapsed_time(ev[2]),
             "render_ms": ev[2].elapsed_time(ev[3])}
    # nearest-valid hole filling of the completed maps (fill_in_tools.fill_depth, the exact second stage of the
    # reference's evaluation-side fill): one frame, and a chunk of frames in one call
    fill = None
    if frames:
        from super_primitive_b200 import fill_in_tools as fit
        res = dc.complete_batch([kf for kf, _ in frames[:CH]], [sp.clone() for _, sp in frames[:CH]], 'median')
        d_stack = torch.stack([r[0] for r in res])
        i_stack = torch.stack([r[1] for r in res])
        # the synthetic segments cover the whole frame; holes like a real completion's are cut in: an unreached border,
        # 24 rectangles of up to 1/8 of each side, 2 % isolated pixels (seeded per frame)
        for f in range(i_stack.shape[0]):
            gen = torch.Generator().manual_seed(1000 + f)
            hole = torch.rand((H, W), generator=gen) < 0.02
            hole[:, :9] = True
            for _ in range(24):
                r0, c0 = int(torch.randint(0, H, (1,), generator=gen)), int(torch.randint(0, W, (1,), generator=gen))
                hole[r0:r0 + int(torch.randint(1, H // 8, (1,), generator=gen)), c0:c0 + int(torch.randint(1, W // 8, (1,), generator=gen))] = True
            i_stack[f] |= hole.to(ctx.device)
        fill = {"holes_frac": float(i_stack.float().mean().item())}
        for name, (dd, ii) in (("one_frame_ms", (d_stack[:1], i_stack[:1])), ("chunk_ms", (d_stack, i_stack))):
            for _ in range(3):
                fit.fill_depth_batch(dd, ii)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fit.fill_depth_batch(dd, ii)
            e1.record()
            torch.cuda.synchronize()
            fill[name] = e0.elapsed_time(e1) / 10
        fill["frames_in_chunk"] = int(d_stack.shape[0])
        fill["frames_per_s"] = d_stack.shape[0] / (fill["chunk_ms"] * 1e-3)
        fill["what"] = ("fill_in_tools.fill_depth_batch: exact Euclidean nearest-valid fill (scipy's tie-breaking), device "
                        "events around the call, completed maps resident; NOT part of `value`")
        if ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline:
            try:        # the reference's own two lines (depth_completion/fill_in_tools.py:5-7) on the host
                from scipy import ndimage as nd
                d_np, i_np = d_stack[0].cpu().numpy(), i_stack[0].cpu().numpy()
                nd.distance_transform_edt(i_np, return_distances=False, return_indices=True)
                t0 = time.perf_counter()
                for _ in range(5):
                    ind = nd.distance_transform_edt(i_np, return_distances=False, return_indices=True)
                    ref_filled = d_np[tuple(ind)]
                fill["cpu_scipy_ms_one_frame"] = (time.perf_counter() - t0) / 5 * 1e3
                fill["equals_scipy"] = bool(np.array_equal(fit.fill_depth_batch(d_stack[:1], i_stack[:1])[0].cpu().numpy(), ref_filled))
            except ImportError:
                fill["cpu_scipy_ms_one_frame"] = None
        del res, d_stack, i_stack
    # gather of the per-frame checksums (the completed maps stay where they were computed)
    if ctx.world > 1:
        pad = torch.full(((n_units + ctx.world - 1) // ctx.world, 2), float('nan'), dtype=torch.float64, device=ctx.device)
        pad[:len(mine)] = sums
        allv = [torch.empty_like(pad) for _ in range(ctx.world)]
        ctx.dist.all_gather(allv, pad)
        gathered = torch.empty((n_units, 2), dtype=torch.float64, device=ctx.device)
        for r in range(ctx.world):
            idx = shard_indices(n_units, r, ctx.world)
            gathered[torch.tensor(idx, device=ctx.device)] = allv[r][:len(idx)]
    else:
        gathered = sums

    def rerun(units):
        out = torch.zeros((len(units), 2), dtype=torch.float64, device=ctx.device)
        for i, u in enumerate(units):
            kf = device_keyframe(H, W, N, ctx.device, seed=u, pad=3, noise=0.0)
            td = 1.5 + 0.5 * torch.sin(torch.linspace(0, 3.0, W, device=ctx.device))[None, :].expand(H, W).contiguous()
            geometry.clear_caches()
            k, vis = depth_init.segment_based_depth_reinit(sparse_depth(td, 1500, u), kf, 'median', return_info=True)
            depth, invalid = dc.render_segments_avg(kf, k, vis)
            out[i, 0] = depth.sum(dtype=torch.float64)
            out[i, 1] = invalid.sum()
        torch.set_grad_enabled(True)
        return (out,)

    chk = shard_check(ctx, n_units, rerun, (gathered,), "per-frame checksums of the completed depth map")
    cpu_baseline = None
    if ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline:
        # the reference's own per-frame path on the host cores (oracle/ref_port.py restates it operation for operation:
        # odometery/depth_init.py:10-67 + depth_completion/segment_based_completion.py:48-54,21-27), a bounded sample
        from oracle import ref_port as port
        kf_c, sp_c = frames[0][0].to("cpu"), frames[0][1].cpu()
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        with torch.no_grad():
            port.completion_render(kf_c, *port.segment_median_reinit(sp_c, kf_c, 'median'))       # warm-up
            t0 = time.perf_counter()
            n_cpu = 3
            for _ in range(n_cpu):
                kk, vis = port.segment_median_reinit(sp_c, kf_c, 'median')
                port.completion_render(kf_c, kk, vis)
            dt = (time.perf_counter() - t0) / n_cpu
        cpu_baseline = {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n_cpu} frames (after 1 warm-up) of per-segment median re-initialisation + dense depth "
                                  f"expansion + average render on the host, {dt * 1e3:.0f} ms per frame"}
    if ctx.rank == 0:
        peak, src = hbm_peak()
        P = geom.P
        bytes_frame = 5 * N * H * W + 8 * P + 8 * P + 8 * H * W
        frames_s = n_units * steps / (total_b * 1e-3)              # the batch entry point is what this batch config calls
        total = total_b
        line = line_base(ctx, "depth-completion frames/sec (640x480, 100 segments per frame)", "frames/s", frames_s, steps,
                         warm, total, f"C4 VOID depth completion: {n_units} frames in total, {N} blob segments, P={P} "
                         "mask pixels/frame, 1500 sparse depths/frame", "strong",
                         {"units_total": n_units, "units_this_rank": len(mine), "stages_ms_one_frame": stage,
                          "per_frame": "mask compaction (5 N H W bytes read) + per-segment lower-median re-initialisation "
                                       "+ average render (8 P + 8 H W), through depth_init / depth_completion",
                          "working_set_bytes_per_gpu": int(len(mine) * 5 * N * H * W)})
        ach = bytes_frame * frames_s / ctx.world / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": None, "peak_source": src, "kernel": "k_row_count + k_row_fill (mask compaction)",
                            "algorithmic_bytes_per_frame": int(bytes_frame),
                            "note": "whole per-frame pipeline time incl. its two host syncs per frame (point count, visible count)"}
        line["batched"] = {"frames_per_call": CH, "bit_equal_to_per_frame": batched_equal,
                           "what": "`value`: depth_completion.complete_batch -- all compactions of a chunk queued, ONE "
                                   "read-back of their point counts, then re-initialisation + render of every frame, one "
                                   "check at the end"}
        line["per_frame_calls"] = {"value": n_units * steps / (total_pf * 1e-3), "unit": "frames/s",
                                   "what": "segment_based_depth_reinit + render_segments_avg frame by frame (the reference's "
                                           "call pattern: two host syncs per frame)"}
        line["shard_check"] = chk
        line["hole_fill"] = fill
        line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)


def run_compaction(ctx):
    """One-time cost per keyframe of the ordered compaction ('report separately'): config 2 and config 5."""
    from super_primitive_b200.geometry import CompactGeometry
    out = []
    for name, (H, W, N, pad) in (("C2", (480, 640, 64, 4)), ("C5", (768, 1024, 256, 2))):
        kf = device_keyframe(H, W, N, ctx.device, seed=1, pad=pad)
        ms = []
        for i in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g = CompactGeometry(kf.keypoint_regions, kf.logdepth_perseg, kf.keypoints, kf.K)
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        img = kf.image
        e0.record()
        g._levels.clear()
        g.level_buffers(img)
        e1.record()
        torch.cuda.synchronize()
        nbytes = 5 * N * H * W
        best = min(ms[1:])
        out.append({"config": name, "H": H, "W": W, "N": N, "P": g.P, "dense_bytes_read": nbytes, "ms_per_keyframe": best,
                    "GBps": nbytes / (best * 1e-3) / 1e9, "level_buffers_ms": e0.elapsed_time(e1),
                    "what": "CompactGeometry(masks, log-depth, keypoints): count + scan + fill + tile table, wall clock "
                            "incl. its one host sync; level_buffers = cached source samples + tile-major stream of one level"})
        # the same keyframe straight from the frontend's hand-over (integrated depth at twice the keyframe resolution,
        # frontend/process_frame.py:231-236): no dense (N,H,W) mask / log-depth tensor is ever built
        if name == "C2":
            from super_primitive_b200.handover import geometry_from_frontend
            dense_depth = torch.exp(kf.logdepth_perseg) * kf.keypoint_regions
            depth_f = dense_depth.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()      # (N, 2H, 2W)
            ms_h = []
            for i in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gh, _, _ = geometry_from_frontend(depth_f, kf.keypoints, kf.K, (H, W))
                torch.cuda.synchronize()
                ms_h.append((time.perf_counter() - t0) * 1e3)
            # what the reference does with the same input before the alignment path can start
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ld = torch.nn.functional.interpolate(depth_f[:, None], size=(H, W), mode='nearest')[:, 0]
            mk = ld > 1e-7
            ld[mk] = torch.log(ld[mk])
            torch.cuda.synchronize()
            out[-1]["handover"] = {"ms_per_keyframe": min(ms_h[1:]), "points_equal_dense_route": bool(gh.P == g.P),
                                   "frontend_bytes_read": int(depth_f.numel() * 4),
                                   "reference_dense_ops_ms": (time.perf_counter() - t0) * 1e3,
                                   "what": "handover.geometry_from_frontend: compact geometry straight from integrated_depth "
                                           "(N,2H,2W) incl. keypoint snap; reference_dense_ops_ms = interpolate + threshold + "
                                           "log on the device with torch (without put_keypoints_back's Python loop)"}
            del dense_depth, depth_f, gh, ld, mk
        del kf, g
    if ctx.rank == 0:
        peak, src = hbm_peak()
        line = {"metric": "one-time compaction per keyframe", "unit": "ms", "value": out[0]["ms_per_keyframe"],
                "n_gpus": ctx.world, "higher_is_better": False, "data": "synthetic", "dtype": "u8/f32",
                "config": {"workload": "ordered mask compaction, configs 2 and 5"}, "compaction": out,
                "roofline": {"bound": "hbm", "achieved": out[1]["GBps"], "peak": peak, "unit": "GB/s",
                             "frac": out[1]["GBps"] / peak, "peak_source": src, "traffic": None}}
        print(json.dumps(line), flush=True)


RUNNERS = {"c2levels": run_c2levels, "c3": run_c3, "c4": run_c4, "c5": run_c5, "compaction": run_compaction}
