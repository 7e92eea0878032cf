#!/usr/bin/env python
"""Benchmark of the dense photometric alignment hot path (BASELINE.json metric: GN-iters/sec,
640x480, 64 primitives, two-frame alignment; HBM GB/s vs roofline for the fused kernel).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores)

A *step* is one Gauss-Newton/LM iteration over one batch of independent two-frame problems
(`--pairs` per GPU, default 64 so the working set is ~0.7 GB >> 126 MB L2): one fused
residual+Jacobian+normal-equation kernel, a fixed-order finalize and the per-problem damped solve.
`value` = problem-iterations per second over all ranks (weak scaling: fixed pairs per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = "GN-iters/sec (640x480, 64 primitives, two-frame alignment)"
UNIT = "GN-iters/s"
WORKLOAD = dict(H=480, W=640, N=64, kind="overlap")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="independent two-frame problems per GPU")
    ap.add_argument("--mode", default="gn", choices=["gn", "grad"],
                    help="gn = IRLS GN/LM iteration; grad = first-order iteration (cost + gradient + Adam update)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=8, help="pairs per ingest launch group of the end-to-end arm")
    ap.add_argument("--workload", default="c2", choices=["c2", "c2levels", "c3", "c4", "c5", "compaction"],
                    help="c2 = the headline (BASELINE config 2, finest level); the others: bench_workloads.py")
    ap.add_argument("--units", type=int, default=0, help="c3: problems per GPU; c4 / c5: units in total (0 = default)")
    ap.add_argument("--segments", type=int, default=0, help="segments per keyframe of c3 / c4 / c5 (0 = the config's)")
    ap.add_argument("--no-numa-pin", action="store_true", help="do not bind the rank to the CPUs next to its GPU")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# host placement: a rank's pinned staging memory should live on the NUMA node its GPU hangs off
# --------------------------------------------------------------------------------------------------
def pin_rank_to_gpu_numa(local_rank, world):
    """Bind this process to the CPUs local to its GPU's PCIe root (sysfs `local_cpulist`), split between the ranks that
    share them, BEFORE the pinned arenas are allocated (first touch then places them on that node).  Round 1's 8-GPU
    end-to-end arm ran every rank on NUMA node 0 and its per-GPU H2D rate fell from 54 to 21 GB/s.  Returns a dict for
    the JSON line (what was found / done); never raises."""
    info = {"pinned": False}
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = []
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        info.update(pci=bus, numa_node=node, local_cpus=len(cpus), allowed_local_cpus=len(allowed))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            info["pinned"] = True
    except Exception as e:      # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"
    return info


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# workload construction
# --------------------------------------------------------------------------------------------------
def build_batch(pairs, device, seed0=0, level=0):
    """`pairs` distinct problems of the C2 shape at the finest pyramid level, every one with its own
    HBM-resident geometry and images (nothing shared, so the working set scales with `pairs`)."""
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.geometry import CompactGeometry, pack_rgba
    from super_primitive_b200.solver import AlignmentBatch
    H, W, N = WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["N"]
    src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind=WORKLOAD["kind"], seed=seed0, noise=0.01)
    src, trg = src.to(device), trg.to(device)
    from super_primitive_b200.frames import image_tt
    g = torch.Generator(device="cpu").manual_seed(seed0)
    problems = []

    def to_u8(img):      # (3,H,W) float in [0,1] -> HWC uint8, what a dataset reader (cv2) hands over
        return (img.clamp(0, 1) * 255.0).round().to(torch.uint8).permute(1, 2, 0).contiguous()

    for i in range(pairs):
        geom = CompactGeometry(src.keypoint_regions, src.get_logdepth(), src.keypoints, src.K)
        noise = (torch.rand(trg.image.shape, generator=g) * 0.02 - 0.01).to(device)
        # frames are 8-bit images; the float frames every arm works on are image_tt of them (tool/etc.py:37-40),
        # computed on the device exactly as the end-to-end arm recomputes them every step
        t_u8, s_u8 = to_u8(trg.image + noise), to_u8(src.image + noise.flip(-1))
        timg, simg = image_tt(t_u8, device), image_tt(s_u8, device)
        src_rgb, pack = geom.level_buffers(simg)
        trg_rgba = pack_rgba(timg)[0].clone()
        dk = (torch.rand(N, generator=g) * 0.1 - 0.05)
        pose = pose0.clone()
        pose[:3, 3] += (torch.rand(3, generator=g) - 0.5) * 0.01
        problems.append(dict(geom=geom, src_rgb=src_rgb, pack=pack, trg_rgba=trg_rgba, K_trg=trg.K, pose=pose.to(device),
                             k=(k0 + dk).to(device), src_image=simg, trg_image=timg, src_u8=s_u8, trg_u8=t_u8))
    batch = AlignmentBatch(problems, with_affine=False, irls_eps=1e-3)
    return batch, problems


class HostStaged:
    """End-to-end arm.  Every step the step's inputs travel from pinned host memory and the results travel back:

    mode "u8" (the headline): the 8-bit source and target FRAMES of every pair (HWC uint8, what the reference's
    dataset readers hand over) + pose + log-depth seeds go up; on the device `spb_ingest_u8` runs the reference's
    `image_tt` conversion and re-derives what the fused kernel streams (RGBA target, cached source samples,
    tile-major level buffer over the resident compact geometry -- keyframe state the frontend produces on the
    device) in three launches per chunk of pairs on the compute stream while later chunks are still in flight on
    a copy stream; then one GN/LM iteration; poses, seeds and LM state come back.
    mode "raw": the same with the frames already converted to float32 on the host (12 bytes per pixel over PCIe,
    what the reference's `image_tt` uploads), re-derived pair by pair.
    mode "packed": the already derived buffers (tile-major level buffer + RGBA target) are uploaded instead.
    mode "params": frames stay resident, only pose + seeds travel."""

    def __init__(self, batch, problems, chunk=8):
        from super_primitive_b200 import _native as nat
        self.nat = nat
        self.batch = batch
        self.problems = problems
        self.raw_dev, self.raw_host, self.packed_dev, self.packed_host = [], [], [], []
        for p in problems:
            self.raw_dev.append((p['src_image'], p['trg_image']))
            self.raw_host.append((p['src_image'].cpu().pin_memory(), p['trg_image'].cpu().pin_memory()))
            self.packed_dev.append((p['pack'], p['trg_rgba']))
            self.packed_host.append((p['pack'].cpu().pin_memory(), p['trg_rgba'].cpu().pin_memory()))
        from super_primitive_b200.frames import FrameIngest
        # SPB_E2E_LEAN=1 (tuning visits with a fused-ingest library only): skip the buffers the iteration does not read
        self.ingest = FrameIngest(problems, batch.geoms, lean=bool(os.environ.get("SPB_E2E_LEAN")))
        self.arena = self.ingest.host_arena()          # pinned arena a loader decodes the 8-bit frames into
        # the odometry case: the source keyframe is resident, only the new (target) frame of every pair arrives
        self.ingest_t = FrameIngest(problems, batch.geoms, target_only=True)
        self.arena_t = self.ingest_t.host_arena()
        for i, p in enumerate(problems):
            self.ingest.fill(self.arena, i, p['src_u8'].cpu(), p['trg_u8'].cpu())
            self.ingest_t.fill(self.arena_t, i, None, p['trg_u8'].cpu())
        self.chunk = max(1, int(chunk))                # pairs per ingest launch group (copy/compute overlap)
        self.chunk_events = [torch.cuda.Event() for _ in range((len(problems) + self.chunk - 1) // self.chunk)]
        self.consumed = []
        self.h_pose = batch.poses.cpu().pin_memory()
        self.h_k = batch.k.cpu().pin_memory()
        self.o_pose = torch.empty_like(self.h_pose).pin_memory()
        self.o_k = torch.empty_like(self.h_k).pin_memory()
        self.o_cost = torch.empty((batch.n, 8), dtype=torch.float32).pin_memory()
        nbytes = lambda pairs: sum(t.numel() * t.element_size() for pr in pairs for t in pr)   # noqa: E731
        self.params_bytes = self.h_pose.numel() * 4 + self.h_k.numel() * 4
        self.h2d = {"u8": self.ingest.offsets[-1] + self.params_bytes,      # bytes actually copied (16-byte aligned frames)
                    "u8t": self.ingest_t.offsets[-1] + self.params_bytes,
                    "raw": nbytes(self.raw_host) + self.params_bytes,
                    "packed": nbytes(self.packed_host) + self.params_bytes, "params": self.params_bytes}
        self.d2h = (self.o_pose.numel() + self.o_k.numel() + self.o_cost.numel()) * 4
        self.copy_stream = torch.cuda.Stream()
        self.events = [torch.cuda.Event() for _ in problems]
        # pose + seeds travel FIRST on the copy stream into their own double-buffered staging and are moved into place on
        # the compute stream: issued on the compute stream they would queue in the H2D engine behind the NEXT step's
        # frame copies (the host runs ahead) and hold the iteration back by a whole step (round 1: 0.7-0.8 of the link)
        self.d_pose_in = [torch.empty_like(batch.poses) for _ in range(2)]
        self.d_k_in = [torch.empty_like(batch.k) for _ in range(2)]
        self.params_ev = [torch.cuda.Event() for _ in range(2)]
        self.params_done = [None, None]            # compute stream has moved staging [parity] into place
        self.parity = 0
        self.launches_per_step = {"u8": 3 * len(self.chunk_events) + 2, "u8t": 3 * len(self.chunk_events) + 2,
                                  "raw": 3 * len(problems) + 2, "packed": 2, "params": 2}

    def step(self, mode="u8"):
        b = self.batch
        main = torch.cuda.current_stream()
        if mode in ("u8", "u8t"):
            cs, n = self.copy_stream, len(self.problems)
            ing, arena = (self.ingest, self.arena) if mode == "u8" else (self.ingest_t, self.arena_t)
            if len(self.consumed) != len(self.chunk_events):
                self.consumed = [None] * len(self.chunk_events)
            par = self.parity
            self.parity ^= 1
            with torch.cuda.stream(cs):
                if self.params_done[par] is not None:
                    cs.wait_event(self.params_done[par])
                self.d_pose_in[par].copy_(self.h_pose, non_blocking=True)
                self.d_k_in[par].copy_(self.h_k, non_blocking=True)
                self.params_ev[par].record(cs)
                for c, ev in enumerate(self.chunk_events):
                    first = c * self.chunk
                    if self.consumed[c] is not None:
                        cs.wait_event(self.consumed[c])     # the previous step's ingest has read this part of the staging buffer
                    ing.upload(arena, first, min(n, first + self.chunk) - first)      # one copy per chunk
                    ev.record(cs)
            for c, ev in enumerate(self.chunk_events):
                main.wait_event(ev)
                first = c * self.chunk
                ing.run(b.d_geoms, first, min(n, first + self.chunk) - first)
                if self.consumed[c] is None:
                    self.consumed[c] = torch.cuda.Event()
                self.consumed[c].record(main)
            # (the copies of the NEXT step may start as soon as its chunk has been ingested: they overlap this step's
            #  iteration; the derived buffers are written and read on the compute stream only, so they need no second copy)
        elif mode == "raw":
            lib, nat = self.nat.lib(), self.nat
            cs = self.copy_stream
            cs.wait_stream(main)                      # the previous step's consumers of the frame buffers are done
            with torch.cuda.stream(cs):
                for (ds, dt), (hs_, ht), ev in zip(self.raw_dev, self.raw_host, self.events):
                    ds.copy_(hs_, non_blocking=True)
                    dt.copy_(ht, non_blocking=True)
                    ev.record(cs)
            st = main.cuda_stream
            for p, (ds, dt), ev in zip(self.problems, self.raw_dev, self.events):
                main.wait_event(ev)
                g = p['geom']
                nat.check(lib.spb_pack_rgba(dt.data_ptr(), dt.stride(0), 1, dt.shape[1], dt.shape[2],
                                            p['trg_rgba'].data_ptr(), st), "spb_pack_rgba")
                nat.check(lib.spb_sample_source(g.cref, ds.data_ptr(), ds.shape[1], ds.shape[2],
                                                p['src_rgb'].data_ptr(), st), "spb_sample_source")
                nat.check(lib.spb_build_tile_pack(g.cref, p['src_rgb'].data_ptr(), p['pack'].data_ptr(), st),
                          "spb_build_tile_pack")
        elif mode == "packed":
            for devs, hosts in zip(self.packed_dev, self.packed_host):
                for d, h in zip(devs, hosts):
                    d.copy_(h, non_blocking=True)
        if mode in ("u8", "u8t"):
            main.wait_event(self.params_ev[par])
            b.poses.copy_(self.d_pose_in[par], non_blocking=True)
            b.k.copy_(self.d_k_in[par], non_blocking=True)
            if self.params_done[par] is None:
                self.params_done[par] = torch.cuda.Event()
            self.params_done[par].record(main)
        else:
            b.poses.copy_(self.h_pose, non_blocking=True)
            b.k.copy_(self.h_k, non_blocking=True)
        b.gn_step()
        self.o_pose.copy_(b.poses, non_blocking=True)
        self.o_k.copy_(b.k, non_blocking=True)
        self.o_cost.copy_(b.lm_state, non_blocking=True)


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch path (live import when mounted, else the pinned port)
# --------------------------------------------------------------------------------------------------
def cpu_reference_iter_fn():
    """Returns (kind, callable running ONE reference iteration: photomeric_cost forward + backward +
    Adam.step, C2 shape at the finest level)."""
    from super_primitive_b200 import synthetic as syn
    H, W, N = WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["N"]
    src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind=WORKLOAD["kind"], seed=0, noise=0.01)
    cfg = {'mode': 'colour', 'collect_stats': 0}
    kind = "port"
    cost = None
    if os.path.isdir("/root/reference/core") and not os.environ.get("SPB_FORCE_PORT"):
        try:
            sys.dont_write_bytecode = True
            sys.path.insert(0, "/root/reference")
            import core.dense_optim as ref_do          # the unmodified reference
            from image.keyframe import KeyFrame as RefKF
            rs = RefKF(src.image, src.K, src.logdepth_perseg, src.keypoints, src.keypoint_regions)
            rt = RefKF(trg.image, trg.K)
            cost = lambda k, pose: ref_do.photomeric_cost(rs, rt, k, pose, cfg)   # noqa: E731
            kind = "reference"
        except Exception:
            cost = None
    if cost is None:
        from oracle import ref_port as port
        cost = lambda k, pose: port.cost_single(src, trg, k, pose, cfg)          # noqa: E731
    torch.set_grad_enabled(True)
    k = torch.nn.Parameter(k0.clone())
    pose = torch.nn.Parameter(pose0.clone())
    # reference learning rates, odometery/two_frame_sfm.py:117-121
    opt = torch.optim.Adam([{'params': [k], 'lr': 1e-3}, {'params': [pose], 'lr': 1e-2}], lr=1e-3)

    def one_iter():
        loss = cost(k, pose)['residual'].mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return kind, one_iter


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu(steps, warmup):
    """Times the reference iteration with the thread count that is FASTEST on this host (the reference gets
    its best shot): candidates 8, 16, 32, 64, ... up to all host cores; escalation stops once a candidate is
    clearly slower than the best so far.  Returns (kind, it/s, ms/iter, threads used)."""
    kind, one_iter = cpu_reference_iter_fn()
    ncpu = host_cores()
    cands = sorted({c for c in (8, 16, 32, 64, 128, ncpu) if c <= ncpu} | {min(ncpu, 8)})
    best_t, best_ms = None, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        one_iter()
        t0 = time.perf_counter()
        one_iter()
        one_iter()
        ms = (time.perf_counter() - t0) / 2 * 1e3
        if ms < best_ms:
            best_t, best_ms = c, ms
        elif ms > 1.5 * best_ms:
            break
    torch.set_num_threads(best_t)
    for _ in range(warmup):
        one_iter()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_iter()
    dt = time.perf_counter() - t0
    return kind, steps / dt, dt / steps * 1e3, best_t


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = max(3, args.warmup) if args.warmup else 3
    kind, its, ms, cores = time_cpu(steps, warm)
    sample = (f"{steps} timed iterations (after {warm} warm-up) of photomeric_cost fwd + backward + Adam.step on one "
              f"640x480 / 64-segment pair at the finest level, float32, {cores} torch threads (fastest of the "
              f"candidate thread counts on {host_cores()} host cores)")
    line = {"metric": METRIC, "value": its, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 two-frame SfM 640x480, 64 segments, finest pyramid level, 1 pair per step",
                       "iteration": "reference PyTorch-CPU: forward + backward + Adam.step"},
            "cpu_baseline": {"value": its, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_dropin_single_pair(device, iters=60, warm=10):
    """photomeric_cost (collect_stats=0) -> residual.mean().backward() -> Adam.step on one C2 pair through
    super_primitive_b200.dense_optim, i.e. exactly what odometery/two_frame_sfm.py does per iteration."""
    from super_primitive_b200 import dense_optim as do, synthetic as syn
    H, W, N = WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["N"]
    src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind=WORKLOAD["kind"], seed=0, noise=0.01)
    src, trg = src.to(device), trg.to(device)
    k = torch.nn.Parameter(k0.to(device))
    pose = torch.nn.Parameter(pose0.to(device))
    opt = torch.optim.Adam([{'params': [k], 'lr': 1e-3}, {'params': [pose], 'lr': 1e-2}], lr=1e-3)
    cfg = {'mode': 'colour', 'collect_stats': 0}

    def one():
        loss = do.photomeric_cost(src, trg, k, pose, cfg)['residual'].mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(warm):
        one()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        one()
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return {"iters_per_s": iters / wall, "ms_per_iter_wall": wall / iters * 1e3,
            "ms_per_iter_device": a.elapsed_time(b) / iters,
            "what": "reference loop through the drop-in API on ONE pair: photomeric_cost (one fused forward+backward launch + "
                    "finalize, finiteness flags read every 16 calls) -> backward -> torch.optim.Adam.step; host-bound by "
                    "PyTorch itself (Adam.step alone ~150 us, the autograd engine ~100 us, this package's forward ~100 us: "
                    "profiles/r02c_dropin_profile.txt)"}


def time_device_loop_single_pair(device, iters=200):
    """ONE C2 pair (the shape of real-time use: nothing to batch over) through the device-resident loops: each
    iteration is two launches with no host synchronisation; eager and replayed from a CUDA graph."""
    batch, _ = build_batch(1, device, seed0=7)
    out = {}
    for name, run, cap in (("gn", batch.run_gn, batch.capture_gn), ("adam", batch.run_adam, batch.capture_adam)):
        run(20)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(iters)
        b.record()
        torch.cuda.synchronize()
        eager_ms = a.elapsed_time(b) / iters
        graph = cap(50)
        graph.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(4):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        graph_ms = a.elapsed_time(b) / 200
        out[name] = {"eager_us_per_iter": 1e3 * eager_ms, "graph_us_per_iter": 1e3 * graph_ms,
                     "iters_per_s": 1e3 / min(eager_ms, graph_ms)}
    out["what"] = ("device-resident GN/LM and Adam loops on ONE 640x480 / 64-segment pair (two launches per iteration, "
                   "no host sync): latency-bound, the real-time tracking shape")
    return out


def time_mapping_windows(device, n_windows=8, iters=60):
    """Windowed mapping (odometery/odometery.py:687-915) through the device-resident window iteration
    (spb_window_iterate): `n_windows` independent windows of 3 keyframes + 3 supporting frames at the TUM shape
    (288x224 after downsample_pow 1, 64 segments), 9 edges each; one iteration = batched gradient kernel over all
    edges + finalize + coupled Adam update / pose bookkeeping (three launches, no host sync)."""
    from super_primitive_b200 import synthetic as syn
    from super_primitive_b200.window import MappingWindows
    wins = []
    for i in range(n_windows):
        w = syn.mapping_window(224, 288, 64, n_kf=3, n_supp=1, kind="overlap", seed=50 + i, affine=True)
        for f in w['frames']:
            for key in ('T', 'image', 'K', 'aff', 'k'):
                f[key] = None if f[key] is None else f[key].to(device)
            if f['kf'] is not None:
                f['kf'] = f['kf'].to(device)
                f['image'] = f['kf'].image
        wins.append(w)
    mw = MappingWindows(wins)
    mw.run(10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    mw.run(iters)
    b.record()
    torch.cuda.synchronize()
    eager_ms = a.elapsed_time(b) / iters
    graph = mw.capture(20)
    graph.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    graph_ms = a.elapsed_time(b) / 60
    best = min(eager_ms, graph_ms)
    return {"windows": n_windows, "edges": mw.n_edges, "frames": mw.n_frames,
            "eager_us_per_iter": 1e3 * eager_ms, "graph_us_per_iter": 1e3 * graph_ms,
            "window_iters_per_s": n_windows * 1e3 / best, "edge_iters_per_s": mw.n_edges * 1e3 / best,
            "algorithmic_GBps": mw.algorithmic_bytes_per_iter() / (best * 1e-3) / 1e9,
            "what": "device-resident mapping windows (3 keyframes + 3 supporting frames, 9 edges, 288x224, 64 segments, "
                    "brightness terms on): gradient kernel over all edges + finalize + coupled Adam/pose update"}


# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = {"pinned": False, "skipped": True} if args.no_numa_pin else pin_rank_to_gpu_numa(local_rank, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.workload != "c2":
        import bench_workloads as bw
        bw.RUNNERS[args.workload](bw.Ctx(args, rank, world, device, dist))
        if world > 1:
            dist.destroy_process_group()
        return
    from super_primitive_b200.shard import gather_results

    steps, warm = max(1, args.steps), max(3, args.warmup)
    batch, problems = build_batch(args.pairs, device, seed0=1000 * rank)
    # both kinds are COMPLETE iterations (derivatives + parameter update + retraction on the device, two launches):
    # gn = IRLS Gauss-Newton/LM; grad = the reference's kind (cost + first-order gradient + torch.optim.Adam update with
    # the reference's learning rates, odometery/two_frame_sfm.py:117-121)
    step_fn = batch.gn_step if args.mode == "gn" else batch.adam_step
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step_fn()
    barrier()
    launches0 = batch.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in kev:           # force creation of the native events before handing out raw handles
        a.record()
        b.record()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        step_fn(kev[i])
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    launches = batch.launches - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    pairs_total = args.pairs * world
    value = pairs_total * steps / (total_ms * 1e-3)

    # ---- the other iteration kind, same batch, reported alongside ("report both, labelled") --------
    other = None
    other_fn = batch.adam_step if args.mode == "gn" else batch.gn_step
    oev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in oev:
        a.record()
        b.record()
    for _ in range(warm):
        other_fn()
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for i in range(steps):
        other_fn(oev[i])
    o1.record()
    barrier()
    ot = torch.tensor([o0.elapsed_time(o1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ot, op=dist.ReduceOp.MAX)
    other_ms = float(ot.item())
    other_kern_ms = float(np.mean([a.elapsed_time(b) for a, b in oev]))

    # ---- end-to-end arm: host buffers in, results out, every step -------------------------------------
    e2e = None
    if not args.no_e2e:
        hs = HostStaged(batch, problems, chunk=args.e2e_chunk)
        e_steps = max(3, min(steps, 10))

        def timed(mode):
            for _ in range(3):
                hs.step(mode)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(e_steps):
                hs.step(mode)
            b.record()
            barrier()
            tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return pairs_total * e_steps / (float(tt.item()) * 1e-3)

        v_u8 = timed("u8")
        # what the link itself delivers: the same pinned arena copied to the staging buffer with nothing else going on
        # every rank copies at the same time (barrier first): with several GPUs this is the CONCURRENT ceiling of the
        # host links, the slowest rank's figure is the one the fraction is taken against
        hs.copy_stream.synchronize()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(5):
            hs.ingest.upload(hs.arena)
        p1.record()
        torch.cuda.synchronize()
        lt = torch.tensor([5 * hs.ingest.offsets[-1] / (p0.elapsed_time(p1) * 1e-3) / 1e9], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MIN)
        h2d_gbps = float(lt.item())
        v_u8t = timed("u8t")            # informational: the odometry case, only the new frame of every pair travels
        v_raw = timed("raw")            # informational: frames converted to float32 on the host (the reference's image_tt)
        v_packed = timed("packed")      # informational: derived buffers uploaded instead of frames
        v_params = timed("params")      # informational: frames resident (as the reference keeps its KeyFrames)
        e2e = {"value": v_u8, "unit": UNIT,
               "h2d_bytes_per_step": int(hs.h2d["u8"]), "d2h_bytes_per_step": int(hs.d2h), "steps": e_steps,
               "gpu_launches_per_step": hs.launches_per_step["u8"],
               "h2d_link_GBps": h2d_gbps,
               "frac_of_link": (hs.h2d["u8"] * v_u8 / pairs_total) / 1e9 / h2d_gbps,
               "what": "per step: H2D (pinned) of the 8-bit source + target frames of every pair (HWC uint8, as the "
                       "reference's dataset readers deliver them) + pose + seeds; on the device spb_ingest_u8 (image_tt, "
                       "RGBA target, cached source samples, tile-major level buffer; three launches per chunk of %d pairs, "
                       "overlapped with the remaining copies), one GN/LM iteration; D2H of poses, seeds and LM state" % hs.chunk,
               "target_frame_only": {"value": v_u8t, "h2d_bytes_per_step": int(hs.h2d["u8t"]),
                                     "frac_of_link": (hs.h2d["u8t"] * v_u8t / pairs_total) / 1e9 / h2d_gbps,
                                     "what": "the odometry case (odometery/odometery.py:323-403): the source keyframe and "
                                             "its derived buffers stay resident, only the 8-bit TARGET frame of every pair "
                                             "+ pose + seeds travel"},
               "numa": numa,
               "frames_f32": {"value": v_raw, "h2d_bytes_per_step": int(hs.h2d["raw"]),
                              "what": "frames converted to float32 on the host as the reference's image_tt does "
                                      "(12 bytes per pixel over PCIe), re-derived pair by pair"},
               "prepacked": {"value": v_packed, "h2d_bytes_per_step": int(hs.h2d["packed"]),
                             "what": "tile-major level buffer + RGBA target uploaded instead of the frames"},
               "params_only": {"value": v_params, "h2d_bytes_per_step": int(hs.h2d["params"]),
                               "what": "frames resident on the device, only poses + seeds uploaded and results read "
                                       "back every step"}}

    # ---- drop-in arm: the reference's own loop (photomeric_cost -> backward -> Adam.step) through the public
    #      Python surface, ONE pair, device-resident inputs: launch/host-bound, reported for context -------------
    dropin = device_loop = mapping = None
    if rank == 0 and not args.no_e2e:
        dropin = time_dropin_single_pair(device)
        device_loop = time_device_loop_single_pair(device)
        try:            # auxiliary figure (a 'next' row of SURVEY 8(f)): its failure must not take the headline line down
            mapping = time_mapping_windows(device)
        except Exception as e:      # noqa: BLE001
            mapping = {"error": f"{type(e).__name__}: {e}"}

    # ---- the same batch size on SAM-like blob segments (the headline keeps round 1's overlapping strips) ----------
    blobs = None
    if rank == 0 and not args.no_e2e:
        try:
            import bench_workloads as bw
            from super_primitive_b200.solver import AlignmentBatch
            bb = AlignmentBatch(bw.build_pair_units(list(range(args.pairs)), WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["N"],
                                                    device, n_geoms=4, pad=4))
            blobs = {"segments": "8x8 grid cells dilated by 4 px (rows of a segment are 88 px wide instead of 18)",
                     "points_per_pair": bb.points_total // bb.n}
            for name, fn, gn in (("gn", bb.gn_step, True), ("first_order", bb.adam_step, False)):
                for _ in range(warm):
                    fn()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                for a, b in evs:
                    a.record()
                    b.record()
                torch.cuda.synchronize()
                for i in range(steps):
                    fn(evs[i])
                torch.cuda.synchronize()
                kms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
                blobs[name] = {"kernel_ms": kms, "algorithmic_bytes_per_launch": int(bb.algorithmic_bytes_per_iter(gn=gn))}
            del bb
        except Exception as e:      # noqa: BLE001
            blobs = {"error": f"{type(e).__name__}: {e}"}

    # ---- the only collective of the path: final gather of poses / seeds / cost -------------------------
    gather_ms = None
    if world > 1:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        costs = batch.lm_state[:, 1] / (3.0 * batch.pts_per_problem)
        gather_results(batch.poses_matrix(), batch.k_padded(), costs, pairs_total)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)

    if rank == 0:
        alg_bytes = batch.algorithmic_bytes_per_iter(gn=(args.mode == "gn"))
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        # DRAM bytes of the fused kernel per launch come from an ncu capture (profiles/traffic.json); the figure is only
        # reported when it was captured on exactly this source tree (hash of csrc/ + the header), else null
        traffic, traffic_note = None, "no ncu capture for this source tree"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                import __graft_entry__ as entry
                tj = json.load(open(tp))
                if tj.get("lib_hash") == entry.source_hash():
                    traffic = tj.get(f"{args.mode}_bytes_per_launch_{args.pairs}pairs")
                    traffic_note = tj.get("source")
                else:
                    traffic_note = "profiles/traffic.json was captured on a different source tree (hash mismatch)"
            except Exception:
                traffic = None
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            kind, its, ms, cores = time_cpu(12, 3)
            cpu_baseline = {"value": its, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": "12 timed iterations (3 warm-up) of the reference iteration (photomeric_cost "
                                      "forward + backward + Adam.step) on ONE 640x480 / 64-segment pair, finest "
                                      f"level, float32, {cores} torch threads (fastest candidate on {host_cores()} host "
                                      f"cores); {ms:.1f} ms/iter"}
        ws = sum(p['pack'].numel() * 4 + p['trg_rgba'].numel() * 4 for p in problems)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2 two-frame SfM 640x480, 64 overlapping segments (P=%d points/pair), finest "
                                       "pyramid level" % (batch.points_total // batch.n),
                           "frames": "8-bit synthetic frames (uint8 HWC); float frames = the reference's image_tt of them",
                           "pairs_per_gpu": args.pairs, "iteration": "IRLS Gauss-Newton/LM" if args.mode == "gn"
                           else "cost + first-order gradient + Adam update (the reference's iteration kind)",
                           "parallelism": f"shard{world}",
                           "working_set_bytes_per_gpu": int(ws), "l2": "inputs larger than L2 (126 MB), no flush"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                             "peak_source": peak_src,
                             "kernel": "k_align_global<GN>" if args.mode == "gn" else "k_align_global<GRAD>",
                             "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(alg_bytes)},
                "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        o_bytes = batch.algorithmic_bytes_per_iter(gn=(args.mode != "gn"))
        line["other_iteration"] = {
            "iteration": "cost + first-order gradient + Adam update + retraction (the reference's iteration kind: "
                         "forward, backward, Adam.step)" if args.mode == "gn" else "IRLS Gauss-Newton/LM",
            "value": pairs_total * steps / (other_ms * 1e-3), "unit": UNIT, "ms_per_step": other_ms / steps,
            "kernel_ms": other_kern_ms, "roofline_frac": o_bytes / (other_kern_ms * 1e-3) / 1e9 / peak}
        if dropin is not None:
            line["dropin_single_pair"] = dropin
            line["device_loop_single_pair"] = device_loop
            line["mapping_windows"] = mapping
        if blobs is not None:
            for name in ("gn", "first_order"):
                if name in blobs:
                    blobs[name]["roofline_frac"] = blobs[name]["algorithmic_bytes_per_launch"] / (blobs[name]["kernel_ms"] * 1e-3) / 1e9 / peak
            line["blob_segments"] = blobs
        if gather_ms is not None:
            line["final_gather_ms"] = gather_ms
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
