#!/usr/bin/env python
"""Host-side profile of the reference's two-frame loop through the drop-in API on one C2 pair (GPU box):
cProfile of `photomeric_cost -> mean -> backward -> Adam.step`, top functions by cumulative time."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from super_primitive_b200 import dense_optim as do, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
H, W, N = bench.WORKLOAD["H"], bench.WORKLOAD["W"], bench.WORKLOAD["N"]
src, trg, k0, pose0 = syn.two_frame_problem(H, W, N, kind=bench.WORKLOAD["kind"], seed=0, noise=0.01)
src, trg = src.to(dev), trg.to(dev)
k = torch.nn.Parameter(k0.to(dev))
pose = torch.nn.Parameter(pose0.to(dev))
opt = torch.optim.Adam([{'params': [k], 'lr': 1e-3}, {'params': [pose], 'lr': 1e-2}], lr=1e-3)
cfg = {'mode': 'colour', 'collect_stats': 0}


def fwd():
    return do.photomeric_cost(src, trg, k, pose, cfg)['residual'].mean()


def one():
    loss = fwd()
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(20):
    one()
torch.cuda.synchronize()
n = 200
for name, fn in (("forward only", fwd), ("full iteration", one)):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / n * 1e6:.1f} us/iter")
t0 = time.perf_counter()
for _ in range(n):
    loss = fwd()
    loss.backward()
torch.cuda.synchronize()
print(f"forward+backward: {(time.perf_counter() - t0) / n * 1e6:.1f} us/iter")
t0 = time.perf_counter()
for _ in range(n):
    opt.step()
torch.cuda.synchronize()
print(f"Adam.step alone: {(time.perf_counter() - t0) / n * 1e6:.1f} us/iter")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    one()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
