#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
drop-in cost + backward, GN/LM and Adam device iterations, one mapping-window iteration, depth splat, re-initialisation.
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from super_primitive_b200 import dense_optim as do, depth_init, depth_render, synthetic as syn  # noqa: E402
from super_primitive_b200.solver import AlignmentBatch, make_problem  # noqa: E402
from super_primitive_b200.window import MappingWindows  # noqa: E402

dev = "cuda"
cfg = {'mode': 'colour', 'collect_stats': 0}
src, trg, k0, pose0 = syn.two_frame_problem(48, 64, 5, kind="rects", seed=1, noise=0.01)
src, trg = src.to(dev), trg.to(dev)
k = k0.to(dev).requires_grad_(True)
pose = pose0.to(dev).requires_grad_(True)
out = do.photomeric_cost(src, trg, k, pose, dict(cfg, collect_stats=2))
out['residual'].mean().backward()
_ = out['residual_raw']
pre = do.unproject_kf(src, k0.to(dev))
do.photomeric_cost_precomputed(pre, trg, pose0.to(dev).requires_grad_(True), cfg)['residual'].mean().backward()
for aff in (False, True):
    a = (torch.tensor([0.01, 0.0], device=dev), torch.tensor([0.0, 0.01], device=dev)) if aff else (None, None)
    b = AlignmentBatch([make_problem(src, trg.image, trg.K, pose0.to(dev), k0.to(dev), aff_src=a[0], aff_trg=a[1])
                        for _ in range(2)], with_affine=aff)
    b.run_gn(3)
    b.run_adam(3)
w = syn.mapping_window(48, 64, 4, n_kf=2, n_supp=1, kind="overlap", seed=2, affine=True)
for f in w['frames']:
    for key in ('T', 'image', 'K', 'aff', 'k'):
        f[key] = None if f[key] is None else f[key].to(dev)
    if f['kf'] is not None:
        f['kf'] = f['kf'].to(dev)
        f['image'] = f['kf'].image
mw = MappingWindows([w])
mw.run(2)
d = depth_render.estimate_depth_kf_native(src, k0.to(dev), pose0.to(dev))
depth_init.segment_based_depth_reinit(d, src, 'median')
torch.cuda.synchronize()
print("sanitize_run ok", float(out['residual']), float(b.costs()[0]))
