"""Golden fixture for the frontend hand-over (build container only, needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_handover.py

The four lines that end `FrontProcessorNew.process_to_kf` (frontend/process_frame.py:231-236 -- the class itself cannot be
imported here: it loads SAM and the normal network) are executed around the reference's OWN `put_keypoints_back`
(image/keyframe.py:151-173, imported from the checkout) on seeded synthetic `integrated_depth` tensors; stored: inputs and
the reference's keypoints / masks / log-depth.  `oracle/frontend_handover.py` is asserted to reproduce them bit for bit.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

from image.keyframe import put_keypoints_back          # noqa: E402  the reference's own function
from oracle import frontend_handover as port            # noqa: E402


def synthetic_depth(N, Hf, Wf, seed):
    g = torch.Generator().manual_seed(seed)
    d = torch.zeros(N, Hf, Wf)
    kps = torch.zeros(N, 2)
    for b in range(N):
        h, w = int(torch.randint(6, Hf // 2, (1,), generator=g)), int(torch.randint(6, Wf // 2, (1,), generator=g))
        r0, c0 = int(torch.randint(0, Hf - h, (1,), generator=g)), int(torch.randint(0, Wf - w, (1,), generator=g))
        yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        d[b, r0:r0 + h, c0:c0 + w] = 0.8 + 0.02 * b + 0.01 * yy + 0.005 * xx
        d[b, r0 + h // 3:r0 + h // 3 + 2, c0 + w // 3:c0 + w // 3 + 3] = 0.0          # a hole: ragged rows
        # keypoints: some inside, some in the hole, some outside the segment (they get snapped)
        kr, kc = (r0 + h // 3, c0 + w // 3) if b % 3 == 0 else ((r0 + h // 2, c0 + w // 2) if b % 3 == 1 else (r0 - 3, c0 + w + 2))
        kps[b] = torch.tensor([2.0 * kr / (Hf - 1) - 1, 2.0 * kc / (Wf - 1) - 1]).clamp(-1, 1)
    # a segment that vanishes under the resampling (one texel at an odd position) and an all-empty one
    d[N - 2] = 0.0
    d[N - 2, 1, 1] = 1.5
    d[N - 1] = 0.0
    return d, kps


def main():
    store = {}
    for name, (N, Hf, Wf, H, W, seed) in {"half": (9, 96, 128, 48, 64, 3), "odd": (8, 90, 122, 40, 56, 5),
                                           "same": (6, 48, 64, 48, 64, 7)}.items():
        depth, kps = synthetic_depth(N, Hf, Wf, seed)
        # frontend/process_frame.py:231-236
        logdepth = torch.nn.functional.interpolate(depth[:, None], size=(H, W), mode='nearest')[:, 0]
        masks = logdepth > 1e-7
        keypoints, masks, logdepth = put_keypoints_back(kps.clone(), masks, logdepth)
        logdepth[masks] = torch.log(logdepth[masks])
        k2, m2, l2, good = port.handover(depth, kps.clone(), (H, W))
        assert torch.equal(k2, keypoints) and torch.equal(m2, masks) and torch.equal(l2, logdepth), name
        assert int(good.sum()) == masks.shape[0] < N or name == "same"
        store.update({f"{name}_depth": depth.numpy(), f"{name}_kps": kps.numpy(), f"{name}_size": np.array([H, W]),
                      f"{name}_keypoints": keypoints.numpy(), f"{name}_masks": masks.numpy(),
                      f"{name}_logdepth": logdepth.numpy(), f"{name}_good": good.numpy()})
        print(name, "segments", N, "->", masks.shape[0], "points", int(masks.sum()))
    np.savez_compressed(os.path.join(HERE, "handover.npz"), **store)


if __name__ == "__main__":
    main()
