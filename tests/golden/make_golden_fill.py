"""Golden fixture for the nearest-valid hole filling (build container only, needs /root/reference and scipy):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fill.py

Runs the reference's OWN `fill_depth` (depth_completion/fill_in_tools.py:5-7, imported from the checkout; scipy's
`distance_transform_edt` underneath) on seeded synthetic maps and stores inputs, scipy's index arrays and the filled maps
for the small cases; for the 480x640 case (a completed VOID frame: most pixels valid, ragged holes) only the bit-packed
mask and the sha256 of scipy's index array are stored (the depth is the pixel's own linear index, so the filled map IS
the index map).  `oracle/fill_oracle.py` is asserted to reproduce all of it bit for bit.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

from scipy import ndimage as nd                                      # noqa: E402
from depth_completion.fill_in_tools import fill_depth               # noqa: E402  the reference's own function
from oracle import fill_oracle as port                              # noqa: E402


def holes(H, W, seed, n_blocks, p_noise):
    rng = np.random.default_rng(seed)
    inv = rng.random((H, W)) < p_noise
    for _ in range(n_blocks):
        r0, c0 = int(rng.integers(0, H)), int(rng.integers(0, W))
        inv[r0:r0 + int(rng.integers(1, max(2, H // 3))), c0:c0 + int(rng.integers(1, max(2, W // 3)))] = True
    return inv


def cases():
    out = {}
    out["blocks"] = holes(48, 64, 1, 6, 0.0)
    out["noise"] = holes(37, 53, 2, 0, 0.9)
    out["mixed"] = holes(64, 80, 3, 5, 0.3)
    out["sparse"] = holes(40, 40, 4, 0, 0.995)
    one = np.ones((21, 34), bool); one[13, 7] = False
    out["one_valid"] = one
    out["all_valid"] = np.zeros((9, 11), bool)
    out["all_invalid"] = np.ones((6, 8), bool)
    col = np.ones((30, 17), bool); col[:, 5] = False                     # one valid column: ties above / below never arise,
    out["one_column"] = col                                              # left / right do not either; rows tie with holes:
    chk = np.ones((16, 16), bool); chk[::4, ::4] = False                 # lattice of valid pixels: many exact ties
    out["lattice"] = chk
    return out


def main():
    store = {}
    for name, inv in cases().items():
        H, W = inv.shape
        rng = np.random.default_rng(len(name) * 7 + H)
        depth = (0.5 + rng.random((H, W))).astype(np.float32)
        ind = nd.distance_transform_edt(inv, return_distances=False, return_indices=True)
        ref = fill_depth(depth, inv)
        mine_ind = port.nearest_valid_indices(inv)
        assert np.array_equal(mine_ind, ind), name
        assert np.array_equal(port.fill_depth(depth, inv), ref), name
        store[name + "_invalid"] = inv
        store[name + "_depth"] = depth
        store[name + "_indices"] = ind.astype(np.int32)
        store[name + "_filled"] = ref
    # a completed 480x640 frame
    H, W = 480, 640
    inv = holes(H, W, 11, 40, 0.02)
    inv[:, :9] = True                                                    # a border no segment reaches
    lin = (np.arange(H * W, dtype=np.float32)).reshape(H, W)             # exact in float32 (< 2^24)
    ind = nd.distance_transform_edt(inv, return_distances=False, return_indices=True).astype(np.int32)
    ref = fill_depth(lin, inv)
    assert np.array_equal(ref, (ind[0] * W + ind[1]).astype(np.float32))
    assert np.array_equal(port.nearest_valid_indices(inv), ind)
    store["vga_invalid_bits"] = np.packbits(inv)
    store["vga_shape"] = np.array([H, W])
    store["vga_indices_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ind).tobytes()).digest(), dtype=np.uint8)
    import scipy
    store["scipy_version"] = np.array(scipy.__version__)
    np.savez_compressed(os.path.join(HERE, "fill_depth.npz"), **store)
    print("written", os.path.join(HERE, "fill_depth.npz"), os.path.getsize(os.path.join(HERE, "fill_depth.npz")), "bytes")


if __name__ == "__main__":
    main()
