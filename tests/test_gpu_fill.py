"""GPU: nearest-valid hole filling (super_primitive_b200/fill_in_tools.py, csrc/spb_fill.cu) against
tests/golden/fill_depth.npz -- the reference's own `fill_depth` (scipy's Euclidean feature transform) -- and against
oracle/fill_oracle.py on seeded maps.  Integer work: bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    return np.load(os.path.join(HERE, "golden", "fill_depth.npz"))


def test_fill_depth_equals_the_reference_on_every_golden_case():
    from super_primitive_b200.fill_in_tools import fill_depth
    z = _golden()
    names = sorted(k[:-len("_invalid")] for k in z.files if k.endswith("_invalid"))
    assert len(names) >= 9
    for name in names:
        inv, depth = z[f"{name}_invalid"], z[f"{name}_depth"]
        out, idx = fill_depth(torch.from_numpy(depth).cuda(), torch.from_numpy(inv).cuda(), return_indices=True)
        assert np.array_equal(idx.cpu().numpy(), z[f"{name}_indices"]), name
        assert np.array_equal(out.cpu().numpy(), z[f"{name}_filled"]), name
        # the reference's calling convention: numpy in, numpy out (evaluate_void.py:124-125)
        out_np = fill_depth(depth, inv)
        assert isinstance(out_np, np.ndarray) and out_np.dtype == depth.dtype
        assert np.array_equal(out_np, z[f"{name}_filled"]), name


def test_fill_depth_vga_frame_matches_scipys_indices():
    from super_primitive_b200.fill_in_tools import fill_depth
    z = _golden()
    H, W = (int(v) for v in z["vga_shape"])
    inv = np.unpackbits(z["vga_invalid_bits"])[:H * W].reshape(H, W).astype(bool)
    lin = torch.arange(H * W, dtype=torch.float32, device="cuda").reshape(H, W)
    out, idx = fill_depth(lin, torch.from_numpy(inv).cuda(), return_indices=True)
    idx = idx.cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(idx).tobytes()).digest() == z["vga_indices_sha256"].tobytes()
    assert np.array_equal(out.cpu().numpy(), (idx[0] * W + idx[1]).astype(np.float32))


@pytest.mark.parametrize("shape,p,seed", [((1, 1), 0.0, 0), ((1, 97), 0.8, 1), ((83, 1), 0.8, 2), ((61, 300), 0.97, 3),
                                            ((200, 333), 0.5, 4), ((128, 2100), 0.999, 5)])
def test_fill_depth_batch_equals_the_oracle(shape, p, seed):
    """ragged sizes (one row, one column, wider than one CTA pass, almost empty) and a batch whose frames differ"""
    from oracle import fill_oracle as port
    from super_primitive_b200.fill_in_tools import fill_depth_batch
    rng = np.random.default_rng(seed)
    F = 3
    inv = rng.random((F,) + shape) < p
    inv[1] = rng.random(shape) < min(1.0, p + 0.0005)
    if shape[0] * shape[1] > 1:
        inv[2] = True                                   # a frame with no valid pixel next to ordinary ones
    depth = (0.5 + rng.random((F,) + shape)).astype(np.float32)
    out, idx = fill_depth_batch(torch.from_numpy(depth).cuda(), torch.from_numpy(inv).cuda(), return_indices=True)
    for f in range(F):
        want = port.nearest_valid_indices(inv[f])
        assert np.array_equal(idx[f].cpu().numpy(), want), f
        assert np.array_equal(out[f].cpu().numpy(), port.fill_depth(depth[f], inv[f])), f


def test_fill_depth_keeps_float64_values_and_refuses_host_tensors():
    from oracle import fill_oracle as port
    from super_primitive_b200.fill_in_tools import fill_depth, fill_single_griddata
    rng = np.random.default_rng(9)
    inv = rng.random((40, 50)) < 0.7
    depth = rng.random((40, 50))                        # float64, not representable in float32
    out = fill_depth(depth, inv)
    assert out.dtype == np.float64 and np.array_equal(out, port.fill_depth(depth, inv))
    with pytest.raises(RuntimeError):
        fill_depth(torch.from_numpy(depth), torch.from_numpy(inv))       # CPU tensors: no CPU path
    with pytest.raises(NotImplementedError):
        fill_single_griddata(depth, inv)


def test_complete_batch_can_fill_the_holes_of_its_maps():
    from oracle import fill_oracle as port
    from super_primitive_b200 import depth_completion as dc, synthetic as syn
    H, W, N = 96, 128, 12
    kfs, sparse = [], []
    for seed in (1, 2):
        kf = syn.make_keyframe(H, W, N, kind="rects", seed=seed).to("cuda")
        sp = torch.zeros(H, W, device="cuda")
        g = torch.Generator().manual_seed(seed)
        rr, cc = torch.randint(0, H, (400,), generator=g), torch.randint(0, W, (400,), generator=g)
        sp[rr, cc] = (1.0 + 0.5 * torch.rand(400, generator=g)).cuda()
        kfs.append(kf)
        sparse.append(sp)
    res = dc.complete_batch(kfs, sparse, 'median', fill_holes=True)
    for depth, invalid, _k, _vis, filled in res:
        assert invalid.any() and not invalid.all()
        want = port.fill_depth(depth.cpu().numpy(), invalid.cpu().numpy())
        assert np.array_equal(filled.cpu().numpy(), want)
        assert torch.equal(filled[~invalid], depth[~invalid])
