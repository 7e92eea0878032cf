"""CPU: the oracles of the device-resident optimiser loops are pinned to the reference's OWN caller code.

tests/golden/make_golden_callers.py executes `Odometery.track_frame` and `Odometery.mapping` from the reference checkout
unmodified (only the absent lietorch is stubbed with its documented semantics) and freezes what they leave behind;
`oracle/adam_loop.py` (oracle of spb_adam_iterate) and `oracle/window_loop.py` (oracle of spb_window_iterate, checked in
tests/test_window_host_cpu.py) must reproduce those results."""
import math
import os

import numpy as np
import torch

from super_primitive_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))


def test_tracker_oracle_is_pinned_to_the_reference_caller():
    """odometery/odometery.py:323-447 -- Adam on the pose increment (track.lr) and the frame's brightness terms (5e-3) over
    photomeric_cost_precomputed, increment folded and re-zeroed every iteration, seeds untouched."""
    from oracle import adam_loop
    z = np.load(os.path.join(HERE, "golden", "tracker.npz"))
    c = {key[4:]: z[key].item() for key in z.files if key.startswith("cfg_")}
    src, trg, k0, pose0 = syn.two_frame_problem(c['H'], c['W'], c['N'], kind=c['kind'], seed=c['seed'], noise=c['noise'])
    aff = (torch.from_numpy(z["aff_src"]), torch.from_numpy(z["aff_trg0"]))
    got = adam_loop.tracker_adam(src, trg, k0, pose0, c['iters'], lr_pose=c['lr'], lr_k=0.0, lr_aff=5e-3, affine=aff,
                                 opt_affine=True)
    # the reference folds T_frame <- T_frame inv(Delta) in camera-to-world form and renormalises at the end; the oracle
    # iterates the relative pose directly: same iteration, different rounding
    np.testing.assert_allclose(got['pose'].numpy(), z["rel_pose"], atol=5e-6)
    np.testing.assert_allclose(got['aff_trg'].numpy(), z["aff_trg"], atol=1e-6)
    assert torch.equal(got['k'], k0)
    # consistency of the fixture itself, and it is not trivial: the pose moved by ~ lr per step at first
    np.testing.assert_allclose(np.linalg.inv(z["frame_pose"].astype(np.float64)) @ z["kf_pose"], z["rel_pose"], atol=1e-6)
    assert np.abs(z["rel_pose"] - pose0.numpy()).max() > 3 * c['lr']
    assert len(got['costs']) == c['iters'] and all(math.isfinite(x) for x in got['costs'])


def test_sfm_oracle_is_pinned_to_the_reference_caller():
    """odometery/two_frame_sfm.py:127-215 (BASELINE config 0 in miniature): 2 pyramid levels x 500 iterations, one source
    keyframe against two supporting frames, poses as never-re-zeroed increments, first iteration without a step."""
    from oracle import adam_loop
    z = np.load(os.path.join(HERE, "golden", "sfm_run.npz"))
    c = {key[4:]: z[key].item() for key in z.files if key.startswith("cfg_")}
    src = syn.make_keyframe(c['H'], c['W'], c['N'], kind=c['kind'], seed=c['seed'], noise=c['noise'])
    trgs = [syn.make_keyframe(c['H'], c['W'], c['N'], shift=(2.0 + j, 1.0 - 0.5 * j), noise=c['noise'],
                              seed=c['seed'] + 1 + j, supporting=True) for j in range(c['n_supp'])]
    T0s = [torch.from_numpy(T) for T in z["T0s"]]
    got = adam_loop.sfm_adam(syn.keyframe_pyramid(src, 0, c['levels']),
                             [syn.keyframe_pyramid(t, 0, c['levels']) for t in trgs], torch.from_numpy(z["k0"]), T0s, 500)
    assert np.array_equal(got['k'].numpy(), z["k"])
    assert np.array_equal(np.stack([d.numpy() for d in got['deltas']]), z["deltas"])
    assert got['losses'][0] == float(z["loss_first"]) and got['losses'][-1] == float(z["loss_last"])
    assert got['losses'][-1] < 0.1 * got['losses'][0]                 # the run is a real optimisation, not a no-op
    assert len(got['losses']) == 500 * c['levels']
