"""GPU: the end-to-end arm of bench.py really rebuilds what the fused kernel streams from the uploaded frames.

`HostStaged.step("raw")` uploads the float32 source / target images and re-derives the RGBA target and the
tile-major level buffer on the device (spb_pack_rgba, spb_sample_source, spb_build_tile_pack).  The rebuilt buffers
must be bit-identical to the ones the batch was created with, and the iteration that follows must give the same
result as the device-resident iteration from the same parameters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_raw_frame_upload_rebuilds_identical_buffers_and_results():
    import bench
    dev = torch.device("cuda:0")
    saved = dict(bench.WORKLOAD)
    bench.WORKLOAD.update(H=96, W=128, N=8)
    try:
        batch, problems = bench.build_batch(3, dev)
        ref_pack = [p['pack'].clone() for p in problems]
        ref_rgba = [p['trg_rgba'].clone() for p in problems]
        hs = bench.HostStaged(batch, problems)
        # reference result: one resident iteration from the initial parameters
        pose0, k0, lm0 = batch.poses.clone(), batch.k.clone(), batch.lm_state.clone()
        batch.gn_step()
        torch.cuda.synchronize()
        want_pose, want_k = batch.poses.clone(), batch.k.clone()
        # wipe everything the raw path has to rebuild, restore the solver state, run the e2e step
        for p in problems:
            p['pack'].zero_()
            p['trg_rgba'].zero_()
            p['src_rgb'].zero_()
            p['src_image'].zero_()
            p['trg_image'].zero_()
        batch.lm_state.copy_(lm0)
        batch.saved_pair.zero_()
        batch.saved_seg.zero_()
        assert torch.equal(hs.h_pose.to(dev), pose0) and torch.equal(hs.h_k.to(dev), k0)
        hs.step("raw")
        torch.cuda.synchronize()
        for p, rp, rr in zip(problems, ref_pack, ref_rgba):
            assert torch.equal(p['pack'], rp)
            assert torch.equal(p['trg_rgba'], rr)
        assert torch.equal(batch.poses, want_pose) and torch.equal(batch.k, want_k)
        assert torch.equal(hs.o_pose, want_pose.cpu()) and torch.equal(hs.o_k, want_k.cpu())
        assert hs.h2d["raw"] == sum(2 * 3 * 96 * 128 * 4 for _ in problems) + hs.params_bytes
    finally:
        bench.WORKLOAD.clear()
        bench.WORKLOAD.update(saved)
