"""GPU: the end-to-end arm of bench.py really rebuilds what the fused kernel streams from the uploaded frames.

`HostStaged.step("u8")` uploads the 8-bit source / target frames and re-derives the float frames (the reference's
image_tt), the RGBA target and the tile-major level buffer on the device in three launches per chunk (spb_ingest_u8);
`step("raw")` does the same from float32 frames pair by pair (spb_pack_rgba, spb_sample_source, spb_build_tile_pack).
The rebuilt buffers must be bit-identical to the ones the batch was created with, and the iteration that follows must
give the same result as the device-resident iteration from the same parameters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_image_tt_is_bit_identical_to_the_reference_helper():
    """tool/etc.py:37-40: (torch.from_numpy(image) / 255.).float() then HWC -> CHW, here on the device."""
    from super_primitive_b200.frames import image_tt
    g = torch.Generator().manual_seed(0)
    for H, W in ((48, 64), (37, 53), (1, 1)):
        img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
        want = (img / 255.).float().permute(2, 0, 1)
        got = image_tt(img.numpy(), "cuda:0")
        assert got.dtype == torch.float32 and tuple(got.shape) == (3, H, W)
        assert torch.equal(got.cpu(), want)
    all_values = torch.arange(256, dtype=torch.uint8).repeat(3).reshape(3, 256).t().contiguous().reshape(16, 16, 3)
    assert torch.equal(image_tt(all_values, "cuda:0").cpu(), (all_values / 255.).float().permute(2, 0, 1))
    with pytest.raises(AssertionError):
        image_tt(torch.zeros((4, 4, 3)), "cuda:0")


@pytest.mark.parametrize("mode", ["u8", "raw", "u8-lean"])
def test_frame_upload_rebuilds_identical_buffers_and_results(mode, monkeypatch):
    import bench
    from super_primitive_b200 import _native as nat
    if mode == "u8-lean":
        # experiment builds only (-DSPB_INGEST_FUSED=1, selected with SPB200_LIB): the tile pack is derived straight from
        # the 8-bit source frame, the planar float frame and the sample array are not materialised
        if not (nat.lib().spb_version() // 1000) & 1:
            pytest.skip("default library: no fused source ingest")
        monkeypatch.setenv("SPB_E2E_LEAN", "1")
        mode = "u8"
        lean = True
    else:
        lean = False
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
        hs.chunk = 2                      # 3 pairs -> two ingest chunks (the chunk boundary is exercised)
        hs.chunk_events = [torch.cuda.Event() for _ in range(2)]
        hs.step(mode)
        torch.cuda.synchronize()
        for p, rp, rr in zip(problems, ref_pack, ref_rgba):
            assert torch.equal(p['pack'], rp)
            assert torch.equal(p['trg_rgba'], rr)
        assert torch.equal(batch.poses, want_pose) and torch.equal(batch.k, want_k)
        assert torch.equal(hs.o_pose, want_pose.cpu()) and torch.equal(hs.o_k, want_k.cpu())
        assert hs.h2d["raw"] == sum(2 * 3 * 96 * 128 * 4 for _ in problems) + hs.params_bytes
        assert hs.h2d["u8"] == sum(2 * 3 * 96 * 128 for _ in problems) + hs.params_bytes
        if mode == "u8" and not lean:     # the float frames themselves were rebuilt too (source planar copy)
            for p, pl in zip(problems, hs.ingest.src_planar):
                # the reference converts on the HOST (tool/etc.py:37-40); torch's CUDA division by a scalar multiplies by
                # the reciprocal and differs in the last bit
                assert torch.equal(pl.cpu(), (p['src_u8'].cpu() / 255.).float().permute(2, 0, 1))
    finally:
        bench.WORKLOAD.clear()
        bench.WORKLOAD.update(saved)
