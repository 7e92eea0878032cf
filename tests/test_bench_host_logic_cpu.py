"""CPU: the host-side logic of bench.py's end-to-end arm (arena layout, chunked uploads, byte accounting, every mode of
HostStaged.step) exercised without a GPU: the native calls and the CUDA stream/event objects are replaced by inert
stand-ins, so a Python-level mistake in the arm shows up here and not on the GPU box."""
import contextlib
import sys
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Lib:
    def __getattr__(self, name):
        return lambda *a: 0


class _Stream:
    cuda_stream = 0

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, **kw):
        pass

    def record(self, s=None):
        pass


class _Geom:
    P_pad, n_tiles, cref = 128, 1, None


class _Batch:
    def __init__(self, n, g):
        self.n, self.geoms = n, [g]
        self.poses, self.k, self.lm_state = torch.zeros(n, 16), torch.zeros(8 * n), torch.zeros(n, 8)
        self.d_geoms = torch.zeros(8)
        self.steps = 0

    def gn_step(self):
        self.steps += 1


@pytest.fixture
def stubbed(monkeypatch):
    import bench
    from super_primitive_b200 import _native as nat, frames
    monkeypatch.setattr(nat, "lib", lambda: _Lib())
    monkeypatch.setattr(frames, "_stream", lambda: 0)
    monkeypatch.setattr(frames, "_struct_array_to_device", lambda arr, dev: torch.zeros(len(bytes(arr)), dtype=torch.uint8))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    return bench, frames


def _problems(n, g, H=6, W=8):
    gen = torch.Generator().manual_seed(0)
    out = []
    for _ in range(n):
        out.append(dict(geom=g, src_rgb=torch.zeros(3, 128), pack=torch.zeros(1, 644, dtype=torch.int32),
                        trg_rgba=torch.zeros(H, W, 4), src_image=torch.zeros(3, H, W), trg_image=torch.zeros(3, H, W),
                        src_u8=torch.randint(0, 256, (H, W, 3), generator=gen, dtype=torch.uint8),
                        trg_u8=torch.randint(0, 256, (H, W, 3), generator=gen, dtype=torch.uint8)))
    return out


def test_frame_arena_layout_and_chunked_upload(stubbed):
    _, frames = stubbed
    g = _Geom()
    probs = _problems(5, g)
    ing = frames.FrameIngest(probs, [g])
    per = 2 * 6 * 8 * 3
    assert ing.offsets == [i * per for i in range(6)] and ing.frame_bytes() == 5 * per
    assert ing.frame_bytes(1, 2) == 2 * per
    arena = ing.host_arena()
    arena.zero_()
    for i, p in enumerate(probs):
        ing.fill(arena, i, p['src_u8'], p['trg_u8'])
    for i, p in enumerate(probs):              # source frame, then target frame, problem after problem
        assert torch.equal(arena[i * per:i * per + per // 2].reshape(6, 8, 3), p['src_u8'])
        assert torch.equal(arena[i * per + per // 2:(i + 1) * per].reshape(6, 8, 3), p['trg_u8'])
    ing.stage.zero_()
    ing.upload(arena, 1, 2)                    # ONE copy moves exactly the frames of problems 1 and 2
    assert torch.equal(ing.stage[per:3 * per], arena[per:3 * per])
    assert int(ing.stage[:per].sum()) == 0 and int(ing.stage[3 * per:].sum()) == 0


def test_host_staged_runs_every_mode_and_accounts_bytes(stubbed):
    bench, _ = stubbed
    g = _Geom()
    probs = _problems(5, g)
    batch = _Batch(5, g)
    hs = bench.HostStaged(batch, probs, chunk=2)
    assert len(hs.chunk_events) == 3
    for mode in ("u8", "raw", "packed", "params"):
        hs.step(mode)
        hs.step(mode)                          # second step takes the "previous chunk consumed" branch
    assert batch.steps == 8
    params = (5 * 16 + 40) * 4
    assert hs.params_bytes == params
    assert hs.h2d["u8"] == 5 * 2 * 6 * 8 * 3 + params
    assert hs.h2d["raw"] == 5 * 2 * 3 * 6 * 8 * 4 + params
    assert hs.h2d["params"] == params and hs.d2h == (5 * 16 + 40 + 5 * 8) * 4
    assert hs.launches_per_step["u8"] == 3 * 3 + 2
    assert torch.equal(hs.ingest.stage, hs.arena)      # every chunk was uploaded


def test_workload_helpers_are_deterministic_in_the_unit_id():
    """bench_workloads: any rank can rebuild any unit from its id (the shard check relies on it)."""
    import bench_workloads as bw
    assert bw.grid_shape(64, 480, 640) == (8, 8) and bw.grid_shape(100, 224, 288) == (10, 10)
    assert bw.grid_shape(256, 768, 1024) == (16, 16) and bw.grid_shape(300, 224, 288) == (20, 15)
    a, b = bw.start_pose(17), bw.start_pose(17)
    assert torch.equal(a, b) and not torch.equal(a, bw.start_pose(18))
    assert set(bw.RUNNERS) == {"c2levels", "c3", "c4", "c5", "compaction"}


def test_grid_sizing_fills_whole_waves():
    """spb_gn_ctas (no GPU needed: the SM count falls back to 148): the CTA count per pair gives >= 3 waves with a last
    wave >= 97 % full whenever the work allows, e.g. 27 CTAs x 64 pairs = 3.89 waves at 3 CTAs/SM, 3 x 1024 pairs = 6.92."""
    from super_primitive_b200 import _native as nat
    lib = nat.lib()
    for tiles, pairs in ((4352, 64), (7100, 1024), (1100, 256), (4352, 128)):
        c = lib.spb_gn_ctas(tiles, pairs)
        assert 1 <= c <= (tiles + 7) // 8
        # spb_gn_ctas is the maximum over the kernel variants (2, 3, 4 CTAs/SM); each variant's own grid is wave-aligned
        assert pairs * c >= 3 * 148 * 2
    assert lib.spb_gn_ctas(384, 1) == 48                         # few pairs: every warp gets one tile
    assert lib.spb_gn_work_stride(4352, 64) >= lib.spb_gn_ctas(4352, 64) * 47 + 4352 * 19
