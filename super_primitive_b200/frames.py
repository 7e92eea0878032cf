"""Frame ingest: 8-bit frames in, everything the fused alignment kernel streams out.

The reference turns a dataset frame (HWC uint8 from cv2: data/replica.py:55, data/tum_undistort.py:112) into a float
tensor on the HOST (`tool.etc.image_tt`, tool/etc.py:37-40: ``(torch.from_numpy(image) / 255.).float().to(device)`` then
HWC -> CHW) and uploads 12 bytes per pixel.  Here the 3-byte pixels are uploaded and `image_tt` runs on the device with
the same float32 division, so the resulting frames are bit-identical (checked on the GPU against torch).

``image_tt``      mirror of the reference helper for one frame (uint8 HWC tensor, host or device).
``FrameIngest``   many (source keyframe, target frame) pairs per launch: `spb_ingest_u8` re-derives the RGBA target, the
                  cached source samples and the tile-major level buffer of every pair in three launches
                  (include/spb200.h) -- what `bench.py`'s end-to-end arm runs every step.
"""
from __future__ import annotations

import torch

from . import _native as nat
from .geometry import _stream
from .solver import _struct_array_to_device


def image_tt(image, device="cuda"):
    """tool/etc.py:37-40: HWC uint8 frame (numpy array or tensor) -> (3,H,W) float32 in [0,1] on ``device``."""
    t = torch.as_tensor(image)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise AssertionError("image_tt expects an HWC uint8 frame with 3 channels")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    t = t.to(dev, non_blocking=True).contiguous()
    H, W = int(t.shape[0]), int(t.shape[1])
    out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    nat.check(nat.lib().spb_image_tt(t.data_ptr(), H, W, out.data_ptr(), _stream()), "spb_image_tt")
    return out


class FrameIngest:
    """Device-side ingest of the frames of ``problems`` (dicts as `solver.AlignmentBatch` takes them: geom, src_rgb, pack,
    trg_rgba; the level size is the target's).

    The 8-bit frames of all problems live in ONE flat device staging buffer ``stage`` (problem after problem: source
    frame, then target frame, each (Hl,Wl,3) uint8), mirrored by a pinned host arena of the same layout
    (`host_arena`) that a loader decodes into -- so a chunk of pairs is ONE host->device copy (`upload`) followed by
    the three launches of `spb_ingest_u8` (`run`)."""

    def __init__(self, problems, geoms, lean=False, target_only=False):
        """``target_only``: only the TARGET frame of every pair arrives (the odometry case: the source keyframe and
        everything derived from it stay resident, a new camera frame is aligned against it; odometery/odometery.py:323-403);
        the staging buffer then holds one frame per problem and the source half of the ingest is skipped.
        ``lean``: do not materialise the planar float source frame and the [3][n_pad] sample array (only the
        statistics path reads them).  Needs a library built with the fused source ingest (experiment switch,
        ``spb_version() // 1000 & 1``); the default library derives the tile pack from those two buffers."""
        import ctypes as C
        if lean and not (nat.lib().spb_version() // 1000) & 1:
            raise RuntimeError("FrameIngest(lean=True) needs a library built with -DSPB_INGEST_FUSED=1")
        dev = problems[0]['trg_rgba'].device
        gidx = {id(g): i for i, g in enumerate(geoms)}
        self.n = len(problems)
        sizes = [3 * int(p['trg_rgba'].shape[0]) * int(p['trg_rgba'].shape[1]) for p in problems]
        self.offsets = [0]
        self.target_only = bool(target_only)
        per = 1 if target_only else 2
        for b in sizes:
            self.offsets.append(self.offsets[-1] + per * ((b + 15) // 16 * 16))      # 16-byte aligned frames
        self.stage = torch.empty(self.offsets[-1], dtype=torch.uint8, device=dev)
        self.src_planar = []
        jarr = (nat.SpbFrameJob * self.n)()
        self.max_pixels = self.max_pad = self.max_tiles = 1
        self._views = []
        for i, p in enumerate(problems):
            Hl, Wl = int(p['trg_rgba'].shape[0]), int(p['trg_rgba'].shape[1])
            g = p['geom']
            half = (self.offsets[i + 1] - self.offsets[i]) // per
            so, to = self.offsets[i], self.offsets[i] + (0 if target_only else half)
            self._views.append((so, to, Hl, Wl))
            pl = None if (lean or target_only) else torch.empty((3, Hl, Wl), dtype=torch.float32, device=dev)
            self.src_planar.append(pl)
            j = jarr[i]
            j.src_u8 = None if target_only else self.stage.data_ptr() + so
            j.trg_u8 = self.stage.data_ptr() + to
            j.src_planar = None if pl is None else pl.data_ptr()
            j.src_rgb = None if (lean or target_only) else p['src_rgb'].data_ptr()
            j.pack = None if target_only else p['pack'].data_ptr()
            j.trg_rgba = p['trg_rgba'].data_ptr()
            j.geom, j.Hl, j.Wl = gidx[id(g)], Hl, Wl
            self.max_pixels = max(self.max_pixels, Hl * Wl)
            self.max_pad = max(self.max_pad, g.P_pad)
            self.max_tiles = max(self.max_tiles, g.n_tiles)
        self.d_jobs = _struct_array_to_device(jarr, dev)
        self.job_bytes = C.sizeof(nat.SpbFrameJob)

    def host_arena(self):
        """Pinned host buffer with the layout of ``stage``."""
        return torch.empty(self.offsets[-1], dtype=torch.uint8).pin_memory()

    def fill(self, arena, i, src_u8, trg_u8):
        """Write the (Hl,Wl,3) uint8 frames of problem ``i`` into ``arena`` (what a loader does when it decodes)."""
        so, to, Hl, Wl = self._views[i]
        n = 3 * Hl * Wl
        if not self.target_only:
            arena[so:so + n].copy_(src_u8.reshape(-1))
        arena[to:to + n].copy_(trg_u8.reshape(-1))

    def frame_bytes(self, first=0, count=None):
        """payload bytes (without alignment padding) of the frames of problems [first, first + count)"""
        count = self.n - first if count is None else count
        return sum((1 if self.target_only else 2) * 3 * v[2] * v[3] for v in self._views[first:first + count])

    def upload(self, arena, first=0, count=None):
        """ONE asynchronous host->device copy of the frames of problems [first, first + count) on the current stream."""
        count = self.n - first if count is None else count
        a, b = self.offsets[first], self.offsets[first + count]
        self.stage[a:b].copy_(arena[a:b], non_blocking=True)

    def run(self, d_geoms, first=0, count=None):
        """Ingest jobs [first, first + count) on the current stream (three launches)."""
        count = self.n - first if count is None else count
        nat.check(nat.lib().spb_ingest_u8(d_geoms.data_ptr(), self.d_jobs.data_ptr() + first * self.job_bytes, count,
                                          self.max_pixels, self.max_pad, self.max_tiles, _stream()), "spb_ingest_u8")
