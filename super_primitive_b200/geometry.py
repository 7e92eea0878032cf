"""Compact per-segment source geometry and its caches (host side of the geometry build).

The reference keeps every keyframe as dense ``(N,H,W)`` tensors and re-derives the point list
with ``torch.where`` on every cost evaluation (core/dense_optim.py:89-114).  Here the masks are
compacted ONCE per keyframe into a point list in the same (segment,row,col) order:

    uv      [n_pad] uint32   u | v<<16 | src_ok<<31
    logd    [n_pad] float32  raw per-segment log-depth at the point
    tiles   [n_tiles,4] int32  {segment, padded start, count, unpadded start}, <=128 points each
    seg_lkp [N] float32      log-depth at each segment's keypoint

Every segment's range is padded to a multiple of 4 points (16-byte aligned float/uint32 runs).
Caches are keyed on tensor identity + in-place version counters and hold weak references only,
so a caller replacing ``logdepth_perseg`` / ``keypoint_regions`` gets a fresh build.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict

import numpy as np
import torch

from . import _native as nat


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class CompactGeometry:
    """Device-resident compacted geometry of one source keyframe."""

    def __init__(self, regions, logd, keypoints, K):
        self._begin(regions, logd, keypoints, K)
        # one host sync per keyframe: three integers size the allocations (the reference syncs on torch.where every
        # iteration); the segment pointers stay on the device and are fetched lazily for the statistics path
        self._finish(*(int(v) for v in self._totals.tolist()))

    @classmethod
    def build_many(cls, keyframes):
        """Compact geometries of several keyframes with ONE host sync for all of them (a batch of independent frames,
        e.g. VOID depth completion): the count + scan passes of every keyframe are queued first, their totals are read
        back together, then every fill pass is queued.  `keyframes`: objects with keypoint_regions / get_logdepth() /
        keypoints / K."""
        pend = []
        for kf in keyframes:
            g = cls.__new__(cls)
            g._begin(kf.keypoint_regions, kf.get_logdepth(), kf.keypoints, kf.K)
            pend.append(g)
        if not pend:
            return []
        totals = torch.stack([g._totals for g in pend]).tolist()          # the one sync
        for g, (P, P_pad, T) in zip(pend, totals):
            g._finish(int(P), int(P_pad), int(T))
        return pend

    def _begin(self, regions, logd, keypoints, K):
        """passes 1 + 2 (row counts, scan): queued on the current stream, nothing read back"""
        if not regions.is_cuda:
            raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
        lib = nat.lib()
        dev = regions.device
        N, H, W = regions.shape
        if keypoints.shape[0] != N:
            raise AssertionError("one keypoint per segment expected")
        per_seg = logd.dim() == 3
        if per_seg and logd.shape[0] != N:
            raise AssertionError("logdepth_perseg and keypoint_regions disagree on N")
        masks = regions.detach()
        if masks.dtype != torch.bool:
            masks = masks != 0
        masks = masks.contiguous()
        m8 = masks.view(torch.uint8)
        logd_c = _f32c(logd)
        kps = _f32c(keypoints)
        self.K = _f32c(K).clone()
        st = _stream()
        i32 = dict(dtype=torch.int32, device=dev)
        row_cnt = torch.empty(N * H, **i32)
        row_off = torch.empty(N * H, **i32)
        csr = torch.empty(3 * (N + 1) + 3, **i32)          # seg_ptr | seg_ptr_pad | seg_tile | totals, one allocation
        seg_ptr, seg_ptr_pad, seg_tile = csr[:N + 1], csr[N + 1:2 * (N + 1)], csr[2 * (N + 1):3 * (N + 1)]
        self._totals = csr[3 * (N + 1):]
        nat.check(lib.spb_compact_count(m8.data_ptr(), N, H, W, row_cnt.data_ptr(), st), "spb_compact_count")
        nat.check(lib.spb_compact_scan(row_cnt.data_ptr(), N, H, row_off.data_ptr(), seg_ptr.data_ptr(),
                                       seg_ptr_pad.data_ptr(), seg_tile.data_ptr(), self._totals.data_ptr(), st),
                  "spb_compact_scan")
        self.N, self.H, self.W = N, H, W
        self._pending = (m8, logd_c, kps, per_seg, row_off, seg_ptr, seg_ptr_pad, seg_tile)
        self._csr = csr

    def _finish(self, P, P_pad, T):
        """pass 3 (ordered scatter) + tile table, once the point / tile counts are known on the host"""
        lib = nat.lib()
        m8, logd_c, kps, per_seg, row_off, seg_ptr, seg_ptr_pad, seg_tile = self._pending
        self._pending = None
        N, H, W = self.N, self.H, self.W
        dev = m8.device
        st = _stream()
        i32 = dict(dtype=torch.int32, device=dev)
        if P <= 0:
            raise AssertionError("keyframe has no segment pixels")
        self.P, self.P_pad = P, P_pad
        both = torch.zeros(2 * P_pad, dtype=torch.int32, device=dev)     # one allocation + one memset for uv | logd
        self.uv = both[:P_pad]                                           # bit pattern of uint32
        self.logd = both[P_pad:].view(torch.float32)
        self.seg_lkp = torch.empty(N, dtype=torch.float32, device=dev)
        self.kp_rc = torch.empty((N, 2), **i32)
        nat.check(lib.spb_compact_fill(m8.data_ptr(), logd_c.data_ptr(), H * W if per_seg else 0, kps.data_ptr(),
                                       N, H, W, row_off.data_ptr(), self.uv.data_ptr(), self.logd.data_ptr(),
                                       self.seg_lkp.data_ptr(), self.kp_rc.data_ptr(), st), "spb_compact_fill")
        # tile table on the device: tiles never straddle segments
        self.n_tiles = T
        self.tiles = torch.empty((T, 4), **i32)
        self.seg_tile = seg_tile
        nat.check(lib.spb_tile_table(seg_ptr.data_ptr(), seg_ptr_pad.data_ptr(), seg_tile.data_ptr(), N,
                                     self.tiles.data_ptr(), st), "spb_tile_table")
        self._finish_host_state(self._csr)

    def _finish_host_state(self, csr):
        """descriptor + caches, once every device array exists"""
        self._csr = csr
        self._pending = None
        self._seg_ptr_host = None
        self._pad_index = None
        self._seg_ids = None
        self.c = nat.SpbGeom(self.uv.data_ptr(), self.logd.data_ptr(), self.tiles.data_ptr(),
                             self.seg_tile.data_ptr(), self.seg_lkp.data_ptr(), self.K.data_ptr(),
                             self.P, self.P_pad, self.N, self.n_tiles, self.H, self.W)
        self.cref = C.byref(self.c)
        self._levels = OrderedDict()     # source-sample caches per (image identity)
        self._work = {}

    # ---- helpers -------------------------------------------------------------------------------
    @property
    def seg_ptr_host(self):
        if self._seg_ptr_host is None:
            h = self._csr[:2 * (self.N + 1)].cpu().numpy().astype(np.int64)
            self._seg_ptr_host = (h[:self.N + 1], h[self.N + 1:])
        return self._seg_ptr_host[0]

    @property
    def seg_ptr_pad_host(self):
        self.seg_ptr_host
        return self._seg_ptr_host[1]

    def pad_index(self):
        """(P,) int64: position of every unpadded point in the padded arrays."""
        if self._pad_index is None:
            cnt = self.seg_ptr_host[1:] - self.seg_ptr_host[:-1]
            seg = np.repeat(np.arange(self.N), cnt)
            idx = np.arange(self.P) + (self.seg_ptr_pad_host[:-1] - self.seg_ptr_host[:-1])[seg]
            self._pad_index = torch.from_numpy(idx).to(self.uv.device)
            self._seg_ids = torch.from_numpy(seg.astype(np.int64)).to(self.uv.device)
        return self._pad_index

    def seg_ids(self):
        self.pad_index()
        return self._seg_ids

    def level_buffers(self, image):
        """(src_rgb, tile_pack) for a source level image, cached per image tensor:
        src_rgb   [3][n_pad]  bilinear samples of the level image at every point's own pixel
        tile_pack [n_tiles][PACK_WORDS] uint32, the tile-major {header, uv, logd, r, g, b} blocks the fused
                  kernel streams with one bulk copy per tile"""
        key = (id(image), image._version, tuple(image.shape))
        hit = self._levels.get(key)
        if hit is not None and hit[0]() is image:
            self._levels.move_to_end(key)
            return hit[1], hit[2]
        img = _f32c(image[:3])
        lib = nat.lib()
        src_rgb = torch.empty((3, self.P_pad), dtype=torch.float32, device=img.device)
        nat.check(lib.spb_sample_source(self.cref, img.data_ptr(), img.shape[1], img.shape[2], src_rgb.data_ptr(),
                                        _stream()), "spb_sample_source")
        pack = torch.empty((self.n_tiles, nat.PACK_WORDS), dtype=torch.int32, device=img.device)
        nat.check(lib.spb_build_tile_pack(self.cref, src_rgb.data_ptr(), pack.data_ptr(), _stream()),
                  "spb_build_tile_pack")
        self._levels[key] = (weakref.ref(image), src_rgb, pack)
        while len(self._levels) > 8:
            self._levels.popitem(last=False)
        return src_rgb, pack

    def source_samples(self, image):
        return self.level_buffers(image)[0]

    def workspace(self, n_pairs):
        """Partial-sum scratch of the fused kernel for ``n_pairs`` pairs over this geometry (stream-ordered reuse)."""
        ws = self._work.get(n_pairs)
        if ws is None:
            ws = torch.empty(nat.lib().spb_workspace_floats(self.cref, n_pairs, 0), dtype=torch.float32,
                             device=self.uv.device)
            self._work[n_pairs] = ws
        return ws

    def bytes(self):
        return self.P_pad * 8 + self.n_tiles * 16 + self.N * 8


# ---------------------------------------------------------------------------------------------------
_GEOM_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_RGBA_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
GEOM_CACHE_SIZE = 16
RGBA_CACHE_SIZE = 32


def geometry_of(kf) -> CompactGeometry:
    """Compact geometry of a keyframe, cached on the identity/version of its geometry tensors (masks, log-depth,
    keypoints).  The intrinsics are a 9-float VALUE of the call, not part of the key: `keyframe_pyramid` gives every
    level its own `K.clone()` (image/keyframe.py:125-146), and with `geo_down=False` -- every caller -- all levels share
    one geometry.  When a hit arrives with a different K tensor its values are copied into the geometry's own device
    buffer (stream-ordered, no host sync), so the three compaction passes and their one host sync run once per keyframe."""
    own = getattr(kf, "_spb_geometry", None)        # handover.CompactKeyFrame: the geometry IS the keyframe
    if own is not None:
        return own
    reg, ld, kp, K = kf.keypoint_regions, kf.get_logdepth(), kf.keypoints, kf.K
    key = (id(reg), reg._version, id(ld), ld._version, id(kp), kp._version)
    hit = _GEOM_CACHE.get(key)
    if hit is not None and hit[0]() is reg and hit[1]() is ld and hit[2]() is kp:
        _GEOM_CACHE.move_to_end(key)
        g = hit[3]
        kref, kver = g._K_src
        if kref() is not K or kver != K._version:
            g.K.copy_(_f32c(K).reshape(g.K.shape), non_blocking=True)
            g._K_src = (weakref.ref(K), K._version)
        return g
    g = CompactGeometry(reg, ld, kp, K)
    g._K_src = (weakref.ref(K), K._version)
    _GEOM_CACHE[key] = (weakref.ref(reg), weakref.ref(ld), weakref.ref(kp), g)
    while len(_GEOM_CACHE) > GEOM_CACHE_SIZE:
        _GEOM_CACHE.popitem(last=False)
    return g


def geometries_of(kfs):
    """`geometry_of` for a batch of keyframes: cache hits are reused, all misses are built with one host sync
    (`CompactGeometry.build_many`)."""
    out = [None] * len(kfs)
    miss = []
    for i, kf in enumerate(kfs):
        if getattr(kf, "_spb_geometry", None) is not None:
            out[i] = kf._spb_geometry
            continue
        reg, ld, kp = kf.keypoint_regions, kf.get_logdepth(), kf.keypoints
        key = (id(reg), reg._version, id(ld), ld._version, id(kp), kp._version)
        hit = _GEOM_CACHE.get(key)
        if hit is not None and hit[0]() is reg and hit[1]() is ld and hit[2]() is kp:
            out[i] = geometry_of(kf)
        else:
            miss.append((i, key))
    built = CompactGeometry.build_many([kfs[i] for i, _ in miss])
    for (i, key), g in zip(miss, built):
        kf = kfs[i]
        g._K_src = (weakref.ref(kf.K), kf.K._version)
        _GEOM_CACHE[key] = (weakref.ref(kf.keypoint_regions), weakref.ref(kf.get_logdepth()), weakref.ref(kf.keypoints), g)
        out[i] = g
    while len(_GEOM_CACHE) > max(GEOM_CACHE_SIZE, len(kfs)):
        _GEOM_CACHE.popitem(last=False)
    return out


def pack_rgba(images):
    """(3,Hl,Wl) or (B,>=3,Hl,Wl) planar float -> (B,Hl,Wl,4) RGBA-interleaved, cached per tensor."""
    key = (id(images), images._version, tuple(images.shape))
    hit = _RGBA_CACHE.get(key)
    if hit is not None and hit[0]() is images:
        _RGBA_CACHE.move_to_end(key)
        return hit[1]
    src = images.detach()
    if src.dim() == 3:
        src = src[None]
    if src.dtype != torch.float32:
        src = src.float()
    B, Cn, Hl, Wl = src.shape
    if Cn < 3:
        raise AssertionError("colour mode needs >= 3 image channels")
    if src.stride(3) != 1 or src.stride(2) != Wl or src.stride(1) != Hl * Wl:
        src = src.contiguous()
    out = torch.empty((B, Hl, Wl, 4), dtype=torch.float32, device=src.device)
    nat.check(nat.lib().spb_pack_rgba(src.data_ptr(), src.stride(0), B, Hl, Wl, out.data_ptr(), _stream()),
              "spb_pack_rgba")
    _RGBA_CACHE[key] = (weakref.ref(images), out)
    while len(_RGBA_CACHE) > RGBA_CACHE_SIZE:
        _RGBA_CACHE.popitem(last=False)
    return out


def clear_caches():
    _GEOM_CACHE.clear()
    _RGBA_CACHE.clear()
