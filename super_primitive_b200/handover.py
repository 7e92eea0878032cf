"""Keyframe hand-over from the frontend without the dense (N,H,W) tensors -- SURVEY.md section 8(f) rank 3, second half.

The reference's frontend ends with (frontend/process_frame.py:231-244)

    logdepth = F.interpolate(integrated_depth[:, None], size=(H_kf, W_kf), mode='nearest')[:, 0]
    masks = logdepth > 1e-7
    keypoints, masks, logdepth = put_keypoints_back(keypoints, masks, logdepth)      # image/keyframe.py:151-173
    logdepth[masks] = torch.log(logdepth[masks])
    kf = KeyFrame(image, K=K_kf, logdepth_perseg=logdepth, keypoints=keypoints, keypoint_regions=masks)

i.e. three dense (N,H,W) tensors (0.8 GB + 0.2 GB per keyframe at 1024x768x256) that the alignment path compacts again
on first use.  `keyframe_from_frontend` produces the compact geometry straight from `integrated_depth` (the same nearest
resampling, threshold, keypoint snap and logarithm, fused into the two compaction passes) and returns a `CompactKeyFrame`
that every entry point of this package accepts in place of a `KeyFrame`; the dense tensors are only materialised if
somebody reads `keypoint_regions` / `logdepth_perseg` (the GUI, `unproject_kf_to_depths`).
"""
from __future__ import annotations

import torch

from . import _native as nat
from .geometry import CompactGeometry, _f32c, _stream

MASK_THRESHOLD = 1e-7        # frontend/process_frame.py:234


def _nearest_maps(Hf, Wf, H, W, device):
    """source row / column of every keyframe row / column under F.interpolate(mode='nearest'): obtained from the operator
    itself on an index ramp, so whatever rounding rule the installed torch uses is reproduced exactly"""
    rows = torch.nn.functional.interpolate(torch.arange(Hf, dtype=torch.float32, device=device).view(1, 1, Hf, 1),
                                           size=(H, 1), mode='nearest').view(H)
    cols = torch.nn.functional.interpolate(torch.arange(Wf, dtype=torch.float32, device=device).view(1, 1, 1, Wf),
                                           size=(1, W), mode='nearest').view(W)
    return rows.to(torch.int32).contiguous(), cols.to(torch.int32).contiguous()


def geometry_from_frontend(integrated_depth, keypoints, K, size):
    """CompactGeometry + snapped keypoints (M,2) + `good` (N,) bool (segments that survive: the reference drops segments
    whose resampled mask is empty) from the frontend's `integrated_depth` (N,Hf,Wf)."""
    if not integrated_depth.is_cuda:
        raise RuntimeError("super_primitive_b200 runs on CUDA tensors only (no CPU fallback)")
    lib = nat.lib()
    dev = integrated_depth.device
    H, W = int(size[0]), int(size[1])
    depth = _f32c(integrated_depth)
    kps = _f32c(keypoints)
    N, Hf, Wf = depth.shape
    if kps.shape[0] != N:
        raise AssertionError("one keypoint per segment expected")
    row_map, col_map = _nearest_maps(Hf, Wf, H, W, dev)
    st = _stream()
    i32 = dict(dtype=torch.int32, device=dev)
    good = torch.ones(N, dtype=torch.bool, device=dev)
    while True:
        row_cnt = torch.empty(N * H, **i32)
        row_off = torch.empty(N * H, **i32)
        csr = torch.empty(3 * (N + 1) + 3, **i32)
        seg_ptr, seg_ptr_pad, seg_tile = csr[:N + 1], csr[N + 1:2 * (N + 1)], csr[2 * (N + 1):3 * (N + 1)]
        totals = csr[3 * (N + 1):]
        nat.check(lib.spb_compact_count_depth(depth.data_ptr(), N, Hf, Wf, row_map.data_ptr(), col_map.data_ptr(), H, W,
                                              MASK_THRESHOLD, row_cnt.data_ptr(), st), "spb_compact_count_depth")
        nat.check(lib.spb_compact_scan(row_cnt.data_ptr(), N, H, row_off.data_ptr(), seg_ptr.data_ptr(),
                                       seg_ptr_pad.data_ptr(), seg_tile.data_ptr(), totals.data_ptr(), st),
                  "spb_compact_scan")
        host = torch.cat([totals, seg_ptr]).tolist()               # the one host sync: sizes + per-segment counts
        P, P_pad, T = host[:3]
        cnt = torch.tensor(host[3:], dtype=torch.int64).diff()
        if bool((cnt > 0).all()):
            break
        # put_keypoints_back drops the segments without a pixel (image/keyframe.py:156-161): rare, redo without them
        keep = (cnt > 0).to(dev)
        idx = torch.nonzero(good).flatten()
        good[idx[~keep]] = False
        depth, kps = depth[keep].contiguous(), kps[keep].contiguous()
        N = int(keep.sum())
        if N == 0:
            raise AssertionError("keyframe has no segment pixels")
    g = CompactGeometry.__new__(CompactGeometry)
    g.K = _f32c(K).clone()
    g.N, g.H, g.W, g.P, g.P_pad = N, H, W, int(P), int(P_pad)
    g.uv = torch.zeros(g.P_pad, dtype=torch.int32, device=dev)
    g.logd = torch.zeros(g.P_pad, dtype=torch.float32, device=dev)
    g.seg_lkp = torch.empty(N, dtype=torch.float32, device=dev)
    g.kp_rc = torch.empty((N, 2), **i32)
    kp_norm = torch.empty((N, 2), dtype=torch.float32, device=dev)
    nat.check(lib.spb_compact_fill_depth(depth.data_ptr(), N, Hf, Wf, row_map.data_ptr(), col_map.data_ptr(), H, W,
                                         MASK_THRESHOLD, row_off.data_ptr(), seg_ptr.data_ptr(), seg_ptr_pad.data_ptr(),
                                         kps.data_ptr(), g.uv.data_ptr(), g.logd.data_ptr(), g.seg_lkp.data_ptr(),
                                         g.kp_rc.data_ptr(), kp_norm.data_ptr(), st), "spb_compact_fill_depth")
    g.n_tiles = int(T)
    g.tiles = torch.empty((g.n_tiles, 4), **i32)
    g.seg_tile = seg_tile
    nat.check(lib.spb_tile_table(seg_ptr.data_ptr(), seg_ptr_pad.data_ptr(), seg_tile.data_ptr(), N, g.tiles.data_ptr(), st),
              "spb_tile_table")
    g._finish_host_state(csr)
    return g, kp_norm, good


class CompactKeyFrame:
    """A keyframe whose geometry exists only in compact form.  Duck-types the reference's `KeyFrame`
    (image/keyframe.py:20-65) for the alignment path; `keypoint_regions` / `logdepth_perseg` are scattered back into dense
    (N,H,W) tensors on first access only."""

    def __init__(self, image, K, geometry, keypoints, K_img=None, id=None):
        self.image, self.K, self.keypoints, self.id = image, K, keypoints, id
        self.K_img = K if K_img is None else K_img
        self._spb_geometry = geometry
        self._dense = None

    def geo_spatial_dim(self):
        return (self._spb_geometry.H, self._spb_geometry.W)

    def num_segments(self):
        return self._spb_geometry.N

    def is_supporting(self):
        return False

    def _materialise(self):
        if self._dense is None:
            g = self._spb_geometry
            idx = g.pad_index()
            uv = g.uv[idx]
            u, v = (uv & 0xffff).long(), ((uv >> 16) & 0x7fff).long()
            seg = g.seg_ids()
            masks = torch.zeros((g.N, g.H, g.W), dtype=torch.bool, device=uv.device)
            logd = torch.zeros((g.N, g.H, g.W), dtype=torch.float32, device=uv.device)
            masks[seg, v, u] = True
            logd[seg, v, u] = g.logd[idx]
            self._dense = (masks, logd)
        return self._dense

    @property
    def keypoint_regions(self):
        return self._materialise()[0]

    @property
    def logdepth_perseg(self):
        return self._materialise()[1]

    def get_logdepth(self):
        return self.logdepth_perseg

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("a CompactKeyFrame lives on the device its geometry was built on")
        return self


def keyframe_from_frontend(image, K_kf, integrated_depth, keypoints, size=None, K_img=None):
    """What `FrontProcessorNew.process_to_kf` builds from `preprocessed['integrated_depth']` and
    `preprocessed['keypoints']` (frontend/process_frame.py:231-244), as a `CompactKeyFrame`.  `size` = (H_kf, W_kf),
    default: the image's."""
    if size is None:
        size = tuple(image.shape[-2:])
    g, kp_norm, _ = geometry_from_frontend(integrated_depth, keypoints, K_kf, size)
    return CompactKeyFrame(image, K_kf, g, kp_norm, K_img=K_img)
