// Geometry-side kernels: mask compaction, image packing, source sampling, dense depth expand,
// point lifting, depth splat.  All once-per-keyframe / once-per-frame work (not per iteration),
// all HBM-streaming; grids are sized from the data, loads are coalesced along image rows.
#include "spb_common.cuh"
#include "spb_frame_stages.cuh"

// sums that several threads add into one pixel are kept in 32.32 fixed point and added with 64-bit integer atomics:
// independent of the order of arrival (bit-reproducible), exact up to 2^-32, rounded to float32 once
#define SPB_AVG_FX 4294967296.0   /* 2^32 */
__device__ __forceinline__ unsigned long long to_fx(float z) { return (unsigned long long)((double)z * SPB_AVG_FX + 0.5); }

// ------------------------------------------------------------------------------------------------
// compaction pass 1: one warp per (segment,row): number of mask pixels in the row.
// The masks are N*H*W bytes (the largest thing the build reads: 31 MB at 640x480x100, 201 MB at 1024x768x256) and are
// almost everywhere zero (a segment covers ~1/N of the image), so both passes over them read 16 mask bytes per lane
// and per load when the rows are 16-byte aligned (W % 16 == 0: every shape the callers produce).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nonzero_bytes(uint32_t w) { return __popc(__vcmpne4(w, 0u) & 0x01010101u); }
// 4-bit mask of the non-zero bytes of a word (bit j = byte j)
__device__ __forceinline__ uint32_t nonzero_mask4(uint32_t w) {
    return (((__vcmpne4(w, 0u) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
}

template <bool VEC>
__global__ void k_row_count(const uint8_t* __restrict__ masks, int rows, int W, int32_t* __restrict__ row_cnt) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint8_t* m = masks + (size_t)row * W;
    int n = 0;
    if constexpr (VEC) {
        const uint4* m16 = reinterpret_cast<const uint4*>(m);
        for (int i = lane; i < (W >> 4); i += 32) {
            const uint4 v = __ldg(m16 + i);
            n += nonzero_bytes(v.x) + nonzero_bytes(v.y) + nonzero_bytes(v.z) + nonzero_bytes(v.w);
        }
    } else {
        for (int x = lane; x < W; x += 32) n += (m[x] != 0);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) row_cnt[row] = n;
}

// pass 2: one CTA; a warp per segment sums its H row counts (coalesced), one thread lays the segments out with their
// padding to SPB_PAD, then a warp per segment turns the row counts into row offsets with a shuffle scan
__global__ void __launch_bounds__(1024)
k_row_scan(const int32_t* __restrict__ row_cnt, int N, int H, int32_t* __restrict__ row_off,
           int32_t* __restrict__ seg_ptr, int32_t* __restrict__ seg_ptr_pad, int32_t* __restrict__ seg_tile,
           int32_t* __restrict__ totals) {
    extern __shared__ int32_t s_seg[];       // [N] counts -> padded starts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int b = warp; b < N; b += nwarps) {
        const int32_t* r = row_cnt + (size_t)b * H;
        int n = 0;
        for (int y = lane; y < H; y += 32) n += r[y];
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) s_seg[b] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, run_pad = 0, run_tile = 0;
        for (int b = 0; b < N; ++b) {
            const int n = s_seg[b];
            seg_ptr[b] = run;
            seg_ptr_pad[b] = run_pad;
            if (seg_tile) seg_tile[b] = run_tile;
            s_seg[b] = run_pad;
            run += n;
            run_pad += (n + SPB_PAD - 1) / SPB_PAD * SPB_PAD;
            run_tile += (n + SPB_TILE - 1) / SPB_TILE;             // tiles never straddle segments
        }
        seg_ptr[N] = run;
        seg_ptr_pad[N] = run_pad;
        if (seg_tile) seg_tile[N] = run_tile;
        totals[0] = run;
        totals[1] = run_pad;
        totals[2] = run_tile;
    }
    __syncthreads();
    for (int b = warp; b < N; b += nwarps) {
        int off = s_seg[b];
        const int32_t* r = row_cnt + (size_t)b * H;
        int32_t* o = row_off + (size_t)b * H;
        for (int y0 = 0; y0 < H; y0 += 32) {
            const int y = y0 + lane;
            const int c = y < H ? r[y] : 0;
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (y < H) o[y] = off + incl - c;
            off += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// pass 3: ordered scatter, one warp per (segment,row)
template <bool VEC>
__global__ void k_row_fill(const uint8_t* __restrict__ masks, const float* __restrict__ logd, int64_t seg_stride,
                           int N, int H, int W, const int32_t* __restrict__ row_off, uint32_t* __restrict__ uv,
                           float* __restrict__ L) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N * H) return;
    const int b = row / H, y = row - b * H;
    const uint8_t* m = masks + (size_t)row * W;
    const float* lrow = logd + (size_t)b * seg_stride + (size_t)y * W;
    // static source validity of the point's own pixel (core/dense_optim.py:128-130 on the
    // re-projected source point): |2 u inv - 1| <= 0.99 on both axes
    const float tiw = 2.0f * (1.0f / (float)(W - 1));
    const float tih = 2.0f * (1.0f / (float)(H - 1));
    const bool yok = fabsf(fmaf((float)y, tih, -1.0f)) <= 0.99f;
    int off = row_off[row];
    if constexpr (VEC) {
        const uint4* m16 = reinterpret_cast<const uint4*>(m);
        const int W16 = W >> 4;
        for (int i0 = 0; i0 < W16; i0 += 32) {
            const int i = i0 + lane;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (i < W16) v = __ldg(m16 + i);
            uint32_t bits = nonzero_mask4(v.x) | (nonzero_mask4(v.y) << 4) | (nonzero_mask4(v.z) << 8) |
                            (nonzero_mask4(v.w) << 12);
            const int c = __popc(bits);
            if (!__any_sync(0xffffffffu, c != 0)) continue;          // 512 empty pixels: the common case
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int dst = off + incl - c;
            off += __shfl_sync(0xffffffffu, incl, 31);
            while (bits) {
                const int j = __ffs((int)bits) - 1;
                bits &= bits - 1u;
                const int x = (i << 4) + j;
                const bool ok = yok && (fabsf(fmaf((float)x, tiw, -1.0f)) <= 0.99f);
                uv[dst] = (uint32_t)x | ((uint32_t)y << 16) | (ok ? 0x80000000u : 0u);
                L[dst] = lrow[x];
                ++dst;
            }
        }
    } else {
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const bool on = (x < W) && (m[x] != 0);
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (on) {
                const int dst = off + __popc(bal & ((1u << lane) - 1u));
                const bool ok = yok && (fabsf(fmaf((float)x, tiw, -1.0f)) <= 0.99f);
                uv[dst] = (uint32_t)x | ((uint32_t)y << 16) | (ok ? 0x80000000u : 0u);
                L[dst] = lrow[x];
            }
            off += __popc(bal);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Compaction straight from the frontend's output (second half).  The frontend hands over
// `integrated_depth` (N,Hf,Wf): per-segment depth from the normal integration, 0 outside the segment, at ITS resolution;
// the reference resamples it to the keyframe grid with nearest-neighbour interpolation, thresholds it into the masks,
// snaps every keypoint to the nearest mask pixel and takes the logarithm -- three dense (N,H,W) tensors
// (frontend/process_frame.py:231-236, image/keyframe.py:151-173).  Here the same result is written directly as the
// compact point list: the index maps of the nearest resampling (row_map[H], col_map[W], produced by the same torch
// operator on an index ramp) select the source texel, mask = depth > thr, L = log(depth).
// ------------------------------------------------------------------------------------------------
__global__ void k_row_count_depth(const float* __restrict__ depth, int Hf, int Wf, const int32_t* __restrict__ row_map,
                                  const int32_t* __restrict__ col_map, int rows, int H, int W, float thr,
                                  int32_t* __restrict__ row_cnt) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = row / H, y = row - b * H;
    const float* src = depth + ((size_t)b * Hf + row_map[y]) * Wf;
    int n = 0;
    for (int x = lane; x < W; x += 32) n += (src[col_map[x]] > thr);
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) row_cnt[row] = n;
}

__global__ void k_row_fill_depth(const float* __restrict__ depth, int Hf, int Wf, const int32_t* __restrict__ row_map,
                                 const int32_t* __restrict__ col_map, int N, int H, int W, float thr,
                                 const int32_t* __restrict__ row_off, uint32_t* __restrict__ uv, float* __restrict__ L) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N * H) return;
    const int b = row / H, y = row - b * H;
    const float* src = depth + ((size_t)b * Hf + row_map[y]) * Wf;
    const float tiw = 2.0f * (1.0f / (float)(W - 1));
    const float tih = 2.0f * (1.0f / (float)(H - 1));
    const bool yok = fabsf(fmaf((float)y, tih, -1.0f)) <= 0.99f;
    int off = row_off[row];
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const float d = (x < W) ? src[col_map[x]] : 0.0f;
        const bool on = (x < W) && (d > thr);
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int dst = off + __popc(bal & ((1u << lane) - 1u));
            const bool ok = yok && (fabsf(fmaf((float)x, tiw, -1.0f)) <= 0.99f);
            uv[dst] = (uint32_t)x | ((uint32_t)y << 16) | (ok ? 0x80000000u : 0u);
            L[dst] = logf(d);                                  // logdepth[masks] = torch.log(logdepth[masks])
        }
        off += __popc(bal);
    }
}

// put_keypoints_back (image/keyframe.py:151-173): every keypoint moves to the mask pixel nearest to its rounded pixel
// position (Euclidean; the first pixel in (row, col) order among equals = argmin's choice).  One CTA per segment over its
// compact points: minimum of (squared distance << 32 | point index).
__global__ void k_snap_keypoints(const uint32_t* __restrict__ uv, const float* __restrict__ L,
                                 const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_ptr_pad,
                                 const float* __restrict__ keypoints, int H, int W, float* __restrict__ seg_lkp,
                                 int32_t* __restrict__ kp_rc, float* __restrict__ kp_norm) {
    const int b = blockIdx.x;
    const int cnt = seg_ptr[b + 1] - seg_ptr[b], start = seg_ptr_pad[b];
    const float hr = 0.5f * ((float)H - 1.0f), hc = 0.5f * ((float)W - 1.0f);
    const long long r0 = (long long)rintf(hr * (keypoints[2 * b] + 1.0f));          // tool/point_utils.py:37-40
    const long long c0 = (long long)rintf(hc * (keypoints[2 * b + 1] + 1.0f));
    unsigned long long best = ~0ull;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t w = uv[start + i];
        const long long dr = (long long)((w >> 16) & 0x7fffu) - r0, dc = (long long)(w & 0xffffu) - c0;
        const unsigned long long key = ((unsigned long long)(dr * dr + dc * dc) << 32) | (unsigned)i;
        best = key < best ? key : best;
    }
    __shared__ unsigned long long s_best[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2) best = s_best[w2] < best ? s_best[w2] : best;
        const int i = (int)(best & 0xffffffffull);
        const uint32_t w = uv[start + i];
        const int r = (int)((w >> 16) & 0x7fffu), cc = (int)(w & 0xffffu);
        kp_rc[2 * b] = r;
        kp_rc[2 * b + 1] = cc;
        seg_lkp[b] = L[start + i];
        // normalise_coordinates (tool/point_utils.py:31-35): 2 x fl32(1 / (dims - 1)) - 1
        kp_norm[2 * b] = __fadd_rn(__fmul_rn((float)(2 * r), __fdiv_rn(1.0f, (float)H - 1.0f)), -1.0f);
        kp_norm[2 * b + 1] = __fadd_rn(__fmul_rn((float)(2 * cc), __fdiv_rn(1.0f, (float)W - 1.0f)), -1.0f);
    }
}

extern "C" int spb_compact_count_depth(const float* depth, int N, int Hf, int Wf, const int32_t* row_map,
                                       const int32_t* col_map, int H, int W, float thr, int32_t* row_cnt, void* stream) {
    if (!depth || !row_map || !col_map || !row_cnt || N < 1 || Hf < 1 || Wf < 1 || H < 2 || W < 2 || H > 32767 || W > 65535)
        return SPB_EINVAL;
    const int rows = N * H;
    k_row_count_depth<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(depth, Hf, Wf, row_map, col_map, rows, H, W, thr,
                                                                        row_cnt);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_compact_fill_depth(const float* depth, int N, int Hf, int Wf, const int32_t* row_map,
                                      const int32_t* col_map, int H, int W, float thr, const int32_t* row_off,
                                      const int32_t* seg_ptr, const int32_t* seg_ptr_pad, const float* keypoints,
                                      uint32_t* uv, float* L, float* seg_lkp, int32_t* kp_rc, float* kp_norm,
                                      void* stream) {
    if (!depth || !row_map || !col_map || !row_off || !seg_ptr || !seg_ptr_pad || !keypoints || !uv || !L || !seg_lkp ||
        !kp_rc || !kp_norm || N < 1)
        return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = N * H;
    k_row_fill_depth<<<(rows + 7) / 8, 256, 0, st>>>(depth, Hf, Wf, row_map, col_map, N, H, W, thr, row_off, uv, L);
    SPB_CHECK_LAUNCH();
    k_snap_keypoints<<<N, 256, 0, st>>>(uv, L, seg_ptr, seg_ptr_pad, keypoints, H, W, seg_lkp, kp_rc, kp_norm);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// keypoint pixel (round half even) and log-depth at the keypoint, core/dense_optim.py:51-64
__global__ void k_keypoints(const float* __restrict__ keypoints, const float* __restrict__ logd, int64_t seg_stride,
                            int N, int H, int W, float* __restrict__ seg_lkp, int32_t* __restrict__ kp_rc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    // 0.5 * (dims - 1) * (x_norm + 1), round, long   (tool/point_utils.py:37-40)
    const float hr = 0.5f * ((float)H - 1.0f), hc = 0.5f * ((float)W - 1.0f);
    int r = (int)rintf(hr * (keypoints[2 * b] + 1.0f));
    int cc = (int)rintf(hc * (keypoints[2 * b + 1] + 1.0f));
    // python negative indices wrap; anything else out of range is a caller bug -> clamp defensively
    if (r < 0) r += H;
    if (cc < 0) cc += W;
    r = min(max(r, 0), H - 1);
    cc = min(max(cc, 0), W - 1);
    kp_rc[2 * b] = r;
    kp_rc[2 * b + 1] = cc;
    seg_lkp[b] = logd[(size_t)b * seg_stride + (size_t)r * W + cc];
}

extern "C" int spb_compact_count(const uint8_t* masks, int N, int H, int W, int32_t* row_cnt, void* stream) {
    if (!masks || !row_cnt || N < 1 || H < 2 || W < 2 || H > 32767 || W > 65535) return SPB_EINVAL;
    const int rows = N * H;
    if (W % 16 == 0 && (reinterpret_cast<uintptr_t>(masks) & 15u) == 0)
        k_row_count<true><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(masks, rows, W, row_cnt);
    else
        k_row_count<false><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(masks, rows, W, row_cnt);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_compact_scan(const int32_t* row_cnt, int N, int H, int32_t* row_off, int32_t* seg_ptr,
                                int32_t* seg_ptr_pad, int32_t* seg_tile, int32_t* totals, void* stream) {
    if (!row_cnt || !row_off || !seg_ptr || !seg_ptr_pad || !totals || N < 1 || H < 1) return SPB_EINVAL;
    if ((size_t)N * sizeof(int32_t) > 48 * 1024) return SPB_ELIMIT;
    k_row_scan<<<1, 1024, N * sizeof(int32_t), (cudaStream_t)stream>>>(row_cnt, N, H, row_off, seg_ptr, seg_ptr_pad,
                                                                       seg_tile, totals);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// tile table: one CTA per segment writes {segment, padded start, count, unpadded start} of its tiles
__global__ void k_tile_table(const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_ptr_pad,
                             const int32_t* __restrict__ seg_tile, int32_t* __restrict__ tiles) {
    const int b = blockIdx.x;
    const int cnt = seg_ptr[b + 1] - seg_ptr[b], t0 = seg_tile[b], nt = seg_tile[b + 1] - t0;
    int4* out = reinterpret_cast<int4*>(tiles) + t0;
    for (int j = threadIdx.x; j < nt; j += blockDim.x)
        out[j] = make_int4(b, seg_ptr_pad[b] + j * SPB_TILE, min(cnt - j * SPB_TILE, SPB_TILE), seg_ptr[b] + j * SPB_TILE);
}

extern "C" int spb_tile_table(const int32_t* seg_ptr, const int32_t* seg_ptr_pad, const int32_t* seg_tile, int N,
                              int32_t* tiles, void* stream) {
    if (!seg_ptr || !seg_ptr_pad || !seg_tile || !tiles || N < 1) return SPB_EINVAL;
    k_tile_table<<<N, 128, 0, (cudaStream_t)stream>>>(seg_ptr, seg_ptr_pad, seg_tile, tiles);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_compact_fill(const uint8_t* masks, const float* logd, int64_t logd_seg_stride,
                                const float* keypoints, int N, int H, int W, const int32_t* row_off, uint32_t* uv,
                                float* L, float* seg_lkp, int32_t* kp_rc, void* stream) {
    if (!masks || !logd || !keypoints || !row_off || !uv || !L || !seg_lkp || !kp_rc) return SPB_EINVAL;
    const int rows = N * H;
    cudaStream_t st = (cudaStream_t)stream;
    if (W % 16 == 0 && (reinterpret_cast<uintptr_t>(masks) & 15u) == 0)
        k_row_fill<true><<<(rows + 7) / 8, 256, 0, st>>>(masks, logd, logd_seg_stride, N, H, W, row_off, uv, L);
    else
        k_row_fill<false><<<(rows + 7) / 8, 256, 0, st>>>(masks, logd, logd_seg_stride, N, H, W, row_off, uv, L);
    SPB_CHECK_LAUNCH();
    k_keypoints<<<(N + 127) / 128, 128, 0, st>>>(keypoints, logd, logd_seg_stride, N, H, W, seg_lkp, kp_rc);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// planar (3,H,W) -> RGBA interleaved
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_rgba(const float* __restrict__ planar, int64_t img_stride, int HW, float4* __restrict__ rgba) {
    const int img = blockIdx.y;
    const float* p = planar + (size_t)img * img_stride;
    float4* o = rgba + (size_t)img * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
        o[i] = make_float4(p[i], p[(size_t)HW + i], p[2 * (size_t)HW + i], 0.f);
}

extern "C" int spb_pack_rgba(const float* planar, int64_t img_stride, int n_img, int Hl, int Wl, float* rgba,
                             void* stream) {
    if (!planar || !rgba || n_img < 1 || Hl < 1 || Wl < 1) return SPB_EINVAL;
    const int HW = Hl * Wl;
    int bx = (HW + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_pack_rgba<<<dim3(bx, n_img), 256, 0, (cudaStream_t)stream>>>(planar, img_stride, HW,
                                                                  reinterpret_cast<float4*>(rgba));
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// cached source samples at a pyramid level: bilinear of the source level image at the point's own
// pixel scaled to the level (reference re-projects the unprojected point, which returns its own
// pixel up to float rounding; core/dense_optim.py:315-317)
// ------------------------------------------------------------------------------------------------
__global__ void k_sample_source(const __grid_constant__ SpbGeom g, const float* __restrict__ img, int Hl, int Wl,
                                float* __restrict__ out) {
    const SourceSampleScale sc = source_sample_scale(g, Hl, Wl);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pad; p += gridDim.x * blockDim.x)
        sample_source_point(g, img, Hl, Wl, sc, p, out);
}

extern "C" int spb_sample_source(const SpbGeom* geom, const float* src_planar, int Hl, int Wl, float* out,
                                 void* stream) {
    if (!geom || !src_planar || !out || Hl < 1 || Wl < 1 || geom->n_pad < 1) return SPB_EINVAL;
    int bx = (geom->n_pad + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_sample_source<<<bx, 256, 0, (cudaStream_t)stream>>>(*geom, src_planar, Hl, Wl, out);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// tile-major level buffer: one warp per tile copies header + the five 128-word arrays (coalesced)
__global__ void k_build_tile_pack(const __grid_constant__ SpbGeom g, const float* __restrict__ rgb,
                                  uint32_t* __restrict__ pack) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= g.n_tiles) return;
    build_tile_pack_tile(g, rgb, pack, t, lane);
}

extern "C" int spb_build_tile_pack(const SpbGeom* geom, const float* src_rgb, uint32_t* pack, void* stream) {
    if (!geom || !src_rgb || !pack || geom->n_tiles < 1) return SPB_EINVAL;
    k_build_tile_pack<<<(geom->n_tiles + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*geom, src_rgb, pack);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// unproject_kf_to_depths: dense exp((L + shift_b) * mask)
// ------------------------------------------------------------------------------------------------
__global__ void k_dense_depths(const uint8_t* __restrict__ masks, const float* __restrict__ logd, int64_t seg_stride,
                               const float* __restrict__ seg_lkp, const float* __restrict__ k, int HW,
                               float* __restrict__ out) {
    const int b = blockIdx.y;
    const float shift = k[b] - seg_lkp[b];
    const uint8_t* m = masks + (size_t)b * HW;
    const float* l = logd + (size_t)b * seg_stride;
    float* o = out + (size_t)b * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
        o[i] = expf((l[i] + shift) * (m[i] ? 1.0f : 0.0f));
}

extern "C" int spb_dense_depths(const uint8_t* masks, const float* logd, int64_t logd_seg_stride,
                                const float* seg_lkp, const float* k, int N, int H, int W, float* out, void* stream) {
    if (!masks || !logd || !seg_lkp || !k || !out || N < 1 || N > 65535) return SPB_EINVAL;
    const int HW = H * W;
    int bx = (HW + 255) / 256;
    if (bx > 1024) bx = 1024;
    k_dense_depths<<<dim3(bx, N), 256, 0, (cudaStream_t)stream>>>(masks, logd, logd_seg_stride, seg_lkp, k, HW, out);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// lifted points (unproject_kf) and the depth splat (estimate_depth_kf_native)
// ------------------------------------------------------------------------------------------------
template <bool SPLAT>
__global__ void k_lift(const __grid_constant__ SpbGeom g, const float* __restrict__ k, const float* __restrict__ pose,
                       float* __restrict__ src_pts, int64_t* __restrict__ seg_ids, uint8_t* __restrict__ src_ok,
                       int mean, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ sum) {
    const int lane = threadIdx.x & 31;
    const int wglobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int wstride = gridDim.x * (blockDim.x >> 5);
    const float ifx = 1.0f / g.K[0], ify = 1.0f / g.K[4], cx = g.K[2], cy = g.K[5];
    const float fx = g.K[0], fy = g.K[4];
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tt[3] = {0, 0, 0};
    if (SPLAT && pose != nullptr) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            R[3 * i] = pose[4 * i]; R[3 * i + 1] = pose[4 * i + 1]; R[3 * i + 2] = pose[4 * i + 2];
            tt[i] = pose[4 * i + 3];
        }
    }
    const int4* tiles = reinterpret_cast<const int4*>(g.tiles);
    for (int t = wglobal; t < g.n_tiles; t += wstride) {
        const int4 td = tiles[t];
        const float shift = k[td.x] - g.seg_lkp[td.x];
        for (int i = lane; i < td.z; i += 32) {
            const int p = td.y + i, q = td.w + i;
            const uint32_t w = g.uv[p];
            const float u = (float)(w & 0xffffu), v = (float)((w >> 16) & 0x7fffu);
            const float z = expf(g.logd[p] + shift);
            const float Xx = (u - cx) * z * ifx, Xy = (v - cy) * z * ify;
            if (!SPLAT) {
                if (src_pts) { src_pts[3 * (size_t)q] = Xx; src_pts[3 * (size_t)q + 1] = Xy; src_pts[3 * (size_t)q + 2] = z; }
                if (seg_ids) seg_ids[q] = td.x;
                if (src_ok) src_ok[q] = ((w >> 31) && z > 1e-7f) ? 1 : 0;
            } else {
                const float Yx = fmaf(R[0], Xx, fmaf(R[1], Xy, R[2] * z)) + tt[0];
                const float Yy = fmaf(R[3], Xx, fmaf(R[4], Xy, R[5] * z)) + tt[1];
                const float Yz = fmaf(R[6], Xx, fmaf(R[7], Xy, R[8] * z)) + tt[2];
                const float zi = fabsf(Yz) > 1e-6f ? 1.0f / Yz : 1e-6f;
                const float up = fmaf(Yx * fx, zi, cx), vp = fmaf(Yy * fy, zi, cy);
                // .long() truncates toward zero (core/ops.py:66); keep NaN/inf out
                if (!(Yz > 1e-6f) || !isfinite(up) || !isfinite(vp)) continue;
                if (fabsf(up) > 1e6f || fabsf(vp) > 1e6f) continue;
                const int col = (int)up, row = (int)vp;
                if (row < 0 || row >= g.H || col < 0 || col >= g.W) continue;
                const int idx = row * g.W + col;
                if (mean) {
                    if (Yz < 2.0e9f) atomicAdd(sum + idx, to_fx(Yz));
                    atomicAdd(keys + idx, 1ull);
                } else {
                    // last writer in point order wins == CPU scatter_ semantics
                    const unsigned long long key = ((unsigned long long)(q + 1) << 32) | __float_as_uint(Yz);
                    atomicMax(keys + idx, key);
                }
            }
        }
    }
}

__global__ void k_splat_resolve(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ sum,
                                int mean, int HW, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (mean)   // scatter_reduce_('mean') with include_self=True: (0 + sum) / (count + 1)
            out[i] = key ? (float)((double)sum[i] * (1.0 / SPB_AVG_FX)) / (float)(key + 1ull) : 0.0f;
        else
            out[i] = key ? __uint_as_float((unsigned)(key & 0xffffffffull)) : 0.0f;
    }
}

static inline int lift_blocks(int n_tiles) {
    int bx = (n_tiles + 7) / 8;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    return bx < 1 ? 1 : bx;
}

extern "C" int spb_lift_points(const SpbGeom* geom, const float* k, float* src_pts, int64_t* seg_ids,
                               uint8_t* src_ok, void* stream) {
    if (!geom || !k || geom->n_tiles < 1) return SPB_EINVAL;
    k_lift<false><<<lift_blocks(geom->n_tiles), 256, 0, (cudaStream_t)stream>>>(*geom, k, nullptr, src_pts, seg_ids,
                                                                               src_ok, 0, nullptr, nullptr);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_depth_splat(const SpbGeom* geom, const float* k, const float* pose, int mean,
                               unsigned long long* keys, unsigned long long* sum, float* out, void* stream) {
    if (!geom || !k || !keys || !out || geom->n_tiles < 1) return SPB_EINVAL;
    if (mean && !sum) return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = geom->H * geom->W;
    cudaError_t e = cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)HW, st);
    if (e != cudaSuccess) return (int)e;
    if (mean) {
        e = cudaMemsetAsync(sum, 0, sizeof(unsigned long long) * (size_t)HW, st);
        if (e != cudaSuccess) return (int)e;
    }
    k_lift<true><<<lift_blocks(geom->n_tiles), 256, 0, st>>>(*geom, k, pose, nullptr, nullptr, nullptr, mean, keys, sum);
    SPB_CHECK_LAUNCH();
    int bx = (HW + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_splat_resolve<<<bx, 256, 0, st>>>(keys, sum, mean, HW, out);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// image pyramid level: 3x3 [1 2 1]^2/16 Gaussian with reflect padding, then [::2, ::2]
// (image/gaussian_pyramid.py:53-85); one thread per output pixel per channel, rows coalesced
// ------------------------------------------------------------------------------------------------
__global__ void k_pyr_down(const float* __restrict__ src, int C, int H, int W, float* __restrict__ dst, int Ho,
                           int Wo) {
    const int c = blockIdx.z;
    const int yo = blockIdx.y;
    const float* s = src + (size_t)c * H * W;
    float* d = dst + ((size_t)c * Ho + yo) * Wo;
    const int y = 2 * yo;
    const int ym = (y == 0) ? (H > 1 ? 1 : 0) : y - 1;                     // reflect: -1 -> 1
    const int yp = (y + 1 >= H) ? (H > 1 ? H - 2 : 0) : y + 1;             // reflect:  H -> H-2
    const float* r0 = s + (size_t)ym * W;
    const float* r1 = s + (size_t)y * W;
    const float* r2 = s + (size_t)yp * W;
    for (int xo = blockIdx.x * blockDim.x + threadIdx.x; xo < Wo; xo += gridDim.x * blockDim.x) {
        const int x = 2 * xo;
        const int xm = (x == 0) ? (W > 1 ? 1 : 0) : x - 1;
        const int xp = (x + 1 >= W) ? (W > 1 ? W - 2 : 0) : x + 1;
        // same association as a 3x3 convolution accumulated row-major: weights w/16
        float acc = 0.0625f * r0[xm];
        acc = fmaf(0.125f, r0[x], acc);
        acc = fmaf(0.0625f, r0[xp], acc);
        acc = fmaf(0.125f, r1[xm], acc);
        acc = fmaf(0.25f, r1[x], acc);
        acc = fmaf(0.125f, r1[xp], acc);
        acc = fmaf(0.0625f, r2[xm], acc);
        acc = fmaf(0.125f, r2[x], acc);
        acc = fmaf(0.0625f, r2[xp], acc);
        d[xo] = acc;
    }
}

extern "C" int spb_pyr_down(const float* src, int C, int H, int W, float* dst, void* stream) {
    if (!src || !dst || C < 1 || C > 65535 || H < 1 || W < 1) return SPB_EINVAL;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    if (Ho > 65535) return SPB_ELIMIT;
    dim3 grid((Wo + 127) / 128, Ho, C);
    k_pyr_down<<<grid, 128, 0, (cudaStream_t)stream>>>(src, C, H, W, dst, Ho, Wo);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ------------------------------------------------------------------------------------------------
// VOID depth completion tail (depth_completion/segment_based_completion.py:21-27,48-54)
// ------------------------------------------------------------------------------------------------
// dense drop-in of render_depth_avg: per pixel over the N stacked depth maps, in place like the reference
__global__ void k_depth_avg_dense(float* __restrict__ depths, int N, int HW, float* __restrict__ out,
                                  uint8_t* __restrict__ invalid) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float mx = -INFINITY, sum = 0.f;
        int cnt = 0;
        for (int b = 0; b < N; ++b) {
            float d = depths[(size_t)b * HW + i];
            mx = fmaxf(mx, d);
            if (d < 1e-6f) { d = 0.f; depths[(size_t)b * HW + i] = 0.f; }
            sum += d;
            cnt += d > 1e-6f;
        }
        out[i] = sum / ((float)cnt + 1e-6f);
        invalid[i] = mx < 1e-6f;
    }
}

// fused variant straight from the compact geometry: no (N,H,W) tensor is ever materialised.
// Overlapping segments add into the same pixel from different warps: the sum is accumulated in 32.32 fixed point with
// 64-bit integer atomics and the count with 32-bit ones, so the result does not depend on the order of arrival
// (bit-reproducible, and the same on any number of GPUs); the exact sum is rounded to float32 once.
__global__ void k_depth_avg_compact(const __grid_constant__ SpbGeom g, const float* __restrict__ k,
                                    const uint8_t* __restrict__ visible, unsigned long long* __restrict__ sum,
                                    uint32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int wglobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int wstride = gridDim.x * (blockDim.x >> 5);
    const int4* tiles = reinterpret_cast<const int4*>(g.tiles);
    for (int t = wglobal; t < g.n_tiles; t += wstride) {
        const int4 td = tiles[t];
        if (visible != nullptr && !visible[td.x]) continue;
        const float shift = k[td.x] - g.seg_lkp[td.x];
        for (int i = lane; i < td.z; i += 32) {
            const int p = td.y + i;
            const uint32_t w = g.uv[p];
            const int u = (int)(w & 0xffffu), v = (int)((w >> 16) & 0x7fffu);
            const float z = expf(g.logd[p] + shift);
            if (z > 1e-6f && z < 2.0e9f) {
                atomicAdd(sum + (size_t)v * g.W + u, to_fx(z));
                atomicAdd(cnt + (size_t)v * g.W + u, 1u);
            }
        }
    }
}

__global__ void k_depth_avg_resolve(const unsigned long long* __restrict__ sum, const uint32_t* __restrict__ cnt, int HW,
                                    float* __restrict__ out, uint8_t* __restrict__ invalid) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        out[i] = (float)((double)sum[i] * (1.0 / SPB_AVG_FX)) / ((float)cnt[i] + 1e-6f);
        invalid[i] = cnt[i] == 0u;
    }
}

extern "C" int spb_depth_avg_dense(float* depths, int N, int H, int W, float* out, uint8_t* invalid, void* stream) {
    if (!depths || !out || !invalid || N < 1 || H < 1 || W < 1) return SPB_EINVAL;
    const int HW = H * W;
    int bx = (HW + 255) / 256;
    if (bx > spb_sm_count() * 16) bx = spb_sm_count() * 16;
    k_depth_avg_dense<<<bx, 256, 0, (cudaStream_t)stream>>>(depths, N, HW, out, invalid);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_depth_avg_compact(const SpbGeom* geom, const float* k, const uint8_t* visible,
                                     unsigned long long* sum, uint32_t* cnt, float* out, uint8_t* invalid, void* stream) {
    if (!geom || !k || !sum || !cnt || !out || !invalid || geom->n_tiles < 1) return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = geom->H * geom->W;
    cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(unsigned long long) * (size_t)HW, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)HW, st);
    if (e != cudaSuccess) return (int)e;
    k_depth_avg_compact<<<lift_blocks(geom->n_tiles), 256, 0, st>>>(*geom, k, visible, sum, cnt);
    SPB_CHECK_LAUNCH();
    int bx = (HW + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_depth_avg_resolve<<<bx, 256, 0, st>>>(sum, cnt, HW, out, invalid);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// estimate_depth_diff for arbitrary points (core/ops.py:59-96): same splat as k_lift<true>, input (P,3)
__global__ void k_splat_points(const float* __restrict__ pts, int P, const float* __restrict__ K, int H, int W,
                               int mean, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ sum,
                               uint8_t* __restrict__ valid) {
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gridDim.x * blockDim.x) {
        const float X = pts[3 * (size_t)q], Y = pts[3 * (size_t)q + 1], Z = pts[3 * (size_t)q + 2];
        const float zi = fabsf(Z) > 1e-6f ? 1.0f / Z : 1e-6f;
        const float up = fmaf(X * fx, zi, cx), vp = fmaf(Y * fy, zi, cy);
        bool ok = (Z > 1e-6f) && isfinite(up) && isfinite(vp) && fabsf(up) < 1e6f && fabsf(vp) < 1e6f;
        int col = 0, row = 0;
        if (ok) {
            col = (int)up; row = (int)vp;                     // .long() truncation toward zero
            ok = row >= 0 && row < H && col >= 0 && col < W;
        }
        if (valid) valid[q] = ok ? 1 : 0;
        if (!ok) continue;
        const int idx = row * W + col;
        if (mean) {
            if (Z < 2.0e9f) atomicAdd(sum + idx, to_fx(Z));
            atomicAdd(keys + idx, 1ull);
        } else {
            atomicMax(keys + idx, ((unsigned long long)(q + 1) << 32) | __float_as_uint(Z));
        }
    }
}

extern "C" int spb_depth_splat_points(const float* pts, int P, const float* K, int H, int W, int mean,
                                      unsigned long long* keys, unsigned long long* sum, float* out, uint8_t* valid,
                                      void* stream) {
    if (!pts || P < 1 || !K || H < 1 || W < 1 || !keys || !out) return SPB_EINVAL;
    if (mean && !sum) return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    cudaError_t e = cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)HW, st);
    if (e != cudaSuccess) return (int)e;
    if (mean) {
        e = cudaMemsetAsync(sum, 0, sizeof(unsigned long long) * (size_t)HW, st);
        if (e != cudaSuccess) return (int)e;
    }
    int bx = (P + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_splat_points<<<bx, 256, 0, st>>>(pts, P, K, H, W, mean, keys, sum, valid);
    SPB_CHECK_LAUNCH();
    int br = (HW + 255) / 256;
    if (br > spb_sm_count() * 8) br = spb_sm_count() * 8;
    k_splat_resolve<<<br, 256, 0, st>>>(keys, sum, mean, HW, out);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_tile_points(void) { return SPB_TILE; }

extern "C" int spb_version(void) { return 106 + 1000 * SPB_INGEST_FUSED; }
