// Per-segment re-initialisation of the log-depth seeds from a rendered / sparse depth map
// (reference odometery/depth_init.py:10-67).
//
// For every segment b: over the segment's pixels whose estimated depth is valid (>= 1e-6),
//   shift = log(est_depth[v,u]) - L[b,v,u];   k_b = (mean | lower median)(shift) + L[b, kp_b]
// and segments with no valid pixel receive the lower median of the visible segments' k.
// The reference loops over segments in Python calling torch.median on a masked dense tensor; here
// one CTA per segment runs an exact 4-pass 8-bit radix select over the compact point list (the
// shifts are recomputed on the fly, 8 B/point/pass, nothing is materialised or sorted).
#include "spb_common.cuh"

#define RI_THREADS 256

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(RI_THREADS)
k_segment_reinit(const __grid_constant__ SpbGeom g, const float* __restrict__ est, int mode,
                 float* __restrict__ seg_val, uint8_t* __restrict__ visible) {
    const int b = blockIdx.x;
    const int4* tiles = reinterpret_cast<const int4*>(g.tiles);
    const int t0 = g.seg_tile[b], t1 = g.seg_tile[b + 1];
    __shared__ int s_hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_k, s_cnt;
    __shared__ double s_sum[RI_THREADS / 32];
    if (t1 <= t0) {            // empty mask
        if (threadIdx.x == 0) { seg_val[b] = 0.f; visible[b] = 0; }
        return;
    }
    const int4 first = tiles[t0], last = tiles[t1 - 1];
    const int start = first.y, count = last.y + last.z - first.y;
    const float eps = 1e-6f;

    auto shift_of = [&](int i, bool& ok) -> float {
        const int p = start + i;
        const uint32_t w = g.uv[p];
        const int u = (int)(w & 0xffffu), v = (int)((w >> 16) & 0x7fffu);
        const float d = est[(size_t)v * g.W + u];
        ok = !(d < eps);                       // reference: invalid = est < eps (NaN counts as valid there too)
        return logf(ok ? d : eps) - g.logd[p];
    };

    if (mode == 0) {           // mean
        double sum = 0.0;
        int cnt = 0;
        for (int i = threadIdx.x; i < count; i += RI_THREADS) {
            bool ok;
            const float s = shift_of(i, ok);
            if (ok) { sum += (double)s; ++cnt; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_hist[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            int n = 0;
            for (int w = 0; w < RI_THREADS / 32; ++w) { tot += s_sum[w]; n += s_hist[w]; }
            visible[b] = n > 0;
            seg_val[b] = n > 0 ? (float)(tot / (double)n) + g.seg_lkp[b] : 0.f;
        }
        return;
    }

    // lower median by radix select, most significant byte first
    if (threadIdx.x == 0) { s_prefix = 0u; s_k = -1; s_cnt = 0; }
    for (int pass = 0; pass < 4; ++pass) {
        for (int i = threadIdx.x; i < 256; i += RI_THREADS) s_hist[i] = 0;
        __syncthreads();
        const int sh = 24 - 8 * pass;
        const uint32_t prefix = s_prefix;
        const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (sh + 8));
        for (int i = threadIdx.x; i < count; i += RI_THREADS) {
            bool ok;
            const float s = shift_of(i, ok);
            if (!ok) continue;
            const uint32_t key = f2key(s);
            if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> sh) & 0xffu], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = s_k;
            if (pass == 0) {
                int n = 0;
                for (int i = 0; i < 256; ++i) n += s_hist[i];
                s_cnt = n;
                k = n > 0 ? (n - 1) / 2 : -1;          // torch.median: lower of the two middle values
            }
            if (k >= 0) {
                int run = 0, bin = 0;
                for (; bin < 256; ++bin) {
                    if (run + s_hist[bin] > k) break;
                    run += s_hist[bin];
                }
                s_prefix = prefix | ((uint32_t)bin << sh);
                k -= run;
            }
            s_k = k;
        }
        __syncthreads();
        if (s_cnt == 0) break;
    }
    if (threadIdx.x == 0) {
        const bool vis = s_cnt > 0;
        visible[b] = vis ? 1 : 0;
        seg_val[b] = vis ? key2f(s_prefix) + g.seg_lkp[b] : 0.f;
    }
}

// invisible segments <- lower median of the visible segments' values (single CTA, O(N^2) ranks)
__global__ void k_reinit_fill(const float* __restrict__ seg_val, const uint8_t* __restrict__ visible, int N,
                              float* __restrict__ out, int* __restrict__ n_visible) {
    __shared__ float s_med;
    __shared__ int s_nvis;
    if (threadIdx.x == 0) { s_nvis = 0; s_med = 0.f; }
    __syncthreads();
    int loc = 0;
    for (int b = threadIdx.x; b < N; b += blockDim.x) loc += visible[b] ? 1 : 0;
    atomicAdd(&s_nvis, loc);
    __syncthreads();
    const int nvis = s_nvis;
    const int kth = nvis > 0 ? (nvis - 1) / 2 : -1;
    for (int b = threadIdx.x; b < N; b += blockDim.x) {
        if (!visible[b]) continue;
        const float v = seg_val[b];
        int rank = 0;
        for (int j = 0; j < N; ++j) {
            if (!visible[j]) continue;
            const float w = seg_val[j];
            rank += (w < v) || (w == v && j < b);
        }
        if (rank == kth) s_med = v;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < N; b += blockDim.x) out[b] = visible[b] ? seg_val[b] : s_med;
    if (threadIdx.x == 0 && n_visible) *n_visible = nvis;
}

extern "C" int spb_segment_reinit(const SpbGeom* geom, const float* est_depth, int mode, float* seg_val,
                                  uint8_t* visible, float* out_k, int32_t* n_visible, void* stream) {
    if (!geom || !est_depth || !seg_val || !visible || !out_k || geom->n_seg < 1) return SPB_EINVAL;
    if (mode != 0 && mode != 1) return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_segment_reinit<<<geom->n_seg, RI_THREADS, 0, st>>>(*geom, est_depth, mode, seg_val, visible);
    SPB_CHECK_LAUNCH();
    k_reinit_fill<<<1, 256, 0, st>>>(seg_val, visible, geom->n_seg, out_k, n_visible);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
