// Shared device helpers for the spb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spb200.h"

// build-time experiment switches (scripts/build_variant.sh); the default library is built with all of them off and
// reports them through spb_version() (thousands digit)
#ifndef SPB_INGEST_FUSED
#define SPB_INGEST_FUSED 0                      // 1: fused source ingest (spb_ingest.cu)
#endif

#ifndef SPB_WARPS
#define SPB_WARPS 8                             // warps per CTA of the fused kernels (tunable)
#endif
#define SPB_THREADS (SPB_WARPS * 32)
#define SPB_PPT (SPB_TILE / 32)   // points per lane per tile

// number of SMs of the current device, queried once per process (every rank sees one kind of GPU); grids of the
// streaming kernels are capped at a multiple of it
static inline int spb_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            return 148;     // B200; only reached without a usable device (sizing queries on a CPU-only host)
    }
    return n;
}

#define SPB_CHECK_LAUNCH()                                  \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

// Per-pair uniform values staged in shared memory once per CTA.
enum CtxSlot {
    C_R = 0,        // 9 floats, row-major rotation
    C_T = 9,        // 3 translation
    C_IFX = 12, C_IFY, C_CX, C_CY,          // source intrinsics (reciprocal focal)
    C_FXT = 16, C_FYT, C_CXT, C_CYT,        // target intrinsics
    C_TIW = 20, C_TIH,                      // 2 * fl32(1/(W-1)), 2 * fl32(1/(H-1))
    C_SX = 22, C_SY,                        // 0.5 * (Wl-1), 0.5 * (Hl-1)
    C_EA = 24, C_BB,                        // exp(-(a_t-a_s)), b_t-b_s
    C_TAU = 26,
    C_KX = 27, C_KY,                        // d ix / d u', d iy / d v'
    C_N = 32
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fill the context from a pair descriptor + geometry grid.  Called by the first warp.
__device__ __forceinline__ void fill_ctx(float* s, const SpbPair& pr, const float* Ksrc, int H, int W) {
    const int l = threadIdx.x;
    if (l < 3) {
        s[C_R + 3 * l + 0] = pr.pose[4 * l + 0];
        s[C_R + 3 * l + 1] = pr.pose[4 * l + 1];
        s[C_R + 3 * l + 2] = pr.pose[4 * l + 2];
        s[C_T + l] = pr.pose[4 * l + 3];
    } else if (l == 3) {
        s[C_IFX] = 1.0f / Ksrc[0];
        s[C_IFY] = 1.0f / Ksrc[4];
        s[C_CX] = Ksrc[2];
        s[C_CY] = Ksrc[5];
    } else if (l == 4) {
        s[C_FXT] = pr.K_trg[0];
        s[C_FYT] = pr.K_trg[4];
        s[C_CXT] = pr.K_trg[2];
        s[C_CYT] = pr.K_trg[5];
    } else if (l == 5) {
        // reference: inv = 1.0f / (dims - 1) in float32; x_norm = 2 * x * inv - 1  (tool/point_utils.py:31-35)
        const float iw = 1.0f / (float)(W - 1);
        const float ih = 1.0f / (float)(H - 1);
        const float sx = 0.5f * (float)(pr.Wl - 1);
        const float sy = 0.5f * (float)(pr.Hl - 1);
        s[C_TIW] = 2.0f * iw;
        s[C_TIH] = 2.0f * ih;
        s[C_SX] = sx;
        s[C_SY] = sy;
        s[C_KX] = sx * (2.0f * iw);
        s[C_KY] = sy * (2.0f * ih);
        s[C_TAU] = pr.tau;
    } else if (l == 6) {
        float a = 0.f, b = 0.f;
        if (pr.aff_src != nullptr && pr.aff_trg != nullptr) {
            a = pr.aff_trg[0] - pr.aff_src[0];
            b = pr.aff_trg[1] - pr.aff_src[1];
        }
        s[C_EA] = expf(-a);
        s[C_BB] = b;
    }
}

struct Taps {
    float4 nw, ne, sw, se;
};

// bilinear blend of one channel + derivatives w.r.t. (ix, iy)
__device__ __forceinline__ void blend(float nw, float ne, float sw, float se, float fx, float fy,
                                      float& val, float& dix, float& diy) {
    const float d0 = ne - nw;
    const float d1 = se - sw;
    const float top = fmaf(fx, d0, nw);
    const float bot = fmaf(fx, d1, sw);
    diy = bot - top;
    val = fmaf(fy, diy, top);
    dix = fmaf(fy, d1 - d0, d0);
}

// zero-padded tap fetch (stats path and source sampling: coordinates may touch the border)
__device__ __forceinline__ float4 tap_rgba(const float4* img, int x, int y, int Wl, int Hl) {
    if (x < 0 || y < 0 || x >= Wl || y >= Hl) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(img + (size_t)y * Wl + x);
}
