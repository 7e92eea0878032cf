// Per-item bodies of the per-frame preparation kernels, shared by the per-pair entry points (spb_geom.cu:
// spb_sample_source, spb_build_tile_pack) and the batched frame ingest (spb_ingest.cu: spb_ingest_u8), so that both
// produce bit-identical buffers by construction.
#pragma once
#include "spb_common.cuh"

// scale factors of the own-pixel sample (geometry grid -> level image), computed once per kernel
struct SourceSampleScale {
    float tiw, tih, sx, sy;
    size_t HW;
};
__device__ __forceinline__ SourceSampleScale source_sample_scale(const SpbGeom& g, int Hl, int Wl) {
    SourceSampleScale s;
    s.tiw = 2.0f * (1.0f / (float)(g.W - 1));
    s.tih = 2.0f * (1.0f / (float)(g.H - 1));
    s.sx = 0.5f * (float)(Wl - 1);
    s.sy = 0.5f * (float)(Hl - 1);
    s.HW = (size_t)Hl * Wl;
    return s;
}

// cached source samples of point p: bilinear of the planar source level image at the point's own pixel scaled to the
// level (the reference re-projects the unprojected point, which returns its own pixel up to float rounding;
// core/dense_optim.py:315-317).  out = [3][n_pad]
__device__ __forceinline__ void sample_source_point(const SpbGeom& g, const float* __restrict__ img, int Hl, int Wl,
                                                    const SourceSampleScale& s, int p, float* __restrict__ out) {
    const uint32_t w = g.uv[p];
    const float u = (float)(w & 0xffffu), v = (float)((w >> 16) & 0x7fffu);
    const float ix = (fmaf(u, s.tiw, -1.0f) + 1.0f) * s.sx;
    const float iy = (fmaf(v, s.tih, -1.0f) + 1.0f) * s.sy;
    const float fxf = floorf(ix), fyf = floorf(iy);
    const int x0 = (int)fxf, y0 = (int)fyf;
    const float fx = ix - fxf, fy = iy - fyf;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float* pl = img + ch * s.HW;
        auto tap = [&](int x, int y) -> float {
            return (x < 0 || y < 0 || x >= Wl || y >= Hl) ? 0.f : pl[(size_t)y * Wl + x];
        };
        float val, d0, d1;
        blend(tap(x0, y0), tap(x0 + 1, y0), tap(x0, y0 + 1), tap(x0 + 1, y0 + 1), fx, fy, val, d0, d1);
        out[(size_t)ch * g.n_pad + p] = val;
    }
}

// tile t of the tile-major level buffer: header {segment, count, unpadded start, 0} + uv | logd | r | g | b, one warp
// per tile (coalesced), zero-filled beyond the tile's count.  rgb = [3][n_pad] cached source samples
__device__ __forceinline__ void build_tile_pack_tile(const SpbGeom& g, const float* __restrict__ rgb,
                                                     uint32_t* __restrict__ pack, int t, int lane) {
    const int4 td = reinterpret_cast<const int4*>(g.tiles)[t];
    uint32_t* o = pack + (size_t)t * SPB_PACK_WORDS;
    if (lane < 4) o[lane] = lane == 0 ? (uint32_t)td.x : (lane == 1 ? (uint32_t)td.z : (lane == 2 ? (uint32_t)td.w : 0u));
    const uint32_t* rgbu = reinterpret_cast<const uint32_t*>(rgb);
    const uint32_t* lu = reinterpret_cast<const uint32_t*>(g.logd);
    for (int i = lane; i < SPB_TILE; i += 32) {
        const bool on = i < td.z;
        const size_t p = (size_t)td.y + (on ? i : 0);
        o[4 + i] = on ? g.uv[p] : 0u;
        o[4 + SPB_TILE + i] = on ? lu[p] : 0u;
        o[4 + 2 * SPB_TILE + i] = on ? rgbu[p] : 0u;
        o[4 + 3 * SPB_TILE + i] = on ? rgbu[(size_t)g.n_pad + p] : 0u;
        o[4 + 4 * SPB_TILE + i] = on ? rgbu[2 * (size_t)g.n_pad + p] : 0u;
    }
}
