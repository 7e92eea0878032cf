// Frame ingest: 8-bit frames as the dataset readers deliver them (HWC uint8: data/replica.py:55, data/tum_undistort.py:112)
// -> everything the fused alignment kernel streams, for MANY (source keyframe, target frame) pairs per launch.
//
// The reference converts on the host and uploads float32 (tool/etc.py:37-40 `image_tt`: (u8 / 255.).float().to(device),
// HWC -> CHW; 12 bytes per pixel over PCIe).  Here the 3-byte pixels travel and `image_tt` runs on the device -- the same
// float32 division, so the frames are bit-identical -- fused with the re-layouts that follow:
//   k_ingest_frames : target u8 HWC -> RGBA-interleaved float4 (spb_pack_rgba of image_tt), source u8 HWC -> planar CHW float
//   k_ingest_sample : cached source samples at every point's own pixel, grid.y = job   } the per-item bodies of
//   k_ingest_pack   : tile-major level buffer, grid.y = job                            } spb_sample_source /
//                                                                                        spb_build_tile_pack (spb_frame_stages.cuh)
// Job descriptors and geometries live in DEVICE memory (uploaded once per batch), so a step's ingest is three launches
// regardless of the number of pairs.  HBM-streaming, coalesced; bandwidth ~45 MB per 640x480 pair.
#include "spb_common.cuh"
#include "spb_frame_stages.cuh"

// image_tt arithmetic: float32(u8) / 255.f, IEEE division (torch divides a uint8 tensor by a Python float in float32)
__device__ __forceinline__ float u8_unit(uint32_t b) { return __fdiv_rn((float)b, 255.0f); }

__global__ void k_ingest_frames(const SpbFrameJob* __restrict__ jobs) {
    const SpbFrameJob jb = jobs[blockIdx.y];
    const int HW = jb.Hl * jb.Wl;
    float4* rgba = reinterpret_cast<float4*>(jb.trg_rgba);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        if (jb.trg_u8 && rgba) {
            const uint8_t* p = jb.trg_u8 + 3 * (size_t)i;
            rgba[i] = make_float4(u8_unit(p[0]), u8_unit(p[1]), u8_unit(p[2]), 0.f);
        }
        if (jb.src_u8 && jb.src_planar) {
            const uint8_t* p = jb.src_u8 + 3 * (size_t)i;
            jb.src_planar[i] = u8_unit(p[0]);
            jb.src_planar[(size_t)HW + i] = u8_unit(p[1]);
            jb.src_planar[2 * (size_t)HW + i] = u8_unit(p[2]);
        }
    }
}

// spb_sample_source for every job: cached source samples at the points' own pixels (spb_frame_stages.cuh)
__global__ void k_ingest_sample(const SpbGeom* __restrict__ geoms, const SpbFrameJob* __restrict__ jobs) {
    const SpbFrameJob jb = jobs[blockIdx.y];
    if (!jb.src_planar || !jb.src_rgb) return;
    const SpbGeom g = geoms[jb.geom];
    const SourceSampleScale sc = source_sample_scale(g, jb.Hl, jb.Wl);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pad; p += gridDim.x * blockDim.x)
        sample_source_point(g, jb.src_planar, jb.Hl, jb.Wl, sc, p, jb.src_rgb);
}

// spb_build_tile_pack for every job: one warp per tile (spb_frame_stages.cuh)
__global__ void k_ingest_pack(const SpbGeom* __restrict__ geoms, const SpbFrameJob* __restrict__ jobs) {
    const SpbFrameJob jb = jobs[blockIdx.y];
    if (!jb.src_rgb || !jb.pack) return;
    const SpbGeom g = geoms[jb.geom];
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < g.n_tiles; t += gridDim.x * (blockDim.x >> 5))
        build_tile_pack_tile(g, jb.src_rgb, jb.pack, t, lane);
}

// SPB_INGEST_FUSED (spb_common.cuh, default 0) -- EXPERIMENT, not yet measured (round-2 plan): the
// source half of the ingest is ONE kernel that samples the 8-bit source frame and writes the tile pack directly
#if SPB_INGEST_FUSED
// One warp per tile: header + uv + logd as k_ingest_pack, and the three colour arrays computed on the fly with the
// arithmetic of k_ingest_sample from taps converted with u8_unit -- the planar float frame and the [3][n_pad] sample
// array are not needed on this path (written only when the job supplies them), 22 MB instead of 45 MB per 640x480 pair.
__global__ void k_ingest_fused(const SpbGeom* __restrict__ geoms, const SpbFrameJob* __restrict__ jobs) {
    const SpbFrameJob jb = jobs[blockIdx.y];
    if (!jb.src_u8 || !jb.pack) return;
    const SpbGeom g = geoms[jb.geom];
    const int lane = threadIdx.x & 31;
    const int Hl = jb.Hl, Wl = jb.Wl;
    const float tiw = 2.0f * (1.0f / (float)(g.W - 1));
    const float tih = 2.0f * (1.0f / (float)(g.H - 1));
    const float sx = 0.5f * (float)(Wl - 1), sy = 0.5f * (float)(Hl - 1);
    const uint32_t* lu = reinterpret_cast<const uint32_t*>(g.logd);
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < g.n_tiles; t += gridDim.x * (blockDim.x >> 5)) {
        const int4 td = reinterpret_cast<const int4*>(g.tiles)[t];
        uint32_t* o = jb.pack + (size_t)t * SPB_PACK_WORDS;
        if (lane < 4) o[lane] = lane == 0 ? (uint32_t)td.x : (lane == 1 ? (uint32_t)td.z : (lane == 2 ? (uint32_t)td.w : 0u));
        for (int i = lane; i < SPB_TILE; i += 32) {
            const bool on = i < td.z;
            const size_t p = (size_t)td.y + (on ? i : 0);
            const uint32_t w = g.uv[p];
            float val[3] = {0.f, 0.f, 0.f};
            if (on) {
                const float u = (float)(w & 0xffffu), v = (float)((w >> 16) & 0x7fffu);
                const float ix = (fmaf(u, tiw, -1.0f) + 1.0f) * sx;
                const float iy = (fmaf(v, tih, -1.0f) + 1.0f) * sy;
                const float fxf = floorf(ix), fyf = floorf(iy);
                const int x0 = (int)fxf, y0 = (int)fyf;
                const float fx = ix - fxf, fy = iy - fyf;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    auto tap = [&](int x, int y) -> float {
                        return (x < 0 || y < 0 || x >= Wl || y >= Hl) ? 0.f : u8_unit(jb.src_u8[3 * ((size_t)y * Wl + x) + ch]);
                    };
                    float d0, d1;
                    blend(tap(x0, y0), tap(x0 + 1, y0), tap(x0, y0 + 1), tap(x0 + 1, y0 + 1), fx, fy, val[ch], d0, d1);
                    if (jb.src_rgb) jb.src_rgb[(size_t)ch * g.n_pad + p] = val[ch];
                }
            }
            o[4 + i] = on ? w : 0u;
            o[4 + SPB_TILE + i] = on ? lu[p] : 0u;
            o[4 + 2 * SPB_TILE + i] = on ? __float_as_uint(val[0]) : 0u;
            o[4 + 3 * SPB_TILE + i] = on ? __float_as_uint(val[1]) : 0u;
            o[4 + 4 * SPB_TILE + i] = on ? __float_as_uint(val[2]) : 0u;
        }
    }
}
#endif

extern "C" int spb_ingest_u8(const SpbGeom* geoms, const SpbFrameJob* jobs, int n_jobs, int max_pixels, int max_pad,
                             int max_tiles, void* stream) {
    if (!geoms || !jobs || n_jobs < 1 || max_pixels < 1 || max_pad < 1 || max_tiles < 1) return SPB_EINVAL;
    if (n_jobs > 65535) return SPB_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    // grid.x sized so that all jobs together fill the 148 SMs a few times over; every kernel loops with a grid stride
    auto gx = [&](int items_per_cta, int items) {
        int want = (items + items_per_cta - 1) / items_per_cta;
        int cap = (spb_sm_count() * 16 + n_jobs - 1) / n_jobs;
        if (cap < 1) cap = 1;
        return want < cap ? want : cap;
    };
    k_ingest_frames<<<dim3(gx(256, max_pixels), n_jobs), 256, 0, st>>>(jobs);
    SPB_CHECK_LAUNCH();
#if SPB_INGEST_FUSED
    (void)max_pad;
    k_ingest_fused<<<dim3(gx(8, max_tiles), n_jobs), 256, 0, st>>>(geoms, jobs);
    SPB_CHECK_LAUNCH();
#else
    k_ingest_sample<<<dim3(gx(256, max_pad), n_jobs), 256, 0, st>>>(geoms, jobs);
    SPB_CHECK_LAUNCH();
    k_ingest_pack<<<dim3(gx(8, max_tiles), n_jobs), 256, 0, st>>>(geoms, jobs);
    SPB_CHECK_LAUNCH();
#endif
    return SPB_OK;
}

// image_tt alone (tool/etc.py:37-40) for one frame: u8 HWC -> float32 CHW in [0,1]
__global__ void k_image_tt(const uint8_t* __restrict__ hwc, int HW, float* __restrict__ chw) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const uint8_t* p = hwc + 3 * (size_t)i;
        chw[i] = u8_unit(p[0]);
        chw[(size_t)HW + i] = u8_unit(p[1]);
        chw[2 * (size_t)HW + i] = u8_unit(p[2]);
    }
}

extern "C" int spb_image_tt(const uint8_t* hwc, int H, int W, float* chw, void* stream) {
    if (!hwc || !chw || H < 1 || W < 1) return SPB_EINVAL;
    const int HW = H * W;
    int bx = (HW + 255) / 256;
    if (bx > spb_sm_count() * 8) bx = spb_sm_count() * 8;
    k_image_tt<<<bx, 256, 0, (cudaStream_t)stream>>>(hwc, HW, chw);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
