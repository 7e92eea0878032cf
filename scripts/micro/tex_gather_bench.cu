// Microbenchmark (tuning evidence, not product code): the bilinear-footprint gather of the fused alignment kernel
// through the LSU (4 x LDG.128 on an RGBA-interleaved float image, what spb_align.cu does) against the texture unit
// (3 x tex2Dgather on planar float CUDA arrays: the 2x2 footprint of one channel per instruction).
// Same access pattern as the benchmark workload: 64 images of 640x480, points of 18-pixel-wide vertical strips in
// (segment,row,col) order, 128-point tiles strided over warps, one point in flight per lane, ~100 dependent FMAs per point.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_gather_bench tex_gather_bench.cu && ./tex_gather_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int W = 640, H = 480, NIMG = 64, SEGW = 18, NSEG = 64, TILE = 128;
constexpr int PTS_PER_SEG = SEGW * H;                 // 8640
constexpr int TILES_PER_SEG = (PTS_PER_SEG + TILE - 1) / TILE;
constexpr int TILES = NSEG * TILES_PER_SEG;

struct Img {
    const float4* rgba;
    cudaTextureObject_t tex[3];
};

__device__ __forceinline__ void point_coords(int tile, int i, float dx, float dy, float& ix, float& iy, bool& ok) {
    const int seg = tile / TILES_PER_SEG, p = (tile % TILES_PER_SEG) * TILE + i;
    ok = p < PTS_PER_SEG;
    const int row = p / SEGW, col = p % SEGW;
    const int u = min(seg * 10 + col, W - 3);          // strips 10 px apart, dilated to 18: overlap like the workload
    ix = (float)u + dx + 0.013f * (float)row * 0.01f;
    iy = fminf((float)row + dy, (float)(H - 2) - 0.5f);
}

template <int EXTRA>
__device__ __forceinline__ float consume(float a, float b, float c, float fx, float fy) {
    float s = a + b * fx + c * fy;
#pragma unroll
    for (int k = 0; k < EXTRA; ++k) s = fmaf(s, 0.999f, fx);
    return s;
}

template <int EXTRA>
__global__ void __launch_bounds__(256, 3) k_ldg(const Img* imgs, int ctas, float dx, float dy, float* out) {
    const Img im = imgs[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int t = blockIdx.x * 8 + warp; t < TILES; t += ctas * 8) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float ix, iy; bool ok;
            point_coords(t, j * 32 + lane, dx, dy, ix, iy, ok);
            if (!ok) continue;
            const float fxf = floorf(ix), fyf = floorf(iy);
            const float fx = ix - fxf, fy = iy - fyf;
            const float4* p0 = im.rgba + ((int)fyf * W + (int)fxf);
            const float4 nw = __ldg(p0), ne = __ldg(p0 + 1), sw = __ldg(p0 + W), se = __ldg(p0 + W + 1);
            float I[3];
            {
                float top = fmaf(fx, ne.x - nw.x, nw.x), bot = fmaf(fx, se.x - sw.x, sw.x); I[0] = fmaf(fy, bot - top, top);
                top = fmaf(fx, ne.y - nw.y, nw.y); bot = fmaf(fx, se.y - sw.y, sw.y); I[1] = fmaf(fy, bot - top, top);
                top = fmaf(fx, ne.z - nw.z, nw.z); bot = fmaf(fx, se.z - sw.z, sw.z); I[2] = fmaf(fy, bot - top, top);
            }
            acc += consume<EXTRA>(I[0], I[1], I[2], fx, fy);
        }
    }
    if (acc == 12345.678f) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(out + 1, acc);
}

template <int EXTRA>
__global__ void __launch_bounds__(256, 3) k_tex(const Img* imgs, int ctas, float dx, float dy, float* out) {
    const Img im = imgs[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int t = blockIdx.x * 8 + warp; t < TILES; t += ctas * 8) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float ix, iy; bool ok;
            point_coords(t, j * 32 + lane, dx, dy, ix, iy, ok);
            if (!ok) continue;
            const float fxf = floorf(ix), fyf = floorf(iy);
            const float fx = ix - fxf, fy = iy - fyf;
            float I[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // gather footprint at the centre of the 2x2 block: w = (x0,y0), z = (x0+1,y0), x = (x0,y0+1), y = (x0+1,y0+1)
                const float4 g = tex2Dgather<float4>(im.tex[c], fxf + 1.0f, fyf + 1.0f, 0);
                const float top = fmaf(fx, g.z - g.w, g.w), bot = fmaf(fx, g.y - g.x, g.x);
                I[c] = fmaf(fy, bot - top, top);
            }
            acc += consume<EXTRA>(I[0], I[1], I[2], fx, fy);
        }
    }
    if (acc == 12345.678f) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(out + 1, acc);
}

int main() {
    std::vector<float> planar((size_t)3 * H * W), rgba((size_t)4 * H * W);
    std::vector<Img> imgs(NIMG);
    for (int n = 0; n < NIMG; ++n) {
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const float v = 0.5f + 0.5f * sinf(0.05f * x + 0.03f * y + c + 0.1f * n);
                    planar[((size_t)c * H + y) * W + x] = v;
                    rgba[((size_t)y * W + x) * 4 + c] = v;
                }
        float* d_rgba;
        CK(cudaMalloc(&d_rgba, rgba.size() * 4));
        CK(cudaMemcpy(d_rgba, rgba.data(), rgba.size() * 4, cudaMemcpyHostToDevice));
        imgs[n].rgba = reinterpret_cast<const float4*>(d_rgba);
        for (int c = 0; c < 3; ++c) {
            cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
            cudaArray_t arr;
            CK(cudaMallocArray(&arr, &desc, W, H, cudaArrayTextureGather));
            CK(cudaMemcpy2DToArray(arr, 0, 0, planar.data() + (size_t)c * H * W, W * 4, W * 4, H, cudaMemcpyHostToDevice));
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = arr;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            CK(cudaCreateTextureObject(&imgs[n].tex[c], &rd, &td, nullptr));
        }
    }
    Img* d_imgs;
    CK(cudaMalloc(&d_imgs, sizeof(Img) * NIMG));
    CK(cudaMemcpy(d_imgs, imgs.data(), sizeof(Img) * NIMG, cudaMemcpyHostToDevice));
    float* d_out;
    CK(cudaMalloc(&d_out, 8));
    const int ctas = 27;
    dim3 grid(ctas, NIMG);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double npts = (double)NIMG * NSEG * PTS_PER_SEG;
    auto timeit = [&](const char* name, auto launch) {
        float h[2] = {0, 0};
        CK(cudaMemcpy(d_out, h, 8, cudaMemcpyHostToDevice));
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(d_out, h, 8, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
        printf("%-28s %.4f ms per launch  %.1f Gpoint/s  checksum %.6e\n", name, ms / reps, npts / (ms / reps) * 1e-6, h[1] / reps);
    };
    printf("points per launch %.0f (%d images x %d tiles)\n", npts, NIMG, TILES);
    timeit("LDG.128 x4, light (8 fma)", [&] { k_ldg<8><<<grid, 256>>>(d_imgs, ctas, 1.3f, 0.7f, d_out); });
    timeit("TLD4 x3,    light (8 fma)", [&] { k_tex<8><<<grid, 256>>>(d_imgs, ctas, 1.3f, 0.7f, d_out); });
    timeit("LDG.128 x4, heavy (120 fma)", [&] { k_ldg<120><<<grid, 256>>>(d_imgs, ctas, 1.3f, 0.7f, d_out); });
    timeit("TLD4 x3,    heavy (120 fma)", [&] { k_tex<120><<<grid, 256>>>(d_imgs, ctas, 1.3f, 0.7f, d_out); });
    CK(cudaGetLastError());
    return 0;
}
