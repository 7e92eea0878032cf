// Nearest-valid hole filling of a depth map (reference depth_completion/fill_in_tools.py:5-7
// `fill_depth`: scipy.ndimage.distance_transform_edt(invalid, return_indices=True) followed by depth[indices]).
//
// Exact Euclidean feature transform in integer arithmetic, separable:
//   pass 1  per column: nearest valid row above / below every pixel (strips of rows in parallel, joined in shared memory);
//   pass 2  per pixel (r, c): the column c' minimising (c - c')^2 + dv(r, c')^2, searched outwards from c and
//           stopped as soon as k^2 exceeds the best squared distance (holes of a completed depth map are small).
// Ties are resolved like scipy 1.18's feature transform (pinned by tests/golden/fill_depth.npz): among equidistant valid
// pixels the smallest column wins, then the smallest row.  A frame without any valid pixel gets scipy's index (-1, 0) =
// numpy's depth[H - 1, 0] everywhere.
#include "spb_common.cuh"

#define FILL_BIG 0x3fffffff                    // "no valid pixel in this column" (k^2 <= 32767^2 adds without overflow)
#define FILL_THREADS 256

// pass 1: a CTA owns 32 columns; the rows are cut into 32 strips, one thread per (column, strip).  Every thread finds the
// first / last valid row of its strip, the strips of a column meet in shared memory (nearest valid row before / after the
// strip), then each thread resolves its own rows: the dependent chain is H/32 rows long instead of H.
#define FC_COLS 32
#define FC_STRIPS 32
__global__ void __launch_bounds__(FC_COLS * FC_STRIPS)
k_fill_cols(const uint8_t* __restrict__ invalid, int H, int W, int32_t* __restrict__ near_row) {
    __shared__ int s_last[FC_STRIPS][FC_COLS + 1], s_first[FC_STRIPS][FC_COLS + 1];
    const int tx = threadIdx.x, s = threadIdx.y;
    const int c = blockIdx.x * FC_COLS + tx;
    const int L = (H + FC_STRIPS - 1) / FC_STRIPS;
    const int r0 = min(H, s * L), r1 = min(H, r0 + L);
    const size_t base = (size_t)blockIdx.y * H * W + min(c, W - 1);
    const uint8_t* m = invalid + base;
    int32_t* o = near_row + base;
    int last = -1, first = -1;
    if (c < W) {
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
            if (!m[(size_t)r * W]) {
                last = r;
                if (first < 0) first = r;
            }
        }
    }
    s_last[s][tx] = last;
    s_first[s][tx] = first;
    __syncthreads();
    if (c >= W) return;
    int above = -1, below = -1;
    for (int q = s - 1; q >= 0; --q) {
        const int v = s_last[q][tx];
        if (v >= 0) { above = v; break; }
    }
    for (int q = s + 1; q < FC_STRIPS; ++q) {
        const int v = s_first[q][tx];
        if (v >= 0) { below = v; break; }
    }
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
        if (!m[(size_t)r * W]) above = r;
        o[(size_t)r * W] = above;
    }
#pragma unroll 4
    for (int r = r1 - 1; r >= r0; --r) {
        if (!m[(size_t)r * W]) below = r;
        const int a = o[(size_t)r * W];
        int best;
        if (a < 0) best = below;
        else if (below < 0) best = a;
        else best = (r - a <= below - r) ? a : below;      // equidistant: the smaller row
        o[(size_t)r * W] = best;
    }
}

__global__ void __launch_bounds__(FILL_THREADS)
k_fill_rows(const float* __restrict__ depth, const int32_t* __restrict__ near_row, int H, int W,
            float* __restrict__ out, int32_t* __restrict__ out_idx) {
    extern __shared__ int32_t s_dv2[];             // [W] squared vertical distance to the column's nearest valid pixel
    const int r = blockIdx.x;
    const size_t frame = (size_t)blockIdx.y * H * W;
    const int32_t* nr = near_row + frame + (size_t)r * W;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        const int a = nr[c];
        const int d = a - r;
        s_dv2[c] = (a < 0) ? FILL_BIG : d * d;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        int best = s_dv2[c], bc = c;
        for (int k = 1; k < W; ++k) {
            const int kk = k * k;
            if (kk > best) break;                  // a tie at distance k^2 is still possible when kk == best
            const int cl = c - k, cr = c + k;
            if (cl >= 0) {
                const int d2 = kk + s_dv2[cl];
                if (d2 <= best) { best = d2; bc = cl; }            // cl < every column seen so far: it wins ties
            }
            if (cr < W) {
                const int d2 = kk + s_dv2[cr];
                if (d2 < best) { best = d2; bc = cr; }             // cr > every column seen so far: it loses ties
            }
        }
        int sr, sc;
        if (best >= FILL_BIG) { sr = H - 1; sc = 0; }              // no valid pixel in the frame (scipy: index (-1, 0))
        else { sr = nr[bc]; sc = bc; }
        const size_t p = frame + (size_t)r * W + c;
        out[p] = depth[frame + (size_t)sr * W + sc];
        if (out_idx) {
            out_idx[2 * frame + (size_t)r * W + c] = (best >= FILL_BIG) ? -1 : sr;
            out_idx[2 * frame + (size_t)(H + r) * W + c] = sc;
        }
    }
}

extern "C" int spb_fill_nearest(const float* depth, const uint8_t* invalid, int n_frames, int H, int W, int32_t* near_row,
                                float* out, int32_t* out_idx, void* stream) {
    if (!depth || !invalid || !near_row || !out || n_frames < 0 || H < 1 || W < 1) return SPB_EINVAL;
    if (H > 32767 || W > 32767 || n_frames > 65535) return SPB_ELIMIT;
    if (n_frames == 0) return SPB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)W * sizeof(int32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fill_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k_fill_cols<<<dim3((W + FC_COLS - 1) / FC_COLS, n_frames), dim3(FC_COLS, FC_STRIPS), 0, st>>>(invalid, H, W, near_row);
    SPB_CHECK_LAUNCH();
    k_fill_rows<<<dim3(H, n_frames), FILL_THREADS, smem, st>>>(depth, near_row, H, W, out, out_idx);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
