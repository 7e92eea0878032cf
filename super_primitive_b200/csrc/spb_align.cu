// Fused per-iteration kernels of the dense photometric alignment path (sm_100a).
//
// One launch evaluates, for every point of every (source geometry, target image) pair:
//   depth = exp(L + k_b - L_kp)           core/dense_optim.py:38-86
//   X = unproject(u, v, depth)            core/dense_optim.py:19-35
//   Y = R X + t                           core/dense_optim.py:117-122, core/ops.py:5-17
//   (u', v') guarded pinhole projection   core/ops.py:19-40
//   normalise by geometry dims, validity  tool/point_utils.py:31-35, core/dense_optim.py:128-162
//   bilinear sample (align_corners, zeros padding) of the level image   core/dense_optim.py:134-136
//   affine brightness, masked residual, L1 mean                          core/dense_optim.py:202-261
// and, in the same pass, either the first-order gradient w.r.t. (pose 4x4, k, affine) that the
// reference obtains by autograd (MODE_GRAD), or the IRLS Gauss-Newton arrowhead blocks (MODE_GN).
//
// Layout: points are stored compacted in (segment,row,col) order; a warp processes one tile
// (<= SPB_TILE points of ONE segment), each lane SPB_PPT points strided by 32 so the streaming
// reads (uv, logd, cached source rgb) are fully coalesced.  The target image is RGBA-interleaved
// so one bilinear tap is one 16-byte load.  Reductions are deterministic: per-warp shuffle ->
// per-CTA partial -> fixed-order finalize kernel; no float atomics.
#include "spb_common.cuh"

enum { MODE_GRAD = 0, MODE_GN = 1 };

struct PairPack {        // up to 16 pair descriptors passed by value (Python per-call path)
    SpbPair p[16];
};

template <int NACC>
__device__ __forceinline__ void block_reduce_store(float (&acc)[NACC], float* s_red, float* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) s_red[warp * NACC + i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < SPB_WARPS; ++w) v += s_red[w * NACC + i];
        dst[i] = v;
    }
}

struct PointOut {   // where the optional per-point statistics go
    const SpbStats* st;
    int pair;
    int n;
};

// ------------------------------------------------------------------------------------------------
// The per-point evaluation shared by every variant.
//   GEOMETRY: X, RX(+t)=Y, projection, validity.  Then taps + residual.
// Accumulators:
//   MODE_GRAD: acc[16] = cost, gt[3], gR[9], ga, gb, nvalid ; seg[1] = gk
//   MODE_GN  : acc[NP*(NP+1)/2 + NP + 3] ; seg[NP + 2]
// ------------------------------------------------------------------------------------------------
template <int MODE, int NP, bool STATS, int NACC, int NSEG>
__device__ __forceinline__ void eval_point(const float* __restrict__ c, const float4* __restrict__ trg,
                                           int Wl, int Hl, float Xx, float Xy, float Xz, bool sok,
                                           float Is0, float Is1, float Is2, float irls_eps,
                                           float (&acc)[NACC], float (&seg)[NSEG],
                                           const PointOut& po, int pidx) {
    const float RXx = fmaf(c[C_R + 0], Xx, fmaf(c[C_R + 1], Xy, c[C_R + 2] * Xz));
    const float RXy = fmaf(c[C_R + 3], Xx, fmaf(c[C_R + 4], Xy, c[C_R + 5] * Xz));
    const float RXz = fmaf(c[C_R + 6], Xx, fmaf(c[C_R + 7], Xy, c[C_R + 8] * Xz));
    const float Yx = RXx + c[C_T + 0];
    const float Yy = RXy + c[C_T + 1];
    const float Yz = RXz + c[C_T + 2];
    const bool live = fabsf(Yz) > 1e-6f;                       // guarded reciprocal, core/ops.py:22,33-34
    const float zi = live ? __fdividef(1.0f, Yz) : 1e-6f;     // MUFU.RCP, <= 1 ulp
    const float up = fmaf(Yx * c[C_FXT], zi, c[C_CXT]);
    const float vp = fmaf(Yy * c[C_FYT], zi, c[C_CYT]);
    const float xn = fmaf(up, c[C_TIW], -1.0f);
    const float yn = fmaf(vp, c[C_TIH], -1.0f);
    const bool tok = (fabsf(xn) <= 0.99f) && (fabsf(yn) <= 0.99f) && (Yz > c[C_TAU]);
    const bool m = tok && sok;
    const float ix = (xn + 1.0f) * c[C_SX];
    const float iy = (yn + 1.0f) * c[C_SY];

    if constexpr (STATS) {
        // slow path: materialise the reference's per-point statistics (core/dense_optim.py:347-361)
        const SpbStats& st = *po.st;
        const size_t n = (size_t)po.n, j = (size_t)po.pair;
        if (st.moved_pts) {
            float* o = st.moved_pts + (j * n + pidx) * 3;
            o[0] = Yx; o[1] = Yy; o[2] = Yz;
        }
        if (st.trg_ok) st.trg_ok[j * n + pidx] = tok ? 1 : 0;
        if (st.full_mask) st.full_mask[j * n + pidx] = m ? 1 : 0;
        float v[3] = {0.f, 0.f, 0.f};
        if (isfinite(ix) && isfinite(iy)) {
            const float fxf = floorf(ix), fyf = floorf(iy);
            // clamp before the int conversion so wild coordinates cannot overflow
            const int x0 = (int)fminf(fmaxf(fxf, -2.0f), (float)Wl + 1.0f);
            const int y0 = (int)fminf(fmaxf(fyf, -2.0f), (float)Hl + 1.0f);
            const float fx = ix - fxf, fy = iy - fyf;
            const float4 nw = tap_rgba(trg, x0, y0, Wl, Hl), ne = tap_rgba(trg, x0 + 1, y0, Wl, Hl);
            const float4 sw = tap_rgba(trg, x0, y0 + 1, Wl, Hl), se = tap_rgba(trg, x0 + 1, y0 + 1, Wl, Hl);
            float d0, d1;
            blend(nw.x, ne.x, sw.x, se.x, fx, fy, v[0], d0, d1);
            blend(nw.y, ne.y, sw.y, se.y, fx, fy, v[1], d0, d1);
            blend(nw.z, ne.z, sw.z, se.z, fx, fy, v[2], d0, d1);
        }
        const float Is[3] = {Is0, Is1, Is2};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float Ia = fmaf(c[C_EA], v[ch], c[C_BB]);
            if (st.trg_px) st.trg_px[(j * 3 + ch) * n + pidx] = Ia;
            if (st.residual_raw) st.residual_raw[(j * 3 + ch) * n + pidx] = m ? (Is[ch] - Ia) : 0.0f;
        }
    }

    if (!m) return;

    // fast path: valid => |xn|,|yn| <= 0.99 => all four taps are inside the image
    const float fxf = floorf(ix), fyf = floorf(iy);
    const int x0 = (int)fxf, y0 = (int)fyf;
    const float fx = ix - fxf, fy = iy - fyf;
    const float4* p0 = trg + (y0 * Wl + x0);                   // 32-bit texel index (image < 2^31 texels)
    const float4 nw = __ldg(p0), ne = __ldg(p0 + 1);
    const float4 sw = __ldg(p0 + Wl), se = __ldg(p0 + Wl + 1);

    float I[3], dx[3], dy[3];
    blend(nw.x, ne.x, sw.x, se.x, fx, fy, I[0], dx[0], dy[0]);
    blend(nw.y, ne.y, sw.y, se.y, fx, fy, I[1], dx[1], dy[1]);
    blend(nw.z, ne.z, sw.z, se.z, fx, fy, I[2], dx[2], dy[2]);
    const float Is[3] = {Is0, Is1, Is2};
    const float ea = c[C_EA], bb = c[C_BB];

    if constexpr (MODE == MODE_GRAD) {
        // d|r|/dI_t = -e^{-a} sign(r);  accumulate sign-weighted image slopes first
        float gx = 0.f, gy = 0.f, ga = 0.f, gb = 0.f, cost = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float r = Is[ch] - fmaf(ea, I[ch], bb);
            cost += fabsf(r);
            const float s = (r > 0.f) ? 1.0f : ((r < 0.f) ? -1.0f : 0.0f);
            gx = fmaf(s, dx[ch], gx);
            gy = fmaf(s, dy[ch], gy);
            ga = fmaf(s, I[ch], ga);
            gb += s;
        }
        const float gu = -ea * c[C_KX] * gx;          // d cost / d u'
        const float gv = -ea * c[C_KY] * gy;
        const float gYx = gu * c[C_FXT] * zi;
        const float gYy = gv * c[C_FYT] * zi;
        const float gYz = live ? -(gYx * Yx + gYy * Yy) * zi : 0.0f;
        acc[0] += cost;
        acc[1] += gYx; acc[2] += gYy; acc[3] += gYz;
        acc[4] = fmaf(gYx, Xx, acc[4]);  acc[5] = fmaf(gYx, Xy, acc[5]);  acc[6] = fmaf(gYx, Xz, acc[6]);
        acc[7] = fmaf(gYy, Xx, acc[7]);  acc[8] = fmaf(gYy, Xy, acc[8]);  acc[9] = fmaf(gYy, Xz, acc[9]);
        acc[10] = fmaf(gYz, Xx, acc[10]); acc[11] = fmaf(gYz, Xy, acc[11]); acc[12] = fmaf(gYz, Xz, acc[12]);
        acc[13] = fmaf(ea, ga, acc[13]);
        acc[14] -= gb;
        acc[15] += 1.0f;
        if constexpr (NSEG > 0) seg[0] += fmaf(gYx, RXx, fmaf(gYy, RXy, gYz * RXz));
    } else {
        // IRLS normal equations: r_c = I_s - (e^{-a} I_c + b); J_c = -(e^{-a}) (dIx kx du' + dIy ky dv')
        float Guu = 0.f, Guv = 0.f, Gvv = 0.f, hu = 0.f, hv = 0.f, cost = 0.f, wcost = 0.f;
        float Aau = 0.f, Aav = 0.f, Abu = 0.f, Abv = 0.f, Aaa = 0.f, Aab = 0.f, Abb = 0.f, ha = 0.f, hb = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float r = Is[ch] - fmaf(ea, I[ch], bb);
            const float ar = fabsf(r);
            const float w = __fdividef(1.0f, fmaxf(ar, irls_eps));
            cost += ar;
            wcost = fmaf(w * r, r, wcost);
            const float wu = w * dx[ch], wv = w * dy[ch];
            Guu = fmaf(wu, dx[ch], Guu);
            Guv = fmaf(wu, dy[ch], Guv);
            Gvv = fmaf(wv, dy[ch], Gvv);
            hu = fmaf(wu, r, hu);
            hv = fmaf(wv, r, hv);
            if constexpr (NP == 8) {
                const float ja = ea * I[ch];      // d r / d a_t ; d r / d b_t = -1
                Aau = fmaf(wu, ja, Aau); Aav = fmaf(wv, ja, Aav);
                Abu -= wu; Abv -= wv;
                Aaa = fmaf(w * ja, ja, Aaa); Aab = fmaf(-w, ja, Aab); Abb += w;
                ha = fmaf(w * ja, r, ha); hb = fmaf(-w, r, hb);
            }
        }
        const float cu = -ea * c[C_KX], cv = -ea * c[C_KY];
        Guu *= cu * cu; Guv *= cu * cv; Gvv *= cv * cv; hu *= cu; hv *= cv;
        // mu = d u'/d(xi,k), mv = d v'/d(xi,k); xi = (tau, phi), left perturbation T <- Exp(xi) T
        const float xb = Yx * zi, yb = Yy * zi;
        const float zl = live ? 1.0f : 0.0f;
        const float fu = c[C_FXT] * zi, fv = c[C_FYT] * zi;
        float mu[7], mv[7];
        mu[0] = fu;  mu[1] = 0.f; mu[2] = -fu * xb * zl;
        mu[3] = mu[2] * Yy;               mu[4] = fu * Yz - mu[2] * Yx;   mu[5] = -fu * Yy;
        mv[0] = 0.f; mv[1] = fv;  mv[2] = -fv * yb * zl;
        mv[3] = -fv * Yz + mv[2] * Yy;    mv[4] = -mv[2] * Yx;            mv[5] = fv * Yx;
        mu[6] = fmaf(mu[0], RXx, mu[2] * RXz);
        mv[6] = fmaf(mv[1], RXy, mv[2] * RXz);
        float pu[7], pv[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            pu[i] = fmaf(Guu, mu[i], Guv * mv[i]);
            pv[i] = fmaf(Guv, mu[i], Gvv * mv[i]);
        }
        // upper triangle of the NPxNP pose block, row-major packed
        int q = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) { acc[q] = fmaf(mu[i], pu[j], fmaf(mv[i], pv[j], acc[q])); ++q; }
            if constexpr (NP == 8) {
                acc[q] = fmaf(mu[i], cu * Aau, fmaf(mv[i], cv * Aav, acc[q])); ++q;
                acc[q] = fmaf(mu[i], cu * Abu, fmaf(mv[i], cv * Abv, acc[q])); ++q;
            }
        }
        if constexpr (NP == 8) { acc[q++] += Aaa; acc[q++] += Aab; acc[q++] += Abb; }
        constexpr int NA = NP * (NP + 1) / 2;
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[NA + i] = fmaf(mu[i], hu, fmaf(mv[i], hv, acc[NA + i]));
        if constexpr (NP == 8) { acc[NA + 6] += ha; acc[NA + 7] += hb; }
        acc[NA + NP + 0] += cost;
        acc[NA + NP + 1] += wcost;
        acc[NA + NP + 2] += 1.0f;
        if constexpr (NSEG > 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) seg[i] = fmaf(mu[i], pu[6], fmaf(mv[i], pv[6], seg[i]));
            if constexpr (NP == 8) {
                seg[6] = fmaf(mu[6], cu * Aau, fmaf(mv[6], cv * Aav, seg[6]));
                seg[7] = fmaf(mu[6], cu * Abu, fmaf(mv[6], cv * Abv, seg[7]));
            }
            seg[NP] = fmaf(mu[6], pu[6], fmaf(mv[6], pv[6], seg[NP]));
            seg[NP + 1] = fmaf(mu[6], hu, fmaf(mv[6], hv, seg[NP + 1]));
        }
    }
}

// per-CTA accumulators (NACC) and per-run partials (NSEG) of each kernel variant.  The 6-column GN variant keeps rows
// 0..2 of the pose block + g_0..g_3 per run and derives the depth column from them (spb_gn_packed.cuh).
template <int MODE, int NP>
struct Sizes {
    static constexpr int NACC = (MODE == MODE_GRAD) ? SPB_PAIR_NOUT : (NP == 6 ? 12 : NP * (NP + 1) / 2 + NP + 3);
    static constexpr int NSEG = (MODE == MODE_GRAD) ? 1 : (NP == 6 ? 19 : NP + 2);
};
#define SPB_MAX_NACC 47        // over the variants: 8-column GN
#define SPB_MAX_NSEG 19        // 6-column GN

// ------------------------------------------------------------------------------------------------
// Compact-geometry variant.  grid = (ctas_per_pair, n_pairs)
//   part_pair : [pair][cta][NACC]      part_seg : [pair][tile][NSEG]
// ------------------------------------------------------------------------------------------------
template <int MODE, int NP, bool STATS>
__device__ __forceinline__ void align_body(const SpbGeom& g, const SpbPair& pr, int pair, float irls_eps,
                                           float* __restrict__ part_pair, float* __restrict__ part_seg,
                                           const SpbStats* stats) {
    constexpr int NACC = Sizes<MODE, NP>::NACC;
    constexpr int NSEG = Sizes<MODE, NP>::NSEG;
    __shared__ float s_ctx[C_N];
    __shared__ float s_red[SPB_WARPS * NACC];
    if (threadIdx.x < 32) fill_ctx(s_ctx, pr, g.K, g.H, g.W);
    __syncthreads();
    const float* c = s_ctx;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* trg = reinterpret_cast<const float4*>(pr.trg_rgba);
    const int Wl = pr.Wl, Hl = pr.Hl;
    const float* sr0 = pr.src_rgb;
    const float* sr1 = pr.src_rgb + g.n_pad;
    const float* sr2 = pr.src_rgb + 2 * (size_t)g.n_pad;
    const int4* tiles = reinterpret_cast<const int4*>(g.tiles);
    PointOut po{stats, pair, g.n_pts};

    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

    const int wstride = gridDim.x * SPB_WARPS;
    for (int t = blockIdx.x * SPB_WARPS + warp; t < g.n_tiles; t += wstride) {
        const int4 td = __ldg(tiles + t);
        const int sidx = td.x, start = td.y, cnt = td.z;
        const float shift = __ldg(pr.k + sidx) - __ldg(g.seg_lkp + sidx);
        float seg[NSEG];
#pragma unroll
        for (int i = 0; i < NSEG; ++i) seg[i] = 0.f;

        // issue all streaming loads of this lane's points first (memory-level parallelism)
        uint32_t w[SPB_PPT];
        float L[SPB_PPT], s0[SPB_PPT], s1[SPB_PPT], s2[SPB_PPT];
#pragma unroll
        for (int j = 0; j < SPB_PPT; ++j) {
            const int i = j * 32 + lane;
            const bool on = i < cnt;
            const int p = start + (on ? i : 0);
            w[j] = __ldg(g.uv + p);
            L[j] = __ldg(g.logd + p);
            s0[j] = __ldg(sr0 + p); s1[j] = __ldg(sr1 + p); s2[j] = __ldg(sr2 + p);
        }
#pragma unroll
        for (int j = 0; j < SPB_PPT; ++j) {
            const int i = j * 32 + lane;
            if (i < cnt) {
                const float u = (float)(w[j] & 0xffffu);
                const float v = (float)((w[j] >> 16) & 0x7fffu);
                const float z = expf(L[j] + shift);
                const float Xx = (u - c[C_CX]) * z * c[C_IFX];
                const float Xy = (v - c[C_CY]) * z * c[C_IFY];
                const bool sok = (w[j] >> 31) && (z > 1e-7f);
                if constexpr (STATS) {
                    if (pair == 0 && stats->src_pts) {
                        float* o = stats->src_pts + (size_t)(td.w + i) * 3;
                        o[0] = Xx; o[1] = Xy; o[2] = z;
                    }
                }
                eval_point<MODE, NP, STATS, NACC, NSEG>(c, trg, Wl, Hl, Xx, Xy, z, sok, s0[j], s1[j], s2[j],
                                                        irls_eps, acc, seg, po, td.w + i);
            }
        }
#pragma unroll
        for (int i = 0; i < NSEG; ++i) {
            const float v = warp_sum(seg[i]);
            if (lane == 0) part_seg[(size_t)t * NSEG + i] = v;
        }
    }
    block_reduce_store<NACC>(acc, s_red, part_pair + (size_t)blockIdx.x * NACC);
}


// ------------------------------------------------------------------------------------------------
// Multi-value warp reduction: N per-lane values -> N warp sums with ~N+2 shuffles instead of 5N.
// Values are folded in chunks of K = 16 / 8 / 4 with a butterfly that halves the value set at every exchange
// (K/2 + K/4 + ... + 1 shuffles), the remaining log2(32/K) exchanges finish the sums; the total of value i of a chunk
// ends up in lane i * (32/K).  Fixed exchange order => deterministic.
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void butterfly_store(const float* v, float* __restrict__ dst, int lane) {
    static_assert(K == 16 || K == 8 || K == 4, "chunk size");
    float a[K];
#pragma unroll
    for (int i = 0; i < K; ++i) a[i] = v[i];
    int n = K;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        if (n > 1) {
            const bool up = lane & o;
            n >>= 1;
#pragma unroll
            for (int i = 0; i < K / 2; ++i) {
                if (i < n) {
                    const float send = up ? a[i] : a[i + n];
                    const float keep = up ? a[i + n] : a[i];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
        } else {
            a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
        }
    }
    constexpr int STEP = 32 / K;
    if ((lane & (STEP - 1)) == 0) dst[lane / STEP] = a[0];
}

template <int N>
__device__ __forceinline__ void tile_reduce_store(float (&v)[N], float* __restrict__ dst, int lane) {
    int done = 0;
    if constexpr (N >= 16) {
        butterfly_store<16>(v, dst, lane);
        done = 16;
    } else if constexpr (N >= 8) {
        butterfly_store<8>(v, dst, lane);
        done = 8;
    }
    if constexpr (N - (N >= 16 ? 16 : (N >= 8 ? 8 : 0)) >= 3) {
        constexpr int D0 = N >= 16 ? 16 : (N >= 8 ? 8 : 0);
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (D0 + i < N) ? v[D0 + i] : 0.f;
        float tmp[4];
        // totals of w[i] land in lane 8 i; store only the real ones
        {
            float a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3];
            const bool up16 = lane & 16, up8 = lane & 8;
            const float s0 = up16 ? a0 : a2, k0 = up16 ? a2 : a0;
            const float s1 = up16 ? a1 : a3, k1 = up16 ? a3 : a1;
            a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
            a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            const float s2 = up8 ? a0 : a1, k2 = up8 ? a1 : a0;
            a0 = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 4);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
            tmp[0] = a0;
        }
        const int vi = lane >> 3;
        if ((lane & 7) == 0 && D0 + vi < N) dst[D0 + vi] = tmp[0];
        done = D0 + 4;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i >= done) {
            const float t = warp_sum(v[i]);
            if (lane == 0) dst[i] = t;
        }
    }
}

#include "spb_fast.cuh"
#include "spb_gn_packed.cuh"
#include "spb_lm.cuh"
#include "spb_adam.cuh"

#ifndef SPB_UNROLL
#define SPB_UNROLL 1                           // unroll factor of the per-lane point loop
#endif
#define SPB_DO_PRAGMA(x) _Pragma(#x)
#define SPB_PRAGMA_UNROLL(n) SPB_DO_PRAGMA(unroll n)

// ------------------------------------------------------------------------------------------------
// Hot-path body: per-warp bulk-async pipelines (see spb_fast.cuh).  grid = (ctas_per_pair, pairs)
//   part_pair : [cta][NACC]      part_seg : [tile][NSEG]
// The warps of a CTA take adjacent tiles (their target footprints share L1 lines), CTAs stride over the pair.
// Measured and dropped in round 2 (profiles/README.md, visits r02a, r02f): groups of 2-8 consecutive tiles per warp with
// one partial-sum reduction per same-segment run (GN unchanged, gradient kernel slower: L1 locality); touching the
// next point's predicted target row ahead of time with a 4-byte cp.async (the gathers already keep the L1 data path
// 57 % busy; one more L1 access per point costs more than the latency it hides: l1tex 87 %, kernel 12 % slower); the
// per-pair context in a __constant__ array indexed by blockIdx.y (nvcc emits vector-indexed LDC, which is slower than
// the broadcast LDS it replaces: GN 0.430 vs 0.442, first-order 0.474 vs 0.599); staging only uv + logd in shared memory
// and reading the cached source colours straight from global memory so that a tile can hold 256 points (r02g: GN
// 0.444 = unchanged, first-order 0.461 vs 0.601: the stream's DRAM latency is no longer hidden by the bulk copy); the
// first-order kernel at 3 CTAs/SM, with or without the projection context hoisted into registers (r02h: 0.558 / 0.544).
// ------------------------------------------------------------------------------------------------
template <int MODE, int NP, bool AFF>
__device__ __forceinline__ void align_body_warp(const SpbGeom& g, const SpbPair& pr, float irls_eps,
                                                float* __restrict__ part_pair, float* __restrict__ part_seg) {
    constexpr int NACC = Sizes<MODE, NP>::NACC;
    constexpr int NSEG = Sizes<MODE, NP>::NSEG;
    extern __shared__ __align__(128) uint32_t s_dyn[];
    __shared__ __align__(16) float s_ctx[F_N];
    __shared__ float s_shift[SPB_NSHIFT];
    __shared__ float s_red[SPB_WARPS * NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* ring = s_dyn + warp * (SPB_WSTAGES * SPB_SLOT_WORDS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + SPB_WARPS * SPB_WSTAGES * SPB_SLOT_WORDS) + warp * SPB_WSTAGES;

    if (threadIdx.x < 32) fill_fast_ctx(s_ctx, pr, g.K, g.H, g.W);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SPB_WSTAGES; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int nshift = min(g.n_seg, SPB_NSHIFT);
    for (int b = threadIdx.x; b < nshift; b += blockDim.x) s_shift[b] = __ldg(pr.k + b) - __ldg(g.seg_lkp + b);
    __syncthreads();
    const float* c = s_ctx;

    const int WS = gridDim.x * SPB_WARPS;                  // tile stride of this warp
    const int t_first = blockIdx.x * SPB_WARPS + warp;
    const int t_end = g.n_tiles;
    const uint32_t* pack = pr.tile_pack;

    // producer (lane 0): ONE bulk copy brings the whole tile block (header + uv + logd + r + g + b)
    auto issue = [&](int t, int slot) {
        const uint32_t bar = smem_u32(bars + slot);
        mbar_expect_tx(bar, SPB_SLOT_WORDS * 4u);
        bulk_g2s(smem_u32(ring + slot * SPB_SLOT_WORDS), pack + (size_t)t * SPB_PACK_WORDS, SPB_SLOT_WORDS * 4u, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SPB_WSTAGES - 1; ++s) {
            const int t = t_first + s * WS;
            if (t < t_end) issue(t, s);
        }
    }

    const float4* trg = reinterpret_cast<const float4*>(pr.trg_rgba);
    const int Wl = pr.Wl;
    constexpr bool PACKED = (MODE == MODE_GN && NP == 6);   // FFMA2 formulation (spb_gn_packed.cuh)
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    GnAcc6 pacc;
    pacc.zero();
    GradAcc gacc;
    gacc.zero();

    int slot = 0, fill = SPB_WSTAGES - 1;                  // slot consumed now / slot refilled now
    uint32_t phase = 0;
    for (int t = t_first; t < t_end; t += WS) {
        if (lane == 0) {
            const int tn = t + (SPB_WSTAGES - 1) * WS;
            if (tn < t_end) issue(tn, fill);
        }
        mbar_wait(smem_u32(bars + slot), phase);
        const uint32_t* sl = ring + slot * SPB_SLOT_WORDS;
        const int sidx = (int)sl[0];
        const float shift = (sidx < SPB_NSHIFT) ? s_shift[sidx] : (__ldg(pr.k + sidx) - __ldg(g.seg_lkp + sidx));
        const uint32_t* s_uv = sl + 4;
        const float* s_f = reinterpret_cast<const float*>(sl + 4);
        float seg[NSEG];
#pragma unroll
        for (int i = 0; i < NSEG; ++i) seg[i] = 0.f;
        GnSeg6 pseg;
        pseg.zero();
        // padding entries of a partial tile are zero words: uv bit 31 clear => invalid, no bounds test needed
        SPB_PRAGMA_UNROLL(SPB_UNROLL)
        for (int j = 0; j < SPB_PPT; ++j) {
            const int i = j * 32 + lane;
            Proj q;
            bool ok = project_point(c, s_uv[i], s_f[SPB_TILE + i], shift, Wl, q);
            if constexpr (PACKED) ok = ok && q.live;       // see point_gn6_packed
            if (ok) {
                Taps4 tp;
                load_taps(trg, Wl, q.off, tp);
                const float i0 = s_f[2 * SPB_TILE + i], i1 = s_f[3 * SPB_TILE + i], i2 = s_f[4 * SPB_TILE + i];
                if constexpr (MODE == MODE_GRAD)
                    point_grad_packed<AFF>(c, tp, q, i0, i1, i2, gacc, seg[0]);
                else if constexpr (PACKED)
                    point_gn6_packed<AFF>(c, tp, q, i0, i1, i2, irls_eps, pacc, pseg);
                else
                    point_gn<NP, NACC, NSEG>(c, tp, q, i0, i1, i2, irls_eps, acc, seg);
            }
        }
        if constexpr (PACKED) pseg.store(seg);
        tile_reduce_store<NSEG>(seg, part_seg + (size_t)t * NSEG, lane);
        __syncwarp();                                      // every lane is done with this slot
        fill = slot;
        if (++slot == SPB_WSTAGES) { slot = 0; phase ^= 1u; }
    }
    if constexpr (PACKED) pacc.store(acc);
    if constexpr (MODE == MODE_GRAD) gacc.store(acc);
    __syncthreads();
    block_reduce_store<NACC>(acc, s_red, part_pair + (size_t)blockIdx.x * NACC);
}


// occupancy target: 3 CTAs/SM (<= 80 registers) for the gradient kernel, 2 for the GN kernel whose
// 38 accumulators do not fit 80 registers without spilling
#ifndef SPB_OCC_GRAD
#define SPB_OCC_GRAD 4                          // CTAs/SM the gradient kernel is compiled for (64 registers)
#endif
#ifndef SPB_OCC_GN
#define SPB_OCC_GN 3                            // 6-column GN kernel: 80 registers
#endif
template <int MODE, int NP>
struct Occ {   // the 8-column (affine) GN variant keeps 47 accumulators: 2 CTAs/SM, no spills
    static constexpr int CTAS = (MODE == MODE_GRAD) ? SPB_OCC_GRAD : (NP == 8 ? 2 : SPB_OCC_GN);
};

// B pairs over one geometry, descriptors by value (Python per-call path: no descriptor upload)
template <int MODE, int NP, bool AFF>
__global__ void __launch_bounds__(SPB_THREADS, Occ<MODE, NP>::CTAS)
k_align_inline(const __grid_constant__ SpbGeom g, const __grid_constant__ PairPack pack, float irls_eps,
               float* __restrict__ work) {
    constexpr int NACC = Sizes<MODE, NP>::NACC;
    constexpr int NSEG = Sizes<MODE, NP>::NSEG;
    const int pair = blockIdx.y;
    const size_t stride = (size_t)gridDim.x * NACC + (size_t)g.n_tiles * NSEG;
    float* base = work + pair * stride;
    align_body_warp<MODE, NP, AFF>(g, pack.p[pair], irls_eps, base, base + (size_t)gridDim.x * NACC);
}

// same, slow path that also materialises the per-point statistics (register-prefetch body)
__global__ void __launch_bounds__(SPB_THREADS, 2)
k_align_stats(const __grid_constant__ SpbGeom g, const __grid_constant__ PairPack pack, float* __restrict__ work,
              SpbStats stats) {
    constexpr int NACC = Sizes<MODE_GRAD, 6>::NACC;
    constexpr int NSEG = Sizes<MODE_GRAD, 6>::NSEG;
    const int pair = blockIdx.y;
    const size_t stride = (size_t)gridDim.x * NACC + (size_t)g.n_tiles * NSEG;
    float* base = work + pair * stride;
    align_body<MODE_GRAD, 6, true>(g, pack.p[pair], pair, 0.f, base, base + (size_t)gridDim.x * NACC, &stats);
}

// n_pairs independent problems, descriptors in device memory (batched solver / benchmark path)
template <int MODE, int NP, bool AFF>
__global__ void __launch_bounds__(SPB_THREADS, Occ<MODE, NP>::CTAS)
k_align_global(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs, float irls_eps,
               float* __restrict__ work, int64_t work_stride) {
    constexpr int NACC = Sizes<MODE, NP>::NACC;
    const int pair = blockIdx.y;
    __shared__ SpbPair s_pr;
    __shared__ SpbGeom s_g;
    if (threadIdx.x == 0) {
        s_pr = pairs[pair];
        s_g = geoms[s_pr.geom];
    }
    __syncthreads();
    float* base = work + pair * work_stride;
    align_body_warp<MODE, NP, AFF>(s_g, s_pr, irls_eps, base, base + (size_t)gridDim.x * NACC);
}


// Fixed-order sum over the per-CTA partials [ctas][NACC] (contiguous): the block's threads form R = blockDim / NACC
// row groups; thread (r, c) adds rows r, r + R, ... of column c (consecutive threads read consecutive floats, 8 loads
// in flight), the groups meet in shared memory and thread c < NACC adds them in group order.  Returns the total to
// threads < NACC (0 elsewhere).  A serial loop over several hundred CTAs costs tens of microseconds when one
// problem owns the whole GPU.  Needs blockDim >= NACC; contains two __syncthreads().
#define SPB_FIN_SMEM 1024
#ifndef SPB_FIN_THREADS
#define SPB_FIN_THREADS 512                     // threads of k_gn_finalize_solve
#endif
#define SPB_FIN_THREADS_MAX 512                 // most threads any GN finalize launch uses
template <int NACC>
__device__ __forceinline__ float sum_cta_partials(const float* pp, int ctas, float* s_part /* [SPB_FIN_SMEM] */) {
    const int R = min((int)blockDim.x, SPB_FIN_SMEM) / NACC;
    const int r = threadIdx.x / NACC, c = threadIdx.x - r * NACC;
    float v = 0.f;
    if (r < R) {
        constexpr int U = 8;
        for (int row0 = r; row0 < ctas; row0 += U * R) {
            float b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * R;
                b[u] = (row < ctas) ? pp[(size_t)row * NACC + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) v += b[u];
        }
        s_part[r * NACC + c] = v;
    }
    __syncthreads();
    float tot = 0.f;
    if (threadIdx.x < NACC) {
        const int groups = min(R, ctas);
        for (int q = 0; q < groups; ++q) tot += s_part[q * NACC + threadIdx.x];
    }
    __syncthreads();
    return tot;
}

// Fixed-order sum of the per-tile partials of TWO segments at once (tiles [ta0,ta1) and [tb0,tb1), NSEG floats per
// tile, contiguous).  Lanes walk the contiguous float range coalesced: lane l < G*NSEG always meets value l % NSEG
// (G = 32 / NSEG tiles per round), so one register per segment accumulates; the G lanes holding the same value
// are then combined in fixed order.  Result valid in lanes < NSEG.  Two segments interleaved = two independent
// load streams in flight per warp (the finalize kernels are latency-bound: one CTA per problem).
template <int NSEG>
__device__ __forceinline__ void seg_sum2(const float* ps, int ta0, int ta1, int tb0, int tb1, int lane, float& ra,
                                         float& rb) {
    constexpr int G = 32 / NSEG, STRIDE = G * NSEG;
    constexpr int U = 16;                       // loads in flight per lane and segment (the kernel is latency-bound)
    float va = 0.f, vb = 0.f;
    if (lane < STRIDE) {
        const float* pa = ps + (size_t)ta0 * NSEG + lane;
        const float* pb = ps + (size_t)tb0 * NSEG + lane;
        const int na = (ta1 - ta0) * NSEG - lane, nb = (tb1 - tb0) * NSEG - lane;   // elements left for this lane
        const int nmax = max(na, nb);
        for (int e0 = 0; e0 < nmax; e0 += U * STRIDE) {
            float ba[U], bb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * STRIDE;
                ba[u] = (e < na) ? pa[e] : 0.f;
                bb[u] = (e < nb) ? pb[e] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                va += ba[u];
                vb += bb[u];
            }
        }
    }
    if constexpr (NSEG == 1) {
        ra = warp_sum(va);
        rb = warp_sum(vb);
    } else {
        ra = va;
        rb = vb;
#pragma unroll
        for (int gi = 1; gi < G; ++gi) {
            ra += __shfl_sync(0xffffffffu, va, (lane + gi * NSEG) & 31);
            rb += __shfl_sync(0xffffffffu, vb, (lane + gi * NSEG) & 31);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// finalize: fixed-order reduction of the partials
// ------------------------------------------------------------------------------------------------
// `scale_cols`: the hot path accumulates d cost / d M with M = R diag(1/fx, 1/fy, 1); d cost / d R scales
// the first two columns by 1/fx, 1/fy.  (The statistics path accumulates d cost / d R directly.)
__device__ __forceinline__ float grad_col_scale(int idx, const float* K, bool scale_cols) {
    if (!scale_cols || idx < 4 || idx > 12) return 1.0f;
    const int col = (idx - 4) % 3;
    return col == 0 ? 1.0f / K[0] : (col == 1 ? 1.0f / K[4] : 1.0f);
}

__global__ void k_finalize_grad(const __grid_constant__ SpbGeom g, const __grid_constant__ PairPack pack, int ctas,
                                int scale_cols, const float* __restrict__ work, float* __restrict__ out_pair,
                                float* __restrict__ out_gk, float* __restrict__ out_pose, float* __restrict__ out_flag) {
    const int pair = blockIdx.x;
    const size_t stride = (size_t)ctas * SPB_PAIR_NOUT + (size_t)g.n_tiles;
    const float* pp = work + pair * stride;
    const float* ps = pp + (size_t)ctas * SPB_PAIR_NOUT;
    const float norm = 1.0f / (3.0f * (float)g.n_pts);
    __shared__ float s_v[SPB_PAIR_NOUT];
    __shared__ float s_part[SPB_FIN_SMEM];
    bool finite = true;
    const float tot = sum_cta_partials<SPB_PAIR_NOUT>(pp, ctas, s_part);
    if (threadIdx.x < SPB_PAIR_NOUT) {
        float v = tot;
        v *= grad_col_scale(threadIdx.x, g.K, scale_cols != 0);
        v = (threadIdx.x == 15) ? v : v * norm;
        out_pair[pair * SPB_PAIR_NOUT + threadIdx.x] = v;
        s_v[threadIdx.x] = v;
        finite = isfinite(v) && isfinite(pack.p[pair].pose[threadIdx.x]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int b = warp; b < g.n_seg; b += 2 * nwarps) {
        const int b2 = b + nwarps;
        const bool two = b2 < g.n_seg;
        float v, v2;
        seg_sum2<1>(ps, g.seg_tile[b], g.seg_tile[b + 1], two ? g.seg_tile[b2] : 0, two ? g.seg_tile[b2 + 1] : 0, lane,
                    v, v2);
        v *= norm;
        v2 *= norm;
        if (lane == 0) {
            out_gk[(size_t)pair * g.n_seg + b] = v;
            if (two) out_gk[(size_t)pair * g.n_seg + b2] = v2;
        }
        finite = finite && isfinite(v) && isfinite(pack.p[pair].k[b]);
        if (two) finite = finite && isfinite(v2) && isfinite(pack.p[pair].k[b2]);
    }
    // the reference's finiteness asserts (core/dense_optim.py:44,78,311,321,340-343), folded into one flag per pair
    const int all_ok = __syncthreads_and(finite ? 1 : 0);
    if (out_flag && threadIdx.x == 0) out_flag[pair] = all_ok ? 1.0f : 0.0f;
    // d cost / d pose as the 4x4 the autograd boundary hands back (bottom row zero)
    if (out_pose && threadIdx.x < 16) {
        const int r = threadIdx.x >> 2, cc = threadIdx.x & 3;
        out_pose[pair * 16 + threadIdx.x] = (r == 3) ? 0.f : (cc == 3 ? s_v[1 + r] : s_v[4 + 3 * r + cc]);
    }
}

template <int NP>
__device__ __forceinline__ void finalize_gn_body(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs,
                                                 const int32_t* __restrict__ seg_off, int ctas,
                                                 const float* __restrict__ work, int64_t work_stride,
                                                 float* __restrict__ out_pair, float* __restrict__ out_seg) {
    constexpr int NACC = Sizes<MODE_GN, NP>::NACC;
    constexpr int NSEG = Sizes<MODE_GN, NP>::NSEG;
    const int pair = blockIdx.x;
    const SpbGeom& g = geoms[pairs[pair].geom];
    const float* pp = work + pair * work_stride;
    const float* ps = pp + (size_t)ctas * NACC;
    float* op = out_pair + (size_t)pair * SPB_GN_PAIR_NOUT;
    __shared__ float s_part[SPB_FIN_SMEM];
    const float tot = sum_cta_partials<NACC>(pp, ctas, s_part);
    const int so = seg_off[pair];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if constexpr (NP == 8) {
        if (threadIdx.x < NACC) op[threadIdx.x] = tot;
        if (threadIdx.x == SPB_GN_PAIR_NOUT - 1) op[threadIdx.x] = 0.f;
        // two segments per warp per round; lane i < NSEG ends up with value i of the segment (seg_sum2)
        for (int b = warp; b < g.n_seg; b += 2 * nwarps) {
            const int b2 = b + nwarps;
            const bool two = b2 < g.n_seg;
            float v, v2;
            seg_sum2<NSEG>(ps, g.seg_tile[b], g.seg_tile[b + 1], two ? g.seg_tile[b2] : 0, two ? g.seg_tile[b2 + 1] : 0,
                           lane, v, v2);
            if (lane < NSEG) {       // fixed 10-float record: B[0..7], D, g_d
                out_seg[(size_t)(so + b) * SPB_GN_SEG_NOUT + lane] = v;
                if (two) out_seg[(size_t)(so + b2) * SPB_GN_SEG_NOUT + lane] = v2;
            }
        }
    } else {
        // 6 pose columns (spb_gn_packed.cuh): the per-run records hold rows 0..2 of the pose block and g_0..g_3; the
        // depth column of a segment is -(t_x J_0 + t_y J_1 + t_z J_2), so B_b, D_b, g_d,b are combinations of the
        // segment's sums (gn6_segment_record) and rows 0..2 of the pose block are the sums over all segments.
        __shared__ float s_rows[SPB_FIN_THREADS_MAX / 32][SPB_GN6_NRUN];
        const float* pose = pairs[pair].pose;
        const double t[3] = {(double)pose[3], (double)pose[7], (double)pose[11]};
        float rows = 0.f;                                       // lane l < 19: running sum of value l over this warp's segments
        for (int b = warp; b < g.n_seg; b += 2 * nwarps) {
            const int b2 = b + nwarps;
            const bool two = b2 < g.n_seg;
            float v, v2;
            seg_sum2<NSEG>(ps, g.seg_tile[b], g.seg_tile[b + 1], two ? g.seg_tile[b2] : 0, two ? g.seg_tile[b2 + 1] : 0,
                           lane, v, v2);
            if (lane >= NSEG) { v = 0.f; v2 = 0.f; }
            if (!two) v2 = 0.f;
            rows += v;
            rows += v2;
            // lane i < 6 derives B[i], lane 6 D, lane 7 g_d; lanes 8, 9 write the zero affine slots of the record
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const float val = which ? v2 : v;
                const int bb = which ? b2 : b;
                const int col = lane < 6 ? lane : 0;
                double r0 = (double)__shfl_sync(0xffffffffu, val, lane < 6 ? gn6_run_index(0, col) : (lane == 7 ? 15 : 0));
                double r1 = (double)__shfl_sync(0xffffffffu, val, lane < 6 ? gn6_run_index(1, col) : (lane == 7 ? 16 : 1));
                double r2 = (double)__shfl_sync(0xffffffffu, val, lane < 6 ? gn6_run_index(2, col) : (lane == 7 ? 17 : 2));
                double rec = -(t[0] * r0 + t[1] * r1 + t[2] * r2);       // B[lane] (lanes 0..5), g_d (lane 7)
                // D = -(t . B[0..2]) = t^T A[0:3,0:3] t
                const double b0 = __shfl_sync(0xffffffffu, rec, 0), b1 = __shfl_sync(0xffffffffu, rec, 1),
                             b2v = __shfl_sync(0xffffffffu, rec, 2);
                if (lane == 6) rec = -(t[0] * b0 + t[1] * b1 + t[2] * b2v);
                if ((which == 0 || two) && lane < 10) {
                    const int slot = lane < 6 ? lane : (lane < 8 ? lane + 2 : lane - 2);   // B[6], B[7] (affine) stay zero
                    out_seg[(size_t)(so + bb) * SPB_GN_SEG_NOUT + slot] = lane < 8 ? (float)rec : 0.f;
                }
            }
        }
        if (lane < SPB_GN6_NRUN) s_rows[warp][lane] = rows;
        __syncthreads();
        if (threadIdx.x < SPB_GN_PAIR_NOUT) op[threadIdx.x] = 0.f;   // affine rows / columns and padding stay zero
        __syncthreads();
        if (threadIdx.x < SPB_GN6_NRUN) {
            float v = 0.f;
            for (int w = 0; w < nwarps; ++w) v += s_rows[w][threadIdx.x];
            const int i = threadIdx.x;
            if (i < 6) op[tri8(0, i)] = v;
            else if (i < 11) op[tri8(1, i - 5)] = v;
            else if (i < 15) op[tri8(2, i - 9)] = v;
            else op[SPB_GN_NA + (i - 15)] = v;                   // g_0..g_3
        }
        // rotation block, g_4, g_5, cost from the per-CTA partials (threads < NACC hold `tot`)
        if (threadIdx.x < 9) {
            const int i = threadIdx.x;
            int idx;
            if (i == 0) idx = tri8(3, 3);
            else if (i < 3) idx = tri8(3, 3 + i);
            else if (i < 5) idx = tri8(4, 4 + (i - 3));
            else if (i == 5) idx = tri8(5, 5);
            else if (i < 8) idx = SPB_GN_NA + 4 + (i - 6);
            else idx = SPB_GN_NA + 8;                            // cost
            op[idx] = tot;
        }
    }
}

template <int NP>
__global__ void k_finalize_gn(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs,
                              const int32_t* __restrict__ seg_off, int ctas, const float* __restrict__ work,
                              int64_t work_stride, float* __restrict__ out_pair, float* __restrict__ out_seg) {
    finalize_gn_body<NP>(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_seg);
}

// finalize + damped solve + retraction in ONE launch (one CTA per problem): the second kernel of a GN iteration
template <int NP>
__global__ void __launch_bounds__(SPB_FIN_THREADS)
k_gn_finalize_solve(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs,
                    const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_cnt, int ctas,
                    const float* __restrict__ work, int64_t work_stride, float* gn_pair, float* gn_seg,
                    int with_affine, int hold_depth, float* __restrict__ poses, float* __restrict__ k,
                    float* __restrict__ aff_trg, float* __restrict__ lm_state, float* __restrict__ saved_pair,
                    float* __restrict__ saved_seg) {
    finalize_gn_body<NP>(geoms, pairs, seg_off, ctas, work, work_stride, gn_pair, gn_seg);
    __threadfence_block();
    __syncthreads();
    lm_update_body(gn_pair, gn_seg, seg_off, seg_cnt, with_affine, hold_depth, poses, k, aff_trg, lm_state, saved_pair,
                   saved_seg);
}

// ------------------------------------------------------------------------------------------------
// pre-lifted points variant (tracking): X given, no log-depth gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPB_THREADS, 2)
k_align_points(const float* __restrict__ src_pts, const float* __restrict__ src_px,
               const uint8_t* __restrict__ src_ok, int P, int H, int W, const __grid_constant__ SpbPair pr,
               float* __restrict__ work) {
    constexpr int NACC = SPB_PAIR_NOUT;
    __shared__ float s_ctx[C_N];
    __shared__ float s_red[SPB_WARPS * NACC];
    __shared__ float s_K[9];
    if (threadIdx.x < 9) s_K[threadIdx.x] = (threadIdx.x % 4 == 0) ? 1.f : 0.f;   // unused source intrinsics
    __syncthreads();
    if (threadIdx.x < 32) fill_ctx(s_ctx, pr, s_K, H, W);
    __syncthreads();
    const float* c = s_ctx;
    const float4* trg = reinterpret_cast<const float4*>(pr.trg_rgba);
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    float none[1] = {0.f};
    PointOut po{nullptr, 0, P};
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        const float Xx = __ldg(src_pts + 3 * (size_t)p), Xy = __ldg(src_pts + 3 * (size_t)p + 1),
                    Xz = __ldg(src_pts + 3 * (size_t)p + 2);
        const bool sok = __ldg(src_ok + p) != 0;
        const float i0 = __ldg(src_px + p), i1 = __ldg(src_px + (size_t)P + p), i2 = __ldg(src_px + 2 * (size_t)P + p);
        eval_point<MODE_GRAD, 6, false, NACC, 0 + 1>(c, trg, pr.Wl, pr.Hl, Xx, Xy, Xz, sok, i0, i1, i2, 0.f, acc,
                                                     none, po, p);
    }
    block_reduce_store<NACC>(acc, s_red, work + (size_t)blockIdx.x * NACC);
}

__global__ void k_finalize_points(int ctas, int P, const float* __restrict__ work, float* __restrict__ out_pair) {
    __shared__ float s_part[SPB_FIN_SMEM];
    const float tot = sum_cta_partials<SPB_PAIR_NOUT>(work, ctas, s_part);
    if (threadIdx.x < SPB_PAIR_NOUT) {
        const float v = tot;
        const float norm = 1.0f / (3.0f * (float)P);
        out_pair[threadIdx.x] = (threadIdx.x == 15) ? v : v * norm;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers (C ABI)
// ------------------------------------------------------------------------------------------------
static inline int ctas_for(int n_tiles, int n_pairs, int occ) {
    // every warp streams a strided set of tiles through its own ring.  Size the grid to (nearly) a whole number of
    // waves of the kernel's occupancy (`occ` CTAs/SM): the smallest CTA count per pair that gives >= 3 waves over all
    // pairs with a last wave >= 97 % full, else the fullest within 8 waves (a 5.3-wave grid wastes a third of its last wave; 1024 pairs
    // with one CTA each would run 2.3 waves at 77 %).
    const int max_ctas = (n_tiles + SPB_WARPS - 1) / SPB_WARPS;
    const int slots = spb_sm_count() * occ;
    if (n_pairs < 1) n_pairs = 1;
    if ((long long)n_pairs * max_ctas <= slots) return max_ctas;
    int best = 1;
    double best_score = -1.0;
    for (int c = 1; c <= max_ctas; ++c) {
        const double w = (double)n_pairs * c / slots;
        if (w > 8.0 && best_score >= 0.0) break;
        const double eff = w / (double)(long long)(w + 0.999999);
        if (w >= 3.0 && eff >= 0.97) return c;                       // the smallest grid that is good enough
        const double score = (w >= 3.0 ? 1.0 : w / 3.0) * eff;       // fewer than 3 waves: the tail weighs more
        if (score > best_score + 1e-9) { best_score = score; best = c; }
    }
    return best;
}
static inline int ctas_grad(int n_tiles, int n_pairs) { return ctas_for(n_tiles, n_pairs, Occ<MODE_GRAD, 6>::CTAS); }
static inline int ctas_gn(int n_tiles, int n_pairs, int with_affine) {
    return ctas_for(n_tiles, n_pairs, with_affine == 1 ? Occ<MODE_GN, 8>::CTAS : Occ<MODE_GN, 6>::CTAS);
}
static inline int ctas_max(int n_tiles, int n_pairs) {
    int a = ctas_grad(n_tiles, n_pairs), b = ctas_gn(n_tiles, n_pairs, 0), c = ctas_gn(n_tiles, n_pairs, 1);
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

template <typename K>
static inline cudaError_t allow_dyn_smem(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPB_FAST_DYN_SMEM);
}

static inline int ctas_for_points(int P) {
    int want = (P + SPB_THREADS * 4 - 1) / (SPB_THREADS * 4);
    if (want > spb_sm_count() * 4) want = spb_sm_count() * 4;
    if (want < 1) want = 1;
    return want;
}

extern "C" int64_t spb_workspace_floats(const SpbGeom* geom, int B, int gn) {
    const int ctas = ctas_max(geom->n_tiles, B);
    const int nacc = gn ? SPB_MAX_NACC : SPB_PAIR_NOUT;
    const int nseg = gn ? SPB_MAX_NSEG : 1;
    return (int64_t)B * ((int64_t)ctas * nacc + (int64_t)geom->n_tiles * nseg);
}

// floats per pair of the batched entry points' workspace (any mode): per-CTA accumulators + per-tile run records
extern "C" int64_t spb_gn_work_stride(int max_tiles, int n_pairs) {
    return (int64_t)ctas_max(max_tiles, n_pairs) * SPB_MAX_NACC + (int64_t)max_tiles * SPB_MAX_NSEG;
}

extern "C" int64_t spb_workspace_floats_points(int P) { return (int64_t)ctas_for_points(P) * SPB_PAIR_NOUT; }

extern "C" int spb_cost_grad(const SpbGeom* geom, const SpbPair* pairs, int B, float* work, float* out_pair,
                             float* out_gk, float* out_pose, float* out_flag, const SpbStats* stats, void* stream) {
    if (!geom || !pairs || B < 1 || !work || !out_pair || !out_gk) return SPB_EINVAL;
    if (B > 16) return SPB_ELIMIT;
    if (geom->n_pts <= 0 || geom->n_tiles <= 0) return SPB_EINVAL;
    for (int j = 0; j < B; ++j)
        if (pairs[j].Wl < 2 || pairs[j].Hl < 2 || !pairs[j].trg_rgba || !pairs[j].src_rgb || !pairs[j].pose ||
            !pairs[j].k || !pairs[j].K_trg)
            return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    PairPack pack;
    for (int j = 0; j < B; ++j) pack.p[j] = pairs[j];
    for (int j = B; j < 16; ++j) pack.p[j] = pairs[0];
    const int ctas = ctas_grad(geom->n_tiles, B);
    dim3 grid(ctas, B);
    if (stats) {
        k_align_stats<<<grid, SPB_THREADS, 0, st>>>(*geom, pack, work, *stats);
    } else {
        bool aff = false;
        for (int j = 0; j < B; ++j) aff = aff || (pairs[j].aff_src != nullptr && pairs[j].aff_trg != nullptr);
        if (aff) {
            cudaError_t e = allow_dyn_smem(k_align_inline<MODE_GRAD, 6, true>);
            if (e != cudaSuccess) return (int)e;
            k_align_inline<MODE_GRAD, 6, true><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(*geom, pack, 0.f, work);
        } else {
            cudaError_t e = allow_dyn_smem(k_align_inline<MODE_GRAD, 6, false>);
            if (e != cudaSuccess) return (int)e;
            k_align_inline<MODE_GRAD, 6, false><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(*geom, pack, 0.f, work);
        }
    }
    SPB_CHECK_LAUNCH();
    k_finalize_grad<<<B, 256, 0, st>>>(*geom, pack, ctas, stats ? 0 : 1, work, out_pair, out_gk, out_pose, out_flag);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_cost_grad_points(const float* src_pts, const float* src_px, const uint8_t* src_ok, int P, int H,
                                    int W, const SpbPair* pair, float* work, float* out_pair, void* stream) {
    if (!src_pts || !src_px || !src_ok || P < 1 || !pair || !work || !out_pair) return SPB_EINVAL;
    if (pair->Wl < 2 || pair->Hl < 2 || !pair->trg_rgba || !pair->pose || !pair->K_trg) return SPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = ctas_for_points(P);
    k_align_points<<<ctas, SPB_THREADS, 0, st>>>(src_pts, src_px, src_ok, P, H, W, *pair, work);
    SPB_CHECK_LAUNCH();
    k_finalize_points<<<1, 256, 0, st>>>(ctas, P, work, out_pair);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_gn_ctas(int max_tiles, int n_pairs) { return ctas_max(max_tiles, n_pairs); }

static int launch_gn_align(const SpbGeom* geoms, const SpbPair* pairs, int n_pairs, int ctas, float irls_eps,
                           int with_affine, float* work, int64_t work_stride, cudaStream_t st) {
    // with_affine: 0 = no brightness terms, 1 = optimise the target affine (8 pose columns),
    //              2 = brightness terms present but fixed (6 pose columns)
    dim3 grid(ctas, n_pairs);
    cudaError_t e;
    if (with_affine == 1) {
        if ((e = allow_dyn_smem(k_align_global<MODE_GN, 8, true>)) != cudaSuccess) return (int)e;
        k_align_global<MODE_GN, 8, true><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(geoms, pairs, irls_eps, work, work_stride);
    } else if (with_affine == 2) {
        if ((e = allow_dyn_smem(k_align_global<MODE_GN, 6, true>)) != cudaSuccess) return (int)e;
        k_align_global<MODE_GN, 6, true><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(geoms, pairs, irls_eps, work, work_stride);
    } else {
        if ((e = allow_dyn_smem(k_align_global<MODE_GN, 6, false>)) != cudaSuccess) return (int)e;
        k_align_global<MODE_GN, 6, false><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(geoms, pairs, irls_eps, work, work_stride);
    }
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

static bool gn_stride_ok(int ctas, int max_tiles, int with_affine, int64_t work_stride) {
    const bool np8 = with_affine == 1;
    const int nacc = np8 ? Sizes<MODE_GN, 8>::NACC : Sizes<MODE_GN, 6>::NACC;
    const int nseg = np8 ? Sizes<MODE_GN, 8>::NSEG : Sizes<MODE_GN, 6>::NSEG;
    return work_stride >= (int64_t)ctas * nacc + (int64_t)max_tiles * nseg;
}

extern "C" int spb_gn_accumulate(const SpbGeom* geoms, const SpbPair* pairs, const int32_t* seg_off, int n_pairs,
                                 int max_tiles, float irls_eps, int with_affine, float* work, int64_t work_stride,
                                 float* out_pair, float* out_seg, void* ev_before, void* ev_after, void* stream) {
    if (!geoms || !pairs || !seg_off || n_pairs < 1 || max_tiles < 1 || !work || !out_pair || !out_seg)
        return SPB_EINVAL;
    if (n_pairs > 65535) return SPB_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = ctas_gn(max_tiles, n_pairs, with_affine);
    if (!gn_stride_ok(ctas, max_tiles, with_affine, work_stride)) return SPB_EINVAL;
    if (ev_before) cudaEventRecord((cudaEvent_t)ev_before, st);
    const int rc = launch_gn_align(geoms, pairs, n_pairs, ctas, irls_eps, with_affine, work, work_stride, st);
    if (rc != SPB_OK) return rc;
    if (ev_after) cudaEventRecord((cudaEvent_t)ev_after, st);
    if (with_affine == 1)
        k_finalize_gn<8><<<n_pairs, 256, 0, st>>>(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_seg);
    else
        k_finalize_gn<6><<<n_pairs, 256, 0, st>>>(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_seg);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// One complete GN/LM iteration in two launches: fused residual+Jacobian+normal-equation kernel, then
// finalize + damped solve + retraction (k_gn_finalize_solve).
extern "C" int spb_gn_iterate(const SpbGeom* geoms, const SpbPair* pairs, const int32_t* seg_off, const int32_t* seg_cnt,
                              int n_pairs, int max_tiles, float irls_eps, int with_affine, int hold_depth, float* work,
                              int64_t work_stride, float* gn_pair, float* gn_seg, float* poses, float* k,
                              float* aff_trg, float* lm_state, float* saved_pair, float* saved_seg, void* ev_before,
                              void* ev_after, void* stream) {
    if (!geoms || !pairs || !seg_off || !seg_cnt || n_pairs < 1 || max_tiles < 1 || !work || !gn_pair || !gn_seg ||
        !poses || !k || !lm_state || !saved_pair || !saved_seg)
        return SPB_EINVAL;
    if (n_pairs > 65535) return SPB_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = ctas_gn(max_tiles, n_pairs, with_affine);
    if (!gn_stride_ok(ctas, max_tiles, with_affine, work_stride)) return SPB_EINVAL;
    if (ev_before) cudaEventRecord((cudaEvent_t)ev_before, st);
    const int rc = launch_gn_align(geoms, pairs, n_pairs, ctas, irls_eps, with_affine, work, work_stride, st);
    if (rc != SPB_OK) return rc;
    if (ev_after) cudaEventRecord((cudaEvent_t)ev_after, st);
    const int opt_aff = with_affine == 1 ? 1 : 0;
    if (with_affine == 1)
        k_gn_finalize_solve<8><<<n_pairs, SPB_FIN_THREADS, 0, st>>>(geoms, pairs, seg_off, seg_cnt, ctas, work, work_stride, gn_pair,
                                                        gn_seg, opt_aff, hold_depth ? 1 : 0, poses, k, aff_trg, lm_state,
                                                        saved_pair, saved_seg);
    else
        k_gn_finalize_solve<6><<<n_pairs, SPB_FIN_THREADS, 0, st>>>(geoms, pairs, seg_off, seg_cnt, ctas, work, work_stride, gn_pair,
                                                        gn_seg, opt_aff, hold_depth ? 1 : 0, poses, k, aff_trg, lm_state,
                                                        saved_pair, saved_seg);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// gradient mode over device-resident descriptors (batched Adam-parity iterations / benchmark)
__device__ __forceinline__ void finalize_grad_global_body(const SpbGeom* __restrict__ geoms,
                                                          const SpbPair* __restrict__ pairs,
                                                          const int32_t* __restrict__ seg_off, int ctas,
                                                          const float* __restrict__ work, int64_t work_stride,
                                                          float* out_pair, float* out_gk) {
    const int pair = blockIdx.x;
    const SpbGeom& g = geoms[pairs[pair].geom];
    const float* pp = work + pair * work_stride;
    const float* ps = pp + (size_t)ctas * SPB_PAIR_NOUT;
    const float norm = 1.0f / (3.0f * (float)g.n_pts);
    __shared__ float s_part[SPB_FIN_SMEM];
    const float tot = sum_cta_partials<SPB_PAIR_NOUT>(pp, ctas, s_part);
    if (threadIdx.x < SPB_PAIR_NOUT) {
        float v = tot;
        v *= grad_col_scale(threadIdx.x, g.K, true);
        out_pair[pair * SPB_PAIR_NOUT + threadIdx.x] = (threadIdx.x == 15) ? v : v * norm;
    }
    const int so = seg_off[pair];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int b = warp; b < g.n_seg; b += 2 * nwarps) {
        const int b2 = b + nwarps;
        const bool two = b2 < g.n_seg;
        float v, v2;
        seg_sum2<1>(ps, g.seg_tile[b], g.seg_tile[b + 1], two ? g.seg_tile[b2] : 0, two ? g.seg_tile[b2 + 1] : 0, lane,
                    v, v2);
        if (lane == 0) {
            out_gk[so + b] = v * norm;
            if (two) out_gk[so + b2] = v2 * norm;
        }
    }
}

__global__ void k_finalize_grad_global(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs,
                                       const int32_t* __restrict__ seg_off, int ctas, const float* __restrict__ work,
                                       int64_t work_stride, float* __restrict__ out_pair, float* __restrict__ out_gk) {
    finalize_grad_global_body(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_gk);
}

// finalize + Adam update + retraction in ONE launch (one CTA per problem): the second kernel of a first-order iteration
__global__ void __launch_bounds__(256)
k_grad_finalize_adam(const SpbGeom* __restrict__ geoms, const SpbPair* __restrict__ pairs,
                     const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_cnt, int ctas,
                     const float* __restrict__ work, int64_t work_stride, float* out_pair, float* out_gk, int with_affine,
                     const SpbAdamHyper h, float* __restrict__ poses, float* __restrict__ k, float* __restrict__ aff_trg,
                     float* __restrict__ adam_pair, float* __restrict__ adam_seg) {
    finalize_grad_global_body(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_gk);
    __threadfence_block();
    __syncthreads();
    adam_update_body(out_pair, out_gk, seg_off, seg_cnt, with_affine, h, poses, k, aff_trg, adam_pair, adam_seg);
}

// the fused gradient-mode launch of the batched entry points
static int launch_grad_align(const SpbGeom* geoms, const SpbPair* pairs, int n_pairs, int ctas, int with_affine,
                             float* work, int64_t work_stride, cudaStream_t st) {
    dim3 grid(ctas, n_pairs);
    if (with_affine) {
        cudaError_t e = allow_dyn_smem(k_align_global<MODE_GRAD, 6, true>);
        if (e != cudaSuccess) return (int)e;
        k_align_global<MODE_GRAD, 6, true><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(geoms, pairs, 0.f, work, work_stride);
    } else {
        cudaError_t e = allow_dyn_smem(k_align_global<MODE_GRAD, 6, false>);
        if (e != cudaSuccess) return (int)e;
        k_align_global<MODE_GRAD, 6, false><<<grid, SPB_THREADS, SPB_FAST_DYN_SMEM, st>>>(geoms, pairs, 0.f, work, work_stride);
    }
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_grad_accumulate(const SpbGeom* geoms, const SpbPair* pairs, const int32_t* seg_off, int n_pairs,
                                   int max_tiles, int with_affine, float* work, int64_t work_stride, float* out_pair,
                                   float* out_gk, void* ev_before, void* ev_after, void* stream) {
    if (!geoms || !pairs || !seg_off || n_pairs < 1 || max_tiles < 1 || !work || !out_pair || !out_gk)
        return SPB_EINVAL;
    if (n_pairs > 65535) return SPB_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = ctas_grad(max_tiles, n_pairs);
    if (work_stride < (int64_t)ctas * SPB_PAIR_NOUT + max_tiles) return SPB_EINVAL;
    if (ev_before) cudaEventRecord((cudaEvent_t)ev_before, st);
    {
        const int rc = launch_grad_align(geoms, pairs, n_pairs, ctas, with_affine, work, work_stride, st);
        if (rc != SPB_OK) return rc;
    }
    if (ev_after) cudaEventRecord((cudaEvent_t)ev_after, st);
    k_finalize_grad_global<<<n_pairs, 256, 0, st>>>(geoms, pairs, seg_off, ctas, work, work_stride, out_pair, out_gk);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

static inline bool adam_hyper_ok(double lr_pose, double lr_k, double lr_aff, double beta1, double beta2, double eps) {
    return lr_pose >= 0.0 && lr_k >= 0.0 && lr_aff >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 &&
           eps >= 0.0;
}

// One complete first-order iteration in two launches: fused residual+gradient kernel, then finalize + Adam update +
// retraction (k_grad_finalize_adam).
extern "C" int spb_adam_iterate(const SpbGeom* geoms, const SpbPair* pairs, const int32_t* seg_off, const int32_t* seg_cnt,
                                int n_pairs, int max_tiles, int with_affine, float* work, int64_t work_stride,
                                float* out_pair, float* out_gk, float* poses, float* k, float* aff_trg, float* adam_pair,
                                float* adam_seg, double lr_pose, double lr_k, double lr_aff, double beta1, double beta2,
                                double eps, void* ev_before, void* ev_after, void* stream) {
    if (!geoms || !pairs || !seg_off || !seg_cnt || n_pairs < 1 || max_tiles < 1 || !work || !out_pair || !out_gk ||
        !poses || !k || !adam_pair || !adam_seg || !adam_hyper_ok(lr_pose, lr_k, lr_aff, beta1, beta2, eps))
        return SPB_EINVAL;
    if (with_affine == 1 && !aff_trg) return SPB_EINVAL;
    if (n_pairs > 65535) return SPB_ELIMIT;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = ctas_grad(max_tiles, n_pairs);
    if (work_stride < (int64_t)ctas * SPB_PAIR_NOUT + max_tiles) return SPB_EINVAL;
    if (ev_before) cudaEventRecord((cudaEvent_t)ev_before, st);
    {
        const int rc = launch_grad_align(geoms, pairs, n_pairs, ctas, with_affine, work, work_stride, st);
        if (rc != SPB_OK) return rc;
    }
    if (ev_after) cudaEventRecord((cudaEvent_t)ev_after, st);
    const SpbAdamHyper h{lr_pose, lr_k, lr_aff, beta1, beta2, eps};
    k_grad_finalize_adam<<<n_pairs, 256, 0, st>>>(geoms, pairs, seg_off, seg_cnt, ctas, work, work_stride, out_pair, out_gk,
                                                  with_affine == 1 ? 1 : 0, h, poses, k, aff_trg, adam_pair, adam_seg);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
