// Packed-FP32 (FFMA2 / FMUL2, sm_100) formulation of the 6-column Gauss-Newton accumulation.
//
// Blackwell issues two independent FP32 FMAs per FFMA2 instruction and accepts a scalar register as a
// broadcast operand (SASS `R.F32`), so every rank-1 update `acc[r][c..c+1] += m_r * p[c..c+1]` of the
// arrowhead blocks is ONE instruction instead of two.  The kernel is issue-bound, not FMA-pipe-bound,
// so halving the instruction count of the accumulation is a direct win.  Same arithmetic as point_gn
// (spb_fast.cuh) up to the association order inside each two-term sum.
#pragma once
#include "spb_fast.cuh"

// Depth-column identity (exact for a live reciprocal): a depth change moves a point along its epipolar line, so the
// Jacobian column of the segment's log-depth seed is a fixed combination of the first three twist columns,
//     j_d = -(t_x J_0 + t_y J_1 + t_z J_2).
// The kernel therefore never forms j_d: it accumulates rows 0..2 of J^T W J and g_0..g_3 PER RUN of same-segment tiles
// (19 values, summed over the segment's points by the finalize kernel, which derives B_b, D_b, g_d,b from them --
// gn6_segment_record below -- and the pose block's rows 0..2 as the sum over the segments), and only the 3x3 rotation
// block, g_4, g_5 and the cost per warp.  28 live accumulators instead of 36 and 24 instead of 33 update instructions.
#define SPB_GN6_NRUN 19        // per-run values: A[0][0..5], A[1][1..5], A[2][2..5], g_0..g_3
#define SPB_GN6_NACC 12        // per-warp values: A[3][3..5], A[4][4..5], A[5][5], g_4, g_5, cost, 0, 0, 0

struct GnAcc6 {            // rotation block of the pose matrix (upper triangle), g_4, g_5, cost
    float a33;
    float2 a34;            // (3,4)(3,5)
    float2 a44;            // (4,4)(4,5)
    float a55;
    float2 g45;
    float cost;            // sum |r| ; the weighted cost and the valid count are not accumulated on this path
    __device__ __forceinline__ void zero() {
        a34 = a44 = g45 = make_float2(0.f, 0.f);
        a33 = a55 = cost = 0.f;
    }
    __device__ __forceinline__ void store(float (&o)[SPB_GN6_NACC]) const {
        o[0] = a33; o[1] = a34.x; o[2] = a34.y; o[3] = a44.x; o[4] = a44.y; o[5] = a55;
        o[6] = g45.x; o[7] = g45.y; o[8] = cost; o[9] = 0.f; o[10] = 0.f; o[11] = 0.f;
    }
};

struct GnSeg6 {            // rows 0..2 of the pose block + g_0..g_3 over the points of one same-segment run
    float2 a00, a02, a04;  // (0,0)(0,1) (0,2)(0,3) (0,4)(0,5)
    float a11;
    float2 a12, a14;       // (1,2)(1,3) (1,4)(1,5)
    float2 a22, a24;       // (2,2)(2,3) (2,4)(2,5)
    float2 g01, g23;
    __device__ __forceinline__ void zero() {
        a00 = a02 = a04 = a12 = a14 = a22 = a24 = g01 = g23 = make_float2(0.f, 0.f);
        a11 = 0.f;
    }
    __device__ __forceinline__ void store(float (&o)[SPB_GN6_NRUN]) const {
        o[0] = a00.x; o[1] = a00.y; o[2] = a02.x; o[3] = a02.y; o[4] = a04.x; o[5] = a04.y;
        o[6] = a11; o[7] = a12.x; o[8] = a12.y; o[9] = a14.x; o[10] = a14.y;
        o[11] = a22.x; o[12] = a22.y; o[13] = a24.x; o[14] = a24.y;
        o[15] = g01.x; o[16] = g01.y; o[17] = g23.x; o[18] = g23.y;
    }
};

// entry (r, i), r < 3, of the symmetric pose block inside a per-run record S[19]
__host__ __device__ inline int gn6_run_index(int r, int i) {
    if (i < r) { const int t = r; r = i; i = t; }
    return r == 0 ? i : (r == 1 ? 5 + i : 9 + i);
}
// B_b[0..5], D_b, g_d,b of one segment from the sums S[19] over its points and the translation t of the pose
// (float64: D is a quadratic form with cancellation between its terms)
__host__ __device__ inline void gn6_segment_record(const double* S, const double* t, double* out /* [8] */) {
    for (int i = 0; i < 6; ++i)
        out[i] = -(t[0] * S[gn6_run_index(0, i)] + t[1] * S[gn6_run_index(1, i)] + t[2] * S[gn6_run_index(2, i)]);
    out[6] = -(t[0] * out[0] + t[1] * out[1] + t[2] * out[2]);
    out[7] = -(t[0] * S[15] + t[1] * S[16] + t[2] * S[17]);
}

// scalar * pair (+ pair): the scalar is a broadcast operand, no extra instruction
__device__ __forceinline__ float2 fma2(float s, float2 b, float2 c) { return __ffma2_rn(make_float2(s, s), b, c); }
__device__ __forceinline__ float2 mul2(float s, float2 b) { return __fmul2_rn(make_float2(s, s), b); }

// value and slope pair (dI/dix, dI/diy) of one channel
__device__ __forceinline__ void blend2(float nw, float ne, float sw, float se, float fx, float fy, float& val,
                                       float2& d) {
    const float d0 = ne - nw;
    const float d1 = se - sw;
    const float top = fmaf(fx, d0, nw);
    const float bot = fmaf(fx, d1, sw);
    d.y = bot - top;
    val = fmaf(fy, d.y, top);
    d.x = fmaf(fy, d1 - d0, d0);
}

// Precondition: q.live (callers treat points behind the guarded-reciprocal threshold |Yz| <= 1e-6 as invalid on
// this path -- a sub-micrometre depth regime where the reference itself samples with a clamped reciprocal).
template <bool AFF>
__device__ __forceinline__ void point_gn6_packed(const float* __restrict__ c, const Taps4& tp, const Proj& q,
                                                 float Is0, float Is1, float Is2, float irls_eps, GnAcc6& A,
                                                 GnSeg6& S) {
    float I[3];
    float2 d[3];
    blend2(tp.nw.x, tp.ne.x, tp.sw.x, tp.se.x, q.fx, q.fy, I[0], d[0]);
    blend2(tp.nw.y, tp.ne.y, tp.sw.y, tp.se.y, q.fx, q.fy, I[1], d[1]);
    blend2(tp.nw.z, tp.ne.z, tp.sw.z, tp.se.z, q.fx, q.fy, I[2], d[2]);
    const float Is[3] = {Is0, Is1, Is2};
    const float ea = c[F_EA], bb = c[F_BB];
    // GA = (Guu, Guv), GB = (Guv, Gvv), H = (hu, hv): image-gradient moments over the channels
    float2 GA = make_float2(0.f, 0.f), GB = GA, H = GA;
    float cost = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float r = AFF ? (Is[ch] - fmaf(ea, I[ch], bb)) : (Is[ch] - I[ch]);
        const float ar = fabsf(r);
        const float wgt = __fdividef(1.0f, fmaxf(ar, irls_eps));     // IRLS weight of the L1 objective
        const float wr = wgt * r;
        cost += ar;
        const float2 wd = mul2(wgt, d[ch]);                          // (w dIx, w dIy)
        GA = fma2(wd.x, d[ch], GA);
        GB = fma2(wd.y, d[ch], GB);
        H = fma2(wr, d[ch], H);
    }
    GA = __fmul2_rn(GA, *reinterpret_cast<const float2*>(c + F_CUU));     // (cu^2, cu cv)
    GB = __fmul2_rn(GB, *reinterpret_cast<const float2*>(c + F_CUV2));    // (cu cv, cv^2)
    H = __fmul2_rn(H, *reinterpret_cast<const float2*>(c + F_CU));        // (cu, cv)
    const float Guu = GA.x, Guv = GA.y, Gvv = GB.y, hu = H.x, hv = H.y;

    // mu = d x_/d xi, mv = d y_/d xi with mu0 = rho, mu1 = 0, mv0 = 0, mv1 = rho
    const float rho = q.rho, xb = q.xb, yb = q.yb;
    const float xy = xb * yb;
    const float2 mu23 = make_float2(-rho * xb, -xy);
    const float2 mu45 = make_float2(fmaf(xb, xb, 1.0f), -yb);
    const float2 mv23 = make_float2(-rho * yb, -fmaf(yb, yb, 1.0f));
    const float2 mv45 = make_float2(xy, xb);
    // pu = Guu mu + Guv mv ; pv = Guv mu + Gvv mv
    const float2 pu01 = mul2(rho, GA);                               // (Guu rho, Guv rho)
    const float pv1 = Gvv * rho;
    const float2 pu23 = fma2(Guu, mu23, mul2(Guv, mv23)), pv23 = fma2(Guv, mu23, mul2(Gvv, mv23));
    const float2 pu45 = fma2(Guu, mu45, mul2(Guv, mv45)), pv45 = fma2(Guv, mu45, mul2(Gvv, mv45));

    // rows 0..2 of the pose block and g_0..g_3: per run (the segment's depth column follows from them)
    S.a00 = fma2(rho, pu01, S.a00);
    S.a02 = fma2(rho, pu23, S.a02);
    S.a04 = fma2(rho, pu45, S.a04);
    S.a11 = fmaf(rho, pv1, S.a11);
    S.a12 = fma2(rho, pv23, S.a12);
    S.a14 = fma2(rho, pv45, S.a14);
    S.a22 = fma2(mu23.x, pu23, fma2(mv23.x, pv23, S.a22));
    S.a24 = fma2(mu23.x, pu45, fma2(mv23.x, pv45, S.a24));
    S.g01 = fma2(rho, H, S.g01);
    S.g23 = fma2(hu, mu23, fma2(hv, mv23, S.g23));
    // rotation block, g_4, g_5, cost: per warp
    A.a33 = fmaf(mu23.y, pu23.y, fmaf(mv23.y, pv23.y, A.a33));
    A.a34 = fma2(mu23.y, pu45, fma2(mv23.y, pv45, A.a34));
    A.a44 = fma2(mu45.x, pu45, fma2(mv45.x, pv45, A.a44));
    A.a55 = fmaf(mu45.y, pu45.y, fmaf(mv45.y, pv45.y, A.a55));
    A.g45 = fma2(hu, mu45, fma2(hv, mv45, A.g45));
    A.cost += cost;
}

// ---- gradient mode, packed ------------------------------------------------------------------------
// canonical order (SPB_PAIR_NOUT): cost, gt[3], gM rows (3x3), ga, gb, nvalid
struct GradAcc {
    float cost;
    float2 gt01;
    float gt2;
    float2 m00;  float m02;     // row 0 of d cost / d M
    float2 m10;  float m12;
    float2 m20;  float m22;
    float ga, gb, nv;
    __device__ __forceinline__ void zero() {
        const float2 z = make_float2(0.f, 0.f);
        gt01 = m00 = m10 = m20 = z;
        cost = gt2 = m02 = m12 = m22 = ga = gb = nv = 0.f;
    }
    __device__ __forceinline__ void store(float (&o)[16]) const {
        o[0] = cost; o[1] = gt01.x; o[2] = gt01.y; o[3] = gt2;
        o[4] = m00.x; o[5] = m00.y; o[6] = m02; o[7] = m10.x; o[8] = m10.y; o[9] = m12;
        o[10] = m20.x; o[11] = m20.y; o[12] = m22; o[13] = ga; o[14] = gb; o[15] = nv;
    }
};

template <bool AFF>
__device__ __forceinline__ void point_grad_packed(const float* __restrict__ c, const Taps4& tp, const Proj& q,
                                                  float Is0, float Is1, float Is2, GradAcc& A, float& gk) {
    float I[3];
    float2 d[3];
    blend2(tp.nw.x, tp.ne.x, tp.sw.x, tp.se.x, q.fx, q.fy, I[0], d[0]);
    blend2(tp.nw.y, tp.ne.y, tp.sw.y, tp.se.y, q.fx, q.fy, I[1], d[1]);
    blend2(tp.nw.z, tp.ne.z, tp.sw.z, tp.se.z, q.fx, q.fy, I[2], d[2]);
    const float Is[3] = {Is0, Is1, Is2};
    float2 gxy = make_float2(0.f, 0.f);            // sum_c sign(r_c) (dIx_c, dIy_c)
    float ga = 0.f, gb = 0.f, cost = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float r = AFF ? (Is[ch] - fmaf(c[F_EA], I[ch], c[F_BB])) : (Is[ch] - I[ch]);
        cost += fabsf(r);
        if constexpr (AFF) {
            const float s = (r > 0.f) ? 1.0f : ((r < 0.f) ? -1.0f : 0.0f);
            gxy = fma2(s, d[ch], gxy);
            ga = fmaf(s, I[ch], ga);
            gb += s;
        } else {
            // sign(r) * (dIx, dIy): XOR r's sign bit into the slope pair (one LOP3 each); sign(0) = 0 contributes
            // nothing, exactly like torch.sign in the reference's backward
            const uint32_t sb = __float_as_uint(r) & 0x80000000u;
            const float2 sd = make_float2(__uint_as_float(__float_as_uint(d[ch].x) ^ sb),
                                          __uint_as_float(__float_as_uint(d[ch].y) ^ sb));
            if (r != 0.f) gxy = __fadd2_rn(gxy, sd);
        }
    }
    // d cost / d (x_, y_) = (cu gx, cv gy);  gY = rho (gx_, gy_, -(gx_ x_ + gy_ y_))
    const float2 gb2 = __fmul2_rn(gxy, *reinterpret_cast<const float2*>(c + F_CU));
    const float2 gY = mul2(q.rho, gb2);
    const float gYz = q.live ? -fmaf(gY.x, q.xb, gY.y * q.yb) : 0.0f;
    const float2 zuv = mul2(q.zs, make_float2(q.uc, q.vc));
    A.cost += cost;
    A.gt01 = __fadd2_rn(A.gt01, gY);
    A.gt2 += gYz;
    A.m00 = fma2(gY.x, zuv, A.m00);  A.m02 = fmaf(gY.x, q.zs, A.m02);
    A.m10 = fma2(gY.y, zuv, A.m10);  A.m12 = fmaf(gY.y, q.zs, A.m12);
    A.m20 = fma2(gYz, zuv, A.m20);   A.m22 = fmaf(gYz, q.zs, A.m22);
    if (AFF) { A.ga = fmaf(c[F_EA], ga, A.ga); A.gb -= gb; }
    A.nv += 1.0f;
    // d cost / d k_b = gY . (R X) = gY . Y - gY . t, and gY . Y == 0 when the reciprocal is live
    gk -= fmaf(gY.x, c[F_TR(0)], fmaf(gY.y, c[F_TR(1)], gYz * c[F_TR(2)]));
    if (!q.live) gk += fmaf(gY.x, q.Yx, gY.y * q.Yy);
}
