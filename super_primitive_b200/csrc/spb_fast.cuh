// Hot-path body of the fused alignment kernel (sm_100a): per-warp bulk-async pipelines + lean math.
//
// Streaming operands (uv, logd, cached source rgb) of a tile are stored tile-major in HBM
// (SpbPair.tile_pack: header + five 128-word arrays = one contiguous 2576-byte block per tile) and travel
// global -> shared memory with ONE cp.async.bulk (TMA 1-D) per tile into a private ring of SPB_WSTAGES
// slots PER WARP, completing on an mbarrier; one elected lane issues the copy SPB_WSTAGES-1 tiles ahead.
// There is no CTA-wide barrier in the loop, the tile descriptor is the block header and the per-segment
// depth shifts live in shared memory, so nothing on a tile's critical path waits on global memory except
// the four bilinear taps of the target image.
//
// Math (same quantities as eval_point in spb_align.cu, reorganised to cut instructions):
//   Y = z (M u~) + t with M = R diag(1/fx, 1/fy, 1), u~ = (u - cx, v - cy, 1)   [folds unproject + rotate]
//   x_ = Yx rho, y_ = Yy rho, rho = 1/Yz (guarded);  xn = ax x_ + bx  (ax = fx' 2/(W-1), bx = cx' 2/(W-1) - 1)
//   d r_c / d x_ = cu dIx_c, d r_c / d y_ = cv dIy_c,  cu = -e^{-a} (Wl-1)/(W-1) fx'
//   d (x_, y_) / d xi  = [rho, 0, -rho x_, -x_ y_, 1 + x_^2, -y_ ; 0, rho, -rho y_, -(1 + y_^2), x_ y_, x_]
//   d (x_, y_) / d k_b = rho (x_ tz - tx, y_ tz - ty)      (exact: depth moves a point along its epipolar line;
//                                                            equals the chain-rule value gY.(R X) without its cancellation)
#pragma once
#include "spb_common.cuh"

#ifndef SPB_WSTAGES
#define SPB_WSTAGES 2                          // ring slots per warp
#endif
#define SPB_SLOT_WORDS SPB_PACK_WORDS           // one tile block: header {seg, cnt, unpadded start, -} + 5 arrays
#define SPB_NSHIFT 512                         // segments whose shift is cached in shared memory
#define SPB_FAST_DYN_SMEM (SPB_WARPS * SPB_WSTAGES * SPB_SLOT_WORDS * 4 + SPB_WARPS * SPB_WSTAGES * 8)

// context words, grouped so the hot loop reads them as eight 16-byte vectors (LDS.128):
//   V0..V2 = rows of [M | t] with M = R diag(1/fx, 1/fy, 1);  V3 = (ax, bx, ay, by);  V4 = (sx, sy, tau, ea);
//   V5 = (bb, -, cu, cv);  V6 = (cu^2, cu cv, cu cv, cv^2);  V7 = (cx, cy, -, -)
enum FastSlot {
    F_AX = 12, F_BX, F_AY, F_BY,   // xn = ax x_ + bx
    F_SX = 16, F_SY, F_TAU, F_EA,
    F_BB = 20, F_PAD0,
    F_CU = 22, F_CV,
    F_CUU = 24, F_CUV,
    F_CUV2 = 26, F_CVV,
    F_CX = 28, F_CY,
    F_N = 32
};
#define F_MAT(i, j) (4 * (i) + (j))
#define F_TR(i) (4 * (i) + 3)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// context for the fast path, filled by the first warp
__device__ __forceinline__ void fill_fast_ctx(float* s, const SpbPair& pr, const float* Ksrc, int H, int W) {
    if (threadIdx.x == 0) {
        const float ifx = 1.0f / Ksrc[0], ify = 1.0f / Ksrc[4], cx = Ksrc[2], cy = Ksrc[5];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float r0 = pr.pose[4 * i], r1 = pr.pose[4 * i + 1], r2 = pr.pose[4 * i + 2];
            s[F_MAT(i, 0)] = r0 * ifx;
            s[F_MAT(i, 1)] = r1 * ify;
            s[F_MAT(i, 2)] = r2;
            s[F_TR(i)] = pr.pose[4 * i + 3];
        }
        s[F_CX] = cx;
        s[F_CY] = cy;
        s[F_CY + 1] = 0.f;
        s[F_CY + 2] = 0.f;
    } else if (threadIdx.x == 1) {
        const float tiw = 2.0f * (1.0f / (float)(W - 1)), tih = 2.0f * (1.0f / (float)(H - 1));
        const float sx = 0.5f * (float)(pr.Wl - 1), sy = 0.5f * (float)(pr.Hl - 1);
        const float fxt = pr.K_trg[0], fyt = pr.K_trg[4], cxt = pr.K_trg[2], cyt = pr.K_trg[5];
        float a = 0.f, b = 0.f;
        if (pr.aff_src != nullptr && pr.aff_trg != nullptr) {
            a = pr.aff_trg[0] - pr.aff_src[0];
            b = pr.aff_trg[1] - pr.aff_src[1];
        }
        const float ea = expf(-a);
        s[F_AX] = fxt * tiw; s[F_BX] = fmaf(cxt, tiw, -1.0f);
        s[F_AY] = fyt * tih; s[F_BY] = fmaf(cyt, tih, -1.0f);
        s[F_SX] = sx; s[F_SY] = sy; s[F_TAU] = pr.tau; s[F_EA] = ea; s[F_BB] = b;
        const float cu = -ea * (sx * tiw) * fxt, cv = -ea * (sy * tih) * fyt;
        s[F_CU] = cu; s[F_CV] = cv; s[F_CUU] = cu * cu; s[F_CUV] = cu * cv; s[F_CUV2] = cu * cv; s[F_CVV] = cv * cv;
        s[F_PAD0] = 0.f;
    }
}


// geometry shared by both modes: returns validity, fills the projected quantities
struct Proj {
    float Yx, Yy, Yz, rho, xb, yb, fx, fy, zs;   // zs = z of the source point
    float uc, vc;                                // centred source pixel (u - cx, v - cy)
    int off;                                     // texel index of the north-west tap
    bool live;
};

__device__ __forceinline__ bool project_point(const float* __restrict__ c, uint32_t w, float logd, float shift, int Wl,
                                              Proj& q) {
    const float4* c4 = reinterpret_cast<const float4*>(c);
    const float4 r0 = c4[0], r1 = c4[1], r2 = c4[2], pa = c4[3], pb = c4[4];
    const float2 cc = *reinterpret_cast<const float2*>(c + F_CX);
    const float u = (float)(w & 0xffffu) - cc.x;
    const float v = (float)((w >> 16) & 0x7fffu) - cc.y;
    const float z = __expf(logd + shift);
    q.uc = u;
    q.vc = v;
    const float qx = fmaf(r0.x, u, fmaf(r0.y, v, r0.z));
    const float qy = fmaf(r1.x, u, fmaf(r1.y, v, r1.z));
    const float qz = fmaf(r2.x, u, fmaf(r2.y, v, r2.z));
    q.zs = z;
    q.Yx = fmaf(z, qx, r0.w);
    q.Yy = fmaf(z, qy, r1.w);
    q.Yz = fmaf(z, qz, r2.w);
    q.live = fabsf(q.Yz) > 1e-6f;                                  // guarded reciprocal, core/ops.py:22,33-34
    q.rho = q.live ? __fdividef(1.0f, q.Yz) : 1e-6f;
    q.xb = q.Yx * q.rho;
    q.yb = q.Yy * q.rho;
    const float xn = fmaf(q.xb, pa.x, pa.y);
    const float yn = fmaf(q.yb, pa.z, pa.w);
    // |xn| <= .99 && |yn| <= .99 && Yz > tau && z > 1e-7 && src_ok, folded into two compares
    const bool ok = (w >> 31) && (fmaxf(fabsf(xn), fabsf(yn)) <= 0.99f) && (fminf(q.Yz - pb.z, z - 1e-7f) > 0.0f);
    const float ix = fmaf(xn, pb.x, pb.x);
    const float iy = fmaf(yn, pb.y, pb.y);
    const float fxf = floorf(ix), fyf = floorf(iy);
    q.fx = ix - fxf;
    q.fy = iy - fyf;
    q.off = (int)fyf * Wl + (int)fxf;
    return ok;
}

// the four RGBA taps of one point; issued early (software pipelining), consumed by point_grad / point_gn
#ifndef SPB_TAP_L2_256
#define SPB_TAP_L2_256 1                        // 1: gathers ask L2 to fill 256-byte granules on a miss
#endif
__device__ __forceinline__ float4 ldg_l2_256(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
struct Taps4 {
    float4 nw, ne, sw, se;
};
__device__ __forceinline__ void load_taps(const float4* __restrict__ trg, int Wl, int off, Taps4& t) {
    const float4* p0 = trg + off;
#if SPB_TAP_L2_256
    t.nw = ldg_l2_256(p0); t.ne = ldg_l2_256(p0 + 1); t.sw = ldg_l2_256(p0 + Wl); t.se = ldg_l2_256(p0 + Wl + 1);
#else
    t.nw = __ldg(p0); t.ne = __ldg(p0 + 1); t.sw = __ldg(p0 + Wl); t.se = __ldg(p0 + Wl + 1);
#endif
}

// ---- Gauss-Newton mode, scalar formulation (used by the 8-column / affine variant) ---------------------
// (the 6-column GN path and the gradient mode live in spb_gn_packed.cuh, written with packed FP32)
// acc layout = upper triangle of the NPxNP pose block (row-major packed), g_p[NP], cost, wcost, nvalid
// seg layout = B column [NP], D, g_d
template <int NP, int NACC, int NSEG>
__device__ __forceinline__ void point_gn(const float* __restrict__ c, const Taps4& tp,
                                         const Proj& q, float Is0, float Is1, float Is2, float irls_eps,
                                         float (&acc)[NACC], float (&seg)[NSEG]) {
    const float4 nw = tp.nw, ne = tp.ne, sw = tp.sw, se = tp.se;
    float I[3], dx[3], dy[3];
    blend(nw.x, ne.x, sw.x, se.x, q.fx, q.fy, I[0], dx[0], dy[0]);
    blend(nw.y, ne.y, sw.y, se.y, q.fx, q.fy, I[1], dx[1], dy[1]);
    blend(nw.z, ne.z, sw.z, se.z, q.fx, q.fy, I[2], dx[2], dy[2]);
    const float Is[3] = {Is0, Is1, Is2};
    const float ea = c[F_EA], bb = c[F_BB];
    float Guu = 0.f, Guv = 0.f, Gvv = 0.f, hu = 0.f, hv = 0.f, cost = 0.f, wcost = 0.f;
    float Aau = 0.f, Aav = 0.f, Abu = 0.f, Abv = 0.f, Aaa = 0.f, Aab = 0.f, Abb = 0.f, ha = 0.f, hb = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float r = Is[ch] - fmaf(ea, I[ch], bb);
        const float ar = fabsf(r);
        const float wgt = __fdividef(1.0f, fmaxf(ar, irls_eps));
        const float wr = wgt * r;
        cost += ar;
        wcost = fmaf(wr, r, wcost);
        const float wu = wgt * dx[ch], wv = wgt * dy[ch];
        Guu = fmaf(wu, dx[ch], Guu);
        Guv = fmaf(wu, dy[ch], Guv);
        Gvv = fmaf(wv, dy[ch], Gvv);
        hu = fmaf(wr, dx[ch], hu);
        hv = fmaf(wr, dy[ch], hv);
        if constexpr (NP == 8) {
            const float ja = ea * I[ch];                               // d r / d a_t ; d r / d b_t = -1
            Aau = fmaf(wu, ja, Aau); Aav = fmaf(wv, ja, Aav);
            Abu -= wu; Abv -= wv;
            Aaa = fmaf(wgt * ja, ja, Aaa); Aab = fmaf(-wgt, ja, Aab); Abb += wgt;
            ha = fmaf(wr, ja, ha); hb -= wr;
        }
    }
    Guu *= c[F_CUU]; Guv *= c[F_CUV]; Gvv *= c[F_CVV]; hu *= c[F_CU]; hv *= c[F_CV];
    // mu = d x_/d(xi,k), mv = d y_/d(xi,k); structural zeros mu[1] = mv[0] = 0 are exploited below
    const float rho = q.rho, xb = q.xb, yb = q.yb;
    float mu2, mu3, mu4, mu5, mu6, mv2, mv3, mv4, mv5, mv6;
    if (q.live) {
        const float xy = xb * yb;
        mu2 = -rho * xb; mu3 = -xy;                 mu4 = fmaf(xb, xb, 1.0f); mu5 = -yb;
        mv2 = -rho * yb; mv3 = -fmaf(yb, yb, 1.0f); mv4 = xy;                 mv5 = xb;
        mu6 = rho * fmaf(xb, c[F_TR(2)], -c[F_TR(0)]);
        mv6 = rho * fmaf(yb, c[F_TR(2)], -c[F_TR(1)]);
    } else {                                         // constant reciprocal: no z-derivative (never taken in practice)
        mu2 = 0.f; mu3 = 0.f; mu4 = rho * q.Yz; mu5 = -rho * q.Yy;
        mv2 = 0.f; mv3 = -rho * q.Yz; mv4 = 0.f; mv5 = rho * q.Yx;
        mu6 = rho * (q.Yx - c[F_TR(0)]);
        mv6 = rho * (q.Yy - c[F_TR(1)]);
    }
    // pu = Guu mu + Guv mv ; pv = Guv mu + Gvv mv   (index 0: mv0 = 0, index 1: mu1 = 0)
    const float pu0 = Guu * rho;
    const float pu1 = Guv * rho, pv1 = Gvv * rho;
    const float pu2 = fmaf(Guu, mu2, Guv * mv2), pv2 = fmaf(Guv, mu2, Gvv * mv2);
    const float pu3 = fmaf(Guu, mu3, Guv * mv3), pv3 = fmaf(Guv, mu3, Gvv * mv3);
    const float pu4 = fmaf(Guu, mu4, Guv * mv4), pv4 = fmaf(Guv, mu4, Gvv * mv4);
    const float pu5 = fmaf(Guu, mu5, Guv * mv5), pv5 = fmaf(Guv, mu5, Gvv * mv5);
    const float pu6 = fmaf(Guu, mu6, Guv * mv6), pv6 = fmaf(Guv, mu6, Gvv * mv6);
    // row r of the packed upper triangle starts at R0(r)
    constexpr int W8 = NP;                           // columns of the pose block
#define TRI(r, cc) ((r) * W8 - (r) * ((r) - 1) / 2 + ((cc) - (r)))
    // row 0: mu0 = rho, mv0 = 0
    acc[TRI(0, 0)] = fmaf(rho, pu0, acc[TRI(0, 0)]);
    acc[TRI(0, 1)] = fmaf(rho, pu1, acc[TRI(0, 1)]);
    acc[TRI(0, 2)] = fmaf(rho, pu2, acc[TRI(0, 2)]);
    acc[TRI(0, 3)] = fmaf(rho, pu3, acc[TRI(0, 3)]);
    acc[TRI(0, 4)] = fmaf(rho, pu4, acc[TRI(0, 4)]);
    acc[TRI(0, 5)] = fmaf(rho, pu5, acc[TRI(0, 5)]);
    // row 1: mu1 = 0, mv1 = rho
    acc[TRI(1, 1)] = fmaf(rho, pv1, acc[TRI(1, 1)]);
    acc[TRI(1, 2)] = fmaf(rho, pv2, acc[TRI(1, 2)]);
    acc[TRI(1, 3)] = fmaf(rho, pv3, acc[TRI(1, 3)]);
    acc[TRI(1, 4)] = fmaf(rho, pv4, acc[TRI(1, 4)]);
    acc[TRI(1, 5)] = fmaf(rho, pv5, acc[TRI(1, 5)]);
    // rows 2..5
    acc[TRI(2, 2)] = fmaf(mu2, pu2, fmaf(mv2, pv2, acc[TRI(2, 2)]));
    acc[TRI(2, 3)] = fmaf(mu2, pu3, fmaf(mv2, pv3, acc[TRI(2, 3)]));
    acc[TRI(2, 4)] = fmaf(mu2, pu4, fmaf(mv2, pv4, acc[TRI(2, 4)]));
    acc[TRI(2, 5)] = fmaf(mu2, pu5, fmaf(mv2, pv5, acc[TRI(2, 5)]));
    acc[TRI(3, 3)] = fmaf(mu3, pu3, fmaf(mv3, pv3, acc[TRI(3, 3)]));
    acc[TRI(3, 4)] = fmaf(mu3, pu4, fmaf(mv3, pv4, acc[TRI(3, 4)]));
    acc[TRI(3, 5)] = fmaf(mu3, pu5, fmaf(mv3, pv5, acc[TRI(3, 5)]));
    acc[TRI(4, 4)] = fmaf(mu4, pu4, fmaf(mv4, pv4, acc[TRI(4, 4)]));
    acc[TRI(4, 5)] = fmaf(mu4, pu5, fmaf(mv4, pv5, acc[TRI(4, 5)]));
    acc[TRI(5, 5)] = fmaf(mu5, pu5, fmaf(mv5, pv5, acc[TRI(5, 5)]));
    constexpr int NA = NP * (NP + 1) / 2;
    if constexpr (NP == 8) {
        // affine columns: A[i][6] = mu_i cu Aau + mv_i cv Aav ; A[i][7] = mu_i cu Abu + mv_i cv Abv
        const float au = c[F_CU] * Aau, av = c[F_CV] * Aav, bu = c[F_CU] * Abu, bv = c[F_CV] * Abv;
        acc[TRI(0, 6)] = fmaf(rho, au, acc[TRI(0, 6)]);  acc[TRI(0, 7)] = fmaf(rho, bu, acc[TRI(0, 7)]);
        acc[TRI(1, 6)] = fmaf(rho, av, acc[TRI(1, 6)]);  acc[TRI(1, 7)] = fmaf(rho, bv, acc[TRI(1, 7)]);
        acc[TRI(2, 6)] = fmaf(mu2, au, fmaf(mv2, av, acc[TRI(2, 6)]));  acc[TRI(2, 7)] = fmaf(mu2, bu, fmaf(mv2, bv, acc[TRI(2, 7)]));
        acc[TRI(3, 6)] = fmaf(mu3, au, fmaf(mv3, av, acc[TRI(3, 6)]));  acc[TRI(3, 7)] = fmaf(mu3, bu, fmaf(mv3, bv, acc[TRI(3, 7)]));
        acc[TRI(4, 6)] = fmaf(mu4, au, fmaf(mv4, av, acc[TRI(4, 6)]));  acc[TRI(4, 7)] = fmaf(mu4, bu, fmaf(mv4, bv, acc[TRI(4, 7)]));
        acc[TRI(5, 6)] = fmaf(mu5, au, fmaf(mv5, av, acc[TRI(5, 6)]));  acc[TRI(5, 7)] = fmaf(mu5, bu, fmaf(mv5, bv, acc[TRI(5, 7)]));
        acc[TRI(6, 6)] += Aaa; acc[TRI(6, 7)] += Aab; acc[TRI(7, 7)] += Abb;
        acc[NA + 6] += ha; acc[NA + 7] += hb;
        seg[6] = fmaf(mu6, au, fmaf(mv6, av, seg[6]));
        seg[7] = fmaf(mu6, bu, fmaf(mv6, bv, seg[7]));
    }
#undef TRI
    acc[NA + 0] = fmaf(rho, hu, acc[NA + 0]);
    acc[NA + 1] = fmaf(rho, hv, acc[NA + 1]);
    acc[NA + 2] = fmaf(mu2, hu, fmaf(mv2, hv, acc[NA + 2]));
    acc[NA + 3] = fmaf(mu3, hu, fmaf(mv3, hv, acc[NA + 3]));
    acc[NA + 4] = fmaf(mu4, hu, fmaf(mv4, hv, acc[NA + 4]));
    acc[NA + 5] = fmaf(mu5, hu, fmaf(mv5, hv, acc[NA + 5]));
    acc[NA + NP + 0] += cost;
    acc[NA + NP + 1] += wcost;
    acc[NA + NP + 2] += 1.0f;
    // depth column of this tile's segment: B[i] = J_i^T W j_d, D = j_d^T W j_d, g_d = j_d^T W r
    seg[0] = fmaf(rho, pu6, seg[0]);
    seg[1] = fmaf(rho, pv6, seg[1]);
    seg[2] = fmaf(mu2, pu6, fmaf(mv2, pv6, seg[2]));
    seg[3] = fmaf(mu3, pu6, fmaf(mv3, pv6, seg[3]));
    seg[4] = fmaf(mu4, pu6, fmaf(mv4, pv6, seg[4]));
    seg[5] = fmaf(mu5, pu6, fmaf(mv5, pv6, seg[5]));
    seg[NP] = fmaf(mu6, pu6, fmaf(mv6, pv6, seg[NP]));
    seg[NP + 1] = fmaf(mu6, hu, fmaf(mv6, hv, seg[NP + 1]));
}
