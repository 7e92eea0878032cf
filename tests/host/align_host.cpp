// TEST INFRASTRUCTURE ONLY.  Host build of the per-point arithmetic of the fused alignment kernel: project_point,
// point_gn6_packed, point_grad_packed and fill_fast_ctx are compiled by g++ from the kernel's own headers (CUDA
// built-ins replaced by tests/host/cuda_shim.h) and driven point by point; every point's contribution is added in
// float64, so what is compared with the float64 closed form (oracle/closed_form.py) is the kernel's per-point math,
// not a summation order.  Never loaded by the product path.
#include "cuda_shim.h"
#include "../../super_primitive_b200/csrc/spb_gn_packed.cuh"

// The kernel keeps rows 0..2 of the pose block + g_0..g_3 per run of same-segment tiles and the rotation block, g_4,
// g_5, cost per warp; the finalize kernel derives the segment's depth column from the run sums (gn6_segment_record).
// Here: per-segment sums in float64, then the same derivation, returned in the canonical order the test expects
// (21 A + 6 g_p + cost; per segment B[6], D, g_d).
template <bool AFF>
static void run_gn(int P, int N, const uint32_t* uv, const float* logd, const float* Is, const int32_t* seg,
                   const float* shift, const float4* trg, int Wl, const float* c, const float* pose16, float irls_eps,
                   double* out_pair, double* out_seg) {
    double* runs = new double[(size_t)N * SPB_GN6_NRUN]();
    double glob[SPB_GN6_NACC] = {0};
    for (int i = 0; i < P; ++i) {
        Proj q;
        bool ok = project_point(c, uv[i], logd[i], shift[seg[i]], Wl, q);
        ok = ok && q.live;
        if (!ok) continue;
        Taps4 tp;
        load_taps(trg, Wl, q.off, tp);
        GnAcc6 A;
        GnSeg6 S;
        A.zero();
        S.zero();
        point_gn6_packed<AFF>(c, tp, q, Is[i], Is[P + i], Is[2 * P + i], irls_eps, A, S);
        float a[SPB_GN6_NACC], s[SPB_GN6_NRUN];
        A.store(a);
        S.store(s);
        for (int j = 0; j < SPB_GN6_NACC; ++j) glob[j] += a[j];
        for (int j = 0; j < SPB_GN6_NRUN; ++j) runs[(size_t)seg[i] * SPB_GN6_NRUN + j] += s[j];
    }
    const double t[3] = {pose16[3], pose16[7], pose16[11]};
    double rows[SPB_GN6_NRUN] = {0};
    for (int b = 0; b < N; ++b) {
        gn6_segment_record(runs + (size_t)b * SPB_GN6_NRUN, t, out_seg + 8 * b);
        for (int j = 0; j < SPB_GN6_NRUN; ++j) rows[j] += runs[(size_t)b * SPB_GN6_NRUN + j];
    }
    // canonical packed upper triangle of the 6x6 block: row 0 (6), row 1 (5), row 2 (4), row 3 (3), row 4 (2), row 5 (1)
    for (int j = 0; j < 15; ++j) out_pair[j] = rows[j];
    for (int j = 0; j < 6; ++j) out_pair[15 + j] = glob[j];
    for (int j = 0; j < 4; ++j) out_pair[21 + j] = rows[15 + j];
    out_pair[25] = glob[6];
    out_pair[26] = glob[7];
    out_pair[27] = glob[8];
    delete[] runs;
}

template <bool AFF>
static void run_grad(int P, const uint32_t* uv, const float* logd, const float* Is, const int32_t* seg, const float* shift,
                     const float4* trg, int Wl, const float* c, double* out_pair, double* out_seg) {
    for (int i = 0; i < P; ++i) {
        Proj q;
        if (!project_point(c, uv[i], logd[i], shift[seg[i]], Wl, q)) continue;
        Taps4 tp;
        load_taps(trg, Wl, q.off, tp);
        GradAcc G;
        G.zero();
        float gk = 0.f;
        point_grad_packed<AFF>(c, tp, q, Is[i], Is[P + i], Is[2 * P + i], G, gk);
        float g[16];
        G.store(g);
        for (int j = 0; j < 16; ++j) out_pair[j] += g[j];
        out_seg[seg[i]] += gk;
    }
}

// mode 0: gradient sums (16 per pair in the kernel's canonical order, 1 per segment); mode 1: GN sums (21 A + 6 g_p + cost,
// 8 per segment: B[6], D, g_d).  Raw sums, before the finalize kernels' normalisation / column scaling.
extern "C" int align_points_host(int mode, int P, const uint32_t* uv, const float* logd, const float* Is,
                                 const int32_t* seg, int N, const float* shift, const float* trg_rgba, int Hl, int Wl,
                                 const float* K_src, const float* K_trg, const float* pose16, const float* aff_src,
                                 const float* aff_trg, float tau, int H, int W, float irls_eps, double* out_pair,
                                 double* out_seg) {
    SpbPair pr;
    memset(&pr, 0, sizeof(pr));
    pr.K_trg = K_trg; pr.pose = pose16; pr.aff_src = aff_src; pr.aff_trg = aff_trg; pr.Hl = Hl; pr.Wl = Wl; pr.tau = tau;
    float c[F_N];
    memset(c, 0, sizeof(c));
    threadIdx.x = 0; fill_fast_ctx(c, pr, K_src, H, W);      // the kernel fills the context with its first two threads
    threadIdx.x = 1; fill_fast_ctx(c, pr, K_src, H, W);
    const float4* trg = reinterpret_cast<const float4*>(trg_rgba);
    const bool aff = aff_src != nullptr && aff_trg != nullptr;
    if (mode == 1) {
        if (aff) run_gn<true>(P, N, uv, logd, Is, seg, shift, trg, Wl, c, pose16, irls_eps, out_pair, out_seg);
        else run_gn<false>(P, N, uv, logd, Is, seg, shift, trg, Wl, c, pose16, irls_eps, out_pair, out_seg);
    } else {
        if (aff) run_grad<true>(P, uv, logd, Is, seg, shift, trg, Wl, c, out_pair, out_seg);
        else run_grad<false>(P, uv, logd, Is, seg, shift, trg, Wl, c, out_pair, out_seg);
    }
    return 0;
}
