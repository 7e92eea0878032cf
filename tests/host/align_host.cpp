// TEST INFRASTRUCTURE ONLY.  Host build of the per-point arithmetic of the fused alignment kernel: project_point,
// point_gn6_packed, point_grad_packed and fill_fast_ctx are compiled by g++ from the kernel's own headers (CUDA
// built-ins replaced by tests/host/cuda_shim.h) and driven point by point; every point's contribution is added in
// float64, so what is compared with the float64 closed form (oracle/closed_form.py) is the kernel's per-point math,
// not a summation order.  Never loaded by the product path.
#include "cuda_shim.h"
#include "../../super_primitive_b200/csrc/spb_gn_packed.cuh"

template <bool AFF>
static void run_gn(int P, const uint32_t* uv, const float* logd, const float* Is, const int32_t* seg, const float* shift,
                   const float4* trg, int Wl, const float* c, float irls_eps, double* out_pair, double* out_seg) {
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
        float a[30], s[8];
        A.store(a);
        S.store(s);
        for (int j = 0; j < 28; ++j) out_pair[j] += a[j];
        for (int j = 0; j < 8; ++j) out_seg[8 * seg[i] + j] += s[j];
    }
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
    (void)N;
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
        if (aff) run_gn<true>(P, uv, logd, Is, seg, shift, trg, Wl, c, irls_eps, out_pair, out_seg);
        else run_gn<false>(P, uv, logd, Is, seg, shift, trg, Wl, c, irls_eps, out_pair, out_seg);
    } else {
        if (aff) run_grad<true>(P, uv, logd, Is, seg, shift, trg, Wl, c, out_pair, out_seg);
        else run_grad<false>(P, uv, logd, Is, seg, shift, trg, Wl, c, out_pair, out_seg);
    }
    return 0;
}
