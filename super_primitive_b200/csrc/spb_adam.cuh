// Device-resident first-order iteration: torch.optim.Adam on (twist increment, log-depth seeds, target affine)
// with the bookkeeping of the reference's tracker (odometery/odometery.py:303-310 optimiser, :386-403 loop): the
// pose parameter is a twist `delta` at the identity (lietorch ordering: translation first), the cost is evaluated at
// Exp(delta) T, and after every optimiser step the increment is folded into the pose (T <- Exp(delta) T) and
// re-zeroed while its Adam moments persist.  The log-depth seeds (odometery/two_frame_sfm.py:117-121) and the
// target brightness terms are plain Adam parameters.  The update arithmetic follows torch.optim.Adam's
// single-tensor path (lerp on the first moment, mul/addcmul on the second, bias corrections in float64 folded
// into float32 scalars, addcdiv) so a float32 torch loop is reproduced to rounding.
// lietorch itself is unpinned upstream: the retraction is the closed-form SE(3) exponential of
// spb_lm.cuh; oracle/adam_loop.py restates the loop with torch.linalg.matrix_exp.
#pragma once
#include "spb_lm.cuh"

struct SpbAdamHyper {
    double lr_pose, lr_k, lr_aff, beta1, beta2, eps;
};

// one parameter: returns the increment  -step_size * m_hat / (sqrt(v_hat) + eps)
__device__ __forceinline__ float adam_increment(float g, float& m, float& v, float b1, float b2, float step_size,
                                                float bc2_sqrt, float eps) {
    m = m + (1.0f - b1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (1.0f - b2) * g * g;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;       // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    return -step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// out_pair / out_gk are deliberately not __restrict__ (the fused kernel writes them earlier in the same launch).
// Block size: >= 64 threads, a multiple of 32.
__device__ __forceinline__ void adam_update_body(const float* out_pair, const float* out_gk,
                                                 const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_cnt,
                                                 int with_affine, const SpbAdamHyper h, float* __restrict__ poses,
                                                 float* __restrict__ k, float* __restrict__ aff_trg,
                                                 float* __restrict__ adam_pair, float* __restrict__ adam_seg) {
    const int p = blockIdx.x;
    const int so = seg_off[p], n = seg_cnt[p];
    float* ap = adam_pair + (size_t)p * SPB_ADAM_PAIR;
    const float* gp = out_pair + (size_t)p * SPB_PAIR_NOUT;
    float* pose = poses + (size_t)p * 16;
    __shared__ float s_sc[4];                            // step sizes (pose, k, affine), sqrt of bias correction 2
    if (threadIdx.x == 0) {
        const float t = ap[0] + 1.0f;
        ap[0] = t;
        const double bc1 = 1.0 - pow(h.beta1, (double)t), bc2 = 1.0 - pow(h.beta2, (double)t);
        s_sc[0] = (float)(h.lr_pose / bc1);
        s_sc[1] = (float)(h.lr_k / bc1);
        s_sc[2] = (float)(h.lr_aff / bc1);
        s_sc[3] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float b1 = (float)h.beta1, b2 = (float)h.beta2, eps = (float)h.eps, bc2s = s_sc[3];
    // log-depth seeds
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        float* as = adam_seg + (size_t)(so + b) * SPB_ADAM_SEG;
        float m = as[0], v = as[1];
        k[so + b] += adam_increment(out_gk[so + b], m, v, b1, b2, s_sc[1], bc2s, eps);
        as[0] = m;
        as[1] = v;
    }
    // target brightness terms (second warp, so it runs beside the pose update)
    if (with_affine && aff_trg && threadIdx.x >= 32 && threadIdx.x < 34) {
        const int i = threadIdx.x - 32;
        float m = ap[1 + 6 + i], v = ap[9 + 6 + i];
        aff_trg[2 * p + i] += adam_increment(gp[13 + i], m, v, b1, b2, s_sc[2], bc2s, eps);
        ap[1 + 6 + i] = m;
        ap[9 + 6 + i] = v;
    }
    // pose: gradient with respect to the twist at the identity, Adam increment, retraction
    if (threadIdx.x == 0) {
        float R[3][3], G[3][3], t0[3], gt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                R[i][j] = pose[4 * i + j];
                G[i][j] = gp[4 + 3 * i + j];                 // d cost / d R
            }
            t0[i] = pose[4 * i + 3];
            gt[i] = gp[1 + i];                               // d cost / d t
        }
        // T(delta) = Exp(delta) T:  dR = [phi]x R, dt = phi x t + tau  =>  g_tau = g_t,
        // g_phi = sum_j R[:,j] x G[:,j] + t x g_t
        float g6[6] = {gt[0], gt[1], gt[2], 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            g6[3] += R[1][j] * G[2][j] - R[2][j] * G[1][j];
            g6[4] += R[2][j] * G[0][j] - R[0][j] * G[2][j];
            g6[5] += R[0][j] * G[1][j] - R[1][j] * G[0][j];
        }
        g6[3] += t0[1] * gt[2] - t0[2] * gt[1];
        g6[4] += t0[2] * gt[0] - t0[0] * gt[2];
        g6[5] += t0[0] * gt[1] - t0[1] * gt[0];
        double xi[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float m = ap[1 + i], v = ap[9 + i];
            xi[i] = (double)adam_increment(g6[i], m, v, b1, b2, s_sc[0], bc2s, eps);
            ap[1 + i] = m;
            ap[9 + i] = v;
        }
        double E[12];
        se3_exp_d(xi, E);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double v = E[4 * i] * (double)(j < 3 ? R[0][j] : t0[0]) + E[4 * i + 1] * (double)(j < 3 ? R[1][j] : t0[1]) +
                           E[4 * i + 2] * (double)(j < 3 ? R[2][j] : t0[2]);
                if (j == 3) v += E[4 * i + 3];
                pose[4 * i + j] = (float)v;
            }
        }
        pose[12] = 0.f; pose[13] = 0.f; pose[14] = 0.f; pose[15] = 1.f;
    }
}
