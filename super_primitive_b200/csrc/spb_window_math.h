// Per-item arithmetic of the mapping-window update (spb_window_update, include/spb200.h): what ONE thread does for
// one edge / one frame / one log-depth seed.  Plain C++ over <math.h>, so the same functions are compiled by nvcc
// into k_window_update (spb_window.cu) and by g++ into the host harness of the CPU tests
// (tests/host/window_host.cpp) -- the kernel adds only the thread mapping and the barriers between the phases.
//
// Reference: odometery/odometery.py:687-915 (mapping loop), :576-648 (optimiser), lie/lie_algebra.py:41-135
// (renormalise_se3 = matrix -> quaternion -> matrix, pytorch3d's conversion).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/spb200.h"

#ifdef __CUDACC__
#define SPB_HD __host__ __device__ __forceinline__
#else
#define SPB_HD static inline
#endif

struct SpbWinHyper {
    double lr_pose, lr_k, lr_aff, beta1, beta2, eps, stop_tol;
};

// step sizes of the iteration that is about to be applied (torch.optim.Adam, single-tensor path: the bias
// corrections are Python floats, folded into float32 scalars)
struct SpbWinStep {
    float step_pose, step_k, step_aff, bc2_sqrt, b1, b2, eps;
};

SPB_HD SpbWinStep win_step_sizes(const SpbWinHyper& h, float t /* 1-based step count */) {
    const double bc1 = 1.0 - pow(h.beta1, (double)t), bc2 = 1.0 - pow(h.beta2, (double)t);
    SpbWinStep s;
    s.step_pose = (float)(h.lr_pose / bc1);
    s.step_k = (float)(h.lr_k / bc1);
    s.step_aff = (float)(h.lr_aff / bc1);
    s.bc2_sqrt = (float)sqrt(bc2);
    s.b1 = (float)h.beta1;
    s.b2 = (float)h.beta2;
    s.eps = (float)h.eps;
    return s;
}

// one parameter: returns the increment -step_size * m / (sqrt(v) / sqrt(bc2) + eps)   (same as spb_adam.cuh)
SPB_HD float win_adam_inc(float g, float* m, float* v, const SpbWinStep& s, float step_size) {
    *m = *m + (1.0f - s.b1) * (g - *m);
    *v = *v * s.b2 + (1.0f - s.b2) * g * g;
    const float denom = sqrtf(*v) / s.bc2_sqrt + s.eps;
    return -step_size * (*m / denom);
}

// SE(3) exponential of a twist (tau, phi) -> 3x4 [R | t] row-major (stride 4), float64
SPB_HD void win_se3_exp(const double* xi, double* T) {
    const double px = xi[3], py = xi[4], pz = xi[5];
    const double th2 = px * px + py * py + pz * pz;
    double A, B, C;
    if (th2 < 1e-16) {
        A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        const double th = sqrt(th2);
        const double sn = sin(th), cs = cos(th);
        A = sn / th; B = (1.0 - cs) / th2; C = (th - sn) / (th2 * th);
    }
    const double K[9] = {0, -pz, py, pz, 0, -px, -py, px, 0};
    double K2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    for (int i = 0; i < 3; ++i) {
        double vt = 0.0;
        for (int j = 0; j < 3; ++j) {
            const double I = (i == j) ? 1.0 : 0.0;
            T[4 * i + j] = I + A * K[3 * i + j] + B * K2[3 * i + j];
            vt += (I + B * K[3 * i + j] + C * K2[3 * i + j]) * xi[j];
        }
        T[4 * i + 3] = vt;
    }
}

// Twist gradients of one edge at delta = 0 from d cost / d [R | t] of its relative pose P = [R | t]:
//   target side, P(delta) = Exp(delta) P   : dR = [phi]x R, dt = phi x t + tau
//       g_tau = g_t,  g_phi = sum_j R[:,j] x G[:,j] + t x g_t
//   source side, P(delta) = P inv(Exp(delta)) = P - P delta^ + ... : dR = -R [phi]x, dt = -R tau
//       g_tau = -R^T g_t,  g_phi = -vee(M - M^T), M = R^T G
// gp = the edge's SPB_PAIR_NOUT gradient row ([1..3] d/dt, [4..12] d/dR row-major), tw = [6 target | 6 source]
SPB_HD void win_edge_twists(const float* gp, const float* pose, float* tw) {
    float R[3][3], G[3][3], t[3], gt[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            R[i][j] = pose[4 * i + j];
            G[i][j] = gp[4 + 3 * i + j];
        }
        t[i] = pose[4 * i + 3];
        gt[i] = gp[1 + i];
    }
    float a[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < 3; ++j) {
        a[0] += R[1][j] * G[2][j] - R[2][j] * G[1][j];
        a[1] += R[2][j] * G[0][j] - R[0][j] * G[2][j];
        a[2] += R[0][j] * G[1][j] - R[1][j] * G[0][j];
    }
    tw[0] = gt[0]; tw[1] = gt[1]; tw[2] = gt[2];
    tw[3] = a[0] + t[1] * gt[2] - t[2] * gt[1];
    tw[4] = a[1] + t[2] * gt[0] - t[0] * gt[2];
    tw[5] = a[2] + t[0] * gt[1] - t[1] * gt[0];
    float M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = R[0][i] * G[0][j] + R[1][i] * G[1][j] + R[2][i] * G[2][j];
    for (int i = 0; i < 3; ++i) tw[6 + i] = -(R[0][i] * gt[0] + R[1][i] * gt[1] + R[2][i] * gt[2]);
    tw[9] = -(M[2][1] - M[1][2]);
    tw[10] = -(M[0][2] - M[2][0]);
    tw[11] = -(M[1][0] - M[0][1]);
}

// renormalise_se3 (lie/lie_algebra.py:41-48): R <- quaternion_to_matrix(_matrix_to_quaternion_t(R)), float32.
// pytorch3d's conversion: four candidate quaternions, the best-conditioned one (largest |component|) wins.
SPB_HD void win_renormalise(float* T) {
    const float m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
    float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    int best = 0;
    for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
    for (int i = 1; i < 4; ++i)
        if (qa[i] > qa[best]) best = i;                   // argmax: first maximum
    float c[4];
    if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
    else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
    const float den = 2.0f * (qa[best] > 0.1f ? qa[best] : 0.1f);
    const float r = c[0] / den, i = c[1] / den, j = c[2] / den, k = c[3] / den;
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    T[0] = 1.0f - two_s * (j * j + k * k); T[1] = two_s * (i * j - k * r);        T[2] = two_s * (i * k + j * r);
    T[4] = two_s * (i * j + k * r);        T[5] = 1.0f - two_s * (i * i + k * k); T[6] = two_s * (j * k - i * r);
    T[8] = two_s * (i * k - j * r);        T[9] = two_s * (j * k + i * r);        T[10] = 1.0f - two_s * (i * i + j * j);
}

// T <- T inv(Exp(xi)) = T Exp(-xi)   (odometery/odometery.py:866), float64 product rounded to float32
SPB_HD void win_fold_increment(float* T, const float* xi) {
    double nx[6], E[12];
    for (int i = 0; i < 6; ++i) nx[i] = -(double)xi[i];
    win_se3_exp(nx, E);
    float out[12];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 4; ++j) {
            double v = (double)T[4 * i] * E[j] + (double)T[4 * i + 1] * E[4 + j] + (double)T[4 * i + 2] * E[8 + j];
            if (j == 3) v += (double)T[4 * i + 3];
            out[4 * i + j] = (float)v;
        }
    }
    for (int i = 0; i < 12; ++i) T[i] = out[i];
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// pose = inv(T_trg) T_src for rigid T:  R = Rt^T Rs,  t = Rt^T (ts - tt)
SPB_HD void win_edge_pose(const float* Tt, const float* Ts, float* pose) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            pose[4 * i + j] = (float)((double)Tt[i] * Ts[j] + (double)Tt[4 + i] * Ts[4 + j] + (double)Tt[8 + i] * Ts[8 + j]);
        pose[4 * i + 3] = (float)((double)Tt[i] * ((double)Ts[3] - Tt[3]) + (double)Tt[4 + i] * ((double)Ts[7] - Tt[7]) +
                                  (double)Tt[8 + i] * ((double)Ts[11] - Tt[11]));
    }
    pose[12] = 0.f; pose[13] = 0.f; pose[14] = 0.f; pose[15] = 1.f;
}

// loss of window w = sum_e w_e cost_e, edges in index order
SPB_HD float win_loss(const SpbWindow& w, int win, const float* out_pair) {
    float loss = 0.f;
    for (int e = w.win_edge_off[win]; e < w.win_edge_off[win + 1]; ++e) loss += w.edge_w[e] * out_pair[(size_t)e * SPB_PAIR_NOUT];
    return loss;
}

// Frame f of window `win`: gather its twist / brightness gradient over the window's edges (index order), apply Adam
// where the frame is optimised, fold the increment into T_f, renormalise (the reference renormalises every frame
// of the window every iteration, optimised or not: :861-868, :876-882).
SPB_HD void win_frame_step(const SpbWindow& w, int win, int f, const float* out_pair, const SpbWinStep& s) {
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = w.win_edge_off[win]; e < w.win_edge_off[win + 1]; ++e) {
        const float we = w.edge_w[e];
        const float* tw = w.edge_tw + (size_t)e * 12;
        const float* gp = out_pair + (size_t)e * SPB_PAIR_NOUT;
        if (w.edge_trg[e] == f) {
            for (int i = 0; i < 6; ++i) g[i] += we * tw[i];
            g[6] += we * gp[13];
            g[7] += we * gp[14];
        }
        if (w.edge_src[e] == f) {
            for (int i = 0; i < 6; ++i) g[i] += we * tw[6 + i];
            g[6] -= we * gp[13];
            g[7] -= we * gp[14];
        }
    }
    const uint8_t fl = w.frame_flags[f];
    float* ad = w.adam_frame + (size_t)f * SPB_WIN_ADAM_FRAME;
    float* T = w.frame_T + (size_t)f * 16;
    if (fl & SPB_WIN_OPT_POSE) {
        float xi[6];
        for (int i = 0; i < 6; ++i) xi[i] = win_adam_inc(g[i], ad + i, ad + 8 + i, s, s.step_pose);
        win_fold_increment(T, xi);
    }
    win_renormalise(T);
    if ((fl & SPB_WIN_OPT_AFF) && w.frame_aff) {
        for (int i = 0; i < 2; ++i)
            w.frame_aff[2 * f + i] += win_adam_inc(g[6 + i], ad + 6 + i, ad + 14 + i, s, s.step_aff);
    }
}

// Seed b of keyframe f: gradient = sum over the keyframe's outgoing edges (index order), Adam
SPB_HD void win_seed_step(const SpbWindow& w, int win, int f, int b, const float* out_gk, const SpbWinStep& s) {
    float g = 0.f;
    for (int e = w.win_edge_off[win]; e < w.win_edge_off[win + 1]; ++e)
        if (w.edge_src[e] == f) g += w.edge_w[e] * out_gk[w.edge_seg_off[e] + b];
    const int i = w.frame_seg_off[f] + b;
    float* as = w.adam_seg + (size_t)i * SPB_ADAM_SEG;
    w.k[i] += win_adam_inc(g, as, as + 1, s, s.step_k);
}
