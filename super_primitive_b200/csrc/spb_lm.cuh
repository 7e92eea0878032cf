// Device code of the damped arrowhead (Schur complement) solve; see spb_solve.cu for the description.
#pragma once
#include "spb_common.cuh"

#define SV_PAIR 80   // floats saved per problem : pose[16] aff[2] pad[14] gn_pair[48]
#define SV_SEG 12    // floats saved per segment : k, pad, gn_seg[10]

static __device__ __forceinline__ int tri8(int r, int c) { return r * 8 - r * (r - 1) / 2 + (c - r); }

static __device__ __forceinline__ void se3_exp_d(const double* xi, double* T /*3x4 row-major R|t*/) {
    const double tx = xi[0], ty = xi[1], tz = xi[2], px = xi[3], py = xi[4], pz = xi[5];
    const double th2 = px * px + py * py + pz * pz;
    double A, B, C;
    if (th2 < 1e-16) {
        A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        const double rth = rsqrt(th2), th = th2 * rth, r2 = rth * rth;
        double sn, cs;
        sincos(th, &sn, &cs);
        A = sn * rth; B = (1.0 - cs) * r2; C = (th - sn) * (r2 * rth);
    }
    const double K[9] = {0, -pz, py, pz, 0, -px, -py, px, 0};
    double K2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    const double tau[3] = {tx, ty, tz};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double vt = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double I = (i == j) ? 1.0 : 0.0;
            T[4 * i + j] = I + A * K[3 * i + j] + B * K2[3 * i + j];
            vt += (I + B * K[3 * i + j] + C * K2[3 * i + j]) * tau[j];
        }
        T[4 * i + 3] = vt;
    }
}

// body shared by the standalone kernel (k_lm_update) and the fused finalize+solve kernel; any block size that
// is a multiple of 32 up to SPB_LM_MAXWARPS warps
#define SPB_LM_MAXWARPS 32
#define SPB_LM_CHUNK 512                         // segments whose 1/D is staged in shared memory per pass
// gn_pair / gn_seg are deliberately NOT __restrict__: the fused kernel writes them earlier in the same launch, so
// they must be read with ordinary (coherent) loads, not through the read-only path.
__device__ __forceinline__ void lm_update_body(const float* gn_pair, const float* gn_seg,
                                               const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_cnt,
                                               int with_affine, int hold_depth, float* __restrict__ poses,
                                               float* __restrict__ k,
                                               float* __restrict__ aff_trg, float* __restrict__ lm_state,
                                               float* __restrict__ saved_pair, float* __restrict__ saved_seg) {
    const int p = blockIdx.x;
    const int so = seg_off[p], n = seg_cnt[p];
    float* st = lm_state + (size_t)p * SPB_LM_NSTATE;
    float* sp = saved_pair + (size_t)p * SV_PAIR;
    float* ss = saved_seg + (size_t)so * SV_SEG;
    const float* gp = gn_pair + (size_t)p * SPB_GN_PAIR_NOUT;
    const float* gs = gn_seg + (size_t)so * SPB_GN_SEG_NOUT;
    float* pose = poses + (size_t)p * 16;
    __shared__ int s_accept;
    __shared__ float s_lam;
    __shared__ double s_red[SPB_LM_MAXWARPS / 2][44];          // one row per group of 64 threads
    __shared__ double s_inv[SPB_LM_CHUNK];
    __shared__ double s_sys[44];
    __shared__ double s_xi[8];

    if (threadIdx.x == 0) {
        const float cost = gp[SPB_GN_NA + 8];
        const bool init = st[2] != 0.f;
        const bool acc = !init || (cost < st[1]);
        float lam = st[0];
        if (acc) {
            st[1] = cost;
            if (init) { lam = fmaxf(lam * 0.25f, 1e-7f); st[3] += 1.f; }
            st[2] = 1.f;
        } else {
            lam = fminf(lam * 8.0f, 1e7f);
            st[4] += 1.f;
        }
        st[0] = lam;
        st[5] = cost;
        s_accept = acc ? 1 : 0;
        s_lam = lam;
    }
    __syncthreads();
    const bool acc = s_accept != 0;
    const double lam = (double)s_lam;
    // accept: snapshot the current parameters and system; reject: roll the parameters back
    if (acc) {
        for (int i = threadIdx.x; i < 16; i += blockDim.x) sp[i] = pose[i];
        if (threadIdx.x < 2) sp[16 + threadIdx.x] = (with_affine && aff_trg) ? aff_trg[2 * p + threadIdx.x] : 0.f;
        for (int i = threadIdx.x; i < SPB_GN_PAIR_NOUT; i += blockDim.x) sp[32 + i] = gp[i];
        for (int b = threadIdx.x; b < n; b += blockDim.x) {
            ss[(size_t)b * SV_SEG] = k[so + b];
            for (int i = 0; i < SPB_GN_SEG_NOUT; ++i) ss[(size_t)b * SV_SEG + 2 + i] = gs[(size_t)b * SPB_GN_SEG_NOUT + i];
        }
    }
    __syncthreads();
    const float* A = sp + 32;          // saved system (== current one after an accept)
    // Schur complement sum over the segments (float64), organised for latency: the CTA is cut into groups of 64
    // threads; thread q < 44 of a group owns ONE entry of  sum_b B_b D_b^-1 [B_b^T | g_d,b]  and walks the
    // group's share of the segments (the per-segment record is a broadcast read), so there is no cross-lane
    // reduction at all; the groups' partial sums meet in shared memory.
    const int ngroups = blockDim.x >> 6, grp = threadIdx.x >> 6, q = threadIdx.x & 63;
    {
        int qi = 0, qj = 0;                                     // q < 36: (i, j) of the triangle; q >= 36: rhs row
        if (q < 36) {
            int rem = q;
            while (rem >= 8 - qi) { rem -= 8 - qi; ++qi; }
            qj = qi + rem;
        } else {
            qi = q - 36; qj = 9;
        }
        double accq = 0.0;
        for (int base = 0; base < n; base += SPB_LM_CHUNK) {
            const int m = min(SPB_LM_CHUNK, n - base);
            for (int b = threadIdx.x; b < m; b += blockDim.x) {
                const double D = (double)ss[(size_t)(base + b) * SV_SEG + 2 + 8] * (1.0 + lam);
                s_inv[b] = (D > 1e-30 && !hold_depth) ? 1.0 / D : 0.0;   // held seeds: no elimination, S = damped A
            }
            __syncthreads();
            if (q < 44 && grp < ngroups) {
                for (int b = grp; b < m; b += ngroups) {
                    const float* sb = ss + (size_t)(base + b) * SV_SEG + 2;
                    accq = fma((double)sb[qi] * s_inv[b], (double)sb[qj], accq);
                }
            }
            __syncthreads();
        }
        if (q < 44 && grp < ngroups) s_red[grp][q] = accq;
        __syncthreads();
        // entry q of the damped Schur system, still one thread per entry:  S = A + lam diag(A) - sum,  rhs = -(g_p - sum)
        if (threadIdx.x < 44) {
            double sub = 0.0;
            for (int w = 0; w < ngroups; ++w) sub += s_red[w][q];
            if (q < 36) {
                const double a = (double)A[q];                  // q enumerates the triangle exactly like tri8(qi, qj)
                s_sys[q] = (qi == qj) ? fma(lam, a, a) - sub : a - sub;
            } else {
                s_sys[q] = -((double)A[SPB_GN_NA + qi] - sub);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // everything below is unrolled with compile-time indices: S, rhs, y, x stay in registers (no local memory)
        double S[8][8], rhs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = i; j < 8; ++j) {
                const double v = s_sys[tri8(i, j)];
                S[i][j] = v; S[j][i] = v;
            }
            rhs[i] = s_sys[36 + i];
        }
        const int np = with_affine ? 8 : 6;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool active = (i < np) && (S[i][i] > 1e-30) && isfinite(S[i][i]);
            if (!active) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { S[i][j] = 0.0; S[j][i] = 0.0; }
                S[i][i] = 1.0; rhs[i] = 0.0;
            }
        }
        // Cholesky S = L L^T (in place, lower); rl[j] = 1 / L[j][j] (one rsqrt per column, no divisions)
        bool ok = true;
        double rl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double d = S[j][j];
#pragma unroll
            for (int qq = 0; qq < j; ++qq) d -= S[j][qq] * S[j][qq];
            if (!(d > 0.0)) { ok = false; d = 1.0; }
            rl[j] = rsqrt(d);
            S[j][j] = d * rl[j];
#pragma unroll
            for (int i = j + 1; i < 8; ++i) {
                double v = S[i][j];
#pragma unroll
                for (int qq = 0; qq < j; ++qq) v -= S[i][qq] * S[j][qq];
                S[i][j] = v * rl[j];
            }
        }
        double x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = rhs[i];
#pragma unroll
            for (int qq = 0; qq < i; ++qq) v -= S[i][qq] * y[qq];
            y[i] = v * rl[i];
        }
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            double v = y[i];
#pragma unroll
            for (int qq = i + 1; qq < 8; ++qq) v -= S[qq][i] * x[qq];
            x[i] = v * rl[i];
        }
        if (!ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.0;
            st[0] = fminf(st[0] * 8.0f, 1e7f);   // not positive definite: damp harder next time
        }
        double nrm = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_xi[i] = x[i]; nrm += x[i] * x[i]; }
        st[6] = (float)sqrt(nrm);
        // T <- Exp(xi) T_saved
        double E[12];
        se3_exp_d(x, E);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double v = E[4 * i] * (double)sp[j] + E[4 * i + 1] * (double)sp[4 + j] + E[4 * i + 2] * (double)sp[8 + j];
                if (j == 3) v += E[4 * i + 3];
                pose[4 * i + j] = (float)v;
            }
        }
        pose[12] = 0.f; pose[13] = 0.f; pose[14] = 0.f; pose[15] = 1.f;
        if (with_affine && aff_trg) {
            aff_trg[2 * p] = sp[16] + (float)x[6];
            aff_trg[2 * p + 1] = sp[17] + (float)x[7];
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        const float* sb = ss + (size_t)b * SV_SEG + 2;
        const double D = (double)sb[8] * (1.0 + lam);
        double dk = 0.0;
        if (D > 1e-30 && !hold_depth) {
            double v = sb[9];
#pragma unroll
            for (int i = 0; i < 8; ++i) v += (double)sb[i] * s_xi[i];
            dk = -v / D;
        }
        k[so + b] = ss[(size_t)b * SV_SEG] + (float)dk;
    }
}

