// Damped arrowhead (Schur complement) solve + SE(3) retraction + LM bookkeeping, one CTA per
// problem, everything device-resident so a whole GN/LM loop runs without host synchronisation
// (CUDA-graph friendly).  No reference counterpart: the reference optimises with torch.optim.Adam
// + autograd (odometery/two_frame_sfm.py:117-124,201-206); this is the GN/LM extension named by
// BASELINE.json, with oracle/closed_form.py as its oracle.
//
// System per problem (8 pose columns = 6 twist + 2 target-affine, N depth columns):
//     [A   B ] [xi]      [g_p]
//     [B^T D ] [dk] = -  [g_d]        D diagonal
// LM: A <- A + lam diag(A), D <- D (1 + lam); Schur: S = A - B D^-1 B^T, solved in float64.
#include "spb_adam.cuh"

__global__ void __launch_bounds__(128)
k_lm_update(const float* __restrict__ gn_pair, const float* __restrict__ gn_seg, const int32_t* __restrict__ seg_off,
            const int32_t* __restrict__ seg_cnt, int with_affine, int hold_depth, float* __restrict__ poses,
            float* __restrict__ k, float* __restrict__ aff_trg, float* __restrict__ lm_state,
            float* __restrict__ saved_pair, float* __restrict__ saved_seg) {
    lm_update_body(gn_pair, gn_seg, seg_off, seg_cnt, with_affine, hold_depth, poses, k, aff_trg, lm_state, saved_pair,
                   saved_seg);
}

extern "C" int spb_lm_saved_floats(int n_pairs, int seg_total, int64_t* pair_floats, int64_t* seg_floats) {
    if (!pair_floats || !seg_floats) return SPB_EINVAL;
    *pair_floats = (int64_t)n_pairs * SV_PAIR;
    *seg_floats = (int64_t)seg_total * SV_SEG;
    return SPB_OK;
}

extern "C" int spb_lm_update(const float* gn_pair, const float* gn_seg, const int32_t* seg_off,
                             const int32_t* seg_cnt, int n_pairs, int with_affine, int hold_depth, float* poses,
                             float* k, float* aff_trg, float* lm_state, float* saved_pair, float* saved_seg,
                             void* stream) {
    if (!gn_pair || !gn_seg || !seg_off || !seg_cnt || n_pairs < 1 || !poses || !k || !lm_state || !saved_pair ||
        !saved_seg)
        return SPB_EINVAL;
    k_lm_update<<<n_pairs, 128, 0, (cudaStream_t)stream>>>(gn_pair, gn_seg, seg_off, seg_cnt, with_affine,
                                                          hold_depth ? 1 : 0, poses, k, aff_trg, lm_state, saved_pair,
                                                          saved_seg);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

// ---- first-order counterpart: torch.optim.Adam + tracker bookkeeping on the device (spb_adam.cuh) ----------------
__global__ void __launch_bounds__(128)
k_adam_update(const float* __restrict__ out_pair, const float* __restrict__ out_gk, const int32_t* __restrict__ seg_off,
              const int32_t* __restrict__ seg_cnt, int with_affine, const SpbAdamHyper h, float* __restrict__ poses,
              float* __restrict__ k, float* __restrict__ aff_trg, float* __restrict__ adam_pair,
              float* __restrict__ adam_seg) {
    adam_update_body(out_pair, out_gk, seg_off, seg_cnt, with_affine, h, poses, k, aff_trg, adam_pair, adam_seg);
}

extern "C" int spb_adam_update(const float* out_pair, const float* out_gk, const int32_t* seg_off, const int32_t* seg_cnt,
                               int n_pairs, int with_affine, float* poses, float* k, float* aff_trg, float* adam_pair,
                               float* adam_seg, double lr_pose, double lr_k, double lr_aff, double beta1, double beta2,
                               double eps, void* stream) {
    if (!out_pair || !out_gk || !seg_off || !seg_cnt || n_pairs < 1 || !poses || !k || !adam_pair || !adam_seg)
        return SPB_EINVAL;
    if (with_affine == 1 && !aff_trg) return SPB_EINVAL;
    if (!(lr_pose >= 0.0 && lr_k >= 0.0 && lr_aff >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 &&
          eps >= 0.0))
        return SPB_EINVAL;
    const SpbAdamHyper h{lr_pose, lr_k, lr_aff, beta1, beta2, eps};
    k_adam_update<<<n_pairs, 128, 0, (cudaStream_t)stream>>>(out_pair, out_gk, seg_off, seg_cnt, with_affine == 1 ? 1 : 0, h,
                                                            poses, k, aff_trg, adam_pair, adam_seg);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}
