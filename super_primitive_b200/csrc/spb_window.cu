// Mapping-window update (coupled problems): per-edge gradients -> per-frame torch.optim.Adam step, pose folding,
// re-normalisation and the relative poses of the next iteration, one CTA per window, nothing leaves the device.
// Replaces the host glue of the reference's mapping loop (odometery/odometery.py:762-915): pose_to_mat / torch.linalg.inv
// products per target per iteration, optim.step(), the folding loop, renormalise_se3 and the loss.item() early-stop test.
// The arithmetic of one edge / frame / seed lives in spb_window_math.h (shared with the host harness of the CPU tests);
// this file is the thread mapping.  Latency-bound by construction (a window has <= ~10 frames and ~40 edges); the
// bandwidth work of the iteration is the fused gradient kernel that runs before it (spb_grad_accumulate).
#include "spb_common.cuh"
#include "spb_window_math.h"

#define SPB_WIN_THREADS 256

// phases: (1) edge twists -> edge_tw ; (2) seeds (all threads) and frames (one thread each) ; (3) edge poses
// out_pair / out_gk are read with ordinary loads (written by the previous launch on the same stream).
__global__ void __launch_bounds__(SPB_WIN_THREADS)
k_window_update(const __grid_constant__ SpbWindow w, const float* out_pair, const float* out_gk, const SpbWinHyper h) {
    const int win = blockIdx.x;
    float* st = w.win_state + (size_t)win * SPB_WIN_NSTATE;
    if (st[3] != 0.f) return;                                 // converged earlier: the window is frozen (uniform)
    const int f0 = w.win_frame_off[win], f1 = w.win_frame_off[win + 1];
    const int e0 = w.win_edge_off[win], e1 = w.win_edge_off[win + 1];
    __shared__ SpbWinStep s_step;
    if (threadIdx.x == 0) s_step = win_step_sizes(h, st[0] + 1.0f);
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x)
        win_edge_twists(out_pair + (size_t)e * SPB_PAIR_NOUT, w.edge_pose + (size_t)e * 16, w.edge_tw + (size_t)e * 12);
    __syncthreads();
    const SpbWinStep s = s_step;
    // frames: the last warp's lanes (so the seed loop of the other warps runs beside the serial pose arithmetic)
    const int ft = (int)threadIdx.x - (SPB_WIN_THREADS - 32);
    if (ft >= 0)
        for (int f = f0 + ft; f < f1; f += 32) win_frame_step(w, win, f, out_pair, s);
    // seeds of the keyframes whose depth is optimised
    for (int f = f0; f < f1; ++f) {
        if (!(w.frame_flags[f] & SPB_WIN_OPT_SEEDS)) continue;
        const int n = w.frame_seg_cnt[f];
        for (int b = threadIdx.x; b < n; b += blockDim.x) win_seed_step(w, win, f, b, out_gk, s);
    }
    __syncthreads();
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x)
        win_edge_pose(w.frame_T + (size_t)w.edge_trg[e] * 16, w.frame_T + (size_t)w.edge_src[e] * 16,
                      w.edge_pose + (size_t)e * 16);
    if (threadIdx.x == 0) {
        const float loss = win_loss(w, win, out_pair);
        const float prev = st[1];
        st[0] += 1.0f;
        st[2] = prev;
        st[1] = loss;
        // early stop of the reference (:907-915): the step of the converging iteration is applied, then the loop ends
        if (h.stop_tol > 0.0 && st[0] > 1.0f && fabsf(loss - prev) / prev < (float)h.stop_tol) st[3] = 1.0f;
    }
}

__global__ void k_window_poses(const __grid_constant__ SpbWindow w) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < w.n_edges; e += gridDim.x * blockDim.x)
        win_edge_pose(w.frame_T + (size_t)w.edge_trg[e] * 16, w.frame_T + (size_t)w.edge_src[e] * 16,
                      w.edge_pose + (size_t)e * 16);
}

static bool window_ok(const SpbWindow* w) {
    return w && w->n_windows >= 1 && w->n_frames >= 1 && w->n_edges >= 1 && w->win_frame_off && w->win_edge_off &&
           w->edge_src && w->edge_trg && w->edge_w && w->edge_seg_off && w->frame_seg_off && w->frame_seg_cnt &&
           w->frame_flags && w->frame_T && w->k && w->edge_pose && w->adam_frame && w->adam_seg && w->win_state && w->edge_tw;
}

extern "C" int spb_window_poses(const SpbWindow* win, void* stream) {
    if (!window_ok(win)) return SPB_EINVAL;
    k_window_poses<<<(win->n_edges + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*win);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_window_update(const SpbWindow* win, const float* out_pair, const float* out_gk, double lr_pose,
                                 double lr_k, double lr_aff, double beta1, double beta2, double eps, double stop_tol,
                                 void* stream) {
    if (!window_ok(win) || !out_pair || !out_gk) return SPB_EINVAL;
    if (!(lr_pose >= 0.0 && lr_k >= 0.0 && lr_aff >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 &&
          eps >= 0.0))
        return SPB_EINVAL;
    if (win->n_windows > 65535) return SPB_ELIMIT;
    const SpbWinHyper h{lr_pose, lr_k, lr_aff, beta1, beta2, eps, stop_tol};
    k_window_update<<<win->n_windows, SPB_WIN_THREADS, 0, (cudaStream_t)stream>>>(*win, out_pair, out_gk, h);
    SPB_CHECK_LAUNCH();
    return SPB_OK;
}

extern "C" int spb_window_iterate(const SpbGeom* geoms, const SpbPair* pairs, const SpbWindow* win, int max_tiles,
                                  int with_affine, float* work, int64_t work_stride, float* out_pair, float* out_gk,
                                  double lr_pose, double lr_k, double lr_aff, double beta1, double beta2, double eps,
                                  double stop_tol, void* ev_before, void* ev_after, void* stream) {
    if (!window_ok(win)) return SPB_EINVAL;
    const int rc = spb_grad_accumulate(geoms, pairs, win->edge_seg_off, win->n_edges, max_tiles, with_affine, work,
                                       work_stride, out_pair, out_gk, ev_before, ev_after, stream);
    if (rc != SPB_OK) return rc;
    return spb_window_update(win, out_pair, out_gk, lr_pose, lr_k, lr_aff, beta1, beta2, eps, stop_tol, stream);
}
