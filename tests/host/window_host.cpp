// TEST INFRASTRUCTURE ONLY.  Host build (g++) of the per-item arithmetic of the mapping-window update
// (super_primitive_b200/csrc/spb_window_math.h) driven by plain loops in the phase order of k_window_update
// (spb_window.cu), so the chain rule, the Adam arithmetic and the pose bookkeeping are checked against
// oracle/window_loop.py without a GPU.  All pointers inside SpbWindow are HOST memory here.  Never loaded by the
// product path.
#include "../../super_primitive_b200/csrc/spb_window_math.h"

extern "C" int window_poses_host(const SpbWindow* w) {
    for (int e = 0; e < w->n_edges; ++e)
        win_edge_pose(w->frame_T + (size_t)w->edge_trg[e] * 16, w->frame_T + (size_t)w->edge_src[e] * 16,
                      w->edge_pose + (size_t)e * 16);
    return 0;
}

extern "C" int window_update_host(const SpbWindow* wp, const float* out_pair, const float* out_gk, double lr_pose,
                                  double lr_k, double lr_aff, double beta1, double beta2, double eps, double stop_tol) {
    const SpbWindow& w = *wp;
    const SpbWinHyper h{lr_pose, lr_k, lr_aff, beta1, beta2, eps, stop_tol};
    for (int win = 0; win < w.n_windows; ++win) {
        float* st = w.win_state + (size_t)win * SPB_WIN_NSTATE;
        if (st[3] != 0.f) continue;
        const int f0 = w.win_frame_off[win], f1 = w.win_frame_off[win + 1];
        const int e0 = w.win_edge_off[win], e1 = w.win_edge_off[win + 1];
        const SpbWinStep s = win_step_sizes(h, st[0] + 1.0f);
        for (int e = e0; e < e1; ++e)
            win_edge_twists(out_pair + (size_t)e * SPB_PAIR_NOUT, w.edge_pose + (size_t)e * 16, w.edge_tw + (size_t)e * 12);
        for (int f = f0; f < f1; ++f) win_frame_step(w, win, f, out_pair, s);
        for (int f = f0; f < f1; ++f) {
            if (!(w.frame_flags[f] & SPB_WIN_OPT_SEEDS)) continue;
            for (int b = 0; b < w.frame_seg_cnt[f]; ++b) win_seed_step(w, win, f, b, out_gk, s);
        }
        for (int e = e0; e < e1; ++e)
            win_edge_pose(w.frame_T + (size_t)w.edge_trg[e] * 16, w.frame_T + (size_t)w.edge_src[e] * 16,
                          w.edge_pose + (size_t)e * 16);
        const float loss = win_loss(w, win, out_pair);
        const float prev = st[1];
        st[0] += 1.0f;
        st[2] = prev;
        st[1] = loss;
        if (h.stop_tol > 0.0 && st[0] > 1.0f && fabsf(loss - prev) / prev < (float)h.stop_tol) st[3] = 1.0f;
    }
    return 0;
}

extern "C" void renormalise_host(float* T) { win_renormalise(T); }
extern "C" void se3_exp_host(const double* xi, double* T12) { win_se3_exp(xi, T12); }
