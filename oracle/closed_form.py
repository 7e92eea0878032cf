"""ORACLE (test infrastructure -- never imported by the product path).

Independent float64 closed-form restatement (numpy) of the photometric alignment cost, its
analytic first derivatives and the IRLS Gauss-Newton normal equations (arrowhead blocks) --
SURVEY.md section 7.4.  Two jobs:

  1. cross-check ``oracle/ref_port.py`` (and through it the reference's autograd gradients);
  2. define the oracle for the GN / LM extension, which has **no reference counterpart**
     (the reference optimises with Adam + autograd only) -- GN-step outputs are
     therefore "parity unpinned" by the reference; they are pinned to this closed form.

Reference lines restated: core/dense_optim.py:19-35 (unproject), :38-86 (depth seeding, exp),
:117-122 (rigid transform), :128-162 (sampling + validity), :202-261 (affine + masked L1);
core/ops.py:19-40 (guarded projection); tool/point_utils.py:31-40 (normalisation).
"""
from __future__ import annotations

import numpy as np


def compact_geometry(regions, logd, keypoints):
    """(N,H,W) masks -> points in (segment,row,col) order (== torch.where order,
    core/dense_optim.py:103).  Returns dict of numpy arrays."""
    regions = np.asarray(regions, dtype=bool)
    logd = np.asarray(logd, dtype=np.float64)
    N, H, W = regions.shape
    if logd.ndim == 2:
        logd = np.broadcast_to(logd, (N, H, W))
    b, r, c = np.nonzero(regions)
    dims = np.array([H, W], dtype=np.float32)
    kp = np.asarray(keypoints, dtype=np.float32)
    rc = np.rint((np.float32(0.5) * (dims - 1)) * (kp + 1)).astype(np.int64)  # point_utils.py:37-40
    return dict(seg=b, u=c.astype(np.float64), v=r.astype(np.float64), L=logd[b, r, c],
                L_kp=logd[np.arange(N), rc[:, 0], rc[:, 1]], kp_rc=rc, N=N, H=H, W=W)


def bilinear_taps(img, ix, iy):
    """Zero-padded bilinear sample and its derivatives w.r.t. (ix, iy).
    img (C,Hl,Wl); ix, iy (P,) pixel coords.  Mirrors ATen grid_sampler_2d
    (align_corners=True, padding zeros)."""
    C, Hl, Wl = img.shape
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    fx = ix - x0
    fy = iy - y0

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < Wl) & (yy >= 0) & (yy < Hl)
        out = np.zeros((C, ix.shape[0]), dtype=np.float64)
        out[:, ok] = img[:, yy[ok], xx[ok]]
        return out

    nw, ne = tap(y0, x0), tap(y0, x0 + 1)
    sw, se = tap(y0 + 1, x0), tap(y0 + 1, x0 + 1)
    val = nw * (1 - fx) * (1 - fy) + ne * fx * (1 - fy) + sw * (1 - fx) * fy + se * fx * fy
    d_ix = (ne - nw) * (1 - fy) + (se - sw) * fy
    d_iy = (sw - nw) * (1 - fx) + (se - ne) * fx
    return val, d_ix, d_iy


def evaluate(geo, src_img, trg_img, K_src, K_trg, k, pose, affine=None, batch_thresholds=False,
             want_gn=False, irls_eps=1e-3, with_affine_cols=False):
    """Cost, analytic gradient and (optionally) IRLS normal equations for one
    (source keyframe, target) pair.

    geo      : compact_geometry() dict
    src_img  : (3,Hl,Wl) source level image; trg_img : (3,Hl,Wl) target level image
    K_src/K_trg : (3,3); k : (N,) log-depth seeds; pose : (4,4)
    affine   : None or (src_ab (2,), trg_ab (2,))
    Returns dict with cost, grads (pose 4x4, k, affine), per-point arrays, GN blocks.
    """
    f64 = np.float64
    src_img = np.asarray(src_img, f64)
    trg_img = np.asarray(trg_img, f64)
    K_src = np.asarray(K_src, f64)
    K_trg = np.asarray(K_trg, f64)
    k = np.asarray(k, f64)
    pose = np.asarray(pose, f64)
    H, W = geo["H"], geo["W"]
    Hl, Wl = trg_img.shape[1:]
    seg, u, v = geo["seg"], geo["u"], geo["v"]
    P = u.shape[0]
    tau = 1e-6 if batch_thresholds else 1e-7

    z = np.exp(geo["L"] + (k - geo["L_kp"])[seg])
    fx, fy, cx, cy = K_src[0, 0], K_src[1, 1], K_src[0, 2], K_src[1, 2]
    X = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)          # (P,3)
    R, t = pose[:3, :3], pose[:3, 3]
    RX = X @ R.T
    Y = RX + t

    inv_w = f64(np.float32(1.0) / np.float32(W - 1))                       # float32 reciprocal
    inv_h = f64(np.float32(1.0) / np.float32(H - 1))

    def project(Pts, Kc):
        zz = Pts[:, 2]
        zi = np.where(np.abs(zz) > 1e-6, 1.0 / np.where(zz == 0, 1.0, zz), 1e-6)
        uu = Pts[:, 0] * Kc[0, 0] * zi + Kc[0, 2]
        vv = Pts[:, 1] * Kc[1, 1] * zi + Kc[1, 2]
        xn = 2 * uu * inv_w - 1
        yn = 2 * vv * inv_h - 1
        ok = (np.abs(xn) <= 0.99) & (np.abs(yn) <= 0.99) & (zz > tau)
        ix = (xn + 1) * 0.5 * (Wl - 1)
        iy = (yn + 1) * 0.5 * (Hl - 1)
        return uu, vv, zi, ok, ix, iy

    # source self-sample (single-path threshold 1e-7 in both entry points)
    _, _, _, m_s, sx_, sy_ = project(X, K_src)
    m_s = m_s & (X[:, 2] > 1e-7)
    I_s, _, _ = bilinear_taps(src_img, sx_, sy_)

    up, vp, zi, m_t, ix, iy = project(Y, K_trg)
    I_t, dIx, dIy = bilinear_taps(trg_img, ix, iy)

    if affine is not None:
        a = f64(affine[1][0]) - f64(affine[0][0])
        bb = f64(affine[1][1]) - f64(affine[0][1])
    else:
        a, bb = 0.0, 0.0
    ea = np.exp(-a)
    I_tc = ea * I_t + bb
    m = (m_s & m_t).astype(f64)
    r = (I_s - I_tc) * m                                                 # (3,P)
    cost = np.abs(r).sum() / (3.0 * P)

    # ---- analytic gradient of the L1 mean (what autograd yields) --------------
    s = np.sign(r) * m / (3.0 * P)
    gI = -ea * s                                                         # d cost / d I_t
    kx = (Wl - 1) * inv_w                                                # d ix / d u'
    ky = (Hl - 1) * inv_h
    g_u = (gI * dIx).sum(0) * kx
    g_v = (gI * dIy).sum(0) * ky
    fxt, fyt = K_trg[0, 0], K_trg[1, 1]
    live = (np.abs(Y[:, 2]) > 1e-6).astype(f64)                          # guarded reciprocal has zero slope
    gY = np.stack([g_u * fxt * zi, g_v * fyt * zi,
                   -(g_u * fxt * Y[:, 0] + g_v * fyt * Y[:, 1]) * zi * zi * live], 1)
    g_pose = np.zeros((4, 4))
    g_pose[:3, :3] = gY.T @ X
    g_pose[:3, 3] = gY.sum(0)
    g_k = np.bincount(seg, weights=(gY * RX).sum(1), minlength=geo["N"])
    g_at = (s * ea * I_t).sum()
    g_bt = -s.sum()
    out = dict(cost=cost, residual_raw=r, mask=m, I_s=I_s, I_t=I_tc, X=X, Y=Y,
               g_pose=g_pose, g_k=g_k, g_aff_trg=np.array([g_at, g_bt]),
               g_aff_src=np.array([-g_at, -g_bt]), m_s=m_s, m_t=m_t, P=P)

    if want_gn:
        # Jacobian of r_c w.r.t. left tangent xi=(tau,phi) (T <- Exp(xi) T), the segment's
        # k_b, and optionally the target affine (a_t, b_t):  r = I_s - (e^{-a} I_t + b)
        # d r_c = -e^{-a} (dIx_c kx du' + dIy_c ky dv') ; du' = fxt zi dYx - fxt Yx zi^2 dYz
        P_ = P
        dudY = np.stack([fxt * zi, np.zeros(P_), -fxt * Y[:, 0] * zi * zi * live], 1)
        dvdY = np.stack([np.zeros(P_), fyt * zi, -fyt * Y[:, 1] * zi * zi * live], 1)
        # dY/dxi = [I | -[Y]x]
        Yx = np.zeros((P_, 3, 3))
        Yx[:, 0, 1], Yx[:, 0, 2] = -Y[:, 2], Y[:, 1]
        Yx[:, 1, 0], Yx[:, 1, 2] = Y[:, 2], -Y[:, 0]
        Yx[:, 2, 0], Yx[:, 2, 1] = -Y[:, 1], Y[:, 0]
        dYdxi = np.concatenate([np.broadcast_to(np.eye(3), (P_, 3, 3)), -Yx], 2)   # (P,3,6)
        Mu = np.einsum('pi,pij->pj', dudY, dYdxi)                                    # (P,6)
        Mv = np.einsum('pi,pij->pj', dvdY, dYdxi)
        du_dk = (dudY * RX).sum(1)
        dv_dk = (dvdY * RX).sum(1)
        npose = 8 if with_affine_cols else 6
        N = geo["N"]
        A = np.zeros((npose, npose))
        Bm = np.zeros((npose, N))
        D = np.zeros(N)
        gp = np.zeros(npose)
        gd = np.zeros(N)
        for c in range(3):
            wgt = m / np.maximum(np.abs(r[c]), irls_eps)                 # IRLS weights for L1
            au = -ea * dIx[c] * kx
            av = -ea * dIy[c] * ky
            Jp = au[:, None] * Mu + av[:, None] * Mv                     # (P,6)
            if with_affine_cols:
                Jp = np.concatenate([Jp, (ea * I_t[c])[:, None], -np.ones((P_, 1))], 1)
            Jp = Jp * m[:, None]
            jd = (au * du_dk + av * dv_dk) * m
            A += (Jp * wgt[:, None]).T @ Jp
            gp += (Jp * (wgt * r[c])[:, None]).sum(0)
            for j in range(npose):
                Bm[j] += np.bincount(seg, weights=wgt * Jp[:, j] * jd, minlength=N)
            D += np.bincount(seg, weights=wgt * jd * jd, minlength=N)
            gd += np.bincount(seg, weights=wgt * jd * r[c], minlength=N)
        out.update(A=A, B=Bm, D=D, g_p=gp, g_d=gd,
                   wcost=sum((m / np.maximum(np.abs(r[c]), irls_eps) * r[c] ** 2).sum()
                             for c in range(3)))
    return out


def lm_step(A, B, D, gp, gd, lam):
    """Damped arrowhead solve by Schur complement on the depth block:
        [A+lam diag(A)   B        ] [xi]   = -[gp]
        [B^T       D + lam D      ] [dk]     [gd]
    Returns (xi, dk).  Segments with D == 0 (no valid observation) get dk = 0."""
    Dd = D * (1.0 + lam)
    ok = Dd > 0
    inv = np.where(ok, 1.0 / np.where(ok, Dd, 1.0), 0.0)
    Ad = A + lam * np.diag(np.diag(A))
    S = Ad - (B * inv) @ B.T
    rhs = -(gp - B @ (inv * gd))
    xi = np.linalg.solve(S, rhs)
    dk = -inv * (gd + B.T @ xi)
    return xi, dk


def se3_exp(xi):
    """Exp of a twist xi = (tau, phi) -> 4x4 (translation first, lietorch ordering)."""
    xi = np.asarray(xi, np.float64)
    tau, phi = xi[:3], xi[3:6]
    th = np.linalg.norm(phi)
    Kx = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]])
    if th < 1e-8:
        R = np.eye(3) + Kx + 0.5 * Kx @ Kx
        V = np.eye(3) + 0.5 * Kx + Kx @ Kx / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Kx + (th - np.sin(th)) / th ** 3 * Kx @ Kx
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ tau
    return T
