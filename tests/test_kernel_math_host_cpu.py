"""CPU: the per-point arithmetic of the fused alignment kernel, compiled by g++ from the kernel's OWN headers.

`project_point`, `point_grad_packed`, `point_gn6_packed` and `fill_fast_ctx` (super_primitive_b200/csrc/spb_fast.cuh,
spb_gn_packed.cuh) are built for the host with the CUDA built-ins replaced by tests/host/cuda_shim.h and driven point by
point (tests/host/align_host.cpp, test infrastructure); every point's contribution is accumulated in float64 and
compared with the float64 closed form (oracle/closed_form.py) on the golden cases of the live reference.  This checks
the folded-context projection, validity tests, bilinear slopes, the sign-bit gradient trick, the exact depth-column
identity and the packed normal-equation update without a GPU; the GPU tests then check the same quantities through the
real kernels (tile pipeline, reductions, finalize)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import closed_form as cf
from tests.common import Golden, assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "super_primitive_b200", "csrc")


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host", "align_host.cpp")
    out_dir = os.path.join(HERE, "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libalign_host.so")
    deps = [src, os.path.join(HERE, "host", "cuda_shim.h")] + [os.path.join(CSRC, h) for h in
                                                                ("spb_fast.cuh", "spb_gn_packed.cuh", "spb_common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, "-D__device__=",
                               "-D__forceinline__=inline", "-D__global__=", "-D__restrict__=", "-DSPB_TAP_L2_256=0",
                               "-o", out, src])
    lib = C.CDLL(out)
    lib.align_points_host.argtypes = ([C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                                               C.c_int] + [C.c_void_p] * 5 +
                                      [C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p])
    return lib


def _run(host, mode, g, lvl, j, with_affine):
    """Returns (closed-form dict, per-pair sums, per-segment sums) for target j of golden case g at pyramid level lvl."""
    z = g.z
    geo = cf.compact_geometry(z["src_regions"], z["src_logdepth"], z["src_keypoints"])
    aff = (z["aff_src"], z["aff_trg"][j]) if with_affine else None
    src_img, trg_img = z[f"L{lvl}_src_image"], z[f"L{lvl}_trg_images"][j]
    r = cf.evaluate(geo, src_img, trg_img, z["src_K"], z["src_K"], z["k"], z["poses"][j], aff, want_gn=True,
                    irls_eps=1e-3)
    P, N = r["P"], geo["N"]
    uv = (geo["u"].astype(np.uint32) | (geo["v"].astype(np.uint32) << 16) | (r["m_s"].astype(np.uint32) << 31))
    uv = np.ascontiguousarray(uv, dtype=np.uint32)
    logd = np.ascontiguousarray(geo["L"], dtype=np.float32)
    Is = np.ascontiguousarray(r["I_s"], dtype=np.float32)                       # (3,P) cached source samples
    seg = np.ascontiguousarray(geo["seg"], dtype=np.int32)
    shift = (z["k"].astype(np.float32) - geo["L_kp"].astype(np.float32)).astype(np.float32)
    Hl, Wl = trg_img.shape[1:]
    rgba = np.zeros((Hl, Wl, 4), np.float32)
    rgba[..., :3] = np.transpose(trg_img, (1, 2, 0))
    K = np.ascontiguousarray(z["src_K"], dtype=np.float32).reshape(9)
    pose = np.ascontiguousarray(z["poses"][j], dtype=np.float32).reshape(16)
    a_s = np.ascontiguousarray(z["aff_src"], dtype=np.float32) if with_affine else None
    a_t = np.ascontiguousarray(z["aff_trg"][j], dtype=np.float32) if with_affine else None
    out_pair = np.zeros(28 if mode == 1 else 16, np.float64)
    out_seg = np.zeros((N, 8) if mode == 1 else (N,), np.float64)
    p = lambda a: None if a is None else a.ctypes.data                           # noqa: E731
    rc = host.align_points_host(mode, P, p(uv), p(logd), p(Is), p(seg), N, p(shift), p(rgba), Hl, Wl, p(K), p(K), p(pose),
                                p(a_s), p(a_t), 1e-7, geo["H"], geo["W"], 1e-3, p(out_pair), p(out_seg))
    assert rc == 0
    return r, out_pair, out_seg


CASES = [("tiny_rects", False), ("tiny_rects", True), ("tiny_strips", False), ("pyr3_rects", True)]


@pytest.mark.parametrize("case,with_affine", CASES)
def test_gradient_arithmetic_matches_closed_form(host, case, with_affine):
    g = Golden(case)
    with_affine = with_affine and g.with_affine
    for lvl in range(g.n_levels):
        for j in range(g.B):
            r, op, gk = _run(host, 0, g, lvl, j, with_affine)
            norm = 1.0 / (3.0 * r["P"])
            fx, fy = float(g.z["src_K"][0, 0]), float(g.z["src_K"][1, 1])
            what = f"{case} level {lvl} target {j}"
            assert_close(op[0] * norm, r["cost"], 2e-5, "cost " + what)
            assert_close(op[1:4] * norm, r["g_pose"][:3, 3], 1e-4, "d/dt " + what)
            gR = op[4:13].reshape(3, 3) * np.array([1.0 / fx, 1.0 / fy, 1.0]) * norm     # d/dM -> d/dR (finalize kernel)
            assert_close(gR, r["g_pose"][:3, :3], 1e-4, "d/dR " + what)
            assert_close(gk * norm, r["g_k"], 1e-4, "d/dk " + what)
            if with_affine:
                assert_close(op[13:15] * norm, r["g_aff_trg"], 1e-4, "d/d(a,b) " + what)
            assert op[15] == (r["m_s"] & r["m_t"]).sum()                                  # valid in both views


@pytest.mark.parametrize("case,with_affine", CASES)
def test_normal_equation_arithmetic_matches_closed_form(host, case, with_affine):
    g = Golden(case)
    with_affine = with_affine and g.with_affine
    for lvl in range(g.n_levels):
        for j in range(g.B):
            r, op, sg = _run(host, 1, g, lvl, j, with_affine)
            A = np.zeros((6, 6))
            q = 0
            for a in range(6):
                for b in range(a, 6):
                    A[a, b] = A[b, a] = op[q]
                    q += 1
            what = f"{case} level {lvl} target {j}"
            assert_close(A, r["A"], 1e-3, "A " + what)
            assert_close(op[21:27], r["g_p"], 1e-3, "g_p " + what)
            assert_close(op[27] / (3.0 * r["P"]), r["cost"], 2e-5, "cost " + what)
            assert_close(sg[:, :6].T, r["B"], 1e-3, "B " + what)
            assert_close(sg[:, 6], r["D"], 1e-3, "D " + what)
            assert_close(sg[:, 7], r["g_d"], 1e-3, "g_d " + what)
            # the depth column is a fixed combination of the first three twist columns:
            # sum_b B_b = -(t_x A[:,0] + t_y A[:,1] + t_z A[:,2])
            t = g.z["poses"][j][:3, 3].astype(np.float64)
            assert_close(sg[:, :6].sum(0), -(A[:, :3] @ t), 2e-3, "depth-column identity " + what)
