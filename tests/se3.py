"""Test-side differentiable SE(3) exponential (stand-in for the absent, unpinned lietorch retraction;
used identically when the golden trajectory was generated and when it is replayed on the GPU)."""
import torch


def se3_exp_t(xi):
    """Differentiable torch SE3 exponential, xi = (tau, phi).  Test-side stand-in for the
    (absent, unpinned) lietorch retraction; used identically on both arms."""
    tau, phi = xi[:3], xi[3:]
    th2 = (phi * phi).sum()
    th = torch.sqrt(th2 + 1e-24)
    zero = torch.zeros((), dtype=xi.dtype, device=xi.device)
    Kx = torch.stack([torch.stack([zero, -phi[2], phi[1]]),
                      torch.stack([phi[2], zero, -phi[0]]),
                      torch.stack([-phi[1], phi[0], zero])])
    A = torch.sin(th) / th
    Bc = (1 - torch.cos(th)) / (th2 + 1e-24)
    Cc = (th - torch.sin(th)) / (th2 * th + 1e-36)
    eye = torch.eye(3, dtype=xi.dtype, device=xi.device)
    small = bool(th2.detach() < 1e-12)
    if small:
        R = eye + Kx + 0.5 * Kx @ Kx
        V = eye + 0.5 * Kx + Kx @ Kx / 6.0
    else:
        R = eye + A * Kx + Bc * Kx @ Kx
        V = eye + Bc * Kx + Cc * Kx @ Kx
    T = torch.eye(4, dtype=xi.dtype, device=xi.device)
    T = T.clone()
    top = torch.cat([R, (V @ tau)[:, None]], 1)
    return torch.cat([top, T[3:4]], 0)


