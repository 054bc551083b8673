"""SE(3)/SO(3) helpers with the reference's conventions (utils/pose.py:6-91): twists are
[rho(3); theta(3)], small-angle cut at 1e-5.  Branch-free in torch (torch.where) so no host sync is
needed to evaluate `angle < 1e-5`."""
import torch


def rt2mat(R, T):
    mat = torch.eye(4, device=R.device, dtype=R.dtype)
    mat[0:3, 0:3] = R
    mat[0:3, 3] = T
    return mat


def skew_sym_mat(x):
    z = torch.zeros((), device=x.device, dtype=x.dtype)
    return torch.stack([torch.stack([z, -x[2], x[1]]), torch.stack([x[2], z, -x[0]]), torch.stack([-x[1], x[0], z])])


def _coeffs(angle):
    small = angle < 1e-5
    safe = torch.where(small, torch.ones_like(angle), angle)
    a1 = torch.where(small, torch.ones_like(angle), torch.sin(safe) / safe)
    a2 = torch.where(small, torch.full_like(angle, 0.5), (1 - torch.cos(safe)) / (safe ** 2))
    b2 = torch.where(small, torch.full_like(angle, 1.0 / 6.0), (safe - torch.sin(safe)) / (safe ** 3))
    return a1, a2, b2


def SO3_exp(theta):
    W = skew_sym_mat(theta)
    a1, a2, _ = _coeffs(torch.norm(theta))
    return torch.eye(3, device=theta.device, dtype=theta.dtype) + a1 * W + a2 * (W @ W)


def V(theta):
    W = skew_sym_mat(theta)
    _, a2, b2 = _coeffs(torch.norm(theta))
    return torch.eye(3, device=theta.device, dtype=theta.dtype) + W * a2 + (W @ W) * b2


def SO3_log(R):
    c = torch.clamp((torch.trace(R) - 1) / 2.0, -1.0, 1.0)
    theta = torch.acos(c)
    if theta.abs() < 1e-5:
        return torch.zeros(3, device=R.device)
    axis = torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * torch.sin(theta))
    return theta * axis


def SE3_exp(deltaT):
    rho, theta = deltaT[:3], deltaT[3:]
    T = torch.eye(4, device=deltaT.device, dtype=deltaT.dtype)
    T[:3, :3] = SO3_exp(theta)
    T[:3, 3] = V(theta) @ rho
    return T
