"""Camera matrices used on the tracking path (mirror of gaussian_splatting/utils/graphics_utils.py:33-101
of the reference; same names and conventions)."""
import math

import torch


def getWorld2View2(R, t, translate=torch.tensor([0.0, 0.0, 0.0]), scale=1.0):
    """4x4 world->camera matrix [R | t].  The reference inverts, recentres and inverts back
    (graphics_utils.py:33-46); with the default translate = 0, scale = 1 that is the identity, so the
    matrix is assembled directly (no linalg.inv, no host sync)."""
    Rt = torch.zeros((4, 4), device=R.device, dtype=R.dtype)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    if scale != 1.0 or bool(torch.any(translate != 0)):
        C2W = torch.linalg.inv(Rt)
        C2W[:3, 3] = (C2W[:3, 3] + translate.to(R.device)) * scale
        Rt = torch.linalg.inv(C2W)
    return Rt


def getProjectionMatrix(znear, zfar, fovX, fovY):
    tan_y, tan_x = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))
