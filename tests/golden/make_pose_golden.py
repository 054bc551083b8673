#!/usr/bin/env python3
"""Generates tests/golden/pose_algebra.npz by importing the REFERENCE's own Python (utils/pose.py,
utils/render_camera/camera.py, gaussian_splatting/utils/graphics_utils.py) from /root/reference on the CPU.
Run in the authoring container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_pose_golden.py
Stored: seeded twists (including the small-angle branch), SE3_exp / SO3_log results, Camera state after
const_vel_model / update_pose / update_vwRT / cal_weighted_velocity, the half-interval velocity
transforms, the projection matrix and the world-view / full-projection matrices of a camera.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GSEVT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))   # munch
sys.path.insert(0, REF)

import torch  # noqa: E402
from utils.pose import SE3_exp, SO3_log  # noqa: E402
from utils.render_camera.camera import Camera  # noqa: E402
from gaussian_splatting.utils.graphics_utils import getProjectionMatrix, focal2fov  # noqa: E402

torch.manual_seed(0)
rng = np.random.default_rng(42)
twists = np.concatenate([rng.normal(0, 0.3, (12, 6)), rng.normal(0, 1e-3, (4, 6)),
                         np.concatenate([rng.normal(0, 0.1, (3, 3)), rng.normal(0, 2e-6, (3, 3))], axis=1),   # small-angle branch
                         np.zeros((1, 6))]).astype(np.float32)
exp = np.stack([SE3_exp(torch.from_numpy(t)).numpy() for t in twists])
logs = np.stack([SO3_log(torch.from_numpy(e[:3, :3].copy())).numpy() for e in exp])

R0 = exp[0][:3, :3].copy()
T0 = np.array([-3.99, 2.07, -1.36], np.float32)
w0 = np.array([0.025, 0.009, 0.010], np.float32)
v0 = np.array([0.0134, 0.1050, 0.0996], np.float32)
W, H, fx, fy = 640, 480, 327.32749, 327.46184


def cam():
    c = Camera(torch.from_numpy(R0.copy()), torch.from_numpy(T0.copy()), torch.from_numpy(w0.copy()), torch.from_numpy(v0.copy()),
               focal2fov(fx, W), focal2fov(fy, H), W, H, delta_tau=0.05, device="cpu")
    c.fx, c.fy = fx, fy
    return c


out = dict(twists=twists, se3_exp=exp, so3_log=logs, R0=R0, T0=T0, w0=w0, v0=v0)
c = cam()
out["last_vel_transform"] = c.last_vel_transform.detach().numpy()
out["next_vel_transform"] = c.next_vel_transform.detach().numpy()
out["last_vel_transform_inv"] = c.last_vel_transform_inv.detach().numpy()
out["world_view_transform"] = c.world_view_transform.detach().numpy()
out["full_proj_transform"] = c.full_proj_transform.detach().numpy()
out["projection_matrix"] = c.projection_matrix.detach().numpy()
out["camera_center"] = c.camera_center.detach().numpy()
out["proj_raw"] = getProjectionMatrix(0.01, 100.0, focal2fov(fx, W), focal2fov(fy, H)).numpy()
c.const_vel_model(0.05)
out["cv_R"], out["cv_T"] = c.R.numpy().copy(), c.T.numpy().copy()
with torch.no_grad():
    c.cam_rot_delta.copy_(torch.tensor([0.004, -0.003, 0.002]))
    c.cam_trans_delta.copy_(torch.tensor([-0.004, 0.004, 0.001]))
    c.cam_w_delta.copy_(torch.tensor([0.002, -0.002, 0.0005]))
    c.cam_v_delta.copy_(torch.tensor([-0.001, 0.002, 0.002]))
    c.update_vwRT()
out["up_R"], out["up_T"] = c.R.detach().numpy().copy(), c.T.detach().numpy().copy()
out["up_w"], out["up_v"] = c.angular_vel.detach().numpy().copy(), c.linear_vel.detach().numpy().copy()
c.cal_weighted_velocity([torch.from_numpy(T0.copy()), torch.from_numpy(R0.copy())], 0.05, 0.5)
out["wv_w"], out["wv_v"] = c.angular_vel.detach().numpy().copy(), c.linear_vel.detach().numpy().copy()
np.savez_compressed(os.path.join(HERE, "pose_algebra.npz"), **out)
print("wrote", os.path.join(HERE, "pose_algebra.npz"), {k: v.shape for k, v in out.items()})
