"""Camera: pose / velocity state of the tracked event camera (mirror of the reference's
utils/render_camera/camera.py:9-219 — same attributes, properties and methods)."""
import torch
from torch import nn

from gsevt.compat import munchify
from utils.pose import SE3_exp, SO3_log
from gaussian_splatting.utils.graphics_utils import getWorld2View2, getProjectionMatrix, focal2fov


def _se3_inv(T):
    """Closed-form inverse of a rigid transform (the reference calls torch.linalg.inv, camera.py:108,118)."""
    R, t = T[:3, :3], T[:3, 3]
    out = torch.eye(4, device=T.device, dtype=T.dtype)
    out[:3, :3] = R.t()
    out[:3, 3] = -(R.t() @ t)
    return out


class Camera(nn.Module):
    def __init__(self, R, t, angular_vel, linear_vel, fovx, fovy, image_width, image_height, delta_tau=0, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.R = R.to(self.device)          # world -> camera rotation
        self.T = t.to(self.device)          # world -> camera translation
        self.angular_vel = angular_vel
        self.linear_vel = linear_vel
        self.FoVx, self.FoVy = fovx, fovy
        self.image_width, self.image_height = image_width, image_height
        self.zfar, self.znear = 100.0, 0.01
        self.delta_tau = delta_tau
        # optimisation variables: left-multiplied twist increments of pose and velocity
        self.cam_rot_delta = nn.Parameter(torch.zeros(3, device=self.device), requires_grad=True)
        self.cam_trans_delta = nn.Parameter(torch.zeros(3, device=self.device), requires_grad=True)
        self.cam_w_delta = nn.Parameter(torch.zeros(3, device=self.device), requires_grad=False)
        self.cam_v_delta = nn.Parameter(torch.zeros(3, device=self.device), requires_grad=False)

    # ---- matrices handed to the rasteriser (row-major torch tensors of the TRANSPOSED matrices) ----
    @property
    def projection_matrix(self):
        return getProjectionMatrix(self.znear, self.zfar, self.FoVx, self.FoVy).transpose(0, 1).to(self.device)

    @property
    def world_view_transform(self):
        return getWorld2View2(self.R, self.T).transpose(0, 1)

    @property
    def full_proj_transform(self):
        return self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0)).squeeze(0)

    @property
    def camera_center(self):
        return -(self.R.t() @ self.T)

    @property
    def curr_pose(self):
        pose = torch.eye(4, device=self.device)
        pose[0:3, 0:3] = self.R
        pose[0:3, 3] = self.T
        return pose

    # ---- constant-velocity half-interval transforms ----
    def compute_motion_vectors(self):
        return self.angular_vel * (self.delta_tau / 2), self.linear_vel * (self.delta_tau / 2)

    @property
    def last_vel_transform(self):
        rot_vec, trans_vec = self.compute_motion_vectors()
        return SE3_exp(torch.cat([-trans_vec, -rot_vec], dim=0))

    @property
    def last_vel_transform_inv(self):
        return _se3_inv(self.last_vel_transform)

    @property
    def next_vel_transform(self):
        rot_vec, trans_vec = self.compute_motion_vectors()
        return SE3_exp(torch.cat([trans_vec, rot_vec], dim=0))

    @property
    def next_vel_transform_inv(self):
        return _se3_inv(self.next_vel_transform)

    # ---- state updates ----
    def update_RT(self, R, t):
        self.R = R.to(device=self.device)
        self.T = t.to(device=self.device)

    def update_pose(self):
        new_w2c = SE3_exp(torch.cat([self.cam_trans_delta, self.cam_rot_delta], dim=0)) @ self.curr_pose
        self.update_RT(new_w2c[0:3, 0:3], new_w2c[0:3, 3])
        self.cam_rot_delta.data.fill_(0)
        self.cam_trans_delta.data.fill_(0)

    def update_velocity(self):
        self.angular_vel += self.cam_w_delta
        self.linear_vel += self.cam_v_delta
        self.cam_w_delta.data.fill_(0)
        self.cam_v_delta.data.fill_(0)

    def update_vwRT(self):
        self.update_velocity()
        self.update_pose()

    def cal_weighted_velocity(self, last_data, delta_tau, weight):
        last_pose = torch.eye(4, device=self.device)
        last_pose[0:3, 0:3] = last_data[1]
        last_pose[0:3, 3] = last_data[0]
        delta_pose = self.curr_pose.detach() @ _se3_inv(last_pose)
        linear_velocity = delta_pose[0:3, 3] / delta_tau
        angular_velocity = SO3_log(delta_pose[0:3, 0:3]) / delta_tau
        self.linear_vel = weight * linear_velocity + (1 - weight) * self.linear_vel
        self.angular_vel = weight * angular_velocity + (1 - weight) * self.angular_vel

    def const_vel_model(self, tau):
        step = SE3_exp(torch.cat([self.linear_vel * tau, self.angular_vel * tau], dim=0))
        new_pose = step @ self.curr_pose
        self.last_R, self.last_T = self.R.clone(), self.T.clone()
        self.update_RT(new_pose[:3, :3], new_pose[:3, 3])

    @staticmethod
    def init_from_yaml(config):
        g = config["Gaussian"]
        device = g["model_params"]["device"]
        calib = munchify(g["calib_params"])
        R = torch.tensor(config["Tracking"]["initial_pose"]["rot"]["data"]).reshape(3, 3)
        t = torch.tensor(config["Tracking"]["initial_pose"]["trans"]["data"]).reshape(3,)
        lin = torch.tensor(config["Tracking"]["initial_vel"]["linear_vel"], device=device, dtype=torch.float32)
        ang = torch.tensor(config["Tracking"]["initial_vel"]["angular_vel"], device=device, dtype=torch.float32)
        cam = Camera(R, t, ang, lin, focal2fov(calib.fx, g["img_width"]), focal2fov(calib.fy, g["img_height"]),
                     g["img_width"], g["img_height"], device=device)
        cam.fx, cam.fy = calib.fx, calib.fy
        cam.update_pose()
        return cam
