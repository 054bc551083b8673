"""RenderFrame: normalised intensity-change frame rendered from the map (mirror of the reference's
utils/render_camera/frame.py:10-94).  Operator-level (autograd) path."""
import torch

from utils.render_camera.camera import Camera
from gaussian_splatting.gaussian_renderer import render1, render2
from gaussian_splatting.utils.graphics_utils import focal2fov

_GRAY = (0.2989, 0.5870, 0.1140)


class RenderFrame:
    def __init__(self, viewpoint, gaussians, pipeline, background, pyramid_level):
        self.viewpoint, self.gaussians, self.pipeline, self.background = viewpoint, gaussians, pipeline, background
        self.sign_delta_Ir, self.unsign_delta_Ir = self.get_delta_Ir(pyramid_level)

    def render(self):
        return render1(self.viewpoint, self.gaussians, self.background)

    @property
    def intensity_frame(self):
        return self.get_intensity_frame(self.render()["render"])

    @property
    def depth_frame(self):
        return self.render()["depth"]

    def get_intensity_frame(self, color_frame):
        w = torch.tensor(_GRAY, device=color_frame.device).view(1, 3, 1, 1)
        return (color_frame * w).sum(dim=1)

    def get_delta_Ir(self, pyramid_level=0):
        vp = self.viewpoint
        assert vp.delta_tau != 0, "delta_tau must equal the time span of the event frame"
        s = 0.5 ** pyramid_level
        new_w, new_h = int(vp.image_width * s), int(vp.image_height * s)
        cams = []
        for T_vel in (vp.last_vel_transform, vp.next_vel_transform):
            pose = T_vel @ vp.curr_pose
            cams.append(Camera(pose[:3, :3], pose[:3, 3], vp.angular_vel, vp.linear_vel, focal2fov(vp.fx * s, new_w),
                               focal2fov(vp.fy * s, new_h), new_w, new_h, delta_tau=vp.delta_tau, device=vp.device))
        last_pkg, next_pkg = render2(cams[0], vp, cams[1], self.gaussians, self.background)
        delta = self.get_intensity_frame(next_pkg["render"]) - self.get_intensity_frame(last_pkg["render"])
        normalized = delta / torch.norm(delta, p=2)
        return normalized, torch.abs(normalized)
