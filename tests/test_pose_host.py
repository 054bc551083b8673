"""Host-side pose / camera algebra (the product's mirror of utils/pose.py, utils/render_camera/camera.py,
gaussian_splatting/utils/graphics_utils.py) and the oracle's numpy restatement, against vectors produced by
the reference's own Python (tests/golden/make_pose_golden.py).  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import oracle as orc

G = np.load(os.path.join(H.GOLDEN, "pose_algebra.npz"))
TOL = 2e-6   # fp32 algebra evaluated in a different but equivalent operation order


def test_oracle_se3_exp_matches_reference():
    for xi, ref in zip(G["twists"], G["se3_exp"]):
        assert np.abs(orc.SE3_exp(xi) - ref).max() < TOL


def test_host_pose_module_matches_reference():
    from utils.pose import SE3_exp, SO3_exp, SO3_log, V
    for xi, ref, lg in zip(G["twists"], G["se3_exp"], G["so3_log"]):
        T = SE3_exp(torch.from_numpy(xi)).numpy()
        assert np.abs(T - ref).max() < TOL
        assert np.abs(SO3_log(torch.from_numpy(ref[:3, :3].copy())).numpy() - lg).max() < 2e-4   # acos near 1 is ill-conditioned
    # small-angle branch: exactly I + W + W^2/2
    th = torch.tensor([2e-6, -1e-6, 3e-6])
    W = np.array([[0, -3e-6, -1e-6], [3e-6, 0, -2e-6], [1e-6, 2e-6, 0]], np.float32)
    assert np.allclose(SO3_exp(th).numpy(), np.eye(3, dtype=np.float32) + W + 0.5 * W @ W, atol=1e-12)
    assert np.allclose(V(th).numpy(), np.eye(3, dtype=np.float32) + 0.5 * W + W @ W / 6, atol=1e-12)


def _cam():
    from utils.render_camera.camera import Camera
    from gaussian_splatting.utils.graphics_utils import focal2fov
    c = Camera(torch.from_numpy(G["R0"].copy()), torch.from_numpy(G["T0"].copy()), torch.from_numpy(G["w0"].copy()),
               torch.from_numpy(G["v0"].copy()), focal2fov(327.32749, 640), focal2fov(327.46184, 480), 640, 480, delta_tau=0.05, device="cpu")
    c.fx, c.fy = 327.32749, 327.46184
    return c


def test_camera_mirror_matches_reference():
    c = _cam()
    for name in ("last_vel_transform", "next_vel_transform", "last_vel_transform_inv", "world_view_transform",
                 "full_proj_transform", "projection_matrix", "camera_center"):
        got = getattr(c, name).detach().numpy()
        assert np.abs(got - G[name]).max() < 5e-6 * max(1.0, np.abs(G[name]).max()), name
    c.const_vel_model(0.05)
    assert np.abs(c.R.numpy() - G["cv_R"]).max() < TOL and np.abs(c.T.numpy() - G["cv_T"]).max() < 5e-6
    with torch.no_grad():
        c.cam_rot_delta.copy_(torch.tensor([0.004, -0.003, 0.002]))
        c.cam_trans_delta.copy_(torch.tensor([-0.004, 0.004, 0.001]))
        c.cam_w_delta.copy_(torch.tensor([0.002, -0.002, 0.0005]))
        c.cam_v_delta.copy_(torch.tensor([-0.001, 0.002, 0.002]))
        c.update_vwRT()
    assert np.abs(c.R.detach().numpy() - G["up_R"]).max() < TOL and np.abs(c.T.detach().numpy() - G["up_T"]).max() < 5e-6
    assert np.array_equal(c.angular_vel.detach().numpy(), G["up_w"]) and np.array_equal(c.linear_vel.detach().numpy(), G["up_v"])
    assert float(c.cam_rot_delta.abs().sum() + c.cam_w_delta.abs().sum()) == 0.0
    c.cal_weighted_velocity([torch.from_numpy(G["T0"].copy()), torch.from_numpy(G["R0"].copy())], 0.05, 0.5)
    assert np.abs(c.angular_vel.detach().numpy() - G["wv_w"]).max() < 2e-4 and np.abs(c.linear_vel.detach().numpy() - G["wv_v"]).max() < 1e-5


def test_oracle_view_setup_matches_reference_camera():
    """The oracle's render2 view construction (used by every tracking-objective parity test) reproduces the
    reference Camera's matrices: vel transforms, projection, view = T_vel * T_cur."""
    views = orc.view_setup(G["R0"], G["T0"], G["w0"], G["v0"], 0.05, 640, 480, 327.32749, 327.46184, 0)
    assert np.abs(views[0]["vel"].reshape(4, 4).T - G["last_vel_transform"]).max() < TOL
    assert np.abs(views[1]["vel"].reshape(4, 4).T - G["next_vel_transform"]).max() < TOL
    assert np.abs(views[0]["vel_inv"].reshape(4, 4).T - G["last_vel_transform_inv"]).max() < TOL
    assert np.abs(views[0]["projmatrix_raw"].reshape(4, 4).T - G["proj_raw"]).max() < 1e-6
    assert views[0]["delta_time"] == -views[1]["delta_time"] == pytest.approx(-0.025)
    cur = np.eye(4, dtype=np.float32)
    cur[:3, :3], cur[:3, 3] = G["R0"], G["T0"]
    assert np.abs(views[1]["viewmatrix"].reshape(4, 4).T - G["next_vel_transform"] @ cur).max() < 5e-6
    # pyramid levels keep the field of view (frame.py:75-82)
    v2 = orc.view_setup(G["R0"], G["T0"], G["w0"], G["v0"], 0.05, 640, 480, 327.32749, 327.46184, 2)
    assert (v2[0]["W"], v2[0]["H"]) == (160, 120) and v2[0]["tanfovx"] == pytest.approx(views[0]["tanfovx"], rel=1e-12)


def test_ply_round_trip_and_config_surface(tmp_path):
    """GaussianModel.load_ply / save_ply (CPU tensors), the vendored PLY reader, munchify, make_config."""
    from gsevt import synth
    from gsevt.compat import munchify, natsorted, read_ply_vertices
    from gaussian_splatting.scene.gaussian_model import GaussianModel
    raw = synth.synth_map(500, seed=3)
    p = str(tmp_path / "m" / "point_cloud.ply")
    synth.save_map_ply(p, raw)
    v = read_ply_vertices(p)
    assert v.shape[0] == 500 and v.dtype.names[:6] == ("x", "y", "z", "nx", "ny", "nz") and "f_rest_44" in v.dtype.names
    gm = GaussianModel(3, device="cpu")
    gm.load_ply(p)
    assert np.array_equal(gm.get_xyz.numpy(), raw["xyz"])
    assert np.array_equal(gm._features_dc.numpy(), raw["f_dc"]) and np.array_equal(gm._features_rest.numpy(), raw["f_rest"])
    act = synth.activate(raw)
    assert np.allclose(gm.get_scaling.numpy(), act["scales"], rtol=1e-6) and np.allclose(gm.get_opacity.numpy(), act["opacities"], atol=1e-7)
    assert np.allclose(gm.get_rotation.numpy(), act["rotations"], atol=1e-6) and gm.get_features.shape == (500, 16, 3)
    p2 = str(tmp_path / "again.ply")
    gm.save_ply(p2)
    assert np.array_equal(read_ply_vertices(p2)["f_rest_17"], v["f_rest_17"])
    cfg = synth.make_config(p, "ev.txt", str(tmp_path / "out"))
    m = munchify(cfg)
    assert m.Gaussian.calib_params.fx == cfg["Gaussian"]["calib_params"]["fx"] and m.Optimizer.max_optim_iter == 200
    assert natsorted(["frame_10.png", "frame_2.png"]) == ["frame_2.png", "frame_10.png"]
    from utils.render_camera.camera import Camera
    cfg["Gaussian"]["model_params"]["device"] = "cpu"
    cam = Camera.init_from_yaml(cfg)
    assert cam.image_width == 640 and abs(cam.FoVx - 2 * math.atan(640 / (2 * 327.32749))) < 1e-12
    assert np.allclose(cam.R.detach().numpy(), np.array(synth.DESK["R"], np.float32).reshape(3, 3), atol=1e-7)


def test_tracker_host_helpers():
    """check_convergence / image_pyramid / tracking_loss keep the reference's semantics (tracker.py:65-103)."""
    from utils.tracker import Tracker
    t = Tracker.__new__(Tracker)
    t.pyramid_lvl = 3
    assert not t.check_convergence([1.0] * 10, 1e-4)                 # needs more than 10 losses
    assert t.check_convergence([1.0] * 11, 1e-4)
    assert not t.check_convergence(list(np.linspace(1, 0, 11)), 1e-4)
    assert t.check_convergence([5.0] + [1.0 + 5e-5 * (i % 2) for i in range(11)], 1e-4)   # only the last 11 count
    img = torch.arange(480 * 640, dtype=torch.float32).view(1, 480, 640)
    pyr = t.image_pyramid(img)
    assert [tuple(p.shape) for p in pyr] == [(1, 480, 640), (1, 240, 320), (1, 120, 160)]
    assert torch.equal(pyr[2], img[:, ::4, ::4])
    a, b = torch.tensor([[3.0, 0.0]]), torch.tensor([[0.0, 4.0]])
    assert float(t.tracking_loss(a, b)) == 5.0
