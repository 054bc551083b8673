"""Tracker: frame loop -> pyramid levels -> two-stage Adam on pose / velocity (mirror of the reference's
utils/tracker.py:33-287: same constructor, same methods, same outputs tracking_pose_tum.txt and
tracking_log.log).

Two execution paths produce the same trajectory:
  * engine (default): the innermost loop (tracker.py:176-240 of the reference) runs entirely on the
    device inside libgsevt (gsevt.engine.TrackingEngine) — no per-iteration host sync;
  * autograd (config["Gsevt"]["use_engine"] = False): the reference's loop structure in PyTorch, with
    RenderFrame / torch.optim.Adam and the drop-in diff_gaussian_rasterization operator.
"""
import copy
import logging
import os
import time
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from utils.auxiliary import Logger
from utils.event_camera.event import EventArray, EventFrame
from utils.render_camera.camera import Camera
from utils.render_camera.frame import RenderFrame


def _quat_xyzw(Rm):
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(np.asarray(Rm, dtype=np.float64)).as_quat()


class Tracker:
    def __init__(self, config, event_arrays: List[EventArray], viewpoint: Camera, gaussians, pipeline, background, device):
        self.device = device
        self.config = config
        self.event_arrays = event_arrays
        self.viewpoint = viewpoint
        self.gaussians = gaussians
        self.pipeline = pipeline
        self.background = background
        ev, opt = config["Event"], config["Optimizer"]
        self.img_width, self.img_height = ev["img_width"], ev["img_height"]
        self.gaussian_kernel_size = ev["gaussian_kernel_size"]
        self.intrinsic = np.array(ev["intrinsic"]["data"]).reshape(3, 3)
        self.distortion_factors = np.array(ev["distortion_factors"])
        self.converged_threshold = opt["converged_threshold"]
        self.max_optim_iter = opt["max_optim_iter"]
        self.save_path = config["Tracking"]["save_path"]
        self.pyramid_lvl = 3
        extra = config.get("Gsevt", {}) or {}
        self.use_engine = bool(extra.get("use_engine", True))
        self.save_frames = bool(extra.get("save_frames", False))
        self.chunk = int(extra.get("iterations_per_poll", 8))
        self.iter_counts = []       # per frame: [(level, coarse_iters, fine_iters, seconds)]
        self.trajectory = []        # per frame: (timestamp, T(3), quat xyzw)
        self.velocities = []        # per frame: (angular_vel(3), linear_vel(3)) after the frame's optimisation (engine path)
        os.makedirs(self.save_path, exist_ok=True)
        self.log = Logger(name="TrackingLogger", log_file=f"{self.save_path}/tracking_log.log", level=logging.INFO)

    # ---- helpers kept for API compatibility ---------------------------------------------------
    def check_convergence(self, losses, threshold=1e-4):
        if len(losses) <= 10:
            return False
        return bool(np.mean(np.abs(np.diff(losses[-11:]))) < threshold)

    def image_pyramid(self, image):
        """Nearest-neighbour pyramid: cv2.resize(image, (int(w s), int(h s)), INTER_NEAREST) with s = 0.5**l, i.e.
        source index min(floor(dst / (dst_size / src_size)), src_size - 1) — image[..., ::2**l, ::2**l] when the size is a
        multiple of 2**l."""
        _, h, w = image.shape
        out = []
        for l in range(self.pyramid_lvl):
            hl, wl = int(h * 0.5 ** l), int(w * 0.5 ** l)
            iy = np.minimum(np.floor(np.arange(hl) * (1.0 / (hl / h))).astype(np.int64), h - 1)
            ix = np.minimum(np.floor(np.arange(wl) * (1.0 / (wl / w))).astype(np.int64), w - 1)
            iy, ix = torch.from_numpy(iy).to(image.device), torch.from_numpy(ix).to(image.device)
            out.append(image[:, iy][:, :, ix].contiguous())
        return out

    def tracking_loss(self, delta_Ir, delta_Ie, mask=None, huber=False):
        residual = delta_Ir * mask - delta_Ie if mask is not None else delta_Ir - delta_Ie
        if huber:
            return torch.sum(F.huber_loss(residual, torch.zeros_like(residual), delta=0.002, reduction="none"))
        return torch.norm(residual)

    # ---- main loop ----------------------------------------------------------------------------
    def tracking(self):
        os.makedirs(self.save_path, exist_ok=True)
        if self.save_frames:
            os.makedirs(f"{self.save_path}/tracking_frames", exist_ok=True)
        t0 = time.time()
        with open(f"{self.save_path}/tracking_pose_tum.txt", "w") as tum:
            if self.use_engine:
                self._tracking_engine(tum)
            else:
                self._tracking_autograd(tum)
        self.log.info(f"totoal tracking time cost: {time.time() - t0:.4f}s")

    def _write_tum(self, tum, frame_idx, Rm, T):
        ts = self.event_arrays[frame_idx].time()
        q = _quat_xyzw(Rm)
        self.trajectory.append((ts, np.array(T, np.float64), q))
        tum.write(f"{ts} {T[0]} {T[1]} {T[2]} {q[0]} {q[1]} {q[2]} {q[3]}\n")
        tum.flush()

    def _make_engine(self):
        from gsevt.engine import TrackingEngine
        opt = self.config["Optimizer"]
        vp = self.viewpoint
        eng = TrackingEngine(self.gaussians.packed(), vp.image_width, vp.image_height, vp.fx, vp.fy,
                             background=[float(b) for b in self.background.detach().cpu().tolist()],
                             levels=self.pyramid_lvl, lr_rot=opt["cam_rot_delta"], lr_trans=opt["cam_trans_delta"],
                             lr_w=opt["cam_w_delta"], lr_v=opt["cam_v_delta"],
                             converged_threshold=self.converged_threshold, max_optim_iter=self.max_optim_iter,
                             znear=vp.znear, zfar=vp.zfar)
        eng.set_state(vp.R.detach().cpu().numpy(), vp.T.detach().cpu().numpy(),
                      vp.angular_vel.detach().cpu().numpy(), vp.linear_vel.detach().cpu().numpy())
        return eng

    def _tracking_engine(self, tum):
        eng = self.engine = self._make_engine()
        vp = self.viewpoint
        last_delta_tau = 0
        n_frames = len(self.event_arrays)
        for frame_idx in range(n_frames):
            ea = self.event_arrays[frame_idx]
            delta_tau = ea.duration()
            vp.delta_tau = delta_tau
            eng.const_vel_model((delta_tau + last_delta_tau) / 2)
            R0, T0, _, _ = eng.get_state()
            self.log.info(f"frame_idx:\t{frame_idx} / {n_frames}")
            self.log.info(f"delta_tau:\t{delta_tau:.4f}")
            eFrame = EventFrame(self.img_width, self.img_height, self.intrinsic, self.distortion_factors,
                                self.gaussian_kernel_size, ea, device=vp.device)
            eng.begin_frame(delta_tau, eFrame.sign_pyramid, eFrame.unsign_pyramid)
            per_level = []
            for lvl in range(self.pyramid_lvl - 1, -1, -1):
                t0 = time.time()
                st = eng.run_level(lvl, opt_vel=(lvl != self.pyramid_lvl - 1), chunk=self.chunk)
                dt = time.time() - t0
                per_level.append((lvl, st.start_vel_opt_iter, st.optim_iter - st.start_vel_opt_iter, dt))
                self.log.info(f"level:\t{lvl}")
                self.log.info(f"optim_iter:\t{st.optim_iter} ({st.start_vel_opt_iter}+{st.optim_iter - st.start_vel_opt_iter})")
                self.log.info(f"opt_time:\t{dt:.4f}")
            self.iter_counts.append(per_level)
            if frame_idx >= 5:
                eng.weighted_velocity(R0, T0, (delta_tau + last_delta_tau) / 2, 0.5)
            last_delta_tau = delta_tau
            Rm, T, w, v = eng.get_state()
            self.velocities.append((np.array(w, np.float64), np.array(v, np.float64)))
            dev = vp.device
            vp.update_RT(torch.from_numpy(Rm.copy()).to(dev), torch.from_numpy(T.copy()).to(dev))
            vp.angular_vel, vp.linear_vel = torch.from_numpy(w.copy()).to(dev), torch.from_numpy(v.copy()).to(dev)
            self.log.info("=" * 20)
            self._write_tum(tum, frame_idx, Rm, T)

    def _tracking_autograd(self, tum):
        cfg_opt = self.config["Optimizer"]
        vp = self.viewpoint
        last_delta_tau = 0
        fraction_num = self.max_optim_iter / 2
        n_frames = len(self.event_arrays)
        names = ("cam_rot_delta", "cam_trans_delta", "cam_w_delta", "cam_v_delta")
        for frame_idx in range(n_frames):
            params = [getattr(vp, n) for n in names]
            optimizer = torch.optim.Adam([{"params": [p], "lr": cfg_opt[n]} for p, n in zip(params, names)])
            delta_tau = self.event_arrays[frame_idx].duration()
            vp.delta_tau = delta_tau
            vp.const_vel_model((delta_tau + last_delta_tau) / 2)
            initial = copy.deepcopy([vp.T.detach(), vp.R.detach()])
            self.log.info(f"frame_idx:\t{frame_idx} / {n_frames}")
            self.log.info(f"delta_tau:\t{delta_tau:.4f}")
            eFrame = EventFrame(self.img_width, self.img_height, self.intrinsic, self.distortion_factors,
                                self.gaussian_kernel_size, self.event_arrays[frame_idx], device=vp.device)
            sign_pyr = self.image_pyramid(eFrame.sign_delta_Ie)
            unsign_pyr = self.image_pyramid(eFrame.unsign_delta_Ie)
            per_level = []
            for lvl in range(self.pyramid_lvl - 1, -1, -1):
                opt_vel = lvl != self.pyramid_lvl - 1
                start_vel_opt_iter = 0
                for group, n in zip(optimizer.param_groups, names):
                    group["lr"] = cfg_opt[n]
                losses, optim_iter = [], 0
                t0 = time.time()
                while True:
                    vp.cam_w_delta.requires_grad_(opt_vel)
                    vp.cam_v_delta.requires_grad_(opt_vel)
                    vp.cam_rot_delta.requires_grad_(True)
                    vp.cam_trans_delta.requires_grad_(True)
                    if opt_vel:
                        k = optim_iter - start_vel_opt_iter
                        fraction = k / fraction_num if 1 <= k <= fraction_num else 1
                        for group, n in zip(optimizer.param_groups, names):
                            group["lr"] = cfg_opt[n] * (fraction if n in names[:2] else (1 - fraction))
                    rFrame = RenderFrame(vp, self.gaussians, self.pipeline, self.background, lvl)
                    if not opt_vel:
                        loss = self.tracking_loss(rFrame.unsign_delta_Ir, unsign_pyr[lvl])
                    else:
                        loss = self.tracking_loss(rFrame.sign_delta_Ir, sign_pyr[lvl])
                    loss.backward()
                    losses.append(loss.item())
                    with torch.no_grad():
                        optimizer.step()
                        converged = self.check_convergence(losses, self.converged_threshold)
                        if not opt_vel:
                            vp.update_pose()
                        else:
                            vp.update_vwRT()
                        optimizer.zero_grad()
                    if converged:
                        if not opt_vel:
                            opt_vel, start_vel_opt_iter = True, optim_iter
                        else:
                            break
                    if not opt_vel:
                        if optim_iter >= self.max_optim_iter:
                            self.log.error("coarse stage optimization iter exceeded the max_optim_iter!")
                            break
                    elif optim_iter >= start_vel_opt_iter + self.max_optim_iter:
                        self.log.error("fine stage optimization iter exceeded the max_optim_iter!")
                        break
                    optim_iter += 1
                dt = time.time() - t0
                per_level.append((lvl, start_vel_opt_iter, optim_iter - start_vel_opt_iter, dt))
                self.log.info(f"level:\t{lvl}")
                self.log.info(f"optim_iter:\t{optim_iter} ({start_vel_opt_iter}+{optim_iter - start_vel_opt_iter})")
                self.log.info(f"opt_time:\t{dt:.4f}")
            self.iter_counts.append(per_level)
            if frame_idx >= 5:
                vp.cal_weighted_velocity(initial, (delta_tau + last_delta_tau) / 2, 0.5)
            last_delta_tau = delta_tau
            self.log.info("=" * 20)
            self._write_tum(tum, frame_idx, vp.R.detach().cpu().numpy(), vp.T.detach().cpu().numpy())
