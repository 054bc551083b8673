"""Seeded synthetic inputs for the GS-EVT tracking path (no dataset is available offline).

Definitions follow SURVEY.md §8(d): desk_normal1 intrinsics / initial pose / velocity / learning rates
(configs/VECTOR/desk_normal1_config.yaml of the reference), maps of P Gaussians drawn in the initial
camera frame, and event packets sampled from a ground-truth intensity-change render.
Everything is numpy + `numpy.random.default_rng(seed)`; nothing here touches the GPU by itself.
"""
import math
import os

import numpy as np

DESK = dict(
    W=640, H=480, fx=327.32749, fy=327.46184, cx=304.97749, cy=235.37621,
    dist=[-0.031982, 0.041966, -0.000507, -0.001031, 0.0],
    R=[0.96273041, 0.13706175, -0.23316138, -0.15101728, 0.98759518, -0.04300630, 0.22437454, 0.07661487, 0.97148661],
    T=[-3.99364391, 2.07062047, -1.35796279],
    angular_vel=[0.0025, 0.0009, 0.0010], linear_vel=[0.0134, 0.1050, 0.0996],
    lr=dict(cam_rot_delta=0.004, cam_trans_delta=0.004, cam_v_delta=0.002, cam_w_delta=0.002),
    converged_threshold=1e-4, max_optim_iter=200, max_events_per_frame=30000,
)

# configs/VECTOR/robot_normal1_config.yaml: same sensor, different initial pose / velocity (config 3).
ROBOT = dict(DESK,
             R=[0.98161380, -0.00045487, 0.19087728, -0.02444478, 0.99146338, 0.12807349, -0.18930609, -0.13038466, 0.97322302],
             T=[-4.80009293, -0.70624608, 2.88821495],
             angular_vel=[0.04814772, 0.04794633, 0.00379134], linear_vel=[0.11414605, 0.00285491, -0.19478575])


def synth_map(P, seed=0, W=DESK["W"], H=DESK["H"], fx=DESK["fx"], fy=DESK["fy"], R=DESK["R"], T=DESK["T"],
              sh_degree=3, structure=0, structure_scale=0.25, fine_opacity_shift=0.0, structure_depth=(3.0, 6.0)):
    """Raw (pre-activation) 3DGS parameters, float32, as a dict:
    xyz (P,3), f_dc (P,1,3), f_rest (P,15,3), opacity (P,1) logits, scaling (P,3) log-scales, rotation (P,4).

    structure > 0: a TRACKABLE scene (DESIGN.md "Sequences").  `structure` of the P Gaussians become large, nearly opaque,
    strongly coloured, view-independent splats (extent `structure_scale` metres at 3..6 m: tens of pixels, well above the
    9x9 blur of the event frame) spread over 1.6x the initial frustum, and the remaining fine splats get their opacity
    logits shifted by `fine_opacity_shift` (negative: a faint texture on top of the structure instead of a noise wall)."""
    if structure > 0:
        return _structured_map(P, seed, W, H, fx, fy, R, T, sh_degree, int(structure), float(structure_scale), float(fine_opacity_shift), structure_depth)
    rng = np.random.default_rng(seed)
    R0 = np.asarray(R, np.float64).reshape(3, 3)
    T0 = np.asarray(T, np.float64).reshape(3)
    tanx, tany = W / (2 * fx), H / (2 * fy)
    n_back = P // 20
    n_front = P - n_back
    z = np.concatenate([rng.uniform(1.0, 6.0, n_front), rng.uniform(-2.0, 0.2, n_back)])
    zz = np.where(z > 0.5, z, 0.5 + np.abs(z))
    pc = np.stack([zz * tanx * rng.uniform(-1.25, 1.25, P), zz * tany * rng.uniform(-1.25, 1.25, P), z], axis=1)
    perm = rng.permutation(P)           # no spatial order in memory, like a trained map
    pc = pc[perm]
    xyz = (pc - T0) @ R0                # p_w = R0^T (p_c - T0)
    mu = math.log(0.02 * (3e5 / P) ** (1.0 / 3.0))
    ncoef = (sh_degree + 1) ** 2
    return dict(
        xyz=xyz.astype(np.float32),
        scaling=rng.normal(mu, 0.6, (P, 3)).astype(np.float32),
        rotation=rng.normal(0.0, 1.0, (P, 4)).astype(np.float32),
        opacity=rng.normal(0.5, 1.5, (P, 1)).astype(np.float32),
        f_dc=rng.normal(0.0, 1.0, (P, 1, 3)).astype(np.float32),
        f_rest=rng.normal(0.0, 0.1, (P, ncoef - 1, 3)).astype(np.float32),
    )


def _structured_map(P, seed, W, H, fx, fy, R, T, sh_degree, n_big, big_scale, fine_shift, depth=(3.0, 6.0)):
    base = synth_map(P, seed=seed, W=W, H=H, fx=fx, fy=fy, R=R, T=T, sh_degree=sh_degree)
    rng = np.random.default_rng(seed + 7919)
    n_big = min(n_big, P)
    R0 = np.asarray(R, np.float64).reshape(3, 3)
    T0 = np.asarray(T, np.float64).reshape(3)
    tanx, tany = W / (2 * fx), H / (2 * fy)
    sel = rng.choice(P, n_big, replace=False)          # spread over the index range like any other Gaussian
    z = rng.uniform(depth[0], depth[1], n_big)
    pc = np.stack([z * tanx * rng.uniform(-1.6, 1.6, n_big), z * tany * rng.uniform(-1.6, 1.6, n_big), z], axis=1)
    base["opacity"] = (base["opacity"] + np.float32(fine_shift)).astype(np.float32)
    # nothing within a metre of the camera plane: a splat that crosses the renderer's 0.2 m near cut covers a third of the
    # image and pops in or out from one iteration to the next (a step in the loss no optimiser can follow)
    cam = base["xyz"].astype(np.float64) @ R0.T + T0
    near = (cam[:, 2] > -1.0) & (cam[:, 2] < 1.0)
    cam[near, 2] -= 2.0
    base["xyz"] = ((cam - T0) @ R0).astype(np.float32)
    base["xyz"][sel] = ((pc - T0) @ R0).astype(np.float32)
    # extent proportional to the depth: every structure splat covers about the same number of pixels
    base["scaling"][sel] = (np.log(big_scale * z / 4.5)[:, None] + rng.normal(0.0, 0.25, (n_big, 1)) + rng.normal(0.0, 0.1, (n_big, 3))).astype(np.float32)
    base["opacity"][sel] = rng.normal(2.5, 0.5, (n_big, 1)).astype(np.float32)
    base["f_dc"][sel] = rng.normal(0.0, 1.6, (n_big, 1, 3)).astype(np.float32)
    base["f_rest"][sel] = 0.0
    return base


def activate(m):
    """The activations the reference applies (gaussian_model.py:75-95), in numpy float32."""
    rot = m["rotation"] / np.maximum(np.linalg.norm(m["rotation"], axis=1, keepdims=True), 1e-12).astype(np.float32)
    return dict(
        xyz=m["xyz"],
        scales=np.exp(m["scaling"]).astype(np.float32),
        rotations=rot.astype(np.float32),
        opacities=(1.0 / (1.0 + np.exp(-m["opacity"].astype(np.float64)))).astype(np.float32),
        shs=np.concatenate([m["f_dc"], m["f_rest"]], axis=1).astype(np.float32),
    )


def save_map_ply(path, m):
    """Standard 3DGS PLY (property order of gaussian_model.py:173-185; f_rest channel-major)."""
    from .compat import write_ply_vertices
    P = m["xyz"].shape[0]
    f_dc = m["f_dc"].transpose(0, 2, 1).reshape(P, -1)
    f_rest = m["f_rest"].transpose(0, 2, 1).reshape(P, -1)
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(f_dc.shape[1])] + \
            [f"f_rest_{i}" for i in range(f_rest.shape[1])] + ["opacity"] + [f"scale_{i}" for i in range(3)] + \
            [f"rot_{i}" for i in range(4)]
    data = np.concatenate([m["xyz"], np.zeros_like(m["xyz"]), f_dc, f_rest, m["opacity"], m["scaling"], m["rotation"]], axis=1)
    arr = np.empty(P, dtype=[(n, "f4") for n in names])
    for i, n in enumerate(names):
        arr[n] = data[:, i]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    write_ply_vertices(path, arr)


def load_map_into(gaussians, m, device="cuda"):
    """Fills a GaussianModel from a synth_map dict without a PLY round trip."""
    import torch
    from torch import nn
    mk = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float, device=device).contiguous(), requires_grad=False)
    gaussians._xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity = (
        mk(m["xyz"]), mk(m["scaling"]), mk(m["rotation"]), mk(m["opacity"]))
    gaussians._features_dc, gaussians._features_rest = mk(m["f_dc"]), mk(m["f_rest"])
    gaussians.active_sh_degree = gaussians.max_sh_degree
    gaussians._packed = None
    return gaussians


# ---- SE(3) in float64 (ground-truth trajectories) ----------------------------------------------
def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def se3_exp(xi):
    rho, th = np.asarray(xi[:3], np.float64), np.asarray(xi[3:], np.float64)
    W = skew(th)
    a = np.linalg.norm(th)
    if a < 1e-10:
        Rm, V = np.eye(3) + W + 0.5 * W @ W, np.eye(3) + 0.5 * W + W @ W / 6
    else:
        Rm = np.eye(3) + math.sin(a) / a * W + (1 - math.cos(a)) / a ** 2 * W @ W
        V = np.eye(3) + (1 - math.cos(a)) / a ** 2 * W + (a - math.sin(a)) / a ** 3 * W @ W
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = Rm, V @ rho
    return T


def ground_truth_trajectory(n_frames, dtau=0.05, R=DESK["R"], T=DESK["T"], lin=DESK["linear_vel"], ang=DESK["angular_vel"],
                            modulation=0.1, period=2.0, mode="drift", orbit_period=12.0):
    """Poses (world->camera 4x4) at frame mid-times and the velocities there.  Frame j spans [j*dtau, (j+1)*dtau].
    mode "drift": constant yaml velocity with +-10 % sinusoidal modulation (the camera leaves a frustum-sized synthetic map
    after a few hundred frames).  mode "orbit": the same speeds, but the body-frame velocity vectors turn at constant
    magnitude with period `orbit_period` seconds about the axis they make with the optical axis — the camera runs a closed
    loop of radius |v| * period / 2 pi (0.28 m at the desk yaml's 0.145 m/s) and never stands still, so a sequence of any
    length stays inside the map."""
    pose = np.eye(4)
    pose[:3, :3], pose[:3, 3] = np.asarray(R, np.float64).reshape(3, 3), np.asarray(T, np.float64)
    lin, ang = np.asarray(lin, np.float64), np.asarray(ang, np.float64)

    def basis(v):
        n = np.linalg.norm(v)
        e1 = v / n if n > 0 else np.array([1.0, 0, 0])
        k = np.cross(e1, np.array([0.0, 0.0, 1.0]))
        if np.linalg.norm(k) < 1e-6:
            k = np.cross(e1, np.array([0.0, 1.0, 0.0]))
        k /= np.linalg.norm(k)
        return n, e1, np.cross(k, e1)

    nl, l1, l2 = basis(lin)
    na, a1, a2 = basis(ang)

    def vel(t):
        if mode == "orbit":
            c, s_ = math.cos(2 * math.pi * t / orbit_period), math.sin(2 * math.pi * t / orbit_period)
            return nl * (c * l1 + s_ * l2), na * (c * a1 + s_ * a2)
        f = 1.0 + modulation * math.sin(2 * math.pi * t / period)
        return lin * f, ang * f

    out, sub = [], 10
    t = 0.0
    for j in range(n_frames):
        for k in range(sub):
            v, w = vel(t)
            if k == sub // 2:
                out.append((pose.copy(), v, w, t))
            h = dtau / sub
            pose = se3_exp(np.concatenate([v * h, w * h])) @ pose
            t += h
    return out


def distort_points(u, v, K, D):
    """Forward distortion (ideal pixel -> sensor pixel), the inverse of what cv2.undistort samples."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    k1, k2, p1, p2, k3 = (list(D) + [0, 0, 0, 0, 0])[:5]
    x, y = (u - cx) / fx, (v - cy) / fy
    r2 = x * x + y * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return xd * fx + cx, yd * fy + cy


def sample_events(delta_I, n_events, t_start_us, t_end_us, K, D, seed):
    """30 000 pixels drawn with probability ~ |delta_I| (rng seed), polarity = sign, forward-distorted to
    sensor pixels, sorted integer microsecond timestamps.  Returns int64 array (n,4): ts x y p."""
    rng = np.random.default_rng(seed)
    H, W = delta_I.shape
    w = np.abs(delta_I).astype(np.float64).ravel()
    if w.sum() <= 0:
        w = np.ones_like(w)
    idx = rng.choice(w.size, size=n_events, p=w / w.sum())
    v, u = np.divmod(idx, W)
    pol = (delta_I.ravel()[idx] > 0).astype(np.int64)
    ud, vd = distort_points(u.astype(np.float64), v.astype(np.float64), np.asarray(K, np.float64).reshape(3, 3), D)
    x = np.clip(np.rint(ud), 0, W - 1).astype(np.int64)
    y = np.clip(np.rint(vd), 0, H - 1).astype(np.int64)
    ts = np.sort(rng.integers(int(t_start_us), int(t_end_us) + 1, n_events))
    ts[0], ts[-1] = int(t_start_us), int(t_end_us)
    return np.stack([ts, x, y, pol], axis=1)


def threshold_events(delta_I, n_events, t_start_us, t_end_us, K, D, seed):
    """Events the way a sensor with a contrast threshold C fires them: a pixel emits floor(|delta_I| / C + u) events of the
    sign of delta_I (u ~ U[0,1): the charge left over from before the frame), with C chosen by bisection so that the frame
    holds n_events — the reference cuts the stream into packets of a fixed COUNT (event.py:69-90), so a frame must hold
    exactly that many; the last few are added / dropped at the pixels nearest to their next threshold crossing.  Same
    output format as sample_events: int64 (n, 4) ts x y p, forward-distorted to sensor pixels."""
    rng = np.random.default_rng(seed)
    H, W = delta_I.shape
    a = np.abs(delta_I).astype(np.float64).ravel()
    if a.sum() <= 0:
        return sample_events(delta_I, n_events, t_start_us, t_end_us, K, D, seed)
    u = rng.uniform(0.0, 1.0, a.size)
    lo, hi = 0.0, a.max() * 4.0 + 1e-30               # counts(C) is decreasing in C
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if np.floor(a / mid + u).sum() > n_events:
            lo = mid
        else:
            hi = mid
    x = a / hi + u
    cnt = np.floor(x).astype(np.int64)
    short = int(n_events - cnt.sum())                 # >= 0 by construction of hi
    if short > 0:
        frac = x - cnt
        cnt[np.argsort(-frac, kind="stable")[:short]] += 1
    elif short < 0:
        frac = np.where(cnt > 0, x - cnt, 2.0)
        cnt[np.argsort(frac, kind="stable")[:-short]] -= 1
    idx = np.repeat(np.arange(a.size), cnt)
    idx = idx[rng.permutation(idx.size)]
    v, uu = np.divmod(idx, W)
    pol = (delta_I.ravel()[idx] > 0).astype(np.int64)
    ud, vd = distort_points(uu.astype(np.float64), v.astype(np.float64), np.asarray(K, np.float64).reshape(3, 3), D)
    xs = np.clip(np.rint(ud), 0, W - 1).astype(np.int64)
    ys = np.clip(np.rint(vd), 0, H - 1).astype(np.int64)
    ts = np.sort(rng.integers(int(t_start_us), int(t_end_us) + 1, idx.size))
    ts[0], ts[-1] = int(t_start_us), int(t_end_us)
    return np.stack([ts, xs, ys, pol], axis=1)


def random_events(n_events, W, H, t_start_us, t_end_us, seed):
    rng = np.random.default_rng(seed)
    ts = np.sort(rng.integers(int(t_start_us), int(t_end_us) + 1, n_events))
    ts[0], ts[-1] = int(t_start_us), int(t_end_us)
    return np.stack([ts, rng.integers(0, W, n_events), rng.integers(0, H, n_events), rng.integers(0, 2, n_events)], axis=1)


def write_events_txt(path, ev):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    try:   # same bytes as np.savetxt(fmt="%d"), twice as fast on tens of millions of rows
        import pandas as pd
        pd.DataFrame(np.asarray(ev, np.int64)).to_csv(path, sep=" ", header=False, index=False)
    except ImportError:
        np.savetxt(path, ev, fmt="%d", delimiter=" ")


def make_config(map_path, events_path, save_path, W=DESK["W"], H=DESK["H"], device="cuda", centered=True, **over):
    """A yaml-compatible dict with the schema of configs/VECTOR/*.yaml.  For synthetic events the
    principal point is the image centre (the renderer's frustum is symmetric)."""
    cx, cy = (W / 2, H / 2) if centered else (DESK["cx"], DESK["cy"])
    s = W / DESK["W"]
    cfg = {
        "Event": {"data_path": events_path, "distortion_factors": list(DESK["dist"]), "filter_threshold": 0,
                  "img_height": H, "img_width": W,
                  "intrinsic": {"cols": 3, "rows": 3, "dt": "d",
                                "data": [DESK["fx"] * s, 0.0, cx, 0.0, DESK["fy"] * s, cy, 0.0, 0.0, 1.0]},
                  "gaussian_kernel_size": 9, "max_events_per_frame": DESK["max_events_per_frame"]},
        "Gaussian": {"calib_params": {"fx": DESK["fx"] * s, "fy": DESK["fy"] * s},
                     "model_params": {"background": [0, 0, 0], "device": device, "model_path": map_path, "sh_degree": 3},
                     "pipeline_params": {"compute_cov3D_python": False, "convert_SHs_python": False},
                     "img_height": H, "img_width": W},
        "Optimizer": {**DESK["lr"], "converged_threshold": DESK["converged_threshold"], "max_optim_iter": DESK["max_optim_iter"]},
        "Tracking": {"initial_pose": {"rot": {"cols": 3, "rows": 3, "dt": "d", "data": list(DESK["R"])},
                                      "trans": {"cols": 1, "rows": 3, "dt": "d", "data": list(DESK["T"])}},
                     "initial_vel": {"angular_vel": list(DESK["angular_vel"]), "linear_vel": list(DESK["linear_vel"])},
                     "save_path": save_path},
    }
    for k, v in over.items():
        sec, key = k.split("__")
        cfg[sec][key] = v
    return cfg


def camera_matrices(R, T, W, H, fx, fy, znear=0.01, zfar=100.0):
    """Float32 matrices the way the reference hands them to the rasteriser (transposed row-major ==
    column-major flat): viewmatrix, projmatrix, projmatrix_raw, campos, tanfovx, tanfovy."""
    R, T = np.asarray(R, np.float32).reshape(3, 3), np.asarray(T, np.float32).reshape(3)
    fovx, fovy = 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy))
    tanx, tany = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    V = np.eye(4, dtype=np.float32)
    V[:3, :3], V[:3, 3] = R, T
    Pm = np.zeros((4, 4), np.float32)
    top, right = math.tan(fovy / 2) * znear, math.tan(fovx / 2) * znear
    Pm[0, 0], Pm[1, 1] = 2.0 * znear / (2 * right), 2.0 * znear / (2 * top)
    Pm[3, 2], Pm[2, 2], Pm[2, 3] = 1.0, zfar / (zfar - znear), -(zfar * znear) / (zfar - znear)
    full = (Pm @ V).astype(np.float32)
    campos = (-(R.T @ T)).astype(np.float32)
    return dict(viewmatrix=np.ascontiguousarray(V.T).ravel(), projmatrix=np.ascontiguousarray(full.T).ravel(),
                projmatrix_raw=np.ascontiguousarray(Pm.T).ravel(), campos=campos, tanfovx=tanx, tanfovy=tany)
