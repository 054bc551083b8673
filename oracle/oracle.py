"""TEST INFRASTRUCTURE — Python face of the CPU oracle (oracle/liboracle.so, oracle/gsevt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product never does.  It restates, on the CPU:
  * the rasteriser forward / backward + pose chain (C, see gsevt_oracle.c for the reference citations);
  * the two-view intensity-change loss and its pixel gradient (numpy; frame.py:86-92, tracker.py:93-103);
  * the SE(3) algebra of utils/pose.py and utils/render_camera/camera.py:100-155 (numpy float32);
so that one full tracking-iteration evaluation (loss + 12 pose/velocity gradients) can be computed
without a GPU.  Parity status: pinned against the live reference in tests/test_parity_reference.py
(GPU box) and against tests/golden/*.npz generated from it.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcScene(C.Structure):
    _fields_ = [("P", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float), ("delta_time", C.c_float),
                ("bg", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
                ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("projmatrix_raw", C.c_void_p), ("campos", C.c_void_p),
                ("vel", C.c_void_p), ("vel_inv", C.c_void_p)]


def build():
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_bin.restype = C.c_int64
    return _LIB


def _p(a):
    """64-bit safe pointer argument (a bare int would be truncated to a C int by ctypes)."""
    return C.c_void_p(None) if a is None else C.c_void_p(a.ctypes.data)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Scene:
    """Holds float32 copies of all inputs and the OrcScene struct pointing at them."""

    def __init__(self, W, H, tanfovx, tanfovy, bg, means3D, opacities, viewmatrix, projmatrix, campos, shs=None,
                 colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0,
                 projmatrix_raw=None, vel=None, vel_inv=None, delta_time=0.0):
        eye = np.eye(4, dtype=np.float32).ravel()
        self.a = dict(bg=_f32(bg), means3D=_f32(means3D), shs=_f32(shs), colors_precomp=_f32(colors_precomp),
                      opacities=_f32(np.asarray(opacities).reshape(-1)), scales=_f32(scales), rotations=_f32(rotations),
                      cov3D_precomp=_f32(cov3D_precomp), viewmatrix=_f32(viewmatrix).ravel(), projmatrix=_f32(projmatrix).ravel(),
                      projmatrix_raw=_f32(projmatrix_raw if projmatrix_raw is not None else eye).ravel(), campos=_f32(campos),
                      vel=_f32(vel if vel is not None else eye).ravel(), vel_inv=_f32(vel_inv if vel_inv is not None else eye).ravel())
        s = OrcScene()
        s.P = self.a["means3D"].shape[0]
        s.D = int(sh_degree)
        s.M = 0 if shs is None else int(self.a["shs"].shape[1])
        s.W, s.H = int(W), int(H)
        s.tanfovx, s.tanfovy = float(tanfovx), float(tanfovy)
        s.scale_modifier, s.delta_time = float(scale_modifier), float(delta_time)
        for k, v in self.a.items():
            setattr(s, k, None if v is None else v.ctypes.data)
        self.s = s
        self.P, self.W, self.H = s.P, s.W, s.H


def forward(sc):
    """Full forward; returns a dict of every intermediate the parity tests compare."""
    L = lib()
    P, W, H = sc.P, sc.W, sc.H
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
             rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
    L.orc_preprocess(C.byref(sc.s), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]),
                     _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["clamped"]), _p(o["tiles_touched"]))
    if sc.a["cov3D_precomp"] is not None:
        o["cov3D"] = sc.a["cov3D_precomp"].reshape(P, 6).copy()
    N = int(o["tiles_touched"].sum())
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o["keys"] = np.zeros(max(N, 1), np.uint64)
    o["point_list"] = np.zeros(max(N, 1), np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    n = L.orc_bin(P, W, H, _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["tiles_touched"]), _p(o["keys"]),
                  _p(o["point_list"]), _p(o["ranges"]))
    assert n == N, (n, N)
    o["keys"], o["point_list"], o["num_rendered"] = o["keys"][:N], o["point_list"][:N], N
    o["color"] = np.zeros((3, H, W), np.float32)
    o["depth"] = np.zeros((1, H, W), np.float32)
    o["opacity"] = np.zeros((1, H, W), np.float32)
    o["final_T"] = np.zeros((H, W), np.float32)
    o["n_contrib"] = np.zeros((H, W), np.uint32)
    o["n_touched"] = np.zeros(P, np.int32)
    colors = o["rgb"]
    L.orc_render_fwd(W, H, _p(o["ranges"]), _p(o["point_list"] if N else np.zeros(1, np.uint32)), _p(o["means2D"]), _p(colors),
                     _p(o["conic_opacity"]), _p(o["depths"]), _p(sc.a["bg"]), _p(o["color"]), _p(o["depth"]), _p(o["opacity"]),
                     _p(o["final_T"]), _p(o["n_contrib"]), _p(o["n_touched"]))
    return o


def backward(sc, fw, dL_dcolor, dL_ddepth=None):
    """Backward from pixel gradients; returns per-Gaussian gradients and the 12 pose sums
    [rho, theta, v, w] (column sums of dL_dtau / dL_dvel, as dgr/.../__init__.py:163-169)."""
    L = lib()
    P, W, H = sc.P, sc.W, sc.H
    g = dict(dL_dmean2D=np.zeros((P, 2), np.float32), dL_dconic=np.zeros((P, 3), np.float32), dL_dopacity=np.zeros(P, np.float32),
             dL_dcolors=np.zeros((P, 3), np.float32), dL_ddepths=np.zeros(P, np.float32), dL_dtau=np.zeros((P, 6), np.float32),
             dL_dvel=np.zeros((P, 6), np.float32), dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32))
    dc = _f32(dL_dcolor).reshape(3, H, W)
    dd = None if dL_ddepth is None else _f32(dL_ddepth).reshape(H, W)
    pl = fw["point_list"] if fw["num_rendered"] else np.zeros(1, np.uint32)
    L.orc_render_bwd(P, W, H, _p(fw["ranges"]), _p(pl), _p(fw["means2D"]), _p(fw["rgb"]), _p(fw["conic_opacity"]),
                     _p(fw["depths"]), _p(sc.a["bg"]), _p(fw["final_T"]), _p(fw["n_contrib"]), _p(dc), _p(dd),
                     _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_ddepths"]))
    cov = np.ascontiguousarray(fw["cov3D"], np.float32)
    L.orc_geom_bwd(C.byref(sc.s), _p(fw["radii"]), _p(cov), _p(fw["clamped"]), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]),
                   _p(g["dL_dcolors"]), _p(g["dL_ddepths"]), _p(g["dL_dtau"]), _p(g["dL_dvel"]), _p(g["dL_dmeans3D"]),
                   _p(g["dL_dcov3D"]))
    tau = g["dL_dtau"].astype(np.float64).sum(0)
    vel = g["dL_dvel"].astype(np.float64).sum(0)
    g["pose_grads"] = np.concatenate([tau, vel]).astype(np.float32)
    return g


def mark_visible(means3D, viewmatrix):
    m = _f32(means3D)
    out = np.zeros(m.shape[0], np.uint8)
    lib().orc_mark_visible(m.shape[0], _p(m), _p(_f32(viewmatrix).ravel()), _p(out))
    return out.astype(bool)


# ---- pose algebra (utils/pose.py:13-91), float32 like torch ---------------------------------------
def skew(x):
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]], np.float32)


def SE3_exp(xi):
    xi = np.asarray(xi, np.float32)
    rho, th = xi[:3], xi[3:]
    W = skew(th)
    W2 = (W @ W).astype(np.float32)
    angle = np.float32(np.sqrt(np.sum(th * th, dtype=np.float32)))
    I = np.eye(3, dtype=np.float32)
    if angle < 1e-5:
        R = I + W + np.float32(0.5) * W2
        V = I + np.float32(0.5) * W + np.float32(1.0 / 6.0) * W2
    else:
        sn, cs = np.float32(math.sin(angle)), np.float32(math.cos(angle))
        R = I + (sn / angle) * W + ((np.float32(1) - cs) / (angle * angle)) * W2
        V = I + W * ((np.float32(1) - cs) / (angle * angle)) + W2 * ((angle - sn) / (angle * angle * angle))
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R
    T[:3, 3] = V.astype(np.float32) @ rho
    return T


def se3_inv(T):
    out = np.eye(4, dtype=np.float32)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -(T[:3, :3].T @ T[:3, 3])
    return out


def projection(znear, zfar, fovX, fovY):
    """getProjectionMatrix, graphics_utils.py:49-69."""
    tY, tX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tY * znear, tX * znear
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def view_setup(R, T, ang_vel, lin_vel, delta_tau, W, H, fx, fy, level, znear=0.01, zfar=100.0):
    """The two GaussianRasterizationSettings of render2 (gaussian_renderer/__init__.py:266-275,313-339)
    for pyramid `level`, as plain dicts of float32 arrays."""
    s = 0.5 ** level
    Wl, Hl = int(W * s), int(H * s)
    fovx, fovy = 2 * math.atan(Wl / (2 * (fx * s))), 2 * math.atan(Hl / (2 * (fy * s)))
    half = np.float32(delta_tau / 2)
    rot, tr = np.asarray(ang_vel, np.float32) * half, np.asarray(lin_vel, np.float32) * half
    cur = np.eye(4, dtype=np.float32)
    cur[:3, :3], cur[:3, 3] = np.asarray(R, np.float32).reshape(3, 3), np.asarray(T, np.float32).reshape(3)
    Pm = projection(znear, zfar, fovx, fovy)
    views = []
    for sign in (-1.0, 1.0):
        Tvel = SE3_exp(np.concatenate([np.float32(sign) * tr, np.float32(sign) * rot]))
        Tk = (Tvel @ cur).astype(np.float32)
        views.append(dict(W=Wl, H=Hl, tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5),
                          viewmatrix=np.ascontiguousarray(Tk.T).ravel(), projmatrix=np.ascontiguousarray((Pm @ Tk).astype(np.float32).T).ravel(),
                          projmatrix_raw=np.ascontiguousarray(Pm.T).ravel(), campos=(-(Tk[:3, :3].T @ Tk[:3, 3])).astype(np.float32),
                          vel=np.ascontiguousarray(Tvel.T).ravel(), vel_inv=np.ascontiguousarray(se3_inv(Tvel).T).ravel(),
                          delta_time=float(sign * half)))
    return views


GRAY = np.array([0.2989, 0.5870, 0.1140], np.float32)


def loss_and_pixel_grad(gray_last, gray_next, E, signed=True):
    """frame.py:86-92 + tracker.py:93-103 and the closed-form dL/d(delta) of SURVEY.md 8(a) a15."""
    d = (gray_next - gray_last).astype(np.float32)
    n = np.float32(np.sqrt(np.sum(d.astype(np.float64) ** 2)))
    u = d / n
    if signed:
        r = u - E
    else:
        r = np.abs(u) - np.abs(E)
    L = np.float32(np.sqrt(np.sum(r.astype(np.float64) ** 2)))
    g = r / L if signed else np.sign(u) * r / L
    dd = (g - u * np.sum(u.astype(np.float64) * g.astype(np.float64))) / n
    return float(L), dd.astype(np.float32)


def tracking_eval(act, R, T, ang_vel, lin_vel, delta_tau, W, H, fx, fy, level, E, signed=True, bg=(0, 0, 0), gray_override=None):
    """One full evaluation of the tracking objective on the CPU: two forwards, loss, two backwards.
    act: dict from gsevt.synth.activate().  Returns (loss, grads12 [rho,theta,v,w], aux).
    gray_override = (gray_last, gray_next): take the loss and its pixel gradient from these images instead of
    the oracle's own renders (used for the unsigned objective, whose gradient flips sign with the rounding
    noise of the render difference)."""
    views = view_setup(R, T, ang_vel, lin_vel, delta_tau, W, H, fx, fy, level)
    fws, scs, grays = [], [], []
    for v in views:
        sc = Scene(v["W"], v["H"], v["tanfovx"], v["tanfovy"], np.asarray(bg, np.float32), act["xyz"], act["opacities"],
                   v["viewmatrix"], v["projmatrix"], v["campos"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"],
                   sh_degree=3, projmatrix_raw=v["projmatrix_raw"], vel=v["vel"], vel_inv=v["vel_inv"], delta_time=v["delta_time"])
        fw = forward(sc)
        scs.append(sc)
        fws.append(fw)
        grays.append(np.tensordot(GRAY, fw["color"], axes=(0, 0)).astype(np.float32))
    lg = grays if gray_override is None else [np.asarray(x, np.float32).reshape(grays[0].shape) for x in gray_override]
    L, dd = loss_and_pixel_grad(lg[0], lg[1], np.asarray(E, np.float32).reshape(grays[0].shape), signed)
    total = np.zeros(12, np.float64)
    for sgn, sc, fw in ((-1.0, scs[0], fws[0]), (1.0, scs[1], fws[1])):
        dcol = (GRAY[:, None, None] * (np.float32(sgn) * dd)[None]).astype(np.float32)
        total += backward(sc, fw, dcol)["pose_grads"].astype(np.float64)
    return L, total.astype(np.float32), dict(gray=grays, fw=fws, views=views)
