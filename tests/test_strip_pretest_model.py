"""CPU model test of the screen-tile split's strip pre-test (csrc/preprocess.cu: strip_may_touch).

The CUDA test runs on 20 bytes per Gaussian — position and the largest eigenvalue of the 3-D covariance — and must be
CONSERVATIVE: whenever the exact projection (the oracle's restatement of preprocessCUDA) gives a Gaussian a tile rect that
reaches a strip of tile rows, the pre-test must say "may touch".  This file restates the device function in numpy, operation
for operation, and checks that implication against the oracle over maps with near-plane, anisotropic and blown-up
Gaussians, perturbed poses, both render views and every strip of 2-, 5- and 8-way splits; it also reports how tight the
bound is (survivors per Gaussian that really touches).  The GPU tests check the same thing end to end (the strips' lists
must equal the unsplit engine's: tests/test_gpu_tilesplit.py)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-evt_b200"), ROOT]

from gsevt import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def strip_may_touch(view, tanfovx, tanfovy, focal_x, focal_y, H, xyz, smax2, row0_px, row1_px):
    """numpy float32 restatement of the device function (same expression order; view is the column-major 4x4)."""
    f = np.float32
    px, py, pz = (xyz[:, i].astype(f) for i in range(3))
    v = view.astype(f)
    tz = ((v[10] * pz + (v[2] * px + v[6] * py).astype(f)).astype(f) + v[14]).astype(f)
    ok = tz > f(0.2)
    tzs = np.where(ok, tz, f(1.0)).astype(f)
    tx = (v[0] * px + v[4] * py + v[8] * pz + v[12]).astype(f)
    ty = (v[1] * px + v[5] * py + v[9] * pz + v[13]).astype(f)
    iz = (f(1.0) / tzs).astype(f)
    limx, limy = f(1.3) * f(tanfovx), f(1.3) * f(tanfovy)
    cx = np.clip(tx * iz, -limx, limx).astype(f)
    cy = np.clip(ty * iz, -limy, limy).astype(f)
    j00, j11 = (f(focal_x) * iz).astype(f), (f(focal_y) * iz).astype(f)
    j02, j12 = (-cx * j00).astype(f), (-cy * j11).astype(f)
    ja, jc, jb = (j00 * j00 + j02 * j02).astype(f), (j11 * j11 + j12 * j12).astype(f), (j02 * j12).astype(f)
    mid, dif = (f(0.5) * (ja + jc)).astype(f), (f(0.5) * (ja - jc)).astype(f)
    lamJ = (mid + np.sqrt(dif * dif + jb * jb).astype(f)).astype(f)
    r = (f(3.0) * np.sqrt(lamJ * smax2.astype(f) * f(1.001) + f(0.3)).astype(f) + f(2.0)).astype(f)
    my = ((ty * iz / f(tanfovy) + f(1.0)) * (f(0.5) * f(H)) - f(0.5)).astype(f)
    return ok & ~(my + r + f(15.0) < f(row0_px)) & ~(my - r >= f(row1_px))


def exact_rows(views_entry, act, W, H):
    """Tile rows [y0, y1) of every Gaussian's rect in one view, from the oracle's preprocess (0, 0 when not visible)."""
    v = views_entry
    sc = orc.Scene(v["W"], v["H"], v["tanfovx"], v["tanfovy"], np.zeros(3, np.float32), act["xyz"], act["opacities"], v["viewmatrix"],
                   v["projmatrix"], v["campos"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"], sh_degree=3,
                   projmatrix_raw=v["projmatrix_raw"], vel=v["vel"], vel_inv=v["vel_inv"], delta_time=v["delta_time"])
    P = act["xyz"].shape[0]
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
             clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
    orc.lib().orc_preprocess(C.byref(sc.s), *(orc._p(o[k]) for k in ("radii", "means2D", "depths", "cov3D", "conic_opacity", "rgb",
                                                                      "clamped", "tiles_touched")))
    gy = (H + 15) // 16
    my, r = o["means2D"][:, 1], o["radii"].astype(np.float32)
    y0 = np.clip(((my - r) * np.float32(0.0625)).astype(np.int64), 0, gy)        # C truncation, as tile_rect does
    y1 = np.clip(((my + r + np.float32(16.0) - np.float32(1.0)) * np.float32(0.0625)).astype(np.int64), 0, gy)
    vis = o["tiles_touched"] > 0
    return np.where(vis, y0, 0), np.where(vis, y1, 0), vis


@pytest.mark.parametrize("scale_mult,seed,dT", [(1.0, 0, 0.0), (4.0, 1, 0.05), (12.0, 2, 0.3), (0.3, 3, 0.02)])
def test_pretest_never_rejects_a_gaussian_that_reaches_the_strip(scale_mult, seed, dT):
    W, H, P = 640, 480, 30000
    D = synth.DESK
    raw = synth.synth_map(P, seed=seed)
    raw["scaling"] = (raw["scaling"] + np.float32(math.log(scale_mult))).astype(np.float32)
    raw["scaling"][::7, 0] += 1.5                                     # strongly anisotropic splats
    act = synth.activate(raw)
    smax2 = (act["scales"].max(axis=1).astype(np.float32) ** 2 * np.float32(1.0001)).astype(np.float32)   # pack_map_kernel
    rng = np.random.default_rng(seed)
    R = np.asarray(D["R"], np.float32).reshape(3, 3)
    T = (np.asarray(D["T"], np.float32) + rng.normal(0, dT, 3).astype(np.float32)).astype(np.float32)
    w = np.asarray(D["angular_vel"], np.float32) * 20
    v = np.asarray(D["linear_vel"], np.float32) * 5
    views = orc.view_setup(R, T, w, v, 0.05, W, H, D["fx"], D["fy"], 0)
    gy = (H + 15) // 16
    touching = surviving = 0
    for ve in views:
        y0, y1, vis = exact_rows(ve, act, W, H)
        fx = W / (2.0 * ve["tanfovx"])
        fy = H / (2.0 * ve["tanfovy"])
        for n in (2, 5, 8):
            bounds = np.linspace(0, gy, n + 1).round().astype(int)
            for s0, s1 in zip(bounds[:-1], bounds[1:]):
                reach = vis & (np.minimum(y1, s1) > np.maximum(y0, s0))
                may = strip_may_touch(ve["viewmatrix"], ve["tanfovx"], ve["tanfovy"], fx, fy, H, act["xyz"], smax2, s0 * 16, s1 * 16)
                missed = np.flatnonzero(reach & ~may)
                assert missed.size == 0, (n, s0, s1, missed[:5], y0[missed[:5]], y1[missed[:5]])
                touching += int(reach.sum())
                surviving += int(may.sum())
    assert touching > 0
    # the bound is loose by design (isotropic radius from the largest 3-D eigenvalue), but not uselessly so
    print(f"scale x{scale_mult}: {surviving / touching:.2f} survivors per Gaussian that reaches its strip")
    assert surviving / touching < (4.0 if scale_mult <= 4 else 8.0)
