"""Trajectory evaluation for the tracker's TUM output.

The reference writes `tracking_pose_tum.txt` rows `ts tx ty tz qx qy qz qw` with the WORLD->CAMERA pose
(utils/tracker.py:258-265 of the reference writes viewpoint.T / viewpoint.R as they are) and ships no
evaluator; this module is the missing piece needed to state the BASELINE.json gates:

  * `compare(est, ref)`        frame-by-frame difference of two trajectories with NO alignment — the parity
                               gate between this repo and the reference (1 mm / 0.05 deg);
  * `ate(est, gt)`             absolute trajectory error against ground truth after a rigid (Horn / Kabsch,
                               no scale) alignment of the camera centres, plus the rotation error of the
                               aligned poses — the usual TUM-benchmark numbers.

    python -m gsevt.ate estimated_tum.txt ground_truth_tum.txt [--max-dt 0.01] [--no-align]
"""
import argparse
import json

import numpy as np


def load_tum(path):
    """-> (ts[n], T[n,3], q_xyzw[n,4]) from a TUM trajectory file ('#' comments allowed)."""
    rows = np.loadtxt(path, ndmin=2, comments="#")
    if rows.size == 0:
        return np.zeros(0), np.zeros((0, 3)), np.zeros((0, 4))
    if rows.shape[1] != 8:
        raise ValueError(f"{path}: expected 8 columns (ts tx ty tz qx qy qz qw), got {rows.shape[1]}")
    return rows[:, 0], rows[:, 1:4], rows[:, 4:8]


def quat_to_matrix(q):
    """xyzw quaternions [n,4] -> rotation matrices [n,3,3] (normalised first)."""
    q = np.asarray(q, np.float64).reshape(-1, 4)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotation_angle_deg(R):
    """Geodesic angle of rotation matrices [n,3,3] in degrees."""
    c = (np.trace(R, axis1=-2, axis2=-1) - 1.0) / 2.0
    sk = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    return np.degrees(np.arctan2(0.5 * np.linalg.norm(sk, axis=-1), c))   # accurate near 0, unlike arccos


def associate(ts_a, ts_b, max_dt=0.01):
    """Greedy nearest-timestamp matching (each stamp used once) -> index arrays (ia, ib)."""
    ts_a, ts_b = np.asarray(ts_a, np.float64), np.asarray(ts_b, np.float64)
    if ts_a.size == 0 or ts_b.size == 0:
        return np.zeros(0, int), np.zeros(0, int)
    cand = []
    order_b = np.argsort(ts_b)
    sb = ts_b[order_b]
    for i, t in enumerate(ts_a):
        k = np.searchsorted(sb, t)
        for kk in (k - 1, k):
            if 0 <= kk < sb.size and abs(sb[kk] - t) <= max_dt:
                cand.append((abs(sb[kk] - t), i, int(order_b[kk])))
    cand.sort()
    used_a, used_b, ia, ib = set(), set(), [], []
    for _, i, j in cand:
        if i in used_a or j in used_b:
            continue
        used_a.add(i); used_b.add(j); ia.append(i); ib.append(j)
    o = np.argsort(ia)
    return np.asarray(ia, int)[o], np.asarray(ib, int)[o]


def camera_centres(T_wc, R_wc):
    """World->camera (R, T) -> camera centres in the world, c = -R^T T."""
    return -np.einsum("nji,nj->ni", R_wc, T_wc)


def horn_align(src, dst):
    """Rigid transform (R, t) minimising sum |R src_i + t - dst_i|^2 (Kabsch with the reflection guard)."""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    mu_s, mu_d = src.mean(0), dst.mean(0)
    Hm = (src - mu_s).T @ (dst - mu_d)
    U, _, Vt = np.linalg.svd(Hm)
    S = np.eye(3)
    if np.linalg.det(Vt.T @ U.T) < 0:
        S[2, 2] = -1
    R = Vt.T @ S @ U.T
    return R, mu_d - R @ mu_s


def compare(est, ref, max_dt=0.01):
    """Unaligned frame-by-frame difference of two (ts, T, q) trajectories in the SAME frame."""
    ia, ib = associate(est[0], ref[0], max_dt)
    if ia.size == 0:
        raise ValueError("no associated poses")
    dT = np.linalg.norm(est[1][ia] - ref[1][ib], axis=1)
    Ra, Rb = quat_to_matrix(est[2][ia]), quat_to_matrix(ref[2][ib])
    ang = rotation_angle_deg(np.einsum("nij,nkj->nik", Ra, Rb))
    return {"pairs": int(ia.size), "trans_max_m": float(dT.max()), "trans_rmse_m": float(np.sqrt(np.mean(dT ** 2))),
            "rot_max_deg": float(ang.max()), "rot_mean_deg": float(ang.mean()),
            "trans_per_frame_m": dT.tolist(), "rot_per_frame_deg": ang.tolist()}


def ate(est, gt, max_dt=0.01, align=True):
    """Absolute trajectory error of `est` against `gt` (both world->camera TUM trajectories)."""
    ia, ib = associate(est[0], gt[0], max_dt)
    if ia.size == 0:
        raise ValueError("no associated poses")
    Re, Rg = quat_to_matrix(est[2][ia]), quat_to_matrix(gt[2][ib])
    ce, cg = camera_centres(est[1][ia], Re), camera_centres(gt[1][ib], Rg)
    if align and ia.size >= 3:
        Ra, ta = horn_align(ce, cg)
    else:
        Ra, ta = np.eye(3), np.zeros(3)
    err = np.linalg.norm(ce @ Ra.T + ta - cg, axis=1)
    # camera-to-world rotations R^T; aligned estimate Ra R_e^T against R_g^T
    rel = np.einsum("ij,nkj,nkl->nil", Ra, Re, Rg)
    ang = rotation_angle_deg(rel)
    return {"pairs": int(ia.size), "aligned": bool(align and ia.size >= 3),
            "ate_rmse_m": float(np.sqrt(np.mean(err ** 2))), "ate_mean_m": float(err.mean()),
            "ate_median_m": float(np.median(err)), "ate_max_m": float(err.max()),
            "rot_rmse_deg": float(np.sqrt(np.mean(ang ** 2))), "rot_max_deg": float(ang.max())}


def write_tum(path, ts, T, q_xyzw):
    with open(path, "w") as f:
        for t, p, q in zip(ts, T, q_xyzw):
            f.write(f"{t} {p[0]} {p[1]} {p[2]} {q[0]} {q[1]} {q[2]} {q[3]}\n")


def matrix_to_quat(R):
    """Rotation matrix (3,3) -> xyzw quaternion with w >= 0 (Shepperd's method)."""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q if q[3] >= 0 else -q


def main(argv=None):
    ap = argparse.ArgumentParser(description="ATE / rotation error of a TUM trajectory against a reference trajectory")
    ap.add_argument("estimate")
    ap.add_argument("reference")
    ap.add_argument("--max-dt", type=float, default=0.01)
    ap.add_argument("--no-align", action="store_true", help="frame-by-frame difference without rigid alignment (parity mode)")
    a = ap.parse_args(argv)
    est, ref = load_tum(a.estimate), load_tum(a.reference)
    out = compare(est, ref, a.max_dt) if a.no_align else ate(est, ref, a.max_dt)
    out.pop("trans_per_frame_m", None)
    out.pop("rot_per_frame_deg", None)
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
