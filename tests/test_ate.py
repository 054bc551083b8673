"""Trajectory evaluator (gsevt.ate): known-answer checks on synthetic trajectories (CPU)."""
import numpy as np
import pytest

from gsevt import ate, synth


def _traj(n=40, seed=0):
    gt = synth.ground_truth_trajectory(n, dtau=0.05)
    ts = np.array([g[3] for g in gt])
    T = np.array([g[0][:3, 3] for g in gt])
    q = np.array([ate.matrix_to_quat(g[0][:3, :3]) for g in gt])
    return ts, T, q


def test_quaternion_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        q = ate.matrix_to_quat(Q)
        assert abs(np.linalg.norm(q) - 1) < 1e-12 and q[3] >= 0
        assert np.abs(ate.quat_to_matrix(q[None])[0] - Q).max() < 1e-12
    # agrees with scipy (what the tracker uses to write the TUM file)
    from scipy.spatial.transform import Rotation
    Rm = Rotation.from_rotvec([0.3, -0.2, 0.9]).as_matrix()
    qs = Rotation.from_matrix(Rm).as_quat()
    qs = qs if qs[3] >= 0 else -qs
    assert np.abs(ate.matrix_to_quat(Rm) - qs).max() < 1e-12


def test_identical_trajectories_have_zero_error(tmp_path):
    tr = _traj()
    c = ate.compare(tr, tr)
    a = ate.ate(tr, tr)
    assert c["pairs"] == 40 and c["trans_max_m"] == 0 and c["rot_max_deg"] < 1e-6
    assert a["ate_rmse_m"] < 1e-12 and a["rot_max_deg"] < 1e-6
    p = tmp_path / "t.txt"
    ate.write_tum(p, *tr)
    back = ate.load_tum(p)
    assert ate.compare(back, tr)["trans_max_m"] < 1e-12


def test_rigid_world_change_is_removed_by_alignment_but_seen_by_compare():
    ts, T, q = _traj()
    R = ate.quat_to_matrix(q)
    # move the world: x_w' = Rw x_w + tw  =>  world->camera pose (R Rw^T, T - R Rw^T tw)
    from scipy.spatial.transform import Rotation
    Rw, tw = Rotation.from_rotvec([0.1, 0.2, -0.3]).as_matrix(), np.array([0.5, -1.0, 2.0])
    R2 = np.einsum("nij,kj->nik", R, Rw)
    T2 = T - np.einsum("nij,j->ni", R2, tw)
    q2 = np.array([ate.matrix_to_quat(r) for r in R2])
    a = ate.ate((ts, T2, q2), (ts, T, q))
    assert a["aligned"] and a["ate_rmse_m"] < 1e-9 and a["rot_max_deg"] < 1e-5
    assert ate.compare((ts, T2, q2), (ts, T, q))["trans_max_m"] > 0.1


def test_known_offsets():
    ts, T, q = _traj()
    # 2 mm along camera x for every pose: the camera centre moves by exactly 2 mm (c = -R^T T)
    T2 = T + np.array([0.002, 0.0, 0.0])
    c = ate.compare((ts, T2, q), (ts, T, q))
    assert abs(c["trans_max_m"] - 0.002) < 1e-12 and c["rot_max_deg"] < 1e-6
    a = ate.ate((ts, T2, q), (ts, T, q), align=False)
    assert abs(a["ate_rmse_m"] - 0.002) < 1e-9
    # 0.5 deg about the camera z axis
    from scipy.spatial.transform import Rotation
    dR = Rotation.from_euler("z", 0.5, degrees=True).as_matrix()
    R = ate.quat_to_matrix(q)
    q3 = np.array([ate.matrix_to_quat(dR @ r) for r in R])
    c = ate.compare((ts, T, q3), (ts, T, q))
    assert abs(c["rot_max_deg"] - 0.5) < 1e-6


def test_association_tolerates_jitter_and_missing_frames():
    ts, T, q = _traj()
    keep = np.array([i for i in range(40) if i % 7 != 3])
    est = (ts[keep] + 0.002, T[keep], q[keep])
    c = ate.compare(est, (ts, T, q), max_dt=0.01)
    assert c["pairs"] == keep.size and c["trans_max_m"] == 0
    with pytest.raises(ValueError):
        ate.compare((ts + 10.0, T, q), (ts, T, q))


def test_cli(tmp_path, capsys):
    tr = _traj(12)
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    ate.write_tum(a, tr[0], tr[1] + 0.001, tr[2])
    ate.write_tum(b, *tr)
    out = ate.main([str(a), str(b), "--no-align"])
    assert abs(out["trans_max_m"] - np.sqrt(3) * 0.001) < 1e-9
    assert "trans_max_m" in capsys.readouterr().out
