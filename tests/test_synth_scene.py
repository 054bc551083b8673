"""CPU tests of the synthetic-scene generators behind the sequence tests and the bench (gsevt.synth): the trackable
scene's invariants, the contrast-threshold event model, the closed-loop trajectory, and the event text writer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-evt_b200"))

from gsevt import synth  # noqa: E402


def _cam(xyz):
    D = synth.DESK
    R0, T0 = np.asarray(D["R"], np.float64).reshape(3, 3), np.asarray(D["T"], np.float64)
    return xyz.astype(np.float64) @ R0.T + T0


def test_default_map_is_unchanged_by_the_new_options():
    a = synth.synth_map(5000, seed=3)
    b = synth.synth_map(5000, seed=3, structure=0, fine_opacity_shift=-9.0)     # ignored without structure
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert (np.abs(_cam(a["xyz"])[:, 2]) < 0.5).sum() > 0                         # the bench map keeps its near-plane splats


def test_structured_map_invariants():
    P, n_big = 20000, 400
    base = synth.synth_map(P, seed=1)
    m = synth.synth_map(P, seed=1, structure=n_big, fine_opacity_shift=-3.0)
    assert all(m[k].shape == base[k].shape and m[k].dtype == np.float32 for k in base)
    z = _cam(m["xyz"])[:, 2]
    assert not np.any((z > -1.0 + 1e-4) & (z < 1.0 - 1e-4))                     # nothing within a metre of the camera plane
    big = np.flatnonzero(np.all(m["f_rest"].reshape(P, -1) == 0, axis=1))       # only the structure splats are view-independent
    assert len(big) == n_big and big.max() > P // 2 and big.min() < P // 2        # spread over the index range
    assert np.all(z[big] >= 3.0 - 1e-4) and np.all(z[big] <= 6.0 + 1e-4)
    assert m["opacity"][big].mean() > 2.0 and m["scaling"][big].mean() > m["scaling"].mean() + 1.0
    fine = np.setdiff1d(np.arange(P), big)
    assert abs((m["opacity"][fine] - base["opacity"][fine]).mean() + 3.0) < 1e-5
    m2 = synth.synth_map(P, seed=1, structure=n_big, fine_opacity_shift=-3.0)
    assert all(np.array_equal(m[k], m2[k]) for k in m)                            # seeded


def test_threshold_events_count_sign_and_determinism():
    rng = np.random.default_rng(0)
    H, W = 48, 64
    dI = rng.normal(size=(H, W)) * (rng.uniform(size=(H, W)) > 0.5)
    K = np.array([60.0, 0, W / 2, 0, 60.0, H / 2, 0, 0, 1]).reshape(3, 3)
    for n in (1, 1000, 5000):
        ev = synth.threshold_events(dI, n, 100, 50099, K, [0, 0, 0, 0, 0], seed=4)
        assert ev.shape == (n, 4) and ev.dtype == np.int64
        assert (n == 1 or ev[0, 0] == 100) and ev[-1, 0] == 50099 and np.all(np.diff(ev[:, 0]) >= 0)
        assert ev[:, 1].min() >= 0 and ev[:, 1].max() < W and ev[:, 2].min() >= 0 and ev[:, 2].max() < H
        # no distortion: the event sits on its pixel, with the sign of the intensity change there
        assert np.array_equal(ev[:, 3], (dI[ev[:, 2], ev[:, 1]] > 0).astype(np.int64))
        assert np.array_equal(ev, synth.threshold_events(dI, n, 100, 50099, K, [0, 0, 0, 0, 0], seed=4))
    # events per pixel follow |dI| / C up to one threshold crossing
    ev = synth.threshold_events(dI, 5000, 0, 49999, K, [0, 0, 0, 0, 0], seed=5)
    cnt = np.zeros((H, W), np.int64)
    np.add.at(cnt, (ev[:, 2], ev[:, 1]), 1)
    a = np.abs(dI)
    C = a.sum() / 5000
    assert np.all(cnt[a == 0] == 0) and np.abs(cnt - a / C).max() < 2.5
    # a frame without any change falls back to the proportional sampler instead of dividing by zero
    assert synth.threshold_events(np.zeros((H, W)), 10, 0, 9, K, [0, 0, 0, 0, 0], seed=1).shape == (10, 4)


def test_orbit_trajectory_is_closed_and_keeps_the_speed():
    D = synth.DESK
    n, dtau, period = 240, 0.05, 12.0
    gt = synth.ground_truth_trajectory(n, dtau, mode="orbit", orbit_period=period)
    assert len(gt) == n and abs(gt[0][3] - dtau / 2) < 1e-12
    speed = np.array([np.linalg.norm(g[1]) for g in gt])
    assert np.allclose(speed, np.linalg.norm(D["linear_vel"]), rtol=1e-9)         # never stands still
    centres = np.array([-(g[0][:3, :3].T @ g[0][:3, 3]) for g in gt])
    radius = np.linalg.norm(D["linear_vel"]) * period / (2 * np.pi)
    assert np.linalg.norm(centres[-1] - centres[0]) < 0.05 * radius               # one period later: back at the start
    assert np.ptp(centres, axis=0).max() < 2.2 * radius
    R = gt[100][0][:3, :3]
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6      # the yaml rotation itself is orthonormal to 1e-8
    drift = synth.ground_truth_trajectory(40, dtau)                                # the old mode is untouched: a straight drift
    c = np.array([-(g[0][:3, :3].T @ g[0][:3, 3]) for g in drift])
    assert np.linalg.norm(c[-1] - c[0]) > 0.8 * np.linalg.norm(D["linear_vel"]) * 39 * dtau


def test_events_text_round_trip(tmp_path):
    ev = synth.random_events(2000, 64, 48, 0, 49999, seed=2)
    p = str(tmp_path / "e.txt")
    synth.write_events_txt(p, ev)
    assert np.array_equal(np.loadtxt(p, dtype=np.int64), ev)
    first = open(p).readline()
    assert first == "%d %d %d %d\n" % tuple(ev[0])
