"""Multi-hypothesis initial-pose search, sharded across GPUs (BASELINE.json configs[3]).

The reference has no multi-GPU path; what shards naturally on this problem is the
set of independent pose / velocity hypotheses: each is its own TrackingEngine against a read-only map,
so hypothesis h runs on rank h mod world with NO data-path collective.  The only exchange is the final
gather of (hypothesis id, loss, state) — 20 floats per hypothesis — to pick the winner, done with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import math

import numpy as np


def assign(n_hypotheses, world):
    """Hypothesis ids per rank: h -> rank h mod world."""
    return [list(range(r, n_hypotheses, world)) for r in range(world)]


def perturb(R, T, angular_vel, linear_vel, h, sigma_t=0.05, sigma_deg=1.0, vel_frac=0.2, seed=7):
    """Hypothesis h: T + N(0, sigma_t^2) per axis, rotation by N(0, sigma_deg^2) degrees about a random
    axis (left-multiplied, world->camera convention), velocities scaled by 1 + vel_frac*U(-1,1) per axis.
    h == 0 is drawn like any other; results depend only on (seed, h), never on the rank layout."""
    rng = np.random.default_rng(seed + int(h))
    R = np.asarray(R, np.float64).reshape(3, 3)
    T = np.asarray(T, np.float64).reshape(3)
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = math.radians(sigma_deg) * rng.normal()
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    dR = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
    w = np.asarray(angular_vel, np.float64) * (1 + vel_frac * rng.uniform(-1, 1, 3))
    v = np.asarray(linear_vel, np.float64) * (1 + vel_frac * rng.uniform(-1, 1, 3))
    return (dR @ R).astype(np.float32), (T + rng.normal(0, sigma_t, 3)).astype(np.float32), w.astype(np.float32), v.astype(np.float32)


def pack_result(h, loss, iterations, R, T, angular_vel, linear_vel):
    """One row of the result table: [h, loss, iterations, R(9), T(3), w(3), v(3)] as float64."""
    return np.concatenate([[float(h), float(loss), float(iterations)], np.asarray(R, np.float64).ravel(), np.asarray(T, np.float64).ravel(),
                           np.asarray(angular_vel, np.float64).ravel(), np.asarray(linear_vel, np.float64).ravel()])


ROW = 21


def gather_results(local_rows, n_hypotheses, dist=None, device="cpu"):
    """All ranks end with the full (n_hypotheses, ROW) table, ordered by hypothesis id.  Each rank fills
    its own rows of a zero table and the tables are summed — one all-reduce of n*21 doubles at the very
    end of the search, nothing during optimisation."""
    import torch
    table = torch.zeros((n_hypotheses, ROW), dtype=torch.float64)
    for row in local_rows:
        table[int(row[0])] = torch.from_numpy(np.asarray(row, np.float64))
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        t = table.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        table = t.cpu()
    return table.numpy()


def best(table):
    """(hypothesis id, loss, row) of the lowest final loss; ties go to the lowest id (deterministic)."""
    losses = table[:, 1]
    h = int(np.flatnonzero(losses == losses.min())[0])
    return h, float(losses[h]), table[h]


def search(make_engine, state, n_hypotheses, run_one, dist=None, device="cpu"):
    """Runs this rank's share of the hypotheses and gathers the table on every rank.
    make_engine() -> engine;  run_one(engine, (R, T, w, v)) -> (loss, iterations, R, T, w, v)."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    eng = make_engine()
    rows = []
    for h in assign(n_hypotheses, world)[rank]:
        loss, iters, R, T, w, v = run_one(eng, perturb(*state, h))
        rows.append(pack_result(h, loss, iters, R, T, w, v))
    return gather_results(rows, n_hypotheses, dist, device)
