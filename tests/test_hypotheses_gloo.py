"""Multi-GPU host logic on CPU: hypothesis sharding + the final result gather over a world_size-2 gloo
group (the GPU run uses the same code over NCCL).  No data-path collective exists on this path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-evt_b200"))

from gsevt import hypotheses as hyp  # noqa: E402
from gsevt import synth  # noqa: E402


def test_assignment_covers_every_hypothesis_once():
    for n, world in ((64, 8), (64, 2), (7, 4), (3, 8), (0, 2)):
        parts = hyp.assign(n, world)
        assert len(parts) == world and sorted(h for p in parts for h in p) == list(range(n))
        assert all(h % world == r for r, p in enumerate(parts) for h in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_perturbation_is_seeded_and_layout_independent():
    D = synth.DESK
    a = hyp.perturb(D["R"], D["T"], D["angular_vel"], D["linear_vel"], 5)
    b = hyp.perturb(D["R"], D["T"], D["angular_vel"], D["linear_vel"], 5)
    c = hyp.perturb(D["R"], D["T"], D["angular_vel"], D["linear_vel"], 6)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not np.array_equal(a[1], c[1])
    R = a[0].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6
    assert np.abs(a[1] - np.array(D["T"])).max() < 0.3
    assert np.all(np.abs(a[3] / np.array(D["linear_vel"]) - 1) <= 0.2 + 1e-6)
    # bench.py draws its hypotheses with the same generator sequence
    sys.path.insert(0, ROOT)
    import bench
    d = dict(R=D["R"], T=D["T"], angular_vel=D["angular_vel"], linear_vel=D["linear_vel"])
    assert all(np.array_equal(x, y) for x, y in zip(bench.perturbed_state(d, 5), a))


def _worker(rank, world, port, n_hyp, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = synth.DESK
    state = (D["R"], D["T"], D["angular_vel"], D["linear_vel"])

    def run_one(_eng, st):  # stand-in for a converged TrackingEngine: loss = distance of T from the truth
        R, T, w, v = st
        return float(np.linalg.norm(T - np.array(D["T"], np.float32))), 10 + rank, R, T, w, v

    table = hyp.search(lambda: None, state, n_hyp, run_one, dist=dist)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_hyp", [8, 5])
def test_sharded_search_over_gloo(n_hyp):
    world, port = 2, 29600 + n_hyp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_hyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds the same full table; rows come from the rank that owns the hypothesis
    assert np.array_equal(got[0], got[1]) and got[0].shape == (n_hyp, hyp.ROW)
    assert np.array_equal(got[0][:, 0], np.arange(n_hyp)) and np.array_equal(got[0][:, 2], 10 + np.arange(n_hyp) % world)
    D = synth.DESK
    single = hyp.search(lambda: None, (D["R"], D["T"], D["angular_vel"], D["linear_vel"]), n_hyp,
                        lambda e, st: (float(np.linalg.norm(st[1] - np.array(D["T"], np.float32))), 0, *st))
    assert np.allclose(single[:, 1], got[0][:, 1]) and np.allclose(single[:, 3:], got[0][:, 3:])   # same answer as one rank
    h, loss, row = hyp.best(got[0])
    assert loss == got[0][:, 1].min() and row[0] == h
