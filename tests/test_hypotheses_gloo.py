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

    table = hyp.search(lambda: None, state, n_hyp, run_one, dist=dist, assignment="static")
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



def _ticket_worker(rank, world, port, n_hyp, q):
    import time
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = synth.DESK
    state = (D["R"], D["T"], D["angular_vel"], D["linear_vel"])

    def run_one(_eng, st):  # rank 0 is 20x slower per hypothesis: with tickets it must end up with far fewer of them
        time.sleep(0.10 if rank == 0 else 0.005)
        R, T, w, v = st
        return float(np.linalg.norm(T - np.array(D["T"], np.float32))), 100 + rank, R, T, w, v

    dist.barrier()
    table = hyp.search(lambda: None, state, n_hyp, run_one, dist=dist)        # dynamic is the default
    table2 = hyp.search(lambda: None, state, n_hyp, run_one, dist=dist)       # a second search draws from a fresh counter
    q.put((rank, table, table2))
    dist.barrier()
    dist.destroy_process_group()


def test_dynamic_assignment_over_gloo_balances_unequal_ranks():
    world, port, n_hyp = 2, 29641, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ticket_worker, args=(r, world, port, n_hyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (a, b) for r, a, b in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in (0, 1):
        t0, t1 = got[0][k], got[1][k]
        assert np.array_equal(t0, t1) and np.array_equal(t0[:, 0], np.arange(n_hyp))        # every id exactly once, same table everywhere
        owners = t0[:, 2] - 100
        assert set(owners) <= {0.0, 1.0} and (owners == 0).sum() >= 1
        assert (owners == 1).sum() > 2 * (owners == 0).sum(), owners                           # the fast rank drew most of the tickets
    single = hyp.search(lambda: None, (synth.DESK["R"], synth.DESK["T"], synth.DESK["angular_vel"], synth.DESK["linear_vel"]), n_hyp,
                        lambda e, st: (float(np.linalg.norm(st[1] - np.array(synth.DESK["T"], np.float32))), 0, *st))
    assert np.allclose(single[:, 1], got[0][0][:, 1])                                          # losses do not depend on who ran what


class _FakeEngine:
    """Stand-in with the TrackingEngine calls track_concurrently makes; a level converges after `need` iterations."""

    def __init__(self, log, name):
        self.log, self.name, self.stream = log, name, None
        self.state, self.level, self.done_iters, self.need, self.flag, self.paused_once = None, None, 0, 0, 1, False

    def set_state(self, R, T, w, v):
        self.state = [np.array(R), np.array(T, np.float64), np.array(w), np.array(v)]

    def begin_frame(self, dt, s, u):
        self.log.append((self.name, "frame"))

    def begin_level(self, level, opt_vel):
        self.level, self.done_iters, self.flag = level, 0, 0
        self.need = 5 + 7 * level + int(abs(self.state[1][0]) * 1000) % 11
        self.log.append((self.name, "level", level, bool(opt_vel)))

    def iterate(self, n):
        assert self.flag in (0,), "iterate on a finished / paused level"
        for _ in range(n):
            if self.flag:
                break
            self.done_iters += 1
            if self.done_iters == 3 and not self.paused_once and self.level == 0:
                self.flag, self.paused_once = 2, True        # one overflow pause per engine lifetime
            elif self.done_iters >= self.need:
                self.flag = 1
                self.state[1] = self.state[1] * 0.5

    def resume(self):
        assert self.flag == 2
        self.flag = 0
        self.log.append((self.name, "resume"))

    def poll_done(self):
        return self.flag

    def status(self):
        class S:
            pass
        s = S()
        s.optim_iter, s.last_loss = self.done_iters, float(np.linalg.norm(self.state[1]))
        return s

    def get_state(self):
        return tuple(self.state)


class _Ev:
    def __init__(self):
        self.polls = 0

    def query(self):           # "finished" on the second poll: exercises the not-ready branch
        self.polls += 1
        return self.polls >= 2


def test_track_concurrently_state_machine():
    D = synth.DESK
    state = (D["R"], D["T"], D["angular_vel"], D["linear_vel"])
    for n_eng in (1, 3):
        log = []
        engines = [_FakeEngine(log, i) for i in range(n_eng)]
        t = hyp.Tickets()
        n_hyp = 7

        def nxt():
            h = t.next()
            return None if h >= n_hyp else (h, hyp.perturb(*state, h))

        rows = hyp.track_concurrently(engines, nxt, 0.05, None, None, levels=3, chunk=4, make_event=lambda e: _Ev())
        assert sorted(int(r[0]) for r in rows) == list(range(n_hyp))
        for r in rows:      # coarse -> fine, the coarsest level without the velocity, iterations summed over the levels
            T0 = hyp.perturb(*state, int(r[0]))[1].astype(np.float64)
            assert np.allclose(r[12:15], T0 * 0.125) and abs(r[1] - np.linalg.norm(T0 * 0.125)) < 1e-9
            assert r[2] >= 3 * 5
        for i in range(n_eng):
            lv = [x for x in log if x[0] == i and x[1] == "level"]
            assert [x[2] for x in lv[:3]] == [2, 1, 0] and [x[3] for x in lv[:3]] == [False, True, True]
            assert sum(1 for x in log if x[0] == i and x[1] == "resume") == 1
        assert hyp.gather_results(rows, n_hyp).shape == (n_hyp, hyp.ROW)
