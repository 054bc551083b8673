"""Multi-hypothesis initial-pose search, sharded across GPUs (BASELINE.json configs[3]).

The reference has no multi-GPU path; what shards naturally on this problem is the
set of independent pose / velocity hypotheses: each is its own TrackingEngine against a read-only map,
with NO data-path collective.  Two things are shared between the ranks, both outside the optimisation:
  * a ticket counter (`Tickets`): hypotheses run to convergence and their iteration counts differ by a factor
    of several (60 .. 400 per pyramid level), so ranks DRAW the next hypothesis id from an atomic counter
    served by the process group's key-value store instead of owning h mod world up front;
  * the final gather of (hypothesis id, loss, iterations, state) — 21 doubles per hypothesis — to pick the
    winner, one all-reduce with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
On one GPU several hypotheses run CONCURRENTLY (`track_concurrently`): one TrackingEngine per hypothesis, each
on its own CUDA stream with its own CUDA graph, so that the kernels of one hypothesis fill the SMs the tail of
another leaves idle; the host only polls the engines' mapped done-flags.
"""
import itertools
import math
import time

import numpy as np


def assign(n_hypotheses, world):
    """Static assignment: hypothesis ids per rank, h -> rank h mod world."""
    return [list(range(r, n_hypotheses, world)) for r in range(world)]


_search_calls = itertools.count()


class Tickets:
    """Atomic ticket counter over all ranks: next() -> 0, 1, 2, ... each value handed out exactly once.
    With a process group it is `store.add(key, 1)` on the group's store (served by rank 0's TCPStore daemon, a
    host-side round trip of tens of microseconds per hypothesis, nothing on the GPU data path); without one it
    is a local counter.  Every rank must construct its Tickets in the same order (the key is numbered)."""

    def __init__(self, dist=None, store=None):
        self._key = f"gsevt/hypothesis_ticket/{next(_search_calls)}"
        self._local = itertools.count()
        self._store = store
        self._stride = None
        if store is None and dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            try:
                from torch.distributed import distributed_c10d as c10d
                self._store = c10d._get_default_store()
            except Exception:
                # no access to the group's store in this torch build: fall back to the static partition h mod world,
                # expressed as a ticket sequence (rank, rank + world, ...), so that callers need no second code path
                self._stride = (dist.get_rank(), dist.get_world_size())

    @property
    def dynamic(self):
        return self._stride is None

    def next(self):
        if self._stride is not None:
            r, w = self._stride
            return r + w * next(self._local)
        if self._store is None:
            return next(self._local)
        return int(self._store.add(self._key, 1)) - 1


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


def search(make_engine, state, n_hypotheses, run_one, dist=None, device="cpu", assignment="dynamic", tickets=None):
    """Runs this rank's share of the hypotheses and gathers the table on every rank.
    make_engine() -> engine;  run_one(engine, (R, T, w, v)) -> (loss, iterations, R, T, w, v).
    assignment "dynamic": ids drawn from the shared ticket counter; "static": h mod world."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    eng = make_engine()
    rows = []
    if assignment == "static":
        ids = iter(assign(n_hypotheses, world)[rank])
    else:
        t = tickets if tickets is not None else Tickets(dist)
        ids = iter(t.next, None)
    for h in ids:
        if h >= n_hypotheses:
            break
        loss, iters, R, T, w, v = run_one(eng, perturb(*state, h))
        rows.append(pack_result(h, loss, iters, R, T, w, v))
    return gather_results(rows, n_hypotheses, dist, device)


def track_concurrently(engines, next_hypothesis, delta_tau, sign_pyramid, unsign_pyramid, levels, chunk=8, make_event=None):
    """Optimises hypotheses to convergence on the engines of ONE GPU, len(engines) at a time.

    engines           TrackingEngine objects over the same PackedMap, each with its own stream;
    next_hypothesis   () -> (h, (R, T, w, v)) or None when there is nothing left to draw;
    levels            pyramid levels, run coarse -> fine with the tracker's rule (utils/tracker.py:116-240: the
                      coarsest level optimises the pose only, the others pose + velocity);
    make_event        engine -> object with .query() that becomes true when the work queued so far on the engine's
                      stream has finished (default: a torch.cuda.Event recorded on engine.stream).
    Returns the result rows (pack_result) of every hypothesis this call ran.

    An engine is never synchronised while another one could be fed: each gets `chunk` graph launches, the host then
    watches the chunk's event and the engine's done-flag (a mapped host word the device writes) and either queues the
    next chunk, moves the hypothesis to the next level, or draws the next hypothesis for the engine."""
    if make_event is None:
        import torch

        def make_event(eng):
            ev = torch.cuda.Event()
            ev.record(eng.stream)
            return ev

    rows, slots = [], [None] * len(engines)

    def feed(i):
        slots[i]["event"] = None
        engines[i].iterate(chunk)
        slots[i]["event"] = make_event(engines[i])

    def start(i):
        nxt = next_hypothesis()
        if nxt is None:
            slots[i] = None
            return
        h, st = nxt
        eng = engines[i]
        eng.set_state(*st)
        eng.begin_frame(delta_tau, sign_pyramid, unsign_pyramid)
        slots[i] = dict(h=h, level=levels - 1, iters=0, event=None)
        eng.begin_level(levels - 1, opt_vel=False)
        feed(i)

    for i in range(len(engines)):
        start(i)
    while any(s is not None for s in slots):
        progressed = False
        for i, s in enumerate(slots):
            if s is None or not s["event"].query():
                continue
            progressed = True
            eng = engines[i]
            flag = eng.poll_done()
            if flag == 0:
                feed(i)
            elif flag == 2:
                eng.resume()
                feed(i)
            elif flag == 3:
                raise RuntimeError("tile-split exchange timed out inside a hypothesis search")
            else:
                st = eng.status()
                s["iters"] += int(st.optim_iter)
                if s["level"] > 0:
                    s["level"] -= 1
                    eng.begin_level(s["level"], opt_vel=True)
                    feed(i)
                else:
                    R, T, w, v = eng.get_state()
                    rows.append(pack_result(s["h"], float(st.last_loss), s["iters"], R, T, w, v))
                    start(i)
        if not progressed:
            time.sleep(2e-5)   # every engine has a chunk in flight: yield the core instead of spinning on the queries
    return rows
