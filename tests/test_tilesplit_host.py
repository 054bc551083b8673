"""Host logic of the screen-tile split on CPU: the strip partition (gsevt_split_balance_rows, pure host code in
libgsevt.so) and the mailbox-handle exchange of TileSplitGroup over a world_size-2 gloo group with a recording
stand-in for the engine (the GPU run uses the same code with real CUDA IPC handles)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-evt_b200"))

from gsevt import tilesplit  # noqa: E402


def test_balance_known_answers(built):
    assert tilesplit.balance_rows([10] * 8, 1) == [0, 8]
    assert tilesplit.balance_rows([10] * 8, 2) == [0, 4, 8]
    assert tilesplit.balance_rows([10] * 8, 4) == [0, 2, 4, 6, 8]
    assert tilesplit.balance_rows([10] * 8, 8) == list(range(9))
    # all the work in the last rows: the first ranks still get one row each
    assert tilesplit.balance_rows([0, 0, 0, 0, 0, 0, 1000, 1000], 4) == [0, 5, 6, 7, 8]
    # heavy head
    assert tilesplit.balance_rows([1000, 1000, 0, 0, 0, 0, 0, 0], 2) == [0, 1, 8]
    # fewer rows than ranks: one row per rank, the rest empty
    assert tilesplit.balance_rows([5, 5, 5], 4) == [0, 1, 2, 3, 3]
    assert tilesplit.balance_rows([], 2) == [0, 0, 0]


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 200000), min_size=1, max_size=255), st.integers(1, 8))
def test_balance_properties(cost, n):
    import gsevt.lib as lib
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libgsevt.so not built")
    b = tilesplit.balance_rows(cost, n)
    rows = len(cost)
    assert len(b) == n + 1 and b[0] == 0 and b[-1] == rows
    assert all(b[k] <= b[k + 1] for k in range(n))
    if rows >= n:
        assert all(b[k] < b[k + 1] for k in range(n))          # nobody idles while rows are left
        # no strip exceeds the ideal share by more than the two rows next to its cuts (+ the "+1" row weights)
        w = np.asarray(cost, np.int64) + 1
        ideal = w.sum() / n
        for k in range(n):
            strip = w[b[k]:b[k + 1]]
            if len(strip) > 1:
                assert strip.sum() <= ideal + 2 * w.max() + 1


def test_balance_rejects_bad_arguments(built):
    from gsevt import lib
    with pytest.raises(lib.GsevtError):
        tilesplit.balance_rows([1] * 300, 2)
    with pytest.raises(lib.GsevtError):
        tilesplit.balance_rows([1] * 10, 9)


class _RecordingLib:
    """Stand-in for libgsevt's split entry points: mailboxes are fake addresses, handles encode (rank, address)."""

    def __init__(self, rank):
        self.rank, self.calls = rank, []
        self.own = 0x10000000 * (rank + 1)

    def gsevt_engine_split_mailbox(self, handle, box_ref):
        box_ref._obj.value = self.own
        return 0

    def gsevt_ipc_export(self, ptr, out):
        blob = np.zeros(64, np.uint8)
        blob[:16] = np.frombuffer(np.array([self.rank, ptr], np.uint64).tobytes(), np.uint8)
        C.memmove(out, blob.ctypes.data, 64)
        return 0

    def gsevt_ipc_open(self, h, ptr_ref):
        rank, addr = np.frombuffer(bytes(h), np.uint64)[:2]
        ptr_ref._obj.value = int(addr) + 0x1000 * (self.rank + 1)     # a peer mapping has its own local address
        self.calls.append(("open", int(rank)))
        return 0

    def gsevt_ipc_close(self, p):
        self.calls.append(("close", int(p)))
        return 0

    def gsevt_engine_split_attach(self, handle, rank, n, boxes, timeout):
        self.calls.append(("attach", rank, n, [boxes[i] for i in range(n)], timeout))
        return 0


class _Engine:
    def __init__(self, rank):
        self._lib, self.handle, self.device = _RecordingLib(rank), 1, None


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = _Engine(rank)
    g = tilesplit.TileSplitGroup(e, rank, world, timeout_s=2.5)
    attach = [c for c in e._lib.calls if c[0] == "attach"][0]
    opened = [c[1] for c in e._lib.calls if c[0] == "open"]
    g.close()
    q.put((rank, attach, opened, [c for c in e._lib.calls if c[0] == "close"], e._lib.calls[-2 if world > 1 else -1]))
    dist.barrier()
    dist.destroy_process_group()


def test_handle_exchange_over_gloo():
    world, port = 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, attach, opened, closed, detach in got:
        _, r, n, boxes, timeout = attach
        assert (r, n, timeout) == (rank, world, 2.5)
        own = 0x10000000 * (rank + 1)
        assert boxes[rank] == own                                        # own box: the local pointer, never re-opened
        assert opened == [p for p in range(world) if p != rank]          # every peer opened once, in rank order
        peer = 1 - rank
        assert boxes[peer] == 0x10000000 * (peer + 1) + 0x1000 * (rank + 1)
        assert [c[1] for c in closed] == [boxes[peer]]                   # close() unmaps exactly what it opened
        assert detach[0] == "attach" and detach[2] == 1                  # ... after detaching (n = 1)


def test_group_argument_checks():
    with pytest.raises(ValueError):
        tilesplit.TileSplitGroup(_Engine(0), 0, 9, all_gather_bytes=lambda b: [b])
    with pytest.raises(ValueError):
        tilesplit.TileSplitGroup(_Engine(0), 2, 2, all_gather_bytes=lambda b: [b, b])
    from gsevt import lib
    with pytest.raises(lib.GsevtError):
        tilesplit.TileSplitGroup(_Engine(0), 0, 2, all_gather_bytes=lambda b: [b])          # table too short
