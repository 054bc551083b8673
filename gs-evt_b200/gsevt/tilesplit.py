"""Screen-tile split of ONE pose hypothesis across the GPUs of a box (BASELINE.json configs[4]).

The reference is single-GPU.  For maps / images too large for one GPU's iteration budget the
tile rows of the screen are split into contiguous strips, one per rank: every rank projects the whole map
(cheap, HBM-bound, no exchange of records), but bins, sorts, blends and back-propagates only its strip.  Two
things cross GPUs per iteration, both tiny and both inside the kernels that produce them (csrc/split_comm.cuh):
  * {sum d^2, sum d*E, sum E^2}  — the loss normalises over the whole image (utils/render_camera/frame.py:90-91);
  * the 12 pose/velocity gradient sums (dgr/diff_gaussian_rasterization/__init__.py:163-167) + the overflow flag.
They travel as stores into the peers' mailboxes over NVLink (CUDA IPC mappings between the one-process-per-GPU
ranks) and are added in rank order, so every rank applies bit-identical Adam steps to its replica of the state.

This module is the host-side plumbing only: export / exchange / open the mailbox handles and attach.  After
`attach`, every TrackingEngine call is collective — all ranks issue the same calls in the same order.
"""
import contextlib
import ctypes as C

from . import lib as _lib

MAX_RANKS = 8


def _on_device(engine):
    dev = getattr(engine, "device", None)
    if dev is None:
        return contextlib.nullcontext()
    import torch
    return torch.cuda.device(dev)


def balance_rows(row_cost, n):
    """Strip bounds [b0=0, b1, ..., bn=rows] for n ranks, balanced on row_cost (tile instances per tile row).
    Same code the engine runs at begin_level (gsevt_split_balance_rows) — pure host arithmetic."""
    L = _lib.load()
    rows = len(row_cost)
    cost = (C.c_uint32 * max(rows, 1))(*[int(c) for c in row_cost])
    out = (C.c_int32 * (n + 1))()
    _lib.check(L.gsevt_split_balance_rows(cost, rows, int(n), out), "gsevt_split_balance_rows")
    return list(out)


def _mailbox(engine):
    box = C.c_void_p()
    with _on_device(engine):
        _lib.check(engine._lib.gsevt_engine_split_mailbox(engine.handle, C.byref(box)), "gsevt_engine_split_mailbox")
    return box.value


def _attach(engine, rank, boxes, timeout_s):
    arr = (C.c_void_p * len(boxes))(*boxes)
    with _on_device(engine):
        _lib.check(engine._lib.gsevt_engine_split_attach(engine.handle, int(rank), len(boxes), arr, float(timeout_s)),
                   "gsevt_engine_split_attach")


def attach_local(engines, timeout_s=5.0):
    """All ranks live in THIS process (one engine per rank, on one device or on peer-enabled devices): the
    mailboxes are addressed directly.  Used by the single-GPU parity test of the split path and by a host that
    drives several GPUs from one process (enable peer access first)."""
    if not 1 <= len(engines) <= MAX_RANKS:
        raise ValueError(f"1..{MAX_RANKS} ranks")
    boxes = [_mailbox(e) for e in engines]
    for r, e in enumerate(engines):
        _attach(e, r, boxes, timeout_s)
    return boxes


class TileSplitGroup:
    """One process per GPU.  `all_gather_bytes(b: bytes) -> list[bytes]` gathers one 64-byte blob per rank in rank
    order; by default torch.distributed.all_gather_object on the default group (NCCL or gloo)."""

    def __init__(self, engine, rank, world, all_gather_bytes=None, barrier=None, timeout_s=5.0):
        if not 1 <= world <= MAX_RANKS or not 0 <= rank < world:
            raise ValueError(f"rank {rank} / world {world}: 1..{MAX_RANKS} ranks")
        self.engine, self.rank, self.world = engine, int(rank), int(world)
        self._opened = []
        L = engine._lib
        own = _mailbox(engine)
        handle = (C.c_uint8 * 64)()
        _lib.check(L.gsevt_ipc_export(own, handle), "gsevt_ipc_export")
        if all_gather_bytes is None:
            all_gather_bytes, barrier = _dist_exchange()
        blobs = all_gather_bytes(bytes(handle))
        if len(blobs) != world or any(len(b) != 64 for b in blobs):
            raise _lib.GsevtError("tile split: handle exchange returned a malformed table")
        boxes = []
        with _on_device(engine):
            for r, b in enumerate(blobs):
                if r == rank:
                    boxes.append(own)
                    continue
                p = C.c_void_p()
                h = (C.c_uint8 * 64).from_buffer_copy(b)
                _lib.check(L.gsevt_ipc_open(h, C.byref(p)), f"gsevt_ipc_open(rank {r})")
                self._opened.append(p.value)
                boxes.append(p.value)
        _attach(engine, rank, boxes, timeout_s)
        if barrier is not None:
            barrier()   # nobody iterates before everybody has reset its mailbox

    def close(self):
        """Detach and unmap the peers' mailboxes (call on every rank before the engines are destroyed)."""
        if self.engine is not None and getattr(self.engine, "handle", None):
            _attach(self.engine, 0, [None], 0.0)
        with _on_device(self.engine):
            for p in self._opened:
                self.engine._lib.gsevt_ipc_close(p)
        self._opened = []
        self.engine = None


def _dist_exchange():
    import torch.distributed as dist
    if not dist.is_initialized():
        raise _lib.GsevtError("tile split: torch.distributed is not initialised and no exchange callback was given")

    def gather(b):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, b)
        return out

    return gather, dist.barrier
