"""Multi-GPU check of the screen-tile split (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/tilesplit_check.py

Every rank builds the same seeded scene, attaches its engine to the group (mailboxes mapped through CUDA IPC,
exchanges by NVLink stores inside the loss / update kernels), and runs an evaluation plus ITERS fine-stage
iterations.  Rank 0 first does the same with an unsplit engine.  Checks: every rank ends with bit-identical
losses and state; split == unsplit to rounding; no exchange timed out.  Prints per-iteration device times of
both for information and "TILESPLIT OK".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gs-evt_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

ITERS = int(os.environ.get("TILESPLIT_ITERS", "40"))
P = int(os.environ.get("TILESPLIT_P", "300000"))
W, Hh = int(os.environ.get("TILESPLIT_W", "640")), int(os.environ.get("TILESPLIT_H", "480"))


def main():
    import torch
    import torch.distributed as dist
    from gsevt import lib, synth, tilesplit
    from gsevt.engine import EventFrameBuilder, PackedMap, TrackingEngine

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # handles travel over the host; data over NVLink
    lib.require_device()

    D = synth.DESK
    s = W / D["W"]
    fx, fy = D["fx"] * s, D["fy"] * s
    act = synth.activate(synth.synth_map(P, seed=4, W=W, H=Hh, fx=fx, fy=fy))
    A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
    pm = PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3)
    R = np.array(D["R"], np.float32).reshape(3, 3)
    T = np.array(D["T"], np.float32)
    w = np.array(D["angular_vel"], np.float32) * 10
    v = np.array(D["linear_vel"], np.float32)
    K = np.array([fx, 0, W / 2, 0, fy, Hh / 2, 0, 0, 1.0]).reshape(3, 3)
    b = EventFrameBuilder(W, Hh, K, D["dist"], levels=3, device=dev)
    ev = synth.random_events(30000 * W // 640, W, Hh, 0, 50000, seed=3)
    sign, unsign = b.build(ev[:, 1].astype(np.int16), ev[:, 2].astype(np.int16), ev[:, 3].astype(np.uint8))
    torch.cuda.synchronize()

    def fresh():
        e = TrackingEngine(pm, W, Hh, fx, fy, levels=3, converged_threshold=0.0, max_optim_iter=ITERS + 20)
        e.set_state(R, T, w, v)
        e.begin_frame(0.05, sign, unsign)
        return e

    def timed(e, n):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e.begin_level(0, True)
        e.iterate(4)
        e.stream.synchronize()
        dist.barrier()
        t0.record(e.stream)
        e.iterate(n - 4)
        t1.record(e.stream)
        e.stream.synchronize()
        return t0.elapsed_time(t1) / (n - 4)

    ref = None
    if rank == 0:
        e0 = fresh()
        L0, g0 = e0.eval(0, True)
    dist.barrier()
    if rank == 0:
        ms0 = timed(e0, ITERS)
        ref = (L0, g0, e0.losses(), e0.get_state(), ms0)
        e0.close()
    else:
        dist.barrier()   # matches the barrier inside timed() on rank 0

    e = fresh()
    grp = tilesplit.TileSplitGroup(e, rank, world, timeout_s=20.0)
    L, g = e.eval(0, True)
    ms = timed(e, ITERS)
    info = e.split_info()
    losses, state = e.losses(), e.get_state()
    st = e.status()
    row = dict(rank=rank, L=L, g=g, losses=losses, state=state, ms=ms, info=info, iters=st.iters_executed)
    table = [None] * world
    dist.all_gather_object(table, row)
    ok = True
    if rank == 0:
        for r in table:
            same = (r["L"] == table[0]["L"] and np.array_equal(r["g"], table[0]["g"]) and np.array_equal(r["losses"], table[0]["losses"])
                    and all(np.array_equal(a, c) for a, c in zip(r["state"], table[0]["state"])))
            print(f"rank {r['rank']}: rows {r['info']['rows']} iters {r['iters']} exchanges {r['info']['exchanges']} "
                  f"comm_error {r['info']['comm_error']} {r['ms']:.4f} ms/iter identical_to_rank0={same}")
            ok &= same and r["info"]["comm_error"] == 0 and r["iters"] == ITERS
        L0, g0, l0, s0, ms0 = ref
        dL = abs(table[0]["L"] - L0) / abs(L0)
        dg = np.abs(table[0]["g"] - g0).max() / np.abs(g0).max()
        dl = np.abs(table[0]["losses"] - l0).max()
        ds = max(np.abs(a - c).max() for a, c in zip(table[0]["state"], s0))
        print(f"split({world}) vs unsplit: loss rel {dL:.2e}, grad rel {dg:.2e}, loss history abs {dl:.2e}, state abs {ds:.2e}")
        print(f"unsplit {ms0:.4f} ms/iter, split({world}) {max(r['ms'] for r in table):.4f} ms/iter at P={P}, {W}x{Hh}")
        d3 = np.abs(table[0]["losses"][:3] - l0[:3]).max()
        # one evaluation must agree to rounding; along the optimisation path rounding differences grow (random event
        # frame = uninformative objective, Adam walks), so the path is held loosely and its first steps tightly
        ok &= dL < 2e-6 and dg < 5e-5 and d3 < 2e-6 and dl < 2e-2 and ds < 2e-2
        print("TILESPLIT OK" if ok else "TILESPLIT FAILED")
    dist.barrier()
    grp.close()
    e.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
