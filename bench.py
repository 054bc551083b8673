#!/usr/bin/env python3
"""bench.py — pose-tracking iterations/s of the GS-EVT tracking hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (libgsevt, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference pipeline (oracle/_ref)
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (config.workload): BASELINE.json configs[2] — robot_normal1-shaped camera / initial pose /
velocity, synthetic 1 M-Gaussian map, 640x480, 30 000-event packets.  A "step" is ONE tracking
iteration at full resolution in the fine stage = everything inside the reference's innermost loop
(utils/tracker.py:176-240): two renders (t -+ dtau/2), normalised intensity-change loss, two backward
passes down to the 12 pose/velocity gradients, Adam, pose + velocity update, convergence bookkeeping.

  value  iterations/s with the event frame already in HBM: K graph launches of the native engine,
         timed with CUDA events on the engine's stream (barrier + synchronize on both sides, max over
         ranks).  Convergence is disabled (threshold 0) so that exactly K iterations execute.
  e2e    the same metric through the public Python API with HOST inputs: per event frame, the 30 000
         events go from pinned host memory to the device, the event frame is built, `--iters-per-frame`
         iterations run, and loss + pose come back to the host; wall clock between synchronizes.
  N>1    one independent pose hypothesis per GPU (BASELINE.json configs[3]); no data-path collective;
         value = N*K / max-over-ranks time; scaling "weak".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gs-evt_b200")


# --------------------------------------------------------------------------------------------------
# shared, implementation-neutral scene description (numpy only)
# --------------------------------------------------------------------------------------------------
def scene_description(args):
    # synth.py is numpy-only scene generation shared by both arms.  It is loaded by file path so that the
    # reference arm never has the product's source root on sys.path (the reference's `utils` is a namespace
    # package and would lose against the product's regular `utils` package).
    import importlib.util
    spec = importlib.util.spec_from_file_location("gsevt_synth_standalone", os.path.join(PKG, "gsevt", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    base = synth.ROBOT if args.camera == "robot" else synth.DESK
    W, H = args.width, args.height
    s = W / base["W"]
    d = dict(W=W, H=H, fx=base["fx"] * s, fy=base["fy"] * s, cx=W / 2.0, cy=H / 2.0, dist=list(base["dist"]),
             R=list(base["R"]), T=list(base["T"]), angular_vel=list(base["angular_vel"]), linear_vel=list(base["linear_vel"]),
             lr=dict(base["lr"]), converged_threshold=0.0, max_optim_iter=1 << 20, max_events_per_frame=args.events,
             delta_tau=0.05, background=[0, 0, 0])
    raw = synth.synth_map(args.gaussians, seed=args.seed, W=W, H=H, fx=d["fx"], fy=d["fy"], R=d["R"], T=d["T"])
    return d, raw, synth


def perturbed_state(d, hyp):
    """Hypothesis `hyp`: N(0, 5 cm) on T, N(0, 1 deg) rotation about a random axis, +-20 % on the velocities
    (BASELINE.json configs[3], seed 7 + hyp)."""
    rng = np.random.default_rng(7 + hyp)
    R = np.asarray(d["R"], np.float64).reshape(3, 3)
    T = np.asarray(d["T"], np.float64)
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = math.radians(1.0) * rng.normal()
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    dR = np.eye(3) + math.sin(ang) * Kx + (1 - math.cos(ang)) * Kx @ Kx
    w = np.asarray(d["angular_vel"], np.float64) * (1 + 0.2 * rng.uniform(-1, 1, 3))
    v = np.asarray(d["linear_vel"], np.float64) * (1 + 0.2 * rng.uniform(-1, 1, 3))
    return (dR @ R).astype(np.float32), (T + rng.normal(0, 0.05, 3)).astype(np.float32), w.astype(np.float32), v.astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.kill()
        self.th.join(timeout=2)
        sm, smax, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        dist = dist_
    else:
        torch.cuda.set_device(local)
    if args.gpus != world:
        if rank == 0:
            print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    return dist, world, rank, local


def barrier_sync(dist, torch):
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(dist, torch, x):
    if dist is None:
        return float(x)
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, torch, x):
    if dist is None:
        return float(x)
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# this repo
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    sys.path.insert(0, PKG)
    import torch
    dist, world, rank, local = dist_setup(args)
    from gsevt import lib
    from gsevt.engine import PackedMap, TrackingEngine
    from utils.event_camera.event import EventArray, EventFrame
    lib.require_device()
    dev = torch.device("cuda", local)
    d, raw, synth = scene_description(args)
    act = synth.activate(raw)
    A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
    pm = PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3)
    del A
    eng = TrackingEngine(pm, d["W"], d["H"], d["fx"], d["fy"], levels=3, lr_rot=d["lr"]["cam_rot_delta"],
                         lr_trans=d["lr"]["cam_trans_delta"], lr_w=d["lr"]["cam_w_delta"], lr_v=d["lr"]["cam_v_delta"],
                         converged_threshold=0.0, max_optim_iter=1 << 20)
    K = np.array([d["fx"], 0, d["cx"], 0, d["fy"], d["cy"], 0, 0, 1.0]).reshape(3, 3)
    Rt = np.asarray(d["R"], np.float32).reshape(3, 3)
    Tt = np.asarray(d["T"], np.float32)
    wt, vt = np.asarray(d["angular_vel"], np.float32), np.asarray(d["linear_vel"], np.float32)

    # ground-truth intensity change at the true state -> synthetic events
    dummy = EventFrame(d["W"], d["H"], K, d["dist"], 9, EventArray(*np.zeros((4, 1), np.int64)), device=dev)
    eng.set_state(Rt, Tt, wt, vt)
    eng.begin_frame(d["delta_tau"], dummy.sign_pyramid, dummy.unsign_pyramid)
    eng.eval(0, True)
    g_last, g_next = eng.gray_images(0)
    delta_gt = (g_next - g_last).cpu().numpy()
    n_frames_e2e = max(1, math.ceil(args.steps / args.iters_per_frame)) + 1
    packets = []
    for j in range(n_frames_e2e):
        tab = synth.sample_events(delta_gt, args.events, j * 50000, (j + 1) * 50000, K, d["dist"], seed=1000 + j)
        packets.append(EventArray(tab[:, 0], tab[:, 1], tab[:, 2], tab[:, 3]))
    ev_bytes = args.events * 5

    hyp = rank  # one hypothesis per GPU
    R0, T0, w0, v0 = perturbed_state(d, hyp)

    # ---- value: device-resident ------------------------------------------------------------------
    ef = EventFrame(d["W"], d["H"], K, d["dist"], 9, packets[0], device=dev)
    eng.set_state(R0, T0, w0, v0)
    eng.begin_frame(d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid)
    eng.begin_level(0, True)
    eng.iterate(max(args.warmup, 3))
    eng.stream.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(dist, torch)
    e0.record(eng.stream)
    eng.iterate(args.steps)
    e1.record(eng.stream)
    eng.stream.synchronize()
    barrier_sync(dist, torch)
    ms = e0.elapsed_time(e1)
    st = eng.status()
    executed = st.iters_executed
    assert executed == max(args.warmup, 3) + args.steps, f"work skipped: {executed} iterations executed"
    assert np.isfinite(st.last_loss)
    ms_max = max_over_ranks(dist, torch, ms)
    wl = eng.workload()

    # ---- e2e: host inputs, public API ---------------------------------------------------------------
    def frame(j, n):
        efj = EventFrame(d["W"], d["H"], K, d["dist"], 9, packets[j], device=dev)        # H2D (pinned) + GPU frame build
        eng.begin_frame(packets[j].duration(), efj.sign_pyramid, efj.unsign_pyramid)
        eng.begin_level(0, True)
        eng.iterate(n)
        s = eng.status()                                                                    # D2H: loss, gradients, counters
        Rm, T, w, v = eng.get_state()                                                       # D2H: pose + velocity
        return s, (Rm, T, w, v)

    eng.set_state(R0, T0, w0, v0)
    frame(0, min(args.iters_per_frame, max(args.warmup, 3)))
    barrier_sync(dist, torch)
    t0 = time.perf_counter()
    left, j, frames_run = args.steps, 1, 0
    while left > 0:
        n = min(args.iters_per_frame, left)
        s, _ = frame(j % len(packets), n)
        assert s.iters_executed == n
        left -= n
        j += 1
        frames_run += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier_sync(dist, torch)
    e2e_s_max = max_over_ranks(dist, torch, e2e_s)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline: per-stage device time, measured live with CUDA events on the engine stream ----------
    eng.set_state(R0, T0, w0, v0)
    eng.begin_frame(d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid)
    eng.begin_level(0, True)
    eng.iterate(3)
    stages = eng.profile(10)
    wl = eng.workload()
    default_wl = (args.gaussians, d["W"], d["H"], args.camera, args.seed) == (1_000_000, 640, 480, "robot", 1)
    roof, stage_table = roofline(stages, wl, args.gaussians, d["W"] * d["H"], clocks, default_wl, tiles=((d["W"] + 15) // 16) * ((d["H"] + 15) // 16))

    # ---- parity of the benchmarked path at the benchmarked size (outside every timed region) ------------
    # the fused engine against the reference-shaped autograd loop through the drop-in operator — the path the parity tests
    # pin to the live reference build: loss, 12 gradients, gray images, and the per-tile lists / ranges bit for bit
    parity = None
    if rank == 0 and not args.no_parity_check:
        from gsevt import selfcheck
        A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
        eng.begin_frame(d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid)
        parity = selfcheck.engine_vs_operator(eng, A, (R0, T0, w0, v0), ef.builder.level_view(ef.sign_pyramid, 0)[0])
        parity["gates"] = {"loss_rel": 1e-5, "grad_rel_max": 1e-3, "gray_rel_max": 1e-4, "lists_bit_identical": True,
                           "n_contrib_final_T_bit_identical": True}
        parity["ok"] = bool(parity["loss_rel"] < 1e-5 and parity["grad_rel_max"] < 1e-3 and parity["gray_rel_max"] < 1e-4
                            and parity["lists_bit_identical"] and parity["n_contrib_final_T_bit_identical"])
        parity["against"] = ("two rasterisations + torch loss + backward through this repo's drop-in diff_gaussian_rasterization (pinned to "
                             "the live reference build by tests/test_gpu_parity.py), same state, same event frame, the camera blocks "
                             "of the engine's pose kernel")
        del A
        assert parity["ok"], f"benchmarked path disagrees with the operator path: {parity}"

    out = None
    if rank == 0:
        value = world * args.steps / (ms_max / 1e3)
        status_bytes = 112 + 72  # GsevtEngineStatus + pose/velocity read-back per frame
        out = {
            "metric": "pose-track iters/sec (fwd+bwd, 640x480, 1M Gaussians)", "value": round(value, 2), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max / args.steps, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: {args.camera}_normal1-shaped synthetic sequence, {args.gaussians}-Gaussian map, "
                                   f"{d['W']}x{d['H']}, {args.events} events/frame, full-resolution fine-stage iterations",
                       "gaussians": args.gaussians, "width": d["W"], "height": d["H"], "events_per_frame": args.events,
                       "parallelism": f"hyp{world} (one independent pose hypothesis per GPU, no collective)",
                       "l2": "working set per iteration (map 232 MB + records/keys > 300 MB) exceeds the 126 MB L2",
                       "iters_per_frame_e2e": args.iters_per_frame},
            "clocks": clocks,
            "e2e": {"value": round(world * args.steps / e2e_s_max, 2), "unit": "iterations/s",
                    "h2d_bytes_per_step": round(ev_bytes * frames_run / args.steps, 1),
                    "d2h_bytes_per_step": round(status_bytes * frames_run / args.steps, 1),
                    "frames": frames_run, "note": "per frame: events pinned-host->device, GPU event frame, iterations, loss+pose read-back"},
            "gpu_launches": eng.launches_per_iteration * args.steps,
            "gpu_launches_note": "kernels per iteration (preprocess_map, bucket_scatter, bucket_sort, blend_fwd, loss_stats, blend_bwd, "
                                 "geom_compact, geom_bwd, engine_update) x steps: all this library's own, no library kernels, one CUDA "
                                 "graph launch per iteration",
            "roofline": roof,
            "parity_check": parity,
            "stages_ms": stage_table,
            "workload_counters": wl,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, d, act, R0, T0, w0, v0, ef)
    # ---- sub-objects, after every timed region of the headline: configs[3] as written, and (N > 1) configs[4] -------------
    if not args.no_extras:
        import argparse as _ap
        extras = {}
        try:
            extras["search"] = hypothesis_search(torch, dist, world, rank, dev, pm, d, ef, args.hypotheses, 4)
        except Exception as ex:   # the headline line must survive a failing extra
            extras["search"] = {"error": f"{type(ex).__name__}: {ex}"}
        eng.close()
        pm.close()
        del eng, pm
        torch.cuda.empty_cache()
        if world > 1:
            ts_args = _ap.Namespace(**vars(args))
            ts_args.gaussians, ts_args.width, ts_args.height, ts_args.steps, ts_args.warmup = 5_000_000, 1280, 720, 100, 5
            try:
                extras["tilesplit"] = tilesplit_line(torch, dist, world, rank, local, ts_args)
            except Exception as ex:
                extras["tilesplit"] = {"error": f"{type(ex).__name__}: {ex}"}
        if rank == 0:
            out.update(extras)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: multi-hypothesis initial-pose search, every hypothesis to convergence
# --------------------------------------------------------------------------------------------------
def hypothesis_search(torch, dist, world, rank, dev, pm, d, ef, n_hyp, concurrent, assignment="dynamic", chunk=8):
    """`n_hyp` perturbed pose/velocity seeds (5 cm / 1 deg / +-20 %, gsevt.hypotheses.perturb) tracked against ONE event
    frame, each through the three pyramid levels with the reference's stopping rule (utils/tracker.py:116-240, threshold
    1e-4, 200-iteration caps).  Ranks draw hypothesis ids from a shared ticket counter (no up-front partition: iteration
    counts differ by hypothesis), `concurrent` hypotheses run at a time per GPU on their own streams, and ONE all-reduce at
    the end gathers the result table.  Device-timed: CUDA events on the default stream, which waits for every engine
    stream; barrier + synchronize on both sides, max over ranks."""
    from gsevt import hypotheses as hyp
    from gsevt.engine import TrackingEngine
    engines = [TrackingEngine(pm, d["W"], d["H"], d["fx"], d["fy"], levels=3, lr_rot=d["lr"]["cam_rot_delta"],
                              lr_trans=d["lr"]["cam_trans_delta"], lr_w=d["lr"]["cam_w_delta"], lr_v=d["lr"]["cam_v_delta"],
                              converged_threshold=1e-4, max_optim_iter=200) for _ in range(concurrent)]
    state = (d["R"], d["T"], d["angular_vel"], d["linear_vel"])

    def one_search(n, first_id):
        tickets = hyp.Tickets(dist)
        static = iter(hyp.assign(n, world)[rank])

        def nxt():
            h = tickets.next() if assignment == "dynamic" else next(static, n)
            return None if h >= n else (h, hyp.perturb(*state, first_id + h))

        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier_sync(dist, torch)
        e0.record(cur)
        for eng in engines:
            eng.stream.wait_stream(cur)
        rows = hyp.track_concurrently(engines, nxt, d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid, levels=3, chunk=chunk)
        for eng in engines:
            cur.wait_stream(eng.stream)
        e1.record(cur)
        torch.cuda.synchronize()
        busy_ms = e0.elapsed_time(e1)        # this rank's own span
        barrier_sync(dist, torch)
        ms = max_over_ranks(dist, torch, busy_ms)
        table = hyp.gather_results(rows, n, dist, dev)
        return table, ms, busy_ms, len(rows)

    one_search(min(n_hyp, 2 * world * concurrent), 1000)          # warm-up: graphs instantiated, buffers grown
    table, ms, busy_ms, mine = one_search(n_hyp, 0)
    per_rank = [None] * world
    if dist is not None:
        dist.all_gather_object(per_rank, dict(rank=rank, hypotheses=mine, busy_ms=round(busy_ms, 2)))
    else:
        per_rank = [dict(rank=0, hypotheses=mine, busy_ms=round(busy_ms, 2))]
    for eng in engines:
        eng.close()
    h, loss, row = hyp.best(table)
    iters = table[:, 2]
    truth_T = np.asarray(d["T"], np.float64)
    return {"workload": f"configs[3]: {n_hyp} perturbed seeds x one {d['W']}x{d['H']} event frame to convergence (3 pyramid levels, "
                        f"threshold 1e-4, caps 200), {len(table)} results gathered",
            "hypotheses": int(n_hyp), "n_gpus": world, "concurrent_per_gpu": concurrent, "assignment": assignment,
            "seconds": round(ms / 1e3, 4), "hypotheses_per_s": round(n_hyp / (ms / 1e3), 3),
            "iterations_total": int(iters.sum()), "iterations_per_s": round(float(iters.sum()) / (ms / 1e3), 1),
            "iterations_per_hypothesis": {"min": int(iters.min()), "median": float(np.median(iters)), "max": int(iters.max())},
            "best": {"hypothesis": h, "loss": round(loss, 6), "trans_error_m": round(float(np.linalg.norm(row[12:15] - truth_T)), 5)},
            "start_trans_error_m_median": round(float(np.median([np.linalg.norm(hyp.perturb(*state, k)[1] - truth_T) for k in range(n_hyp)])), 5),
            "final_trans_error_m_median": round(float(np.median(np.linalg.norm(table[:, 12:15] - truth_T, axis=1))), 5),
            "per_rank": per_rank, "timing": "CUDA events on the default stream joined with every engine stream; max over ranks"}


def run_search(args):
    """`--mode search`: configs[3] on its own (hypotheses/s); the default mode carries the same object under "search"."""
    sys.path.insert(0, PKG)
    import torch
    dist, world, rank, local = dist_setup(args)
    from gsevt import lib
    from gsevt.engine import PackedMap, TrackingEngine
    from utils.event_camera.event import EventArray, EventFrame
    lib.require_device()
    dev = torch.device("cuda", local)
    d, raw, synth = scene_description(args)
    act = synth.activate(raw)
    A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
    pm = PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3)
    del A
    K = np.array([d["fx"], 0, d["cx"], 0, d["fy"], d["cy"], 0, 0, 1.0]).reshape(3, 3)
    ef = ground_truth_event_frame(torch, dev, pm, d, K, synth, args.events)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    res = {}
    for c in args.concurrent:
        res[f"concurrent_{c}"] = hypothesis_search(torch, dist, world, rank, dev, pm, d, ef, args.hypotheses, c, args.assignment)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        best = max(res.values(), key=lambda r: r["hypotheses_per_s"])
        out = {"metric": "hypotheses/s (64 perturbed seeds, one event frame each to convergence)", "value": best["hypotheses_per_s"],
               "unit": "hypotheses/s", "n_gpus": world, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": best["workload"], "gaussians": args.gaussians, "width": d["W"], "height": d["H"],
                                               "parallelism": f"hypotheses over {world} GPUs, ticket counter, no data-path collective"},
               "clocks": clocks, "runs": res}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ground_truth_event_frame(torch, dev, pm, d, K, synth, n_events, seed=1000):
    """Event frame sampled from the intensity change rendered at the TRUE pose / velocity."""
    from gsevt.engine import TrackingEngine
    from utils.event_camera.event import EventArray, EventFrame
    eng = TrackingEngine(pm, d["W"], d["H"], d["fx"], d["fy"], levels=1)
    dummy = EventFrame(d["W"], d["H"], K, d["dist"], 9, EventArray(*np.zeros((4, 1), np.int64)), device=dev)
    eng.set_state(np.asarray(d["R"], np.float32).reshape(3, 3), np.asarray(d["T"], np.float32),
                  np.asarray(d["angular_vel"], np.float32), np.asarray(d["linear_vel"], np.float32))
    eng.begin_frame(d["delta_tau"], dummy.sign_pyramid, dummy.unsign_pyramid)
    eng.eval(0, True)
    g_last, g_next = eng.gray_images(0)
    tab = synth.sample_events((g_next - g_last).cpu().numpy(), n_events, 0, 50000, K, d["dist"], seed=seed)
    eng.close()
    return EventFrame(d["W"], d["H"], K, d["dist"], 9, EventArray(tab[:, 0], tab[:, 1], tab[:, 2], tab[:, 3]), device=dev)


# --------------------------------------------------------------------------------------------------
# this repo, screen-tile split: ONE hypothesis over all ranks (BASELINE configs[4]; strong scaling)
# --------------------------------------------------------------------------------------------------
def run_tilesplit(args):
    """`--mode tilesplit`: the tile rows of one hypothesis are split over the N ranks; the loss sums and the 12-float
    gradient are exchanged inside the kernels over NVLink peer memory (gsevt/tilesplit.py).  The driver's default line
    (hypothesis-parallel weak scaling) carries the same measurement as its "tilesplit" sub-object when N > 1; run it on its
    own with e.g.
      torchrun --nproc-per-node 8 bench.py --gpus 8 --mode tilesplit --gaussians 5000000 --width 1280 --height 720"""
    sys.path.insert(0, PKG)
    import torch
    dist, world, rank, local = dist_setup(args)
    out = tilesplit_line(torch, dist, world, rank, local, args)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def tilesplit_line(torch, dist, world, rank, local, args):
    """One hypothesis, tile rows split over `world` ranks (BASELINE.json configs[4] when called with 5 M / 1280x720).
    Every rank returns; rank 0 gets the JSON object, the others None."""
    from gsevt import lib, tilesplit
    from gsevt.engine import PackedMap, TrackingEngine
    lib.require_device()
    dev = torch.device("cuda", local)
    d, raw, synth = scene_description(args)
    act = synth.activate(raw)
    A = {k: torch.from_numpy(v).to(dev) for k, v in act.items()}
    pm = PackedMap(A["xyz"], A["scales"], A["rotations"], A["opacities"], A["shs"], 3)
    del A, act, raw
    eng = TrackingEngine(pm, d["W"], d["H"], d["fx"], d["fy"], levels=3, lr_rot=d["lr"]["cam_rot_delta"],
                         lr_trans=d["lr"]["cam_trans_delta"], lr_w=d["lr"]["cam_w_delta"], lr_v=d["lr"]["cam_v_delta"],
                         converged_threshold=0.0, max_optim_iter=1 << 20)
    K = np.array([d["fx"], 0, d["cx"], 0, d["fy"], d["cy"], 0, 0, 1.0]).reshape(3, 3)
    # events from the ground-truth intensity change, rendered unsplit on every rank (identical everywhere)
    ef = ground_truth_event_frame(torch, dev, pm, d, K, synth, args.events)
    R0, T0, w0, v0 = perturbed_state(d, 0)          # the same hypothesis on every rank
    # the unsplit answer at the start state, on every rank, before the group exists
    eng.set_state(R0, T0, w0, v0)
    eng.begin_frame(d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid)
    loss_u, grad_u = eng.eval(0, True)
    grp = tilesplit.TileSplitGroup(eng, rank, world, timeout_s=30.0) if world > 1 else None
    eng.set_state(R0, T0, w0, v0)
    eng.begin_frame(d["delta_tau"], ef.sign_pyramid, ef.unsign_pyramid)
    loss_s, grad_s = eng.eval(0, True)               # collective: every rank scores its strip, sums exchanged in-kernel
    gmax = float(np.abs(grad_u).max())
    vs_unsplit = {"loss_rel": abs(loss_s - loss_u) / max(abs(loss_u), 1e-30), "grad_rel_max": float(np.abs(grad_s - grad_u).max() / max(gmax, 1e-30))}
    eng.begin_level(0, True)
    eng.iterate(max(args.warmup, 3))
    eng.stream.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(dist, torch)
    e0.record(eng.stream)
    eng.iterate(args.steps)
    e1.record(eng.stream)
    eng.stream.synchronize()
    barrier_sync(dist, torch)
    ms_max = max_over_ranks(dist, torch, e0.elapsed_time(e1))
    st = eng.status()
    assert st.iters_executed == max(args.warmup, 3) + args.steps, f"work skipped: {st.iters_executed} iterations executed"
    info = eng.split_info()
    assert info["comm_error"] == 0
    clocks = sampler.stop() if rank == 0 else None
    state_bytes = b"".join(np.ascontiguousarray(x, np.float32).tobytes() for x in eng.get_state())
    stages = eng.profile(10)                          # collective: every rank replays 10 un-graphed iterations
    wl = eng.workload()
    rows = [None] * world
    mine = dict(rank=rank, rows=info["rows"], instances=sum(wl["instances"]), pairs_walked=sum(wl["pairs_walked"]),
                stages_ms={k: round(v, 4) for k, v in stages.items()}, loss=float(st.last_loss), state=state_bytes.hex())
    if dist is not None:
        dist.all_gather_object(rows, mine)
    else:
        rows = [mine]
    out = None
    if rank == 0:
        identical = all(r["loss"] == rows[0]["loss"] and r["state"] == rows[0]["state"] for r in rows)
        for r in rows:
            del r["state"]
        out = {"metric": "pose-track iters/sec (fwd+bwd), one hypothesis, screen-tile split", "value": round(args.steps / (ms_max / 1e3), 2),
               "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": round(ms_max / args.steps, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"configs[4]-shaped: {args.gaussians}-Gaussian map, {d['W']}x{d['H']}, one hypothesis, tile rows split "
                                      f"over {world} GPUs", "gaussians": args.gaussians, "width": d["W"], "height": d["H"],
                          "parallelism": f"tilesplit{world}: in-kernel exchange of 3 loss sums + 12 gradient sums over NVLink peer memory",
                          "l2": "map alone exceeds the 126 MB L2"},
               "clocks": clocks, "gpu_launches": eng.launches_per_iteration * args.steps,
               "ranks_bit_identical": bool(identical),
               "ranks_bit_identical_what": f"loss and the 18 floats of pose + velocity after {max(args.warmup, 3) + args.steps} optimiser steps, byte compare over all ranks",
               "vs_unsplit": {k: float(f"{v:.3e}") for k, v in vs_unsplit.items()}, "ranks": rows}
        assert identical, "ranks diverged"
    if grp is not None:
        dist.barrier()
        grp.close()
    eng.close()
    pm.close()
    return out


def load_traffic():
    """DRAM bytes per launch of each stage from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); {} when absent."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return {k: v["traffic_bytes"] for k, v in j["stages"].items()}, j.get("source")
    except Exception:
        return {}, None


def roofline(stages, wl, P, HW, clocks, default_workload=True, tiles=1200):
    """Algorithmic bytes / flops per launch (DESIGN.md "Kernels and rooflines") over the measured stage time."""
    hbm_peak, sm_max, how = load_peaks()
    traffic, traffic_src = load_traffic() if default_workload else ({}, None)
    Pv, N, S = sum(wl["visible"]), sum(wl["instances"]), sum(wl["pairs_walked"])
    Pg = wl["gaussians_with_grad"]
    bytes_alg = {
        # projection, SURVEY.md 8(d) row A: 44 B of geometry per Gaussian, 192 B of SH per Gaussian visible in >= 1 view, out
        # 32 B per visible pair (depth, pixel centre, conic + opacity, gray) and 4 B per Gaussian (tiles touched)
        "preprocess_map": 44 * P + 192 * max(wl["visible"]) + 32 * Pv + 4 * P,
        # binning (csrc/bucketbin.cu), irreducible bytes: every visible pair's (rect, depth bits, id) must be read once
        # (12 B) and every tile instance's id written once (4 B); the bucket keys in between are this design's own traffic
        "bucket_scatter": 12 * Pv,
        "bucket_sort": 4 * N,
        # compaction scan (rect word per pair + 24 B of the accumulator per visible pair) + per active pair: list entry
        # out/in, accumulator in and cleared, xyz/opacity + covariance, SH (AoS copy), clamp byte
        "geom_bwd_pose": 2 * P * 4 + Pv * 24 + Pg * (8 + 32 + 32 + 40 + 192 + 1),
        "loss_stats": 3 * 4 * HW,
    }
    flops_alg = {"blend_fwd_gray": 30.0 * S, "blend_bwd_gray": 100.0 * S}
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s at the clock seen during the run
    table = {}
    for name, ms in stages.items():
        row = {"ms": round(ms, 4)}
        if name in bytes_alg and ms > 0.002:   # (a stage folded into its neighbour reports ~0: no rate for it)
            row.update(bound="hbm", achieved_gbs=round(bytes_alg[name] / (ms * 1e-3) / 1e9, 1),
                       frac=round(bytes_alg[name] / (ms * 1e-3) / 1e9 / hbm_peak, 4), alg_bytes=int(bytes_alg[name]))
        if name in flops_alg and ms > 0:
            row.update(bound="fp32", achieved_tflops=round(flops_alg[name] / (ms * 1e-3) / 1e12, 3),
                       frac=round(flops_alg[name] / (ms * 1e-3) / 1e12 / fp32_peak, 4), alg_flops=int(flops_alg[name]))
        if name in traffic:
            row["traffic_bytes"] = int(traffic[name])
        table[name] = row
    # the dominant HBM-bound kernel carries the contract's `roofline` object (every stage is listed with its own fraction
    # in `stages_ms`)
    hb = max((n for n in table if table[n].get("bound") == "hbm"), key=lambda n: table[n]["ms"])
    top = max(table, key=lambda n: table[n]["ms"])
    r = table[hb]
    roof = {"kernel": hb, "bound": "hbm", "achieved": r["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": r["frac"],
            "traffic": r.get("traffic_bytes"), "traffic_source": f"profiles/{traffic_src} (ncu --set full, same workload)" if traffic_src else None,
            "alg_bytes": r["alg_bytes"], "peak_source": how, "ms_per_launch": r["ms"],
            "top_stage": top, "top_stage_ms": table[top]["ms"],
            "fp32_peak_tflops_at_clock": round(fp32_peak, 2)}
    if table[top].get("bound") == "fp32":
        roof["top_stage_fp32"] = {"achieved": table[top]["achieved_tflops"], "unit": "TFLOP/s", "frac": table[top]["frac"]}
    return roof, table


def cpu_baseline(args, d, act, R0, T0, w0, v0, ef):
    """One tracking-objective evaluation (2 renders + loss + 2 backward passes to the 12 gradients) of the
    SAME workload on ONE host core with the plain-C oracle (oracle/liboracle.so): the reported CPU baseline."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    P = act["xyz"].shape[0]
    frac = 1.0
    sub = act
    budget_P = args.cpu_sample_gaussians
    if P > budget_P:  # bounded sample: a prefix of the (randomly ordered) map, time scaled linearly in P
        sub = {k: v[:budget_P] for k, v in act.items()}
        frac = budget_P / P
    E = ef.sign_delta_Ie[0].cpu().numpy()
    t0 = time.perf_counter()
    orc.tracking_eval(sub, R0, T0, w0, v0, d["delta_tau"], d["W"], d["H"], d["fx"], d["fy"], 0, E, True)
    dt = time.perf_counter() - t0
    est = dt / frac
    return {"value": round(1.0 / est, 5), "unit": "iterations/s", "cores": 1, "kind": "port",
            "sample": f"1 iteration (2 views fwd+bwd+loss) on a {sub['xyz'].shape[0]}-Gaussian prefix of the map at {d['W']}x{d['H']}: "
                      f"{dt:.2f} s, scaled x{1 / frac:.2f} to {P} Gaussians", "host_cpus": os.cpu_count()}


# --------------------------------------------------------------------------------------------------
# the unmodified reference pipeline
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs and prints the reference arm
    sys.path.insert(0, ROOT)
    from oracle import ref_runner
    if not ref_runner.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return
    ref_runner.activate()
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    d, raw, synth = scene_description(args)
    K = np.array([d["fx"], 0, d["cx"], 0, d["fy"], d["cy"], 0, 0, 1.0]).reshape(3, 3)
    it = ref_runner.RefIterations(d, raw, save_path="/tmp/gsevt_ref_bench")
    vp = it.viewpoint
    # ground-truth intensity change at the true state, rendered by the reference itself
    from utils.render_camera.frame import RenderFrame
    vp.delta_tau = d["delta_tau"]
    with torch.no_grad():
        delta_gt = RenderFrame(vp, it.gaussians, it.pipeline, it.background, 0).sign_delta_Ir[0].cpu().numpy()
    n_frames = max(1, math.ceil(args.steps / args.iters_per_frame)) + 1
    tabs = [synth.sample_events(delta_gt, args.events, j * 50000, (j + 1) * 50000, K, d["dist"], seed=1000 + j) for j in range(n_frames)]
    arrays = [ref_runner.event_arrays_from_table(t, args.events)[0] for t in tabs]
    R0, T0, w0, v0 = perturbed_state(d, 0)
    dev = vp.device
    vp.update_RT(torch.from_numpy(R0).to(dev), torch.from_numpy(T0).to(dev))
    vp.angular_vel, vp.linear_vel = torch.from_numpy(w0).to(dev), torch.from_numpy(v0).to(dev)

    warm = max(args.warmup, 3)
    it.new_frame(arrays[0])
    it.iterate(0, min(warm, args.iters_per_frame), opt_vel=True, want_grads=False)
    torch.cuda.synchronize()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    # same structure as our e2e leg: per frame EventFrame (host numpy/OpenCV + H2D), then iterations with the
    # reference's own loss.item() read-back every iteration
    t_frames, t_iters = 0.0, 0.0
    t0 = time.perf_counter()
    left, j, frames_run = args.steps, 1, 0
    losses = []
    while left > 0:
        n = min(args.iters_per_frame, left)
        ta = time.perf_counter()
        it.new_frame(arrays[j % len(arrays)])
        tb = time.perf_counter()
        ls, _ = it.iterate(0, n, opt_vel=True, want_grads=False)   # statement for statement tracker.py:176-222, nothing added
        torch.cuda.synchronize()
        tc = time.perf_counter()
        t_frames += tb - ta
        t_iters += tc - tb
        losses += ls
        left -= n
        j += 1
        frames_run += 1
    total = time.perf_counter() - t0
    clocks = sampler.stop()
    assert len(losses) == args.steps and np.all(np.isfinite(losses))
    t_pack, t_frame_cpu = ref_runner.time_event_side(tabs[0], d, repeats=1)
    value = args.steps / total
    out = {
        "impl": "reference", "metric": "pose-track iters/sec (fwd+bwd, 640x480, 1M Gaussians)", "value": round(value, 3),
        "unit": "iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": round(1e3 * total / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2]: {args.camera}_normal1-shaped synthetic sequence, {args.gaussians}-Gaussian map, "
                               f"{d['W']}x{d['H']}, {args.events} events/frame, full-resolution fine-stage iterations",
                   "gaussians": args.gaussians, "width": d["W"], "height": d["H"], "events_per_frame": args.events,
                   "parallelism": "single process, single GPU (the reference has no multi-GPU path)",
                   "iters_per_frame_e2e": args.iters_per_frame,
                   "what_runs": "unmodified reference: RenderFrame -> tracking_loss -> backward -> Adam -> update_vwRT with its own CUDA "
                                "rasteriser (sm_100a build) and torch autograd; EventFrame on the host (numpy loop + OpenCV)"},
        "clocks": clocks,
        "e2e": {"value": round(value, 3), "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "iterations_only": {"value": round(args.steps / t_iters, 3), "unit": "iterations/s"},
        "cpu_baseline": {"value": round(value, 3), "unit": "iterations/s", "kind": "reference", "cores": os.cpu_count(),
                         "sample": f"{args.steps} iterations over {frames_run} event frames; host side per frame: EventFrame "
                                   f"{1e3 * t_frames / frames_run:.1f} ms in-loop ({1e3 * t_frame_cpu:.1f} ms device='cpu' alone), "
                                   f"event-object packetisation {1e3 * t_pack:.1f} ms/frame"},
        "gpu_launches": None,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--events", type=int, default=30000)
    ap.add_argument("--camera", choices=["robot", "desk"], default="robot")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--iters-per-frame", type=int, default=50)
    ap.add_argument("--cpu-sample-gaussians", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the engine-vs-operator-path check after the timed region")
    ap.add_argument("--hypotheses", type=int, default=64, help="seeds of the configs[3] search")
    ap.add_argument("--concurrent", type=lambda t: [int(x) for x in t.split(",")], default=[1, 4],
                    help="--mode search: hypotheses in flight per GPU (comma list = one run each)")
    ap.add_argument("--assignment", choices=["dynamic", "static"], default="dynamic")
    ap.add_argument("--no-extras", action="store_true", help="default mode: skip the configs[3] search / configs[4] tile-split sub-objects")
    ap.add_argument("--mode", choices=["hypotheses", "tilesplit", "search"], default="hypotheses",
                    help="hypotheses: one independent hypothesis per GPU (weak scaling, the driver's line); "
                         "tilesplit: one hypothesis, screen tiles split over the GPUs (strong scaling, configs[4])")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 200 if args.impl == "ours" else 60
    if args.warmup is None:
        args.warmup = 10 if args.impl == "ours" else 3
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "tilesplit":
        run_tilesplit(args)
    elif args.mode == "search":
        run_search(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
