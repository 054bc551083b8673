"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference pipeline (oracle/_ref/pipeline = the reference's
utils/, gaussian_splatting/, main.py, configs/ as installed by oracle/build_ref.sh; oracle/_ref/ext = its
rasteriser extension built for sm_100a) through the reference's own classes.  Used by
  * bench.py --impl reference   (the timed reference arm), and
  * tests/ + tests/golden/make_golden.py (parity vectors, trajectories),
never by the product.  Because the reference's module names (utils, gaussian_splatting,
diff_gaussian_rasterization) are the ones the product mirrors, the reference must be imported in a
process whose sys.path puts oracle/_ref first: call `activate()` before anything else, or run this
file as a script (subprocess) — `python oracle/ref_runner.py <mode> --in x.npz --out y.npz`.

What is driven, statement for statement as in the reference's Tracker.tracking inner loop
(utils/tracker.py:176-240): RenderFrame(viewpoint, gaussians, pipeline, background, lvl) ->
Tracker.tracking_loss -> loss.backward() -> loss.item() -> optimizer.step() -> check_convergence ->
update_pose / update_vwRT -> optimizer.zero_grad().
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    ext = os.path.join(REF, "ext", "diff_gaussian_rasterization")
    return (os.path.isfile(os.path.join(REF, "pipeline", "utils", "tracker.py")) and os.path.isdir(ext)
            and any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(ext)))


def _load_by_path(name, path):
    """Imports the package at `path` under `name` without touching sys.path."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def activate(operator="reference"):
    """Puts the reference (and the shims for munch/plyfile/natsort/open3d/rosbag) first on sys.path.

    operator="reference": the reference's own extension (oracle/_ref/ext).
    operator="ours": INTEGRATION.md section 1(b) — the reference's UNMODIFIED Python pipeline with only the rasteriser
    operator replaced: `diff_gaussian_rasterization` resolves to this repo's drop-in package (and its ctypes binding
    `gsevt`), loaded by file path because the product's source root must stay off sys.path here (its `utils` package would
    shadow the reference's namespace package of the same name)."""
    product = os.path.join(os.path.dirname(HERE), "gs-evt_b200")
    for name in ("utils", "gaussian_splatting", "diff_gaussian_rasterization"):
        m = sys.modules.get(name)
        if m is not None:
            src = getattr(m, "__file__", None) or (list(getattr(m, "__path__", [])) or [""])[0]
            if not os.path.abspath(src).startswith(os.path.abspath(REF)):
                raise RuntimeError(f"module '{name}' was already imported from the product; run the reference in its own process")
    # The reference's `utils` has no __init__.py (namespace package): any regular `utils` package anywhere
    # on sys.path would shadow it, so the product's source root must not be importable in this process.
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != product]
    paths = [os.path.join(REF, "pipeline"), os.path.join(HERE, "shims")]
    if operator == "reference":
        paths.insert(0, os.path.join(REF, "ext"))
    for p in paths:
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if operator == "ours":
        _load_by_path("gsevt", os.path.join(product, "gsevt"))
        _load_by_path("diff_gaussian_rasterization", os.path.join(product, "diff_gaussian_rasterization"))
    elif operator != "reference":
        raise ValueError(operator)
    import importlib
    for name in ("utils.tracker", "utils.event_camera.event", "gaussian_splatting", "diff_gaussian_rasterization"):
        m = importlib.import_module(name)
        src = os.path.abspath(getattr(m, "__file__", None) or list(m.__path__)[0])
        want = os.path.abspath(product if (operator == "ours" and name == "diff_gaussian_rasterization") else REF)
        if not src.startswith(want):
            raise RuntimeError(f"'{name}' resolved to {src}, not under {want}")


def make_config(d, save_path, map_path="", events_path="", device="cuda"):
    """yaml-shaped dict (schema of configs/VECTOR/*.yaml) from a plain description dict."""
    return {
        "Event": {"data_path": events_path, "distortion_factors": list(d["dist"]), "filter_threshold": 0,
                  "img_height": d["H"], "img_width": d["W"],
                  "intrinsic": {"cols": 3, "rows": 3, "dt": "d", "data": [d["fx"], 0.0, d["cx"], 0.0, d["fy"], d["cy"], 0.0, 0.0, 1.0]},
                  "gaussian_kernel_size": 9, "max_events_per_frame": d.get("max_events_per_frame", 30000)},
        "Gaussian": {"calib_params": {"fx": d["fx"], "fy": d["fy"]},
                     "model_params": {"background": list(d.get("background", [0, 0, 0])), "device": device, "model_path": map_path, "sh_degree": 3},
                     "pipeline_params": {"compute_cov3D_python": False, "convert_SHs_python": False},
                     "img_height": d["H"], "img_width": d["W"]},
        "Optimizer": {"cam_rot_delta": d["lr"]["cam_rot_delta"], "cam_trans_delta": d["lr"]["cam_trans_delta"],
                      "cam_v_delta": d["lr"]["cam_v_delta"], "cam_w_delta": d["lr"]["cam_w_delta"],
                      "converged_threshold": d.get("converged_threshold", 1e-4), "max_optim_iter": d.get("max_optim_iter", 200)},
        "Tracking": {"initial_pose": {"rot": {"cols": 3, "rows": 3, "dt": "d", "data": [float(x) for x in d["R"]]},
                                      "trans": {"cols": 1, "rows": 3, "dt": "d", "data": [float(x) for x in d["T"]]}},
                     "initial_vel": {"angular_vel": [float(x) for x in d["angular_vel"]], "linear_vel": [float(x) for x in d["linear_vel"]]},
                     "save_path": save_path},
    }


def load_map(raw, sh_degree=3):
    """Reference GaussianModel filled with raw (pre-activation) parameters — the same tensors
    load_ply (gaussian_model.py:298-323) would create, without a PLY round trip."""
    import torch
    from torch import nn
    from gaussian_splatting.scene.gaussian_model import GaussianModel
    g = GaussianModel(sh_degree)
    mk = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float, device="cuda").contiguous().requires_grad_(False))
    g._xyz, g._scaling, g._rotation, g._opacity = mk(raw["xyz"]), mk(raw["scaling"]), mk(raw["rotation"]), mk(raw["opacity"])
    g._features_dc, g._features_rest = mk(raw["f_dc"]), mk(raw["f_rest"])
    g.active_sh_degree = g.max_sh_degree
    return g


def event_arrays_from_table(table, max_events_per_frame):
    """Reference EventArray objects from an (n,4) int table [ts x y p] — same packetisation as
    load_events_from_txt (event.py:20-37) without the text round trip."""
    from utils.event_camera.event import Event, EventArray
    out, cur = [], EventArray()
    for ts, x, y, p in table.tolist():
        cur.callback(Event(x=x, y=y, ts=ts, polarity=p))
        if cur.size() >= max_events_per_frame:
            out.append(cur)
            cur = EventArray()
    return out


class RefIterations:
    """The reference's inner optimisation loop on one event frame at one pyramid level."""

    def __init__(self, desc, raw_map, save_path="/tmp/gsevt_ref_run"):
        import torch
        from munch import munchify
        from utils.render_camera.camera import Camera
        from utils.tracker import Tracker
        os.makedirs(save_path, exist_ok=True)
        self.torch = torch
        self.config = make_config(desc, save_path)
        self.viewpoint = Camera.init_from_yaml(self.config)
        self.gaussians = load_map(raw_map)
        self.pipeline = munchify(self.config["Gaussian"]["pipeline_params"])
        self.background = torch.tensor(self.config["Gaussian"]["model_params"]["background"], dtype=torch.float32, device="cuda")
        self.tracker = Tracker(self.config, [], self.viewpoint, self.gaussians, self.pipeline, self.background, "cuda")
        self.optimizer = None
        self.record_states, self.states = False, []

    def new_frame(self, event_array, delta_tau=None):
        """tracker.py:117-147: fresh Adam, delta_tau, EventFrame + pyramids (host-side numpy/OpenCV)."""
        torch, vp, cfg = self.torch, self.viewpoint, self.config["Optimizer"]
        from utils.event_camera.event import EventFrame
        self.optimizer = torch.optim.Adam([{"params": [vp.cam_rot_delta], "lr": cfg["cam_rot_delta"]},
                                           {"params": [vp.cam_trans_delta], "lr": cfg["cam_trans_delta"]},
                                           {"params": [vp.cam_w_delta], "lr": cfg["cam_w_delta"]},
                                           {"params": [vp.cam_v_delta], "lr": cfg["cam_v_delta"]}])
        vp.delta_tau = event_array.duration() if delta_tau is None else delta_tau
        t = self.tracker
        eFrame = EventFrame(t.img_width, t.img_height, t.intrinsic, t.distortion_factors, t.gaussian_kernel_size, event_array)
        self.sign_pyr = t.image_pyramid(eFrame.sign_delta_Ie)
        self.unsign_pyr = t.image_pyramid(eFrame.unsign_delta_Ie)
        self.eFrame = eFrame

    def iterate(self, lvl, n, opt_vel=True, k0=0, step=True, want_grads=True):
        """n iterations of tracker.py:176-222 at level `lvl` (fine stage when opt_vel).  k0 = optim_iter -
        start_vel_opt_iter of the first one (drives the LR cross-fade).  Returns per-iteration losses and
        the gradients [rho, theta, v, w] of each iteration (want_grads=False: no gradient read-back — the reference's loop has
        none, so the timed reference arm of bench.py runs without it)."""
        from utils.render_camera.frame import RenderFrame
        torch, vp, cfg, t = self.torch, self.viewpoint, self.config["Optimizer"], self.tracker
        fraction_num = t.max_optim_iter / 2
        losses, grads = [], []
        for i in range(n):
            vp.cam_w_delta.requires_grad_(bool(opt_vel))
            vp.cam_v_delta.requires_grad_(bool(opt_vel))
            vp.cam_rot_delta.requires_grad_(True)
            vp.cam_trans_delta.requires_grad_(True)
            if opt_vel:
                k = k0 + i
                fraction = k / fraction_num if 1 <= k <= fraction_num else 1
                for g in self.optimizer.param_groups:
                    p = g["params"][0]
                    if p is vp.cam_rot_delta:
                        g["lr"] = cfg["cam_rot_delta"] * fraction
                    if p is vp.cam_trans_delta:
                        g["lr"] = cfg["cam_trans_delta"] * fraction
                    if p is vp.cam_w_delta:
                        g["lr"] = cfg["cam_w_delta"] * (1 - fraction)
                    if p is vp.cam_v_delta:
                        g["lr"] = cfg["cam_v_delta"] * (1 - fraction)
            rFrame = RenderFrame(vp, self.gaussians, self.pipeline, self.background, lvl)
            if not opt_vel:
                loss = t.tracking_loss(rFrame.unsign_delta_Ir, self.unsign_pyr[lvl], huber=False)
            else:
                loss = t.tracking_loss(rFrame.sign_delta_Ir, self.sign_pyr[lvl], huber=False)
            loss.backward()
            losses.append(loss.item())
            if want_grads:
                z = torch.zeros(3, device="cuda")
                gg = [vp.cam_trans_delta.grad, vp.cam_rot_delta.grad, vp.cam_v_delta.grad, vp.cam_w_delta.grad]
                grads.append(torch.cat([z if x is None else x.detach().reshape(-1) for x in gg]).cpu().numpy())
            with torch.no_grad():
                if step:
                    self.optimizer.step()
                    t.check_convergence(losses, t.converged_threshold)   # tracker.py:217 (its result is not acted on here)
                    if not opt_vel:
                        vp.update_pose()
                    else:
                        vp.update_vwRT()
                self.optimizer.zero_grad()
            self.rFrame = rFrame
            if self.record_states:
                self.states.append(np.concatenate([x.reshape(-1) for x in self.state()]))
        return losses, grads

    def state(self):
        vp = self.viewpoint
        return (vp.R.detach().cpu().numpy(), vp.T.detach().cpu().numpy(), vp.angular_vel.detach().cpu().numpy(),
                vp.linear_vel.detach().cpu().numpy())


def run_tracker(desc, raw_map, table, save_path):
    """The reference's whole Tracker.tracking() (unmodified) on in-memory synthetic inputs.  Returns the
    TUM trajectory rows and the per-level iteration counts parsed from its own log."""
    import re
    import numpy as np
    import torch
    from munch import munchify
    from utils.render_camera.camera import Camera
    from utils.tracker import Tracker
    os.makedirs(save_path, exist_ok=True)
    config = make_config(desc, save_path)
    viewpoint = Camera.init_from_yaml(config)
    gaussians = load_map(raw_map)
    pipeline = munchify(config["Gaussian"]["pipeline_params"])
    background = torch.tensor(config["Gaussian"]["model_params"]["background"], dtype=torch.float32, device="cuda")
    arrays = event_arrays_from_table(table, config["Event"]["max_events_per_frame"])
    tracker = Tracker(config, arrays, viewpoint, gaussians, pipeline, background, "cuda")
    t0 = time.time()
    try:
        tracker.tracking()
    except Exception as e:  # the video/gif tail of tracking() may fail without codecs; the TUM file is complete by then
        print("[ref_runner] tracking() tail raised:", repr(e))
    wall = time.time() - t0
    tum = np.loadtxt(os.path.join(save_path, "tracking_pose_tum.txt"), ndmin=2)
    log = open(os.path.join(save_path, "tracking_log.log")).read()
    iters = [(int(a), int(b), int(c)) for a, b, c in re.findall(r"optim_iter:\s+(\d+) \((\d+)\+(\d+)\)", log)]
    times = [float(x) for x in re.findall(r"opt_time:\s+([0-9.]+)", log)]
    return dict(tum=tum, iters=np.array(iters), opt_time=np.array(times), wall=wall)


def time_event_side(table, desc, repeats=5):
    """CPU side of the reference: per-event Python objects + EventFrame (numpy loop + OpenCV), event.py:11-128.
    Returns seconds per frame for packetisation and for EventFrame(device='cpu')."""
    import numpy as np
    from utils.event_camera.event import EventFrame
    n = desc.get("max_events_per_frame", 30000)
    t0 = time.perf_counter()
    arrays = event_arrays_from_table(table[:n * max(1, min(repeats, table.shape[0] // n))], n)
    t_pack = (time.perf_counter() - t0) / max(1, len(arrays))
    K = np.array([desc["fx"], 0.0, desc["cx"], 0.0, desc["fy"], desc["cy"], 0.0, 0.0, 1.0]).reshape(3, 3)
    D = np.array(desc["dist"])
    ts = []
    for a in arrays:
        t0 = time.perf_counter()
        EventFrame(desc["W"], desc["H"], K, D, 9, a, device="cpu")
        ts.append(time.perf_counter() - t0)
    return t_pack, float(np.median(ts))


def _main():
    import argparse
    import numpy as np
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["iterations", "tracker"])
    ap.add_argument("--inp", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--operator", choices=["reference", "ours"], default="reference",
                    help="ours: the reference's Python with this repo's drop-in diff_gaussian_rasterization (INTEGRATION.md 1b)")
    a = ap.parse_args()
    activate(a.operator)
    z = np.load(a.inp, allow_pickle=True)
    desc = z["desc"].item()
    raw = {k: z[k] for k in ("xyz", "scaling", "rotation", "opacity", "f_dc", "f_rest")}
    table = z["events"]
    if a.mode == "tracker":
        r = run_tracker(desc, raw, table, desc.get("save_path", "/tmp/gsevt_ref_track"))
        np.savez(a.out, **r)
        return
    it = RefIterations(desc, raw)
    it.record_states = True
    arrays = event_arrays_from_table(table, desc.get("max_events_per_frame", 30000))
    if "start" in desc:   # start state different from the yaml pose (perturbed hypothesis)
        import torch
        R0, T0, w0, v0 = (np.asarray(x, np.float32) for x in desc["start"])
        vp = it.viewpoint
        vp.update_RT(torch.from_numpy(R0.reshape(3, 3)).to(vp.device), torch.from_numpy(T0).to(vp.device))
        vp.angular_vel, vp.linear_vel = torch.from_numpy(w0).to(vp.device), torch.from_numpy(v0).to(vp.device)
    it.new_frame(arrays[0], desc.get("delta_tau"))
    out = {}
    for lvl, opt_vel, n in desc["plan"]:
        losses, grads = it.iterate(int(lvl), int(n), bool(opt_vel), k0=0, step=bool(desc.get("step", True)))
        out[f"loss_L{lvl}_{int(opt_vel)}"] = np.array(losses)
        out[f"grad_L{lvl}_{int(opt_vel)}"] = np.array(grads)
    R, T, w, v = it.state()
    out.update(R=R, T=T, w=w, v=v, sign_Ie=it.eFrame.sign_delta_Ie.cpu().numpy(), states=np.array(it.states))
    np.savez(a.out, **out)


if __name__ == "__main__":
    _main()
