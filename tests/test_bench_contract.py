"""CPU tests of bench.py's pure parts: the roofline object of the measurement contract (keys, units, the byte / flop
models of DESIGN.md section 3 applied to known counters), the peak table, and the command-line defaults the driver
relies on (N = 1, a K / W that finish within minutes)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

STAGES = {"preprocess_map": 0.073, "bucket_scatter": 0.053, "bucket_sort": 0.070, "blend_fwd_gray": 0.175, "loss_stats": 0.013,
          "blend_bwd_gray": 0.260, "geom_bwd_pose": 0.040, "engine_update": 0.018}
WL = dict(visible=[684694, 680251], instances=[3210677, 3196205], pairs_walked=[76125691, 77018806], gaussians_with_grad=66362,
          sorted_slots=3759552)


def test_roofline_object_follows_the_contract():
    roof, table = bench.roofline(dict(STAGES), WL, 1_000_000, 640 * 480, {"sm_mhz": 1965.0}, default_workload=True)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof, k
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["kernel"] == "preprocess_map"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-3
    # SURVEY 8(d) row A applied to the bench scene: 44 P + 192 max(visible) + 32 Pv + 4 P
    assert roof["alg_bytes"] == 44 * 1_000_000 + 192 * 684694 + 32 * (684694 + 680251) + 4 * 1_000_000
    assert abs(roof["achieved"] - roof["alg_bytes"] / 0.073e-3 / 1e9) < 1.0
    assert roof["top_stage"] == "blend_bwd_gray" and roof["top_stage_fp32"]["unit"] == "TFLOP/s"
    fp32_peak = 148 * 128 * 2 * 1965.0e6 / 1e12
    assert abs(roof["fp32_peak_tflops_at_clock"] - fp32_peak) < 0.01
    assert abs(roof["top_stage_fp32"]["frac"] - 100.0 * sum(WL["pairs_walked"]) / 0.260e-3 / 1e12 / fp32_peak) < 1e-3
    # the committed ncu capture feeds `traffic` (bytes per launch of the same kernel) for the default workload only
    if os.path.exists(os.path.join(ROOT, "profiles", "ncu_traffic.json")):
        assert roof["traffic"] and roof["traffic"] > 0.5 * roof["alg_bytes"] and "profiles/" in roof["traffic_source"]
    other, _ = bench.roofline(dict(STAGES), WL, 1_000_000, 640 * 480, {"sm_mhz": 1965.0}, default_workload=False)
    assert other["traffic"] is None
    assert set(table) == set(STAGES) and table["blend_fwd_gray"]["bound"] == "fp32" and "frac" not in table["engine_update"]
    json.dumps(roof), json.dumps(table)       # serialisable as they are


def test_a_stage_folded_into_its_neighbour_gets_no_rate():
    st = dict(STAGES, loss_stats=0.0003)        # e.g. GSEVT_FUSE_LOSS=1: the launch is gone, the profile reports ~0
    _, table = bench.roofline(st, WL, 1_000_000, 640 * 480, None)
    assert "frac" not in table["loss_stats"] and table["loss_stats"]["ms"] == 0.0003


def test_measured_peaks_are_used_when_present():
    hbm, sm_max, how = bench.load_peaks()
    assert hbm > 1000 and sm_max > 1000
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        assert "MEASURED_PEAKS" in how and abs(hbm - float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])) < 1e-6
    else:
        assert "fallback" in how


def test_command_line_defaults(monkeypatch):
    seen = {}
    monkeypatch.setattr(bench, "run_ours", lambda a: seen.update(vars(a), which="ours"))
    monkeypatch.setattr(bench, "run_reference", lambda a: seen.update(vars(a), which="reference"))
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    bench.main()
    assert seen["which"] == "ours" and seen["gpus"] == 1 and seen["steps"] == 200 and seen["warmup"] == 10
    assert (seen["gaussians"], seen["width"], seen["height"], seen["events"]) == (1_000_000, 640, 480, 30000)   # BASELINE configs[2]
    assert seen["hypotheses"] == 64 and seen["mode"] == "hypotheses" and not seen["no_extras"]
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "4", "--steps", "7", "--warmup", "3"])
    bench.main()
    assert seen["which"] == "reference" and seen["gpus"] == 4 and seen["steps"] == 7 and seen["warmup"] == 3
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference"])
    bench.main()
    assert seen["steps"] == 60 and seen["warmup"] == 3       # a bounded sample: the reference needs 17-19 ms per step


def test_reference_arm_reports_unavailable_without_the_build(monkeypatch, capsys):
    from oracle import ref_runner
    monkeypatch.setattr(ref_runner, "available", lambda: False)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line
    monkeypatch.setenv("RANK", "1")            # under torchrun only rank 0 runs and prints the reference arm
    bench.main()
    assert capsys.readouterr().out.strip() == ""
