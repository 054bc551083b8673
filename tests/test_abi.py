"""C-ABI surface: libgsevt.so loads without a GPU, exports every symbol include/gsevt.h declares, and
the ctypes mirrors of the ABI structs have the C compiler's layout.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsevt.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"GSEVT_API\s+[\w\s\*]+?\b(gsevt_\w+)\s*\(", src)))


def test_header_declares_expected_groups():
    syms = declared_symbols()
    assert len(syms) >= 35
    for must in ("gsevt_raster_forward_geometry", "gsevt_raster_forward_render", "gsevt_raster_backward",
                 "gsevt_mark_visible", "gsevt_event_accumulate", "gsevt_event_frame", "gsevt_engine_iterate",
                 "gsevt_engine_eval", "gsevt_map_create"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    from gsevt import lib
    L = C.CDLL(lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/gsevt.h but not exported: {missing}"
    # and the Python prototype table covers the header one to one
    assert sorted(lib.PROTOTYPES) == declared_symbols()


def test_no_unexpected_exports(built):
    from gsevt import lib
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    ours = {s for s in exported if not s.startswith("_") and "cudart" not in s}
    assert ours <= set(declared_symbols()), ours - set(declared_symbols())


def test_struct_layouts_match_the_c_compiler(built, tmp_path):
    from gsevt import lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gsevt.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(GsevtRasterArgs), offsetof(GsevtRasterArgs, background), offsetof(GsevtRasterArgs, pose_grads),"
                    "sizeof(GsevtEngineConfig), sizeof(GsevtEngineStatus), offsetof(GsevtEngineStatus, pose_grads));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(lib.GsevtRasterArgs), lib.GsevtRasterArgs.background.offset, lib.GsevtRasterArgs.pose_grads.offset,
            C.sizeof(lib.GsevtEngineConfig), C.sizeof(lib.GsevtEngineStatus), lib.GsevtEngineStatus.pose_grads.offset]
    assert got == want


def test_host_only_entry_points(built):
    """Entry points that need no device: version, size queries, argument validation, the undistort map."""
    import numpy as np
    from gsevt import lib
    L = built
    assert L.gsevt_abi_version() == 1
    gb, ib = C.c_size_t(0), C.c_size_t(0)
    assert L.gsevt_raster_sizes(1000, 640, 480, C.byref(gb), C.byref(ib)) == 0
    assert gb.value >= 1000 * (32 + 16 + 24 + 4 + 1 + 4 + 4) and ib.value >= 640 * 480 * 8 + 1200 * 8
    assert L.gsevt_raster_sizes(-1, 640, 480, C.byref(gb), C.byref(ib)) < 0
    assert b"bad sizes" in L.gsevt_last_error()
    assert L.gsevt_raster_binning_size(100000) >= 100000 * 24
    assert L.gsevt_raster_geom_offset(b"rec", 1000) > 0 and L.gsevt_raster_geom_offset(b"nope", 1000) == -1
    # argument validation mirrors the reference's exceptions (dgr/.../__init__.py:229-233)
    a = lib.GsevtRasterArgs()
    a.P, a.width, a.height = 10, 64, 48
    a.means3D = 1  # non-null
    assert L.gsevt_raster_forward_geometry(C.byref(a), None) == -1
    assert b"SHs or precomputed colors" in L.gsevt_last_error()
    a.shs = 1
    a.sh_coeffs, a.sh_degree = 16, 3
    assert L.gsevt_raster_forward_geometry(C.byref(a), None) == -1
    assert b"scale/rotation pair or precomputed 3D covariance" in L.gsevt_last_error()
    # the fixed-point undistort map is host code: check it against the numpy oracle
    from oracle import event_oracle as eo
    K = np.array([327.3, 0, 305.0, 0, 327.5, 235.4, 0, 0, 1.0])
    D = np.array([-0.031982, 0.041966, -0.000507, -0.001031, 0.0])
    ix = np.zeros((48, 64), np.int32)
    iy = np.zeros((48, 64), np.int32)
    assert L.gsevt_event_undistort_map(K.ctypes.data_as(C.POINTER(C.c_double)), D.ctypes.data_as(C.POINTER(C.c_double)), 64, 48,
                                       ix.ctypes.data, iy.ctypes.data) == 0
    ex, ey = eo.undistort_map(K.reshape(3, 3), D, 64, 48)
    assert np.array_equal(ix, ex) and np.array_equal(iy, ey)


def test_product_fails_loudly_without_gpu(built):
    import torch
    from gsevt import lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.GsevtError):
        lib.require_device()
    import diff_gaussian_rasterization as dgr
    s = dgr.GaussianRasterizationSettings(48, 64, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), torch.eye(4), 3,
                                          torch.zeros(3), False, torch.zeros(3), torch.zeros(3), torch.eye(4), torch.eye(4), 0.0, False)
    r = dgr.GaussianRasterizer(s)
    with pytest.raises(Exception):
        r(means3D=torch.zeros(4, 3), means2D=torch.zeros(4, 3), opacities=torch.ones(4, 1), shs=torch.zeros(4, 16, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_product_never_imports_the_oracle():
    """The product path (package sources + csrc) must not reference oracle/ at all."""
    pkg = os.path.join(ROOT, "gs-evt_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "liboracle" in txt or "oracle/" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_reference_install_is_not_tracked():
    """oracle/_ref holds the unmodified reference (built by oracle/build_ref.sh): it travels to the GPU box with the snapshot
    but must never enter the history."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, ".git")):
        pytest.skip("not a git checkout")
    out = subprocess.run(["git", "ls-files", "oracle/_ref", "baseline/_ref"], cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "", out.stdout


def test_logger_follows_the_log_file(tmp_path):
    """Two Trackers in one process with different save paths: each gets its own tracking_log.log."""
    from utils.auxiliary import Logger
    a, b = tmp_path / "a" / "tracking_log.log", tmp_path / "b" / "tracking_log.log"
    Logger(name="T_gsevt_test", log_file=str(a)).info("first")
    Logger(name="T_gsevt_test", log_file=str(b)).info("second")
    assert "first" in a.read_text() and "second" in b.read_text() and "second" not in a.read_text()
