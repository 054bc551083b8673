"""ctypes binding of libgsevt.so — the C ABI declared in include/gsevt.h.

The product path has NO fallback: if the shared library is missing, or the GPU is not sm_100, every
entry point raises.  PyTorch is used only for device memory and streams; tensors cross the boundary
as raw pointers (tensor.data_ptr()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libgsevt.so")

c_float_p = C.POINTER(C.c_float)
c_void_p = C.c_void_p


class GsevtRasterArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float), ("delta_time", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("want_n_touched", C.c_int32), ("reserved0", C.c_int32),
        ("background", c_void_p), ("means3D", c_void_p), ("shs", c_void_p), ("colors_precomp", c_void_p),
        ("opacities", c_void_p), ("scales", c_void_p), ("rotations", c_void_p), ("cov3D_precomp", c_void_p),
        ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("projmatrix_raw", c_void_p), ("campos", c_void_p),
        ("vel_transform", c_void_p), ("vel_transform_inv", c_void_p),
        ("geom_buffer", c_void_p), ("geom_bytes", C.c_size_t),
        ("img_buffer", c_void_p), ("img_bytes", C.c_size_t),
        ("binning_buffer", c_void_p), ("binning_bytes", C.c_size_t),
        ("out_color", c_void_p), ("out_depth", c_void_p), ("out_opacity", c_void_p),
        ("radii", c_void_p), ("n_touched", c_void_p),
        ("dL_dout_color", c_void_p), ("dL_dout_depth", c_void_p),
        ("num_rendered", C.c_int32), ("reserved1", C.c_int32),
        ("pose_grads", c_void_p),
        ("bwd_workspace", c_void_p), ("bwd_workspace_bytes", C.c_size_t),
        ("dL_dmeans2D", c_void_p), ("dL_dmeans3D", c_void_p), ("dL_dopacity", c_void_p), ("dL_dcolors", c_void_p),
        ("dL_dcov3D", c_void_p), ("dL_dsh", c_void_p), ("dL_dscales", c_void_p), ("dL_drotations", c_void_p),
        ("dL_dtau", c_void_p), ("dL_dvel", c_void_p),
    ]


class GsevtEngineConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("levels", C.c_int32),
        ("fx", C.c_float), ("fy", C.c_float), ("znear", C.c_float), ("zfar", C.c_float),
        ("background", C.c_float * 3),
        ("lr_rot", C.c_float), ("lr_trans", C.c_float), ("lr_w", C.c_float), ("lr_v", C.c_float),
        ("converged_threshold", C.c_float), ("max_optim_iter", C.c_int32), ("instance_capacity", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class GsevtEngineStatus(C.Structure):
    _fields_ = [
        ("level_done", C.c_int32), ("optim_iter", C.c_int32), ("start_vel_opt_iter", C.c_int32),
        ("opt_vel", C.c_int32), ("iters_executed", C.c_int32), ("overflow", C.c_int32),
        ("num_rendered", C.c_int32 * 2), ("last_loss", C.c_float), ("pose_grads", C.c_float * 12),
        ("reserved", C.c_float * 4),
    ]


# name -> (restype, argtypes); mirrors include/gsevt.h one to one (tests check the export list).
PROTOTYPES = {
    "gsevt_last_error": (C.c_char_p, []),
    "gsevt_abi_version": (C.c_int, []),
    "gsevt_device_arch": (C.c_int, []),
    "gsevt_raster_sizes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gsevt_raster_binning_size": (C.c_size_t, [C.c_int32]),
    "gsevt_raster_backward_workspace_size": (C.c_size_t, [C.c_int32]),
    "gsevt_raster_forward_geometry": (C.c_int, [C.POINTER(GsevtRasterArgs), c_void_p]),
    "gsevt_raster_forward_render": (C.c_int, [C.POINTER(GsevtRasterArgs), c_void_p]),
    "gsevt_raster_backward": (C.c_int, [C.POINTER(GsevtRasterArgs), c_void_p]),
    "gsevt_mark_visible": (C.c_int, [C.c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gsevt_raster_geom_offset": (C.c_int64, [C.c_char_p, C.c_int32]),
    "gsevt_raster_binning_offset": (C.c_int64, [C.c_char_p, C.c_int32]),
    "gsevt_raster_img_offset": (C.c_int64, [C.c_char_p, C.c_int32, C.c_int32]),
    "gsevt_event_accumulate": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int32, C.c_int32, C.c_int32, c_void_p,
                                         C.c_int32, c_void_p, c_void_p]),
    "gsevt_event_undistort_map": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_int32,
                                            c_void_p, c_void_p]),
    "gsevt_event_frame": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int32, C.c_int32, C.c_int32, c_void_p, c_void_p,
                                    c_void_p, C.c_size_t, c_void_p]),
    "gsevt_event_frame_k": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_void_p, c_void_p,
                                      c_void_p, C.c_size_t, c_void_p]),
    "gsevt_event_frame_scratch_size": (C.c_size_t, [C.c_int32, C.c_int32]),
    "gsevt_map_create": (C.c_int, [C.c_int32, C.c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_float,
                                   c_void_p, C.POINTER(c_void_p)]),
    "gsevt_map_destroy": (None, [c_void_p]),
    "gsevt_map_size": (C.c_int32, [c_void_p]),
    "gsevt_map_bytes": (C.c_size_t, [c_void_p]),
    "gsevt_engine_create": (C.c_int, [c_void_p, C.POINTER(GsevtEngineConfig), C.POINTER(c_void_p)]),
    "gsevt_engine_destroy": (None, [c_void_p]),
    "gsevt_engine_set_state": (C.c_int, [c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, c_void_p]),
    "gsevt_engine_get_state": (C.c_int, [c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, c_void_p]),
    "gsevt_engine_begin_frame": (C.c_int, [c_void_p, C.c_double, c_void_p, c_void_p, c_void_p]),
    "gsevt_engine_begin_level": (C.c_int, [c_void_p, C.c_int32, C.c_int32, c_void_p]),
    "gsevt_engine_iterate": (C.c_int, [c_void_p, C.c_int32, c_void_p]),
    "gsevt_engine_poll_done": (C.c_int, [c_void_p]),
    "gsevt_engine_resume": (C.c_int, [c_void_p, c_void_p]),
    "gsevt_engine_status": (C.c_int, [c_void_p, C.POINTER(GsevtEngineStatus), c_void_p]),
    "gsevt_engine_losses": (C.c_int, [c_void_p, c_float_p, C.c_int32, c_void_p]),
    "gsevt_engine_const_vel_model": (C.c_int, [c_void_p, C.c_double, c_void_p]),
    "gsevt_engine_weighted_velocity": (C.c_int, [c_void_p, c_float_p, c_float_p, C.c_double, C.c_double, c_void_p]),
    "gsevt_engine_render_delta": (C.c_int, [c_void_p, C.c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gsevt_engine_eval": (C.c_int, [c_void_p, C.c_int32, C.c_int32, c_float_p, c_float_p, c_void_p]),
    "gsevt_engine_image_state": (C.c_int, [c_void_p, C.c_int32, c_void_p, c_void_p, c_void_p]),
    "gsevt_engine_view_params": (C.c_int, [c_void_p, C.c_int32, C.POINTER(C.c_float), c_void_p]),
    "gsevt_engine_launches_per_iteration": (C.c_int, [c_void_p]),
    "gsevt_engine_set_binning": (C.c_int, [c_void_p, C.c_int32]),
    "gsevt_parse_int_table": (C.c_int64, [c_void_p, C.c_size_t, c_void_p, C.c_size_t, C.c_int32]),
    "gsevt_engine_binning": (C.c_int, [c_void_p, C.c_int32, c_void_p, c_void_p, c_void_p, C.c_int32, c_void_p]),
    "gsevt_engine_stage_count": (C.c_int, []),
    "gsevt_engine_stage_name": (C.c_char_p, [C.c_int32]),
    "gsevt_engine_profile": (C.c_int, [c_void_p, C.c_int32, c_float_p, c_void_p]),
    "gsevt_engine_workload": (C.c_int, [c_void_p, C.POINTER(C.c_int64), c_void_p]),
    "gsevt_engine_split_mailbox": (C.c_int, [c_void_p, C.POINTER(c_void_p)]),
    "gsevt_split_mailbox_bytes": (C.c_size_t, []),
    "gsevt_ipc_export": (C.c_int, [c_void_p, C.POINTER(C.c_uint8)]),
    "gsevt_ipc_open": (C.c_int, [C.POINTER(C.c_uint8), C.POINTER(c_void_p)]),
    "gsevt_ipc_close": (C.c_int, [c_void_p]),
    "gsevt_engine_split_attach": (C.c_int, [c_void_p, C.c_int32, C.c_int32, C.POINTER(c_void_p), C.c_double]),
    "gsevt_engine_split_info": (C.c_int, [c_void_p, C.POINTER(C.c_int32), c_void_p]),
    "gsevt_split_balance_rows": (C.c_int, [C.POINTER(C.c_uint32), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
}

_lib = None
_ARCH = {}


class GsevtError(RuntimeError):
    pass


def load():
    """Loads libgsevt.so (once).  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsevtError(
            f"{LIB_PATH} not found. Build it with `make -C gs-evt_b200/csrc` (or __graft_entry__.build()); "
            "gsevt has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    """Turns a negative return code into a RuntimeError carrying gsevt_last_error()."""
    if rc < 0:
        msg = load().gsevt_last_error()
        raise GsevtError(f"{what}: {msg.decode() if msg else 'error'} (code {rc})")
    return rc


def require_device():
    """Fails loudly unless the current CUDA device is a B200-class (sm_100) GPU."""
    import torch
    if not torch.cuda.is_available():
        raise GsevtError("gsevt needs a CUDA device (sm_100a); no CPU fallback exists")
    dev = torch.cuda.current_device()
    arch = _ARCH.get(dev)
    if arch is None:  # cudaGetDeviceProperties is slow (milliseconds): ask once per device
        arch = _ARCH[dev] = check(load().gsevt_device_arch(), "gsevt_device_arch")
    if arch // 10 != 10:
        raise GsevtError(f"libgsevt.so only carries sm_100a code; current device is sm_{arch}")
    return arch


def ptr(t):
    """Device (or host) pointer of a tensor; None / empty tensors map to NULL like the reference's
    `torch.Tensor([])` arguments (dgr/diff_gaussian_rasterization/__init__.py:235-253)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
