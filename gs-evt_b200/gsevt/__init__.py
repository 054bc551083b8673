"""gsevt — B200-native (sm_100a) tracking hot path of GS-EVT behind the reference's Python surface.

Layout of this source root (put it on sys.path, exactly like the reference's repository root):
  diff_gaussian_rasterization/   drop-in operator package (GaussianRasterizer, rasterize_gaussians)
  utils/, gaussian_splatting/    host-side mirror of the reference modules on the tracking path
  gsevt/                         ctypes binding (lib), native engine handles (engine), synthetic data (synth)
  csrc/ -> ../libgsevt.so        hand-written CUDA kernels + C ABI (include/gsevt.h)
"""
from . import lib  # noqa: F401

__version__ = "0.1.0"
