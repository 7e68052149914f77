"""rt_b200 -- B200-native (sm_100a) path-tracing hot path behind marzer/rt's renderer plugin interface.

Layout: csrc/ (CUDA kernels + the C ABI of include/rtcu.h), _native (ctypes binding), scene (containers +
TOML loader restating scene.cpp), camera (viewport matrices), renderer (renderer_interface mirror,
registry, Context), synth (synthetic benchmark scenes), dist (multi-GPU sample-range partition).
Importing the package does not import torch and does not load the CUDA library; using it does, and
fails loudly when the library or a B200-class device is missing (there is no CPU fallback).
"""
from .scene import Scene, Camera, SceneError, load as load_scene, loads as loads_scene  # noqa: F401
from .renderer import (Context, ImageView, RendererInterface, Description, renderers, register_renderer,  # noqa: F401
                       cuda_path_tracer, CudaPathTracer, ProgressiveRenderer, make_view, render_multi, DEFAULT_SEED)
from ._native import (RtcuError, MODE_MG, MODE_SM, ACCEL_AUTO, ACCEL_LINEAR, ACCEL_BVH,  # noqa: F401
                      PIPE_AUTO, PIPE_MEGAKERNEL, PIPE_WAVEFRONT, PRIM_MISS, PRIM_PLANE)

__version__ = "0.1.0"
