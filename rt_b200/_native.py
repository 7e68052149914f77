"""ctypes binding of rt_b200/lib/librtcu.so (the C ABI in include/rtcu.h).

There is no fallback: if the library is missing or no sm_100 device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib

import numpy as np

import os

# RTCU_LIB overrides the library path (kernel-variant experiments only; the default is the in-tree build)
LIB_PATH = pathlib.Path(os.environ.get("RTCU_LIB") or (pathlib.Path(__file__).resolve().parent / "lib" / "librtcu.so"))

RTCU_OK, RTCU_ERR_INVALID, RTCU_ERR_CUDA, RTCU_ERR_STATE, RTCU_ERR_NOMEM = 0, -1, -2, -3, -4
MODE_MG, MODE_SM = 0, 1
ACCEL_AUTO, ACCEL_LINEAR, ACCEL_BVH = 0, 1, 2
PIPE_AUTO, PIPE_MEGAKERNEL, PIPE_WAVEFRONT = 0 << 4, 1 << 4, 2 << 4
FLAG_ACCUMULATE = 0x100
PRIM_MISS, PRIM_PLANE, PRIM_BOX = 0xFFFFFFFF, 0x80000000, 0x40000000

# every symbol include/rtcu.h declares (tests check the .so exports exactly these)
EXPORTS = (
    "rtcu_abi_version", "rtcu_device_count", "rtcu_create", "rtcu_destroy", "rtcu_last_error", "rtcu_bvh_threshold",
    "rtcu_upload_scene", "rtcu_render", "rtcu_render_device", "rtcu_resolve_device", "rtcu_sync", "rtcu_render_multi",
    "rtcu_intersect_batch", "rtcu_primary_rays", "rtcu_scatter_batch", "rtcu_philox_batch", "rtcu_get_stats", "rtcu_measure_fp32_peak", "rtcu_bvh_build_host",
    "rtcu_rasterize", "rtcu_rasterize_device", "rtcu_selftest_math",
    "rtcu_ipc_alloc", "rtcu_ipc_open", "rtcu_ipc_release", "rtcu_reduce_resolve_rows", "rtcu_bvh4_build_host",
    "rtcu_reload_env", "rtcu_upload_scene_multi", "rtcu_exchange_reduce_resolve", "rtcu_exchange_check",
    "rtcu_accum_download", "rtcu_accum_upload", "rtcu_set_output_pinning",
)


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("albedo", C.c_float * 4), ("roughness", C.c_float), ("reflectivity", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("spheres", C.c_void_p), ("sphere_material", C.c_void_p), ("n_spheres", C.c_uint32),
        ("planes", C.c_void_p), ("plane_material", C.c_void_p), ("n_planes", C.c_uint32),
        ("materials", C.c_void_p), ("n_materials", C.c_uint32),
        ("boxes", C.c_void_p), ("box_material", C.c_void_p), ("n_boxes", C.c_uint32),
    ]


class View(C.Structure):
    _fields_ = [
        ("inv_view_proj", C.c_float * 16),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("samples_per_pixel", C.c_uint32), ("max_bounces", C.c_uint32),
        ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
        ("tile_x0", C.c_uint32), ("tile_y0", C.c_uint32), ("tile_x1", C.c_uint32), ("tile_y1", C.c_uint32),
        ("seed", C.c_uint64),
        ("material_mode", C.c_uint32), ("flags", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("segments", C.c_uint64), ("samples", C.c_uint64), ("sphere_tests", C.c_uint64), ("node_visits", C.c_uint64),
        ("ms_render", C.c_float), ("ms_resolve", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
        ("kernel_launches", C.c_uint32), ("pipeline", C.c_uint32), ("accel", C.c_uint32), ("reserved", C.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class RtcuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rtcu error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Loads librtcu.so or raises -- the product path has no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m rt_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback for the path-tracing hot path.")
    lib = C.CDLL(str(LIB_PATH))
    p, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "rtcu_abi_version": (i, []),
        "rtcu_device_count": (i, []),
        "rtcu_create": (p, [i]),
        "rtcu_destroy": (None, [p]),
        "rtcu_last_error": (C.c_char_p, []),
        "rtcu_bvh_threshold": (u32, []),
        "rtcu_upload_scene": (i, [p, C.POINTER(SceneDesc)]),
        "rtcu_render": (i, [p, C.POINTER(View), p, p]),
        "rtcu_render_device": (i, [p, C.POINTER(View), p, i, p]),
        "rtcu_resolve_device": (i, [p, p, u32, u32, u32, p, p]),
        "rtcu_sync": (i, [p]),
        "rtcu_render_multi": (i, [C.POINTER(p), u32, C.POINTER(View), p, p]),
        "rtcu_intersect_batch": (i, [p, p, p, u32, p, p, p, p, u32]),
        "rtcu_primary_rays": (i, [p, C.POINTER(View), p, p, p, u32, p, p]),
        "rtcu_scatter_batch": (i, [p, u32, u64, u32, p, p, p, p, p, p, p, p, p, p, p, p]),
        "rtcu_philox_batch": (i, [p, p, u32, u64, p]),
        "rtcu_get_stats": (i, [p, C.POINTER(Stats)]),
        "rtcu_measure_fp32_peak": (i, [p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "rtcu_bvh_build_host": (i, [p, u32, p, p, u32, C.POINTER(u32), C.POINTER(u32)]),
        "rtcu_rasterize": (i, [p, C.POINTER(View), p, p, p]),
        "rtcu_rasterize_device": (i, [p, C.POINTER(View), p, p]),
        "rtcu_selftest_math": (i, [p, p, u32, p]),
        "rtcu_bvh4_build_host": (i, [p, u32, p, u32, p, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "rtcu_ipc_alloc": (i, [p, u64, C.POINTER(p), p]),
        "rtcu_ipc_open": (i, [p, p, C.POINTER(p)]),
        "rtcu_ipc_release": (i, [p, p]),
        "rtcu_reduce_resolve_rows": (i, [p, C.POINTER(p), u32, u32, u32, u32, u32, p, p]),
        "rtcu_reload_env": (i, [p]),
        "rtcu_upload_scene_multi": (i, [C.POINTER(p), u32, C.POINTER(SceneDesc)]),
        "rtcu_exchange_reduce_resolve": (i, [p, C.POINTER(p), C.POINTER(p), u32, u32, u32, u32, u32, u32, u32, u32, p, p]),
        "rtcu_exchange_check": (i, [p, p, p]),
        "rtcu_accum_download": (i, [p, u32, u32, p]),
        "rtcu_accum_upload": (i, [p, u32, u32, p]),
        "rtcu_set_output_pinning": (i, [p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().rtcu_last_error().decode(errors="replace")


def check(rc: int) -> None:
    if rc != RTCU_OK:
        raise RtcuError(rc, last_error())


def ptr(a: np.ndarray | None) -> int | None:
    return None if a is None else a.ctypes.data


def contiguous(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))
