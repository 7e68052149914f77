#!/usr/bin/env python3
"""Times what a scene change costs for the 100 001-sphere scene (C4): the host BVH build (binary SAH tree), build + collapse to the
4-wide device layout, and the whole rtcu_upload_scene (validation, build, pack, one H2D transfer), per builder thread count.
One JSON line per thread count; `first` is the first call of the process, `best` the fastest of five."""
import ctypes as C, json, os, pathlib, sys, time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rt_b200 import _native as nat, synth  # noqa: E402
from rt_b200.renderer import Context  # noqa: E402

lib = nat.load_library()
scene = synth.grid_scene()
sph = nat.contiguous(scene.spheres, np.float32)
n = len(sph)
gpu = "--no-gpu" not in sys.argv
ctx = Context(0) if gpu else None
prepared = ctx.prepare_scene(scene) if gpu else None


def binary():
    a, d = C.c_uint32(), C.c_uint32()
    nat.check(lib.rtcu_bvh_build_host(nat.ptr(sph), n, None, None, 0, C.byref(a), C.byref(d)))


def wide():
    a, b, d = C.c_uint32(), C.c_uint32(), C.c_uint32()
    nat.check(lib.rtcu_bvh4_build_host(nat.ptr(sph), n, None, 0, None, 0, C.byref(a), C.byref(b), C.byref(d)))


def upload():
    ctx.upload_prepared(prepared)


def ms(fn):
    t = time.perf_counter()
    fn()
    return (time.perf_counter() - t) * 1e3


print(json.dumps({"cores": os.cpu_count(), "spheres": n}), flush=True)
for threads in sys.argv[1:] if [a for a in sys.argv[1:] if a.isdigit()] else ("1", "2", "4", "8", "16"):
    if not threads.isdigit():
        continue
    os.environ["RTCU_BVH_THREADS"] = threads
    row = {"threads": int(threads)}
    for name, fn in (("build", binary), ("build_pack", wide)) + ((("upload_scene", upload),) if gpu else ()):
        t = [ms(fn) for _ in range(6)]
        row[name + "_ms"] = {"first": round(t[0], 2), "best": round(min(t[1:]), 2)}
    print(json.dumps(row), flush=True)
