#!/usr/bin/env python3
"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck): linear + BVH megakernel, straggler
pass, wavefront pipeline, batch entry points.  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from rt_b200 import _native as nat, scene as S, synth  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402

os.environ["RTCU_WF_RAYS"] = str(96 * 54 * 2)
ctx = Context(0)
for name, sc in (("c2", S.load(ROOT / "scenes/dielectric.toml")), ("c3", synth.rtiow_scene()), ("grid", synth.grid_scene(nx=40, nz=30))):
    ctx.upload_scene(sc)
    for flags in (nat.ACCEL_LINEAR, nat.ACCEL_BVH, nat.ACCEL_LINEAR | nat.PIPE_WAVEFRONT, nat.ACCEL_BVH | nat.PIPE_WAVEFRONT):
        v = make_view(sc, 97, 53, samples_per_pixel=5, max_bounces=12, flags=flags)  # odd size: partial tiles
        rgba8, accum = ctx.render(v, want_accum=True)
        assert (accum[..., 3] == 5).all()
    o, d = synth.random_rays(sc, 5000, seed=1)
    for accel in (nat.ACCEL_LINEAR, nat.ACCEL_BVH):
        ctx.intersect_batch(o, d, accel=accel)
    print(name, "ok", ctx.stats()["segments"])
v = make_view(sc, 64, 48)
ctx.primary_rays(v, np.arange(100) % 64, np.arange(100) % 48, np.arange(100) % 7)
ctx.philox_batch(np.arange(400, dtype=np.uint32).reshape(100, 4), 99)
ctx.close()
print("sanitize smoke done")
