#!/usr/bin/env python3
"""Measures the linear-scan / BVH crossover: the first N spheres of the C3 scene at 1920x1080, 16 spp, both traversals."""
import json, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rt_b200 import _native as nat, synth  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402

full = synth.rtiow_scene()
ctx = Context(0)
for n in (8, 16, 24, 32, 48, 64, 96, 128, 256, 484):
    sc = synth.rtiow_scene()
    keep = list(range(n - 3)) + list(range(len(full.spheres) - 3, len(full.spheres)))  # ground + first small ones + the three big ones
    keep = keep[:n] if n < len(full.spheres) else list(range(len(full.spheres)))
    sc.spheres = full.spheres[keep]
    sc.sphere_material = full.sphere_material[keep]
    ctx.upload_scene(sc)
    out = {"n": len(sc.spheres)}
    for name, accel in (("linear", nat.ACCEL_LINEAR), ("bvh", nat.ACCEL_BVH)):
        best = None
        for _ in range(3):
            ctx.render(make_view(sc, 1920, 1080, samples_per_pixel=16, max_bounces=50, flags=accel), want_rgba8=False)
            ms = ctx.stats()["ms_render"]
            best = ms if best is None or ms < best else best
        out[name + "_ms"] = round(best, 3)
    print(json.dumps(out), flush=True)
ctx.close()
