#!/usr/bin/env python3
"""Device A/B of the BVH kernels at progressive-refinement sample counts (4 .. 12 per call): the default against 2 / 4 lanes sharing
a pixel forced (RTCU_BVH_LANES).  gpurun -- 'python tools/sweep_bvh_low_spp.py'   (profiles/r2_bvh_low_spp.txt)"""
import os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "tools"))
from rt_b200 import _native as nat, synth
from rt_b200.renderer import Context, make_view
import run_configs
cfgs = run_configs.configs()
ctx = Context(0)
for name, sc in (("c3", cfgs["c3"][0]), ("c4", synth.grid_scene())):
    ctx.upload_scene(sc)
    depth = 50 if name == "c3" else 10
    for size in ("800x600", "1920x1080", "3840x2160"):
        w, h = (int(x) for x in size.split("x"))
        for spp in (4, 6, 8, 12):
            v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=depth, material_mode=nat.MODE_SM)
            cells = []
            for label, kv in (("default", {}), ("lanes2", {"RTCU_BVH_LANES": "2"}), ("lanes4", {"RTCU_BVH_LANES": "4"})):
                os.environ.pop("RTCU_BVH_LANES", None); os.environ.update(kv); ctx.reload_env(); ctx.upload_scene(sc)
                best = 1e30
                for _ in range(5):
                    ctx.render(v, want_rgba8=False, want_accum=False); best = min(best, ctx.stats()["ms_render"])
                cells.append(f"{label}={best:.3f}")
            print(name, size, f"spp={spp}", *cells, flush=True)
