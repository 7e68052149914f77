#!/usr/bin/env python3
"""C5: the C3 scene at 3840x2160, 4096 spp, split by sample range over G GPUs of one box in a single process
(rtcu_render_multi: one context per device, NVLink peer-load reduce fused into the resolve kernel).  usage: run_c5.py G"""
import json, pathlib, sys, time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rt_b200 import _native as nat, synth  # noqa: E402
from rt_b200.renderer import Context, make_view, render_multi  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sc = synth.rtiow_scene()
ctxs = [Context(i) for i in range(g)]
for c in ctxs:
    c.upload_scene(sc)
view = make_view(sc, 3840, 2160, samples_per_pixel=spp, max_bounces=50, material_mode=nat.MODE_SM)
warm = make_view(sc, 3840, 2160, samples_per_pixel=g, max_bounces=50, material_mode=nat.MODE_SM)
render_multi(ctxs, warm)
t0 = time.perf_counter()
rgba8, _ = render_multi(ctxs, view)
wall = time.perf_counter() - t0
st = ctxs[0].stats()
samples = 3840 * 2160 * spp
print(json.dumps({"config": "c5", "gpus": g, "spp": spp, "samples": samples, "wall_s": round(wall, 4), "msamples_per_s_wall": round(samples / wall / 1e6, 1),
                  "ms_render_root": round(st["ms_render"], 2), "ms_reduce_resolve": round(st["ms_resolve"], 3), "segments": st["segments"]}))
