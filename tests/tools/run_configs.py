#!/usr/bin/env python3
"""Times the BASELINE.json configs on cuda:0 through the C ABI (device-resident kernel time from CUDA events
inside rtcu_render, best of `--reps`) and prints one JSON line per config.  Optionally times the CPU oracle's
fast build beside it (`--cpu`, bounded row subsets).  Usage: python tests/tests/tools/run_configs.py [c1 c2 c2mg c3 c4 c5slice] """
import argparse, json, os, pathlib, sys, time

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from rt_b200 import _native as nat, scene as S, synth  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402


def configs():
    c1 = S.load(ROOT / "scenes/basic.toml")
    c2 = S.load(ROOT / "scenes/dielectric.toml")
    c3 = synth.rtiow_scene()
    return {
        "c1": (c1, dict(width=800, height=600, spp=30, depth=10, mode=nat.MODE_MG)),
        "c2": (c2, dict(width=1920, height=1080, spp=64, depth=50, mode=nat.MODE_SM)),
        "c2mg": (c2, dict(width=1920, height=1080, spp=64, depth=50, mode=nat.MODE_MG)),
        "c3": (c3, dict(width=1920, height=1080, spp=256, depth=50, mode=nat.MODE_SM)),
        "c3s": (c3, dict(width=1920, height=1080, spp=16, depth=50, mode=nat.MODE_SM)),  # short C3 for profiling
        "c4": (None, dict(width=3840, height=2160, spp=64, depth=10, mode=nat.MODE_SM)),
        "c4s": (None, dict(width=3840, height=2160, spp=4, depth=10, mode=nat.MODE_SM)),
        # profiling sizes that take the default BVH path (lanes share a pixel's samples: from 16 samples per call, 16 lanes from 32)
        "c3p": (c3, dict(width=1920, height=1080, spp=32, depth=50, mode=nat.MODE_SM)),
        "c4p": (None, dict(width=1920, height=1080, spp=32, depth=10, mode=nat.MODE_SM)),
        "c5slice": (c3, dict(width=3840, height=2160, spp=4096, depth=50, mode=nat.MODE_SM, sample_range=(0, 64))),
        "c5n8": (c3, dict(width=3840, height=2160, spp=4096, depth=50, mode=nat.MODE_SM, sample_range=(0, 512))),  # one GPU's share at N = 8
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=["c1", "c2", "c2mg", "c3"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--depth", type=int, default=0, help="override max_bounces (0 = the config's own)")
    ap.add_argument("--size", default="", help="override the frame size, WxH")
    ap.add_argument("--spp", type=int, default=0, help="override the samples per pixel")
    args = ap.parse_args()
    cfgs = configs()
    ctx = Context(0)
    ffma, ffma2 = ctx.measure_fp32_peak()
    for name in args.names:
        sc, c = cfgs[name]
        if sc is None:
            sc = synth.grid_scene()
        t0 = time.perf_counter()
        ctx.upload_scene(sc)
        upload_ms = (time.perf_counter() - t0) * 1e3
        if args.depth:
            c = dict(c, depth=args.depth)
        if args.size:
            c = dict(c, width=int(args.size.split("x")[0]), height=int(args.size.split("x")[1]))
        if args.spp:
            c = dict(c, spp=args.spp)
        v = make_view(sc, c["width"], c["height"], samples_per_pixel=c["spp"], max_bounces=c["depth"], material_mode=c["mode"],
                      sample_range=c.get("sample_range"), flags=args.flags)
        best, st = None, None
        for _ in range(args.reps):
            t0 = time.perf_counter()
            ctx.render(v, want_accum=False)
            wall = (time.perf_counter() - t0) * 1e3
            s = ctx.stats()
            if best is None or s["ms_render"] < best:
                best, st, best_wall = s["ms_render"], s, wall
        n = len(sc.spheres)
        flops = 18 * st["sphere_tests"] + 60 * (st["segments"] + st["samples"])
        out = {"config": name, "n_spheres": n, **{k: c[k] for k in ("width", "height", "spp", "depth", "mode")},
               "samples": st["samples"], "segments": st["segments"], "sphere_tests": st["sphere_tests"], "node_visits": st["node_visits"],
               "kernel_ms": round(best, 3), "e2e_ms": round(best_wall, 3), "upload_ms": round(upload_ms, 2),
               "msamples_per_s": round(st["samples"] / best / 1e3, 1), "gsegments_per_s": round(st["segments"] / best / 1e6, 2),
               "gtests_per_s": round(st["sphere_tests"] / best / 1e6, 1), "alg_tflops": round(flops / best / 1e9, 2),
               "frac_of_ffma_peak": round(flops / best / 1e9 / ffma, 4), "ffma_tflops": round(ffma, 1), "ffma2_tflops": round(ffma2, 1),
               "pipeline": st["pipeline"], "accel": st["accel"], "launches": st["kernel_launches"]}
        if args.cpu:
            from oracle.binding import Oracle
            o = Oracle("fast")
            cores = os.cpu_count()
            tile_rows = c["height"]
            step = max(1, tile_rows // (8 if n < 100 else 2))
            v1 = make_view(sc, c["width"], c["height"], samples_per_pixel=c["spp"], max_bounces=c["depth"], material_mode=c["mode"],
                           sample_range=(0, min(c["spp"], 8 if n >= 100 else c["spp"])))
            t0 = time.perf_counter()
            _, _, segs = o.render(sc, v1, threads=cores, row_step=step, want_accum=False)
            dt = time.perf_counter() - t0
            smp = len(range(0, c["height"], step)) * c["width"] * (v1.sample_end - v1.sample_begin)
            out["cpu_msamples_per_s"] = round(smp / dt / 1e6, 3)
            out["cpu_cores"] = cores
            out["cpu_sample"] = f"rows 0::{step}, samples [0,{v1.sample_end}) = {smp/1e6:.2f} Msamples in {dt:.2f}s"
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
