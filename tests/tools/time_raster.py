#!/usr/bin/env python3
"""Times rtcu_rasterize_device (cuda_rasterizer, reference rasterizer.cpp:22-88) on full-HD frames; CUDA events on torch's
current stream, L2 flushed between launches.  Prints one JSON line per scene; the CPU oracle is timed beside it with
--cpu (all host threads, a bounded sample of rows)."""
import argparse, json, pathlib, sys, time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from rt_b200 import scene as S, synth  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    scenes = {"boxes": S.load(ROOT / "scenes" / "boxes.toml"), "c2": S.load(ROOT / "scenes" / "dielectric.toml"), "c3": synth.rtiow_scene(),
              "c4s": synth.grid_scene()}
    ctx = Context(0)
    w, h = 1920, 1080
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    for name, sc in scenes.items():
        ctx.upload_scene(sc)
        v = make_view(sc, w, h)
        for _ in range(3):
            ctx.rasterize_device(v, out.data_ptr(), stream=st)
        ms = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.rasterize_device(v, out.data_ptr(), stream=st)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = float(np.median(ms))
        n_prim = len(sc.spheres) + len(sc.planes) + len(sc.boxes)
        st_ = ctx.stats()
        rec = {"scene": name, "accel": "bvh" if st_["accel"] == 2 else "linear", "n_spheres": len(sc.spheres), "n_planes": len(sc.planes), "n_boxes": len(sc.boxes), "width": w, "height": h,
               "ms": round(ms, 4), "mpixels_per_s": round(w * h / ms / 1e3, 1), "gtests_per_s": round(w * h * n_prim / ms / 1e6, 1)}
        if a.cpu:
            sys.path.insert(0, str(ROOT))
            from oracle.binding import ReferenceBuild

            ref = ReferenceBuild("fast")  # the reference's own rasterizer.cpp, its -O3 -ffast-math flags, all host threads
            step = max(1, int(n_prim / 60))
            t0 = time.perf_counter()
            ref.render(sc, w, h, 1, 1, 0, "rasterizer", threads=0, row_step=step)
            dt = time.perf_counter() - t0
            rows = len(range(0, h, step))
            rec["cpu_mpixels_per_s"] = round(w * rows / dt / 1e6, 2)
            rec["cpu_sample"] = f"reference rasterizer.cpp, every {step}th row, all host threads"
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
