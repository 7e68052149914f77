#!/usr/bin/env python3
"""Device A/B of the scan kernels over frame sizes and sample counts, in one process (knobs are re-read with rtcu_reload_env):
thread-per-pixel grid (first frame of a view, and best of the measured tile orders) against lanes sharing a pixel.
gpurun -- 'python tools/sweep_scan_sizes.py [--configs c1 c2] [--sizes 800x600 ...] [--spp 30 64] > gpurun_out/sweep.txt'"""
import argparse, os, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "tools"))
from rt_b200 import _native as nat  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402
import run_configs  # noqa: E402

VARIANTS = [("tpp-first", {"RTCU_SCAN_DIRECT": "0", "RTCU_TILE_ORDER": "0"}), ("tpp-tuned", {"RTCU_SCAN_DIRECT": "0"}),
            ("flat8", {"RTCU_SCAN_DIRECT": "8", "RTCU_SCAN_NESTED": "0"}), ("flat16", {"RTCU_SCAN_DIRECT": "16", "RTCU_SCAN_NESTED": "0"}),
            ("nest2", {"RTCU_SCAN_DIRECT": "2", "RTCU_SCAN_NESTED": "1"}), ("nest4", {"RTCU_SCAN_DIRECT": "4", "RTCU_SCAN_NESTED": "1"}),
            ("nest8", {"RTCU_SCAN_DIRECT": "8", "RTCU_SCAN_NESTED": "1"}), ("nest16", {"RTCU_SCAN_DIRECT": "16", "RTCU_SCAN_NESTED": "1"}),
            ("default", {})]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=["c1", "c2", "c2mg"])
    ap.add_argument("--sizes", nargs="*", default=["320x240", "800x600", "1280x720", "1920x1080", "3840x2160"])
    ap.add_argument("--spp", nargs="*", type=int, default=[16, 30, 64, 256])
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    cfgs = run_configs.configs()
    ctx = Context(0)
    keys = sorted({k for _, kv in VARIANTS for k in kv})
    for name in args.configs:
        sc, c = cfgs[name]
        ctx.upload_scene(sc)
        for size in args.sizes:
            w, h = (int(x) for x in size.split("x"))
            for spp in args.spp:
                v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=c["depth"], material_mode=c["mode"])
                cells = []
                for label, kv in VARIANTS:
                    for k in keys:
                        os.environ.pop(k, None)
                    os.environ.update(kv)
                    ctx.reload_env()
                    ctx.upload_scene(sc)  # forgets the view's tile-order history
                    best = 1e30
                    for _ in range(args.reps + (2 if label == "tpp-tuned" else 0)):  # (the tuned order exists from the third frame)
                        ctx.render(v, want_rgba8=False, want_accum=False)
                        best = min(best, ctx.stats()["ms_render"])
                    cells.append(f"{label}={best:.3f}")
                print(f"{name} {size} spp={spp}  " + "  ".join(cells), flush=True)


if __name__ == "__main__":
    main()
