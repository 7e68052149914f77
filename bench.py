#!/usr/bin/env python3
"""bench.py -- Msamples/s of the path-tracing hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1], "C2"): scenes/dielectric.toml at 1920x1080, 64 spp, max depth 50,
sm scatter table (lambert / metal / dielectric with Schlick).  A *step* is one frame.  At N > 1 the frame
is split by sample range: every rank renders 64 samples per pixel of a 64*N-spp frame (weak scaling, fixed
work per GPU), the fp32 accumulation buffers are reduce-scattered by row band over NCCL, every rank resolves its band and
rank 0 gathers the packed bands.

  value  whole-job Msamples/s with the scene resident on the device, timed with CUDA events per step on the
         launching stream; L2 is flushed between steps (outside the events); max over ranks.
  e2e    the same metric through the reference-facing C-ABI call with HOST buffers: every step re-uploads the
         scene (rtcu_upload_scene, H2D) and reads the packed image back into pinned host memory (D2H),
         wall clock around K steps.
  roofline      dominant kernel (k_render_mega) against the non-tensor FP32 peak: SURVEY.md 8d names the FP32
                pipe, not HBM or tensor cores, as the bound of this path.
  cpu_baseline  the reference's own sm_ray_tracer.cpp compiled against the muu stand-in (oracle/_ref, kind "reference";
                the oracle port when that library is absent) on the host cores, bounded row subset of the same frame.
                `--impl reference` prints that arm as its own line.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WIDTH, HEIGHT, SPP, MAX_BOUNCES = 1920, 1080, 64, 50
SCENE_FILE = "scenes/dielectric.toml"
WORKLOAD = "C2: scenes/dielectric.toml 1920x1080 64spp depth50 sm-table (lambert/metal/dielectric)"
METRIC, UNIT = "Msamples/sec (paths*spp/s)", "Msamples/s"
FLOP_PER_TEST, FLOP_PER_SEGMENT, FLOP_PER_SAMPLE = 18, 60, 60  # SURVEY.md 8d algorithmic work unit
# dram__bytes_read.sum + dram__bytes_write.sum of one k_render_mega launch on this workload, from the committed
# `ncu --set full` capture (profiles/r1_final_ncu_summary.txt, first two lines: 0.040 MB read, 1.0-1.7 MB written -- the mean
# of the two launches; the 33 MB accumulation buffer stays in the 126 MB L2 for the duration of the launch)
NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH = 40_700 + 1_350_000


def load_scene():
    from rt_b200 import scene as S

    return S.load(ROOT / SCENE_FILE)


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def cpu_baseline(scene, view, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times the reference's CPU implementation of the path on all host cores, on a bounded row subset of the frame.

    kind "reference": oracle/_ref/librt_ref_fast.so -- marzer/rt's own sm_ray_tracer.cpp compiled (in the dev container,
    with the reference's -O3 -ffast-math flags, x86-64-v3) against the muu stand-in; used whenever that library travelled.
    kind "port": the oracle's fast build (oracle/rtref.c), when it did not.
    Returns (mean Msamples/s, best Msamples/s, cores, kind, description, ms_per_step)."""
    from oracle.binding import Oracle, ReferenceBuild

    cores = os.cpu_count() or 1
    spp = view.sample_end - view.sample_begin
    if ReferenceBuild.FAST_PATH.exists():
        ref = ReferenceBuild("fast")
        kind = "reference"
        what = "marzer/rt sm_ray_tracer.cpp compiled against the muu stand-in (oracle/_ref, -O3 -march=x86-64-v3 -ffast-math)"

        def run(row_step):
            ref.render(scene, view.width, view.height, spp, view.max_bounces, view.seed, "sm_ray_tracer", threads=cores, row_step=row_step)
    else:
        oracle = Oracle("fast")
        kind = "port"
        what = "oracle fast build (oracle/rtref.c, -O3 -march=native -ffast-math)"

        def run(row_step):
            oracle.render(scene, view, threads=cores, row_step=row_step, want_rgba8=True, want_accum=False)

    t0 = time.perf_counter()
    run(120)  # probe: every 120th row
    probe = time.perf_counter() - t0
    per_row = probe / len(range(0, view.height, 120))
    rows_wanted = max(1, min(view.height, int(target_seconds / max(per_row, 1e-6))))
    row_step = max(1, view.height // rows_wanted)
    rows = len(range(0, view.height, row_step))
    samples = rows * view.width * spp
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run(row_step)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    desc = (f"{what}, {cores} threads, rows 0::{row_step} of the {view.width}x{view.height} frame "
            f"({rows} rows, {samples / 1e6:.1f} Msamples per step), same RNG streams")
    return samples / mean / 1e6, samples / best / 1e6, cores, kind, desc, mean * 1e3


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from rt_b200.renderer import make_view
    from rt_b200 import _native as nat

    scene = load_scene()
    view = make_view(scene, WIDTH, HEIGHT, samples_per_pixel=SPP, max_bounces=MAX_BOUNCES, material_mode=nat.MODE_SM)
    budget = 150.0 / max(1, args.steps + args.warmup)
    mean_v, best_v, cores, kind, desc, ms = cpu_baseline(scene, view, target_seconds=min(20.0, budget), steps=args.steps, warmup=args.warmup)
    # the same frame once more with the reference's OWN generator (src/random.cpp: thread_local mt19937) instead of the
    # counter-based stand-in: the renderer exactly as shipped -- shows that the RNG swap flatters neither side
    own_rng = None
    from oracle.binding import ReferenceBuild
    if kind == "reference" and ReferenceBuild.FAST_MT_PATH.exists():
        mt = ReferenceBuild("fast_mt")
        row_step = int(desc.split("rows 0::")[1].split(" ")[0])
        rows = len(range(0, HEIGHT, row_step))
        best = None
        for _ in range(2):  # first call warms the library up
            t0 = time.perf_counter()
            mt.render(scene, WIDTH, HEIGHT, SPP, MAX_BOUNCES, 0, "sm_ray_tracer", threads=cores, row_step=row_step)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        own_rng = {"value": round(rows * WIDTH * SPP / best / 1e6, 3), "unit": UNIT,
                   "note": "same sample with marzer/rt's own src/random.cpp (thread_local std::mt19937, random_device seed) linked in place of the counter-based stream"}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mean_v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": round(mean_v, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "own_rng": own_rng},
        "e2e": {"value": round(mean_v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as tdist

    from rt_b200 import _native as nat, build, dist as rdist
    from rt_b200.renderer import Context, make_view

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs one process per GPU: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=dev)

    if rank == 0:
        build.build_cuda()
    if world > 1:
        tdist.barrier()

    scene = load_scene()
    ctx = Context(local_rank)
    ctx.upload_scene(scene)
    total_spp = SPP * world
    view = make_view(scene, WIDTH, HEIGHT, samples_per_pixel=total_spp, max_bounces=MAX_BOUNCES, material_mode=nat.MODE_SM)
    mine = rdist.partition_view(view, rank, world, by="samples")
    gr = rdist.GpuRank(ctx, WIDTH, HEIGHT, dev, world=world)
    peer = world > 1 and args.exchange == "peer"
    if peer:
        # every rank must agree: if CUDA IPC / peer access is unavailable anywhere, all ranks use the NCCL exchange
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            gr.enable_peer_exchange(rank)
        except Exception as e:  # noqa: BLE001
            print(f"[rank {rank}] peer exchange unavailable ({e}); using --exchange nccl", file=sys.stderr, flush=True)
            ok.zero_()
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        peer = bool(ok.item())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def step_device():
        if peer:
            # trace into an IPC-shared buffer -> barrier -> ONE kernel per rank: peer-load sum of its row band + resolve +
            # store into rank 0's image over NVLink -> barrier
            gr.render_peer_reduce_resolve(mine, total_spp)
        elif world > 1:
            # trace -> reduce-scatter fp32 row bands (NCCL) -> resolve the band on every rank -> gather RGBA8 bands on rank 0
            gr.render_reduce_resolve(mine, rank, total_spp)
        else:
            gr.resolve(gr.render_accum(mine), total_spp)

    # ---- value: device-resident, CUDA events per step, L2 flushed between steps ----
    for _ in range(max(3, args.warmup)):
        step_device()
        torch.cuda.synchronize(dev)  # lets the library time the first frames of this view and settle its tile issue order
    sampler = ClockSampler(local_rank)
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize(dev)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        if world > 1:
            tdist.barrier()
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        if peer:
            ctx.render_device(mine, gr.peer_accum, accumulate=False, stream=stream.cuda_stream)
        else:
            accum = gr.render_accum(mine)
        kev[i][1].record(stream)
        if peer:
            tdist.all_reduce(gr._peer_sync)
            row0, row1 = rdist.row_band_for_rank(0, HEIGHT, rank, world)
            ctx.reduce_resolve_rows(gr.peer_accums, WIDTH, row0, row1 - row0, total_spp, gr.peer_img, stream=stream.cuda_stream)
            tdist.all_reduce(gr._peer_sync)
        elif world > 1:
            band = rdist.sum_row_bands(gr.accum_padded, rank, world)
            rdist.gather_bands(gr.resolve_band(band, total_spp), HEIGHT, rank, world)
        else:
            gr.resolve(accum, total_spp)
        ev[i][1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        tdist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(total_ms, op=tdist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    stats = ctx.stats()  # segments of the last render on this rank (deterministic per step)
    seg_t = torch.tensor([stats["segments"]], dtype=torch.int64, device=dev)
    if world > 1:
        tdist.all_reduce(seg_t, op=tdist.ReduceOp.SUM)
    segments_all = int(seg_t.item())
    samples_per_step = WIDTH * HEIGHT * total_spp
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the C-ABI, scene re-upload + image read-back every step ----
    host_rgba = torch.empty((HEIGHT, WIDTH), dtype=torch.int32).pin_memory()
    host_np = host_rgba.numpy().view(np.uint32)
    scene_bytes = int(scene.spheres.nbytes + scene.sphere_material.nbytes + scene.planes.nbytes + scene.plane_material.nbytes + scene.materials.nbytes)

    prepared = ctx.prepare_scene(scene)  # the descriptor over the host arrays; the upload itself happens every step

    def step_e2e():
        ctx.upload_prepared(prepared)
        if world == 1:
            ctx.render(view, rgba8=host_np, want_accum=False)  # rtcu_render: launch + D2H into the pinned buffer + sync
        else:
            img = gr.render_peer_reduce_resolve(mine, total_spp) if peer else gr.render_reduce_resolve(mine, rank, total_spp)
            if rank == 0:
                host_rgba.copy_(img, non_blocking=False)
            torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step_e2e()
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    if world > 1:
        tdist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(e2e_s, op=tdist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(e2e_s.item()) / 1e6

    if rank == 0:
        # ---- roofline of the dominant kernel (k_render_mega on this rank) ----
        ffma_tf, ffma2_tf = ctx.measure_fp32_peak()
        peaks = measured_peaks()
        sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_nominal = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        n_sph = len(scene.spheres)
        my_samples = WIDTH * HEIGHT * (mine.sample_end - mine.sample_begin)
        flops = FLOP_PER_TEST * n_sph * stats["segments"] + FLOP_PER_SEGMENT * stats["segments"] + FLOP_PER_SAMPLE * my_samples
        kern_avg_ms = sum(kern_ms) / len(kern_ms)
        achieved = flops / (kern_avg_ms * 1e-3) / 1e12
        alg_bytes = WIDTH * HEIGHT * 16
        roofline = {
            "bound": "fp32", "kernel": "k_render_mega", "achieved": round(achieved, 3), "peak": round(peak_nominal, 2), "unit": "TFLOP/s",
            "frac": round(achieved / peak_nominal, 4), "traffic": NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH if world == 1 else None,
            "peak_source": f"non-tensor FP32: {sm_count} SMs x 128 lanes x 2 x sm_max_mhz {sm_max_mhz:.0f} (MEASURED_PEAKS.json clock{'' if peaks else ', fallback'}); "
                           "SURVEY.md 8d: this path is FP32-pipe bound, not HBM/tensor",
            "peak_measured_ffma": round(ffma_tf, 2), "peak_measured_ffma2": round(ffma2_tf, 2),
            "frac_of_measured_ffma": round(achieved / ffma_tf, 4) if ffma_tf > 0 else None,
            "algorithmic_flops_per_launch": flops, "segments_per_launch": stats["segments"], "kernel_ms": round(kern_avg_ms, 4),
            "segments_per_s": round(stats["segments"] / (kern_avg_ms * 1e-3) / 1e9, 3), "segments_per_s_unit": "G/s",
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": round(alg_bytes / (kern_avg_ms * 1e-3) / 1e9, 2),
                    "peak_gbs": peaks.get("hbm_gbs", 6650.0), "note": "accumulation buffer write only; paths live in registers"},
        }
        # ---- CPU baseline on the host cores (bounded sample) ----
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cview = make_view(scene, WIDTH, HEIGHT, samples_per_pixel=SPP, max_bounces=MAX_BOUNCES, material_mode=nat.MODE_SM)
            mean_v, best_v, cores, kind, desc, _ = cpu_baseline(scene, cview, target_seconds=12.0, steps=1, warmup=0)
            cpu = {"value": round(mean_v, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "spp_per_gpu": SPP, "spp_total": total_spp, "max_bounces": MAX_BOUNCES,
                       "n_spheres": n_sph, "partition": "sample-range" if world > 1 else "single",
                       "exchange": ("fused peer-load reduce+resolve kernel over CUDA IPC / NVLink" if peer else "NCCL reduce-scatter + gather") if world > 1 else None,
                       "seed": view.seed,
                       "l2": "flushed between steps (256 MiB fill, outside the per-step CUDA events)",
                       "pipeline": "megakernel", "accel": "linear",
                       "tile_order": "chosen by the library per view: row-major vs descending cost of the previous frame's tiles, whichever it timed faster "
                                     "in the warm-up frames (scheduling only: every sample is traced every step, the image is bit-identical)"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": WIDTH * HEIGHT * 4},
            # per step and rank: the trace kernels the library counted for the last frame (k_render_mega, k_render_stragglers, and
            # k_tile_order when the cost-sorted tile order won) + the resolve / fused reduce-resolve kernel
            "gpu_launches": args.steps * (int(stats.get("kernel_launches") or 2) + 1) * world,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "wall_ms_per_step_incl_flush": round(wall * 1e3 / args.steps, 3),
            "segments_per_sample": round(segments_all / samples_per_step, 4),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--exchange", choices=("peer", "nccl"), default="peer",
                    help="N > 1: 'peer' = IPC-shared buffers + one fused NVLink reduce/resolve kernel per rank; 'nccl' = reduce-scatter + gather")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
