#!/usr/bin/env python3
"""bench.py -- Msamples/s of the path-tracing hot path on N B200s (one process per GPU).

Workload (the same at every N, so that the driver's scaling arithmetic compares like with like): BASELINE.json configs[4],
"C5" -- the synthetic random-spheres scene of configs[2] (484 spheres, mixed lambert / metal / dielectric, BVH traversal) at
3840x2160 and 4096 samples per pixel, depth 50: 33.97 G samples per frame.  It is the largest configuration of the list and it
fits one GPU (133 MB of fp32 sums), so N = 1 renders the whole frame; at N > 1 the frame is split by sample range -- rank g
renders global samples [g*4096/N, (g+1)*4096/N) of every pixel -- and the fp32 sums are added over NVLink: "scaling": "strong".
`--config c1..c4` selects the other BASELINE configs; at N = 1 the default run also times each of them for a few frames and
reports them under "configs" (C4, the 100 001-sphere scene, has no CPU arm that finishes: its brute-force reference needs
~10^15 sphere tests per frame, which is why it is not the headline).

A *step* is one frame.
  value   whole-job Msamples/s with the scene resident on the device: CUDA events per step on the launching stream, summed
          over exactly K steps, max over ranks; between steps L2 is flushed (256 MiB fill, outside the events).
  e2e     the same metric through the reference-facing C-ABI calls with HOST buffers: every step re-sends the scene
          (rtcu_upload_scene: validation, BVH build, H2D) and delivers the packed image into PAGEABLE 64-byte aligned host
          memory -- what the reference's image.cpp allocates; wall clock around K steps, max over ranks.
  roofline      dominant kernel against the non-tensor FP32 peak (the path is FP32-issue bound, not HBM or tensor);
                for BVH configs the algorithmic work is counted from the kernel's own node-visit and sphere-test counters.
  cpu_baseline  marzer/rt's own sm_ray_tracer.cpp / mg_ray_tracer.cpp compiled against the muu stand-in (oracle/_ref, kind
                "reference"; the oracle port when that library is absent) on all host cores, on a bounded sample of the same
                frame (a row subset at a reduced sample count: samples/s does not depend on the samples per pixel).
  parity_max_lsb  in-run check, untimed: N > 1 -- a row band of rank 0's reduced image against the same band rendered by one GPU
                over the whole sample range; N = 1 -- a row band through the BVH path against the same band through the
                reference's O(N) scan.  Largest difference of any 8-bit channel.
`--impl reference` prints the CPU arm as its own line.
"""
from __future__ import annotations

import argparse
import dataclasses
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

METRIC, UNIT = "Msamples/sec (paths*spp/s)", "Msamples/s"
# SURVEY.md 8d algorithmic work unit: 18 flop per ray-sphere test, 60 per path segment (shade) and per sample (generate); a visit of a
# 4-wide BVH node is four slab tests of 24 flop (3 sub, 3 mul, 6 fma, 4 min/max, 2 compares)
FLOP_PER_TEST, FLOP_PER_SEGMENT, FLOP_PER_SAMPLE, FLOP_PER_NODE_VISIT = 18, 60, 60, 96


@dataclasses.dataclass(frozen=True)
class Config:
    key: str
    label: str
    scene: str          # "file:<path>" | "rtiow" | "grid"
    width: int
    height: int
    spp: int
    depth: int
    mode: str           # scatter table: "sm" | "mg"
    cpu_renderer: str   # the reference renderer the CPU arm runs
    cpu_spp: int        # samples per pixel of the CPU sample


CONFIGS = {
    "c1": Config("c1", "C1: scenes/basic.toml 800x600 30spp depth10 mg-table (3 spheres, linear scan)", "file:scenes/basic.toml", 800, 600, 30, 10, "mg",
                 "mg_ray_tracer", 30),
    "c2": Config("c2", "C2: scenes/dielectric.toml 1920x1080 64spp depth50 sm-table (7 spheres, lambert/metal/dielectric, linear scan)",
                 "file:scenes/dielectric.toml", 1920, 1080, 64, 50, "sm", "sm_ray_tracer", 64),
    "c3": Config("c3", "C3: synthetic RTiOW random-spheres scene (484 spheres, seed 20260118) 1920x1080 256spp depth50 sm-table (BVH)", "rtiow",
                 1920, 1080, 256, 50, "sm", "sm_ray_tracer", 16),
    "c4": Config("c4", "C4: synthetic 100001-sphere grid scene (seed 20260119) 3840x2160 64spp depth10 sm-table (BVH)", "grid", 3840, 2160, 64, 10, "sm",
                 "sm_ray_tracer", 1),
    "c5": Config("c5", "C5: synthetic RTiOW random-spheres scene (484 spheres, seed 20260118) 3840x2160 4096spp depth50 sm-table (BVH), "
                       "split by sample range across the GPUs", "rtiow", 3840, 2160, 4096, 50, "sm", "sm_ray_tracer", 16),
}


def load_scene(cfg: Config):
    from rt_b200 import scene as S, synth

    if cfg.scene.startswith("file:"):
        return S.load(ROOT / cfg.scene[5:])
    return synth.rtiow_scene() if cfg.scene == "rtiow" else synth.grid_scene()


def material_mode(cfg: Config) -> int:
    from rt_b200 import _native as nat

    return nat.MODE_SM if cfg.mode == "sm" else nat.MODE_MG


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return {}


_RESULT_FD = None


def claim_stdout() -> None:
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner on fd 1 when the box
    sets NCCL_DEBUG=VERSION).  From here on everything written to fd 1 goes to stderr; emit() writes the line to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def kernel_facts(cfg_key: str) -> dict:
    """What only a profiler sees (DRAM bytes per launch, active lanes per warp instruction) for the dominant kernel of a config:
    read from profiles/kernel_facts.json, which names the ncu capture and the git commit it was taken at.  Absent -> nulls."""
    p = ROOT / "profiles" / "kernel_facts.json"
    try:
        return json.loads(p.read_text()).get(cfg_key, {})
    except Exception:
        return {}


def aligned_pageable(height: int, width: int, align: int = 64) -> np.ndarray:
    """(height, width) uint32 in pageable host memory whose first byte is `align`-aligned: what image.cpp:9-13 allocates"""
    raw = np.zeros(height * width * 4 + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + height * width * 4].view(np.uint32).reshape(height, width)


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
def cpu_baseline(cfg: Config, scene, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times the reference's CPU implementation of the path on all host cores, on a bounded sample of the config's frame: rows
    0::k of the full-width frame at cfg.cpu_spp samples per pixel (the cost of a sample does not depend on how many follow it).

    kind "reference": oracle/_ref/librt_ref_fast.so -- marzer/rt's own renderer sources compiled (in the dev container, with the
    reference's -O3 -ffast-math flags, x86-64-v3) against the muu stand-in; used whenever that library travelled.
    kind "port": the oracle's fast build (oracle/rtref.c), when it did not.
    Returns (mean Msamples/s, best Msamples/s, cores, kind, description, ms per step, samples per step)."""
    from oracle.binding import Oracle, ReferenceBuild
    from rt_b200.renderer import make_view

    cores = os.cpu_count() or 1
    spp = min(cfg.spp, cfg.cpu_spp)
    if ReferenceBuild.FAST_PATH.exists():
        ref = ReferenceBuild("fast")
        kind = "reference"
        what = f"marzer/rt {cfg.cpu_renderer}.cpp compiled against the muu stand-in (oracle/_ref, -O3 -march=x86-64-v3 -ffast-math)"

        def run(row_step):
            ref.render(scene, cfg.width, cfg.height, spp, cfg.depth, 0x5EED, cfg.cpu_renderer, threads=cores, row_step=row_step)
    else:
        oracle = Oracle("fast")
        kind = "port"
        what = "oracle fast build (oracle/rtref.c, -O3 -march=native -ffast-math)"
        view = make_view(scene, cfg.width, cfg.height, samples_per_pixel=spp, max_bounces=cfg.depth, material_mode=material_mode(cfg))

        def run(row_step):
            oracle.render(scene, view, threads=cores, row_step=row_step, want_rgba8=True, want_accum=False)

    # probe with a sparse row set, then choose the row stride that fills the time budget
    probe_step = max(1, cfg.height // max(cores // 2, 4))
    t0 = time.perf_counter()
    run(probe_step)
    probe = time.perf_counter() - t0
    per_row = probe / len(range(0, cfg.height, probe_step))
    rows_wanted = max(1, min(cfg.height, int(target_seconds / max(per_row, 1e-6))))
    row_step = max(1, cfg.height // rows_wanted)
    # the sparse probe under-estimates frames whose rows differ in cost: one untimed pass over the chosen rows (it doubles as the
    # first warm-up step) and a coarser stride when that pass overshoots the budget
    t0 = time.perf_counter()
    run(row_step)
    dt = time.perf_counter() - t0
    if dt > 1.25 * target_seconds and row_step < cfg.height:
        row_step = min(cfg.height, int(row_step * dt / target_seconds) + 1)
    warmup = max(0, warmup - 1)
    rows = len(range(0, cfg.height, row_step))
    samples = rows * cfg.width * spp
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run(row_step)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    desc = (f"{what}, {cores} threads, rows 0::{row_step} of the {cfg.width}x{cfg.height} frame at {spp} spp "
            f"({rows} rows, {samples / 1e6:.1f} Msamples per step), same RNG streams")
    return samples / mean / 1e6, samples / best / 1e6, cores, kind, desc, mean * 1e3, samples


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    scene = load_scene(cfg)
    budget = 150.0 / max(1, args.steps + args.warmup)
    mean_v, best_v, cores, kind, desc, ms, samples = cpu_baseline(cfg, scene, target_seconds=min(20.0, budget), steps=args.steps, warmup=args.warmup)
    # the same sample once more with the reference's OWN generator (src/random.cpp: thread_local mt19937) instead of the
    # counter-based stand-in: the renderer exactly as shipped -- shows that the RNG swap flatters neither side
    own_rng = None
    from oracle.binding import ReferenceBuild
    if kind == "reference" and ReferenceBuild.FAST_MT_PATH.exists():
        mt = ReferenceBuild("fast_mt")
        row_step = int(desc.split("rows 0::")[1].split(" ")[0])
        spp = min(cfg.spp, cfg.cpu_spp)
        best = None
        for _ in range(2):  # first call warms the library up
            t0 = time.perf_counter()
            mt.render(scene, cfg.width, cfg.height, spp, cfg.depth, 0, cfg.cpu_renderer, threads=cores, row_step=row_step)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        own_rng = {"value": round(samples / best / 1e6, 3), "unit": UNIT,
                   "note": "same sample with marzer/rt's own src/random.cpp (thread_local std::mt19937, random_device seed) linked in place of the counter-based stream"}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mean_v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg.label, "sample": desc},
        "cpu_baseline": {"value": round(mean_v, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "own_rng": own_rng},
        "e2e": {"value": round(mean_v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
def algorithmic_flops(stats: dict, n_spheres: int, samples: int) -> tuple[int, dict]:
    """SURVEY.md 8d work unit from the kernel's own exact counters"""
    from rt_b200 import _native as nat

    segs = stats["segments"]
    if stats["accel"] == nat.ACCEL_BVH:
        flops = FLOP_PER_TEST * stats["sphere_tests"] + FLOP_PER_NODE_VISIT * stats["node_visits"] + FLOP_PER_SEGMENT * segs + FLOP_PER_SAMPLE * samples
        extra = {"node_visits_per_segment": round(stats["node_visits"] / max(segs, 1), 3),
                 "sphere_test_slots_per_segment": round(stats["sphere_tests"] / max(segs, 1), 3),
                 "equivalent_brute_force_flops": FLOP_PER_TEST * n_spheres * segs}
    else:
        flops = FLOP_PER_TEST * n_spheres * segs + FLOP_PER_SEGMENT * segs + FLOP_PER_SAMPLE * samples
        extra = {}
    return flops, extra


def time_config_once(torch, ctx, gr_cls, cfg: Config, dev, flush, target_s: float = 0.4, min_steps: int = 5) -> dict:
    """A few device-resident frames of one of the other BASELINE configs at N = 1: CUDA events per frame on the launching stream,
    L2 flushed between frames.  Returns the per-config record of the "configs" block."""
    from rt_b200 import _native as nat
    from rt_b200.renderer import make_view

    scene = load_scene(cfg)
    t0 = time.perf_counter()
    ctx.upload_scene(scene)
    upload_ms = (time.perf_counter() - t0) * 1e3
    gr = gr_cls(ctx, cfg.width, cfg.height, dev, world=1)
    view = make_view(scene, cfg.width, cfg.height, samples_per_pixel=cfg.spp, max_bounces=cfg.depth, material_mode=material_mode(cfg))
    stream = torch.cuda.current_stream(dev)

    def frame(seed):
        view.seed = seed
        gr.resolve(gr.render_accum(view), cfg.spp)

    for i in range(3):
        frame(0x5EED + i)
        torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); frame(0x5EED + 3); b.record(stream)
    torch.cuda.synchronize(dev)
    steps = max(min_steps, min(200, int(target_s * 1e3 / max(a.elapsed_time(b), 1e-3))))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record(stream)
        frame(0x5EED + 4 + i)
        ev[i][1].record(stream)
    torch.cuda.synchronize(dev)
    ms = [x.elapsed_time(y) for x, y in ev]
    stats = ctx.stats()
    samples = cfg.width * cfg.height * cfg.spp
    flops, extra = algorithmic_flops(stats, len(scene.spheres), samples)
    mean = sum(ms) / len(ms)
    del gr
    rec = {"workload": cfg.label, "steps": steps, "ms_per_step": round(mean, 4), "ms_min": round(min(ms), 4), "value": round(samples / mean / 1e3, 1), "unit": UNIT,
           "n_spheres": len(scene.spheres), "accel": "bvh" if stats["accel"] == nat.ACCEL_BVH else "linear",
           "segments_per_sample": round(stats["segments"] / samples, 4), "algorithmic_tflops": round(flops / mean / 1e9, 3),
           "scene_upload_ms": round(upload_ms, 2), **{k: v for k, v in extra.items() if k != "equivalent_brute_force_flops"}}
    if "equivalent_brute_force_flops" in extra:
        rec["equivalent_brute_force_tflops"] = round(extra["equivalent_brute_force_flops"] / mean / 1e9, 1)
    facts = kernel_facts(cfg.key)
    if facts:
        rec["ncu"] = facts
    return rec


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

    cfg = CONFIGS[args.config]
    W, H = cfg.width, cfg.height
    scene = load_scene(cfg)
    ctx = Context(local_rank)
    ctx.upload_scene(scene)
    weak = args.scaling == "weak"
    total_spp = cfg.spp * world if weak else cfg.spp
    if total_spp < world:
        raise SystemExit(f"{cfg.key}: {total_spp} samples per pixel cannot be split over {world} GPUs")
    mode = material_mode(cfg)
    view = make_view(scene, W, H, samples_per_pixel=total_spp, max_bounces=cfg.depth, material_mode=mode)
    mine = rdist.partition_view(view, rank, world, by="samples")
    gr = rdist.GpuRank(ctx, W, H, dev, world=world)
    exchange = args.exchange if world > 1 else None
    if world > 1 and exchange != "nccl":
        # every rank must agree: if CUDA IPC / peer access is unavailable anywhere, all ranks use the NCCL exchange
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            gr.enable_peer_exchange(rank)
        except Exception as e:  # noqa: BLE001
            print(f"[rank {rank}] peer exchange unavailable ({e}); using --exchange nccl", file=sys.stderr, flush=True)
            ok.zero_()
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        if not bool(ok.item()):
            exchange = "nccl"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    base_seed = int(view.seed)

    def set_seed(i):
        # a new seed every step: no step re-traces the previous step's paths (the tile-order heuristic of the scan kernels then
        # sees the previous frame of the same VIEW, as in progressive refinement, never the identical frame)
        view.seed = mine.seed = base_seed + i

    def step_device(i, kev=None):
        set_seed(i)
        if kev:
            kev[0].record(stream)
        if exchange == "peer":
            # trace into this frame's IPC-shared buffer, then ONE kernel: flag handshake + peer-load sum of this rank's row band +
            # resolve + store into rank 0's image over NVLink
            epoch, buf = gr.next_frame()
            ctx.render_device(mine, gr.peer_accums[buf][rank], accumulate=False, stream=stream.cuda_stream)
            if kev:
                kev[1].record(stream)
            gr.exchange(epoch, buf, total_spp, stream.cuda_stream)
        elif exchange == "peer-barrier":
            gr.render_peer_barrier_reduce_resolve(mine, total_spp)
            if kev:
                kev[1].record(stream)
        elif exchange == "nccl":
            # trace -> reduce-scatter fp32 row bands (NCCL) -> resolve the band on every rank -> gather RGBA8 bands on rank 0
            gr.render_accum(mine)
            if kev:
                kev[1].record(stream)
            band = rdist.sum_row_bands(gr.accum_padded, rank, world)
            rdist.gather_bands(gr.resolve_band(band, total_spp), H, rank, world)
        else:
            accum = gr.render_accum(mine)
            if kev:
                kev[1].record(stream)
            gr.resolve(accum, total_spp)

    # ---- value: device-resident, CUDA events per step, L2 flushed between steps ----
    warmup = max(3, args.warmup)
    for i in range(warmup):
        step_device(i)
        torch.cuda.synchronize(dev)  # (lets the library time the first frames of a view and settle its tile issue order)
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
        ev[i][0].record(stream)
        step_device(warmup + i, kev[i])
        ev[i][1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        tdist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    if exchange == "peer":
        gr.check_exchange()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(total_ms, op=tdist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    stats = ctx.stats()  # counters of the last frame on this rank
    seg_t = torch.tensor([stats["segments"]], dtype=torch.int64, device=dev)
    if world > 1:
        tdist.all_reduce(seg_t, op=tdist.ReduceOp.SUM)
    segments_all = int(seg_t.item())
    samples_per_step = W * H * total_spp
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- parity, untimed ----
    parity = None
    band_rows = 16 if cfg.height >= 64 else cfg.height
    y0 = min(cfg.height - band_rows, (cfg.height * 5 // 8) & ~3)
    if rank == 0:
        band = make_view(scene, W, H, samples_per_pixel=total_spp, max_bounces=cfg.depth, material_mode=mode, tile=(0, y0, W, y0 + band_rows), seed=int(view.seed))
        if world > 1:
            reduced = (gr.peer_rgba8 if exchange in ("peer", "peer-barrier") else None)
            if reduced is not None:
                got = reduced[y0:y0 + band_rows].clone()
                gr.resolve(gr.render_accum(band), total_spp)
                want = gr.rgba8[y0:y0 + band_rows]
                what = f"rows [{y0},{y0 + band_rows}) of the {world}-GPU image vs the same rows over all {total_spp} samples on one GPU"
        else:
            gr.resolve(gr.render_accum(mine), total_spp)
            got = gr.rgba8[y0:y0 + band_rows].clone()
            band.flags = nat.ACCEL_LINEAR
            gr.resolve(gr.render_accum(band), total_spp)
            want = gr.rgba8[y0:y0 + band_rows]
            what = f"rows [{y0},{y0 + band_rows}): the default path ({'BVH' if stats['accel'] == nat.ACCEL_BVH else 'scan'}) vs the reference's O(N) scan"
        if world == 1 or reduced is not None:
            shifts = torch.tensor([24, 16, 8], device=dev, dtype=torch.int32)
            d = (((got.unsqueeze(-1) >> shifts) & 255) - ((want.unsqueeze(-1) >> shifts) & 255)).abs()
            parity = {"max_lsb": int(d.max().item()), "pixels_differing": int((d.amax(-1) > 0).sum().item()), "pixels": int(got.numel()), "what": what}
    if world > 1:
        tdist.barrier()

    # ---- e2e: host buffers through the C-ABI, scene re-upload + image delivery into pageable memory every step ----
    host_img = aligned_pageable(H, W)
    host_t = torch.from_numpy(host_img.view(np.int32))
    scene_bytes = int(scene.spheres.nbytes + scene.sphere_material.nbytes + scene.planes.nbytes + scene.plane_material.nbytes + scene.materials.nbytes)
    prepared = ctx.prepare_scene(scene)  # the descriptor over the host arrays; the upload itself happens every step
    ctx.set_output_pinning(True)  # as plugin/cuda_path_tracer.cpp does: the frame's image is page-locked once, on first sight

    def step_e2e(i, dst=host_img):
        set_seed(i)
        ctx.upload_prepared(prepared)
        if world == 1:
            ctx.render(view, rgba8=dst, want_accum=False)  # rtcu_render: launch + delivery into the caller's host image + sync
        else:
            if exchange == "peer":
                img = gr.render_peer_reduce_resolve(mine, total_spp)
            elif exchange == "peer-barrier":
                img = gr.render_peer_barrier_reduce_resolve(mine, total_spp)
            else:
                img = gr.render_reduce_resolve(mine, rank, total_spp)
            if rank == 0:
                host_t.copy_(img, non_blocking=False)
            torch.cuda.synchronize(dev)

    e2e_steps = args.steps
    for i in range(3):
        step_e2e(i)
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(3 + i)
    torch.cuda.synchronize(dev)
    if world > 1:
        tdist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(e2e_s, op=tdist.ReduceOp.MAX)
    e2e_value = samples_per_step * e2e_steps / float(e2e_s.item()) / 1e6
    e2e_stats = ctx.stats() if world == 1 else None
    e2e_pinned = None
    if world == 1:
        # secondary figure: the same loop into a caller-pinned image
        pinned = torch.empty((H, W), dtype=torch.int32).pin_memory()
        pin_np = pinned.numpy().view(np.uint32)
        n_pin = max(3, min(e2e_steps, 5))
        step_e2e(0, pin_np)
        t0 = time.perf_counter()
        for i in range(n_pin):
            step_e2e(1 + i, pin_np)
        e2e_pinned = round(samples_per_step * n_pin / (time.perf_counter() - t0) / 1e6, 2)

    if rank == 0:
        # ---- roofline of the dominant kernel on this rank ----
        ffma_tf, ffma2_tf = ctx.measure_fp32_peak()
        peaks = measured_peaks()
        sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_nominal = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        n_sph = len(scene.spheres)
        my_samples = W * H * (mine.sample_end - mine.sample_begin)
        flops, extra = algorithmic_flops(stats, n_sph, my_samples)
        kern_avg_ms = sum(kern_ms) / len(kern_ms)
        achieved = flops / (kern_avg_ms * 1e-3) / 1e12
        is_bvh = stats["accel"] == nat.ACCEL_BVH
        facts = kernel_facts(cfg.key)
        roofline = {
            "bound": "fp32", "kernel": ("k_render_bvh (lanes share a pixel's samples)" if is_bvh else
                                       "k_render_scan_shared (k_render_stragglers<scan>: lanes share a pixel's samples)" if int(stats.get("kernel_launches") or 0) == 1 else "k_render_mega"),
            "achieved": round(achieved, 3), "peak": round(peak_nominal, 2), "unit": "TFLOP/s", "frac": round(achieved / peak_nominal, 4),
            "traffic": facts.get("dram_bytes_per_launch") if world == 1 else None,
            "traffic_source": facts.get("source") if world == 1 else None,
            "peak_source": f"non-tensor FP32: {sm_count} SMs x 128 lanes x 2 x sm_max_mhz {sm_max_mhz:.0f} (MEASURED_PEAKS.json clock{'' if peaks else ', fallback'}); "
                           "SURVEY.md 8d: this path is FP32-issue bound, not HBM/tensor",
            "peak_measured_ffma": round(ffma_tf, 2), "peak_measured_ffma2": round(ffma2_tf, 2),
            "frac_of_measured_ffma": round(achieved / ffma_tf, 4) if ffma_tf > 0 else None,
            "algorithmic_flops_per_launch": flops, "segments_per_launch": stats["segments"], "kernel_ms": round(kern_avg_ms, 4),
            "segments_per_s": round(stats["segments"] / (kern_avg_ms * 1e-3) / 1e9, 3), "segments_per_s_unit": "G/s",
            "work_unit": ("18 flop x sphere-test slots + 96 x node visits (both counted by the kernel) + 60 x (segments + samples)" if is_bvh
                          else "18 flop x spheres x segments + 60 x (segments + samples)"),
            "active_lanes_per_warp_instruction": facts.get("active_lanes"),
            "hbm": {"algorithmic_bytes": W * H * 16, "achieved_gbs": round(W * H * 16 / (kern_avg_ms * 1e-3) / 1e9, 2),
                    "peak_gbs": peaks.get("hbm_gbs", 6650.0), "note": "accumulation buffer write only; paths live in registers"},
        }
        for k, v in extra.items():
            if k == "equivalent_brute_force_flops":
                roofline["equivalent_brute_force_tflops"] = round(v / (kern_avg_ms * 1e-3) / 1e12, 1)
            else:
                roofline[k] = v
        # ---- CPU baseline on the host cores (bounded sample), N = 1 only ----
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            mean_v, best_v, cores, kind, desc, _, _ = cpu_baseline(cfg, scene, target_seconds=12.0, steps=1, warmup=0)
            cpu = {"value": round(mean_v, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        # ---- the other BASELINE configs, a few frames each (N = 1 default run only) ----
        others = None
        if world == 1 and not args.no_configs:
            others = {}
            for key in ("c1", "c2", "c3", "c4"):  # (C5 is 2-3 s per frame: only as the headline)
                if key != cfg.key:
                    others[key] = time_config_once(torch, ctx, rdist.GpuRank, CONFIGS[key], dev, flush)
        launches_per_step = int(stats.get("kernel_launches") or 1) + 1  # trace kernels the library counted + resolve / exchange
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg.label, "width": W, "height": H, "spp_total": total_spp, "spp_per_gpu": mine.sample_end - mine.sample_begin,
                       "max_bounces": cfg.depth, "n_spheres": n_sph, "partition": "sample-range" if world > 1 else "single",
                       "exchange": {"peer": "one kernel per rank: release/acquire flag handshake in IPC-shared device memory + NVLink peer-load sum of the rank's row band "
                                            "+ resolve + store into rank 0's image; no collective, no barrier on the frame path",
                                    "peer-barrier": "two NCCL one-element all-reduce barriers around the peer-load reduce+resolve kernel (round-1 form)",
                                    "nccl": "NCCL reduce-scatter + gather", None: None}[exchange],
                       "seed": f"{base_seed} + step index", "l2": "flushed between steps (256 MiB fill, outside the per-step CUDA events)",
                       "pipeline": "megakernel", "accel": "bvh" if is_bvh else "linear", "timed_region_s": round(total_ms * 1e-3, 3)},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": scene_bytes * world, "d2h_bytes_per_step": W * H * 4,
                    "destination": "pageable 64-byte aligned host memory (as image.cpp:9-13), page-locked on first sight (rtcu_set_output_pinning, as the plugin does)" if world == 1
                                   else "pageable host memory on rank 0 (torch copy from the reduced device image)",
                    "steps": e2e_steps, "value_pinned_destination": e2e_pinned,
                    "ms_d2h_last_frame": round(e2e_stats["ms_d2h"], 4) if e2e_stats else None},
            "gpu_launches": args.steps * launches_per_step * world,
            "parity_max_lsb": parity["max_lsb"] if parity else None, "parity": parity,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "wall_ms_per_step_incl_flush": round(wall * 1e3 / args.steps, 3),
            "segments_per_sample": round(segments_all / samples_per_step, 4),
            "configs": others,
        }
        emit(line)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c5", help="BASELINE.json config (default c5: 3840x2160, 4096 spp, the same frame at every N)")
    ap.add_argument("--scaling", choices=("strong", "weak"), default="strong",
                    help="N > 1: 'strong' splits the config's samples per pixel over the GPUs (fixed frame); 'weak' gives every GPU the config's samples")
    ap.add_argument("--exchange", choices=("peer", "peer-barrier", "nccl"), default="peer",
                    help="N > 1: 'peer' = IPC-shared buffers + ONE kernel per rank (flag handshake, NVLink peer-load reduce, resolve); "
                         "'peer-barrier' = the same reduce kernel between two NCCL barriers; 'nccl' = reduce-scatter + gather")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table of the N = 1 run")
    args = ap.parse_args()
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    claim_stdout()
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
