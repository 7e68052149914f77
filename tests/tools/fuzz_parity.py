#!/usr/bin/env python3
"""Randomised parity campaign on a GPU box: N random scenes (1-300 spheres, 0-3 planes, all eight material types, random
cameras, both scatter tables) rendered by the CUDA path (linear scan, BVH, wavefront pipeline, ray-pool kernel) and by the
strict CPU oracle; closest hits of random + silhouette-grazing ray batches compared bit for bit; the preview renderer
(rtcu_rasterize, scan and BVH, with random boxes) compared bit for bit (pixels, primitive ids, depth); every third scene also rendered through the lanes-share-a-pixel BVH path.
Prints one JSON summary.
usage: python tests/tests/tools/fuzz_parity.py [n_scenes] [seed]"""
import json, os, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from oracle.binding import Oracle  # noqa: E402
from rt_b200 import _native as nat, scene as S, synth  # noqa: E402
from rt_b200.renderer import Context, make_view  # noqa: E402

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
oracle, ctx = Oracle("strict"), Context(0)
stats = dict(scenes=0, renders=0, rays=0, seg_mismatch=0, accum_max_rel=0.0, rgba_max_lsb=0, hit_mismatch=0, bvh_vs_linear_mismatch=0,
             raster_frames=0, raster_mismatch=0, failures=[])


def random_scene(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 3, 5, 8, 17, 40, 90, 300]))
    k = int(rng.integers(1, 9))
    mats = [(int(rng.integers(0, 8)), tuple(rng.uniform(0.05, 1.0, 3)), float(rng.uniform(0, 0.7)), float(rng.choice([0.3, 0.5, 0.8, 1.0, 1.31, 1.333, 1.52, 2.4]))) for _ in range(k)]
    sc = S.Scene(samples_per_pixel=int(rng.integers(1, 6)), max_bounces=int(rng.choice([1, 2, 5, 12, 50])))
    sc.materials = S.make_materials(mats)
    spread = float(rng.choice([2.0, 6.0, 20.0]))
    sph = np.concatenate([rng.normal(0, spread, (n, 3)), rng.uniform(0.05, 1.5, (n, 1)) * rng.choice([0.2, 1.0, 3.0])], axis=1)
    if rng.random() < 0.7:
        sph[0] = [0, -1000.5, 0, 1000]
    sc.spheres = sph.astype(np.float32)
    sc.sphere_material = rng.integers(0, k, n).astype(np.uint32)
    npl = int(rng.integers(0, 4)) if rng.random() < 0.4 else 0
    if npl:
        nrm = rng.normal(size=(npl, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        sc.planes = np.concatenate([nrm, rng.uniform(0.5, 8, (npl, 1))], axis=1).astype(np.float32)
        sc.plane_material = rng.integers(0, k, npl).astype(np.uint32)
    nb = int(rng.integers(0, 5)) if rng.random() < 0.5 else 0  # drawn by the rasterizer only
    sc.boxes = np.concatenate([rng.normal(0, spread, (nb, 3)), rng.uniform(0.05, 2.0, (nb, 3))], axis=1).astype(np.float32)
    sc.box_material = rng.integers(0, k, nb).astype(np.uint32)
    d = rng.normal(size=3); d[1] *= 0.3
    sc.camera = S.Camera(position=tuple(float(x) for x in rng.normal(0, spread * 1.5, 3) + [0, 1, 0]), direction=tuple(float(x) for x in d))
    return sc, rng


for i in range(n_scenes):
    sc, rng = random_scene(seed0 * 100003 + i)
    try:
        ctx.upload_scene(sc)
        w, h = int(rng.integers(9, 70)), int(rng.integers(5, 50))
        mode = int(rng.integers(0, 2))
        ref = None
        for label, flags, env in (("linear", nat.ACCEL_LINEAR, None), ("bvh", nat.ACCEL_BVH, None), ("wavefront", nat.ACCEL_AUTO | nat.PIPE_WAVEFRONT, None),
                                  ("pool", nat.ACCEL_BVH, ("RTCU_BVH_KERNEL", "pool"))):
            if env:
                os.environ[env[0]] = env[1]
                ctx.reload_env()
            v = make_view(sc, w, h, material_mode=mode, seed=1000 + i, flags=flags)
            rgba8, accum = ctx.render(v, want_accum=True)
            segs = ctx.stats()["segments"]
            if env:
                del os.environ[env[0]]
                ctx.reload_env()
            if ref is None:
                ref = oracle.render(sc, v)
            r_rgba8, r_accum, r_segs = ref
            stats["renders"] += 1
            if segs != r_segs:
                stats["seg_mismatch"] += 1
                stats["failures"].append((i, label, "segments", segs, r_segs))
            denom = np.maximum(np.abs(r_accum[..., :3]), 1e-3)
            stats["accum_max_rel"] = max(stats["accum_max_rel"], float((np.abs(accum[..., :3] - r_accum[..., :3]) / denom).max()))
            lsb = int(np.abs(((rgba8[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int) - ((r_rgba8[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int)).max())
            stats["rgba_max_lsb"] = max(stats["rgba_max_lsb"], lsb)
        for o, d in (synth.random_rays(sc, 4096, seed=i, spread=10.0), synth.grazing_rays(sc, 4096, seed=i + 7)):
            lin = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
            bvh = ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)
            orc = oracle.intersect_batch(sc, o, d)
            stats["rays"] += len(o)
            stats["hit_mismatch"] += int(sum((a != b).sum() for a, b in zip(lin, orc)))
            stats["bvh_vs_linear_mismatch"] += int(sum((a != b).sum() for a, b in zip(bvh, lin)))
        # the lanes-share-a-pixel BVH path (>= 16 samples per call; 8 or 16 lanes per pixel), every third scene
        if i % 3 == 0:
            spp_d = int(rng.choice([16, 24, 33, 70]))
            vd = make_view(sc, w, h, samples_per_pixel=spp_d, material_mode=mode, seed=2000 + i, flags=nat.ACCEL_BVH)
            rgba8, accum = ctx.render(vd, want_accum=True)
            segs = ctx.stats()["segments"]
            r_rgba8, r_accum, r_segs = oracle.render(sc, vd)
            stats["renders"] += 1
            stats["direct_renders"] = stats.get("direct_renders", 0) + 1
            if segs != r_segs or ctx.stats()["kernel_launches"] != 1:
                stats["seg_mismatch"] += 1
                stats["failures"].append((i, "bvh_direct", "segments", segs, r_segs))
            denom = np.maximum(np.abs(r_accum[..., :3]), 1e-3)
            stats["accum_max_rel"] = max(stats["accum_max_rel"], float((np.abs(accum[..., :3] - r_accum[..., :3]) / denom).max()))
            lsb = int(np.abs(((rgba8[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int) - ((r_rgba8[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int)).max())
            stats["rgba_max_lsb"] = max(stats["rgba_max_lsb"], lsb)
        pv = make_view(sc, w, h)
        expect = oracle.rasterize(sc, pv)
        for accel in (nat.ACCEL_LINEAR, nat.ACCEL_BVH):
            pv.flags = accel
            got = ctx.rasterize(pv, want_prim=True, want_depth=True)
            stats["raster_frames"] += 1
            bad = int((got[0] != expect[0]).sum() + (got[1] != expect[1]).sum() + (got[2].view(np.uint32) != expect[2].view(np.uint32)).sum())
            if bad:
                stats["raster_mismatch"] += bad
                stats["failures"].append((i, "raster", accel, bad))
        stats["scenes"] += 1
    except Exception as e:  # keep going, report at the end
        stats["failures"].append((i, "exception", str(e)[:200]))
stats["failures"] = stats["failures"][:20]
print(json.dumps(stats))
