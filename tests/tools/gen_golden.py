#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the strict CPU oracle (oracle/rtref.c).

The reference ships no golden vectors, so these fixtures pin the oracle's own
output: a change to either the oracle or the CUDA path that alters results shows up against them.
Run from the repo root: `python tests/tests/tools/gen_golden.py`.
"""
import pathlib, sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle.binding import Oracle  # noqa: E402
from rt_b200 import scene as S, synth  # noqa: E402
from rt_b200.renderer import make_view  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
O = Oracle("strict")


def planes_scene():
    s = S.loads("""
camera = { position = [0, 1.5, 6], direction = [0, -0.1, -1] }
materials = [ { type = 'lambert', albedo = [0.8, 0.8, 0.2] }, { type = 'metal', albedo = [0.9,0.9,0.9], roughness = 0.1 },
              { type = 'dielectric', albedo = [0.6, 0.6, 0.6] } ]
planes = [ { material = 0 }, { material = 1, position = [0, 0, -6], normal = [0, 0, 1] }, { material = 0, position = [0, 0.25, 0] } ]
spheres = [ { material = 2, position = [0, 1, 0], radius = 1.0 }, { material = 1, position = [2.2, 0.5, 0.5] }, { material = 0, position = [-2, 0.5, 1] } ]
""")
    return s


SCENES = {
    "c1": (S.load(ROOT / "scenes" / "basic.toml"), 10),
    "c2": (S.load(ROOT / "scenes" / "dielectric.toml"), 50),
    "c3": (synth.rtiow_scene(), 50),
    "planes": (planes_scene(), 12),
}

REFBUILD_CASES = [  # (scene, reference renderer, material mode, width, height, spp, max_bounces)
    ("c1", "mg_ray_tracer", 0, 160, 120, 8, 10),
    ("c2", "sm_ray_tracer", 1, 192, 108, 16, 50),
    ("c2", "mg_ray_tracer", 0, 192, 108, 16, 50),
    ("planes", "sm_ray_tracer", 1, 128, 96, 8, 12),
    ("planes", "mg_ray_tracer", 0, 128, 96, 8, 12),
    ("c3", "sm_ray_tracer", 1, 96, 54, 4, 50),
]


def gen_reference_build_goldens():
    """Outputs of the reference's OWN renderer sources (oracle/_ref/librt_ref.so: mg_ray_tracer.cpp / sm_ray_tracer.cpp
    compiled where they lie against the muu stand-in, RNG replaced by the counter-based stream).  Only possible where
    /root/reference exists; the fixtures are committed so the pin travels."""
    from oracle.binding import ReferenceBuild

    if not ReferenceBuild.available():
        print("reference build unavailable: refbuild_* fixtures not regenerated")
        return
    ref = ReferenceBuild()
    for name, renderer, mode, w, h, spp, depth in REFBUILD_CASES:
        sc = SCENES[name][0]
        rgba8, ivp = ref.render(sc, w, h, spp, depth, 0x5EED, renderer, threads=0)
        np.savez_compressed(OUT / f"refbuild_{name}_{renderer}.npz", rgba8=rgba8, inv_view_proj=ivp, width=w, height=h, spp=spp,
                            max_bounces=depth, mode=mode, seed=np.uint64(0x5EED))
        print("reference build", name, renderer, rgba8.shape)


RASTER_SCENES = {**{k: v[0] for k, v in SCENES.items()}, "boxes": S.load(ROOT / "scenes" / "boxes.toml")}
RASTER_CASES = [("c1", 160, 120), ("c2", 192, 108), ("planes", 128, 96), ("c3", 96, 54), ("boxes", 200, 125)]  # (scene, width, height)


def gen_raster_goldens():
    """rgba8 = the reference's OWN rasterizer.cpp (compiled against the muu stand-in); prim / depth = the oracle's per-pixel
    record for the same matrix (the reference has no such output)."""
    from oracle.binding import ReferenceBuild

    if not ReferenceBuild.available():
        print("reference build unavailable: raster_* fixtures not regenerated")
        return
    ref = ReferenceBuild()
    for name, w, h in RASTER_CASES:
        sc = RASTER_SCENES[name]
        rgba8, ivp = ref.render(sc, w, h, 1, 1, 0, "rasterizer", threads=0)
        v = make_view(sc, w, h)
        v.inv_view_proj[:] = ivp.tolist()
        o_rgba8, prim, depth = O.rasterize(sc, v, threads=0)
        assert np.array_equal(o_rgba8, rgba8), (name, int((o_rgba8 != rgba8).sum()))
        np.savez_compressed(OUT / f"raster_{name}.npz", rgba8=rgba8, prim=prim, depth=depth, inv_view_proj=ivp, width=w, height=h)
        print("raster", name, rgba8.shape, "hit fraction", float((prim != 0xFFFFFFFF).mean()))


if __name__ == "__main__":
    if "--raster-only" in sys.argv:
        gen_raster_goldens()
        sys.exit(0)
    gen_reference_build_goldens()
    gen_raster_goldens()
    for name, (sc, depth) in SCENES.items():
        o, d = synth.random_rays(sc, 4096, seed=7)
        hit, prim, t, nrm = O.intersect_batch(sc, o, d)
        np.savez_compressed(OUT / f"rays_{name}.npz", o=o, d=d, hit=hit, prim=prim, t=t, normal=nrm)
        print(name, "rays hit fraction", hit.mean())
        w, h, spp = (64, 48, 8) if name != "c3" else (64, 36, 4)
        for mode in (0, 1):
            v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=depth, material_mode=mode, seed=0x5EED)
            rgba8, accum, segs = O.render(sc, v, threads=0)
            np.savez_compressed(OUT / f"image_{name}_{'mg' if mode == 0 else 'sm'}.npz", rgba8=rgba8, accum=accum, segments=np.uint64(segs),
                                width=w, height=h, spp=spp, max_bounces=depth, mode=mode, seed=np.uint64(0x5EED))
            print(name, mode, "segments/sample", segs / (w * h * spp))
