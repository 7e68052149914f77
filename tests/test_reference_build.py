"""Pins the oracle against marzer/rt's OWN renderer sources.

oracle/_ref/librt_ref.so is src/renderers/mg_ray_tracer.cpp, src/renderers/sm_ray_tracer.cpp and src/renderer.cpp of the
reference, compiled where they lie (oracle/Makefile `ref`) against oracle/ref_shim -- a stand-in for the un-vendored muu
math library -- with rt::detail::random_float() replaced by the counter-based stream.  Everything the reference's renderer
files themselves decide (closest-hit loops and tie rules, select, the scatter tables, lambert / metal / dielectric /
Schlick, draw order, recursion and attenuation nesting, per-pixel accumulation, gamma, packing, colour quirks) is therefore
the reference's code, not a restatement.  The oracle must reproduce its packed images bit for bit.  What stays unverified is
muu itself (vector / matrix / ray primitives), which both sides take from the same numbered SPEC.

tests/golden/refbuild_*.npz are outputs of that build (tests/tools/gen_golden.py), committed so the pin holds where the reference
tree is absent; the live tests run wherever the prebuilt library (or the reference tree) is available."""
import numpy as np
import pytest

from rt_b200 import scene as S
from rt_b200.renderer import make_view

from conftest import GOLDEN


def _cases():
    import sys
    sys.path.insert(0, str(GOLDEN.parent / "tools"))
    import gen_golden

    return gen_golden.REFBUILD_CASES


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c[0]}-{c[1]}")
def test_oracle_reproduces_reference_build_fixtures_bit_for_bit(oracle, scenes, case):
    name, renderer, mode, w, h, spp, depth = case
    g = np.load(GOLDEN / f"refbuild_{name}_{renderer}.npz")
    sc = scenes[name][0]
    v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=depth, material_mode=mode, seed=int(g["seed"]))
    v.inv_view_proj[:] = g["inv_view_proj"].tolist()  # the matrix the reference's own camera produced
    rgba8, _, _ = oracle.render(sc, v, threads=2, want_accum=False)
    np.testing.assert_array_equal(rgba8, g["rgba8"])


@pytest.fixture(scope="module")
def refbuild():
    from oracle.binding import ReferenceBuild

    if not ReferenceBuild.available():
        pytest.skip("oracle/_ref/librt_ref.so not present and no reference tree to build it from")
    return ReferenceBuild()


def test_reference_build_registers_both_ray_tracers(refbuild):
    assert refbuild.renderers() == ["mg_ray_tracer", "sm_ray_tracer", "rasterizer"]  # REGISTER_RENDERER ran at load time


def _random_scene(seed: int) -> S.Scene:
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 12))
    mats = [(int(rng.integers(0, 8)), tuple(rng.uniform(0.1, 1.0, 3)), float(rng.uniform(0, 0.6)), float(rng.choice([0.5, 0.8, 1.0, 1.31, 1.52])))
            for _ in range(int(rng.integers(1, 6)))]
    sc = S.Scene(samples_per_pixel=4, max_bounces=int(rng.integers(1, 30)))
    sc.materials = S.make_materials(mats)
    sph = np.concatenate([rng.uniform(-3, 3, (n, 3)), rng.uniform(0.2, 1.2, (n, 1))], axis=1)
    sph[0] = [0, -100.5, 0, 100]
    sc.spheres = sph.astype(np.float32)
    sc.sphere_material = rng.integers(0, len(mats), n).astype(np.uint32)
    if seed % 2:
        sc.planes = np.array([[0, 1, 0, 1.0], [0, 0, 1, 6.0]], np.float32)
        sc.plane_material = rng.integers(0, len(mats), 2).astype(np.uint32)
    sc.camera = S.Camera(position=(float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 2)), 6.0), direction=(0.0, float(rng.uniform(-0.3, 0.1)), -1.0))
    return sc


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("renderer,mode", [("mg_ray_tracer", 0), ("sm_ray_tracer", 1)])
def test_oracle_equals_reference_build_on_random_scenes(oracle, refbuild, seed, renderer, mode):
    sc = _random_scene(seed)
    w, h = 72, 48
    ref, ivp = refbuild.render(sc, w, h, sc.samples_per_pixel, sc.max_bounces, 1234 + seed, renderer, threads=2)
    v = make_view(sc, w, h, material_mode=mode, seed=1234 + seed)
    v.inv_view_proj[:] = ivp.tolist()
    rgba8, _, _ = oracle.render(sc, v, threads=2, want_accum=False)
    np.testing.assert_array_equal(rgba8, ref)


def test_reference_camera_matches_harness_camera(refbuild, scenes):
    # rt_b200.camera restates camera::viewport in float64; the stand-in matrix code runs the reference's camera.hpp.
    # Both describe the same projection: unprojected points agree to float precision of the far plane (1000).
    from rt_b200.camera import inverse_view_projection

    for name in ("c1", "c2", "c3"):
        sc = scenes[name][0]
        a = inverse_view_projection(sc.camera, 640, 360).reshape(4, 4).T.astype(np.float64)
        b = refbuild.inverse_view_projection(sc, 640, 360).reshape(4, 4).T.astype(np.float64)
        for ndc in ([0, 0, 0, 1], [0.7, -0.4, 0, 1], [-1, 1, 1, 1], [0.3, 0.2, 1, 1]):
            pa, pb = a @ ndc, b @ ndc
            np.testing.assert_allclose(pa[:3] / pa[3], pb[:3] / pb[3], rtol=2e-3, atol=2e-3)


def test_reference_build_with_its_own_mt19937_agrees_statistically(refbuild, scenes):
    """oracle/_ref/librt_ref_fast_mt.so links the reference's own src/random.cpp (thread_local mt19937, random_device seed)
    instead of the counter-based stand-in: the renderer exactly as shipped.  Non-deterministic, so only the statistics can
    agree: same mean colour, small mean absolute difference at 64 spp."""
    from oracle.binding import ReferenceBuild

    if not ReferenceBuild.FAST_MT_PATH.exists():
        pytest.skip("librt_ref_fast_mt.so not built")
    sc = scenes["c2"][0]
    w, h, spp = 160, 90, 64
    a, _ = ReferenceBuild("fast_mt").render(sc, w, h, spp, 50, 0, "sm_ray_tracer", threads=2)
    b, _ = refbuild.render(sc, w, h, spp, 50, 0x5EED, "sm_ray_tracer", threads=2)
    ca = np.stack([(a >> s) & 255 for s in (24, 16, 8)], -1).astype(float)
    cb = np.stack([(b >> s) & 255 for s in (24, 16, 8)], -1).astype(float)
    assert np.abs(ca.mean(axis=(0, 1)) - cb.mean(axis=(0, 1))).max() < 1.0
    assert np.abs(ca - cb).mean() < 6.0
    assert (a != b).any()  # a different stream, not the same image
