"""BASELINE.json's larger configurations at their full sizes, through size-independent properties (the CPU oracle cannot
finish them): determinism, tile and sample-range composition, BVH == linear scan on sub-rectangles, and a strided row
subset against the oracle where that is affordable."""
import numpy as np
import pytest

from rt_b200 import _native as nat, synth
from rt_b200.renderer import make_view

from conftest import unpack_rgba

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return synth.rtiow_scene()


@pytest.fixture(scope="module")
def c4():
    return synth.grid_scene()  # 100 001 spheres


def test_c3_full_size(ctx, oracle, c3):
    # configs[2]: ~484 spheres, 1920x1080, 256 spp, depth 50 (530.8 M samples)
    assert 470 <= len(c3.spheres) <= 500
    ctx.upload_scene(c3)
    kw = dict(samples_per_pixel=256, max_bounces=50, material_mode=nat.MODE_SM)
    v = make_view(c3, 1920, 1080, **kw)
    rgba_a, accum_a = ctx.render(v, want_accum=True)
    st = ctx.stats()
    assert st["accel"] == nat.ACCEL_BVH and st["samples"] == 1920 * 1080 * 256
    assert (accum_a[..., 3] == 256).all() and np.isfinite(accum_a).all()
    # deterministic
    rgba_b, accum_b = ctx.render(v, want_accum=True)
    np.testing.assert_array_equal(accum_a, accum_b)
    assert ctx.stats()["segments"] == st["segments"]
    # a tile through the linear scan equals the same pixels of the BVH frame (paths are bit-identical; the straggler pass
    # may re-order the sum of a few pixels, hence the tolerance instead of equality)
    tile = (800, 600, 928, 664)
    _, accum_t = ctx.render(make_view(c3, 1920, 1080, tile=tile, flags=nat.ACCEL_LINEAR, **kw), want_accum=True)
    np.testing.assert_allclose(accum_t[600:664, 800:928], accum_a[600:664, 800:928], rtol=2e-5, atol=1e-5)
    # sample ranges: [0,100) + [100,256) compose to the frame
    total = np.zeros_like(accum_t[600:664, 800:928])
    for rng in ((0, 100), (100, 256)):
        _, part = ctx.render(make_view(c3, 1920, 1080, tile=tile, sample_range=rng, **kw), want_accum=True)
        total += part[600:664, 800:928]
    np.testing.assert_allclose(total, accum_a[600:664, 800:928], rtol=2e-5, atol=1e-5)  # 256-term fp32 sums, different order
    # three rows against the oracle at the full 256 spp / depth 50
    r_rgba8, r_accum, _ = oracle.render(c3, v, row_step=400)
    rows = np.arange(0, 1080, 400)
    np.testing.assert_allclose(accum_a[rows][..., :3], r_accum[rows][..., :3], rtol=5e-5, atol=1e-4)
    assert np.abs(unpack_rgba(rgba_a[rows]) - unpack_rgba(r_rgba8[rows])).max() <= 1


def test_c4_full_size(ctx, c4, knobs):
    # configs[3]: 100 001 spheres, 3840x2160, 64 spp, depth 10 (530.8 M samples) -- BVH traversal + divergence
    assert len(c4.spheres) == 100001
    ctx.upload_scene(c4)
    kw = dict(samples_per_pixel=64, max_bounces=10, material_mode=nat.MODE_SM)
    v = make_view(c4, 3840, 2160, **kw)
    rgba_a, _ = ctx.render(v)
    st = ctx.stats()
    assert st["accel"] == nat.ACCEL_BVH and st["samples"] == 3840 * 2160 * 64
    assert st["sphere_tests"] < 30 * st["segments"] and st["node_visits"] < 60 * st["segments"]  # it culls: 100 001 tests/segment otherwise
    assert ((rgba_a & 0xFF) == 0xFF).all() and len(np.unique(rgba_a)) > 1000
    rgba_b, _ = ctx.render(v)
    np.testing.assert_array_equal(rgba_a, rgba_b)
    assert ctx.stats()["segments"] == st["segments"]
    # BVH == brute-force scan over all 100 001 spheres on a tile (the reference's O(N) loop, mg_ray_tracer.cpp:62-87)
    tile = (1900, 1200, 1964, 1232)
    _, acc_bvh = ctx.render(make_view(c4, 3840, 2160, tile=tile, flags=nat.ACCEL_BVH, **kw), want_accum=True)
    segs_bvh = ctx.stats()["segments"]
    rgba_lin, acc_lin = ctx.render(make_view(c4, 3840, 2160, tile=tile, flags=nat.ACCEL_LINEAR, **kw), want_accum=True)
    assert ctx.stats()["segments"] == segs_bvh and ctx.stats()["sphere_tests"] == segs_bvh * 100001  # the very same paths
    # at >= 64 spp the BVH path gives every pixel to a warp whose lanes share its samples (launch_render, direct mode): the
    # per-pixel sum is a butterfly over 32 lane sums instead of the scan kernel's sequential sum -- fp32 order only
    np.testing.assert_array_equal(acc_bvh[..., 3], acc_lin[..., 3])
    np.testing.assert_allclose(acc_bvh[..., :3], acc_lin[..., :3], rtol=4e-6, atol=1e-6)
    assert np.abs(unpack_rgba(rgba_lin[1200:1232, 1900:1964]) - unpack_rgba(rgba_a[1200:1232, 1900:1964])).max() <= 1
    # with the thread-per-pixel kernel on both sides the buffers are bit-identical (a tile this small would otherwise take the
    # lanes-share-a-pixel kernel on the scan side too)
    knobs(RTCU_BVH_DIRECT="0", RTCU_SCAN_DIRECT="0")
    _, acc_bvh_tpp = ctx.render(make_view(c4, 3840, 2160, tile=tile, flags=nat.ACCEL_BVH, **kw), want_accum=True)
    _, acc_lin_tpp = ctx.render(make_view(c4, 3840, 2160, tile=tile, flags=nat.ACCEL_LINEAR, **kw), want_accum=True)
    knobs(RTCU_BVH_DIRECT=None, RTCU_SCAN_DIRECT=None)
    np.testing.assert_array_equal(acc_bvh_tpp, acc_lin_tpp)
    np.testing.assert_allclose(acc_lin[..., :3], acc_lin_tpp[..., :3], rtol=4e-6, atol=1e-6)
    # closest hits of 2^16 random + 2^16 silhouette-grazing rays: BVH == scan
    for o, d in (synth.random_rays(c4, 1 << 16, seed=3, spread=60.0), synth.grazing_rays(c4, 1 << 16, seed=4)):
        lin = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
        bvh = ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)
        for a, b in zip(bvh, lin):
            np.testing.assert_array_equal(a, b)


def test_c5_sample_range_slices_compose_at_4k(ctx, c3):
    # configs[4]: the C3 scene at 3840x2160, 4096 spp split by sample range; two of the 8-GPU slices, checked on a band
    ctx.upload_scene(c3)
    kw = dict(samples_per_pixel=4096, max_bounces=50, material_mode=nat.MODE_SM, tile=(0, 1000, 3840, 1016))
    parts = []
    segs = 0
    for g in (0, 1):
        _, acc = ctx.render(make_view(c3, 3840, 2160, sample_range=(g * 512, (g + 1) * 512), **kw), want_accum=True)
        parts.append(acc[1000:1016])
        segs += ctx.stats()["segments"]
        assert (parts[-1][..., 3] == 512).all()
    _, both = ctx.render(make_view(c3, 3840, 2160, sample_range=(0, 1024), **kw), want_accum=True)
    assert ctx.stats()["segments"] == segs
    np.testing.assert_allclose(parts[0] + parts[1], both[1000:1016], rtol=5e-5, atol=1e-4)  # 1024-term fp32 sums
