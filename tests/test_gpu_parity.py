"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): hit/miss flags and closest-hit primitive indices bit-exact, hit t within
4 ulp (we observe and require 0: both sides implement the same numbered IEEE spec), images within a stated
RMSE at equal spp with the same RNG streams.  Image tolerance: the paths are bit-identical (asserted through the
exact segment count); only the radiance product differs -- the oracle nests it right-to-left like the reference's
recursion (mg_ray_tracer.cpp:171), the kernel folds it left-to-right -- so the fp32 sums agree to ~1e-6 relative
and the packed RGBA8 image may differ by 1 LSB in isolated pixels.  Stated bounds: accum rtol 2e-5, RGBA8
max |diff| <= 1 on <= 0.5 % of channel values, RMSE on the sqrt-encoded [0,1] image <= 1e-4.
"""
import numpy as np
import pytest

from rt_b200 import _native as nat, scene as S, synth
from rt_b200.renderer import ImageView, cuda_path_tracer, make_view

from conftest import GOLDEN, ulp_diff, unpack_rgba

pytestmark = pytest.mark.gpu

ACCUM_RTOL = 2e-5
RMSE_BOUND = 1e-4


def assert_images_match(rgba_gpu, accum_gpu, rgba_ref, accum_ref, spp):
    np.testing.assert_array_equal(accum_gpu[..., 3], accum_ref[..., 3])
    np.testing.assert_allclose(accum_gpu[..., :3], accum_ref[..., :3], rtol=ACCUM_RTOL, atol=1e-6)
    d = np.abs(unpack_rgba(rgba_gpu) - unpack_rgba(rgba_ref))
    assert d.max() <= 1, f"RGBA8 differs by {d.max()} LSB"
    assert (d > 0).mean() <= 0.005, f"{(d > 0).mean():.4%} of channel values differ"
    enc = lambda a: np.sqrt(np.clip(a[..., :3] / spp, 0, 1))
    rmse = float(np.sqrt(np.mean((enc(accum_gpu) - enc(accum_ref)) ** 2)))
    assert rmse <= RMSE_BOUND, rmse


# ---- level 0: RNG -------------------------------------------------------------------------------------
def test_philox_matches_published_vectors_and_oracle(ctx, oracle):
    from test_oracle_kat import PHILOX_KAT

    for ctr, key, expect in PHILOX_KAT:
        out = ctx.philox_batch(np.array([ctr], np.uint32), key[0] | (key[1] << 32))
        assert tuple(int(x) for x in out[0]) == expect
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2 ** 32, (4096, 4), dtype=np.uint64).astype(np.uint32)
    out = ctx.philox_batch(ctr, 0x0123456789ABCDEF)
    for i in range(0, 4096, 97):
        np.testing.assert_array_equal(out[i], oracle.philox(ctr[i], 0x0123456789ABCDEF))


# ---- level 1: closest hit on fixed ray batches ------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c3", "planes"])
def test_intersect_golden_batches(ctx, scenes, name):
    g = np.load(GOLDEN / f"rays_{name}.npz")
    ctx.upload_scene(scenes[name][0])
    hit, prim, t, nrm = ctx.intersect_batch(g["o"], g["d"], accel=nat.ACCEL_LINEAR)
    np.testing.assert_array_equal(hit, g["hit"])
    np.testing.assert_array_equal(prim, g["prim"])
    assert ulp_diff(t, g["t"]).max() == 0
    np.testing.assert_array_equal(nrm, g["normal"])


@pytest.mark.parametrize("name,n", [("c1", 1 << 20), ("c2", 1 << 20), ("c3", 1 << 18), ("planes", 1 << 18)])
def test_intersect_large_seeded_batches_vs_oracle(ctx, oracle, scenes, name, n):
    sc = scenes[name][0]
    o, d = synth.random_rays(sc, n, seed=11)
    ctx.upload_scene(sc)
    hit, prim, t, nrm = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
    rh, rp, rt, rn = oracle.intersect_batch(sc, o, d)
    np.testing.assert_array_equal(hit, rh)
    np.testing.assert_array_equal(prim, rp)
    assert ulp_diff(t, rt).max() <= 4
    assert ulp_diff(t, rt).max() == 0
    np.testing.assert_array_equal(nrm, rn)
    assert 0.2 < hit.mean() < 0.95


def test_intersect_edge_cases(ctx, oracle):
    from test_oracle_kat import _scene

    # empty scene, ragged batch sizes (not a multiple of the block), n = 0
    ctx.upload_scene(_scene([]))
    hit, prim, t, _ = ctx.intersect_batch(np.zeros((5, 3), np.float32), np.tile(np.float32([0, 0, -1]), (5, 1)))
    assert not hit.any() and (prim == nat.PRIM_MISS).all() and (t == -1).all()
    hit, prim, t, _ = ctx.intersect_batch(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(hit) == 0
    # ties, inside / behind / tangent cases: identical to the oracle
    sc = _scene([(0, 0, -5, 1), (0, 0, -5, 1), (0, 0, -9, 1), (0, 0, 0, 2), (0, 0.5, 0, 0.5)], planes=[(0, 1, 0, 3), (0, 1, 0, 3)])
    ctx.upload_scene(sc)
    o = np.float32([(0, 0, 0), (0, 0, 0), (0, 1, 3), (0, 0, 10), (0, 0, 2.0001), (0, 20, 0), (1, 1, 1)])
    d = np.float32([(0, 0, -1), (1, 0, 0), (0, 0, -1), (0, 0, 1), (0, 0, -1), (0, -1, 0), (0, -1, 0)])
    got = ctx.intersect_batch(o, d)
    ref = oracle.intersect_batch(sc, o, d)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


def test_many_spheres_fall_back_to_global_memory_scan(ctx, oracle):
    # more primitives than the shared-memory staging budget (200 KB / 16 B): the unstaged kernel variant
    sc = synth.grid_scene(nx=130, nz=100, seed=5)
    assert len(sc.spheres) * 16 > 200 * 1024
    o, d = synth.random_rays(sc, 2048, seed=5, spread=30.0)
    ctx.upload_scene(sc)
    got = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
    ref = oracle.intersect_batch(sc, o, d)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


# ---- level 1b: primary rays and scatter events, bit-exact --------------------------------------------------
def test_primary_rays_bit_exact(ctx, oracle, scenes):
    sc = scenes["c2"][0]
    v = make_view(sc, 1920, 1080, samples_per_pixel=64, max_bounces=50)
    rng = np.random.default_rng(5)
    n = 3000
    px = rng.integers(0, 1920, n); py = rng.integers(0, 1080, n); smp = rng.integers(0, 64, n)
    smp[:200] = 0  # pixel-centre rule
    o, d = ctx.primary_rays(v, px, py, smp)
    ro, rd = oracle.primary_rays(v, px, py, smp)
    np.testing.assert_array_equal(o, ro)
    np.testing.assert_array_equal(d, rd)


@pytest.mark.parametrize("mode", [nat.MODE_MG, nat.MODE_SM])
def test_scatter_bit_exact(ctx, oracle, scenes, mode):
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    rng = np.random.default_rng(8 + mode)
    n = 4000
    # realistic inputs: rays that hit something, with the oracle's t / normal / material
    o, d = synth.random_rays(sc, 4 * n, seed=21)
    hit, prim, t, nrm = oracle.intersect_batch(sc, o, d)
    sel = np.flatnonzero(hit)[:n]
    o, d, t, nrm, prim = o[sel], d[sel], t[sel], nrm[sel], prim[sel]
    mat = sc.sphere_material[prim]
    d = (d * rng.uniform(0.5, 2.0, (len(sel), 1))).astype(np.float32)  # dielectric bounces do not renormalise
    pixel = rng.integers(0, 2 ** 21, len(sel)); smp = rng.integers(0, 64, len(sel)); blk = rng.integers(1, 50, len(sel))
    s, att, oo, do = ctx.scatter_batch(mode, 0x5EED, mat, o, d, t, nrm, pixel, smp, blk)
    rs, ratt, roo, rdo = oracle.scatter_batch(sc, mode, 0x5EED, mat, o, d, t, nrm, pixel, smp, blk)
    np.testing.assert_array_equal(s, rs)
    np.testing.assert_array_equal(att, ratt)
    ok = s.astype(bool)
    np.testing.assert_array_equal(oo[ok], roo[ok])
    np.testing.assert_array_equal(do[ok], rdo[ok])
    kinds = {int(k) for k in np.unique(sc.materials["type"][mat])}
    assert {S.LAMBERT, S.METAL, S.DIELECTRIC} <= kinds


# ---- level 2: images --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c3", "planes"])
@pytest.mark.parametrize("mode", ["mg", "sm"])
def test_render_matches_golden_images(ctx, scenes, name, mode):
    g = np.load(GOLDEN / f"image_{name}_{mode}.npz")
    sc = scenes[name][0]
    ctx.upload_scene(sc)
    v = make_view(sc, int(g["width"]), int(g["height"]), samples_per_pixel=int(g["spp"]), max_bounces=int(g["max_bounces"]),
                  material_mode=int(g["mode"]), seed=int(g["seed"]))
    rgba8, accum = ctx.render(v, want_accum=True)
    st = ctx.stats()
    assert st["segments"] == int(g["segments"])  # identical paths, segment for segment
    assert st["samples"] == int(g["width"]) * int(g["height"]) * int(g["spp"])
    assert_images_match(rgba8, accum, g["rgba8"], g["accum"], int(g["spp"]))


def test_render_c1_default_frame_vs_oracle(ctx, oracle, scenes):
    # BASELINE configs[0] at its real size: 800x600, 30 spp, depth 10, mg table
    sc = scenes["c1"][0]
    ctx.upload_scene(sc)
    v = make_view(sc, 800, 600, material_mode=nat.MODE_MG)
    rgba8, accum = ctx.render(v, want_accum=True)
    r_rgba8, r_accum, r_segs = oracle.render(sc, v)
    assert ctx.stats()["segments"] == r_segs
    assert_images_match(rgba8, accum, r_rgba8, r_accum, 30)


def test_render_c2_full_size_properties(ctx, oracle, scenes, knobs):
    # BASELINE configs[1] at full size (1920x1080, 64 spp, depth 50, sm table): size-independent properties
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    v = make_view(sc, 1920, 1080, samples_per_pixel=64, max_bounces=50, material_mode=nat.MODE_SM)
    rgba_a, accum_a = ctx.render(v, want_accum=True)
    segs_a = ctx.stats()["segments"]
    rgba_b, accum_b = ctx.render(v, want_accum=True)
    # (1) deterministic: bit-identical repeat, identical segment count
    np.testing.assert_array_equal(accum_a, accum_b)
    np.testing.assert_array_equal(rgba_a, rgba_b)
    assert ctx.stats()["segments"] == segs_a
    assert (accum_a[..., 3] == 64).all() and np.isfinite(accum_a).all() and ((rgba_a & 0xFF) == 0xFF).all()
    # (2) tiles compose: a tile render equals the same pixels of the full frame -- the same paths; a tile this small shares a pixel
    # between 16 lanes and the frame between 8 (rtcu.cu, scan_direct_lanes), so the sums differ in fp32 order only, and bit for bit
    # when tile and frame are rendered the same way
    vt = make_view(sc, 1920, 1080, samples_per_pixel=64, max_bounces=50, material_mode=nat.MODE_SM, tile=(1000, 500, 1100, 540))
    rgba_t, accum_t = ctx.render(vt, want_accum=True)
    assert ctx.stats()["kernel_launches"] == 1
    np.testing.assert_array_equal(accum_t[500:540, 1000:1100, 3], accum_a[500:540, 1000:1100, 3])
    np.testing.assert_allclose(accum_t[500:540, 1000:1100, :3], accum_a[500:540, 1000:1100, :3], rtol=4e-6, atol=1e-6)
    assert np.abs(unpack_rgba(rgba_t[500:540, 1000:1100]) - unpack_rgba(rgba_a[500:540, 1000:1100])).max() <= 1
    assert (rgba_t[:500] == 0).all()
    for how in ("8", "0"):  # 8 lanes per pixel; the thread-per-pixel grid
        knobs(RTCU_SCAN_DIRECT=how)
        rgba_f, accum_f = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs_a
        rgba_t, accum_t = ctx.render(vt, want_accum=True)
        knobs(RTCU_SCAN_DIRECT=None)
        np.testing.assert_array_equal(accum_t[500:540, 1000:1100], accum_f[500:540, 1000:1100])
        np.testing.assert_array_equal(rgba_t[500:540, 1000:1100], rgba_f[500:540, 1000:1100])
        assert (rgba_t[:500] == 0).all()
        np.testing.assert_allclose(accum_f[..., :3], accum_a[..., :3], rtol=4e-6, atol=1e-6)
    # (3) a strided row subset against the oracle at the full spp / depth
    r_rgba8, r_accum, r_segs = oracle.render(sc, v, row_step=90)
    rows = np.arange(0, 1080, 90)
    assert_images_match(rgba_a[rows], accum_a[rows], r_rgba8[rows], r_accum[rows], 64)
    # (4) every sample is bounded by the brightest possible radiance (attenuations > 1 exist: IOR quirk)
    assert accum_a[..., :3].min() >= 0


def test_sample_ranges_compose(ctx, scenes):
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    full = make_view(sc, 160, 90, samples_per_pixel=16, max_bounces=50)
    _, accum = ctx.render(full, want_accum=True)
    total = np.zeros_like(accum)
    segs = 0
    for b, e in ((0, 5), (5, 6), (6, 16)):
        v = make_view(sc, 160, 90, samples_per_pixel=16, max_bounces=50, sample_range=(b, e))
        _, part = ctx.render(v, want_accum=True)
        assert (part[..., 3] == e - b).all()
        total += part
        segs += ctx.stats()["segments"]
    ctx.render(full, want_rgba8=False, want_accum=False)
    assert segs == ctx.stats()["segments"]  # same paths, partitioned
    np.testing.assert_allclose(total, accum, rtol=2e-6, atol=1e-7)  # only the summation order differs
    # an empty sample range renders nothing
    v = make_view(sc, 160, 90, samples_per_pixel=16, sample_range=(4, 4))
    _, part = ctx.render(v, want_accum=True)
    assert (part == 0).all() and ctx.stats()["segments"] == 0


def test_odd_sizes_and_one_pixel(ctx, oracle, scenes):
    sc = scenes["c1"][0]
    ctx.upload_scene(sc)
    for w, h in ((1, 1), (17, 5), (33, 31)):
        v = make_view(sc, w, h, samples_per_pixel=3, max_bounces=4, material_mode=nat.MODE_MG)
        rgba8, accum = ctx.render(v, want_accum=True)
        r_rgba8, r_accum, r_segs = oracle.render(sc, v)
        assert ctx.stats()["segments"] == r_segs
        assert_images_match(rgba8, accum, r_rgba8, r_accum, 3)


# ---- the boundary: plugin-style calls and error behaviour -------------------------------------------------
def test_plugin_render_into_image_view(ctx, oracle, scenes):
    r = cuda_path_tracer(0)
    r.material_mode = nat.MODE_MG
    sc = S.load("scenes/basic.toml")
    sc.samples_per_pixel = 4
    img = ImageView.allocate(96, 64).clear(0x000000FF)
    r.render(sc, img)
    assert r.last_error is None
    v = make_view(sc, 96, 64, material_mode=nat.MODE_MG)
    ref, _, _ = oracle.render(sc, v)
    d = np.abs(unpack_rgba(img.data) - unpack_rgba(ref))
    assert d.max() <= 1
    # scene change is detected by content (no dirty flag in the reference, main.cpp:123-125)
    sc.spheres[1, 1] += 1.0
    before = img.data.copy()
    r.render(sc, img)
    assert (img.data != before).any()


def test_plugin_render_is_noexcept_and_leaves_image_black(ctx, capsys):
    r = cuda_path_tracer(0)
    sc = S.load("scenes/basic.toml")
    sc.sphere_material = np.array([0, 1, 9], np.uint32)  # out-of-range material
    img = ImageView.allocate(32, 16).clear(0x000000FF)
    r.render(sc, img)  # must not raise
    assert r.last_error and "out-of-range" in r.last_error
    assert (img.data == 0x000000FF).all()
    assert "cuda_path_tracer" in capsys.readouterr().err


def test_error_codes(ctx, scenes):
    from rt_b200.renderer import Context

    fresh = Context(0)
    v = make_view(scenes["c1"][0], 16, 16)
    with pytest.raises(nat.RtcuError) as e:
        fresh.render(v)
    assert e.value.code == nat.RTCU_ERR_STATE
    fresh.upload_scene(scenes["c1"][0])
    for bad in (dict(tile=(4, 4, 4, 8)), dict(tile=(0, 0, 17, 16)), dict(sample_range=(3, 2)), dict(max_bounces=0), dict(samples_per_pixel=0)):
        with pytest.raises(nat.RtcuError) as e:
            fresh.render(make_view(scenes["c1"][0], 16, 16, **bad))
        assert e.value.code == nat.RTCU_ERR_INVALID
    fresh.close()


# ---- BVH traversal: the linear scan's exact result ----------------------------------------------------------
def _grid():
    return synth.grid_scene(nx=120, nz=80, seed=9)  # 9601 spheres + ground


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "planes", "grid"])
def test_bvh_matches_linear_scan_exactly(ctx, oracle, scenes, name):
    sc = _grid() if name == "grid" else scenes[name][0]
    ctx.upload_scene(sc)
    batches = [synth.random_rays(sc, 1 << 17, seed=31, spread=40.0 if name == "grid" else 12.0), synth.grazing_rays(sc, 1 << 17, seed=32)]
    for o, d in batches:
        lin = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
        bvh = ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)
        st = ctx.stats()
        for a, b in zip(bvh, lin):
            np.testing.assert_array_equal(a, b)
        assert st["accel"] == nat.ACCEL_BVH and st["node_visits"] > 0
        if len(sc.spheres) > 400:
            assert st["sphere_tests"] < 0.2 * len(o) * len(sc.spheres)  # it actually culls
    # and against the oracle on a subset (the linear kernel is already pinned to it)
    o, d = batches[1][0][:20000], batches[1][1][:20000]
    got = ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)
    ref = oracle.intersect_batch(sc, o, d)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


def test_bvh_non_unit_directions_fall_back_to_the_scan(ctx, scenes):
    sc = scenes["c3"][0]
    ctx.upload_scene(sc)
    o, d = synth.random_rays(sc, 4096, seed=5)
    d = (d * np.float32(1.7)).astype(np.float32)
    lin = ctx.intersect_batch(o, d, accel=nat.ACCEL_LINEAR)
    bvh = ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)
    for a, b in zip(bvh, lin):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("kernel", ["pool", "mega"])
@pytest.mark.parametrize("name,mode", [("c3", nat.MODE_SM), ("c2", nat.MODE_SM), ("planes", nat.MODE_SM), ("grid", nat.MODE_SM), ("grid", nat.MODE_MG)])
def test_bvh_render_matches_linear_render(ctx, scenes, knobs, name, mode, kernel):
    # Both BVH kernels trace exactly the linear scan's paths (same segment count).  "mega" (one segment per iteration) also
    # sums every pixel's samples in the same order -> bit-identical buffers; "pool" (warp-local ray pool, not the default) sums
    # them in completion order -> identical up to fp32 summation order, and deterministic.  (The scan side is held to its
    # thread-per-pixel kernel: frames this small otherwise share each pixel between lanes, which re-orders the sums.)
    knobs(RTCU_BVH_KERNEL=kernel, RTCU_SCAN_DIRECT="0", RTCU_BVH_DIRECT="0")
    sc = _grid() if name == "grid" else scenes[name][0]
    ctx.upload_scene(sc)
    for w, h, spp in ((192, 108, 4), (61, 37, 9)):  # the second size leaves partial 8x4 patches and partial CTAs
        kw = dict(samples_per_pixel=spp, max_bounces=20, material_mode=mode)
        _, lin = ctx.render(make_view(sc, w, h, flags=nat.ACCEL_LINEAR, **kw), want_accum=True)
        segs_lin = ctx.stats()["segments"]
        rgba, bvh = ctx.render(make_view(sc, w, h, flags=nat.ACCEL_BVH, **kw), want_accum=True)
        st = ctx.stats()
        assert st["accel"] == nat.ACCEL_BVH and st["segments"] == segs_lin
        if kernel == "mega":
            np.testing.assert_array_equal(bvh, lin)
        else:
            np.testing.assert_array_equal(bvh[..., 3], lin[..., 3])
            np.testing.assert_allclose(bvh[..., :3], lin[..., :3], rtol=2e-6, atol=1e-6)
            rgba2, bvh2 = ctx.render(make_view(sc, w, h, flags=nat.ACCEL_BVH, **kw), want_accum=True)
            np.testing.assert_array_equal(bvh2, bvh)  # deterministic


def test_bvh_pool_tiles_and_sample_ranges(ctx, scenes, knobs):
    knobs(RTCU_BVH_KERNEL="pool")
    sc = scenes["c3"][0]
    ctx.upload_scene(sc)
    kw = dict(samples_per_pixel=12, max_bounces=50, material_mode=nat.MODE_SM, flags=nat.ACCEL_BVH)
    _, full = ctx.render(make_view(sc, 160, 96, **kw), want_accum=True)
    segs = ctx.stats()["segments"]
    _, tile = ctx.render(make_view(sc, 160, 96, tile=(21, 10, 77, 59), **kw), want_accum=True)
    np.testing.assert_allclose(tile[10:59, 21:77], full[10:59, 21:77], rtol=2e-6, atol=1e-6)
    assert (tile[:10] == 0).all() and (tile[:, :21] == 0).all() and (tile[59:] == 0).all()
    total, s = np.zeros_like(full), 0
    for rng in ((0, 5), (5, 12)):
        _, part = ctx.render(make_view(sc, 160, 96, sample_range=rng, **kw), want_accum=True)
        total += part
        s += ctx.stats()["segments"]
    assert s == segs
    np.testing.assert_allclose(total, full, rtol=2e-6, atol=1e-6)


def test_bvh_auto_threshold(ctx, scenes):
    thr = nat.load_library().rtcu_bvh_threshold()
    for name in ("c2", "c3"):
        sc = scenes[name][0]
        ctx.upload_scene(sc)
        ctx.render(make_view(sc, 32, 32, samples_per_pixel=1, max_bounces=2), want_accum=False)
        assert ctx.stats()["accel"] == (nat.ACCEL_BVH if len(sc.spheres) >= thr else nat.ACCEL_LINEAR)


# ---- the CUDA path against outputs of the reference's own renderer sources (oracle/_ref, see test_reference_build.py) --
def test_render_matches_reference_build_fixtures(ctx, scenes):
    import sys
    sys.path.insert(0, str(GOLDEN.parent / "tools"))
    import gen_golden

    for name, renderer, mode, w, h, spp, depth in gen_golden.REFBUILD_CASES:
        g = np.load(GOLDEN / f"refbuild_{name}_{renderer}.npz")
        sc = scenes[name][0]
        ctx.upload_scene(sc)
        v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=depth, material_mode=mode, seed=int(g["seed"]))
        v.inv_view_proj[:] = g["inv_view_proj"].tolist()
        rgba8, _ = ctx.render(v)
        d = np.abs(unpack_rgba(rgba8) - unpack_rgba(g["rgba8"]))
        # identical paths; only the left- vs right-nested radiance product differs (S12): isolated 1-LSB differences
        assert d.max() <= 1, (name, renderer, int(d.max()))
        assert (d > 0).mean() <= 0.005, (name, renderer, float((d > 0).mean()))


# ---- wavefront pipeline (generate / intersect + material sort / shade + compact over HBM queues) ------------------
@pytest.mark.parametrize("name,mode,accel", [("c2", nat.MODE_SM, nat.ACCEL_LINEAR), ("c1", nat.MODE_MG, nat.ACCEL_LINEAR),
                                             ("planes", nat.MODE_SM, nat.ACCEL_LINEAR), ("c3", nat.MODE_SM, nat.ACCEL_BVH),
                                             ("c3", nat.MODE_MG, nat.ACCEL_LINEAR)])
def test_wavefront_is_bit_identical_to_the_megakernel(ctx, scenes, knobs, name, mode, accel):
    # same paths, and per-pixel sums in sample order in both pipelines (the megakernel's straggler pass, which re-orders the
    # sum of the few pixels it takes over, is switched off for this comparison)
    knobs(RTCU_STRAGGLER_BUDGET="0", RTCU_SCAN_DIRECT="0", RTCU_BVH_DIRECT="0")  # (and so are the lanes-share-a-pixel kernels)
    knobs(RTCU_WF_RAYS=str(200 * 120 * 3))  # 3 samples per wave: 6 spp = 2 waves, 7 spp = 2 full + 1 partial
    sc = scenes[name][0]
    ctx.upload_scene(sc)
    for spp in (6, 7):
        kw = dict(samples_per_pixel=spp, max_bounces=scenes[name][1], material_mode=mode)
        rgba_m, acc_m = ctx.render(make_view(sc, 200, 120, flags=accel | nat.PIPE_MEGAKERNEL, **kw), want_accum=True)
        st_m = ctx.stats()
        rgba_w, acc_w = ctx.render(make_view(sc, 200, 120, flags=accel | nat.PIPE_WAVEFRONT, **kw), want_accum=True)
        st_w = ctx.stats()
        assert st_w["pipeline"] == nat.PIPE_WAVEFRONT and st_m["pipeline"] == nat.PIPE_MEGAKERNEL
        assert st_w["segments"] == st_m["segments"] and st_w["kernel_launches"] > 5
        np.testing.assert_array_equal(acc_w, acc_m)
        np.testing.assert_array_equal(rgba_w, rgba_m)


def test_wavefront_tiles_sample_ranges_and_golden(ctx, scenes):
    g = np.load(GOLDEN / "image_c2_sm.npz")
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    w, h, spp = int(g["width"]), int(g["height"]), int(g["spp"])
    kw = dict(samples_per_pixel=spp, max_bounces=int(g["max_bounces"]), material_mode=nat.MODE_SM, seed=int(g["seed"]), flags=nat.PIPE_WAVEFRONT)
    rgba8, accum = ctx.render(make_view(sc, w, h, **kw), want_accum=True)
    assert ctx.stats()["segments"] == int(g["segments"])
    assert_images_match(rgba8, accum, g["rgba8"], g["accum"], spp)
    # odd tile (not a multiple of the 8x4 patch) and a partial sample range
    _, part = ctx.render(make_view(sc, w, h, tile=(3, 5, 40, 31), sample_range=(2, 7), **kw), want_accum=True)
    _, ref = ctx.render(make_view(sc, w, h, tile=(3, 5, 40, 31), sample_range=(2, 7), **{**kw, "flags": nat.PIPE_MEGAKERNEL}), want_accum=True)
    np.testing.assert_allclose(part, ref, rtol=2e-6, atol=1e-7)
    assert (part[5:31, 3:40, 3] == 5).all() and (part[:5] == 0).all()
    # empty range
    _, none = ctx.render(make_view(sc, w, h, sample_range=(3, 3), **kw), want_accum=True)
    assert (none == 0).all()


def test_progressive_refinement_converges_to_the_single_render(ctx, scenes):
    from rt_b200.renderer import ProgressiveRenderer

    sc = scenes["c2"][0]
    prog = ProgressiveRenderer(ctx, sc, 160, 90, samples_per_step=4, max_bounces=50)
    first = prog.refine()
    for _ in range(3):
        img = prog.refine()
    assert prog.samples_done == 16 and (img != first).any()
    rgba8, accum = ctx.render(make_view(sc, 160, 90, samples_per_pixel=16, max_bounces=50), want_accum=True)
    np.testing.assert_allclose(prog.accum, accum, rtol=2e-6, atol=1e-7)
    assert np.abs(unpack_rgba(img) - unpack_rgba(rgba8)).max() <= 1
    prog.reset()
    assert prog.samples_done == 0 and (prog.refine() == first).all()


def test_cheaper_exact_sqrt_rcp_div_equal_the_ieee_intrinsics_for_every_input(ctx):
    """spec.cuh replaces nvcc's __fsqrt_rn/__frcp_rn/__fdiv_rn expansions by the same fast-path sequences behind fewer range
    tests; rtcu_selftest_math compares them on the device over all 2^32 float patterns (sqrt, 1/sqrt) and over every float
    in {0} u [2^-24, 2^24] divided by a set of image sides"""
    sides = [1, 2, 3, 7, 33, 97, 131, 160, 200, 256, 320, 600, 800, 1080, 1920, 2160, 3840, 4096, 8191, 65535, 2 ** 23]
    r = ctx.selftest_math(sides)
    assert r["div_pairs"] == len(sides) * (0x4B800000 - 0x33800000 + 2)
    assert (r["sqrt"], r["rcp_of_sqrt"], r["div"]) == (0, 0, 0), r


def test_tile_issue_order_never_changes_the_image(ctx, scenes, knobs):
    """The library times row-major against cost-sorted tile order over the first frames of a view and keeps the faster
    (rtcu.cu: launch_render).  Whatever it picks, and whichever phase a frame falls in, accum and pixels are bit-identical."""
    knobs(RTCU_SCAN_DIRECT="0", RTCU_BVH_DIRECT="0")  # (these frames would otherwise take the kernels that have no tile order)
    for name, spp, flags in (("c2", 16, 0), ("c3", 8, 0), ("c3", 8, nat.ACCEL_LINEAR)):
        sc, depth = scenes[name]
        ctx.upload_scene(sc)  # resets the per-view history
        v = make_view(sc, 1280, 720, samples_per_pixel=spp, max_bounces=depth, material_mode=nat.MODE_SM, flags=flags)
        knobs(RTCU_TILE_ORDER="0")
        ref_rgba8, ref_accum = ctx.render(v, want_accum=True)
        segs = ctx.stats()["segments"]
        knobs(RTCU_TILE_ORDER=None)
        launches = []
        for _ in range(5):  # row-major (timed), sorted (timed), then the winner
            rgba8, accum = ctx.render(v, want_accum=True)
            np.testing.assert_array_equal(accum, ref_accum)
            np.testing.assert_array_equal(rgba8, ref_rgba8)
            assert ctx.stats()["segments"] == segs
            launches.append(ctx.stats()["kernel_launches"])
        beams = 1 if ctx.stats()["accel"] == nat.ACCEL_BVH else 0  # a BVH frame is preceded by k_beam_lists
        assert launches[0] == 2 + beams and launches[1] == 3 + beams  # second frame: k_tile_order + megakernel + stragglers
        knobs(RTCU_TILE_ORDER="1")
        rgba8, accum = ctx.render(v, want_accum=True)
        np.testing.assert_array_equal(accum, ref_accum)
        assert ctx.stats()["kernel_launches"] == 3 + beams
        knobs(RTCU_TILE_ORDER=None)


def test_bvh_direct_mode_tiles_ranges_and_device_accumulate(ctx, oracle, scenes, knobs):
    """From 4 samples per call a BVH scene is rendered with 2 .. 16 lanes sharing each pixel's samples (k_render_stragglers in
    direct mode): same paths as the thread-per-pixel kernel (equal segment counts), per-pixel sums equal up to fp32 order, partial tiles / ragged 8x4 patches /
    sample ranges / accumulate-onto-a-device-buffer all behave like the other kernels, and the oracle agrees."""
    import torch

    sc, depth = scenes["c3"]
    ctx.upload_scene(sc)
    w, h = 203, 117  # ragged against the 8x4 patches
    for spp in (5, 9, 20, 96):  # 2 / 4 / 8 / 16 lanes per pixel (16 / 8 / 4 / 2 pixels per warp)
        kw = dict(samples_per_pixel=spp, max_bounces=depth, material_mode=nat.MODE_SM)
        v = make_view(sc, w, h, **kw)
        rgba8, accum = ctx.render(v, want_accum=True)
        segs = ctx.stats()["segments"]
        assert ctx.stats()["kernel_launches"] == 2 and (accum[..., 3] == spp).all()  # k_beam_lists + the trace kernel
        # without the patch beams every primary ray traverses: the same hits, hence the same paths (equal segment counts); which
        # lane traces which sample shifts (a lane with a list advances two segments per iteration), so the sums agree to fp32 order
        knobs(RTCU_BVH_BEAM="0")
        rgba8_nb, accum_nb = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs and ctx.stats()["kernel_launches"] == 1
        knobs(RTCU_BVH_BEAM=None)
        np.testing.assert_array_equal(accum_nb[..., 3], accum[..., 3])
        np.testing.assert_allclose(accum_nb[..., :3], accum[..., :3], rtol=4e-6, atol=1e-6)
        assert np.abs(unpack_rgba(rgba8_nb) - unpack_rgba(rgba8)).max() <= 1
        knobs(RTCU_BVH_DIRECT="0")
        rgba8_t, accum_t = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs and ctx.stats()["kernel_launches"] >= 2
        knobs(RTCU_BVH_DIRECT=None)
        np.testing.assert_allclose(accum[..., :3], accum_t[..., :3], rtol=4e-6, atol=1e-6)
        assert np.abs(unpack_rgba(rgba8) - unpack_rgba(rgba8_t)).max() <= 1
        r_rgba8, r_accum, r_segs = oracle.render(sc, v, threads=0)
        assert r_segs == segs
        np.testing.assert_allclose(accum[..., :3], r_accum[..., :3], rtol=3e-5, atol=1e-5)
        assert np.abs(unpack_rgba(rgba8) - unpack_rgba(r_rgba8)).max() <= 1
        rgba8_b, accum_b = ctx.render(v, want_accum=True)  # deterministic
        np.testing.assert_array_equal(accum_b, accum)
    # images and tiles smaller than one 8x4 patch
    for tw, th, tile in ((3, 2, None), (40, 30, (17, 11, 18, 12))):
        tkw = dict(samples_per_pixel=16, max_bounces=depth, material_mode=nat.MODE_SM)
        tv = make_view(sc, tw, th, tile=tile, **tkw) if tile else make_view(sc, tw, th, **tkw)
        _, small = ctx.render(tv, want_accum=True)
        assert ctx.stats()["kernel_launches"] == 2
        knobs(RTCU_BVH_DIRECT="0")
        _, small_t = ctx.render(tv, want_accum=True)
        knobs(RTCU_BVH_DIRECT=None)
        np.testing.assert_array_equal(small[..., 3], small_t[..., 3])
        np.testing.assert_allclose(small[..., :3], small_t[..., :3], rtol=4e-6, atol=1e-6)
        assert small[..., 3].sum() == 16 * (1 if tile else tw * th)
    # a partial tile writes only the tile, and equals the same pixels of the frame (same paths; the tile's 8x4 patches start at
    # the tile's corner, so which patches have a beam list -- and with it the order of a pixel's partial sums -- may differ)
    img = np.full((h, w), 0xDEADBEEF, np.uint32)
    tv = make_view(sc, w, h, tile=(13, 9, 150, 100), **kw)
    part, pacc = ctx.render(tv, rgba8=img, want_accum=True)
    np.testing.assert_array_equal(pacc[9:100, 13:150, 3], accum[9:100, 13:150, 3])
    np.testing.assert_allclose(pacc[9:100, 13:150, :3], accum[9:100, 13:150, :3], rtol=4e-6, atol=1e-6)
    assert (part[:9] == 0xDEADBEEF).all() and (part[:, 150:] == 0xDEADBEEF).all()
    # rtcu_render_device with accumulate: two calls of 64 samples onto one device buffer = one call of 128 (fp32 order)
    dev = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    for rng in ((0, 64), (64, 128)):
        ctx.render_device(make_view(sc, w, h, samples_per_pixel=128, max_bounces=depth, material_mode=nat.MODE_SM, sample_range=rng),
                          dev.data_ptr(), accumulate=rng[0] > 0, stream=st)
    torch.cuda.synchronize()
    _, whole = ctx.render(make_view(sc, w, h, samples_per_pixel=128, max_bounces=depth, material_mode=nat.MODE_SM), want_accum=True)
    got = dev.cpu().numpy()
    np.testing.assert_array_equal(got[..., 3], whole[..., 3])
    np.testing.assert_allclose(got[..., :3], whole[..., :3], rtol=4e-6, atol=1e-6)
