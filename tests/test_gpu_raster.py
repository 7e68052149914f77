"""cuda_rasterizer (rtcu_rasterize) against the oracle and against the reference's own rasterizer.cpp fixtures, through the C
ABI.  No RNG on this path: every comparison is bit for bit (packed pixels, primitive ids, and the float32 depth)."""
import ctypes as C

import numpy as np
import pytest

from rt_b200 import _native as nat
from rt_b200 import scene as S
from rt_b200 import synth
from rt_b200.renderer import ImageView, cuda_rasterizer, make_view, renderers

from conftest import GOLDEN
from test_raster_oracle import _raster_cases, _scene, random_raster_scene, GREY

pytestmark = pytest.mark.gpu


def _same(gpu, cpu):
    for g, c, name in zip(gpu, cpu, ("rgba8", "prim", "depth")):
        if name == "depth":
            g, c = g.view(np.uint32), c.view(np.uint32)
        np.testing.assert_array_equal(g, c, err_msg=name)


@pytest.mark.parametrize("case", _raster_cases().RASTER_CASES, ids=lambda c: c[0])
def test_reference_rasterizer_fixtures_bit_for_bit(ctx, case):
    name, w, h = case
    g = np.load(GOLDEN / f"raster_{name}.npz")  # rgba8 = the reference's rasterizer.cpp, prim / depth = the oracle
    sc = _raster_cases().RASTER_SCENES[name]
    ctx.upload_scene(sc)
    v = make_view(sc, w, h)
    v.inv_view_proj[:] = g["inv_view_proj"].tolist()
    _same(ctx.rasterize(v, want_prim=True, want_depth=True), (g["rgba8"], g["prim"], g["depth"]))


@pytest.mark.parametrize("seed", range(16))
def test_random_scenes_equal_the_oracle(ctx, oracle, seed):
    sc = random_raster_scene(seed)
    ctx.upload_scene(sc)
    v = make_view(sc, 131, 77)  # ragged against the 32x8 tile
    _same(ctx.rasterize(v, want_prim=True, want_depth=True), oracle.rasterize(sc, v))


def test_known_answer_scenes(ctx, oracle):
    cases = [
        _scene(GREY),                                                                                  # sky only
        _scene(GREY, planes=[(0, 0, 1, 5, 0)]),
        _scene(GREY, planes=[(0, 0, 1, 5, 0)], boxes=[(0, 0, -3, 0.5, 0.5, 0.5, 0)]),                   # box inherits the wall normal
        _scene(GREY, boxes=[(0, 0, -3, 0.5, 0.5, 0.5, 0)]),                                            # box keeps `up`
        _scene(GREY, spheres=[(0, 0, 4, 1, 0)], planes=[(0, 0, 1, 5, 0)]),                             # negative distance
        _scene(GREY * 2, spheres=[(0, 0, -4, 1, 0), (0, 0, -4, 1, 1)], planes=[(0, 0, 1, 9, 0), (0, 0, 1, 9, 1)],
               boxes=[(2.5, 0, -4, 0.5, 0.5, 0.5, 0), (2.5, 0, -4, 0.5, 0.5, 0.5, 1)]),               # ties
        _scene(GREY, spheres=[(0, 0, -4, 1, 0)] * 5),                                                  # odd sphere count
        _scene(GREY, boxes=[(0, 0, -3, 1, 1, 0, 0)], cam=((0, 0, 0), (0, 0, -1))),                     # zero-thickness box
    ]
    for sc in cases:
        ctx.upload_scene(sc)
        for w, h in ((33, 17), (64, 8), (1, 2)):
            v = make_view(sc, w, h)
            _same(ctx.rasterize(v, want_prim=True, want_depth=True), oracle.rasterize(sc, v))


def test_many_spheres_equal_the_oracle(ctx, oracle):
    sc = synth.rtiow_scene()  # C3: ~485 spheres, the packed pair sweep dominates
    ctx.upload_scene(sc)
    v = make_view(sc, 320, 180)
    _same(ctx.rasterize(v, want_prim=True, want_depth=True), oracle.rasterize(sc, v))


def _cloud_scene(seed: int, n: int) -> S.Scene:
    """spheres all around the camera (in front, behind, enclosing it), duplicates for ties, a wall and a box"""
    rng = np.random.default_rng(seed)
    mats = [(0, tuple(rng.uniform(0.1, 1.0, 3)), 0.0, 1.0) for _ in range(5)]
    sph = np.concatenate([rng.uniform(-8, 8, (n, 3)), rng.uniform(0.05, 0.9, (n, 1))], axis=1).astype(np.float32)
    sph[n // 2:n // 2 + 8] = sph[:8]                 # exact duplicates: the lower index must win
    sph[5] = [0.1, 1.3, 9.0, 0.8]                    # behind the camera: accepted with a negative distance
    if seed % 2:
        sph[3] = [0.2, 1.0, 5.5, 3.0]                # the camera sits inside this one
    spheres = [(*row, int(rng.integers(0, 5))) for row in sph.tolist()]
    return _scene(mats, spheres=spheres, planes=[(0, 0, 1, 7.5, 1)], boxes=[(1.0, 0.5, 1.0, 0.7, 0.7, 0.7, 2)], cam=((0.0, 1.0, 5.0), (0.05, -0.1, -1.0)))


@pytest.mark.parametrize("seed,n", [(0, 40), (1, 257), (2, 1500), (3, 6000)])
def test_bvh_traversal_equals_the_index_ordered_loop(ctx, oracle, seed, n):
    sc = _cloud_scene(seed, n)
    ctx.upload_scene(sc)
    v = make_view(sc, 200, 120)
    expect = oracle.rasterize(sc, v)
    assert (expect[2] < 0).any()  # negative distances are part of the case
    for accel in (nat.ACCEL_LINEAR, nat.ACCEL_BVH, nat.ACCEL_AUTO):
        v.flags = accel
        _same(ctx.rasterize(v, want_prim=True, want_depth=True), expect)
        assert ctx.stats()["accel"] == (nat.ACCEL_LINEAR if accel == nat.ACCEL_LINEAR else nat.ACCEL_BVH)


def test_tile_writes_only_the_tile(ctx, oracle, scenes):
    sc = S.load("scenes/boxes.toml")
    ctx.upload_scene(sc)
    v = make_view(sc, 200, 125)
    full, fprim, fdepth = ctx.rasterize(v, want_prim=True, want_depth=True)
    _same((full, fprim, fdepth), oracle.rasterize(sc, v))
    v.tile_x0, v.tile_y0, v.tile_x1, v.tile_y1 = 37, 11, 150, 99
    img = np.full((125, 200), 0xDEADBEEF, np.uint32)
    part, prim, depth = ctx.rasterize(v, rgba8=img, want_prim=True, want_depth=True)
    np.testing.assert_array_equal(part[11:99, 37:150], full[11:99, 37:150])
    np.testing.assert_array_equal(prim[11:99, 37:150], fprim[11:99, 37:150])
    np.testing.assert_array_equal(depth[11:99, 37:150], fdepth[11:99, 37:150])
    outside = np.ones_like(part, bool)
    outside[11:99, 37:150] = False
    assert (part[outside] == 0xDEADBEEF).all() and (prim[outside] == nat.PRIM_MISS).all()


def test_pinned_destination_is_written_zero_copy_and_device_destination_matches(ctx, scenes):
    import torch

    sc = S.load("scenes/boxes.toml")
    ctx.upload_scene(sc)
    v = make_view(sc, 640, 360)
    staged, _, _ = ctx.rasterize(v)
    pinned = torch.zeros((360, 640), dtype=torch.int32).pin_memory()
    out, _, _ = ctx.rasterize(v, rgba8=pinned.numpy().view(np.uint32))
    np.testing.assert_array_equal(out, staged)
    dev = torch.zeros((360, 640), dtype=torch.int32, device="cuda:0")
    ctx.rasterize_device(v, dev.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev.cpu().numpy().view(np.uint32), staged)
    st = ctx.stats()
    assert st["kernel_launches"] == 1 and st["samples"] == 640 * 360


def test_renderer_interface_and_registry(oracle):
    assert renderers.find_by_name("cuda_rasterizer").create is cuda_rasterizer and renderers.find("cuda_r").create is cuda_rasterizer
    r = cuda_rasterizer()
    sc = S.load("scenes/boxes.toml")
    img = ImageView.allocate(160, 100).clear(0x000000FF)
    r.render(sc, img)
    assert r.last_error is None
    expect, _, _ = oracle.rasterize(sc, make_view(sc, 160, 100))
    np.testing.assert_array_equal(img.data, expect)
    # a changed scene (one box moved) is re-uploaded: boxes are part of the fingerprint
    sc.boxes[0, 0] += 1.0
    r.render(sc, img)
    expect2, _, _ = oracle.rasterize(sc, make_view(sc, 160, 100))
    np.testing.assert_array_equal(img.data, expect2)
    assert (expect2 != expect).any()
    r.ctx.close()


def test_full_hd_frame_equals_the_oracle_on_sampled_rows(ctx, oracle, scenes):
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    v = make_view(sc, 1920, 1080)
    gpu = ctx.rasterize(v, want_prim=True, want_depth=True)
    cpu = oracle.rasterize(sc, v, row_step=9)
    rows = np.arange(0, 1080, 9)
    _same([a[rows] for a in gpu], [a[rows] for a in cpu])
    assert len(np.unique(gpu[0])) > 200


def test_errors(ctx):
    from rt_b200.renderer import Context

    fresh = Context(0)
    v = make_view(S.load("scenes/basic.toml"), 16, 16)
    with pytest.raises(nat.RtcuError, match="rtcu_upload_scene"):
        fresh.rasterize(v)
    fresh.close()
    sc = S.load("scenes/basic.toml")
    ctx.upload_scene(sc)
    v.tile_x1 = 17
    with pytest.raises(nat.RtcuError, match="bad tile"):
        ctx.rasterize(v)
    bad = S.load("scenes/boxes.toml")
    bad.box_material[0] = 99
    with pytest.raises(nat.RtcuError, match="box 0"):
        ctx.upload_scene(bad)
