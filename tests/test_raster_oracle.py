"""The preview renderer (reference src/renderers/rasterizer.cpp:22-88) on the CPU side: known-answer tests of the oracle's
restatement (oracle/rtref.c: rtref_rasterize, S13 ray against box) and its pin against the reference's OWN rasterizer.cpp
compiled against the muu stand-in (oracle/_ref/librt_ref.so; fixtures tests/golden/raster_*.npz where the tree is absent)."""
import numpy as np
import pytest

from rt_b200 import _native as nat
from rt_b200 import scene as S
from rt_b200.renderer import make_view

from conftest import GOLDEN

MISS, PLANE, BOX = nat.PRIM_MISS, nat.PRIM_PLANE, nat.PRIM_BOX


def _raster_cases():
    import sys
    sys.path.insert(0, str(GOLDEN.parent / "tools"))
    import gen_golden

    return gen_golden


# ---- S13: ray against box --------------------------------------------------------------------------
@pytest.mark.parametrize("o,d,box,expect", [
    ((0, 0, 5), (0, 0, -1), (0, 0, 0, 1, 1, 1), 4.0),          # front face
    ((0, 0, 0.5), (0, 0, -1), (0, 0, 0, 1, 1, 1), 1.5),        # origin inside: the exit distance
    ((0, 0, -5), (0, 0, -1), (0, 0, 0, 1, 1, 1), None),        # box behind the origin (tmax < 0)
    ((3, 0, 5), (0, 0, -1), (0, 0, 0, 1, 1, 1), None),         # parallel to the x slab and outside it (+-inf, tmin > tmax)
    ((0.5, 0.25, 5), (0, 0, -1), (0, 0, 0, 1, 1, 1), 4.0),     # parallel and inside: the infinite slabs drop out
    ((1, 0, 5), (0, 0, -1), (0, 0, 0, 1, 1, 1), None),         # on a slab plane with zero direction: {-inf, 0/0 = NaN} -> fmax drops the NaN, tmax = -inf
    ((-4, 0, 0), (1, 0, 0), (2, 0, 0, 0.5, 3, 3), 5.5),
    ((0, 0, 5), (0, 0, -1), (0, 0, 0, 1, 1, 0), 5.0),          # zero-thickness box: tmin == tmax still hits
])
def test_box_known_answers(oracle, o, d, box, expect):
    hit, t = oracle.ray_hits_box(o, d, box)
    assert hit == (expect is not None)
    if expect is not None:
        assert t == np.float32(expect)


def test_box_diagonal_matches_float32_slab_arithmetic(oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        o = rng.uniform(-4, 4, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d /= np.float32(np.sqrt((d * d).sum(dtype=np.float32)))
        box = np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(0.1, 1.5, 3)]).astype(np.float32)
        lo, hi = box[:3] - box[3:], box[:3] + box[3:]
        with np.errstate(all="ignore"):
            t1, t2 = (lo - o) / d, (hi - o) / d
        tmin, tmax = np.minimum(t1, t2).max(), np.maximum(t1, t2).min()
        expect = None if (tmax < 0 or tmin > tmax) else (tmax if tmin < 0 else tmin)
        hit, t = oracle.ray_hits_box(o, d, box)
        assert hit == (expect is not None)
        if hit:
            assert np.float32(t) == np.float32(expect)


# ---- whole-pixel known answers ----------------------------------------------------------------------
def _scene(materials, spheres=(), planes=(), boxes=(), cam=((0, 0, 0), (0, 0, -1))):
    sc = S.Scene()
    sc.materials = S.make_materials(materials)
    sc.spheres = np.array([s[:4] for s in spheres], np.float32).reshape(-1, 4)
    sc.sphere_material = np.array([s[4] for s in spheres], np.uint32)
    sc.planes = np.array([p[:4] for p in planes], np.float32).reshape(-1, 4)
    sc.plane_material = np.array([p[4] for p in planes], np.uint32)
    sc.boxes = np.array([b[:6] for b in boxes], np.float32).reshape(-1, 6)
    sc.box_material = np.array([b[6] for b in boxes], np.uint32)
    sc.camera = S.Camera(position=cam[0], direction=cam[1])
    return sc


GREY = [(0, (0.5, 0.5, 0.5), 0.0, 1.0)]


def _bytes(px):
    return [(int(px) >> s) & 0xFF for s in (24, 16, 8, 0)]


def test_empty_scene_is_the_saturated_sky(oracle):
    # rasterizer.cpp:65-66: colour{238,245,255} / colour{208,228,255} are int-constructed and saturate to white
    sc = _scene(GREY)
    img, prim, depth = oracle.rasterize(sc, make_view(sc, 33, 17))
    assert (img == 0xFFFFFFFF).all() and (prim == MISS).all()
    assert (depth > 1.0).all()  # max_dist + 1 (:35)


def test_wall_facing_the_camera_shades_to_quarter_plus_three_quarter_albedo(oracle):
    # plane z = -5 with normal +z: n.l = 1 at the centre pixel -> 0.25 + 0.5*0.75 = 0.625 -> byte 159 (no gamma on this path)
    sc = _scene(GREY, planes=[(0, 0, 1, 5, 0)])
    img, prim, depth = oracle.rasterize(sc, make_view(sc, 33, 17))
    assert (prim == PLANE).all()
    assert _bytes(img[8, 16]) == [159, 159, 159, 255]
    assert img[8, 16] >= img[0, 0]  # grazing corners are darker
    assert abs(float(depth[8, 16]) - 5.0) < 0.2  # measured from the near plane


def test_box_keeps_the_normal_of_the_last_accepted_plane(oracle):
    # rasterizer.cpp:55-58 has no box branch: a box in front of the wall is lit with the WALL's normal (bright); alone it keeps
    # `up` and, seen head-on, is lit at n.l ~ 0 -> 0.25 -> byte 63/64
    box = (0, 0, -3, 0.5, 0.5, 0.5, 0)
    with_wall = _scene(GREY, planes=[(0, 0, 1, 5, 0)], boxes=[box])
    alone = _scene(GREY, boxes=[box])
    a, pa, _ = oracle.rasterize(with_wall, make_view(with_wall, 33, 17))
    b, pb, _ = oracle.rasterize(alone, make_view(alone, 33, 17))
    assert pa[8, 16] == BOX and pb[8, 16] == BOX
    assert _bytes(a[8, 16])[0] == 159
    assert _bytes(b[8, 16])[0] in (63, 64)


def test_sphere_behind_the_camera_is_accepted_with_a_negative_distance(oracle):
    # `!hit || *hit >= dist` (:48) has no lower bound and S4 returns a - f for an origin outside the sphere
    sc = _scene(GREY, spheres=[(0, 0, 4, 1, 0)], planes=[(0, 0, 1, 5, 0)])
    _, prim, depth = oracle.rasterize(sc, make_view(sc, 33, 17))
    assert prim[8, 16] == 0 and depth[8, 16] < 0
    assert (prim == PLANE).any()  # outside the silhouette the wall is still there


def test_first_index_wins_ties_within_a_category(oracle):
    sc = _scene(GREY * 2, spheres=[(0, 0, -4, 1, 0), (0, 0, -4, 1, 1)], planes=[(0, 0, 1, 9, 0), (0, 0, 1, 9, 1)],
                boxes=[(2.5, 0, -4, 0.5, 0.5, 0.5, 0), (2.5, 0, -4, 0.5, 0.5, 0.5, 1)])
    _, prim, _ = oracle.rasterize(sc, make_view(sc, 65, 33))
    assert set(np.unique(prim).tolist()) == {0, PLANE, BOX}


def test_tile_and_row_step_write_only_their_pixels(oracle, scenes):
    sc = scenes["planes"][0]
    v = make_view(sc, 64, 40)
    full, fprim, _ = oracle.rasterize(sc, v)
    v.tile_x0, v.tile_y0, v.tile_x1, v.tile_y1 = 5, 7, 41, 29
    part, pprim, _ = oracle.rasterize(sc, v, row_step=3)
    rows = np.arange(7, 29, 3)
    np.testing.assert_array_equal(part[rows, 5:41], full[rows, 5:41])
    mask = np.zeros_like(part, bool)
    mask[rows[:, None], np.arange(5, 41)[None, :]] = True
    assert (part[~mask] == 0).all() and (pprim[~mask] == MISS).all()


# ---- the pin: the reference's own rasterizer.cpp ----------------------------------------------------
@pytest.mark.parametrize("case", _raster_cases().RASTER_CASES, ids=lambda c: c[0])
def test_oracle_reproduces_reference_rasterizer_fixtures_bit_for_bit(oracle, case):
    name, w, h = case
    g = np.load(GOLDEN / f"raster_{name}.npz")
    sc = _raster_cases().RASTER_SCENES[name]
    v = make_view(sc, w, h)
    v.inv_view_proj[:] = g["inv_view_proj"].tolist()
    rgba8, prim, depth = oracle.rasterize(sc, v, threads=2)
    np.testing.assert_array_equal(rgba8, g["rgba8"])
    np.testing.assert_array_equal(prim, g["prim"])
    np.testing.assert_array_equal(depth.view(np.uint32), g["depth"].view(np.uint32))


def random_raster_scene(seed: int) -> S.Scene:
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(0, 14))
    mats = [(int(rng.integers(0, 8)), tuple(rng.uniform(0.0, 1.0, 3)), 0.0, 1.0) for _ in range(int(rng.integers(1, 6)))]
    sc = S.Scene()
    sc.materials = S.make_materials(mats)
    sc.spheres = np.concatenate([rng.uniform(-3, 3, (n, 3)), rng.uniform(0.2, 1.2, (n, 1))], axis=1).astype(np.float32)
    sc.sphere_material = rng.integers(0, len(mats), n).astype(np.uint32)
    if seed % 2:
        sc.planes = np.array([[0, 1, 0, 1.0], [0, 0, 1, 6.0], [0.6, 0.8, 0, 4.0]], np.float32)[: 1 + seed % 3]
        sc.plane_material = rng.integers(0, len(mats), len(sc.planes)).astype(np.uint32)
    nb = int(rng.integers(0, 6))
    sc.boxes = np.concatenate([rng.uniform(-3, 3, (nb, 3)), rng.uniform(0.05, 1.0, (nb, 3))], axis=1).astype(np.float32)
    sc.box_material = rng.integers(0, len(mats), nb).astype(np.uint32)
    sc.camera = S.Camera(position=(float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 2)), float(rng.uniform(2, 7))),
                         direction=(float(rng.uniform(-0.2, 0.2)), float(rng.uniform(-0.3, 0.1)), -1.0))
    return sc


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_reference_rasterizer_on_random_scenes(oracle, seed):
    from oracle.binding import ReferenceBuild

    if not ReferenceBuild.available():
        pytest.skip("oracle/_ref/librt_ref.so not present and no reference tree to build it from")
    ref = ReferenceBuild()
    sc = random_raster_scene(seed)
    w, h = 97, 61
    cpu, ivp = ref.render(sc, w, h, 1, 1, 0, "rasterizer", threads=2)
    v = make_view(sc, w, h)
    v.inv_view_proj[:] = ivp.tolist()
    rgba8, _, _ = oracle.rasterize(sc, v, threads=2)
    np.testing.assert_array_equal(rgba8, cpu)
