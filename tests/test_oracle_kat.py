"""Known-answer tests that pin the CPU oracle (SURVEY.md section 8c "pins the new repo must create").

The reference has no tests or golden vectors, so these are: published Philox4x32 vectors for 7 rounds -- the SPEC's stream -- and 10 (Random123
kat_vectors), hand-derived intersection cases, and pack / colour identities read off the reference source.
"""
import math

import numpy as np
import pytest

from rt_b200 import scene as S
from rt_b200.renderer import make_view

from conftest import GOLDEN


# ---- RNG: Philox4x32, Random123 kat_vectors (`philox4x32 7 ...` and `philox4x32 10 ...` lines) -----------------
# PHILOX_KAT = the SPEC's stream (7 rounds); the 10-round vectors pin the same round function a second time.
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A)),
]
PHILOX_KAT_10 = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,expect", PHILOX_KAT)
def test_philox_published_vectors(oracle, ctr, key, expect):
    out = oracle.philox(ctr, key[0] | (key[1] << 32))
    assert tuple(int(x) for x in out) == expect
    assert tuple(int(x) for x in oracle.philox_rounds(ctr, key[0] | (key[1] << 32), 7)) == expect


@pytest.mark.parametrize("ctr,key,expect", PHILOX_KAT_10)
def test_philox_published_vectors_ten_rounds(oracle, ctr, key, expect):
    assert tuple(int(x) for x in oracle.philox_rounds(ctr, key[0] | (key[1] << 32), 10)) == expect


def test_stream_is_uniform_over_the_counters_a_pixel_uses(oracle):
    """the renderer walks the counter (pixel, sample, block, retry) sequentially in `sample`; the seven-round stream must look
    uniform and uncorrelated along exactly that axis (chi-square over 64 bins per output word, lag-1 correlation, and the
    correlation between the two jitter words of a sample)"""
    n = 1 << 14
    u = np.array([[oracle.u01(int(w)) for w in oracle.philox((12345, s, 0, 0), 0x5EED)] for s in range(n)])
    for w in range(4):
        counts = np.bincount((u[:, w] * 64).astype(int), minlength=64)
        chi2 = float(((counts - n / 64) ** 2 / (n / 64)).sum())
        assert chi2 < 120.0, (w, chi2)  # 63 degrees of freedom: mean 63, 99.99th percentile ~ 113
        assert abs(np.corrcoef(u[:-1, w], u[1:, w])[0, 1]) < 0.04  # lag 1 along the sample index (sigma = 1 / sqrt(n) ~ 0.008)
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.04
    assert abs(u.mean() - 0.5) < 0.01


def test_u01_mapping(oracle):
    assert oracle.u01(0) == 0.0
    assert oracle.u01(0xFF) == 0.0  # low 8 bits dropped
    assert oracle.u01(0x100) == 2.0 ** -24
    assert oracle.u01(0xFFFFFFFF) == 1.0 - 2.0 ** -24  # never 1.0
    assert oracle.u01(0x80000000) == 0.5


# ---- intersection: hand-derived cases (mg_ray_tracer.cpp:62-87 + muu ray::hits spec S4) ----------------
def _scene(spheres, planes=(), materials=None):
    s = S.Scene()
    s.materials = S.make_materials(materials or [(S.LAMBERT, (1, 1, 1), 0.5, 0.5)])
    s.spheres = np.array(spheres, np.float32).reshape(-1, 4)
    s.sphere_material = np.zeros(len(s.spheres), np.uint32)
    s.planes = np.array(planes, np.float32).reshape(-1, 4)
    s.plane_material = np.zeros(len(s.planes), np.uint32)
    return s


def _hit(oracle, scene, o, d):
    hit, prim, t, n = oracle.intersect_batch(scene, np.array([o], np.float32), np.array([d], np.float32))
    return int(hit[0]), int(prim[0]), float(t[0]), n[0]


def test_tangent_ray_basic_scene(oracle):
    # ray (0,1,3)->(0,0,-1) vs sphere c=(0,.5,0) r=.5: e=(0,-.5,-3), a=3, e2=9.25, disc=.25-.25=0 -> t=3
    h, prim, t, n = _hit(oracle, _scene([(0, 0.5, 0, 0.5)]), (0, 1, 3), (0, 0, -1))
    assert (h, prim, t) == (1, 0, 3.0)
    np.testing.assert_array_equal(n, [0, 1, 0])


def test_centre_hit_ground_sphere(oracle):
    h, prim, t, n = _hit(oracle, _scene([(0, -1000, 0, 1000)]), (0, 5, 0), (0, -1, 0))
    assert (h, prim, t) == (1, 0, 5.0)
    np.testing.assert_array_equal(n, [0, 1, 0])


def test_origin_inside_sphere_takes_far_root(oracle):
    # e2 < r2 -> t = a + f: from the centre every direction exits at t = r
    h, prim, t, n = _hit(oracle, _scene([(0, 0, 0, 2)]), (0, 0, 0), (1, 0, 0))
    assert (h, prim, t) == (1, 0, 2.0)
    np.testing.assert_array_equal(n, [1, 0, 0])  # normal stays outward (mg_ray_tracer.cpp:85)


def test_sphere_behind_origin_rejected(oracle):
    # a = -5, f = 1 -> t = -6 < min_hit_dist
    h, prim, t, _ = _hit(oracle, _scene([(0, 0, 5, 1)]), (0, 0, 0), (0, 0, -1))
    assert (h, prim, t) == (0, 0xFFFFFFFF, -1.0)


def test_grazing_miss(oracle):
    h, prim, t, _ = _hit(oracle, _scene([(0, 0, -5, 1)]), (0, 1.001, 0), (0, 0, -1))
    assert h == 0


def test_near_root_below_min_dist_skips_sphere_entirely(oracle):
    # origin just outside the surface, pointing inward: near root ~1e-4 < 0.001 -> the single returned root is
    # rejected and the far side is never reported (one-root quirk)
    h, prim, t, _ = _hit(oracle, _scene([(0, 0, 0, 1)]), (0, 0, 1.0001), (0, 0, -1))
    assert h == 0


def test_equal_distance_tie_lowest_index_wins(oracle):
    sc = _scene([(0, 0, -5, 1), (0, 0, -5, 1), (0, 0, -9, 1)])
    h, prim, t, _ = _hit(oracle, sc, (0, 0, 0), (0, 0, -1))
    assert (h, prim, t) == (1, 0, 4.0)


def test_closest_of_many(oracle):
    sc = _scene([(0, 0, -9, 1), (0, 0, -5, 1), (0, 0, -7, 1)])
    h, prim, t, _ = _hit(oracle, sc, (0, 0, 0), (0, 0, -1))
    assert (h, prim, t) == (1, 1, 4.0)


def test_plane_one_sided_and_sphere_wins_tie(oracle):
    # plane y=0 (n=+Y, d=0): hit from above, missed from below (nd >= 0)
    sc = _scene([], planes=[(0, 1, 0, 0)])
    h, prim, t, n = _hit(oracle, sc, (0, 2, 0), (0, -1, 0))
    assert (h, prim, t) == (1, 0x80000000, 2.0)
    np.testing.assert_array_equal(n, [0, 1, 0])
    assert _hit(oracle, sc, (0, -2, 0), (0, 1, 0))[0] == 0
    # sphere touching the plane at the same distance: select(spheres, planes) keeps the sphere on a tie
    sc2 = _scene([(0, 1, 0, 1)], planes=[(0, 1, 0, -2)])  # plane y=2, sphere top at y=2
    h, prim, t, _ = _hit(oracle, sc2, (0, 5, 0), (0, -1, 0))
    assert (h, prim, t) == (1, 0, 3.0)


def test_empty_scene_is_all_miss(oracle):
    h, prim, t, _ = _hit(oracle, _scene([]), (0, 0, 0), (0, 0, -1))
    assert (h, prim, t) == (0, 0xFFFFFFFF, -1.0)


# ---- resolve / pack (mg_ray_tracer.cpp:195-200, colour.hpp:100-106) ------------------------------------
def test_pack_known_answers(oracle):
    assert oracle.pack_pixel(0, 0, 0, 1) == 0x000000FF
    assert oracle.pack_pixel(1, 1, 1, 1) == 0xFFFFFFFF
    assert oracle.pack_pixel(30, 30, 30, 30) == 0xFFFFFFFF
    # 0.5 -> sqrt -> 0.70710677 * 255.99999 = 181.0 -> 0xB5
    assert oracle.pack_pixel(0.5, 0.5, 0.5, 1) == 0xB5B5B5FF
    assert oracle.pack_pixel(4.0, 0.25, 0.0, 1) == 0xFF7F00FF  # clamp; sqrt(.25)=.5 -> 127
    assert oracle.pack_pixel(float("nan"), 0, 0, 1) & 0xFF0000FF == 0x000000FF  # NaN clamps to 0


# ---- colour semantics (colour.hpp:72-98, scene.cpp:347-356) --------------------------------------------
def test_named_colours_are_binarised():
    assert S.named_colour("gray_33") == (1.0, 1.0, 1.0, 1.0)
    assert S.named_colour("fuchsia") == (1.0, 0.0, 1.0, 1.0)
    assert S.named_colour("aquamarine") == (1.0, 1.0, 1.0, 1.0)
    assert S.named_colour("black") == (0.0, 0.0, 0.0, 1.0)
    assert S.named_colour("teal") == (0.0, 1.0, 1.0, 1.0)
    with pytest.raises(S.SceneError):
        S.named_colour("octarine")


# ---- camera / primary rays (camera.hpp:42-48, :122-137; mg_ray_tracer.cpp:189-193) ---------------------
def test_primary_ray_centre_pixel_looks_forward(oracle):
    sc = S.load("scenes/basic.toml")
    v = make_view(sc, 801, 601)  # odd size: pixel (400,300) + 0.5 is the exact image centre
    o, d = oracle.primary_rays(v, [400], [300], [0])
    np.testing.assert_allclose(d[0], [0, 0, -1], atol=2e-6)
    np.testing.assert_allclose(o[0], [0, 1, 3 - 0.01], atol=2e-5)  # origin on the near plane (0.01)


def test_primary_ray_fov_and_orientation(oracle):
    sc = S.load("scenes/basic.toml")
    v = make_view(sc, 800, 600)
    o, d = oracle.primary_rays(v, [0, 799, 400, 400], [300, 300, 0, 599], [0, 0, 0, 0])
    assert d[0][0] < 0 < d[1][0]  # left / right
    assert d[2][1] > 0 > d[3][1]  # row 0 is the top of the image
    # vertical half-angle ~ vfov/2 = pi/8 at the top row centre
    ang = math.atan2(d[2][1], -d[2][2])
    assert abs(ang - math.pi / 8) < 2e-3


def test_sample_zero_is_pixel_centre_and_others_jitter(oracle):
    sc = S.load("scenes/basic.toml")
    v = make_view(sc, 64, 48)
    o0, d0 = oracle.primary_rays(v, [10, 10, 10], [20, 20, 20], [0, 1, 2])
    assert not np.array_equal(d0[0], d0[1]) and not np.array_equal(d0[1], d0[2])
    # jitter draws are (pixel, sample, block 0) of the Philox stream
    u = oracle.philox((20 * 64 + 10, 1, 0, 0), 0x5EED)
    jx, jy = oracle.u01(u[0]), oracle.u01(u[1])
    assert 0 <= jx < 1 and 0 <= jy < 1


# ---- scatter table (mg_ray_tracer.cpp:142-152 vs sm_ray_tracer.cpp:221-236) ----------------------------
def test_scatter_tables(oracle):
    mats = [(t, (0.5, 0.6, 0.7), 0.1, 1.5) for t in range(8)]
    sc = _scene([(0, 0, 0, 1)], materials=mats)
    o, d, n = [(0, 0, 2)], [(0, 0, -1)], [(0, 0, 1)]
    for mode in (0, 1):
        for mtype in range(8):
            s, att, oo, do = oracle.scatter_batch(sc, mode, 1, [mtype], o, d, [1.0], n, [5], [3], [1])
            np.testing.assert_array_equal(att[0], np.float32([0.5, 0.6, 0.7]) * np.float32(1.5))  # albedo * reflectivity quirk
            np.testing.assert_array_equal(oo[0], [0, 0, 1])
            if mtype == S.METAL:
                continue
            dielectric = mode == 1 and S.DIELECTRIC <= mtype <= S.ICE
            length = float(np.linalg.norm(do[0].astype(np.float64)))
            if dielectric:
                # head-on hit: reflected = (0,0,1) or refracted = (0,0,-1), unit either way
                assert abs(abs(do[0][2]) - 1.0) < 1e-6 and do[0][0] == 0 and do[0][1] == 0
            else:
                # lambert: normalize(n + positive-octant unit vector) -> all components >= 0, z dominant
                assert abs(length - 1) < 1e-6 and (do[0] >= 0).all()


def test_lambert_positive_octant_quirk(oracle):
    sc = _scene([(0, 0, 0, 1)])
    n = 256
    s, att, oo, do = oracle.scatter_batch(sc, 0, 9, [0] * n, [(0, 3, 0)] * n, [(0, -1, 0)] * n, [2.0] * n, [(0, -1, 0)] * n,
                                          np.arange(n), [1] * n, [1] * n)
    # scatter = normalize(n + u), u in the positive octant: x,z >= 0 always, even for a downward normal
    assert (do[:, 0] >= 0).all() and (do[:, 2] >= 0).all() and s.all()


def test_metal_absorbs_when_scatter_goes_below_surface(oracle):
    sc = _scene([(0, 0, 0, 1)], materials=[(S.METAL, (1, 1, 1), 0.0, 0.8)])
    # mirror reflection of a grazing ray stays above; a ray arriving from below the normal's hemisphere is absorbed
    s, *_ = oracle.scatter_batch(sc, 0, 1, [0], [(0, 0, 2)], [(0, 0, -1)], [1.0], [(0, 0, 1)], [0], [0], [1])
    assert s[0] == 1
    s, *_ = oracle.scatter_batch(sc, 0, 1, [0], [(0, 0, 2)], [(0, 0, -1)], [1.0], [(0, 0, -1)], [0], [0], [1])
    assert s[0] == 0  # reflect about -n gives (0,0,1); dot with n=(0,0,-1) <= 0


def test_dielectric_total_internal_reflection(oracle):
    sc = _scene([(0, 0, 0, 1)], materials=[(S.DIELECTRIC, (1, 1, 1), 0.0, 1.5)])
    # exiting (dot(d,n) > 0) at a grazing angle: sin2_t = 1.5^2 * (1 - cos^2) > 1 -> always reflect
    d = np.float32([0.9, 0, 0.43588989])
    for pix in range(16):
        s, att, oo, do = oracle.scatter_batch(sc, 1, 3, [0], [(0, 0, 0)], [d], [1.0], [(0, 0, 1)], [pix], [0], [1])
        assert s[0] == 1 and do[0][2] < 0 and abs(do[0][0] - d[0]) < 1e-6


# ---- trace (mg_ray_tracer.cpp:154-174) -------------------------------------------------------------------
def test_trace_miss_returns_sky(oracle):
    sc = _scene([])
    v = make_view(sc, 8, 8, samples_per_pixel=1, max_bounces=5)
    rad, nseg = oracle.trace_sample(sc, v, 4, 0, 0)
    assert nseg == 1
    o, d = oracle.primary_rays(v, [4], [0], [0])
    a = np.float32(0.5) * (d[0][1] + np.float32(1))
    expect = np.float32([0.5, 0.7, 1.0]) * a + (np.float32(1) - a)
    np.testing.assert_allclose(rad, expect, rtol=1e-6)


def test_trace_bounce_budget(oracle):
    # camera inside a closed lambert sphere: the primary segment always hits.  With a budget of 1 the recursion
    # returns {} right after the scatter (mg_ray_tracer.cpp:157-158): one segment, radiance 0.  Larger budgets
    # never exceed max_bounces segments (a bounce can still escape through the one-root quirk of ray::hits).
    sc = _scene([(0, 1, 0, 50)])
    v = make_view(sc, 8, 8, samples_per_pixel=1, max_bounces=1)
    rad, nseg = oracle.trace_sample(sc, v, 3, 3, 1)
    assert nseg == 1
    np.testing.assert_array_equal(rad, [0, 0, 0])
    for depth in (3, 10):
        v = make_view(sc, 8, 8, samples_per_pixel=1, max_bounces=depth)
        for px in range(8):
            rad, nseg = oracle.trace_sample(sc, v, px, 3, 1)
            assert 1 <= nseg <= depth
            if nseg == depth:
                # either the last segment missed (sky) or the budget ran out (0)
                assert (rad >= 0).all()


# ---- the oracle still reproduces the committed golden fixtures -----------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c3", "planes"])
def test_oracle_matches_golden_rays(oracle, scenes, name):
    g = np.load(GOLDEN / f"rays_{name}.npz")
    hit, prim, t, nrm = oracle.intersect_batch(scenes[name][0], g["o"], g["d"])
    np.testing.assert_array_equal(hit, g["hit"])
    np.testing.assert_array_equal(prim, g["prim"])
    np.testing.assert_array_equal(t, g["t"])
    np.testing.assert_array_equal(nrm, g["normal"])


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "planes"])
@pytest.mark.parametrize("mode", ["mg", "sm"])
def test_oracle_matches_golden_images(oracle, scenes, name, mode):
    g = np.load(GOLDEN / f"image_{name}_{mode}.npz")
    sc = scenes[name][0]
    v = make_view(sc, int(g["width"]), int(g["height"]), samples_per_pixel=int(g["spp"]), max_bounces=int(g["max_bounces"]),
                  material_mode=int(g["mode"]), seed=int(g["seed"]))
    rgba8, accum, segs = oracle.render(sc, v, threads=2)
    assert segs == int(g["segments"])
    np.testing.assert_array_equal(accum, g["accum"])
    np.testing.assert_array_equal(rgba8, g["rgba8"])


def test_oracle_thread_count_and_tiles_do_not_change_results(oracle, scenes):
    sc = scenes["c2"][0]
    v = make_view(sc, 40, 30, samples_per_pixel=4, max_bounces=50, material_mode=1)
    a1 = oracle.render(sc, v, threads=1)
    a5 = oracle.render(sc, v, threads=5)
    np.testing.assert_array_equal(a1[1], a5[1])
    assert a1[2] == a5[2]
    vt = make_view(sc, 40, 30, samples_per_pixel=4, max_bounces=50, material_mode=1, tile=(8, 4, 24, 20))
    at = oracle.render(sc, vt, threads=3)
    np.testing.assert_array_equal(at[1][4:20, 8:24], a1[1][4:20, 8:24])
    assert (at[1][:4] == 0).all() and (at[1][:, :8] == 0).all()
