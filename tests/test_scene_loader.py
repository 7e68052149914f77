"""Scene loader: defaults, clamps, aliases and errors of scene::load (reference src/scene.cpp:483-618)."""
import numpy as np
import pytest

from rt_b200 import scene as S, synth


def test_defaults_of_an_empty_scene():
    s = S.loads("")
    assert (s.samples_per_pixel, s.max_bounces) == (30, 10)  # scene.hpp:10-11
    assert s.camera.position == (0.0, 1.0, 0.0) and s.camera.direction == (0.0, 0.0, -1.0)
    # fallback material: lambert fuchsia rough .05 refl .5 (scene.cpp:565-566)
    assert len(s.materials) == 1
    m = s.materials[0]
    assert int(m["type"]) == S.LAMBERT and tuple(m["albedo"]) == (1.0, 0.0, 1.0, 1.0)
    assert m["roughness"] == np.float32(0.05) and m["reflectivity"] == np.float32(0.5)
    assert len(s.spheres) == len(s.planes) == len(s.boxes) == 0


def test_clamps():
    assert S.loads("samples_per_pixel = 4096").samples_per_pixel == 1000  # scene.cpp:531
    assert S.loads("samples_per_pixel = 0").samples_per_pixel == 1
    assert S.loads("max_bounces = 0").max_bounces == 1
    assert S.loads("max_bounces = 5000").max_bounces == 1000


def test_basic_scene_file():
    s = S.load("scenes/basic.toml")
    assert len(s.spheres) == 3 and len(s.materials) == 3
    np.testing.assert_array_equal(s.spheres[1], [0, 0.5, 0, 0.5])  # default radius 0.5
    np.testing.assert_array_equal(s.sphere_material, [0, 1, 2])
    assert tuple(s.materials[0]["albedo"]) == (1, 1, 1, 1)  # gray_33 binarised
    assert s.materials[2]["reflectivity"] == np.float32(0.8) and s.materials[2]["roughness"] == np.float32(0.05)
    assert s.camera.position == (0.0, 1.0, 3.0) and s.camera.direction == (0.0, 0.0, -1.0)


def test_dielectric_scene_file_per_type_ior_defaults():
    s = S.load("scenes/dielectric.toml")
    assert s.samples_per_pixel == 200 and len(s.spheres) == 7
    ior = {int(m["type"]): float(m["reflectivity"]) for m in s.materials}
    assert ior[S.DIELECTRIC] == np.float32(1.52) and ior[S.AIR] == np.float32(1.000293)
    assert ior[S.VACUUM] == 1.0 and ior[S.WATER] == np.float32(1.333) and ior[S.ICE] == np.float32(1.31)
    assert ior[S.LAMBERT] == 0.5 and ior[S.METAL] == np.float32(0.8)


def test_material_type_by_name_and_int_and_roughness_default():
    s = S.loads("materials = [ {type = 2}, {type = 'metal'}, {type = 'diamond'} ]")
    assert [int(m["type"]) for m in s.materials] == [2, 1, 7]
    assert s.materials[0]["roughness"] == 0.0 and s.materials[1]["roughness"] == 0.5  # 0.0 only for 'dielectric'
    assert s.materials[2]["reflectivity"] == 0.5  # diamond has no IOR default (scene.cpp:546-556)
    with pytest.raises(S.SceneError):
        S.loads("materials = [ {type = 8} ]")
    with pytest.raises(S.SceneError):
        S.loads("materials = [ {type = 'glass'} ]")


def test_vector_syntax():
    s = S.loads("camera = { position = 'up', direction = 'left' }")
    assert s.camera.position == (0, 1, 0) and s.camera.direction == (-1, 0, 0)
    s = S.loads("spheres = [ {position = 2}, {position = [7]}, {position = [1, 2]} ]")
    np.testing.assert_array_equal(s.spheres[0], [2, 2, 2, 0.5])      # scalar broadcast
    np.testing.assert_array_equal(s.spheres[1], [7, 1, -3, 0.5])     # missing components keep the default (0,1,-3)
    np.testing.assert_array_equal(s.spheres[2], [1, 2, -3, 0.5])
    with pytest.raises(S.SceneError):
        S.loads("spheres = [ {position = [1,2,3,4]} ]")
    with pytest.raises(S.SceneError):
        S.loads("spheres = [ {position = 'sideways'} ]")


def test_colour_syntax():
    s = S.loads("materials = [ {albedo = [0.2, 0.3, 0.4]}, {albedo = [0.5]}, {albedo = [1,1,1,0.25]}, {albedo = 'teal'}, {} ]")
    np.testing.assert_array_equal(s.materials[0]["albedo"], np.float32([0.2, 0.3, 0.4, 1.0]))
    np.testing.assert_array_equal(s.materials[1]["albedo"], np.float32([0.5, 0, 0, 1.0]))  # starts from zero, not from the default
    np.testing.assert_array_equal(s.materials[2]["albedo"], np.float32([1, 1, 1, 0.25]))
    np.testing.assert_array_equal(s.materials[3]["albedo"], np.float32([0, 1, 1, 1]))
    np.testing.assert_array_equal(s.materials[4]["albedo"], np.float32([1, 0, 1, 1]))  # default fuchsia


def test_errors():
    with pytest.raises(S.SceneError, match="out-of-range"):
        S.loads("spheres = [ {material = 1} ]")
    with pytest.raises(S.SceneError, match="NaN"):
        S.loads("spheres = [ {radius = nan} ]")
    with pytest.raises(S.SceneError, match="NaN"):
        S.loads("spheres = [ {position = [inf, 0, 0]} ]")
    with pytest.raises(S.SceneError):
        S.loads("spheres = 3")
    with pytest.raises(S.SceneError):
        S.load("scenes/does_not_exist.toml")
    with pytest.raises(S.SceneError):
        S.load("")


def test_planes_are_normalised_and_boxes_parse():
    s = S.loads("planes = [ {normal = [0, 2, 0], position = [0, 3, 0]}, {} ]\nboxes = [ {}, {extents = 2} ]")
    np.testing.assert_array_equal(s.planes[0], [0, 1, 0, -3])
    np.testing.assert_array_equal(s.planes[1], [0, 1, 0, 0])
    np.testing.assert_array_equal(s.boxes[0], [0, 1, -3, 0.5, 0.5, 0.5])
    np.testing.assert_array_equal(s.boxes[1], [0, 1, -3, 2, 2, 2])


def test_plane_arithmetic_follows_the_spec_dot_and_normalize():
    """S1 / S2: dot3 = fma(z, z', fma(y, y', x * x')), normalize = v * (1 / sqrt(dot3(v, v))), d = -dot3(n, p) -- the
    arithmetic of the muu stand-in the reference's sources are compiled against, so the loader's planes are the ones the
    reference's scene::load would hand to the renderers."""
    import math
    from fractions import Fraction

    # the fused multiply-add helper rounds once: against exact rational arithmetic on random operands, and on the classic
    # double-rounding trap (a * b + c evaluated in binary64 and then rounded to binary32 gives the other neighbour)
    rng = np.random.default_rng(5)
    for a, b, c in rng.standard_normal((300, 3)).astype(np.float32) * np.float32(3):
        exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
        got = S._fma32(a, b, c)
        lo, hi = np.nextafter(got, np.float32(-np.inf)), np.nextafter(got, np.float32(np.inf))
        assert abs(Fraction(float(got)) - exact) <= min(abs(Fraction(float(lo)) - exact), abs(Fraction(float(hi)) - exact))
    a, b, c = np.float32(1 + 2.0**-12), np.float32(1 + 2.0**-12), np.float32(2.0**-60)  # exact: 1 + 2^-11 + 2^-24 + 2^-60
    assert S._fma32(a, b, c) == np.float32(1 + 2.0**-11 + 2.0**-23)                       # above the tie: rounds up ...
    assert np.float32(float(a) * float(b) + float(c)) == np.float32(1 + 2.0**-11)         # ... binary64 first: ties to even, down
    assert math.copysign(1.0, float(S._fma32(-0.5, 0.0, -0.0))) == -1.0                   # (-0) + (-0) = -0
    assert math.copysign(1.0, float(S._fma32(0.5, 0.5, -0.25))) == 1.0                    # exact cancellation = +0

    s = S.loads("planes = [ {normal = [1.6840e-01, -0.46002, 0.6859540551275798], position = [2.5, -1, 0.75]}, {normal = -0.6}, {normal = [0, 0, 0]} ]")
    n = np.float32([1.6840e-01, -0.46002, 0.6859540551275798])
    inv = np.float32(1) / np.sqrt(S._dot3(n, n), dtype=np.float32)
    unit = (n * inv).astype(np.float32)
    np.testing.assert_array_equal(s.planes[0, :3], unit)
    assert s.planes[0, 3] == -S._dot3(unit, np.float32([2.5, -1, 0.75]))
    assert s.planes[1, 3] == 0 and not np.signbit(s.planes[1, 3])  # -((-0) + (-0) + (-0)) = +0
    assert np.isnan(s.planes[2]).all()                              # a zero normal normalises to NaN (as muu's normalize would)


def test_dumps_round_trips_every_primitive_table():
    s = S.load("scenes/boxes.toml")
    assert len(s.boxes) and len(s.planes) and len(s.spheres)
    t = S.loads(S.dumps(s))
    np.testing.assert_array_equal(t.boxes, s.boxes)
    np.testing.assert_array_equal(t.box_material, s.box_material)
    np.testing.assert_array_equal(t.spheres, s.spheres)
    np.testing.assert_array_equal(t.materials, s.materials)
    np.testing.assert_array_equal(t.planes[:, :3], s.planes[:, :3])
    np.testing.assert_allclose(t.planes[:, 3], s.planes[:, 3], rtol=1e-6)
    assert (t.samples_per_pixel, t.max_bounces, t.camera.position, t.camera.direction) == (s.samples_per_pixel, s.max_bounces, s.camera.position, s.camera.direction)


def test_relative_path_search(tmp_path, monkeypatch):
    (tmp_path / "scenes").mkdir()
    (tmp_path / "scenes" / "x.toml").write_text("samples_per_pixel = 7")
    (tmp_path / "sub").mkdir()
    monkeypatch.chdir(tmp_path / "sub")
    assert S.load("x.toml").samples_per_pixel == 7  # found through ../scenes/ (scene.cpp:479-480)


def test_synthetic_scenes_round_trip_through_toml():
    sc = synth.rtiow_scene()
    assert 470 <= len(sc.spheres) <= 500 and len(sc.materials) == len(sc.spheres)
    sc.samples_per_pixel, sc.max_bounces = 256, 50
    back = S.loads(S.dumps(sc))
    np.testing.assert_array_equal(back.spheres, sc.spheres)
    np.testing.assert_array_equal(back.sphere_material, sc.sphere_material)
    np.testing.assert_array_equal(back.materials, sc.materials)
    assert back.camera == sc.camera and back.samples_per_pixel == 256 and back.max_bounces == 50
    g = synth.grid_scene(nx=20, nz=10)
    assert len(g.spheres) == 201 and int(g.sphere_material.max()) < len(g.materials)
    # determinism of the generators
    np.testing.assert_array_equal(synth.rtiow_scene().spheres, sc.spheres)


def test_scene_files_describe_the_reference_scenes():
    # the two shipped scene files restate the reference's scenes by value (they are written in a different TOML style);
    # expected values read off reference scenes/basic.toml:1-19 and scenes/dielectric.toml:1-30
    b = S.load("scenes/basic.toml")
    np.testing.assert_array_equal(b.spheres, np.float32([[0, -1000, 0, 1000], [0, 0.5, 0, 0.5], [1, 0.5, 0, 0.5]]))
    assert [int(t) for t in b.materials["type"]] == [S.LAMBERT, S.LAMBERT, S.METAL] and (b.samples_per_pixel, b.max_bounces) == (30, 10)
    d = S.load("scenes/dielectric.toml")
    np.testing.assert_array_equal(d.spheres[:, :3], np.float32([[0, -1000, 0], [-3, 0.5, 0], [-3, 2, 0], [-1, 0.5, 0], [-1, 2, 0], [1, 0.5, 0], [1, 2, 0]]))
    assert [int(t) for t in d.materials["type"]] == [S.LAMBERT, S.VACUUM, S.METAL, S.DIELECTRIC, S.AIR, S.WATER, S.ICE]
    np.testing.assert_array_equal(d.sphere_material, np.arange(7))
    assert d.samples_per_pixel == 200 and d.camera.position == (0.0, 1.0, 7.0)
