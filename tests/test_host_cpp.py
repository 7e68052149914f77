"""The C++ host (rt_b200/host: TOML reader, scene loader restating scene.cpp, viewport, rt_headless CLI) against its Python
twin (rt_b200/scene.py, camera.py) and, on a GPU, against the Python path through the same C ABI."""
import json
import pathlib
import subprocess
import sys

import numpy as np
import pytest

from rt_b200 import build, scene as S, synth
from rt_b200.camera import inverse_view_projection


@pytest.fixture(scope="module")
def cli():
    build.build_cuda()
    return str(build.build_host())


def run(cli, *args, check=True):
    r = subprocess.run([cli, *args], capture_output=True, text=True)
    if check:
        assert r.returncode == 0, r.stderr
    return r


def dumped(cli, path) -> dict:
    return json.loads(run(cli, "--scene", str(path), "--dump-scene").stdout)


def assert_same_scene(d: dict, s: S.Scene):
    assert d["samples_per_pixel"] == s.samples_per_pixel and d["max_bounces"] == s.max_bounces
    np.testing.assert_array_equal(np.float32(d["camera"]["position"]), np.float32(s.camera.position))
    np.testing.assert_array_equal(np.float32(d["camera"]["direction"]), np.float32(s.camera.direction))
    assert len(d["materials"]) == len(s.materials)
    for a, b in zip(d["materials"], s.materials):
        assert a["type"] == int(b["type"])
        np.testing.assert_array_equal(np.float32(a["albedo"]), b["albedo"])
        assert np.float32(a["roughness"]) == b["roughness"] and np.float32(a["reflectivity"]) == b["reflectivity"]
    np.testing.assert_array_equal(np.float32(d["spheres"]).reshape(-1, 4), s.spheres)
    np.testing.assert_array_equal(np.uint32(d["sphere_material"]), s.sphere_material)
    np.testing.assert_array_equal(np.float32(d["planes"]).reshape(-1, 4), s.planes)
    np.testing.assert_array_equal(np.uint32(d["plane_material"]), s.plane_material)
    np.testing.assert_array_equal(np.float32(d["boxes"]).reshape(-1, 6), s.boxes)


def test_list_and_renderer_lookup(cli):
    assert run(cli, "--list").stdout.split() == ["cuda_path_tracer", "cuda_rasterizer"]
    r = run(cli, "--scene", "scenes/basic.toml", "--renderer", "nope", check=False)
    assert r.returncode == 1 and "error: unknown renderer 'nope'" in r.stderr
    assert run(cli, "--scene", "scenes/basic.toml", "--renderer", "cuda", "--dump-view").returncode == 0  # prefix match (main.cpp:68-81)


@pytest.mark.parametrize("name", ["basic.toml", "dielectric.toml", "boxes.toml"])
def test_shipped_scenes_load_like_the_python_loader(cli, name):
    assert_same_scene(dumped(cli, f"scenes/{name}"), S.load(f"scenes/{name}"))


def test_synthetic_scene_toml_round_trip(cli, tmp_path):
    sc = synth.rtiow_scene()
    p = tmp_path / "c3.toml"
    p.write_text(S.dumps(sc))
    assert_same_scene(dumped(cli, p), S.loads(p.read_text()))
    g = synth.grid_scene(nx=60, nz=40)
    p2 = tmp_path / "grid.toml"
    p2.write_text(S.dumps(g))
    d = dumped(cli, p2)
    assert len(d["spheres"]) == 2401
    assert_same_scene(d, S.loads(p2.read_text()))


SYNTAX = '''
# every value syntax the loader accepts (scene.cpp:113-166, :184-356, :381-404)
samples_per_pixel = 4_096        # clamps to 1000
max_bounces = 0                  # clamps to 1
camera.position = 'up'
camera.direction = [0.25, -1]    # missing z keeps the default (-1)

[[materials]]
type = 2
albedo = [0.5]                   # starts from zero, alpha 1
[[materials]]
type = 'metal'
albedo = "teal"
roughness = 1e-1
[[materials]]
name = "default-everything"

[[spheres]]
position = 2                     # scalar broadcast
material = 1
[[spheres]]
radius = 0x2

[[planes]]
normal = [0, 2.0, 0]
position = [0, 3, 0]

[[boxes]]
extents = 2
'''


def test_value_syntax_matches_the_python_loader(cli, tmp_path):
    p = tmp_path / "syntax.toml"
    p.write_text(SYNTAX)
    d = dumped(cli, p)
    assert_same_scene(d, S.loads(SYNTAX))
    assert d["samples_per_pixel"] == 1000 and d["max_bounces"] == 1
    assert d["camera"]["direction"] == [0.25, -1, -1]


@pytest.mark.parametrize("text,message", [
    ("spheres = [ {material = 3} ]", "material index 3 out-of-range"),
    ("spheres = [ {radius = nan} ]", "Infinities and NaNs are not allowed."),
    ("spheres = [ {position = 'sideways'} ]", "unknown vector alias 'sideways'"),
    ("materials = [ {albedo = 'octarine'} ]", "unknown colour alias 'octarine'"),
    ("materials = [ {type = 'glass'} ]", "was not a member of enum material_type"),
    ("materials = [ {type = 8} ]", "integer value 8 was not a member of enum material_type"),
    ("spheres = 3", "expected array at key 'spheres'"),
    ("spheres = [ {position = [1,2,3,4]} ]", "No mapping from TOML array"),
    ("spheres = [ {radius = 1 ]", "TOML parse error"),
])
def test_loader_errors_exit_1_with_the_reference_messages(cli, tmp_path, text, message):
    p = tmp_path / "bad.toml"
    p.write_text(text)
    r = run(cli, "--scene", str(p), "--dump-scene", check=False)
    assert r.returncode == 1 and r.stderr.strip().splitlines()[-1].startswith("error: ") and message in r.stderr
    with pytest.raises(S.SceneError):
        S.loads(text)


NUMBERS_OK = ["0", "+0", "-0", "7", "+7", "-17", "1_000", "1_2_3", "0x1F", "0xdead_beef", "0o17", "0b1_01", "1.5", "-0.25", "+1.0", "1e3", "1E-02", "1.2e-01",
              "6.02e+23", "1_0.2_5e1_0", "9223372036854775807", "-9223372036854775808", "0e0", "0.0", "-0.0"]
NUMBERS_BAD = ["007", "+007", "-01", "1.", ".5", "1._5", "1_.5", "_1", "1_", "1__0", "0x", "0x_1", "0xG", "+0x1", "-0b1", "0o8", "1e", "1e+", "1.5e", "1e1.5",
               "0x1.8p3", "infinity", "in_f", "NaN", "Inf", "1-2", "1979-05-27", "07:32:00", "--1", "+-1", "1e_5"]


NUMBERS_OUT_OF_RANGE = ["9223372036854775808", "-9223372036854775809", "0xFFFFFFFFFFFFFFFF"]  # TOML integers are int64 (tomllib keeps big ints)


def test_number_grammar_is_tomls(cli, tmp_path):
    """toml_lite accepts exactly TOML's integers and floats (found by tests/tools/fuzz_loader.py: `1.2e-01` was taken for a date;
    strtod / strtoll alone would also take `1.`, `.5`, `007`, hex floats); tomllib is the yardstick."""
    import tomllib

    for tok in NUMBERS_OK + NUMBERS_BAD:
        text = f"spheres = [ {{ position = {tok} }} ]\n"  # scalar broadcast: a plain cast of any TOML number (scene.cpp:146-157)
        try:
            v = tomllib.loads(text)["spheres"][0]["position"]
            expect_ok = isinstance(v, (int, float)) and not isinstance(v, bool)
        except tomllib.TOMLDecodeError:
            expect_ok = False
        assert expect_ok == (tok in NUMBERS_OK), tok
        p = tmp_path / "n.toml"
        p.write_text(text)
        r = run(cli, "--scene", str(p), "--dump-scene", check=False)
        if expect_ok:
            assert r.returncode == 0, (tok, r.stderr)
            assert np.float32(json.loads(r.stdout)["spheres"][0][0]) == np.float32(v), tok
        else:
            assert r.returncode == 1 and "error: " in r.stderr, (tok, r.stdout, r.stderr)
    for tok in NUMBERS_OUT_OF_RANGE:
        p = tmp_path / "n.toml"
        p.write_text(f"spheres = [ {{ position = {tok} }} ]\n")
        r = run(cli, "--scene", str(p), "--dump-scene", check=False)
        assert r.returncode == 1 and "out of range" in r.stderr, (tok, r.stderr)


def test_string_forms_decode_like_tomllib(cli, tmp_path):
    """multi-line basic / literal strings and every escape: the decoded text selects aliases, colours and material types, so
    the loaded values tell whether both readers saw the same string"""
    text = (pathlib.Path(__file__).parent / "golden" / "strings.toml").read_text()
    p = tmp_path / "strings.toml"
    p.write_text(text)
    s = S.loads(text)
    assert tuple(s.camera.position) == (0, 1, 0) and tuple(s.camera.direction) == (0, -1, 0) and int(s.materials[1]["type"]) == 1
    assert_same_scene(dumped(cli, p), s)
    for bad, message in [('name = "\\q"', "escape"), ('name = "\\u12"', "unicode escape"), ('name = "\\uD800"', "scalar value"),
                         ('name = "a\x01b"', "control character"), ('name = """never closed', "unterminated"), ('"""k""" = 1', "key")]:
        p.write_text(bad + "\n")
        r = run(cli, "--scene", str(p), "--dump-scene", check=False)
        assert r.returncode == 1 and "TOML parse error" in r.stderr and message in r.stderr, (bad, r.stderr)
        with pytest.raises(S.SceneError):
            S.loads(bad + "\n")


@pytest.mark.parametrize("text,expect", [
    # toml++'s permissive node.value<T>() as recalled (UNVERIFIED, see scene_loader.hpp): what converts ...
    ("samples_per_pixel = 4.0", {"samples_per_pixel": 4}), ("samples_per_pixel = true", {"samples_per_pixel": 1}), ("max_bounces = 7e0", {"max_bounces": 7}),
    ("spheres = [ {radius = 16777216} ]", {"radius": 16777216.0}), ("spheres = [ {radius = -3} ]", {"radius": -3.0}),
    ("spheres = [ {radius = 3.4028234e38} ]", {"radius": 3.4028234663852886e38}), ("spheres = [ {material = 0.0} ]", {"material": 0}),
    # ... and what has no mapping: out-of-range integers do not wrap, fractions do not truncate, floats do not overflow to inf
    ("samples_per_pixel = -3", "No mapping from TOML integer to unsigned"), ("samples_per_pixel = 4294967296", "No mapping from TOML integer to unsigned"),
    ("samples_per_pixel = 4.5", "to unsigned"), ("max_bounces = inf", "to unsigned"), ("spheres = [ {material = -1} ]", "No mapping from TOML integer to unsigned"),
    ("spheres = [ {radius = 16777217} ]", "No mapping from TOML integer to float"), ("spheres = [ {radius = 1e39} ]", "to float"),
    ("spheres = [ {radius = true} ]", "No mapping from TOML boolean to float"), ("spheres = [ {radius = inf} ]", "Infinities and NaNs are not allowed."),
    ("spheres = [ {position = [1, 1e39]} ]", "to float"),
])
def test_numeric_conversions_follow_tomlplusplus_value(cli, tmp_path, text, expect):
    p = tmp_path / "v.toml"
    p.write_text(text + "\n")
    r = run(cli, "--scene", str(p), "--dump-scene", check=False)
    if isinstance(expect, str):
        assert r.returncode == 1 and expect in r.stderr, r.stderr
        with pytest.raises(S.SceneError) as e:
            S.loads(text)
        assert expect in str(e.value)
        return
    assert r.returncode == 0, r.stderr
    d, s = json.loads(r.stdout), S.loads(text)
    assert_same_scene(d, s)
    for key, want in expect.items():
        got = {"radius": lambda: d["spheres"][0][3], "material": lambda: d["sphere_material"][0]}.get(key, lambda: d[key])()
        assert np.float32(got) == np.float32(want), (key, got)


def test_non_table_elements_read_as_defaults(cli, tmp_path):
    text = "materials = [1, {type = 'metal'}]\nspheres = [3, 'x', {radius = 2}]\nplanes = [[1, 2]]\nboxes = [true]\n"
    p = tmp_path / "d.toml"
    p.write_text(text)
    d, s = dumped(cli, p), S.loads(text)
    assert_same_scene(d, s)
    assert len(s.materials) == 2 and int(s.materials[0]["type"]) == 0 and len(s.spheres) == 3 and len(s.planes) == 1 and len(s.boxes) == 1
    np.testing.assert_array_equal(s.spheres[0], [0, 1, -3, 0.5])
    np.testing.assert_array_equal(s.planes[0], [0, 1, 0, 0])


def test_stdin_and_first_available_scene(cli, tmp_path, monkeypatch):
    """scene::load("-") reads standard input (scene.cpp:490-493); without --scene the app loads the first *.toml of the first
    search directory that has one (scene.cpp:620-643, main.cpp:121-125)"""
    text = "samples_per_pixel = 7\nspheres = [ {radius = 2} ]\n"
    r = subprocess.run([cli, "--scene", "-", "--dump-scene"], input=text, capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["samples_per_pixel"] == 7
    (tmp_path / "work").mkdir()
    (tmp_path / "scenes").mkdir()
    (tmp_path / "scenes" / "notes.txt").write_text("not a scene")
    (tmp_path / "scenes" / "only.toml").write_text("samples_per_pixel = 11\n")
    (tmp_path / "work" / "here.toml").write_text("samples_per_pixel = 13\n")  # the working directory itself is not searched
    r = subprocess.run([cli, "--dump-scene"], capture_output=True, text=True, cwd=tmp_path / "work")
    assert r.returncode == 0 and json.loads(r.stdout)["samples_per_pixel"] == 11, r.stderr
    monkeypatch.chdir(tmp_path / "work")
    assert S.load_first_available().samples_per_pixel == 11
    (tmp_path / "scenes" / "only.toml").unlink()
    r = subprocess.run([cli, "--dump-scene"], capture_output=True, text=True, cwd=tmp_path / "work")
    assert r.returncode == 1 and "error: no scene files found" in r.stderr
    with pytest.raises(S.SceneError, match="no scene files found"):
        S.load_first_available()
    d = tmp_path / "work" / "dir.toml"
    d.mkdir()  # a directory is not a scene file
    r = subprocess.run([cli, "--scene", "dir.toml", "--dump-scene"], capture_output=True, text=True, cwd=tmp_path / "work")
    assert r.returncode == 1 and "did not exist or was not a file" in r.stderr
    with pytest.raises(S.SceneError, match="did not exist or was not a file"):
        S.load("dir.toml")


@pytest.mark.parametrize("args,message", [
    (["--spp", "abc"], "bad --spp 'abc'"), (["--spp", "0"], "bad --spp '0'"), (["--spp", "-5"], "bad --spp '-5'"), (["--bounces", "3x"], "bad --bounces '3x'"),
    (["--gpus", "9"], "bad --gpus '9' (expected a whole number in [1, 8])"), (["--device", "x"], "bad --device 'x'"), (["--seed", "zz"], "bad --seed 'zz'"),
    (["--size", "0x0"], "bad --size '0x0' (expected WxH)"), (["--size"], "option '--size' needs a value"), (["--bogus"], "unknown option '--bogus'"),
    (["--mode", "xx"], "bad --mode 'xx'"),
])
def test_cli_argument_errors_name_the_option(cli, args, message):
    r = run(cli, "--scene", "scenes/basic.toml", *args, check=False)
    assert r.returncode == 1 and r.stderr.strip().splitlines()[-1].startswith("error: ") and message in r.stderr, r.stderr


def test_documents_must_be_utf8(cli, tmp_path):
    p = tmp_path / "u.toml"
    p.write_bytes('[[materials]]\nname = "caf\u00e9 \U0001F600"  # \u00e9 in a comment\n'.encode())
    assert len(dumped(cli, p)["materials"]) == 1
    for bad in (b'a = 1 # \x80 stray continuation\n', b'a = "\xc0\xaf"\n', b'a = "\xed\xa0\x80"\n', b'a = "\xf4\x90\x80\x80"\n', b'a = "\xe2\x82"\n'):
        p.write_bytes(bad)
        r = run(cli, "--scene", str(p), "--dump-scene", check=False)
        assert r.returncode == 1 and "not valid UTF-8" in r.stderr, (bad, r.stderr)
        with pytest.raises(UnicodeDecodeError):
            bad.decode("utf-8")


def test_deeply_nested_values_are_a_parse_error_not_a_crash(cli, tmp_path):
    """arrays / inline tables nested 200 000 deep used to overflow the C++ parser's stack (exit 139); toml++ caps nesting at 256
    and reports a parse error, which main.cpp:370-379 prints as `error: ...` with exit code 1"""
    from rt_b200 import scene as S

    p = tmp_path / "deep.toml"
    for doc in ("x = " + "[" * 200000 + "]" * 200000 + "\n", "x = " + "{a = " * 100000 + "1" + "}" * 100000 + "\n", "x = " + "[" * 300 + "]" * 300 + "\n"):
        p.write_text(doc)
        r = run(cli, "--scene", str(p), "--dump-scene", check=False)
        assert r.returncode == 1 and "nested" in r.stderr and r.stderr.strip().splitlines()[-1].startswith("error: "), (r.returncode, r.stderr[-200:])
    with pytest.raises(S.SceneError, match="nested"):
        S.loads("x = " + "[" * 200000 + "]" * 200000)
    p.write_text("x = " + "[" * 200 + "]" * 200 + "\n[[spheres]]\nradius = 2\n")  # within the limit: an ordinary document
    assert dumped(cli, p)["spheres"][0][3] == 2.0 and len(S.loads(p.read_text()).spheres) == 1


def test_dump_keeps_signed_zeros_and_non_finite_values(cli, tmp_path):
    p = tmp_path / "z.toml"
    p.write_text("planes = [ {normal = -0.6}, {normal = [0, 0, 0]}, {position = [0, 2, 0]} ]\nspheres = [ {position = [-0.0, 0, 1e-46]} ]\n")
    d = dumped(cli, p)
    s = S.load(p)
    for key in ("planes", "spheres"):
        a, b = np.float32(d[key]).ravel(), getattr(s, key).ravel()
        assert ((np.isnan(a) & np.isnan(b)) | (a.view(np.uint32) == b.view(np.uint32))).all(), key
    assert np.signbit(np.float32(d["spheres"][0][0])) and np.isnan(np.float32(d["planes"][1])).all()


def test_viewport_matrix_matches_the_python_camera(cli, tmp_path):
    """both hosts build inverse(P * view) as world * inverse(P) with the same float64 sums: bit-identical matrices, exact
    structural zeros (the w row is (0, 0, m11, m15), which lets the kernels take the cheaper perspective divide)"""
    def check(path, size):
        m = np.float32(json.loads(run(cli, "--scene", str(path), "--size", f"{size[0]}x{size[1]}", "--dump-view").stdout))
        ref = inverse_view_projection(S.load(path).camera, *size)
        assert np.array_equal(m.view(np.uint32), ref.view(np.uint32)), (path, m, ref)
        assert m[3] == 0 and m[7] == 0 and np.isfinite(m).all()

    for name, size in (("scenes/basic.toml", (800, 600)), ("scenes/dielectric.toml", (1920, 1080)), ("scenes/boxes.toml", (320, 200))):
        check(name, size)
    rng = np.random.default_rng(3)
    directions = [(0, 1, 0), (0, -1, 0), (0, -1, 1e-5), (1e-3, 5, 2e-3), (3, 0, 0), (0, -0.35, -1), (-13, -2, -3)] + [tuple(rng.normal(size=3)) for _ in range(12)]
    for i, d in enumerate(directions):
        p = tmp_path / f"cam{i}.toml"
        pos = rng.uniform(-20, 20, 3)
        p.write_text(f"camera = {{ position = [{float(pos[0])!r}, {float(pos[1])!r}, {float(pos[2])!r}], direction = [{float(d[0])!r}, {float(d[1])!r}, {float(d[2])!r}] }}\n")
        check(p, [(640, 360), (800, 600), (97, 131)][i % 3])


def test_no_device_fails_loudly(cli):
    from rt_b200 import _native as nat

    if nat.load_library().rtcu_device_count() > 0:
        pytest.skip("device present")
    r = run(cli, "--scene", "scenes/basic.toml", "--size", "32x24", check=False)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_render_equals_the_python_path_byte_for_byte(cli, ctx, tmp_path):
    from rt_b200.renderer import make_view

    out = tmp_path / "img.ppm"
    r = run(cli, "--scene", "scenes/dielectric.toml", "--size", "320x180", "--spp", "8", "--bounces", "50", "--out", str(out))
    assert "Msamples/s" in r.stderr
    data = out.read_bytes()
    header = b"P6\n320 180\n255\n"
    assert data.startswith(header)
    img = np.frombuffer(data[len(header):], np.uint8).reshape(180, 320, 3)
    sc = S.load("scenes/dielectric.toml")
    v = make_view(sc, 320, 180, samples_per_pixel=8, max_bounces=50)
    v.inv_view_proj[:] = json.loads(run(cli, "--scene", "scenes/dielectric.toml", "--size", "320x180", "--dump-view").stdout)
    ctx.upload_scene(sc)
    rgba8, _ = ctx.render(v)
    ref = np.stack([(rgba8 >> 24) & 255, (rgba8 >> 16) & 255, (rgba8 >> 8) & 255], axis=-1).astype(np.uint8)
    np.testing.assert_array_equal(img, ref)
    # two GPUs when present: the sample-range split changes only the summation order
    if nat_device_count() >= 2:
        out2 = tmp_path / "img2.ppm"
        run(cli, "--scene", "scenes/dielectric.toml", "--size", "320x180", "--spp", "8", "--bounces", "50", "--gpus", "2", "--out", str(out2))
        img2 = np.frombuffer(out2.read_bytes()[len(header):], np.uint8).reshape(180, 320, 3)
        assert np.abs(img2.astype(int) - img.astype(int)).max() <= 1


@pytest.mark.gpu
def test_cli_rasterizer_equals_the_oracle(cli, oracle, tmp_path):
    from rt_b200.renderer import make_view

    out = tmp_path / "preview.ppm"
    r = run(cli, "--scene", "scenes/boxes.toml", "--renderer", "cuda_rasterizer", "--size", "320x200", "--out", str(out))
    assert "Mpixels/s" in r.stderr
    header = b"P6\n320 200\n255\n"
    img = np.frombuffer(out.read_bytes()[len(header):], np.uint8).reshape(200, 320, 3)
    sc = S.load("scenes/boxes.toml")
    v = make_view(sc, 320, 200)
    v.inv_view_proj[:] = json.loads(run(cli, "--scene", "scenes/boxes.toml", "--size", "320x200", "--dump-view").stdout)
    rgba8, _, _ = oracle.rasterize(sc, v)
    ref = np.stack([(rgba8 >> 24) & 255, (rgba8 >> 16) & 255, (rgba8 >> 8) & 255], axis=-1).astype(np.uint8)
    np.testing.assert_array_equal(img, ref)


def nat_device_count() -> int:
    from rt_b200 import _native as nat

    return nat.load_library().rtcu_device_count()


def test_loader_twins_agree_on_generated_and_mutated_files(cli):
    """a short run of the differential fuzz (tests/tools/fuzz_loader.py; the campaigns of record are in profiles/): the Python and
    the C++ loader accept and reject the same files and load the same values, and the C++ host never crashes"""
    tool = pathlib.Path(__file__).parent / "tools" / "fuzz_loader.py"
    r = subprocess.run([sys.executable, "-W", "ignore", str(tool), "--cases", "240", "--seed", "314"], capture_output=True, text=True)
    stats = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert r.returncode == 0, r.stdout[-2000:]
    assert stats["cases"] == 240 and stats["crashes"] == stats["value_mismatch"] == stats["accept_mismatch_generated"] == 0
    assert stats["both_accept"] >= 60 and stats["both_reject"] >= 60  # the generator exercises both outcomes
