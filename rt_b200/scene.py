"""Scene containers and the TOML scene loader (host side, feeds the C ABI).

Restates the reference loader `scene::load` (reference src/scene.cpp:483-618) and the containers in
src/scene.hpp:8-25 / src/soa.toml with the same defaults, clamps, aliases and error behaviour, so
that `rt --scene <toml>` and this harness flatten a scene file to the same columns.  Quirks kept
on purpose:

* named colours are binarised: `colour(uint32)` clamps the integer byte to [0,1] without /255
  (src/colour.hpp:72-98), so any non-zero byte becomes 1.0;
* a colour given as an array starts from zero (not from the default) and gets alpha 1 when it has
  fewer than four components (src/scene.cpp:347-356);
* samples_per_pixel and max_bounces are clamped to [1,1000] (src/scene.cpp:531-532);
* dielectric-class materials store their IOR in `reflectivity` (src/scene.cpp:546-556).
"""
from __future__ import annotations

import dataclasses
import fractions
import math
import os
import pathlib
import sys
import tomllib
from typing import Any, Sequence

import numpy as np

from .colour_table import NAMED_COLOURS

# rt::material_type, src/common.hpp:105-115
MATERIAL_TYPES = ("lambert", "metal", "dielectric", "air", "vacuum", "water", "ice", "diamond")
LAMBERT, METAL, DIELECTRIC, AIR, VACUUM, WATER, ICE, DIAMOND = range(8)

# src/scene.cpp:546-556
_DEFAULT_REFLECTIVITY = {
    METAL: 0.8, DIELECTRIC: 1.52, AIR: 1.000293, VACUUM: 1.0, ICE: 1.31, WATER: 1.333,
}

# muu vector constants reachable through the loader's aliases (src/scene.cpp:113-144);
# right-handed, forward = -Z (UNVERIFIED against muu, consistent with scenes/basic.toml)
_VECTOR_ALIASES = {
    "origin": (0, 0, 0), "zero": (0, 0, 0), "one": (1, 1, 1),
    "forward": (0, 0, -1), "back": (0, 0, 1), "backward": (0, 0, 1),
    "up": (0, 1, 0), "down": (0, -1, 0), "left": (-1, 0, 0), "right": (1, 0, 0),
    "x": (1, 0, 0), "x_axis": (1, 0, 0), "y": (0, 1, 0), "y_axis": (0, 1, 0),
    "z": (0, 0, 1), "z_axis": (0, 0, 1),
}

MATERIAL_DTYPE = np.dtype([("type", "<u4"), ("albedo", "<f4", (4,)), ("roughness", "<f4"), ("reflectivity", "<f4")])
assert MATERIAL_DTYPE.itemsize == 28


class SceneError(RuntimeError):
    """Mirrors the std::runtime_error thrown by the reference loader."""


def named_colour(name: str) -> tuple[float, float, float, float]:
    """colours::<name> as the reference constructs it: `colour{uint32}` -> to_component_value(int)
    clamps the *byte value* to [0,1] (src/colour.hpp:72-98), i.e. byte != 0 -> 1.0."""
    try:
        rgb = NAMED_COLOURS[name]
    except KeyError:
        raise SceneError(f"unknown colour alias '{name}'") from None
    comps = ((rgb >> 16) & 0xFF, (rgb >> 8) & 0xFF, rgb & 0xFF, 0xFF)
    return tuple(min(max(float(c), 0.0), 1.0) for c in comps)  # type: ignore[return-value]


_FLT_MAX = float(np.finfo(np.float32).max)


def _toml_type(node: Any) -> str:
    """the node type as toml++ prints it (mismatch_error, src/scene.cpp:68-87)"""
    for t, name in ((bool, "boolean"), (int, "integer"), (float, "floating-point"), (str, "string"), (list, "array"), (dict, "table")):
        if isinstance(node, t):
            return name
    return type(node).__name__


def _finite_float(node: Any, what: str) -> float:
    """`node.value<float>()` + the infinity / NaN check of src/scene.cpp:89-102.  toml++ (un-vendored dependency,
    subprojects/tomlplusplus.wrap @ f1a38d23) converts permissively; its rules as recalled, UNVERIFIED like the muu arithmetic:
    an integer converts when it lies in [-2^24, 2^24], a finite float when it lies inside float's range, inf / nan pass through
    (and are then refused by the reference's own check); booleans, strings ... have no mapping."""
    no_mapping = SceneError(f"No mapping from TOML {_toml_type(node)} to float ({what})")
    if isinstance(node, bool) or not isinstance(node, (int, float)):
        raise no_mapping
    if isinstance(node, int):
        if not -(1 << 24) <= node <= (1 << 24):
            raise no_mapping
        return float(node)
    if math.isnan(node) or math.isinf(node):
        raise SceneError("Infinities and NaNs are not allowed.")
    if not -_FLT_MAX <= node <= _FLT_MAX:
        raise no_mapping
    return float(np.float32(node))


def _vector(node: Any, default: Sequence[float], what: str) -> tuple[float, ...]:
    """src/scene.cpp:113-166"""
    n = len(default)
    if node is None:
        return tuple(float(x) for x in default)
    if isinstance(node, str):
        if node not in _VECTOR_ALIASES:
            raise SceneError(f"unknown vector alias '{node}'")
        return tuple(float(x) for x in _VECTOR_ALIASES[node][:n])
    if isinstance(node, (int, float)) and not isinstance(node, bool):
        return tuple(float(np.float32(node)) for _ in range(n))  # scalar broadcast (no NaN check, :146-157)
    if not isinstance(node, list) or len(node) > n:
        raise SceneError(f"No mapping from TOML {_toml_type(node)} to vector<float, {n}> ({what})")
    out = [float(x) for x in default]
    for i, c in enumerate(node):
        out[i] = _finite_float(c, what)
    return tuple(out)


def _colour(node: Any, default: tuple[float, float, float, float]) -> tuple[float, float, float, float]:
    """src/scene.cpp:184-356"""
    if node is None:
        return default
    if isinstance(node, str):
        return named_colour(node)
    if not isinstance(node, list) or len(node) > 4:
        raise SceneError(f"No mapping from TOML {_toml_type(node)} to colour")
    out = [0.0, 0.0, 0.0, 0.0]
    for i, c in enumerate(node):
        out[i] = _finite_float(c, "colour")
    if len(node) < 4:
        out[3] = 1.0
    return tuple(out)  # type: ignore[return-value]


def _unsigned(node: Any, default: int, what: str) -> int:
    """`node.value<unsigned>()` (toml++, rules as recalled, UNVERIFIED): an integer inside [0, UINT_MAX] -- out-of-range values
    have no mapping, they do not wrap --, a float holding a whole number inside that range, a boolean as 0 / 1"""
    if node is None:
        return default
    no_mapping = SceneError(f"No mapping from TOML {_toml_type(node)} to unsigned ({what})")
    if isinstance(node, bool):
        v = int(node)
    elif isinstance(node, int):
        v = node
    elif isinstance(node, float):
        if not math.isfinite(node) or not -9.2e18 <= node <= 9.2e18 or float(int(node)) != node:
            raise no_mapping
        v = int(node)
    else:
        raise no_mapping
    if not 0 <= v <= 0xFFFFFFFF:
        raise no_mapping
    return v


def _material_type(node: Any) -> int:
    """src/scene.cpp:381-404 (magic_enum by integer or by name)"""
    if node is None:
        return LAMBERT
    if isinstance(node, bool):
        raise SceneError("No mapping from TOML boolean to material_type")
    if isinstance(node, int):
        if not 0 <= node < len(MATERIAL_TYPES):
            raise SceneError(f"integer value {node} was not a member of enum material_type")
        return node
    if isinstance(node, str):
        if node not in MATERIAL_TYPES:
            raise SceneError(f"string value '{node}' was not a member of enum material_type")
        return MATERIAL_TYPES.index(node)
    raise SceneError(f"No mapping from TOML {_toml_type(node)} to material_type")


def _table_array(cfg: dict, key: str) -> list:
    node = cfg.get(key)
    if node is None:
        return []
    if not isinstance(node, list):
        raise SceneError(f"expected array at key '{key}', got {_toml_type(node)}")
    # an element that is not a table reads as an empty one, as in the reference: its lookups go through toml::node_view's
    # operator[] (src/scene.cpp:420-430), which yields an empty view for a non-table parent, so every field takes its default
    return [t if isinstance(t, dict) else {} for t in node]


def _f32(x: float) -> np.float32:
    return np.float32(x)


def _round_f32(x: fractions.Fraction) -> np.float32:
    """The binary32 nearest to the exact rational x (ties to even): one rounding, as a hardware fma performs it."""
    f = np.float32(float(x))  # within one ulp; decide exactly between it and its neighbours
    if not np.isfinite(f):
        return f
    best, best_err = f, abs(fractions.Fraction(float(f)) - x)
    for g in (np.nextafter(f, np.float32(-np.inf)), np.nextafter(f, np.float32(np.inf))):
        if not np.isfinite(g):
            continue
        err = abs(fractions.Fraction(float(g)) - x)
        if err < best_err or (err == best_err and (int(g.view(np.uint32)) & 1) == 0 and (int(best.view(np.uint32)) & 1) == 1):
            best, best_err = g, err
    return best


def _fma32(a, b, c) -> np.float32:
    """fmaf(a, b, c): the product is not rounded (DESIGN.md SPEC, S1-S3)."""
    a, b, c = np.float32(a), np.float32(b), np.float32(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        with np.errstate(invalid="ignore"):
            return np.float32(a * b + c)
    exact = fractions.Fraction(float(a)) * fractions.Fraction(float(b)) + fractions.Fraction(float(c))
    if exact == 0:  # the sign of an exact zero follows IEEE addition; the binary64 product of two binary32 values is exact
        return np.float32(float(a) * float(b) + float(c))
    return _round_f32(exact)


def _dot3(a, b) -> np.float32:
    """S1: dot3(a, b) = fma(a.z, b.z, fma(a.y, b.y, a.x * b.x))"""
    return _fma32(a[2], b[2], _fma32(a[1], b[1], np.float32(a[0]) * np.float32(b[0])))


@dataclasses.dataclass
class Camera:
    """rt::camera (src/camera.hpp:51-138): vfov pi/4, near 0.01, far 1000 are fixed (private, no setter)."""
    position: tuple[float, float, float] = (0.0, 1.0, 0.0)
    direction: tuple[float, float, float] = (0.0, 0.0, -1.0)
    vfov: float = math.pi / 4
    near: float = 0.01
    far: float = 1000.0


@dataclasses.dataclass
class Scene:
    """rt::scene (src/scene.hpp:8-25) with the soagen tables as numpy columns (src/soa.toml)."""
    samples_per_pixel: int = 30
    max_bounces: int = 10
    path: str = ""
    camera: Camera = dataclasses.field(default_factory=Camera)
    materials: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, MATERIAL_DTYPE))
    material_names: list = dataclasses.field(default_factory=list)
    spheres: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0, 4), np.float32))  # spheres.value()
    sphere_material: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.uint32))
    planes: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0, 4), np.float32))  # planes.value()
    plane_material: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.uint32))
    boxes: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0, 6), np.float32))  # centre, extents
    box_material: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.uint32))

    def validate(self) -> None:
        if len(self.materials) == 0:
            raise SceneError("scene has no materials")
        for name, col in (("sphere", self.sphere_material), ("plane", self.plane_material)):
            if len(col) and int(col.max()) >= len(self.materials):
                raise SceneError(f"{name} material index {int(col.max())} out-of-range")


def make_materials(rows: Sequence[tuple]) -> np.ndarray:
    """rows of (type, albedo rgba, roughness, reflectivity)"""
    m = np.zeros(len(rows), MATERIAL_DTYPE)
    for i, (t, albedo, rough, refl) in enumerate(rows):
        a = list(albedo) + [1.0] * (4 - len(albedo))
        m[i] = (t, a, rough, refl)
    return m


def _nesting(value) -> int:
    """levels of arrays / tables below and including `value` (iterative: the point is documents nested hundreds deep)"""
    deepest, todo = 0, [(value, 1)]
    while todo:
        v, d = todo.pop()
        if isinstance(v, (list, dict)):
            deepest = max(deepest, d)
            todo.extend((c, d + 1) for c in (v.values() if isinstance(v, dict) else v))
    return deepest


def loads(text: str, path: str = "") -> Scene:
    """scene::load on TOML text (src/scene.cpp:527-618)."""
    try:
        cfg = tomllib.loads(text)
    except tomllib.TOMLDecodeError as e:
        raise SceneError(f"TOML parse error: {e}") from None
    except RecursionError:  # arrays / inline tables nested thousands deep: a parse error in toml++ too (256 levels)
        raise SceneError("TOML parse error: exceeded maximum nested value depth of 256") from None
    if _nesting(cfg) > 256 + 1:  # (+1: the document itself is a table) -- the limit of toml++ and of the C++ twin's parser
        raise SceneError("TOML parse error: exceeded maximum nested value depth of 256")

    s = Scene(path=path)
    s.samples_per_pixel = min(max(_unsigned(cfg.get("samples_per_pixel"), 30, "samples_per_pixel"), 1), 1000)
    s.max_bounces = min(max(_unsigned(cfg.get("max_bounces"), 10, "max_bounces"), 1), 1000)

    cam = cfg.get("camera")
    if cam is not None:
        if not isinstance(cam, dict):
            raise SceneError(f"expected table at key 'camera', got {_toml_type(cam)}")
        s.camera = Camera(position=_vector(cam.get("position"), (0, 1, 0), "camera.position"),
                          direction=_vector(cam.get("direction"), (0, 0, -1), "camera.direction"))

    rows, names = [], []
    for tbl in _table_array(cfg, "materials"):
        t = _material_type(tbl.get("type"))
        refl_default = _DEFAULT_REFLECTIVITY.get(t, 0.5)
        name = tbl.get("name", "")
        if not isinstance(name, str):
            raise SceneError("No mapping from TOML value to string (name)")
        albedo = _colour(tbl.get("albedo"), named_colour("fuchsia"))
        rough = _finite_float(tbl["roughness"], "roughness") if "roughness" in tbl else (0.0 if t == DIELECTRIC else 0.5)
        refl = _finite_float(tbl["reflectivity"], "reflectivity") if "reflectivity" in tbl else refl_default
        rows.append((t, albedo, rough, refl))
        names.append(name)
    if not rows:  # src/scene.cpp:565-566
        rows.append((LAMBERT, named_colour("fuchsia"), 0.05, 0.5))
        names.append("")
    s.materials = make_materials(rows)
    s.material_names = names

    def material_of(tbl: dict) -> int:
        m = _unsigned(tbl.get("material"), 0, "material")
        if m >= len(rows):
            raise SceneError(f"material index {m} out-of-range")
        return m

    planes, plane_mat = [], []
    for tbl in _table_array(cfg, "planes"):
        pos = np.array(_vector(tbl.get("position"), (0, 0, 0), "plane.position"), np.float32)
        n = np.array(_vector(tbl.get("normal"), (0, 1, 0), "plane.normal"), np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):  # a zero normal normalises to NaN, as in the reference
            n = (n * (_f32(1.0) / np.sqrt(_dot3(n, n), dtype=np.float32))).astype(np.float32)  # S2: v * (1 / sqrt(dot3(v, v)))
        d = -_dot3(n, pos)  # muu plane{position, normal}: dot(n, p) + d == 0 (UNVERIFIED), dot as S1
        planes.append((n[0], n[1], n[2], d))
        plane_mat.append(material_of(tbl))

    spheres, sphere_mat = [], []
    for tbl in _table_array(cfg, "spheres"):
        pos = _vector(tbl.get("position"), (0, 1, -3), "sphere.position")
        radius = _finite_float(tbl["radius"], "radius") if "radius" in tbl else 0.5
        spheres.append((*pos, radius))
        sphere_mat.append(material_of(tbl))

    boxes, box_mat = [], []
    for tbl in _table_array(cfg, "boxes"):
        pos = _vector(tbl.get("position"), (0, 1, -3), "box.position")
        ext = _vector(tbl.get("extents"), (0.5, 0.5, 0.5), "box.extents")
        boxes.append((*pos, *ext))
        box_mat.append(material_of(tbl))

    s.planes = np.array(planes, np.float32).reshape(-1, 4)
    s.plane_material = np.array(plane_mat, np.uint32)
    s.spheres = np.array(spheres, np.float32).reshape(-1, 4)
    s.sphere_material = np.array(sphere_mat, np.uint32)
    s.boxes = np.array(boxes, np.float32).reshape(-1, 6)
    s.box_material = np.array(box_mat, np.uint32)
    return s


_SEARCH_PREFIXES = ("scenes/", "../scenes/", "../../scenes/", "", "../", "../../")  # src/scene.cpp:479-480


def load(path: str | pathlib.Path) -> Scene:
    """scene::load(file) including the relative-path search (src/scene.cpp:483-525); "-" reads standard input."""
    if not str(path):
        raise SceneError("no scene file path provided")
    if str(path) == "-":
        return loads(sys.stdin.read(), "")
    p = pathlib.Path(path)
    found = None
    if not p.is_absolute():
        for root in _SEARCH_PREFIXES:
            q = pathlib.Path(root) / p if root else p
            if q.is_file():
                found = q
                break
    elif p.is_file():
        found = p
    if found is None:
        raise SceneError(f"scene path '{p}' did not exist or was not a file")
    return loads(found.read_text(), str(found))


def load_first_available() -> Scene:
    """scene::load_first_available (src/scene.cpp:620-643; what the app loads when --scene is not given, main.cpp:121-125): the
    first regular *.toml file of the first search directory that has one, in the directory's own iteration order."""
    for root in _SEARCH_PREFIXES:
        d = pathlib.Path(root)
        if not root or not d.is_dir():  # the empty prefix is not a directory (fs::status("") is not_found): the cwd is not searched
            continue
        with os.scandir(d) as it:  # the same (unsorted) order as std::filesystem::directory_iterator
            for entry in it:
                f = pathlib.Path(entry.path)
                if not f.stem or f.suffix != ".toml" or not f.is_file():
                    continue
                return load(f)
    raise SceneError("no scene files found")


def dumps(s: Scene) -> str:
    """Emit a scene as TOML the reference loader accepts (albedos as float arrays, types by name)."""
    def fl(x: float) -> str:
        r = repr(float(np.float32(x)))
        return r if any(c in r for c in ".en") else r + ".0"

    def vec(v: Sequence[float]) -> str:
        return "[" + ", ".join(fl(x) for x in v) + "]"

    out = [f"samples_per_pixel = {s.samples_per_pixel}", f"max_bounces = {s.max_bounces}", "",
           f"camera = {{ position = {vec(s.camera.position)}, direction = {vec(s.camera.direction)} }}", "", "materials = ["]
    for m in s.materials:
        out.append(f"    {{ type = '{MATERIAL_TYPES[int(m['type'])]}', albedo = {vec(m['albedo'])}, "
                   f"roughness = {fl(m['roughness'])}, reflectivity = {fl(m['reflectivity'])} }},")
    out += ["]", "", "spheres = ["]
    for sp, mat in zip(s.spheres, s.sphere_material):
        out.append(f"    {{ material = {int(mat)}, position = {vec(sp[:3])}, radius = {fl(sp[3])} }},")
    out += ["]", ""]
    if len(s.planes):
        out.append("planes = [")
        for pl, mat in zip(s.planes, s.plane_material):
            n = pl[:3].astype(np.float64)
            pos = -float(pl[3]) * n  # a point on the plane: the reloaded offset -dot3(n, pos) equals d up to rounding (exact for axis-aligned planes)
            out.append(f"    {{ material = {int(mat)}, position = {vec(pos)}, normal = {vec(pl[:3])} }},")
        out += ["]", ""]
    boxes, box_material = getattr(s, "boxes", ()), getattr(s, "box_material", ())
    if len(boxes):
        out.append("boxes = [")
        for bx, mat in zip(boxes, box_material):
            out.append(f"    {{ material = {int(mat)}, position = {vec(bx[:3])}, extents = {vec(bx[3:6])} }},")
        out += ["]", ""]
    return "\n".join(out)
