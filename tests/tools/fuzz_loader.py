#!/usr/bin/env python3
"""Differential fuzz of the two scene loaders that restate scene.cpp:483-618: rt_b200/scene.py (tomllib) and the C++ host
(rt_b200/host: toml_lite.hpp + scene_loader.hpp, through `rt_headless --dump-scene`).

Two families of inputs:
  * generated: random scene files written in every value syntax the reference loader accepts (aliases, scalar broadcast, short
    arrays, ints / hex / underscores / exponents, inline tables vs [[tables]], dotted keys, both string kinds, comments), with a
    share of deliberate semantic errors (out-of-range material, NaN, unknown alias, wrong type, over-long array).  Both loaders
    must agree on accept / reject, and on every loaded value bit for bit.
  * mutated: the same files with random byte edits.  The C++ host must never crash (exit code 0 or 1, `error: ...` on stderr);
    where both parsers accept the file the loaded values must still agree.

`--asan` runs a build of the C++ host with -fsanitize=address,undefined (made under /tmp).  Usage:
    python tests/tools/fuzz_loader.py [--cases N] [--seed S] [--asan] [--out profiles/fuzz_loader_rN.json]"""
import argparse, json, pathlib, random, subprocess, sys, tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from rt_b200 import build, scene as S  # noqa: E402
from rt_b200.colour_table import NAMED_COLOURS as COLOURS  # noqa: E402

VEC_ALIASES = ["origin", "zero", "one", "forward", "back", "backward", "up", "down", "left", "right", "x", "x_axis", "y", "y_axis", "z", "z_axis"]
TYPES = ["lambert", "metal", "dielectric", "air", "vacuum", "water", "ice", "diamond"]


class Gen:
    def __init__(self, rng: random.Random, p_error: float):
        self.r = rng
        self.p_error = p_error

    def num(self, lo=-5.0, hi=5.0, allow_int=True) -> str:
        r = self.r
        k = r.random()
        if allow_int and k < 0.25:
            v = r.randint(int(lo), int(hi))
            form = r.random()
            if form < 0.15 and v >= 0:
                return hex(v)
            if form < 0.25 and v >= 0:
                return "+" + str(v)
            return str(v)
        v = r.uniform(lo, hi)
        if k < 0.5:
            return repr(round(v, r.randint(0, 6)) + 0.0)
        if k < 0.7:
            return f"{v:.{r.randint(1, 8)}e}"
        if k < 0.8:
            return f"{v:.3f}".replace(".", "_0.", 1) if False else f"{v:.3f}"
        return repr(v)

    def string(self, s: str) -> str:
        """`s` in one of TOML's string forms; the loaders see the same text (aliases and colour names are looked up by it)"""
        r = self.r
        k = r.random()
        if k < 0.4:
            return f"'{s}'"
        if k < 0.8 or not s:
            return f'"{s}"'
        if k < 0.88:  # one character as a unicode escape
            i = r.randrange(len(s))
            esc = f"\\u{ord(s[i]):04x}" if r.random() < 0.5 else f"\\U{ord(s[i]):08X}"
            return '"' + s[:i] + esc + s[i + 1:] + '"'
        if k < 0.94:  # multi-line basic: leading newline dropped, line-ending backslash joins
            i = r.randrange(len(s) + 1)
            return '"""' + r.choice(["", "\n"]) + s[:i] + r.choice(["", "\\\n   ", "\\  \n\n\t"]) + s[i:] + '"""'
        return "'''" + r.choice(["", "\n"]) + s + "'''"

    def name(self) -> str:
        """a material name (loaded and dropped): every string form, escapes, quotes inside multi-line strings"""
        r = self.r
        word = "m" + str(r.randint(0, 99))
        k = r.random()
        if k < 0.5:
            return self.string(word)
        if k < 0.6:
            return '"' + word + r.choice(["\\t", "\\n", "\\\\", '\\"', "\\u00e9", "\\U0001F600", "\\b\\f\\r"]) + '"'
        if k < 0.7:
            return '"' + word + r.choice(["\\q", "\\x41", "\\u12", "\\UD8000000", "\\uD800"]) + '"' if r.random() < self.p_error else self.string(word)
        if k < 0.85:
            return '"""' + r.choice(["", "\n"]) + word + r.choice(["", " \\\n    joined", "\nsecond line", ' "quoted" ', ' ""two'] ) + r.choice(['"""', '""""', '"""""'])
        return "'''" + r.choice(["", "\n"]) + word + r.choice(["", "\nraw \\n line", " 'q' ", " ''two"]) + r.choice(["'''", "''''"])

    def unsigned(self, lo: int, hi: int) -> str:
        """a value for an unsigned field: mostly integers, sometimes the other number-like forms toml++'s value<unsigned>() is asked to
        convert (whole / fractional floats, booleans, negative and > 32-bit integers)"""
        r = self.r
        k = r.random()
        if k < 0.7:
            return str(r.randint(lo, hi))
        return r.choice([f"{r.randint(lo, hi)}.0", f"{r.randint(lo, hi)}e0", f"{r.randint(lo, hi)}.5", "true", "false", str(-r.randint(1, 9)), "4294967295", "4294967296",
                         "9223372036854775807", "1e19", "-0.0", "inf"])

    def bad_number(self) -> str:
        return self.r.choice(["nan", "inf", "-inf", "+nan"])

    def vec(self, n=3, lo=-5.0, hi=5.0) -> str:
        r = self.r
        if r.random() < self.p_error:
            return r.choice([self.string("sideways"), "[" + ", ".join(self.num() for _ in range(n + 1)) + "]", "true", "[1, " + self.bad_number() + "]",
                             self.bad_number(), "{ x = 1 }"])
        k = r.random()
        if k < 0.15:
            return self.string(r.choice(VEC_ALIASES))
        if k < 0.3:
            return self.num(lo, hi)
        m = n if k < 0.8 else r.randint(0, n)
        nl = "\n   " if r.random() < 0.1 else " "
        body = ("," + nl).join(self.num(lo, hi) for _ in range(m))
        return "[" + body + (", " if m and r.random() < 0.2 else "") + "]"

    def colour(self) -> str:
        r = self.r
        if r.random() < self.p_error:
            return r.choice([self.string("octarine"), "[0.1, 0.2, 0.3, 0.4, 0.5]", "3", "[0.5, " + self.bad_number() + "]"])
        if r.random() < 0.4:
            return self.string(r.choice(sorted(COLOURS)))
        return "[" + ", ".join(self.num(0, 1, allow_int=False) for _ in range(r.randint(0, 4))) + "]"

    def material(self) -> list[tuple[str, str]]:
        r = self.r
        kv = []
        if r.random() < 0.3:
            kv.append(("name", self.name()))
        if r.random() < 0.8:
            if r.random() < self.p_error:
                kv.append(("type", r.choice(["8", "-1", self.string("glass"), "1.5", "[1]"])))
            else:
                t = r.randrange(8)
                kv.append(("type", str(t) if r.random() < 0.4 else self.string(TYPES[t])))
        if r.random() < 0.8:
            kv.append(("albedo", self.colour()))
        if r.random() < 0.5:
            kv.append(("roughness", self.num(0, 1) if r.random() > self.p_error else r.choice([self.bad_number(), self.string("rough")])))
        if r.random() < 0.5:
            kv.append(("reflectivity", self.num(0, 2)))
        r.shuffle(kv)
        return kv

    def prim(self, kind: str, n_mats: int) -> list[tuple[str, str]]:
        r = self.r
        kv = []
        if r.random() < 0.85:
            kv.append(("position", self.vec()))
        if kind == "spheres" and r.random() < 0.8:
            kv.append(("radius", self.num(0.05, 3) if r.random() > self.p_error else r.choice([self.bad_number(), "[1]", self.string("big"), "16777216", "16777217",
                                                                                                "-16777217", "1e39", "3.4028234e38", "3.4028236e38", "true"])))
        if kind == "planes" and r.random() < 0.7:
            kv.append(("normal", self.vec(lo=-1, hi=1)))
        if kind == "boxes" and r.random() < 0.7:
            kv.append(("extents", self.vec(lo=0.1, hi=2)))
        if r.random() < 0.7:
            m = r.randrange(max(n_mats, 1))
            if r.random() < self.p_error:
                m = r.choice([n_mats + r.randint(0, 3) if n_mats else 1 + r.randint(0, 3), -1])
            kv.append(("material", str(m) if r.random() < 0.9 else r.choice([f"{m}.0", "false", f"{m}.25"])))
        r.shuffle(kv)
        return kv

    def table_list(self, key: str, rows: list[list[tuple[str, str]]], out_top: list[str], out_tail: list[str]):
        """Writes `rows` either as an inline array of inline tables (top of the file) or as [[key]] sections (tail)."""
        r = self.r
        if not rows:
            if r.random() < 0.3:
                out_top.append(f"{key} = []")
            return
        if r.random() < 0.5:
            sep = ",\n  " if r.random() < 0.5 else ", "
            items = ["{ " + ", ".join(f"{k} = {v}" for k, v in row if "\n" not in v or k == "name") + " }" for row in rows]
            if r.random() < 0.05:  # an element that is not a table reads as all defaults
                items.insert(r.randrange(len(items) + 1), r.choice(["1", "'x'", "[1, 2]", "true", "0.5"]))
            out_top.append(f"{key} = [{sep.join(items)}{',' if r.random() < 0.2 else ''}]")
        else:
            for row in rows:
                out_tail.append(f"[[{key}]]" + ("  # " + key if r.random() < 0.2 else ""))
                out_tail.extend(f"{k} = {v}" for k, v in row)
                if r.random() < 0.3:
                    out_tail.append("")

    def document(self) -> str:
        r = self.r
        top, tail = [], []
        if r.random() < 0.2:
            top.append("# generated by tests/tools/fuzz_loader.py")
        if r.random() < 0.6:
            v = self.unsigned(0, 2000) if r.random() < 0.5 else str(r.randint(1, 64))
            top.append(f"samples_per_pixel = {v if r.random() > self.p_error else self.r.choice(['1.5', self.string('many'), '[4]'])}")
        if r.random() < 0.6:
            top.append(f"max_bounces = {self.unsigned(0, 2000) if r.random() > 0.2 else f'{r.randint(1, 9)}_000'}")
        cam = []
        if r.random() < 0.7:
            cam.append(("position", self.vec()))
        if r.random() < 0.7:
            cam.append(("direction", self.vec(lo=-1, hi=1)))
        if cam:
            style = r.random()
            if style < 0.33:
                top.extend(f"camera.{k} = {v}" for k, v in cam)
            elif style < 0.66 and all("\n" not in v for _, v in cam):
                top.append("camera = { " + ", ".join(f"{k} = {v}" for k, v in cam) + " }")
            else:
                tail.append("[camera]")
                tail.extend(f"{k} = {v}" for k, v in cam)
        n_mats = r.choice([0, 1, 1, 2, 3, 5])
        self.table_list("materials", [self.material() for _ in range(n_mats)], top, tail)
        for kind in ("planes", "spheres", "boxes"):
            self.table_list(kind, [self.prim(kind, n_mats) for _ in range(r.choice([0, 0, 1, 2, 4, 9]))], top, tail)
        r.shuffle(top)
        return "\n".join(top + [""] + tail) + "\n"


def mutate(text: str, r: random.Random) -> str:
    b = bytearray(text.encode())
    for _ in range(r.randint(1, 4)):
        if not b:
            break
        k, i = r.random(), r.randrange(len(b))
        if k < 0.3:
            del b[i:i + r.randint(1, 6)]
        elif k < 0.6:
            b[i] = r.choice(b"[]{}=,.'\"#\n\\ \t0_-+xeE\x00\xff\x80ab")
        elif k < 0.8:
            j = r.randrange(len(b))
            b[i:i] = b[j:j + r.randint(1, 12)]
        else:
            b.insert(i, r.choice(b"[]{}=,\"'\n"))
    if r.random() < 0.02:  # values nested beyond any parser's patience (toml++ gives up at 256 levels): a parse error on both sides
        depth = r.choice([257, 300, 5000, 200000])
        b += ("\ndeep = " + ("[" * depth + "]" * depth if r.random() < 0.5 else "{a = " * depth + "1" + "}" * depth) + "\n").encode()
    return b.decode("utf-8", errors="surrogateescape")


def python_load(text: str):
    try:
        return S.loads(text), None
    except S.SceneError as e:
        return None, str(e)
    except Exception as e:  # tomllib rejects the file (syntax, invalid UTF-8, duplicate keys ...)
        return None, f"{type(e).__name__}: {e}"


def same_scene(d: dict, s: S.Scene) -> str | None:
    def eq(a, b, dtype=np.float32):
        a, b = np.asarray(a, dtype).ravel(), np.asarray(b, dtype).ravel()
        if a.shape != b.shape:
            return False
        if dtype is np.float32:  # bit for bit (signed zeros included); a NaN matches any NaN
            both_nan = np.isnan(a) & np.isnan(b)
            return bool(np.all(both_nan | (a.view(np.uint32) == b.view(np.uint32))))
        return np.array_equal(a, b)
    if d["samples_per_pixel"] != s.samples_per_pixel or d["max_bounces"] != s.max_bounces:
        return "sampling fields"
    if not eq(d["camera"]["position"], s.camera.position) or not eq(d["camera"]["direction"], s.camera.direction):
        return "camera"
    if len(d["materials"]) != len(s.materials):
        return "material count"
    for a, b in zip(d["materials"], s.materials):
        if a["type"] != int(b["type"]) or not eq(a["albedo"], b["albedo"]) or not eq([a["roughness"], a["reflectivity"]], [b["roughness"], b["reflectivity"]]):
            return "material"
    for key, width in (("spheres", 4), ("planes", 4), ("boxes", 6)):
        if not eq(d[key], getattr(s, key)):
            return key
    if not eq(d["sphere_material"], s.sphere_material, np.uint32) or not eq(d["plane_material"], s.plane_material, np.uint32):
        return "material index"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--asan", action="store_true")
    ap.add_argument("--out")
    args = ap.parse_args()

    build.build_cuda()
    cli = str(build.build_host())
    if args.asan:
        cli = "/tmp/rt_headless_asan"
        src = ROOT / "rt_b200" / "host" / "rt_headless.cpp"
        subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", f"-I{ROOT / 'include'}", "-o", cli,
                        str(src), f"-L{ROOT / 'rt_b200' / 'lib'}", "-lrtcu", f"-Wl,-rpath,{ROOT / 'rt_b200' / 'lib'}"], check=True)

    rng = random.Random(args.seed)
    stats = {"cases": 0, "generated": 0, "mutated": 0, "both_accept": 0, "both_reject": 0, "value_mismatch": 0, "accept_mismatch_generated": 0,
             "syntax_disagreement_mutated": 0, "crashes": 0, "asan": bool(args.asan), "seed": args.seed}
    failures = []
    with tempfile.TemporaryDirectory() as tmp:
        path = pathlib.Path(tmp) / "case.toml"
        for case in range(args.cases):
            gen = Gen(rng, p_error=rng.choice([0.0, 0.0, 0.03, 0.1]))
            text = gen.document()
            mutated = case % 3 == 2
            if mutated:
                text = mutate(text, rng)
            path.write_bytes(text.encode("utf-8", errors="surrogateescape"))
            try:
                text_py = path.read_bytes().decode("utf-8")
                py_scene, py_err = python_load(text_py)
            except UnicodeDecodeError as e:
                py_scene, py_err = None, f"invalid UTF-8: {e}"
            r = subprocess.run([cli, "--scene", str(path), "--dump-scene"], capture_output=True, text=True, errors="replace",
                               env={"ASAN_OPTIONS": "detect_leaks=0", "PATH": "/usr/bin:/bin"})
            stats["cases"] += 1
            stats["mutated" if mutated else "generated"] += 1
            ok_exit = r.returncode == 0 or (r.returncode == 1 and any(ln.startswith("error: ") for ln in r.stderr.splitlines()) and "Sanitizer" not in r.stderr
                                            and "runtime error" not in r.stderr)
            if not ok_exit:
                stats["crashes"] += 1
                failures.append({"case": case, "kind": "crash", "rc": r.returncode, "stderr": r.stderr[-400:], "text": text})
                continue
            cpp_ok = r.returncode == 0
            if cpp_ok and py_scene is not None:
                why = same_scene(json.loads(r.stdout), py_scene)
                stats["both_accept"] += 1
                if why:
                    stats["value_mismatch"] += 1
                    failures.append({"case": case, "kind": "value:" + why, "text": text})
            elif not cpp_ok and py_scene is None:
                stats["both_reject"] += 1
            elif mutated:
                # the two TOML readers may disagree on what is well-formed at the edges of the grammar (toml_lite is a subset reader)
                stats["syntax_disagreement_mutated"] += 1
                failures.append({"case": case, "kind": "syntax-disagreement", "cpp": r.stderr.strip()[-200:], "py": py_err, "text": text})
            else:
                stats["accept_mismatch_generated"] += 1
                failures.append({"case": case, "kind": "accept-mismatch", "cpp": r.stderr.strip()[-200:], "py": py_err, "text": text})
    print(json.dumps(stats))
    hard = [f for f in failures if f["kind"] != "syntax-disagreement"]
    for f in (hard + [f for f in failures if f["kind"] == "syntax-disagreement"])[:12]:
        print("----", f["kind"], "case", f["case"], {k: v for k, v in f.items() if k not in ("text", "kind", "case")})
        print(f["text"])
    if args.out:
        pathlib.Path(args.out).write_text(json.dumps(stats) + "\n")
    return 1 if hard else 0


if __name__ == "__main__":
    sys.exit(main())
