"""The device's BVH traversal replayed on the CPU (tests/tools/bvh_replay.c: margins, culls, visiting order, (t, index) rule, S4 leaf
arithmetic) over the very arrays the library uploads, against the oracle's linear scan (mg_ray_tracer.cpp:62-87): hit flag, sphere index
and t must be identical -- for the default tree and for every experimental builder / collapse variant, on random, silhouette-grazing and
on-surface rays.  The GPU tests check the same claim on the device; this one checks the algorithm where there is time for many rays."""
import ctypes as C
import pathlib
import subprocess

import numpy as np
import pytest

from rt_b200 import renderer as R, scene as S, synth

ROOT = pathlib.Path(__file__).resolve().parent.parent


def build_replay(tmp_path_factory, *defines):
    so = tmp_path_factory.mktemp("bvh_replay") / "libbvh_replay.so"
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-mfma", *defines, "-o", str(so),
                    str(ROOT / "tests" / "tools" / "bvh_replay.c"), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.bvh_replay_batch.restype = C.c_int
    lib.bvh_replay_batch.argtypes = [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 5 + [C.c_int]

    def run(nodes, leaves, o, d, any_t=False):
        n = len(o)
        hit, prim, t, skipped = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        counters = np.zeros(2, np.uint64)
        depth = lib.bvh_replay_batch(nodes.ctypes.data, leaves.ctypes.data, o.ctypes.data, d.ctypes.data, n, hit.ctypes.data, prim.ctypes.data,
                                     t.ctypes.data, skipped.ctypes.data, counters.ctypes.data, int(any_t))
        assert depth >= 0, "traversal stack overflow"
        return hit, prim, t, skipped.astype(bool), counters, depth
    return run


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    """trav_step: leaf children tested inside the node visit (the preview renderer; RTCU_BVH_TRAV=0)"""
    return build_replay(tmp_path_factory)


@pytest.fixture(scope="module")
def replay2(tmp_path_factory):
    """closest_sphere_bvh2: leaves deferred and ordered with the inner children (the path tracer's default)"""
    return build_replay(tmp_path_factory, "-DBVH_REPLAY_TRAV2")


def unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def rays_for(sph: np.ndarray, n: int, rng) -> tuple[np.ndarray, np.ndarray]:
    """a quarter each: random rays through the scene's bounds; rays grazing a sphere's silhouette from 0.3 - 400 units away (the
    discriminant near zero is where rounding decides hit or miss); rays leaving a sphere's surface (origin on the surface: the
    near-root-below-epsilon rule); rays from inside a sphere"""
    lo, hi = (sph[1:, :3] - sph[1:, 3:4]).min(axis=0), (sph[1:, :3] + sph[1:, 3:4]).max(axis=0)  # sphere 0 may be a huge ground sphere
    q = n // 4
    o1 = rng.uniform(lo - 5, hi + 5, (q, 3))
    d1 = unit(rng.uniform(lo, hi, (q, 3)) - o1)
    pick = rng.integers(0, len(sph), q)
    c, r = sph[pick, :3].astype(np.float64), sph[pick, 3:4].astype(np.float64)
    away = unit(rng.normal(size=(q, 3))).astype(np.float64)
    dist = np.exp(rng.uniform(np.log(0.3), np.log(400.0), (q, 1)))
    o2 = c + away * (r + dist)
    side = np.cross(away, unit(rng.normal(size=(q, 3))).astype(np.float64))
    side /= np.linalg.norm(side, axis=1, keepdims=True)
    rim = c + side * r * (1.0 + rng.choice([-1e-6, -1e-7, 0.0, 1e-7, 1e-6], (q, 1)))  # just inside / on / just outside the silhouette
    d2 = unit(rim - o2)
    nrm = unit(rng.normal(size=(q, 3))).astype(np.float64)
    pick3 = rng.integers(0, len(sph), q)
    o3 = sph[pick3, :3] + nrm * sph[pick3, 3:4]
    d3 = unit(nrm + unit(rng.uniform(0, 1, (q, 3))) * rng.choice([1.0, -1.0], (q, 1)))  # leaving or entering, like a scatter / refraction
    pick4 = rng.integers(0, len(sph), n - 3 * q)
    o4 = sph[pick4, :3] + unit(rng.normal(size=(n - 3 * q, 3))) * sph[pick4, 3:4] * rng.uniform(0, 0.99, (n - 3 * q, 1))
    d4 = unit(rng.normal(size=(n - 3 * q, 3)))
    o = np.concatenate([o1, o2, o3, o4]).astype(np.float32)
    d = np.concatenate([d1, d2, d3, d4]).astype(np.float32)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def scene_of(sph: np.ndarray) -> S.Scene:
    sc = S.loads("")
    sc.spheres = np.ascontiguousarray(sph, np.float32)
    sc.sphere_material = np.zeros(len(sph), np.uint32)
    return sc


def cases():
    rng = np.random.default_rng(11)
    cloud = np.concatenate([rng.uniform(-30, 30, (3000, 3)), rng.uniform(0.05, 1.5, (3000, 1))], axis=1).astype(np.float32)
    cloud[0] = [0, -1000, 0, 1000]
    cloud[1] = cloud[2]  # two identical spheres: the lower index must win
    nested = np.concatenate([np.zeros((40, 3)), np.linspace(0.5, 20, 40)[:, None]], axis=1).astype(np.float32)  # concentric shells
    return {"rtiow": (synth.rtiow_scene().spheres, 60000), "cloud": (cloud, 40000), "nested": (nested, 20000),
            "grid9601": (synth.grid_scene(nx=120, nz=80).spheres, 24000), "grid100k": (synth.grid_scene().spheres, 4000)}


KNOBS = [{}, {"RTCU_BVH_LEAF_COST": "1"}, {"RTCU_BVH_COLLAPSE": "sah"}, {"RTCU_BVH_SWEEP": "512", "RTCU_BVH_LEAF_COST": "1", "RTCU_BVH_COLLAPSE": "sah"}]


@pytest.mark.parametrize("trav", [2, 1], ids=["deferred-leaves", "leaves-in-visit"])
@pytest.mark.parametrize("knobs", KNOBS, ids=["default", "leafcost", "sah-collapse", "sweep+leafcost+sah-collapse"])
@pytest.mark.parametrize("name", ["rtiow", "cloud", "nested", "grid9601", "grid100k"])
def test_replayed_traversal_equals_the_oracle_scan(replay, replay2, oracle, name, knobs, trav, monkeypatch):
    if knobs and name == "grid100k":
        pytest.skip("variants: the smaller scenes")
    replay = replay2 if trav == 2 else replay
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    sph, n = cases()[name]
    sph = np.ascontiguousarray(sph, np.float32)
    nodes, leaves, depth = R.bvh4_build_host(sph)
    o, d = rays_for(sph, n, np.random.default_rng(hash(name) % 1000))
    hit, prim, t, skipped, counters, stack = replay(nodes, leaves, o, d)
    ref_hit, ref_prim, ref_t, _ = oracle.intersect_batch(scene_of(sph), o, d)
    assert not skipped.any()                     # all directions are unit length to within the kernel's 1e-3
    assert stack <= 3 * depth + 2 <= 64          # what rtcu_upload_scene validates against BVH_STACK
    assert np.array_equal(hit, ref_hit)
    assert np.array_equal(prim[hit == 1], ref_prim[hit == 1])
    assert np.array_equal(t.view(np.uint32), ref_t.view(np.uint32))
    assert 0.2 < hit.mean() < 0.999 and counters[0] < len(sph) * n / 4  # the rays do hit things, and the tree does cull


def test_replay_declines_non_unit_directions(replay):
    sph = synth.rtiow_scene().spheres
    nodes, leaves, _ = R.bvh4_build_host(sph)
    o = np.zeros((3, 3), np.float32)
    d = np.float32([[0, 0, -1], [0, 0, -1.01], [0, 0.5, -0.5]])
    _, _, _, skipped, _, _ = replay(nodes, leaves, o, d)
    assert skipped.tolist() == [False, True, True]  # |d.d - 1| > 1e-3: the kernel falls back to the scan


def test_replayed_preview_traversal_equals_the_oracle_rasterizer(replay, oracle):
    """the preview renderer's variant of the traversal (ANY_T: no minimum distance, negative distances count -- a sphere behind the
    camera hides what is in front, rasterizer.cpp:41-52) against the oracle's rasterizer.cpp restatement: spheres in front of,
    behind and around the camera, one ray per pixel"""
    from rt_b200.renderer import make_view

    rng = np.random.default_rng(8)
    sph = np.concatenate([rng.uniform(-12, 12, (600, 3)), rng.uniform(0.1, 1.2, (600, 1))], axis=1).astype(np.float32)
    sph[0] = [0, 0, 0, 3.0]        # the camera sits inside this one
    sph[1] = [0.2, 0.1, 6.0, 1.0]  # behind the camera: negative distance
    sph[2] = sph[3]                # identical spheres: the lower index wins
    sc = scene_of(sph)
    sc.camera.position, sc.camera.direction = (0.0, 0.0, 1.0), (0.1, -0.05, -1.0)
    view = make_view(sc, 160, 100)
    _, ref_prim, ref_depth = oracle.rasterize(sc, view)
    px, py = np.meshgrid(np.arange(160), np.arange(100))
    o, d = oracle.primary_rays(view, px.ravel(), py.ravel(), np.zeros(160 * 100, np.uint32))  # sample 0 = the pixel centre
    nodes, leaves, _ = R.bvh4_build_host(sph)
    hit, prim, t, skipped, _, _ = replay(nodes, leaves, np.ascontiguousarray(o), np.ascontiguousarray(d), any_t=True)
    assert not skipped.any()
    ref_prim, ref_depth = ref_prim.ravel(), ref_depth.ravel()
    ref_hit = ref_prim != 0xFFFFFFFF
    assert np.array_equal(hit.astype(bool), ref_hit) and ref_hit.mean() > 0.9
    assert np.array_equal(prim[ref_hit], ref_prim[ref_hit])
    assert np.array_equal(t[ref_hit].view(np.uint32), ref_depth[ref_hit].view(np.uint32))
    assert (t[ref_hit] < 0).any() and len(np.unique(prim[ref_hit])) > 20  # negative distances do occur, and many spheres are seen


@pytest.mark.parametrize("trav_define", [(), ("-DBVH_REPLAY_TRAV2",)], ids=["leaves-in-visit", "deferred-leaves"])
def test_the_rays_need_the_margins(tmp_path_factory, oracle, trav_define):
    """the same traversal with the conservative margins switched off loses grazing hits on these rays: the equality above is the
    margins' doing, not the rays' leniency"""
    no_margin = build_replay(tmp_path_factory, "-DBVH_REPLAY_NO_MARGIN", *trav_define)
    sph = np.ascontiguousarray(synth.rtiow_scene().spheres, np.float32)
    nodes, leaves, _ = R.bvh4_build_host(sph)
    o, d = rays_for(sph, 240000, np.random.default_rng(5))
    hit, prim, _, _, _, _ = no_margin(nodes, leaves, o, d)
    ref_hit, ref_prim, _, _ = oracle.intersect_batch(scene_of(sph), o, d)
    wrong = int((hit != ref_hit).sum() + ((prim != ref_prim) & (hit == 1) & (ref_hit == 1)).sum())
    assert 0 < wrong < 200, wrong
