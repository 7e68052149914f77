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


KNOBS = [{}, {"RTCU_BVH_LEAF_COST": "1"}, {"RTCU_BVH_SWEEP": "0", "RTCU_BVH_LEAF_COST": "0", "RTCU_BVH_COLLAPSE": "greedy"},
         {"RTCU_BVH_SWEEP": "512", "RTCU_BVH_LEAF_COST": "1", "RTCU_BVH_COLLAPSE": "sah"}]


@pytest.mark.parametrize("trav", [2, 1], ids=["deferred-leaves", "leaves-in-visit"])
@pytest.mark.parametrize("knobs", KNOBS, ids=["default", "leafcost", "binned-greedy", "sweep+leafcost+sah-collapse"])
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


# ---- patch beams ----------------------------------------------------------------------------------------------------------------
def build_beam_replay(tmp_path_factory, *defines):
    so = tmp_path_factory.mktemp("bvh_beam") / "libbvh_beam.so"
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-mfma", *defines, "-o", str(so),
                    str(ROOT / "tests" / "tools" / "bvh_replay.c"), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.bvh_replay_beam_collect.restype = C.c_int
    lib.bvh_replay_beam_collect.argtypes = [C.c_void_p] * 4
    lib.bvh_replay_beam_closest.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_uint32] + [C.c_void_p] * 4
    return lib


def beam_pixel_rays(oracle, view, px, py, rng, n_random, pw=8, ph=4):
    """screen positions of the pw x ph pixel patch at (px, py): centre, the four corners (the far ones are the closed end of the
    jitter range), points on the edges, the largest jitter a sample of the last pixel can draw (1 - 2^-24), pixel centres (sample
    0), random positions; and the primary rays through them -- the first five are what k_beam_lists evaluates"""
    top = np.float32(1.0) - np.float32(2.0) ** -24
    w, h = np.float32(pw), np.float32(ph)
    offs = [(w / 2, h / 2), (0, 0), (w, 0), (0, h), (w, h), (w - 1 + top, h - 1 + top), (0, h - 1 + top), (w - 1 + top, 0), (w / 2, 0), (0, h / 2),
            (w - 1 + top, h / 2), (w / 2, h - 1 + top), (0.5, 0.5), (w - 0.5, h - 0.5), (w - 0.5, 0.5), (0.5, h - 0.5)]
    offs += [tuple(x) for x in (rng.random((n_random, 2)) * [pw, ph]).astype(np.float32)]
    sx = np.float32(px) + np.array([a for a, _ in offs], np.float32)
    sy = np.float32(py) + np.array([b for _, b in offs], np.float32)
    return oracle.screen_rays(view, sx, sy)


def beam_cases():
    from rt_b200.scene import Camera

    rtiow, grid = synth.rtiow_scene(), synth.grid_scene(nx=120, nz=80)
    rng = np.random.default_rng(21)
    cloud = np.concatenate([rng.uniform(-8, 8, (900, 3)), rng.uniform(0.05, 0.7, (900, 1))], axis=1).astype(np.float32)
    cloud[0] = [0, -1000, 0, 1000]
    cloud[1] = cloud[2]  # identical spheres: the lower index wins, in a list as in a scan
    # (spheres, camera, image size): fine pixels (the benchmark's 4K), coarse pixels (fat beams), the camera inside / between spheres
    return {
        "rtiow-4k": (rtiow.spheres, rtiow.camera, (3840, 2160)),
        "rtiow-coarse": (rtiow.spheres, rtiow.camera, (160, 90)),
        "rtiow-low": (rtiow.spheres, Camera(position=(3.0, 0.45, 2.0), direction=(-1.0, -0.05, -0.6)), (640, 360)),
        "grid-4k": (grid.spheres, grid.camera, (3840, 2160)),
        "cloud": (cloud, Camera(position=(0.3, 0.2, 0.1), direction=(0.2, -0.1, -1.0)), (320, 200)),
    }


@pytest.mark.parametrize("name", ["rtiow-4k", "rtiow-coarse", "rtiow-low", "grid-4k", "cloud"])
@pytest.mark.parametrize("patch", [(8, 4), (5, 3), (1, 1)], ids=["8x4", "ragged-5x3", "1x1"])
def test_replayed_patch_beams_equal_the_oracle_scan(tmp_path_factory, oracle, name, patch):
    """every primary ray of a patch of pixels (the kernels' 8x4, a ragged patch at a tile edge, a single pixel) -- corners, edges,
    extreme and random jitters -- finds in the patch's candidate list exactly what the reference's scan over ALL spheres finds: hit
    flag, sphere index and t, bit for bit"""
    from rt_b200.renderer import make_view

    lib = build_beam_replay(tmp_path_factory)
    sph, cam, (w, h) = beam_cases()[name]
    sph = np.ascontiguousarray(sph, np.float32)
    sc = scene_of(sph)
    sc.camera = cam
    view = make_view(sc, w, h)
    nodes, leaves, _ = R.bvh4_build_host(sph)
    pw, ph = patch
    rng = np.random.default_rng(sum(map(ord, name)) + pw)
    n_pixels = 200 if len(sph) < 2000 else 100
    px = rng.integers(0, w - pw + 1, n_pixels)
    py = np.concatenate([rng.integers(0, h - ph + 1, n_pixels // 2), rng.integers(h // 3, h - ph + 1, n_pixels - n_pixels // 2)])  # sky is boring: favour the lower rows
    lists, tested, lengths = 0, 0, []
    list_tn, list_leaf = np.zeros(32, np.float32), np.zeros(32, np.uint32)
    for x, y in zip(px, py):
        o, d = beam_pixel_rays(oracle, view, int(x), int(y), rng, 12, pw, ph)
        rays5 = np.ascontiguousarray(np.concatenate([o[:5], d[:5]], axis=1).reshape(5, 2, 3), np.float32)  # centre + corners as (o, d)
        n_list = lib.bvh_replay_beam_collect(nodes.ctypes.data, rays5.ctypes.data, list_tn.ctypes.data, list_leaf.ctypes.data)
        if n_list < 0:
            continue  # this pixel traverses
        lists += 1
        lengths.append(n_list)
        n = len(o)
        hit, prim, t, usable = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        lib.bvh_replay_beam_closest(leaves.ctypes.data, list_tn.ctypes.data, list_leaf.ctypes.data, n_list, o.ctypes.data, d.ctypes.data, n,
                                    hit.ctypes.data, prim.ctypes.data, t.ctypes.data, usable.ctypes.data)
        ref_hit, ref_prim, ref_t, _ = oracle.intersect_batch(sc, o, d)
        ok = usable.astype(bool)
        assert ok.all()  # normalised directions are always within BEAM_EPS_D of unit length
        assert np.array_equal(hit[ok], ref_hit[ok]), (name, int(x), int(y))
        assert np.array_equal(prim[ok], ref_prim[ok]), (name, int(x), int(y))
        assert np.array_equal(t[ok].view(np.uint32), ref_t[ok].view(np.uint32)), (name, int(x), int(y))
        tested += int(ok.sum())
    assert lists >= n_pixels // 5, (lists, n_pixels)  # the lists are used, not just declined
    assert tested > 800 and max(lengths) >= 2


def test_the_beams_need_the_footprint_margins(tmp_path_factory, oracle):
    """with the pixel-footprint terms of the margin switched off (the centre ray's own margins only), coarse pixels lose hits of
    their corner rays: the equality above is the margins' doing"""
    from rt_b200.renderer import make_view

    lib = build_beam_replay(tmp_path_factory, "-DBVH_REPLAY_NO_BEAM_MARGIN")
    sph, cam, _ = beam_cases()["rtiow-coarse"]
    w, h = 64, 36  # pixels a quarter of a small sphere wide: the corner rays see spheres the centre ray passes by
    sph = np.ascontiguousarray(sph, np.float32)
    sc = scene_of(sph)
    sc.camera = cam
    view = make_view(sc, w, h)
    nodes, leaves, _ = R.bvh4_build_host(sph)
    rng = np.random.default_rng(4)
    wrong = 0
    list_tn, list_leaf = np.zeros(32, np.float32), np.zeros(32, np.uint32)
    for x, y in zip(rng.integers(0, w, 500), rng.integers(h // 3, h, 500)):
        o, d = beam_pixel_rays(oracle, view, int(x), int(y), rng, 4, 1, 1)
        rays5 = np.ascontiguousarray(np.concatenate([o[:5], d[:5]], axis=1).reshape(5, 2, 3), np.float32)
        n_list = lib.bvh_replay_beam_collect(nodes.ctypes.data, rays5.ctypes.data, list_tn.ctypes.data, list_leaf.ctypes.data)
        if n_list < 0:
            continue
        n = len(o)
        hit, prim, t, usable = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        lib.bvh_replay_beam_closest(leaves.ctypes.data, list_tn.ctypes.data, list_leaf.ctypes.data, n_list, o.ctypes.data, d.ctypes.data, n,
                                    hit.ctypes.data, prim.ctypes.data, t.ctypes.data, usable.ctypes.data)
        ref_hit, ref_prim, _, _ = oracle.intersect_batch(sc, o, d)
        wrong += int((hit != ref_hit).sum() + ((prim != ref_prim) & (hit == 1) & (ref_hit == 1)).sum())
    assert wrong > 0
