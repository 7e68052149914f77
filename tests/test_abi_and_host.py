"""The C-ABI library loads and exports every symbol include/rtcu.h declares; host-side mirrors of the
reference interface (renderer registry, image_view, views, partitioning) behave like the reference's.
No compute is launched here (these run without a GPU)."""
import ctypes
import pathlib
import re
import subprocess
import sys

import numpy as np
import pytest

from rt_b200 import _native as nat, build, dist, renderer as R, scene as S

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return nat.load_library()


def test_header_symbols_are_all_exported(lib):
    header = (ROOT / "include" / "rtcu.h").read_text()
    declared = set(re.findall(r"\b(rtcu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    nm = subprocess.run(["nm", "-D", "--defined-only", str(nat.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (rtcu_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(nat.Material) == 28
    assert ctypes.sizeof(nat.View) == 64 + 4 * 10 + 8 + 8
    assert nat.View.seed.offset == 104 and nat.View.material_mode.offset == 112
    assert ctypes.sizeof(nat.SceneDesc) == 88 and nat.SceneDesc.boxes.offset == 64  # ABI 2 appended the box columns
    assert ctypes.sizeof(nat.Stats) == 64
    assert S.MATERIAL_DTYPE.itemsize == ctypes.sizeof(nat.Material)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", str(nat.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_threshold(lib):
    assert lib.rtcu_abi_version() == 3
    assert lib.rtcu_bvh_threshold() > 0


def test_null_arguments_are_rejected_without_a_device(lib):
    # argument validation happens before any CUDA call
    assert lib.rtcu_upload_scene(None, None) == nat.RTCU_ERR_INVALID
    assert lib.rtcu_render(None, None, None, None) == nat.RTCU_ERR_INVALID
    assert lib.rtcu_rasterize(None, None, None, None, None) == nat.RTCU_ERR_INVALID
    assert lib.rtcu_get_stats(None, None) == nat.RTCU_ERR_INVALID
    assert b"null" in lib.rtcu_last_error()
    lib.rtcu_destroy(None)  # no-op


def test_no_cpu_fallback_without_a_device(lib):
    if lib.rtcu_device_count() > 0:
        pytest.skip("a device is present; the failure path is covered on CPU-only hosts")
    assert not lib.rtcu_create(0)
    assert b"no CPU fallback" in lib.rtcu_last_error()
    with pytest.raises(nat.RtcuError):
        R.Context(0)


def test_registry_semantics():
    reg = R._Registry()

    class a(R.RendererInterface):
        pass

    class b(R.RendererInterface):
        pass

    reg.install(R.Description("k1", "alpha", a))
    reg.install(R.Description("k2", "beta", b))
    reg.install(R.Description("k1", "alpha2", b))  # same key overwrites in place (renderer.cpp:27-34)
    assert [d.name for d in reg.all()] == ["alpha2", "beta"]
    assert reg.find_by_key("k2").create is b and reg.find_by_key("") is None and reg.find_by_name("nope") is None
    assert reg.find("be").name == "beta"  # prefix match like main.cpp:68-81
    # the plugin registers under its type name
    assert R.renderers.find_by_name("cuda_path_tracer") is not None
    assert R.renderers.find("cuda").create is R.cuda_path_tracer  # first registered wins a shared prefix
    assert R.renderers.find("cuda_r").create is R.cuda_rasterizer


def test_image_view():
    img = R.ImageView.allocate(5, 3)
    assert img.size() == (5, 3) and img.position_of(7) == (2, 1)
    img.clear(0x000000FF)
    assert img(4, 2) == 0xFF
    with pytest.raises(ValueError):
        R.ImageView(np.zeros((3, 5), np.float32))


def test_make_view_defaults_follow_the_scene():
    sc = S.load("scenes/dielectric.toml")
    v = R.make_view(sc, 320, 200)
    assert (v.samples_per_pixel, v.max_bounces) == (200, 10)
    assert (v.sample_begin, v.sample_end) == (0, 200)
    assert (v.tile_x0, v.tile_y0, v.tile_x1, v.tile_y1) == (0, 0, 320, 200)
    assert v.seed == R.DEFAULT_SEED and v.material_mode == nat.MODE_SM
    m = np.array(list(v.inv_view_proj), np.float32).reshape(4, 4).T  # column-major -> matrix
    # the image centre at depth 0 unprojects onto the near plane in front of the camera
    p = m @ np.array([0, 0, 0, 1], np.float32)
    np.testing.assert_allclose(p[:3] / p[3], [0, 1, 7 - 0.01], atol=1e-4)
    p = m @ np.array([0, 0, 1, 1], np.float32)
    np.testing.assert_allclose(p[:3] / p[3], [0, 1, 7 - 1000], rtol=1e-3)


def test_sample_range_partition_is_exact_cover():
    for total, world in ((64, 1), (64, 8), (4096, 8), (7, 3), (3, 8)):
        ranges = [dist.sample_range_for_rank(5, 5 + total, r, world) for r in range(world)]
        assert ranges[0][0] == 5 and ranges[-1][1] == 5 + total
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    sc = S.load("scenes/basic.toml")
    v = R.make_view(sc, 64, 48, samples_per_pixel=64)
    parts = [dist.partition_view(v, r, 4) for r in range(4)]
    assert [(p.sample_begin, p.sample_end) for p in parts] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert all(p.samples_per_pixel == 64 and p.seed == v.seed for p in parts) and (v.sample_begin, v.sample_end) == (0, 64)
    rows = [dist.partition_view(v, r, 4, by="rows") for r in range(4)]
    assert [(p.tile_y0, p.tile_y1) for p in rows] == [(0, 12), (12, 24), (24, 36), (36, 48)]


def test_scene_fingerprint_detects_changes():
    a = S.load("scenes/basic.toml")
    b = S.load("scenes/basic.toml")
    assert R.scene_fingerprint(a) == R.scene_fingerprint(b)
    b.spheres[1, 0] += 0.25
    assert R.scene_fingerprint(a) != R.scene_fingerprint(b)


# ---- host BVH builder (no device): every sphere in exactly one leaf, child boxes enclose their subtrees ------------
@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 9, 484, 5000])
def test_bvh_builder_invariants(lib, n):
    from rt_b200 import synth

    if n == 484:
        sph = synth.rtiow_scene().spheres
    else:
        rng = np.random.default_rng(n)
        sph = np.concatenate([rng.uniform(-20, 20, (n, 3)), rng.uniform(0.05, 2.0, (n, 1))], axis=1).astype(np.float32)
        if n >= 5:
            sph[0] = [0, -1000, 0, 1000]   # a huge ground sphere like the benchmark scenes
            sph[1] = sph[2]                # duplicates (identical centroids)
    n = len(sph)
    nodes, order, depth = R.bvh_build_host(sph)
    assert sorted(order.tolist()) == list(range(n))
    assert depth <= 62 and len(nodes) >= 1
    seen = np.zeros(n, bool)

    def check(node_index):
        nd = nodes[node_index]
        lo_hi = []
        for side in range(2):
            lo = np.array([nd["x"][2 * side], nd["y"][2 * side], nd["z"][2 * side]])
            hi = np.array([nd["x"][2 * side + 1], nd["y"][2 * side + 1], nd["z"][2 * side + 1]])
            child, count = int(nd["child"][side]), int(nd["count"][side])
            if child < 0:
                first = ~child
                assert count <= 4
                prims = order[first:first + count]
                assert list(prims) == sorted(prims)  # ascending original index inside a leaf
                for p in prims:
                    assert not seen[p]
                    seen[p] = True
                    assert (lo <= sph[p, :3] - sph[p, 3]).all() and (hi >= sph[p, :3] + sph[p, 3]).all()
                sub = (lo, hi) if count else None
            else:
                sub = check(child)
                assert (lo <= sub[0]).all() and (hi >= sub[1]).all()
                sub = (lo, hi)
            if sub is not None:
                lo_hi.append(sub)
        if not lo_hi:
            return np.full(3, np.inf), np.full(3, -np.inf)
        los = np.min([b[0] for b in lo_hi], axis=0)
        his = np.max([b[1] for b in lo_hi], axis=0)
        return los, his

    import sys
    sys.setrecursionlimit(10000)
    check(0)
    assert seen.all()


# ---- the 4-wide device tree (collapsed on upload), built on the host: same invariants in the device layout ---------
@pytest.mark.parametrize("knobs", [{}, {"RTCU_BVH_SWEEP": "512"}, {"RTCU_BVH_LEAF_COST": "1"}, {"RTCU_BVH_SWEEP": "64", "RTCU_BVH_LEAF_COST": "1"},
                                   {"RTCU_BVH_COLLAPSE": "sah"}, {"RTCU_BVH_COLLAPSE": "sah", "RTCU_BVH_LEAF_COST": "1"},
                                   {"RTCU_BVH_SWEEP": "0", "RTCU_BVH_LEAF_COST": "0", "RTCU_BVH_COLLAPSE": "greedy"},  # round 1's tree
                                   {"RTCU_BVH_SWEEP": "0", "RTCU_BVH_LEAF_COST": "0"}, {"RTCU_BVH_COLLAPSE": "greedy"}],
                         ids=["default", "sweep", "leafcost", "sweep+leafcost", "sah-collapse", "sah-collapse+leafcost", "binned-greedy", "binned", "greedy"])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 9, 33, 484, 5000, 100001])
def test_bvh4_device_tree_invariants(lib, n, knobs, monkeypatch):
    """every tree the builder can produce -- the default and the experimental split / collapse rules (bvh.h: RTCU_BVH_SWEEP,
    RTCU_BVH_LEAF_COST; rtcu.cu pack_bvh4: RTCU_BVH_COLLAPSE=sah) -- is a valid input for the traversal: coverage, containment,
    leaf packing, stack bound"""
    from rt_b200 import synth

    if knobs and n in (1, 3, 4, 9, 5000):
        pytest.skip("knob variants: a subset of sizes")
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)

    if n == 484:
        sph = synth.rtiow_scene().spheres
    elif n == 100001:
        sph = synth.grid_scene().spheres
    else:
        rng = np.random.default_rng(100 + n)
        sph = np.concatenate([rng.uniform(-20, 20, (n, 3)), rng.uniform(0.05, 2.0, (n, 1))], axis=1).astype(np.float32)
        if n >= 5:
            sph[0] = [0, -1000, 0, 1000]
            sph[1] = sph[2]
    n = len(sph)
    nodes, leaves, depth = R.bvh4_build_host(sph)
    assert len(nodes) >= 1 and 3 * depth + 2 <= 64  # the traversal stack (BVH_STACK) holds 3 entries per level
    refs = nodes[:, 6, :].copy().view(np.uint32)    # four child references per node
    H = nodes[:, 7, :]
    lo_s = (sph[:, :3].astype(np.float64) - sph[:, 3:4]).astype(np.float64)
    hi_s = (sph[:, :3].astype(np.float64) + sph[:, 3:4]).astype(np.float64)
    seen = np.zeros(n, np.int64)
    visited = np.zeros(len(nodes), bool)
    max_depth = 0
    stack = [(0, 1)]
    # returns nothing: containment is checked child by child against the spheres below it, gathered iteratively
    below = {}  # node -> (lo, hi) of everything below it, filled in post-order

    order = []
    while stack:
        node, d = stack.pop()
        assert not visited[node]
        visited[node] = True
        max_depth = max(max_depth, d)
        order.append(node)
        for c in range(4):
            if not (refs[node, c] & 0x80000000) and np.isfinite(H[node, c]):
                stack.append((int(refs[node, c]), d + 1))
    assert visited.all() and max_depth == depth
    for node in reversed(order):
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        for c in range(4):
            pair, slot = c // 2, c % 2
            cen = nodes[node, 3 * pair:3 * pair + 3, slot].astype(np.float64)
            half = nodes[node, 3 * pair:3 * pair + 3, 2 + slot].astype(np.float64)
            if not np.isfinite(H[node, c]):
                assert (half == -np.inf).all() and H[node, c] == -np.inf  # empty slot: can never be hit
                continue
            assert H[node, c] >= half.sum() * (1 - 1e-7)
            if refs[node, c] & 0x80000000:
                blk = leaves[int(refs[node, c] & 0x7FFFFFFF)]
                idx = blk[4].view(np.uint32)
                real = idx[idx != 0x7FFFFFFF]
                assert 1 <= len(real) <= 4 and list(real) == sorted(real)
                assert (idx[:len(real)] != 0x7FFFFFFF).all()  # padding comes last: a padded third slot means an empty second pair (leaf_second_pair)
                # the packed pairs hold exactly those spheres: {cx0,cx1,cy0,cy1},{cz0,cz1,r2_0,r2_1}, padding r2 = -inf
                for k, i in enumerate(idx):
                    a, b = blk[2 * (k // 2)], blk[2 * (k // 2) + 1]
                    got = np.array([a[k % 2], a[2 + k % 2], b[k % 2], b[2 + k % 2]], np.float32)
                    if i == 0x7FFFFFFF:
                        assert got[3] == -np.inf
                    else:
                        assert (got[:3] == sph[i, :3]).all() and got[3] == np.float32(sph[i, 3]) * np.float32(sph[i, 3])
                seen[real] += 1
                sub_lo, sub_hi = lo_s[real].min(axis=0), hi_s[real].max(axis=0)
            else:
                sub_lo, sub_hi = below[int(refs[node, c])]
            assert (cen - half <= sub_lo).all() and (cen + half >= sub_hi).all()  # [c - h, c + h] encloses the subtree
            lo_all, hi_all = np.minimum(lo_all, sub_lo), np.maximum(hi_all, sub_hi)
        below[node] = (lo_all, hi_all)
    assert (seen == 1).all()  # every sphere in exactly one leaf


# ---- the builder runs on worker threads: the tree must not depend on how many -------------------------------------------
@pytest.mark.parametrize("case", ["grid100k", "rtiow", "random20k", "identical10k", "line9k", "clusters30k"])
def test_bvh_build_is_independent_of_thread_count(lib, case, monkeypatch):
    from rt_b200 import synth

    rng = np.random.default_rng(7)
    if case == "grid100k":
        sph = synth.grid_scene().spheres
    elif case == "rtiow":
        sph = synth.rtiow_scene().spheres  # below the builder's threading threshold: the serial path under every setting
    elif case == "random20k":
        sph = np.concatenate([rng.uniform(-50, 50, (20000, 3)), rng.uniform(0.01, 3.0, (20000, 1))], axis=1).astype(np.float32)
        sph[0] = [0, -1000, 0, 1000]
    elif case == "identical10k":
        sph = np.tile(np.array([[1.0, 2.0, 3.0, 0.5]], np.float32), (10000, 1))  # all centroids coincide: index-order splits
    elif case == "line9k":
        sph = np.zeros((9000, 4), np.float32)  # one axis only, many equal centroids
        sph[:, 0] = rng.integers(0, 40, 9000)
        sph[:, 3] = 0.25
    else:
        centres = rng.uniform(-200, 200, (30, 3))
        sph = np.concatenate([centres[rng.integers(0, 30, 30000)] + rng.normal(0, 0.5, (30000, 3)), rng.uniform(0.01, 0.2, (30000, 1))],
                             axis=1).astype(np.float32)
    if case in ("grid100k", "clusters30k"):  # the other split / collapse rules are order-independent as well
        monkeypatch.setenv("RTCU_BVH_SWEEP", "512" if case == "grid100k" else "0")
        monkeypatch.setenv("RTCU_BVH_LEAF_COST", "0" if case == "grid100k" else "1")
        monkeypatch.setenv("RTCU_BVH_COLLAPSE", "greedy")
        check_threads(sph, monkeypatch, ("8",))
        monkeypatch.delenv("RTCU_BVH_SWEEP")
        monkeypatch.delenv("RTCU_BVH_LEAF_COST")
        monkeypatch.delenv("RTCU_BVH_COLLAPSE")
    check_threads(sph, monkeypatch, ("2", "3", "8", "16"))


def check_threads(sph, monkeypatch, counts):
    monkeypatch.setenv("RTCU_BVH_THREADS", "1")
    ref_nodes, ref_leaves, ref_depth = R.bvh4_build_host(sph)
    for threads in counts:
        monkeypatch.setenv("RTCU_BVH_THREADS", threads)
        nodes, leaves, depth = R.bvh4_build_host(sph)
        assert depth == ref_depth
        assert np.array_equal(nodes.view(np.uint32), ref_nodes.view(np.uint32)), threads
        assert np.array_equal(leaves.view(np.uint32), ref_leaves.view(np.uint32)), threads


NOMEM_SCRIPT = r"""
import ctypes as C, resource, sys
import numpy as np
from rt_b200 import _native as nat
lib = nat.load_library()
n = 3_000_000
rng = np.random.default_rng(1)
sph = np.concatenate([rng.uniform(-50, 50, (n, 3)), rng.uniform(0.01, 0.2, (n, 1))], axis=1).astype(np.float32)
vm = int([l for l in open('/proc/self/status') if l.startswith('VmSize')][0].split()[1]) * 1024
resource.setrlimit(resource.RLIMIT_AS, (vm + (48 << 20), vm + (48 << 20)))  # the build needs ~250 MB more
a, b, d = C.c_uint32(), C.c_uint32(), C.c_uint32()
print(lib.rtcu_bvh4_build_host(nat.ptr(sph), n, None, 0, None, 0, C.byref(a), C.byref(b), C.byref(d)), nat.last_error())
print(lib.rtcu_bvh_build_host(nat.ptr(sph), n, None, None, 0, C.byref(a), C.byref(d)), nat.last_error())
"""


def test_a_failed_host_allocation_is_an_error_code_not_an_exception(lib):
    """C++ exceptions never cross the C ABI: under an address-space limit the builder (its vectors, its worker threads' stacks)
    runs out of memory and the entry point returns RTCU_ERR_NOMEM"""
    r = subprocess.run([sys.executable, "-c", NOMEM_SCRIPT], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = r.stdout.strip().splitlines()
    assert lines == [f"{nat.RTCU_ERR_NOMEM} rtcu_bvh4_build_host: out of host memory", f"{nat.RTCU_ERR_NOMEM} rtcu_bvh_build_host: out of host memory"]


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: the product (rt_b200/, include/, plugin/, tools/) never imports, links or names it;
    only tests/ (incl. tests/tools), __graft_entry__.smoke() and bench.py's CPU legs load it.  rt_b200/build.py may *build* it."""
    allowed_py = {ROOT / "bench.py", ROOT / "__graft_entry__.py"}
    offenders = []
    for p in ROOT.rglob("*"):
        rel = p.relative_to(ROOT)
        if not p.is_file() or rel.parts[0] in ("tests", "oracle", "gpurun_out", ".git", "profiles") or p.suffix not in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c"):
            continue
        text = p.read_text(errors="replace")
        uses = re.search(r"from oracle|import oracle|oracle\.binding|rtref_|librtref|librt_ref", text) is not None
        if uses and p not in allowed_py and rel != pathlib.Path("rt_b200/build.py"):
            offenders.append(str(rel))
    assert offenders == [], offenders
    build_py = (ROOT / "rt_b200" / "build.py").read_text()
    assert "oracle.binding" not in build_py and "CDLL" not in build_py  # it only runs `make -C oracle`


# ---- checkpoint / resume of a progressive render (host logic; the oracle stands in for the device) ------------------------
class _OracleContext:
    """what ProgressiveRenderer needs from a Context, answered by the CPU oracle: the same (pixel, sample)-indexed sums"""

    def __init__(self, oracle):
        self.oracle, self.scene = oracle, None

    def upload_scene(self, scene):
        self.scene = scene

    def render(self, view, want_rgba8=True, want_accum=False, **_):
        rgba8, accum, _ = self.oracle.render(self.scene, view, want_rgba8=want_rgba8, want_accum=want_accum)
        return rgba8, accum


def test_progressive_render_checkpoint_and_resume(oracle, tmp_path):
    from rt_b200.renderer import ProgressiveRenderer

    sc = S.load("scenes/dielectric.toml")
    whole = ProgressiveRenderer(_OracleContext(oracle), sc, 48, 27, samples_per_step=3, max_bounces=12)
    for _ in range(4):
        final = whole.refine()

    first = ProgressiveRenderer(_OracleContext(oracle), sc, 48, 27, samples_per_step=3, max_bounces=12)
    first.refine()
    first.refine()
    ckpt = tmp_path / "render.npz"
    first.save(ckpt)
    first.refine()           # work after the checkpoint is lost with the process ...
    del first

    second = ProgressiveRenderer(_OracleContext(oracle), sc, 48, 27, samples_per_step=3, max_bounces=12)
    assert second.restore(ckpt) == 6
    second.refine()
    resumed = second.refine()  # ... and redone: the samples are global indices, so the sums come out the same, bit for bit
    assert second.samples_done == whole.samples_done == 12
    assert np.array_equal(second.accum.view(np.uint32), whole.accum.view(np.uint32)) and np.array_equal(resumed, final)
    assert [p.name for p in tmp_path.iterdir()] == ["render.npz"]  # no temporary file left behind

    # a checkpoint is only good for the render it came from
    for kwargs, what in (({"seed": 7}, "seed"), ({"material_mode": nat.MODE_MG}, "material_mode"), ({"max_bounces": 13}, "max_bounces")):
        other = ProgressiveRenderer(_OracleContext(oracle), sc, 48, 27, samples_per_step=3, **{"max_bounces": 12, **kwargs})
        with pytest.raises(ValueError, match=what):
            other.restore(ckpt)
    moved = S.load("scenes/dielectric.toml")
    moved.camera.position = (0.0, 2.0, 5.0)
    with pytest.raises(ValueError, match="camera"):
        ProgressiveRenderer(_OracleContext(oracle), moved, 48, 27, samples_per_step=3, max_bounces=12).restore(ckpt)
    edited = S.load("scenes/dielectric.toml")
    edited.spheres[0, 3] *= 2
    with pytest.raises(ValueError, match="scene"):
        ProgressiveRenderer(_OracleContext(oracle), edited, 48, 27, samples_per_step=3, max_bounces=12).restore(ckpt)
    with pytest.raises(ValueError, match="width"):
        ProgressiveRenderer(_OracleContext(oracle), sc, 64, 27, samples_per_step=3, max_bounces=12).restore(ckpt)


def test_worker_pool_sleep_and_wake_protocol_under_stress(tmp_path):
    """bvh.h's pool with its spin shortened to a few iterations, so nearly every hand-over goes through the condition variable:
    160 builds at 1 - 11 workers must terminate (a lost wake-up would hang) and equal the one-thread tree"""
    exe = tmp_path / "pool_stress"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-DRTCU_BVH_POOL_SPIN=4", "-o", str(exe), str(ROOT / "tests" / "tools" / "pool_stress.cpp")], check=True)
    r = subprocess.run([str(exe), "20"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok 160 builds", r.stdout + r.stderr
