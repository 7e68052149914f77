"""Round-2 parity cases on the device: boxes in a path-traced scene (a-5), scenes beyond the BVH's safe coordinate range, frames
of one context launched on two streams, the page-locked destination (incl. a buffer re-mapped at the same address), the multi-
device scene upload, and the plugin as it lives in the application (one renderer instance across frames, all GPUs of the box)."""
import ctypes as C
import dataclasses
import mmap

import numpy as np
import pytest

from rt_b200 import _native as nat, scene as S, synth
from rt_b200.renderer import Context, make_view, render_multi, upload_scene_multi

from conftest import unpack_rgba

pytestmark = pytest.mark.gpu


def _device_count():
    return nat.load_library().rtcu_device_count()


# ---- a-5: test_boxes is a stub (mg_ray_tracer.cpp:89-93) -- boxes load, upload, and never change a path-traced image ------------
@pytest.mark.parametrize("mode", [nat.MODE_MG, nat.MODE_SM])
def test_boxes_scene_path_traced_against_the_oracle(ctx, oracle, mode):
    sc = S.load("scenes/boxes.toml")
    assert len(sc.boxes) >= 3 and len(sc.planes) >= 2 and len(sc.spheres) >= 4
    ctx.upload_scene(sc)
    v = make_view(sc, 320, 200, samples_per_pixel=8, max_bounces=8, material_mode=mode)
    rgba8, accum = ctx.render(v, want_accum=True)
    r_rgba8, r_accum, r_segs = oracle.render(sc, v)
    assert ctx.stats()["segments"] == r_segs  # the same paths, segment for segment: no box was ever hit
    np.testing.assert_allclose(accum[..., :3], r_accum[..., :3], rtol=2e-5, atol=1e-6)
    assert np.abs(unpack_rgba(rgba8) - unpack_rgba(r_rgba8)).max() <= 1


def test_appended_boxes_do_not_change_the_frame(ctx, scenes):
    sc = scenes["c2"][0]
    kw = dict(samples_per_pixel=16, max_bounces=50, material_mode=nat.MODE_SM)
    ctx.upload_scene(sc)
    rgba8, accum = ctx.render(make_view(sc, 400, 225, **kw), want_accum=True)
    segs = ctx.stats()["segments"]
    # boxes right in front of the camera, around spheres, enclosing the whole scene
    boxes = np.array([[0, 1, 2, 0.5, 0.5, 0.5], [0, 0.5, 0, 1, 1, 1], [0, 0, 0, 50, 50, 50], [-1, 0.5, -1, 0.2, 2, 0.2]], np.float32)
    with_boxes = dataclasses.replace(sc, boxes=boxes, box_material=np.zeros(len(boxes), np.uint32))
    ctx.upload_scene(with_boxes)
    rgba8_b, accum_b = ctx.render(make_view(with_boxes, 400, 225, **kw), want_accum=True)
    assert ctx.stats()["segments"] == segs
    np.testing.assert_array_equal(accum_b, accum)
    np.testing.assert_array_equal(rgba8_b, rgba8)


# ---- coordinates beyond the range in which the traversal equals the scan: the scan is kept ---------------------------------------
def test_huge_coordinates_keep_the_reference_scan(ctx, oracle):
    rng = np.random.default_rng(5)
    n = 48  # above the BVH threshold
    sph = np.zeros((n, 4), np.float32)
    sph[:, :3] = rng.uniform(-3, 3, (n, 3))
    sph[:, 3] = rng.uniform(0.2, 0.6, n)
    sph[7, :3] = [6e18, 0.5, -8e18]   # beyond 2^62: one more order of magnitude and |e|^2 overflows
    sph[19] = [0, -9e18, 0, 9e18]     # a "ground" whose r^2 is within a factor four of FLT_MAX
    base = synth.rtiow_scene()
    sc = dataclasses.replace(base, spheres=sph, sphere_material=(np.arange(n) % len(base.materials)).astype(np.uint32))
    ctx.upload_scene(sc)
    v = make_view(sc, 96, 64, samples_per_pixel=4, max_bounces=6, material_mode=nat.MODE_SM)
    rgba8, accum = ctx.render(v, want_accum=True)
    st = ctx.stats()
    assert st["accel"] == nat.ACCEL_LINEAR  # no BVH for this scene, although it has more than rtcu_bvh_threshold() spheres
    r_rgba8, r_accum, r_segs = oracle.render(sc, v)
    assert st["segments"] == r_segs
    np.testing.assert_allclose(accum[..., :3], r_accum[..., :3], rtol=2e-5, atol=1e-6)
    assert np.abs(unpack_rgba(rgba8) - unpack_rgba(r_rgba8)).max() <= 1
    # where S4 does overflow (inf - inf = NaN distances, which the scan's `best <= t` rule accepts): still the scan's answer
    sph[7, :3] = [1e20, 0.5, -2e20]
    sph[19, 3] = 3e19
    sc = dataclasses.replace(sc, spheres=sph)
    ctx.upload_scene(sc)
    o, d = synth.random_rays(sc, 4096, seed=11)
    o[::5] *= np.float32(1e19)
    for accel in (nat.ACCEL_AUTO, nat.ACCEL_LINEAR):
        got = ctx.intersect_batch(o, d, accel=accel)
        want = oracle.intersect_batch(sc, o, d)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    with pytest.raises(nat.RtcuError, match="no BVH"):
        ctx.intersect_batch(o, d, accel=nat.ACCEL_BVH)


def test_far_origins_fall_back_to_the_scan_inside_a_bvh_scene(ctx, oracle, scenes):
    sc = scenes["c3"][0]
    ctx.upload_scene(sc)
    o, d = synth.random_rays(sc, 8192, seed=12)
    o[::3] *= np.float32(3e19)
    o[1::7, 0] = np.float32(np.nan)
    for accel in (nat.ACCEL_BVH, nat.ACCEL_LINEAR):
        hit, prim, t, _ = ctx.intersect_batch(o, d, accel=accel)
        r_hit, r_prim, r_t, _ = oracle.intersect_batch(sc, o, d)
        np.testing.assert_array_equal(hit, r_hit)
        np.testing.assert_array_equal(prim, r_prim)
        np.testing.assert_array_equal(np.isnan(t), np.isnan(r_t))  # (a NaN's payload is the hardware's business)
        ok = ~np.isnan(r_t)
        np.testing.assert_array_equal(t.view(np.uint32)[ok], r_t.view(np.uint32)[ok])


# ---- one frame in flight per context, whatever the streams ---------------------------------------------------------------------
def test_frames_of_one_context_on_two_streams_do_not_interfere(ctx, scenes):
    import torch

    sc = scenes["c3"][0]
    ctx.upload_scene(sc)
    dev = torch.device("cuda", 0)
    w, h = 320, 180
    views = [make_view(sc, w, h, samples_per_pixel=spp, max_bounces=50, material_mode=nat.MODE_SM, flags=flags, seed=seed)
             for spp, flags, seed in ((8, nat.ACCEL_BVH, 1), (24, nat.ACCEL_BVH, 2), (6, nat.ACCEL_LINEAR, 3))]
    want = []
    for v in views:  # one at a time
        buf = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        ctx.render_device(v, buf.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        want.append((buf.cpu().numpy(), ctx.stats()["segments"]))
    streams = [torch.cuda.Stream(dev) for _ in views]
    bufs = [torch.zeros((h, w, 4), dtype=torch.float32, device=dev) for _ in views]
    torch.cuda.synchronize(dev)
    for v, s, b in zip(views, streams, bufs):  # back to back on three streams, no synchronisation in between
        ctx.render_device(v, b.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize(dev)
    for b, (w_accum, _) in zip(bufs, want):
        np.testing.assert_array_equal(b.cpu().numpy(), w_accum)
    assert ctx.stats()["segments"] == want[-1][1]


# ---- the caller's pageable image: page-locked once, verified every frame -------------------------------------------------------
def _anonymous_pages(nbytes):
    libc = C.CDLL(None, use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
    return libc


@pytest.mark.parametrize("name,spp", [("c2", 8), ("c3", 16)])  # zero-copy stores (scan kernels) and the DMA path (direct-mode BVH)
def test_pageable_destination_is_registered_once_and_survives_a_remap(scenes, name, spp):
    sc = scenes[name][0]
    w, h = 512, 256
    nbytes = w * h * 4
    libc = _anonymous_pages(nbytes)
    prot, flags_ = mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS
    addr = libc.mmap(None, nbytes, prot, flags_, -1, 0)
    assert addr not in (None, C.c_void_p(-1).value)
    with Context(0) as c:  # its own context: registrations are per context
        c.set_output_pinning(True)  # what the plugin does: the image is handed over frame after frame
        c.upload_scene(sc)
        v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=20, material_mode=nat.MODE_SM)
        want, _ = c.render(v, rgba8=np.zeros((h, w), np.uint32))  # a throw-away pageable array: staged or registered, either way right
        img = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint32)), shape=(h, w))
        img[...] = 0x000000FF  # main.cpp:318 clears to black first
        c.render(v, rgba8=img)
        np.testing.assert_array_equal(img, want)
        first = c.stats()["ms_d2h"]
        c.render(v, rgba8=img)
        np.testing.assert_array_equal(img, want)
        later = c.stats()["ms_d2h"]
        if name == "c2":
            assert later < 0.05, (first, later)  # the kernels store into the image themselves: nothing follows them
        # the application frees its image and the allocator maps NEW pages at the same address: the registration is stale
        fixed = libc.mmap(addr, nbytes, prot, flags_ | 0x10, -1, 0)  # MAP_FIXED (Linux): replaces the mapping in place
        assert fixed == addr
        img = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint32)), shape=(h, w))
        assert not img.any()
        c.render(v, rgba8=img)
        np.testing.assert_array_equal(img, want)  # detected, delivered by the staged copy
        c.render(v, rgba8=img)
        np.testing.assert_array_equal(img, want)
        # a different size at the same address (window resize) registers anew
        half = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint32)), shape=(h // 2, w))
        v2 = make_view(sc, w, h // 2, samples_per_pixel=spp, max_bounces=20, material_mode=nat.MODE_SM)
        want2, _ = c.render(v2, rgba8=np.zeros((h // 2, w), np.uint32))
        c.render(v2, rgba8=half)
        np.testing.assert_array_equal(half, want2)
    libc.munmap(addr, nbytes)


# ---- single-process multi-GPU: one scene build for all devices -----------------------------------------------------------------
@pytest.mark.parametrize("ngpu", [2, 8])
def test_upload_scene_multi_equals_per_device_uploads(scenes, ngpu):
    if _device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, this box has {_device_count()}")
    sc = scenes["c3"][0]
    view = make_view(sc, 480, 270, samples_per_pixel=16 * ngpu, max_bounces=50, material_mode=nat.MODE_SM)
    ctxs = [Context(g) for g in range(ngpu)]
    try:
        for c in ctxs:
            c.upload_scene(sc)
        want, want_accum = render_multi(ctxs, view, want_accum=True)
        segs = ctxs[0].stats()["segments"]
        upload_scene_multi(ctxs, scenes["c2"][0])  # something else in between
        upload_scene_multi(ctxs, sc)
        got, got_accum = render_multi(ctxs, view, want_accum=True)
        np.testing.assert_array_equal(got_accum, want_accum)
        np.testing.assert_array_equal(got, want)
        st = ctxs[0].stats()
        assert st["segments"] == segs and st["accel"] == nat.ACCEL_BVH
        assert st["node_visits"] > 0 and st["sphere_tests"] != segs * len(sc.spheres)  # counted per device, not segments x spheres
    finally:
        for c in ctxs:
            c.close()


# ---- the plugin as the application holds it: one instance across frames, every GPU of the box ---------------------------------
def _plugin():
    from oracle.binding import ReferenceBuild

    if not (ReferenceBuild.PLUGIN_PATH.exists() or ReferenceBuild.available()):
        pytest.skip("oracle/_ref/librt_ref_plugin.so not present and no reference tree to build it from")
    return ReferenceBuild("plugin")


def test_plugin_instance_across_frames_writes_the_applications_image_directly(monkeypatch):
    monkeypatch.setenv("RT_CUDA_DEVICES", "1")
    monkeypatch.delenv("RT_CUDA_MATERIAL_MODE", raising=False)
    ref = _plugin()
    sc = S.load("scenes/dielectric.toml")
    w, h, spp, depth = 640, 360, 16, 50
    cpu, _ = ref.render(sc, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
    handle = ref.open("cuda_path_tracer")
    try:
        from bench import aligned_pageable  # the reference's image: pageable, 64-byte aligned (image.cpp:9-13)

        img = aligned_pageable(h, w)
        d2h = []
        for frame in range(3):
            img[...] = 0x000000FF
            ref.render_with(handle, sc, img, spp, depth, 0x5EED)
            d = np.abs(unpack_rgba(img) - unpack_rgba(cpu))
            assert d.max() <= 1 and (d > 0).mean() <= 0.005
            d2h.append(ref.plugin_last_stats()["ms_d2h"])
        assert d2h[1] < 0.05 and d2h[2] < 0.05, d2h  # from the second frame on nothing follows the kernels
        # a camera move and a scene edit between frames (main.cpp:271-299, :123-125): detected by content, not by pointer
        moved = dataclasses.replace(sc, camera=dataclasses.replace(sc.camera, position=(0.5, 1.2, 3.5)))
        cpu_moved, _ = ref.render(moved, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
        ref.render_with(handle, moved, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(cpu_moved)).max() <= 1
        edited = dataclasses.replace(moved, spheres=moved.spheres * np.float32([1, 1, 1, 0.8]))
        cpu_edited, _ = ref.render(edited, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
        ref.render_with(handle, edited, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(cpu_edited)).max() <= 1
    finally:
        ref.close(handle)


@pytest.mark.parametrize("ngpu", [2, 8])
def test_plugin_splits_the_frame_over_all_gpus(monkeypatch, ngpu):
    if _device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, this box has {_device_count()}")
    monkeypatch.setenv("RT_CUDA_DEVICES", str(ngpu))
    monkeypatch.delenv("RT_CUDA_MATERIAL_MODE", raising=False)
    ref = _plugin()
    sc = S.load("scenes/dielectric.toml")
    w, h, spp, depth = 320, 200, 16 * ngpu, 50
    cpu, _ = ref.render(sc, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
    handle = ref.open("cuda_path_tracer")
    try:
        img = np.zeros((h, w), np.uint32)
        for _ in range(2):
            ref.render_with(handle, sc, img, spp, depth, 0x5EED)
            d = np.abs(unpack_rgba(img) - unpack_rgba(cpu))
            assert d.max() <= 1 and (d > 0).mean() <= 0.005
            st = ref.plugin_last_stats()
            assert st["kernel_launches"] >= ngpu + 1  # a trace kernel per device + the fused reduce / resolve
        # too few samples to be worth splitting: device 0 alone, same image rules
        ref.render_with(handle, sc, img, 8, depth, 0x5EED)
        cpu8, _ = ref.render(sc, w, h, 8, depth, 0x5EED, "sm_ray_tracer", threads=0)
        assert np.abs(unpack_rgba(img) - unpack_rgba(cpu8)).max() <= 1
        assert ref.plugin_last_stats()["kernel_launches"] <= 3
    finally:
        ref.close(handle)


def test_plugin_refines_an_unchanged_view_when_asked_to(monkeypatch):
    """RT_CUDA_PROGRESSIVE=1 (f-3): frames of an unchanged scene and camera add their samples to the sums on the device; any change
    starts over.  k frames of n samples equal one frame of k n samples (global sample indices), within fp32 summation order."""
    monkeypatch.setenv("RT_CUDA_DEVICES", "1")
    monkeypatch.setenv("RT_CUDA_PROGRESSIVE", "1")
    monkeypatch.delenv("RT_CUDA_MATERIAL_MODE", raising=False)
    ref = _plugin()
    sc = S.load("scenes/dielectric.toml")
    w, h, spp, depth = 320, 200, 8, 50
    one, _ = ref.render(sc, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
    three, _ = ref.render(sc, w, h, 3 * spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
    handle = ref.open("cuda_path_tracer")
    try:
        img = np.zeros((h, w), np.uint32)
        ref.render_with(handle, sc, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(one)).max() <= 1      # frame 1: samples [0, 8)
        ref.render_with(handle, sc, img, spp, depth, 0x5EED)
        ref.render_with(handle, sc, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(three)).max() <= 1    # frame 3: the mean over samples [0, 24)
        moved = dataclasses.replace(sc, camera=dataclasses.replace(sc.camera, position=(0.5, 1.2, 3.5)))
        moved_one, _ = ref.render(moved, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
        ref.render_with(handle, moved, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(moved_one)).max() <= 1  # the camera moved: samples [0, 8) of the new view
        edited = dataclasses.replace(moved, spheres=moved.spheres * np.float32([1, 1, 1, 0.8]))
        edited_one, _ = ref.render(edited, w, h, spp, depth, 0x5EED, "sm_ray_tracer", threads=0)
        ref.render_with(handle, edited, img, spp, depth, 0x5EED)
        assert np.abs(unpack_rgba(img) - unpack_rgba(edited_one)).max() <= 1  # the scene changed: start over too
    finally:
        ref.close(handle)


def test_accumulate_flag_needs_sums_of_the_same_size(ctx, scenes):
    sc = scenes["c2"][0]
    ctx.upload_scene(sc)
    kw = dict(samples_per_pixel=8, max_bounces=20, material_mode=nat.MODE_SM)
    whole, whole_accum = ctx.render(make_view(sc, 200, 120, **kw), want_accum=True)
    ctx.render(make_view(sc, 200, 120, **{**kw, "samples_per_pixel": 3}, sample_range=(0, 3)))
    img, accum = ctx.render(make_view(sc, 200, 120, sample_range=(3, 8), flags=nat.FLAG_ACCUMULATE, **kw), want_accum=True)
    np.testing.assert_allclose(accum, whole_accum, rtol=2e-6, atol=1e-7)   # [0,3) + [3,8) == [0,8)
    assert np.abs(unpack_rgba(img) - unpack_rgba(whole)).max() <= 1
    np.testing.assert_array_equal(ctx.accum_download(200, 120), accum)
    with pytest.raises(nat.RtcuError, match="holds no sums"):
        ctx.render(make_view(sc, 160, 120, flags=nat.FLAG_ACCUMULATE, **kw))
    ctx.accum_upload(np.zeros((120, 200, 4), np.float32))
    img0, accum0 = ctx.render(make_view(sc, 200, 120, flags=nat.FLAG_ACCUMULATE, **kw), want_accum=True)
    np.testing.assert_array_equal(accum0, whole_accum)                      # onto zeros: the plain frame


@pytest.mark.gpu
@pytest.mark.parametrize("run", ["4", "8"])
def test_runs_kernel_traces_the_same_paths_as_direct_mode(ctx, oracle, scenes, knobs, run):
    """k_render_runs (RTCU_BVH_RUNS=1, an experiment that is not the default): lane groups walk runs of pixels instead of one
    pixel at a time.  Same paths (equal segment counts), per-pixel sums equal up to fp32 order, deterministic, ragged patches /
    sub-patch tiles / sample ranges included, and the oracle agrees."""
    sc, depth = scenes["c3"]
    ctx.upload_scene(sc)
    w, h = 203, 117  # ragged against the 8x4 patches
    for spp, rng in ((20, None), (96, None), (64, (7, 40))):  # 8 lanes per pixel, 16 lanes per pixel, a sample range
        kw = dict(samples_per_pixel=spp, max_bounces=depth, material_mode=nat.MODE_SM)
        if rng:
            kw["sample_range"] = rng
        v = make_view(sc, w, h, **kw)
        _, want = ctx.render(v, want_accum=True)
        segs = ctx.stats()["segments"]
        knobs(RTCU_BVH_RUNS="1", RTCU_BVH_RUN=run)
        img, got = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs and ctx.stats()["kernel_launches"] == 2
        _, again = ctx.render(v, want_accum=True)
        np.testing.assert_array_equal(again, got)
        knobs(RTCU_BVH_BEAM="0")
        _, nobeam = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs and ctx.stats()["kernel_launches"] == 1
        knobs(RTCU_BVH_RUNS=None, RTCU_BVH_RUN=None, RTCU_BVH_BEAM=None)
        np.testing.assert_array_equal(got[..., 3], want[..., 3])
        np.testing.assert_allclose(got[..., :3], want[..., :3], rtol=4e-6, atol=1e-6)
        np.testing.assert_allclose(nobeam[..., :3], want[..., :3], rtol=4e-6, atol=1e-6)
        if rng is None:
            r_rgba8, r_accum, r_segs = oracle.render(sc, v, threads=0)
            assert r_segs == segs
            assert np.abs(unpack_rgba(img) - unpack_rgba(r_rgba8)).max() <= 1
    knobs(RTCU_BVH_RUNS="1", RTCU_BVH_RUN=run)
    for tw, th, tile in ((3, 2, None), (40, 30, (17, 11, 18, 12)), (w, h, (13, 9, 150, 100))):
        tkw = dict(samples_per_pixel=16, max_bounces=depth, material_mode=nat.MODE_SM)
        tv = make_view(sc, tw, th, tile=tile, **tkw) if tile else make_view(sc, tw, th, **tkw)
        knobs(RTCU_BVH_RUNS="1", RTCU_BVH_RUN=run)
        _, small = ctx.render(tv, want_accum=True)
        knobs(RTCU_BVH_RUNS=None)
        _, small_d = ctx.render(tv, want_accum=True)
        np.testing.assert_array_equal(small[..., 3], small_d[..., 3])
        np.testing.assert_allclose(small[..., :3], small_d[..., :3], rtol=4e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("c1", nat.MODE_MG), ("c2", nat.MODE_SM), ("planes", nat.MODE_SM)])
def test_small_scan_frames_share_pixels_between_lanes(ctx, oracle, scenes, knobs, name, mode):
    """Scan (non-BVH) scenes: below 3840x2160, from 16 samples per call, a frame is rendered by k_render_stragglers in direct mode
    (4 / 8 / 16 lanes share a pixel's samples, pixel-sized work items handed out dynamically) instead of the thread-per-pixel grid,
    which cannot balance a few thousand tiles (C1 0.55 -> 0.35 ms).  Same paths (segment counts equal the oracle's and the other
    kernel's), sums equal up to fp32 order, deterministic; tiles of such a frame compose bit for bit while the lane count is the
    same.  RTCU_SCAN_NESTED=0 (lanes claim samples by ballot rank, as the BVH kernels do) is held to the same."""
    sc, depth = scenes[name]
    ctx.upload_scene(sc)
    w, h = 203, 117
    for spp in (20, 70):  # 8 lanes per pixel, 16 lanes per pixel
        kw = dict(samples_per_pixel=spp, max_bounces=depth, material_mode=mode)
        v = make_view(sc, w, h, **kw)
        rgba8, accum = ctx.render(v, want_accum=True)
        st = ctx.stats()
        assert st["kernel_launches"] == 1 and st["accel"] == nat.ACCEL_LINEAR and (accum[..., 3] == spp).all()
        rgba8_b, accum_b = ctx.render(v, want_accum=True)
        np.testing.assert_array_equal(accum_b, accum)
        np.testing.assert_array_equal(rgba8_b, rgba8)
        r_rgba8, r_accum, r_segs = oracle.render(sc, v, threads=0)
        assert r_segs == st["segments"]
        np.testing.assert_allclose(accum[..., :3], r_accum[..., :3], rtol=2e-5, atol=1e-6)
        assert np.abs(unpack_rgba(rgba8) - unpack_rgba(r_rgba8)).max() <= 1
        knobs(RTCU_SCAN_DIRECT="0")  # the thread-per-pixel grid: sequential sums, bit-identical to the oracle's order
        rgba8_t, accum_t = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == st["segments"] and ctx.stats()["kernel_launches"] >= 2
        knobs(RTCU_SCAN_DIRECT=None)
        np.testing.assert_allclose(accum[..., :3], accum_t[..., :3], rtol=4e-6, atol=1e-6)
        assert np.abs(unpack_rgba(rgba8) - unpack_rgba(rgba8_t)).max() <= 1
        # a tile of the frame (ragged against the 8x4 patches) equals the frame's pixels bit for bit and leaves the rest alone
        img = np.full((h, w), 0xDEADBEEF, np.uint32)
        part, pacc = ctx.render(make_view(sc, w, h, tile=(13, 9, 150, 100), **kw), rgba8=img, want_accum=True)
        np.testing.assert_array_equal(pacc[9:100, 13:150], accum[9:100, 13:150])
        np.testing.assert_array_equal(part[9:100, 13:150], rgba8[9:100, 13:150])
        assert (part[:9] == 0xDEADBEEF).all() and (part[:, :13] == 0xDEADBEEF).all() and (part[100:] == 0xDEADBEEF).all()
        # sample ranges: the same paths, partitioned; accumulated on the device they resolve to the frame
        lo = spp // 3
        ctx.render(make_view(sc, w, h, sample_range=(0, lo), **kw))
        segs_lo = ctx.stats()["segments"]
        img2, acc2 = ctx.render(make_view(sc, w, h, sample_range=(lo, spp), flags=nat.FLAG_ACCUMULATE, **kw), want_accum=True)
        assert segs_lo + ctx.stats()["segments"] == st["segments"]
        np.testing.assert_array_equal(acc2[..., 3], accum[..., 3])
        np.testing.assert_allclose(acc2[..., :3], accum[..., :3], rtol=4e-6, atol=1e-6)
        assert np.abs(unpack_rgba(img2) - unpack_rgba(rgba8)).max() <= 1
    # the other sample-to-lane assignment, and the lane counts the bigger frames take (forced here on the small frame)
    v = make_view(sc, w, h, samples_per_pixel=70, max_bounces=depth, material_mode=mode)
    _, want = ctx.render(v, want_accum=True)
    segs = ctx.stats()["segments"]
    for kn in (dict(RTCU_SCAN_NESTED="0"), dict(RTCU_SCAN_DIRECT="4"), dict(RTCU_SCAN_DIRECT="8"), dict(RTCU_SCAN_DIRECT="8", RTCU_SCAN_NESTED="0"),
               dict(RTCU_SCAN_DIRECT="2")):
        knobs(**kn)
        _, got = ctx.render(v, want_accum=True)
        assert ctx.stats()["segments"] == segs and ctx.stats()["kernel_launches"] == 1
        _, again = ctx.render(v, want_accum=True)
        knobs(**{k: None for k in kn})
        np.testing.assert_array_equal(again, got)
        np.testing.assert_array_equal(got[..., 3], want[..., 3])
        np.testing.assert_allclose(got[..., :3], want[..., :3], rtol=4e-6, atol=1e-6)
    # 1920x1080: 4 lanes per pixel below 64 samples, still one launch; 4 / 2 lanes for 8 / 4 samples per call; below 4 samples
    # per call the thread-per-pixel grid stays
    ctx.render(make_view(sc, 1920, 1080, samples_per_pixel=16, max_bounces=depth, material_mode=mode))
    assert ctx.stats()["kernel_launches"] == 1
    for spp in (5, 9):
        v = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=depth, material_mode=mode)
        rgba8, accum = ctx.render(v, want_accum=True)
        assert ctx.stats()["kernel_launches"] == 1
        r_rgba8, r_accum, r_segs = oracle.render(sc, v, threads=0)
        assert r_segs == ctx.stats()["segments"]
        np.testing.assert_array_equal(accum[..., 3], r_accum[..., 3])
        np.testing.assert_allclose(accum[..., :3], r_accum[..., :3], rtol=2e-5, atol=1e-6)
        assert np.abs(unpack_rgba(rgba8) - unpack_rgba(r_rgba8)).max() <= 1
    ctx.render(make_view(sc, w, h, samples_per_pixel=3, max_bounces=depth, material_mode=mode))
    assert ctx.stats()["kernel_launches"] >= 2
    # 3840x2160: scenes of at most 8 spheres (fixed-pair kernels) share pixels there too, the others keep the grid
    v4k = make_view(sc, 3840, 2160, samples_per_pixel=16, max_bounces=depth, material_mode=mode)
    rgba8_4k, _ = ctx.render(v4k, want_accum=True)
    segs_4k = ctx.stats()["segments"]
    assert (ctx.stats()["kernel_launches"] == 1) == (len(sc.spheres) <= 8)
    knobs(RTCU_SCAN_FIXED_PAIRS="0")  # the generic pair loop: the grid at 4K, the same paths everywhere
    rgba8_g, _ = ctx.render(v4k, want_accum=True)
    assert ctx.stats()["kernel_launches"] >= 2 and ctx.stats()["segments"] == segs_4k
    assert np.abs(unpack_rgba(rgba8_g) - unpack_rgba(rgba8_4k)).max() <= 1
    v = make_view(sc, w, h, samples_per_pixel=70, max_bounces=depth, material_mode=mode)
    _, got = ctx.render(v, want_accum=True)
    knobs(RTCU_SCAN_FIXED_PAIRS=None)
    assert ctx.stats()["segments"] == segs
    np.testing.assert_array_equal(got, want)  # fixed or generic pair loop: the same tests in the same order, bit for bit
