"""Host-side mirror of the reference's renderer plugin interface, on top of the C ABI.

Mirrors, with the same names and argument meaning:

* `renderer_interface::render(const scene&, image_view&, muu::thread_pool&) noexcept` -- reference
  src/renderer.hpp:9-14  -> `RendererInterface.render(scene, pixels, threads=None)`
* `renderers::description / install / all / find_by_key / find_by_name` -- src/renderer.hpp:16-31,
  src/renderer.cpp:21-69 -> `Description`, `renderers.install(...)` etc.
* `REGISTER_RENDERER(T)` -- src/renderer.hpp:34-41 -> `@register_renderer`
* `rt::image_view` -- src/image.hpp:112-163 -> `ImageView`

`CudaPathTracer` is the Python twin of plugin/cuda_path_tracer.cpp: it flattens the scene into the POD
descriptor, calls `rtcu_upload_scene` + `rtcu_render`, and like a `noexcept` C++ override it never raises
from `render` -- it logs to stderr and leaves the (pre-cleared) image untouched ("Errors").
The lower-level `Context` class raises `RtcuError` and is what tests and bench.py use.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import hashlib
import sys
from typing import Callable, Optional, Sequence

import numpy as np

from . import _native as nat
from .camera import inverse_view_projection
from .scene import MATERIAL_DTYPE, Scene

DEFAULT_SEED = 0x5EED


# --------------------------------------------------------------------------------------------------
# rt::image_view
class ImageView:
    """Non-owning view over a row-major uint32 RGBA8888 buffer, row 0 = top (src/image.hpp:112-163)."""

    def __init__(self, data: np.ndarray):
        if data.dtype != np.uint32 or data.ndim != 2 or not data.flags.c_contiguous:
            raise ValueError("ImageView needs a C-contiguous (height, width) uint32 array")
        self.data = data

    @classmethod
    def allocate(cls, width: int, height: int) -> "ImageView":
        return cls(np.zeros((height, width), np.uint32))

    def size(self) -> tuple[int, int]:
        return (self.data.shape[1], self.data.shape[0])  # vec2u{x, y}

    def position_of(self, idx: int) -> tuple[int, int]:
        w = self.data.shape[1]
        return (idx % w, idx // w)  # src/image.hpp:156-159

    def clear(self, colour: int = 0x000000FF) -> "ImageView":
        self.data[...] = colour
        return self

    def __call__(self, x: int, y: int) -> int:
        return int(self.data[y, x])


# --------------------------------------------------------------------------------------------------
# renderer_interface + registry
class RendererInterface:
    """src/renderer.hpp:9-14"""

    def render(self, scene: Scene, pixels: ImageView, threads=None) -> None:  # noexcept in the reference
        raise NotImplementedError


@dataclasses.dataclass(frozen=True)
class Description:
    """renderers::description (src/renderer.hpp:18-25)"""
    key: str
    name: str
    create: Callable[[], RendererInterface]


class _Registry:
    """src/renderer.cpp:9-69"""

    def __init__(self) -> None:
        self._all: list[Description] = []

    def install(self, desc: Description) -> None:
        assert desc.key and desc.name and desc.create
        for i, r in enumerate(self._all):
            if r.key == desc.key:  # same key overwrites (renderer.cpp:27-34)
                self._all[i] = desc
                return
        self._all.append(desc)

    def all(self) -> Sequence[Description]:
        return tuple(self._all)

    def find_by_key(self, key: str) -> Optional[Description]:
        if not key:
            return None
        return next((r for r in self._all if r.key == key), None)

    def find_by_name(self, name: str) -> Optional[Description]:
        if not name:
            return None
        return next((r for r in self._all if r.name == name), None)

    def find(self, name: str) -> Optional[Description]:
        """CLI lookup: exact name, then prefix (src/main.cpp:68-81)."""
        return self.find_by_name(name) or next((r for r in self._all if name and r.name.startswith(name)), None)


renderers = _Registry()


def register_renderer(cls):
    """REGISTER_RENDERER(T): the CLI name is the type name (src/renderer.hpp:34-41)."""
    renderers.install(Description(key=f"{cls.__module__}:{cls.__qualname__}", name=cls.__name__, create=cls))
    return cls


# --------------------------------------------------------------------------------------------------
def make_view(scene: Scene, width: int, height: int, *, samples_per_pixel: Optional[int] = None,
              max_bounces: Optional[int] = None, sample_range: Optional[tuple[int, int]] = None,
              tile: Optional[tuple[int, int, int, int]] = None, seed: int = DEFAULT_SEED,
              material_mode: int = nat.MODE_SM, flags: int = 0) -> nat.View:
    """Fills an rtcu_view for `scene` (camera -> inverse_view_projection, defaults = the reference's single call)."""
    v = nat.View()
    v.inv_view_proj[:] = inverse_view_projection(scene.camera, width, height).tolist()
    v.width, v.height = width, height
    v.samples_per_pixel = scene.samples_per_pixel if samples_per_pixel is None else samples_per_pixel
    v.max_bounces = scene.max_bounces if max_bounces is None else max_bounces
    v.sample_begin, v.sample_end = (0, v.samples_per_pixel) if sample_range is None else sample_range
    v.tile_x0, v.tile_y0, v.tile_x1, v.tile_y1 = (0, 0, width, height) if tile is None else tile
    v.seed = seed
    v.material_mode = material_mode
    v.flags = flags
    return v


class Context:
    """One rtcu_ctx (one device).  Raises RtcuError; no CPU fallback."""

    def __init__(self, device: int = 0):
        self._lib = nat.load_library()
        self._h = self._lib.rtcu_create(device)
        if not self._h:
            raise nat.RtcuError(nat.RTCU_ERR_CUDA, nat.last_error())
        self.device = device
        self._keep: tuple = ()

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rtcu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> int:
        return self._h

    # -- scene ----------------------------------------------------------------------------------
    def prepare_scene(self, scene: Scene):
        """Flattens `scene` into the C descriptor once: (rtcu_scene, arrays that must outlive it).  A caller that re-sends an
        unchanged scene every frame (bench.py's end-to-end loop) skips the numpy conversions, not the upload."""
        sph = nat.contiguous(scene.spheres, np.float32).reshape(-1, 4)
        smat = nat.contiguous(scene.sphere_material, np.uint32)
        pl = nat.contiguous(scene.planes, np.float32).reshape(-1, 4)
        pmat = nat.contiguous(scene.plane_material, np.uint32)
        mats = np.ascontiguousarray(scene.materials.astype(MATERIAL_DTYPE))
        bx = nat.contiguous(getattr(scene, "boxes", np.zeros((0, 6))), np.float32).reshape(-1, 6)  # rasterizer only
        bmat = nat.contiguous(getattr(scene, "box_material", np.zeros(0)), np.uint32)
        d = nat.SceneDesc(nat.ptr(sph) if len(sph) else None, nat.ptr(smat) if len(smat) else None, len(sph),
                          nat.ptr(pl) if len(pl) else None, nat.ptr(pmat) if len(pmat) else None, len(pl),
                          nat.ptr(mats) if len(mats) else None, len(mats),
                          nat.ptr(bx) if len(bx) else None, nat.ptr(bmat) if len(bmat) else None, len(bx))
        return d, (sph, smat, pl, pmat, mats, bx, bmat)

    def upload_prepared(self, prepared) -> None:
        """rtcu_upload_scene with a descriptor from prepare_scene: validation, BVH build and the H2D transfer all happen"""
        nat.check(self._lib.rtcu_upload_scene(self._h, C.byref(prepared[0])))

    def upload_scene(self, scene: Scene) -> None:
        self.upload_prepared(self.prepare_scene(scene))

    # -- render ---------------------------------------------------------------------------------
    def render(self, view: nat.View, rgba8: Optional[np.ndarray] = None, accum: Optional[np.ndarray] = None,
               want_rgba8: bool = True, want_accum: bool = False):
        """rtcu_render into (new or caller-provided) host arrays; returns (rgba8 (H,W) u32, accum (H,W,4) f32)."""
        h, w = view.height, view.width
        if rgba8 is None and want_rgba8:
            rgba8 = np.zeros((h, w), np.uint32)
        if accum is None and want_accum:
            accum = np.zeros((h, w, 4), np.float32)
        nat.check(self._lib.rtcu_render(self._h, C.byref(view), nat.ptr(rgba8), nat.ptr(accum)))
        return rgba8, accum

    def rasterize(self, view: nat.View, rgba8: Optional[np.ndarray] = None, want_prim: bool = False, want_depth: bool = False):
        """rtcu_rasterize (rasterizer.cpp:22-88) -> (rgba8 (H,W) u32, prim (H,W) u32 | None, depth (H,W) f32 | None)"""
        h, w = view.height, view.width
        if rgba8 is None:
            rgba8 = np.zeros((h, w), np.uint32)
        prim = np.full((h, w), nat.PRIM_MISS, np.uint32) if want_prim else None
        depth = np.zeros((h, w), np.float32) if want_depth else None
        nat.check(self._lib.rtcu_rasterize(self._h, C.byref(view), nat.ptr(rgba8), nat.ptr(prim), nat.ptr(depth)))
        return rgba8, prim, depth

    def rasterize_device(self, view: nat.View, d_rgba8_ptr: int, stream: int = 0) -> None:
        nat.check(self._lib.rtcu_rasterize_device(self._h, C.byref(view), d_rgba8_ptr, stream or None))

    def render_device(self, view: nat.View, d_accum_ptr: int, accumulate: bool = False, stream: int = 0) -> None:
        nat.check(self._lib.rtcu_render_device(self._h, C.byref(view), d_accum_ptr, int(accumulate), stream or None))

    def resolve_device(self, d_accum_ptr: int, width: int, height: int, spp: int, d_rgba8_ptr: int, stream: int = 0) -> None:
        nat.check(self._lib.rtcu_resolve_device(self._h, d_accum_ptr, width, height, spp, d_rgba8_ptr, stream or None))

    def accum_download(self, width: int, height: int) -> np.ndarray:
        """the fp32 sums the context holds from its last render(s): (H, W, 4) {sum_r, sum_g, sum_b, n}"""
        out = np.zeros((height, width, 4), np.float32)
        nat.check(self._lib.rtcu_accum_download(self._h, width, height, nat.ptr(out)))
        return out

    def accum_upload(self, accum: np.ndarray) -> None:
        a = nat.contiguous(accum, np.float32)
        nat.check(self._lib.rtcu_accum_upload(self._h, a.shape[1], a.shape[0], nat.ptr(a)))

    def sync(self) -> None:
        nat.check(self._lib.rtcu_sync(self._h))

    def set_output_pinning(self, enable: bool) -> None:
        """rtcu_set_output_pinning: page-lock a pageable image that is handed over frame after frame (what the plugin does)"""
        nat.check(self._lib.rtcu_set_output_pinning(self._h, int(enable)))

    def reload_env(self) -> None:
        """the RTCU_* knobs are read at rtcu_create; re-read them for this context (A/B tests that change os.environ)"""
        nat.check(self._lib.rtcu_reload_env(self._h))

    def stats(self) -> dict:
        s = nat.Stats()
        nat.check(self._lib.rtcu_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # -- one process per GPU: IPC-shared buffers and the fused exchange ---------------------------------
    def ipc_alloc(self, nbytes: int) -> tuple[int, bytes]:
        """(device pointer, 64-byte CUDA IPC handle) of a zeroed buffer owned by this context"""
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        nat.check(self._lib.rtcu_ipc_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle.raw)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        nat.check(self._lib.rtcu_ipc_open(self._h, C.create_string_buffer(handle, 64), C.byref(ptr)))
        return int(ptr.value)

    def ipc_release(self, ptr: int) -> None:
        nat.check(self._lib.rtcu_ipc_release(self._h, ptr))

    def reduce_resolve_rows(self, accum_ptrs: Sequence[int], width: int, row0: int, rows: int, spp: int, d_rgba8_ptr: int, stream: int = 0) -> None:
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        nat.check(self._lib.rtcu_reduce_resolve_rows(self._h, arr, len(accum_ptrs), width, row0, rows, spp, d_rgba8_ptr, stream or None))

    def exchange_reduce_resolve(self, accum_ptrs: Sequence[int], flag_ptrs: Sequence[int], rank: int, dst: int, epoch: int, width: int,
                                row0: int, rows: int, spp: int, d_rgba8_ptr: int, stream: int = 0) -> None:
        """rtcu_exchange_reduce_resolve: handshake + peer-load sum + resolve + store, one launch"""
        a = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        f = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        nat.check(self._lib.rtcu_exchange_reduce_resolve(self._h, a, f, len(accum_ptrs), rank, dst, epoch, width, row0, rows, spp, d_rgba8_ptr, stream or None))

    def exchange_check(self, flags_ptr: int, stream: int = 0) -> None:
        nat.check(self._lib.rtcu_exchange_check(self._h, flags_ptr, stream or None))

    def selftest_math(self, divisors=(800, 600, 1920, 1080, 3840, 2160)) -> dict:
        """rtcu_selftest_math: mismatches of the kernels' cheaper exact sqrt / rcp / division against the IEEE intrinsics"""
        d = nat.contiguous(divisors, np.float32)
        out = np.zeros(4, np.uint64)
        nat.check(self._lib.rtcu_selftest_math(self._h, nat.ptr(d), len(d), nat.ptr(out)))
        return {"sqrt": int(out[0]), "rcp_of_sqrt": int(out[1]), "div": int(out[2]), "div_pairs": int(out[3])}

    def measure_fp32_peak(self) -> tuple[float, float]:
        """(FFMA TFLOP/s, FFMA2 TFLOP/s) achieved by a register-only stream on this device right now."""
        a, b = C.c_float(0), C.c_float(0)
        nat.check(self._lib.rtcu_measure_fp32_peak(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    # -- step-wise parity entry points -------------------------------------------------------------
    def intersect_batch(self, o: np.ndarray, d: np.ndarray, accel: int = nat.ACCEL_AUTO, want_normal: bool = True):
        o = nat.contiguous(o, np.float32).reshape(-1, 3)
        d = nat.contiguous(d, np.float32).reshape(-1, 3)
        n = len(o)
        hit = np.zeros(n, np.uint8)
        prim = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32) if want_normal else None
        nat.check(self._lib.rtcu_intersect_batch(self._h, nat.ptr(o), nat.ptr(d), n, nat.ptr(hit), nat.ptr(prim), nat.ptr(t), nat.ptr(nrm), accel))
        return hit, prim, t, nrm

    def primary_rays(self, view: nat.View, px, py, sample):
        px = nat.contiguous(px, np.uint32); py = nat.contiguous(py, np.uint32); sample = nat.contiguous(sample, np.uint32)
        n = len(px)
        o = np.zeros((n, 3), np.float32)
        d = np.zeros((n, 3), np.float32)
        nat.check(self._lib.rtcu_primary_rays(self._h, C.byref(view), nat.ptr(px), nat.ptr(py), nat.ptr(sample), n, nat.ptr(o), nat.ptr(d)))
        return o, d

    def scatter_batch(self, mode: int, seed: int, material, o, d, t, normal, pixel, sample, block):
        material = nat.contiguous(material, np.uint32)
        n = len(material)
        o = nat.contiguous(o, np.float32).reshape(n, 3); d = nat.contiguous(d, np.float32).reshape(n, 3)
        t = nat.contiguous(t, np.float32); normal = nat.contiguous(normal, np.float32).reshape(n, 3)
        pixel = nat.contiguous(pixel, np.uint32); sample = nat.contiguous(sample, np.uint32); block = nat.contiguous(block, np.uint32)
        sc = np.zeros(n, np.uint8)
        att = np.zeros((n, 3), np.float32); oo = np.zeros((n, 3), np.float32); do = np.zeros((n, 3), np.float32)
        nat.check(self._lib.rtcu_scatter_batch(self._h, mode, seed, n, nat.ptr(material), nat.ptr(o), nat.ptr(d), nat.ptr(t), nat.ptr(normal),
                                               nat.ptr(pixel), nat.ptr(sample), nat.ptr(block), nat.ptr(sc), nat.ptr(att), nat.ptr(oo), nat.ptr(do)))
        return sc, att, oo, do

    def philox_batch(self, ctr: np.ndarray, key: int) -> np.ndarray:
        ctr = nat.contiguous(ctr, np.uint32).reshape(-1, 4)
        out = np.zeros_like(ctr)
        nat.check(self._lib.rtcu_philox_batch(self._h, nat.ptr(ctr), len(ctr), key, nat.ptr(out)))
        return out


BVH_NODE_DTYPE = np.dtype([("x", "<f4", (4,)), ("y", "<f4", (4,)), ("z", "<f4", (4,)), ("child", "<i4", (2,)), ("count", "<u4", (2,))])


def bvh_build_host(spheres: np.ndarray):
    """The library's host BVH builder (no device needed): returns (nodes structured array, order, depth)."""
    lib = nat.load_library()
    sph = nat.contiguous(spheres, np.float32).reshape(-1, 4)
    n = len(sph)
    n_nodes, depth = C.c_uint32(0), C.c_uint32(0)
    nat.check(lib.rtcu_bvh_build_host(nat.ptr(sph) if n else None, n, None, None, 0, C.byref(n_nodes), C.byref(depth)))
    nodes = np.zeros(n_nodes.value, BVH_NODE_DTYPE)
    order = np.zeros(n, np.uint32)
    nat.check(lib.rtcu_bvh_build_host(nat.ptr(sph) if n else None, n, nat.ptr(nodes), nat.ptr(order) if n else None, len(nodes), C.byref(n_nodes), C.byref(depth)))
    return nodes, order, int(depth.value)


def bvh4_build_host(spheres: np.ndarray):
    """The 4-wide device tree built on the host (no device needed): (nodes (N, 8, 4) f32, leaves (L, 5, 4) f32, depth)."""
    lib = nat.load_library()
    sph = nat.contiguous(spheres, np.float32).reshape(-1, 4)
    n_nodes, n_leaves, depth = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    nat.check(lib.rtcu_bvh4_build_host(nat.ptr(sph), len(sph), None, 0, None, 0, C.byref(n_nodes), C.byref(n_leaves), C.byref(depth)))
    nodes = np.zeros((n_nodes.value, 8, 4), np.float32)
    leaves = np.zeros((n_leaves.value, 5, 4), np.float32)
    nat.check(lib.rtcu_bvh4_build_host(nat.ptr(sph), len(sph), nat.ptr(nodes), len(nodes), nat.ptr(leaves), len(leaves),
                                       C.byref(n_nodes), C.byref(n_leaves), C.byref(depth)))
    return nodes, leaves, int(depth.value)


def upload_scene_multi(contexts: Sequence[Context], scene: Scene) -> None:
    """rtcu_upload_scene_multi: one validation / BVH build, one copy per device"""
    lib = nat.load_library()
    arr = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    prepared = contexts[0].prepare_scene(scene)
    nat.check(lib.rtcu_upload_scene_multi(arr, len(contexts), C.byref(prepared[0])))


def render_multi(contexts: Sequence[Context], view: nat.View, want_accum: bool = False, rgba8: Optional[np.ndarray] = None):
    """rtcu_render_multi: single-process sample-range split over several devices."""
    lib = nat.load_library()
    arr = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    if rgba8 is None:
        rgba8 = np.zeros((view.height, view.width), np.uint32)
    accum = np.zeros((view.height, view.width, 4), np.float32) if want_accum else None
    nat.check(lib.rtcu_render_multi(arr, len(contexts), C.byref(view), nat.ptr(rgba8), nat.ptr(accum)))
    return rgba8, accum


class ProgressiveRenderer:
    """Progressive refinement behind the same boundary: every `refine()` call traces the next
    `samples_per_step` global samples of the frame and returns the image resolved over the samples so far -- what an
    interactive `render()` would show while the camera rests.  The fp32 sums stay ON THE DEVICE (RTCU_FLAG_ACCUMULATE):
    a step moves a few hundred bytes in and the packed image out, and the divide / sqrt / pack runs in the kernels.  Because
    sample indices are global (counter-based RNG), the image after k steps equals a single render of k*samples_per_step
    samples up to fp32 summation order.  `accum` reads the sums back on demand (checkpoints, tests)."""

    def __init__(self, ctx: "Context", scene: Scene, width: int, height: int, *, samples_per_step: int = 4, max_bounces: Optional[int] = None,
                 material_mode: int = nat.MODE_SM, seed: int = DEFAULT_SEED, flags: int = 0):
        self.ctx, self.scene, self.width, self.height = ctx, scene, width, height
        self.samples_per_step, self.max_bounces = samples_per_step, max_bounces
        self.material_mode, self.seed, self.flags = material_mode, seed, flags
        self.samples_done = 0
        # a Context keeps the sums on its device; anything else (the CPU stand-in of the host-logic tests) is summed here
        self._on_device = hasattr(ctx, "accum_download")
        self._host_accum = None if self._on_device else np.zeros((height, width, 4), np.float32)
        self._image = np.zeros((height, width), np.uint32)
        ctx.upload_scene(scene)

    @property
    def accum(self) -> np.ndarray:
        if not self._on_device:
            return self._host_accum
        if self.samples_done == 0:
            return np.zeros((self.height, self.width, 4), np.float32)
        return self.ctx.accum_download(self.width, self.height)

    def reset(self) -> None:
        """camera moved / scene reloaded (main.cpp:233-313): start over"""
        self.samples_done = 0
        if not self._on_device:
            self._host_accum[...] = 0
        self.ctx.upload_scene(self.scene)

    def refine(self) -> np.ndarray:
        begin, end = self.samples_done, self.samples_done + self.samples_per_step
        flags = self.flags | (nat.FLAG_ACCUMULATE if self._on_device and begin > 0 else 0)
        view = make_view(self.scene, self.width, self.height, samples_per_pixel=end, max_bounces=self.max_bounces,
                         sample_range=(begin, end), seed=self.seed, material_mode=self.material_mode, flags=flags)
        if self._on_device:
            self.ctx.render(view, rgba8=self._image, want_accum=False)  # sums added and resolved on the device
            self.samples_done = end
            return self._image.copy()
        _, part = self.ctx.render(view, want_rgba8=False, want_accum=True)
        self._host_accum += part
        self.samples_done = end
        return self.resolve()

    # -- checkpoint / resume (the reference is stateless per frame; a long render here is a sum over global
    #    sample indices, so the accumulation buffer and the number of samples done are all there is to save) -------------------
    def _identity(self) -> dict:
        cam = self.scene.camera
        return {"width": self.width, "height": self.height, "seed": int(self.seed), "material_mode": int(self.material_mode),
                "max_bounces": int(self.scene.max_bounces if self.max_bounces is None else self.max_bounces),
                "camera": [float(x) for x in (*cam.position, *cam.direction)], "scene": scene_fingerprint(self.scene).hex()}

    def save(self, path) -> None:
        """Writes the state of the render to `path` (.npz): fp32 sums, samples done, and what they are sums *of* (scene content,
        camera, size, seed, scatter table, depth).  Written to a temporary file and renamed, so an interrupted save leaves the
        previous checkpoint intact."""
        import json, os, tempfile
        path = os.fspath(path)
        fd, tmp = tempfile.mkstemp(dir=os.path.dirname(os.path.abspath(path)), suffix=".tmp")
        try:
            with os.fdopen(fd, "wb") as f:
                np.savez(f, accum=self.accum, samples_done=np.int64(self.samples_done), identity=np.frombuffer(json.dumps(self._identity()).encode(), np.uint8))
            os.replace(tmp, path)
        except BaseException:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise

    def restore(self, path) -> int:
        """Continues from a checkpoint written by `save`; returns the samples already done.  Refuses a checkpoint of a different
        scene, camera, size, seed, scatter table or depth: its sums would not be sums of this render's samples."""
        import json
        with np.load(path) as z:
            identity = json.loads(bytes(z["identity"]).decode())
            accum, done = z["accum"], int(z["samples_done"])
        mine = self._identity()
        differs = sorted(k for k in mine if identity.get(k) != mine[k])
        if differs:
            raise ValueError(f"checkpoint {path} belongs to a different render ({', '.join(differs)} differ)")
        if accum.shape != (self.height, self.width, 4) or accum.dtype != np.float32 or done < 0:
            raise ValueError(f"checkpoint {path} is malformed")
        if self._on_device:
            self.ctx.accum_upload(accum)
        else:
            self._host_accum[...] = accum
        self.samples_done = done
        return done

    def resolve(self) -> np.ndarray:
        """divide / sqrt / pack over the samples so far (mg_ray_tracer.cpp:195-200) from a host copy of the sums (refine() returns
        the image the device resolved; this is the same arithmetic for a stand-in context and for checks)"""
        n = np.float32(max(self.samples_done, 1))
        c = np.sqrt(self.accum[..., :3] / n, dtype=np.float32)
        c = np.minimum(np.maximum(c, np.float32(0)), np.float32(1))
        b = (c * np.float32(255.99999)).astype(np.uint32)
        return (b[..., 0] << 24) | (b[..., 1] << 16) | (b[..., 2] << 8) | np.uint32(255)


def scene_fingerprint(scene: Scene) -> bytes:
    """The reference scene has no dirty flag: detect changes by content."""
    h = hashlib.blake2b(digest_size=16)
    for a in (scene.spheres, scene.sphere_material, scene.planes, scene.plane_material, scene.materials,
              getattr(scene, "boxes", ()), getattr(scene, "box_material", ())):
        h.update(np.ascontiguousarray(a).tobytes())
        h.update(b"|")
    return h.digest()


@register_renderer
class cuda_path_tracer(RendererInterface):  # noqa: N801 -- the CLI name is the type name (renderer.hpp:34-41)
    """Drop-in for mg_ray_tracer / sm_ray_tracer behind the renderer interface.

    `material_mode` picks the scatter table: MODE_SM (default; lambert/metal/dielectric, sm_ray_tracer.cpp:221-236)
    or MODE_MG (mg_ray_tracer.cpp:142-152, no dielectric).  The two agree on scenes without dielectric-class
    materials."""

    material_mode = nat.MODE_SM
    seed = DEFAULT_SEED
    flags = 0

    def __init__(self, device: int = 0):
        self.ctx = Context(device)  # may raise, like a throwing C++ constructor inside create()
        self._fingerprint: Optional[bytes] = None
        self.last_error: Optional[str] = None

    def render(self, scene: Scene, pixels: ImageView, threads=None) -> None:
        try:
            fp = scene_fingerprint(scene)
            if fp != self._fingerprint:
                self.ctx.upload_scene(scene)
                self._fingerprint = fp
            w, h = pixels.size()
            view = make_view(scene, w, h, seed=self.seed, material_mode=self.material_mode, flags=self.flags)
            self.ctx.render(view, rgba8=pixels.data, want_accum=False)
            self.last_error = None
        except Exception as e:  # noexcept: log, leave the pre-cleared image alone
            self.last_error = str(e)
            print(f"cuda_path_tracer: {e}", file=sys.stderr)


@register_renderer
class cuda_rasterizer(RendererInterface):  # noqa: N801
    """Drop-in for `rasterizer` (reference src/renderers/rasterizer.cpp:22-88): the one-ray-per-pixel N.L preview.
    Unlike the path tracers it also draws `scene.boxes`."""

    def __init__(self, device: int = 0):
        self.ctx = Context(device)
        self._fingerprint: Optional[bytes] = None
        self.last_error: Optional[str] = None

    def render(self, scene: Scene, pixels: ImageView, threads=None) -> None:
        try:
            fp = scene_fingerprint(scene)
            if fp != self._fingerprint:
                self.ctx.upload_scene(scene)
                self._fingerprint = fp
            w, h = pixels.size()
            self.ctx.rasterize(make_view(scene, w, h), rgba8=pixels.data)
            self.last_error = None
        except Exception as e:  # noexcept: log, leave the pre-cleared image alone
            self.last_error = str(e)
            print(f"cuda_rasterizer: {e}", file=sys.stderr)


CudaPathTracer = cuda_path_tracer
CudaRasterizer = cuda_rasterizer
