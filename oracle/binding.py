"""ctypes binding of the CPU oracle (oracle/rtref.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under rt_b200/ does.  PARITY UNPINNED against a reference binary (see rtref.h).
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
PRIM_MISS, PRIM_PLANE = 0xFFFFFFFF, 0x80000000


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("albedo", C.c_float * 4), ("roughness", C.c_float), ("reflectivity", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("spheres", C.c_void_p), ("sphere_material", C.c_void_p), ("n_spheres", C.c_uint32),
                ("planes", C.c_void_p), ("plane_material", C.c_void_p), ("n_planes", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32),
                ("boxes", C.c_void_p), ("box_material", C.c_void_p), ("n_boxes", C.c_uint32)]


class View(C.Structure):
    _fields_ = [("inv_view_proj", C.c_float * 16), ("width", C.c_uint32), ("height", C.c_uint32),
                ("samples_per_pixel", C.c_uint32), ("max_bounces", C.c_uint32),
                ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("tile_x0", C.c_uint32), ("tile_y0", C.c_uint32), ("tile_x1", C.c_uint32), ("tile_y1", C.c_uint32),
                ("seed", C.c_uint64), ("material_mode", C.c_uint32), ("flags", C.c_uint32)]


def build(flavour: str = "strict") -> pathlib.Path:
    out = HERE / "_build" / f"librtref_{flavour}.so"
    r = subprocess.run(["make", "-C", str(HERE), flavour], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return out


_MATERIAL_DTYPE = np.dtype([("type", "<u4"), ("albedo", "<f4", (4,)), ("roughness", "<f4"), ("reflectivity", "<f4")])


def _scene_desc(scene):
    """(SceneDesc, arrays that must outlive it)"""
    sph = np.ascontiguousarray(scene.spheres, np.float32).reshape(-1, 4)
    smat = np.ascontiguousarray(scene.sphere_material, np.uint32)
    pl = np.ascontiguousarray(scene.planes, np.float32).reshape(-1, 4)
    pmat = np.ascontiguousarray(scene.plane_material, np.uint32)
    bx = np.ascontiguousarray(getattr(scene, "boxes", np.zeros((0, 6))), np.float32).reshape(-1, 6)
    bmat = np.ascontiguousarray(getattr(scene, "box_material", np.zeros(0)), np.uint32)
    mats = np.ascontiguousarray(scene.materials.astype(_MATERIAL_DTYPE))
    sd = SceneDesc(sph.ctypes.data if len(sph) else None, smat.ctypes.data if len(smat) else None, len(sph),
                   pl.ctypes.data if len(pl) else None, pmat.ctypes.data if len(pmat) else None, len(pl),
                   mats.ctypes.data, len(mats),
                   bx.ctypes.data if len(bx) else None, bmat.ctypes.data if len(bmat) else None, len(bx))
    return sd, (sph, smat, pl, pmat, mats, bx, bmat)


class Oracle:
    """flavour 'strict' = the parity checker; 'fast' = the CPU timing baseline (reference's -ffast-math flags)."""

    def __init__(self, flavour: str = "strict"):
        self.lib = C.CDLL(str(build(flavour)))
        L, p, u32, u64, i = self.lib, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.rtref_philox_stream.argtypes = [p, p, p]
        L.rtref_philox4x32_r.argtypes = [p, p, C.c_int, p]
        L.rtref_u01.restype = C.c_float
        L.rtref_u01.argtypes = [u32]
        L.rtref_intersect_batch.argtypes = [C.POINTER(SceneDesc), p, p, u32, p, p, p, p]
        L.rtref_primary_ray.argtypes = [C.POINTER(View), u32, u32, u32, p, p]
        L.rtref_screen_ray.argtypes = [C.POINTER(View), C.c_float, C.c_float, p, p]
        L.rtref_scatter.argtypes = [C.POINTER(SceneDesc), u32, u32, p, p, C.c_float, p, u64, u32, u32, u32, p, p, p]
        L.rtref_pack_pixel.restype = u32
        L.rtref_pack_pixel.argtypes = [C.c_float, C.c_float, C.c_float, u32]
        L.rtref_render.argtypes = [C.POINTER(SceneDesc), C.POINTER(View), p, p, p, i, u32]
        L.rtref_trace_sample.restype = u32
        L.rtref_trace_sample.argtypes = [C.POINTER(SceneDesc), C.POINTER(View), u32, u32, u32, p]
        L.rtref_rasterize.argtypes = [C.POINTER(SceneDesc), C.POINTER(View), p, p, p, i, u32]
        L.rtref_ray_hits_box.argtypes = [p, p, p, p]
        L.rtref_build_flavour.restype = C.c_char_p
        assert L.rtref_build_flavour().decode() == flavour
        self._keep = None

    # ---- scene ------------------------------------------------------------------------------------
    def scene_desc(self, scene) -> SceneDesc:
        sd, self._keep = _scene_desc(scene)
        return sd

    @staticmethod
    def view_from(v) -> View:
        """copy an rt_b200 View (same field layout) into the oracle's own struct"""
        out = View()
        for name, _ in View._fields_:
            val = getattr(v, name)
            if name == "inv_view_proj":
                out.inv_view_proj[:] = list(val)
            else:
                setattr(out, name, val)
        return out

    # ---- entry points -------------------------------------------------------------------------------
    def philox(self, ctr, key: int) -> np.ndarray:
        c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
        k = (C.c_uint32 * 2)(key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF)
        o = (C.c_uint32 * 4)()
        self.lib.rtref_philox_stream(c, k, o)
        return np.array(list(o), np.uint32)

    def philox_rounds(self, ctr, key: int, rounds: int) -> np.ndarray:
        """Philox4x32 with an explicit round count (the published vectors exist for 7 and 10)"""
        c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
        k = (C.c_uint32 * 2)(key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF)
        o = (C.c_uint32 * 4)()
        self.lib.rtref_philox4x32_r(c, k, int(rounds), o)
        return np.array(list(o), np.uint32)

    def u01(self, x: int) -> float:
        return float(self.lib.rtref_u01(int(x)))

    def intersect_batch(self, scene, o, d):
        sd = self.scene_desc(scene)
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = len(o)
        hit = np.zeros(n, np.uint8); prim = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); nrm = np.zeros((n, 3), np.float32)
        rc = self.lib.rtref_intersect_batch(C.byref(sd), o.ctypes.data, d.ctypes.data, n, hit.ctypes.data, prim.ctypes.data, t.ctypes.data, nrm.ctypes.data)
        assert rc == 0
        return hit, prim, t, nrm

    def primary_rays(self, view, px, py, sample):
        v = self.view_from(view)
        n = len(px)
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        oo = (C.c_float * 3)(); dd = (C.c_float * 3)()
        for i in range(n):
            self.lib.rtref_primary_ray(C.byref(v), int(px[i]), int(py[i]), int(sample[i]), oo, dd)
            o[i] = list(oo); d[i] = list(dd)
        return o, d

    def screen_rays(self, view, sx, sy):
        """primary rays through explicit screen positions (pixel corners, edges, ...)"""
        v = self.view_from(view)
        n = len(sx)
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        oo = (C.c_float * 3)(); dd = (C.c_float * 3)()
        for i in range(n):
            self.lib.rtref_screen_ray(C.byref(v), float(sx[i]), float(sy[i]), oo, dd)
            o[i] = list(oo); d[i] = list(dd)
        return o, d

    def scatter_batch(self, scene, mode, seed, material, o, d, t, normal, pixel, sample, block):
        sd = self.scene_desc(scene)
        n = len(material)
        o = np.ascontiguousarray(o, np.float32).reshape(n, 3); d = np.ascontiguousarray(d, np.float32).reshape(n, 3)
        normal = np.ascontiguousarray(normal, np.float32).reshape(n, 3)
        sc = np.zeros(n, np.uint8); att = np.zeros((n, 3), np.float32); oo = np.zeros((n, 3), np.float32); do = np.zeros((n, 3), np.float32)
        for i in range(n):
            sc[i] = self.lib.rtref_scatter(C.byref(sd), mode, int(material[i]), o[i].ctypes.data, d[i].ctypes.data, float(t[i]), normal[i].ctypes.data,
                                           seed, int(pixel[i]), int(sample[i]), int(block[i]), att[i].ctypes.data, oo[i].ctypes.data, do[i].ctypes.data)
        return sc, att, oo, do

    def pack_pixel(self, r: float, g: float, b: float, spp: int) -> int:
        return int(self.lib.rtref_pack_pixel(r, g, b, spp))

    def render(self, scene, view, threads: int = 0, row_step: int = 1, want_rgba8: bool = True, want_accum: bool = True):
        sd = self.scene_desc(scene)
        v = self.view_from(view)
        rgba8 = np.zeros((v.height, v.width), np.uint32) if want_rgba8 else None
        accum = np.zeros((v.height, v.width, 4), np.float32) if want_accum else None
        segs = C.c_uint64(0)
        rc = self.lib.rtref_render(C.byref(sd), C.byref(v), rgba8.ctypes.data if want_rgba8 else None,
                                   accum.ctypes.data if want_accum else None, C.byref(segs), threads, row_step)
        if rc != 0:
            raise RuntimeError(f"rtref_render failed: {rc}")
        return rgba8, accum, int(segs.value)

    def rasterize(self, scene, view, threads: int = 0, row_step: int = 1):
        """rasterizer.cpp:22-88 -> (rgba8, prim, depth)"""
        sd = self.scene_desc(scene)
        v = self.view_from(view)
        rgba8 = np.zeros((v.height, v.width), np.uint32)
        prim = np.full((v.height, v.width), 0xFFFFFFFF, np.uint32)
        depth = np.zeros((v.height, v.width), np.float32)
        rc = self.lib.rtref_rasterize(C.byref(sd), C.byref(v), rgba8.ctypes.data, prim.ctypes.data, depth.ctypes.data, threads, row_step)
        if rc != 0:
            raise RuntimeError(f"rtref_rasterize failed: {rc}")
        return rgba8, prim, depth

    def ray_hits_box(self, o, d, box):
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32); box = np.ascontiguousarray(box, np.float32)
        t = C.c_float(0)
        hit = self.lib.rtref_ray_hits_box(o.ctypes.data, d.ctypes.data, box.ctypes.data, C.byref(t))
        return bool(hit), float(t.value)

    def trace_sample(self, scene, view, px: int, py: int, sample: int):
        sd = self.scene_desc(scene)
        v = self.view_from(view)
        out = (C.c_float * 3)()
        nseg = self.lib.rtref_trace_sample(C.byref(sd), C.byref(v), px, py, sample, out)
        return np.array(list(out), np.float32), int(nseg)


class ReferenceBuild:
    """oracle/_ref/librt_ref.so: marzer/rt's own mg_ray_tracer.cpp / sm_ray_tracer.cpp / renderer.cpp compiled against the
    ref_shim stand-in for muu (see oracle/Makefile `ref`).  TEST INFRASTRUCTURE.  The library is built in the dev
    container (where /root/reference exists) and travels to the GPU box prebuilt."""

    PATH = HERE / "_ref" / "librt_ref.so"
    FAST_PATH = HERE / "_ref" / "librt_ref_fast.so"  # -O3 -ffast-math flavour for CPU timing
    FAST_MT_PATH = HERE / "_ref" / "librt_ref_fast_mt.so"  # the same with the reference's own src/random.cpp (mt19937): non-deterministic
    PLUGIN_PATH = HERE / "_ref" / "librt_ref_plugin.so"  # + plugin/cuda_path_tracer.cpp against the reference's headers, linked to librtcu.so

    @classmethod
    def available(cls, build_if_possible: bool = True) -> bool:
        if build_if_possible and pathlib.Path("/root/reference/src/renderers/mg_ray_tracer.cpp").exists():
            r = subprocess.run(["make", "-C", str(HERE), "ref"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference build failed:\n" + r.stdout + r.stderr)
        return cls.PATH.exists()

    def __init__(self, flavour: str = "strict"):
        if not self.available():
            raise RuntimeError(f"{self.PATH} is missing and /root/reference is not present to build it")
        path = {"fast": self.FAST_PATH, "fast_mt": self.FAST_MT_PATH, "plugin": self.PLUGIN_PATH}.get(flavour, self.PATH)
        if not path.exists():
            raise RuntimeError(f"{path} is missing")
        self.lib = C.CDLL(str(path))
        p, u32 = C.c_void_p, C.c_uint32
        self.lib.refbin_list.argtypes = [C.c_char_p, u32]
        self.lib.refbin_last_error.restype = C.c_char_p
        self.lib.refbin_render.argtypes = [C.POINTER(SceneDesc), p, p, u32, u32, u32, u32, C.c_uint64, C.c_char_p, p, C.c_int, u32, p]
        self.lib.refbin_open.restype = p
        self.lib.refbin_open.argtypes = [C.c_char_p]
        self.lib.refbin_close.argtypes = [p]
        self.lib.refbin_render_with.argtypes = [p, C.POINTER(SceneDesc), p, p, u32, u32, u32, u32, C.c_uint64, p, C.c_int]
        self._keep = None

    # a renderer that lives across frames, as in the application (main.cpp:49-53)
    def open(self, renderer: str) -> int:
        h = self.lib.refbin_open(renderer.encode())
        if not h:
            raise RuntimeError("renderer construction failed: " + self.lib.refbin_last_error().decode(errors="replace"))
        return h

    def close(self, handle: int) -> None:
        self.lib.refbin_close(handle)

    def render_with(self, handle: int, scene, rgba8: np.ndarray, spp: int, max_bounces: int, seed: int, threads: int = 0) -> np.ndarray:
        """one frame of the open renderer into `rgba8` ((H, W) uint32, any host memory)"""
        sd = self._desc(scene)
        pos = np.array(scene.camera.position, np.float32); d = np.array(scene.camera.direction, np.float32)
        h, w = rgba8.shape
        rc = self.lib.refbin_render_with(handle, C.byref(sd), pos.ctypes.data, d.ctypes.data, w, h, spp, max_bounces, seed, rgba8.ctypes.data, threads)
        assert rc == 0
        return rgba8

    def plugin_last_stats(self) -> dict:
        """rt_cuda_last_stats of plugin/cuda_path_tracer.cpp (plugin flavour only): the library's statistics of the last frame"""
        from rt_b200._native import Stats  # the struct layout of include/rtcu.h

        st = Stats()
        self.lib.rt_cuda_last_stats(C.byref(st))
        return st.as_dict()

    def renderers(self) -> list:
        buf = C.create_string_buffer(1024)
        self.lib.refbin_list(buf, 1024)
        return [x for x in buf.value.decode().split("\n") if x]

    def _desc(self, scene) -> SceneDesc:
        sd, self._keep = _scene_desc(scene)
        return sd

    def inverse_view_projection(self, scene, width: int, height: int) -> np.ndarray:
        """the matrix the reference's own camera::viewport produces (through the stand-in matrix code)"""
        sd = self._desc(scene)
        pos = np.array(scene.camera.position, np.float32); d = np.array(scene.camera.direction, np.float32)
        out = np.zeros(16, np.float32)
        rc = self.lib.refbin_render(C.byref(sd), pos.ctypes.data, d.ctypes.data, width, height, 1, 1, 0, b"mg_ray_tracer", None, 1, 1, out.ctypes.data)
        assert rc == 0
        return out

    def render(self, scene, width: int, height: int, spp: int, max_bounces: int, seed: int, renderer: str = "mg_ray_tracer",
               threads: int = 0, row_step: int = 1):
        sd = self._desc(scene)
        pos = np.array(scene.camera.position, np.float32); d = np.array(scene.camera.direction, np.float32)
        rgba8 = np.zeros((height, width), np.uint32)
        ivp = np.zeros(16, np.float32)
        rc = self.lib.refbin_render(C.byref(sd), pos.ctypes.data, d.ctypes.data, width, height, spp, max_bounces, seed, renderer.encode(),
                                    rgba8.ctypes.data, threads, row_step, ivp.ctypes.data)
        if rc == -2:
            raise RuntimeError("renderer construction failed: " + self.lib.refbin_last_error().decode(errors="replace"))
        if rc != 0:
            raise RuntimeError(f"reference has no renderer named {renderer!r}")
        return rgba8, ivp
