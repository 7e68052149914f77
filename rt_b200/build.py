"""In-tree build of the CUDA library (sm_100a only) and of the C++ plugin compile check.

`python -m rt_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.  The
`.so` lands in rt_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "rt_b200" / "csrc"
LIBDIR = ROOT / "rt_b200" / "lib"
LIB = LIBDIR / "librtcu.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # every fused multiply-add in the kernels is written explicitly (arithmetic SPEC)
    "-Xcompiler", "-fPIC,-pthread", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and pathlib.Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def sources() -> list[pathlib.Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "rtcu.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build_cuda(force: bool = False, verbose: bool = False) -> pathlib.Path:
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "rtcu.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


def build_plugin_check() -> None:
    """Compile the reference-side plugin TU against the minimal accessor stub (plugin/stub)."""
    plugin = ROOT / "plugin"
    src = plugin / "cuda_path_tracer.cpp"
    if not src.exists():
        return
    out = plugin / "_build"
    out.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++20", "-Wall", "-Wextra", "-fsyntax-only", "-DRTCU_PLUGIN_STUB_CHECK", f"-I{plugin / 'stub'}", f"-I{ROOT / 'include'}", str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("plugin compile check failed:\n" + r.stdout + r.stderr)


def build_oracle(fast: bool = False) -> pathlib.Path:
    """Builds the checker (oracle/); building it is not using it."""
    target = "fast" if fast else "strict"
    r = subprocess.run(["make", "-C", str(ROOT / "oracle"), target], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return ROOT / "oracle" / "_build" / f"librtref_{target}.so"


HOST_BIN = ROOT / "rt_b200" / "host" / "rt_headless"


def build_host() -> pathlib.Path:
    """The headless C++ host (scene loader + CLI over the C ABI): rt_b200/host/rt_headless, linked to librtcu.so."""
    src = ROOT / "rt_b200" / "host" / "rt_headless.cpp"
    deps = [src, src.with_name("scene_loader.hpp"), src.with_name("toml_lite.hpp"), src.with_name("colour_table.inc"), ROOT / "include" / "rtcu.h"]
    if HOST_BIN.exists() and all(d.stat().st_mtime <= HOST_BIN.stat().st_mtime for d in deps):
        return HOST_BIN
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", f"-I{ROOT / 'include'}", "-o", str(HOST_BIN), str(src),
           f"-L{LIBDIR}", "-lrtcu", "-Wl,-rpath,$ORIGIN/../lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + r.stdout + r.stderr)
    return HOST_BIN


def build_reference() -> bool:
    """oracle/_ref/librt_ref*.so from the reference's own sources (oracle/Makefile `ref`); a no-op where the reference
    tree is absent (the prebuilt libraries travel with the repo snapshot)."""
    if not pathlib.Path("/root/reference/src/renderers/mg_ray_tracer.cpp").exists():
        return False
    r = subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference build failed:\n" + r.stdout + r.stderr)
    return True


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    build_plugin_check()
    print(build_host())
    print(build_oracle())
    print("reference build:", build_reference())
