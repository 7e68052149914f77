"""Device arithmetic that is a pure FMA algorithm can be proven on the CPU as well: spec.cuh's `div_by_const` (pixel coordinate / image
side without a MUFU or a range test) against IEEE division for every image side a window can have, not only the sides the device
self-test samples (tests/test_gpu_parity.py::test_cheaper_exact_sqrt_rcp_div_equal_the_ieee_intrinsics_for_every_input)."""
import ctypes as C
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_div_by_const_is_ieee_division_for_every_image_side(tmp_path):
    so = tmp_path / "libdivcheck.so"
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-mfma", "-o", str(so),
                    str(ROOT / "tests" / "tools" / "div_by_const_check.c"), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.div_by_const_check.restype = C.c_uint64
    lib.div_by_const_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.c_float * 2]
    checked, first_bad = C.c_uint64(0), (C.c_float * 2)()
    # every side 1 .. 8192 (the formats up to 8K), 3 + 5 coordinates per pixel column: 2.7e8 divisions
    bad = lib.div_by_const_check(1, 8192, 5, C.byref(checked), first_bad)
    assert bad == 0, (bad, list(first_bad))
    assert checked.value == (8192 * 8193 // 2) * 8
    # a sparse set of larger sides up to the ABI's limit of 2^23 pixels per side (rtcu.cu make_params)
    for side in (8193, 10000, 16384, 65535, 65536, 100003, 1 << 20, (1 << 23) - 1, 1 << 23):
        bad = lib.div_by_const_check(side, side, 2 if side > 70000 else 20, C.byref(checked), first_bad)
        assert bad == 0, (side, bad, list(first_bad))
