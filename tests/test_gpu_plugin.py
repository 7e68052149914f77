"""The drop-in, end to end: plugin/cuda_path_tracer.cpp compiled against the reference's REAL headers (scene.hpp, soa.hpp,
camera.hpp, image.hpp, renderer.hpp -- with the muu stand-in) and linked to librtcu.so registers itself in the reference's
own registry; `renderers::find_by_name("cuda_path_tracer")` then renders on the B200 through the C ABI, next to the
reference's CPU renderers in the same binary (oracle/_ref/librt_ref_plugin.so, built by `make -C oracle ref`)."""
import numpy as np
import pytest

from rt_b200 import scene as S

from conftest import unpack_rgba


def _plugin_build():
    from oracle.binding import ReferenceBuild

    if not (ReferenceBuild.PLUGIN_PATH.exists() or ReferenceBuild.available()):
        pytest.skip("oracle/_ref/librt_ref_plugin.so not present and no reference tree to build it from")
    return ReferenceBuild("plugin")


def test_plugin_registers_under_its_type_name_and_fails_loudly_without_a_device():
    ref = _plugin_build()
    # REGISTER_RENDERER(cuda_path_tracer) / REGISTER_RENDERER(cuda_rasterizer) next to the reference's own three
    assert ref.renderers() == ["mg_ray_tracer", "sm_ray_tracer", "rasterizer", "cuda_path_tracer", "cuda_rasterizer"]
    from rt_b200 import _native as nat

    if nat.load_library().rtcu_device_count() > 0:
        pytest.skip("device present: covered by the gpu test below")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ref.render(S.load("scenes/basic.toml"), 32, 24, 1, 2, 0x5EED, "cuda_path_tracer")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ref.render(S.load("scenes/basic.toml"), 32, 24, 1, 2, 0x5EED, "cuda_rasterizer")


@pytest.mark.gpu
@pytest.mark.parametrize("scene_file,cpu_renderer,env_mode", [("scenes/dielectric.toml", "sm_ray_tracer", None), ("scenes/basic.toml", "mg_ray_tracer", "mg"),
                                                               ("scenes/dielectric.toml", "mg_ray_tracer", "mg")])
def test_plugin_renders_like_the_reference_renderers(monkeypatch, scene_file, cpu_renderer, env_mode):
    ref = _plugin_build()
    if env_mode:
        monkeypatch.setenv("RT_CUDA_MATERIAL_MODE", env_mode)
    else:
        monkeypatch.delenv("RT_CUDA_MATERIAL_MODE", raising=False)
    sc = S.load(scene_file)
    w, h, spp, depth = 320, 200, 16, 50
    cpu, _ = ref.render(sc, w, h, spp, depth, 0x5EED, cpu_renderer, threads=0)    # the reference's own CPU loops
    gpu, _ = ref.render(sc, w, h, spp, depth, 0x5EED, "cuda_path_tracer")          # same registry, same scene object, B200
    d = np.abs(unpack_rgba(gpu) - unpack_rgba(cpu))
    assert d.max() <= 1, int(d.max())
    assert (d > 0).mean() <= 0.005
    assert len(np.unique(gpu)) > 100  # a real image, not the cleared buffer


@pytest.mark.gpu
@pytest.mark.parametrize("scene_file", ["scenes/basic.toml", "scenes/dielectric.toml", "scenes/boxes.toml"])
def test_plugin_rasterizer_equals_the_reference_rasterizer(scene_file):
    """cuda_rasterizer against rasterizer.cpp in the same binary, same rt::scene: no RNG on this path, so bit for bit"""
    ref = _plugin_build()
    sc = S.load(scene_file)
    w, h = 400, 250
    cpu, _ = ref.render(sc, w, h, 1, 1, 0, "rasterizer", threads=0)
    gpu, _ = ref.render(sc, w, h, 1, 1, 0, "cuda_rasterizer")
    np.testing.assert_array_equal(gpu, cpu)
    assert len(np.unique(gpu)) > 50
