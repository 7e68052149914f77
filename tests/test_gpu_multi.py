"""Multi-GPU composition on real devices (skipped on a single-GPU box): rtcu_render_multi splits the sample
range over the devices and sums the fp32 buffers through NVLink peer loads inside the resolve kernel."""
import numpy as np
import pytest

from rt_b200 import _native as nat
from rt_b200.renderer import Context, make_view, render_multi

from conftest import unpack_rgba

pytestmark = pytest.mark.gpu


def _device_count():
    return nat.load_library().rtcu_device_count()


@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_render_multi_matches_single_device(ctx, scenes, ngpu):
    if _device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    sc = scenes["c2"][0]
    view = make_view(sc, 640, 360, samples_per_pixel=64, max_bounces=50, material_mode=nat.MODE_SM)
    ctx.upload_scene(sc)
    rgba_1, accum_1 = ctx.render(view, want_accum=True)
    segs_1 = ctx.stats()["segments"]
    ctxs = [Context(g) for g in range(ngpu)]
    try:
        for c in ctxs:
            c.upload_scene(sc)
        rgba_n, accum_n = render_multi(ctxs, view, want_accum=True)
        assert ctxs[0].stats()["segments"] == segs_1  # same paths, partitioned by sample range
        assert (accum_n[..., 3] == 64).all()
        np.testing.assert_allclose(accum_n[..., :3], accum_1[..., :3], rtol=2e-6, atol=1e-7)  # summation order only
        assert np.abs(unpack_rgba(rgba_n) - unpack_rgba(rgba_1)).max() <= 1
        # deterministic: the peer sum order is fixed (device 0, 1, 2, ...)
        rgba_m, accum_m = render_multi(ctxs, view, want_accum=True)
        np.testing.assert_array_equal(accum_m, accum_n)
    finally:
        for c in ctxs:
            c.close()
