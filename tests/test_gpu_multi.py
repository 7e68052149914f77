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
        pytest.skip(f"needs {ngpu} GPUs, this box has {_device_count()}")
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


def _peer_worker(rank: int, world: int, port: int, out_path: str):
    """one process per GPU: the fused peer exchange of rt_b200.dist (IPC-shared buffers, rtcu_reduce_resolve_rows)"""
    import os

    import torch
    import torch.distributed as dist

    from rt_b200 import dist as rdist, scene as S

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sc = S.load("scenes/dielectric.toml")
        ctx = Context(rank)
        ctx.upload_scene(sc)
        w, h, spp = 640, 360, 64
        view = make_view(sc, w, h, samples_per_pixel=spp, max_bounces=50, material_mode=nat.MODE_SM)
        mine = rdist.partition_view(view, rank, world)
        gr = rdist.GpuRank(ctx, w, h, dev, world=world)
        gr.enable_peer_exchange(rank)
        frames = []
        for i in range(5):  # the two buffers per rank alternate frame after frame; frames 3 and 4 are queued back to back
            img = gr.render_peer_reduce_resolve(mine, spp)  # one launch: flag handshake + peer-load sum + resolve + store
            if i != 3:
                torch.cuda.synchronize(dev)
            if rank == 0:
                frames.append(img.cpu().numpy().view(np.uint32).copy())  # (stream-ordered after the frame's kernel)
        gr.check_exchange()  # no rank timed out
        barrier = gr.render_peer_barrier_reduce_resolve(mine, spp)  # round-1 form: the same reduce kernel between two NCCL barriers
        torch.cuda.synchronize(dev)
        barrier = barrier.cpu().numpy().view(np.uint32).copy() if rank == 0 else None
        nccl = gr.render_reduce_resolve(mine, rank, spp)  # the NCCL exchange on the same share
        torch.cuda.synchronize(dev)
        if rank == 0:
            np.savez(out_path, peer=np.stack(frames), barrier=barrier, nccl=nccl.cpu().numpy().view(np.uint32))
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_peer_exchange_one_process_per_gpu(ctx, scenes, tmp_path, ngpu):
    if _device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, this box has {_device_count()}")
    import torch.multiprocessing as mp

    out = str(tmp_path / "peer.npz")
    mp.spawn(_peer_worker, args=(ngpu, 29650 + ngpu, out), nprocs=ngpu, join=True)
    got = np.load(out)
    sc = scenes["c2"][0]
    view = make_view(sc, 640, 360, samples_per_pixel=64, max_bounces=50, material_mode=nat.MODE_SM)
    # the single-process peer reduce (rtcu_render_multi) sums in the same device order: bit-identical
    ctxs = [Context(g) for g in range(ngpu)]
    try:
        for c in ctxs:
            c.upload_scene(sc)
        ref, _ = render_multi(ctxs, view)
    finally:
        for c in ctxs:
            c.close()
    for frame in got["peer"]:
        np.testing.assert_array_equal(frame, ref)
    np.testing.assert_array_equal(got["barrier"], ref)
    assert np.abs(unpack_rgba(got["nccl"]) - unpack_rgba(ref)).max() <= 1  # NCCL's sum order is its own
