"""World-size-2 (and 3) gloo runs of the multi-GPU composition logic on CPU: sample-range partition,
reduce of the fp32 accumulation buffers onto rank 0, resolve on rank 0.  The per-rank
renderer here is the CPU oracle (tests may use it); on the GPU box the same code path runs with the CUDA
renderer over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from rt_b200 import dist, scene as S
from rt_b200.renderer import make_view


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, by: str, out_path: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle

        oracle = Oracle("strict")
        sc = S.load("scenes/dielectric.toml")
        view = make_view(sc, 48, 32, samples_per_pixel=8, max_bounces=50, material_mode=1)

        def render_accum(v):
            _, accum, _ = oracle.render(sc, v, threads=1, want_rgba8=False)
            return torch.from_numpy(accum)

        def resolve(accum):
            a = accum.numpy()
            return np.array([[oracle.pack_pixel(*a[y, x, :3], view.samples_per_pixel) for x in range(a.shape[1])] for y in range(a.shape[0])], np.uint32)

        accum, rgba8 = dist.render_distributed(render_accum, view, rank=rank, world=world, by=by, resolve=resolve)
        if rank == 0:
            np.savez(out_path, accum=accum.numpy(), rgba8=rgba8)
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("world,by", [(2, "samples"), (3, "samples"), (2, "rows")])
def test_partition_and_reduce_match_single_process(tmp_path, oracle, world, by):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), by, out), nprocs=world, join=True)
    got = np.load(out)
    sc = S.load("scenes/dielectric.toml")
    view = make_view(sc, 48, 32, samples_per_pixel=8, max_bounces=50, material_mode=1)
    rgba8, accum, _ = oracle.render(sc, view, threads=2)
    assert (got["accum"][..., 3] == 8).all()  # every pixel received all 8 samples exactly once
    if by == "rows":
        np.testing.assert_array_equal(got["accum"], accum)  # disjoint pixels: bit-exact
        np.testing.assert_array_equal(got["rgba8"], rgba8)
    else:
        # partial sums are added in a different order: last-bit differences only
        np.testing.assert_allclose(got["accum"], accum, rtol=2e-6, atol=1e-7)
        d = np.abs(((got["rgba8"] >> 8) & 0xFF).astype(int) - ((rgba8 >> 8) & 0xFF).astype(int))
        assert d.max() <= 1


def _band_worker(rank: int, world: int, port: int, out_path: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle

        oracle = Oracle("strict")
        sc = S.load("scenes/dielectric.toml")
        w, h = 40, 27  # 27 rows over 2 or 4 ranks: the last band hangs over the frame
        view = make_view(sc, w, h, samples_per_pixel=8, max_bounces=50, material_mode=1)
        mine = dist.partition_view(view, rank, world)
        band = dist.band_rows(h, world)
        padded = torch.zeros((band * world, w, 4), dtype=torch.float32)
        _, accum, _ = oracle.render(sc, mine, threads=1, want_rgba8=False)
        padded[:h] = torch.from_numpy(accum)
        mine_band = dist.sum_row_bands(padded, rank, world)
        packed = torch.tensor([[oracle.pack_pixel(*mine_band[y, x, :3].tolist(), 8) for x in range(w)] for y in range(band)], dtype=torch.int64)
        img = dist.gather_bands(packed, h, rank, world)
        if rank == 0:
            np.savez(out_path, rgba8=img.numpy().astype(np.uint32))
        else:
            assert img is None
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_band_reduce_resolve_gather(tmp_path, oracle, world):
    out = str(tmp_path / "bands.npz")
    mp.spawn(_band_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)["rgba8"]
    sc = S.load("scenes/dielectric.toml")
    view = make_view(sc, 40, 27, samples_per_pixel=8, max_bounces=50, material_mode=1)
    rgba8, _, _ = oracle.render(sc, view, threads=2)
    assert got.shape == rgba8.shape
    d = np.abs(((got[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int) - ((rgba8[..., None] >> np.uint32([24, 16, 8])) & 255).astype(int))
    assert d.max() <= 1
