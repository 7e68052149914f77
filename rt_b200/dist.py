"""Multi-GPU composition: one process per GPU, sample-range (and optional row-band) partition, one
reduce of the fp32 accumulation buffers.

Every (pixel, sample) is independent under the counter-based RNG (key = seed, counter = pixel, sample,
block), so rank g of G renders global sample indices [g*spp/G, (g+1)*spp/G) of the whole image and the
only exchange is `torch.distributed.reduce(accum, dst=0, SUM)` over NCCL/NVLink; rank 0 then resolves
(divide, sqrt, pack -- mg_ray_tracer.cpp:195-200).  Sample 0's pixel-centre rule (mg_ray_tracer.cpp:189)
applies to *global* sample 0 only, which lives on rank 0.

torch is imported lazily: it is plumbing (device memory, streams, process group), not the product.
"""
from __future__ import annotations

from typing import Callable, Optional

from . import _native as nat


def sample_range_for_rank(sample_begin: int, sample_end: int, rank: int, world: int) -> tuple[int, int]:
    """Even split of [sample_begin, sample_end) -- the same formula rtcu_render_multi uses on the device side."""
    total = sample_end - sample_begin
    return (sample_begin + total * rank // world, sample_begin + total * (rank + 1) // world)


def row_band_for_rank(y0: int, y1: int, rank: int, world: int) -> tuple[int, int]:
    """Tile split by row bands (no reduction needed, only a gather); used when spp < world."""
    rows = y1 - y0
    return (y0 + rows * rank // world, y0 + rows * (rank + 1) // world)


def partition_view(view: nat.View, rank: int, world: int, by: str = "samples") -> nat.View:
    v = type(view).from_buffer_copy(bytes(view))  # plain-data struct: byte copy
    if by == "samples":
        v.sample_begin, v.sample_end = sample_range_for_rank(view.sample_begin, view.sample_end, rank, world)
    elif by == "rows":
        v.tile_y0, v.tile_y1 = row_band_for_rank(view.tile_y0, view.tile_y1, rank, world)
    else:
        raise ValueError(by)
    return v


def render_distributed(render_accum: Callable[[nat.View], "object"], view: nat.View, *, rank: int, world: int,
                       by: str = "samples", group=None, dst: int = 0, resolve: Optional[Callable] = None):
    """Render this rank's share and reduce onto `dst`.

    render_accum(view) -> torch tensor (H, W, 4) fp32 holding {sum_r,sum_g,sum_b,n} for the rank's share and
    zeros elsewhere (CUDA tensor under NCCL, CPU tensor under gloo).  Returns (accum, resolved) on dst
    (resolved = resolve(accum) when given), (accum_partial, None) elsewhere.
    """
    import torch.distributed as dist

    mine = partition_view(view, rank, world, by)
    accum = render_accum(mine)
    if world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank == dst and resolve is not None:
        return accum, resolve(accum)
    return accum, None


class GpuRank:
    """One rank's device state for bench.py / multi-GPU runs: a Context plus torch-owned frame buffers."""

    def __init__(self, ctx, width: int, height: int, device=None):
        import torch

        self.torch = torch
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device) if device is None else device
        self.accum = torch.zeros((height, width, 4), dtype=torch.float32, device=self.device)
        self.rgba8 = torch.zeros((height, width), dtype=torch.int32, device=self.device)  # bit pattern of uint32
        self.width, self.height = width, height

    def render_accum(self, view: nat.View):
        """Trace `view` into self.accum on torch's current stream (zero-filled first when the tile is partial)."""
        full = view.tile_x0 == 0 and view.tile_y0 == 0 and view.tile_x1 == view.width and view.tile_y1 == view.height
        if not full:
            self.accum.zero_()
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.render_device(view, self.accum.data_ptr(), accumulate=False, stream=stream)
        return self.accum

    def resolve(self, accum, spp: int):
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.resolve_device(accum.data_ptr(), self.width, self.height, spp, self.rgba8.data_ptr(), stream=stream)
        return self.rgba8
