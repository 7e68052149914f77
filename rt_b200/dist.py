"""Multi-GPU composition: one process per GPU, sample-range (and optional row-band) partition, one
reduce of the fp32 accumulation buffers.

Two exchanges are implemented.  "peer" (default on NVLink): every rank renders into a buffer shared through CUDA IPC, then
ONE kernel per rank (rtcu_exchange_reduce_resolve) hand-shakes with the peers through release / acquire flags in device memory,
sums all ranks' buffers over its row band through NVLink peer loads, resolves, and stores the packed pixels straight into
rank 0's image -- no staging, no collective and no barrier on the frame path, deterministic sum order.  "nccl": reduce-scatter of the fp32 row bands, resolve per rank,
gather of the packed bands (also what the gloo CPU tests exercise through all-reduce).

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


def band_rows(height: int, world: int) -> int:
    """rows per rank when the frame is cut into `world` equal row bands (the last band may hang over the frame)"""
    return -(-height // world)


def sum_row_bands(accum_padded, rank: int, world: int, group=None):
    """Sum the ranks' accumulation buffers and leave each rank with *its* row band of the total: a reduce-scatter
    (NCCL; over NVSwitch every rank sends and receives (world-1)/world of the buffer, instead of rank 0 receiving all of
    it), or all-reduce + slice on backends without reduce-scatter (gloo, CPU tests).  accum_padded: (world*band, W, 4)."""
    import torch
    import torch.distributed as dist

    band = accum_padded.shape[0] // world
    if world == 1:
        return accum_padded[:band]
    if dist.get_backend(group) == "nccl":
        out = torch.empty_like(accum_padded[:band])
        dist.reduce_scatter_tensor(out, accum_padded, op=dist.ReduceOp.SUM, group=group)
        return out
    dist.all_reduce(accum_padded, op=dist.ReduceOp.SUM, group=group)
    return accum_padded[rank * band:(rank + 1) * band].clone()


def gather_bands(band_rgba8, height: int, rank: int, world: int, dst: int = 0, group=None):
    """Gather the packed row bands onto `dst` (4x fewer bytes than the fp32 buffers); returns (H, W) there, None elsewhere."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return band_rgba8[:height]
    parts = [torch.empty_like(band_rgba8) for _ in range(world)] if rank == dst else None
    dist.gather(band_rgba8, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat(parts, dim=0)[:height]


class _DeviceArray:
    """__cuda_array_interface__ over a raw device pointer, so torch can view library-owned memory without copying"""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


class GpuRank:
    """One rank's device state for bench.py / multi-GPU runs: a Context plus torch-owned frame buffers."""

    def __init__(self, ctx, width: int, height: int, device=None, world: int = 1):
        import torch

        self.torch = torch
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device) if device is None else device
        self.world = world
        self.band = band_rows(height, world)
        # the buffer is padded to world equal row bands (rows past the frame stay zero) for the reduce-scatter
        self.accum_padded = torch.zeros((self.band * world, width, 4), dtype=torch.float32, device=self.device)
        self.accum = self.accum_padded[:height]
        self.rgba8 = torch.zeros((height, width), dtype=torch.int32, device=self.device)  # bit pattern of uint32
        self.band_rgba8 = torch.zeros((self.band, width), dtype=torch.int32, device=self.device)
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

    def resolve_band(self, band_accum, spp: int):
        """divide / sqrt / pack this rank's summed row band (mg_ray_tracer.cpp:195-200 over band rows)"""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.resolve_device(band_accum.data_ptr(), self.width, self.band, spp, self.band_rgba8.data_ptr(), stream=stream)
        return self.band_rgba8

    # ---- "peer" exchange: IPC-shared buffers + the fused handshake / reduce / resolve kernel ---------------------
    def enable_peer_exchange(self, rank: int, dst: int = 0, group=None) -> None:
        """Allocates this rank's TWO accumulation buffers (frames alternate between them, see k_exchange_reduce_resolve), its
        128-byte flag block and, on `dst`, the packed image as CUDA-IPC-exportable memory in the library, swaps the handles
        between the ranks once (all_gather_object -- plumbing, not on the frame path) and maps the peers' memory (NVLink peer
        access)."""
        import torch.distributed as dist

        torch = self.torch
        npix = self.width * self.height
        self.peer_rank, self.peer_dst = rank, dst
        # every rank takes part in the handle swap even if its own allocation failed, so that nobody waits forever
        mine = {"accum": [None, None], "flags": None, "img": None, "err": None}
        try:
            own_accum = [self.ctx.ipc_alloc(npix * 16) for _ in range(2)]
            own_flags = self.ctx.ipc_alloc(128)
            mine["accum"] = [h for _, h in own_accum]
            mine["flags"] = own_flags[1]
            if rank == dst:
                self.peer_img, mine["img"] = self.ctx.ipc_alloc(npix * 4)
        except Exception as e:  # noqa: BLE001
            mine["err"] = str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        failed = [(g, h["err"]) for g, h in enumerate(handles) if h["err"] is not None or h["flags"] is None]
        if failed:
            raise RuntimeError(f"CUDA IPC allocation failed on rank(s) {failed}")
        # peer_accums[b][g]: buffer b of rank g, as this rank's device can address it
        self.peer_accums = [[own_accum[b][0] if g == rank else self.ctx.ipc_open(handles[g]["accum"][b]) for g in range(self.world)] for b in range(2)]
        self.peer_flags = [own_flags[0] if g == rank else self.ctx.ipc_open(handles[g]["flags"]) for g in range(self.world)]
        self.peer_accum = self.peer_accums[0][rank]  # (first buffer; kept for callers that render into it directly)
        if rank != dst:
            self.peer_img = self.ctx.ipc_open(handles[dst]["img"])
        self._epoch = 0
        self._peer_sync = torch.zeros(1, dtype=torch.int32, device=self.device)
        if rank == dst:  # the image as a torch tensor over the library's memory (for read-back through torch)
            self.peer_rgba8 = torch.as_tensor(_DeviceArray(self.peer_img, (self.height, self.width), "<i4"), device=self.device)

    def next_frame(self) -> tuple[int, int]:
        """(epoch, buffer index) of the next multi-GPU frame -- the same sequence on every rank"""
        self._epoch += 1
        return self._epoch, self._epoch & 1

    def exchange(self, epoch: int, buf: int, spp_total: int, stream: int) -> None:
        """the exchange step of frame `epoch` whose share this rank has just rendered into buffer `buf`: ONE launch"""
        row0, row1 = row_band_for_rank(0, self.height, self.peer_rank, self.world)
        self.ctx.exchange_reduce_resolve(self.peer_accums[buf], self.peer_flags, self.peer_rank, self.peer_dst, epoch, self.width, row0, row1 - row0,
                                         spp_total, self.peer_img, stream=stream)

    def render_peer_reduce_resolve(self, view: nat.View, spp_total: int, group=None):
        """One multi-GPU frame with neither a collective nor a barrier on the frame path: trace into this frame's IPC-shared
        buffer, then one kernel that hand-shakes with the peers through flags in device memory, sums their buffers over this
        rank's row band through NVLink peer loads, resolves and stores into dst's image.  Returns the (H, W) packed image on
        dst (complete when the stream reaches this point), None elsewhere."""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        epoch, buf = self.next_frame()
        self.ctx.render_device(view, self.peer_accums[buf][self.peer_rank], accumulate=False, stream=stream)
        self.exchange(epoch, buf, spp_total, stream)
        return self.peer_rgba8 if self.peer_rank == self.peer_dst else None

    def check_exchange(self) -> None:
        """raises if a rank failed to arrive at an exchange (synchronises the stream)"""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.exchange_check(self.peer_flags[self.peer_rank], stream=stream)

    def render_peer_barrier_reduce_resolve(self, view: nat.View, spp_total: int, group=None):
        """The round-1 form of the same exchange (kept for A/B timing): two stream-ordered one-element all-reduces as barriers
        around rtcu_reduce_resolve_rows."""
        import torch.distributed as dist

        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.render_device(view, self.peer_accums[0][self.peer_rank], accumulate=False, stream=stream)
        dist.all_reduce(self._peer_sync, group=group)  # stream-ordered barrier: every rank's buffer is complete
        row0, row1 = row_band_for_rank(0, self.height, self.peer_rank, self.world)
        self.ctx.reduce_resolve_rows(self.peer_accums[0], self.width, row0, row1 - row0, spp_total, self.peer_img, stream=stream)
        dist.all_reduce(self._peer_sync, group=group)  # every band is stored; the buffers may be overwritten again
        return self.peer_rgba8 if self.peer_rank == self.peer_dst else None

    def render_reduce_resolve(self, view: nat.View, rank: int, spp_total: int, dst: int = 0):
        """One multi-GPU frame: trace this rank's share, reduce-scatter the fp32 sums by row band, resolve the band here,
        gather the packed bands on `dst`.  Returns the (H, W) packed image on dst, None elsewhere."""
        self.render_accum(view)
        band = sum_row_bands(self.accum_padded, rank, self.world)
        packed = self.resolve_band(band, spp_total)
        return gather_bands(packed, self.height, rank, self.world, dst)
