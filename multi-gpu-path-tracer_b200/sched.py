"""One-process-per-GPU frame rendering: dynamic tile claims + NCCL framebuffer gather.

Replaces, for the multi-process launch that bench.py uses (torchrun, one rank per GPU), what
RenderManager + StreamThread + TaskGenerator do inside one process in the reference
(src/RenderManager.h:76-112,410-431; src/StreamThread.h:64-104; src/Scheduling/TaskGenerator.h:58-80):
there, one fixed rectangle per GPU per frame written into a cudaMallocManaged framebuffer; here, tiles
claimed from a node-wide atomic counter (capi.TileQueue, POSIX shared memory), rendered into each rank's
private device framebuffer and summed onto rank 0 with ONE collective at frame end (the claimed tile sets
are disjoint and unclaimed pixels are zero, so a uint8 SUM reduce is the gather).  The path has no other
exchange step: pixels are independent and the RNG is keyed by the global pixel index.

torch is plumbing here (device buffers, streams, torch.distributed); pixels come from libptcore.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from .capi import PathTracer, TileQueue

Tile = Tuple[int, int, int, int]  # offset_x, offset_y, width, height (bottom-up pixel space, as RenderTask)


def make_tiles(width: int, height: int, tile_w: int, tile_h: int) -> List[Tile]:
    return [(x, y, min(tile_w, width - x), min(tile_h, height - y)) for y in range(0, height, tile_h) for x in range(0, width, tile_w)]


def interleave(tiles: List[Tile], stride: int) -> List[Tile]:
    """Order tiles so that any contiguous run of `stride`-spaced claims samples the whole image
    (cheap black tiles and expensive duck tiles end up in every rank's share)."""
    if stride <= 1:
        return list(tiles)
    out = []
    for s in range(stride):
        out.extend(tiles[s::stride])
    return out


def lpt_levels(world: int) -> int:
    return 8 if world <= 2 else 4


def _z_order(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    def spread(v):
        v = v & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


def lpt_block_order(costs: torch.Tensor, bw: int, levels: int, z: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Order of the 8x4 blocks of a bw-wide grid for the LPT frame: most expensive cost CLASS first (class = cost * levels //
    (max cost + 1)), along the Z-order curve within a class.  A full sort by cost scatters neighbouring blocks over the launch;
    classes keep "long chains start first" and leave warps that run side by side on neighbouring pixels, i.e. on the same
    triangles: 728 -> 700 ms per 1080p / 1024 spp frame on one GPU (profiles/r02_block_order_ab.txt).  Same function as
    TaskGenerator::lptBlockOrder (csrc/host/TaskGenerator.h); no host synchronisation."""
    if z is None:
        i = torch.arange(costs.numel(), device=costs.device, dtype=torch.int64)
        z = _z_order(i % bw, i // bw)
    c = costs.to(torch.int64)
    cls = (c * levels) // (c.max() + 1)
    return torch.argsort(((levels - 1 - cls) << 32) | z)  # keys are unique


@dataclass
class FramePlan:
    width: int
    height: int
    tiles: List[Tile]
    claim: int  # tiles per claim


class RankRenderer:
    """Per-rank state: tracer handle, private device framebuffer (RGB8 + I420 in ONE allocation, so the gather is one
    collective), two launch streams, per-stage CUDA events of the last frames."""

    def __init__(self, pt: PathTracer, width: int, height: int, device: torch.device, n_streams: int = 2):
        self.pt = pt
        self.device = device
        self.width, self.height = width, height
        n_rgb, n_yuv = width * height * 3, width * height * 3 // 2
        self.fb = torch.zeros((n_rgb + n_yuv + 15) // 16 * 16, dtype=torch.uint8, device=device)
        self.rgb, self.yuv = self.fb[:n_rgb], self.fb[n_rgb:n_rgb + n_yuv]
        self.fb_words = self.fb.view(torch.int32)  # what travels: see _gather
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
        pt.bind_framebuffer(self.rgb.data_ptr(), self.yuv.data_ptr(), width, height)
        self.launches = 0
        self.stage_events = []  # one list of (name, start event, end event) per LPT frame since the last reset_stage_times()

    def render_frame(self, plan: FramePlan, queue: Optional[TileQueue], rank: int, world: int, gather: bool = True) -> None:
        """Renders this rank's dynamically claimed share of the frame; after it returns, rank 0's
        self.rgb / self.yuv hold the whole frame (when gather=True and world > 1)."""
        cur = torch.cuda.current_stream(self.device)
        if world > 1:
            self.fb.zero_()
        for s in self.streams:
            s.wait_stream(cur)
        n = len(plan.tiles)
        i = 0
        if world == 1 or queue is None:
            self.pt.render_tiles_async(plan.tiles, self.streams[0].cuda_stream)
            self.launches += (n + 47) // 48
        else:
            while True:
                first = queue.claim(plan.claim, n)
                if first < 0:
                    break
                chunk = plan.tiles[first:first + plan.claim]
                self.pt.render_tiles_async(chunk, self.streams[i % len(self.streams)].cuda_stream)
                self.launches += (len(chunk) + 47) // 48
                i += 1
        for s in self.streams:
            cur.wait_stream(s)
        if world > 1 and gather:
            self._gather()

    def _gather(self) -> None:
        import torch.distributed as dist
        # claimed pixel sets are disjoint and the rest is zero, so a SUM is the gather; summed as 32-bit words (every byte is non-zero on
        # at most one rank: no carries), which NCCL reduces several times faster than 8-bit elements (15.6 -> ms at 8 ranks, r02_scale_n8)
        dist.reduce(self.fb_words, dst=0, op=dist.ReduceOp.SUM)

    def reset_stage_times(self) -> None:
        self.stage_events = []

    def stage_times_ms(self) -> dict:
        """Mean device milliseconds per stage (pilot, sort, render, gather) over the frames since reset_stage_times()."""
        if not self.stage_events:
            return {}
        torch.cuda.synchronize(self.device)
        acc = {}
        for frame in self.stage_events:
            for name, e0, e1 in frame:
                acc.setdefault(name, []).append(e0.elapsed_time(e1))
        return {k: sum(v) / len(v) for k, v in acc.items()}

    def render_frame_lpt(self, rank: int, world: int, pilot_spp: int = 4, gather: bool = True, emulated: bool = False) -> None:
        """Cost-sorted block scheduling (longest processing time first).

        Every pixel is one sequential chain of spp samples (its XORWOW stream), so the unit of work cannot be split and
        a frame ends when the last chain ends.  A pilot pass (pilot_spp samples per pixel, a few per mille of the frame)
        measures rays per 8x4 block — each rank traces 1/world of the blocks and one all-reduce (SUM of a 260 KB map)
        gives every rank the whole map; blocks are ordered by that cost (lpt_block_order), dealt round-robin to the ranks (equal cost per
        GPU: every rank computes the same order) and each GPU's persistent kernel takes its blocks most-expensive-first,
        which keeps the tail of the frame short.  `emulated`: this process stands in for rank `rank` of `world` on one
        GPU (experiments): the pilot covers the whole frame and nothing is gathered.
        """
        import torch.distributed as dist
        cur = torch.cuda.current_stream(self.device)
        s = self.streams[0]
        s.wait_stream(cur)
        bw, bh = (self.width + 7) // 8, (self.height + 3) // 4
        n = bw * bh
        if getattr(self, "costs", None) is None or self.costs.numel() != n:
            self.costs = torch.zeros(n, dtype=torch.int32, device=self.device)
            i = torch.arange(n, device=self.device, dtype=torch.int64)
            self.z_order = _z_order(i % bw, i // bw)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        distributed = world > 1 and not emulated
        with torch.cuda.stream(s):
            ev[0].record(s)
            if world > 1:
                self.fb.zero_()
            if distributed:
                per = (n + world - 1) // world
                self.pt.block_costs_range_async(pilot_spp, self.costs.data_ptr(), rank * per, per, s.cuda_stream)
                dist.all_reduce(self.costs, op=dist.ReduceOp.SUM)  # NCCL enqueues on the current stream (s)
            else:
                self.pt.block_costs_async(pilot_spp, self.costs.data_ptr(), s.cuda_stream)
            ev[1].record(s)
            order = lpt_block_order(self.costs, bw, lpt_levels(world), self.z_order)
            mine = order[rank::world]
            self.blocks = ((mine % bw) | ((mine // bw) << 16)).to(torch.int32).contiguous()  # kept alive until the next frame
            ev[2].record(s)
            self.pt.render_blocks_async(self.blocks.data_ptr(), int(self.blocks.numel()), s.cuda_stream)
            ev[3].record(s)
            if distributed and gather:
                self._gather()
            ev[4].record(s)
        self.launches += 2
        self.stage_events.append([("pilot", ev[0], ev[1]), ("sort", ev[1], ev[2]), ("render", ev[2], ev[3]), ("gather", ev[3], ev[4])])
        cur.wait_stream(s)

    def render_frame_keyed(self, rank: int, world: int, n_chunks: int = 16, pilot_spp: int = 4, gather: bool = True, emulated: bool = False) -> None:
        """PT_RNG_SAMPLE_KEYED frame (capi.PT_RNG_SAMPLE_KEYED; parity is statistical in this mode, so it is never the bench line).

        The stream is keyed by (pixel, sample), so a pixel's spp samples are cut into n_chunks independent work items.  Rank r
        traces chunks r, r + world, ... of EVERY pixel — statistically identical shares, no pilot pass needed for balance between
        ranks; within a GPU the blocks are taken most-expensive-first (same cost map as the stream mode).  The chunk sums travel as
        one float reduce (every entry is written by exactly one rank: the sum is exact) and rank 0 adds the chunks of every pixel in
        chunk order: the image does not depend on world, on n_chunks' assignment to ranks or on launch order.
        """
        import torch.distributed as dist
        cur = torch.cuda.current_stream(self.device)
        s = self.streams[0]
        s.wait_stream(cur)
        bw, bh = (self.width + 7) // 8, (self.height + 3) // 4
        n = bw * bh
        npix = self.width * self.height
        if getattr(self, "costs", None) is None or self.costs.numel() != n:
            self.costs = torch.zeros(n, dtype=torch.int32, device=self.device)
            i = torch.arange(n, device=self.device, dtype=torch.int64)
            self.z_order = _z_order(i % bw, i // bw)
        if getattr(self, "accum", None) is None or self.accum.numel() != n_chunks * npix * 3:
            self.accum = torch.zeros(n_chunks * npix * 3, dtype=torch.float32, device=self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        distributed = world > 1 and not emulated
        with torch.cuda.stream(s):
            ev[0].record(s)
            self.accum.zero_()
            if distributed:
                per = (n + world - 1) // world
                self.pt.block_costs_range_async(pilot_spp, self.costs.data_ptr(), rank * per, per, s.cuda_stream)
                dist.all_reduce(self.costs, op=dist.ReduceOp.SUM)
            else:
                self.pt.block_costs_async(pilot_spp, self.costs.data_ptr(), s.cuda_stream)
            ev[1].record(s)
            order = lpt_block_order(self.costs, bw, lpt_levels(1), self.z_order)
            self.blocks = ((order % bw) | ((order // bw) << 16)).to(torch.int32).contiguous()
            ev[2].record(s)
            self.pt.render_keyed_async(self.accum.data_ptr(), n_chunks, rank, world, self.blocks.data_ptr(), int(self.blocks.numel()), s.cuda_stream)
            ev[3].record(s)
            if distributed and gather:
                dist.reduce(self.accum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0 or not distributed:
                self.pt.resolve_keyed_async(self.accum.data_ptr(), n_chunks, s.cuda_stream)
            ev[4].record(s)
        self.launches += 3
        self.stage_events.append([("pilot", ev[0], ev[1]), ("sort", ev[1], ev[2]), ("render", ev[2], ev[3]), ("gather", ev[3], ev[4])])
        cur.wait_stream(s)
