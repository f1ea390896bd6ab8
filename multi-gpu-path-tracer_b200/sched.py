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


@dataclass
class FramePlan:
    width: int
    height: int
    tiles: List[Tile]
    claim: int  # tiles per claim


class RankRenderer:
    """Per-rank state: tracer handle, private device framebuffer, two launch streams."""

    def __init__(self, pt: PathTracer, width: int, height: int, device: torch.device, n_streams: int = 2):
        self.pt = pt
        self.device = device
        self.width, self.height = width, height
        self.rgb = torch.zeros(width * height * 3, dtype=torch.uint8, device=device)
        self.yuv = torch.zeros(width * height * 3 // 2, dtype=torch.uint8, device=device)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
        pt.bind_framebuffer(self.rgb.data_ptr(), self.yuv.data_ptr(), width, height)
        self.launches = 0

    def render_frame(self, plan: FramePlan, queue: Optional[TileQueue], rank: int, world: int, gather: bool = True) -> None:
        """Renders this rank's dynamically claimed share of the frame; after it returns, rank 0's
        self.rgb / self.yuv hold the whole frame (when gather=True and world > 1)."""
        cur = torch.cuda.current_stream(self.device)
        if world > 1:
            self.rgb.zero_()
            self.yuv.zero_()
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
        dist.reduce(self.rgb, dst=0, op=dist.ReduceOp.SUM)
        dist.reduce(self.yuv, dst=0, op=dist.ReduceOp.SUM)

    def render_frame_lpt(self, rank: int, world: int, pilot_spp: int = 4, gather: bool = True) -> None:
        """Cost-sorted block scheduling (longest processing time first).

        Every pixel is one sequential chain of spp samples (its XORWOW stream), so the unit of work cannot be split and
        a frame ends when the last chain ends.  A pilot pass (pilot_spp samples of every pixel, a few per mille of the
        frame) measures rays per 8x4 block; blocks are sorted by that cost, dealt round-robin to the ranks (equal cost
        per GPU without any exchange: every rank computes the same map and the same order) and each GPU's persistent
        kernel takes its blocks most-expensive-first, which keeps the tail of the frame short.
        """
        cur = torch.cuda.current_stream(self.device)
        s = self.streams[0]
        s.wait_stream(cur)
        bw, bh = (self.width + 7) // 8, (self.height + 3) // 4
        if getattr(self, "costs", None) is None or self.costs.numel() != bw * bh:
            self.costs = torch.zeros(bw * bh, dtype=torch.int32, device=self.device)
        with torch.cuda.stream(s):
            if world > 1:
                self.rgb.zero_()
                self.yuv.zero_()
            self.pt.block_costs_async(pilot_spp, self.costs.data_ptr(), s.cuda_stream)
            order = torch.argsort(self.costs, descending=True, stable=True)
            mine = order[rank::world]
            self.blocks = ((mine % bw) | ((mine // bw) << 16)).to(torch.int32).contiguous()  # kept alive until the next frame
            self.pt.render_blocks_async(self.blocks.data_ptr(), int(self.blocks.numel()), s.cuda_stream)
        self.launches += 2
        cur.wait_stream(s)
        if world > 1 and gather:
            self._gather()
