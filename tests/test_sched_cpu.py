"""Host-side multi-rank logic on CPU (gloo, world_size 2): dynamic tile claims from the node-wide counter, disjoint
private framebuffers, one SUM reduce as the gather.  Pixels come from the oracle here (no GPU in this container);
the same claim/gather code drives the CUDA tracer in bench.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLD, ROOT


def test_make_tiles_cover_the_frame_exactly(ptb):
    sched = ptb.sched
    for (w, h, tw, th) in [(1920, 1080, 64, 32), (100, 50, 64, 32), (37, 23, 8, 4), (8, 4, 64, 32)]:
        cover = np.zeros((h, w), np.int32)
        for (x, y, cw, ch) in sched.make_tiles(w, h, tw, th):
            assert cw > 0 and ch > 0
            cover[y:y + ch, x:x + cw] += 1
        assert (cover == 1).all()
    tiles = sched.make_tiles(256, 128, 32, 32)
    inter = sched.interleave(tiles, 8)
    assert sorted(inter) == sorted(tiles) and inter != tiles
    assert sched.interleave(tiles, 1) == tiles


def test_tile_queue_claims_are_disjoint_and_bounded(ptb, core_lib):
    name = f"/ptb200_test_{os.getpid()}"
    q = ptb.TileQueue(name, create=True)
    q2 = ptb.TileQueue(name, create=False)
    got = []
    while True:
        a = q.claim(3, 20)
        if a >= 0:
            got.extend(range(a, min(a + 3, 20)))
        b = q2.claim(2, 20)
        if b >= 0:
            got.extend(range(b, min(b + 2, 20)))
        if a < 0 and b < 0:
            break
    assert sorted(got) == list(range(20))
    q.reset()
    assert q2.claim(1, 20) == 0
    q2.close()
    q.close()


def _worker(rank, world, port, qname, out_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import _oracle
    import ptb200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sched = ptb200.sched
    W, H, SPP, DEPTH = 48, 27, 2, 4
    scene = ptb200.load_scene_file(GOLD / "cornell_duck.ptscene.gz")
    orc = _oracle.load()
    world_h = orc.world(scene)
    tiles = sched.interleave(sched.make_tiles(W, H, 16, 8), world * 2)
    if rank == 0:
        q = ptb200.TileQueue(qname, create=True)
    dist.barrier()
    if rank != 0:
        q = ptb200.TileQueue(qname, create=False)
    dist.barrier()
    rgb = np.zeros((H, W, 3), np.uint8)
    mine = 0
    while True:
        first = q.claim(2, len(tiles))
        if first < 0:
            break
        for (x, y, w, h) in tiles[first:first + 2]:
            part, _, _ = orc.render(world_h, W, H, SPP, DEPTH, rect=(x, y, w, h), threads=1)
            rgb |= part  # untouched pixels are zero in `part`
            mine += 1
    t = torch.from_numpy(rgb.reshape(-1).copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    counts = torch.tensor([mine])
    dist.all_reduce(counts)
    if rank == 0:
        full, _, _ = orc.render(world_h, W, H, SPP, DEPTH, threads=1)
        np.save(out_path, np.stack([t.numpy().reshape(H, W, 3), full]))
        assert int(counts[0]) == len(tiles)
    dist.barrier()
    q.close()
    dist.destroy_process_group()


def test_two_ranks_dynamic_claims_and_reduce_gather_equal_single_render(core_lib, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "frames.npy"
    qname = f"/ptb200_test_mp_{os.getpid()}"
    mp.spawn(_worker, args=(2, port, qname, str(out)), nprocs=2, join=True)
    gathered, full = np.load(out)
    assert np.array_equal(gathered, full)


def _lpt_worker(rank, world, port, out_path):
    """The N > 1 frame of sched.RankRenderer.render_frame_lpt, step by step, with the CPU oracle standing in for the kernels: pilot cost of
    this rank's 1/N of the blocks (rays of one pilot sample per pixel), all-reduce of the cost map, sched.lpt_block_order on every rank,
    round-robin deal, this rank's blocks rendered, ONE reduce of the fused RGB + I420 frame summed as 32-bit words."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import _oracle
    import ptb200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sched = ptb200.sched
    W, H, SPP, DEPTH = 52, 30, 2, 4  # ragged: 7 x 8 blocks of 8x4, the last column 4 wide, the last row 2 high
    scene = ptb200.load_scene_file(GOLD / "cornell_duck.ptscene.gz")
    orc = _oracle.load()
    world_h = orc.world(scene)
    bw, bh = (W + 7) // 8, (H + 3) // 4
    n = bw * bh

    def block_rect(b):
        x, y = (b % bw) * 8, (b // bw) * 4
        return (x, y, min(8, W - x), min(4, H - y))
    per = (n + world - 1) // world
    costs = torch.zeros(n, dtype=torch.int32)
    for b in range(rank * per, min(n, (rank + 1) * per)):
        costs[b] = int(orc.render(world_h, W, H, 1, DEPTH, rect=block_rect(b), threads=1)[2]["rays"])
    dist.all_reduce(costs, op=dist.ReduceOp.SUM)
    order = sched.lpt_block_order(costs, bw, sched.lpt_levels(world))
    mine = order[rank::world].tolist()
    n_rgb, n_yuv = W * H * 3, W * H * 3 // 2
    fb = np.zeros((n_rgb + n_yuv + 15) // 16 * 16, np.uint8)
    for b in mine:
        rgb, yuv, _ = orc.render(world_h, W, H, SPP, DEPTH, rect=block_rect(b), threads=1)
        fb[:n_rgb] |= rgb.reshape(-1)
        fb[n_rgb:n_rgb + n_yuv] |= yuv
    words = torch.from_numpy(fb).view(torch.int32)
    dist.reduce(words, dst=0, op=dist.ReduceOp.SUM)
    share = torch.tensor([float(costs[order[rank::world]].sum()), float(len(mine))])
    shares = [torch.zeros(2) for _ in range(world)]
    dist.all_gather(shares, share)
    if rank == 0:
        full_rgb, full_yuv, _ = orc.render(world_h, W, H, SPP, DEPTH, threads=1)
        got = words.view(torch.uint8).numpy()
        np.save(out_path, np.array([np.array_equal(got[:n_rgb], full_rgb.reshape(-1)), np.array_equal(got[n_rgb:n_rgb + n_yuv], full_yuv),
                                    sum(int(s[1]) for s in shares) == n, max(float(s[0]) for s in shares) <= 1.25 * float(costs.sum()) / world]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_lpt_frame_sharded_pilot_class_order_and_word_reduce_equal_single_render(core_lib, tmp_path):
    """gloo, world size 2: the multi-GPU frame's host logic (sharded pilot + all-reduce, cost-class / Z-order block order, round-robin
    deal, one 32-bit-word SUM reduce of the fused RGB + I420 frame) reproduces the single render byte for byte, every block is rendered
    exactly once and the two shares of the pilot cost are balanced."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "lpt.npy"
    mp.spawn(_lpt_worker, args=(2, port, str(out)), nprocs=2, join=True)
    rgb_ok, yuv_ok, all_blocks, balanced = np.load(out)
    assert rgb_ok and yuv_ok and all_blocks and balanced
