#!/usr/bin/env python3
"""Block order of the LPT frame on one GPU (1080p cornell_duck): full cost sort vs cost classes that keep the spatial order
within a class vs plain spatial order.  tools/order_ab.py [spp] [world]   (world > 1: this GPU stands in for rank 0 of `world`)"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200, torch  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w, h = 1920, 1080
pt = ptb200.PathTracer(0)
pt.upload_scene(ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"))
pt.set_camera(); pt.set_params(spp, 10)
rgb = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
pt.bind_framebuffer(rgb.data_ptr(), 0, w, h)
bw, bh = (w + 7) // 8, (h + 3) // 4
costs = torch.zeros(bw * bh, dtype=torch.int32, device="cuda")
pt.block_costs_async(4, costs.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
cmax = int(costs.max().item())


def morton(ix, iy):
    def part(v):
        v = v & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return part(ix) | (part(iy) << 1)


idx = torch.arange(bw * bh, device="cuda")
mort = morton(idx % bw, idx // bw)
orders = {"sorted": torch.argsort(costs, descending=True, stable=True), "spatial": idx}
for levels in (4, 8, 16, 32, 64):
    cls = (costs.to(torch.int64) * levels) // (cmax + 1)
    orders[f"classes{levels}"] = torch.argsort(cls, descending=True, stable=True)
    key = cls * (1 << 32) + ((1 << 32) - 1 - mort)
    orders[f"classes{levels}_morton"] = torch.argsort(key, descending=True, stable=True)
ref = None
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, order in orders.items():
    if only and name not in only:
        continue
    per_rank, share = [], []
    for rank in range(world):
        mine = order[rank::world]
        share.append(float(costs[mine].sum().item()))
        blocks = ((mine % bw) | ((mine // bw) << 16)).to(torch.int32).contiguous()
        best = 1e9
        for _ in range(2):
            rgb.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pt.render_blocks_async(blocks.data_ptr(), int(blocks.numel()), torch.cuda.current_stream().cuda_stream); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        per_rank.append(round(best, 2))
    img = rgb.cpu().numpy().copy()
    if ref is None and world == 1:
        ref = img
    ms = max(per_rank)
    print(json.dumps({"order": name, "world": world, "ms": ms, "msamples_per_s_whole_job": round(w * h * spp / ms / 1e3, 1), "per_rank_ms": per_rank,
                      "pilot_cost_imbalance": round(max(share) / (sum(share) / world), 4), "identical": bool(world > 1 or np.array_equal(img, ref))}), flush=True)
pt.close()
