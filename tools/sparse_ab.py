#!/usr/bin/env python3
"""How fast does ONE pixel's chain of samples advance when its warp carries fewer pixels?  Renders the N most expensive 8x4
blocks of the frame (pilot pass) with 32 / 16 / 8 / 4 pixel-carrying lanes per warp; N is chosen so that every variant is a
single round of the persistent grid, so the time of a launch is the time of its slowest pixel chain."""
import argparse, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--blocks", type=int, default=592)
ap.add_argument("--lanes", default="32,16,8,4")
ap.add_argument("--smem-nodes", type=int, default=0)
ap.add_argument("--bvh-width", type=int, default=0)
args = ap.parse_args()
import torch
dev = torch.device("cuda", 0)
w, h = 1920, 1080
pt = ptb200.PathTracer(0)
if args.bvh_width:
    pt.set_option(ptb200.PT_OPT_BVH_WIDTH, args.bvh_width)
pt.upload_scene(ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"))
pt.set_camera()
pt.set_params(args.spp, 10)
rgb = torch.zeros(w * h * 3, dtype=torch.uint8, device=dev)
pt.bind_framebuffer(rgb.data_ptr(), 0, w, h)
stream = torch.cuda.current_stream(dev)
bw, bh = (w + 7) // 8, (h + 3) // 4
costs = torch.zeros(bw * bh, dtype=torch.int32, device=dev)
pt.block_costs_async(4, costs.data_ptr(), stream.cuda_stream)
order = torch.argsort(costs, descending=True, stable=True)
mine = order[:args.blocks]
blocks = ((mine % bw) | ((mine // bw) << 16)).to(torch.int32).contiguous()
c = costs.cpu().numpy().astype(np.float64)
print(json.dumps({"blocks": args.blocks, "mean_cost_all": c.mean() / 128, "mean_cost_selected": float(costs[mine].double().mean()) / 128, "max_cost": c.max() / 128, "unit": "rays per sample (pilot, 4 spp)"}))
pt.set_option(ptb200.PT_OPT_SMEM_NODES, args.smem_nodes)
ref = None
for L in [int(x) for x in args.lanes.split(",")]:
    pt.set_option(ptb200.PT_OPT_LANES_PER_WARP, L)
    best = None
    for _ in range(2):
        rgb.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pt.render_blocks_async(blocks.data_ptr(), int(blocks.numel()), stream.cuda_stream)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    img = rgb.cpu().numpy().copy()
    if ref is None:
        ref = img
    print(json.dumps({"lanes_per_warp": L, "spp": args.spp, "ms": best, "ms_per_1024spp": best * 1024 / args.spp, "identical": bool(np.array_equal(img, ref))}), flush=True)
pt.close()
