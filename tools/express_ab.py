#!/usr/bin/env python3
"""Express lane for the longest chains of an N-GPU frame, emulated on one GPU (1080p cornell_duck, 1024 spp): every rank's share is rendered
one after the other, either in ONE launch (baseline) or as an express launch — the first n_sms * warps blocks of the rank's list, one block
per warp, on n_sms small CTAs — beside a main launch on the other SMs.   tools/express_ab.py [world] [spp] [n_sms:warps[:refill_at[:node_burst]],...]   (n_sms 0 = ONE launch with `warps` warps per SM, 0 = the core's rule)"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200, torch  # noqa: E402
from ptb200 import sched  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
configs = [(0, 0, 0, 2)] + [(tuple(int(x) for x in a.split(":")) + (0, 2))[:4] for a in (sys.argv[3] if len(sys.argv) > 3 else "8:4,16:4,24:4,16:8").split(",")]  # n_sms:warps[:refill_at[:node_burst]]
w, h = 1920, 1080
sms = torch.cuda.get_device_properties(0).multi_processor_count
pt = ptb200.PathTracer(0)
pt.upload_scene(ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"))
pt.set_camera(); pt.set_params(spp, 10)
rgb = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
pt.bind_framebuffer(rgb.data_ptr(), 0, w, h)
bw, bh = (w + 7) // 8, (h + 3) // 4
costs = torch.zeros(bw * bh, dtype=torch.int32, device="cuda")
main, side = torch.cuda.current_stream(), torch.cuda.Stream()
pt.block_costs_async(4, costs.data_ptr(), main.cuda_stream)
order = sched.lpt_block_order(costs, bw, sched.lpt_levels(world))
packed = ((order % bw) | ((order // bw) << 16)).to(torch.int32)
torch.cuda.synchronize()
ref = None
for n_sms, warps, refill, burst in configs:
    pt.set_option(ptb200.PT_OPT_REFILL_AT, refill); pt.set_option(ptb200.PT_OPT_NODE_BURST, burst)
    rgb.zero_()
    per_rank, express_ms = [], []
    for rank in range(world):
        mine = packed[rank::world].contiguous()
        best, best_x = 1e9, 0.0
        for _ in range(2):
            torch.cuda.synchronize()
            e0, e1, x0, x1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record(main)
            if n_sms:
                n_x = min(n_sms * warps, mine.numel())
                ex, rest = mine[:n_x].contiguous(), mine[n_x:].contiguous()
                side.wait_stream(main)
                pt.set_option(ptb200.PT_OPT_GRID_CTAS, n_sms); pt.set_option(ptb200.PT_OPT_CTA_WARPS, warps)
                x0.record(side); pt.render_blocks_async(ex.data_ptr(), int(ex.numel()), side.cuda_stream); x1.record(side)
                pt.set_option(ptb200.PT_OPT_GRID_CTAS, sms - n_sms); pt.set_option(ptb200.PT_OPT_CTA_WARPS, 0)
                pt.render_blocks_async(rest.data_ptr(), int(rest.numel()), main.cuda_stream)
                main.wait_stream(side)
            else:
                pt.set_option(ptb200.PT_OPT_GRID_CTAS, 0); pt.set_option(ptb200.PT_OPT_CTA_WARPS, warps)  # 0:W = ONE launch with W warps per SM
                pt.render_blocks_async(mine.data_ptr(), int(mine.numel()), main.cuda_stream)
            e1.record(main); e1.synchronize()
            t = e0.elapsed_time(e1)
            if t < best:
                best, best_x = t, (x0.elapsed_time(x1) if n_sms else 0.0)
        per_rank.append(round(best, 2)); express_ms.append(round(best_x, 2))
    img = rgb.cpu().numpy().copy()
    if ref is None:
        ref = img
    ms = max(per_rank)
    print(json.dumps({"express_sms": n_sms, "express_warps_per_sm": warps, "refill_at": refill, "node_burst": burst, "world": world, "ms": ms, "msamples_per_s_whole_job": round(w * h * spp / ms / 1e3, 1),
                      "per_rank_ms": per_rank, "express_launch_ms": express_ms, "identical": bool(np.array_equal(img, ref))}), flush=True)
pt.close()
