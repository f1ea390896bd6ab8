#!/usr/bin/env python3
"""Scheduling knobs of the wavefront kernel on one frame (1080p cornell_duck): tools/knob_sweep.py [spp] [refill:burst,...] [nosmem]"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200, torch  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w, h = 1920, 1080
pt = ptb200.PathTracer(0)
pt.upload_scene(ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"))
pt.set_camera(); pt.set_params(spp, 10)
rgb = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
pt.bind_framebuffer(rgb.data_ptr(), 0, w, h)
smem = 0 if 'nosmem' in sys.argv[3:] else 1
pt.set_option(ptb200.PT_OPT_SMEM_NODES, smem)
ref = None
for refill, burst in [tuple(int(x) for x in a.split(':')) for a in (sys.argv[2].split(',') if len(sys.argv) > 2 else '24:2,20:2,28:2,16:2,24:3,24:1,30:2'.split(','))]:
    pt.set_option(ptb200.PT_OPT_REFILL_AT, refill); pt.set_option(ptb200.PT_OPT_NODE_BURST, burst)
    best = 1e9
    for _ in range(3):
        rgb.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pt.render_tile_async(0, 0, w, h); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    img = rgb.cpu().numpy().copy()
    if ref is None:
        ref = img
    print(json.dumps({"smem_nodes": smem, "refill_at": refill, "node_burst": burst, "ms": round(best, 2), "msamples_per_s": round(w * h * spp / best / 1e3, 1), "identical": bool(np.array_equal(img, ref))}), flush=True)
pt.close()
