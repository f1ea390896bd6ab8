#!/usr/bin/env python3
"""GPU box: renders cornell_duck 1080p with both node formats and the thread-per-pixel kernel, lists pixels that differ and,
for each, the first ray whose closest hit differs (per-ray event log of ptcore_debug_trace_pixel).  This is how the exact-t
ties at shared triangle edges were found.   usage: tools/tie_probe.py"""
import sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200
dev = torch.device("cuda", 0)
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
W, H, SPP, D = 1920, 1080, 128, 10
pt = ptb200.PathTracer(0)
pt.upload_scene(duck); pt.set_camera(); pt.set_params(SPP, D)
rr = ptb200.sched.RankRenderer(pt, W, H, dev)
imgs = {}
for fmt in (1, 2):
    pt.set_option(ptb200.PT_OPT_NODE_FORMAT, fmt)
    rr.render_frame_lpt(0, 1); torch.cuda.synchronize()
    imgs[fmt] = rr.rgb.cpu().numpy().reshape(H, W, 3).astype(np.int32)
# direct kernel as the third opinion
pt.set_option(ptb200.PT_OPT_KERNEL, ptb200.PT_KERNEL_DIRECT)
pt.set_option(ptb200.PT_OPT_NODE_FORMAT, 1)
rgb, _ = pt.render_frame_host(W, H)
imgs[0] = np.asarray(rgb).reshape(H, W, 3).astype(np.int32)
pt.set_option(ptb200.PT_OPT_KERNEL, ptb200.PT_KERNEL_PERSISTENT)
pt.bind_framebuffer(rr.rgb.data_ptr(), rr.yuv.data_ptr(), W, H)
d12 = np.abs(imgs[1] - imgs[2]).max(axis=2)
d10 = np.abs(imgs[1] - imgs[0]).max(axis=2)
d20 = np.abs(imgs[2] - imgs[0]).max(axis=2)
print("pixels differing fmt1 vs fmt2:", int((d12 > 0).sum()), "max", int(d12.max()), "| fmt1 vs direct:", int((d10 > 0).sum()), "| fmt2 vs direct:", int((d20 > 0).sum()))
ys, xs = np.nonzero(d12)
for k in range(min(4, len(ys))):
    row, x = int(ys[k]), int(xs[k])   # image row (top-down) -> pixel y bottom-up
    y = H - 1 - row
    print("pixel x", x, "y", y, "fmt1", imgs[1][row, x], "fmt2", imgs[2][row, x])
    evs = {}
    for fmt in (1, 2):
        pt.set_option(ptb200.PT_OPT_NODE_FORMAT, fmt)
        ev, col = pt.trace_pixel(W, H, x, y, 16384)
        evs[fmt] = ev
        print("  fmt", fmt, "events", len(ev), "col", col)
    n = min(len(evs[1]), len(evs[2]))
    for i in range(n):
        if not np.array_equal(evs[1][i].view(np.uint32), evs[2][i].view(np.uint32)):
            print("  first divergent event", i)
            for fmt in (1, 2):
                e = evs[fmt][i]
                print("   fmt", fmt, "sample", e[0], "bounce", e[1], "prim", e[2], "t", repr(float(e[3])), "u", repr(float(e[4])), "v", repr(float(e[5])), "o", e[6:9], "d", e[9:12])
            break
