#!/usr/bin/env python3
"""GPU box: A/B of the two node formats (64-byte float planes / 32-byte quantised planes) on four scenes: time, box and
triangle tests per ray, leaf-box inflation, and whether the images are identical.   usage: tools/node_format_ab.py"""
import sys, json, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200
dev = torch.device("cuda", 0)
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
def run(name, scene, w, h, spp, depth, cam=None):
    pt = ptb200.PathTracer(0)
    pt.upload_scene(scene); pt.set_camera(**(cam or {})); pt.set_params(spp, depth)
    rr = ptb200.sched.RankRenderer(pt, w, h, dev)
    ref = None
    for fmt in (1, 2):
        pt.set_option(ptb200.PT_OPT_NODE_FORMAT, fmt)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rr.render_frame_lpt(0, 1); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        img = rr.rgb.clone()
        same = True if ref is None else bool(torch.equal(ref, img))
        ref = img if ref is None else ref
        pt.set_option(ptb200.PT_OPT_COUNT_TESTS, 1); pt.set_params(2, depth); pt.reset_stats(); pt.render_tiles_async([(0, 0, w, h)]); c = pt.stats()
        pt.set_option(ptb200.PT_OPT_COUNT_TESTS, 0); pt.set_params(spp, depth)
        print(name, "fmt", fmt, "ms %.2f" % best, "Msamples/s %.1f" % (w * h * spp / best / 1e3), "box/ray %.2f tri/ray %.2f" % (c["box_tests"] / c["rays"], c["tri_tests"] / c["rays"]),
              "inflation %.3f" % c["quant_inflation"], "same_image", same, flush=True)
    pt.close()
run("duck1080p_s128", duck, 1920, 1080, 128, 10)
sf, cam = ptb200.scenes.rtow_sphere_field()
run("rtow1080p_s64", sf, 1920, 1080, 64, 10, cam)
mesh = ptb200.scenes.displaced_sphere_in_cornell(duck, n=1000)
run("mesh2M_4k_s16", mesh, 3840, 2160, 16, 10)
mesh = ptb200.scenes.displaced_sphere_in_cornell(duck, n=300)
run("mesh180k_4k_s16", mesh, 3840, 2160, 16, 10)
