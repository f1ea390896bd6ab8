#!/usr/bin/env python3
"""GPU box: sweep of the two BVH build knobs (leaf size, SAH cost of a primitive test) on cornell_duck 1080p / 128 spp and on the
180 K-triangle mesh at 4K / 16 spp.   usage: tools/bvh_sweep.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200
dev = torch.device("cuda", 0)
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
mesh = ptb200.scenes.displaced_sphere_in_cornell(duck, n=300)
for name, scene, w, h, spp in (("duck1080p_s128", duck, 1920, 1080, 128), ("mesh180k_4k_s16", mesh, 3840, 2160, 16)):
    for leaf, cost in ((4, 120), (4, 80), (6, 120), (6, 80), (8, 120), (8, 80), (8, 50), (2, 120)):
        pt = ptb200.PathTracer(0)
        pt.set_option(ptb200.PT_OPT_BVH_LEAF_MAX, leaf)
        pt.set_option(ptb200.PT_OPT_SAH_INTERSECT_COST, cost)
        pt.upload_scene(scene); pt.set_camera(); pt.set_params(spp, 10)
        rr = ptb200.sched.RankRenderer(pt, w, h, dev)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rr.render_frame_lpt(0, 1); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        st = pt.stats()
        print(f"{name} leaf_max {leaf} isect_cost {cost / 100:.2f}: {best:.2f} ms  {w * h * spp / best / 1e3:.1f} Msamples/s  nodes {st['bvh_nodes']} depth {st['bvh_depth']} inflation {st['quant_inflation']:.3f}", flush=True)
        pt.close()
