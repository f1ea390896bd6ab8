#!/usr/bin/env python3
"""GPU box: the BASELINE configs other than the bench line (parity-test cases measured for the record).
Writes gpurun_out/configs.json.   usage: tools/bench_configs.py [--quick] [--only config4] [--refill 24]"""
import json, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200

quick = "--quick" in sys.argv
refill = int(sys.argv[sys.argv.index("--refill") + 1]) if "--refill" in sys.argv else 0
cta_warps = int(sys.argv[sys.argv.index("--cta-warps") + 1]) if "--cta-warps" in sys.argv else 0
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
dev = torch.device("cuda", 0)
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
out = {}


def run(name, scene, w, h, spp, depth, cam=None, reps=2):
    if only and only not in name:
        return
    pt = ptb200.PathTracer(0)
    pt.set_option(ptb200.PT_OPT_REFILL_AT, refill)
    pt.set_option(ptb200.PT_OPT_CTA_WARPS, cta_warps)
    t0 = time.perf_counter(); pt.upload_scene(scene); up = time.perf_counter() - t0
    pt.set_camera(**(cam or {})); pt.set_params(spp, depth)
    rr = ptb200.sched.RankRenderer(pt, w, h, dev)
    best = 1e9
    for _ in range(reps + 1):
        pt.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rr.render_frame_lpt(0, 1); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = pt.stats()
    pt.set_option(ptb200.PT_OPT_COUNT_TESTS, 1); pt.set_params(2, depth); pt.reset_stats(); pt.render_tiles_async([(0, 0, w, h)]); c = pt.stats()
    samples = w * h * spp
    out[name] = dict(width=w, height=h, spp=spp, depth=depth, ms=best, msamples_per_s=samples / best / 1e3, mrays_per_s=st["rays"] / best / 1e3, rays_per_sample=st["rays"] / samples,
                     box_per_ray=c["box_tests"] / c["rays"], tri_per_ray=c["tri_tests"] / c["rays"], bvh_nodes=st["bvh_nodes"], bvh_depth=st["bvh_depth"], scene_upload_s=up,
                     bvh_build_ms=st["bvh_build_ms"], scene_mb=st["scene_bytes"] / 1e6, mean_rgb=float(rr.rgb.float().mean()))
    print(name, json.dumps(out[name]), flush=True)
    pt.close()


run("config1_duck_640x360_s64_d8", duck, 640, 360, 64, 8)
sf, cam = ptb200.scenes.rtow_sphere_field()
run("config3_sphere_field_1080p_s256", sf, 1920, 1080, 64 if quick else 256, 10, cam)
mesh = ptb200.scenes.displaced_sphere_in_cornell(duck, n=300 if quick else 1000)
run("config4_2Mtri_mesh_4k_s256", mesh, 3840, 2160, 16 if quick else 256, 10)
run("config5_duck_4k_s4096" if not quick else "config5_duck_4k_s256", duck, 3840, 2160, 256 if quick else 4096, 10, reps=0 if not quick else 1)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / ("configs.json" if not (only or refill or cta_warps) else f"configs_{only}_refill{refill}_warps{cta_warps}.json")).write_text(json.dumps(out, indent=1))
