#!/usr/bin/env python3
"""BASELINE config 4 (2 M-triangle mesh, 3840x2160) on one GPU: tree width / kernel variants of the loaded library.
usage: tools/config4_ab.py [spp] [n]     (A/B builds: PTB200_LIBPTCORE=_variants/<name>/libptcore.so)"""
import json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
scene = ptb200.scenes.displaced_sphere_in_cornell(duck, n=n)
w, h = 3840, 2160
fb = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
ref = None
for name, opts in (("bvh2 wavefront", {}), ("bvh2 wavefront + nodes persisting in L2", {ptb200.PT_OPT_L2_PERSIST_NODES: 1}), ("bvh4 wavefront", {ptb200.PT_OPT_BVH_WIDTH: 4}),
                   ("bvh2 pool", {ptb200.PT_OPT_KERNEL: ptb200.PT_KERNEL_POOL}), ("bvh2 pool + nodes persisting in L2", {ptb200.PT_OPT_KERNEL: ptb200.PT_KERNEL_POOL, ptb200.PT_OPT_L2_PERSIST_NODES: 1}),
                   ("bvh2 wavefront float nodes", {ptb200.PT_OPT_NODE_FORMAT: ptb200.PT_NODES_FULL})):
    pt = ptb200.PathTracer(0)
    for k, v in opts.items():
        pt.set_option(k, v)
    t0 = time.perf_counter(); pt.upload_scene(scene); up = time.perf_counter() - t0
    pt.set_camera(); pt.set_params(spp, 10)
    pt.bind_framebuffer(fb.data_ptr(), 0, w, h)
    best = 1e9
    for i in range(3):
        fb.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pt.render_tile_async(0, 0, w, h); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = pt.stats()
    img = fb.cpu()
    if ref is None:
        ref = img
    print(json.dumps({"variant": name, "lib": str(ptb200.load_library()._name).split("/")[-2], "spp": spp, "ms": best, "msamples_per_s": w * h * spp / best / 1e3, "upload_s": up,
                      "scene_mb": st["scene_bytes"] / 1e6, "n_vertices": st["n_vertices"], "bvh_nodes": st["bvh_nodes"], "identical": bool(torch.equal(img, ref))}), flush=True)
    pt.close()
