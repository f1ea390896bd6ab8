#!/usr/bin/env python3
"""Render one BASELINE config once through the C ABI (for ncu captures of the non-bench configs).
usage: tools/render_config.py {config3|config4} [spp] [n]"""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200

which = sys.argv[1] if len(sys.argv) > 1 else "config4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
duck = ptb200.load_scene_file(ROOT / "tests/golden/cornell_duck.ptscene.gz")
if which == "config3":
    scene, cam = ptb200.scenes.rtow_sphere_field()
    w, h = 1920, 1080
else:
    scene, cam = ptb200.scenes.displaced_sphere_in_cornell(duck, n=int(sys.argv[3]) if len(sys.argv) > 3 else 1000), {}
    w, h = 3840, 2160
pt = ptb200.PathTracer(0)
pt.upload_scene(scene); pt.set_camera(**cam); pt.set_params(spp, 10)
fb = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda"); fy = torch.zeros(w * h * 3 // 2, dtype=torch.uint8, device="cuda")
pt.bind_framebuffer(fb.data_ptr(), fy.data_ptr(), w, h)
for i in range(2):
    t0 = time.perf_counter(); pt.render_tile_async(0, 0, w, h); pt.wait(); dt = time.perf_counter() - t0
st = pt.stats()
print(which, "spp", spp, "ms", round(dt * 1e3, 2), "Msamples/s", round(w * h * spp / dt / 1e6, 1), "Mrays/s", round(st["rays"] / 2 / dt / 1e6, 1), "scene MB", round(st["scene_bytes"] / 1e6, 1))
