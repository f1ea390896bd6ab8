#!/usr/bin/env python3
"""GPU box: run the reference CUDA renderer on a grid of (size, spp) and compare with ours (diagnostic)."""
import json, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ptb200, _oracle

scene = ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz")
pt = ptb200.PathTracer(0); pt.upload_scene(scene); pt.set_camera()
orc = _oracle.load(); world = orc.world(scene)
ref_gpu = ROOT / "oracle" / "_ref" / "ref_gpu"
with tempfile.TemporaryDirectory() as td:
    flat = Path(td) / "duck.ptscene"; flat.write_bytes(scene.to_ptscene_bytes())
    for (w, h, spp, depth) in [(192, 108, 12, 10), (192, 108, 8, 10), (200, 112, 12, 10), (200, 112, 8, 10), (160, 90, 12, 10), (160, 90, 8, 10), (240, 135, 6, 10), (128, 72, 24, 6)]:
        ppm = Path(td) / "r.ppm"
        r = subprocess.run([str(ref_gpu), str(flat), str(w), str(h), str(spp), str(depth), str(ppm)], capture_output=True, text=True, timeout=600)
        ref = np.array(Image.open(ppm).convert("RGB"))
        pt.set_params(spp, depth); g, _ = pt.render_frame_host(w, h)
        o, _, _ = orc.render(world, w, h, spp, depth)
        print(dict(w=w, h=h, spp=spp, ref_mean=float(ref.mean()), ours_mean=float(g.mean()), oracle_mean=float(o.mean()), ours_eq_ref=float((g == ref).all(axis=2).mean()),
                   oracle_eq_ref=float((o == ref).all(axis=2).mean()), ref_white=int((ref.min(axis=2) == 255).sum()), ours_white=int((g.min(axis=2) == 255).sum())), flush=True)
# camera fixture: which pixels differ, per kernel
m = json.loads((ROOT / "tests/golden/ref_gpu_images.json").read_text())["images"]["duck_64x48_s16_d3_cam"]
v = [float(x) for x in m["extra"][1:9]]
cam = dict(look_from=tuple(v[0:3]), front=tuple(v[3:6]), vfov=v[6], hfov=v[7])
ref = np.array(Image.open(ROOT / "tests/golden/ref_gpu_duck_64x48_s16_d3_cam.png").convert("RGB"))
pt.set_camera(**cam); pt.set_params(16, 3)
for k in (0, 1, 2):
    pt.set_option(ptb200.PT_OPT_KERNEL, k)
    g, _ = pt.render_frame_host(64, 48)
    bad = np.argwhere((g != ref).any(axis=2))
    print("kernel", k, "differing pixels", [(int(r), int(c), g[r, c].tolist(), ref[r, c].tolist()) for r, c in bad], flush=True)
o, _, _ = orc.render(world, 64, 48, 16, 3, camera=cam)
print("oracle differing vs ref", int((o != ref).any(axis=2).sum()))
for r, c in bad[:2]:
    eg, cg = pt.trace_pixel(64, 48, int(c), int(47 - r))
    eo, co = orc.trace_pixel(world, 64, 48, 16, 3, int(c), int(47 - r), camera=cam)
    print("pixel", r, c, "gpu col", cg.tolist(), "oracle col", co.tolist(), "n events", len(eg), len(eo))
    for i in range(min(len(eg), len(eo))):
        if eg[i][2] != eo[i][2] or abs(eg[i][3] - eo[i][3]) > 1e-3 * abs(eo[i][3]):
            print("  first differing event", i, eg[i].tolist(), eo[i].tolist()); break
