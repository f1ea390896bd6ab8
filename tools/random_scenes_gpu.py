#!/usr/bin/env python3
"""GPU box: the CUDA core on the random-scene fixtures (tests/golden/random, images from the reference's own headers on the CPU): differing
pixels per case and per kernel.  tools/random_scenes_gpu.py"""
import json, sys
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402
G = ROOT / "tests" / "golden" / "random"
meta = json.loads((G / "cases.json").read_text())
pt = ptb200.PathTracer(0)
for name, m in sorted(meta.items()):
    sc = ptb200.load_scene_file(G / f"{name}.ptscene.gz")
    ref = np.array(Image.open(G / f"{name}.png").convert("RGB")).astype(np.int32)
    cam = dict(look_from=tuple(m["camera"]["look_from"]), front=tuple(m["camera"]["front"]), vfov=m["camera"]["vfov"], hfov=m["camera"]["hfov"])
    out = {}
    for kname, k in (("wavefront", ptb200.PT_KERNEL_PERSISTENT), ("direct", ptb200.PT_KERNEL_DIRECT), ("pool", ptb200.PT_KERNEL_POOL)):
        pt.set_option(ptb200.PT_OPT_KERNEL, k)
        pt.upload_scene(sc); pt.set_camera(**cam); pt.set_params(m["spp"], m["depth"])
        rgb, _ = pt.render_frame_host(m["width"], m["height"])
        d = np.abs(rgb.astype(np.int32) - ref).max(axis=2)
        out[kname] = dict(differ=int((d > 0).sum()), over1=int((d > 1).sum()), max=int(d.max()))
    print(json.dumps(dict(case=name, pixels=m["width"] * m["height"], spp=m["spp"], depth=m["depth"], **out)), flush=True)
pt.close()
