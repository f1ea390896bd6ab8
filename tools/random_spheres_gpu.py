#!/usr/bin/env python3
"""GPU box: the CUDA core against the CPU restatement (the checker, device draw order) on the sphere fixtures of tests/golden/spheres (whose images
pin the restatement to the reference's dead classes): differing pixels per case.  tools/random_spheres_gpu.py"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ptb200  # noqa: E402
import _oracle  # noqa: E402
G = ROOT / "tests" / "golden" / "spheres"
meta = json.loads((G / "cases.json").read_text())
orc = _oracle.load()
pt = ptb200.PathTracer(0)
for name, m in sorted(meta.items()):
    sc = ptb200.load_scene_file(G / m["scene"])
    cam = dict(look_from=tuple(m["camera"]["look_from"]), front=tuple(m["camera"]["front"]), vfov=m["camera"]["vfov"], hfov=m["camera"]["hfov"])
    ref, _, _ = orc.render(sc, m["width"], m["height"], m["spp"], m["depth"], camera=cam)
    out = {}
    for kname, k in (("wavefront", ptb200.PT_KERNEL_PERSISTENT), ("direct", ptb200.PT_KERNEL_DIRECT), ("pool", ptb200.PT_KERNEL_POOL)):
        pt.set_option(ptb200.PT_OPT_KERNEL, k)
        pt.upload_scene(sc); pt.set_camera(**cam); pt.set_params(m["spp"], m["depth"])
        rgb, _ = pt.render_frame_host(m["width"], m["height"])
        d = np.abs(rgb.astype(np.int32) - ref.astype(np.int32)).max(axis=2)
        out[kname] = dict(differ=int((d > 0).sum()), over1=int((d > 1).sum()), max=int(d.max()))
    print(json.dumps(dict(case=name, pixels=m["width"] * m["height"], spp=m["spp"], depth=m["depth"], **out)), flush=True)
pt.close()
