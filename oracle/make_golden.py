#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference (run in the build container, where /root/reference exists).

TEST INFRASTRUCTURE.  Steps:
  1. `make -C oracle ref port`   (host-compiled reference headers -> oracle/_ref/*)
  2. models/cornell_duck.glb --(our SceneLoader)--> tests/golden/cornell_duck.ptscene.gz
  3. oracle/_ref/ref_kat                          -> tests/golden/ref_kats.json
  4. oracle/_ref/ref_cpu on the duck scene        -> tests/golden/ref_cpu_*.png (+ .yuv.gz), lossless
The reference has no tests/golden vectors of its own; these are produced by its own code.
"""
import gzip
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

# (name, W, H, spp, depth, extra args)
IMAGES = [
    ("duck_160x90_s8_d10", 160, 90, 8, 10, []),
    ("duck_96x54_s64_d8", 96, 54, 64, 8, []),
    ("duck_64x48_s16_d3_cam", 64, 48, 16, 3, ["--cam", "-120", "40", "-300", "0.25", "-0.1", "-1", "60", "80"]),
]


def main():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref", "port"], check=True)
    import ptb200

    GOLD.mkdir(parents=True, exist_ok=True)
    scene = ptb200.load_scene_file(REF / "models" / "cornell_duck.glb")
    scene.save_ptscene(GOLD / "cornell_duck.ptscene.gz")
    box = ptb200.load_scene_file(REF / "models" / "cornell_box.glb")
    box.save_ptscene(GOLD / "cornell_box.ptscene.gz")

    kats = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_kat")], check=True, capture_output=True, text=True).stdout
    json.loads(kats)
    (GOLD / "ref_kats.json").write_text(kats)

    meta = {}
    with tempfile.TemporaryDirectory() as td:
        flat = Path(td) / "duck.ptscene"
        flat.write_bytes(scene.to_ptscene_bytes())
        for name, w, h, spp, depth, extra in IMAGES:
            ppm, yuv = Path(td) / f"{name}.ppm", Path(td) / f"{name}.yuv"
            out = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_cpu"), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--yuv", str(yuv), *extra],
                                 check=True, capture_output=True, text=True).stdout
            Image.open(ppm).save(GOLD / f"ref_cpu_{name}.png", optimize=True)
            (GOLD / f"ref_cpu_{name}.yuv.gz").write_bytes(gzip.compress(yuv.read_bytes(), 9, mtime=0))
            meta[name] = dict(width=w, height=h, spp=spp, depth=depth, extra=extra, ref_cpu=json.loads(out.strip().splitlines()[-1]))
            print(name, meta[name]["ref_cpu"]["seconds"], "s")
    (GOLD / "ref_cpu_images.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
