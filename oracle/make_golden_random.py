#!/usr/bin/env python3
"""Whole images of RANDOM scenes from the reference's own code (oracle/_ref/ref_cpu: its unmodified headers compiled for the host), so that
the CPU restatement is pinned on geometry, materials, textures and cameras the cornell_duck fixtures do not have.

    python oracle/make_golden_random.py     (needs /root/reference for `make -C oracle ref`; writes tests/golden/random/)

Scenes come from tools/fuzz_oracle.py::random_scene with fixed seeds (triangle soups, optionally a tilted floor and wall, 1-5 universal
materials with emitters, 0-2 textures bound as base-colour and emissive maps, texture coordinates the reference can look up without leaving
its texel array).  Per case: `case_K.ptscene.gz`, `case_K.png`, `case_K.yuv.gz` and an entry in `cases.json` (size, spp, depth, camera).
TEST INFRASTRUCTURE."""
import gzip, importlib.util, json, subprocess, tempfile
from pathlib import Path
import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "random"
REF = ROOT / "oracle" / "_ref" / "ref_cpu"
spec = importlib.util.spec_from_file_location("fuzz_oracle", ROOT / "tools" / "fuzz_oracle.py")


def main():
    if not REF.exists():
        raise SystemExit("oracle/_ref/ref_cpu missing: make -C oracle ref (needs /root/reference)")
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    OUT.mkdir(parents=True, exist_ok=True)
    meta = {}
    for k, seed in enumerate((3, 8, 14, 21, 33, 47, 52, 60)):
        rng = np.random.default_rng(900000 + seed)
        sc = fz.random_scene(rng)
        w, h = 2 * int(rng.integers(6, 20)), 2 * int(rng.integers(4, 14))
        spp, depth = int(rng.integers(2, 9)), int(rng.integers(1, 9))
        cam = dict(look_from=[float(v) for v in rng.uniform(-0.5, 0.5, 3) + np.array([0, 0, 0.5])], front=[float(v) for v in rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, -1.0])],
                   vfov=float(rng.uniform(25, 80)), hfov=float(rng.uniform(25, 80)))
        name = f"case_{k}"
        sc.save_ptscene(OUT / f"{name}.ptscene.gz")
        with tempfile.TemporaryDirectory() as td:
            flat, ppm, yuv = Path(td) / "s.ptscene", Path(td) / "r.ppm", Path(td) / "r.yuv"
            flat.write_bytes(sc.to_ptscene_bytes())
            subprocess.run([str(REF), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--yuv", str(yuv),
                            "--cam", *[repr(v) for v in (*cam["look_from"], *cam["front"], cam["vfov"], cam["hfov"])]], check=True, capture_output=True, text=True)
            Image.open(ppm).save(OUT / f"{name}.png", optimize=True)
            (OUT / f"{name}.yuv.gz").write_bytes(gzip.compress(yuv.read_bytes(), 9, mtime=0))
        meta[name] = dict(width=w, height=h, spp=spp, depth=depth, camera=cam, triangles=int(len(sc.tri_mat)), materials=int(len(sc.mats)), textures=len(sc.textures),
                          lit_pixels=int((np.array(Image.open(OUT / f"{name}.png")).max(axis=2) > 0).sum()))
        print(name, meta[name])
    (OUT / "cases.json").write_text(json.dumps(meta, indent=1) + "\n")


if __name__ == "__main__":
    main()
