#!/usr/bin/env python3
"""GPU-box half of the golden fixtures: images rendered by the reference's OWN CUDA renderer
(oracle/_ref/ref_gpu = its RenderManager/DevicePathTracer/kernels, unmodified, sm_100), frame >= 2.

    gpurun -- python oracle/make_golden_gpu.py        -> gpurun_out/golden_refgpu/*.png, *.yuv.gz, ref_gpu_images.json
then copy that directory's content into tests/golden/ and commit it.  TEST INFRASTRUCTURE.
"""
import gzip
import json
import subprocess
import sys
import tempfile
from pathlib import Path

from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out" / "golden_refgpu"

IMAGES = [
    ("duck_160x90_s8_d10", 160, 90, 8, 10, []),
    ("duck_96x54_s64_d8", 96, 54, 64, 8, []),
    ("duck_64x48_s16_d3_cam", 64, 48, 16, 3, ["--cam", "-120", "40", "-300", "0.25", "-0.1", "-1", "60", "80"]),
    ("duck_320x180_s16_d10", 320, 180, 16, 10, []),
    ("duck_37x23_s5_d1", 37, 23, 5, 1, []),
    ("duck_64x36_s4096_d10", 64, 36, 4096, 10, []),  # converged frame: gate A at 4096 spp and the RMSE bound of the keyed RNG mode
]


def main():
    import ptb200
    OUT.mkdir(parents=True, exist_ok=True)
    scene = ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz")
    meta = {}
    with tempfile.TemporaryDirectory() as td:
        flat = Path(td) / "duck.ptscene"
        flat.write_bytes(scene.to_ptscene_bytes())
        for name, w, h, spp, depth, extra in IMAGES:
            ppm, yuv = Path(td) / f"{name}.ppm", Path(td) / f"{name}.yuv"
            r = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_gpu"), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--yuv", str(yuv), *extra],
                               capture_output=True, text=True, timeout=900)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_GPU_JSON ")]
            if r.returncode != 0 or not line:
                raise SystemExit(f"ref_gpu failed for {name}: {(r.stderr or r.stdout)[-400:]}")
            Image.open(ppm).save(OUT / f"ref_gpu_{name}.png", optimize=True)
            (OUT / f"ref_gpu_{name}.yuv.gz").write_bytes(gzip.compress(yuv.read_bytes(), 9, mtime=0))
            meta[name] = dict(width=w, height=h, spp=spp, depth=depth, extra=extra, ref_gpu=json.loads(line[-1][len("REF_GPU_JSON "):]))
            print(name, meta[name]["ref_gpu"]["msamples_per_s"], "Msamples/s", flush=True)
    gpu = subprocess.run(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    (OUT / "ref_gpu_images.json").write_text(json.dumps(dict(gpu=gpu, images=meta), indent=1))


if __name__ == "__main__":
    main()
