#!/usr/bin/env python3
"""Gate A at BASELINE config 2's FULL size: cornell_duck 1920x1080, depth 10, 1024 spp, our CUDA core against the reference's own CUDA
renderer (oracle/_ref/ref_gpu, unmodified kernels) on the same box.  About 6 minutes of the reference.  Writes gpurun_out/r02_gate_a_full.json
(copy it to profiles/).   usage: tools/gate_a_full.py [spp]"""
import json, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w, h, depth = 1920, 1080, 10
duck = ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz")
pt = ptb200.PathTracer(0)
pt.upload_scene(duck); pt.set_camera(); pt.set_params(spp, depth)
t0 = time.perf_counter(); rgb, yuv = pt.render_frame_host(w, h); ours_s = time.perf_counter() - t0
pt.close()
with tempfile.TemporaryDirectory() as td:
    flat, ppm, yv = Path(td) / "duck.ptscene", Path(td) / "ref.ppm", Path(td) / "ref.yuv"
    flat.write_bytes(duck.to_ptscene_bytes())
    r = subprocess.run([str(ROOT / "oracle" / "_ref" / "ref_gpu"), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--yuv", str(yv)], capture_output=True, text=True, timeout=3000)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_GPU_JSON ")]
    if r.returncode != 0 or not line:
        raise SystemExit(f"ref_gpu failed: {(r.stderr or r.stdout)[-400:]}")
    ref = np.array(Image.open(ppm).convert("RGB"))
    ref_yuv = np.frombuffer(yv.read_bytes(), np.uint8)
meta = json.loads(line[-1][len("REF_GPU_JSON "):])
d = np.abs(rgb.astype(np.int32) - ref.astype(np.int32)).max(axis=2)
ys, xs = np.nonzero(d)
out = {"workload": f"cornell_duck {w}x{h} spp={spp} depth={depth}", "pixels": w * h, "pixels_differing": int((d > 0).sum()), "max_abs_diff_levels": int(d.max()),
       "differing_pixels_xy_diff": [[int(x), int(y), int(d[y, x])] for y, x in zip(ys[:64], xs[:64])],
       "y_plane_bytes_differing": int((yuv[: w * h] != ref_yuv[: w * h]).sum()), "reference_valid": bool(ref.mean() > 0.5 * rgb.mean()),
       "ours_seconds_incl_host_copies": ours_s, "reference_seconds": meta["seconds"], "reference_msamples_per_s": meta["msamples_per_s"],
       "note": "differences are rays that meet the shared edge of two triangles at exactly the same t (DESIGN.md section 2): the reference keeps whichever its own tree reaches first"}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r02_gate_a_full.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
