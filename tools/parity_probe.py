#!/usr/bin/env python3
"""Parity probe (GPU box): where do the CUDA core and the checkers part ways?

  1. CUDA core vs CPU oracle: identical-pixel fractions over (spp, depth); for differing pixels the per-ray event
     traces of both sides are walked to the first divergent event and classified.
  2. CUDA core vs the reference's own CUDA renderer (oracle/_ref/ref_gpu, frame >= 2) on the same parameters.
Writes gpurun_out/parity_probe.json.  Test tooling: the oracle is used as the checker only.
"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import _oracle  # noqa: E402
import ptb200  # noqa: E402


def cmp(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return dict(identical=float((d.max(axis=2) == 0).mean()), within1=float((d.max(axis=2) <= 1).mean()), mad=float(d.mean()),
                rmse=float(np.sqrt((d.astype(np.float64) ** 2).mean())), max=int(d.max()))


def first_divergence(eg, eo):
    n = min(len(eg), len(eo))
    for i in range(n):
        g, o = eg[i], eo[i]
        if g[0] != o[0] or g[1] != o[1]:
            return i, "sample/bounce index (draw count already diverged)"
        if not np.array_equal(g[6:12].view(np.uint32), o[6:12].view(np.uint32)):
            # ray differs in low bits: not a divergence unless the decision differs
            pass
        if g[2] != o[2]:
            return i, f"different primitive: gpu {int(g[2])} t={g[3]!r} vs oracle {int(o[2])} t={o[3]!r}"
    if len(eg) != len(eo):
        return n, f"event count {len(eg)} vs {len(eo)}"
    return -1, "none"


def main():
    out = {}
    orc = _oracle.load()
    scene = ptb200.load_scene_file(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz")
    world = orc.world(scene)
    pt = ptb200.PathTracer(0)
    pt.upload_scene(scene)
    pt.set_camera()
    W, H = 160, 90
    table = []
    for spp, depth in [(1, 1), (1, 2), (1, 3), (1, 10), (8, 10), (32, 10), (64, 8)]:
        pt.set_params(spp, depth)
        g, _ = pt.render_frame_host(W, H)
        o, _, _ = orc.render(world, W, H, spp, depth)
        row = dict(spp=spp, depth=depth, **cmp(g, o))
        table.append(row)
        print(row, flush=True)
    out["gpu_vs_oracle_160x90"] = table

    # classify first divergences at spp=8 depth=10
    spp, depth = 8, 10
    pt.set_params(spp, depth)
    g, _ = pt.render_frame_host(W, H)
    o, _, _ = orc.render(world, W, H, spp, depth)
    bad = np.argwhere(np.abs(g.astype(int) - o.astype(int)).max(axis=2) > 0)
    kinds = {}
    details = []
    for (row, x) in bad[:200]:
        y = H - 1 - row
        eg, cg = pt.trace_pixel(W, H, int(x), int(y))
        eo, co = orc.trace_pixel(world, W, H, spp, depth, int(x), int(y))
        i, why = first_divergence(eg, eo)
        key = why.split(":")[0]
        kinds[key] = kinds.get(key, 0) + 1
        if len(details) < 25 and i >= 0 and i < min(len(eg), len(eo)):
            details.append(dict(x=int(x), y=int(y), event=int(i), why=why, gpu=[float(v) for v in eg[i]], oracle=[float(v) for v in eo[i]],
                                prev_gpu=[float(v) for v in eg[i - 1]] if i else None, prev_oracle=[float(v) for v in eo[i - 1]] if i else None))
    out["divergence_kinds_spp8_d10"] = kinds
    out["divergence_details"] = details
    print(kinds, flush=True)
    for d in details[:8]:
        print(json.dumps(d), flush=True)

    # vs the reference's own CUDA renderer
    ref_gpu = ROOT / "oracle" / "_ref" / "ref_gpu"
    if ref_gpu.exists():
        rows = []
        with tempfile.TemporaryDirectory() as td:
            flat = Path(td) / "duck.ptscene"
            flat.write_bytes(scene.to_ptscene_bytes())
            for (w, h, spp, depth) in [(160, 90, 8, 10), (320, 180, 16, 10), (640, 360, 64, 8)]:
                ppm = Path(td) / "ref.ppm"
                r = subprocess.run([str(ref_gpu), str(flat), str(w), str(h), str(spp), str(depth), str(ppm)], capture_output=True, text=True, timeout=900)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_GPU_JSON ")]
                if r.returncode != 0 or not line:
                    rows.append(dict(w=w, h=h, spp=spp, depth=depth, error=(r.stderr or r.stdout)[-300:]))
                    continue
                ref = np.array(Image.open(ppm).convert("RGB"))
                pt.set_params(spp, depth)
                g, _ = pt.render_frame_host(w, h)
                o, _, _ = orc.render(world, w, h, spp, depth)
                rows.append(dict(w=w, h=h, spp=spp, depth=depth, ours_vs_refgpu=cmp(g, ref), oracle_vs_refgpu=cmp(o, ref), ours_vs_oracle=cmp(g, o),
                                 refgpu=json.loads(line[-1][len("REF_GPU_JSON "):])))
                print(rows[-1], flush=True)
        out["vs_reference_cuda"] = rows
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "parity_probe.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
