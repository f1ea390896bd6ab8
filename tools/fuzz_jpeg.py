#!/usr/bin/env python3
"""Random JPEG files (Pillow: random sizes, quality, chroma subsampling, progressive / sequential, optimised tables, restart intervals, grey)
decoded by the repo's loader (csrc/host/JpegDecoder.h) and by the reference's own decoder (oracle/_ref/ref_stb, build container only): every
texel must agree.  tools/fuzz_jpeg.py [cases]"""
import base64, io, json, random, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402
STB = ROOT / "oracle" / "_ref" / "ref_stb"


def gltf(tmp, image_bytes):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32); uv = np.array([[0, 0], [1, 0], [0, 1]], np.float32); blob = pos.tobytes() + uv.tobytes()
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}], "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "material": 0}]}],
         "materials": [{"name": "photo", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}], "textures": [{"source": 0}],
         "images": [{"uri": "data:image/jpeg;base64," + base64.b64encode(image_bytes).decode()}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    p = tmp / "x.gltf"; p.write_text(json.dumps(g)); return p


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rnd = random.Random(11)
    bad = 0
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for it in range(cases):
            w, h = rnd.randint(1, 90), rnd.randint(1, 70)
            rng = np.random.default_rng(it)
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([127 + 120 * np.sin(x / (2.0 + it % 7) + it), 127 + 120 * np.cos(y / (3.0 + it % 5)), (x * 9 + y * 5 + it * 31) % 256], axis=2) + rng.normal(0, rnd.choice([0, 5, 30]), (h, w, 3))
            img = np.clip(img, 0, 255).astype(np.uint8)
            grey = rnd.random() < 0.15
            kw = dict(quality=rnd.choice([1, 5, 20, 50, 75, 90, 95, 100]), progressive=rnd.random() < 0.4, optimize=rnd.random() < 0.3)
            if not grey:
                kw["subsampling"] = rnd.choice([0, 1, 2])
            if rnd.random() < 0.3:
                kw["restart_marker_blocks"] = rnd.randint(1, 7)
            buf = io.BytesIO()
            Image.fromarray(img[:, :, 0] if grey else img, "L" if grey else "RGB").save(buf, "JPEG", **kw)
            data = buf.getvalue()
            f = td / "x.jpg"; f.write_bytes(data); raw = td / "o.raw"
            r = subprocess.run([str(STB), str(f), str(raw), "3"], capture_output=True, text=True)
            if r.returncode != 0:
                print("stb refused", w, h, kw, r.stderr.strip()); bad += 1
                continue
            head, body = raw.read_bytes().split(b"\n", 1)
            W, H, _ = map(int, head.split())
            ref = np.frombuffer(body, np.uint8).reshape(H, W, 3)
            tex = ptb200.load_scene_file(gltf(td, data)).textures[0]
            if tex.shape != (H, W, 3) or not np.array_equal(tex, ref.astype(np.float32)):
                bad += 1
                print("MISMATCH", w, h, "grey" if grey else "rgb", kw, tex.shape)
    print("cases", cases, "bad", bad)


if __name__ == "__main__":
    main()
