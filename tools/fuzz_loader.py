#!/usr/bin/env python3
"""Malformed scene files must raise, never crash: random byte mutations / truncations of models/cornell_duck.glb (container only), of a small
.gltf with data URIs and of an .obj + .mtl, each loaded in a child process (a crash shows as a negative return code).
tools/fuzz_loader.py [cases]"""
import base64, json, random, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = """
import sys
sys.path.insert(0, %r)
import ptb200
ok = bad = 0
for p in sys.argv[1:]:
    try:
        ptb200.load_scene_file(p); ok += 1
    except Exception:
        bad += 1
print(ok, bad)
""" % str(ROOT)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rnd = random.Random(3)
    seeds = []
    duck = Path("/root/reference/models/cornell_duck.glb")
    if duck.exists():
        seeds.append(("glb", duck.read_bytes()))
    tri = b"\x00\x00\x00\x00" * 3 + b"\x00\x00\x80\x3f" + b"\x00" * 8 + b"\x00" * 4 + b"\x00\x00\x80\x3f" + b"\x00" * 4
    gltf = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0, "translation": [0, 0, -3]}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "material": 0}]}], "materials": [{"name": "m", "emissiveFactor": [1, 1, 1]}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}], "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}],
            "buffers": [{"byteLength": 36, "uri": "data:application/octet-stream;base64," + base64.b64encode(tri).decode()}]}
    seeds.append(("gltf", json.dumps(gltf).encode()))
    seeds.append(("obj", b"mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nusemtl a\nf 1/1 2/2 3/3\nf 1 3 4\nf -1 -2 -3\n"))
    crashes = total = 0
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "m.mtl").write_text("newmtl a\nKd 1 0 0\nKe 1 1 1\n")
        batch = []
        for it in range(cases):
            ext, data = seeds[it % len(seeds)]
            b = bytearray(data)
            mode = rnd.random()
            if mode < 0.3:
                b = b[:rnd.randint(0, len(b))]
            elif mode < 0.8:
                for _ in range(rnd.randint(1, 12)):
                    if b:
                        b[rnd.randrange(min(len(b), 4000) if rnd.random() < 0.7 else len(b))] = rnd.randrange(256)
            else:
                i = rnd.randrange(max(1, min(len(b), 4000)))
                b[i:i + rnd.randint(1, 16)] = bytes(rnd.randrange(256) for _ in range(rnd.randint(0, 24)))
            p = td / f"c{it}.{ext}"
            p.write_bytes(bytes(b))
            batch.append(str(p))
            if len(batch) == 25 or it == cases - 1:
                r = subprocess.run([sys.executable, "-c", CHILD, *batch], capture_output=True, text=True, timeout=600)
                total += len(batch)
                if r.returncode != 0:  # find the file
                    for f in batch:
                        r1 = subprocess.run([sys.executable, "-c", CHILD, f], capture_output=True, text=True, timeout=300)
                        if r1.returncode != 0:
                            crashes += 1
                            keep = ROOT / "gpurun_out" / ("crash_" + Path(f).name)
                            keep.parent.mkdir(exist_ok=True)
                            keep.write_bytes(Path(f).read_bytes())
                            print("CRASH rc", r1.returncode, keep, r1.stderr[-200:].replace("\n", " | "))
                batch = []
    print("cases", total, "crashes", crashes)


if __name__ == "__main__":
    main()
