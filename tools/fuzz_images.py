#!/usr/bin/env python3
"""Random BMP / TGA variants (the byte-level writers of oracle/make_golden_images.py with random sizes, depths, headers, masks, palettes, RLE,
orientation) decoded by the repo's loader and by the reference's own decoder (oracle/_ref/ref_stb, build container only): every texel must
agree and what stb_image refuses must be refused.  tools/fuzz_images.py   -> "cases 400 bad 0" on the committed build."""
import sys, subprocess, tempfile, json, base64, importlib.util, random
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT/'tests'))
import ptb200
spec = importlib.util.spec_from_file_location("mg", ROOT/"oracle/make_golden_images.py"); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
spec2 = importlib.util.spec_from_file_location("thl", ROOT/"tests/test_host_logic.py")
STB = ROOT/'oracle/_ref/ref_stb'
def gltf(tmp, name, image_bytes):
    pos = np.array([[0,0,0],[1,0,0],[0,1,0]], np.float32); uv = np.array([[0,0],[1,0],[0,1]], np.float32); blob = pos.tobytes()+uv.tobytes()
    g = {"asset":{"version":"2.0"},"scene":0,"scenes":[{"nodes":[0]}],"nodes":[{"mesh":0}],"meshes":[{"primitives":[{"attributes":{"POSITION":0,"TEXCOORD_0":1},"material":0}]}],
      "materials":[{"name":"photo","pbrMetallicRoughness":{"baseColorTexture":{"index":0}}}],"textures":[{"source":0}],"images":[{"uri":"data:image/x;base64,"+base64.b64encode(image_bytes).decode()}],
      "accessors":[{"bufferView":0,"componentType":5126,"count":3,"type":"VEC3"},{"bufferView":1,"componentType":5126,"count":3,"type":"VEC2"}],
      "bufferViews":[{"buffer":0,"byteOffset":0,"byteLength":36},{"buffer":0,"byteOffset":36,"byteLength":24}],
      "buffers":[{"byteLength":len(blob),"uri":"data:application/octet-stream;base64,"+base64.b64encode(blob).decode()}]}
    p = tmp/f"{name}.gltf"; p.write_text(json.dumps(g)); return p
rnd = random.Random(7)
bad = 0; n = 0
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    for it in range(400):
        w, h = rnd.randint(1, 40), rnd.randint(1, 24)
        if rnd.random() < 0.5:
            bpp = rnd.choice([1,4,8,16,24,32]); kw = {}
            if bpp <= 8:
                kw['palette_n'] = rnd.randint(2, 1 << bpp) if bpp > 1 else 2
                kw['hsz'] = rnd.choice([40,108,124])
                if rnd.random() < 0.2: kw = dict(os2=True, palette_n=1 << bpp, index_n=max(1, ((14+12+3*(1<<bpp)) - 38)//3))
            elif bpp == 24:
                kw['hsz'] = rnd.choice([40,56,108,124]); kw['top_down'] = rnd.random() < 0.3
                if rnd.random() < 0.15: kw = dict(os2=True)
            else:
                kw['hsz'] = rnd.choice([40,108,124]); kw['top_down'] = rnd.random() < 0.3
                if rnd.random() < 0.6:
                    kw['compress'] = 3
                    def mk(bits_total):
                        # random non-overlapping masks
                        widths = [rnd.randint(1, 8 if bits_total == 32 else 5) for _ in range(3)]
                        pos = 0; ms = []
                        order = [0,1,2]; rnd.shuffle(order)
                        res = [0,0,0]
                        for k in order:
                            gap = rnd.randint(0, 2)
                            pos += gap
                            res[k] = ((1 << widths[k]) - 1) << pos
                            pos += widths[k]
                        if pos > bits_total: return None
                        return (res[0], res[1], res[2], 0)
                    m = None
                    while m is None: m = mk(bpp)
                    kw['masks'] = m
            if rnd.random() < 0.15 and not kw.get('os2'): kw['gap'] = rnd.randint(1, 40)
            data = mg.bmp(w, h, bpp, it, **kw); ext = 'bmp'; desc = ('bmp', w, h, bpp, kw)
        else:
            kind = rnd.choice(['rgb', 'grey', 'indexed']); kw = dict(rle=rnd.random() < 0.5, top_down=rnd.random() < 0.5)
            if kind == 'rgb': bits = rnd.choice([15,16,24,32])
            elif kind == 'grey': bits = rnd.choice([8,16])
            else:
                bits = 0; kw.update(pal_bits=rnd.choice([8,15,16,24,32]), idx_bits=rnd.choice([8,16]))
                kw['pal_n'] = rnd.randint(1, 256 if kw['idx_bits'] == 8 else 600); kw['bad_index'] = rnd.random() < 0.3 and kw['pal_n'] < 250
            if rnd.random() < 0.3: kw['image_id'] = bytes(rnd.randint(0,255) for _ in range(rnd.randint(1, 30)))
            data = mg.tga(w, h, kind, bits, it, **kw); ext = 'tga'; desc = ('tga', w, h, kind, bits, kw)
        f = td/f"x.{ext}"; f.write_bytes(data); raw = td/"o.raw"
        r = subprocess.run([str(STB), str(f), str(raw), "3"], capture_output=True, text=True)
        sc = ptb200.load_scene_file(gltf(td, "x", data)); tex = sc.textures[0]
        n += 1
        if r.returncode != 0:
            if tex.shape[0] != 0: bad += 1; print("stb refused, we decoded:", desc, r.stderr.strip())
            continue
        head, body = raw.read_bytes().split(b"\n", 1); W, H, C = map(int, head.split()); ref = np.frombuffer(body, np.uint8).reshape(H, W, 3)
        if tex.shape != (H, W, 3) or not np.array_equal(tex, ref.astype(np.float32)):
            bad += 1; print("MISMATCH", desc, tex.shape, (H, W))
print("cases", n, "bad", bad)
