"""Host-side logic that needs no GPU: the C ABI library loads and exports what include/ptcore.h declares, scene
ingestion (own GLB / PNG / OBJ readers, .ptscene round trip), the host SAH BVH, float-vs-double folding used by the
kernels, and the loud failure without a CUDA device."""
import base64
import ctypes as C
import json
import re
import struct
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLD, ROOT


def declared_functions():
    text = (ROOT / "include" / "ptcore.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ptcore|ptscene|pt)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(core_lib):
    names = declared_functions()
    assert len(names) >= 28 and "ptcore_render_tile_async" in names and "pt_tileq_claim" in names
    for n in names:
        assert hasattr(core_lib, n), f"libptcore.so does not export {n}"
    assert core_lib.ptcore_abi_version() == 1


def test_no_cpu_fallback_without_a_device(ptb, core_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ptb.PtError) as e:
        ptb.PathTracer(0)
    assert "no CPU fallback" in str(e.value)


def test_package_never_imports_the_oracle():
    """The product may not import, link, call or execute anything under oracle/ (comments naming it are fine)."""
    pkg = ROOT / "multi-gpu-path-tracer_b200"
    forbidden = ("pt_oracle", "libpt_oracle", "import _oracle", "pto_", "oracle/_ref", "oracle/_build", "ref_cpu", "ref_gpu")
    files = [f for ext in ("*.py", "*.h", "*.cpp", "*.cu", "*.cuh") for f in pkg.rglob(ext)]
    assert len(files) > 15
    for f in files:
        text = f.read_text(errors="replace")
        for bad in forbidden:
            assert bad not in text, (f, bad)
        assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f


def test_struct_layouts_match_the_header(ptb):
    assert C.sizeof(ptb.PtMaterial) == 44 and C.sizeof(ptb.PtCamera) == 32 and C.sizeof(ptb.PtTile) == 16
    assert C.sizeof(ptb.PtStats) == 6 * 8 + 4 * 4 + 8 + 8 + 8 + 2 * 4 + 8 + 2 * 4


# ------------------------------------------------------------------ scene ingestion
def test_duck_fixture_matches_survey_appendix_a(duck):
    # SURVEY Appendix A pins (the reference pins nothing at the assimp boundary)
    assert duck.tri_pos.shape == (4224, 9) and len(duck.mats) == 5 and len(duck.textures) == 1
    assert np.bincount(duck.tri_mat, minlength=5).tolist() == [2, 6, 2, 2, 4212]
    lo, hi = duck.tri_pos.reshape(-1, 3).min(0), duck.tri_pos.reshape(-1, 3).max(0)
    assert np.allclose(lo, [-298.21, -215.13, -1246.35], atol=0.01) and np.allclose(hi, [257.79, 339.58, -687.15], atol=0.01)
    assert np.allclose(duck.tri_pos[8], [109.791084, 338.58173, -861.65204, -150.20908, 338.5817, -861.65204, -150.20908, 338.58167, -1071.6521], atol=2e-4)
    assert np.allclose(duck.tri_uv[12], [0.866606, 0.398924, 0.871384, 0.397619, 0.87416, 0.398826], atol=1e-6)
    assert np.allclose(duck.mats["base"], [[0, 0.5, 0], [0.4, 0.4, 0.4], [0, 0, 0], [0.5, 0, 0], [1, 1, 1]])
    assert duck.mats["emis"].tolist() == [[0, 0, 0], [0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0]]
    assert duck.mats["base_tex"].tolist() == [-1, -1, -1, -1, 0] and duck.textures[0].shape == (512, 512, 3)
    p = duck.tri_pos[8].reshape(3, 3).astype(np.float64)
    assert abs(0.5 * np.linalg.norm(np.cross(p[1] - p[0], p[2] - p[0])) - 27300.02) < 0.05


def test_box_fixture_has_no_emitter(box):
    assert box.tri_pos.shape == (36, 9) and np.bincount(box.tri_mat, minlength=5).tolist() == [6, 2, 2, 24, 2]
    assert not box.mats["emis"].any()


def _png(w, h, rgb, palette=False):
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    if palette:
        colours = sorted({tuple(p) for p in rgb.reshape(-1, 3)})
        idx = np.array([colours.index(tuple(p)) for p in rgb.reshape(-1, 3)], np.uint8).reshape(h, w)
        raw = b"".join(b"\x00" + idx[y].tobytes() for y in range(h))
        body = chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 3, 0, 0, 0)) + chunk(b"PLTE", bytes(sum(colours, ()))) + chunk(b"IDAT", zlib.compress(raw))
    else:
        # filter type 1 (sub) on odd rows to exercise the unfilter
        rows = []
        for y in range(h):
            row = rgb[y].astype(np.int32).reshape(-1)
            if y % 2:
                prev = np.concatenate([np.zeros(3, np.int32), row[:-3]])
                rows.append(b"\x01" + ((row - prev) % 256).astype(np.uint8).tobytes())
            else:
                rows.append(b"\x00" + row.astype(np.uint8).tobytes())
        body = chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"".join(rows)))
    return b"\x89PNG\r\n\x1a\n" + body + chunk(b"IEND", b"")


def _glb(tmp_path, palette):
    """two nodes (parent scale+translate, child rotate) with one textured quad and one emissive triangle"""
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tri = np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float32)
    tex = (np.arange(4 * 3 * 3) * 7 % 256).astype(np.uint8).reshape(3, 4, 3)
    png = _png(4, 3, tex, palette)
    blobs, views = b"", []
    for b in (pos.tobytes(), uv.tobytes(), idx.tobytes(), tri.tobytes(), png):
        views.append({"buffer": 0, "byteOffset": len(blobs), "byteLength": len(b)})
        blobs += b + b"\x00" * (-len(b) % 4)
    s = float(np.sqrt(0.5))
    gltf = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 2]}],
        "nodes": [{"children": [1], "scale": [2, 2, 2], "translation": [10, 0, 0]}, {"mesh": 0, "rotation": [0, 0, s, s]}, {"mesh": 1}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 1}]},
                   {"primitives": [{"attributes": {"POSITION": 3}, "material": 0}]}],
        "materials": [{"name": "Light", "emissiveFactor": [1, 0.5, 0.25], "pbrMetallicRoughness": {"baseColorFactor": [0, 0, 0, 1]}},
                      {"name": "tex", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
        "textures": [{"source": 0}], "images": [{"bufferView": 4, "mimeType": "image/png"}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC2"},
                      {"bufferView": 2, "componentType": 5123, "count": 6, "type": "SCALAR"}, {"bufferView": 3, "componentType": 5126, "count": 3, "type": "VEC3"}],
        "bufferViews": views, "buffers": [{"byteLength": len(blobs)}],
    }
    js = json.dumps(gltf).encode()
    js += b" " * (-len(js) % 4)
    glb = struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(blobs)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blobs), 0x004E4942) + blobs
    path = tmp_path / ("scene_p.glb" if palette else "scene.glb")
    path.write_bytes(glb)
    return path, tex


@pytest.mark.parametrize("palette", [False, True])
def test_glb_loader_bakes_transforms_orders_by_material_flips_v_and_decodes_png(ptb, core_lib, tmp_path, palette):
    path, tex = _glb(tmp_path, palette)
    sc = ptb.load_scene_file(path)
    # material-index order (assimp PreTransformVertices): the emissive triangle (material 0) first, then the quad (material 1)
    assert sc.tri_mat.tolist() == [0, 1, 1]
    assert np.allclose(sc.tri_pos[0], [0, 0, 1, 1, 0, 1, 0, 1, 1])
    # quad: rotate 90 deg about z, scale 2, translate (10,0,0): (x,y,0) -> (10 - 2y, 2x, 0)
    assert np.allclose(sc.tri_pos[1], [10, 0, 0, 10, 2, 0, 8, 2, 0], atol=1e-5)
    assert np.allclose(sc.tri_uv[1], [0, 1, 1, 1, 1, 0])  # V flipped to 1 - v
    assert sc.mats["emis"][0].tolist() == [1, 0.5, 0.25] and sc.mats["base"][1].tolist() == [1, 1, 1]  # baseColorFactor default (1,1,1)
    assert sc.mats["base_tex"].tolist() == [-1, 0]
    assert np.array_equal(sc.textures[0], tex.astype(np.float32))
    ok, msg, _ = ptb.bvh_selftest(sc)
    assert ok, msg


def test_obj_loader_routes_materials_by_name_prefix(ptb, core_lib, tmp_path):
    # reference README.md:60-76 / src/obj_loader.h:65-96
    (tmp_path / "m.mtl").write_text("newmtl lambertian_red\nKa 0.5 0.1 0.1\nnewmtl metal_a\nKa 0.8 0.8 0.9\nNs 0.3\nnewmtl dielectric_g\nNi 1.5\n"
                                    "newmtl diffuse_light_top\nKd 4 4 4\nnewmtl plain\nKd 0.2 0.3 0.4\nKe 1 1 1\n")
    (tmp_path / "s.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
                                    "usemtl lambertian_red\nf 1/1 2/2 3/3 4/4\nusemtl metal_a\nf 1 2 3\nusemtl dielectric_g\nf -4 -3 -2\n"
                                    "usemtl diffuse_light_top\nf 1 3 4\nusemtl plain\nf 1//1 2//1 4//1\nl 1 2\n")
    sc = ptb.load_scene_file(tmp_path / "s.obj")
    assert sc.tri_mat.tolist() == [0, 0, 1, 2, 3, 4]  # the quad is fan-triangulated, the line is dropped
    assert sc.mats["type"].tolist() == [ptb.PT_MAT_LAMBERTIAN, ptb.PT_MAT_METAL, ptb.PT_MAT_DIELECTRIC, ptb.PT_MAT_DIFFUSE_LIGHT, ptb.PT_MAT_UNIVERSAL]
    assert np.allclose(sc.mats["base"][0], [0.5, 0.1, 0.1]) and np.isclose(sc.mats["fuzz"][1], 0.3) and np.isclose(sc.mats["ior"][2], 1.5)
    assert sc.mats["emis"][3].tolist() == [4, 4, 4] and sc.mats["emis"][4].tolist() == [1, 1, 1] and np.allclose(sc.mats["base"][4], [0.2, 0.3, 0.4])
    assert np.allclose(sc.tri_uv[0], [0, 0, 1, 0, 1, 1])


def test_gltf_with_a_jpeg_texture_loads_with_the_placeholder_texture(ptb, core_lib, tmp_path):
    """The reference decodes textures with stb_image (JPEG, BMP, TGA ... besides PNG, src/HostScene.cpp:10-51); only PNG is restated
    here.  A JPEG image must not abort the load: the material keeps its texture slot and the texture has no texels,
    which the device shades with the reference's placeholder colour for a texture without data (src/Texture.h:33-35)."""
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    blob = pos.tobytes() + uv.tobytes()
    jpeg = bytes([0xFF, 0xD8, 0xFF, 0xE0, 0, 16]) + b"JFIF\x00" + bytes(32)
    gltf = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "material": 0}]}],
        "materials": [{"name": "photo", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
        "textures": [{"source": 0}], "images": [{"uri": "data:image/jpeg;base64," + base64.b64encode(jpeg).decode()}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
    }
    path = tmp_path / "jpeg.gltf"
    path.write_text(json.dumps(gltf))
    sc = ptb.load_scene_file(path)
    assert sc.tri_mat.tolist() == [0] and sc.mats["base_tex"].tolist() == [0]
    assert len(sc.textures) == 1 and sc.textures[0].shape[0] == 0  # no texels -> placeholder colour on the device
    again = ptb.Scene.from_ptscene_bytes(sc.to_ptscene_bytes())
    assert again.textures[0].shape[0] == 0 and again.mats["base_tex"].tolist() == [0]


def test_unknown_and_broken_files_raise(ptb, core_lib, tmp_path):
    (tmp_path / "a.fbx").write_text("x")
    (tmp_path / "b.glb").write_bytes(b"glTF" + b"\x00" * 30)
    for f in ("a.fbx", "b.glb", "missing.glb"):
        with pytest.raises(ptb.PtError):
            ptb.load_scene_file(tmp_path / f)


def test_ptscene_round_trip_python_and_native(ptb, core_lib, duck, tmp_path):
    raw = duck.to_ptscene_bytes()
    again = ptb.Scene.from_ptscene_bytes(raw)
    assert again.to_ptscene_bytes() == raw
    flat = tmp_path / "duck.ptscene"
    flat.write_bytes(raw)
    native = ptb.load_scene_file(flat)  # .ptscene through Python
    tool = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "ptscene_tool"
    if tool.exists():
        out = tmp_path / "again.ptscene"
        info = json.loads(subprocess.run([str(tool), "convert", str(flat), str(out)], check=True, capture_output=True, text=True).stdout)
        assert info["triangles"] == 4224 and info["per_material"] == [2, 6, 2, 2, 4212]
        assert out.read_bytes() == raw  # native reader + writer reproduce the file byte for byte
    assert np.array_equal(native.tri_pos, duck.tri_pos)


# ------------------------------------------------------------------ BVH
def test_bvh_is_structurally_valid_for_fixtures_and_edge_cases(ptb, core_lib, duck, box):
    for scene, leaf in ((duck, 4), (duck, 1), (duck, 8), (box, 4), (box, 2)):
        ok, msg, st = ptb.bvh_selftest(scene, leaf)
        assert ok, msg
        assert st["bvh_depth"] <= 48 and st["bvh_leaves"] >= len(scene.tri_mat) / leaf
    ok, msg, st = ptb.bvh_selftest(duck, 4)
    assert 1500 < st["bvh_nodes"] < 4224 and st["bvh_depth"] < 30
    # quantised nodes (the self-test also checks that every 15-bit box contains its float box with half a cell to spare):
    # close to the float boxes for the duck, far too coarse for the sphere field inside its 5000-unit sky sphere
    assert 1.0 <= st["quant_inflation"] < 1.05
    field, _ = ptb.scenes.rtow_sphere_field()
    ok, msg, st = ptb.bvh_selftest(field, 4)
    assert ok and st["quant_inflation"] > 3.0, (msg, st["quant_inflation"])
    rng = np.random.default_rng(1)
    mats = np.zeros(1, ptb.MAT_DTYPE)
    mats["type"] = ptb.PT_MAT_UNIVERSAL
    cases = {
        "empty": np.zeros((0, 9), np.float32),
        "single": rng.random((1, 9), dtype=np.float32),
        "all coincident": np.tile(rng.random((1, 9), dtype=np.float32), (37, 1)),
        "collinear centroids": np.stack([np.concatenate([[i, 0, 0], [i + .5, 0, 0], [i, .5, 0]]) for i in range(200)]).astype(np.float32),
        "huge range": (rng.random((500, 9), dtype=np.float32) * np.float32(1e6)) ** 2,
        # exponentially spaced slivers: the SAH peels them off one by one; the builder must switch to median splits in time
        "exponential spacing": np.stack([np.concatenate([[x, 0, 0], [x * 1.0001, 0, 0], [x, x * 0.0001, 0]]) for x in 1.02 ** np.arange(3000)]).astype(np.float32),
    }
    for name, pos in cases.items():
        sc = ptb.Scene(tri_pos=pos, tri_uv=np.zeros((len(pos), 6), np.float32), tri_mat=np.zeros(len(pos), np.int32), mats=mats)
        ok, msg, st = ptb.bvh_selftest(sc, 4)
        assert ok and st["bvh_depth"] <= 48, (name, msg, st["bvh_depth"])
    sph = ptb.Scene(sph=np.array([[0, 0, 0, 1], [3, 0, 0, 0.5], [0, -1000, 0, 999]], np.float32), sph_mat=np.zeros(3, np.int32), mats=mats)
    ok, msg, st = ptb.bvh_selftest(sph, 4)
    assert ok, msg


# ------------------------------------------------------------------ arithmetic folds used by the kernels
def test_double_literal_comparisons_fold_to_float():
    # pt_device.cuh: x < 1e-8 (double)  <=>  x < 1.00000008274037e-08f ;  x > 0.0001 (double)  <=>  x > 0.0001f
    eps_up = np.float32(1.00000008274037e-08)
    around = np.float32(1e-8)
    xs = [np.nextafter(around, np.float32(0), dtype=np.float32), around, np.nextafter(around, np.float32(1), dtype=np.float32), eps_up]
    x = xs[0]
    for _ in range(8):
        x = np.nextafter(x, np.float32(0), dtype=np.float32)
        xs.append(x)
    for x in xs:
        for sgn in (1, -1):
            v = np.float32(sgn) * x
            assert (float(v) < 1e-8 and float(v) > -1e-8) == bool(v < eps_up and v > -eps_up), v
    t = np.float32(0.0001)
    for x in (np.nextafter(t, np.float32(0), dtype=np.float32), t, np.nextafter(t, np.float32(1), dtype=np.float32)):
        assert (float(x) > 0.0001) == bool(x > t)


def test_exhaustive_exactness_of_reciprocal_and_div_by_pi(tmp_path):
    """(float)(1.0/(double)x) == 1.0f/x for every positive normal float; the 3-op double evaluation of c / M_PI
    equals the IEEE double division for every float c in [0, 4]: tests/c/exactness_check.c (about 3 G cases, OpenMP)."""
    exe = tmp_path / "exactness_check"
    subprocess.run(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fopenmp", str(ROOT / "tests" / "c" / "exactness_check.c"), "-lm", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=600).stdout
    assert "double mismatches 0, float-result mismatches 0" in out and "rcp:" in out and out.strip().endswith("mismatches 0"), out


def test_gpu_monitor_accumulators_and_wire_format_without_a_gpu(core_lib):
    """csrc/host/GPUMonitor.h (reference src/Profiling/GPUMonitor.{h,cpp}): NVML is opened at run time, so the test binary
    runs here too (no devices: empty statistics) — accumulators, RENDER_STATS# prefix, monitor thread start/stop."""
    import subprocess
    exe = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "gpu_monitor_test"
    assert exe.exists(), "run __graft_entry__.build()"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "GPU_MONITOR_TEST_OK" in r.stdout, r.stdout


def test_quantised_planes_never_reject_what_float_planes_accept(ptb, core_lib, duck):
    """Host emulation of trav_node_step's two slab tests (same float operations) on pseudo-random rays — aimed at boxes, on their
    faces and edges, axis-parallel, tiny components, finite far limits — against the trees of three scenes."""
    field, _ = ptb.scenes.rtow_sphere_field()
    mesh = ptb.scenes.displaced_sphere_in_cornell(duck, n=60)
    for scene, rays, seed in ((duck, 1500, 1), (field, 2000, 2), (mesh, 800, 3)):
        ok, msg, (tests, on_float, on_quant) = ptb.quant_selftest(scene, rays, seed)
        assert ok, msg
        assert tests > 1_000_000 and on_float > 5_000 and on_quant >= on_float, (tests, on_float, on_quant)


def test_task_generator_and_time_driven_schedulers_tile_the_frame(core_lib):
    """csrc/host/TaskGenerator.h: equal tasks, DYNAMIC tiles, DSFL border nudging (bounded steps, converges to the equal-time
    border) and DSDL time-weighted bisection (reference src/RenderManager.h:264-408,546-639) — every result tiles the frame."""
    import subprocess
    exe = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "task_generator_test"
    assert exe.exists(), "run __graft_entry__.build()"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "TASK_GENERATOR_TEST_OK" in r.stdout, r.stdout[-800:]


def test_lpt_block_order_is_cost_classes_then_z_order(ptb):
    """sched.lpt_block_order (the torch side of TaskGenerator::lptBlockOrder, whose self-test holds the same literal): a permutation,
    highest cost class first, along the Z-order curve within a class; independent of the rank that computes it."""
    import torch
    from ptb200 import sched
    costs = torch.tensor([79, 10, 50, 30, 40, 20, 60, 0], dtype=torch.int32)
    assert sched.lpt_block_order(costs, 4, 4).tolist() == [0, 6, 4, 2, 5, 3, 1, 7]
    g = torch.Generator().manual_seed(7)
    bw, bh = 37, 23
    c = torch.randint(0, 5000, (bw * bh,), generator=g, dtype=torch.int32)
    for levels in (1, 4, 8):
        order = sched.lpt_block_order(c, bw, levels)
        assert sorted(order.tolist()) == list(range(bw * bh))
        cls = (c.long() * levels // (int(c.max()) + 1))[order]
        assert bool((cls[:-1] >= cls[1:]).all())
        z = sched._z_order(order % bw, order // bw)
        same = cls[:-1] == cls[1:]
        assert bool((z[:-1][same] < z[1:][same]).all())
    assert sched.lpt_levels(1) == 8 and sched.lpt_levels(8) == 4


def test_argument_loader_positionals_flags_and_errors(core_lib):
    """csrc/host/ArgumentLoader.h: the reference's two positionals with their defaults (src/ArgumentLoader.h:10-13), every added
    flag, and exceptions that name the offending argument."""
    import subprocess
    exe = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "argument_loader_test"
    assert exe.exists(), "run __graft_entry__.build()"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ARGUMENT_LOADER_TEST_OK" in r.stdout, r.stdout[-800:]


def test_host_restatement_of_the_walk_finds_the_closest_hit_on_both_node_formats(ptb, core_lib, duck, box):
    """pt_walk_selftest: the kernel's resumable walk (held leaf, sentinel stack, two-primitive leaf step, tie rule) restated on
    the host, on float planes, on quantised planes and on the four-wide nodes, against a test of every triangle — camera-like,
    bounce-like (origin on a triangle), vertex-aimed and axis-parallel rays.  A lane state that neither of the kernel's two votes
    would step (a leaf waiting in `cur` with no leaf held: the stall an accepted unused four-wide slot once caused) is an error."""
    mesh = ptb.scenes.displaced_sphere_in_cornell(duck, n=40)
    for scene, rays, seed in ((duck, 12000, 1), (box, 6000, 2), (mesh, 6000, 3)):
        ok, msg, (n, hits, steps_float, steps_quant) = ptb.walk_selftest(scene, rays, seed)
        assert ok, msg
        assert n == rays and hits > 0.9 * rays and steps_float > 5 * rays
        assert steps_quant < 1.05 * steps_float  # the looser boxes cost a few per cent more node steps, not more
    # one triangle: the root's second child is absent, the four-wide root has three unused slots
    import numpy as np
    one = ptb.Scene(tri_pos=np.array([[-1, -1, -3, 1, -1, -3, 0, 1, -3]], np.float32), tri_uv=np.zeros((1, 6), np.float32), tri_mat=np.zeros(1, np.int32),
                    mats=np.array([(ptb.PT_MAT_UNIVERSAL, (0.5, 0.5, 0.5), (1, 1, 1), -1, -1, 0, 1.5)], ptb.MAT_DTYPE))
    ok, msg, (n, hits, _, _) = ptb.walk_selftest(one, 4000, 4)
    assert ok and n == 4000 and hits > 0, msg


# ------------------------------------------------------------------ JPEG textures: pinned to the reference's own decoder (stb_image)
def _gltf_with_image(tmp_path, name, image_bytes, mime):
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    blob = pos.tobytes() + uv.tobytes()
    gltf = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "material": 0}]}],
        "materials": [{"name": "photo", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
        "textures": [{"source": 0}], "images": [{"uri": f"data:{mime};base64," + base64.b64encode(image_bytes).decode()}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
    }
    path = tmp_path / f"{name}.gltf"
    path.write_text(json.dumps(gltf))
    return path


def test_jpeg_textures_decode_to_the_texels_of_the_references_own_decoder(ptb, core_lib, tmp_path):
    """tests/golden/jpeg/*.jpg with, beside each, what the reference's image decoder (third-party/stb_image.h through oracle/_ref/ref_stb,
    oracle/make_golden_jpeg.py) makes of it.  The repo's JPEG reader (csrc/host/JpegDecoder.h: stb_image's inverse DCT, chroma upsampling and
    YCbCr -> RGB restated) must give the SAME bytes: 4:4:4 / 4:2:2 / 4:2:0, sizes that are not multiples of the MCU down to 1x1, restart
    intervals, optimised tables, quality 5 .. 100, grey, progressive (spectral selection + successive approximation).  When the
    reference-derived tool is present (build container) it is also run live."""
    import gzip
    jdir = GOLD / "jpeg"
    names = sorted(p.stem for p in jdir.glob("*.jpg"))
    assert len(names) >= 17
    stb = ROOT / "oracle" / "_ref" / "ref_stb"
    for name in names:
        raw = gzip.decompress((jdir / f"{name}.raw.gz").read_bytes())
        head, body = raw.split(b"\n", 1)
        w, h, c = (int(x) for x in head.split())
        ref = np.frombuffer(body, np.uint8).reshape(h, w, c)
        if stb.exists():
            out = tmp_path / "live.raw"
            subprocess.run([str(stb), str(jdir / f"{name}.jpg"), str(out)], check=True)
            assert out.read_bytes() == raw, name  # the fixture is what the reference's decoder says today
        sc = ptb.load_scene_file(_gltf_with_image(tmp_path, name, (jdir / f"{name}.jpg").read_bytes(), "image/jpeg"))
        tex = sc.textures[0]
        expect = ref if c == 3 else np.repeat(ref, 3, axis=2)  # grey is replicated (stbi_load(path, ..., 3) of the reference's file path)
        assert tex.shape == (h, w, 3), (name, tex.shape)
        assert np.array_equal(tex, expect.astype(np.float32)), (name, int((tex != expect).sum()))


def test_bmp_and_tga_textures_decode_to_the_texels_of_the_references_own_decoder(ptb, core_lib, tmp_path):
    """tests/golden/images/*.bmp|tga|gif|pgm|ppm (GIF: first frame, sub-rectangles, background index, local tables, transparency, interlace;
    PNM: P5 / P6, 8- and 16-bit, comments in the header) with, beside each, what the reference's decoder returns for a texture FILE (stb_image with three
    requested channels, src/HostScene.cpp:29; oracle/make_golden_images.py).  csrc/host/BmpTgaDecoder.h must give the SAME bytes for every
    header size, bit depth, palette, mask, orientation and RLE variant there — including stb_image's own quirk for a gap between a BMP
    header and its pixels — and must refuse (placeholder texture) what stb_image refuses.  When the reference-derived tool is present
    (build container) it is also run live."""
    import gzip
    idir = GOLD / "images"
    files = sorted(p for p in idir.iterdir() if p.suffix in (".bmp", ".tga", ".gif", ".pgm", ".ppm"))
    refused = json.loads((idir / "refused.json").read_text())
    assert len(files) >= 55 and len(refused) >= 1 and sum(p.suffix == ".gif" for p in files) >= 12
    stb = ROOT / "oracle" / "_ref" / "ref_stb"
    for path in files:
        sc = ptb.load_scene_file(_gltf_with_image(tmp_path, path.stem, path.read_bytes(), "image/" + path.suffix[1:]))
        tex = sc.textures[0]
        if path.name in refused:
            assert tex.shape[0] == 0, path.name  # no texels: shaded with the reference's placeholder colour
            continue
        raw = gzip.decompress((idir / f"{path.stem}.raw.gz").read_bytes())
        head, body = raw.split(b"\n", 1)
        w, h, c = (int(x) for x in head.split())
        assert c == 3
        ref = np.frombuffer(body, np.uint8).reshape(h, w, 3)
        if stb.exists():
            out = tmp_path / "live.raw"
            subprocess.run([str(stb), str(path), str(out), "3"], check=True)
            assert out.read_bytes() == raw, path.name  # the fixture is what the reference's decoder says today
        assert tex.shape == (h, w, 3), (path.name, tex.shape)
        assert np.array_equal(tex, ref.astype(np.float32)), (path.name, int((tex != ref).sum()))
    # truncated and corrupted files never crash the loader: they decode (missing bytes read as zero, as in stb_image) or fall back to the placeholder
    rng = np.random.default_rng(5)
    for path in files[::3]:
        data = path.read_bytes()
        for cut in (0, 1, 10, 19, len(data) // 2, len(data) - 1):
            sc = ptb.load_scene_file(_gltf_with_image(tmp_path, "cut", data[:cut] or b"\0", "image/" + path.suffix[1:]))
            assert len(sc.textures) == 1
        noisy = bytearray(data)
        for k in rng.integers(0, min(len(noisy), 64), 6):
            noisy[int(k)] = int(rng.integers(0, 256))
        noisy[12:16] = b"\x10\x00\x10\x00" if path.suffix == ".tga" else noisy[12:16]  # keep a corrupted TGA small (16 x 16)
        if path.suffix == ".bmp":
            noisy[18:26] = (16).to_bytes(4, "little") + (16).to_bytes(4, "little")
        sc = ptb.load_scene_file(_gltf_with_image(tmp_path, "noisy", bytes(noisy), "image/" + path.suffix[1:]))
        assert len(sc.textures) == 1
    # a corrupted header that announces a gigantic image is refused at once (stb_image's limits), not allocated: PNG raises like the reference
    # ("Cannot load texture data"), JPEG falls back to the placeholder
    png = bytearray((GOLD / "png" / "type_0_depth_8_interlace_false_37x27.png").read_bytes())
    png[16:20] = bytes([0x6D, 0, 0, 0x25])  # width 1 828 716 581
    with pytest.raises(Exception, match="Very large image"):
        ptb.load_scene_file(_gltf_with_image(tmp_path, "huge", bytes(png), "image/png"))
    jpg = bytearray((GOLD / "jpeg" / "c444_q90_33x21.jpg").read_bytes())
    sof = jpg.find(bytes([0xFF, 0xC0]))
    jpg[sof + 5:sof + 9] = bytes([0xFF, 0xFF, 0xFF, 0xFF])  # 65535 x 65535
    assert ptb.load_scene_file(_gltf_with_image(tmp_path, "hugej", bytes(jpg), "image/jpeg")).textures[0].shape[0] == 0


def test_obj_material_texture_files_follow_the_references_binding(ptb, core_lib, tmp_path):
    """An .mtl with map_Kd (a BMP) and map_Ke (a TGA): the reference loads every texture file with stbi_load(path, ..., 3) but binds
    only assimp's BASE_COLOR and EMISSIVE types (src/HostScene.cpp:52-72,174-184) — for an OBJ that is map_Ke; map_Kd takes a slot in the
    texture list and is bound to nothing.  A texture file that does not exist aborts the load, as there."""
    import gzip, shutil
    idir = GOLD / "images"
    shutil.copy(idir / "bmp24_hdr40_33x17.bmp", tmp_path / "kd.bmp")
    shutil.copy(idir / "tga24_rle_29x14.tga", tmp_path / "ke.tga")
    (tmp_path / "m.mtl").write_text("newmtl glow\nKd 0.2 0.3 0.4\nKe 1 1 1\nmap_Kd -s 1 1 1 kd.bmp\nmap_Ke ke.tga\nnewmtl plain\nKd 0.5 0.5 0.5\n")
    (tmp_path / "s.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl glow\nf 1/1 2/2 3/3\nusemtl plain\nf 1/1 3/3 4/4\n")
    sc = ptb.load_scene_file(tmp_path / "s.obj")
    assert len(sc.textures) == 2 and sc.textures[0].shape == (17, 33, 3) and sc.textures[1].shape == (14, 29, 3)
    glow, plain = sc.mats[0], sc.mats[1]
    assert int(glow["base_tex"]) == -1 and int(glow["emis_tex"]) == 1 and int(plain["base_tex"]) == -1 and int(plain["emis_tex"]) == -1
    raw = gzip.decompress((idir / "tga24_rle_29x14.raw.gz").read_bytes()).split(b"\n", 1)[1]
    assert np.array_equal(sc.textures[1], np.frombuffer(raw, np.uint8).reshape(14, 29, 3).astype(np.float32))
    (tmp_path / "m.mtl").write_text("newmtl glow\nmap_Ke missing.png\n")
    with pytest.raises(Exception, match="Cannot load texture data"):
        ptb.load_scene_file(tmp_path / "s.obj")


def test_png_textures_decode_to_the_texels_of_the_references_own_decoder(ptb, core_lib, tmp_path):
    """tests/golden/png/*.png — written byte by byte (oracle/make_golden_png.py): every colour type and bit depth, a random filter type per
    row, Adam7 interlacing, tRNS for palettes and as a colour key, split IDAT — with what the reference's decoder returns for an embedded
    texture (stb_image, native channel count).  The reference then walks that buffer three bytes per texel WHATEVER the channel count
    (src/HostScene.cpp:37-46: alpha and grey bytes become texels; with fewer than three channels it runs off the buffer — only the texels
    that lie inside it are compared).  The repo's PNG reader must give the same bytes."""
    import gzip
    pdir = GOLD / "png"
    files = sorted(pdir.glob("*.png"))
    assert len(files) >= 26
    stb = ROOT / "oracle" / "_ref" / "ref_stb"
    channels_seen = set()
    for path in files:
        raw = gzip.decompress((pdir / f"{path.stem}.raw.gz").read_bytes())
        head, body = raw.split(b"\n", 1)
        w, h, c = (int(x) for x in head.split())
        channels_seen.add(c)
        if stb.exists():
            out = tmp_path / "live.raw"
            subprocess.run([str(stb), str(path), str(out)], check=True)
            assert out.read_bytes() == raw, path.name
        flat = np.frombuffer(body, np.uint8)
        n_tex = min(w * h, len(flat) // 3)
        tex = ptb.load_scene_file(_gltf_with_image(tmp_path, path.stem, path.read_bytes(), "image/png")).textures[0]
        assert tex.shape == (h, w, 3), (path.name, tex.shape)
        assert np.array_equal(tex.reshape(-1, 3)[:n_tex], flat[:n_tex * 3].reshape(n_tex, 3).astype(np.float32)), path.name
    assert channels_seen == {1, 2, 3, 4}


def test_malformed_gltf_raises_and_never_crashes(ptb, core_lib, tmp_path):
    """Found by tools/fuzz_loader.py (byte mutations of the reference's models, each load in a child process): accessors without
    `componentType` / `type` / `count`, texture slots without `index`, negative or absurd counts, offsets and strides, TEXCOORD_0 shorter than
    POSITION.  Every one raises; none dereferences a missing member or wraps a size around."""
    import base64, copy
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32).tobytes() + bytes([0, 1, 2, 0])
    good = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}]}],
            "materials": [{"name": "m", "emissiveFactor": [1, 1, 1], "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
            "textures": [{"source": 0}], "images": [{"uri": "data:image/png;base64,"}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC2"},
                          {"bufferView": 1, "componentType": 5121, "count": 3, "type": "SCALAR"}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 4}],
            "buffers": [{"byteLength": 40, "uri": "data:application/octet-stream;base64," + base64.b64encode(tri).decode()}]}

    def load(mutate):
        g = copy.deepcopy(good)
        mutate(g)
        path = tmp_path / "m.gltf"
        path.write_text(json.dumps(g))
        return ptb.load_scene_file(path)

    def drop(path_keys):
        def f(g):
            o = g
            for k in path_keys[:-1]:
                o = o[k]
            del o[path_keys[-1]]
        return f

    def put(path_keys, value):
        def f(g):
            o = g
            for k in path_keys[:-1]:
                o = o[k]
            o[path_keys[-1]] = value
        return f
    bad = [drop(["accessors", 0, "componentType"]), drop(["accessors", 0, "type"]), drop(["accessors", 0, "count"]), drop(["accessors", 2, "componentType"]),
           drop(["accessors", 2, "count"]), drop(["accessors", 2, "bufferView"]),
           put(["accessors", 0, "count"], -1), put(["accessors", 0, "count"], 1e30), put(["accessors", 0, "count"], 2 ** 61), put(["accessors", 2, "count"], 2 ** 62),
           put(["accessors", 0, "byteOffset"], -4), put(["accessors", 0, "byteOffset"], 2 ** 63), put(["bufferViews", 0, "byteStride"], 2 ** 62),
           put(["bufferViews", 0, "byteLength"], -36), put(["bufferViews", 0, "byteOffset"], 1e18), put(["bufferViews", 0, "buffer"], 7),
           put(["accessors", 1, "count"], 2), put(["accessors", 0, "componentType"], 1234), put(["accessors", 0, "type"], "MAT9"), put(["accessors", 2, "count"], "three")]
    bad.append(put(["nodes", 0, "children"], [0]))  # a node that is its own child
    for k, m in enumerate(bad):
        with pytest.raises(Exception):
            load(m)
    deep = tmp_path / "deep.gltf"
    deep.write_text('{"asset": ' + "[" * 100000 + "]" * 100000 + "}")  # 100 000 nested arrays: a parse error, not a stack overflow
    with pytest.raises(Exception, match="nesting too deep"):
        ptb.load_scene_file(deep)
    # (the unmutated file is fine apart from its empty image: that one loads with a placeholder texture)
    sc = load(lambda g: None)
    assert len(sc.tri_mat) == 1 and len(sc.textures) == 1
    sc = load(drop(["materials", 0, "pbrMetallicRoughness", "baseColorTexture", "index"]))  # a slot without an index is no texture
    assert len(sc.tri_mat) == 1 and len(sc.textures) == 0


def test_corrupt_ptscene_files_raise_before_anything_is_allocated(ptb, core_lib, tmp_path):
    """The native .ptscene reader (csrc/host/SceneLoader.cpp::read_ptscene, the input of cuda_project and ptscene_tool): counts in a
    corrupted header are checked against the bytes that are there before any table is sized, material types and every index against their
    ranges.  Found with 600 mutated files under ASan (50 findings before the checks: multi-GB reservations, invalid enum loads)."""
    import gzip, struct
    exe = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "ptscene_tool"
    good = gzip.decompress((GOLD / "random" / "case_1.ptscene.gz").read_bytes())
    magic, ver, nt, ns, nm, ntex = struct.unpack_from("<4sIIIII", good, 0)
    assert magic == b"PTSC" and nt > 0 and nm > 0 and ntex > 0

    def run(data):
        f = tmp_path / "x.ptscene"
        f.write_bytes(data)
        return subprocess.run([str(exe), "info", str(f)], capture_output=True, text=True, timeout=60)
    assert run(good).returncode == 0
    bad = []
    for field, value in ((8, 0x7fffffff), (12, 0x40000000), (16, 0xffffffff), (20, 0x10000000)):  # n_tris, n_spheres, n_mats, n_tex
        b = bytearray(good)
        struct.pack_into("<I", b, field, value)
        bad.append(bytes(b))
    b = bytearray(good); struct.pack_into("<i", b, 24 + 60, nm + 5); bad.append(bytes(b))          # first triangle's material index
    b = bytearray(good); struct.pack_into("<i", b, 24 + nt * 68 + ns * 20, 77); bad.append(bytes(b))  # first material's type
    b = bytearray(good); struct.pack_into("<i", b, 24 + nt * 68 + ns * 20 + 28, ntex + 1); bad.append(bytes(b))  # its base-colour texture
    b = bytearray(good); struct.pack_into("<ii", b, 24 + nt * 68 + ns * 20 + nm * 44, 1 << 20, 1 << 20); bad.append(bytes(b))  # first texture: 2^40 texels
    bad.append(good[:len(good) // 2])
    for data in bad:
        r = run(data)
        assert r.returncode not in (0, -11, -6) and "ptscene" in (r.stderr + r.stdout), (r.returncode, (r.stderr + r.stdout)[-200:])


def test_the_abi_header_is_plain_c(tmp_path):
    """include/ptcore.h is the drop-in boundary: it must compile as C99 (pedantic) and as C++11 with nothing but itself — no torch, CUDA or
    STL type in any signature."""
    src = tmp_path / "t.c"
    src.write_text('#include "ptcore.h"\nint main(void) { PtCamera c; PtStats s; (void)c; (void)s; return ptcore_last_error(0) == 0; }\n')
    inc = str(ROOT / "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)
    import re
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "ptcore.h").read_text(), flags=re.S)  # declarations only, comments removed
    for banned in ("#include <cuda", "torch", "std::", "cudaStream_t", "at::"):
        assert banned not in text, banned
