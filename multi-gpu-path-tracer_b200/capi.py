"""ctypes binding of include/ptcore.h (libptcore.so) — host-side plumbing only.

The product path is the CUDA library; nothing here computes pixels.  If the library is
missing or no CUDA device is present the calls fail loudly (PtError) — there is no CPU
fallback (the CPU restatement under oracle/ is test infrastructure and is never imported
from this package).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "_lib" / "libptcore.so"

PT_MAT_LAMBERTIAN, PT_MAT_METAL, PT_MAT_DIELECTRIC, PT_MAT_DIFFUSE_LIGHT, PT_MAT_UNIVERSAL = range(5)
PT_OPT_KERNEL, PT_OPT_COUNT_TESTS, PT_OPT_BVH_LEAF_MAX, PT_OPT_BLOCKS_PER_SM, _PT_OPT_RESERVED_5, PT_OPT_REFILL_AT, PT_OPT_NODE_BURST, PT_OPT_MIN_BLOCKS, PT_OPT_BVH_WIDTH = range(1, 10)
PT_KERNEL_PERSISTENT, PT_KERNEL_DIRECT, PT_KERNEL_LOCKSTEP, PT_KERNEL_POOL = 0, 1, 2, 3
PT_OPT_NODE_FORMAT = 10
PT_OPT_SAH_INTERSECT_COST = 11
PT_OPT_POOL_SLOTS, PT_OPT_POOL_IDLE_AT, PT_OPT_WATCHDOG, PT_OPT_POOL_PERIOD, PT_OPT_POOL_CARVEOUT, PT_OPT_SMEM_NODES, PT_OPT_LANES_PER_WARP, PT_OPT_STICKY_TEXTURES, PT_OPT_RNG_MODE, PT_OPT_RNG_CHUNKS, PT_OPT_L2_PERSIST_NODES = 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22
PT_OPT_GRID_CTAS, PT_OPT_CTA_WARPS = 23, 24
PT_RNG_STREAM, PT_RNG_SAMPLE_KEYED = 0, 1
PT_NODES_AUTO, PT_NODES_FULL, PT_NODES_QUANTISED = 0, 1, 2


class PtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ptcore error {code}: {msg}")
        self.code = code


class PtMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("base", C.c_float * 3), ("emis", C.c_float * 3), ("base_tex", C.c_int32),
                ("emis_tex", C.c_int32), ("fuzz", C.c_float), ("ior", C.c_float)]


class PtTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_float))]


class PtSceneDesc(C.Structure):
    _fields_ = [("n_tris", C.c_int32), ("tri_pos", C.POINTER(C.c_float)), ("tri_uv", C.POINTER(C.c_float)), ("tri_mat", C.POINTER(C.c_int32)),
                ("n_spheres", C.c_int32), ("sph", C.POINTER(C.c_float)), ("sph_mat", C.POINTER(C.c_int32)),
                ("n_mats", C.c_int32), ("mats", C.POINTER(PtMaterial)), ("n_tex", C.c_int32), ("tex", C.POINTER(PtTexture))]


class PtCamera(C.Structure):
    _fields_ = [("look_from", C.c_float * 3), ("front", C.c_float * 3), ("vfov", C.c_float), ("hfov", C.c_float)]


class PtTile(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset_x", C.c_int32), ("offset_y", C.c_int32)]


class PtStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64), ("light_tests", C.c_uint64),
                ("launches", C.c_uint64), ("bvh_nodes", C.c_uint32), ("bvh_leaves", C.c_uint32), ("bvh_depth", C.c_uint32), ("n_lights", C.c_uint32),
                ("bvh_build_ms", C.c_double), ("sah_cost", C.c_double), ("scene_bytes", C.c_uint64), ("bvh4_nodes", C.c_uint32), ("bvh4_depth", C.c_uint32), ("quant_inflation", C.c_double),
                ("n_vertices", C.c_uint32), ("reserved0", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


MAT_DTYPE = np.dtype([("type", "<i4"), ("base", "<f4", (3,)), ("emis", "<f4", (3,)), ("base_tex", "<i4"), ("emis_tex", "<i4"), ("fuzz", "<f4"), ("ior", "<f4")])
assert MAT_DTYPE.itemsize == C.sizeof(PtMaterial) == 44

# default camera of the reference: src/main.cu:40, src/CameraConfig.h:6
DEFAULT_CAMERA = dict(look_from=(0.0, 0.0, 0.5), front=(0.0, 0.0, -0.5), vfov=45.0, hfov=45.0)


def make_camera(look_from=(0.0, 0.0, 0.5), front=(0.0, 0.0, -0.5), vfov=45.0, hfov=45.0) -> PtCamera:
    return PtCamera((C.c_float * 3)(*look_from), (C.c_float * 3)(*front), vfov, hfov)


@dataclass
class Scene:
    """Flat host scene (numpy), the Python-side twin of HostScene (reference src/HostScene.h:54-58)."""
    tri_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    tri_uv: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    tri_mat: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    sph: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    sph_mat: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    mats: np.ndarray = field(default_factory=lambda: np.zeros((0,), MAT_DTYPE))
    textures: List[np.ndarray] = field(default_factory=list)  # each (h, w, 3) float32 in 0..255

    def normalised(self) -> "Scene":
        return Scene(np.ascontiguousarray(self.tri_pos, np.float32).reshape(-1, 9), np.ascontiguousarray(self.tri_uv, np.float32).reshape(-1, 6),
                     np.ascontiguousarray(self.tri_mat, np.int32).reshape(-1), np.ascontiguousarray(self.sph, np.float32).reshape(-1, 4),
                     np.ascontiguousarray(self.sph_mat, np.int32).reshape(-1), np.ascontiguousarray(self.mats, MAT_DTYPE).reshape(-1),
                     [np.ascontiguousarray(t, np.float32) for t in self.textures])

    def desc(self):
        """Returns (PtSceneDesc, keepalive) — keep `keepalive` referenced while the desc is in use."""
        s = self.normalised()
        n_tex = len(s.textures)
        texs = (PtTexture * max(n_tex, 1))()
        for i, t in enumerate(s.textures):
            texs[i].height, texs[i].width = int(t.shape[0]), int(t.shape[1])
            texs[i].rgb = t.ctypes.data_as(C.POINTER(C.c_float))
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)) if a.size else None
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32)) if a.size else None
        d = PtSceneDesc(len(s.tri_mat), fp(s.tri_pos), fp(s.tri_uv), ip(s.tri_mat), len(s.sph_mat), fp(s.sph), ip(s.sph_mat),
                        len(s.mats), s.mats.ctypes.data_as(C.POINTER(PtMaterial)) if len(s.mats) else None, n_tex, texs)
        return d, (s, texs)

    # ---- .ptscene / .ptscene.gz (layout: oracle/ptscene_io.h) ----
    @staticmethod
    def from_ptscene_bytes(b: bytes) -> "Scene":
        magic, ver, nt, ns, nm, ntex = struct.unpack_from("<4sIIIII", b, 0)
        if magic != b"PTSC" or ver != 1:
            raise ValueError("not a ptscene v1 file")
        off = 24
        tri = np.frombuffer(b, dtype=np.dtype([("pos", "<f4", (9,)), ("uv", "<f4", (6,)), ("mat", "<i4"), ("tex", "<i4")]), count=nt, offset=off)
        off += nt * 68
        sp = np.frombuffer(b, dtype=np.dtype([("c", "<f4", (4,)), ("mat", "<i4")]), count=ns, offset=off)
        off += ns * 20
        mats = np.frombuffer(b, dtype=MAT_DTYPE, count=nm, offset=off).copy()
        off += nm * 44
        texs = []
        for _ in range(ntex):
            w, h = struct.unpack_from("<ii", b, off)
            off += 8
            texs.append(np.frombuffer(b, dtype=np.uint8, count=w * h * 3, offset=off).astype(np.float32).reshape(h, w, 3))
            off += w * h * 3
        return Scene(tri["pos"].copy(), tri["uv"].copy(), tri["mat"].copy(), sp["c"].copy(), sp["mat"].copy(), mats, texs)

    @staticmethod
    def load_ptscene(path) -> "Scene":
        path = Path(path)
        raw = path.read_bytes()
        if path.suffix == ".gz":
            raw = gzip.decompress(raw)
        return Scene.from_ptscene_bytes(raw)

    def to_ptscene_bytes(self) -> bytes:
        s = self.normalised()
        out = [struct.pack("<4sIIIII", b"PTSC", 1, len(s.tri_mat), len(s.sph_mat), len(s.mats), len(s.textures))]
        tri = np.zeros(len(s.tri_mat), dtype=np.dtype([("pos", "<f4", (9,)), ("uv", "<f4", (6,)), ("mat", "<i4"), ("tex", "<i4")]))
        tri["pos"], tri["uv"], tri["mat"], tri["tex"] = s.tri_pos, s.tri_uv, s.tri_mat, -1
        out.append(tri.tobytes())
        sp = np.zeros(len(s.sph_mat), dtype=np.dtype([("c", "<f4", (4,)), ("mat", "<i4")]))
        sp["c"], sp["mat"] = s.sph, s.sph_mat
        out.append(sp.tobytes())
        out.append(s.mats.tobytes())
        for t in s.textures:
            out.append(struct.pack("<ii", t.shape[1], t.shape[0]))
            out.append(np.clip(t, 0, 255).astype(np.uint8).tobytes())
        return b"".join(out)

    def save_ptscene(self, path) -> None:
        path = Path(path)
        raw = self.to_ptscene_bytes()
        path.write_bytes(gzip.compress(raw, 9, mtime=0) if path.suffix == ".gz" else raw)

    @staticmethod
    def from_desc(d: PtSceneDesc) -> "Scene":
        arr = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n and p else np.zeros((0,), dt)
        mats = np.zeros(d.n_mats, MAT_DTYPE)
        if d.n_mats:
            mats = np.frombuffer(C.string_at(d.mats, d.n_mats * 44), dtype=MAT_DTYPE).copy()
        texs = []
        for i in range(d.n_tex):
            t = d.tex[i]
            texs.append(np.ctypeslib.as_array(t.rgb, shape=(t.height, t.width, 3)).astype(np.float32).copy() if t.rgb else np.zeros((0, t.width, 3), np.float32))  # no texels: height 0 = the reference's 'texture without data' (placeholder colour)
        return Scene(arr(d.tri_pos, d.n_tris * 9, np.float32).reshape(-1, 9), arr(d.tri_uv, d.n_tris * 6, np.float32).reshape(-1, 6),
                     arr(d.tri_mat, d.n_tris, np.int32), arr(d.sph, d.n_spheres * 4, np.float32).reshape(-1, 4), arr(d.sph_mat, d.n_spheres, np.int32), mats, texs)


_lib: Optional[C.CDLL] = None


def load_library(path: Optional[Path] = None) -> C.CDLL:
    """Loads libptcore.so (building nothing: see _build.py). Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("PTB200_LIBPTCORE", LIB_PATH))  # the variable: A/B builds of the same ABI
    if not p.exists():
        raise FileNotFoundError(f"{p} is missing — run __graft_entry__.build() (nvcc, sm_100a); there is no fallback path")
    lib = C.CDLL(str(p))
    vp, i32, u32, i64, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64
    sig = {
        "ptcore_abi_version": (C.c_int, []),
        "ptcore_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "ptcore_destroy": (C.c_int, [vp]),
        "ptcore_last_error": (C.c_char_p, [vp]),
        "ptcore_upload_scene": (C.c_int, [vp, C.POINTER(PtSceneDesc)]),
        "ptcore_reupload_scene": (C.c_int, [vp, vp, C.POINTER(u64)]),
        "ptcore_set_camera": (C.c_int, [vp, C.POINTER(PtCamera)]),
        "ptcore_set_params": (C.c_int, [vp, u32, u32]),
        "ptcore_set_thread_block_size": (C.c_int, [vp, u32, u32]),
        "ptcore_set_option": (C.c_int, [vp, C.c_int, i64]),
        "ptcore_bind_framebuffer": (C.c_int, [vp, vp, vp, u32, u32]),
        "ptcore_render_tile_async": (C.c_int, [vp, i32, i32, i32, i32, vp]),
        "ptcore_render_tiles_async": (C.c_int, [vp, C.POINTER(PtTile), i32, vp]),
        "ptcore_render_blocks_async": (C.c_int, [vp, vp, u32, vp]),
        "ptcore_block_costs_async": (C.c_int, [vp, u32, vp, vp]),
        "ptcore_block_costs_range_async": (C.c_int, [vp, u32, vp, u32, u32, vp]),
        "ptcore_set_retire_log": (C.c_int, [vp, vp, u32]),
        "ptcore_gather_blocks_async": (C.c_int, [vp, vp, vp, vp, u32, vp]),
        "ptcore_render_keyed_async": (C.c_int, [vp, vp, u32, vp, u32, u32, u32, vp]),
        "ptcore_resolve_keyed_async": (C.c_int, [vp, vp, u32, vp]),
        "ptcore_sync": (C.c_int, [vp, vp]),
        "ptcore_wait": (C.c_int, [vp]),
        "ptcore_render_frame_host": (C.c_int, [vp, u32, u32, vp, vp]),
        "ptcore_get_stats": (C.c_int, [vp, C.POINTER(PtStats)]),
        "ptcore_reset_stats": (C.c_int, [vp]),
        "ptcore_debug_trace_pixel": (C.c_int, [vp, u32, u32, i32, i32, vp, i32, C.POINTER(i32), vp]),
        "pt_bvh_selftest": (C.c_int, [C.POINTER(PtSceneDesc), i32, C.POINTER(PtStats), C.c_char_p, C.c_size_t]),
        "pt_quant_selftest": (C.c_int, [C.POINTER(PtSceneDesc), u32, u32, C.POINTER(u64), C.c_char_p, C.c_size_t]),
        "pt_walk_selftest": (C.c_int, [C.POINTER(PtSceneDesc), u32, u32, C.POINTER(u64), C.c_char_p, C.c_size_t]),
        "pt_tileq_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "pt_tileq_claim": (i64, [vp, i64, i64]),
        "pt_tileq_reset": (C.c_int, [vp]),
        "pt_tileq_close": (C.c_int, [vp, C.c_int]),
        "ptscene_load": (C.c_int, [C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_size_t]),
        "ptscene_desc": (C.POINTER(PtSceneDesc), [vp]),
        "ptscene_save": (C.c_int, [vp, C.c_char_p]),
        "ptscene_free": (None, [vp]),
        "pt_write_ppm": (C.c_int, [C.c_char_p, vp, u32, u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def load_scene_file(path) -> Scene:
    """Any supported scene file -> Scene.  .ptscene(.gz) is read in Python; .glb/.gltf/.obj go through
    the library's loader (SceneLoader::load of the host API)."""
    path = Path(path)
    if path.name.endswith(".ptscene") or path.name.endswith(".ptscene.gz"):
        return Scene.load_ptscene(path)
    lib = load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.ptscene_load(str(path).encode(), C.byref(h), err, 512)
    if rc != 0:
        raise PtError(rc, err.value.decode(errors="replace"))
    try:
        return Scene.from_desc(lib.ptscene_desc(h).contents)
    finally:
        lib.ptscene_free(h)


class TileQueue:
    """Node-wide dynamic tile counter (POSIX shared memory) shared by the ranks of one box."""

    def __init__(self, name: str, create: bool):
        self.lib = load_library()
        self.h = C.c_void_p()
        self.name = name
        self.owner = create
        rc = self.lib.pt_tileq_open(name.encode(), 1 if create else 0, C.byref(self.h))
        if rc != 0:
            raise PtError(rc, f"pt_tileq_open({name})")

    def claim(self, count: int, limit: int) -> int:
        return int(self.lib.pt_tileq_claim(self.h, count, limit))

    def reset(self) -> None:
        self.lib.pt_tileq_reset(self.h)

    def close(self) -> None:
        if self.h:
            self.lib.pt_tileq_close(self.h, 1 if self.owner else 0)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PathTracer:
    """Thin RAII wrapper over a ptcore handle (the C ABI twin of the reference's DevicePathTracer)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.ptcore_create(device, C.byref(self.h))
        if rc != 0:
            raise PtError(rc, (self.lib.ptcore_last_error(None) or b"").decode())
        self.device = device
        self._keep = None

    def _ck(self, rc: int) -> None:
        if rc != 0:
            raise PtError(rc, (self.lib.ptcore_last_error(self.h) or b"").decode())

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.ptcore_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene: Scene) -> None:
        d, keep = scene.desc()
        self._ck(self.lib.ptcore_upload_scene(self.h, C.byref(d)))

    def reupload_scene(self, stream: int = 0) -> int:
        n = C.c_uint64()
        self._ck(self.lib.ptcore_reupload_scene(self.h, stream, C.byref(n)))
        return n.value

    def set_camera(self, **kw) -> None:
        cam = make_camera(**{**DEFAULT_CAMERA, **kw})
        self._ck(self.lib.ptcore_set_camera(self.h, C.byref(cam)))

    def set_params(self, spp: int, depth: int) -> None:
        self._ck(self.lib.ptcore_set_params(self.h, spp, depth))

    def set_option(self, key: int, value: int) -> None:
        self._ck(self.lib.ptcore_set_option(self.h, key, value))

    def bind_framebuffer(self, rgb_ptr: int, yuv_ptr: int, width: int, height: int) -> None:
        self._ck(self.lib.ptcore_bind_framebuffer(self.h, rgb_ptr, yuv_ptr or None, width, height))

    def render_tile_async(self, ox: int, oy: int, w: int, h: int, stream: int = 0) -> None:
        self._ck(self.lib.ptcore_render_tile_async(self.h, ox, oy, w, h, stream or None))

    def render_tiles_async(self, tiles: Sequence[Sequence[int]], stream: int = 0) -> None:
        """tiles: iterable of (offset_x, offset_y, width, height)."""
        arr = (PtTile * max(1, len(tiles)))()
        for i, (ox, oy, w, h) in enumerate(tiles):
            arr[i] = PtTile(w, h, ox, oy)
        self._ck(self.lib.ptcore_render_tiles_async(self.h, arr, len(tiles), stream or None))

    def render_blocks_async(self, blocks_dev_ptr: int, n_blocks: int, stream: int = 0) -> None:
        """blocks: DEVICE uint32 array, entry = bx | (by << 16) of an 8x4-pixel block; lanes take them in list order."""
        self._ck(self.lib.ptcore_render_blocks_async(self.h, blocks_dev_ptr, n_blocks, stream or None))

    def block_costs_async(self, pilot_spp: int, costs_dev_ptr: int, stream: int = 0) -> None:
        """Pilot pass: rays per 8x4 block over pilot_spp samples into a DEVICE uint32[ceil(W/8)*ceil(H/4)] array."""
        self._ck(self.lib.ptcore_block_costs_async(self.h, pilot_spp, costs_dev_ptr, stream or None))

    def block_costs_range_async(self, pilot_spp: int, costs_dev_ptr: int, first_block: int, n_blocks: int, stream: int = 0) -> None:
        """Pilot pass over blocks [first_block, first_block + n_blocks) of the row-major block grid only."""
        self._ck(self.lib.ptcore_block_costs_range_async(self.h, pilot_spp, costs_dev_ptr, first_block, n_blocks, stream or None))

    def render_keyed_async(self, accum_dev_ptr: int, n_chunks: int, first_chunk: int = 0, chunk_step: int = 1, blocks_dev_ptr: int = 0, n_blocks: int = 0, stream: int = 0) -> None:
        """PT_RNG_SAMPLE_KEYED: chunks first_chunk, first_chunk + chunk_step, ... of every pixel (of the listed blocks) into accum[chunk][pixel][3]."""
        self._ck(self.lib.ptcore_render_keyed_async(self.h, blocks_dev_ptr or None, n_blocks, accum_dev_ptr, n_chunks, first_chunk, chunk_step, stream or None))

    def resolve_keyed_async(self, accum_dev_ptr: int, n_chunks: int, stream: int = 0) -> None:
        self._ck(self.lib.ptcore_resolve_keyed_async(self.h, accum_dev_ptr, n_chunks, stream or None))

    def gather_blocks_async(self, dst_rgb_ptr: int, dst_yuv_ptr: int, blocks_dev_ptr: int, n_blocks: int, stream: int = 0) -> None:
        """Copies the listed blocks of the bound framebuffer into another frame of the same size (peer memory allowed)."""
        self._ck(self.lib.ptcore_gather_blocks_async(self.h, dst_rgb_ptr, dst_yuv_ptr or None, blocks_dev_ptr, n_blocks, stream or None))

    def set_retire_log(self, log_dev_ptr: int, n_warps: int) -> None:
        self._ck(self.lib.ptcore_set_retire_log(self.h, log_dev_ptr or None, n_warps))

    def sync(self, stream: int = 0) -> None:
        self._ck(self.lib.ptcore_sync(self.h, stream or None))

    def wait(self) -> None:
        self._ck(self.lib.ptcore_wait(self.h))

    def render_frame_host(self, width: int, height: int, want_yuv: bool = True, rgb: Optional[np.ndarray] = None, yuv: Optional[np.ndarray] = None):
        if rgb is None:
            rgb = np.empty((height, width, 3), np.uint8)
        if want_yuv and yuv is None:
            yuv = np.empty((width * height * 3 // 2,), np.uint8)
        self._ck(self.lib.ptcore_render_frame_host(self.h, width, height, rgb.ctypes.data, yuv.ctypes.data if want_yuv else None))
        return rgb, (yuv if want_yuv else None)

    def trace_pixel(self, width: int, height: int, x: int, y: int, max_events: int = 4096):
        """Parity probe: (events[n,16] float32, col[3]) for one pixel with the current scene/camera/params."""
        ev = np.zeros((max_events, 16), np.float32)
        col = np.zeros(3, np.float32)
        n = C.c_int32()
        self._ck(self.lib.ptcore_debug_trace_pixel(self.h, width, height, x, y, ev.ctypes.data, max_events, C.byref(n), col.ctypes.data))
        return ev[: min(n.value, max_events)], col

    def stats(self) -> dict:
        st = PtStats()
        self._ck(self.lib.ptcore_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def reset_stats(self) -> None:
        self._ck(self.lib.ptcore_reset_stats(self.h))


def quant_selftest(scene: Scene, n_rays: int = 2000, seed: int = 1):
    """(ok, message, (tests, accepted on float planes, accepted on quantised planes)): host emulation of the kernel's slab tests."""
    d, keep = scene.desc()
    counts = (C.c_uint64 * 3)()
    msg = C.create_string_buffer(256)
    rc = load_library().pt_quant_selftest(C.byref(d), n_rays, seed, counts, msg, 256)
    return rc == 0, msg.value.decode(), tuple(int(c) for c in counts)


def walk_selftest(scene: Scene, n_rays: int = 2000, seed: int = 1):
    """(ok, message, (rays, hits, node steps on float planes, node steps on quantised planes)): host restatement of the kernel's
    walk on both node formats against a test of every triangle (triangle-only scenes)."""
    d, keep = scene.desc()
    counts = (C.c_uint64 * 4)()
    msg = C.create_string_buffer(256)
    rc = load_library().pt_walk_selftest(C.byref(d), n_rays, seed, counts, msg, 256)
    return rc == 0, msg.value.decode(), tuple(int(c) for c in counts)


def bvh_selftest(scene: Scene, leaf_max: int = 4):
    """(ok, message, stats) of the host-only BVH build + validation (no GPU needed)."""
    d, keep = scene.desc()
    st = PtStats()
    msg = C.create_string_buffer(256)
    rc = load_library().pt_bvh_selftest(C.byref(d), leaf_max, C.byref(st), msg, 256)
    return rc == 0, msg.value.decode(), st.as_dict()


def write_ppm(path, rgb: np.ndarray) -> None:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    rc = load_library().pt_write_ppm(str(path).encode(), rgb.ctypes.data, rgb.shape[1], rgb.shape[0])
    if rc != 0:
        raise PtError(rc, f"pt_write_ppm({path})")
