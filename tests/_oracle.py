"""ctypes binding of oracle/pt_oracle.h — TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_build" / "libpt_oracle.so"


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def t(self):
        return (self.x, self.y, self.z)


class Rng(C.Structure):
    _fields_ = [("d", C.c_uint32), ("v", C.c_uint32 * 5)]


class Hit(C.Structure):
    _fields_ = [("hit", C.c_int), ("t", C.c_float), ("p", Vec3), ("normal", Vec3), ("u", C.c_float), ("v", C.c_float), ("bu", C.c_float), ("bv", C.c_float), ("mat", C.c_int32), ("prim", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "rays", "draws", "emitter_paths", "miss_paths", "depth_paths", "absorbed_paths")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build():
    src = [ROOT / "oracle" / "pt_oracle.c", ROOT / "oracle" / "pt_oracle.h", ROOT / "include" / "ptcore.h"]
    if LIB.exists() and all(LIB.stat().st_mtime >= s.stat().st_mtime for s in src):
        return
    LIB.parent.mkdir(exist_ok=True)
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", str(src[0]), "-lm", "-o", str(LIB)], check=True)


class Oracle:
    def __init__(self, lib):
        import ptb200
        self.lib = lib
        self.ptb = ptb200
        f, vp, u32, i32 = C.c_float, C.c_void_p, C.c_uint32, C.c_int32
        P = C.POINTER
        fa9, fa6, fa4, fa3 = f * 9, f * 6, f * 4, f * 3
        sig = {
            "pto_rng_init": (None, [P(Rng), C.c_uint64]),
            "pto_rng_next": (u32, [P(Rng)]),
            "pto_uniform": (f, [P(Rng)]),
            "pto_random_cosine_direction": (Vec3, [P(Rng)]),
            "pto_random_in_unit_sphere": (Vec3, [P(Rng)]),
            "pto_onb": (None, [Vec3, P(Vec3 * 3)]),
            "pto_camera_ray": (None, [P(ptb200.PtCamera), f, f, P(Vec3), P(Vec3)]),
            "pto_triangle_hit": (C.c_int, [P(fa9), P(fa6), Vec3, Vec3, f, f, P(Hit)]),
            "pto_triangle_area": (f, [P(fa9)]),
            "pto_triangle_pdf_value": (f, [P(fa9), Vec3, Vec3]),
            "pto_triangle_random": (Vec3, [P(fa9), Vec3, P(Rng)]),
            "pto_sphere_hit": (C.c_int, [P(fa4), Vec3, Vec3, f, f, P(Hit)]),
            "pto_texture_value": (Vec3, [P(ptb200.PtTexture), f, f]),
            "pto_cosine_pdf_value": (f, [Vec3, Vec3]),
            "pto_scattering_pdf": (f, [Vec3, Vec3]),
            "pto_refract": (C.c_int, [Vec3, Vec3, f, P(Vec3)]),
            "pto_schlick": (f, [f, f]),
            "pto_quantise": (None, [P(fa3), u32, P(C.c_uint8 * 3)]),
            "pto_yuv": (None, [P(C.c_uint8 * 3), P(C.c_uint8), P(C.c_uint8), P(C.c_uint8)]),
            "pto_world_create": (vp, [P(ptb200.PtSceneDesc)]),
            "pto_world_destroy": (None, [vp]),
            "pto_world_n_lights": (C.c_int, [vp]),
            "pto_world_hit": (C.c_int, [vp, Vec3, Vec3, f, f, P(Hit)]),
            "pto_ray_color": (Vec3, [vp, Vec3, Vec3, u32, P(Rng), P(Stats)]),
            "pto_trace_pixel": (C.c_int, [vp, P(ptb200.PtCamera), u32, u32, u32, u32, i32, i32, vp, i32, vp]),
            "pto_render": (C.c_int, [vp, P(ptb200.PtCamera), u32, u32, u32, u32, i32, i32, i32, i32, vp, vp, vp, P(Stats), C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args

    def world(self, scene):
        d, keep = scene.desc()
        w = self.lib.pto_world_create(C.byref(d))
        return w

    def set_triple_draw_order_zyx(self, on):
        """Test only: the argument-evaluation order of the HOST build of the reference (see pt_oracle.c)."""
        self.lib.pto_set_triple_draw_order_zyx(int(bool(on)))

    def render(self, scene_or_world, width, height, spp, depth, camera=None, rect=None, threads=0, want_accum=False, keyed_chunks=0):
        """Returns (rgb[h,w,3] u8, yuv[w*h*3/2] u8, stats dict[, accum]).  keyed_chunks > 0: the (pixel, sample)-keyed RNG mode."""
        self.lib.pto_set_keyed_chunks(keyed_chunks)
        own = not isinstance(scene_or_world, int)
        w = self.world(scene_or_world) if own else scene_or_world
        cam = self.ptb.make_camera(**{**self.ptb.DEFAULT_CAMERA, **(camera or {})})
        rgb = np.zeros((height, width, 3), np.uint8)
        yuv = np.zeros((width * height * 3 // 2 + 2,), np.uint8)
        acc = np.zeros((height, width, 3), np.float32) if want_accum else None
        st = Stats()
        ox, oy, tw, th = rect or (0, 0, width, height)
        rc = self.lib.pto_render(w, C.byref(cam), width, height, spp, depth, ox, oy, tw, th, rgb.ctypes.data, yuv.ctypes.data,
                                 acc.ctypes.data if want_accum else None, C.byref(st), threads)
        if own:
            self.lib.pto_world_destroy(w)
        assert rc == 0
        out = (rgb, yuv[: width * height * 3 // 2], st.as_dict())
        return out + (acc,) if want_accum else out


    def trace_pixel(self, world, width, height, spp, depth, x, y, camera=None, max_events=4096):
        cam = self.ptb.make_camera(**{**self.ptb.DEFAULT_CAMERA, **(camera or {})})
        ev = np.zeros((max_events, 16), np.float32)
        col = np.zeros(3, np.float32)
        n = self.lib.pto_trace_pixel(world, C.byref(cam), width, height, spp, depth, x, y, ev.ctypes.data, max_events, col.ctypes.data)
        return ev[: min(n, max_events)], col


def load():
    build()
    return Oracle(C.CDLL(str(LIB)))
