"""Known answers dumped from the reference's own headers (oracle/ref_kat.cpp -> tests/golden/ref_kats.json)
replayed through the CPU restatement (oracle/pt_oracle.c).  Bit-exact unless a comment says why not."""
import ctypes as C
import json
import math

import numpy as np
import pytest

import _oracle
from _oracle import Hit, Rng, Vec3
from conftest import GOLD

K = json.loads((GOLD / "ref_kats.json").read_text())


def f32(x):
    if isinstance(x, str):
        return np.float32({"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[x])
    return np.float32(x)


def same(a, b):
    a, b = f32(a), f32(b)
    return (np.isnan(a) and np.isnan(b)) or a == b


def v3(v):
    return Vec3(*[float(f32(c)) for c in v])


def assert_v3(got, want, what=""):
    assert all(same(g, w) for g, w in zip(got.t(), want)), f"{what}: {got.t()} != {want}"


def test_survey_starter_kats(oracle):
    # SURVEY §8c "Starter KATs [probe, host-compiled reference headers]"
    s = Rng()
    oracle.lib.pto_rng_init(C.byref(s), 1984)
    assert (s.d, list(s.v)) == (237688853, [999278866, 564768280, 4171507460, 3705206908, 881605398])
    assert [oracle.lib.pto_rng_next(C.byref(s)) for _ in range(4)] == [841754470, 1949948301, 1541868453, 3110210077]
    oracle.lib.pto_rng_init(C.byref(s), 1984)
    got = [oracle.lib.pto_uniform(C.byref(s)) for _ in range(4)]
    assert [float(f32(x)) for x in got] == [float(f32(x)) for x in (0.195986241, 0.454007715, 0.358994216, 0.724152207)]
    oracle.lib.pto_rng_init(C.byref(s), 1984)
    assert_v3(oracle.lib.pto_random_cosine_direction(C.byref(s)), (0.448618054, 1.27073705, 0.73891288))
    cam = oracle.ptb.make_camera()
    o, d = Vec3(), Vec3()
    oracle.lib.pto_camera_ray(C.byref(cam), 0.25, 0.75, C.byref(o), C.byref(d))
    assert_v3(o, (0, 0, 0.5))
    assert_v3(d, (-0.207106784, 0.207106799, -1))


@pytest.mark.parametrize("case", K["rng"], ids=lambda c: str(c["seed"]))
def test_rng(oracle, case):
    s = Rng()
    oracle.lib.pto_rng_init(C.byref(s), case["seed"])
    assert s.d == case["d"] and list(s.v) == case["v"]
    a = Rng.from_buffer_copy(s)
    assert [oracle.lib.pto_rng_next(C.byref(a)) for _ in range(8)] == case["raw"]
    a = Rng.from_buffer_copy(s)
    for want in case["uniform"]:
        assert same(oracle.lib.pto_uniform(C.byref(a)), want)


def test_random_cosine_direction(oracle):
    for case in K["random_cosine_direction"]:
        s = Rng()
        oracle.lib.pto_rng_init(C.byref(s), case["seed"])
        for want in case["out"]:
            assert_v3(oracle.lib.pto_random_cosine_direction(C.byref(s)), want)


def test_random_in_unit_sphere(oracle):
    # helper_math.h:1505-1507 builds float3(U(),U(),U()): the argument evaluation order is the compiler's.
    # g++ (the dump) evaluates right-to-left, so the dump's (x,y,z) is our (z,y,x); the port fixes x,y,z
    # left-to-right (builder-defined: the function is dead code in the reference).
    for case in K["random_in_unit_sphere"]:
        s = Rng()
        oracle.lib.pto_rng_init(C.byref(s), case["seed"])
        for want in case["out"]:
            got = oracle.lib.pto_random_in_unit_sphere(C.byref(s))
            assert_v3(Vec3(got.z, got.y, got.x), want)


def test_onb(oracle):
    for case in K["onb"]:
        axis = (Vec3 * 3)()
        oracle.lib.pto_onb(v3(case["n"]), C.byref(axis))
        assert_v3(axis[0], case["u"], "u")
        assert_v3(axis[1], case["v"], "v")
        assert_v3(axis[2], case["w"], "w")


def test_camera(oracle):
    for case in K["camera"]:
        cam = oracle.ptb.make_camera(tuple(case["look_from"]), tuple(case["front"]), case["vfov"], case["hfov"])
        o, d = Vec3(), Vec3()
        oracle.lib.pto_camera_ray(C.byref(cam), float(f32(case["u"])), float(f32(case["v"])), C.byref(o), C.byref(d))
        assert_v3(o, case["o"])
        assert_v3(d, case["d"])


def test_triangle(oracle):
    n_hit = 0
    for case in K["triangle"]:
        pos = (C.c_float * 9)(*case["pos"])
        uv = (C.c_float * 6)(*case["uv"])
        rec = Hit()
        h = oracle.lib.pto_triangle_hit(C.byref(pos), C.byref(uv), v3(case["o"]), v3(case["d"]), 0.001, float(np.finfo(np.float32).max), C.byref(rec))
        assert h == case["hit"]
        if h:
            n_hit += 1
            assert same(rec.t, case["t"])
            assert_v3(rec.p, case["p"])
            assert_v3(rec.normal, case["normal"])
            assert same(rec.u, case["tex"][0]) and same(rec.v, case["tex"][1])
        assert same(oracle.lib.pto_triangle_area(C.byref(pos)), case["area"])
        assert same(oracle.lib.pto_triangle_pdf_value(C.byref(pos), v3(case["o"]), v3(case["d"])), case["pdf_value"])
        s = Rng()
        oracle.lib.pto_rng_init(C.byref(s), case["seed"])
        assert_v3(oracle.lib.pto_triangle_random(C.byref(pos), v3(case["o"]), C.byref(s)), case["random"])
    assert n_hit >= 20


def test_sphere(oracle):
    n_hit = 0
    for case in K["sphere"]:
        sph = (C.c_float * 4)(*case["sph"])
        rec = Hit()
        h = oracle.lib.pto_sphere_hit(C.byref(sph), v3(case["o"]), v3(case["d"]), 0.001, float(np.finfo(np.float32).max), C.byref(rec))
        assert h == case["hit"]
        if h:
            n_hit += 1
            assert same(rec.t, case["t"])
            assert_v3(rec.p, case["p"])
            assert_v3(rec.normal, case["normal"])
    assert n_hit >= 8


def test_texture(oracle):
    t = K["texture"]
    data = np.array(t["data"], np.float32)
    tex = oracle.ptb.PtTexture(t["width"], t["height"], data.ctypes.data_as(C.POINTER(C.c_float)))
    for case in t["cases"]:
        assert_v3(oracle.lib.pto_texture_value(C.byref(tex), float(f32(case["u"])), float(f32(case["v"]))), case["value"])


def test_pdfs(oracle):
    for case in K["pdf"]:
        assert same(oracle.lib.pto_cosine_pdf_value(v3(case["normal"]), v3(case["dir"])), case["cosine_pdf"])
        assert same(oracle.lib.pto_scattering_pdf(v3(case["normal"]), v3(case["dir"])), case["scattering_pdf"])


def test_light_list_and_mixture(oracle):
    # two-triangle emitter of cornell_duck: hitable_list::random / pdf_value and mixture_pdf (pdf.h:57-75)
    # are exercised end to end by test_oracle_parity; here the light pdf alone is pinned.
    m = K["mixture"]
    lights = [(C.c_float * 9)(*p) for p in m["lights"]]
    for case in m["cases"]:
        lv = np.float32(0.0)
        for p in lights:
            lv = np.float32(lv + np.float32(np.float32(0.5) * f32(oracle.lib.pto_triangle_pdf_value(C.byref(p), v3(case["p"]), v3(case["dir"])))))
        assert same(lv, case["light_value"])
        cv = f32(oracle.lib.pto_cosine_pdf_value(v3(case["normal"]), v3(case["dir"])))
        assert same(np.float32(np.float32(0.5) * lv + np.float32(0.5) * cv), case["value"])


def test_rtow_pieces(oracle):
    for case in K["rtow"]:
        out = Vec3()
        ok = oracle.lib.pto_refract(v3(case["d_in"]), v3(case["normal"]), float(np.float32(1.0) / f32(case["ior"])), C.byref(out))
        assert ok == case["refract_ok"]
        if ok:
            assert_v3(out, case["refracted"])
        # schlick: the reference's pow(1-c,5) is libm's; ours is x*x*x*x*x (builder-defined) — equal to a few ulp
        nd = np.array(case["d_in"], np.float32)
        n = np.array(case["normal"], np.float32)
        want = float(f32(case["schlick"]))
        c = abs(float(np.dot(nd / np.linalg.norm(nd), n)))
        got = oracle.lib.pto_schlick(c, float(f32(case["ior"])))
        assert got == pytest.approx(want, rel=2e-5, abs=1e-7)


def test_quantiser_and_i420(oracle):
    for case in K["quantise"]:
        col = (C.c_float * 3)(*case["col"])
        rgb = (C.c_uint8 * 3)()
        oracle.lib.pto_quantise(C.byref(col), case["spp"], C.byref(rgb))
        if max(case["col"]) / case["spp"] > 8e6:
            # float->int overflow: x86 (the dump) yields INT_MIN -> byte 0, CUDA saturates -> 255.  The port follows CUDA.
            assert list(rgb)[0] == 255 and list(rgb)[1:] == case["rgb"][1:]
            continue
        assert list(rgb) == case["rgb"]
        y, u, v = C.c_uint8(), C.c_uint8(), C.c_uint8()
        oracle.lib.pto_yuv(C.byref(rgb), C.byref(y), C.byref(u), C.byref(v))
        assert [y.value, u.value, v.value] == case["yuv"]
