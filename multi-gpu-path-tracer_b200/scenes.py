"""Synthetic scenes for the BASELINE configs that have no model file.

Everything is generated with numpy's MT19937 (`np.random.RandomState(seed)`), float32 throughout, so the oracle and the
CUDA core read identical floats; a scene can be written with `Scene.save_ptscene` for the C++ hosts.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

from .capi import MAT_DTYPE, PT_MAT_DIELECTRIC, PT_MAT_DIFFUSE_LIGHT, PT_MAT_LAMBERTIAN, PT_MAT_METAL, PT_MAT_UNIVERSAL, Scene


def _mat(type_, base=(0, 0, 0), emis=(0, 0, 0), fuzz=0.0, ior=1.5):
    m = np.zeros(1, MAT_DTYPE)
    m["type"], m["base"], m["emis"], m["base_tex"], m["emis_tex"], m["fuzz"], m["ior"] = type_, base, emis, -1, -1, fuzz, ior
    return m


def rtow_sphere_field(seed: int = 1984, grid: int = 11, sky=(0.7, 0.8, 1.0)) -> Tuple[Scene, dict]:
    """Config 3: "Ray Tracing in One Weekend" final scene — ground sphere r=1000, (2*grid)^2 jittered r=0.2 spheres,
    3 spheres r=1 (grid=11: 488 spheres) with lambertian / metal / dielectric materials (the reference's dead classes,
    src/material.h:110-179, src/sphere.h), lit by an enclosing DIFFUSE_LIGHT sky sphere because the reference's
    background is hard-wired black (src/camera.h:109).  Returns (scene, camera kwargs)."""
    rs = np.random.RandomState(seed)
    sph, mats, sph_mat = [], [], []

    def add(c, r, m):
        sph.append([c[0], c[1], c[2], r])
        mats.append(m)
        sph_mat.append(len(mats) - 1)

    add((0, -1000, 0), 1000, _mat(PT_MAT_LAMBERTIAN, (0.5, 0.5, 0.5)))
    for a in range(-grid, grid):
        for b in range(-grid, grid):
            choose = rs.random_sample()
            c = (a + 0.9 * rs.random_sample(), 0.2, b + 0.9 * rs.random_sample())
            if math.dist(c, (4, 0.2, 0)) <= 0.9:
                rs.random_sample(3)
                continue
            if choose < 0.8:
                u = rs.random_sample(6)
                add(c, 0.2, _mat(PT_MAT_LAMBERTIAN, (u[0] * u[1], u[2] * u[3], u[4] * u[5])))
            elif choose < 0.95:
                u = rs.random_sample(4)
                add(c, 0.2, _mat(PT_MAT_METAL, (0.5 * (1 + u[0]), 0.5 * (1 + u[1]), 0.5 * (1 + u[2])), fuzz=0.5 * u[3]))
            else:
                add(c, 0.2, _mat(PT_MAT_DIELECTRIC, (1, 1, 1), ior=1.5))
    add((0, 1, 0), 1.0, _mat(PT_MAT_DIELECTRIC, (1, 1, 1), ior=1.5))
    add((-4, 1, 0), 1.0, _mat(PT_MAT_LAMBERTIAN, (0.4, 0.2, 0.1)))
    add((4, 1, 0), 1.0, _mat(PT_MAT_METAL, (0.7, 0.6, 0.5), fuzz=0.0))
    add((0, 0, 0), 5000.0, _mat(PT_MAT_DIFFUSE_LIGHT, emis=sky))
    scene = Scene(sph=np.array(sph, np.float32), sph_mat=np.array(sph_mat, np.int32), mats=np.concatenate(mats))
    look_from, look_at, vfov = (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 20.0
    front = tuple(look_at[k] - look_from[k] for k in range(3))
    hfov = math.degrees(2 * math.atan(math.tan(math.radians(vfov) / 2) * 16 / 9))  # hfov is independent of the aspect (camera.h:24-35)
    return scene, dict(look_from=look_from, front=front, vfov=vfov, hfov=hfov)


def displaced_sphere_in_cornell(duck: Scene, n: int = 1000, seed: int = 1984, center=(-8.0, -60.0, -1035.0), radius: float = 150.0) -> Scene:
    """Config 4: a displaced UV sphere of 2*n*n triangles (n=1000: 2.0 M) standing where the duck stands, inside the
    cornell_duck box (its 12 wall/light triangles and materials are kept, the 4212 duck triangles dropped).
    Displacement = radius * 0.12 * sum of 6 seeded sinusoids in (theta, phi)."""
    rs = np.random.RandomState(seed)
    keep = duck.tri_mat != 4
    th = np.linspace(0.0, math.pi, n + 1, dtype=np.float64)[:, None]
    ph = np.linspace(0.0, 2.0 * math.pi, n + 1, dtype=np.float64)[None, :]
    disp = np.zeros((n + 1, n + 1))
    for _ in range(6):
        a, b = rs.randint(2, 40, 2)
        p1, p2 = rs.random_sample(2) * 2 * math.pi
        disp += np.sin(a * th + p1) * np.sin(b * ph + p2) / 6.0
    disp[:, -1] = disp[:, 0]  # seam
    r = radius * (1.0 + 0.12 * disp) * np.ones_like(th * ph)
    x = center[0] + r * np.sin(th) * np.cos(ph)
    y = center[1] + r * np.cos(th) * np.ones_like(ph)
    z = center[2] + r * np.sin(th) * np.sin(ph)
    P = np.stack([x, y, z], -1).astype(np.float32)
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.concatenate([a, b, c], -1).reshape(-1, 9)
    t2 = np.concatenate([a, c, d], -1).reshape(-1, 9)
    tris = np.concatenate([t1, t2])
    # drop the degenerate triangles at the poles (two coincident corners), as the loaders do
    ok = ~((tris[:, 0:3] == tris[:, 3:6]).all(1) | (tris[:, 3:6] == tris[:, 6:9]).all(1) | (tris[:, 0:3] == tris[:, 6:9]).all(1))
    tris = tris[ok]
    mats = duck.mats.copy()
    mats[4]["base"], mats[4]["base_tex"] = (0.73, 0.73, 0.73), -1
    return Scene(tri_pos=np.concatenate([duck.tri_pos[keep], tris]), tri_uv=np.zeros((int(keep.sum()) + len(tris), 6), np.float32),
                 tri_mat=np.concatenate([duck.tri_mat[keep], np.full(len(tris), 4, np.int32)]), mats=mats)


def mixed_material_test_scene(seed: int = 7) -> Tuple[Scene, dict]:
    """Small scene touching every primitive / material kind at once (tests): a UNIVERSAL floor quad and emitter quad
    (importance-sampled light), lambertian / metal / dielectric / diffuse_light spheres, one textured triangle pair."""
    rs = np.random.RandomState(seed)
    tex = (rs.random_sample((8, 8, 3)) * 255).astype(np.uint8).astype(np.float32)
    mats = np.concatenate([
        _mat(PT_MAT_UNIVERSAL, (0.6, 0.6, 0.6)),                       # 0 floor
        _mat(PT_MAT_UNIVERSAL, (0, 0, 0), emis=(1, 0.9, 0.8)),         # 1 area light (x50)
        _mat(PT_MAT_LAMBERTIAN, (0.8, 0.3, 0.3)),                      # 2
        _mat(PT_MAT_METAL, (0.8, 0.8, 0.9), fuzz=0.15),                # 3
        _mat(PT_MAT_DIELECTRIC, (1, 1, 1), ior=1.5),                   # 4
        _mat(PT_MAT_DIFFUSE_LIGHT, emis=(2.0, 2.0, 4.0)),              # 5
        _mat(PT_MAT_UNIVERSAL, (1, 1, 1)),                             # 6 textured
    ])
    mats[6]["base_tex"] = 0

    def quad(p0, p1, p2, p3):
        return [list(p0) + list(p1) + list(p2), list(p0) + list(p2) + list(p3)]
    tri = quad((-6, 0, -6), (6, 0, -6), (6, 0, 6), (-6, 0, 6)) + quad((-1.5, 5, -1.5), (-1.5, 5, 1.5), (1.5, 5, 1.5), (1.5, 5, -1.5)) + \
        quad((-3, 0, -4), (3, 0, -4), (3, 3, -4), (-3, 3, -4))
    uv = [[0] * 6] * 4 + [[0.05, 0.05, 0.95, 0.05, 0.95, 0.95], [0.05, 0.05, 0.95, 0.95, 0.05, 0.95]]
    scene = Scene(tri_pos=np.array(tri, np.float32), tri_uv=np.array(uv, np.float32), tri_mat=np.array([0, 0, 1, 1, 6, 6], np.int32),
                  sph=np.array([[-2.2, 1, 0, 1], [0, 1, 0.5, 1], [2.2, 1, 0, 1], [0, 0.4, 2.5, 0.4]], np.float32), sph_mat=np.array([2, 4, 3, 5], np.int32),
                  mats=mats, textures=[tex])
    cam = dict(look_from=(0.0, 2.5, 9.0), front=(0.0, -0.15, -1.0), vfov=35.0, hfov=55.0)
    return scene, cam
