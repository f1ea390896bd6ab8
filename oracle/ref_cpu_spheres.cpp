// ref_cpu_spheres.cpp — the reference's dead classes (sphere, lambertian, metal, dielectric, diffuse_light) compiled as host
// C++ and driven over a whole image: the CPU twin of oracle/ref_gpu_spheres.cu, so that the restatement of those rows (oracle/pt_oracle.c)
// can be pinned bit for bit here, without a GPU.
//
// TEST INFRASTRUCTURE (oracle/).  Built by oracle/Makefile into oracle/_ref/ref_cpu_spheres; executed by tools/ and oracle/make_golden_*.py only.
//
// Reference code used where it lies (-iquote /root/reference/src): `sphere::hit` (src/sphere.h:21-50), `lambertian` / `metal` / `dielectric` /
// `diffuse_light` with their virtual `scatter` / `emitted` (src/material.h:110-217), `camera` (src/camera.h:21-36,95-97), the cuRAND XORWOW
// generator of the CUDA headers compiled for the host.  Ours (no counterpart in the reference, whose `ray_color` only knows UniversalMaterial):
// the pixel loop of `render` (src/DevicePathTracer.h:73-120), the every-sphere loop in the pattern of hitable_list::hit
// (src/hitable_list.h:38-52) and the glue — a scattered ray multiplies the throughput by `attenuation`, a hit that does not scatter ends
// the path with throughput * emitted(), a miss or an exhausted depth contributes (0,0,0) * throughput / (0,0,0) — the same as in
// oracle/ref_gpu_spheres.cu, oracle/pt_oracle.c and the CUDA core.
#include <float.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "ray.h"
#include "helper_math.h"
#include "hitable_list.h"
#include "sphere.h"
#include "hitable.h"
#include "camera.h"
#include "material.h"
#include "triangle.h"
#include "bvh.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptscene_io.h"

static bool world_hit(const std::vector<hitable *> &list, const ray &r, interval ray_t, hit_record &rec) {
    hit_record temp_rec;
    bool hit_anything = false;
    for (hitable *h : list) {
        if (h->hit(r, ray_t, temp_rec)) {
            hit_anything = true;
            ray_t.max = temp_rec.t;
            rec = temp_rec;
        }
    }
    return hit_anything;
}

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: ref_cpu_spheres <scene.ptscene> <W> <H> <spp> <depth> <out.ppm> [--cam lx ly lz fx fy fz vfov hfov]\n");
        return 2;
    }
    int W = atoi(argv[2]), H = atoi(argv[3]), spp = atoi(argv[4]), depth = atoi(argv[5]);
    float cam[8] = {0, 0, 0.5f, 0, 0, -0.5f, 45.f, 45.f};
    int trace_x = -1, trace_y = -1;  // --trace-pixel x y (bottom-up y): every bounce of that pixel on stderr
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--cam") && i + 8 < argc) { for (int k = 0; k < 8; k++) cam[k] = (float)atof(argv[i + 1 + k]); i += 8; }
        else if (!strcmp(argv[i], "--trace-pixel") && i + 2 < argc) { trace_x = atoi(argv[i + 1]); trace_y = atoi(argv[i + 2]); i += 2; }
    }
    pts_scene ps;
    if (pts_load(argv[1], &ps) != 0 || ps.n_spheres == 0) { fprintf(stderr, "cannot load %s (or it has no spheres)\n", argv[1]); return 1; }
    std::vector<material *> mats;
    for (uint32_t i = 0; i < ps.n_mats; i++) {
        const pts_mat &m = ps.mats[i];
        switch (m.type) {  // enum material_type, src/HostScene.h:20-26
            case 0: mats.push_back(new lambertian(make_float3(m.base[0], m.base[1], m.base[2]))); break;
            case 1: mats.push_back(new metal(make_float3(m.base[0], m.base[1], m.base[2]), m.fuzz)); break;
            case 2: mats.push_back(new dielectric(m.ior)); break;
            default: mats.push_back(new diffuse_light(make_float3(m.emis[0], m.emis[1], m.emis[2]))); break;
        }
    }
    std::vector<hitable *> list;
    for (uint32_t i = 0; i < ps.n_spheres; i++)  // sphere keeps its material as UniversalMaterial* (src/sphere.h:19): carried through that field
        list.push_back(new sphere(make_float3(ps.spheres[i].c[0], ps.spheres[i].c[1], ps.spheres[i].c[2]), ps.spheres[i].r, (UniversalMaterial *)(void *)mats[(size_t)ps.spheres[i].mat]));
    camera cam_obj;
    CameraConfig cfg(make_float3(cam[0], cam[1], cam[2]), make_float3(cam[3], cam[4], cam[5]), cam[6], cam[7]);
    cam_obj.recalculate_camera_params(cfg);
    std::vector<uint8_t> fb((size_t)W * H * 3, 0);
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            int pixel_index = (H - j - 1) * W + i;
            curandState local_rand_state;
            curand_init(1984 + pixel_index, 0, 0, &local_rand_state);
            float3 col = make_float3(0, 0, 0);
            for (int s = 0; s < spp; s++) {
                float u = float(i + curand_uniform(&local_rand_state)) / float(W);
                float v = float(j + curand_uniform(&local_rand_state)) / float(H);
                ray cur = cam_obj.get_ray(u, v);
                float3 att = make_float3(1.0f, 1.0f, 1.0f), out = make_float3(0, 0, 0);
                for (int d = 0; d < depth; d++) {
                    hit_record rec;
                    const bool any = world_hit(list, cur, interval(0.001f, FLT_MAX), rec);
                    if (i == trace_x && j == trace_y)
                        fprintf(stderr, "TRACE s %d d %d hit %d t %.9g o %.9g %.9g %.9g dir %.9g %.9g %.9g att %.9g %.9g %.9g\n", s, d, (int)any, any ? rec.t : 0.f, cur.origin().x, cur.origin().y,
                                cur.origin().z, cur.direction().x, cur.direction().y, cur.direction().z, att.x, att.y, att.z);
                    if (!any) { out = make_float3(0, 0, 0) * att; break; }
                    const material *m = (const material *)(const void *)rec.mat_ptr;
                    float3 attenuation;
                    ray scattered;
                    if (m->scatter(cur, rec, attenuation, scattered, &local_rand_state)) {
                        att = att * attenuation;
                        cur = scattered;
                    } else {
                        out = att * m->emitted();
                        break;
                    }
                }
                col += out;
            }
            int3 color = make_int3(255.99f * col / float(spp));
            fb[3 * (size_t)pixel_index] = (uint8_t)min(255, color.x);
            fb[3 * (size_t)pixel_index + 1] = (uint8_t)min(255, color.y);
            fb[3 * (size_t)pixel_index + 2] = (uint8_t)min(255, color.z);
        }
    if (pts_write_ppm(argv[6], fb.data(), W, H) != 0) { fprintf(stderr, "cannot write %s\n", argv[6]); return 1; }
    printf("{\"impl\": \"ref_cpu_spheres\", \"spheres\": %u, \"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %d}\n", ps.n_spheres, W, H, spp, depth);
    return 0;
}
