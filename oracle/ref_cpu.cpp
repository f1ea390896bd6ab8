// ref_cpu.cpp — "Oracle C": the reference's own __device__ headers compiled as host C++.
//
// TEST INFRASTRUCTURE (oracle/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may execute the binary this builds
// (oracle/_ref/ref_cpu).  The product never links or calls it.
//
// What is the reference's and what is ours:
//   * every #include "…" below resolves (via -iquote) to the UNMODIFIED files under
//     /root/reference/src — camera.h (ray_color), bvh.h, triangle.h, material.h,
//     pdf.h, onb.h, hitable_list.h, Texture.h, helper_math.h — so the integrator,
//     BVH, samplers and RNG draw order are the reference's own code;
//   * the only restated piece is the ~30-line body of the `render` kernel
//     (reference src/DevicePathTracer.h:73-120) turned into an OpenMP pixel loop, and the
//     scene upload of DevicePathTracer::loadTextures/loadMaterials/
//     loadTrianglesWithTextures/reloadWorld (src/DevicePathTracer.h:241-340) done with
//     std::vector instead of thrust::device_vector.
// Built by oracle/Makefile with g++ -fopenmp; nothing from /root/reference is copied.
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

#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ptscene_io.h"

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr,
                "usage: ref_cpu <scene.ptscene> <W> <H> <spp> <depth> <out.ppm|-> [--cam lx ly lz fx fy fz vfov hfov]\n"
                "               [--yuv out.yuv] [--threads n] [--rect ox oy w h] [--row-stride k]\n");
        return 2;
    }
    const char *scene_path = argv[1];
    int W = atoi(argv[2]), H = atoi(argv[3]), spp = atoi(argv[4]);
    unsigned depth = (unsigned)atoi(argv[5]);
    const char *out_path = argv[6];
    // defaults of reference src/main.cu:40
    float cam[8] = {0, 0, 0.5f, 0, 0, -0.5f, 45.f, 45.f};
    const char *yuv_path = nullptr;
    int threads = omp_get_max_threads();
    int rect[4] = {0, 0, W, H};
    int trace_x = -1, trace_y = -1;  // --trace-pixel x y: print what ray_color returns for every sample of that pixel (debugging the restatement)
    int row_stride = 1;  // > 1: only every k-th row of the rect is rendered (a uniform sample of the frame for timing)
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--cam") && i + 8 < argc) { for (int k = 0; k < 8; k++) cam[k] = (float)atof(argv[i + 1 + k]); i += 8; }
        else if (!strcmp(argv[i], "--yuv") && i + 1 < argc) yuv_path = argv[++i];
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--rect") && i + 4 < argc) { for (int k = 0; k < 4; k++) rect[k] = atoi(argv[i + 1 + k]); i += 4; }
        else if (!strcmp(argv[i], "--trace-pixel") && i + 2 < argc) { trace_x = atoi(argv[i + 1]); trace_y = atoi(argv[i + 2]); i += 2; }
        else if (!strcmp(argv[i], "--row-stride") && i + 1 < argc) row_stride = atoi(argv[++i]) > 0 ? atoi(argv[i]) : 1;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }

    pts_scene ps;
    if (pts_load(scene_path, &ps) != 0) { fprintf(stderr, "cannot load %s\n", scene_path); return 1; }

    // --- scene "upload" (src/DevicePathTracer.h:241-340) -------------------------------
    std::vector<BaseColorTexture> textures;
    textures.reserve(ps.n_tex);
    for (uint32_t i = 0; i < ps.n_tex; i++) textures.emplace_back(ps.tex[i].w, ps.tex[i].h, (float3 *)ps.tex[i].rgb);

    std::vector<UniversalMaterial> materials;
    materials.reserve(ps.n_mats);
    BaseColorTexture *baseTex = nullptr, *emisTex = nullptr;  // sticky across materials, as in :269-279
    for (uint32_t i = 0; i < ps.n_mats; i++) {
        const pts_mat &m = ps.mats[i];
        if (m.base_tex >= 0) baseTex = &textures[(size_t)m.base_tex];
        if (m.emis_tex >= 0) emisTex = &textures[(size_t)m.emis_tex];
        materials.emplace_back(make_float3(m.base[0], m.base[1], m.base[2]), baseTex, make_float3(m.emis[0], m.emis[1], m.emis[2]), emisTex);
    }

    std::vector<triangle> faces;
    faces.reserve(ps.n_tris);  // no reallocation => light pointers stay valid (the reference's dangle, SURVEY §0.9b, is not reproduced)
    std::vector<triangle *> light_faces;
    for (uint32_t i = 0; i < ps.n_tris; i++) {
        const pts_tri &t = ps.tris[i];
        Vertex v[3];
        for (int k = 0; k < 3; k++) {
            v[k].position = make_float3(t.pos[3 * k], t.pos[3 * k + 1], t.pos[3 * k + 2]);
            v[k].texCoords = make_float2(t.uv[2 * k], t.uv[2 * k + 1]);
        }
        faces.emplace_back(v[0], v[1], v[2], &materials[(size_t)t.mat]);
        const pts_mat &m = ps.mats[t.mat];
        if (m.emis[0] > 0.0001 || m.emis[1] > 0.0001 || m.emis[2] > 0.0001) light_faces.push_back(&faces.back());
    }
    if (light_faces.empty()) {
        fprintf(stderr, "scene has no emissive triangle: the reference dereferences an empty light list (SURVEY §0.3)\n");
        return 3;
    }

    auto t_build0 = std::chrono::high_resolution_clock::now();
    BVH *world = new BVH(faces.data(), (int)faces.size());
    auto t_build1 = std::chrono::high_resolution_clock::now();
    hitable_list *lights = new hitable_list(light_faces.data(), (int)light_faces.size());
    camera *cam_obj = new camera();
    CameraConfig cfg(make_float3(cam[0], cam[1], cam[2]), make_float3(cam[3], cam[4], cam[5]), cam[6], cam[7]);
    {   // frame >= 2 semantics: the camera object holds valid parameters before the first get_ray
        cam_obj->recalculate_camera_params(cfg);
    }

    std::vector<uint8_t> fb_rgb((size_t)W * (size_t)H * 3, 0), fb_yuv((size_t)W * (size_t)H * 3 / 2 + (size_t)W + (size_t)H + 64, 0);  // slack: odd sizes overrun the reference's own layout
    Resolution res{(unsigned)W, (unsigned)H};

    omp_set_num_threads(threads);
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < rect[3]; j++) {
        if (j % row_stride) continue;
        for (int i = 0; i < rect[2]; i++) {
            // --- body of `render`, src/DevicePathTracer.h:76-119 ---
            int x = rect[0] + i;
            int y = rect[1] + j;
            int pixel_index = ((int)res.height - y - 1) * (int)res.width + x;
            curandState local_rand_state;
            curand_init(1984 + pixel_index, 0, 0, &local_rand_state);  // render_init, :46-55
            float3 col = make_float3(0, 0, 0);
            CameraConfig cc = cfg;  // passed by value at launch (:210)
            for (int s = 0; s < spp; s++) {
                float u = float(x + curand_uniform(&local_rand_state)) / float(res.width);
                float v = float(y + curand_uniform(&local_rand_state)) / float(res.height);
                ray r = cam_obj->get_ray(u, v);
                const float3 one = cam_obj->ray_color(r, &world, cc, depth, &lights, &local_rand_state);
                if (x == trace_x && y == trace_y) fprintf(stderr, "TRACE sample %d colour %.9g %.9g %.9g\n", s, one.x, one.y, one.z);
                col += one;
            }
            float3 color_modifier = make_float3(1, 1, 1);
            int3 color = make_int3(255.99 * col / float(spp) * color_modifier);
            color.x = min(255, color.x);
            color.y = min(255, color.y);
            color.z = min(255, color.z);
            fb_rgb[3 * (size_t)pixel_index] = (uint8_t)color.x;
            fb_rgb[3 * (size_t)pixel_index + 1] = (uint8_t)color.y;
            fb_rgb[3 * (size_t)pixel_index + 2] = (uint8_t)color.z;
            fb_yuv[(size_t)pixel_index] = (uint8_t)(((66 * color.x + 129 * color.y + 25 * color.z + 128) >> 8) + 16);
            int blockRow = pixel_index / (int)res.width;
            int blockCol = pixel_index % (int)res.width;
            if (blockRow % 2 == 0 && blockCol % 2 == 0) {
                int totalPixels = (int)(res.width * res.height);
                int uvSize = totalPixels / 4;
                int uOffset = totalPixels;
                int vOffset = totalPixels + uvSize;
                int uvIndex = (blockRow / 2) * ((int)res.width / 2) + (blockCol / 2);
                fb_yuv[(size_t)(uOffset + uvIndex)] = (uint8_t)(((-38 * color.x - 74 * color.y + 112 * color.z + 128) >> 8) + 128);
                fb_yuv[(size_t)(vOffset + uvIndex)] = (uint8_t)(((112 * color.x - 94 * color.y - 18 * color.z + 128) >> 8) + 128);
            }
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    double sec = std::chrono::duration<double>(t1 - t0).count();
    double bsec = std::chrono::duration<double>(t_build1 - t_build0).count();

    if (strcmp(out_path, "-") != 0 && pts_write_ppm(out_path, fb_rgb.data(), W, H) != 0) { fprintf(stderr, "cannot write %s\n", out_path); return 1; }
    if (yuv_path) {
        FILE *f = fopen(yuv_path, "wb");
        if (f) { fwrite(fb_yuv.data(), 1, (size_t)W * (size_t)H * 3 / 2, f); fclose(f); }
    }
    double samples = (double)rect[2] * (double)((rect[3] + row_stride - 1) / row_stride) * (double)spp;
    printf("{\"impl\": \"ref_cpu\", \"seconds\": %.6f, \"bvh_build_seconds\": %.6f, \"samples\": %.0f, \"msamples_per_s\": %.6f, \"threads\": %d, "
           "\"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %u, \"rect\": [%d, %d, %d, %d], \"row_stride\": %d}\n",
           sec, bsec, samples, samples / sec / 1e6, threads, W, H, spp, depth, rect[0], rect[1], rect[2], rect[3], row_stride);
    pts_free(&ps);
    return 0;
}
