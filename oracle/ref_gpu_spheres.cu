// ref_gpu_spheres.cu — the "reference CUDA" number for BASELINE config 3 (sphere field, lambertian / metal / dielectric).
//
// TEST INFRASTRUCTURE (oracle/).  Built by oracle/Makefile into oracle/_ref/ref_gpu_spheres (git-ignored, shipped to the GPU box).
// Executed by tools/ and tests/ only — never by the product.
//
// The reference cannot render this configuration: `sphere`, `lambertian`, `metal`, `dielectric`, `diffuse_light` and
// `hitable_list::hit` are dead code there (src/sphere.h:8, src/material.h:110-217) and `ray_color` only
// knows UniversalMaterial.  What a reference build of config 3 WOULD execute is assembled here from those very classes, unmodified,
// compiled where they lie (-iquote /root/reference/src): one thread per pixel (8x8 blocks, as `render`, src/DevicePathTracer.h:73-120),
// one curandState per pixel seeded 1984 + pixel_index (:54), the reference's `camera` (src/camera.h:21-36,95-97), every ray tested
// against every sphere through the virtual `hitable::hit` in the pattern of hitable_list::hit (src/hitable_list.h:38-52), material
// response through the virtual `material::scatter` / `emitted` (src/material.h:16-38).  Only the glue between them is ours (it has
// no counterpart in the reference) and is the same as in the core and in oracle/pt_oracle.c: a scattered ray
// multiplies the throughput by `attenuation`; a hit that does not scatter ends the path with throughput * emitted(); a miss or an
// exhausted depth contributes (0,0,0) (src/camera.h:82,109); quantiser as :98-101.
//
// sphere keeps its material as `UniversalMaterial *` (src/sphere.h:19) although the RTOW materials derive from `material`: the pointer
// is carried through that field and cast back, which is what the dead code would have needed too.
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
#include "cuda_utils.h"
#include "bvh.h"

#include <chrono>
#include <vector>

#include "ptscene_io.h"

HostScene SceneLoader::load(std::string &) { throw std::runtime_error("not available in the oracle build"); }

struct DevMat { int type; float3 base, emis; float fuzz, ior; };

__global__ void build_world(hitable **list, material **mats, const float4 *sph, const int *sph_mat, int n, const DevMat *dm, int nm) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < nm; i++) {
        switch (dm[i].type) {  // enum material_type, src/HostScene.h:20-26
            case 0: mats[i] = new lambertian(dm[i].base); break;
            case 1: mats[i] = new metal(dm[i].base, dm[i].fuzz); break;
            case 2: mats[i] = new dielectric(dm[i].ior); break;
            default: mats[i] = new diffuse_light(dm[i].emis); break;
        }
    }
    for (int i = 0; i < n; i++) list[i] = new sphere(make_float3(sph[i].x, sph[i].y, sph[i].z), sph[i].w, (UniversalMaterial *)(void *)mats[sph_mat[i]]);
}

__device__ bool world_hit(hitable **list, int n, const ray &r, interval ray_t, hit_record &rec) {  // src/hitable_list.h:38-52 over hitable*
    hit_record temp_rec;
    bool hit_anything = false;
    for (int i = 0; i < n; i++) {
        if (list[i]->hit(r, ray_t, temp_rec)) {
            hit_anything = true;
            ray_t.max = temp_rec.t;
            rec = temp_rec;
        }
    }
    return hit_anything;
}

__global__ void render_spheres(uint8_t *fb, int nx, int ny, int spp, int depth, hitable **list, int n, CameraConfig cfg) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    int j = threadIdx.y + blockIdx.y * blockDim.y;
    if (i >= nx || j >= ny) return;
    int pixel_index = (ny - j - 1) * nx + i;
    curandState local_rand_state;
    curand_init(1984 + pixel_index, 0, 0, &local_rand_state);
    camera cam;
    cam.recalculate_camera_params(cfg);
    float3 col = make_float3(0, 0, 0);
    for (int s = 0; s < spp; s++) {
        float u = float(i + curand_uniform(&local_rand_state)) / float(nx);
        float v = float(j + curand_uniform(&local_rand_state)) / float(ny);
        ray cur = cam.get_ray(u, v);
        float3 att = make_float3(1.0f, 1.0f, 1.0f), out = make_float3(0, 0, 0);
        for (int d = 0; d < depth; d++) {
            hit_record rec;
            if (!world_hit(list, n, cur, interval(0.001f, FLT_MAX), rec)) { out = make_float3(0, 0, 0) * att; break; }
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
    fb[3 * pixel_index] = (uint8_t)min(255, color.x);
    fb[3 * pixel_index + 1] = (uint8_t)min(255, color.y);
    fb[3 * pixel_index + 2] = (uint8_t)min(255, color.z);
}

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: ref_gpu_spheres <scene.ptscene> <W> <H> <spp> <depth> <out.ppm|-> [--cam lx ly lz fx fy fz vfov hfov] [--frames n]\n");
        return 2;
    }
    int W = atoi(argv[2]), H = atoi(argv[3]), spp = atoi(argv[4]), depth = atoi(argv[5]);
    const char *out_path = argv[6];
    float cam[8] = {0, 0, 0.5f, 0, 0, -0.5f, 45.f, 45.f};
    int frames = 2;
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--cam") && i + 8 < argc) { for (int k = 0; k < 8; k++) cam[k] = (float)atof(argv[i + 1 + k]); i += 8; }
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    }
    if (frames < 1) frames = 1;
    pts_scene ps;
    if (pts_load(argv[1], &ps) != 0 || ps.n_spheres == 0) { fprintf(stderr, "cannot load %s (or it has no spheres)\n", argv[1]); return 1; }
    std::vector<float4> sph(ps.n_spheres);
    std::vector<int> sm(ps.n_spheres);
    for (uint32_t i = 0; i < ps.n_spheres; i++) { sph[i] = make_float4(ps.spheres[i].c[0], ps.spheres[i].c[1], ps.spheres[i].c[2], ps.spheres[i].r); sm[i] = ps.spheres[i].mat; }
    std::vector<DevMat> dm(ps.n_mats);
    for (uint32_t i = 0; i < ps.n_mats; i++) {
        const pts_mat &m = ps.mats[i];
        dm[i] = {m.type, make_float3(m.base[0], m.base[1], m.base[2]), make_float3(m.emis[0], m.emis[1], m.emis[2]), m.fuzz, m.ior};
    }
    checkCudaErrors(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
    float4 *d_sph; int *d_sm; DevMat *d_dm; hitable **d_list; material **d_mats; uint8_t *d_fb;
    checkCudaErrors(cudaMalloc(&d_sph, sizeof(float4) * sph.size()));
    checkCudaErrors(cudaMalloc(&d_sm, sizeof(int) * sm.size()));
    checkCudaErrors(cudaMalloc(&d_dm, sizeof(DevMat) * dm.size()));
    checkCudaErrors(cudaMalloc(&d_list, sizeof(hitable *) * sph.size()));
    checkCudaErrors(cudaMalloc(&d_mats, sizeof(material *) * dm.size()));
    checkCudaErrors(cudaMalloc(&d_fb, (size_t)W * H * 3));
    checkCudaErrors(cudaMemcpy(d_sph, sph.data(), sizeof(float4) * sph.size(), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_sm, sm.data(), sizeof(int) * sm.size(), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_dm, dm.data(), sizeof(DevMat) * dm.size(), cudaMemcpyHostToDevice));
    build_world<<<1, 1>>>(d_list, d_mats, d_sph, d_sm, (int)sph.size(), d_dm, (int)dm.size());
    checkCudaErrors(cudaDeviceSynchronize());
    CameraConfig cfg(make_float3(cam[0], cam[1], cam[2]), make_float3(cam[3], cam[4], cam[5]), cam[6], cam[7]);
    dim3 block(8, 8), grid(W / 8 + 1, H / 8 + 1);
    std::vector<double> secs;
    for (int f = 0; f < frames; f++) {
        auto t0 = std::chrono::high_resolution_clock::now();
        render_spheres<<<grid, block>>>(d_fb, W, H, spp, depth, d_list, (int)sph.size(), cfg);
        checkCudaErrors(cudaGetLastError());
        checkCudaErrors(cudaDeviceSynchronize());
        secs.push_back(std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count());
    }
    std::vector<uint8_t> fb((size_t)W * H * 3);
    checkCudaErrors(cudaMemcpy(fb.data(), d_fb, fb.size(), cudaMemcpyDeviceToHost));
    if (strcmp(out_path, "-") != 0) pts_write_ppm(out_path, fb.data(), W, H);
    double best = 1e30;
    for (double s : secs) best = std::min(best, s);
    double samples = (double)W * H * spp;
    printf("REF_GPU_JSON {\"impl\": \"ref_gpu_spheres\", \"seconds\": %.6f, \"samples\": %.0f, \"msamples_per_s\": %.6f, \"spheres\": %u, \"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %d, \"frames\": %d}\n",
           best, samples, samples / best / 1e6, ps.n_spheres, W, H, spp, depth, frames);
    return 0;
}
