// ref_gpu.cu — "Oracle G": the reference's own CUDA renderer, UNMODIFIED, driven headless.
//
// TEST INFRASTRUCTURE (oracle/). Built by oracle/Makefile into oracle/_ref/ref_gpu (git-ignored,
// shipped to the GPU box).  Executed only by tests/ (parity), bench.py (the "reference CUDA
// renderer on the same box" figure) — never by the product.
//
// Everything that renders is the reference's: RenderManager.h pulls in StreamThread.h,
// DevicePathTracer.h (all six __global__ kernels), camera.h, bvh.h, triangle.h, material.h, pdf.h …
// straight from /root/reference/src (via -iquote).  This file only
//   * reads a .ptscene into the reference's HostScene struct (assimp is absent, so
//     SceneLoader::load is a throwing stub the render path never calls),
//   * fills RendererConfig / CameraConfig (the reference's CLI cannot set them: SURVEY §0.4),
//   * calls RenderManager::renderFrame() and times frames >= 2 (frame 1 reads the camera
//     object before any thread has written it: SURVEY §0.9a),
//   * dumps getCurrentFrame() as P6.
#include <float.h>
#include "RenderManager.h"

#include <unistd.h>

#include "ptscene_io.h"

HostScene SceneLoader::load(std::string &) { throw std::runtime_error("SceneLoader::load is not available in the oracle build"); }

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr,
                "usage: ref_gpu <scene.ptscene> <W> <H> <spp> <depth> <out.ppm|-> [--cam lx ly lz fx fy fz vfov hfov]\n"
                "               [--gpus n] [--frames n] [--block bx by] [--heap-mb n] [--stack n] [--yuv out.yuv]\n");
        return 2;
    }
    const char *scene_path = argv[1];
    unsigned W = (unsigned)atoi(argv[2]), H = (unsigned)atoi(argv[3]);
    unsigned spp = (unsigned)atoi(argv[4]), depth = (unsigned)atoi(argv[5]);
    const char *out_path = argv[6];
    float cam[8] = {0, 0, 0.5f, 0, 0, -0.5f, 45.f, 45.f};  // src/main.cu:40
    unsigned gpus = 1, frames = 2, bx = 8, by = 8;
    size_t heap_mb = 0, stack = 0;
    const char *yuv_path = nullptr;
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--cam") && i + 8 < argc) { for (int k = 0; k < 8; k++) cam[k] = (float)atof(argv[i + 1 + k]); i += 8; }
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--block") && i + 2 < argc) { bx = (unsigned)atoi(argv[i + 1]); by = (unsigned)atoi(argv[i + 2]); i += 2; }
        else if (!strcmp(argv[i], "--heap-mb") && i + 1 < argc) heap_mb = (size_t)atol(argv[++i]);
        else if (!strcmp(argv[i], "--stack") && i + 1 < argc) stack = (size_t)atol(argv[++i]);
        else if (!strcmp(argv[i], "--yuv") && i + 1 < argc) yuv_path = argv[++i];
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (frames < 2) frames = 2;

    pts_scene ps;
    if (pts_load(scene_path, &ps) != 0) { fprintf(stderr, "cannot load %s\n", scene_path); return 1; }

    HostScene scene;
    for (uint32_t i = 0; i < ps.n_tex; i++) {
        HostTexture t;
        t.width = ps.tex[i].w;
        t.height = ps.tex[i].h;
        t.data.resize((size_t)t.width * (size_t)t.height);
        memcpy(t.data.data(), ps.tex[i].rgb, t.data.size() * sizeof(float3));
        scene.textures.push_back(std::move(t));
    }
    bool any_light = false;
    for (uint32_t i = 0; i < ps.n_mats; i++) {
        const pts_mat &m = ps.mats[i];
        HostMaterial hm;
        hm.baseColor = make_float3(m.base[0], m.base[1], m.base[2]);
        hm.emissiveFactor = make_float3(m.emis[0], m.emis[1], m.emis[2]);
        if (m.base_tex >= 0) hm.baseColorTextureIdx = m.base_tex;
        if (m.emis_tex >= 0) hm.emissiveTextureIdx = m.emis_tex;
        scene.materials.push_back(hm);
    }
    scene.triangles.reserve(ps.n_tris);
    for (uint32_t i = 0; i < ps.n_tris; i++) {
        const pts_tri &t = ps.tris[i];
        Triangle tr;
        Vertex *v[3] = {&tr.v0, &tr.v1, &tr.v2};
        for (int k = 0; k < 3; k++) {
            v[k]->position = make_float3(t.pos[3 * k], t.pos[3 * k + 1], t.pos[3 * k + 2]);
            v[k]->texCoords = make_float2(t.uv[2 * k], t.uv[2 * k + 1]);
        }
        tr.textureIdx = t.tex;
        tr.materialIdx = t.mat;
        scene.triangles.push_back(tr);
        const pts_mat &m = ps.mats[t.mat];
        if (m.emis[0] > 0.0001 || m.emis[1] > 0.0001 || m.emis[2] > 0.0001) any_light = true;
    }
    if (!any_light) { fprintf(stderr, "scene has no emissive triangle: the reference cannot render it (SURVEY §0.3)\n"); return 3; }

    for (unsigned g = 0; g < gpus; g++) {
        cudaSetDevice((int)g);
        if (heap_mb) checkCudaErrors(cudaDeviceSetLimit(cudaLimitMallocHeapSize, heap_mb << 20));
        if (stack) checkCudaErrors(cudaDeviceSetLimit(cudaLimitStackSize, stack));
    }
    cudaSetDevice(0);

    RendererConfig cfg;
    cfg.samplesPerPixel = spp;
    cfg.recursionDepth = depth;
    cfg.gpuNumber = gpus;
    cfg.streamsPerGpu = 1;
    cfg.resolution = {W, H};
    cfg.algorithmType = FSFL;
    cfg.threadBlockSize = dim3(bx, by);
    cfg.showTasks = false;
    CameraConfig cameraConfig(make_float3(cam[0], cam[1], cam[2]), make_float3(cam[3], cam[4], cam[5]), cam[6], cam[7]);
    SceneLoader loader;

    auto t_init0 = std::chrono::high_resolution_clock::now();
    RenderManager manager(cfg, scene, cameraConfig, loader);
    auto t_init1 = std::chrono::high_resolution_clock::now();

    std::vector<double> frame_s;
    for (unsigned f = 0; f < frames; f++) {
        auto t0 = std::chrono::high_resolution_clock::now();
        manager.renderFrame();
        auto t1 = std::chrono::high_resolution_clock::now();
        frame_s.push_back(std::chrono::duration<double>(t1 - t0).count());
    }
    for (unsigned g = 0; g < gpus; g++) { cudaSetDevice((int)g); cudaDeviceSynchronize(); }

    if (strcmp(out_path, "-") != 0) pts_write_ppm(out_path, manager.getCurrentFrame(), (int)W, (int)H);
    if (yuv_path) {
        FILE *f = fopen(yuv_path, "wb");
        if (f) { fwrite(manager.getYUVFrame(), 1, (size_t)W * H * 3 / 2, f); fclose(f); }
    }
    double best = 1e30, sum = 0;
    for (size_t f = 1; f < frame_s.size(); f++) { best = std::min(best, frame_s[f]); sum += frame_s[f]; }
    double mean = sum / (double)(frame_s.size() - 1);
    double samples = (double)W * H * spp;
    fflush(stdout);
    printf("\nREF_GPU_JSON {\"impl\": \"ref_gpu\", \"seconds\": %.6f, \"seconds_best\": %.6f, \"first_frame_seconds\": %.6f, \"init_seconds\": %.6f, "
           "\"samples\": %.0f, \"msamples_per_s\": %.6f, \"gpus\": %u, \"width\": %u, \"height\": %u, \"spp\": %u, \"depth\": %u, \"frames_timed\": %zu}\n",
           mean, best, frame_s[0], std::chrono::duration<double>(t_init1 - t_init0).count(), samples, samples / mean / 1e6, gpus, W, H, spp, depth,
           frame_s.size() - 1);
    fflush(stdout);
    _exit(0);  // skip the reference's racy teardown (detach while workers sit in the barrier)
}
