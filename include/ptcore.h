/* ptcore.h — C ABI of the B200-native path-tracing core (libptcore.so).
 *
 * This is the drop-in boundary for the reference's device layer.  Each entry point replaces one
 * piece of `class DevicePathTracer` (reference src/DevicePathTracer.h:167-392); the C++ shim
 * multi-gpu-path-tracer_b200/csrc/host/DevicePathTracer.h keeps the reference's class signature
 * and forwards here.  Plain pointers and sizes only: no STL, no torch, no exceptions cross it.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a cudaError_t value or a PT_ERR_* code;
 *     ptcore_last_error(h) returns a static/handle-owned message for the last failure.
 *     (The reference has no return codes: checkCudaErrors prints and exit(99)s,
 *     src/cuda_utils.h:6-16.  The C++ shim reproduces that on a non-zero return.)
 *   - `stream` parameters are `cudaStream_t` passed as void* (NULL = legacy default stream).
 *   - a handle is bound to one device; calls may come from any host thread, and
 *     ptcore_render_* may be called concurrently on different streams (as StreamThread does,
 *     reference src/StreamThread.h:83-84); setters must not race with renders (same contract as
 *     the reference: setters run while the workers are parked, src/RenderManager.h:146-183).
 *   - image space is the reference's: RenderTask offsets are in bottom-up pixel space and the
 *     framebuffer row 0 is the TOP image row (src/DevicePathTracer.h:77-79).
 */
#ifndef PTCORE_H
#define PTCORE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTCORE_ABI_VERSION 1

typedef struct ptcore ptcore_t;

/* error codes outside the cudaError_t range */
enum {
    PT_OK = 0,
    PT_ERR_INVALID_ARGUMENT = 100001,
    PT_ERR_NO_SCENE = 100002,
    PT_ERR_NO_FRAMEBUFFER = 100003,
    PT_ERR_UNSUPPORTED = 100004,
    PT_ERR_SYSTEM = 100005
};

/* Same enumerators and order as `enum material_type`, reference src/HostScene.h:20-26. */
enum {
    PT_MAT_LAMBERTIAN = 0,   /* reference src/material.h:110-127 (dead code there, live here) */
    PT_MAT_METAL = 1,        /* src/material.h:130-144 */
    PT_MAT_DIELECTRIC = 2,   /* src/material.h:146-179 */
    PT_MAT_DIFFUSE_LIGHT = 3,/* src/material.h:210-217 */
    PT_MAT_UNIVERSAL = 4     /* src/material.h:40-102 — the only class the reference instantiates */
};

/* 44 bytes; identical to the material record of the .ptscene file. */
typedef struct PtMaterial {
    int32_t type;      /* PT_MAT_* */
    float base[3];     /* baseColorFactor / albedo */
    float emis[3];     /* emissiveFactor (UNIVERSAL: multiplied by 50, src/material.h:80-86) / light colour */
    int32_t base_tex;  /* texture index or -1 */
    int32_t emis_tex;  /* texture index or -1 */
    float fuzz;        /* METAL */
    float ior;         /* DIELECTRIC */
} PtMaterial;

/* float3 texels in 0..255, row 0 = top row of the image file: HostTexture, src/HostScene.h:48-52 */
typedef struct PtTexture {
    int32_t width, height;
    const float *rgb; /* width*height*3 */
} PtTexture;

/* Flat view of HostScene (src/HostScene.h:54-58). All pointers are HOST pointers and are only
 * read during ptcore_upload_scene. */
typedef struct PtSceneDesc {
    int32_t n_tris;
    const float *tri_pos;   /* [n_tris][3 vertices][xyz] */
    const float *tri_uv;    /* [n_tris][3 vertices][uv]  */
    const int32_t *tri_mat; /* [n_tris] material index   */
    int32_t n_spheres;
    const float *sph;       /* [n_spheres][cx,cy,cz,r]   */
    const int32_t *sph_mat; /* [n_spheres]               */
    int32_t n_mats;
    const PtMaterial *mats;
    int32_t n_tex;
    const PtTexture *tex;
} PtSceneDesc;

/* CameraConfig, reference src/CameraConfig.h:5-17 (pitch/yaw are UI state and not part of the path). */
typedef struct PtCamera {
    float look_from[3];
    float front[3];
    float vfov; /* degrees */
    float hfov; /* degrees, independent of the aspect ratio (src/camera.h:24-35) */
} PtCamera;

/* RenderTask, reference src/DevicePathTracer.h:19-25 (without the host-side `time`). */
typedef struct PtTile {
    int32_t width, height, offset_x, offset_y;
} PtTile;

enum { /* ptcore_set_option keys */
    PT_OPT_KERNEL = 1,       /* PT_KERNEL_* */
    PT_OPT_COUNT_TESTS = 2,  /* 1: count box/triangle tests per ray (stats build of the kernel, slower) */
    PT_OPT_BVH_LEAF_MAX = 3, /* max triangles per leaf for the next upload (default 4) */
    PT_OPT_BLOCKS_PER_SM = 4,/* persistent grid = SMs x this (0 = auto) */
    /* 5 is reserved */
    PT_OPT_REFILL_AT = 6,    /* wavefront kernel: finished lanes per warp that trigger a shade/refill pass (1..32; 0 = the default, 20) */
    PT_OPT_NODE_BURST = 7,   /* wavefront kernel: node steps per warp vote (1..4) */
    PT_OPT_MIN_BLOCKS = 8,   /* accepted for compatibility: only the __launch_bounds__(128, 8) (64-register) build is shipped */
    PT_OPT_BVH_WIDTH = 9,    /* wavefront kernel: walk the 2-wide (64 B nodes, default) or the collapsed 4-wide (128 B nodes) tree; 4-wide measured 20 % slower on cornell_duck */
    PT_OPT_NODE_FORMAT = 10, /* wavefront kernel, 2-wide tree: PT_NODES_* */
    PT_OPT_SAH_INTERSECT_COST = 11, /* next ptcore_upload_scene: cost of one primitive test relative to one node visit, in hundredths (default 120) */
    PT_OPT_POOL_SLOTS = 12,  /* pool kernel: pixel slots per warp, 32..96 (0 = auto: pixels of the launch / warps of the grid, clamped) */
    PT_OPT_POOL_IDLE_AT = 13,/* pool kernel: with no ready ray left, hits waiting per warp that trigger a shade pass (1..32, default 8) */
    PT_OPT_POOL_PERIOD = 15, /* pool kernel: traverse iterations between two rounds of retiring finished rays / pulling new ones (1, 2, 4, 8; default 2) */
    PT_OPT_POOL_CARVEOUT = 16,/* pool kernel: preferred shared-memory carve-out in percent of 228 KB (default 28 = the 64 KB configuration; -1 = driver default) */
    PT_OPT_SMEM_NODES = 17,  /* wavefront kernel: 1 (default) = run as one 1024-thread CTA per SM that keeps the quantised node array in shared memory when it fits 160 KB (cornell_duck: 67 KB, +5 %), 0 = always fetch nodes through L1 */
    PT_OPT_LANES_PER_WARP = 18, /* wavefront kernel: lanes of every warp that take pixels (1..32, default 32) */
    PT_OPT_STICKY_TEXTURES = 19, /* next ptcore_upload_scene: 1 (default) = a UNIVERSAL material without a texture inherits the last texture index seen, as the reference's loadMaterials does (src/DevicePathTracer.h:269-279); 0 = indices as given */
    PT_OPT_RNG_MODE = 20,    /* PT_RNG_*: which random stream ptcore_render_frame_host uses (the explicit ptcore_render_keyed_async is always keyed) */
    PT_OPT_RNG_CHUNKS = 21,  /* PT_RNG_SAMPLE_KEYED through ptcore_render_frame_host: work items per pixel (default 16) */
    PT_OPT_GRID_CTAS = 23,   /* wavefront kernels: CTAs per launch (0 = default: one per SM, or SMs x occupancy for the 128-thread kernels) */
    PT_OPT_CTA_WARPS = 24,   /* shared-memory-node kernel: warps per CTA = per SM, 1..32 (0 = default: by the launch's pixel count — 32 for a full frame, down to 12 for launches with under ~3 pixels per lane, whose sequential sample chains end sooner on fewer warps) */
    PT_OPT_L2_PERSIST_NODES = 22, /* 1 = mark the node array persisting in L2 (access-policy window on the launch stream); for scenes larger than L2 */
    PT_OPT_WATCHDOG = 14     /* pool kernel, debugging aid: bound on the traverse iterations of a warp (0 = none); a launch that hits it renders garbage instead of hanging */
};
enum {
    PT_RNG_STREAM = 0,       /* default, the parity mode: one XORWOW stream per pixel, curand_init(1984 + pixel, 0, 0), samples consumed in
                                sequence exactly as the reference does (src/DevicePathTracer.h:54,80-87): bit-exact images, but a pixel's
                                spp samples form ONE sequential chain */
    PT_RNG_SAMPLE_KEYED = 1  /* throughput mode: the stream is keyed by (pixel, sample): sample s of pixel p draws from
                                curand_init(splitmix64(1984 + p + s * W * H), 0, 0); a pixel's samples are independent work items.  Same estimator,
                                different random numbers: parity is statistical (converged RMSE), images are still independent of how the
                                work is split over launches and GPUs */
};
enum {
    PT_NODES_AUTO = 0,       /* default: quantised when PtStats.quant_inflation <= 1.3, else full */
    PT_NODES_FULL = 1,       /* 64-byte nodes, float planes */
    PT_NODES_QUANTISED = 2   /* 32-byte nodes, 15-bit planes on a scene-wide grid, rounded outwards (same pixels; smaller working set, looser boxes) */
};
enum {
    PT_KERNEL_PERSISTENT = 0, /* persistent-thread wavefront: per-lane pixel refill + warp-voted uniform traversal steps (default) */
    PT_KERNEL_DIRECT = 1,     /* one thread per pixel, no refill: the plain parity slice */
    PT_KERNEL_LOCKSTEP = 2,   /* persistent threads with per-lane refill but a per-lane traversal loop (A/B baseline) */
    PT_KERNEL_POOL = 3        /* persistent warps with a pool of pixels each: rays are pulled by lanes from a shared-memory ring, hits are shaded one material class at a time (pt_pool.cuh) */
};

typedef struct PtStats {
    uint64_t samples;     /* camera paths started */
    uint64_t rays;        /* closest-hit queries (camera + bounce segments) */
    uint64_t box_tests;   /* only with PT_OPT_COUNT_TESTS */
    uint64_t tri_tests;   /* only with PT_OPT_COUNT_TESTS (BVH leaf tests; light-pdf tests excluded) */
    uint64_t light_tests; /* only with PT_OPT_COUNT_TESTS */
    uint64_t launches;    /* kernel launches issued by this handle */
    uint32_t bvh_nodes, bvh_leaves, bvh_depth, n_lights;
    double bvh_build_ms;  /* host time of the last scene compile */
    double sah_cost;
    uint64_t scene_bytes; /* size of the compiled device blob */
    uint32_t bvh4_nodes, bvh4_depth; /* the collapsed four-wide tree */
    double quant_inflation; /* mean over the leaves of (box area in the quantised nodes / in the float nodes); 1 = nothing lost */
    uint32_t n_vertices;    /* unique vertex positions of the compiled scene (indexed primitive layout; 0 with the direct layout) */
    uint32_t reserved0;
} PtStats;

/* ---- lifetime (DevicePathTracer ctor/dtor, src/DevicePathTracer.h:169-192,372-377) ---- */
int ptcore_abi_version(void);
int ptcore_create(int device, ptcore_t **out);
int ptcore_destroy(ptcore_t *h);
const char *ptcore_last_error(const ptcore_t *h); /* h may be NULL: last error of ptcore_create */

/* ---- scene (reloadWorld + loadTextures/loadMaterials/loadTrianglesWithTextures, :241-340;
 *      replaces create_world<<<1,1>>> / create_lights<<<1,1>>> and the device-side BVH build of
 *      src/bvh.h:20-176 with a host SAH build + one bulk upload) ---- */
int ptcore_upload_scene(ptcore_t *h, const PtSceneDesc *scene);
/* re-copies the already compiled scene blob host->device on `stream` (end-to-end timing, hot reload) */
int ptcore_reupload_scene(ptcore_t *h, void *stream, uint64_t *bytes_copied);

/* ---- camera (reloadCamera :230-239 + the by-value CameraConfig of every launch :210) ---- */
int ptcore_set_camera(ptcore_t *h, const PtCamera *cam);

/* ---- parameters (setSamplesPerPixel / setRecursionDepth / setThreadBlockSize :360-370) ---- */
int ptcore_set_params(ptcore_t *h, uint32_t samples_per_pixel, uint32_t recursion_depth);
int ptcore_set_thread_block_size(ptcore_t *h, uint32_t bx, uint32_t by); /* accepted for API parity; the persistent kernel ignores it */
int ptcore_set_option(ptcore_t *h, int key, int64_t value);

/* ---- framebuffer (setFramebuffer :342-358; replaces render_init<<<>>> and the 48 B/pixel
 *      curandState array: XORWOW states are derived from the pixel index on the fly) ----
 * rgb: W*H*3 bytes, yuv: W*H*3/2 bytes (I420), DEVICE or MANAGED pointers; yuv may be NULL. */
int ptcore_bind_framebuffer(ptcore_t *h, uint8_t *rgb, uint8_t *yuv, uint32_t width, uint32_t height);

/* ---- render (renderTaskAsync :194-214 / synchronizeStream :222-226 / waitForRenderTask :216-220) ---- */
int ptcore_render_tile_async(ptcore_t *h, int32_t offset_x, int32_t offset_y, int32_t width, int32_t height, void *stream);
int ptcore_render_tiles_async(ptcore_t *h, const PtTile *tiles, int32_t n_tiles, void *stream);
/* Explicit work list: 8x4-pixel blocks, blocks[i] = bx | (by << 16) with block origin (8*bx, 4*by) in the same bottom-up
 * pixel space as RenderTask.  `blocks_dev` is a DEVICE pointer that must stay valid until the launch has finished.  Lanes take
 * blocks in list order, so a list sorted by descending cost gives longest-processing-time-first scheduling. */
int ptcore_render_blocks_async(ptcore_t *h, const uint32_t *blocks_dev, uint32_t n_blocks, void *stream);
/* Pilot pass: traces `pilot_spp` samples of every pixel (same RNG streams, nothing is stored) and accumulates the number of
 * rays per 8x4 block into costs_dev[by * ceil(W/8) + bx] (DEVICE uint32 array, zeroed by the call). */
int ptcore_block_costs_async(ptcore_t *h, uint32_t pilot_spp, uint32_t *costs_dev, void *stream);
/* The same for blocks [first_block, first_block + n_blocks) of the row-major block grid only (the rest of costs_dev stays zero):
 * N ranks each measure a slice and sum the maps (one small all-reduce) instead of all tracing the whole pilot frame. */
int ptcore_block_costs_range_async(ptcore_t *h, uint32_t pilot_spp, uint32_t *costs_dev, uint32_t first_block, uint32_t n_blocks, void *stream);
/* PT_RNG_SAMPLE_KEYED, explicit form.  A pixel's spp samples are cut into n_chunks chunks of ceil(spp / n_chunks); this call traces chunks
 * first_chunk, first_chunk + chunk_step, ... of every pixel of the listed 8x4 blocks (blocks_dev == NULL: the whole bound framebuffer) and
 * stores each chunk's colour sum to accum_dev[(chunk * W * H + pixel) * 3 .. + 3) (DEVICE floats, zeroed by the caller).  N ranks take
 * first_chunk = rank, chunk_step = N and sum their buffers (each entry is written by exactly one rank, so the float sum is exact).
 * ptcore_resolve_keyed_async then adds the chunks of every pixel in chunk order and stores RGB8 + I420 into the bound framebuffer. */
int ptcore_render_keyed_async(ptcore_t *h, const uint32_t *blocks_dev, uint32_t n_blocks, float *accum_dev, uint32_t n_chunks, uint32_t first_chunk, uint32_t chunk_step,
                              void *stream);
int ptcore_resolve_keyed_async(ptcore_t *h, const float *accum_dev, uint32_t n_chunks, void *stream);
/* Multi-GPU gather (replaces the page migration of the reference's cudaMallocManaged framebuffer, src/Framebuffer.h:27-35): copies the
 * pixels of the listed 8x4 blocks (RGB, Y and the U / V samples they own) from the framebuffer bound to `h` into another frame of the
 * same size — typically the master copy on GPU 0, written with peer stores over NVLink (peer access must be enabled by the caller) —
 * on `stream`, i.e. right behind the ptcore_render_blocks_async that produced them.  dst_yuv may be NULL. */
int ptcore_gather_blocks_async(ptcore_t *h, uint8_t *dst_rgb, uint8_t *dst_yuv, const uint32_t *blocks_dev, uint32_t n_blocks, void *stream);
/* Measurement aid: while set, every warp of the wavefront kernel stores the GPU's nanosecond timer at its start and at its exit into
 * log_dev[2 * warp] / [2 * warp + 1] (DEVICE array of 2 * n_warps uint64; NULL switches it off).  bench.py derives from it when
 * 50 / 90 / 99 % of the lanes of a launch had retired. */
int ptcore_set_retire_log(ptcore_t *h, uint64_t *log_dev, uint32_t n_warps);
int ptcore_sync(ptcore_t *h, void *stream);
int ptcore_wait(ptcore_t *h);

/* Whole-frame convenience with HOST output buffers: binds an internal device framebuffer, renders
 * every pixel, copies RGB (and I420 if yuv_host != NULL) back.  What `e2e` in bench.py times. */
int ptcore_render_frame_host(ptcore_t *h, uint32_t width, uint32_t height, uint8_t *rgb_host, uint8_t *yuv_host);

/* ---- statistics ---- */
int ptcore_get_stats(ptcore_t *h, PtStats *out); /* synchronises the device */
int ptcore_reset_stats(ptcore_t *h);

/* ---- parity probe: walks ONE pixel (current spp/depth/camera/scene) with one device thread and records every
 *      ray as 16 floats: sample, bounce, original primitive id (-1 = miss), t, bary u, bary v, origin xyz,
 *      direction xyz, throughput xyz before shading, 0.  col receives the un-quantised pixel sum. ---- */
int ptcore_debug_trace_pixel(ptcore_t *h, uint32_t width, uint32_t height, int32_t x, int32_t y, float *events, int32_t max_events,
                             int32_t *n_events, float *col);

/* Host-only structural self-test of the scene compiler: builds the BVH for `scene` exactly as ptcore_upload_scene does and
 * validates it (every primitive in exactly one leaf, child boxes inside parents and around their primitives, references in
 * range, depth within the device stack, leaves <= leaf_max).  Needs no GPU.  msg receives the reason on failure. */
int pt_bvh_selftest(const PtSceneDesc *scene, int32_t leaf_max, PtStats *out, char *msg, size_t msg_len);
/* Host emulation of the kernel's two slab tests (float planes / quantised planes) on pseudo-random rays against the boxes of
 * the scene's tree: a box accepted on float planes must be accepted on quantised planes.
 * counts = {box tests, accepted on float planes, accepted on quantised planes}.  No GPU needed. */
int pt_quant_selftest(const PtSceneDesc *scene, uint32_t n_rays, uint32_t seed, uint64_t counts[3], char *msg, size_t msg_len);
/* Host restatement of the kernel's resumable walk (node steps with a held leaf and a sentinel stack, two-primitive leaf steps,
 * tie rule) on the scene's triangles: for n_rays pseudo-random rays the closest hit of the walk over float planes, of the walk
 * over quantised planes and of testing every triangle must coincide (same triangle, same t, u, v).
 * counts = {rays, rays that hit, node steps on float planes, node steps on quantised planes}.  No GPU needed. */
int pt_walk_selftest(const PtSceneDesc *scene, uint32_t n_rays, uint32_t seed, uint64_t counts[4], char *msg, size_t msg_len);

/* ---- multi-GPU plumbing: a tile counter shared by the ranks of one node (POSIX shared memory).
 *      Replaces the per-frame static rectangles of RenderManager/TaskGenerator
 *      (src/RenderManager.h:42-59, src/Scheduling/TaskGenerator.h:58-80) with dynamic claims. ---- */
typedef struct pt_tileq pt_tileq_t;
int pt_tileq_open(const char *name, int create, pt_tileq_t **out);
int64_t pt_tileq_claim(pt_tileq_t *q, int64_t count, int64_t limit); /* returns first claimed index, or -1 when >= limit */
int pt_tileq_reset(pt_tileq_t *q);
int pt_tileq_close(pt_tileq_t *q, int unlink_name);

/* ---- scene files (SceneLoader::load, src/HostScene.cpp:98-139, without assimp) ----
 * Loads .glb/.gltf/.obj/.ptscene into library-owned memory exposed as a PtSceneDesc. */
typedef struct ptscene ptscene_t;
int ptscene_load(const char *path, ptscene_t **out, char *err, size_t err_len);
const PtSceneDesc *ptscene_desc(const ptscene_t *s);
int ptscene_save(const ptscene_t *s, const char *path);
void ptscene_free(ptscene_t *s);

/* P6 writer for an RGB8 framebuffer (the README's out.ppm, README.md:52-58; the reference has no writer) */
int pt_write_ppm(const char *path, const uint8_t *rgb, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif /* PTCORE_H */
