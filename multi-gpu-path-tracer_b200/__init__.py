"""multi-gpu-path-tracer_b200 — B200-native path-tracing core behind the reference's device API.

Layout
    csrc/core   CUDA kernels (sm_100a) + the C ABI of include/ptcore.h  -> _lib/libptcore.so
    csrc/host   C++ mirror of the reference's host API (DevicePathTracer, RenderManager, ...)
    capi.py     ctypes binding (plumbing for tests / bench / multi-rank scheduling)
    scenes.py   synthetic scene generators for the BASELINE configs without a model file
    sched.py    one-process-per-GPU dynamic tile scheduling + NCCL framebuffer gather

The directory name is the project's; import it as `import ptb200` (shim at the repo root) or
`importlib.import_module("multi-gpu-path-tracer_b200")`.
"""
from .capi import (  # noqa: F401
    DEFAULT_CAMERA, MAT_DTYPE, PT_KERNEL_DIRECT, PT_KERNEL_PERSISTENT, PT_KERNEL_LOCKSTEP, PT_KERNEL_POOL, PT_OPT_POOL_SLOTS, PT_OPT_POOL_IDLE_AT, PT_OPT_WATCHDOG, PT_OPT_POOL_PERIOD, PT_OPT_POOL_CARVEOUT, PT_OPT_SMEM_NODES, PT_OPT_LANES_PER_WARP, PT_OPT_STICKY_TEXTURES, PT_OPT_RNG_MODE, PT_OPT_RNG_CHUNKS, PT_OPT_L2_PERSIST_NODES, PT_OPT_GRID_CTAS, PT_OPT_CTA_WARPS, PT_RNG_STREAM, PT_RNG_SAMPLE_KEYED, PT_OPT_REFILL_AT, PT_OPT_NODE_BURST, PT_OPT_MIN_BLOCKS, PT_OPT_BVH_WIDTH, PT_OPT_NODE_FORMAT, PT_OPT_SAH_INTERSECT_COST, PT_NODES_AUTO, PT_NODES_FULL, PT_NODES_QUANTISED, PT_MAT_DIELECTRIC, PT_MAT_DIFFUSE_LIGHT, PT_MAT_LAMBERTIAN,
    PT_MAT_METAL, PT_MAT_UNIVERSAL, PT_OPT_BLOCKS_PER_SM, PT_OPT_BVH_LEAF_MAX, PT_OPT_COUNT_TESTS, PT_OPT_KERNEL,
    PathTracer, PtCamera, PtError, PtMaterial, PtSceneDesc, PtStats, PtTexture, PtTile, Scene, TileQueue, LIB_PATH,
    load_library, load_scene_file, make_camera, write_ppm, bvh_selftest, quant_selftest, walk_selftest,
)

__all__ = [n for n in dir() if not n.startswith("_")]


def __getattr__(name):
    # torch-dependent / heavier submodules are imported on first use: ptb200.sched, ptb200.scenes
    if name in ("sched", "scenes"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
