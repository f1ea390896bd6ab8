"""Build recipe for the native pieces (explicit nvcc / g++ commands, in-tree outputs).

    python -m "multi-gpu-path-tracer_b200._build"        (or __graft_entry__.build())

Outputs (git-ignored, shipped to the GPU box by gpurun):
    multi-gpu-path-tracer_b200/_lib/libptcore.so     the C ABI of include/ptcore.h (CUDA kernels, sm_100a)
    multi-gpu-path-tracer_b200/_lib/ptscene_tool     scene converter
    multi-gpu-path-tracer_b200/_lib/cuda_project     host executable mirroring the reference's CLI
    multi-gpu-path-tracer_b200/_lib/host_api_test, gpu_monitor_test, task_generator_test, argument_loader_test   C++ tests of the host mirror
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
NVCC = str(CUDA_HOME / "bin" / "nvcc")
HOST_CXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else (shutil.which("g++") or "g++")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", *ARCH, "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread"]

CORE_SOURCES = [
    CSRC / "core" / "ptcore.cu",
    CSRC / "core" / "bvh_builder.cpp",
    CSRC / "core" / "tileq.cpp",
    CSRC / "core" / "scene_capi.cpp",
    CSRC / "host" / "SceneLoader.cpp",
]
CORE_DEPS = [
    ROOT / "include" / "ptcore.h",
    CSRC / "core" / "pt_device.cuh",
    CSRC / "core" / "pt_kernels.cuh",
    CSRC / "core" / "bvh_builder.h",
    CSRC / "host" / "HostScene.h",
]


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).exists() and Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(str(c) for c in cmd), flush=True)
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(str(c) for c in cmd)}")
    return r.stdout + r.stderr


def build_core(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "libptcore.so"
    if force or _stale(out, [*CORE_SOURCES, *CORE_DEPS, Path(__file__)]):
        cmd = [NVCC, *NVCC_FLAGS, "-shared", *CORE_SOURCES, "-lz", "-lrt", "-o", out]
        if ptxas_info:
            cmd[1:1] = ["-Xptxas", "-v"]
        log = _run(cmd, verbose)
        if ptxas_info:
            print(log)
    return out


def build_tools(force: bool = False, verbose: bool = False) -> None:
    LIBDIR.mkdir(exist_ok=True)
    tool = LIBDIR / "ptscene_tool"
    srcs = [CSRC / "tools" / "ptscene_tool.cpp", CSRC / "host" / "SceneLoader.cpp"]
    if force or _stale(tool, [*srcs, CSRC / "host" / "HostScene.h"]):
        _run([HOST_CXX, "-std=c++17", "-O2", "-ffp-contract=off", f"-I{CUDA_HOME}/include", *srcs, "-lz", "-o", tool], verbose)
    test_src = CSRC / "host" / "host_api_test.cpp"
    if test_src.exists():
        exe = LIBDIR / "host_api_test"
        deps = list((CSRC / "host").glob("*.h")) + [test_src, CSRC / "host" / "SceneLoader.cpp", ROOT / "include" / "ptcore.h"]
        if force or _stale(exe, deps) or _stale(exe, [LIBDIR / "libptcore.so"]):
            _run([HOST_CXX, "-std=c++17", "-O2", "-pthread", f"-I{CUDA_HOME}/include", f"-I{ROOT / 'include'}", test_src,
                  f"-L{LIBDIR}", "-lptcore", f"-L{CUDA_HOME}/lib64", "-lcudart", "-ldl", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{CUDA_HOME}/lib64", "-o", exe], verbose)
    mon_src = CSRC / "host" / "gpu_monitor_test.cpp"
    if mon_src.exists():
        exe = LIBDIR / "gpu_monitor_test"
        if force or _stale(exe, [mon_src, CSRC / "host" / "GPUMonitor.h", CSRC / "host" / "Renderer.h"]):
            _run([HOST_CXX, "-std=c++17", "-O2", "-pthread", mon_src, "-ldl", "-o", exe], verbose)
    tg_src = CSRC / "host" / "task_generator_test.cpp"
    if tg_src.exists():
        exe = LIBDIR / "task_generator_test"
        if force or _stale(exe, [tg_src, CSRC / "host" / "TaskGenerator.h", CSRC / "host" / "RenderTask.h"]):
            _run([HOST_CXX, "-std=c++17", "-O2", tg_src, "-o", exe], verbose)
    al_src = CSRC / "host" / "argument_loader_test.cpp"
    if al_src.exists():
        exe = LIBDIR / "argument_loader_test"
        if force or _stale(exe, [al_src, CSRC / "host" / "ArgumentLoader.h", CSRC / "host" / "RendererConfig.h"]):
            _run([HOST_CXX, "-std=c++17", "-O2", f"-I{CUDA_HOME}/include", al_src, "-o", exe], verbose)
    cli_src = CSRC / "host" / "main.cpp"
    if cli_src.exists():
        cli = LIBDIR / "cuda_project"
        deps = list((CSRC / "host").glob("*.h")) + [cli_src, ROOT / "include" / "ptcore.h"]
        if force or _stale(cli, deps) or _stale(cli, [LIBDIR / "libptcore.so"]):
            _run([HOST_CXX, "-std=c++17", "-O2", "-pthread", f"-I{CUDA_HOME}/include", f"-I{ROOT / 'include'}", cli_src,
                  f"-L{LIBDIR}", "-lptcore", f"-L{CUDA_HOME}/lib64", "-lcudart", "-ldl", f"-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{CUDA_HOME}/lib64", "-o", cli], verbose)


def build_all(force: bool = False, verbose: bool = False) -> Path:
    lib = build_core(force=force, verbose=verbose)
    build_tools(force=force, verbose=verbose)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIBDIR)
