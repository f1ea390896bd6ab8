#!/usr/bin/env python3
"""bench.py — Msamples/s (and Mrays/s) of the path-tracing hot path on BASELINE config 2:
models/cornell_duck.glb, 1920x1080, 1024 spp, max depth 10, the reference's default camera.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA core (N > 1: under torchrun)
    python bench.py --impl reference [...]                         the reference's own CPU implementation

A "step" is one whole frame (W*H*spp camera paths).  `value` is whole-job throughput with the compiled
scene resident in HBM; `e2e` is the same frame through the C ABI with HOST buffers (scene blob H2D from
pinned memory + camera + render + RGB/I420 D2H inside the timed region).  N > 1 splits the SAME frame
over the ranks (strong scaling): a pilot pass (1/N of the blocks per rank, one small all-reduce) measures rays per 8x4 block, the
blocks are sorted by cost and dealt round-robin (LPT), one NCCL reduce gathers the frame (`--sched tiles`: dynamic tile claims).
Beside `value` the line carries `stages_ms`, `warp_retire`, `rng_keyed` (the (pixel, sample)-keyed RNG mode, statistical parity
only) and `ref_gpu` (the reference's own CUDA renderer with gpuNumber = N on the same box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WIDTH, HEIGHT, SPP, DEPTH = 1920, 1080, 1024, 10
WORKLOAD = f"cornell_duck {WIDTH}x{HEIGHT} spp={SPP} depth={DEPTH} (BASELINE configs[1])"
SCENE = ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"
# algorithmic work per ray on OUR tree, from the core's own counters
FLOPS_PER_RAY = lambda n_box, n_tri: 24.0 * n_box + 45.0 * n_tri + 180.0
BYTES_PER_RAY = lambda n_box, n_tri: 32.0 * n_box + 48.0 * n_tri + 168.0


NCU_SUMMARY = {"persistent": "r02_ncu_wavefront_1080p_1024spp.txt", "pool": "r02_ncu_pool_1080p_1024spp.txt"}


def ncu_counters(kernel: str):
    """Per-launch counters of the dominant kernel on the default workload (1 GPU, 1080p, 1024 spp) from the committed
    `ncu --set full` summary of the CURRENT build (profiles/r02_ncu_*_1080p_1024spp.txt, written by tools/ncu_summary.py):
    executed warp instructions, threads per instruction, L1 / L2 / DRAM bytes.  They are properties of the (deterministic)
    workload and the build, so dividing them by the duration measured live gives achieved rates.  None if the file is absent."""
    p = ROOT / "profiles" / NCU_SUMMARY.get(kernel, "-")
    if not p.exists():
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1.0, "": 1.0, "ms": 1.0, "%": 1.0}
    out = {}
    for ln in p.read_text().splitlines():
        f = ln.split()
        if ln.startswith("== ") and out:
            break  # first kernel of the file only
        if len(f) == 3:
            try:
                out[f[0]] = float(f[2]) * scale.get(f[1], 1.0)
            except ValueError:
                pass
        elif len(f) == 2:
            try:
                out[f[0]] = float(f[1])
            except ValueError:
                pass
    return out or None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + ",".join(self.FIELDS), "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def flat_scene_file(tmpdir: Path) -> Path:
    import ptb200
    flat = tmpdir / "cornell_duck.ptscene"
    if not flat.exists():
        flat.write_bytes(ptb200.load_scene_file(SCENE).to_ptscene_bytes())
    return flat


def cpu_reference_sample(threads: int, rect=None, spp=4, row_stride=9):
    """Times the reference's CPU implementation (oracle/_ref/ref_cpu: its own headers host-compiled, OpenMP) on a bounded sample of the
    workload.  Default: every `row_stride`-th row of the WHOLE 1920x1080 frame at `spp` samples, depth 10 — a uniform sample of the frame's
    pixels, so Msamples/s of the sample is an unbiased estimate of the frame's (the centre rect used before is the most
    expensive region and flattered the ratio by 1.2 - 1.5x).  `rect` = (x, y, w, h): that rectangle only, every row.
    Falls back to the plain-C port (oracle/_build) when the reference-derived binary is not there."""
    import tempfile
    ref_cpu = ROOT / "oracle" / "_ref" / "ref_cpu"
    if rect is None and not ref_cpu.exists():
        rect = (640, 360, 640, 360)  # the port renders rectangles only
    if rect is None:
        rect, stride = (0, 0, WIDTH, HEIGHT), row_stride
        sample = f"{WIDTH}x{HEIGHT} frame, every {stride}th row ({(HEIGHT + stride - 1) // stride} full rows), spp={spp}, depth={DEPTH}"
    else:
        stride = 1
        sample = f"{WIDTH}x{HEIGHT} frame, rect {rect[2]}x{rect[3]} at ({rect[0]},{rect[1]}), spp={spp}, depth={DEPTH}"
    n_samples = rect[2] * ((rect[3] + stride - 1) // stride) * spp
    if ref_cpu.exists():
        with tempfile.TemporaryDirectory() as td:
            flat = flat_scene_file(Path(td))
            out = subprocess.run([str(ref_cpu), str(flat), str(WIDTH), str(HEIGHT), str(spp), str(DEPTH), "-", "--rect", *map(str, rect), "--row-stride", str(stride),
                                  "--threads", str(threads)], check=True, capture_output=True, text=True).stdout
        j = json.loads(out.strip().splitlines()[-1])
        return dict(value=j["msamples_per_s"], unit="Msamples/s", cores=threads, kind="reference", sample=sample, seconds=j["seconds"], samples=n_samples)
    sys.path.insert(0, str(ROOT / "tests"))
    import _oracle
    import ptb200
    orc = _oracle.load()
    scene = ptb200.load_scene_file(SCENE)
    t0 = time.perf_counter()
    orc.render(scene, WIDTH, HEIGHT, spp, DEPTH, rect=rect, threads=threads)
    sec = time.perf_counter() - t0
    return dict(value=n_samples / sec / 1e6, unit="Msamples/s", cores=threads, kind="port", sample=sample, seconds=sec, samples=n_samples)


def gpu_reference_sample(spp=16, gpus=1):
    """The reference's own CUDA renderer (oracle/_ref/ref_gpu: RenderManager + DevicePathTracer + kernels, unmodified,
    compiled for sm_100) on the same box: full 1920x1080 frame at reduced spp, frame >= 2 timed; gpus > 1 = the reference's own
    multi-GPU mode (gpuNumber = N, fixed equal tasks, one managed framebuffer, src/RenderManager.h:42-59)."""
    import tempfile
    ref_gpu = ROOT / "oracle" / "_ref" / "ref_gpu"
    if not ref_gpu.exists():
        return {"unavailable": "oracle/_ref/ref_gpu not built (needs /root/reference at build time)"}
    try:
        with tempfile.TemporaryDirectory() as td:
            flat = flat_scene_file(Path(td))
            r = subprocess.run([str(ref_gpu), str(flat), str(WIDTH), str(HEIGHT), str(spp), str(DEPTH), "-", "--frames", "3", "--gpus", str(gpus)], capture_output=True, text=True, timeout=900)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_GPU_JSON ")]
        if r.returncode != 0 or not line:
            return {"unavailable": f"ref_gpu exited {r.returncode}: {(r.stderr or r.stdout)[-200:]}"}
        j = json.loads(line[-1][len("REF_GPU_JSON "):])
        return dict(value=j["msamples_per_s"], unit="Msamples/s", n_gpus=gpus, kind="reference CUDA renderer (unmodified kernels, sm_100, 8x8 blocks" + (f", gpuNumber={gpus} FSFL)" if gpus > 1 else ")"),
                    sample=f"{WIDTH}x{HEIGHT} spp={spp} depth={DEPTH}, mean of frames 2-3", seconds=j["seconds"], init_seconds=j["init_seconds"])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = host_cores()
    for _ in range(args.warmup):
        cpu_reference_sample(threads, rect=(880, 500, 160, 80), spp=1)
    vals, secs = [], []
    last = None
    for _ in range(args.steps):
        last = cpu_reference_sample(threads, **({"rect": (880, 500, 160, 80), "spp": 1} if args.ref_sample == "small" else {}))
        vals.append(last["value"]); secs.append(last["seconds"])
    value = sum(last["samples"] for _ in vals) / sum(secs) / 1e6
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "reference model cornell_duck (converted fixture), fixed XORWOW seeds",
            "config": {"workload": WORKLOAD, "sample_per_step": last["sample"]},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    global WIDTH, HEIGHT, SPP, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="override for quick experiments (a non-default value is flagged in config)")
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "direct", "lockstep", "pool"])
    ap.add_argument("--cta-warps", type=int, default=0, help="shared-memory-node kernel: warps per SM (0 = library default, chosen per launch from its pixel count)")
    ap.add_argument("--smem-nodes", type=int, default=-1, help="1024-thread wavefront kernel with the quantised nodes in shared memory (-1 = library default)")
    ap.add_argument("--refill-at", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--node-burst", type=int, default=0)
    ap.add_argument("--min-blocks", type=int, default=0)
    ap.add_argument("--bvh-width", type=int, default=0)
    ap.add_argument("--node-format", type=int, default=0, help="1 = 64-byte float nodes, 2 = 32-byte quantised nodes (0 = library default)")
    ap.add_argument("--sched", default="lpt", choices=["lpt", "tiles"], help="lpt: pilot pass + cost-sorted 8x4 blocks dealt round-robin to the ranks; tiles: dynamic tile claims")
    ap.add_argument("--pilot-spp", type=int, default=4)
    ap.add_argument("--emulate-world", type=int, default=0, help="experiments: render only rank 0's share of an N-rank frame on one GPU")
    ap.add_argument("--tile", default="64x32")
    ap.add_argument("--claim", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-retire-log", action="store_true")
    ap.add_argument("--keyed-steps", type=int, default=3, help="extra frames in the (pixel, sample)-keyed RNG mode, reported under `rng_keyed` (0 = skip)")
    ap.add_argument("--keyed-chunks", type=int, default=32)
    ap.add_argument("--ref-sample", default="default", choices=["default", "small"], help="--impl reference: size of the bounded CPU sample (small: for tests)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 (default, the bench line): 1080p/1024 spp; config5: 3840x2160/4096 spp, the strong-scaling case of BASELINE.json (a parity-test case, not the bench line)")
    args = ap.parse_args()
    if args.workload == "config5":
        WIDTH, HEIGHT, SPP = 3840, 2160, 4096
        WORKLOAD = f"cornell_duck {WIDTH}x{HEIGHT} spp={SPP} depth={DEPTH} (BASELINE configs[4])"
        if args.spp == 1024:
            args.spp = SPP

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import ptb200
    sched = ptb200.sched

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    spp = args.spp
    scene = ptb200.load_scene_file(SCENE)
    pt = ptb200.PathTracer(local_rank)
    pt.upload_scene(scene)
    pt.set_camera()
    pt.set_params(spp, DEPTH)
    pt.set_option(ptb200.PT_OPT_KERNEL, {"persistent": ptb200.PT_KERNEL_PERSISTENT, "direct": ptb200.PT_KERNEL_DIRECT, "lockstep": ptb200.PT_KERNEL_LOCKSTEP, "pool": ptb200.PT_KERNEL_POOL}[args.kernel])
    if args.smem_nodes >= 0:
        pt.set_option(ptb200.PT_OPT_SMEM_NODES, args.smem_nodes)
    if args.cta_warps:
        pt.set_option(ptb200.PT_OPT_CTA_WARPS, args.cta_warps)
    if args.refill_at:
        pt.set_option(ptb200.PT_OPT_REFILL_AT, args.refill_at)
    if args.blocks_per_sm:
        pt.set_option(ptb200.PT_OPT_BLOCKS_PER_SM, args.blocks_per_sm)
    if args.node_burst:
        pt.set_option(ptb200.PT_OPT_NODE_BURST, args.node_burst)
    if args.min_blocks:
        pt.set_option(ptb200.PT_OPT_MIN_BLOCKS, args.min_blocks)
    if args.bvh_width:
        pt.set_option(ptb200.PT_OPT_BVH_WIDTH, args.bvh_width)
    if args.node_format:
        pt.set_option(ptb200.PT_OPT_NODE_FORMAT, args.node_format)


    tw, th = (int(x) for x in args.tile.split("x"))
    if world == 1:
        tiles = [(0, 0, WIDTH, HEIGHT)]
        claim = 1
    else:
        tiles = sched.interleave(sched.make_tiles(WIDTH, HEIGHT, tw, th), world * 4)
        claim = args.claim or max(1, len(tiles) // (world * 8))
    plan = sched.FramePlan(WIDTH, HEIGHT, tiles, claim)
    rr = sched.RankRenderer(pt, WIDTH, HEIGHT, device)

    queue = None
    if world > 1:
        qname = f"/ptb200_tileq_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            queue = ptb200.TileQueue(qname, create=True)
        dist.barrier()
        if rank != 0:
            queue = ptb200.TileQueue(qname, create=False)
        dist.barrier()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def render_frame():
        if args.sched == "lpt" and args.kernel in ("persistent", "pool"):
            rr.render_frame_lpt(rank, args.emulate_world or world, args.pilot_spp, gather=not args.emulate_world, emulated=bool(args.emulate_world))
        else:
            rr.render_frame(plan, queue, rank, world)

    def step():
        """one frame; returns device milliseconds (CUDA events on the launching stream)"""
        flush.zero_()
        if world > 1:
            barrier()
            if rank == 0:
                queue.reset()
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        render_frame()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    # ---- algorithmic work per ray on our tree (stats build of the kernel, low spp, outside the timed region) ----
    per_ray = None
    if rank == 0:
        pt.set_option(ptb200.PT_OPT_COUNT_TESTS, 1)
        pt.set_params(4, DEPTH)
        pt.reset_stats()
        pt.render_tiles_async([(0, 0, WIDTH, HEIGHT)])
        st = pt.stats()
        st_build = dict(st)
        per_ray = dict(box=st["box_tests"] / st["rays"], tri=st["tri_tests"] / st["rays"], light=st["light_tests"] / st["rays"], rays_per_sample=st["rays"] / st["samples"])
        pt.set_option(ptb200.PT_OPT_COUNT_TESTS, 0)
        pt.set_params(spp, DEPTH)
    pt.reset_stats()

    for _ in range(args.warmup):
        step()
    pt.reset_stats()
    rr.launches = 0
    rr.reset_stage_times()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    ms = [step() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(ms))
    st = pt.stats()
    stages = rr.stage_times_ms()  # pilot / sort / render / gather, mean device ms per step on this rank
    render_ms = stages.get("render", total_ms / args.steps)
    if world > 1:
        t = torch.tensor([total_ms, float(st["rays"]), float(st["launches"]), render_ms, stages.get("pilot", 0.0), stages.get("gather", 0.0)], dtype=torch.float64, device=device)
        tmax = t.clone()
        tmin = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, rays, launches = float(tmax[0]), float(t[1]), int(t[2])
        stages = dict(stages, render_max_over_ranks=float(tmax[3]), render_min_over_ranks=float(tmin[3]), pilot_max_over_ranks=float(tmax[4]), gather_max_over_ranks=float(tmax[5]))
        render_ms = float(tmax[3])
    else:
        rays, launches = float(st["rays"]), int(st["launches"])

    # ---- when did the warps of the render launch retire?  One extra, untimed frame with the kernel's retire log switched on ----
    retire = None
    if args.sched == "lpt" and args.kernel == "persistent" and not args.no_retire_log:
        n_log = 148 * 64
        log = torch.zeros(2 * n_log, dtype=torch.int64, device=device)
        pt.set_retire_log(log.data_ptr(), n_log)
        rr.render_frame_lpt(rank, args.emulate_world or world, args.pilot_spp, gather=False, emulated=bool(args.emulate_world))
        torch.cuda.synchronize(device)
        pt.set_retire_log(0, 0)
        lg = log.view(n_log, 2)
        # the pilot launch of the same frame logs too (and may run more warps than the render launch, whose entries then survive): a pilot
        # warp ended before the render launch began, i.e. before the latest start in the log; a render warp ends after it
        used = lg[:, 1] > lg[:, 0].max()
        if bool(used.any()):
            t0 = lg[used, 0].min()
            ends = ((lg[used, 1] - t0).double() / 1e6).sort().values
            q = lambda f: float(ends[min(len(ends) - 1, int(f * len(ends)))])
            retire = {"warps": int(used.sum()), "p50_ms": q(0.50), "p90_ms": q(0.90), "p99_ms": q(0.99), "last_ms": float(ends[-1]),
                      "note": "rank 0, one untimed frame: time from the first warp's start until 50 / 90 / 99 / 100 % of the render launch's warps had exited"}
        barrier()
    samples_per_step = WIDTH * HEIGHT * spp
    value = samples_per_step * args.steps / (total_ms / 1e3) / 1e6
    mrays = rays / (total_ms / 1e3) / 1e6

    # ---- the same frame in the (pixel, sample)-keyed RNG mode (statistical parity only: reported separately, never the bench value) ----
    keyed = None
    if args.keyed_steps > 0 and args.kernel == "persistent" and args.sched == "lpt":
        def keyed_step():
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rr.render_frame_keyed(rank, args.emulate_world or world, args.keyed_chunks, args.pilot_spp, gather=not args.emulate_world, emulated=bool(args.emulate_world))
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1)
        keyed_step()
        rr.reset_stage_times()
        kms = float(sum(keyed_step() for _ in range(args.keyed_steps)))
        kstages = rr.stage_times_ms()
        if world > 1:
            kt = torch.tensor([kms], dtype=torch.float64, device=device)
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
            kms = float(kt[0])
        keyed = {"value": WIDTH * HEIGHT * spp * args.keyed_steps / (kms / 1e3) / 1e6, "unit": "Msamples/s", "ms_per_step": kms / args.keyed_steps, "steps": args.keyed_steps,
                 "chunks_per_pixel": args.keyed_chunks, "stages_ms": kstages,
                 "rng": "XORWOW keyed by (pixel, sample): curand_init(splitmix64(1984 + pixel + sample * W * H), 0, 0); parity with the reference is statistical in this mode (converged RMSE <= 1/255, tests/test_gpu_parity.py)",
                 "parallelism": "rank r traces chunks r, r + N, ... of every pixel; one float reduce of the chunk sums; rank 0 adds them in chunk order"}
        rr.reset_stage_times()

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        rgb_host = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.uint8).pin_memory()
        yuv_host = torch.empty((WIDTH * HEIGHT * 3 // 2,), dtype=torch.uint8).pin_memory()
        rgb_np, yuv_np = rgb_host.numpy(), yuv_host.numpy()
        h2d = d2h = 0

        def e2e_step():
            nonlocal h2d, d2h
            if world > 1:
                barrier()
                if rank == 0:
                    queue.reset()
                barrier()
            t0 = time.perf_counter()
            h2d = pt.reupload_scene() + 32  # compiled scene blob from pinned host memory + the camera struct
            pt.set_camera()
            render_frame()
            if rank == 0:
                rgb_host.view(-1).copy_(rr.rgb, non_blocking=True)
                yuv_host.copy_(rr.yuv, non_blocking=True)
                d2h = rgb_np.nbytes + yuv_np.nbytes
            torch.cuda.synchronize(device)
            barrier()
            return time.perf_counter() - t0

        e2e_step()
        secs = [e2e_step() for _ in range(args.steps)]
        tsum = sum(secs)
        if world > 1:
            tt = torch.tensor([tsum], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tsum = float(tt[0])
        e2e = {"value": samples_per_step * args.steps / tsum / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1000.0 * tsum / args.steps}

    if rank == 0:
        peaks = measured_peaks()
        n_kernel_launches = max(1, launches)
        kernel_ms = render_ms                             # the dominant launch (ptcore_render_blocks_async), CUDA events on its stream, mean over the timed steps (max over ranks)
        rays_per_launch = rays / (args.steps * world)      # per GPU
        abytes = BYTES_PER_RAY(per_ray["box"], per_ray["tri"]) * rays_per_launch
        aflops = FLOPS_PER_RAY(per_ray["box"], per_ray["tri"]) * rays_per_launch
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        kernel_s = kernel_ms / 1e3
        kname = {"persistent": "pt_wavefront_kernel", "pool": "pt_pool_kernel", "direct": "pt_direct_kernel", "lockstep": "pt_persistent_kernel"}[args.kernel]
        # Three candidate roofs ("not tensor cores, not HBM for configs 2/3/5"); the one with the largest fraction binds.
        #  (a) fp32: algorithmic FLOPs (24 per box test, 45 per triangle test, 180 per ray) against 148 SMs x 128 lanes x 2 x clock
        #  (b) L1 / L2: bytes the kernel moves through l1tex / lts (ncu, current build) against 128 B/clk/SM and the measured
        #      ~6300 B/clk L2 cap (B300_MICROARCH.md), at the clock sustained in this run
        #  (c) issue: executed warp instructions (ncu, current build; the workload is deterministic) against 4 schedulers x 1
        #      instruction per clock per SM; times threads-per-instruction / 32 = the share of lane-issue slots doing work
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        cand = {"fp32": {"bound": "fp32", "achieved": aflops / kernel_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": aflops / kernel_s / 1e12 / fp32_peak,
                         "algorithmic_flops_per_ray": FLOPS_PER_RAY(per_ray["box"], per_ray["tri"]), "algorithmic_flops_per_launch": aflops}}
        nc = ncu_counters(args.kernel) if (world == 1 and spp == SPP and not args.emulate_world) else None
        traffic = None
        if nc:
            traffic = nc.get("dram__bytes_read.sum", 0.0) + nc.get("dram__bytes_write.sum", 0.0)
            issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions / s
            inst = nc.get("smsp__inst_executed.sum")
            tpi = nc.get("smsp__thread_inst_executed_per_inst_executed.ratio")
            if inst:
                cand["issue"] = {"bound": "issue", "achieved": inst / kernel_s / 1e9, "peak": issue_peak, "unit": "G warp-instructions/s", "frac": inst / kernel_s / 1e9 / issue_peak,
                                 "warp_instructions_per_launch": inst, "threads_per_instruction": tpi, "useful_lane_issue_frac": (inst / kernel_s / 1e9 / issue_peak) * (tpi / 32.0) if tpi else None}
            # L1 / L2: ncu's own fraction of the unit's peak during the captured launch, rescaled to the duration measured live
            t_ncu = nc.get("gpu__time_duration.sum", 0.0) / 1e3  # the summary stores ms
            scale = (t_ncu / kernel_s) if t_ncu > 0 else 1.0
            if nc.get("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"):
                f = nc["l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"] / 100.0 * scale
                cand["l1"] = {"bound": "l1", "achieved": f * 100.0, "peak": 100.0, "unit": "% of the L1 LSU data pipe's wavefront rate", "frac": f}
            if nc.get("lts__throughput.avg.pct_of_peak_sustained_elapsed"):
                f = nc["lts__throughput.avg.pct_of_peak_sustained_elapsed"] / 100.0 * scale
                cand["l2"] = {"bound": "l2", "achieved": f * 100.0, "peak": 100.0, "unit": "% of L2 (lts) throughput", "frac": f,
                              "bytes_per_launch": nc.get("lts__t_sectors.sum", 0.0) * 32.0}
            if traffic:
                cand["hbm"] = {"bound": "hbm", "achieved": traffic / kernel_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": traffic / kernel_s / 1e9 / peaks["hbm_gbs"],
                               "note": "real DRAM bytes of the launch (ncu), not algorithmic bytes: the 5 MB scene is cache-resident"}
        binding = max(cand.values(), key=lambda c: c["frac"])
        roofline = dict(binding)
        roofline.update({"traffic": traffic, "kernel": kname, "peak_source": peaks["source"] + "; SM clock sampled during the timed region",
                         "counters_from": ("profiles/" + NCU_SUMMARY[args.kernel]) if nc else None,
                         "algorithmic_bytes_per_ray": BYTES_PER_RAY(per_ray["box"], per_ray["tri"]), "algorithmic_bytes_per_launch": abytes,
                         "note": "the scene (5 MB) is L1/L2-resident, so HBM is not the roof: `bound` names the largest of the fp32 / L1 / L2 / issue / hbm fractions in `roofs`; algorithmic node+triangle bytes (SURVEY 8d) are served by L1/L2 and are reported only as algorithmic_bytes_*",
                         "nodes": ("32-byte quantised" if (args.node_format == 2 or (args.node_format == 0 and 0 < st_build.get("quant_inflation", 0) <= 1.3)) else "64-byte float") if args.kernel in ("persistent", "pool") and not args.bvh_width == 4 else "64-byte float",
                         "roofs": {k: {kk: vv for kk, vv in v.items() if kk in ("achieved", "peak", "unit", "frac", "useful_lane_issue_frac", "threads_per_instruction")} for k, v in cand.items()}})
        roofline_fp32 = cand["fp32"]
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "reference model cornell_duck (converted fixture), fixed XORWOW seeds",
                "config": {"workload": WORKLOAD if spp == SPP else WORKLOAD + f" [spp overridden to {spp}]", "kernel": args.kernel, "l2": "flushed between steps (256 MiB write)",
                           "parallelism": ("1 GPU" if world == 1 else f"{world} ranks") + (f", pilot pass ({args.pilot_spp} spp) + cost-sorted 8x4 blocks dealt round-robin (LPT), NCCL reduce gather" if args.sched == "lpt" and args.kernel in ("persistent", "pool") else f", image tiles {tw}x{th}, dynamic claims of {claim}, NCCL reduce gather"),
                           "rng": "XORWOW per pixel, reference stream order"},
                "mrays_per_s": mrays, "rays_per_sample": rays / (samples_per_step * args.steps), "per_ray": per_ray, "wall_s_timed_region": t_wall,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_fp32": roofline_fp32,
                "stages_ms": stages, "warp_retire": retire, "rng_keyed": keyed}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample(host_cores())
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_ref_gpu:
            pt.close()
            line["ref_gpu"] = gpu_reference_sample(gpus=world)  # the reference's own CUDA renderer with gpuNumber = N (FSFL, managed framebuffer) beside our N-GPU number
        print(json.dumps(line), flush=True)
    if queue is not None:
        barrier()
        queue.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
