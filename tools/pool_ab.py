#!/usr/bin/env python3
"""A/B of the render kernels on one GPU: pt_pool_kernel against pt_wavefront_kernel.

    python tools/pool_ab.py [--spp 128] [--size 1920x1080] [--fixtures] [--share N]

1. (--fixtures) gate A: every tests/golden/ref_gpu_* fixture rendered with the pool kernel must equal the reference's
   CUDA renderer byte for byte.
2. the same frame with both kernels: images must be identical; device time of each (CUDA events).
3. (--share N) rank 0's share of an N-rank frame (cost-sorted blocks, every N-th), both kernels.
Prints one JSON object per measurement.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=128)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--fixtures", action="store_true")
    ap.add_argument("--share", type=int, default=0)
    ap.add_argument("--watchdog", type=int, default=200_000_000)
    ap.add_argument("--variants", default="0:8:2:28", help="comma-separated pool variants slots:idle_at:period:carveout")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--scene", default=str(GOLD / "cornell_duck.ptscene.gz"))
    args = ap.parse_args()
    import torch
    from PIL import Image

    dev = torch.device("cuda", 0)
    scene = ptb200.load_scene_file(args.scene)
    pt = ptb200.PathTracer(0)
    pt.upload_scene(scene)
    pt.set_option(ptb200.PT_OPT_WATCHDOG, args.watchdog)

    if args.fixtures:
        meta = json.loads((GOLD / "ref_gpu_images.json").read_text())["images"]
        for name, m in sorted(meta.items()):
            ex = m["extra"]
            cam = {}
            if ex:
                v = [float(x) for x in ex[1:9]]
                cam = dict(look_from=tuple(v[0:3]), front=tuple(v[3:6]), vfov=v[6], hfov=v[7])
            pt.set_camera(**cam)
            pt.set_params(m["spp"], m["depth"])
            ref = np.array(Image.open(GOLD / f"ref_gpu_{name}.png").convert("RGB"))
            out = {}
            for kname, k in (("wavefront", ptb200.PT_KERNEL_PERSISTENT), ("pool", ptb200.PT_KERNEL_POOL)):
                pt.set_option(ptb200.PT_OPT_KERNEL, k)
                rgb, _ = pt.render_frame_host(m["width"], m["height"])
                out[kname] = int((np.abs(rgb.astype(int) - ref.astype(int)).max(axis=2) > 0).sum())
            print(json.dumps({"fixture": name, "pixels_differing_from_ref_gpu": out}), flush=True)

    w, h = (int(x) for x in args.size.split("x"))
    pt.set_camera()
    pt.set_params(args.spp, 10)
    rgb = torch.zeros(w * h * 3, dtype=torch.uint8, device=dev)
    yuv = torch.zeros(w * h * 3 // 2, dtype=torch.uint8, device=dev)
    pt.bind_framebuffer(rgb.data_ptr(), yuv.data_ptr(), w, h)
    stream = torch.cuda.current_stream(dev)

    def timed(fn):
        best = None
        for _ in range(args.reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best

    images = {}
    variants = [("wavefront", ptb200.PT_KERNEL_PERSISTENT, (0, 8, 2, 28)), ("wavefront+smem_nodes", ptb200.PT_KERNEL_PERSISTENT, (0, 8, 2, 28, 1))]
    for v in args.variants.split(","):
        variants.append((f"pool[{v}]", ptb200.PT_KERNEL_POOL, tuple(int(x) for x in v.split(":"))))

    def select(k, v):
        pt.set_option(ptb200.PT_OPT_KERNEL, k)
        pt.set_option(ptb200.PT_OPT_POOL_SLOTS, v[0])
        pt.set_option(ptb200.PT_OPT_POOL_IDLE_AT, v[1])
        pt.set_option(ptb200.PT_OPT_POOL_PERIOD, v[2])
        pt.set_option(ptb200.PT_OPT_POOL_CARVEOUT, v[3])
        pt.set_option(ptb200.PT_OPT_SMEM_NODES, v[4] if len(v) > 4 else 0)

    for kname, k, v in variants:
        select(k, v)
        rgb.zero_()
        pt.reset_stats()
        ms = timed(lambda: pt.render_tiles_async([(0, 0, w, h)], stream.cuda_stream))
        st = pt.stats()
        images[kname] = rgb.cpu().numpy().copy()
        same = bool(np.array_equal(images[kname], images["wavefront"]))
        print(json.dumps({"frame": f"{w}x{h} spp={args.spp}", "kernel": kname, "ms": ms, "msamples_per_s": w * h * args.spp / ms / 1e3,
                          "rays": st["rays"] // args.reps, "identical_to_wavefront": same,
                          "n_diff": int((images[kname] != images["wavefront"]).sum())}), flush=True)

    if args.share:
        bw, bh = (w + 7) // 8, (h + 3) // 4
        costs = torch.zeros(bw * bh, dtype=torch.int32, device=dev)
        pt.set_option(ptb200.PT_OPT_KERNEL, ptb200.PT_KERNEL_PERSISTENT)
        pt.block_costs_async(4, costs.data_ptr(), stream.cuda_stream)
        order = torch.argsort(costs, descending=True, stable=True)
        mine = order[0::args.share]
        blocks = ((mine % bw) | ((mine // bw) << 16)).to(torch.int32).contiguous()
        ref_img = None
        for kname, k, v in variants:
            select(k, v)
            rgb.zero_()
            ms = timed(lambda: pt.render_blocks_async(blocks.data_ptr(), int(blocks.numel()), stream.cuda_stream))
            img = rgb.cpu().numpy().copy()
            if ref_img is None:
                ref_img = img
            print(json.dumps({"share": f"1/{args.share} of {w}x{h} spp={args.spp} (cost-sorted blocks)", "kernel": kname, "ms": ms,
                              "ideal_ms_at_full_rate": None, "identical": bool(np.array_equal(img, ref_img))}), flush=True)
    pt.close()


if __name__ == "__main__":
    main()
