#!/usr/bin/env python3
"""One launch per requested kernel on one frame (for `ncu -k regex:pt_ ...`): tools/prof_one.py --kernels wavefront,pool --spp 32"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernels", default="pool")
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--scene", default=str(ROOT / "tests" / "golden" / "cornell_duck.ptscene.gz"))
ap.add_argument("--pool-slots", type=int, default=0)
ap.add_argument("--idle-at", type=int, default=8)
args = ap.parse_args()
w, h = (int(x) for x in args.size.split("x"))
K = {"wavefront": ptb200.PT_KERNEL_PERSISTENT, "pool": ptb200.PT_KERNEL_POOL, "direct": ptb200.PT_KERNEL_DIRECT, "lockstep": ptb200.PT_KERNEL_LOCKSTEP}
pt = ptb200.PathTracer(0)
pt.upload_scene(ptb200.load_scene_file(args.scene))
pt.set_camera()
pt.set_params(args.spp, 10)
pt.set_option(ptb200.PT_OPT_POOL_SLOTS, args.pool_slots)
pt.set_option(ptb200.PT_OPT_POOL_IDLE_AT, args.idle_at)
for k in args.kernels.split(","):
    pt.set_option(ptb200.PT_OPT_KERNEL, K[k])
    rgb, _ = pt.render_frame_host(w, h)
    print(k, int(rgb.sum()))
pt.close()
