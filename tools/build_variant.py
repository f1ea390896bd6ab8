#!/usr/bin/env python3
"""A/B builds of libptcore.so with extra compile-time definitions (same ABI):
    tools/build_variant.py NAME -DPT_POOL_MINB=3 ...   ->  _variants/NAME/libptcore.so   (select it with PTB200_LIBPTCORE=...)"""
import importlib
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
b = importlib.import_module("multi-gpu-path-tracer_b200._build")
name, flags = sys.argv[1], sys.argv[2:]
out = ROOT / "_variants" / name
out.mkdir(parents=True, exist_ok=True)
cmd = [b.NVCC, *b.NVCC_FLAGS, *flags, "-shared", *map(str, b.CORE_SOURCES), "-lz", "-lrt", "-o", str(out / "libptcore.so")]
subprocess.run(cmd, check=True)
print(out / "libptcore.so")
