"""bench.py pieces that run without a GPU: the reference arm (the reference's own CPU code on a bounded sample) and the helpers."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "small"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cornell_duck 1920x1080 spp=1024 depth=10")
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_helpers():
    sys.path.insert(0, str(ROOT))
    import bench
    assert bench.FLOPS_PER_RAY(26.6, 4.1) == 24 * 26.6 + 45 * 4.1 + 180
    assert bench.BYTES_PER_RAY(26.6, 4.1) == 32 * 26.6 + 48 * 4.1 + 168
    p = bench.measured_peaks()
    assert p["hbm_gbs"] > 1000 and "source" in p
    for kernel in ("persistent", "pool"):
        c = bench.ncu_counters(kernel)  # None until the round's capture of that kernel is committed under profiles/
        if c is not None:
            assert c["smsp__inst_executed.sum"] > 1e9 and 1 <= c["smsp__thread_inst_executed_per_inst_executed.ratio"] <= 32
            assert 1e6 < c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"] < 1e12
