#!/bin/bash
# The two multi-GPU numbers of the final build in one short gpurun call (charged N x):   gpurun --gpus 8 -- bash tools/multigpu_lean.sh 8
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/r02_scale_n$N.log 2>&1
tail -1 $OUT/r02_scale_n$N.log > $OUT/r02_scale_n$N.json
python -c "import gzip; open('$OUT/duck.ptscene','wb').write(gzip.decompress(open('tests/golden/cornell_duck.ptscene.gz','rb').read()))"
timeout 120 multi-gpu-path-tracer_b200/_lib/cuda_project 0 $OUT/duck.ptscene --width 1920 --height 1080 --spp 1024 --depth 10 --gpus $N --streams 1 --scheduler lpt --frames 3 --show-tasks 0 --out $OUT/cli_n$N.ppm > $OUT/r02_cuda_project_lpt_n$N.log 2>&1
grep CUDA_PROJECT_JSON $OUT/r02_cuda_project_lpt_n$N.log | sed 's/CUDA_PROJECT_JSON //' > $OUT/r02_cuda_project_lpt_n$N.json
rm -f $OUT/duck.ptscene $OUT/cli_n$N.ppm
python - "$OUT/r02_scale_n$N.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "stages", {k: round(v, 2) for k, v in (d.get("stages_ms") or {}).items()}, "keyed", (d.get("rng_keyed") or {}).get("value"), "ref_gpu", (d.get("ref_gpu") or {}).get("value"))
PY
cat $OUT/r02_cuda_project_lpt_n$N.json
