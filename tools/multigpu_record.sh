#!/bin/bash
# Everything that needs N GPUs of one box, in one gpurun call (charged N x):   gpurun --gpus 8 -- bash tools/multigpu_record.sh 8
# Writes gpurun_out/r02_*_n$N.* (copy what is to be kept to profiles/).
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > $OUT/r02_scale_n$N.log 2>&1
tail -1 $OUT/r02_scale_n$N.log > $OUT/r02_scale_n$N.json
if [ -z "$SKIP_C5" ]; then
$TR --master-port 29522 bench.py --gpus $N --steps 3 --warmup 2 --workload config5 --no-ref-gpu > $OUT/r02_config5_n$N.log 2>&1
tail -1 $OUT/r02_config5_n$N.log > $OUT/r02_config5_n$N.json
fi
# the in-process C++ host (RenderManager + worker threads + P2P tile gather), the reference's model of multi-GPU
python -c "import gzip,sys; open('$OUT/duck.ptscene','wb').write(gzip.decompress(open('tests/golden/cornell_duck.ptscene.gz','rb').read()))"
CLI=multi-gpu-path-tracer_b200/_lib/cuda_project
for sched in lpt dynamic fsfl; do
  timeout 300 $CLI 0 $OUT/duck.ptscene --width 1920 --height 1080 --spp 1024 --depth 10 --gpus $N --streams 2 --scheduler $sched --tile 128x64 --frames 3 --show-tasks 0 --out $OUT/cli_n$N.ppm > $OUT/r02_cuda_project_${sched}_n$N.log 2>&1
  grep CUDA_PROJECT_JSON $OUT/r02_cuda_project_${sched}_n$N.log | sed 's/CUDA_PROJECT_JSON //' > $OUT/r02_cuda_project_${sched}_n$N.json
done
$CLI 0 $OUT/duck.ptscene --width 1920 --height 1080 --spp 1024 --depth 10 --gpus 1 --frames 2 --show-tasks 0 --out $OUT/cli_n1.ppm > $OUT/r02_cuda_project_n1.log 2>&1
cmp $OUT/cli_n$N.ppm $OUT/cli_n1.ppm && echo "cuda_project: $N-GPU frame == 1-GPU frame" > $OUT/r02_cuda_project_cmp_n$N.txt
rm -f $OUT/duck.ptscene $OUT/cli_n$N.ppm $OUT/cli_n1.ppm
timeout 600 python -m pytest tests -x -q -m gpu -k "several_gpus or lpt" > $OUT/r02_pytest_n$N.log 2>&1
tail -3 $OUT/r02_pytest_n$N.log
for f in $OUT/r02_scale_n$N.json $OUT/r02_config5_n$N.json; do [ -s $f ] && python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print(sys.argv[1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "stages", {k: round(v, 2) for k, v in (d.get("stages_ms") or {}).items()}, "retire", d.get("warp_retire"), "keyed", (d.get("rng_keyed") or {}).get("value"), "ref_gpu", (d.get("ref_gpu") or {}).get("value"))
PY
done
cat $OUT/r02_cuda_project_*_n$N.json
