timeout 300 python -m pytest tests/test_gpu_host_api.py -m gpu -x -q -k monitor 2>&1 | tail -n 3
for b in 8 6 5 4 3; do for r in 24 12; do
timeout 300 python bench.py --steps 3 --warmup 1 --emulate-world 8 --blocks-per-sm $b --refill-at $r --no-cpu-baseline --no-ref-gpu --no-e2e 2>&1 | tail -n 1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('emu8 blocks $b refill $r ms', round(j['ms_per_step'],1))"
done; done
