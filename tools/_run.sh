set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 3
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -n 1 gpurun_out/bench_default.log | cut -c 1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -n 1 gpurun_out/bench_reference.log | cut -c 1-400
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-e2e > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pt_wavefront -s 2 -c 1 -f -o gpurun_out/prof_r01_wavefront_final3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-e2e > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 1 --spp 16 --no-cpu-baseline --no-ref-gpu > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --spp 16 --no-cpu-baseline --no-ref-gpu > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*final3*
