#!/bin/bash
# round 2 final single-GPU record: smoke, whole -m gpu suite, default bench, reference arm, launch list of one bench step
mkdir -p gpurun_out
(lscpu | grep -E "Model name|^CPU\(s\)"; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv) > gpurun_out/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches_ncu.csv python tools/prove_once.py 4096 2 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/smoke.log; tail -12 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.err; wc -c gpurun_out/bench.log gpurun_out/bench_ref.log
