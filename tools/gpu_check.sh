#!/bin/bash
# GPU round-trip: smoke, parity tests, bench (and optionally the ncu launch list).  Usage: tools/gpu_check.sh [ncu]
mkdir -p gpurun_out
(lscpu | grep -E "Model name|^CPU\(s\)"; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv) > gpurun_out/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --saturated-pairs 0 > gpurun_out/ncu_bench.log 2>&1
fi
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
