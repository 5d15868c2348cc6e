#!/bin/bash
# round 2: whole -m gpu suite, default bench, reference arm, Poseidon lab (shipped file + open variants)
mkdir -p gpurun_out
(lscpu | grep -E "Model name|^CPU\(s\)"; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv) > gpurun_out/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
bash tools/probe/run_poseidon_lab.sh > /dev/null 2>&1
tail -3 gpurun_out/smoke.log; tail -16 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_ref.log
grep -E "===|this variant|full round" gpurun_out/poseidon_lab.txt
