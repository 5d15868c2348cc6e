#!/bin/bash
# Poseidon lab on the box's host CPU (all variants) + one default bench
mkdir -p gpurun_out
(lscpu | grep -E "Model name|^CPU\(s\)"; grep -o -E "avx512[a-z0-9_]*" /proc/cpuinfo | sort -u | tr '\n' ' '; echo; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv) > gpurun_out/box.txt 2>&1
timeout 600 bash tools/probe/run_poseidon_lab.sh "$@" > gpurun_out/lab.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
grep -E "===|IFMA|this variant|full round, chained" gpurun_out/poseidon_lab.txt; tail -2 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.log
