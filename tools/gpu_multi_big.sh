#!/bin/bash
# BASELINE configs 3 and 4 on N GPUs (run under `gpurun --gpus N`): n = 2^16 and n = 2^20, strided shards; bench.py asserts that the
# sharded proof is byte-identical to the single-GPU proof of the same inputs.  Usage: tools/gpu_multi_big.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus $N --pairs 65536 --steps 2 --quick --no-cpu-baseline --saturated-pairs 0 > gpurun_out/bench_n65536_${N}gpu.log 2> gpurun_out/bench_n65536_${N}gpu.err
tail -1 gpurun_out/bench_n65536_${N}gpu.log | cut -c1-1800; tail -2 gpurun_out/bench_n65536_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --pairs 1048576 --steps 1 --quick --no-cpu-baseline --saturated-pairs 0 > gpurun_out/bench_n1048576_${N}gpu.log 2> gpurun_out/bench_n1048576_${N}gpu.err
tail -1 gpurun_out/bench_n1048576_${N}gpu.log | cut -c1-1800; tail -2 gpurun_out/bench_n1048576_${N}gpu.err
