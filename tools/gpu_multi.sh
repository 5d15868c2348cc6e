#!/bin/bash
# multi-GPU measurements (run under `gpurun --gpus N`): config 5 (batched instances, sharded by instance), config 3 (n = 2^16, strided
# shards) and the default weak-scaling line.  Usage: tools/gpu_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --workload batch --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_batch_${N}gpu.log 2> gpurun_out/bench_batch_${N}gpu.err
tail -1 gpurun_out/bench_batch_${N}gpu.log | cut -c1-1500; tail -2 gpurun_out/bench_batch_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --pairs 65536 --steps 2 --warmup 3 --no-cpu-baseline --saturated-pairs 0 > gpurun_out/bench_n65536_${N}gpu.log 2> gpurun_out/bench_n65536_${N}gpu.err
tail -1 gpurun_out/bench_n65536_${N}gpu.log | cut -c1-1500; tail -2 gpurun_out/bench_n65536_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-1200; tail -2 gpurun_out/bench_${N}gpu.err
