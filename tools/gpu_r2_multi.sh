#!/bin/bash
# multi-GPU round trip (run with gpurun --gpus N): NCCL tests of the library's sharded prover, then the bench at N ranks.
# Usage: tools/gpu_r2_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout 600 -k "nccl" > gpurun_out/pytest_nccl_${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl_${N}.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/bench_${N}gpu.err
tail -8 gpurun_out/pytest_nccl_${N}.log; cat gpurun_out/bench_${N}gpu.log; tail -8 gpurun_out/bench_${N}gpu.err
