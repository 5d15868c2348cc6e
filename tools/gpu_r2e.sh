#!/bin/bash
# round 2 profiles: ncu --set full captures (traffic + summary), launch list of the default bench line
mkdir -p gpurun_out
bash tools/ncu_traffic.sh r02 > gpurun_out/ncu_traffic.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --saturated-pairs 0 --batch-instances 0 --large '' > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_traffic.log; cat gpurun_out/ncu_traffic.json; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
