#!/bin/bash
# ncu captures: (1) launch list of one bench prove, (2) --set full of the saturated Miller kernels.  1 GPU only.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --saturated-pairs 0 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lines|k_accum' -s 2 -c 2 -f -o gpurun_out/prof_miller python tools/sat_miller.py 131072 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
