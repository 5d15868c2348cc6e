#!/bin/bash
# ncu captures, round 1 v6: (1) launch list of the default bench prove (n = 2^12), (2) launch list of a 4096 x 128 batch,
# (3) --set full of the batch kernels (round 1 of a 1024-instance batch), exported to CSV on the box (the report itself
# exceeds what gpurun copies back).  1 GPU only.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/v6_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --saturated-pairs 0 --batch-instances 0 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/v6_batch4096_launches.csv python tools/batch_bench.py 4096 128 1 > gpurun_out/ncu_batch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_lines_batch|k_accum|k_fe_batch|k_fold_straus|k_tr_round|k_tr_absorb' -c 9 -f -o /tmp/prof_batch python tools/batch_bench.py 1024 128 1 > gpurun_out/ncu_full_batch.log 2>&1
ncu -i /tmp/prof_batch.ncu-rep --page raw --csv > gpurun_out/v6_ncu_full_batch_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_full_batch.log
ls -la gpurun_out/ | tail -8
