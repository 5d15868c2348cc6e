#!/bin/bash
# One `ncu --set full` capture of the dominant kernels of (a) the saturated Miller leg (2^17 pairs: k_lines + k_accum) and (b) the
# headline prove (n = 2^12: the first, large launches and the latency-regime kernels of the late rounds).  Exports the raw page to
# CSV on the box and condenses it into gpurun_out/ncu_traffic.json (copy to profiles/ncu_traffic.json: bench.py reads it for
# `roofline.traffic`) and gpurun_out/<tag>_ncu_full_*.csv (copy to profiles/).  1 GPU only.  Usage: tools/ncu_traffic.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lines$|k_accum$' -c 4 -f -o /tmp/prof_sat python tools/sat_miller.py 131072 2 > gpurun_out/${TAG}_ncu_sat.log 2>&1
ncu -i /tmp/prof_sat.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_saturated_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'k_lines|k_accum|k_reduce_fe_eng|k_fold_wide|k_fold_split|k_validate_points|k_mat_|k_qlines|k_eval_lines' -c 70 -f -o /tmp/prof_head python tools/prove_once.py 4096 2 > gpurun_out/${TAG}_ncu_head.log 2>&1
ncu -i /tmp/prof_head.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_headline_raw.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/${TAG}_ncu_full_saturated_raw.csv gpurun_out/${TAG}_ncu_full_headline_raw.csv gpurun_out/ncu_traffic.json gpurun_out/${TAG}_ncu_full_summary.csv
tail -2 gpurun_out/${TAG}_ncu_sat.log gpurun_out/${TAG}_ncu_head.log
cat gpurun_out/ncu_traffic.json
