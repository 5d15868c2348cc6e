#!/bin/bash
# saturated Miller leg only: timing without the profiler, then one ncu --set full capture of k_lines / k_accum (DRAM traffic)
mkdir -p gpurun_out
python tools/sat_miller.py 131072 3 > gpurun_out/sat_miller.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_lines$|k_accum$' -c 2 -f -o /tmp/prof_sat python tools/sat_miller.py 131072 1 > gpurun_out/ncu_sat_only.log 2>&1
ncu -i /tmp/prof_sat.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); n=rows[h]; u=rows[h+1]
for r in rows[h+2:]:
    d=dict(zip(n,r)); print(d['Kernel Name'].split('(')[0], 'read', d['dram__bytes_read.sum'], u[n.index('dram__bytes_read.sum')], 'write', d['dram__bytes_write.sum'], u[n.index('dram__bytes_write.sum')], 'time', d['gpu__time_duration.sum'], u[n.index('gpu__time_duration.sum')])
" > gpurun_out/ncu_sat_only.txt
cat gpurun_out/sat_miller.txt gpurun_out/ncu_sat_only.txt
