#!/bin/bash
# round 2, pairing-matrix tail: targeted parity tests, A/B timing, launch list of one n = 2^12 prove
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "matrix_tail or prove_golden or n64 or n128 or n4096 or round_api" > gpurun_out/pytest_tail.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tail.log
timeout 300 python tools/tail_ab.py 4096 5 > gpurun_out/tail_ab.txt 2>&1
timeout 300 python tools/tail_ab.py 128 5 >> gpurun_out/tail_ab.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_tail_launches.csv python tools/prove_once.py 4096 2 > gpurun_out/ncu_tail.log 2>&1
tail -15 gpurun_out/pytest_tail.log; cat gpurun_out/tail_ab.txt
