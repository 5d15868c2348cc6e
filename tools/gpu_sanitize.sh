#!/bin/bash
# compute-sanitizer over small end-to-end runs (memcheck, then racecheck for the shared-memory exchanges)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
tail -4 gpurun_out/sanitize_memcheck.log; tail -4 gpurun_out/sanitize_racecheck.log
