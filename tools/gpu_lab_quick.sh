#!/bin/bash
# whole-permutation timings of every Poseidon variant on the box's host CPU (no GPU work)
mkdir -p gpurun_out
LAB_QUICK=1 timeout 600 bash tools/probe/run_poseidon_lab.sh "$@" > gpurun_out/lab.log 2>&1
grep -E "===|IFMA partial|this variant|error" gpurun_out/poseidon_lab.txt
