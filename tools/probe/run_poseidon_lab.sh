#!/bin/bash
# builds and runs tools/probe/poseidon_lab.cc on this machine (the GPU box's host CPU is the one that counts); output -> gpurun_out/
mkdir -p gpurun_out
g++ -O3 -march=x86-64-v3 -std=c++17 -fPIC -c -o /tmp/transcript_lab.o sipp_b200/csrc/transcript.cc && \
g++ -O3 -march=x86-64-v3 -std=c++17 -o /tmp/poseidon_lab tools/probe/poseidon_lab.cc /tmp/transcript_lab.o && \
(lscpu | grep -E "Model name|^CPU\(s\)|MHz|L2|L3"; grep -o -E "avx512[a-z0-9_]*" /proc/cpuinfo | sort -u | tr '\n' ' '; echo; taskset -c 2 /tmp/poseidon_lab) > gpurun_out/poseidon_lab.txt 2>&1
cat gpurun_out/poseidon_lab.txt
