#!/bin/bash
# builds tools/probe/poseidon_lab.cc once per variant of the AVX-512 / IFMA permutation (the shipped file plus tools/probe/variants/*.cc,
# which tools/probe/make_ifma_variants.py generates) and runs them on this machine (the GPU box's host CPU is the one that counts);
# output -> gpurun_out/poseidon_lab.txt
# Usage: tools/probe/run_poseidon_lab.sh [variant files...]   (default: the shipped file + every variant)
mkdir -p gpurun_out
OUT=gpurun_out/poseidon_lab.txt
python tools/probe/make_ifma_variants.py > /dev/null
g++ -O3 -march=x86-64-v3 -std=c++17 -fPIC -c -o /tmp/transcript_lab.o sipp_b200/csrc/transcript.cc || exit 1
(lscpu | grep -E "Model name|^CPU\(s\)"; grep -o -E "avx512[a-z0-9_]*" /proc/cpuinfo | sort -u | tr '\n' ' '; echo) > $OUT
LIST="$@"
[ -z "$LIST" ] && LIST="sipp_b200/csrc/poseidon_avx512.cc tools/probe/variants/*.cc"
for v in $LIST; do
  [ -f "$v" ] || continue
  echo "=== $v" >> $OUT
  if g++ -O3 -march=x86-64-v3 -std=c++17 -Isipp_b200/csrc -DPOS_IMPL="\"$PWD/$v\"" -o /tmp/poseidon_lab_v tools/probe/poseidon_lab.cc /tmp/transcript_lab.o 2>> $OUT; then
    for rep in 1 2; do taskset -c 2 /tmp/poseidon_lab_v | grep -v "^  \[" >> $OUT; done
  fi
done
cat $OUT
