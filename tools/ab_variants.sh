#!/bin/bash
# A/B: saturated Miller leg (2^17 pairs) and a 2^12 prove for each library variant under sipp_b200/variants/
for v in sipp_b200/variants/*.so; do
  echo "== $v"
  SIPP_LIB=$PWD/$v timeout 300 python tools/sat_miller.py 131072 3 2>&1 | tail -2
done
