#!/bin/bash
# usage: bash tools/run_debug.sh case1 case2 ...   (each case in its own process, 120 s timeout)
for c in "$@"; do
  echo "=== $c"
  timeout 120 python tools/debug_gemm.py $c 2>&1 | tail -40
  echo "--- rc=$?"
done
