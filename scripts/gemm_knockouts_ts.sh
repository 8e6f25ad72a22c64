#!/bin/bash
# knock-out timing of the TS GEMM (results are wrong when a bit is set): 1 no TMEM read-back, 2 no split (zeros),
# 4 one MMA term, 8 no write-out, 16 no tcgen05.st
for dbg in 0 1 2 4 8 16 31; do
  echo "== SEGGER_B200_TC_DBG=$dbg"
  SEGGER_B200_TC_DBG=$dbg python scripts/bench_gemm.py --reps 3 2>&1 | grep "fwd exact=0\|dgrad" | grep -v "N=  64 K=  64"
done
