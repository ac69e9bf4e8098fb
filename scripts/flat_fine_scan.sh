#!/bin/bash
# kernel time of a 1/8 share of the npl=1e5 triangle for different graded-schedule tails (development aid)
for fine in "2,4,8,16" "1,2,4,8" "1,2,4,4" "1,2,2,4" "2,2,4,4" "1,1,2,4" "0,2,4,8" "1,2,4,0" "2,4,8,0" "1,2,0,0" "0,0,0,0"; do
  for q in 0 2 4; do
  echo -n "emulate 8 ranks, fine=$fine nsplit=$q: "; SWCU_KICK_NSPLIT=$q SWCU_FLAT_EMULATE_RANKS=8 SWCU_FLAT_FINE=$fine python scripts/kick_bench.py 100000 6 2>&1 | grep -E "^flat lclose=1" | awk '{print $3, $4}'
  done
done
