#!/bin/bash
# times the third-law kernel at npl=1e5 for the compiled (CTAs/SM, unroll) variants (development aid)
for cfg in ${@:-38 34 32 28 24 22}; do
  echo "== SWCU_FLAT_CFG=$cfg"; SWCU_FLAT_CFG=$cfg python scripts/kick_bench.py 100000 4 2>&1 | grep -E "^flat"
done
