#!/bin/bash
# tp kernels at ntp = 1e6 for the compiled CTAs-per-SM variants of the drift family (development aid)
for m in ${MINBS:-10 8 6}; do echo "== SWCU_DRIFT_MINB=$m"; SWCU_DRIFT_MINB=$m python scripts/tp_bench.py ${1:-1000000} 6 | grep -v "pl->tp"; done
