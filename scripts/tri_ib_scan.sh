#!/bin/bash
# full-row kernel: rows per thread (SWCU_KICK_IB) vs npl (development aid)
for ib in 1 2 4; do echo "== SWCU_KICK_IB=$ib"; SWCU_KICK_IB=$ib python scripts/crossover_scan.py 4096 10000 20000 40000 70000 100000 | grep "^| [0-9]" | awk -F'|' '{print $2, "tri us:", $4}'; done
