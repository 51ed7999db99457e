#!/bin/bash
# tuning experiment: field path with class rows vs direct rows at a few image caps
cd "$(dirname "$0")/.."
echo "== class rows"; PGW_DIRECT_ROWS=0 timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"
for kb in 128 176 200; do echo "== direct rows, image cap $kb KB"; PGW_SMEM_IMAGE_KB=$kb timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"; done
