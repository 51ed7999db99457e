#!/bin/bash
# tuning experiment: run tools/exp_throughput.py against every pingoo_b200/libpgw_*.so variant present
cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms "
for lib in pingoo_b200/libpgw_*.so; do echo "== $lib"; PGW_LIB=$PWD/$lib timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms "; done
