#!/bin/bash
# tuning experiment: run tools/exp_throughput.py for the default library and every pingoo_b200/libpgw_*.so variant
cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"
for lib in pingoo_b200/libpgw_*.so; do [ -e "$lib" ] || continue; echo "== $lib"; PGW_LIB=$PWD/$lib timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"; done
