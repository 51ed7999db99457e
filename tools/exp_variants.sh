#!/bin/bash
# tuning experiment: run tools/exp_throughput.py for the default library (each kernel path) and every
# pingoo_b200/libpgw_*.so variant present (PGW_VARIANT_KERNEL selects the path used for the variants)
cd "$(dirname "$0")/.."
for k in lane field; do echo "== default PGW_KERNEL=$k"; PGW_KERNEL=$k timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"; done
for lib in pingoo_b200/libpgw_*.so; do [ -e "$lib" ] || continue; echo "== $lib"; PGW_KERNEL=${PGW_VARIANT_KERNEL:-field} PGW_LIB=$PWD/$lib timeout 300 python tools/exp_throughput.py 2>&1 | grep -E " ms |rror"; done
