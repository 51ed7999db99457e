#!/bin/bash
# tuning experiment: kernel variants (threads x chunk) on the main workloads
for lib in libpgw_t1024_c32 libpgw_t768_c32 libpgw_t512_c32 libpgw_t1024_c16 libpgw_t768_c16; do
  echo "== $lib"
  PGW_LIB=$PWD/pingoo_b200/$lib.so timeout 300 python exp1.py 2>&1 | grep -E "ms|Error|error" 
done
