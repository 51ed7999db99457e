#!/bin/bash
# One GPU call: the -m gpu suite, the default bench line, and the ncu launch list of config 3 (4 M requests, two batches).
mkdir -p gpurun_out
(time timeout 480 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/pytest_gpu.txt
(time timeout 420 python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -4 gpurun_out/bench_default.err
cut -c1-700 gpurun_out/bench_default.json
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg3_4M.csv \
    python tools/prof_config.py 3 4000000 2 > gpurun_out/launches.log 2>&1
tail -3 gpurun_out/launches.log
grep -c waf_ gpurun_out/launches_cfg3_4M.csv
