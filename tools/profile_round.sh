#!/bin/bash
# Profiling recipe of this repo (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md).
#   tools/profile_round.sh <tag>
# Writes into gpurun_out/: launches_<tag>.csv (every launch of a short bench run with its device time),
# scan_<tag>.ncu-rep (ncu --set full of the scan kernel), scan_<tag>_raw.csv (its raw metrics page).
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
export BMAGWA_BENCH_DIR=/tmp/bmagwa_bench
# 1. launch list: one warm-up step + two timed steps of the C2 bench (a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 3000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
# 2. the top kernel, full set, three launches after warm-up
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_dots -s 3 -c 3 -f -o $OUT/scan_$TAG \
    python tools/scan_probe.py 5000 100000 3 > $OUT/scan_probe_under_ncu_$TAG.log 2>&1
ncu -i $OUT/scan_$TAG.ncu-rep --page raw --csv > $OUT/scan_${TAG}_raw.csv 2>/dev/null
ls -la $OUT
