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
# 1. launch list: one warm-up step + two timed steps of the C2 bench (a number printed under ncu is never a bench value).
#    ncu serialises kernels, which a persistent kernel does not survive, so the per-move column statistics run in their
#    launch-per-move form here (BMG_COLSTATS_SERVER=0: same work items, one k_column_stats_inline launch per move).
BMG_COLSTATS_SERVER=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 3000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > $OUT/bench_under_ncu_$TAG.log 2>&1
# 2. the scan kernel (default variant: integer tensor cores), full set, one launch after the probe's warm-up,
#    at the bench size (C2) and at a size that does not fit in L2 ten times over
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_dots_imma -s 3 -c 1 -f -o $OUT/scan_$TAG \
    python tools/scan_probe.py 5000 100000 1 > $OUT/scan_probe_under_ncu_$TAG.log 2>&1
ncu -i $OUT/scan_$TAG.ncu-rep --page raw --csv > $OUT/scan_${TAG}_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_dots_imma -s 3 -c 1 -f -o $OUT/scan1m_$TAG \
    python tools/scan_probe.py 5000 1000000 1 >> $OUT/scan_probe_under_ncu_$TAG.log 2>&1
ncu -i $OUT/scan1m_$TAG.ncu-rep --page raw --csv > $OUT/scan1m_${TAG}_raw.csv 2>/dev/null
# 3. the per-move column-statistics work item, captured in its launch-per-move form (BMG_COLSTATS_SERVER=0): the
#    persistent server runs the same device function but never ends, so ncu cannot time it per request
BMG_COLSTATS_SERVER=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_column_stats_inline -s 200 -c 1 -f -o $OUT/colstats_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub > $OUT/colstats_under_ncu_$TAG.log 2>&1
ncu -i $OUT/colstats_$TAG.ncu-rep --page raw --csv > $OUT/colstats_${TAG}_raw.csv 2>/dev/null
ls -la $OUT
# 4. the two-residual scan of shard groups (k_scan_dots_imma2), on one GPU: a one-chain group scans its residual as both
#    halves of a pair (BMG_GROUP_SELF_PAIR, development switch of group.cu)
BMG_GROUP_SELF_PAIR=1 BMG_COLSTATS_SERVER=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_scan_dots_imma2 -s 2 -c 1 -f -o $OUT/scan2_$TAG \
    python bench.py --workload C4s --steps 3 --warmup 1 --burnin 0 --no-e2e --no-cpu-baseline > $OUT/scan2_under_ncu_$TAG.log 2>&1
ncu -i $OUT/scan2_$TAG.ncu-rep --page raw --csv > $OUT/scan2_${TAG}_raw.csv 2>/dev/null
