#!/bin/bash
# 1-GPU legs of BASELINE configs 4 and 5 plus a 10 M-row launch list.
#   gpurun --timeout 1500 -- 'bash tools/gpu_ns3_single.sh <tag>'
TAG=${1:-ns3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 500 python tools/ns3_measure.py encode > $OUT/encode_1gpu.log 2>&1
echo "encode rc=$?"; grep "^encode" $OUT/encode_1gpu.log
timeout 600 python tools/ns3_measure.py deep > $OUT/deep_50m.log 2>&1
echo "deep rc=$?"; tail -c 1500 $OUT/deep_50m.log
PROF_ROWS=10000000 PROF_DATA=mix timeout 300 python tools/profile_iter.py > $OUT/profile_iter_10m.log 2>&1
echo "10m rc=$?"; tail -3 $OUT/profile_iter_10m.log
PROF_ROWS=10000000 PROF_DATA=mix timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/launches_10m.csv python tools/profile_iter.py > $OUT/ncu_10m.log 2>&1
echo "10m launch list rc=$?"
