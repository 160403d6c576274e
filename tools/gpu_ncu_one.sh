#!/bin/bash
# Usage: bash tools/gpu_ncu_one.sh <tag> <prof_kernels-case> <kernel-regex> <skip> <count>
TAG=$1; K=$2; PAT=$3; SKIP=$4; CNT=$5; O=gpurun_out; mkdir -p $O
timeout 300 python tools/prof_kernels.py $K > $O/plain_$K.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "$PAT" -s $SKIP -c $CNT -f -o $O/${TAG}_prof_$K python tools/prof_kernels.py $K > $O/ncu_$K.log 2>&1
echo "ncu $K rc=$?"
ncu -i $O/${TAG}_prof_$K.ncu-rep --page raw --csv > $O/${TAG}_prof_${K}_raw.csv 2>/dev/null
ncu -i $O/${TAG}_prof_$K.ncu-rep --page source --csv > $O/${TAG}_prof_${K}_source.csv 2>/dev/null
sz=$(stat -c %s $O/${TAG}_prof_$K.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 12000000 ]; then rm -f $O/${TAG}_prof_$K.ncu-rep; fi
