#!/bin/bash
# ncu --set full of the launches matching a kernel regex in eager steps; CSV pages come back in gpurun_out
# usage: bash tools/gpu_ncu_kernel.sh <tag> <kernel regex> <skip> <count> [script args]
TAG=$1; RE=$2; SKIP=${3:-0}; CNT=${4:-2}
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o /tmp/${TAG} python tools/one_step.py 3 > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${TAG}_ncu.log
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page details --csv > $O/${TAG}_details.csv 2>/dev/null
ls -la $O/${TAG}_*
