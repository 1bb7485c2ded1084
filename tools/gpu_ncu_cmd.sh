#!/bin/bash
# ncu --set full of kernels matching a regex in an arbitrary command; usage: bash tools/gpu_ncu_cmd.sh <tag> <regex> <skip> <count> <cmd...>
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift; shift; shift; shift
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o /tmp/${TAG} "$@" > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${TAG}_ncu.log
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ls -la $O/${TAG}_*
