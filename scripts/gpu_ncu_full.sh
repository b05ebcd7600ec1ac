#!/bin/bash
# usage: gpu_ncu_full.sh <tag> <kernel regex> <skip> <count> <cmd...>
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -o gpurun_out/prof_$TAG "$@" > gpurun_out/ncufull_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log | tail -12; ls -la gpurun_out/prof_$TAG.ncu-rep
