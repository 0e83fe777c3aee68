#!/bin/bash
# ncu --set full of named kernels of one solver run (developer aid)
# usage: scratch/ncu_kernels.sh TAG 'regex' skip count [t_perf1 args...]
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/ncu_$TAG -f python scratch/t_perf1.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu $TAG rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
