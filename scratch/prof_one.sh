#!/bin/bash
# one group of scratch/prof_r2.sh: usage scratch/prof_one.sh TAG name regex count solver what warm steps
TAG=$1; shift
REP=/tmp/ncu_$TAG; mkdir -p $REP gpurun_out/profiles_$TAG
ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -c $3 -o $REP/prof_${TAG}_$1 -f \
    python scratch/t_prof_solver.py $4 $5 $6 $7 > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
echo "full $1 rc=$? $(tail -1 gpurun_out/ncu_full_${TAG}_$1.log)"
python scratch/prof_summary_r2.py $TAG $REP gpurun_out/profiles_$TAG
