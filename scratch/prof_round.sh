#!/bin/bash
# Round profile: plain bench (must exit 0), ncu launch list of the same command, ncu --set full of the DFSPH sweeps.
# usage: scratch/prof_round.sh r1d
TAG=${1:-rX}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -20 gpurun_out/bench_$TAG.err; exit 1; }
tail -c 1500 gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_df_|k_build_lists' -s 68 -c 12 -o gpurun_out/prof_$TAG -f \
    python scratch/t_perf1.py 100 2 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
