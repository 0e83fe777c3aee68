#!/bin/bash
# Round-2 profile set: plain bench (must exit 0), ncu launch list of the same command, ncu --set full of one warm step
# of every solver family at 10^6 particles (profiling starts after the warm-up: cudaProfilerStart in t_prof_solver.py).
# usage: scratch/prof_r2.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-also > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -20 gpurun_out/bench_$TAG.err; exit 1; }
tail -c 600 gpurun_out/bench_$TAG.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
# the .ncu-rep files stay on the box (gpurun merges at most 64 MiB back): the summaries are written there
REP=/tmp/ncu_$TAG
mkdir -p $REP
full() { # name regex count solver what warm steps
  ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -c $3 -o $REP/prof_${TAG}_$1 -f \
      python scratch/t_prof_solver.py $4 $5 $6 $7 > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
  echo "full $1 rc=$? $(tail -1 gpurun_out/ncu_full_${TAG}_$1.log)"
}
full dfsph 'k_df_|k_build_lists' 95 dfsph 100 6 1
full grid 'k_hash|k_scan|k_scatter|k_cell_fix|k_gather|k_reorder' 16 dfsph 100 6 1
full pcisph 'k_pc_' 40 pcisph 100 400 1
full iisph 'k_ii_' 40 iisph 100 250 1
full wcsph 'k_wc_|k_build_lists' 8 wcsph 100 20 1
full rigid 'k_rigid' 30 dfsph dam_flush_cube 20 1
mkdir -p gpurun_out/profiles_$TAG
python scratch/prof_summary_r2.py $TAG $REP gpurun_out/profiles_$TAG
ls -la gpurun_out/profiles_$TAG
