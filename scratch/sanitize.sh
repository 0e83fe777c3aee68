#!/bin/bash
# compute-sanitizer over small scenes: memcheck (global / shared out-of-bounds, misaligned),
# racecheck (shared-memory hazards in the block reductions and the scan), synccheck, initcheck on the solver sweeps;
# memcheck on a 2-rank slab run when 2 GPUs are visible.  Summaries -> gpurun_out/sanitize_*.log
# usage: scratch/sanitize.sh [TAG]
TAG=${1:-r2}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool, args...
  local name=$1 tool=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 "$@" > gpurun_out/sanitize_${TAG}_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${TAG}_${name}_${tool}.log | tail -1)"
}
run solvers memcheck python scratch/t_sanitize.py solvers 2
run solvers racecheck python scratch/t_sanitize.py solvers 2
run solvers synccheck python scratch/t_sanitize.py solvers 1
run rigid memcheck python scratch/t_sanitize.py rigid 2
run rigid racecheck python scratch/t_sanitize.py rigid 2
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  run slab memcheck --target-processes all python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scratch/t_sanitize.py slab 5
fi
