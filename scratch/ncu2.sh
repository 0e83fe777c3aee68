#!/bin/bash
# ncu --set full on two DFSPH sweeps for the in-tree build and the given variants
for v in main "$@"; do
  if [ "$v" = main ]; then unset SPH_B200_LIB; else export SPH_B200_LIB=$PWD/scratch/variants/$v/libsph_b200.so; fi
  ncu --set full --clock-control none --import-source on -k regex:'k_df_(drho|div_iter)' -s 40 -c 4 -o gpurun_out/ncu_$v -f python scratch/t_perf1.py 100 2 > gpurun_out/ncu_$v.log 2>&1
  tail -2 gpurun_out/ncu_$v.log
done
