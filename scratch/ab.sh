#!/bin/bash
# A/B the variant builds under scratch/variants/* against the in-tree build (developer aid)
mkdir -p gpurun_out
for v in main "$@"; do
  if [ "$v" = main ]; then unset SPH_B200_LIB; else export SPH_B200_LIB=$PWD/scratch/variants/$v/libsph_b200.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open('gpurun_out/ab_%s.json'%v))
    k=d['kernel_ms']
    print(v, 'ms/step %.3f'%d['ms_per_step'], 'Mps %.1f'%(d['value']/1e6), 'it',d['config']['iterations'], ' '.join('%s=%.1f'%(n,1e3*k[n]['ms']/k[n]['launches']) for n in k), 'clk',d['clocks']['sm_mhz'])
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/ab_%s.err'%v).read()[-2000:])
PY
done
