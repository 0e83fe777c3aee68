#!/bin/bash
# weak scaling at 8 M particles per GPU (BASELINE configs[4]); $1 = GPUs
N=$1
time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --n-side 200 --steps 5 --warmup 3 > gpurun_out/mgbig_$N.json 2> gpurun_out/mgbig_$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/mgbig_%s.json'%n) if l.startswith('{')][-1])
    print(n,'x 8M: ms/step %.3f'%d['ms_per_step'],'Mps %.1f'%(d['value']/1e6),'it',d['config']['iterations'],'e2e %.1f'%(d['e2e']['value']/1e6))
except Exception as e:
    print('FAILED',e); print(open('gpurun_out/mgbig_%s.err'%n).read()[-3000:])
PY
tail -4 gpurun_out/mgbig_$N.err
