#!/bin/bash
# multi-GPU bench: $1 = number of GPUs; runs the peer-window transport and the NCCL transport
N=$1
mkdir -p gpurun_out
for tr in ${TRANSPORTS:-p2p nccl}; do
  SPH_MG_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_${N}_$tr.json 2> gpurun_out/mg_${N}_$tr.err
  python - $N $tr <<'PY'
import json,sys
n,tr=sys.argv[1:]
try:
    d=json.loads([l for l in open('gpurun_out/mg_%s_%s.json'%(n,tr)) if l.startswith('{')][-1])
    print(n,tr,'ms/step %.3f'%d['ms_per_step'],'Mps %.1f'%(d['value']/1e6),'it',d['config']['iterations'],'e2e %.1f'%(d['e2e']['value']/1e6))
    k=d['kernel_ms']; print('   per step us:',' '.join('%s=%.0f(%d)'%(c,1e3*k[c]['ms']/d['steps'],k[c]['launches']/d['steps']) for c in k), 'sum=%.0f'%(sum(1e3*v['ms'] for v in k.values())/d['steps']))
except Exception as e:
    print(n,tr,'FAILED',e); print(open('gpurun_out/mg_%s_%s.err'%(n,tr)).read()[-3000:])
PY
done
