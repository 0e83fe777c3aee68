#!/bin/bash
# cost of the per-launch profiling events: bench with and without them, 1 GPU and N GPUs
N=${1:-1}
for np in 0 1; do
  if [ $np = 1 ]; then export SPH_BENCH_NO_PROFILE=1; else unset SPH_BENCH_NO_PROFILE; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 gpu noprofile=$np ms/step %.3f Mps %.1f'%(d['ms_per_step'],d['value']/1e6))"
  if [ $N -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('$N gpu noprofile=$np ms/step %.3f Mps %.1f'%(d['ms_per_step'],d['value']/1e6))"
  fi
done
