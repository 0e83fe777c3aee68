import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from cfd_taichi_b200 import selfcheck
rank = int(os.environ['RANK']); torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
for steps in (6, 10, 25):
    res = selfcheck.slab_vs_single(solver='dfsph', steps=steps, strict=False)
    if rank == 0:
        print('steps', steps, 'exact', res['slab_vs_single_bit_exact'], 'iters_ok', res['iters_ok'], 'dpos %.3e dvel %.3e' % (res['max_abs_dpos'], res['max_abs_dvel']), flush=True)
dist.destroy_process_group()
