import sys, time, numpy as np, torch
sys.path.insert(0,'.')
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200 import scenes
n_side=int(sys.argv[1]) if len(sys.argv)>1 else 100
steps=int(sys.argv[2]) if len(sys.argv)>2 else 20
strict=len(sys.argv)>3 and sys.argv[3]=='strict'
cfg=scenes.breaking_dam(n_side)
t0=time.time(); ps=ParticleSystem(cfg, strict=strict); sol=dfsph_solver(ps,cfg); torch.cuda.synchronize(); print('init s',time.time()-t0)
for i in range(5): sol.step()
torch.cuda.synchronize()
s=sol.stats(); print('warm', s.div_iters, s.den_iters, s.delta_time, s.error_flags, s.max_neighbors_seen, s.max_boundary_neighbors_seen)
ev0=torch.cuda.Event(enable_timing=True); ev1=torch.cuda.Event(enable_timing=True)
ev0.record()
its=[]
for i in range(steps):
    sol.step()
ev1.record(); torch.cuda.synchronize()
ms=ev0.elapsed_time(ev1)/steps
s=sol.stats()
print('N',ps.particle_num,'ms/step',ms,'Mps/s',ps.particle_num/ms/1e3,'div',s.div_iters,'den',s.den_iters,'dt',s.delta_time,'flags',s.error_flags,'maxn',s.max_neighbors_seen, s.max_boundary_neighbors_seen, 'launches', s.kernel_launches)
