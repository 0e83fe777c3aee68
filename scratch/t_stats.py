import sys, numpy as np, torch, contextlib, io
sys.path.insert(0,'.')
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200 import scenes
name=sys.argv[1] if len(sys.argv)>1 else 'breaking_dam_30k'
nsteps=int(sys.argv[2]) if len(sys.argv)>2 else 1000
cfg=scenes.shipped(name,'dfsph')
sims=[]
for strict in (True, False):
    with contextlib.redirect_stdout(io.StringIO()):
        ps=ParticleSystem(cfg, strict=strict); sol=dfsph_solver(ps,cfg)
    sims.append((ps,sol))
def stat(ps,sol):
    v=ps._vel4[:ps.particle_num,:3].double(); ke=0.5*0.125*(v*v).sum().item()
    rho=sol.rho.to_torch().double(); err=torch.clamp(rho-1000,min=0).mean().item()
    y=ps._pos4[:ps.particle_num,1].double().mean().item()
    return ke, rho.mean().item(), err, y
for step in range(1,nsteps+1):
    for ps,sol in sims: sol.step()
    if step%50==0 or step in (1,2,5,10,20):
        a=stat(*sims[0]); b=stat(*sims[1]); s0=sims[0][1].stats(); s1=sims[1][1].stats()
        print(step, 'KE %.5g %.5g (%.3f%%)'%(a[0],b[0],100*(b[0]-a[0])/a[0]), 'rho %.4f %.4f'%(a[1],b[1]), 'err %.4g %.4g'%(a[2],b[2]), 'y %.5f %.5f'%(a[3],b[3]), 'it',s0.div_iters,s0.den_iters,s1.div_iters,s1.den_iters,'dt %.3g %.3g'%(s0.delta_time,s1.delta_time),'maxn',s0.max_neighbors_seen, flush=True)
