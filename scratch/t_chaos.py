import sys, numpy as np, torch, contextlib, io
sys.path.insert(0,'.')
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200 import scenes
name=sys.argv[1] if len(sys.argv)>1 else 'breaking_dam_30k'
nsteps=int(sys.argv[2]) if len(sys.argv)>2 else 1000
cfg=scenes.shipped(name,'dfsph')
sims=[]
for tag,strict,pert in (('strict',True,False),('strict+1ulp',True,True),('fast',False,False)):
    with contextlib.redirect_stdout(io.StringIO()):
        ps=ParticleSystem(cfg, strict=strict); sol=dfsph_solver(ps,cfg)
    if pert:
        x=ps._pos4[1234,0].item(); ps._pos4[1234,0]=float(np.nextafter(np.float32(x),np.float32(10)))
    sims.append((tag,ps,sol,[],[0.0]))
def stat(ps,sol):
    v=ps._vel4[:ps.particle_num,:3].double(); ke=0.5*0.125*(v*v).sum().item()
    y=ps._pos4[:ps.particle_num,1].double().mean().item()
    x=ps._pos4[:ps.particle_num,0].double().mean().item()
    return ke,y,x
for step in range(1,nsteps+1):
    for tag,ps,sol,hist,t in sims:
        sol.step()
        if step%10==0:
            dt=sol.stats().delta_time; 
        else: dt=None
        hist.append((step,)+stat(ps,sol) if step%10==0 else None)
        if step%10==0: t.append(dt)
# equal-step comparison
print('equal-step comparison (KE, mean y, mean x) relative to strict')
for k in range(9,nsteps,100):
    a=sims[0][3][k]
    print(a[0], ' '.join('%s: KE %+.2f%% y %+.3f%% x %+.3f%%'%(s[0],100*(s[3][k][1]-a[1])/a[1],100*(s[3][k][2]-a[2])/a[2],100*(s[3][k][3]-a[3])/a[3]) for s in sims[1:]))
