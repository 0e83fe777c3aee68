import sys, numpy as np, torch, contextlib, io
sys.path.insert(0,'.')
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200 import scenes
name=sys.argv[1] if len(sys.argv)>1 else 'breaking_dam_30k'
nsteps=int(sys.argv[2]) if len(sys.argv)>2 else 1000
cfg=scenes.shipped(name,'dfsph')
runs={}
for tag,strict,pert in (('strict',True,False),('strict+1ulp',True,True),('fast',False,False)):
    with contextlib.redirect_stdout(io.StringIO()):
        ps=ParticleSystem(cfg, strict=strict); sol=dfsph_solver(ps,cfg)
    if pert:
        x=ps._pos4[1234,0].item(); ps._pos4[1234,0]=float(np.nextafter(np.float32(x),np.float32(10)))
    t=0.0; rows=[]
    for step in range(nsteps):
        sol.step(); st=sol.stats(); t+=st.delta_time
        v=ps._vel4[:ps.particle_num,:3].double(); ke=0.5*0.125*(v*v).sum().item()
        y=ps._pos4[:ps.particle_num,1].double().mean().item(); x=ps._pos4[:ps.particle_num,0].double().mean().item()
        rho=sol.rho.to_torch().double(); 
        rows.append((t,ke,y,x,rho.mean().item(),torch.clamp(rho-1000,min=0).mean().item()))
    runs[tag]=np.array(rows); ps.close()
T=min(r[-1,0] for r in runs.values())
print('end times', {k:float(v[-1,0]) for k,v in runs.items()})
a=runs['strict']
for frac in (0.1,0.25,0.5,0.75,1.0):
    tt=T*frac
    ref=[np.interp(tt,a[:,0],a[:,c]) for c in range(1,6)]
    for k in ('strict+1ulp','fast'):
        b=runs[k]; val=[np.interp(tt,b[:,0],b[:,c]) for c in range(1,6)]
        print('t=%.4f %-12s KE %+.3f%% y %+.3f%% x %+.3f%% rho %+.4f%% derr %.4g vs %.4g'%(tt,k,100*(val[0]-ref[0])/ref[0],100*(val[1]-ref[1])/ref[1],100*(val[2]-ref[2])/ref[2],100*(val[3]-ref[3])/ref[3],val[4],ref[4]))
