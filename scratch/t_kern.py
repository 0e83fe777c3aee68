import sys, json, ctypes, contextlib, io, torch
sys.path.insert(0,'.')
from cfd_taichi_b200 import _lib, scenes
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
n_side=int(sys.argv[1]) if len(sys.argv)>1 else 100
cfg=scenes.breaking_dam(n_side)
with contextlib.redirect_stdout(io.StringIO()):
    ps=ParticleSystem(cfg); sol=dfsph_solver(ps,cfg)
for i in range(4): sol.step()
L,h=ps._lib,ps._h
_lib.check(L.sph_profile_begin(h),h)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(6): sol.step()
e1.record(); torch.cuda.synchronize()
nk=len(_lib.KERNEL_CLASSES); ms=(ctypes.c_float*nk)(); cnt=(ctypes.c_int32*nk)()
_lib.check(L.sph_profile_end(h,ms,cnt,nk),h)
st=sol.stats()
print('ms/step %.3f'%(e0.elapsed_time(e1)/6), 'D',st.div_iters,'C',st.den_iters, {_lib.KERNEL_CLASSES[k]: round(ms[k]/cnt[k]*1000,1) for k in range(nk) if cnt[k]>0})
