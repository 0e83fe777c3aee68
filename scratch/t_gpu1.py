import json, sys, time, numpy as np, torch
sys.path.insert(0,'.')
from oracle import oracle as O
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200 import _lib
from cfd_taichi_b200 import scenes; cfg=scenes.shipped('small_block','dfsph')
for strict in (True, False):
    ps=ParticleSystem(cfg, strict=strict); sol=dfsph_solver(ps,cfg)
    o=O.Oracle(cfg, threads=8)
    print('bvol equal', np.array_equal(ps.boundary_particles.volume.to_numpy(), o.field('bvol')), np.abs(ps.boundary_particles.volume.to_numpy()-o.field('bvol')).max())
    print('cell1 equal', np.array_equal(ps.cell_indices_1d().cpu().numpy(), o.field('cell1')))
    print('cell_start equal', np.array_equal(ps.cell_start().cpu().numpy(), o.field('cell_start')), 'items', np.array_equal(ps.sorted_index().cpu().numpy(), o.field('cell_items')))
    for step in range(5):
        sol.step(); o.step()
        s=sol.stats()
        def cmp(name,a,b):
            d=np.abs(a-b).max(); sc=np.abs(b).max()
            print('  %s maxabs %.3e rel %.3e equal %s'%(name,d,d/(sc+1e-30),np.array_equal(a,b)))
        print('step',step,'strict',strict,'gpu div',s.div_iters,s.div_first_err,s.div_err,'den',s.den_iters,s.den_err,'dt',s.delta_time,'flags',s.error_flags,'maxn',s.max_neighbors_seen)
        print('          oracle div',o.scalar('df_div_iters'),o.scalar('df_div_first_err'),o.scalar('df_div_err'),'den',o.scalar('df_den_iters'),o.scalar('df_den_err'),'dt',o.scalar('delta_time'))
        cmp('pos',ps.fluid_particles.pos.to_numpy(),o.field('pos')); cmp('vel',ps.fluid_particles.vel.to_numpy(),o.field('vel'))
        cmp('rho',sol.rho.to_numpy(),o.field('rho')); cmp('alpha',sol.alpha.to_numpy(),o.field('alpha'))
        cmp('k',sol.warm_start_k.to_numpy(),o.field('warm_start_k'))
        print('  nbr equal', np.array_equal(ps.neighbour_counts().cpu().numpy(), o.field('nbr_count')))
    ps.close()
