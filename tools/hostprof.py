import cProfile, pstats, sys, time
sys.path.insert(0, '/root/repo')
import torch
import aggforce_b200 as agf
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
coords, forces = synth_trajectory_device(topo, 200_000, seed=1)
def step():
    cons = agf.guess_pairwise_constraints(coords)
    r1 = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, method=agf.constraint_aware_uni_map)
    r2 = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=1e3)
for _ in range(3): step()
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t)*100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
