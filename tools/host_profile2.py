import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo')
import torch
import aggforce_b200 as agf
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
coords, forces = synth_trajectory_device(topo, 200_000, seed=1)
cons = agf.guess_pairwise_constraints(coords)
calls = {
    "guess": lambda: agf.guess_pairwise_constraints(coords),
    "uni": lambda: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, method=agf.constraint_aware_uni_map),
    "opt": lambda: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=1e3),
}
for name, f in calls.items():
    for _ in range(5): f()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): f()
    pr.disable()
    out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(22)
    print("=====", name); print("\n".join(l for l in out.getvalue().splitlines()[4:] if l.strip())[:3800])
