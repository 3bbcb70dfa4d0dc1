"""Where project_forces spends its wall time at the config-4 shape (host profile + per-entry kernel times)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
topo = protein_like_topology(500)
c, f = synth_trajectory_device(topo, T, seed=3)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
run = lambda: agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds="auto", l2_regularization=1e3)
run(); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print(f"wall {(time.perf_counter()-t0)*1e3:.1f} ms")
_lib.timing(True); run(); recs = _lib.timing_records(); _lib.timing(False)
print("kernels:", ", ".join(f"{n} {ms:.2f}" for n, ms in recs), " sum", sum(ms for _, ms in recs))
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
