"""Where project_forces spends its wall time at config 3 (featurised fit): host profile + per-entry kernel times."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat, qp_feat_linear_map
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
from aggforce_b200.util import Curry

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
topo = chignolin_topology()
c, f = synth_trajectory_device(topo, T, seed=2)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
feat = Multifeaturize([id_feat, Curry(gb_feat, inner=0, outer=8, width=1, n_basis=7)])
run = lambda: agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds=topo.xh_constraints,
                                 method=qp_feat_linear_map, featurizer=feat, kbt=0.6955215, l2_regularization=1e3,
                                 constraint_frames=np.arange(20))
run(); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print(f"wall {(time.perf_counter()-t0)*1e3:.1f} ms")
_lib.timing(True); run(); recs = _lib.timing_records(); _lib.timing(False)
agg = {}
for n, ms in recs: agg[n] = agg.get(n, 0) + ms
print("kernels:", ", ".join(f"{n} {ms:.2f}" for n, ms in agg.items()), " sum", sum(agg.values()))
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
