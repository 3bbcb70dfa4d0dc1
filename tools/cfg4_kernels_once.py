"""One tiled Gram and one int8 application at the config-4 shape (for ncu captures).  usage: [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
topo = protein_like_topology(500)
_, forces = synth_trajectory_device(topo, T, seed=3)
cols = reduced_columns(topo.n_sites, topo.xh_constraints)
n_red = int(cols.max()) + 1
lm = agf.LinearMap(np.random.default_rng(0).normal(size=(500, n_red))[:, cols])
for _ in range(2):
    g = _engine.gram_linear(_engine.Frames(forces), cols, n_red)
    o = lm(forces)
torch.cuda.synchronize()
print("done", float(g[0, 0]), float(o[0, 0, 0]))
