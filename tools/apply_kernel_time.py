import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
_, forces = synth_trajectory_device(topo, 1_000_000, seed=1, want_coords=False)
cols = reduced_columns(175, topo.xh_constraints)
lm = agf.LinearMap(np.random.default_rng(0).normal(size=(10, 97))[:, cols])
for _ in range(3): lm(forces)
_lib.timing(True)
for _ in range(5): lm(forces)
print("apply kernel ms", np.mean([ms for n, ms in _lib.timing_records() if n == "agf_map_apply"]))
