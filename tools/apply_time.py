import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
_, forces = synth_trajectory_device(topo, 1_000_000, seed=1, want_coords=False)
cols = reduced_columns(175, topo.xh_constraints)
rng = np.random.default_rng(0)
lm = agf.LinearMap(rng.normal(size=(10, 97))[:, cols])
for _ in range(3): lm(forces)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): lm(forces)
e1.record(); torch.cuda.synchronize()
print("apply ms", e0.elapsed_time(e1)/5)
