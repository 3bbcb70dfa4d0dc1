import sys
sys.path.insert(0, '/root/repo')
import torch
from aggforce_b200 import _engine
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
cols = reduced_columns(175, topo.xh_constraints)
_, f = synth_trajectory_device(topo, 1_000_000, seed=5, want_coords=False)
fr = _engine.Frames(f)
for _ in range(2):
    _engine.gram_linear_raw(fr, cols, 97)
torch.cuda.synchronize()
