import sys, time
sys.path.insert(0, '/root/repo')
import torch
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
_, forces = synth_trajectory_device(topo, 1_000_000, seed=1, want_coords=False)
cols = reduced_columns(175, topo.xh_constraints)
fr = _engine.Frames(forces)
for _ in range(3): _engine.gram_linear(fr, cols, 97)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): _engine.gram_linear(fr, cols, 97)
e1.record(); torch.cuda.synchronize()
print("gram ms", e0.elapsed_time(e1)/5)
