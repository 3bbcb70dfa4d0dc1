import sys, ctypes, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
coords, _ = synth_trajectory_device(topo, 1_000_000, seed=1, want_forces=False)
gran = int(os.environ.get("L2G", "0"))
if gran:
    rt = ctypes.CDLL("libcudart.so.12")
    val = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5); print("before", val.value)
    print("set rc", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)))
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5); print("after", val.value)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
for _ in range(3): cmap(coords)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): cmap(coords)
e1.record(); torch.cuda.synchronize()
print("sparse apply ms", e0.elapsed_time(e1)/10)
