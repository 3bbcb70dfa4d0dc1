"""Wall time of each API call of the end-to-end (pinned host arrays in, numpy out) step, per step."""
import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
topo = chignolin_topology()
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
coords, forces = synth_trajectory_device(topo, T, seed=1)
hc = torch.empty(coords.shape, dtype=coords.dtype, pin_memory=True); hc.copy_(coords)
hf = torch.empty(forces.shape, dtype=forces.dtype, pin_memory=True); hf.copy_(forces)
torch.cuda.synchronize()
nc, nf = hc.numpy(), hf.numpy()
del coords, forces
print("affinity", len(os.sched_getaffinity(0)), "cpus;", open("/proc/self/status").read().split("Mems_allowed_list:")[1].split()[0])
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, f = agf.Frames(nc), agf.Frames(nf)
    cons = agf.guess_pairwise_constraints(c); torch.cuda.synchronize(); t1 = time.perf_counter()
    r1 = agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds=cons, method=agf.constraint_aware_uni_map)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r2 = agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds=cons, l2_regularization=1e3)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    del r1, r2
    print(f"step {it}: guess(+2.1 GB up) {1e3*(t1-t0):7.1f} ms | uni(+2.1 GB up, 0.48 GB down) {1e3*(t2-t1):7.1f} ms | "
          f"opt(0.48 GB down) {1e3*(t3-t2):7.1f} ms | total {1e3*(t3-t0):7.1f} ms")
# raw copies for comparison
d = torch.empty((T, 175, 3), dtype=torch.float32, device="cuda")
out = torch.empty((T, 10, 3), dtype=torch.float64, device="cuda")
ho = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hc, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    ho.copy_(out, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"raw H2D 2.1 GB {1e3*(t1-t0):.1f} ms ({2.1/(t1-t0):.1f} GB/s); raw D2H 0.24 GB {1e3*(t2-t1):.1f} ms ({0.24/(t2-t1):.1f} GB/s)")
t0 = time.perf_counter(); x = torch.empty((T, 10, 3), dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
print(f"fresh pinned allocation of 0.24 GB: {1e3*(t1-t0):.1f} ms")
