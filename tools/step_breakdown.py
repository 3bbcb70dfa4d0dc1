"""Where a bench step's wall time goes: per API call wall time (synchronised) vs its kernels' time,
plus a cProfile of the host side sorted by own time."""
import cProfile, pstats, sys, time
sys.path.insert(0, '/root/repo')
import torch
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
topo = chignolin_topology()
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
coords, forces = synth_trajectory_device(topo, 1_000_000, seed=1)
calls = {
    "guess": lambda st: st.__setitem__("cons", agf.guess_pairwise_constraints(coords)),
    "uni": lambda st: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=st["cons"], method=agf.constraint_aware_uni_map),
    "qp": lambda st: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=st["cons"], l2_regularization=1e3),
}
st = {}
for _ in range(3):
    for f in calls.values(): f(st)
torch.cuda.synchronize()
N = 20
wall = {k: 0.0 for k in calls}; kern = {k: 0.0 for k in calls}
for _ in range(N):
    for k, f in calls.items():
        _lib.timing(True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(st); torch.cuda.synchronize(); wall[k] += time.perf_counter() - t0
        kern[k] += sum(ms for _, ms in _lib.timing_records()); _lib.timing(False)
for k in calls:
    print(f"{k:6s} wall {wall[k]/N*1e3:.3f} ms  kernels {kern[k]/N:.3f} ms  host-exposed {wall[k]/N*1e3-kern[k]/N:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(N):
    for f in calls.values(): f(st)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
