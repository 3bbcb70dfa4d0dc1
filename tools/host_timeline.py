"""Host-side timeline of the three API calls of a bench step: when each C-ABI launch is issued and
how long the synchronising reads wait (perf_counter, microseconds from the call's entry)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine, _lib
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device

topo = chignolin_topology()
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
coords, forces = synth_trajectory_device(topo, 1_000_000, seed=1)
log = []
real_call = _lib.call
def call(name, *a):
    log.append((time.perf_counter(), "launch " + name)); real_call(name, *a)
_lib.call = call
_engine._lib.call = call
real_sync = torch.cuda.Stream.synchronize
def sync(self):
    t = time.perf_counter(); real_sync(self); log.append((t, f"sync waited {1e6*(time.perf_counter()-t):.0f} us"))
torch.cuda.Stream.synchronize = sync
st = {}
calls = {
    "guess": lambda: st.__setitem__("cons", agf.guess_pairwise_constraints(coords)),
    "uni": lambda: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=st["cons"], method=agf.constraint_aware_uni_map),
    "opt": lambda: agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=st["cons"], l2_regularization=1e3),
}
for _ in range(5):
    for f in calls.values(): f()
torch.cuda.synchronize()
for name, f in calls.items():
    best = None
    for _ in range(7):
        torch.cuda.synchronize(); log.clear(); t0 = time.perf_counter(); f(); t1 = time.perf_counter()
        if best is None or t1 - t0 < best[0]:
            best = (t1 - t0, [(1e6 * (t - t0), what) for t, what in log])
    print(f"== {name}: {1e6*best[0]:.0f} us")
    for t, what in best[1]:
        print(f"   {t:7.0f} us  {what}")
