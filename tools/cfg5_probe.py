"""Config-5 shape probe (2000 atoms, 200 beads): joptgauss_map fit + application on device-resident
synthetic frames, augmented arrays generated slab by slab; parity of the fit on a sub-sample."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import oracle
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
topo = protein_like_topology(200)
coords, forces = synth_trajectory_device(topo, T, seed=5)
cons = topo.xh_constraints
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
print("n_sites", topo.n_sites, "beads", len(topo.bead_atoms), "constraints", len(cons), "T", T)
traj = agf.Trajectory(coords=coords, forces=forces)
var, kbt = 0.25, 0.6955215
def fit():
    return agf.joptgauss_map(traj, cmap, var=var, kbt=kbt, constraints=cons, seed=42100, l2_regularization=1e3)
fit(); torch.cuda.synchronize()
_lib.timing(True)
t0 = time.perf_counter(); tmap = fit(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
recs = _lib.timing_records(); _lib.timing(False)
agg = {}
for n, ms in recs: agg[n] = agg.get(n, 0.0) + ms
print(f"fit: {dt*1e3:.1f} ms ({T/dt:.3e} frames/s); entry points:", ", ".join(f"{n} {ms:.1f} ms" for n, ms in agg.items()))
n_red = tmap.tmap.force_map.standard_matrix.shape[1]
t0 = time.perf_counter(); out = tmap(traj); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"apply: {dt*1e3:.1f} ms ({T/dt:.3e} frames/s), mapped forces {tuple(out.forces.shape)}")
# parity on a sub-sample with injected noise
n = 256
sub_c, sub_f = coords[:n].cpu().numpy(), forces[:n].cpu().numpy()
noise = np.random.default_rng(1).standard_normal((n, len(topo.bead_atoms), 3)).astype(np.float32)
small = agf.joptgauss_map(agf.Trajectory(coords=sub_c, forces=sub_f), cmap, var=var, kbt=kbt, constraints=cons,
                          noise=noise, l2_regularization=1e3)
full_c, full_f = oracle.gauss_augment(sub_c, sub_f, cmap.standard_matrix, var, kbt, noise)
n_all = topo.n_sites + len(topo.bead_atoms)
aug_cm = np.zeros((len(topo.bead_atoms), n_all)); aug_cm[np.arange(len(topo.bead_atoms)), topo.n_sites + np.arange(len(topo.bead_atoms))] = 1
ref_w = oracle.qp_linear_weights(full_f.astype(np.float32), aug_cm, cons, 1e3)
w = small.tmap.force_map.standard_matrix
print("weights rel err vs oracle (f32 augmented arrays):", np.linalg.norm(w - ref_w) / np.linalg.norm(ref_w))
