"""Config-4 shape probe (5000 atoms, 500 beads, n_red 2600): parity on a sub-sample + kernel timings."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import oracle
import aggforce_b200 as agf
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

topo = protein_like_topology(500)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
coords, forces = synth_trajectory_device(topo, T, seed=3)
cons = topo.xh_constraints
cols = reduced_columns(topo.n_sites, cons)
n_red = int(cols.max()) + 1
print("n_sites", topo.n_sites, "n_red", n_red, "T", T)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fr = _engine.Frames(forces)
ms = timeit(lambda: _engine.gram_linear(fr, cols, n_red))
flop = (3 * n_red * (n_red + 1) + 3 * topo.n_sites) * T
print(f"gram: {ms:.2f} ms  {T/ms*1e3:.3e} frames/s  {flop/ms/1e9:.2f} TFLOP/s algorithmic ({flop/ms/1e9/37.15*100:.1f}% of DMMA peak)")
g = _engine.gram_linear(fr, cols, n_red).cpu().numpy()
sub = forces[:64].cpu().numpy()
g64 = _engine.gram_linear(_engine.Frames(forces[:64]), cols, n_red).cpu().numpy()
ref = oracle.gram_linear(sub, cons)
print("gram parity (64 frames) rel fro", np.linalg.norm(g64 - ref) / np.linalg.norm(ref))
# constraints
t0 = time.perf_counter(); got = agf.guess_pairwise_constraints(coords); torch.cuda.synchronize()
print("constraints", len(got), "== topology:", got == cons, f"{(time.perf_counter()-t0)*1e3:.1f} ms")
# dense apply with a random dense 500 x 5000 map whose columns repeat within constraint groups
rng = np.random.default_rng(0)
red = rng.normal(size=(500, n_red))
lm = agf.LinearMap(red[:, cols])
ms = timeit(lambda: lm(forces))
aflop = 6 * 500 * n_red * T
print(f"dense apply (n_ucol {n_red}): {ms:.2f} ms  {T/ms*1e3:.3e} frames/s  {aflop/ms/1e9:.2f} TFLOP/s ({aflop/ms/1e9/37.15*100:.1f}% of DMMA peak)")
_lib.timing(True); lm(forces); recs = _lib.timing_records(); _lib.timing(False)
print("  entry points:", ", ".join(f"{n} {ms:.2f} ms" for n, ms in recs))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); lm(forces); lm(forces); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
lmn = agf.LinearMap(red[:, cols], handle_nans=False)
ms = timeit(lambda: lmn(forces))
print(f"dense apply, handle_nans=False: {ms:.2f} ms")
fn = forces[:4096].clone(); fn[5, 7, 1] = float("nan"); fn[4000, 4999, 2] = float("nan")
red0 = red.copy(); red0[:, cols[7]] = 0.0; red0[:, cols[4999]] = 0.0
lm0 = agf.LinearMap(red0[:, cols])
on = lm0(fn).cpu().numpy()
fz = fn.cpu().numpy().copy(); fz[np.isnan(fz)] = 0.0
refn = oracle.apply_map(fz, red0[:, cols])
print("apply with NaNs under zero columns: rel fro", np.linalg.norm(on - refn) / np.linalg.norm(refn))
out = lm(forces[:32]).cpu().numpy()
refo = oracle.apply_map(forces[:32].cpu().numpy(), red[:, cols])
print("apply parity rel fro", np.linalg.norm(out - refo) / np.linalg.norm(refo))
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
ms = timeit(lambda: cmap(coords))
print(f"slice apply: {ms:.3f} ms  {T/ms*1e3:.3e} frames/s")

# whole path at this shape: constraints + Gram + host QP (n_red 2600, 500 beads) + both applications
t0 = time.perf_counter()
res = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds="auto", l2_regularization=1e3)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"project_forces end to end: {dt*1e3:.0f} ms ({T/dt:.3e} frames/s)")
pr = cProfile.Profile(); pr.enable()
res = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=1e3)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
sub_w = oracle.qp_linear_weights(forces.cpu().numpy(), cmap.standard_matrix, cons, 1e3) if T <= 4096 else None
if sub_w is not None:
    w = res["tmap"].force_map.standard_matrix
    print("weights rel err vs oracle:", np.linalg.norm(w - sub_w) / np.linalg.norm(sub_w))
