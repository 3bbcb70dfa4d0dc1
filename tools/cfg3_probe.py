"""Config-3 probe: featurised Gram (id_feat + gb_feat(0,8,1,n_basis=7)) on cln025, timing + fit end to end."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine
from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat, qp_feat_linear_map
from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable
from aggforce_b200.util import Curry
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
topo = chignolin_topology()
coords, forces = synth_trajectory_device(topo, T, seed=2)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
feat = Multifeaturize([id_feat, Curry(gb_feat, inner=0, outer=8, width=1, n_basis=7)])
ctx = _FusedContext(cmap, topo.xh_constraints, _fusable(feat))
c, f = _engine.Frames(coords), _engine.Frames(forces)
ctx.grams(c, f, 0.6955215); torch.cuda.synchronize()
t0 = time.perf_counter(); g = ctx.grams(c, f, 0.6955215); torch.cuda.synchronize(); dt = time.perf_counter() - t0
nf = g.shape[1]
flop = 10 * 3 * nf * (nf + 1) * T
from aggforce_b200 import _lib
_lib.timing(True); ctx.grams(c, f, 0.6955215)
recs = _lib.timing_records(); _lib.timing(False)
kms = sum(ms for name, ms in recs if name.startswith("agf_gram_feat"))
print(f"feat gram kernels only: {kms:.1f} ms  {T/kms*1e3:.3e} frames/s  {flop/kms/1e9:.2f} TFLOP/s algorithmic ({flop/kms/1e9/37.15*100:.1f}% of DMMA peak)")
print(f"feat gram: n_feat {nf}, T {T}: {dt*1e3:.1f} ms  {T/dt:.3e} frames/s  {flop/dt/1e12:.2f} TFLOP/s algorithmic ({flop/dt/1e12/37.15*100:.1f}% of DMMA peak)")
t0 = time.perf_counter()
res = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=topo.xh_constraints,
                         method=qp_feat_linear_map, featurizer=feat, kbt=0.6955215, l2_regularization=1e3)
torch.cuda.synchronize()
print(f"project_forces(qp_feat_linear_map) end to end: {(time.perf_counter()-t0)*1e3:.1f} ms, residual {res['residual']:.6g}")
