"""Featurised Gram at config 3 (cln025, n_feat 769, 10 beads): int8 tensor-core path against the FP64 DMMA path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat
from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device
from aggforce_b200.util import Curry

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
topo = chignolin_topology()
c, f = synth_trajectory_device(topo, T, seed=2)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
feat = Multifeaturize([id_feat, Curry(gb_feat, inner=0, outer=8, width=1, n_basis=7)])
ctx = _FusedContext(cmap, topo.xh_constraints, _fusable(feat))
fc, ff = _engine.Frames(c), _engine.Frames(f)
flop = 10 * 3 * 769 * 770 * T

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
for on in (True, False):
    _engine._GRAM_I8[0] = on
    _lib.timing(True); g = ctx.grams(fc, ff, 0.6955215, on_device=True); names = sorted({n for n, _ in _lib.timing_records()}); _lib.timing(False)
    ms = timeit(lambda: ctx.grams(fc, ff, 0.6955215, on_device=True))
    res[on] = g.cpu().numpy()
    print(f"{'int8' if on else 'FP64 DMMA'}: {ms:.3f} ms  {T/ms*1e3:.3e} frames/s  {flop/ms/1e9:.1f} float64-equivalent TFLOP/s  entries {names}", flush=True)
a, b = res[True], res[False]
print("rel fro int8 vs DMMA per bead:", [float(f"{np.linalg.norm(a[i]-b[i])/np.linalg.norm(b[i]):.2e}") for i in range(10)])
