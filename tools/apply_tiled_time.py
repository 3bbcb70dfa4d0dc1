"""Large dense application: agf_map_apply_i8 (tcgen05 int8) against agf_map_apply_ws (FP64 DMMA) at the config-4
shape (500 beads x n_red 2 600) or the config-5 shape.  usage: apply_tiled_time.py [frames] [beads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
beads = int(sys.argv[2]) if len(sys.argv) > 2 else 500
topo = protein_like_topology(beads)
_, forces = synth_trajectory_device(topo, T, seed=3)
cols = reduced_columns(topo.n_sites, topo.xh_constraints)
n_red = int(cols.max()) + 1
rng = np.random.default_rng(0)
lm = agf.LinearMap(rng.normal(size=(beads, n_red))[:, cols])
print("n_sites", topo.n_sites, "n_red", n_red, "beads", beads, "T", T, flush=True)
flop = 6 * beads * n_red * T

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
    _lib.timing(True); o = lm(forces); names = sorted({n for n, _ in _lib.timing_records()}); _lib.timing(False)
    ms = timeit(lambda: lm(forces))
    res[on] = o.cpu().numpy()
    print(f"{'int8' if on else 'FP64 DMMA'}: {ms:.3f} ms  {T/ms*1e3:.3e} frames/s  {flop/ms/1e9:.1f} float64-equivalent TFLOP/s  entries {names}", flush=True)
a, b = res[True], res[False]
print("rel fro int8 vs DMMA:", np.linalg.norm(a - b) / np.linalg.norm(b))
