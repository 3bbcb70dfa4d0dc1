"""Tiled tensor-core Gram (agf_gram_linear_i8t) against the FP64 DMMA packed-panel kernel (agf_gram_linear_ws)
at the config-4 shape (5 000 atoms, n_red 2 600) and the config-5 shape: times, float64-equivalent TFLOP/s,
agreement.  usage: gram_tiled_time.py [frames] [beads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
beads = int(sys.argv[2]) if len(sys.argv) > 2 else 500
topo = protein_like_topology(beads)
_, forces = synth_trajectory_device(topo, T, seed=3)
cols = reduced_columns(topo.n_sites, topo.xh_constraints)
n_red = int(cols.max()) + 1
print("n_sites", topo.n_sites, "n_red", n_red, "T", T, flush=True)
fr = _engine.Frames(forces)
flop = (3 * n_red * (n_red + 1) + 3 * topo.n_sites) * T

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
    _lib.timing(True); g = _engine.gram_linear(fr, cols, n_red); names = sorted({n for n, _ in _lib.timing_records()}); _lib.timing(False)
    ms = timeit(lambda: _engine.gram_linear(fr, cols, n_red))
    res[on] = g.cpu().numpy()
    print(f"{'int8 tiled' if on else 'FP64 DMMA'}: {ms:.3f} ms  {T/ms*1e3:.3e} frames/s  {flop/ms/1e9:.1f} float64-equivalent TFLOP/s"
          f" ({flop/ms/1e9/37.1*100:.0f}% of the DMMA peak)  entries {names}", flush=True)
a, b = res[True], res[False]
print("rel fro int8 vs DMMA:", np.linalg.norm(a - b) / np.linalg.norm(b))
d = np.sqrt(np.diag(b))
print("max |err| / sqrt(Gxx Gyy):", (np.abs(a - b) / np.outer(d, d)).max())
