"""int8/tcgen05 Gram (agf_gram_linear_i8) against the FP64 DMMA kernel: error and time."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _engine, _lib
from aggforce_b200.qp.qplinear import reduced_columns
from aggforce_b200.synth import chignolin_topology, synth_trajectory_device

topo = chignolin_topology()
cols = reduced_columns(175, topo.xh_constraints)
n_red = int(cols.max()) + 1


def gram(frames, use_i8, cols=cols, n_red=n_red):
    _engine._GRAM_I8[0] = use_i8
    g, order = _engine.gram_linear_raw(_engine.Frames(frames), cols, n_red)
    torch.cuda.synchronize()
    u = torch.triu(g).cpu().numpy()
    full = u + np.triu(u, 1).T  # symmetric, internal column order
    out = np.empty_like(full)
    out[np.ix_(order, order)] = full  # caller's column order (the two kernels use different internal orders)
    return out, order


def timeit(frames, use_i8, n=5):
    _engine._GRAM_I8[0] = use_i8
    fr = _engine.Frames(frames)
    for _ in range(2):
        _engine.gram_linear_raw(fr, cols, n_red)
    _lib.timing(True)
    for _ in range(n):
        _engine.gram_linear_raw(fr, cols, n_red)
    recs = _lib.timing_records(); _lib.timing(False)
    return np.mean([ms for name, ms in recs if name.startswith("agf_gram_linear")])


for T in (8192, 10001, 65536, 1_000_000):
    _, f = synth_trajectory_device(topo, T, seed=5, want_coords=False)
    a, _ = gram(f, True)
    b, _ = gram(f, False)
    err = np.linalg.norm(a - b) / np.linalg.norm(b)
    print(f"T {T:8d}: rel Frobenius (upper triangle) i8 vs DMMA {err:.2e}", flush=True)
    if T == 10001:  # unaligned view + leftovers: huge and non-finite values far beyond the sampled scale
        g = f[3:].clone()
        g[7000, 5, 1] = 3.0e7
        g[9000, 100, 2] = -1.0e9
        a, _ = gram(g, True)
        b, _ = gram(g, False)
        print(f"   unaligned view + 2 out-of-range frames: {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}", flush=True)
        g[8000, 17, 0] = float('nan')
        a, _ = gram(g, True)
        b, _ = gram(g, False)
        print("   NaN frame: same NaN pattern:", np.array_equal(np.isnan(a), np.isnan(b)), int(np.isnan(a).sum()), flush=True)
# a smaller system: first 100 sites, its own constraints (n_red < 96)
sub_cons = {c for c in topo.xh_constraints if max(c) < 100}
scols = reduced_columns(100, sub_cons)
sn = int(scols.max()) + 1
_, f = synth_trajectory_device(topo, 50000, seed=6, want_coords=False)
fs = f[:, :100].contiguous()
a, _ = gram(fs, True, scols, sn)
b, _ = gram(fs, False, scols, sn)
print(f"100-site subsystem, n_red {sn}: {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}", flush=True)
_, f = synth_trajectory_device(topo, 1_000_000, seed=5, want_coords=False)
print(f"1 M frames: i8 {timeit(f, True):.3f} ms, DMMA {timeit(f, False):.3f} ms")
