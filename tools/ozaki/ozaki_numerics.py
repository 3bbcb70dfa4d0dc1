"""Numerics of an int8-sliced (Ozaki-style) Gram on the workloads of BASELINE.json -- CPU study.

G = sum_k v_k v_k' with v the float64 group sums of float32 forces.  Scheme evaluated here:
  * rows are processed in chunks; inside a chunk every column x gets a power-of-two scale
    2^e_x >= max_k |v[k][x]|;
  * v[k][x] / 2^e_x is rounded to B = S*W bits and split into S signed W-bit slices
    (two's-complement digits: slice 0 most significant, digits in [-2^(W-1), 2^(W-1)) );
  * slice products A_s' A_t are EXACT in int32 (what tcgen05.mma kind::i8 computes); pairs with
    s + t >= S are dropped (their weight is below the rounding already made);
  * chunk results are recombined in float64: G += sum_{s,t} 2^(e_x + e_y - W (s + t + 2)) (A_s' A_t).
Reported: relative Frobenius error against the float64 Gram (bar: 1e-9), number of int8 GEMMs.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from aggforce_b200.qp.qplinear import reduced_columns  # noqa: E402
from aggforce_b200.synth import chignolin_topology, protein_like_topology, synth_trajectory_host  # noqa: E402


def group_rows(forces, cols, n_red):
    T, n, _ = forces.shape
    v = np.zeros((T, 3, n_red))
    np.add.at(v, (slice(None), slice(None), cols), np.swapaxes(forces.astype(np.float64), 1, 2))
    return v.reshape(T * 3, n_red)


def sliced_gram(v, n_slices, width, chunk_rows, drop=True):
    n = v.shape[1]
    gram = np.zeros((n, n))
    n_gemm = 0
    for r0 in range(0, v.shape[0], chunk_rows):
        blk = v[r0:r0 + chunk_rows]
        amax = np.abs(blk).max(axis=0)
        e = np.where(amax > 0, np.ceil(np.log2(np.maximum(amax, 1e-300))) + 1, 0.0)  # |v| / 2^e < 1/2
        total_bits = n_slices * width
        q = np.rint(blk / 2.0 ** e * 2.0 ** (total_bits - 1)).astype(np.int64)  # signed fixed point, B bits
        # balanced digits, most significant first: q = sum_s d_s 2^(W (S-1-s)), d_s in [-2^(W-1), 2^(W-1))
        digits = []
        rem = q.copy()
        for s in range(n_slices - 1, -1, -1):
            d = ((rem + (1 << (width - 1))) & ((1 << width) - 1)) - (1 << (width - 1))
            rem = (rem - d) >> width
            digits.append(d)
        assert np.all(rem == 0) or np.all(np.abs(rem) <= 1)
        digits = digits[::-1]  # digits[0] most significant
        assert all(np.abs(d).max() <= (1 << (width - 1)) for d in digits)
        part = np.zeros((n, n))
        n_gemm = 0
        for s in range(n_slices):
            for t in range(n_slices):
                if drop and s + t >= n_slices:
                    continue
                prod = digits[s].T.astype(np.int64) @ digits[t].astype(np.int64)  # exact; int32 range checked below
                assert np.abs(prod).max() < 2**31
                part += prod.astype(np.float64) * 2.0 ** (-width * (s + t))
                n_gemm += 1
        scale = 2.0 ** (e[:, None] + e[None, :] - 2 * (width - 1) - 0) * 2.0 ** (-2 * (total_bits - width))
        # q = sum_s d_s 2^(W(S-1-s))  ->  v ~ q 2^(e - (B-1)); product weight 2^(e_x+e_y-2(B-1)) 2^(W(2S-2-s-t))
        gram += part * 2.0 ** (e[:, None] + e[None, :] - 2 * (total_bits - 1) + 2 * width * (n_slices - 1))
        del scale
    return gram, n_gemm


def study(name, v):
    ref = v.T @ v
    print(f"{name}: rows {v.shape[0]}, n_red {v.shape[1]}")
    for n_slices, width in ((4, 8), (5, 7), (5, 8), (6, 7)):
        for chunk in (1536, 49152):
            if chunk > v.shape[0] and chunk != 1536:
                continue
            g, n_gemm = sliced_gram(v, n_slices, width, chunk)
            err = np.linalg.norm(g - ref) / np.linalg.norm(ref)
            print(f"  {n_slices} slices x {width} bit, chunks of {chunk:6d} rows, {n_gemm:2d} int8 GEMMs: "
                  f"rel Frobenius error {err:.2e}  ({'PASS' if err < 1e-9 else 'FAIL'} at 1e-9)")


if __name__ == "__main__":
    topo = chignolin_topology()
    _, f = synth_trajectory_host(topo, 20000, seed=3)
    cols = reduced_columns(topo.n_sites, topo.xh_constraints)
    study("cln025 synthetic (config 2 shape)", group_rows(f, cols, int(cols.max()) + 1))
    wd = np.load(Path(__file__).resolve().parents[2] / "tests/golden/waterdimer.npz")
    study("water dimer (reference test data, real forces)", group_rows(wd["Fs"], np.arange(6), 6))
    topo4 = protein_like_topology(60)  # config-4-like topology, smaller (600 atoms, n_red 312)
    _, f4 = synth_trajectory_host(topo4, 3000, seed=4)
    cols4 = reduced_columns(topo4.n_sites, topo4.xh_constraints)
    study("protein-like 600 atoms (config 4 topology, reduced size)", group_rows(f4, cols4, int(cols4.max()) + 1))
