"""GPU parity of the tiled tensor-core Gram (agf_gram_linear_i8t, csrc/gram_i8t.cu) through the C ABI:
float64 oracle at the north-star bar (1e-9 relative Frobenius), ragged frame counts (partial chunks, partial
slices, several slabs), column counts around the 96 / 128 tile edges, constraint groups of several sizes,
out-of-range and non-finite frames (float64 leftover pass), agreement with the FP64 DMMA kernel.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from conftest import rel_fro

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _groups(rng, n_sites, n_groups, max_size=5):
    cons = set()
    free = list(rng.permutation(n_sites))
    for _ in range(n_groups):
        k = int(rng.integers(2, max_size + 1))
        if len(free) < k:
            break
        cons.add(frozenset(int(free.pop()) for _ in range(k)))
    return cons


def _tiled(forces_dev, n_sites, ptr_, sites, n_red):
    from aggforce_b200 import _lib

    n_frames = forces_dev.shape[0]
    need = int(_lib.lib().agf_gram_linear_i8t_workspace_bytes(n_sites, n_red, n_frames))
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    gram = torch.zeros((n_red, n_red), dtype=torch.float64, device="cuda")
    d_ptr, d_sites = torch.as_tensor(ptr_, device="cuda"), torch.as_tensor(sites, device="cuda")
    _lib.call("agf_gram_linear_i8t", _p(forces_dev), _lib.F32, n_frames, n_sites, _p(d_ptr), _p(d_sites), n_red,
              _p(gram), _p(ws), C.c_size_t(need), _stream())
    _lib.call("agf_symmetrize", _p(gram), n_red, _stream())
    return gram.cpu().numpy()


def _case(n_sites, n_groups, n_frames, seed, scale_spread=True):
    from aggforce_b200 import _engine

    rng = np.random.default_rng(seed)
    cons = _groups(rng, n_sites, n_groups)
    cols = oracle.group_columns(n_sites, cons)
    n_red = int(cols.max()) + 1
    forces = rng.normal(0, 40.0, size=(n_frames, n_sites, 3))
    if scale_spread:  # per-site magnitudes over six decades: the per-column scales matter
        forces *= 10.0 ** rng.uniform(-3, 3, size=(1, n_sites, 1))
    forces = forces.astype(np.float32)
    ptr_, sites = _engine.csr_from_labels(cols, n_red)
    return forces, cons, ptr_, sites, n_red


@pytest.mark.parametrize("n_sites,n_groups,n_frames", [
    (98, 0, 31),        # smallest supported n_red, less than one chunk
    (130, 9, 100),      # two row blocks, two column blocks, ragged last chunk
    (200, 4, 4129),     # n_pad = 288 (column blocks end past the row blocks), one slice + 3 frames
    (385, 30, 1500),    # 3 x 128 + 1
    (700, 60, 700),
])
def test_tiled_gram_matches_the_float64_oracle(n_sites, n_groups, n_frames):
    forces, cons, ptr_, sites, n_red = _case(n_sites, n_groups, n_frames, n_sites + n_frames)
    got = _tiled(torch.as_tensor(forces, device="cuda"), n_sites, ptr_, sites, n_red)
    ref = oracle.gram_linear(forces, cons)
    assert rel_fro(got, ref) < 1e-9
    # element-wise against the column scales: |err_xy| <= 1e-9 sqrt(G_xx G_yy)
    d = np.sqrt(np.diag(ref))
    assert (np.abs(got - ref) / np.outer(d, d)).max() < 1e-9


def test_tiled_gram_over_several_slabs_and_an_unaligned_view():
    """More frames than one slab (16 384) and a device view that starts 3 frames in."""
    n_sites, n_frames = 150, 16384 * 2 + 77
    forces, cons, ptr_, sites, n_red = _case(n_sites, 12, n_frames + 3, 5)
    dev = torch.as_tensor(forces, device="cuda")[3:]
    got = _tiled(dev, n_sites, ptr_, sites, n_red)
    ref = oracle.gram_linear(forces[3:], cons)
    assert rel_fro(got, ref) < 1e-9


def test_tiled_gram_agrees_with_the_dmma_kernel_and_handles_outliers():
    from aggforce_b200 import _lib

    n_sites, n_frames = 333, 5000
    forces, cons, ptr_, sites, n_red = _case(n_sites, 25, n_frames, 11, scale_spread=False)
    forces[1234, 7, 2] = 5.0e8      # far outside the sampled scale
    forces[4999, 300, 0] = -2.0e9
    dev = torch.as_tensor(forces, device="cuda")
    got = _tiled(dev, n_sites, ptr_, sites, n_red)
    d_ptr, d_sites = torch.as_tensor(ptr_, device="cuda"), torch.as_tensor(sites, device="cuda")
    plain = torch.zeros((n_red, n_red), dtype=torch.float64, device="cuda")
    _lib.call("agf_gram_linear", _p(dev), _lib.F32, n_frames, n_sites, _p(d_ptr), _p(d_sites), n_red, _p(plain), _stream())
    _lib.call("agf_symmetrize", _p(plain), n_red, _stream())
    want = plain.cpu().numpy()
    assert rel_fro(got, want) < 1e-12  # the two huge frames dominate and are exact
    mask = np.abs(want) < 1e12
    assert np.abs(got - want)[mask].max() < 1e-9 * np.abs(want[mask]).max()
    forces[2000, 17, 1] = np.nan
    dev = torch.as_tensor(forces, device="cuda")
    got = _tiled(dev, n_sites, ptr_, sites, n_red)
    plain.zero_()
    _lib.call("agf_gram_linear", _p(dev), _lib.F32, n_frames, n_sites, _p(d_ptr), _p(d_sites), n_red, _p(plain), _stream())
    _lib.call("agf_symmetrize", _p(plain), n_red, _stream())
    want = plain.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).any()
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * np.abs(want[ok]).max()


def test_tiled_gram_is_what_a_large_fit_uses():
    """gram_linear_raw routes float32 input with n_red > 97 through the tiled tensor-core kernel."""
    from aggforce_b200 import _engine, _lib

    n_sites, n_frames = 260, 3000
    forces, cons, _, _, n_red = _case(n_sites, 20, n_frames, 3)
    cols = oracle.group_columns(n_sites, cons)
    _lib.timing(True)
    gram = _engine.gram_linear(_engine.Frames(forces), cols, n_red)
    names = {n for n, _ in _lib.timing_records()}
    _lib.timing(False)
    assert "agf_gram_linear_i8t" in names
    assert rel_fro(gram.cpu().numpy(), oracle.gram_linear(forces, cons)) < 1e-9


# --------------------------------------------------------------------------------------
# agf_map_apply_i8: the large dense application on the tensor cores (csrc/apply_i8.cu)
# --------------------------------------------------------------------------------------
def _apply_both(forces_dev, cmap, nan_mode, out_dtype=torch.float64, atol=1e-8):
    """(out, sumsq, flags) of agf_map_apply_i8 and of the FP64 DMMA entry agf_map_apply_ws on the same input."""
    from aggforce_b200 import _lib

    n_frames, n_sites = forces_dev.shape[0], forces_dev.shape[1]
    res = []
    for entry in ("agf_map_apply_i8", "agf_map_apply_ws"):
        out = torch.full((n_frames, cmap.n_cg, 3), float("nan"), dtype=out_dtype, device="cuda")
        sumsq = torch.zeros(1, dtype=torch.float64, device="cuda")
        flags = torch.zeros(2, dtype=torch.int32, device="cuda")
        code = _lib.F64 if out_dtype == torch.float64 else _lib.F32
        if entry == "agf_map_apply_i8":
            need = int(_lib.lib().agf_map_apply_i8_workspace_bytes(n_sites, cmap.n_ucol, cmap.n_cg, n_frames))
            assert need > 0
            ws = torch.empty(need, dtype=torch.uint8, device="cuda")
            _lib.call(entry, _p(forces_dev), _lib.F32, n_frames, n_sites, _p(cmap.ucol_ptr), _p(cmap.ucol_sites),
                      cmap.n_ucol, _p(cmap.umat_t), cmap.n_cg, _p(out), code, _p(sumsq), nan_mode, atol, _p(flags),
                      _p(ws), C.c_size_t(need), _stream())
        else:
            need = int(_lib.lib().agf_map_apply_workspace_bytes(_lib.F32, n_sites, cmap.n_ucol, cmap.nnz, cmap.n_cg, n_frames))
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda")
            _lib.call(entry, _p(forces_dev), _lib.F32, n_frames, n_sites, _p(cmap.ucol_ptr), _p(cmap.ucol_sites),
                      cmap.n_ucol, cmap.nnz, _p(cmap.umat_t), cmap.n_cg, _p(out), code, _p(sumsq), nan_mode, atol,
                      _p(flags), _p(ws), C.c_size_t(ws.numel()), _stream())
        res.append((out.cpu().numpy(), float(sumsq.item()), flags.cpu().numpy().tolist()))
    return res


def _dense_map(rng, n_cg, n_sites, cons, zero_sites=()):
    cols = oracle.group_columns(n_sites, cons)
    n_red = int(cols.max()) + 1
    red = rng.normal(size=(n_cg, n_red)) * 10.0 ** rng.uniform(-2, 2, size=(n_cg, 1))
    w = red[:, cols]
    for z in zero_sites:
        w[:, z] = 0.0
    return w


@pytest.mark.parametrize("n_cg,n_sites,n_groups,n_frames", [
    (65, 120, 6, 31),       # one bead block, less than one frame block
    (130, 300, 20, 100),    # two bead blocks, ragged frame block
    (500, 700, 60, 2085),
    (200, 1100, 100, 4129),
])
def test_i8_apply_matches_numpy_float64(n_cg, n_sites, n_groups, n_frames):
    from aggforce_b200 import _engine

    rng = np.random.default_rng(n_cg + n_frames)
    forces, cons, _, _, _ = _case(n_sites, n_groups, n_frames, n_cg)
    w = _dense_map(rng, n_cg, n_sites, cons)
    cmap = _engine.CompiledMap(w, keep_zero_columns=False)
    assert not cmap.sparse
    (got, sq, flags), (want, sq_w, flags_w) = _apply_both(torch.as_tensor(forces, device="cuda"), cmap, 1)
    ref = oracle.apply_map(forces, w)
    assert rel_fro(got, ref) < 1e-9  # north-star bar for mapped forces: 1e-6
    assert rel_fro(want, ref) < 1e-12
    assert abs(sq / float((ref ** 2).sum()) - 1) < 1e-9 and flags == flags_w == [0, 0]
    # every bead on its own (rows and columns spread over several decades each): two decades under the bar
    num = np.sqrt(((got - ref) ** 2).sum(axis=(0, 2)))
    assert (num / np.sqrt((ref ** 2).sum(axis=(0, 2)))).max() < 1e-8


def test_i8_apply_nan_protocol_outliers_and_float32_output():
    from aggforce_b200 import _engine

    n_cg, n_sites, n_frames = 150, 400, 3000
    rng = np.random.default_rng(4)
    forces, cons, _, _, _ = _case(n_sites, 30, n_frames, 9, scale_spread=False)
    free = sorted(set(range(n_sites)) - {i for c in cons for i in c})
    zero = free[:3]
    w = _dense_map(rng, n_cg, n_sites, cons, zero_sites=zero)
    forces[100, zero[0], 1] = np.nan       # under an all-zero column: ignored, no flag at all (the column is dropped)
    forces[2500, free[10], 0] = 7.0e8      # far outside the sampled scale: float64 pass, exact
    cmap = _engine.CompiledMap(w, keep_zero_columns=False)
    dev = torch.as_tensor(forces, device="cuda")
    (got, sq, flags), (want, sq_w, flags_w) = _apply_both(dev, cmap, 1)
    assert flags == flags_w and np.isfinite(got).all()
    assert rel_fro(got, want) < 1e-9 and abs(sq / sq_w - 1) < 1e-9
    assert np.array_equal(got[2500], want[2500])  # the same float64 arithmetic order is not promised, the values are
    # a NaN under a live column: counts as 0, both flags raised exactly like the FP64 entry
    forces[700, free[20], 2] = np.nan
    dev = torch.as_tensor(forces, device="cuda")
    (got, sq, flags), (want, sq_w, flags_w) = _apply_both(dev, cmap, 1)
    assert flags == flags_w == [1, 1] and np.isfinite(got).all()
    assert rel_fro(got, want) < 1e-9 and abs(sq / sq_w - 1) < 1e-9
    # plain mode (zero columns kept): NaN propagates to every bead of that frame and component, 0 * NaN included
    cmap0 = _engine.CompiledMap(w, keep_zero_columns=True)
    (got, _, _), (want, _, _) = _apply_both(dev, cmap0, 0)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[700][:, 2]).all() and np.isnan(got[100][:, 1]).all()
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * np.abs(want[ok]).max()
    # float32 output: values as stored, residual sum of the stored values
    forces = np.nan_to_num(forces, nan=0.0)
    dev = torch.as_tensor(forces, device="cuda")
    (got, sq, _), (want, sq_w, _) = _apply_both(dev, cmap, 1, out_dtype=torch.float32)
    assert got.dtype == np.float32 and rel_fro(got, want) < 1e-6 and abs(sq / sq_w - 1) < 1e-6


def test_i8_apply_is_what_a_large_linear_map_uses():
    from aggforce_b200 import LinearMap, _lib

    n_cg, n_sites, n_frames = 100, 500, 2500
    rng = np.random.default_rng(8)
    forces, cons, _, _, _ = _case(n_sites, 40, n_frames, 2)
    w = _dense_map(rng, n_cg, n_sites, cons)
    lm = LinearMap(w)
    _lib.timing(True)
    out = lm(forces)
    names = {n for n, _ in _lib.timing_records()}
    _lib.timing(False)
    assert "agf_map_apply_i8" in names
    assert rel_fro(out, oracle.apply_map(forces, w)) < 1e-9
    mapped, sumsq = lm.apply_with_sumsq(torch.as_tensor(forces, device="cuda"))
    assert abs(sumsq / float((mapped.double() ** 2).sum().item()) - 1) < 1e-12


def test_tiled_gram_with_a_reordered_column_table():
    """98 <= n_red <= 128: the plan orders the columns by group size (the DMMA kernel's layout), the tiled kernel
    takes them as they come and gram_linear maps the result back to the caller's order."""
    from aggforce_b200 import _engine, _lib

    n_sites, n_frames = 131, 3000
    forces, cons, _, _, n_red = _case(n_sites, 9, n_frames, 21)
    assert 98 <= n_red <= 128
    cols = oracle.group_columns(n_sites, cons)
    _lib.timing(True)
    gram = _engine.gram_linear(_engine.Frames(forces), cols, n_red)
    names = {n for n, _ in _lib.timing_records()}
    _lib.timing(False)
    assert "agf_gram_linear_i8t" in names
    assert rel_fro(gram.cpu().numpy(), oracle.gram_linear(forces, cons)) < 1e-9
