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
