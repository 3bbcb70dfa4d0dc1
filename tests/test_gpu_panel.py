"""GPU parity of the packed-panel entry points (agf_gram_linear_ws, agf_map_apply_ws,
agf_gram_feat_ws) called directly through the C ABI: slab processing with deliberately small
workspaces, narrow / exact / ragged last blocks, the NaN redo path of the GEMM apply, and
agreement with the workspace-free kernels.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from conftest import rel_fro

pytestmark = pytest.mark.gpu

PANEL_BYTES = 24 * 132 * 8


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _groups(rng, n_sites, n_groups):
    cons = set()
    free = list(rng.permutation(n_sites))
    for _ in range(n_groups):
        k = int(rng.integers(2, 5))
        if len(free) < k:
            break
        cons.add(frozenset(int(free.pop()) for _ in range(k)))
    return cons


@pytest.mark.parametrize("n_sites,n_groups,n_frames", [(129, 0, 21), (140, 5, 64), (256, 0, 9), (300, 20, 77),
                                                       (640, 60, 33)])
@pytest.mark.parametrize("slab_chunks", [1, 3, 1000])
def test_gram_linear_ws_slabs(n_sites, n_groups, n_frames, slab_chunks):
    from aggforce_b200 import _engine, _lib

    rng = np.random.default_rng(n_sites + n_frames)
    cons = _groups(rng, n_sites, n_groups)
    cols = oracle.group_columns(n_sites, cons)
    n_red = int(cols.max()) + 1
    forces = rng.normal(0, 40.0, size=(n_frames, n_sites, 3)).astype(np.float32)
    ptr_, sites = _engine.csr_from_labels(cols, n_red)
    d_f = torch.as_tensor(forces, device="cuda")
    d_ptr, d_sites = torch.as_tensor(ptr_, device="cuda"), torch.as_tensor(sites, device="cuda")
    gram = torch.zeros((n_red, n_red), dtype=torch.float64, device="cuda")
    n_blocks = (n_red + 127) // 128
    ws = torch.empty(slab_chunks * n_blocks * PANEL_BYTES, dtype=torch.uint8, device="cuda")
    _lib.call("agf_gram_linear_ws", _p(d_f), _lib.F32, n_frames, n_sites, _p(d_ptr), _p(d_sites), n_red, _p(gram),
              _p(ws), C.c_size_t(ws.numel()), _stream())
    _lib.call("agf_symmetrize", _p(gram), n_red, _stream())
    ref = oracle.gram_linear(forces, cons)
    assert rel_fro(gram.cpu().numpy(), ref) < 1e-9
    # the workspace-free kernel computes the same matrix
    plain = torch.zeros_like(gram)
    _lib.call("agf_gram_linear", _p(d_f), _lib.F32, n_frames, n_sites, _p(d_ptr), _p(d_sites), n_red, _p(plain),
              _stream())
    _lib.call("agf_symmetrize", _p(plain), n_red, _stream())
    assert rel_fro(gram.cpu().numpy(), plain.cpu().numpy()) < 1e-13


def _apply_ws(x, m, ws_bytes, nan_mode=0, out_dtype=torch.float64, want_sumsq=True):
    from aggforce_b200 import _engine, _lib

    cm = _engine.CompiledMap(m, keep_zero_columns=nan_mode == 0)
    assert not cm.sparse
    d_x = torch.as_tensor(x, device="cuda")
    out = torch.full((x.shape[0], m.shape[0], 3), 7.0, dtype=out_dtype, device="cuda")
    sumsq = torch.zeros(1, dtype=torch.float64, device="cuda") if want_sumsq else None
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    need = int(_lib.lib().agf_map_apply_workspace_bytes(_engine.dtype_code(d_x), x.shape[1], cm.n_ucol, cm.nnz,
                                                        cm.n_cg, x.shape[0]))
    ws = None
    if ws_bytes is not None:
        ws = torch.empty(need if ws_bytes == "full" else ws_bytes(cm), dtype=torch.uint8, device="cuda")
    _lib.call("agf_map_apply_ws", _p(d_x), _engine.dtype_code(d_x), x.shape[0], x.shape[1], _p(cm.ucol_ptr),
              _p(cm.ucol_sites), cm.n_ucol, cm.nnz, _p(cm.umat_t), cm.n_cg, _p(out), _engine.dtype_code(out),
              _p(sumsq), nan_mode, 1e-6, _p(flags), _p(ws), C.c_size_t(0 if ws is None else ws.numel()), _stream())
    return out.cpu().numpy(), None if sumsq is None else float(sumsq.item()), flags.cpu().numpy(), need


def _one_block_ws(cm):
    kch = (cm.n_ucol + 23) // 24
    return 256 + ((cm.n_cg + 127) // 128) * kch * PANEL_BYTES + 3 * kch * PANEL_BYTES


@pytest.mark.parametrize("n_cg,n_fg,n_frames", [(70, 90, 33), (65, 700, 128), (129, 260, 300), (300, 700, 419)])
@pytest.mark.parametrize("in_dtype", [np.float32, np.float64])
def test_map_apply_ws_gemm(n_cg, n_fg, n_frames, in_dtype):
    rng = np.random.default_rng(n_cg + n_fg)
    m = rng.normal(size=(n_cg, n_fg))
    m[:, n_fg // 3] = m[:, 1]
    x = rng.normal(0, 30, size=(n_frames, n_fg, 3)).astype(in_dtype)
    ref = oracle.apply_map(x, m)
    full, sumsq, flags, need = _apply_ws(x, m, "full")
    assert need > 0
    assert rel_fro(full, ref) < 1e-12
    assert abs(sumsq / float((ref ** 2).sum()) - 1) < 1e-12
    assert not flags.any()
    slabbed, sumsq2, _, _ = _apply_ws(x, m, _one_block_ws)  # one 128-frame block per slab
    assert np.array_equal(slabbed, full)
    assert abs(sumsq2 / sumsq - 1) < 1e-12
    fallback, sumsq3, _, _ = _apply_ws(x, m, None)  # no workspace: DFMA fallback kernel
    assert rel_fro(fallback, ref) < 1e-12 and abs(sumsq3 / sumsq - 1) < 1e-12
    f32, _, _, _ = _apply_ws(x, m, "full", out_dtype=torch.float32, want_sumsq=False)
    assert f32.dtype == np.float32 and rel_fro(f32, ref) < 1e-6


def test_map_apply_ws_nan_protocol():
    """NaNs under all-zero map columns are ignored (redo path of the GEMM apply); NaNs under
    non-zero columns raise the violation flag; plain mode propagates NaN like numpy."""
    rng = np.random.default_rng(3)
    n_cg, n_fg, n_frames = 80, 200, 300
    m = rng.normal(size=(n_cg, n_fg))
    m[:, 17] = 0.0
    m[:, 150] = 0.0
    x = rng.normal(0, 10, size=(n_frames, n_fg, 3)).astype(np.float32)
    m[:, 60] = 1e-9  # referenced, but its weight mass is below the protocol's atol
    x[5, 17, 1] = np.nan
    x[299, 150, :] = np.nan
    x[10, 60, 2] = np.nan  # read by the pack kernel -> the slab is redone with the masking kernel
    clean = np.nan_to_num(x, nan=0.0)
    out, sumsq, flags, _ = _apply_ws(x, m, "full", nan_mode=1)
    ref = oracle.apply_map(clean, m)
    assert rel_fro(out, ref) < 1e-12 and flags[0] == 1 and not flags[1]
    assert abs(sumsq / float((ref ** 2).sum()) - 1) < 1e-12
    x[7, 3, 0] = np.nan  # under a non-zero column: result would depend on the NaN
    out, _, flags, _ = _apply_ws(x, m, "full", nan_mode=1)
    assert flags[0] == 1 and flags[1] == 1
    plain, _, _, _ = _apply_ws(x, m, "full", nan_mode=0)
    assert np.isnan(plain[7, :, 0]).all() and np.isnan(plain[5, :, 1]).all()
    assert np.isfinite(plain[8]).all()


@pytest.mark.parametrize("slab_chunks", [1, 2, 1000])
@pytest.mark.parametrize("drop_last,in_dtype", [(True, np.float32), (False, np.float32), (True, np.float64)])
def test_gram_feat_ws_matches_fused_kernel_and_oracle(slab_chunks, drop_last, in_dtype):
    from aggforce_b200 import LinearMap, _engine, _lib
    from aggforce_b200.qp.featlinearmap import id_feat
    from aggforce_b200.synth import chignolin_topology, synth_trajectory_host

    topo = chignolin_topology()
    n_frames, nb, kbt = 21, 5, 0.6955215
    coords, forces = synth_trajectory_host(topo, n_frames, seed=77)
    beads = topo.bead_atoms[:3]
    cmap = LinearMap([[i] for i in beads], n_fg_sites=topo.n_sites)
    labels = id_feat(None, cmap, topo.xh_constraints, return_ids=True)
    G = int(labels.max()) + 1
    n_ch = G - 1 if drop_last else G
    n_feat = G + nb * n_ch
    ptr_, sites = _engine.csr_from_labels(labels, G)
    centers = np.linspace(0.0, 8.0 ** 0.5, nb) ** 2
    dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device="cuda")  # noqa: E731
    d_c, d_f = dev(coords, in_dtype), dev(forces, in_dtype)
    code = _lib.F32 if in_dtype == np.float32 else _lib.F64
    d_ptr, d_sites = dev(ptr_, np.int32), dev(sites, np.int32)
    d_bptr, d_bsites = dev(np.arange(len(beads) + 1), np.int32), dev(beads, np.int32)
    d_bw, d_cent = dev(np.ones(len(beads)), np.float64), dev(centers, np.float64)
    common = (_p(d_ptr), _p(d_sites), G, n_ch, _p(d_bptr), _p(d_bsites), _p(d_bw), len(beads), _p(d_cent), nb, 1.0,
              1e-3, kbt)
    fused = torch.zeros((len(beads), n_feat, n_feat), dtype=torch.float64, device="cuda")
    _lib.call("agf_gram_feat", _p(d_c), _p(d_f), code, n_frames, topo.n_sites, *common, _p(fused), _stream())
    _lib.call("agf_symmetrize_batch", _p(fused), n_feat, len(beads), _stream())
    packed = torch.zeros_like(fused)
    n_blocks = (n_feat + 127) // 128
    ws = torch.empty(slab_chunks * len(beads) * n_blocks * PANEL_BYTES, dtype=torch.uint8, device="cuda")
    _lib.call("agf_gram_feat_ws", _p(d_c), _p(d_f), code, n_frames, topo.n_sites, *common, _p(packed), _p(ws),
              C.c_size_t(ws.numel()), _stream())
    _lib.call("agf_symmetrize_batch", _p(packed), n_feat, len(beads), _stream())
    a, b = packed.cpu().numpy(), fused.cpu().numpy()
    assert np.isfinite(a).all()
    for bead in range(len(beads)):
        assert rel_fro(a[bead], b[bead]) < 1e-12
    # id-id block is the linear Gram of the label groups, identical for every bead
    cons = {frozenset(int(s) for s in np.nonzero(labels == g)[0]) for g in range(G) if (labels == g).sum() > 1}
    lin = oracle.gram_linear(forces, cons)
    order = oracle.group_columns(topo.n_sites, cons)
    perm = np.array([order[np.nonzero(labels == g)[0][0]] for g in range(G)])
    assert rel_fro(a[0][:G, :G], lin[np.ix_(perm, perm)]) < 1e-9


def test_large_fit_uses_device_solve_and_matches_oracle():
    """n_red >= 512: Gram stays on the device, exact solve through cuSOLVER; weights vs the oracle's
    host solve of the same problem (<= 1e-6 relative), and the singular case falls back to the host."""
    from aggforce_b200 import LinearMap, project_forces
    from aggforce_b200.synth import protein_like_topology, synth_trajectory_host

    topo = protein_like_topology(120)  # 1200 atoms -> n_red 624
    coords, forces = synth_trajectory_host(topo, 600, seed=4)
    cmap = LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
    cons = topo.xh_constraints
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=1e2)
    w = res["tmap"].force_map.standard_matrix
    ref = oracle.qp_linear_weights(forces, cmap.standard_matrix, cons, 1e2)
    assert rel_fro(w, ref) < 1e-6
    assert rel_fro(res["mapped_forces"], oracle.apply_map(forces, ref)) < 1e-6
    # 3 * 40 rows < n_red and no regularisation: P is singular -> host null-space solve
    res0 = project_forces(coords=coords[:40], forces=forces[:40], coord_map=cmap, constrained_inds=cons)
    w0 = res0["tmap"].force_map.standard_matrix
    assert np.isfinite(w0).all()
    assert np.abs(w0[:, topo.bead_atoms] @ np.eye(len(topo.bead_atoms)) - 0).shape == (len(topo.bead_atoms),) * 2
    ref0 = oracle.qp_linear_weights(forces[:40], cmap.standard_matrix, cons, 0.0)
    f0 = forces[:40]
    assert abs(np.mean(oracle.apply_map(f0, w0) ** 2) - np.mean(oracle.apply_map(f0, ref0) ** 2)) < 1e-6 * max(
        1.0, np.mean(oracle.apply_map(f0, ref0) ** 2))


@pytest.mark.parametrize("in_dtype", [np.float32, np.float64])
def test_slice_map_kernel(in_dtype):
    """agf_map_apply_slice (one site per bead): values, f32 maps, residual sum, NaN protocol."""
    from aggforce_b200 import LinearMap

    rng = np.random.default_rng(11)
    n_fg, beads = 60, [3, 17, 18, 59, 0]
    x = rng.normal(0, 20, size=(1031, n_fg, 3)).astype(in_dtype)
    lm = LinearMap([[b] for b in beads], n_fg_sites=n_fg)
    out, sumsq = lm.apply_with_sumsq(x)
    assert out.dtype == np.float64 and np.array_equal(out, x[:, beads, :].astype(np.float64))
    assert abs(sumsq / float((out ** 2).sum()) - 1) < 1e-12
    w = np.zeros((len(beads), n_fg))
    w[np.arange(len(beads)), beads] = [0.5, -2.0, 1.0, 3.0, 1e-9]
    scaled = LinearMap(w)(x)
    assert rel_fro(scaled, oracle.apply_map(x, w)) < 1e-14
    if in_dtype == np.float32:
        out32 = LinearMap(w.astype(np.float32))(x)
        assert out32.dtype == np.float32 and rel_fro(out32, oracle.apply_map(x, w)) < 1e-6
    xn = x.copy()
    xn[5, 7, :] = np.nan  # unreferenced site: ignored
    xn[9, 0, 1] = np.nan  # referenced with weight 1e-9 < atol: counts as 0
    clean = np.nan_to_num(xn, nan=0.0)
    assert rel_fro(LinearMap(w)(xn), oracle.apply_map(clean, w)) < 1e-14
    xn[11, 17, 2] = np.nan  # referenced with weight -2: the result depends on the NaN
    with pytest.raises(ValueError):
        LinearMap(w)(xn)
    plain = LinearMap(w, handle_nans=False)(xn)
    assert np.isnan(plain[11, :, 2]).all()  # numpy semantics: 0 * NaN = NaN in every bead


def test_config4_shape_properties():
    """BASELINE config 4 shape (5 000 atoms, 500 beads, n_red 2 600): oracle parity on a sub-sample
    and size-independent properties on more frames -- Gram additivity over frame shards, linearity
    of the packed-panel application, constraints recovered exactly."""
    from aggforce_b200 import LinearMap, _engine, guess_pairwise_constraints
    from aggforce_b200.qp.qplinear import force_gram, reduced_columns
    from aggforce_b200.synth import protein_like_topology, synth_trajectory_device

    topo = protein_like_topology(500)
    coords, forces = synth_trajectory_device(topo, 3000, seed=17)
    cons = topo.xh_constraints
    cols = reduced_columns(topo.n_sites, cons)
    n_red = int(cols.max()) + 1
    assert (topo.n_sites, n_red) == (5000, 2600)
    sub = forces[:48].cpu().numpy()
    g_sub, _ = force_gram(forces[:48], topo.n_sites, cons)
    assert rel_fro(g_sub, oracle.gram_linear(sub, cons)) < 1e-9
    whole, _ = force_gram(forces, topo.n_sites, cons)
    parts = sum(force_gram(forces[a:b], topo.n_sites, cons)[0] for a, b in [(0, 1001), (1001, 2050), (2050, 3000)])
    # 3 000 frames take the tiled tensor-core kernel (int8 digit planes, ~1e-11), the shards of ~1 000 frames the
    # FP64 DMMA kernel: additivity at the north-star bar across the two, at rounding level within the DMMA path
    assert rel_fro(whole, parts) < 1e-9 and np.array_equal(whole, whole.T)
    _engine._GRAM_I8[0] = False
    try:
        whole64, _ = force_gram(forces, topo.n_sites, cons)
    finally:
        _engine._GRAM_I8[0] = True
    assert rel_fro(whole64, parts) < 1e-12 and rel_fro(whole, whole64) < 1e-9
    rng = np.random.default_rng(2)
    w = rng.normal(size=(500, n_red))[:, cols]
    lm = LinearMap(w)
    out = lm(forces)  # 3 000 float32 frames: the int8 tensor-core application (csrc/apply_i8.cu), ~1e-11
    assert rel_fro(out[:16].cpu().numpy(), oracle.apply_map(forces[:16].cpu().numpy(), w)) < 1e-9
    assert rel_fro((-1.5 * lm)(forces).cpu().numpy(), -1.5 * out.cpu().numpy()) < 1e-9
    _engine._GRAM_I8[0] = False  # the FP64 DMMA GEMM on the same input: rounding level
    try:
        out64 = lm(forces)
        assert rel_fro(out64[:16].cpu().numpy(), oracle.apply_map(forces[:16].cpu().numpy(), w)) < 1e-12
        assert rel_fro((-1.5 * lm)(forces).cpu().numpy(), -1.5 * out64.cpu().numpy()) < 1e-12
    finally:
        _engine._GRAM_I8[0] = True
    assert rel_fro(out.cpu().numpy(), out64.cpu().numpy()) < 1e-9
    mapped, sumsq = lm.apply_with_sumsq(forces)
    assert abs(sumsq / float((mapped.double() ** 2).sum().item()) - 1) < 1e-12
    assert guess_pairwise_constraints(coords) == cons


# --------------------------------------------------------------------------------------
# device QP for small reduced problems (agf_qp_equality_small) and the one-read project_forces
# --------------------------------------------------------------------------------------
def _qp_small(gram_upper, diag, a_mat, x_index, u_index=None, n_ucol=0):
    import ctypes as C

    from aggforce_b200 import _engine, _lib

    n, m = gram_upper.shape[0], a_mat.shape[0]
    dev = "cuda"
    g = torch.as_tensor(gram_upper, device=dev).contiguous()
    d = None if diag is None else torch.as_tensor(diag, device=dev).contiguous()
    a = torch.as_tensor(a_mat, device=dev).contiguous()
    xi = torch.as_tensor(np.asarray(x_index, dtype=np.int32), device=dev)
    x = torch.full((m, n), float("nan"), dtype=torch.float64, device=dev)
    ui = None if u_index is None else torch.as_tensor(np.asarray(u_index, dtype=np.int32), device=dev)
    u = None if u_index is None else torch.full((n_ucol, m), float("nan"), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    p = _engine.ptr
    _lib.call("agf_qp_equality_small", p(g), n, p(d), p(a), m, p(xi), p(x), p(ui), p(u), p(status), _engine.stream_ptr())
    return x.cpu().numpy(), None if u is None else u.cpu().numpy(), int(status.item())


@pytest.mark.parametrize("n,m", [(97, 10), (6, 2), (128, 32), (33, 1)])
def test_qp_equality_small_matches_oracle(n, m):
    rng = np.random.default_rng(n * 100 + m)
    r = rng.normal(size=(3 * n + 5, n)) * rng.uniform(0.5, 300.0, size=n)
    gram = r.T @ r
    diag = rng.uniform(0.0, 1e3, size=n)
    a = np.zeros((m, n))
    a[np.arange(m), rng.choice(n, size=m, replace=False)] = 1.0
    a += 0.1 * rng.normal(size=(m, n)) * (rng.random((m, n)) < 0.2)
    want = oracle.solve_equality_qp(gram + np.diag(diag), a, np.eye(m)).T  # [m, n]
    upper = np.triu(gram) + np.tril(np.full((n, n), np.nan), k=-1)  # the kernel must not read below the diagonal
    perm = rng.permutation(n).astype(np.int32)
    uidx = rng.permutation(n).astype(np.int32)
    x, u, status = _qp_small(upper, diag, a, perm, uidx, n)
    assert status == 0
    got = np.empty_like(x)
    got[:, np.arange(n)] = x[:, perm]  # x_out[c, x_index[p]] = X[p, c]
    assert rel_fro(got, want) < 1e-9  # north star: weights within 1e-6
    assert np.abs(a @ got.T - np.eye(m)).max() < 1e-9
    assert np.array_equal(u[uidx].T, got)
    x0, _, status0 = _qp_small(np.triu(gram), None, a, np.arange(n))
    assert status0 == 0 and rel_fro(x0, oracle.solve_equality_qp(gram, a, np.eye(m)).T) < 1e-7


def test_qp_equality_small_reports_failures():
    rng = np.random.default_rng(0)
    n, m = 12, 3
    r = rng.normal(size=(4, n))  # rank 4 < n: P is singular
    a = np.zeros((m, n))
    a[np.arange(m), [0, 5, 9]] = 1.0
    _, _, status = _qp_small(np.triu(r.T @ r), None, a, np.arange(n))
    assert status != 0
    r = rng.normal(size=(40, n))
    a[2] = a[1]  # dependent equality rows: Schur complement singular
    _, _, status = _qp_small(np.triu(r.T @ r), None, a, np.arange(n))
    assert status != 0
    g = np.triu(r.T @ r)
    g[3, 7] = np.nan
    _, _, status = _qp_small(g, None, a, np.arange(n))
    assert status != 0


def test_project_forces_falls_back_to_the_host_solver_when_the_device_solve_declines():
    """Fewer frames than reduced columns and no l2: P is singular, agf_qp_equality_small reports it,
    the host solver's null-space branch answers and the forces are re-applied with ITS map."""
    from aggforce_b200 import LinearMap, project_forces, qp_linear_map
    from aggforce_b200.trajectory import Trajectory

    rng = np.random.default_rng(5)
    coords = rng.normal(size=(2, 9, 3)).astype(np.float32)
    forces = rng.normal(size=(2, 9, 3)).astype(np.float32)
    cmap = LinearMap([[0], [4]], n_fg_sites=9)
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=set())
    w = res["tmap"].force_map.standard_matrix
    ref = oracle.qp_linear_weights(forces, cmap.standard_matrix, set(), 0.0)
    assert np.abs(cmap.standard_matrix @ w.T - np.eye(2)).max() < 1e-9
    # the null-space solution maps these two frames to (numerically) zero forces: absolute bars
    assert np.abs(res["mapped_forces"] - oracle.apply_map(forces, w)).max() < 1e-12
    assert abs(res["residual"] - oracle.force_smoothness(oracle.apply_map(forces, w))) < 1e-20
    assert oracle.force_smoothness(oracle.apply_map(forces, ref)) < 1e-20
    # a direct call resolves (and falls back) before returning
    tm = qp_linear_map(Trajectory(coords=coords, forces=forces), cmap, constraints=set())
    assert np.abs(cmap.standard_matrix @ tm.force_map.standard_matrix.T - np.eye(2)).max() < 1e-9


def test_device_fit_is_lazy_and_editable():
    """project_forces keeps the fitted matrix on the device until somebody asks; an in-place edit of
    the downloaded matrix must still reach later applications."""
    from aggforce_b200 import LinearMap, project_forces

    rng = np.random.default_rng(6)
    coords = rng.normal(size=(300, 12, 3)).astype(np.float32)
    forces = rng.normal(size=(300, 12, 3)).astype(np.float32)
    cmap = LinearMap([[0], [4], [8]], n_fg_sites=12)
    cons = {frozenset({0, 1}), frozenset({4, 5})}
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=10.0)
    fm = res["tmap"].force_map
    w = oracle.qp_linear_weights(forces, cmap.standard_matrix, cons, 10.0)
    assert fm.n_cg_sites == 3 and fm.n_fg_sites == 12
    assert rel_fro(fm.standard_matrix, w) < 1e-9
    assert rel_fro(fm(forces), oracle.apply_map(forces, w)) < 1e-9
    fm.standard_matrix[1, 3] += 2.0
    assert rel_fro(fm(forces), oracle.apply_map(forces, fm.standard_matrix)) < 1e-9
    assert rel_fro((2.0 * fm).standard_matrix, 2.0 * fm.standard_matrix) < 1e-15
    # an edit made BEFORE the fitted map is applied for the first time must be seen too
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, l2_regularization=10.0)
    fm = res["tmap"].force_map
    fm.standard_matrix[2, 7] -= 1.5
    assert rel_fro(fm(forces), oracle.apply_map(forces, fm.standard_matrix)) < 1e-9
    # ... also for a large fit, whose coefficients stay on the device until somebody asks
    from aggforce_b200.synth import protein_like_topology, synth_trajectory_host

    topo = protein_like_topology(130)
    c2, f2 = synth_trajectory_host(topo, 64, seed=5)
    cm2 = LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
    big = project_forces(coords=c2, forces=f2, coord_map=cm2, constrained_inds=topo.xh_constraints,
                         l2_regularization=10.0)["tmap"].force_map
    assert big._matrix is None  # n_red >= 512: device solve, nothing downloaded yet
    m = big.standard_matrix
    m[0, 0] += 3.0
    assert rel_fro(big(f2), oracle.apply_map(f2, m)) < 1e-9


# --------------------------------------------------------------------------------------
# int8 / tcgen05 Gram (agf_gram_linear_i8): the Blackwell tensor-core path of kernel (a)
# --------------------------------------------------------------------------------------
def _gram_caller_order(forces, cols, n_red, use_i8):
    from aggforce_b200 import _engine

    prev = _engine._GRAM_I8[0]
    _engine._GRAM_I8[0] = use_i8
    try:
        g, order = _engine.gram_linear_raw(_engine.Frames(forces), cols, n_red)
    finally:
        _engine._GRAM_I8[0] = prev
    u = torch.triu(g).cpu().numpy()
    full = u + np.triu(u, 1).T
    out = np.empty_like(full)
    out[np.ix_(order, order)] = full
    return out


def test_gram_i8_matches_the_float64_oracle():
    """North-star bar for the Gram: 1e-9 relative Frobenius against float64 numpy.  The int8-sliced
    tensor-core kernel (5 signed 8-bit digits, 15 exact int32 products) sits near 1e-11."""
    from aggforce_b200 import _lib
    from aggforce_b200.qp.qplinear import reduced_columns
    from aggforce_b200.synth import chignolin_topology, synth_trajectory_host

    topo = chignolin_topology()
    cols = reduced_columns(175, topo.xh_constraints)
    n_red = int(cols.max()) + 1
    _, forces = synth_trajectory_host(topo, 20011, seed=21)  # odd count: head / tail chunks, odd sub-chunk totals
    ref = oracle.gram_linear(forces, topo.xh_constraints)
    n0 = _lib.LAUNCHES["count"]
    _lib.timing(True)
    got = _gram_caller_order(forces, cols, n_red, True)
    names = {n for n, _ in _lib.timing_records()}
    _lib.timing(False)
    assert "agf_gram_linear_i8" in names and _lib.LAUNCHES["count"] > n0
    assert rel_fro(got, ref) < 1e-9
    assert rel_fro(got, ref) < 1e-10  # what the scheme delivers on this data
    # exact integer accumulation across CTAs and launches: bit-identical from run to run
    for _ in range(3):
        assert np.array_equal(_gram_caller_order(forces, cols, n_red, True), got)
    assert rel_fro(_gram_caller_order(forces, cols, n_red, False), ref) < 1e-13  # the FP64 DMMA kernel
    # an unaligned device view (frames 3..) and a smaller system (n_red < 96: no ride-along column)
    dev = torch.as_tensor(forces, device="cuda")
    assert rel_fro(_gram_caller_order(dev[3:], cols, n_red, True), oracle.gram_linear(forces[3:], topo.xh_constraints)) < 1e-9
    sub = {c for c in topo.xh_constraints if max(c) < 100}
    scols = reduced_columns(100, sub)
    fsub = np.ascontiguousarray(forces[:, :100])
    assert rel_fro(_gram_caller_order(fsub, scols, int(scols.max()) + 1, True), oracle.gram_linear(fsub, sub)) < 1e-9


def test_gram_i8_hands_out_of_range_frames_to_the_float64_pass():
    """Values far outside the scale taken from the sample (and non-finite ones) must not be clipped: their
    frames are added exactly by the leftover kernel."""
    from aggforce_b200.qp.qplinear import reduced_columns
    from aggforce_b200.synth import chignolin_topology, synth_trajectory_host

    topo = chignolin_topology()
    cols = reduced_columns(175, topo.xh_constraints)
    n_red = int(cols.max()) + 1
    _, forces = synth_trajectory_host(topo, 12000, seed=22)
    forces[7000, 5, 1] = 3.0e7
    forces[9000, 100, 2] = -1.0e9
    forces[11999, 174, 0] = 4.0e8
    ref = oracle.gram_linear(forces, topo.xh_constraints)
    got = _gram_caller_order(forces, cols, n_red, True)
    assert rel_fro(got, ref) < 1e-12  # the three huge frames dominate and are exact
    small = ref.copy()
    mask = np.abs(ref) < 1e12  # entries the outliers do not touch: still at the 1e-9 bar
    assert np.abs(got - ref)[mask].max() < 1e-9 * np.abs(ref[mask]).max()
    forces[8000, 17, 0] = np.nan
    got = _gram_caller_order(forces, cols, n_red, True)
    want = _gram_caller_order(forces, cols, n_red, False)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).any()
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * np.abs(want[ok]).max()
    del small
