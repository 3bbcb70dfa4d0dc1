"""GPU: edge cases the reference's tests exercise implicitly (empty / ragged / mixed dtype inputs)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_fro

pytestmark = pytest.mark.gpu


def test_empty_trajectory():
    from aggforce_b200 import LinearMap, guess_pairwise_constraints
    from aggforce_b200.qp.qplinear import force_gram

    empty = np.zeros((0, 7, 3), dtype=np.float32)
    gram, cols = force_gram(empty, 7, {frozenset((1, 2))})
    assert gram.shape == (6, 6) and not gram.any()
    out = LinearMap(np.ones((2, 7)))(empty)
    assert out.shape == (0, 2, 3)
    assert guess_pairwise_constraints(empty) == set()


def test_single_frame_and_single_site():
    from aggforce_b200 import LinearMap, guess_pairwise_constraints
    from aggforce_b200.qp.qplinear import force_gram

    rng = np.random.default_rng(0)
    f = rng.normal(size=(1, 1, 3)).astype(np.float32)
    gram, _ = force_gram(f, 1, set())
    assert rel_fro(gram, oracle.gram_linear(f)) < 1e-12
    x = rng.normal(size=(1, 5, 3))
    # one frame: every distance has zero variance -> every pair is "constrained" (reference semantics)
    assert guess_pairwise_constraints(x) == oracle.guess_pairwise_constraints(x)
    assert rel_fro(LinearMap(np.ones((1, 5)))(x), oracle.apply_map(x, np.ones((1, 5)))) < 1e-13


def test_shape_errors_match_reference_types():
    from aggforce_b200 import LinearMap, Trajectory, project_forces

    lm = LinearMap(np.ones((2, 6)))
    with pytest.raises(ValueError):
        lm(np.zeros((3, 5, 3)))  # wrong number of sites
    with pytest.raises(ValueError):
        Trajectory(coords=np.zeros((3, 6, 3)), forces=np.zeros((3, 5, 3)))
    with pytest.raises(ValueError):
        project_forces(coords=None, forces=np.zeros((3, 6, 3)), coord_map=lm, constrained_inds="auto")


def test_mixed_dtypes_and_noncontiguous_views():
    from aggforce_b200 import LinearMap, guess_pairwise_constraints
    from aggforce_b200.qp.qplinear import force_gram

    rng = np.random.default_rng(1)
    big = rng.normal(size=(40, 30, 3))
    view = big[::2, 5:25]  # non-contiguous float64 view
    cons = {frozenset((0, 1)), frozenset((3, 4)), frozenset((4, 9))}
    gram, _ = force_gram(view, 20, cons)
    assert rel_fro(gram, oracle.gram_linear(view, cons)) < 1e-12
    m = rng.normal(size=(3, 20)).astype(np.float32)
    out = LinearMap(m)(view)  # f64 points x f32 matrix -> f64 (numpy promotion)
    assert out.dtype == np.float64 and rel_fro(out, oracle.apply_map(view, m)) < 1e-12
    x32, o64 = view.astype(np.float32), big[::2, :7].copy()
    got = guess_pairwise_constraints(x32, cross_xyz=o64, threshold=5.0)
    assert got == oracle.guess_pairwise_constraints(x32, cross_xyz=o64, threshold=5.0)


def test_many_members_per_group_uses_generic_paths():
    """Groups with more than 4 members (beyond the register-resident fast paths)."""
    from aggforce_b200 import LinearMap
    from aggforce_b200.qp.qplinear import force_gram

    rng = np.random.default_rng(2)
    n = 40
    cons = {frozenset(range(0, 7)), frozenset(range(10, 16)), frozenset((20, 21))}
    f = rng.normal(0, 10, size=(37, n, 3)).astype(np.float32)
    gram, cols = force_gram(f, n, cons)
    assert rel_fro(gram, oracle.gram_linear(f, cons)) < 1e-9
    red = rng.normal(size=(4, int(cols.max()) + 1))
    out = LinearMap(red[:, cols])(f)
    assert rel_fro(out, oracle.apply_map(f, red[:, cols])) < 1e-12


def test_device_and_host_paths_agree_bitwise():
    from aggforce_b200 import LinearMap
    from aggforce_b200.qp.qplinear import force_gram

    rng = np.random.default_rng(3)
    f = rng.normal(0, 50, size=(301, 33, 3)).astype(np.float32)
    m = rng.normal(size=(5, 33))
    host = LinearMap(m)(f)
    dev = LinearMap(m)(torch.as_tensor(f, device="cuda")).cpu().numpy()
    assert np.array_equal(host, dev)
    g_host, _ = force_gram(f, 33, set())
    g_dev, _ = force_gram(torch.as_tensor(f, device="cuda"), 33, set())
    assert rel_fro(g_host, g_dev) < 1e-14  # atomics: order may differ in the last bits


def test_constraints_second_pruning_point(monkeypatch):
    """A weak first screen (2 frames) leaves more than 4 n candidate pairs; the rescreen after a
    few hundred frames prunes them exactly (partial M2 never exceeds the total)."""
    from aggforce_b200 import _engine, guess_pairwise_constraints
    from aggforce_b200.synth import protein_like_topology, synth_trajectory_host

    topo = protein_like_topology(30)
    coords, _ = synth_trajectory_host(topo, 3000, seed=8)
    monkeypatch.setattr(_engine, "_SCREEN_FRAMES", 2)
    monkeypatch.setattr(_engine, "_RESCREEN_FRAMES", 256)
    got = guess_pairwise_constraints(coords)
    assert got == topo.xh_constraints == oracle.guess_pairwise_constraints(coords[:, :, :].astype(np.float64))


# --------------------------------------------------------------------------------------
# round-1 advisor findings
# --------------------------------------------------------------------------------------
def test_project_forces_hands_methods_the_callers_arrays():
    """Methods other than the built-in fast paths (user callables, generic featurizers) must see
    plain arrays -- not a proxy -- for numpy inputs and for CUDA tensors."""
    import torch

    from aggforce_b200 import LinearMap, project_forces
    from aggforce_b200.map import SeperableTMap

    rng = np.random.default_rng(0)
    coords = rng.normal(size=(30, 6, 3)).astype(np.float32)
    forces = rng.normal(size=(30, 6, 3)).astype(np.float32)
    cmap = LinearMap([[0, 1], [4]], n_fg_sites=6)
    seen = {}

    def method(traj, coord_map, constraints):
        seen["types"] = (type(traj.coords), type(traj.forces))
        doubled = traj.forces * 2  # scalar arithmetic must work on what we are handed
        neg = -traj.forces
        assert doubled.shape == neg.shape == traj.forces.shape
        return SeperableTMap(coord_map=coord_map, force_map=LinearMap(np.asarray(coord_map.standard_matrix) * 2))

    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=set(), method=method)
    assert seen["types"] == (np.ndarray, np.ndarray)
    assert np.allclose(res["mapped_forces"], 2 * oracle.apply_map(forces, cmap.standard_matrix))
    dc, df = torch.as_tensor(coords, device="cuda"), torch.as_tensor(forces, device="cuda")
    res = project_forces(coords=dc, forces=df, coord_map=cmap, constrained_inds=set(), method=method)
    assert seen["types"] == (torch.Tensor, torch.Tensor) and res["mapped_forces"].is_cuda


def test_staged_force_map_and_generic_featurizer_accept_cuda_tensors():
    """The paths the proxy used to break: np.zeros_like(traj.forces) in the staged force map and a
    user featurizer inside project_forces."""
    import torch

    from aggforce_b200 import LinearMap, project_forces
    from aggforce_b200.qp import id_feat, qp_feat_linear_map

    rng = np.random.default_rng(1)
    coords = (rng.normal(size=(40, 8, 3)) * 3).astype(np.float32)
    forces = rng.normal(size=(40, 8, 3)).astype(np.float32)
    cmap = LinearMap([[0], [5]], n_fg_sites=8)
    cons = {frozenset({0, 1})}

    def custom(points, cm, constraints):
        assert isinstance(points, (np.ndarray, torch.Tensor))
        return id_feat(points, cm, constraints)

    frames = np.arange(20)
    a = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, method=qp_feat_linear_map,
                       featurizer=custom, kbt=0.6, l2_regularization=1.0, constraint_frames=frames)
    b = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, method=qp_feat_linear_map,
                       featurizer=id_feat, kbt=0.6, l2_regularization=1.0, constraint_frames=frames)
    assert rel_fro(a["mapped_forces"], b["mapped_forces"]) < 1e-9


def test_paired_pieces_share_one_frame_schedule(monkeypatch):
    """coords on the device + forces on the host (and two host arrays of different dtypes) used to be
    cut at different frame boundaries; the featurised Gram must not depend on where the arrays live."""
    import torch

    from aggforce_b200 import LinearMap, _engine
    from aggforce_b200.qp import id_feat
    from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable

    rng = np.random.default_rng(2)
    T, n = 300, 10
    coords = (rng.normal(size=(T, n, 3)) * 3).astype(np.float32)
    forces = rng.normal(size=(T, n, 3)).astype(np.float32)
    cmap = LinearMap([[0], [5]], n_fg_sites=n)
    ctx = _FusedContext(cmap, {frozenset({0, 1})}, _fusable(id_feat))
    ref = ctx.grams(_engine.Frames(torch.as_tensor(coords, device="cuda")),
                    _engine.Frames(torch.as_tensor(forces, device="cuda")), 0.6)
    monkeypatch.setattr(_engine, "_PIECE_BYTES", 64 * n * 12)  # 64 float32 frames per piece
    monkeypatch.setattr(_engine, "_RESIDENT_FRACTION", 0.0)  # host arrays are streamed, never cached whole
    cases = [
        (_engine.Frames(torch.as_tensor(coords, device="cuda")), _engine.Frames(forces)),
        (_engine.Frames(coords), _engine.Frames(torch.as_tensor(forces, device="cuda"))),
        (_engine.Frames(coords), _engine.Frames(forces.astype(np.float64))),  # 64 vs 32 frames per piece
        (_engine.Frames(coords), _engine.Frames(forces)),
    ]
    for c, f in cases:
        starts = [(t0, pc.shape[0], pf.shape[0]) for t0, pc, pf in _engine.paired_pieces(c, f)]
        assert all(a == b for _, a, b in starts) and sum(a for _, a, _ in starts) == T
        assert rel_fro(ctx.grams(c, f, 0.6), ref) < 1e-12
        out, _ = ctx.apply(c, f, np.ones((2, ref.shape[1])))
        out_ref, _ = ctx.apply(cases[-1][0], cases[-1][1], np.ones((2, ref.shape[1])))
        assert torch.equal(out, out_ref)
    with pytest.raises(ValueError):
        list(_engine.paired_pieces(_engine.Frames(coords), _engine.Frames(forces[:-1])))


def test_streamed_pieces_survive_block_reuse(monkeypatch):
    """Streaming path (array not kept resident): staging blocks are recycled while earlier kernels may
    still be reading them; the result must equal the resident path."""
    from aggforce_b200 import LinearMap, _engine

    rng = np.random.default_rng(3)
    T, n = 4000, 50
    x = rng.normal(size=(T, n, 3)).astype(np.float32)
    lm = LinearMap(rng.normal(size=(7, n)))
    want = lm(x)
    monkeypatch.setattr(_engine, "_PIECE_BYTES", 128 * n * 12)
    monkeypatch.setattr(_engine, "_RESIDENT_FRACTION", 0.0)
    for _ in range(3):
        assert np.array_equal(lm(x), want)


def test_in_place_edit_of_a_large_map_is_seen():
    """The reference re-reads standard_matrix on every call (core.py:240); an edit of ANY element of
    a large matrix must reach the device copy (round 1 hashed a strided sample only)."""
    from aggforce_b200 import LinearMap

    rng = np.random.default_rng(4)
    mat = rng.normal(size=(600, 700))
    x = rng.normal(size=(5, 700, 3))
    lm = LinearMap(mat)
    first = lm(x)
    lm.standard_matrix[301, 377] += 1.0  # neither index is a multiple of the old 2-element stride
    second = lm(x)
    assert rel_fro(second, oracle.apply_map(x, lm.standard_matrix)) < 1e-12
    assert np.abs(second - first).max() > 1e-3
    frozen = rng.normal(size=(600, 700))
    frozen.setflags(write=False)
    lf = LinearMap(frozen)
    assert rel_fro(lf(x), oracle.apply_map(x, frozen)) < 1e-12 and lf._frozen_digest is not None
