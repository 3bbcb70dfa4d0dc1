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
