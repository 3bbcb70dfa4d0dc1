"""GPU parity: the CUDA path (Python API -> C ABI -> sm_100a kernels) against the numpy oracle
and the reference's golden vectors.  Tolerances are BASELINE.json's: constraint sets
identical, Grams <= 1e-9 relative Frobenius, weights / mapped forces <= 1e-6 relative.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import pairs_to_set, rel_fro

pytestmark = pytest.mark.gpu

GRAM_TOL = 1e-9
MAP_TOL = 1e-6


def _cmap(topo, **kw):
    from aggforce_b200 import LinearMap

    return LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites, **kw)


def _slice_matrix(topo):
    cm = np.zeros((len(topo.bead_atoms), topo.n_sites))
    cm[np.arange(len(topo.bead_atoms)), topo.bead_atoms] = 1
    return cm


@pytest.fixture(scope="module")
def topo():
    from aggforce_b200.synth import chignolin_topology

    return chignolin_topology()


# ------------------------------------------------------------------ kernel (a)
@pytest.mark.parametrize("n_frames", [1, 3, 16, 47, 48, 200, 1025])
def test_gram_linear_cln(topo, n_frames):
    from aggforce_b200.qp.qplinear import force_gram
    from aggforce_b200.synth import synth_trajectory_host

    _, forces = synth_trajectory_host(topo, n_frames, seed=5 + n_frames)
    gram, cols = force_gram(forces, topo.n_sites, topo.xh_constraints)
    assert np.array_equal(cols, oracle.group_columns(topo.n_sites, topo.xh_constraints))
    ref = oracle.gram_linear(forces, topo.xh_constraints)
    assert gram.shape == (97, 97)
    assert np.array_equal(gram, gram.T)
    assert rel_fro(gram, ref) < GRAM_TOL


def test_gram_golden_reference(small_cln):
    from aggforce_b200.qp.qplinear import force_gram

    cons = pairs_to_set(small_cln["cons10"])
    gram, _ = force_gram(small_cln["forces"], 175, cons)
    assert rel_fro(gram, small_cln["gram_raw"]) < GRAM_TOL
    gram0, _ = force_gram(small_cln["forces"], 175, set())
    assert rel_fro(gram0, small_cln["gram_nocons"]) < GRAM_TOL


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("offset", [0, 1, 2, 3, 5])
def test_gram_unaligned_views_and_dtypes(topo, dtype, offset):
    """Device tensors that start at any frame offset (TMA alignment head/tail handling)."""
    from aggforce_b200.qp.qplinear import force_gram
    from aggforce_b200.synth import synth_trajectory_host

    _, forces = synth_trajectory_host(topo, 150, seed=11)
    forces = forces.astype(dtype)
    dev = torch.as_tensor(forces, device="cuda")[offset:]
    gram, _ = force_gram(dev, topo.n_sites, topo.xh_constraints)
    assert rel_fro(gram, oracle.gram_linear(forces[offset:], topo.xh_constraints)) < GRAM_TOL


@pytest.mark.parametrize("n_sites,n_groups", [(6, 0), (40, 7), (130, 20), (300, 60), (777, 150)])
def test_gram_general_sizes(n_sites, n_groups):
    """n_red below / above one 128-column block, ragged last block, random constraint groups."""
    from aggforce_b200.qp.qplinear import force_gram

    rng = np.random.default_rng(n_sites)
    cons = set()
    for _ in range(n_groups):
        cons.add(frozenset(int(v) for v in rng.choice(n_sites, size=int(rng.integers(2, 5)), replace=False)))
    forces = rng.normal(0, 50.0, size=(77, n_sites, 3)).astype(np.float32)
    gram, cols = force_gram(forces, n_sites, cons)
    assert np.array_equal(cols, oracle.group_columns(n_sites, cons))
    assert rel_fro(gram, oracle.gram_linear(forces, cons)) < GRAM_TOL


def test_gram_is_additive_over_frame_shards(topo):
    """Size-independent property: Gram(all frames) == sum of Grams of disjoint shards."""
    from aggforce_b200.qp.qplinear import force_gram
    from aggforce_b200.synth import synth_trajectory_device

    _, forces = synth_trajectory_device(topo, 50_000, seed=3, want_coords=False)
    whole, _ = force_gram(forces, topo.n_sites, topo.xh_constraints)
    parts = sum(force_gram(forces[a:b], topo.n_sites, topo.xh_constraints)[0]
                for a, b in [(0, 12_345), (12_345, 30_001), (30_001, 50_000)])
    # (the int8 / tcgen05 kernel scales every call by its own sample: additive to its 1e-11, not to the last bit)
    assert rel_fro(whole, parts) < 1e-10
    sub = forces[:3000].cpu().numpy()
    assert rel_fro(force_gram(forces[:3000], topo.n_sites, topo.xh_constraints)[0],
                   oracle.gram_linear(sub, topo.xh_constraints)) < GRAM_TOL


# ------------------------------------------------------------------ fit + apply
def test_waterdimer_known_answer(golden):
    """Reference tests/test_agg.py: optimised O-only map sums each molecule's forces."""
    from aggforce_b200 import LinearMap, project_forces

    forces = np.load(golden / "waterdimer.npz")["Fs"]
    cmap = LinearMap([[0], [3]], n_fg_sites=6, handle_nans=False)
    coords = np.full_like(forces, np.nan)
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=set(),
                         solver_args={"solver": "scs"})
    w = res["tmap"].force_map.standard_matrix
    assert np.allclose(w, np.array([[1, 1, 1, 0, 0, 0], [0, 0, 0, 1, 1, 1]], dtype=float), atol=5e-3)
    assert rel_fro(w, oracle.qp_linear_weights(forces, cmap.standard_matrix, (), 0.0)) < MAP_TOL
    assert np.isnan(res["mapped_coords"]).all()  # handle_nans=False: NaN propagates like numpy
    assert rel_fro(res["mapped_forces"], oracle.apply_map(forces, w)) < MAP_TOL


def test_project_forces_matches_reference_outputs(topo, small_cln):
    from aggforce_b200 import project_forces

    cons = pairs_to_set(small_cln["cons10"])
    res = project_forces(coords=small_cln["coords"], forces=small_cln["forces"], coord_map=_cmap(topo),
                         constrained_inds=cons, l2_regularization=1e3)
    w = res["tmap"].force_map.standard_matrix
    assert rel_fro(w, small_cln["W_l2_1e3"]) < MAP_TOL
    assert rel_fro(res["mapped_forces"], small_cln["mapped_forces"]) < MAP_TOL
    assert rel_fro(res["mapped_coords"], small_cln["mapped_coords"]) < MAP_TOL
    assert res["mapped_forces"].dtype == np.float64
    assert abs(res["residual"] / float(small_cln["residual"]) - 1) < MAP_TOL
    assert res["constraints"] == cons


def test_project_forces_auto_constraints_and_uni_map(topo, small_cln, golden):
    from aggforce_b200 import constraint_aware_uni_map, project_forces

    res = project_forces(coords=small_cln["coords"], forces=small_cln["forces"], coord_map=_cmap(topo),
                         constrained_inds="auto", method=constraint_aware_uni_map)
    assert res["constraints"] == pairs_to_set(small_cln["cons_all"])
    assert np.array_equal(res["tmap"].force_map.standard_matrix, small_cln["uni_matrix"])
    assert ((res["tmap"].force_map.standard_matrix - np.loadtxt(golden / "cln_basic_force_mat.txt")) ** 2).sum() < 1e-5
    assert rel_fro(res["mapped_forces"], small_cln["uni_mapped_forces"]) < MAP_TOL
    assert abs(res["residual"] / float(small_cln["uni_residual"]) - 1) < MAP_TOL


def test_project_forces_device_resident(topo):
    """Torch CUDA tensors in -> CUDA tensors out, same numbers as the host path."""
    from aggforce_b200 import project_forces
    from aggforce_b200.synth import synth_trajectory_device

    coords, forces = synth_trajectory_device(topo, 4096, seed=9)
    res_d = project_forces(coords=coords, forces=forces, coord_map=_cmap(topo),
                           constrained_inds=topo.xh_constraints, l2_regularization=1e3)
    assert res_d["mapped_forces"].is_cuda
    c, f = coords.cpu().numpy(), forces.cpu().numpy()
    w = oracle.qp_linear_weights(f, _slice_matrix(topo), topo.xh_constraints, 1e3)
    assert rel_fro(res_d["tmap"].force_map.standard_matrix, w) < MAP_TOL
    assert rel_fro(res_d["mapped_forces"].cpu().numpy(), oracle.apply_map(f, w)) < MAP_TOL
    assert rel_fro(res_d["mapped_coords"].cpu().numpy(), oracle.apply_map(c, _slice_matrix(topo))) < 1e-12
    assert abs(res_d["residual"] / oracle.force_smoothness(oracle.apply_map(f, w)) - 1) < MAP_TOL


# ------------------------------------------------------------------ kernel (d)
def test_linearmap_reference_fixtures(golden):
    """Reference tests/test_linearmap.py data (seed 42100) and recorded outputs."""
    from aggforce_b200 import LinearMap

    ref = np.load(golden / "ref_linearmap.npz")
    lm = LinearMap(mapping=ref["mat"])
    assert rel_fro(lm(ref["pos"]), ref["mapped"]) < 1e-13
    out32 = lm(ref["pos"].astype(np.float32))
    assert out32.dtype == np.float64 and rel_fro(out32, ref["mapped_f32in"]) < 1e-13
    lm32 = lm.astype(np.float32)
    o = lm32(ref["pos"].astype(np.float32))
    assert o.dtype == np.float32
    assert np.sqrt(((o - ref["mapped"]) ** 2).sum()) / o.size < 1e-4  # reference test_linearmap.py:133-150
    flat = lm.flat_call(ref["pos"].reshape(20, 45))
    assert np.allclose(flat, ref["flat"]) and flat.shape == (20, 15)


def test_linearmap_nan_protocol(golden):
    from aggforce_b200 import LinearMap

    ref = np.load(golden / "ref_linearmap.npz")
    pos_nan = ref["pos_nan"].copy()
    sl = LinearMap(ref["slice_matrix"])
    out = sl(pos_nan)
    assert np.allclose(out, ref["nan_out"], rtol=0, atol=1e-12)
    assert np.isnan(pos_nan[:, [0, 3, 5]]).all()  # caller's array untouched
    with pytest.raises(ValueError):
        LinearMap(ref["mat"])(pos_nan)  # dense map touches the NaN sites
    plain = LinearMap(ref["slice_matrix"], handle_nans=False)(pos_nan)
    assert np.isnan(plain).all()  # numpy semantics: 0 * NaN = NaN in every bead


@pytest.mark.parametrize("n_cg,n_fg,n_frames", [(1, 5, 9), (10, 175, 333), (17, 64, 50), (64, 300, 41),
                                               (70, 90, 33), (300, 700, 19)])
@pytest.mark.parametrize("in_dtype", [np.float32, np.float64])
def test_map_apply_dense_shapes(n_cg, n_fg, n_frames, in_dtype):
    """Dense maps through the DMMA path (n_cg <= 64) and the large fallback; duplicated and
    zero columns exercise the unique-column compression."""
    from aggforce_b200 import LinearMap

    rng = np.random.default_rng(n_cg * 1000 + n_fg)
    m = rng.normal(size=(n_cg, n_fg))
    m[:, n_fg // 3] = m[:, 0]
    m[:, n_fg // 2] = 0.0
    x = rng.normal(0, 30, size=(n_frames, n_fg, 3)).astype(in_dtype)
    out = LinearMap(m)(x)
    assert out.shape == (n_frames, n_cg, 3) and out.dtype == np.float64
    assert rel_fro(out, oracle.apply_map(x, m)) < 1e-12


def test_map_apply_sumsq_and_big_stream(topo):
    from aggforce_b200 import LinearMap
    from aggforce_b200.synth import synth_trajectory_device

    _, forces = synth_trajectory_device(topo, 100_000, seed=21, want_coords=False)
    rng = np.random.default_rng(0)
    w = rng.normal(size=(10, 175))
    lm = LinearMap(w)
    mapped, sumsq = lm.apply_with_sumsq(forces)
    ref = oracle.apply_map(forces[:5000].cpu().numpy(), w)
    assert rel_fro(mapped[:5000].cpu().numpy(), ref) < 1e-12
    # linearity (size independent): map(a*F) = a*map(F); residual = mean of squares
    assert rel_fro((2.5 * lm)(forces).cpu().numpy(), 2.5 * mapped.cpu().numpy()) < 1e-12
    assert abs(sumsq / float((mapped.double() ** 2).sum().item()) - 1) < 1e-12


# ------------------------------------------------------------------ kernel (c)
def test_constraints_match_reference_sets(small_cln):
    from aggforce_b200 import guess_pairwise_constraints

    coords = small_cln["coords"]
    assert guess_pairwise_constraints(coords[0:10], threshold=1e-3) == pairs_to_set(small_cln["cons10"])
    assert guess_pairwise_constraints(coords) == pairs_to_set(small_cln["cons_all"])
    cross = guess_pairwise_constraints(coords[:, :60], cross_xyz=coords[:, 40:90])
    assert cross == {(int(i), int(j)) for i, j in small_cln["cross_pairs"]}


def test_constraints_threshold_straddle():
    """Adversarial: pair fluctuations placed just below / above the threshold."""
    from aggforce_b200 import guess_pairwise_constraints

    rng = np.random.default_rng(1)
    T, n = 400, 12
    x = rng.uniform(0, 20, size=(1, n, 3)) + rng.normal(0, 0.2, size=(T, n, 3))
    amp = {1: 0.9e-3, 3: 0.999e-3, 5: 1.001e-3, 7: 1.1e-3, 9: 0.0}
    phase = np.where(np.arange(T) % 2 == 0, 1.0, -1.0)  # sd of (1 + a*phase) is exactly a
    for j, a in amp.items():
        x[:, j] = x[:, j - 1] + np.array([1.0, 0, 0]) * (1.0 + a * phase)[:, None]
    got = guess_pairwise_constraints(x, threshold=1e-3)
    assert got == oracle.guess_pairwise_constraints(x, threshold=1e-3)
    assert frozenset((0, 1)) in got and frozenset((8, 9)) in got and frozenset((6, 7)) not in got
    assert guess_pairwise_constraints(np.full((5, 4, 3), np.nan)) == set()


def test_constraints_long_trajectory_with_pruning(topo):
    """Full-size style run: all frames streamed for the survivors; the set is the X-H bonds."""
    from aggforce_b200 import guess_pairwise_constraints
    from aggforce_b200.synth import synth_trajectory_device

    coords, _ = synth_trajectory_device(topo, 200_000, seed=4, want_forces=False)
    got = guess_pairwise_constraints(coords)
    assert got == topo.xh_constraints
    sub = coords[:600].cpu().numpy()
    assert guess_pairwise_constraints(coords[:600]) == oracle.guess_pairwise_constraints(sub)


# ------------------------------------------------------------------ C ABI error behaviour
def test_abi_rejects_bad_arguments():
    from aggforce_b200 import _lib

    with pytest.raises(_lib.AgfError):
        _lib.call("agf_gram_linear", None, 0, 10, 5, None, None, 3, None, None)
    assert "null pointer" in _lib.lib().agf_last_error().decode()


# ------------------------------------------------------------------ BASELINE-size properties
def test_full_size_properties_1m_frames(topo):
    """configs[1] size (1 M frames): size-independent checks -- additivity of the Gram over frame
    shards, the constraint set, the equality constraints of the fitted map, residual = mean square
    of the mapped forces, linearity of the application."""
    from aggforce_b200 import LinearMap, project_forces
    from aggforce_b200.qp.qplinear import force_gram
    from aggforce_b200.synth import synth_trajectory_device

    T = 1_000_000
    coords, forces = synth_trajectory_device(topo, T, seed=1234)
    cmap = _cmap(topo)
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds="auto",
                         l2_regularization=1e3)
    assert res["constraints"] == topo.xh_constraints
    w = res["tmap"].force_map.standard_matrix
    assert np.abs(_slice_matrix(topo) @ w.T - np.eye(10)).max() < 1e-9
    mf = res["mapped_forces"]
    assert mf.shape == (T, 10, 3)
    assert abs(res["residual"] / float((mf ** 2).mean().item()) - 1) < 1e-12
    whole, _ = force_gram(forces, topo.n_sites, topo.xh_constraints)
    halves = sum(force_gram(forces[a:b], topo.n_sites, topo.xh_constraints)[0]
                 for a, b in [(0, 500_001), (500_001, T)])
    assert rel_fro(whole, halves) < 1e-12
    sub = slice(123_456, 123_456 + 2_000)
    assert rel_fro(mf[sub].cpu().numpy(), oracle.apply_map(forces[sub].cpu().numpy(), w)) < MAP_TOL
    assert rel_fro(LinearMap(3.0 * w)(forces[sub]).cpu().numpy(), 3.0 * mf[sub].cpu().numpy()) < 1e-12


# ------------------------------------------------------------------ grid cross-validation
def test_grid_cv_one_pass_grams_match_per_fold_fits(topo):
    """Fast path (one Gram per fold, train = total - fold, hold-out score from the fold's Gram)
    against the reference's procedure (fit on the training frames, apply to the hold-out frames),
    and against a numpy restatement with the oracle's solver."""
    from aggforce_b200 import project_forces_grid_cv, qp_linear_map
    from aggforce_b200.synth import synth_trajectory_host

    coords, forces = synth_trajectory_host(topo, 600, seed=31)
    cmap, cons = _cmap(topo), topo.xh_constraints
    grid = {"l2_regularization": [1.0, 1e3]}
    fast = project_forces_grid_cv(grid, coords, forces, n_folds=4, rng=np.random.default_rng(5), coord_map=cmap,
                                  constrained_inds=cons)
    wrapped = lambda **kw: qp_linear_map(**kw)  # noqa: E731  (not `is qp_linear_map` -> generic path)
    slow = project_forces_grid_cv(grid, coords, forces, n_folds=4, rng=np.random.default_rng(5), coord_map=cmap,
                                  constrained_inds=cons, method=wrapped)
    assert list(fast["scores"]) == list(slow["scores"]) and len(fast["scores"]) == 2
    for key in fast["scores"]:
        assert key._fields == ("l2_regularization",)
        assert fast["n_runs"][key] == slow["n_runs"][key] == 4
        assert abs(fast["scores"][key] / slow["scores"][key] - 1) < 1e-9
        assert abs(fast["sds"][key] / slow["sds"][key] - 1) < 1e-6
    # numpy restatement of one grid point
    frames = np.arange(600)
    np.random.default_rng(5).shuffle(frames)
    folds = np.array_split(frames, 4)
    cm = _slice_matrix(topo)
    scores = []
    for i, val in enumerate(folds):
        train = np.concatenate([f for j, f in enumerate(folds) if j != i])
        w = oracle.qp_linear_weights(forces[train], cm, cons, 1e3)
        scores.append(np.mean(oracle.apply_map(forces[val], w) ** 2))
    key = [k for k in fast["scores"] if k.l2_regularization == 1e3][0]
    assert abs(fast["scores"][key] / np.mean(scores) - 1) < 1e-6
