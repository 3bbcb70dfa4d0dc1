"""CPU: the numpy oracle against the reference's golden vectors / recorded outputs.

The fixtures under tests/golden were produced by running the unmodified reference
(tests/golden/make_golden.py); this file is what "pins" the oracle.
"""
import json

import numpy as np
import pytest

import oracle
from conftest import pairs_to_set, rel_fro


def test_pair_sd_and_constraint_sets(small_cln):
    coords = small_cln["coords"]
    sds = oracle.pair_distance_sd(coords)
    assert np.allclose(sds, small_cln["sds"], rtol=1e-9, atol=1e-12)
    assert oracle.guess_pairwise_constraints(coords[:10]) == pairs_to_set(small_cln["cons10"])
    assert oracle.guess_pairwise_constraints(coords) == pairs_to_set(small_cln["cons_all"])
    # float32 evaluation of the reference agrees on this (well separated) data: Q10
    assert pairs_to_set(small_cln["cons10_f32"]) == pairs_to_set(small_cln["cons10"])
    cross = oracle.guess_pairwise_constraints(coords[:, :60], cross_xyz=coords[:, 40:90])
    assert cross == {(int(i), int(j)) for i, j in small_cln["cross_pairs"]}
    assert len(cross) > 0


def test_constraints_are_the_xh_bonds(small_cln):
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    assert pairs_to_set(small_cln["cons10"]) == topo.xh_constraints
    assert len(topo.xh_constraints) == 78


def test_gram_linear_matches_reference(small_cln):
    forces = small_cln["forces"]
    cons = pairs_to_set(small_cln["cons10"])
    assert rel_fro(oracle.gram_linear(forces, cons), small_cln["gram_raw"]) < 1e-13
    assert rel_fro(oracle.gram_linear(forces, ()), small_cln["gram_nocons"]) < 1e-13
    p = oracle.gram_linear(forces, cons) + 1e3 * oracle.l2_linear_term(forces.shape[1], cons)
    assert rel_fro(p, small_cln["gram_l2_1e3"]) < 1e-13
    assert p.shape == (97, 97)


def test_linear_weights_match_reference(small_cln):
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    cm = np.zeros((10, 175))
    cm[np.arange(10), topo.bead_atoms] = 1
    cons = pairs_to_set(small_cln["cons10"])
    w = oracle.qp_linear_weights(small_cln["forces"], cm, cons, 1e3)
    assert rel_fro(w, small_cln["W_l2_1e3"]) < 1e-8
    assert np.allclose(cm @ w.T, np.eye(10), atol=1e-10)
    mapped = oracle.apply_map(small_cln["forces"], w)
    assert rel_fro(mapped, small_cln["mapped_forces"]) < 1e-8
    assert abs(oracle.force_smoothness(mapped) / small_cln["residual"] - 1) < 1e-8
    assert rel_fro(oracle.apply_map(small_cln["coords"], cm), small_cln["mapped_coords"]) < 1e-12


def test_uni_map_matches_reference_goldens(small_cln, golden):
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    cm = np.zeros((10, 175))
    cm[np.arange(10), topo.bead_atoms] = 1
    uni = oracle.uni_map_matrix(cm, topo.xh_constraints)
    assert np.array_equal(uni, small_cln["uni_matrix"])
    # the reference's own golden file (tests/test_forces.py:154-157): exact
    assert ((uni - np.loadtxt(golden / "cln_basic_force_mat.txt")) ** 2).sum() < 1e-5
    assert uni.sum() == 21


def test_opt_golden_structure(golden):
    """cln_opt_force_mat.txt came from the real trajectory (not available): check the
    invariants any correct solution has -- it lies in range(C) for the X-H groups and
    satisfies the equality constraints."""
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    w = np.loadtxt(golden / "cln_opt_force_mat.txt")
    cols = oracle.group_columns(175, topo.xh_constraints)
    assert cols.max() + 1 == 97
    for g in range(97):
        members = np.nonzero(cols == g)[0]
        assert np.abs(w[:, members] - w[:, members[:1]]).max() < 1e-12
    cm = np.zeros((10, 175))
    cm[np.arange(10), topo.bead_atoms] = 1
    a = cm @ oracle.bond_constraint_matrix(175, topo.xh_constraints)
    x = np.stack([w[:, np.nonzero(cols == g)[0][0]] for g in range(97)], axis=1)
    assert np.abs(a @ x.T - np.eye(10)).max() < 1e-12


def test_waterdimer_known_answer(golden):
    """tests/test_agg.py:17-44 of the reference with the exact solve."""
    forces = np.load(golden / "waterdimer.npz")["Fs"]
    cm = np.zeros((2, 6))
    cm[0, 0] = cm[1, 3] = 1
    w = oracle.qp_linear_weights(forces, cm, (), 0.0)
    expected = np.array([[1, 1, 1, 0, 0, 0], [0, 0, 0, 1, 1, 1]], dtype=float)
    assert np.allclose(w, expected, atol=5e-3)


def test_id_feature_gram_matches_reference(small_cln, golden):
    ref = np.load(golden / "ref_idfeat.npz")
    feats, divs = oracle.id_features(small_cln["forces"].shape[0], ref["ids"])
    p = oracle.feat_gram(small_cln["forces"], feats, divs, float(ref["kbt"]), 1e1)
    # the reference builds this Gram in float32: agreement to f32 accuracy
    for bead in range(ref["P"].shape[0]):
        assert rel_fro(ref["P"][bead], p) < 5e-6
    from aggforce_b200.synth import chignolin_topology

    cm = np.zeros((10, 175))
    cm[np.arange(10), chignolin_topology().bead_atoms] = 1
    a, b = oracle.feat_constraint_rows(feats, cm, 3, ref["frame_choice"])
    assert np.array_equal(a, ref["A"][3]) and np.array_equal(b, ref["b"][3])
    coefs = [oracle.solve_equality_qp(p, *oracle.feat_constraint_rows(feats, cm, c, ref["frame_choice"]))
             for c in range(10)]
    assert rel_fro(np.stack(coefs), ref["coefs"]) < 1e-4
    mapped = oracle.feat_map_apply(small_cln["forces"], [feats] * 10, [divs] * 10, coefs)
    assert rel_fro(mapped, ref["mapped_forces"]) < 1e-4


def test_id_gram_equals_linear_gram(small_cln, golden):
    """One-hot id features collapse the atom contraction to the group sum."""
    ref = np.load(golden / "ref_idfeat.npz")
    feats, divs = oracle.id_features(small_cln["forces"].shape[0], ref["ids"])
    p = oracle.feat_gram(small_cln["forces"], feats, divs, 0.7, 0.0)
    g = oracle.gram_linear(small_cln["forces"], pairs_to_set(small_cln["cons10"]))
    cols = oracle.group_columns(175, pairs_to_set(small_cln["cons10"]))
    perm = np.array([ref["ids"][np.nonzero(cols == c)[0][0]] for c in range(97)])
    assert rel_fro(p[np.ix_(perm, perm)], g) < 1e-13


def test_gb_divergence_against_finite_differences():
    rng = np.random.default_rng(3)
    n, T = 9, 4
    x = rng.uniform(0, 6, size=(T, n, 3))
    cons = {frozenset((0, 1)), frozenset((1, 2)), frozenset((5, 6))}
    labels = oracle.canonical_labels(n, cons)
    cm = np.zeros((2, n))
    cm[0, 3] = 1.0
    cm[1, [0, 7]] = 0.5
    kw = dict(outer=8.0, inner=0.0, n_basis=7, width=1.0)
    for bead in (0, 1):
        for drop in (True, False):
            feats, divs = oracle.gb_features(x, cm, cons, labels, bead, drop_last_channel=drop, **kw)
            fd = oracle.gb_divergence_fd(x, cm, cons, labels, bead, drop_last_channel=drop, **kw)
            ok = np.isfinite(divs)
            assert np.abs(divs[ok] - fd[ok]).max() < 1e-6
            # an unconstrained bead atom sits on the bead: NaN there, finite elsewhere
            if bead == 1:
                assert ok.all()
            n_ch = labels.max() + (0 if drop else 1)
            assert feats.shape == (T, n, n_ch * 7) and divs.shape == (T, n_ch * 7, 3)
    mu = oracle.gb_centers(0.0, 8.0, 7)
    assert np.allclose(mu, 8 * np.arange(7) ** 2 / 36.0)


def test_condnormal_closed_form(golden):
    ref = np.load(golden / "ref_condnormal.npz")
    var = float(ref["var"])
    src, gen = ref["source"].astype(np.float64), ref["generated"].astype(np.float64)
    noise = (gen - src) / np.sqrt(var)
    eye = np.eye(src.shape[1])
    coords, forces = oracle.gauss_augment(src, np.zeros_like(src), eye, var, 1.0, noise)
    n = src.shape[1]
    assert np.allclose(coords[:, n:], gen, atol=1e-5)
    assert np.allclose(forces[:, :n], ref["lg_source"], rtol=2e-5, atol=2e-5)
    assert np.allclose(forces[:, n:], ref["lg_generated"], rtol=2e-5, atol=2e-5)


def test_merge_groups_matches_reference(golden):
    cases = json.loads((golden / "ref_sets.json").read_text())["cases"]
    for case in cases:
        merged = oracle.merge_constraint_groups(case["constraints"])
        assert sorted(list(g) for g in merged) == case["reduced"]


def test_apply_nan_protocol(golden):
    ref = np.load(golden / "ref_linearmap.npz")
    out = oracle.apply_map_nan_protocol(ref["pos_nan"], ref["slice_matrix"])
    assert np.allclose(out, ref["nan_out"], rtol=0, atol=1e-12)
    dense = ref["mat"]
    with pytest.raises(ValueError):
        oracle.apply_map_nan_protocol(ref["pos_nan"], dense)
    assert np.allclose(oracle.apply_map(ref["pos"], dense), ref["mapped"], rtol=1e-13)
