"""CPU: host-side logic of the package (no kernel launches) and the C ABI surface."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from aggforce_b200 import _lib

    header = (ROOT / "include" / "agf_b200.h").read_text()
    declared = set(re.findall(r"\b(agf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    handle = _lib.lib()
    for name in declared:
        assert getattr(handle, name) is not None
    assert handle.agf_version() == 100
    assert isinstance(handle, ctypes.CDLL)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aggforce_b200 import LinearMap, _lib

    lm = LinearMap(np.eye(3))
    with pytest.raises(_lib.AgfError):
        lm(np.zeros((2, 3, 3)))


def test_package_never_imports_the_oracle():
    for path in (ROOT / "aggforce_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_reduce_and_lookup_match_reference(golden):
    from aggforce_b200.constraints import constraint_lookup_dict, reduce_constraint_sets

    cases = json.loads((golden / "ref_sets.json").read_text())["cases"]
    for case in cases:
        cons = {frozenset(g) for g in case["constraints"]}
        red = reduce_constraint_sets(cons)
        assert sorted(sorted(g) for g in red) == case["reduced"]
        assert {str(k): v for k, v in constraint_lookup_dict(red).items()} == case["lookup"]
    assert reduce_constraint_sets({frozenset((1, 2)), frozenset((2, 3)), frozenset((4, 5))}) == {
        frozenset((1, 2, 3)), frozenset((4, 5))}
    assert reduce_constraint_sets(set()) == set()


def test_reduced_columns_and_constraint_matrix():
    from aggforce_b200.qp import make_bond_constraint_matrix
    from aggforce_b200.qp.qplinear import reduced_columns

    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(3, 80))
        cons = {frozenset(int(v) for v in rng.choice(n, size=2, replace=False)) for _ in range(int(rng.integers(0, n)))}
        assert np.array_equal(reduced_columns(n, cons), oracle.group_columns(n, cons))
        assert np.array_equal(make_bond_constraint_matrix(n, cons), oracle.bond_constraint_matrix(n, cons))
    m = make_bond_constraint_matrix(5, {frozenset((1, 2))})
    assert np.array_equal(m, np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]]))


def test_uni_map_host(golden, small_cln):
    from aggforce_b200 import LinearMap, constraint_aware_uni_map
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    cmap = LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=175)
    tm = constraint_aware_uni_map(None, cmap, topo.xh_constraints)
    assert np.array_equal(tm.force_map.standard_matrix, small_cln["uni_matrix"])
    assert ((tm.force_map.standard_matrix - np.loadtxt(golden / "cln_basic_force_mat.txt")) ** 2).sum() < 1e-5


def test_linearmap_construction_and_algebra(golden):
    from aggforce_b200 import LinearMap

    ref = np.load(golden / "ref_linearmap.npz")
    lst = LinearMap([[0, 2, 3], [4]], n_fg_sites=6)
    assert np.array_equal(lst.standard_matrix, ref["list_matrix"])
    assert lst.n_cg_sites == 2 and lst.n_fg_sites == 6
    assert [list(map(int, g)) for g in lst.participating_fg] == [[0, 2, 3], [4]]
    lm = LinearMap(ref["mat"])
    assert np.array_equal(lm.T.standard_matrix, ref["transpose"])
    assert np.array_equal((2.5 * lm).standard_matrix, ref["scaled"])
    assert np.array_equal((lm + lm).standard_matrix, ref["summed"])
    assert np.array_equal((LinearMap(ref["mat"][:, :5]) @ lm).standard_matrix, ref["composed"])
    assert (lm.astype(np.float32).standard_matrix == ref["mat"].astype(np.float32)).all()
    with pytest.raises(ValueError):
        LinearMap(ref["mat"], n_fg_sites=15)
    with pytest.raises(ValueError):
        LinearMap([[0]])
    with pytest.raises(ValueError):
        LinearMap(np.array([[np.nan, 1.0]]))
    LinearMap(np.array([[np.nan, 1.0]]), handle_nans=False)
    with pytest.raises(ValueError):
        lm.flat_call(np.zeros((2, 5, 3)))
    with pytest.raises(ValueError):
        lm.flat_call(np.zeros((2, 44)))
    assert LinearMap(np.eye(4)).close_to_identity() and not lm.close_to_identity()


def test_trajectory_containers():
    from aggforce_b200 import Trajectory
    from aggforce_b200.trajectory import AugmentedTrajectory, CoordsTrajectory, SimpleCondNormal

    c, f = np.zeros((5, 4, 3), dtype=np.float32), np.ones((5, 4, 3), dtype=np.float32)
    t = Trajectory(coords=c, forces=f)
    assert len(t) == 5 and t.n_sites == 4 and t.n_dim == 3 and len(t[1:3]) == 2
    with pytest.raises(ValueError):
        t[0]
    with pytest.raises(ValueError):
        Trajectory(coords=c, forces=f[:, :3])
    with pytest.raises(ValueError):
        CoordsTrajectory(coords=np.zeros((3, 3)))
    assert t.astype(np.float64).coords.dtype == np.float64
    aug = AugmentedTrajectory.from_trajectory(t, kbt=0.6, augmenter=SimpleCondNormal(var=0.5, seed=1))
    assert aug.coords.shape == (5, 8, 3) and aug.n_real_sites == 4 and aug.n_aug_sites == 4
    y = aug.coords[:, aug.aug_slice]
    assert np.allclose(aug.forces[:, aug.aug_slice], -0.6 * (y - c) / 0.5, atol=1e-5)
    assert np.allclose(aug.forces[:, aug.real_slice], f + 0.6 * (y - c) / 0.5, atol=1e-5)
    assert aug[0:2].coords.shape == (2, 8, 3)


def test_condnormal_host_matches_reference(golden):
    from aggforce_b200.trajectory import SimpleCondNormal

    ref = np.load(golden / "ref_condnormal.npz")
    aug = SimpleCondNormal(var=float(ref["var"]), seed=5)
    gen = aug.sample(ref["source"])
    assert np.array_equal(gen, ref["generated"])
    a, b = aug.log_gradient(ref["source"], gen)
    assert np.array_equal(a, ref["lg_source"]) and np.array_equal(b, ref["lg_generated"])


def test_exact_qp_solver():
    from aggforce_b200.qp import solve_equality_qp

    rng = np.random.default_rng(2)
    m = rng.normal(size=(40, 12))
    p = m.T @ m
    a = rng.normal(size=(3, 12))
    b = np.eye(3)
    x = solve_equality_qp(p, a, b)
    assert np.allclose(a @ x, b, atol=1e-10)
    assert np.allclose(x, oracle.solve_equality_qp(p, a, b), rtol=1e-9, atol=1e-12)
    # stationarity: P x is in the row space of A
    lam = np.linalg.lstsq(a.T, p @ x, rcond=None)[0]
    assert np.allclose(a.T @ lam, p @ x, atol=1e-8)
    # singular P (fewer rows than unknowns): still a feasible minimiser
    m2 = rng.normal(size=(5, 12))
    x2 = solve_equality_qp(m2.T @ m2, a, b[:, 0])
    assert np.allclose(a @ x2, b[:, 0], atol=1e-8)
    assert solve_equality_qp(p, np.zeros((1, 12)), np.ones(1)) is None


def test_curry_and_flatten():
    from aggforce_b200.util import Curry, curry, flatten

    def f(x, y, z=0):
        return (x, y, z)

    assert Curry(f, 2, z=3)(1) == (1, 2, 3) and curry(f, 2, z=3)(1) == (1, 2, 3)
    c = Curry(f, z=5)
    assert c.func is f and c.kwargs == {"z": 5} and c.args == ()
    assert "Curry" in repr(c)
    assert flatten([[1, 2], [3, 4]]) == [1, 2, 3, 4]


def test_chunk_schedule_properties():
    """Mirror of csrc/frame_pipe.cuh make_schedule: chunks tile [0, T) exactly once."""
    def schedule(addr, n_frames, frame_bytes, kf):
        head = 0
        while head < 16 and head < n_frames and (addr + head * frame_bytes) % 16:
            head += 1
        rest = n_frames - head
        chunks = [(0, head)]
        for c in range((rest + kf - 1) // kf):
            s = head + c * kf
            chunks.append((s, min(kf, n_frames - s)))
        return chunks

    for n_sites in (1, 6, 175, 176, 5000):
        for off in range(0, 7):
            for T in (0, 1, 3, 16, 17, 100):
                fb = n_sites * 12
                ch = schedule(4096 + off * fb, T, fb, 16)
                covered = sum(c for _, c in ch)
                assert covered == T
                for s, c in ch[1:]:
                    assert (4096 + off * fb + s * fb) % 16 == 0 or c == 0 or s == ch[1][0] and ch[0][1] >= 16


def test_id_labels_match_reference_order(golden):
    """the label ORDER (which decides the dropped gb channel, Q5) equals the reference's."""
    from aggforce_b200 import LinearMap
    from aggforce_b200.qp import id_feat
    from aggforce_b200.synth import chignolin_topology

    ref = json.loads((golden / "ref_sets.json").read_text())
    for case in ref["cases"]:
        cons = {frozenset(g) for g in case["constraints"]}
        lab = id_feat(None, LinearMap([[0]], n_fg_sites=case["n"]), cons, return_ids=True)
        assert lab.dtype == np.int32 and [int(v) for v in lab] == case["ref_labels"]
    topo = chignolin_topology()
    cmap = LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=175)
    assert [int(v) for v in id_feat(None, cmap, topo.xh_constraints, return_ids=True)] == ref["cln_ids"]
    out = id_feat(np.zeros((3, 175, 3)), cmap, topo.xh_constraints)
    assert out["feats"][0].shape == (3, 175, 97) and out["feats"][0].dtype == np.float32
    assert out["feats"][0] is out["feats"][9] and out["divs"][0].shape == (3, 97, 3) and out["names"] is None
    assert (out["feats"][0].sum(axis=2) == 1).all()


def test_featzipper_and_multifeaturize():
    from aggforce_b200.qp import FeatZipper, Multifeaturize

    def f1(points, cmap, cons):
        return {"feats": (np.full((2, 3, 1), b) for b in range(2)), "divs": (np.zeros((2, 1, 3)) for _ in range(2)),
                "names": None}

    def f2(points, cmap, cons):
        return {"feats": [np.full((2, 3, 2), 10 + b) for b in range(2)], "divs": [np.ones((2, 2, 3))] * 2,
                "names": None}

    z = Multifeaturize([f1, f2])(None, None, None)
    assert isinstance(z, FeatZipper) and z["names"] is None and set(z.keys()) == {"feats", "divs", "names"}
    feats = list(z["feats"])
    divs = list(z["divs"])
    assert len(feats) == 2 and feats[1].shape == (2, 3, 3) and (feats[1][..., 0] == 1).all() and (feats[1][..., 1:] == 11).all()
    assert divs[0].shape == (2, 3, 3)
    with pytest.raises(KeyError):
        z["nope"]
    assert "Multifeaturize" in str(Multifeaturize([f1])) and "C0:" in repr(Multifeaturize([f1]))


def test_fusable_featurizer_recognition():
    import functools

    from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat
    from aggforce_b200.qp.featlinearmap import _fusable
    from aggforce_b200.util import Curry

    plan = _fusable(Multifeaturize([id_feat, Curry(gb_feat, inner=0, outer=8, width=1, n_basis=7)]))
    assert plan is not None and [b[0] for b in plan.blocks] == ["id", "gb"]
    assert plan.spec.n_basis == 7 and plan.spec.drop_last_channel and plan.spec.n_channels(97) == 96
    assert np.allclose(plan.spec.centers(), oracle.gb_centers(0, 8, 7))
    assert [b[0] for b in _fusable(Multifeaturize([functools.partial(gb_feat, outer=6.0), id_feat])).blocks] == ["gb", "id"]
    assert _fusable(id_feat) is not None
    assert _fusable(lambda p, c, k: None) is None
    assert _fusable(Multifeaturize([id_feat, id_feat])) is None
    assert _fusable(Curry(gb_feat, 8.0)) is None  # positional args are not introspected


def test_cv_helpers_match_reference_semantics():
    from aggforce_b200.agg import mean, process_cvargs, sample_sd

    grid = process_cvargs({"l2_regularization": [1.0, 10.0], "n_folds_unused": ["a"]})
    assert [g[1] for g in grid] == [{"l2_regularization": 1.0, "n_folds_unused": "a"},
                                    {"l2_regularization": 10.0, "n_folds_unused": "a"}]
    assert grid[0][0]._fields == ("l2_regularization", "n_folds_unused") and grid[1][0].l2_regularization == 10.0
    assert mean([]) is None and sample_sd([]) is None
    assert mean([1.0, 2.0, 6.0]) == 3.0
    assert abs(sample_sd([1.0, 2.0, 6.0]) - np.std([1.0, 2.0, 6.0], ddof=1)) < 1e-15


def test_ignored_solver_options_are_reported_once():
    import warnings

    from aggforce_b200.qp import solver

    solver._WARNED[0] = False
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        solver.warn_if_solver_ignored(solver.DEFAULT_SOLVER_OPTIONS)  # the default dictionary: silent
        solver.warn_if_solver_ignored({"backend": "qpsolvers", "solver": "scs"})  # honoured: silent
        assert not seen
        solver.warn_if_solver_ignored({"solver": "scs"})
        solver.warn_if_solver_ignored({"solver": "scs"})
    assert len(seen) == 1 and "qpsolvers" in str(seen[0].message)
    solver._WARNED[0] = False
