"""GPU parity for the featurised path (kernel b), the Gaussian maps (config 5) and sharding helpers."""
import functools

import numpy as np
import pytest
import torch

import oracle
from conftest import pairs_to_set, rel_fro

pytestmark = pytest.mark.gpu

KBT = 0.6955215
GB = dict(outer=8.0, inner=0.0, n_basis=7, width=1.0)


@pytest.fixture(scope="module")
def topo():
    from aggforce_b200.synth import chignolin_topology

    return chignolin_topology()


@pytest.fixture(scope="module")
def data(topo):
    from aggforce_b200.synth import synth_trajectory_host

    return synth_trajectory_host(topo, 40, seed=77)


def _cmap(topo):
    from aggforce_b200 import LinearMap

    return LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)


def _oracle_features(coords, cm, cons, labels, bead, drop=True):
    idf, idd = oracle.id_features(coords.shape[0], labels)
    gf, gd = oracle.gb_features(coords, cm, cons, labels, bead, drop_last_channel=drop, **GB)
    return np.concatenate([idf, gf], axis=2), np.concatenate([idd, gd], axis=1)


def _featurizer():
    from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat
    from aggforce_b200.util import Curry

    return Multifeaturize([id_feat, Curry(gb_feat, **GB)])


def test_feat_gram_matches_oracle(topo, data):
    from aggforce_b200 import _engine
    from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable

    coords, forces = data
    cmap = _cmap(topo)
    cons = topo.xh_constraints
    ctx = _FusedContext(cmap, cons, _fusable(_featurizer()))
    assert ctx.n_groups == 97 and ctx.n_channels == 96 and len(ctx.columns) == 97 + 7 * 96 == 769
    grams = ctx.grams(_engine.Frames(coords), _engine.Frames(forces), KBT)
    assert grams.shape == (10, 769, 769)
    for bead in (0, 4, 9):
        feats, divs = _oracle_features(coords, cmap.standard_matrix, cons, ctx.labels, bead)
        ref = oracle.feat_gram(forces, feats, divs, KBT)
        assert rel_fro(grams[bead], ref) < 1e-9
        assert np.array_equal(grams[bead], grams[bead].T)
    # equality rows for a handful of frames
    frames = np.array([3, 17, 0, 39, 21])
    for bead in (2, 7):
        feats, _ = _oracle_features(coords, cmap.standard_matrix, cons, ctx.labels, bead)
        a_ref, _ = oracle.feat_constraint_rows(feats, cmap.standard_matrix, bead, frames)
        a = ctx.constraint_rows(_engine.Frames(coords), bead, frames)
        assert np.abs(a - a_ref).max() < 1e-12


def test_id_only_gram_matches_reference_recording(small_cln, golden, topo):
    """qp_feat_linear_map(id_feat): Gram recorded from the REFERENCE (float32 there)."""
    from aggforce_b200 import _engine
    from aggforce_b200.qp import id_feat
    from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable

    ref = np.load(golden / "ref_idfeat.npz")
    cons = pairs_to_set(small_cln["cons10"])
    ctx = _FusedContext(_cmap(topo), cons, _fusable(id_feat))
    assert np.array_equal(ctx.labels, ref["ids"])
    grams = ctx.grams(_engine.Frames(small_cln["coords"]), _engine.Frames(small_cln["forces"]), float(ref["kbt"]))
    for bead in range(10):
        assert rel_fro(grams[bead] + 10.0 * np.eye(97), ref["P"][bead]) < 5e-6
    a = ctx.constraint_rows(_engine.Frames(small_cln["coords"]), 3, ref["frame_choice"])
    assert np.array_equal(a, ref["A"][3])


def test_qp_feat_linear_map_id_matches_reference(small_cln, golden, topo):
    from aggforce_b200 import Trajectory
    from aggforce_b200.qp import id_feat, qp_feat_linear_map

    ref = np.load(golden / "ref_idfeat.npz")
    cons = pairs_to_set(small_cln["cons10"])
    traj = Trajectory(coords=small_cln["coords"], forces=small_cln["forces"])
    tmap = qp_feat_linear_map(traj, _cmap(topo), id_feat, float(ref["kbt"]), constraints=cons, l2_regularization=1e1,
                              constraint_frames=ref["frame_choice"])
    coefs = np.stack(tmap.force_map.tags["coef_list"])
    assert rel_fro(coefs, ref["coefs"]) < 1e-4  # reference Gram is float32
    mapped = tmap(traj)
    assert rel_fro(mapped.forces, ref["mapped_forces"]) < 1e-4
    assert rel_fro(mapped.coords, small_cln["mapped_coords"]) < 1e-12


@pytest.mark.parametrize("drop", [True, False])
def test_qp_feat_linear_map_gb_end_to_end(topo, data, drop):
    from aggforce_b200 import Trajectory, project_forces
    from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat, qp_feat_linear_map

    coords, forces = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    feat = Multifeaturize([id_feat, functools.partial(gb_feat, drop_last_channel=drop, **GB)])
    frames = np.random.default_rng(42100).choice(len(coords), size=20, replace=False)
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons,
                         method=qp_feat_linear_map, featurizer=feat, kbt=KBT, l2_regularization=1e3,
                         constraint_frames=frames)
    coefs = np.stack(res["tmap"].force_map.tags["coef_list"])
    labels = id_feat(None, cmap, cons, return_ids=True)
    cm = cmap.standard_matrix
    ref_coefs, fs, ds = [], [], []
    for bead in range(10):
        feats, divs = _oracle_features(coords, cm, cons, labels, bead, drop)
        p = oracle.feat_gram(forces, feats, divs, KBT, 1e3)
        a, b = oracle.feat_constraint_rows(feats, cm, bead, frames)
        ref_coefs.append(oracle.solve_equality_qp(p, a, b))
        fs.append(feats)
        ds.append(divs)
    assert rel_fro(coefs, np.stack(ref_coefs)) < 1e-6
    ref_mapped = oracle.feat_map_apply(forces, fs, ds, ref_coefs)
    assert rel_fro(res["mapped_forces"], ref_mapped) < 1e-6
    assert abs(res["residual"] / oracle.force_smoothness(ref_mapped) - 1) < 1e-6
    # the CLAMap's scale/trans (reference-style materialised weights) give the same map
    fm = res["tmap"].force_map
    t = Trajectory(coords=coords[:6], forces=forces[:6])
    from aggforce_b200.util import trjdot

    slow = trjdot(t.forces, fm.scale(t.coords)) + fm.trans(t.coords)
    assert rel_fro(slow, ref_mapped[:6]) < 1e-5  # float32 materialised features


def test_generic_featurizer_path_equals_fused(small_cln, golden, topo):
    from aggforce_b200 import Trajectory
    from aggforce_b200.qp import id_feat, qp_feat_linear_map

    ref = np.load(golden / "ref_idfeat.npz")
    cons = pairs_to_set(small_cln["cons10"])
    traj = Trajectory(coords=small_cln["coords"], forces=small_cln["forces"])

    def custom(points, cmap, constraints):  # not recognised -> generic device path
        return id_feat(points, cmap, constraints)

    kw = dict(kbt=0.7, constraints=cons, l2_regularization=5.0, constraint_frames=ref["frame_choice"])
    a = qp_feat_linear_map(traj, _cmap(topo), custom, **kw)
    b = qp_feat_linear_map(traj, _cmap(topo), id_feat, **kw)
    ca, cb = np.stack(a.force_map.tags["coef_list"]), np.stack(b.force_map.tags["coef_list"])
    assert rel_fro(ca, cb) < 1e-9
    assert rel_fro(a(traj).forces, b(traj).forces) < 1e-9


def test_gb_feat_direct_api(topo, data):
    from aggforce_b200.qp import gb_feat, id_feat

    coords, _ = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    out = gb_feat(coords[:5], cmap, cons, lazy=False, **GB)
    labels = id_feat(None, cmap, cons, return_ids=True)
    assert out["names"] is None and len(out["feats"]) == 10
    for bead in (0, 9):
        gf, gd = oracle.gb_features(coords[:5], cmap.standard_matrix, cons, labels, bead, **GB)
        assert out["feats"][bead].dtype == np.float32 and out["feats"][bead].shape == gf.shape
        assert np.abs(out["feats"][bead] - gf).max() < 1e-6
        assert np.abs(out["divs"][bead] - gd).max() < 1e-5


def test_joptgauss_with_injected_noise(topo, data):
    from aggforce_b200 import Trajectory, joptgauss_map

    coords, forces = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    var, kbt = 0.25, KBT
    noise = np.random.default_rng(5).standard_normal((len(coords), 10, 3))
    tmap = joptgauss_map(Trajectory(coords=coords, forces=forces), cmap, var=var, kbt=kbt, constraints=cons,
                         noise=noise, l2_regularization=1e2)
    w = tmap.tmap.force_map.standard_matrix
    assert w.shape == (10, 185)
    full_c, full_f = oracle.gauss_augment(coords, forces, cmap.standard_matrix, var, kbt, noise.astype(np.float32))
    aug_cm = np.zeros((10, 185))
    aug_cm[np.arange(10), 175 + np.arange(10)] = 1
    # the augmenter works in float32 like the reference's JCondNormal: compare on its own arrays
    ref_w = oracle.qp_linear_weights(full_f.astype(np.float32), aug_cm, cons, 1e2)
    assert rel_fro(w, ref_w) < 1e-4
    out = tmap(Trajectory(coords=coords, forces=forces))  # fresh noise: shapes / finiteness only
    assert out.coords.shape == (len(coords), 10, 3) and np.isfinite(out.forces).all()
    # noised coordinates scatter around the mapped coordinates with the requested variance
    resid = out.coords - oracle.apply_map(coords, cmap.standard_matrix)
    assert abs(resid.var() / var - 1) < 0.2


def test_condnormal_device_matches_closed_form(topo, data):
    from aggforce_b200.trajectory import AugmentedTrajectory, CondNormal, Trajectory

    coords, forces = data
    cmap = _cmap(topo)
    noise = np.random.default_rng(9).standard_normal((len(coords), 10, 3)).astype(np.float32)
    aug = CondNormal(cov=0.4, premap=cmap, noise=noise)
    at = AugmentedTrajectory.from_trajectory(Trajectory(coords=coords, forces=forces), kbt=0.6, augmenter=aug)
    rc, rf = oracle.gauss_augment(coords, forces, cmap.standard_matrix, 0.4, 0.6, noise)
    assert rel_fro(at.coords, rc) < 1e-6 and rel_fro(at.forces, rf) < 1e-6
    # sample / log_gradient pair (reference tests/test_simplegausstraj.py:13-29 style, atol 2e-6 relative)
    aug2 = CondNormal(cov=0.4, premap=cmap, noise=noise)
    y = aug2.sample(coords)
    gx, gy = aug2.log_gradient(coords, y)
    assert np.allclose(gy, -(y - oracle.apply_map(coords, cmap.standard_matrix)) / 0.4, atol=1e-4)
    assert np.allclose(gx[:, topo.bead_atoms], -gy, atol=1e-4)


def test_joptgauss_slabs_reproduce_the_same_noise(topo, data, monkeypatch):
    """The augmented arrays are generated slab by slab (never materialised): Philox noise keyed by
    (seed, global frame, bead, draw) makes the fit independent of the slab size, and the fused
    kernel agrees with the closed form on device-resident input."""
    from aggforce_b200 import Trajectory, joptgauss_map
    from aggforce_b200.trajectory import gausstraj

    coords, forces = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    traj = Trajectory(coords=coords, forces=forces)
    whole = joptgauss_map(traj, cmap, var=0.3, kbt=KBT, constraints=cons, seed=11, l2_regularization=1e2)
    monkeypatch.setattr(gausstraj, "_AUG_SLAB_BYTES", 185 * 12 * 40)  # 40 frames per slab
    slabbed = joptgauss_map(traj, cmap, var=0.3, kbt=KBT, constraints=cons, seed=11, l2_regularization=1e2)
    w0, w1 = whole.tmap.force_map.standard_matrix, slabbed.tmap.force_map.standard_matrix
    assert rel_fro(w1, w0) < 1e-10
    other = joptgauss_map(traj, cmap, var=0.3, kbt=KBT, constraints=cons, seed=12, l2_regularization=1e2)
    assert rel_fro(other.tmap.force_map.standard_matrix, w0) > 1e-6  # a different seed is a different draw
    # application: same draw for coordinates and forces, device tensors in -> device tensors out
    dc, df = torch.as_tensor(coords, device="cuda"), torch.as_tensor(forces, device="cuda")
    out = slabbed(Trajectory(coords=dc, forces=df))
    assert out.coords.is_cuda and tuple(out.forces.shape) == (len(coords), 10, 3)
    noise = np.random.default_rng(2).standard_normal((len(coords), 10, 3)).astype(np.float32)
    aug = gausstraj.CondNormal(cov=0.3, premap=cmap, noise=noise)
    draw = aug.new_draw()
    got_c = torch.cat([p for _, p in gausstraj.AugmentedFrames(dc, aug, KBT, draw, "coords").pieces()]).cpu().numpy()
    got_f = torch.cat([p for _, p in gausstraj.AugmentedFrames(df, aug, KBT, draw, "forces").pieces()]).cpu().numpy()
    rc, rf = oracle.gauss_augment(coords, forces, cmap.standard_matrix, 0.3, KBT, noise)
    assert rel_fro(got_c, rc) < 1e-6 and rel_fro(got_f, rf) < 1e-6


def test_staged_gauss_maps(topo, data):
    """Staged Gaussian maps (reference jgauss.py:143-650) against a numpy restatement of their
    composition (parity with the JAX reference itself is unpinned)."""
    import warnings

    from aggforce_b200 import Trajectory, stagedjforcegauss_map, stagedjoptgauss_map, stagedjslicegauss_map
    from aggforce_b200.trajectory import CoordsTrajectory

    coords, forces = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    cm = np.asarray(cmap.standard_matrix)
    T, n_cg, var, kbt = len(coords), 10, 0.25, KBT
    z = np.random.default_rng(3).standard_normal((T, n_cg, 3)).astype(np.float32)
    traj = Trajectory(coords=coords, forces=forces)
    staged = stagedjoptgauss_map(traj, cmap, var=var, kbt=kbt, constraints=cons, seed=1, premap_l2_regularization=1e2,
                                 noise=z, l2_regularization=1e1)
    pre, post = staged[1], staged[0]
    w0 = oracle.qp_linear_weights(forces, cm, cons, 1e2)
    assert rel_fro(pre.force_map.standard_matrix, w0) < 1e-6
    full_c, full_f = oracle.gauss_augment(coords, forces, cm, var, kbt, z)
    pm_f = np.concatenate([oracle.apply_map(full_f[:, :175], w0), full_f[:, 175:]], axis=1).astype(np.float32)
    sl = np.zeros((n_cg, 2 * n_cg))
    sl[np.arange(n_cg), n_cg + np.arange(n_cg)] = 1
    w1 = oracle.qp_linear_weights(pm_f, sl, set(), 1e1)
    got_w1 = post.tmap.force_map.standard_matrix
    assert got_w1.shape == (n_cg, 2 * n_cg) and rel_fro(got_w1, w1) < 1e-4
    # application with an injected draw: Y_final = W1 [W0 F + kbt/var P eps ; -kbt/var eps],  P = W0 A^T
    z2 = np.random.default_rng(4).standard_normal((T, n_cg, 3)).astype(np.float32)
    post.augmenter._noise = z2
    out = staged(traj)
    eps = np.sqrt(var) * z2.astype(np.float64)
    x_cg, y_cg = oracle.apply_map(coords, cm), oracle.apply_map(forces, w0)
    p = w0 @ cm.T
    f_aug = np.concatenate([y_cg + kbt / var * oracle.apply_map(eps, p), -kbt / var * eps], axis=1)
    assert rel_fro(out.coords, x_cg + eps) < 1e-5
    assert rel_fro(out.forces, oracle.apply_map(f_aug, w1)) < 1e-4
    # slice map: forces are the noise-site forces -kbt (y - A x) / var, also for force-free input
    sliced = stagedjslicegauss_map(CoordsTrajectory(coords=coords), cmap, var=var, kbt=kbt, seed=2)
    o = sliced(CoordsTrajectory(coords=coords))
    assert np.allclose(o.forces, -kbt * (o.coords - x_cg) / var, rtol=1e-3, atol=1e-3)
    assert abs((o.coords - x_cg).var() / var - 1) < 0.2
    # force map: W0 A^T = I, so the noise contributions cancel exactly with W1 = [I | I]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        fmap = stagedjforcegauss_map(traj, cmap, var=var, kbt=kbt, constraints=cons, seed=5,
                                     premap_l2_regularization=1e2)
    assert np.allclose(fmap[0].tmap.force_map.standard_matrix, np.hstack([np.eye(n_cg), np.eye(n_cg)]), atol=1e-4)
    of = fmap(traj)
    assert rel_fro(of.forces, y_cg) < 1e-4  # real mapped forces, noise forces cancelled


def test_project_forces_with_gaussian_methods(topo, data):
    """project_forces drives the noised maps like any other method (reference agg.py:120-128)."""
    from aggforce_b200 import joptgauss_map, project_forces, stagedjoptgauss_map

    coords, forces = data
    cmap, cons = _cmap(topo), topo.xh_constraints
    x_cg = oracle.apply_map(coords, cmap.standard_matrix)
    for method in (joptgauss_map, stagedjoptgauss_map):
        res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons, method=method,
                             var=0.2, kbt=KBT, seed=3, l2_regularization=1e2)
        assert res["mapped_coords"].shape == (len(coords), 10, 3) and np.isfinite(res["mapped_forces"]).all()
        assert abs((res["mapped_coords"] - x_cg).var() / 0.2 - 1) < 0.2
        assert abs(res["residual"] - np.mean(res["mapped_forces"] ** 2)) < 1e-9 * res["residual"]
        assert res["constraints"] == cons


# --------------------------------------------------------------------------------------
# featurised Gram on the tensor cores (agf_gram_feat_i8: int8 digit planes, batched tcgen05 SYRK)
# --------------------------------------------------------------------------------------
def _feat_grams(coords, forces, topo, use_i8, min_frames=None, featurizer=None):
    from aggforce_b200 import _engine, _lib
    from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable

    ctx = _FusedContext(_cmap(topo), topo.xh_constraints, _fusable(featurizer or _featurizer()))
    old = (_engine._GRAM_I8[0], _engine._GRAM_I8T_MIN_FRAMES)
    _engine._GRAM_I8[0] = use_i8
    if min_frames is not None:
        _engine._GRAM_I8T_MIN_FRAMES = min_frames
    try:
        _lib.timing(True)
        grams = ctx.grams(_engine.Frames(coords), _engine.Frames(forces), KBT)
        names = {n for n, _ in _lib.timing_records()}
        _lib.timing(False)
    finally:
        _engine._GRAM_I8[0], _engine._GRAM_I8T_MIN_FRAMES = old
    assert ("agf_gram_feat_i8" in names) == use_i8 and ("agf_gram_feat_ws" in names) != use_i8
    return grams, ctx


def test_feat_gram_i8_matches_oracle(topo, data):
    """40 frames through the int8 path (its frame threshold lowered) against the float64 oracle: the north-star bar."""
    coords, forces = data
    grams, ctx = _feat_grams(coords, forces, topo, True, min_frames=1)
    cmap = _cmap(topo)
    for bead in (0, 5, 9):
        feats, divs = _oracle_features(coords, cmap.standard_matrix, topo.xh_constraints, ctx.labels, bead)
        ref = oracle.feat_gram(forces, feats, divs, KBT)
        assert rel_fro(grams[bead], ref) < 1e-9
        assert np.array_equal(grams[bead], grams[bead].T)


def test_feat_gram_i8_agrees_with_the_dmma_kernel(topo):
    """3 001 frames (ragged chunk, several slices), float32 and float64 input, an out-of-range frame and a NaN."""
    from aggforce_b200.synth import synth_trajectory_host

    coords, forces = synth_trajectory_host(topo, 3001, seed=78)
    want, _ = _feat_grams(coords, forces, topo, False)
    got, _ = _feat_grams(coords, forces, topo, True)
    for bead in range(10):
        assert rel_fro(got[bead], want[bead]) < 1e-9
    got64, _ = _feat_grams(coords.astype(np.float64), forces.astype(np.float64), topo, True)
    assert rel_fro(got64, want) < 1e-9
    forces = forces.copy()
    forces[1500, 20, 1] = 4.0e8  # far outside the sampled scale: float64 pass
    want, _ = _feat_grams(coords, forces, topo, False)
    got, _ = _feat_grams(coords, forces, topo, True)
    assert rel_fro(got, want) < 1e-12  # the huge frame dominates and is exact
    mask = np.abs(want) < 1e-6 * np.abs(want).max()
    assert np.abs(got - want)[mask].max() < 1e-9 * np.abs(want[mask]).max()
    forces[2000, 3, 0] = np.nan
    want, _ = _feat_grams(coords, forces, topo, False)
    got, _ = _feat_grams(coords, forces, topo, True)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).any()


def test_condnormal_full_covariance(topo, data):
    """``JCondNormal`` accepts the full covariance of the flattened noise vector (jaxgausstraj.py:148-150):
    a diagonal matrix reproduces the scalar kernel for the same injected draw, a dense one follows the closed
    form (eps = L z, gradients through cov^-1), and the draw does not depend on the slab decomposition."""
    from aggforce_b200.trajectory import AugmentedTrajectory, CondNormal, Trajectory
    from aggforce_b200.trajectory.gausstraj import NoiseDraw

    coords, forces = data
    cmap = _cmap(topo)
    cm = cmap.standard_matrix
    rng = np.random.default_rng(10)
    noise = rng.standard_normal((len(coords), 10, 3))
    diag = CondNormal(cov=0.4 * np.eye(30), premap=cmap, noise=noise, dtype=np.float64)
    scal = CondNormal(cov=0.4, premap=cmap, noise=noise, dtype=np.float64)
    traj = Trajectory(coords=coords, forces=forces)
    a = AugmentedTrajectory.from_trajectory(traj, kbt=0.6, augmenter=diag)
    b = AugmentedTrajectory.from_trajectory(traj, kbt=0.6, augmenter=scal)
    assert rel_fro(a.coords, b.coords) < 1e-12 and rel_fro(a.forces, b.forces) < 1e-12
    # dense covariance against numpy
    q = rng.standard_normal((30, 30))
    sigma = q @ q.T / 30 + 0.2 * np.eye(30)
    chol = np.linalg.cholesky(sigma)
    aug = CondNormal(cov=sigma, premap=cmap, noise=noise)  # float64 taken from the matrix
    at = AugmentedTrajectory.from_trajectory(traj, kbt=0.6, augmenter=aug)
    eps = (noise.reshape(-1, 30) @ chol.T).reshape(-1, 10, 3)
    mean = oracle.apply_map(coords.astype(np.float64), cm)
    scaled = np.linalg.solve(sigma, eps.reshape(-1, 30).T).T.reshape(-1, 10, 3)
    want_c = np.concatenate([coords, mean + eps], axis=1)
    back = np.einsum("cf,tcd->tfd", cm, scaled)
    want_f = np.concatenate([forces + 0.6 * back, -0.6 * scaled], axis=1)
    assert at.coords.dtype == np.float64 and rel_fro(at.coords, want_c) < 1e-12 and rel_fro(at.forces, want_f) < 1e-10
    y = mean + eps
    gx, gy = aug.log_gradient(coords, y)
    assert rel_fro(gy, -scaled) < 1e-10 and rel_fro(gx, back) < 1e-10
    # generated noise: independent of the slabs, and with the requested covariance
    gen = CondNormal(cov=sigma, premap=cmap, seed=3)
    dev = torch.as_tensor(coords, device="cuda")
    whole = gen.augment_device(dev, None, 0.0, NoiseDraw(1, None))[0]
    parts = torch.cat([gen.augment_device(dev[a0:a1], None, 0.0, NoiseDraw(1, None), frame0=a0)[0]
                       for a0, a1 in [(0, 7), (7, 33), (33, len(coords))]])
    assert float((whole - parts).abs().max()) < 1e-12 * float(whole.abs().max())  # same draw (GEMM shapes differ)
    big = torch.zeros((20000, topo.n_sites, 3), dtype=torch.float64, device="cuda")
    ys = gen.augment_device(big, None, 0.0, NoiseDraw(2, None))[0][:, topo.n_sites:].reshape(20000, 30)
    emp = (ys.T @ ys / 20000).cpu().numpy()
    assert np.abs(emp - sigma).max() < 0.05 * np.abs(sigma).max()
    with pytest.raises(ValueError):
        CondNormal(cov=np.eye(7), premap=cmap)


def test_gauss_augment_row_kernel_equals_the_per_site_kernel(monkeypatch):
    """Systems whose frame rows are whole 16-byte vectors (config 5: 2 000 + 200 sites) take the row-wise
    augmentation kernel: bit-identical to the per-site kernel (same Philox draw, same float64 arithmetic), for both
    dtypes, with a site that belongs to two beads, and equal to the numpy closed form for an injected draw."""
    from aggforce_b200 import LinearMap
    from aggforce_b200.trajectory import CondNormal
    from aggforce_b200.trajectory.gausstraj import NoiseDraw

    rng = np.random.default_rng(12)
    n_sites, n_cg, n_frames = 176, 12, 301
    groups = [[int(i) for i in rng.choice(n_sites, size=rng.integers(1, 4), replace=False)] for _ in range(n_cg)]
    groups[3].append(groups[2][0])  # one site in two beads: two corrections on one force
    mat = np.zeros((n_cg, n_sites))
    for c, g in enumerate(groups):
        mat[c, g] = rng.uniform(0.2, 1.0, size=len(g))
    cmap = LinearMap(mat)
    for dtype in (np.float32, np.float64):
        coords = rng.normal(0, 5, size=(n_frames, n_sites, 3)).astype(dtype)
        forces = rng.normal(0, 40, size=(n_frames, n_sites, 3)).astype(dtype)
        dc, df = torch.as_tensor(coords, device="cuda"), torch.as_tensor(forces, device="cuda")
        aug = CondNormal(cov=0.3, premap=cmap, seed=5, dtype=dtype)
        oc, of = aug.augment_device(dc, df, 0.7, NoiseDraw(1, None), frame0=40)
        monkeypatch.setenv("AGF_AUGMENT_ROWS_OFF", "1")
        oc0, of0 = aug.augment_device(dc, df, 0.7, NoiseDraw(1, None), frame0=40)
        monkeypatch.delenv("AGF_AUGMENT_ROWS_OFF")
        assert torch.equal(oc, oc0) and torch.equal(of, of0)
        only_f = aug.augment_device(None, df, 0.7, NoiseDraw(1, None), frame0=40)[1]
        assert torch.equal(only_f, of)
        noise = rng.standard_normal((n_frames, n_cg, 3)).astype(dtype)
        inj = CondNormal(cov=0.3, premap=cmap, noise=noise, dtype=dtype)
        ic, jf = inj.augment_device(dc, df, 0.7, inj.new_draw())
        rc, rf = oracle.gauss_augment(coords, forces, mat, 0.3, 0.7, noise)
        tol = 1e-6 if dtype == np.float32 else 1e-12
        assert rel_fro(ic.cpu().numpy(), rc) < tol and rel_fro(jf.cpu().numpy(), rf) < tol


def test_feat_gram_i8_with_a_padded_width_that_is_no_multiple_of_128(topo):
    """n_basis = 4: 481 features per bead -> 576 padded columns (six 96-column blocks, 4.5 row blocks): the
    last 128-column pass of the digits kernel is partial."""
    from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat
    from aggforce_b200.synth import synth_trajectory_host
    from aggforce_b200.util import Curry

    feat = Multifeaturize([id_feat, Curry(gb_feat, inner=0.0, outer=8.0, width=1.0, n_basis=4)])
    coords, forces = synth_trajectory_host(topo, 2100, seed=79)
    want, _ = _feat_grams(coords, forces, topo, False, featurizer=feat)
    got, _ = _feat_grams(coords, forces, topo, True, featurizer=feat)
    assert got.shape == (10, 481, 481)
    for bead in range(10):
        assert rel_fro(got[bead], want[bead]) < 1e-9
