"""GPU parity against the REFERENCE's JAX half: gb_feat / JCondNormal / joptgauss_map / jaxmapval.

The fixtures (``tests/golden/ref_gbfeat.npz``, ``ref_jcondnormal.npz``, ``ref_mapval.npz``) hold the
outputs of the reference's own modules run unmodified behind the torch-backed jax stand-in
(``tests/golden/make_golden_jax.py``).  The reference computes these in float32, the CUDA path in
float64: bars are float32 bars, written at each assert.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import pairs_to_set, rel_fro

pytestmark = pytest.mark.gpu

F32_ABS = 5e-6


@pytest.fixture(scope="module")
def gbfix(golden):
    return dict(np.load(golden / "ref_gbfeat.npz"))


@pytest.fixture(scope="module")
def jcn(golden):
    return dict(np.load(golden / "ref_jcondnormal.npz"))


def _slice_map(beads, n_fg):
    from aggforce_b200 import LinearMap

    return LinearMap([[int(b)] for b in beads], n_fg_sites=n_fg)


@pytest.mark.parametrize("nb", [4, 7])
@pytest.mark.parametrize("dm", ["reorder", "basic"])
def test_gb_feat_matches_reference(gbfix, nb, dm):
    from aggforce_b200 import LinearMap
    from aggforce_b200.qp import gb_feat, id_feat

    x = gbfix["small_x"]
    cons = pairs_to_set(gbfix["small_cons"])
    cmap = LinearMap(gbfix["small_cmap_slice"])
    assert np.array_equal(id_feat(x, cmap, cons, return_ids=True), gbfix["small_ids"])
    out = gb_feat(x, cmap, cons, outer=8, inner=0, n_basis=nb, width=1.0, div_method=dm, lazy=False)
    ref_f, ref_d = gbfix[f"small_slice_nb{nb}_{dm}_feats"], gbfix[f"small_slice_nb{nb}_{dm}_divs"]
    for bead in range(3):
        assert out["feats"][bead].shape == ref_f[bead].shape and out["feats"][bead].dtype == np.float32
        assert np.abs(out["feats"][bead] - ref_f[bead]).max() < F32_ABS
        # bead 1 sits on the (dropped) top-label site: NaN everywhere under "reorder", nowhere under "basic"
        assert np.array_equal(np.isnan(out["divs"][bead]), np.isnan(ref_d[bead]))
        ok = ~np.isnan(ref_d[bead])
        assert not ok.any() or np.abs(out["divs"][bead][ok] - ref_d[bead][ok]).max() < F32_ABS


def test_gb_feat_cln025_matches_reference(gbfix, small_cln):
    from aggforce_b200.qp import gb_feat
    from aggforce_b200.synth import chignolin_topology

    topo = chignolin_topology()
    cmap = _slice_map(topo.bead_atoms, 175)
    out = gb_feat(small_cln["coords"][:4], cmap, pairs_to_set(small_cln["cons10"]), outer=8, inner=0, n_basis=7,
                  width=1.0, lazy=False)
    for bead in (0, 7):
        assert np.abs(out["feats"][bead] - gbfix[f"cln_feats_b{bead}"]).max() < F32_ABS
        assert np.abs(out["divs"][bead] - gbfix[f"cln_divs_b{bead}"]).max() < 2e-5


def test_fused_featurised_fit_matches_reference(gbfix):
    """Kernel (b) + equality rows + solve + featurised application against the reference's
    qp_feat_linear_map(Multifeaturize([id_feat, Curry(gb_feat, n_basis=4)]))."""
    from aggforce_b200 import Trajectory, _engine
    from aggforce_b200.qp import Multifeaturize, gb_feat, id_feat, qp_feat_linear_map
    from aggforce_b200.qp.featlinearmap import _FusedContext, _fusable
    from aggforce_b200.util import Curry

    c, f = gbfix["fit_coords"], gbfix["fit_forces"]
    cons = pairs_to_set(gbfix["fit_cons"])
    cmap = _slice_map(gbfix["fit_beads"], c.shape[1])
    kbt, l2 = float(gbfix["fit_kbt"]), float(gbfix["fit_l2"])
    feat = Multifeaturize([id_feat, Curry(gb_feat, inner=0.0, outer=8.0, width=1.0, n_basis=4)])
    ctx = _FusedContext(cmap, cons, _fusable(feat))
    assert np.array_equal(ctx.labels, gbfix["fit_ids"])
    grams = ctx.grams(_engine.Frames(c), _engine.Frames(f), kbt)
    n_feat = grams.shape[1]
    assert grams.shape[1:] == gbfix["fit_P"].shape[1:]
    for bead in range(len(gbfix["fit_beads"])):
        # the reference accumulates this Gram in float32
        assert rel_fro(grams[bead] + l2 * np.eye(n_feat), gbfix["fit_P"][bead]) < 2e-6
        a = ctx.constraint_rows(_engine.Frames(c), bead, gbfix["fit_frame_choice"])
        assert np.abs(a - gbfix["fit_A"][bead]).max() < F32_ABS
    traj = Trajectory(coords=c, forces=f)
    tmap = qp_feat_linear_map(traj, cmap, feat, kbt, constraints=cons, l2_regularization=l2,
                              constraint_frames=gbfix["fit_frame_choice"])
    mapped = tmap(traj)
    # 72 regression rows for 106 features: the coefficients are ill-conditioned against the float32
    # Gram of the reference, the mapped forces are not
    assert rel_fro(mapped.forces, gbfix["fit_mapped_forces"]) < 1e-4
    assert rel_fro(mapped.coords, gbfix["fit_mapped_coords"]) < 1e-7
    # the application kernel with the REFERENCE's coefficients
    out, _ = ctx.apply(_engine.Frames(c), _engine.Frames(f), gbfix["fit_coefs"])
    assert rel_fro(out.cpu().numpy(), gbfix["fit_mapped_forces"]) < 1e-5


@pytest.mark.parametrize("name", ["slice", "avg"])
def test_condnormal_matches_reference_jcondnormal(jcn, name):
    from aggforce_b200 import LinearMap
    from aggforce_b200.trajectory import CondNormal

    src, var = jcn["source"], float(jcn["var"])
    lm = LinearMap(jcn[f"{name}_matrix"])
    aug = CondNormal(cov=var, premap=lm, noise=jcn[f"{name}_z"])
    y = aug.sample(src)
    assert y.dtype == np.float32 and np.abs(y - jcn[f"{name}_generated"]).max() < 1e-5
    gx, gy = aug.log_gradient(src, jcn[f"{name}_generated"])
    scale = max(1.0, np.abs(jcn[f"{name}_lg_source"]).max())
    assert np.abs(gx - jcn[f"{name}_lg_source"]).max() < 2e-5 * scale
    assert np.abs(gy - jcn[f"{name}_lg_generated"]).max() < 2e-5 * scale


def test_joptgauss_matches_reference(jcn, gbfix):
    from aggforce_b200 import Trajectory, joptgauss_map

    c, f = gbfix["fit_coords"], gbfix["fit_forces"]
    cons = pairs_to_set(jcn["jopt_cons"])
    cmap = _slice_map(jcn["jopt_beads"], c.shape[1])
    var, kbt, l2 = float(jcn["jopt_var"]), float(jcn["jopt_kbt"]), float(jcn["jopt_l2"])
    traj = Trajectory(coords=c, forces=f)
    tmap = joptgauss_map(traj, cmap, var=var, kbt=kbt, constraints=cons, noise=jcn["jopt_z_fit"], l2_regularization=l2)
    assert np.array_equal(tmap.tmap.coord_map.standard_matrix, jcn["jopt_coord_matrix"])
    w = tmap.tmap.force_map.standard_matrix
    assert rel_fro(w, jcn["jopt_W"]) < 1e-4  # the reference's augmented forces and Gram inputs are float32
    tmap.augmenter._noise = jcn["jopt_z_apply"]
    out = tmap(traj)
    assert rel_fro(out.coords, jcn["jopt_mapped_coords"]) < 1e-6
    assert rel_fro(out.forces, jcn["jopt_mapped_forces"]) < 1e-4


def test_validation_projections_match_reference(golden):
    from aggforce_b200 import jaxmapval as mv

    ref = dict(np.load(golden / "ref_mapval.npz"))
    c, f = ref["cg_coords"], ref["cg_forces"]
    kw = dict(inner=6.0, outer=12.0, width=0.5)
    for s in range(3):
        got = mv.rsqpg_forces(c, randg=np.random.default_rng(s), **kw)
        r0 = ref["rsqpg_forces"][s]
        assert got.shape == r0.shape and np.abs(got - r0).max() < 1e-4 * max(1.0, np.abs(r0).max())
        off, w = oracle.rsqpg_offset(6.0, 12.0, 0.5, np.random.default_rng(s))
        assert rel_fro(mv.sq_gaussian_forces(c, off, w), oracle.sq_gaussian_forces(c, off, w)) < 1e-12
    got = mv.rsqpg_forces(c, inner=30.0, outer=80.0, width=9.0, randg=np.random.default_rng(1), sq_args=False)
    assert np.abs(got - ref["rsqpg_forces_nosq"]).max() < 1e-4 * max(1.0, np.abs(ref["rsqpg_forces_nosq"]).max())
    proj = mv.random_force_proj(coords=c, forces=f, n_samples=16, randg=np.random.default_rng(42100), average=False, **kw)
    scale = np.abs(ref["proj"]).max()
    assert np.abs(np.asarray(proj) - ref["proj"]).max() < 1e-4 * scale
    avg = mv.random_force_proj(coords=c, forces=f, n_samples=16, randg=np.random.default_rng(42100), **kw)
    assert abs(avg - ref["proj_avg"]) < 1e-4 * scale
    shift = mv.random_residual_shift(coords=c, forces=f, n_samples=16, randg=np.random.default_rng(42100), **kw)
    bar = 1e-5 * float(np.mean(f.astype(np.float64) ** 2)) + 1e-4 * np.abs(ref["shift"]).max()
    assert np.abs(np.asarray(shift) - ref["shift"]).max() < bar
    # float64 oracle: tight
    o_proj = oracle.random_force_proj(c, f, 16, np.random.default_rng(42100), **kw)
    o_shift = oracle.random_residual_shift(c, f, 16, np.random.default_rng(42100), **kw)
    assert rel_fro(proj, o_proj) < 1e-11 and np.abs(np.asarray(shift) - o_shift).max() < 1e-9 * np.mean(f.astype(float) ** 2)
    # device tensors in, many samples (several per thread and more than one launch's worth of threads)
    dc, df = torch.as_tensor(c, device="cuda"), torch.as_tensor(f, device="cuda")
    many = mv.random_force_proj(coords=dc, forces=df, n_samples=700, randg=np.random.default_rng(3), average=False, **kw)
    o_many = oracle.random_force_proj(c, f, 700, np.random.default_rng(3), **kw)
    assert rel_fro(many, o_many) < 1e-11
    # a user-supplied method runs sample by sample like the reference
    uni = mv.random_force_proj(coords=c, forces=f, n_samples=4, randg=np.random.default_rng(3),
                               method=mv.random_uniform_forces, average=False, scale=2.0)
    r3 = np.random.default_rng(3)
    want = [oracle.mscg_ip(f, mv.random_uniform_forces(c, scale=2.0, randg=r3)) for _ in range(4)]
    assert np.allclose(uni, want, rtol=1e-12)
    assert np.allclose(mv.random_uniform_forces(c, scale=2.0, randg=np.random.default_rng(3)), ref["uniform"])
