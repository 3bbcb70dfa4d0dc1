"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE.

Run once in the build container (``python tests/golden/make_golden.py``); the GPU box
has no ``/root/reference`` so tests only ever read the committed outputs.

The reference (``/root/reference/src``) is imported unmodified.  Its two unavailable
third-party imports are handled like this:
  * ``qpsolvers`` -- replaced by a shim whose ``solve_qp`` (a) RECORDS the ``P, A, b``
    the reference hands to the solver, which is how the Gram matrices built by the
    reference's own numpy lines are captured, and (b) returns the exact equality-QP
    minimiser;
  * ``jax`` -- absent; the reference's ``try/except ImportError`` blocks drop gb_feat and
    the Gaussian maps, so nothing from those is pinned here.

Outputs:
  waterdimer.npz                  copy of the reference's known-answer INPUT data
  cln_basic_force_mat.txt         reference golden (tests/test_forces.py:154-157)
  cln_opt_force_mat.txt           reference golden (tests/test_forces.py:182-185)
  ../../aggforce_b200/data/cln025_topology.json   topology derived from tests/data/cln025.pdb
  ref_small_cln.npz               seeded synthetic inputs + reference outputs for kernels a, c, d
  ref_idfeat.npz                  reference id-feature Grams / labels / equality rows
  ref_sets.json                   reference set-algebra outputs on random constraint sets
  ref_linearmap.npz               reference LinearMap outputs (tests/test_linearmap.py data, NaN protocol)
  ref_condnormal.npz              reference SimpleCondNormal.log_gradient
"""
from __future__ import annotations

import json
import shutil
import sys
import types
from pathlib import Path

import numpy as np
import scipy.sparse as ss

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")

RECORD: list = []


def _solve_qp(P, q, A=None, b=None, **_):
    Pd = P.toarray() if ss.issparse(P) else np.asarray(P, float)
    Ad = A.toarray() if ss.issparse(A) else np.asarray(A, float)
    bd = np.asarray(b, float)
    RECORD.append((Pd.copy(), Ad.copy(), bd.copy()))
    PiAt = np.linalg.solve(Pd, Ad.T)
    return PiAt @ np.linalg.lstsq(Ad @ PiAt, bd, rcond=None)[0]


shim = types.ModuleType("qpsolvers")
shim.solve_qp = _solve_qp
sys.modules["qpsolvers"] = shim
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REPO))

import aggforce  # noqa: E402  (the reference)
from aggforce import LinearMap, guess_pairwise_constraints, project_forces  # noqa: E402
from aggforce.constraints import constraint_lookup_dict, reduce_constraint_sets  # noqa: E402
from aggforce.qp import constraint_aware_uni_map, id_feat, qp_feat_linear_map, qp_linear_map  # noqa: E402
from aggforce.qp import featlinearmap as ref_flm  # noqa: E402
from aggforce.trajectory import Trajectory  # noqa: E402
from aggforce.trajectory.simplegausstraj import SimpleCondNormal  # noqa: E402

assert "/root/reference" in aggforce.__file__


def make_topology() -> dict:
    names, elems, xyz, resid = [], [], [], []
    for line in (REF / "tests/data/cln025.pdb").read_text().splitlines():
        if line.startswith("ATOM"):
            names.append(line[12:16].strip())
            resid.append(int(line[22:26]))
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            elems.append(line[76:78].strip())
    xyz = np.asarray(xyz)
    heavy = [i for i, e in enumerate(elems) if e != "H"]
    pairs = []
    for i, e in enumerate(elems):
        if e == "H":
            d = np.linalg.norm(xyz[heavy] - xyz[i], axis=1)
            pairs.append([int(heavy[int(np.argmin(d))]), i])
    ca = [i for i, n in enumerate(names) if n == "CA"]
    topo = {
        "source": "derived from the reference's tests/data/cln025.pdb (chignolin CLN025, 175 atoms)",
        "n_atoms": len(names),
        "names": names,
        "elements": elems,
        "residue": resid,
        "ca_indices": ca,
        "xh_pairs": pairs,
        "positions_angstrom": np.round(xyz, 3).tolist(),
    }
    out = REPO / "aggforce_b200" / "data" / "cln025_topology.json"
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(topo))
    return topo


def main() -> None:
    for name in ("waterdimer.npz", "cln_basic_force_mat.txt", "cln_opt_force_mat.txt"):
        shutil.copyfile(REF / "tests/data" / name, HERE / name)
    topo = make_topology()
    assert topo["ca_indices"] == [8, 29, 50, 70, 76, 91, 105, 112, 126, 150]

    from aggforce_b200.synth import chignolin_topology, synth_trajectory_host

    t = chignolin_topology()
    coords, forces = synth_trajectory_host(t, n_frames=48, seed=1234)
    n_fg = coords.shape[1]
    cmap = LinearMap([[i] for i in t.bead_atoms], n_fg_sites=n_fg)

    # ---- kernel (c): reference evaluated in float64 and in float32
    cons10 = guess_pairwise_constraints(coords[0:10].astype(np.float64), threshold=1e-3)
    cons10_f32 = guess_pairwise_constraints(coords[0:10], threshold=1e-3)
    cons_all = guess_pairwise_constraints(coords.astype(np.float64), threshold=1e-3)
    cross = guess_pairwise_constraints(
        coords[:, :60].astype(np.float64), cross_xyz=coords[:, 40:90].astype(np.float64), threshold=1e-3
    )
    from aggforce.util import distances

    sds = np.sqrt(np.var(distances(coords.astype(np.float64)), axis=0))

    def pairs(s):
        return np.asarray(sorted(tuple(sorted(int(v) for v in p)) for p in s), dtype=np.int64).reshape(-1, 2)

    # ---- kernel (a) + solve + (d): the reference's project_forces, recording P/A
    RECORD.clear()
    res = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons10,
                         l2_regularization=1e3)
    P_l2, A_lin, _ = RECORD[0]
    W = res["tmap"].force_map.standard_matrix
    RECORD.clear()
    res0 = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons10)
    P_raw = RECORD[0][0]
    RECORD.clear()
    res_nc = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=set())
    P_nocons = RECORD[0][0]
    uni = project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds=cons10,
                         method=constraint_aware_uni_map)
    np.savez_compressed(
        HERE / "ref_small_cln.npz",
        coords=coords, forces=forces,
        cons10=pairs(cons10), cons10_f32=pairs(cons10_f32), cons_all=pairs(cons_all),
        cross_pairs=np.asarray(sorted(cross), dtype=np.int64).reshape(-1, 2),
        sds=sds,
        gram_raw=P_raw, gram_l2_1e3=P_l2, gram_nocons=P_nocons, A_lin=A_lin,
        W_l2_1e3=W, W_raw=res0["tmap"].force_map.standard_matrix,
        W_nocons=res_nc["tmap"].force_map.standard_matrix,
        mapped_forces=res["mapped_forces"], mapped_coords=res["mapped_coords"],
        residual=np.float64(res["residual"]),
        uni_matrix=uni["tmap"].force_map.standard_matrix,
        uni_mapped_forces=uni["mapped_forces"], uni_residual=np.float64(uni["residual"]),
    )

    # ---- id features: labels, Grams, equality rows with an injected frame choice
    ids = id_feat(coords, cmap, cons10, return_ids=True)
    frame_choice = np.random.default_rng(42100).choice(len(coords), size=20, replace=False)

    class _FixedRng:
        def choice(self, n, size, replace=False):
            assert size == 20 and n == len(coords)
            return frame_choice

    ref_flm.default_rng = lambda *a, **k: _FixedRng()
    RECORD.clear()
    kbt = 0.6955215
    fmap = qp_feat_linear_map(Trajectory(coords=coords, forces=forces), cmap, id_feat, kbt,
                              constraints=cons10, l2_regularization=1e1)
    id_P = np.stack([r[0] for r in RECORD])
    id_A = np.stack([r[1] for r in RECORD])
    id_b = np.stack([r[2] for r in RECORD])
    id_coefs = np.stack(fmap.force_map.tags["coef_list"])
    id_mapped = fmap(Trajectory(coords=coords, forces=forces)).forces
    np.savez_compressed(HERE / "ref_idfeat.npz", ids=ids, frame_choice=frame_choice, P=id_P, A=id_A, b=id_b,
                        coefs=id_coefs, mapped_forces=id_mapped, kbt=np.float64(kbt))

    # ---- set algebra on random inputs (including chains/stars) + label orders
    rng = np.random.default_rng(7)
    cases = []
    for trial in range(40):
        n = int(rng.integers(4, 60))
        k = int(rng.integers(0, n))
        cons = set()
        for _ in range(k):
            sz = int(rng.integers(2, 4))
            cons.add(frozenset(int(v) for v in rng.choice(n, size=sz, replace=False)))
        red = reduce_constraint_sets(cons)
        look = constraint_lookup_dict(red)
        dummy = LinearMap([[0]], n_fg_sites=n)
        lab = id_feat(np.zeros((1, n, 3)), dummy, cons, return_ids=True)
        cases.append({
            "n": n,
            "constraints": sorted(sorted(g) for g in cons),
            "reduced": sorted(sorted(int(v) for v in g) for g in red),
            "lookup": {str(k2): int(v2) for k2, v2 in look.items()},
            "ref_labels": [int(v) for v in lab],
        })
    (HERE / "ref_sets.json").write_text(json.dumps({"cases": cases, "cln_ids": [int(v) for v in ids]}))

    # ---- LinearMap behaviour (tests/test_linearmap.py fixtures + NaN protocol)
    r2 = np.random.default_rng(seed=42100)
    pos = 100 * (r2.random(size=(20, 15, 3)) - 0.5)
    r3 = np.random.default_rng(seed=42100)
    mat = r3.random(size=(5, 15))
    lm = LinearMap(mapping=mat)
    lst = LinearMap([[0, 2, 3], [4]], n_fg_sites=6)
    pos_nan = pos.copy()
    sl = LinearMap([[1], [7], [9]], n_fg_sites=15)
    pos_nan[:, [0, 3, 5], :] = np.nan  # untouched by the slice map -> allowed
    out_nan = sl(pos_nan.copy())
    np.savez_compressed(
        HERE / "ref_linearmap.npz",
        pos=pos, mat=mat, mapped=lm(pos), mapped_f32in=lm(pos.astype(np.float32)),
        mapped_f32map=lm.astype(np.float32)(pos.astype(np.float32)),
        flat=lm.flat_call(pos.reshape(20, 45)),
        list_matrix=lst.standard_matrix,
        pos_nan=pos_nan, nan_out=out_nan, slice_matrix=sl.standard_matrix,
        transpose=lm.T.standard_matrix, scaled=(2.5 * lm).standard_matrix, summed=(lm + lm).standard_matrix,
        composed=(LinearMap(mat[:, :5]) @ lm).standard_matrix,
    )

    # ---- Gaussian score closed form (reference numpy augmenter)
    aug = SimpleCondNormal(var=0.37, seed=5)
    src = pos[:, :5].astype(np.float32)
    gen = aug.sample(src)
    lg_src, lg_gen = aug.log_gradient(src, gen)
    np.savez_compressed(HERE / "ref_condnormal.npz", source=src, generated=gen, lg_source=lg_src,
                        lg_generated=lg_gen, var=np.float64(0.37))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
