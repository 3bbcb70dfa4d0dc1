"""CPU, world_size 2, gloo: the host-side combine logic of frame sharding.

Kernels need a GPU, so these tests feed the collectives with per-rank partial results computed
by the oracle and check that what comes out equals the single-process answer.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import oracle
    from aggforce_b200 import _engine
    from aggforce_b200.agg import _global_mean_sq

    rng = np.random.default_rng(0)  # same data on both ranks; each takes its shard
    T, n = 90, 7
    x = rng.normal(size=(T, n, 3)) + 5 * rng.normal(size=(1, n, 3))
    f = rng.normal(size=(T, n, 3))
    lo, hi = (0, 37) if rank == 0 else (37, T)  # ragged split
    with _engine.frame_sharding(True):
        assert _engine.sharded()
        assert _engine.global_count(hi - lo) == T
        # pair moments: per-rank (count, mean, M2) -> merged sd == full-data sd
        ii, jj = np.triu_indices(n, k=1)
        d = np.linalg.norm(x[lo:hi, ii] - x[lo:hi, jj], axis=-1)
        rec = np.concatenate([[float(hi - lo)], d.mean(0), ((d - d.mean(0)) ** 2).sum(0)])
        sd = _engine.merge_moments(_engine.allgather_host(rec), len(ii))
        full = oracle.pair_distance_sd(x)[ii, jj]
        assert np.allclose(sd, full, rtol=1e-12, atol=1e-14)
        # additive accumulators: Gram shards sum to the full Gram
        g = torch.as_tensor(oracle.gram_linear(f[lo:hi]))
        _engine.allreduce_sum_(g)
        assert np.allclose(g.numpy(), oracle.gram_linear(f), rtol=1e-12)
        # alive masks combine with MIN (dead on any rank = dead)
        m = torch.tensor([1, rank, 1 - rank, 0], dtype=torch.uint8)
        _engine.allreduce_min_(m)
        assert m.tolist() == [1, 0, 0, 0]
        # residual: global mean of squares from per-rank sums
        local = f[lo:hi]
        assert abs(_global_mean_sq(float((local**2).sum()), local.shape) - float((f**2).mean())) < 1e-12
        # host helpers: sum / broadcast / rank
        assert _engine.shard_rank() == rank
        assert np.array_equal(_engine.allreduce_host_sum(np.full((2, 3), rank + 1.0)), np.full((2, 3), 3.0))
        assert np.array_equal(_engine.broadcast_host(np.arange(4.0) + 10 * rank), np.arange(4.0))
        # equality-row frames of a featurised fit are chosen ONCE for all ranks, in global frame
        # indices, and every rank assembles the same rows from the frames it owns
        from aggforce_b200.qp.featlinearmap import _ConstraintFrames

        chosen = _ConstraintFrames(hi - lo, n_beads=3, n_frames=20, constraint_frames=None)
        assert chosen.picks.shape == (3, 20) and chosen.offset == lo
        both = _engine.allgather_host(chosen.picks.astype(np.float64).reshape(-1))
        assert np.array_equal(both[0], both[1])  # unseeded draw, still identical on both ranks
        feat = rng.normal(size=(T, 2, 5))  # "features" (frame, bead', feature), known to both ranks

        def local_rows(idx):
            return feat[lo:hi][idx].reshape(-1, 5)

        rows = chosen.rows(1, local_rows, n_cg=2)
        assert np.allclose(rows, feat[chosen.picks[1]].reshape(-1, 5), rtol=0, atol=0)
        fixed = _ConstraintFrames(hi - lo, 1, 4, np.array([0, 36, 37, 89]))  # straddles the shard boundary
        assert np.array_equal(fixed.rows(0, local_rows, 2), feat[[0, 36, 37, 89]].reshape(-1, 5))
    assert not _engine.sharded()
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_frame_sharding_combine_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
