"""Config 4 at one rank's full share (default 1.25 M frames x 5000 atoms, n_red 2600, 500 beads):
the trajectory (75 GB per array) is never materialised -- SynthFrames regenerate slabs of frames
from the counter-based generator for every pass (constraint detection, Gram, two applications)."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
import aggforce_b200 as agf
from aggforce_b200 import _lib
from aggforce_b200.synth import make_synth_frames, protein_like_topology

import contextlib, os
import torch.distributed as dist
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:  # torchrun: every rank owns frames [rank*T, (rank+1)*T) of the same synthetic trajectory
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    _print = print
    def print(*a, **k):  # noqa: A001
        if rank == 0: _print(*a, **k)
topo = protein_like_topology(500)
coords = make_synth_frames(topo, T, "coords", seed=3, frame0=rank * T, slab_bytes=4 << 30)
forces = make_synth_frames(topo, T, "forces", seed=3, frame0=rank * T, slab_bytes=4 << 30)
cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
print(f"n_sites {topo.n_sites} beads {len(topo.bead_atoms)} frames {T}  ({T * topo.n_sites * 12 / 1e9:.1f} GB per array, virtual)")
if world > 1:  # warm NCCL and cuSOLVER so the timed pass is the steady state
    with agf.frame_sharding():
        agf.project_forces(coords=make_synth_frames(topo, 4096, "coords", seed=3, frame0=rank * 4096),
                           forces=make_synth_frames(topo, 4096, "forces", seed=3, frame0=rank * 4096), coord_map=cmap,
                           constrained_inds="auto", l2_regularization=1e3)
    dist.barrier()
torch.cuda.synchronize()
_lib.timing(True)
t0 = time.perf_counter()
with (agf.frame_sharding() if world > 1 else contextlib.nullcontext()):
    res = agf.project_forces(coords=coords, forces=forces, coord_map=cmap, constrained_inds="auto", l2_regularization=1e3)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
recs = _lib.timing_records(); _lib.timing(False)
agg = {}
for n, ms in recs: agg[n] = agg.get(n, 0.0) + ms
print(f"project_forces(auto constraints, qp_linear_map), {world} GPU(s): {dt:.2f} s = {world * T/dt:.3e} frames/s "
      f"({world * T} frames in total)")
print("  entry points [ms]:", ", ".join(f"{n} {ms:.0f}" for n, ms in sorted(agg.items(), key=lambda kv: -kv[1])))
print("  constraints", len(res["constraints"]), "== topology:", res["constraints"] == topo.xh_constraints,
      " residual", res["residual"], " mapped forces", tuple(res["mapped_forces"].shape), res["mapped_forces"].dtype)
flop = (3 * 2600 * 2601 + 3 * 5000) * T
for name in ("agf_gram_linear_ws", "agf_gram_linear_i8t"):
    g = agg.get(name, 0.0)
    if g: print(f"  Gram ({name}): {g:.0f} ms = {flop / g / 1e9:.1f} float64-equivalent TFLOP/s ({flop / g / 1e9 / 37.15 * 100:.0f}% of the DMMA peak)")
print("  peak memory allocated: %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
if world > 1:
    w = torch.as_tensor(np.ascontiguousarray(res["tmap"].force_map.standard_matrix), device="cuda")
    ref = w.clone(); dist.broadcast(ref, 0)
    print("  force map identical on every rank:", bool(torch.equal(w, ref)))
    dist.destroy_process_group()
