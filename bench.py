#!/usr/bin/env python
"""Benchmark of the force-map fit+apply hot path (BASELINE.json metric: frames/s).

Workload (BASELINE.json configs[1]): a synthetic replica of chignolin (cln025: 175 atoms,
10 C-alpha beads, 78 X-H constraints -> 97 reduced columns), 1 M frames per GPU.  One
"step" is one pass of the whole path over those frames:

    constraints = guess_pairwise_constraints(coords)                        # kernel (c), all frames
    project_forces(..., method=constraint_aware_uni_map)                    # kernel (d) x2
    project_forces(..., method=qp_linear_map, l2_regularization=1e3)        # kernel (a) + host QP + (d) x2

  value : frames/s with inputs resident in HBM (torch CUDA tensors in, CUDA tensors out)
  e2e   : the same calls on pinned HOST arrays; every step uploads coords+forces once and reads
          all four mapped arrays back (numpy out)
  N > 1 : every rank owns its own 1 M frames (weak scaling); Gram and pair moments are
          all-reduced (NCCL), the map is applied locally.

`--impl reference` times the CPU path (the numpy oracle port of the reference's lines; the
reference itself is Python and cannot travel to the GPU box without jax/qpsolvers) on a bounded
frame sample with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "frames/sec through force-map fit+apply"
UNIT = "frames/s"
L2_REG = 1e3
FP64_DMMA_PEAK_TFLOPS = 37.15  # measured here: profiles/r01_fp64_hbm_microbench.json (dmma884)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1_000_000, help="frames per GPU")
    ap.add_argument("--cpu-frames", type=int, default=4000, help="frames of the CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arm
def cpu_step(oracle, topo, coords, forces, cm):
    cons = oracle.guess_pairwise_constraints_literal(coords)
    uni = oracle.uni_map_matrix(cm, cons)
    out = [oracle.apply_map(coords, cm), oracle.apply_map(forces, uni)]
    w = oracle.qp_linear_weights(forces, cm, cons, L2_REG)
    out += [oracle.apply_map(coords, cm), oracle.apply_map(forces, w)]
    return oracle.force_smoothness(out[-1])


def cpu_run(n_frames: int, steps: int, warmup: int):
    import oracle
    from aggforce_b200.synth import chignolin_topology, synth_trajectory_host

    topo = chignolin_topology()
    coords, forces = synth_trajectory_host(topo, n_frames, seed=1234)
    cm = np.zeros((len(topo.bead_atoms), topo.n_sites))
    cm[np.arange(len(topo.bead_atoms)), topo.bead_atoms] = 1
    for _ in range(warmup):
        cpu_step(oracle, topo, coords, forces, cm)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(oracle, topo, coords, forces, cm)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return n_frames / dt, dt


def reference_arm(args) -> None:
    if int(os.environ.get("RANK", "0")) != 0:
        return
    value, dt = cpu_run(args.cpu_frames, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = (f"{args.cpu_frames} frames of the same synthetic cln025 workload per step "
              "(float64 numpy oracle port of the reference lines; OpenBLAS threads = all cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.cpu_frames),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(frames_per_gpu: int) -> dict:
    return {
        "workload": "cln025 (175 atoms, 10 CA beads, 78 X-H constraints -> n_red 97): "
                    "guess_pairwise_constraints(all frames) + constraint_aware_uni_map + optimised "
                    "qp_linear_map(l2=1e3) via project_forces, synthetic replica",
        "frames_per_gpu": frames_per_gpu,
        "cache": "inputs (2.1 GB per array per GPU) are larger than L2; no explicit flush",
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """The contract is ONE JSON line on stdout: route everything libraries print to fd 1 (NCCL's
    version banner, ...) to stderr and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import aggforce_b200 as agf
    from aggforce_b200 import _lib
    from aggforce_b200.synth import chignolin_topology, synth_trajectory_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T = args.frames
    topo = chignolin_topology()
    cmap = agf.LinearMap([[i] for i in topo.bead_atoms], n_fg_sites=topo.n_sites)
    coords, forces = synth_trajectory_device(topo, T, seed=1234, frame0=rank * T)
    torch.cuda.synchronize()

    def step(c, f):
        cons = agf.guess_pairwise_constraints(c)
        r1 = agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds=cons,
                                method=agf.constraint_aware_uni_map)
        r2 = agf.project_forces(coords=c, forces=f, coord_map=cmap, constrained_inds=cons,
                                l2_regularization=L2_REG)
        return cons, r1, r2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        # the sampler runs across warm-up AND the timed steps: the timed region alone lasts tens of
        # milliseconds, less than one nvidia-smi sampling period
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None  # one nvidia-smi poller per box
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.LAUNCHES["count"]
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        launches = _lib.LAUNCHES["count"] - n0
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        return float(ms.item()) / steps, launches, clocks

    ctx = agf.frame_sharding(world > 1)
    with ctx:
        # ---- device-resident throughput
        ms_dev, launches, clocks = timed(lambda: step(coords, forces), args.steps, args.warmup, sample_clocks=True)

        # ---- per-kernel device times inside one extra (untimed-for-value) pass
        _lib.timing(True)
        step(coords, forces)
        per_kernel = {}
        for name, ms in _lib.timing_records():
            per_kernel.setdefault(name, []).append(ms)
        _lib.timing(False)

        # ---- end to end from pinned host memory
        e2e = None
        if not args.no_e2e:
            hc = torch.empty(coords.shape, dtype=coords.dtype, pin_memory=True)
            hf = torch.empty(forces.shape, dtype=forces.dtype, pin_memory=True)
            hc.copy_(coords)
            hf.copy_(forces)
            torch.cuda.synchronize()
            nc, nf = hc.numpy(), hf.numpy()
            d2h = {"n": 0}

            def host_step():
                c, f = agf.Frames(nc), agf.Frames(nf)  # fresh wrappers: uploaded once per step
                _, r1, r2 = step(c, f)
                d2h["n"] = sum(r[k].nbytes for r in (r1, r2) for k in ("mapped_coords", "mapped_forces"))

            ms_e2e, _, _ = timed(host_step, max(1, min(args.steps, 5)), 3)
            e2e = {"value": world * T / (ms_e2e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(nc.nbytes + nf.nbytes), "d2h_bytes_per_step": int(d2h["n"]),
                   "ms_per_step": ms_e2e}
            del hc, hf, nc, nf

    # ---- roofline of the dominant kernel (algorithmic work per launch / measured time)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    hbm_peak, hbm_src = 6650.0, "fallback"
    if peaks_path.exists():
        hbm_peak, hbm_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured"
    n, n_red, n_cg = topo.n_sites, 97, len(topo.bead_atoms)
    uni_nnz = 21
    algo = {  # per launch over T frames: (kind, amount, unit)
        "agf_gram_linear": ("tensor", (3 * n_red * (n_red + 1) + 3 * n) * T, "flop"),
        "agf_pair_moments": ("hbm", 12 * n * T, "B"),
        "agf_map_apply": ("hbm", (12 * n + 24 * n_cg) * T, "B"),
        "agf_map_apply_sparse": ("hbm", None, "B"),
    }
    # DRAM bytes per frame of each kernel from `ncu --set full` captures of this workload
    # (dram__bytes_read.sum + dram__bytes_write.sum divided by the frames of the captured launch):
    # profiles/r01_ncu_gram_ws.txt, profiles/r01_ncu_apply_v2.txt, profiles/r01_ncu_all_kernels_v1.txt
    traffic_per_frame = {"agf_gram_linear": 2111.0, "agf_map_apply": 2333.3, "agf_pair_moments": 2109.6}
    kernels = {}
    for name, times in per_kernel.items():
        tot = float(np.sum(times))
        entry = {"launches": len(times), "ms_total": tot}
        if name in algo and algo[name][1] is not None:
            kind, amount, _ = algo[name]
            per = amount / (np.mean(times) * 1e-3)
            if kind == "tensor":
                entry.update(bound="tensor", achieved=per / 1e12, peak=FP64_DMMA_PEAK_TFLOPS, unit="TFLOP/s")
            else:
                entry.update(bound="hbm", achieved=per / 1e9, peak=hbm_peak, unit="GB/s")
            entry["frac"] = entry["achieved"] / entry["peak"]
        elif name in ("agf_map_apply_sparse", "agf_map_apply_slice"):
            # slice: the coordinate map of both project_forces calls (10 sites); sparse: the uniform force
            # map (21 nnz).  Algorithmic bytes = referenced sites (12 B each) + f64 outputs, per launch
            nnz = n_cg if name == "agf_map_apply_slice" else uni_nnz
            amount = (nnz * 12 + 24 * n_cg) * T * len(times)
            entry.update(bound="hbm", achieved=amount / (tot * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
            entry["frac"] = entry["achieved"] / entry["peak"]
        if name in traffic_per_frame:
            entry["traffic"] = traffic_per_frame[name] * T
        kernels[name] = entry
    dominant = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    roofline = None
    if dominant and "frac" in kernels[dominant]:
        k = kernels[dominant]
        roofline = {"kernel": dominant, "bound": k["bound"], "achieved": k["achieved"], "peak": k["peak"],
                    "unit": k["unit"], "frac": k["frac"], "traffic": k.get("traffic"),
                    "peak_source": ("FP64 DMMA microbenchmark measured on this pool "
                                    "(profiles/r01_fp64_hbm_microbench.json)") if k["bound"] == "tensor"
                    else f"MEASURED_PEAKS.json hbm_gbs ({hbm_src})"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt = cpu_run(args.cpu_frames, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{args.cpu_frames} frames of the same workload, 3 timed steps "
                         f"({dt:.2f} s each), float64 numpy oracle, BLAS threads = all cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * T / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(T), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
