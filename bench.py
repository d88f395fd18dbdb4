#!/usr/bin/env python
"""Benchmark of the hot path: Lanczos (full re-orthogonalisation) forward + adjoint.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): sparse SPD operator, n = 1M rows, 11 entries per row
(COO layout of `suite_sparse_load`), Krylov depth 100, fp32, cotangents on the tridiagonal
coefficients (the SLQ case).  One bench "step" = one forward + one adjoint sweep = 100 Krylov
steps; the metric is Krylov steps per second, `depth / (t_fwd + t_adj)`.

N > 1 (launched by torchrun, one rank per GPU): every rank runs the forward + adjoint of its
own probe vector on a replicated operator (probe sharding) and the step ends with a
single NCCL all-reduce of the parameter cotangent; weak scaling.

`--impl reference`: the reference's algorithm on the host CPUs (the NumPy/SciPy oracle port —
JAX is not installed in this image, so the reference itself cannot run), bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = int(os.environ.get("BL_BENCH_N", 1_000_000))
DEPTH = int(os.environ.get("BL_BENCH_DEPTH", 100))
BANDS = 5
METRIC = "Lanczos fwd+adjoint steps/sec at n=1M, K=100; achieved HBM GB/s vs peak"
UNIT = "krylov_steps/s"


# DRAM traffic of the dominant kernel from one `ncu --set full` capture (profiles/r1_prof_r1_fused.md):
# k_fused_tma<float,128>, adjoint step idx = 14 (15 resident + 170 streamed basis rows + 4 vectors + out),
# dram__bytes_read.sum + dram__bytes_write.sum per launch, next to the algorithmic bytes of that launch.
NCU_TRAFFIC = {"k_xdots_tma": {"traffic": 396.5e6 + 11.7e6, "algorithmic": 100 * 4.0e6,
                               "launch": "forward pass B, i=95 (96 streamed rows + 3 terms + out), fp32, n=1M",
                               "source": "profiles/r1c_prof_xdots.md"},
               "k_fused_tma": {"traffic": 395.3e6 + 6.8e6, "algorithmic": 98 * 4.0e6,
                               "launch": "forward pass B, i=95 (96 resident rows), fp32, n=1M (before k_xdots_tma)",
                               "source": "profiles/r1b_prof_sym_fused.md"}}


def algorithmic_bytes(n, nnz, K, w):
    """compulsory traffic with the best legal fusion, active columns only."""
    fwd = 3 * n * w * K * (K + 1) / 2 + K * (nnz * (w + 4) + 4 * (n + 1)) + 8 * K * n * w
    adj = 3 * n * w * K * (K + 1) + K * (nnz * (3 * w + 4) + 4 * (n + 1)) + 10 * K * n * w
    return fwd, adj


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _sample(self):
        nv = self._nvml
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            names = {
                "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
            }
            for name, bit in names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self._sample()
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "source": "unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml"}  # fmt: skip


def build_workload(seed=0):
    from experiments_lanczos_adjoints_b200 import synthetic

    row, col, data = synthetic.banded_spd_coo(N_ROWS, bands=BANDS, seed=seed)
    rng = np.random.default_rng(seed + 1)
    dalpha, dbeta = rng.standard_normal(DEPTH), rng.standard_normal(DEPTH - 1)
    return row, col, data, dalpha, dbeta


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(row, col, data, dalpha, dbeta, depth, dtype, seed=0):
    """One forward + adjoint of the reference's algorithm (oracle port) on the host CPUs."""
    from oracle import krylov, operators

    n = N_ROWS
    op = operators.CsrFastOperator(row, col, (n, n))
    v = np.random.default_rng(seed + 7).standard_normal(n).astype(dtype)
    alg = krylov.tridiag(op, depth, reortho="full")
    t0 = time.perf_counter()
    ((Qt, _), (q_rem, b_rem)), pull = alg.vjp(v, data.astype(dtype))
    pull(((np.zeros_like(Qt), (dalpha[:depth].astype(dtype), dbeta[: depth - 1].astype(dtype))),
          (np.zeros_like(q_rem), np.zeros((), dtype))))  # fmt: skip
    return time.perf_counter() - t0


def cpu_baseline(row, col, data, dalpha, dbeta, sample_depth=16):
    cores = os.cpu_count() or 1
    t = cpu_reference_run(row, col, data, dalpha, dbeta, sample_depth, np.float32)
    return {
        "value": sample_depth / t, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"n={N_ROWS}, nnz={len(data)}, fp32, one forward+adjoint at Krylov depth {sample_depth} "
                  f"({t:.1f} s; cost per Krylov step grows with depth, so depth {DEPTH} is slower per step); "
                  "NumPy/SciPy oracle port, BLAS threads = all cores (JAX not installed: the reference "
                  "itself cannot run)",
    }  # fmt: skip


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    row, col, data, dalpha, dbeta = build_workload()
    depth = int(os.environ.get("BL_REF_DEPTH", 12))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(row, col, data, dalpha, dbeta, depth, np.float32)
    times = [cpu_reference_run(row, col, data, dalpha, dbeta, depth, np.float32) for _ in range(max(1, args.steps))]
    t = float(np.mean(times))
    value = depth / t
    cores = os.cpu_count() or 1
    sample = (f"each step = one forward+adjoint at Krylov depth {depth} (bounded sample of depth {DEPTH}) on "
              f"n={N_ROWS}, nnz={len(data)}, fp32; NumPy/SciPy oracle port of the reference algorithm "
              "(JAX is not installed, the reference itself cannot be imported)")  # fmt: skip
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"sparse SPD COO operator n={N_ROWS} nnz={len(data)} ({2 * BANDS + 1}/row), Lanczos full "
                               f"reortho depth {DEPTH}, forward + adjoint (cotangents on alpha/beta)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))  # fmt: skip


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--probes", type=int, default=int(os.environ.get("BL_BENCH_PROBES", 4)),
                    help="independent probe vectors in flight per GPU, each on its own stream (one step = all of them)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-timed steps only (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    # stdout carries exactly ONE line, the JSON: whatever libraries print on file descriptor 1 while the
    # benchmark runs (NCCL's version banner, for one) goes to stderr; the descriptor is restored for the result
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import experiments_lanczos_adjoints_b200 as bl
    from experiments_lanczos_adjoints_b200 import plan as bl_plan
    from experiments_lanczos_adjoints_b200 import synthetic

    bl.set_device(local_rank)
    dtype = np.float32 if args.dtype == "f32" else np.float64
    w = np.dtype(dtype).itemsize
    row, col, data, dalpha, dbeta = build_workload()
    nnz = len(data)
    # P independent probes per GPU (the Hutchinson / SLQ workload: probes are independent runs), each with its own
    # operator handle, plan and stream: while one run sits in a kernel's ramp or grid-wide reduction tail, the other
    # runs' kernels keep the memory system busy.  One "step" = one forward + adjoint of every probe.
    from experiments_lanczos_adjoints_b200 import device as bl_dev

    P = max(1, args.probes)
    # three or more runs in flight: one block per SM and kernel, so that kernels of different runs share an SM and
    # fill each other's ramps and reduction tails (4 probes: 4.8k -> 5.15k steps/s); one run alone wants two
    blocks_in_flight = 1 if P >= 3 and "BL_BLOCKS_PER_SM" not in os.environ else 0
    bl.set_blocks_per_sm(blocks_in_flight)
    plans = []
    for p in range(P):
        op = bl.operators.SparseOperator(row, col, (N_ROWS, N_ROWS))
        plans.append(bl_plan.TridiagAdjointPlan(op, DEPTH, dtype, stream=bl_dev.Stream()))
    plan = plans[0]

    # pinned host buffers for the end-to-end path; P probe vectors per rank
    p_host = bl_plan.pinned_empty((nnz,), dtype)
    p_host[:] = data
    dH_host = bl_plan.pinned_empty((DEPTH, DEPTH), dtype)
    dH_host[:] = synthetic.slq_cotangent_dH(dalpha, dbeta, dtype)
    v_hosts, outs = [], []
    for p, pl in enumerate(plans):
        v_host = bl_plan.pinned_empty((N_ROWS,), dtype)
        v_host[:] = np.random.default_rng(100 + rank * P + p).standard_normal(N_ROWS)
        v_hosts.append(v_host)
        outs.append((bl_plan.pinned_empty((DEPTH, DEPTH), dtype), bl_plan.pinned_empty((N_ROWS,), dtype),
                     [bl_plan.pinned_empty((nnz,), dtype)]))  # fmt: skip
        pl.set_vector(v_host)
        pl.set_params(p_host)
        pl.set_cotangent(dH_host)
    grad_ts = []
    if dist is not None:
        import torch

        grad_ts = [torch.as_tensor(pl.grads[0], device=f"cuda:{local_rank}") for pl in plans]

    def reduce_grads():  # probe sharding: the rank's probes are summed, then ONE all-reduce per step (hutchinson.py:54)
        for pl in plans:
            pl.stream.synchronize()
        total = grad_ts[0] if P == 1 else torch.stack(grad_ts).sum(0)
        dist.all_reduce(total)

    def step_device(active=None):
        for pl in active or plans:
            pl.run()
        if dist is not None:
            reduce_grads()

    def barrier():
        if dist is not None:
            import torch

            torch.cuda.synchronize()
            dist.barrier()
        bl.synchronize()

    def timed(fn, steps, active=None):
        active = active or plans
        barrier()
        e0, ends = bl.Event(), [bl.Event() for _ in active]
        launches0 = bl.launch_count()
        e0.record(active[0].stream)
        for _ in range(steps):
            fn()
        if dist is not None:
            import torch

            torch.cuda.synchronize()
        for pl, e1 in zip(active, ends):
            e1.record(pl.stream)
        for e1 in ends:
            e1.synchronize()
        barrier()
        ms = max(e0.elapsed_ms(e1) for e1 in ends)  # first stream's start -> last stream's end
        if dist is not None:
            import torch

            t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, bl.launch_count() - launches0

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, launches = timed(step_device, args.steps)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = world * P * DEPTH / (ms_per_step * 1e-3)

    if args.quick:
        if rank == 0:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
            print(json.dumps({"value": value, "ms_per_step": ms_per_step, "gpu_launches": launches, "probes": P, "quick": True}),
                  flush=True)
        return
    # end-to-end through the host-buffer entry point: H2D (v, params, dH) + fwd + adjoint + D2H
    io = {}

    def step_host():
        h2d = d2h = 0
        for pl, v_host, (out_H, out_dv, out_g) in zip(plans, v_hosts, outs):
            a, b = pl.run_host(v_host, [p_host], dH_host, out_H, out_dv, out_g, sync=False)
            h2d, d2h = h2d + a, d2h + b
        for pl in plans:
            pl.stream.synchronize()  # the step's results (H, dv, dparams of every probe) are on the host
        io["h2d"], io["d2h"] = h2d, d2h
        if dist is not None:
            reduce_grads()

    for _ in range(2):
        step_host()
    e2e_steps = max(2, args.steps // 2)
    ms_e2e, _ = timed(step_host, e2e_steps)
    e2e_value = world * P * DEPTH / (ms_e2e / e2e_steps * 1e-3)

    # one probe alone (the latency of a single forward + adjoint), for transparency next to the P-probe throughput
    bl.set_blocks_per_sm(0)  # a run alone: the default two blocks per SM
    step_device(plans[:1])
    ms_single, _ = timed(lambda: step_device(plans[:1]), 3, plans[:1])
    ms_single /= 3

    # per-kernel-class timing of one more step (events around every launch)
    prof = bl_plan.profile(plan.run)
    peak, peak_src = measured_peaks()
    fwd_b, adj_b = (P * b for b in algorithmic_bytes(N_ROWS, nnz, DEPTH, w))
    step_gbs = (fwd_b + adj_b) / (ms_per_step * 1e-3) / 1e9
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    dom_gbs = d["algorithmic_bytes"] / max(d["ms"], 1e-9) / 1e6
    prof_total = sum(c["ms"] for c in prof.values())
    streamed_b = P * sum(c["algorithmic_bytes"] for c in prof.values())  # per step = P probes
    symmetric = all(os.environ.get(k, "1")[:1] != "0" for k in ("BL_SYMMETRIC_FORWARD", "BL_SYMMETRIC_ADJOINT", "BL_XDOTS"))
    KERNEL_OF = {"dots": "k_dots_few" if symmetric else "k_dots_tma", "combine": "k_combine_tma",
                 "matvec": "k_sell_spmv_normalised", "vjp": "k_sell_vjp", "other": "k_scale_copy",
                 "fused": "k_xdots_tma" if symmetric else "k_fused_tma"}  # fmt: skip
    roofline = {
        "bound": "hbm", "kernel": KERNEL_OF[dom],
        "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "peak_source": peak_src,
        "traffic": NCU_TRAFFIC.get(KERNEL_OF[dom], {}).get("traffic"),
        "traffic_capture": NCU_TRAFFIC.get(KERNEL_OF[dom]), "launches_per_step": P * d["launches"], "avg_launch_ms": d["ms"] / max(1, d["launches"]),
        "share_of_step": d["ms"] / max(prof_total, 1e-9),
        # streamed = the bytes this build's kernels account for; contract = SURVEY 8(d)'s figure for the general
        # (non-symmetric) loops.  The symmetric loops of tridiag(reortho="full") skip part of the contract's traffic
        # (local first Gram-Schmidt pass, one Lambda row, banded Gamma), so contract_frac may exceed 1: the saving is
        # reported as such, the bandwidth claim is `frac` (streamed bytes / time / peak).
        "whole_step": {"streamed_gb": streamed_b / 1e9, "achieved": streamed_b / (ms_per_step * 1e-3) / 1e9,
                       "frac": streamed_b / (ms_per_step * 1e-3) / 1e9 / peak,
                       "contract_gb": (fwd_b + adj_b) / 1e9, "contract_equivalent_gbs": step_gbs,
                       "contract_frac": step_gbs / peak},
        "classes": {k: {"launches": c["launches"], "ms": round(c["ms"], 4),
                        "gbs": c["algorithmic_bytes"] / max(c["ms"], 1e-9) / 1e6} for k, c in prof.items()},
    }  # fmt: skip

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"sparse SPD COO operator n={N_ROWS} nnz={nnz} ({2 * BANDS + 1}/row), Lanczos full "
                            f"reortho depth {DEPTH}, forward + adjoint (cotangents on alpha/beta)",
                "per_gpu": f"{P} independent probe vectors per GPU per step, each on its own stream (one step = forward + "
                           f"adjoint of all {P}); their parameter cotangents are summed and all-reduced once per step",
                "probes_in_flight": P,
                "blocks_per_sm": {"timed_region": blocks_in_flight or 2, "single_probe_and_kernel_profile": 2},
                "single_probe": {"ms_per_forward_adjoint": ms_single, "krylov_steps_per_s": DEPTH / (ms_single * 1e-3)},
                "loops": ("symmetric loops of tridiag(reortho=full): BL_FWD_SYMMETRIC, BL_ADJ_SYMMETRIC, "
                          "BL_ADJ_TRIDIAG_COTANGENT (include/b200_lanczos.h); switched off by BL_SYMMETRIC_FORWARD=0 / "
                          "BL_SYMMETRIC_ADJOINT=0: " + ",".join(
                              f"{k}={os.environ[k]}" for k in ("BL_SYMMETRIC_FORWARD", "BL_SYMMETRIC_ADJOINT") if k in os.environ)),
                "l2": f"inputs larger than L2 (basis Q {DEPTH * N_ROWS * w / 1e6:.0f} MB + adjoint basis, 126 MB L2)",
            },
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                    "ms_per_step": ms_e2e / e2e_steps},
            "roofline": roofline,
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(row, col, data, dalpha, dbeta)
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
