#!/usr/bin/env python
"""Benchmark of the LM hot path (BASELINE.json metric: LM iterations/s + HBM GB/s of build / Schur / PCG on a
BAL-Venice-shaped bundle adjustment).

    python bench.py --gpus N --steps K --warmup W          # our CUDA backend (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W  # the reference itself (oracle/_ref/libg2o_ref_core.so, else the oracle port) on the host cores

A step is one outer Levenberg-Marquardt iteration (`OptimizationAlgorithmLevenberg::solve`, all its trials) of
config C3: 1778 cameras / 993 923 points / 5 001 946 observations, BlockSolver<9,3> + PCG, Huber(1.0), synthetic.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lm_iterations_per_second"
UNIT = "LM iterations/s"


def workload(name: str, scale: float):
    from g2o_b200 import workloads as W
    if name == "bal_venice":
        if scale == 1.0:
            return W.bal_venice(), "C3 bal_venice: 1778 cameras / 993923 points / 5001946 observations, Huber(1.0), BlockSolver<9,3>+PCG"
        nc = max(16, int(1778 * scale ** 0.5)); npnt = int(993_923 * scale); nobs = int(5_001_946 * scale)
        return (W.bal_synthetic(n_cameras=nc, n_points=npnt, n_obs=nobs, k_max=min(500, nc)),
                f"C3 bal_venice scaled x{scale}: {nc} cameras / {npnt} points / {nobs} observations")
    if name == "bal_large":
        return W.bal_large(), "C4 bal_large: 10000 cameras / 4000000 points / 20000000 observations"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """SM clock / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe).  NVML through pynvml (a sample
    every few ms; the timed region of the default run is ~80 ms), nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index; self.samples = []; self._stop = threading.Event(); self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        n = self._nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append([str(mhz), str(self._max), "", *["Active" if r & b else "Not Active" for b in bits.values()]])
            except Exception:
                pass
            self._stop.wait(0.004)

    def _run(self):
        if self._nvml is not None:
            return self._run_nvml()
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_tensor_peak():
    """FP64 tensor-pipe (DMMA) peak in TFLOP/s.  MEASURED_PEAKS.json only holds the bf16 figure, so the denominator is the number
    tools/fp64_pipes.cu measured on this pool's B200 (committed under profiles/); 148 SMs x 64 FMA/clk x 1.965 GHz x 2 = 37.2."""
    p = os.path.join(ROOT, "profiles", "fp64_peaks_b200.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["dmma_tflops"]), "measured (profiles/fp64_peaks_b200.json, tools/fp64_pipes.cu)"
    return 37.2, "nominal (148 SMs x 64 FP64 FMA/clk x 1.965 GHz)"


def run_reference(args):
    """The reference's own CPU implementation on the host cores.  When oracle/_ref/libg2o_ref_core.so exists (built from /root/reference by
    `make -C oracle ref_core`: g2o/core + BlockSolver + LinearSolverPCG + the types, unmodified, against the stand-in for the absent Eigen3,
    OpenMP on) that is the real reference: SparseOptimizer::optimize with OptimizationAlgorithmLevenberg over BlockSolver<9,3> + PCG, as
    examples/bal/bal_example.cpp sets it up.  Otherwise the oracle port.  Bounded sample: at most 1 warm-up + 3 timed LM iterations of the
    full workload (iteration 0 also pays buildStructure and is never timed); per-iteration times are G2OBatchStatistics::timeIteration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    if orc.reference_core() is not None and not os.environ.get("G2O_BENCH_REFERENCE_CHILD"):
        # the leg that drives the compiled reference runs in a child process; if it dies, this process still reports the oracle port
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], capture_output=True, text=True, timeout=1500,
                               env=dict(os.environ, G2O_BENCH_REFERENCE_CHILD="1", RANK="0"))
            if r.returncode == 0 and r.stdout.strip():
                print(r.stdout.strip().splitlines()[-1], flush=True)
                return
            sys.stderr.write(f"bench: reference child failed (exit {r.returncode}): {r.stderr[-400:]}\n")
        except Exception as e:
            sys.stderr.write(f"bench: reference child failed: {e}\n")
        orc.reference_core = lambda: None
    g, desc = workload(args.workload, args.scale)
    threads = orc.max_threads()
    warm, steps = 1, max(1, min(args.steps, 3))

    def timed_part(stats):
        times = [s["timeIteration"] for s in stats]
        timed = times[warm:] if len(times) > warm else times
        return timed, (len(timed) / sum(timed) if timed and sum(timed) > 0 else 0.0)

    kind, solver, extra = "port", "CPU oracle port of BlockSolver + LinearSolverPCG (oracle/g2o_oracle.cpp)", {}
    stats = None
    if orc.reference_core() is not None:
        vt = set(int(t) for t in np.unique(g.v_type)); marg = bool(np.any(g.v_marginalized))
        bs = "9_3" if marg and 6 in vt else "6_3" if marg and 4 in vt else "3_2" if marg and 1 in vt else "var"
        try:
            ref = orc.ReferenceG2o(g, "lm", bs, threads=threads)
            if ref.initialize_optimization():
                n, stats = ref.optimize(warm + steps)
                kind = "reference"
                solver = f"the reference itself: SparseOptimizer + OptimizationAlgorithmLevenberg + BlockSolver<{bs.replace("_", ",")}> + LinearSolverPCG compiled from /root/reference (Eigen3 replaced by oracle/eigen_shim, OpenMP, -O3)"
        except Exception as e:      # fall back to the port, say why
            extra["reference_error"] = str(e)[:200]
            stats = None
    if stats is None:
        o = orc.Oracle(g, "lm", "pcg", threads=threads)
        o.initialize_optimization()
        n, stats = o.optimize(warm + steps)
    timed, value = timed_part(stats)
    if kind == "reference" and not args.no_cpu:      # the port on the same sample, for comparison
        try:
            o = orc.Oracle(g, "lm", "pcg", threads=threads)
            o.initialize_optimization()
            _, pstats = o.optimize(warm + 1)          # one timed iteration is enough for the side-by-side
            ptimed, pvalue = timed_part(pstats)
            extra["port"] = {"value": pvalue, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": 1e3 * sum(ptimed) / max(len(ptimed), 1),
                             "chi2": [s["chi2"] for s in pstats]}
        except Exception as e:
            extra["port_error"] = str(e)[:200]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed), "warmup": min(warm, len(stats)),
            "ms_per_step": 1e3 * sum(timed) / max(len(timed), 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": desc, "solver": solver},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{len(timed)} LM iteration(s) of the full workload after {min(warm, len(stats))} warm-up iteration(s) (requested steps={args.steps}, warmup={args.warmup}; capped to keep the run within minutes)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "chi2": [s["chi2"] for s in stats],
            "phases_s": {k: stats[-1][k] for k in ("timeResiduals", "timeQuadraticForm", "timeSchurComplement", "timeLinearSolver", "timeUpdate")} if stats else None}
    line.update(extra)
    print(json.dumps(line), flush=True)


def cpu_sample(g):
    """cpu_baseline of the main arm: LM iteration 1 of the same graph on all host threads (steady state: iteration 0 additionally pays
    buildStructure, which the GPU arm also keeps outside its timed region) - the real reference when oracle/_ref/libg2o_ref_core.so is there
    (kind "reference"), else the oracle port (kind "port").  Never raises: a failure of the reference leg falls back to the port."""
    import numpy as np
    from oracle import oracle as orc
    threads = orc.max_threads()
    t0 = time.perf_counter()
    kind, cstats, note = "port", None, ""
    if orc.reference_core() is not None:
        try:
            vt = set(int(t) for t in np.unique(g.v_type)); marg = bool(np.any(g.v_marginalized))
            bs = "9_3" if marg and 6 in vt else "6_3" if marg and 4 in vt else "3_2" if marg and 1 in vt else "var"
            ref = orc.ReferenceG2o(g, "lm", bs, threads=threads)
            if ref.initialize_optimization():
                _, cstats = ref.optimize(2)
                kind = "reference"
                note = f"; the reference itself (BlockSolver<{bs.replace('_', ',')}> + LinearSolverPCG compiled from /root/reference against oracle/eigen_shim, OpenMP)"
        except Exception as e:
            cstats, note = None, f"; reference leg failed ({str(e)[:80]}), oracle port timed instead"
    if not cstats:
        kind = "port"
        o = orc.Oracle(g, "lm", "pcg", threads=threads)
        o.initialize_optimization()
        _, cstats = o.optimize(2)
    secs = time.perf_counter() - t0
    it = cstats[-1]["timeIteration"] if cstats else float("nan")
    out = {"value": 1.0 / it, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"LM iteration 1 of the same graph ({it:.2f} s; graph construction + optimize(2) took {secs:.1f} s incl. buildStructure in iteration 0){note}",
           "chi2": [c["chi2"] for c in cstats]}
    if cstats:      # G2OBatchStatistics of the timed iteration (both the reference and the port fill the same fields)
        out["phases_s"] = {k: cstats[-1][k] for k in ("timeResiduals", "timeQuadraticForm", "timeSchurComplement", "timeLinearSolver", "timeUpdate")}
    return out


def cpu_sample_isolated(args, g):
    """cpu_sample of the reference in a child process (same workload, regenerated there from its seed): whatever happens in that leg - it
    builds 6 million g2o objects for C3 - cannot take the GPU measurement down with it.  Falls back to the oracle port in this process."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu_sample", "--workload", args.workload, "--scale", str(args.scale)],
                           capture_output=True, text=True, timeout=900)
        if r.returncode == 0 and r.stdout.strip():
            return json.loads(r.stdout.strip().splitlines()[-1])
        why = f"exit code {r.returncode}: {r.stderr[-160:]}"
    except Exception as e:
        why = str(e)[:160]
    from oracle import oracle as orc
    saved = orc.reference_core
    orc.reference_core = lambda: None          # the port only
    try:
        out = cpu_sample(g)
    finally:
        orc.reference_core = saved
    out["sample"] += f"; reference leg in a child process failed ({why})"
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from g2o_b200.binding import CudaSolver

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    g, desc = workload(args.workload, args.scale)
    est0 = g.v_estimate.copy()
    s = CudaSolver(g, "lm_fix9_3_cuda", device=local)
    if world > 1:
        from g2o_b200.dist import install_nccl, install_torch_allreduce
        if os.environ.get("G2O_BENCH_COLLECTIVES", "nccl") == "hook":
            install_torch_allreduce(s, rank, world)      # every collective through the Python callback (debug)
        else:
            install_nccl(s, rank, world)                 # collectives issued by libg2ocu.so itself
    s.initialize_optimization()
    s.init()
    if world > 1 and os.environ.get("G2O_BENCH_P2P", "1") != "0":
        from g2o_b200.dist import install_p2p
        s.build_structure()
        install_p2p(s, rank, world)                      # q = A d of the slab PCG is exchanged through NVLink peer memory

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_run(e2e: bool):
        """W untimed + K timed LM iterations from the initial estimates; returns (seconds, stats, launches, phases)."""
        pin_in = torch.from_numpy(est0.copy()).pin_memory()
        pin_out = torch.empty_like(pin_in).pin_memory()
        s.set_estimates(est0)
        s.init()
        host_in, host_out = pin_in.numpy(), pin_out.numpy()      # two pinned host buffers used in turn: a step's output is the next step's input
        stats = []
        for i in range(args.warmup):
            if e2e:
                s.set_estimates(host_in)
            stats.append(s.solver_iteration(i))
            if e2e:
                s.get_estimates(host_out); host_in, host_out = host_out, host_in
        s.reset_counters()
        sampler = ClockSampler(local); sampler.start()
        barrier()
        l0 = s.launch_count()
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            if e2e:
                s.set_estimates(host_in)                       # H2D of this step's inputs (pinned)
            stats.append(s.solver_iteration(i))
            if e2e:
                s.get_estimates(host_out); host_in, host_out = host_out, host_in   # D2H of the step's result (estimates + chi2); it feeds the next step
        barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        phases = {ph: s.phase_time(ph) for ph in ["errors", "build", "schur", "schur_coeff", "schur_pairs", "schur_tiles", "schur_exchange", "pcg_setup", "pcg_spmv",
                                                  "pcg_exchange", "pcg_vec", "linear_solver", "backsub", "update"]}
        return dt, stats, s.launch_count() - l0, phases, clocks

    dt, stats, launches, _, clocks = timed_run(False)          # headline: phase-level events only
    dt_e, stats_e, _, _, _ = timed_run(True)
    s.set_property("kernelTiming", 1.0)                         # breakdown pass: the same steps with CUDA events around every kernel
    dt_k, _, _, phases, _ = timed_run(False)
    s.set_property("kernelTiming", 0.0)
    if world > 1:
        t = torch.tensor([dt, dt_e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    dims = s.get_i32("dims"); nnzS = int(s.get_i32("hschur_colptr")[-1]); nc, npnt = int(dims[0]), int(dims[1]); ne = g.n_edges
    peak, peak_src = measured_peaks()
    # algorithmic bytes per launch: P=9, L=3, E=2
    bytes_spmv = nnzS * 81 * 8 + nc * 81 * 8 + 10 * nc * 9 * 8
    bytes_build = ne * (16 + 8 + 216) + npnt * (24 + 72 + 24) + nc * (72 + 648 + 72)
    bytes_schur = ne * 216 + npnt * (72 + 24) + nc * (648 + 72) + npnt * 72 + nnzS * 648 + nc * 72
    # Schur tile kernel: FP64 tensor pipe.  Algorithmic flops = 2 P P L per (landmark, camera pair i <= j) of the tracks it handles (>= 8 observations)
    k = np.bincount(np.asarray(g.e_v1) - nc, minlength=npnt).astype(np.int64)
    pairs_all = int((k * (k + 1) // 2).sum()); kt = k[k >= 8]; pairs_tiles = int((kt * (kt + 1) // 2).sum())
    flops_tiles = 2.0 * 81 * 3 * pairs_tiles
    bytes_coeff = ne * 216 * 2 + npnt * (72 + 24) + nc * 72        # read Hpl, write W = Hpl Dinv, read Dinv/db, update b_schur
    per = {}
    for name, nbytes in [("pcg_spmv", bytes_spmv), ("build", bytes_build), ("schur", bytes_schur), ("schur_coeff", bytes_coeff)]:
        sec, _, calls = phases[name]
        if calls:
            per[name] = {"seconds_total": sec, "calls": calls, "avg_ms": 1e3 * sec / calls, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (sec / calls) / 1e9,
                         "frac_of_hbm_peak": nbytes / (sec / calls) / 1e9 / peak}
    if "pcg_spmv" in per:   # launches issued after convergence return at once: rate per ACTIVE product = per PCG iteration
        its = sum(st["iterations_linear_solver"] for st in stats[args.warmup:])
        if its:
            per["pcg_spmv"].update({"active_products": its, "avg_ms_active": 1e3 * per["pcg_spmv"]["seconds_total"] / its,
                                    "achieved_gbs": bytes_spmv / (per["pcg_spmv"]["seconds_total"] / its) / 1e9,
                                    "frac_of_hbm_peak": bytes_spmv / (per["pcg_spmv"]["seconds_total"] / its) / 1e9 / peak})
    tpeak, tpeak_src = fp64_tensor_peak()
    for name, fl in [("schur_tiles", flops_tiles), ("schur_pairs", 2.0 * 81 * 3 * (pairs_all - pairs_tiles))]:
        sec, _, calls = phases[name]
        if calls:
            per[name] = {"seconds_total": sec, "calls": calls, "avg_ms": 1e3 * sec / calls, "algorithmic_flops": fl, "achieved_tflops": fl / (sec / calls) / 1e12,
                         "frac_of_fp64_peak": fl / (sec / calls) / 1e12 / tpeak}
    kernels = {"pcg_spmv": "spmv_sym_kernel<9>", "build": "build_pl_kernel<BAL> + pose_accum_kernel<BAL>", "schur_coeff": "coeff_w_kernel<9,3>",
               "schur_tiles": "schur_mma_kernel<9,3>", "schur_pairs": "schur_pairs_kernel<9,3>"}
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            traffic = json.load(fh)
    cand = [k2 for k2 in per if k2 in kernels]
    dominant = max(cand, key=lambda k2: per[k2]["seconds_total"]) if cand else None
    roof = None
    if dominant and "algorithmic_flops" in per[dominant]:
        a = per[dominant]["achieved_tflops"]
        roof = {"kernel": kernels[dominant], "bound": "tensor", "achieved": a, "peak": tpeak, "unit": "TFLOP/s", "frac": a / tpeak, "traffic": traffic.get(kernels[dominant]), "peak_source": tpeak_src,
                "note": "FP64 tensor pipe (DMMA m8n8k4); FP64 FMA shares the same pipe on B200 (tools/dmma_dfma_mix.cu), so this is the only FP64 roof",
                "avg_launch_ms": per[dominant]["avg_ms"], "phases": per}
    elif dominant:
        a = per[dominant]["achieved_gbs"]
        roof = {"kernel": kernels[dominant], "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": traffic.get(kernels[dominant]), "peak_source": peak_src,
                "avg_launch_ms": per[dominant]["avg_ms"], "phases": per}
    timed_stats = stats[args.warmup:]
    value = args.steps / dt
    est_bytes = int(est0.nbytes)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "solver": "lm_fix9_3_cuda (Schur + block-Jacobi PCG)", "l2_policy": "working set (Hpl 1.08 GB, Hschur %.2f GB) exceeds the 126 MB L2" % (nnzS * 648 / 1e9),
                       "nnz_hschur_blocks": nnzS, "parallelism": f"landmark-sharded x{world}" if world > 1 else "single GPU"},
            "e2e": {"value": args.steps / dt_e, "unit": UNIT, "h2d_bytes_per_step": est_bytes, "d2h_bytes_per_step": est_bytes + 8,
                    "note": "per step: host vertex estimates -> device (pinned), one LM iteration through g2ocu_solver_iteration, estimates + chi2 back to the host"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "phase_ms_per_step": {ph: round(1e3 * v[0] / args.steps, 4) for ph, v in phases.items()},
            "phase_note": "from a separate pass of the same steps with per-kernel CUDA events (%.3f ms per step in that pass)" % (1e3 * dt_k / args.steps),
            "lm": {"chi2": [st["chi2"] for st in stats], "lambda": [st["lambda"] for st in stats], "trials": [st["levenberg_iterations"] for st in stats],
                   "pcg_iterations": [st["iterations_linear_solver"] for st in stats]}}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_sample_isolated(args, g)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu_sample"])   # cpu_sample: internal, the CPU leg of the main arm in a child process
    ap.add_argument("--workload", default="bal_venice")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("G2O_BENCH_ALLOW_SHORT_WARMUP"):
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cpu_sample":
        g, _ = workload(args.workload, args.scale)
        print(json.dumps(cpu_sample(g)), flush=True)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
