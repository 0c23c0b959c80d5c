"""Benchmark of the frequency-sweep S-parameter path (BASELINE.json metric: S-matrix
frequency-point solves/sec; % FP64 peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c1|c2|c3|c4] [--impl ours|reference]

Headline workload = BASELINE.json configs[4]: the ~200-node ladder / transmission-line network with 8
ports on a 1 000 000-point sweep, STRONG-scaled: the million points are split over the N ranks
(torchrun, one rank per GPU; frequency slabs, no data-path collective).  One "step" = R back-to-back
passes of the hot path (stamp assembly -> complex solve -> S) over this rank's share, R chosen during
warm-up so that a step lasts >= 10 ms on the slowest rank (K = 20 steps then time >= 0.2 s); R is in
`config.passes_per_step`, `value` counts every pass.

  value     whole-job solves/s with inputs resident in HBM and S left in HBM (kernels only), CUDA events on
            the launching stream, max over ranks, L2 flushed between timed steps
  e2e       the same metric through the public API with host buffers -- N = 1: `Model.run_sp(f)`;
            N > 1: `sharding.run_sp_distributed` (every rank's GPU copies its block over its own PCIe link
            into ONE page-locked shared array, rank 0 ends up holding the whole S) -- H2D of f, kernels and
            D2H of S inside the timed region
  roofline  dominant kernel: executed FP64 flops of the factorisation / port recurrences against the DFMA
            peak measured in this run (MEASURED_PEAKS.json has no FP64 entry); `roofline_hbm`: algorithmic
            bytes 8 + 16 P^2 per solve against the measured HBM copy rate; `traffic` =
            ncu dram bytes per launch from profiles/traffic.json (capture of this build)
  parity    untimed: a strided subset of the timed e2e result against oracle.run_sp(..., "lu_batched")
  cpu_baseline   the oracle port of the reference's algorithm (numba prange + zgelsd per port per frequency,
            oracle/mna_oracle_numba.py) on the host cores, bounded sample
  per_config     the same measurements (shorter) for configs[0..4], configs[3] as 10 000 samples split over N

`--impl reference` times the CPU path alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "S-matrix frequency-point solves/sec (freq x MC samples)"
UNIT = "solves/s"
TARGET_STEP_MS = 10.0

DESCS = {
    "c1": "configs[0]: README demo (5th-order Cauer band-pass + SMD resistor), 2001 points 1.8-2.2 GHz",
    "c2": "configs[1]: demo2 Cauer band/low/high-pass set, 2-port, 100001-point sweep each",
    "c3": "configs[2]: synthetic 8-port Touchstone block in LC matching network, 1M freq points",
    "c4": "configs[3]: Monte-Carlo Cauer band-pass, 10k samples x 2001 freqs",
    "c5": "configs[4]: ladder/TL network (~200 nodes, 8 ports), 1M-point sweep",
}


# --------------------------------------------------------------------------------------------
# workloads: whole jobs; the caller shards them
# --------------------------------------------------------------------------------------------
def make_jobs(hv, workload: str):
    """-> list of (model, f[nF], drivers | None, values[nS, nR] | None): the WHOLE job."""
    import circuits
    if workload == "c2":
        return [(circuits.build_c2(hv, band, nF=8)[0], np.linspace(2e9, 3e9, 100_001), None, None)
                for band in ("bandpass", "lowpass", "highpass")]
    if workload == "c1":
        return [(circuits.build_c1(hv, nF=8)[0], np.linspace(1.8e9, 2.2e9, 2001), None, None)]
    if workload == "c3":
        npts = int(os.environ.get("HVB_C3_POINTS", "1000000"))
        return [(circuits.build_c3(hv, nF=8)[0], np.linspace(1e9, 6e9, npts), None, None)]
    if workload == "c4":
        m, f, mc = circuits.build_c4(hv)
        nS = int(os.environ.get("HVB_C4_SAMPLES", "10000"))
        rs = np.random.RandomState(7)
        vals = np.stack([rs.normal(r._mean, r._std, size=nS) for r in mc._random_numbers], axis=1)
        return [(m, np.ascontiguousarray(f), mc._random_numbers, vals)]
    if workload == "c5":
        npts = int(os.environ.get("HVB_C5_POINTS", "1000000"))
        return [(circuits.build_c5(hv, nF=8)[0], np.linspace(0.5e9, 6e9, npts), None, None)]
    raise SystemExit("unknown workload " + workload)


def shard_job(job, world: int, rank: int):
    """This rank's share of a job (strong scaling): whole samples when nS >= world, else a frequency slab."""
    from heavi_b200 import sharding as sh
    m, f, drivers, values = job
    nS = 1 if values is None else values.shape[0]
    s = sh.plan_shard(nS, len(f), world, rank)
    if s.axis == "sample":
        return (m, f, drivers, None if values is None else np.ascontiguousarray(values[s.lo:s.hi]))
    return (m, np.ascontiguousarray(f[s.lo:s.hi]), drivers, values)


def job_points(job) -> int:
    m, f, drivers, values = job
    return len(f) * (1 if values is None else values.shape[0])


# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                mhz = float(parts[1]); smax = float(parts[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(mhz)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm and self.rows:
            try:
                sm = [float(self.rows[-1][1].split(",")[1])]
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def ncu_traffic(workload: str):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this build, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------
def cpu_job_sample(job, budget_points: int):
    m, f, drivers, values = job
    n = max(1, min(len(f), budget_points))
    return np.ascontiguousarray(f[:: max(1, len(f) // n)][:n])


def time_cpu(jobs, steps: int, warmup: int, per_step_seconds: float = 4.0):
    """Bounded sample of the same workload through the oracle port; returns (solves/s, ms/step, info)."""
    from oracle import mna_oracle_numba as onb
    recs = [(j[0].records(), j[0].n_nodes, len(j[0].sources)) for j in jobs]
    for (r, N, M), j in zip(recs, jobs):          # the first call also triggers the JIT once
        onb.run_sp(r, N, M, cpu_job_sample(j, 4))
    t = time.perf_counter()
    for (r, N, M), j in zip(recs, jobs):
        onb.run_sp(r, N, M, cpu_job_sample(j, 32))
    rate = 32 * len(jobs) / max(time.perf_counter() - t, 1e-6)
    budget = int(max(32, min(max(len(j[1]) for j in jobs), rate * per_step_seconds / len(jobs))))
    samples = [cpu_job_sample(j, budget) for j in jobs]
    pts = sum(len(s) for s in samples)
    for _ in range(warmup):
        for (r, N, M), s in zip(recs, samples):
            onb.run_sp(r, N, M, s)
    t0 = time.perf_counter()
    for _ in range(steps):
        for (r, N, M), s in zip(recs, samples):
            onb.run_sp(r, N, M, s)
    dt = time.perf_counter() - t0
    info = {"cores": onb.threads(), "kind": "port",
            "sample": "%d of %d points per step (strided subset of the same sweep; one nominal sample for Monte-Carlo "
                      "workloads), oracle/mna_oracle_numba.py: numpy stamp assembly + numba prange zgelsd per port per "
                      "frequency like solvers/spfn.py:145-156" % (pts, sum(job_points(j) for j in jobs))}
    return pts * steps / dt, dt / steps * 1e3, info


# --------------------------------------------------------------------------------------------
def kernel_name(ki, cn) -> str:
    kc = ki["kernel_class"]
    if kc == 0:
        return "hvb_warp%s_kernel<%d,%d>" % ("2" if cn.rows_per_lane == 2 else "", ki["group"], ki["pmax"])
    if kc == 1:
        return "hvb_block_kernel<%d>" % ki["group"]
    if kc == 2:
        return "hvb_bandf_kernel<%d>" % ki["group"]
    form = ("row loop%s, %d shapes" % (" walked from both ends" if ki.get("gen_two_sided") else "", ki["gen_cases"])) if ki["gen_rolled"] == 1 else "straight-line"
    if ki.get("gen_hub"):
        return "hvb_gen_kernel (NVRTC, kernel E: N-port block + chains, n=%d, P=%d)" % (cn.plan.n_real, cn.plan.P)
    return "hvb_gen_kernel (NVRTC, kernel D: chain, %s, n=%d, P=%d)" % (form, cn.plan.n_real, cn.plan.P)


def executed_flops_per_solve(ki, plan) -> tuple[float, str]:
    """FP64 flops of the factorisation + substitutions the kernel really executes for one point (stamp
    values excluded)."""
    n, P = plan.n_real, plan.P
    if ki["kernel_class"] == 3:
        if ki.get("gen_hub"):
            return float(ki["gen_flops_per_solve"]), "counted by the generator: S->Y solve + chain eliminations + core solve + back substitution"
        return float(ki["gen_flops_per_solve"]), "counted by the generator: tridiagonal LU with partial pivoting + forward port recurrences"
    if ki["kernel_class"] == 2:
        B = ki["group"]
        return 8.0 * n * B * (2 * B + P) + 8.0 * n * 2 * B * P, "banded LU n*B*(2B+P) + back substitution n*2B*P complex FMAs"
    return plan.flops_per_solve(), "dense complex LU + P substitutions: 8/3 n^3 + 8 n^2 P, n = dimension factored"


class Measurement:
    """One workload on this rank's GPU: device-resident state, kernel-only timing, e2e timing."""

    def __init__(self, hv, workload, world, rank, dev, dist):
        import torch
        self.torch, self.dist, self.dev, self.world, self.rank, self.workload, self.hv = torch, dist, dev, world, rank, workload, hv
        self.jobs_whole = make_jobs(hv, workload)
        self.jobs = [shard_job(j, world, rank) for j in self.jobs_whole]
        self.points_whole = sum(job_points(j) for j in self.jobs_whole)
        self.points_rank = sum(job_points(j) for j in self.jobs)
        self.state = []
        for (m, f, drivers, values) in self.jobs:
            cn = m._compiled()
            whole = self.points_whole                       # every rank picks the kernel variant of the WHOLE job
            plan, handle = cn.variant(whole)
            consts, prefs, tables, samples = cn.resolve(f, drivers, values, plan)
            P = plan.P
            self.state.append(dict(cn=cn, plan=plan, handle=handle, consts=consts, prefs=prefs,
                                   d_f=torch.from_numpy(f).to(dev),
                                   d_tab=torch.from_numpy(tables).to(dev) if tables.shape[0] else None,
                                   d_smp=torch.from_numpy(samples).to(dev),
                                   d_S=torch.empty((samples.shape[0], P, P, len(f)), dtype=torch.complex128, device=dev),
                                   d_status=torch.zeros(1, dtype=torch.int32, device=dev)))

        # a two-sided chain kernel may refuse a point (status bit 4, include/heavi_b200.h): such a plan is put back on the
        # one-sided kernel BEFORE anything is timed, exactly what the host entries of the library do on their own
        self.launch_all()
        torch.cuda.synchronize()
        for st in self.state:
            if int(st["d_status"].item()) & 4:
                st["handle"].set_tunable("gen_two_sided", 0)
            st["d_status"].zero_()

    def launch_all(self):
        for st in self.state:
            if st["d_f"].numel():
                st["handle"].run_device(st["d_f"], st["consts"], st["prefs"], st["d_tab"], st["d_smp"], st["d_S"], st["d_status"])

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def launches(self) -> int:
        return sum(st["handle"].kernel_info()["launches"] for st in self.state)

    def time_kernels(self, steps, warmup, flush, sampler=None):
        torch = self.torch
        for _ in range(warmup):
            self.launch_all()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); self.launch_all(); b.record(); torch.cuda.synchronize()
        t_pass = self.allmax(a.elapsed_time(b))
        passes = max(1, int(math.ceil(TARGET_STEP_MS / max(t_pass, 1e-4))))
        passes = min(passes, 4096)
        l0 = self.launches()
        if sampler:
            sampler.start()
            time.sleep(0.25)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        t_wall0 = time.time()
        for k in range(steps):
            flush.zero_()                                                   # evict L2 (untimed)
            ev[k][0].record()
            for _ in range(passes):
                self.launch_all()
            ev[k][1].record()
        self.barrier()
        t_wall1 = time.time()
        dev_ms = self.allmax(sum(x.elapsed_time(y) for x, y in ev))
        status = int(sum(int(st["d_status"].item()) for st in self.state))
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        return dict(value=self.points_whole * passes * steps / (dev_ms * 1e-3), ms_per_step=dev_ms / steps, passes=passes,
                    ms_per_pass=dev_ms / steps / passes, launches=self.launches() - l0, status=status, clocks=clocks)

    # ---- e2e: public API, host buffers ----------------------------------------------------------
    def e2e_once(self):
        """One pass of the whole job through the public API; returns the results rank 0 holds."""
        from heavi_b200 import sharding
        if self.world == 1 and len(self.jobs_whole) > 1 and all(j[3] is None for j in self.jobs_whole):
            # several independent sweeps (configs[1]'s filter set): queued together, collected together
            return [S._S[None] for S in self.hv.run_sp_many([(j[0], j[1]) for j in self.jobs_whole])]
        out = []
        for (m, f, drivers, values) in self.jobs_whole:
            if self.world > 1:
                out.append(sharding.run_sp_distributed(m, f, drivers, values))
            elif values is None:
                out.append(m.run_sp(f)._S[None])
            else:
                out.append(m.run_sp_batch(f, drivers, values))
        return out

    def time_e2e(self, min_seconds=0.25, min_steps=3, max_steps=400):
        for _ in range(2):
            res = self.e2e_once()
        self.barrier()
        t0 = time.perf_counter()
        res = self.e2e_once()
        acc = float(res[0].reshape(-1)[-1].real) if res[0] is not None else 0.0
        self.barrier()
        t_one = self.allmax(time.perf_counter() - t0)
        steps = int(min(max_steps, max(min_steps, math.ceil(min_seconds / max(t_one, 1e-6)))))
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = self.e2e_once()
            if res[0] is not None:
                acc += float(res[0].reshape(-1)[-1].real)                   # the host reads the step's result
        self.barrier()
        dt = self.allmax(time.perf_counter() - t0)
        h2d = sum(j[1].nbytes + (0 if j[3] is None else j[3].nbytes) for j in self.jobs_whole)
        d2h = sum(job_points(j) * st["plan"].P ** 2 * 16 for j, st in zip(self.jobs_whole, self.state))
        api = ("sharding.run_sp_distributed(model, f) -> S on rank 0 (one page-locked shared array, each GPU copies its block)"
               if self.world > 1 else ("heavi_b200.run_sp_many([(model, f), ...]) -> [Sparameters] (sweeps queued together, pinned host results)"
                                       if len(self.jobs_whole) > 1 else "Model.run_sp(f) / run_sp_batch -> Sparameters (pinned host result)"))
        return dict(value=self.points_whole * steps / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                    steps=steps, ms_per_step=dt / steps * 1e3, api=api), res

    # ---- parity spot check of the timed e2e result (rank 0, untimed) ----------------------------
    def parity(self, results, max_points):
        from oracle import mna_oracle as orc
        worst, checked, worst_rel = 0.0, 0, 0.0
        for (m, f, drivers, values), S in zip(self.jobs_whole, results):
            if S is None:
                continue
            n = max(1, min(len(f), max_points))
            idx = np.unique(np.linspace(0, len(f) - 1, n).astype(np.int64))
            rows = [0] if values is None else sorted({0, values.shape[0] // 2, values.shape[0] - 1})
            saved = None if values is None else [d._value for d in drivers]
            for s in rows:
                if values is not None:
                    for d, v in zip(drivers, values[s]):
                        d._value = float(v)
                ref = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f[idx], method="lu_batched")
                err = np.abs(S[s][..., idx] - ref)
                worst = max(worst, float(err.max()))
                worst_rel = max(worst_rel, float((err / (1e-12 + 1e-10 * np.abs(ref))).max()))
                checked += len(idx)
            if saved is not None:
                for d, v in zip(drivers, saved):
                    d._value = v
        return {"checked_points": checked, "max_abs_err": worst, "max_err_over_tolerance": worst_rel,
                "tolerance": "1e-12 + 1e-10 |S| against oracle.run_sp(method='lu_batched')", "ok": bool(worst_rel <= 1.0)}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    def roofline(self, flush, peak_tflops, reps=8):
        torch = self.torch
        cost = [executed_flops_per_solve(st["handle"].kernel_info(), st["plan"])[0] * job_points(j) for j, st in zip(self.jobs, self.state)]
        top = int(np.argmax(cost))
        st, job = self.state[top], self.jobs[top]
        h = st["handle"]
        single = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h.run_device(st["d_f"], st["consts"], st["prefs"], st["d_tab"], st["d_smp"], st["d_S"], st["d_status"])
            b.record()
            torch.cuda.synchronize()
            single.append(a.elapsed_time(b))
        k_ms = float(np.mean(single))
        ki = h.kernel_info()
        plan = st["plan"]
        pts = job_points(job)
        fl, how = executed_flops_per_solve(ki, plan)
        achieved = fl * pts / (k_ms * 1e-3) / 1e12
        mp = measured_peaks()
        hbm_peak = mp["hbm_gbs"] if mp else 6650.0
        bytes_alg = 8 + 16 * plan.P ** 2
        roof = {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": ncu_traffic(self.workload), "kernel": kernel_name(ki, st["cn"]), "launch_ms": k_ms,
                "n_factored": plan.n_real, "solver_mode": plan.mode, "flops_per_solve": fl, "flops_basis": how,
                "dense_lu_equivalent_flops_per_solve": plan.flops_per_solve(), "solves_per_launch": pts,
                "peak_source": "DFMA probe measured in this run (hvb_measure_dfma_peak); MEASURED_PEAKS.json has no FP64 entry",
                "regs": ki["regs"], "ctas_per_sm": ki["ctas_per_sm"], "smem_bytes": ki["smem_bytes"], "local_bytes": ki["local_bytes"]}
        roof_hbm = {"bound": "hbm", "achieved": bytes_alg * pts / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg * pts / (k_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_solve": bytes_alg,
                    "traffic": ncu_traffic(self.workload),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if mp else "fallback 6650 GB/s (B200_PROFILING.md)"}
        return roof, roof_hbm


def workload_config(workload, world, m: Measurement, passes):
    return {"workload": DESCS[workload] + (" -- STRONG-scaled over %d GPUs" % world if world > 1 else ""),
            "solves_per_pass": m.points_whole, "solves_per_pass_per_gpu": m.points_rank, "passes_per_step": passes,
            "l2": "flushed between timed steps (256 MiB memset); S of one pass (%d MB per GPU) is rewritten by the next"
                  % (sum(st["d_S"].numel() * 16 for st in m.state) // 1000000),
            "sharding": ("frequency slab / sample block per rank, no data-path collective" if world > 1 else "single GPU")}


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("HVB_WORKLOAD", "c5"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-per-config", action="store_true", help="headline workload only")
    args = ap.parse_args()

    # the contract is ONE JSON line on stdout: whatever libraries print to fd 1 (NCCL's version banner, ...)
    # goes to stderr instead; the line itself is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        # one process = one GPU here; N > 1 is launched by torchrun (one rank per GPU).  The single-process
        # multi-device mode of Model.run_sp is measured by tools/multi_device_bench.py.
        os.environ["HVB_DEVICES"] = "1"
        if args.gpus > 1:
            print("bench.py: --gpus %d without torchrun: measuring 1 GPU" % args.gpus, file=sys.stderr)
    import heavi_b200 as hv

    if args.impl == "reference":
        if rank != 0:
            return
        jobs = make_jobs(hv, args.workload)
        steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
        val, ms, info = time_cpu(jobs, steps, warmup)
        info["value"] = val
        emit({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": DESCS[args.workload], "cpu_path": "reference algorithm (zgelsd per port per frequency) on host cores"},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (heavi_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from heavi_b200 import engine

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    peak = engine.measure_dfma_peak(local_rank)

    def run_workload(wl, k_steps, k_warm, headline):
        m = Measurement(hv, wl, world, rank, dev, dist)
        sampler = ClockSampler(local_rank) if headline else None
        kt = m.time_kernels(k_steps, k_warm, flush, sampler)
        e2e, results = m.time_e2e(min_seconds=0.25 if headline else 0.1)
        out = None
        if rank == 0:
            roof, roof_hbm = m.roofline(flush, peak, reps=8 if headline else 4)
            par = m.parity(results, 48 if wl in ("c5", "c3") else 512)
            cpu = None
            if not args.no_cpu and world == 1:
                v, ms, info = time_cpu(m.jobs_whole, 2 if headline else 1, 1 if headline else 0, per_step_seconds=4.0 if headline else 2.0)
                info["value"], info["unit"] = v, UNIT
                cpu = info
            out = {"value": kt["value"], "unit": UNIT, "ms_per_step": kt["ms_per_step"], "ms_per_pass": kt["ms_per_pass"],
                   "config": workload_config(wl, world, m, kt["passes"]), "e2e": e2e, "gpu_launches": int(kt["launches"]),
                   "status_word": kt["status"], "roofline": roof, "roofline_hbm": roof_hbm, "parity": par, "cpu_baseline": cpu,
                   "clocks": kt["clocks"]}
        del m, results
        torch.cuda.empty_cache()
        return out

    head = run_workload(args.workload, steps, warmup, True)
    per = {}
    if not args.no_per_config:
        for wl in ("c1", "c2", "c3", "c4", "c5"):
            if wl == args.workload:
                continue
            try:
                r = run_workload(wl, min(steps, 10), 3, False)
            except Exception as exc:                              # a side workload must not take the headline line down
                if world > 1:
                    raise                                          # ranks would diverge: fail loudly under torchrun
                r = {"error": "%s: %s" % (type(exc).__name__, exc)}
            if rank == 0:
                r.pop("clocks", None)
                per[wl] = r
    if rank == 0:
        brief = {k: head[k] for k in ("value", "unit", "ms_per_step", "ms_per_pass", "config", "e2e", "gpu_launches", "status_word",
                                      "roofline", "roofline_hbm", "parity", "cpu_baseline")}
        per[args.workload] = brief
        out = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic", "config": head["config"], "clocks": head["clocks"], "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "status_word": head["status_word"], "roofline": head["roofline"],
            "roofline_hbm": head["roofline_hbm"], "parity": head["parity"], "cpu_baseline": head["cpu_baseline"],
            "per_config": {k: per[k] for k in sorted(per)},
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
