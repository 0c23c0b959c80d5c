"""Benchmark of the frequency-sweep S-parameter path (BASELINE.json metric: S-matrix
frequency-point solves/sec; % FP64 peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference]

One "step" = one pass of the hot path (stamp assembly -> complex solve -> S) over one batch of
synthetic input.  Default workload = BASELINE.json configs[1]: the demo2 Cauer filter set (band-,
low- and high-pass, 2 ports) on a 100 001-point sweep each = 300 003 solves per step per GPU.
With N > 1 (launched by torchrun, one rank per GPU) the sweep is N times longer and rank r takes
frequency block r (weak scaling, no data-path collective).

  value     whole-job solves/s with inputs resident in HBM and S left in HBM (kernel only),
            CUDA events on the launching stream, L2 flushed between timed steps
  e2e       the same metric through the public API `Model.run_sp(f)` with host buffers: H2D of f,
            kernels, D2H of S into pinned memory all inside the timed region
  roofline  dominant kernel against the FP64 FMA pipe (the bound SURVEY.md section 8d derives);
            `peak` is a DFMA probe measured in this run because MEASURED_PEAKS.json has no FP64
            entry; `roofline_hbm` is the same launch against the measured HBM copy rate
  cpu_baseline   the oracle port of the reference's algorithm (numba prange + zgelsd per port per
            frequency, oracle/mna_oracle_numba.py) on the host cores, bounded sample

`--impl reference` times that CPU path alone (rank 0 only under torchrun).
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "S-matrix frequency-point solves/sec (freq x MC samples)"
UNIT = "solves/s"


# --------------------------------------------------------------------------------------------
# workloads; each returns a list of jobs (model, f, drivers, values)
# --------------------------------------------------------------------------------------------
def make_jobs(hv, workload: str, rank: int, world: int):
    import circuits
    jobs = []

    def slab(f_lo, f_hi, n_per_gpu):
        full = np.linspace(f_lo, f_hi, n_per_gpu * world)
        return np.ascontiguousarray(full[rank * n_per_gpu:(rank + 1) * n_per_gpu])

    if workload == "c2":
        for band in ("bandpass", "lowpass", "highpass"):
            m, _ = circuits.build_c2(hv, band, nF=8)
            jobs.append((m, slab(2e9, 3e9, 100_001), None, None))
        desc = "configs[1]: demo2 Cauer band/low/high-pass set, 2-port, 100001-point sweep each (per GPU)"
    elif workload == "c1":
        m, _ = circuits.build_c1(hv, nF=8)
        jobs.append((m, slab(1.8e9, 2.2e9, 2001), None, None))
        desc = "configs[0]: README demo, 2001 points (per GPU)"
    elif workload == "c3":
        m, _ = circuits.build_c3(hv, nF=8)
        npts = int(os.environ.get("HVB_C3_POINTS", "1000000"))
        jobs.append((m, slab(1e9, 6e9, npts), None, None))
        desc = "configs[2]: synthetic 8-port Touchstone block in LC matching network, %d points (per GPU)" % npts
    elif workload == "c4":
        m, f, mc = circuits.build_c4(hv)
        nS = 10_000 // 8                      # the config shards 10k samples over 8 GPUs
        rs = np.random.RandomState(7 + rank)
        vals = np.stack([rs.normal(r._mean, r._std, size=nS) for r in mc._random_numbers], axis=1)
        jobs.append((m, np.ascontiguousarray(f), mc._random_numbers, vals))
        desc = "configs[3]: Monte-Carlo Cauer band-pass, 1250 samples x 2001 freqs per GPU"
    elif workload == "c5":
        m, _ = circuits.build_c5(hv, nF=8)
        jobs.append((m, slab(0.5e9, 6e9, int(os.environ.get("HVB_C5_POINTS", "1000000"))), None, None))
        desc = "configs[4]: ladder/TL network n=207, 8 ports, %s points per GPU" % os.environ.get("HVB_C5_POINTS", "1000000")
    else:
        raise SystemExit("unknown workload " + workload)
    return jobs, desc


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
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
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
    sel = np.ascontiguousarray(f[:: max(1, len(f) // n)][:n])
    return sel


def time_cpu(jobs, steps: int, warmup: int, per_step_seconds: float = 4.0):
    """Bounded sample of the same workload through the oracle port; returns (solves/s, info)."""
    from oracle import mna_oracle_numba as onb
    recs = [(j[0].records(), j[0].n_nodes, len(j[0].sources)) for j in jobs]
    # size the sample: calibrate on tiny runs (the first also triggers the JIT once)
    for (r, N, M), j in zip(recs, jobs):
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("HVB_WORKLOAD", "c2"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import heavi_b200 as hv

    if args.impl == "reference":
        if rank != 0:
            return
        jobs, desc = make_jobs(hv, args.workload, 0, max(world, args.gpus, 1))
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        val, ms, info = time_cpu(jobs, min(steps, 3), min(warmup, 1))
        info["value"] = val
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": min(steps, 3),
            "warmup": min(warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": desc, "cpu_path": "reference algorithm (zgelsd per port per frequency) on host cores"},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
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

    jobs, desc = make_jobs(hv, args.workload, rank, world)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    pts_step = sum(job_points(j) for j in jobs)

    # ---- device-resident state per job ------------------------------------------------------
    state = []
    for (m, f, drivers, values) in jobs:
        cn = m._compiled()
        consts, prefs, tables, samples = cn.resolve(f, drivers, values)
        P = cn.plan.P
        st = dict(cn=cn, consts=consts, prefs=prefs,
                  d_f=torch.from_numpy(f).to(dev),
                  d_tab=torch.from_numpy(tables).to(dev) if tables.shape[0] else None,
                  d_smp=torch.from_numpy(samples).to(dev),
                  d_S=torch.empty((samples.shape[0], P, P, len(f)), dtype=torch.complex128, device=dev),
                  d_status=torch.zeros(1, dtype=torch.int32, device=dev), nT=tables.shape[0])
        state.append(st)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def launch_all():
        for st in state:
            st["cn"].handle.run_device(st["d_f"], st["consts"], st["prefs"], st["d_tab"], st["d_smp"], st["d_S"], st["d_status"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: kernels only ---------------------------------------------------------------
    for _ in range(warmup):
        launch_all()
    torch.cuda.synchronize()
    launches0 = sum(st["cn"].handle.kernel_info()["launches"] for st in state)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall0 = time.time()
    for k in range(steps):
        flush.zero_()                                                   # evict L2 (untimed)
        ev[k][0].record()
        launch_all()
        ev[k][1].record()
    barrier()
    t_wall1 = time.time()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = sum(st["cn"].handle.kernel_info()["launches"] for st in state) - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    status = int(sum(int(st["d_status"].item()) for st in state))

    # ---- e2e: public API, host buffers -------------------------------------------------------
    def e2e_step():
        acc = 0.0
        for (m, f, drivers, values) in jobs:
            if values is None:
                S = m.run_sp(f)
                acc += float(S._S[0, 0, -1].real)
            else:
                S4 = m.run_sp_batch(f, drivers, values)
                acc += float(S4[-1, 0, 0, -1].real)
        return acc

    for _ in range(3):
        e2e_step()
    e2e_steps = max(3, min(steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = sum(j[1].nbytes + (0 if j[3] is None else j[3].nbytes) for j in jobs)
    d2h = sum(job_points(j) * st["cn"].plan.P ** 2 * 16 for j, st in zip(jobs, state))

    # ---- reduce over ranks -----------------------------------------------------------------
    tot = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max = float(tot[0]), float(tot[1])
    value = world * pts_step * steps / (dev_ms_max * 1e-3)
    e2e_value = world * pts_step * e2e_steps / e2e_s_max

    if rank == 0:
        # dominant kernel = the job with the most algorithmic flops; time it alone
        flops = [st["cn"].plan.flops_per_solve() * job_points(j) for j, st in zip(jobs, state)]
        top = int(np.argmax(flops))
        st = state[top]
        h = st["cn"].handle
        single = []
        for _ in range(max(5, steps)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h.run_device(st["d_f"], st["consts"], st["prefs"], st["d_tab"], st["d_smp"], st["d_S"], st["d_status"])
            b.record()
            torch.cuda.synchronize()
            single.append(a.elapsed_time(b))
        k_ms = float(np.mean(single))
        peak = engine.measure_dfma_peak(local_rank)
        achieved = flops[top] / (k_ms * 1e-3) / 1e12
        ki = h.kernel_info()
        plan = st["cn"].plan
        bytes_alg = (8 + 16 * plan.P ** 2) * job_points(jobs[top])
        mp = measured_peaks()
        hbm_peak = mp["hbm_gbs"] if mp else 6650.0
        roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload),
                "kernel": ("hvb_warp%s_kernel<%d,%d>" % ("2" if ki["threads"] and st["cn"].rows_per_lane == 2 else "", ki["group"], ki["pmax"]))
                          if ki["kernel_class"] == 0 else "hvb_block_kernel<%d>" % ki["group"],
                "launch_ms": k_ms, "n_factored": plan.n, "solver_mode": plan.mode,
                "flops_per_solve": plan.flops_per_solve(), "solves_per_launch": job_points(jobs[top]),
                "peak_source": "DFMA probe measured in this run (hvb_measure_dfma_peak); MEASURED_PEAKS.json has no FP64 entry",
                "regs": ki["regs"], "ctas_per_sm": ki["ctas_per_sm"], "smem_bytes": ki["smem_bytes"]}
        roof_hbm = {"bound": "hbm", "achieved": bytes_alg / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_solve": 8 + 16 * plan.P ** 2,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if mp else "fallback 6650 GB/s"}
        if ki["kernel_class"] == 2:
            # banded kernel: thread per point, pivot-row records stream through HBM (write once, read once);
            # its flop count is O(n B (2B+P)), so the bound is memory, and the dense-LU flop figure is only
            # reported as an equivalent
            per_solve = 8 + 16 * (2 * ki["stream_words"] + plan.P ** 2)
            gbs = per_solve * job_points(jobs[top]) / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": ncu_traffic(args.workload), "kernel": "hvb_bandf_kernel<%d,false>" % ki["group"],
                    "launch_ms": k_ms, "n_factored": plan.n, "solver_mode": plan.mode, "half_bandwidth": plan.band,
                    "bytes_per_solve": per_solve, "solves_per_launch": job_points(jobs[top]),
                    "dense_lu_equivalent_tflops": achieved,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if mp else "fallback 6650 GB/s",
                    "regs": ki["regs"], "ctas_per_sm": ki["ctas_per_sm"], "smem_bytes": ki["smem_bytes"]}
        cpu = None
        if not args.no_cpu:
            v, ms, info = time_cpu(jobs, 2, 1)
            info["value"], info["unit"] = v, UNIT
            cpu = info
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": desc, "solves_per_step_per_gpu": pts_step, "l2": "flushed between timed steps (256 MiB memset)",
                       "sharding": "frequency block per rank, no collective" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "Model.run_sp(f) -> Sparameters (pinned host result)"},
            "gpu_launches": int(launches), "status_word": status,
            "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
