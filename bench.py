#!/usr/bin/env python
"""Headline benchmark: accepted trajectory-steps/s of the batched Dopri5 on a Van der Pol sweep
(BASELINE.json north_star: ">=1M-member Dopri5 Van der Pol ensemble ... accepted trajectory-steps/sec (whole box)
at 1/2/4/8 B200; % fp64 FMA peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

A "step" is one pass of the hot path over the whole ensemble: every member integrated from t=0 to t=2 with
Dopri5(rtol=atol=1e-6) -- reset (device to device) + one kernel launch.  Weak scaling: each rank owns
MEMBERS_PER_GPU members; the global mu law 10^(-1+3i/(N-1)) is dealt to the ranks in 32-member chunks round-robin so
that every GPU sees the same cost histogram.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

METRIC = "accepted trajectory-steps/sec (Dopri5 Van der Pol sweep, whole box)"
UNIT = "accepted_steps/s"
TF = 2.0
RTOL = ATOL = 1e-6
# algorithmic fp64 flops per ACCEPTED Dopri5 step for VdP (n=2, F_rhs=5), SURVEY.md 8(d) convention
# (add/sub/mul/div/sqrt/pow = 1, FMA = 2, abs/max/compare = 0; rejected attempts earn nothing):
#   stages 46n + error vector 12n + error norm 5n + 6 F_rhs + 20 controller   = 63n + 6F + 20 = 176
# The reference additionally always builds the dense-output coefficients (12n + 6n = 36 flops -> 212); the
# final-state-only kernel (n_out = 0, no events) does not compute them, so they are not credited.
FLOPS_PER_STEP = 63 * 2 + 6 * 5 + 20


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the headline kernel, from the committed ncu
    capture of this same command (profiles/r*_ncu.json, written by tools/profile_summary.py); None if absent."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu.json"))):
        try:
            m = json.load(open(f)).get("dopri5_vdp")
            if m:
                best = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], os.path.basename(f))
        except Exception:
            pass
    return best


def ncu_fp64_counters():
    """The counters BASELINE.json's metric names, for ONE launch of the headline kernel at 2^23 members, from the committed
    ncu capture (profiles/r*_ncu.json): thread-level DFMA / DMUL / DADD and the flops they amount to; None if absent."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu.json"))):
        try:
            m = json.load(open(f)).get("dopri5_vdp")
            k = "sm__sass_thread_inst_executed_op_%s_pred_on.sum"
            if m and (k % "dfma") in m:
                best = {"dfma": m[k % "dfma"], "dmul": m[k % "dmul"], "dadd": m[k % "dadd"],
                        "executed_flops": 2 * m[k % "dfma"] + m[k % "dmul"] + m[k % "dadd"],
                        "achieved_occupancy_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                        "fp64_pipe_active_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                        "source": "profiles/" + os.path.basename(f)}
        except Exception:
            pass
    return best


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and the page-locked buffers it is about to allocate) to the CPUs next to its GPU: the e2e leg
    streams the host arrays over that GPU's PCIe link, and with 8 ranks on a two-socket host unbound processes send
    half of that traffic across the socket interconnect.  Best effort (NVML); returns the CPU list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            bind_to_gpu_numa_node.original = os.sched_getaffinity(0)
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def vdp_members(n_total, rank, world, per_rank):
    """Members of this rank: 32-member chunks of the global sweep dealt round-robin."""
    chunks = np.arange(rank, n_total // 32, world)[: per_rank // 32]
    idx = (chunks[:, None] * 32 + np.arange(32)[None, :]).reshape(-1)
    mu = 10.0 ** (-1.0 + 3.0 * idx / max(n_total - 1, 1))
    return idx, mu


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t_begin - 0.05 <= ts <= t_end + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take everything we saw
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU legs: the reference's Python loop over its own solver classes (oracle/_ref), all host cores
# ------------------------------------------------------------------------------------------------------
def _ref_worker(mus):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import io
    import contextlib
    with contextlib.redirect_stderr(io.StringIO()):
        import ref_models
        from assimulo.solvers import Dopri5
    nsteps = 0
    for mu in mus:
        sim = Dopri5(ref_models.vdp(float(mu)))
        sim.verbosity = 50
        sim.rtol, sim.atol = RTOL, ATOL
        sim.simulate(TF)
        nsteps += int(sim.statistics["nsteps"])
    return nsteps


def _ref_warm(_):
    return _ref_worker([1.0])


def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "assimulo", "solvers", "runge_kutta.py")) and \
        os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdopri5_c.so"))


def run_cpu_reference(n_total, sample, cores=None):
    """Times the reference's per-member Python loop (Explicit_Problem -> Dopri5 -> simulate) on `sample` members
    spread evenly over the global sweep, in a multiprocessing.Pool over all host cores.  Returns
    (steps_per_s, cores, total_steps, wall, kind)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    idx = np.linspace(0, n_total - 1, sample).astype(np.int64)
    mu = 10.0 ** (-1.0 + 3.0 * idx / max(n_total - 1, 1))
    if reference_available():
        parts = [mu[i::cores * 4] for i in range(cores * 4)]
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_ref_warm, range(cores))  # imports + first-call costs outside the timer
            t0 = time.perf_counter()
            steps = sum(pool.map(_ref_worker, parts))
            wall = time.perf_counter() - t0
        return steps / wall, cores, steps, wall, "reference"
    # the reference build does not exist on this box: fall back to the C oracle port (OpenMP, all cores)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ens_oracle as O
    y0 = np.tile(np.array([2.0, -0.6]), (len(mu), 1))
    O.simulate("vdp", "Dopri5", y0[:8], mu[:8].reshape(-1, 1), tf=TF, rtol=RTOL, atol=ATOL)
    t0 = time.perf_counter()
    out = O.simulate("vdp", "Dopri5", y0, mu.reshape(-1, 1), tf=TF, rtol=RTOL, atol=ATOL)
    wall = time.perf_counter() - t0
    steps = int(out["stats"]["nsteps"].sum())
    return steps / wall, cores, steps, wall, "port"


def main_reference(a, rank, world):
    if rank != 0:
        return 0
    n_total = a.members * a.gpus
    per_step = []
    info = None
    for it in range(a.warmup + a.steps):
        info = run_cpu_reference(n_total, a.cpu_sample)
        if it >= a.warmup:
            per_step.append(info)
    v = float(np.mean([p[0] for p in per_step]))
    wall = float(np.mean([p[3] for p in per_step]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, n_total),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info[1], "kind": info[4],
                         "sample": "%d members evenly spaced over the %d-member sweep, Python loop "
                                   "Explicit_Problem->Dopri5->simulate in a %d-process pool (%d accepted steps per pass)"
                                   % (a.cpu_sample, n_total, info[1], info[2])},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(a, n_total):
    return {"workload": "Dopri5 Van der Pol sweep: y0=[2,-0.6], t in [0,2], rtol=atol=1e-6, "
                        "mu_i=10^(-1+3i/(N-1)), N=%d members (%d per GPU)" % (n_total, a.members),
            "members_total": n_total, "members_per_gpu": a.members, "solver": "Dopri5", "model": "vdp",
            "arith": "fast", "l2": "per-GPU inputs+outputs %.0f MB > 126 MB L2 (no flush needed)"
                                   % (a.members * (3 + 3 + 1.5 + 4) * 8 / 1e6),
            "fetch": "descending member index (cost grows with mu), 32-member batches per warp",
            "parallelism": "members sharded over %d GPU(s), 32-member chunks dealt round-robin, no data-path "
                           "collective; ONE ncclAllGather of {t, y, status} at the end of every e2e step" % a.gpus}


# ------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations: in the headline line under "configs" (N = 1), or one at a time at any
# number of GPUs with --config NAME (weak scaling, members dealt like the headline's)
# ------------------------------------------------------------------------------------------------------
def deal(n_total, rank, world, per_rank):
    """global member indices of this rank: 32-member chunks of the global sweep dealt round-robin"""
    chunks = np.arange(rank, n_total // 32, world)[: per_rank // 32]
    return (chunks[:, None] * 32 + np.arange(32)[None, :]).reshape(-1)


def config_table():
    """name -> dict(members per GPU, model, solver, builder(n_total, idx) -> (y0, p, sw0), tf, flops per accepted
    step, options).  Flop counts follow the SURVEY 8(d) convention; Radau5ODE's is the
    instrumented count of the CPU oracle on this very ensemble (oracle/ens_oracle.c, AENS_ORACLE_FLOPS)."""
    import cases
    grid100 = np.linspace(0.0, 0.1, 101)[1:]
    rob_opts = dict(rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    gyro_flops = 81 * 12 + 6 * 39 + 20
    return {
        "cfg2_rk4": dict(title="cfg2 RungeKutta4 h=1e-3 (strict arithmetic, bit-exact vs reference)", members=1 << 20,
                         model="vdp", solver="RungeKutta4", build=cases.vdp_sweep, tf=2.0, flops=15 * 2 + 4 * 5 + 7,
                         opts=dict(h=1e-3), reverse=True),
        "cfg2_rk34": dict(title="cfg2 RungeKutta34 rtol=atol=1e-6", members=1 << 20, model="vdp", solver="RungeKutta34",
                          build=cases.vdp_sweep, tf=2.0, flops=30 * 2 + 5 * 5 + 12,
                          opts=dict(rtol=1e-6, atol=1e-6, inith=0.01), reverse=True),
        # per accepted step 81n + 6F + 20 = 182 (n = 2, F = 0), plus the event locator's probes ("182 (+ event
        # probes)"): every Illinois iteration of src/ode_event_locator.c:64-142 is one dense-output evaluation (8n + 3 =
        # 19) and 14 flops of its own (bracket width, alpha, the secant formula, the two nudge tests).  Iterations =
        # nstatefcns - one g per accepted step - two per segment (segment start + DOPRI5's first SOLOUT call).
        "cfg3": dict(title="cfg3 Dopri5 bouncing ball, state events on device", members=1 << 20, model="ball",
                     solver="Dopri5", build=cases.ball_sweep, tf=3.0, flops=81 * 2 + 20, opts=dict(rtol=1e-6, atol=1e-6),
                     extra_flops=lambda sums, members: 33.0 * (sums["nstatefcns"] - sums["nsteps"]
                                                               - 2 * (members + sums["nstateevents"]))),
        "cfg4": dict(title="cfg4 Radau5ODE Robertson, analytic Jacobian", members=1 << 18, model="robertson",
                     solver="Radau5ODE", build=cases.robertson_sweep, tf=40.0, flops=RADAU_ROBERTSON_FLOPS, opts=rob_opts),
        "cfg4_rodas": dict(title="cfg4' RodasODE Robertson, analytic Jacobian", members=1 << 18, model="robertson",
                           solver="RodasODE", build=cases.robertson_sweep, tf=40.0, flops=RODAS_ROBERTSON_FLOPS, opts=rob_opts),
        "cfg5": dict(title="cfg5 Dopri5 gyro (12 states), dense output every 1e-3", members=1 << 22, model="gyro",
                     solver="Dopri5", build=cases.gyro_sweep, tf=0.1, flops=gyro_flops, t_grid=grid100,
                     opts=dict(rtol=1e-8, atol=1e-8)),
        "cfg5_reduced": dict(title="cfg5' the same with the 100 rows reduced over the ensemble in the kernel "
                                   "(mean/var/min/max, no dense buffer)", members=1 << 22, model="gyro", solver="Dopri5",
                             build=cases.gyro_sweep, tf=0.1, flops=gyro_flops, t_grid=grid100, reduce_rows=True,
                             opts=dict(rtol=1e-8, atol=1e-8)),
    }


# Radau5ODE on the cfg4 Robertson ensemble: fp64 flops per ACCEPTED step counted by the instrumented CPU oracle over
# every 64th member (tools/radau_flops.py; SURVEY 8(d) asked for a count instead of the 1040 estimate of round 1)
RADAU_ROBERTSON_FLOPS, RODAS_ROBERTSON_FLOPS = 1040, 420      # (the estimates of round 1, if the counts are absent)
try:
    _fl = json.load(open(os.path.join(ROOT, "profiles", "r2_radau_flops.json")))
    RADAU_ROBERTSON_FLOPS = float(_fl["flops_per_accepted_step"])
    RODAS_ROBERTSON_FLOPS = float(_fl["rodas_flops_per_accepted_step"])
except Exception:
    pass


def run_config(device, cfg, rank, world, peak_tflops, hbm_gbs, steps=3, warmup=1):
    """One configuration on this rank's GPU, inputs resident in HBM: `warmup` untimed passes, then `steps` passes of
    (reset + kernel) timed with CUDA events on the launching stream.  Returns the per-rank record."""
    from assimulo_b200.gpu import _core
    per = cfg["members"]
    n_total = per * world
    idx = deal(n_total, rank, world, per) if world > 1 else None
    arrs = cfg["build"](n_total, idx)
    y0, p = arrs[0], arrs[1]
    sw0 = arrs[2] if len(arrs) > 2 else None
    h = _core.Handle(device)
    h.set_model_builtin(cfg["model"])
    h.set_ensemble(y0, p, sw0, 0.0)
    d = {"solver": _core.SOLVER_IDS[cfg["solver"]]}
    d.update(cfg["opts"])
    h.set_opts(d)
    h.set_row_reduction(bool(cfg.get("reduce_rows")))
    h.set_output(cfg.get("t_grid"), 16)
    h.set_fetch_reverse(bool(cfg.get("reverse")))
    for _ in range(warmup):
        h.reset()
        h.simulate(cfg["tf"])
    l0 = h.launch_count()
    for _ in range(steps):
        h.reset()
        h.simulate_async(cfg["tf"])
    h.sync()
    launches = h.launch_count() - l0
    ms_all = h.kernel_ms_history(steps)
    ms = float(np.mean(ms_all))
    sums, nbad = h.get_summary()
    n = y0.shape[1]
    flops = cfg["flops"]
    if cfg.get("extra_flops"):  # work the solver does besides its steps (event location), spread over the accepted steps
        flops = flops + cfg["extra_flops"](sums, int(y0.shape[0])) / max(int(sums["nsteps"]), 1)
    rec = {"name": cfg["title"], "members": int(y0.shape[0]), "model": cfg["model"], "solver": cfg["solver"],
           "kernel_ms": ms, "stat": "mean of %d launches after %d warm-up" % (steps, warmup),
           "kernel_ms_min": float(min(ms_all)), "accepted_steps": int(sums["nsteps"]),
           "value": sums["nsteps"] / (ms * 1e-3), "unit": UNIT,
           "not_complete": int(nbad), "state_events": int(sums["nstateevents"]), "state_event_probes": int(sums["nstatefcns"]),
           "flops_per_accepted_step": flops, "gpu_launches": int(launches),
           "fp64_frac": sums["nsteps"] * flops / (ms * 1e-3) / 1e12 / peak_tflops}
    if cfg.get("reduce_rows"):   # the rows never reach HBM: [rows, n] statistics come back instead of [rows, N, n]
        rs = h.get_row_reduction()
        rec["rows_reduced_in_kernel"] = int(len(cfg["t_grid"]))
        rec["members_in_last_row"] = int(rs["count"][-1])
    elif cfg.get("t_grid") is not None:
        nbytes = float(len(cfg["t_grid"])) * n * 8 * y0.shape[0]
        rec["dense_bytes"] = nbytes
        rec["hbm_write_frac"] = nbytes / (ms * 1e-3) / 1e9 / hbm_gbs
    h.close()
    return rec


def run_other_configs(device, peak_tflops, hbm_gbs):
    """cfg 2-5 of BASELINE.json at full size on one GPU (the `configs` key of the headline line)."""
    return [run_config(device, cfg, 0, 1, peak_tflops, hbm_gbs) for cfg in config_table().values()]


def main_config(a, rank, world, local_rank):
    """--config NAME: one of the other BASELINE configurations as the timed workload, at any number of GPUs (weak
    scaling: every GPU integrates the configuration's member count, dealt from the world-sized sweep)."""
    import torch
    import torch.distributed as dist
    from assimulo_b200.gpu import _core
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        _watchdog(600.0)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak_tflops, _ = _core.measure_fp64_peak(local_rank, 0.5)
    hbm = measured_hbm()
    names = list(config_table()) if a.config == "all" else [a.config]
    for name in names:      # "all": one JSON line per configuration, one process group for all of them
        cfg = config_table()[name]
        _config_line(a, cfg, rank, world, local_rank, peak_tflops, hbm, json_fd)
    if world > 1:
        _leave(0)
    return 0


def _config_line(a, cfg, rank, world, local_rank, peak_tflops, hbm, json_fd):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rec = run_config(local_rank, cfg, rank, world, peak_tflops, hbm, steps=a.steps, warmup=max(a.warmup, 1))
    vals = torch.tensor([rec["kernel_ms"]], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(rec["accepted_steps"]), float(rec["not_complete"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms = float(vals[0])
    steps_all, nbad = float(work[0]), int(work[1])
    if rank == 0:
        hbm_bound = "hbm_write_frac" in rec
        line = {"metric": "accepted trajectory-steps/sec (%s, whole box)" % cfg["title"], "value": steps_all / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": cfg["title"], "members_per_gpu": cfg["members"], "members_total": cfg["members"] * world,
                           "solver": cfg["solver"], "model": cfg["model"],
                           "timing": "kernel launches only (CUDA events on the launching stream), max over ranks of the "
                                     "per-rank mean; inputs resident in HBM"},
                "not_complete": nbad, "gpu_launches": rec["gpu_launches"], "rank0": rec,
                "roofline": ({"bound": "hbm", "achieved": rec["dense_bytes"] / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": rec["dense_bytes"] / (ms * 1e-3) / 1e9 / hbm, "traffic": None} if hbm_bound else
                             {"bound": "fp64_fma", "achieved": rec["accepted_steps"] * rec["flops_per_accepted_step"] / (ms * 1e-3) / 1e12,
                              "peak": peak_tflops, "unit": "TFLOP/s",
                              "frac": rec["accepted_steps"] * rec["flops_per_accepted_step"] / (ms * 1e-3) / 1e12 / peak_tflops,
                              "traffic": None, "flops_per_accepted_step": rec["flops_per_accepted_step"]})}
        os.write(json_fd, (json.dumps(line) + "\n").encode())


def _leave(code):
    """End of a multi-rank run.  Every rank is past its last collective; nothing is left that needs the peers.  The
    process ends here, without the interpreter's teardown: destroying NCCL communicators (torch's and the library's)
    while the other ranks are at different points of their own exit has been seen to block for minutes on an 8-GPU box
    (profiles/r2_summary.md), and there is nothing to flush but the two standard streams."""
    sys.stdout.flush()
    sys.stderr.flush()
    # The other ranks wait (on the rendezvous store, not on a collective) until rank 0 has printed its line, so that no
    # peer disappears under rank 0 while it still works; whatever goes wrong here, the process ends.
    try:
        import datetime
        import torch.distributed as dist
        store = dist.distributed_c10d._get_default_store()
        if dist.get_rank() == 0:
            store.set("bench_rank0_done", b"1")
        else:
            store.wait(["bench_rank0_done"], datetime.timedelta(seconds=240))
    except BaseException:
        pass
    os._exit(code)


def _watchdog(seconds):
    """A multi-rank run that is still alive after `seconds` leaves a traceback of every thread on stderr and ends: a
    hang must cost minutes of an 8-GPU box, not the caller's whole time limit."""
    import faulthandler
    faulthandler.dump_traceback_later(max(30.0, seconds - 20.0), exit=False)
    t = threading.Timer(seconds, lambda: os._exit(1))
    t.daemon = True
    t.start()


def measured_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6457.4


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main_gpu(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from assimulo_b200.gpu import _core
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        _watchdog(600.0)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_total = a.members * world
    idx, mu = vdp_members(n_total, rank, world, a.members)
    N = len(mu)
    # host inputs in pinned memory (the e2e leg copies them to the device every step)
    y0_h = _core.pinned_empty((N, 2))
    p_h = _core.pinned_empty((N, 1))
    y0_h[:, 0] = 2.0
    y0_h[:, 1] = -0.6
    p_h[:, 0] = mu

    # measured fp64 FMA peak of THIS device (MEASURED_PEAKS.json has no fp64 entry)
    peak_tflops, peak_mhz = _core.measure_fp64_peak(local_rank, 1.0)

    h = _core.Handle(local_rank)
    h.set_model_builtin("vdp")
    opts = {"solver": _core.SOLVER_IDS["Dopri5"], "arith": 1, "rtol": RTOL, "atol": ATOL}
    h.set_ensemble(y0_h, p_h, None, 0.0)
    h.set_opts(opts)
    h.set_output(None, 0)
    # most expensive members first (mu ascending within the rank's deal => fetch from the last member down)
    h.set_schedule(32, None)
    h.set_fetch_reverse(True)

    # ---- value: inputs resident in HBM, K x (reset + kernel) ----
    for _ in range(a.warmup):
        h.reset()
        h.simulate(TF)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    launches0 = h.launch_count()
    stream = torch.cuda.ExternalStream(h.stream())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        h.reset()
        h.simulate_async(TF)
    ev1.record(stream)
    h.sync()
    barrier()
    t_end = time.time()
    total_ms = ev0.elapsed_time(ev1)
    launches = h.launch_count() - launches0
    sums = h.get_stats_sum()
    _, _, _, status = h.get_final(want_y=False, want_sw=False, want_t=False)
    assert (status == 0).all(), "some members failed"
    steps_rank = int(sums["nsteps"])
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None

    # ---- roofline: the duration of every kernel launch OF THE TIMED REGION above (each launch records its own
    # CUDA event pair on the launching stream; the library keeps the last 64 pairs) ----
    kernel_ms = h.kernel_ms_history(min(a.steps, 64))

    # ---- what the host link sustains with all ranks busy at once (the ceiling e2e is judged against) ----
    link = {}
    for mode in range(6):
        barrier()
        link[mode] = _core.measure_hostlink(local_rank, 1 << 27, mode, 0.12)
    link_t = torch.tensor([link[m] for m in range(6)], dtype=torch.float64, device="cuda")
    link_min, link_sum = link_t.clone(), link_t.clone()
    if world > 1:
        dist.all_reduce(link_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(link_sum, op=dist.ReduceOp.SUM)
    link_min, link_sum = [float(v) for v in link_min.cpu()], [float(v) for v in link_sum.cpu()]

    # ---- e2e: the public solver API with HOST buffers; H2D + kernel + D2H inside every step; at N > 1 the step ends
    # with the path's one collective, the all-gather of {t, y, status} over NCCL issued by the library ----
    from assimulo_b200.gpu import parallel
    GATHER_WHAT = _core.GATHER_T | _core.GATHER_Y | _core.GATHER_STATUS

    def make_sim(problem):
        sm = Dopri5(problem, device=local_rank)
        sm.rtol, sm.atol = RTOL, ATOL
        sm.order = "two-ended"  # expensive and cheap 32-member batches alternate: even PCIe demand in zero-copy mode
        sm.options["reuse_buffers"] = True
        if a.host_chunks:
            sm.options["host_chunks"] = a.host_chunks
        return sm

    def make_step(sm):
        def step():
            sm.reset()                    # back to (t0, y0): the next simulate() uploads the inputs again
            sm.simulate(TF)               # H2D inputs / kernel(s) / D2H final states + summary, mode = sm.host_mode
            assert sm.n_not_complete == 0                 # at N > 1 the library issued the all-gather behind the last
            return sm.statistics["nsteps"]                # kernel (gather_in_simulate) and simulate() waited for it
        return step

    def time_steps(step, k):
        barrier()
        t0 = time.perf_counter()
        n = 0
        for _ in range(k):
            n += step()
        barrier()
        return time.perf_counter() - t0, n

    def tune(sm, step):
        """pick the host mode by measurement: which side of the link the SMs or the copy engines drive better depends
        on the machine and on how many GPUs share it (profiles/r2_pcie.md)"""
        res = {}
        for mode in ("zero-copy", "zero-copy-in", "zero-copy-out", "staged"):
            sm.host_mode = mode
            step()
            w, _ = time_steps(step, 2)
            tw = torch.tensor([w], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            res[mode] = float(tw[0]) / 2 * 1e3
        best = min(res, key=res.get)
        sm.host_mode = best
        return best, res

    # the sweep as a loop over Explicit_Problem(rhs_i, y0) holds it: N identical y0 rows + mu[N].  The problem class
    # notices the shared row, so only mu crosses PCIe (aens_set_shared_y0); results: y[N, 2] every step.
    prob = Batched_Explicit_Problem(model="vdp", y0=y0_h, p0=p_h)
    assert prob.y0_row is not None
    sim = make_sim(prob)
    if world > 1:
        parallel.gather_in_simulate(sim, GATHER_WHAT)
    e2e_step = make_step(sim)
    e2e_mode, e2e_modes = tune(sim, e2e_step)
    for _ in range(max(1, a.warmup // 2)):
        e2e_step()
    e2e_wall, e2e_steps = time_steps(e2e_step, a.steps)
    gather_ms = sim._handle.last_gather_ms() if world > 1 else 0.0      # inside the step: includes waiting for the last rank
    gather_alone_ms = 0.0
    if world > 1:   # the exchange by itself, all ranks ready: what it costs when nothing hides it
        ga = []
        for _ in range(3):
            barrier()
            sim._handle.allgather_results(GATHER_WHAT)
            sim._handle.sync()
            ga.append(sim._handle.last_gather_ms())
        gather_alone_ms = float(np.median(ga))
    # the general case: every member with its own initial state (y0[N, 2] uploaded every step); extra key
    prob2 = Batched_Explicit_Problem(model="vdp", y0=y0_h, p0=p_h, detect_shared_y0=False)
    sim2 = make_sim(prob2)
    if world > 1:
        parallel.gather_in_simulate(sim2, GATHER_WHAT)
    e2e2_step = make_step(sim2)
    e2e2_mode, e2e2_modes = tune(sim2, e2e2_step)
    e2e2_wall, e2e2_steps = time_steps(e2e2_step, a.steps)
    h2d = N * 8 + 16                     # p[N,1] + the shared y0 row
    h2d2 = N * 3 * 8                     # y0[N,2], p[N,1]
    d2h = N * 2 * 8 + 96                 # y[N,2] + statistic sums and the not-complete count (t, status: on demand)

    # ---- the all-gather checked on hardware: a strided sample of the GATHERED members against a fresh single-GPU
    # integration of the same members on rank 0 (bit-identical: members do not interact) ----
    gather_check = None
    if world > 1:
        sim._handle.allgather_results(GATHER_WHAT | _core.GATHER_STATS)   # (the timed steps gather {t, y, status} only)
        g = sim._handle.get_gathered(n_total, want_stats=True)
        if rank == 0:
            S = 4096
            gi = np.linspace(0, n_total - 1, S).astype(np.int64)            # global member indices
            chunk, lane = gi // 32, gi % 32
            pos = (chunk % world) * N + (chunk // world) * 32 + lane         # where the deal put them (rank-major)
            mu_s = 10.0 ** (-1.0 + 3.0 * gi / max(n_total - 1, 1))
            hs = _core.Handle(local_rank)
            hs.set_model_builtin("vdp")
            hs.set_ensemble(np.tile([2.0, -0.6], (S, 1)), mu_s.reshape(-1, 1), None, 0.0)
            hs.set_opts(opts)
            hs.set_output(None, 0)
            hs.simulate(TF)
            ts, ys, _, sts = hs.get_final()
            ns = hs.get_stats("nsteps")
            hs.close()
            gather_check = {"members": S, "max_abs_diff_y": float(np.max(np.abs(g["y"][pos] - ys))),
                            "nsteps_equal": bool(np.array_equal(g["stats"][0][pos], ns)),
                            "status_equal": bool(np.array_equal(g["status"][pos], sts)),
                            "t_equal": bool(np.array_equal(g["t"][pos], ts)),
                            "all_complete": bool((g["status"] == 0).all()),
                            "accepted_steps_gathered": int(g["stats"][0].astype(np.int64).sum())}
            assert gather_check["max_abs_diff_y"] == 0.0 and gather_check["nsteps_equal"] and \
                gather_check["status_equal"] and gather_check["t_equal"], gather_check
        del g

    # ---- reduce over ranks: MAX time, SUM work ----
    vals = torch.tensor([total_ms, e2e_wall, float(np.mean(kernel_ms)), e2e2_wall, gather_ms, gather_alone_ms],
                        dtype=torch.float64, device="cuda")
    work = torch.tensor([float(steps_rank), float(e2e_steps), float(e2e2_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    total_ms, e2e_wall, kern_ms, e2e2_wall, gather_ms, gather_alone_ms = [float(v) for v in vals.cpu()]
    steps_all, e2e_steps_all, e2e2_steps_all = [float(v) for v in work.cpu()]

    if rank == 0:
        value = steps_all * a.steps / (total_ms * 1e-3)
        e2e_value = e2e_steps_all / e2e_wall
        achieved = steps_rank * FLOPS_PER_STEP / (kern_ms * 1e-3) / 1e12
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12

        def floor(h2d_b, d2h_b):
            """lower bound of one e2e step on this box: the kernel followed by the all-gather (the exchange needs the
            results in HBM), or the step's bytes at the best rate each direction of the host link sustained with all
            ranks busy (copy engine or SM-initiated; the slowest GPU's link, since the step ends when the last rank is
            done), whichever is longer"""
            up = max(link_min[0], link_min[3])
            down = max(link_min[1], link_min[4])
            link_ms = max(h2d_b / up, d2h_b / down) / 1e6
            return {"kernel_ms": kern_ms, "link_ms": link_ms, "gather_alone_ms": gather_alone_ms,
                    "floor_ms": max(kern_ms + gather_alone_ms, link_ms)}
        f1, f2 = floor(h2d, d2h), floor(h2d2, d2h)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, n_total),
            "accepted_steps_per_pass": steps_all, "kernel_ms": kern_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_wall / a.steps * 1e3, "host_mode": e2e_mode, "host_mode_probe_ms": e2e_modes,
                    "gather_ms": gather_ms, "gather_alone_ms": gather_alone_ms,
                    "gather_bytes_per_rank": int(N * 28 + 32) if world > 1 else 0,
                    "floor": f1, "frac_of_floor": f1["floor_ms"] / (e2e_wall / a.steps * 1e3),
                    "path": "Dopri5(Batched_Explicit_Problem(y0[N,2], mu[N])).simulate() -> aens_simulate_host through "
                            "the C ABI with pinned HOST arrays; the problem class detects that all y0 rows are equal, so "
                            "mu[N] + one row go up and y[N,2] + the summary come down every step; host mode picked by "
                            "measurement (host_mode_probe_ms)" +
                            ("; the step ends with ONE ncclAllGather of {t, y, status} issued by the library "
                             "(aens_allgather_results)" if world > 1 else "")},
            "e2e_per_member_y0": {"value": e2e2_steps_all / e2e2_wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d2),
                                  "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e2_wall / a.steps * 1e3,
                                  "host_mode": e2e2_mode, "host_mode_probe_ms": e2e2_modes, "floor": f2,
                                  "frac_of_floor": f2["floor_ms"] / (e2e2_wall / a.steps * 1e3),
                                  "path": "the same call when every member has its own initial state: y0[N,2] is "
                                          "uploaded every step as well (shared-row detection off).  Extra information"},
            "hostlink": {"what": "GB/s per GPU with all %d rank(s) transferring at once, 128 MiB pinned buffers "
                                 "(aens_measure_hostlink): min over ranks / sum over ranks" % world,
                         "modes": {_core.HOSTLINK_MODES[m]: [link_min[m], link_sum[m]] for m in range(6)}},
            "gather_check": gather_check,
            "gpu_launches": int(launches),
            "cpu_affinity": ("rank 0 bound to %d CPUs next to its GPU (NVML)" % len(numa)) if numa else "unbound",
            "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                         "traffic": (ncu_traffic() or (None, None))[0] if a.members == (1 << 23) else None,
                         "traffic_source": "dram bytes read+written by one launch at 2^23 members, ncu --set full, "
                                           "profiles/%s" % (ncu_traffic() or (None, "absent"))[1],
                         "algorithmic_bytes": a.members * ((2 * 2 + 1 + 1 + 1) * 8 + 4 + 9 * 4),
                         "algorithmic_flops_per_launch": steps_rank * FLOPS_PER_STEP,
                         "ncu_fp64_thread_inst": ncu_fp64_counters() if a.members == (1 << 23) else None,
                         "peak_source": "measured live on this GPU by this library (builder's number): 1 s of 8-chain "
                                        "DFMA in the full-rate operand form (aens_measure_fp64_peak, "
                                        "profiles/r1_fp64_pipe.md); MEASURED_PEAKS.json has no fp64 entry; "
                                        "peak_nominal = 148 SM x 64 lanes x 2 x 1.965 GHz",
                         "kernel": "aens_dopri5_d0_vdp_x_f", "flops_per_accepted_step": FLOPS_PER_STEP,
                         "implied_sm_mhz": peak_mhz},
            "clocks": clk,
        }
        if a.other_configs and world == 1:
            del sim, prob, sim2, prob2
            h.close()
            line["configs"] = run_other_configs(local_rank, peak_tflops, measured_hbm())
        if a.cpu_baseline:
            if getattr(bind_to_gpu_numa_node, "original", None):  # the CPU leg uses every host core again
                os.sched_setaffinity(0, bind_to_gpu_numa_node.original)
            v, cores, steps, wall, kind = run_cpu_reference(n_total, a.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d members evenly spaced over the sweep, %d accepted steps in %.2f s"
                                              % (a.cpu_sample, steps, wall)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        _leave(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--members", type=int, default=1 << 23, help="members per GPU (multiple of 32)")
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--host-chunks", type=int, default=0, help="chunks of the e2e host pipeline (0 = library default)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false",
                    help="skip the BASELINE.json configs 2-5 inside the headline line (N = 1 only)")
    ap.add_argument("--config", default="headline",
                    help="headline (default) or one of the other BASELINE configurations as the timed workload, at any "
                         "--gpus: cfg2_rk4 cfg2_rk34 cfg3 cfg4 cfg4_rodas cfg5 cfg5_reduced, or all (one line each)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "graft" else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        a.gpus = max(a.gpus, world)
        return main_reference(a, rank, world)
    if world != a.gpus and world == 1 and a.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        return subprocess.call(cmd)
    if a.config != "headline":
        return main_config(a, rank, world, local_rank)
    return main_gpu(a, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
