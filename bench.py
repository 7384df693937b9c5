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
                           "collective; one all_gather of final states at the end (not in the timed region)" % a.gpus}


# ------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (N = 1 only; reported in the same JSON line under "configs")
# ------------------------------------------------------------------------------------------------------
def run_other_configs(device, peak_tflops, hbm_gbs):
    """cfg 2-5 of BASELINE.json at full size, inputs resident in HBM, best of 3 kernel launches each (CUDA events
    on the launching stream).  flops/step follow DESIGN.md section 3."""
    import cases
    from assimulo_b200.gpu import _core
    out = []

    def run(name, model, solver, y0, p, sw0, tf, flops, t_grid=None, reverse=False, reduce_rows=False, **opts):
        h = _core.Handle(device)
        h.set_model_builtin(model)
        h.set_ensemble(y0, p, sw0, 0.0)
        d = {"solver": _core.SOLVER_IDS[solver]}
        d.update(opts)
        h.set_opts(d)
        h.set_row_reduction(reduce_rows)
        h.set_output(t_grid, 16)
        h.set_fetch_reverse(reverse)
        ms = []
        for _ in range(4):
            h.reset()
            h.simulate(tf)
            ms.append(h.last_kernel_ms())
        ms = min(ms[1:])
        sums, nbad = h.get_summary()
        n = y0.shape[1]
        rec = {"name": name, "members": int(y0.shape[0]), "model": model, "solver": solver, "kernel_ms": ms,
               "accepted_steps": int(sums["nsteps"]), "value": sums["nsteps"] / (ms * 1e-3), "unit": UNIT,
               "not_complete": int(nbad), "state_events": int(sums["nstateevents"]),
               "flops_per_accepted_step": flops,
               "fp64_frac": sums["nsteps"] * flops / (ms * 1e-3) / 1e12 / peak_tflops}
        if reduce_rows:   # the rows never reach HBM: [rows, n] statistics come back instead of [rows, N, n]
            rs = h.get_row_reduction()
            rec["rows_reduced_in_kernel"] = int(len(t_grid))
            rec["members_in_last_row"] = int(rs["count"][-1])
        elif t_grid is not None:
            nbytes = float(len(t_grid)) * n * 8 * y0.shape[0]
            rec["dense_bytes"] = nbytes
            rec["hbm_write_frac"] = nbytes / (ms * 1e-3) / 1e9 / hbm_gbs
        h.close()
        out.append(rec)

    N = 1 << 20
    y0, mu = cases.vdp_sweep(N)
    run("cfg2 RungeKutta4 h=1e-3 (strict arithmetic, bit-exact vs reference)", "vdp", "RungeKutta4", y0, mu, None, 2.0,
        15 * 2 + 4 * 5 + 7, h=1e-3, reverse=True)
    run("cfg2 RungeKutta34 rtol=atol=1e-6", "vdp", "RungeKutta34", y0, mu, None, 2.0, 30 * 2 + 5 * 5 + 12,
        rtol=1e-6, atol=1e-6, inith=0.01, reverse=True)
    y0, p, sw0 = cases.ball_sweep(N)
    run("cfg3 Dopri5 bouncing ball, state events on device", "ball", "Dopri5", y0, p, sw0, 3.0, 81 * 2 + 20,
        rtol=1e-6, atol=1e-6)
    y0, p = cases.robertson_sweep(1 << 18)
    run("cfg4 Radau5ODE Robertson, analytic Jacobian", "robertson", "Radau5ODE", y0, p, None, 40.0, 1040,
        rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    # the same ensemble with RodasODE: ~420 flops per accepted step by the same counting convention
    # (6 x (RHS + back-substitution), LU, Jacobian, df/dt difference, stage combinations, CONTRO coefficients)
    run("cfg4' RodasODE Robertson, analytic Jacobian", "robertson", "RodasODE", y0, p, None, 40.0, 420,
        rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    y0, p = cases.gyro_sweep(1 << 22)
    run("cfg5 Dopri5 gyro (12 states), dense output every 1e-3", "gyro", "Dopri5", y0, p, None, 0.1,
        81 * 12 + 6 * 39 + 20, t_grid=np.linspace(0.0, 0.1, 101)[1:], rtol=1e-8, atol=1e-8)
    run("cfg5' the same with the 100 rows reduced over the ensemble in the kernel (mean/var/min/max, no dense buffer)",
        "gyro", "Dopri5", y0, p, None, 0.1, 81 * 12 + 6 * 39 + 20, t_grid=np.linspace(0.0, 0.1, 101)[1:],
        reduce_rows=True, rtol=1e-8, atol=1e-8)
    return out


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

    # ---- e2e: the public solver API with HOST buffers; H2D + kernel + D2H inside every step ----
    prob = Batched_Explicit_Problem(model="vdp", y0=y0_h, p0=p_h)
    sim = Dopri5(prob, device=local_rank)
    sim.rtol, sim.atol = RTOL, ATOL
    sim.order = "two-ended"  # expensive and cheap 32-member batches alternate: even PCIe demand in zero-copy mode
    sim.options["reuse_buffers"] = True
    if a.host_chunks:
        sim.options["host_chunks"] = a.host_chunks

    def e2e_step():
        sim.reset()                    # back to (t0, y0): the next simulate() uploads y0 and p again
        sim.simulate(TF)               # chunk-pipelined H2D inputs / kernels / D2H final states + summary
        assert sim.n_not_complete == 0
        return sim.statistics["nsteps"]

    for _ in range(max(1, a.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    for _ in range(a.steps):
        e2e_steps += e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t0
    # the same sweep given the way it is specified -- ONE initial state [2, -0.6] for all members plus mu[N]: the row
    # travels in the kernel's argument block (aens_set_shared_y0), only mu crosses PCIe (extra key e2e_shared_y0)
    prob2 = Batched_Explicit_Problem(model="vdp", y0=[2.0, -0.6], p0=p_h)
    sim2 = Dopri5(prob2, device=local_rank)
    sim2.rtol, sim2.atol = RTOL, ATOL
    sim2.order = "two-ended"
    sim2.options["reuse_buffers"] = True

    def e2e2_step():
        sim2.reset()
        sim2.simulate(TF)
        assert sim2.n_not_complete == 0
        return sim2.statistics["nsteps"]

    for _ in range(max(1, a.warmup // 2)):
        e2e2_step()
    barrier()
    t0 = time.perf_counter()
    e2e2_steps = 0
    for _ in range(a.steps):
        e2e2_steps += e2e2_step()
    barrier()
    e2e2_wall = time.perf_counter() - t0
    h2d = N * 3 * 8                      # y0[N,2], p[N,1]
    d2h = N * 2 * 8 + 72                 # y[N,2] + statistic sums and the not-complete count (t, status: on demand)

    # ---- reduce over ranks: MAX time, SUM work ----
    vals = torch.tensor([total_ms, e2e_wall, float(np.mean(kernel_ms)), e2e2_wall], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(steps_rank), float(e2e_steps), float(e2e2_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    total_ms, e2e_wall, kern_ms, e2e2_wall = [float(v) for v in vals.cpu()]
    steps_all, e2e_steps_all, e2e2_steps_all = [float(v) for v in work.cpu()]

    # final all-gather of the results (the path's only collective; outside the timed region)
    if world > 1:
        from assimulo_b200.gpu import parallel
        g = parallel.gather_final(sim, n_total)
        assert g["t"].shape[0] == n_total

    if rank == 0:
        value = steps_all * a.steps / (total_ms * 1e-3)
        e2e_value = e2e_steps_all / e2e_wall
        achieved = steps_rank * FLOPS_PER_STEP / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, n_total),
            "accepted_steps_per_pass": steps_all, "kernel_ms": kern_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_wall / a.steps * 1e3,
                    "path": "Dopri5(Batched_Explicit_Problem).simulate() -> aens_simulate_host through the C ABI, "
                            "zero-copy mode: ONE launch whose warps read their members' y0/p rows from pinned host "
                            "memory over PCIe when they fetch them and write the final y rows back when they finish"},
            "e2e_shared_y0": {"value": e2e2_steps_all / e2e2_wall, "unit": UNIT, "h2d_bytes_per_step": int(N * 8 + 16),
                              "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e2_wall / a.steps * 1e3,
                              "path": "the same call with the problem given as specified: ONE initial state for all "
                                      "members (y0=[2,-0.6]) + mu[N]; the row travels in the kernel's argument block "
                                      "(aens_set_shared_y0), only mu crosses PCIe.  Extra information: the `e2e` key "
                                      "above keeps the general per-member y0[N,2] upload"},
            "gpu_launches": int(launches),
            "cpu_affinity": ("rank 0 bound to %d CPUs next to its GPU (NVML)" % len(numa)) if numa else "unbound",
            "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops,
                         "traffic": (ncu_traffic() or (None, None))[0] if a.members == (1 << 23) else None,
                         "traffic_source": "dram bytes read+written by one launch at 2^23 members, ncu --set full, "
                                           "profiles/%s" % (ncu_traffic() or (None, "absent"))[1],
                         "algorithmic_bytes": a.members * ((2 * 2 + 1 + 1 + 1) * 8 + 4 + 9 * 4),
                         "peak_source": "measured live on this GPU: 1 s of 8-chain DFMA in the full-rate operand "
                                        "form (aens_measure_fp64_peak, profiles/r1_fp64_pipe.md); "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "kernel": "aens_dopri5_d0_vdp_x_f", "flops_per_accepted_step": FLOPS_PER_STEP,
                         "implied_sm_mhz": peak_mhz},
            "clocks": clk,
        }
        if a.other_configs and world == 1:
            hbm = 6457.4
            try:
                hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            del sim, prob
            h.close()
            line["configs"] = run_other_configs(local_rank, peak_tflops, hbm)
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
        dist.destroy_process_group()
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
                    help="skip the BASELINE.json configs 2-5 (they run only at --gpus 1)")
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
    return main_gpu(a, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
