"""N>1 path on CPU: world_size=2 `gloo` process group.  Each rank takes its contiguous block of the ensemble
(parallel.shard_problem), "integrates" it -- here with the C oracle standing in for the device, which is exactly
what the oracle is for -- and the single end-of-run all_gather (parallel.gather_final) must reassemble the
whole-ensemble result on every rank."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

WORLD = 2
SLOTS = 16
N = 37  # deliberately not divisible by WORLD: the last block is padded and trimmed after the gather


class _OracleBackedSolver(object):
    """Quacks like a Batched_Explicit_ODE after simulate() for gather_final's host (gloo) path."""

    def __init__(self, problem):
        import ens_oracle as O
        if problem.model == "ball":
            out = O.simulate("ball", "Dopri5", problem.y0, problem.p0, problem.sw0, tf=3.0, max_ev=SLOTS,
                             rtol=1e-6, atol=1e-6)
            self.event_log = (out["stats"]["nstateevents"], out["ev_t"], out["ev_info"])
        else:
            out = O.simulate("vdp", "Dopri5", problem.y0, problem.p0, tf=2.0, rtol=1e-6, atol=1e-6)
        self.options = {"event_slots": SLOTS}
        self.problem, self.N = problem, problem.N
        self.t_members, self.y, self.status = out["t"], out["y"], out["status"]
        self._stats = out["stats"]
        self._handle = None
        outer = self

        class _S(dict):
            def per_member(self, k):
                return outer._stats[k]
        self.statistics = _S()


GRID = np.linspace(0.0, 2.0, 9)[1:]


def _oracle_rows(y0, mu):
    """dense-output rows [rows, members, n] of a block, from the oracle"""
    import ens_oracle as O
    return np.transpose(O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, t_grid=GRID, rtol=1e-6, atol=1e-6)["dense"], (1, 0, 2))


class _RowStats(object):
    """what a solver with reduce_rows = True holds after simulate(): the statistics of ALL members of the shard
    problem it was built on (copies appended by padding included -- which is why padded shards are refused)"""

    def __init__(self, problem):
        rows = _oracle_rows(problem.y0, problem.p0)
        self.problem, self.N = problem, problem.N
        self.row_statistics = {"mean": rows.mean(axis=1), "var": rows.var(axis=1), "min": rows.min(axis=1),
                               "max": rows.max(axis=1), "count": np.full(rows.shape[0], rows.shape[1]), "t": GRID}


def _worker(rank, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import cases
    from assimulo_b200.gpu import Batched_Explicit_Problem, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    y0, mu = cases.vdp_sweep(N)
    full = Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu)
    shard = parallel.shard_problem(full, rank, WORLD)
    lo, hi, B = parallel.shard_bounds(N, rank, WORLD)
    assert shard.N == B and np.array_equal(shard.p0[: hi - lo, 0], mu[lo:hi, 0])
    g = parallel.gather_final(_OracleBackedSolver(shard), N)
    # the same exchange for a model with state events: event counts / times / info words travel too
    y0b, pb, swb = cases.ball_sweep(N)
    ball = parallel.shard_problem(Batched_Explicit_Problem(model="ball", y0=y0b, p0=pb, sw0=swb), rank, WORLD)
    gb = parallel.gather_final(_OracleBackedSolver(ball), N)
    # row reduction mode: each rank contributes the [rows, n] statistics of its block, never the rows
    # (reduced over the members of the shard problem itself: the shard must be unpadded, a padded one is refused)
    padded_refused = True
    if shard.n_valid != shard.N:
        try:
            parallel.gather_row_statistics(_RowStats(shard))
            padded_refused = False
        except ValueError:
            pass
    assert padded_refused
    exact = parallel.shard_problem(full, rank, WORLD, pad=False)
    assert exact.N == hi - lo == exact.n_valid
    rs = parallel.gather_row_statistics(_RowStats(exact))
    # more ranks than members: the empty rank contributes nothing
    tiny = Batched_Explicit_Problem(model="vdp", y0=y0[:1], p0=mu[:1])
    tshard = parallel.shard_problem(tiny, rank, WORLD, pad=False)
    assert (tshard is None) == (rank > 0)
    rs1 = parallel.gather_row_statistics(_RowStats(tshard) if tshard is not None else None)
    assert (rs1["count"] == 1).all()
    q.put((rank, g["y"], g["t"], g["status"], g["statistics"]["nsteps"], g["stats_per_member"]["nsteps"],
           gb["y"], gb["event_count"], gb["event_t"], gb["event_info"], rs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    import torch.multiprocessing as mp
    import cases
    import ens_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    y0, mu = cases.vdp_sweep(N)
    ref = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, rtol=1e-6, atol=1e-6)
    y0b, pb, swb = cases.ball_sweep(N)
    refb = O.simulate("ball", "Dopri5", y0b, pb, swb, tf=3.0, max_ev=SLOTS, rtol=1e-6, atol=1e-6)
    rows = _oracle_rows(y0, mu)
    for rank, y, t, st, nsteps, per, yb, evc, evt, evi, rs in res:
        assert (rs["count"] == N).all() and np.array_equal(rs["t"], GRID)
        assert np.array_equal(rs["min"], rows.min(axis=1)) and np.array_equal(rs["max"], rows.max(axis=1))
        assert np.allclose(rs["mean"], rows.mean(axis=1), rtol=1e-13, atol=1e-14)
        assert np.allclose(rs["var"], rows.var(axis=1), rtol=1e-11, atol=1e-14)
        assert np.array_equal(yb, refb["y"]) and np.array_equal(evc, refb["stats"]["nstateevents"])
        m = np.isfinite(refb["ev_t"])
        assert np.array_equal(evt[m], refb["ev_t"][m]) and np.array_equal(evi[m], refb["ev_info"][m])
        assert y.shape == (N, 2)
        assert np.array_equal(y, ref["y"]) and np.array_equal(t, ref["t"])
        assert np.array_equal(per, ref["stats"]["nsteps"]) and nsteps == int(ref["stats"]["nsteps"].sum())
        assert (st == 0).all()


def test_multi_rank_bench_exit_path():
    """bench.py at N > 1 ends without tearing communicators down: the peers wait on the rendezvous store until rank 0 has
    printed, then every rank leaves (an 8-GPU run was seen to block for minutes in NCCL teardown, profiles/r2_pcie.md).
    Exercised here with a 2-rank gloo group under torchrun."""
    import subprocess
    import time
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 100),
                        os.path.join(ROOT, "tests", "_leave_helper.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("LINE from rank 0") == 1 and time.time() - t0 < 120
