"""Multi-GPU execution of one ensemble: contiguous member blocks per rank, ONE all-gather at the end.

The members are independent, so the path shards with no data-path collective: rank r of G
integrates members [r*B, min((r+1)*B, N)) with B = ceil(N/G) on its own GPU, and the final
``{t, y, status, statistics, event times / info words}`` are exchanged once: ONE ``ncclAllGather`` over NVLink of the
library's packed SoA result buffers, issued by the library itself (``aens_comm_init`` / ``aens_allgather_results`` in
include/assimulo_cuda.h -- a non-Python host gathers the same way); through host tensors with gloo in the CPU tests.
Dense-output rows stay on the producing GPU.

One process per GPU (torchrun); ``torch.distributed`` is plumbing only.
"""
import numpy as np

from .solvers import STAT_NAMES


def shard_bounds(N, rank, world):
    """[lo, hi) of this rank's contiguous block and the padded block size B = ceil(N / world)."""
    B = -(-int(N) // int(world))
    lo = min(rank * B, N)
    hi = min(lo + B, N)
    return lo, hi, B


def cost_balanced_order(cost, world=1):
    """Fetch order for ``solver.order``: most expensive members first (longest-processing-time rule), dealt
    warp-wise so that every 32-lane chunk holds members of similar cost.  `cost` is any per-member proxy
    (e.g. the stiffness parameter)."""
    cost = np.asarray(cost)
    return np.argsort(-cost, kind="stable").astype(np.intc)


def shard_problem(problem, rank, world, pad=True):
    """A Batched_Explicit_Problem holding only this rank's block.  pad=True (what gather_final needs: all ranks
    exchange equal-sized records) fills the last block up to B = ceil(N / world) by repeating its last member, and a
    rank beyond the last member runs a copy of the last member; the copies are dropped after the gather.
    pad=False gives the block as it is -- required for anything reduced over the shard's members on the device
    (reduce_rows, ensemble_statistics), where a repeated member would be counted twice -- and None for a rank that
    holds no member.  ``n_valid`` on the returned problem is the number of genuine members."""
    from .problem import Batched_Explicit_Problem
    lo, hi, B = shard_bounds(problem.N, rank, world)
    n_valid = hi - lo
    if not pad:
        if hi <= lo:
            return None
        idx = np.arange(lo, hi)
    else:
        if hi <= lo:  # more ranks than members: run a copy of the last member, dropped after the gather
            lo, hi = problem.N - 1, problem.N
        idx = np.concatenate([np.arange(lo, hi), np.full(B - (hi - lo), hi - 1, dtype=np.int64)])
    kw = dict(y0=problem.y0[idx], p0=problem.p0[idx] if problem.n_p else None,
              sw0=problem.sw0[idx] if problem.n_sw else None, t0=problem.t0, name=problem.name)
    if problem.model is not None:
        sp = Batched_Explicit_Problem(model=problem.model, **kw)
    else:
        sp = Batched_Explicit_Problem(source=problem.source, n=problem.n, n_p=problem.n_p, n_g=problem.n_g,
                                      n_sw=problem.n_sw, has_jac=problem.has_jac, time_events=problem.time_events, **kw)
    sp.n_valid = max(n_valid, 0)
    return sp


def comm_init(solver, group=None):
    """Create the library's own NCCL communicator on this rank's handle (aens_comm_init): rank 0 draws the id, the
    process group -- any backend -- only carries its 128 bytes to the other ranks.  Idempotent per solver."""
    import torch.distributed as dist
    if getattr(solver, "_comm_ready", None) == id(solver._get_handle()):
        return
    from . import _core
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [_core.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    solver._get_handle().comm_init(rank, world, box[0])
    solver._comm_ready = id(solver._handle)


def gather_in_simulate(solver, what, group=None):
    """Make every simulate() of `solver` end with the all-gather of the fields `what` (a mask of _core.GATHER_*; 0 = off),
    issued by the library right behind the last integration kernel (aens_set_gather_in_simulate): with the chunked host
    modes the exchange over NVLink overlaps the device->host copies of the last chunks over PCIe."""
    if what:
        comm_init(solver, group)
    solver._get_handle().set_gather_in_simulate(int(what))


def allgather_device(solver, what=0):
    """ONE ncclAllGather of this rank's packed final results, issued by the library on the handle's stream
    (aens_allgather_results; asynchronous).  The gathered records stay on the device (handle.gathered_ptr());
    returns nothing -- pair with solver._handle.sync() or gather_final(...)."""
    solver._handle.allgather_results(int(what))


def gather_final(solver, N_total, group=None, events=None):
    """All-gather the final state of a sharded run.  `solver` is this rank's solver after simulate() on its
    shard.  Returns dict(t[N], y[N,n], status[N], statistics{name:total}, stats_per_member{name:[N]}) on every rank,
    plus event_count[N], event_t[N,slots], event_info[N,slots] (ring order) for models with state or time events.

    With an NCCL process group the exchange is the library's own single ncclAllGather over its packed SoA result
    buffers (aens_comm_init / aens_allgather_results / aens_get_gathered: torch.distributed only distributes the
    communicator id); with gloo (CPU tests of the host logic) the same record travels through host tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = solver.problem.n
    B = solver.N
    backend = dist.get_backend(group)
    has_events = solver.problem.n_g > 0 or solver.problem.time_events
    if events is not None:
        has_events = has_events and bool(events)
    if backend == "nccl":
        comm_init(solver, group)
        h = solver._handle
        from . import _core
        what = _core.GATHER_T | _core.GATHER_Y | _core.GATHER_STATUS | _core.GATHER_STATS
        if has_events:
            what |= _core.GATHER_EVENTS
        h.allgather_results(what)
        g = h.get_gathered(min(int(N_total), world * B), want_events=has_events)
        solver.last_gather_ms = h.last_gather_ms()
        stats = g["stats"]
        res = {"t": g["t"], "y": g["y"], "status": g["status"],
               "stats_per_member": {k: stats[i] for i, k in enumerate(STAT_NAMES)},
               "statistics": {k: int(stats[i].astype(np.int64).sum()) for i, k in enumerate(STAT_NAMES)}}
        if has_events:
            res["event_t"], res["event_info"] = g["ev_t"], g["ev_info"]
            res["event_count"] = res["stats_per_member"]["nstateevents"] + res["stats_per_member"]["ntimeevents"]
        return res
    t_loc = torch.from_numpy(np.ascontiguousarray(solver.t_members))
    y_loc = torch.from_numpy(np.ascontiguousarray(solver.y.T))
    st_loc = torch.from_numpy(np.ascontiguousarray(solver.status.astype(np.int32)))
    stats_loc = torch.from_numpy(np.stack([solver.statistics.per_member(k) for k in STAT_NAMES]).astype(np.int32))
    locs = [t_loc, y_loc, st_loc, stats_loc]
    if has_events:
        _, te, info = solver.event_log
        locs += [torch.from_numpy(np.ascontiguousarray(te.T)), torch.from_numpy(np.ascontiguousarray(info.T.astype(np.int32)))]
    outs = []
    for loc in locs:
        lst = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(lst, loc, group=group)
        outs.append(torch.stack(lst).numpy())
    t_all, y_all, st_all, stats_all = outs[:4]
    ev_all = outs[4:]
    res = assemble(t_all, y_all, st_all, stats_all, N_total)
    if has_events:  # [world, slots, B] -> [N, slots]; ring order as in Batched_Explicit_ODE.event_log
        S = ev_all[0].shape[1]
        res["event_t"] = np.transpose(ev_all[0], (0, 2, 1)).reshape(world * B, S)[:N_total]
        res["event_info"] = np.transpose(ev_all[1], (0, 2, 1)).reshape(world * B, S)[:N_total]
        spm = res["stats_per_member"]
        res["event_count"] = spm["nstateevents"] + spm["ntimeevents"]
    return res


def assemble(t_all, y_all, st_all, stats_all, N_total):
    """[world, ...] per-rank blocks (SoA) -> member-major arrays of the first N_total members."""
    world, n, B = y_all.shape
    t = t_all.reshape(world * B)[:N_total]
    y = np.transpose(y_all, (0, 2, 1)).reshape(world * B, n)[:N_total]
    status = st_all.reshape(world * B)[:N_total]
    stats = np.transpose(stats_all, (1, 0, 2)).reshape(stats_all.shape[1], world * B)[:, :N_total]
    names = STAT_NAMES
    return {"t": t, "y": y, "status": status, "stats_per_member": {k: stats[i] for i, k in enumerate(names)},
            "statistics": {k: int(stats[i].astype(np.int64).sum()) for i, k in enumerate(names)}}


def combine_row_statistics(parts):
    """Merge per-shard row statistics (dicts with mean/var/min/max [rows, n] and count [rows], as returned by
    ``solver.row_statistics``) into the statistics of the whole ensemble: counts add, means and variances combine with
    the pairwise update of Chan, Golub & LeVeque, extrema by min/max.  Rows no member of a shard reached (count 0,
    NaN entries) do not contribute."""
    rows, n = np.asarray(parts[0]["mean"]).shape
    cnt = np.zeros(rows)
    mean = np.zeros((rows, n))
    m2 = np.zeros((rows, n))
    mn = np.full((rows, n), np.inf)
    mx = np.full((rows, n), -np.inf)
    for p in parts:
        cb = np.asarray(p["count"], dtype=np.float64)
        ok = cb > 0
        mb = np.where(ok[:, None], np.asarray(p["mean"]), 0.0)
        m2b = np.where(ok[:, None], np.asarray(p["var"]) * cb[:, None], 0.0)
        tot = cnt + cb
        w = np.divide(cb, tot, out=np.zeros_like(tot), where=tot > 0)
        d = mb - mean
        mean = mean + d * w[:, None]
        m2 = m2 + m2b + d * d * (cnt * w)[:, None]
        cnt = tot
        mn = np.where(ok[:, None], np.minimum(mn, np.where(ok[:, None], p["min"], np.inf)), mn)
        mx = np.where(ok[:, None], np.maximum(mx, np.where(ok[:, None], p["max"], -np.inf)), mx)
    none = (cnt == 0)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        var = m2 / cnt[:, None]
    out = {"mean": np.where(none, np.nan, mean), "var": np.where(none, np.nan, var),
           "min": np.where(none, np.nan, mn), "max": np.where(none, np.nan, mx), "count": cnt.astype(np.int64)}
    if "t" in parts[0]:
        out["t"] = np.asarray(parts[0]["t"])
    return out


def gather_row_statistics(solver, group=None):
    """Row statistics of a sharded run with ``reduce_rows``: every rank contributes the [rows, n] reductions of its
    block (a few KB -- the rows themselves never leave the GPUs) and gets the statistics of the whole ensemble.
    The shard must be unpadded (``shard_problem(..., pad=False)``): a repeated member would be counted twice.  A rank
    without members passes ``solver=None``."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = None
    if solver is not None:
        nv = getattr(solver.problem, "n_valid", solver.N)
        if nv != solver.N:
            raise ValueError("row statistics of a padded shard are biased (%d of its %d members are copies): build the "
                             "shard with shard_problem(..., pad=False)" % (solver.N - nv, solver.N))
        mine = solver.row_statistics
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    parts = [p for p in parts if p is not None]
    if not parts:
        raise ValueError("no rank holds row statistics")
    return combine_row_statistics(parts)
