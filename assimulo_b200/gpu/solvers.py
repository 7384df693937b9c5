"""Batched solver classes with the reference's option surface.

Each class mirrors one ``Explicit_ODE`` subclass of the reference (same option names, defaults, validation
messages and statistics keys) but integrates all N members of a Batched_Explicit_Problem in ONE kernel launch
through libassimulo_cuda.so:

    Dopri5         src/solvers/runge_kutta.py:26-428    (core thirdparty/hairer/dopri5.f)
    RungeKutta34   src/solvers/runge_kutta.py:430-743
    RungeKutta4    src/solvers/runge_kutta.py:746-872
    ExplicitEuler  src/solvers/euler.pyx:485-692
    Radau5ODE      src/solvers/radau5.py:69-430, src/lib/radau_core.py (core thirdparty/radau5/radau5.c)

``simulate(tfinal, ncp)`` follows ODE.simulate (src/ode.pyx:192-332): output grid ``linspace(t0, tfinal, ncp+1)[1:]``
(or ``ncp_list``), continuation from the current time on a second call, statistics reset per call.  It returns
``(t_sol, y_sol)`` with ``y_sol[k]`` of shape ``[N, n]``; with ``ncp=0`` the rows are the initial and the final
state (the members' internal steps differ, so there is no common internal grid to report).

There is no CPU fallback: creating the device handle raises if no CUDA device is usable.
"""
import numpy as np

from .exception import (AssimuloException, Dopri5_Exception, Explicit_ODE_Exception, Radau_Exception,
                        Rodas_Exception, TimeLimitExceeded)
from .problem import Batched_Explicit_Problem

ARITH = {"strict": 0, "fast": 1}
STAT_NAMES = ("nsteps", "nfcns", "nerrfails", "njacs", "nlus", "nstatefcns", "nstateevents", "nstepstotal",
              "ntimeevents", "nniters", "nnfails")

# status codes (include/assimulo_cuda.h)
ST_COMPLETE, ST_TERMINATED, ST_TIME_LIMIT = 0, 1, -9
#: how simulate() moves host arrays when it takes the one-call host path (aens_set_host_mode)
HOST_MODES = {"auto": 0, "staged": 1, "zero-copy": 2, "zero-copy-in": 3, "zero-copy-out": 4}


def _core():
    from . import _core as c  # the Cython module; import lazily so that option handling works without it
    return c


class Statistics(dict):
    """Totals over the ensemble under the reference's key names (src/support.pyx:33-83, src/ode.pyx:153-172);
    ``per_member(key)`` gives the int32 array over members."""

    def __init__(self, solver):
        dict.__init__(self, {k: 0 for k in STAT_NAMES})
        self._solver = solver

    def reset(self):
        for k in self:
            self[k] = 0

    def per_member(self, key):
        return self._solver._stats_per_member(key)


class Batched_Explicit_ODE(object):
    """Base class: the batched counterpart of Explicit_ODE (src/explicit_ode.pyx:69-135) + ODE (src/ode.pyx)."""
    _solver_name = None
    _exc = Explicit_ODE_Exception
    #: the reference's core of this solver carries the direction of integration (POSNEG)
    _supports_backward = False

    def __init__(self, problem, device=None, comm=None):
        if not isinstance(problem, Batched_Explicit_Problem):
            raise Explicit_ODE_Exception("The problem needs to be a subclass of a Batched_Explicit_Problem.")
        self.problem = problem
        # the reference's general options (src/ode.pyx:45-52) followed by the batched solver's own
        self.options = {"report_continuously": False, "display_progress": True, "verbosity": 30, "backward": False,
                        "store_event_points": True, "time_limit": 0, "clock_step": False, "num_threads": 1,
                        "arith": "fast", "refill_threshold": 32, "order": None, "event_slots": 16,
                        "reuse_buffers": False, "host_chunks": None, "reduce_rows": False, "host_mode": "auto",
                        "lanes_per_member": 0}
        self.supports = {"state_events": False, "interpolated_output": False, "report_continuously": False}
        self.problem_info = {"dim": problem.n, "dimRoot": problem.n_g, "state_events": problem.n_g > 0,
                             "jac_fcn": problem.has_jac, "switches": problem.n_sw > 0}
        self.statistics = Statistics(self)
        self._device = device
        self._comm = comm          # a parallel.Shard describing this rank's block, or None
        self._handle = None
        self._leny = problem.n
        self._dirty = True         # ensemble must be (re)uploaded
        self._launched = False
        self.t0 = problem.t0
        self.t = problem.t0
        self.N = problem.N
        # no defensive copies of ensemble-sized arrays (they are only read when uploaded): if the caller built the problem
        # on page-locked arrays, the very first simulate() can already hand them to the device as they are
        self.y = problem.y0
        self._y_row = problem.y0_row   # set while every member still sits at one shared initial state
        self.sw = problem.sw0.astype(bool)
        self._sw_in = problem.sw0      # the uint8 array the device reads (self.sw is the reference's bool view of it)
        self._status = None        # per-member arrays are materialised on first access (see the properties)
        self._t_members = None
        self._events = None
        self.n_not_complete = 0
        self.t_sol, self.y_sol = [], []
        self.elapsed_kernel_ms = None

    # ---- device plumbing -------------------------------------------------------------------------
    def _get_handle(self):
        if self._handle is None:
            c = _core()
            dev = self._device
            if dev is None:
                dev = 0
            self._handle = c.Handle(int(dev))
            if self.problem.model is not None:
                self._handle.set_model_builtin(self.problem.model)
            else:
                p = self.problem
                self._handle.set_model_source(p.source, p.n, p.n_p, p.n_g, p.n_sw,
                                              int(bool(p.has_jac)) | (2 if p.time_events else 0) | (4 if p.fused_rhs else 0))
        return self._handle

    def _opts_dict(self):
        c = _core()
        d = {"solver": c.SOLVER_IDS[self._solver_name], "arith": ARITH[self.options["arith"]]}
        for k in ("maxsteps", "newt", "rtol", "h", "inith", "safe", "fac1", "fac2", "beta", "thet", "quot1", "quot2"):
            if k in self.options and self.options[k] is not None:
                d[k] = self.options[k]
        if "atol" in self.options:
            d["atol"] = np.asarray(self.options["atol"], dtype=np.float64)
        if self.options.get("fnewt"):
            d["fnewt"] = self.options["fnewt"]
        maxh = self.options.get("maxh")
        if maxh is not None and np.isfinite(maxh):
            d["maxh"] = maxh  # inf / None -> 0 = unset (dopri5.f:301-305, radau5.c:210-212)
        if "usejac" in self.options:
            d["usejac"] = int(bool(self.options["usejac"]))
        return d

    def _stats_per_member(self, key):
        if self._handle is None or not self._launched:
            return np.zeros(self.N, dtype=np.intc)
        return self._handle.get_stats(key)

    # ---- reference surface ------------------------------------------------------------------------
    def re_init(self, t0, y0, sw0=None):
        """Reinitiates the solver (src/explicit_ode.pyx:108-135)."""
        y0 = np.asarray(y0, dtype=np.float64)
        if y0.shape[-1] != self._leny:
            raise Explicit_ODE_Exception("y0 must be of the same length as the original problem.")
        self.t = float(t0)
        # no defensive copies of ensemble-sized arrays: they are only read when uploaded to the device
        self._y_row = y0.reshape(-1).copy() if (y0.size == self._leny and self.N > 1) else None
        if self._y_row is not None:   # one shared initial state: a stride-0 view, nothing ensemble-sized is built
            self.y = np.broadcast_to(self._y_row, (self.N, self._leny))
        else:
            self.y = np.ascontiguousarray(y0.reshape(-1, self._leny))
            if self.y.shape[0] != self.N:
                raise Explicit_ODE_Exception("y0 must have shape [n] or [N, n] with N = %d." % self.N)
        if sw0 is not None:
            if sw0 is self.problem.sw0:
                self._sw_in = sw0
            else:
                self._sw_in = np.ascontiguousarray(np.broadcast_to(
                    np.asarray(sw0, dtype=bool).reshape(-1, self.problem.n_sw), (self.N, self.problem.n_sw))).astype(np.uint8)
            self.sw = self._sw_in.astype(bool)
        self._status, self._t_members, self._events = None, None, None
        self.n_not_complete = 0
        self._dirty = True

    def reset(self):
        y0 = self.problem.y0_row if self.problem.y0_row is not None else self.problem.y0
        self.re_init(self.problem.t0, y0, self.problem.sw0 if self.problem.n_sw else None)

    def initialize(self):
        self.statistics.reset()

    def finalize(self):
        pass

    def simulate(self, tfinal, ncp=0, ncp_list=None):
        t0 = self.t
        self.t_sol, self.y_sol = [], []
        try:
            tfinal = float(tfinal)
        except (ValueError, TypeError):
            raise AssimuloException("Final time must be an integer or float.")
        backward = bool(self.options["backward"])
        if self.t > tfinal and not backward:
            raise AssimuloException("Final time {} must be greater than start time {}.\n Perhaps you should consider "
                                    "to reset the integration. If the intention is to integrate backwards, please use "
                                    "the backwards option.".format(tfinal, self.t))
        if backward and self.t < tfinal:
            raise AssimuloException("Final time {} must be smaller than start time {} when integrating "
                                    "backwards.".format(tfinal, self.t))
        if not isinstance(ncp, int):
            raise AssimuloException("Number of communication points must be an integer")
        if ncp < 0:
            ncp = 0
        if (ncp != 0 or ncp_list is not None) and not self.supports["interpolated_output"]:
            ncp, ncp_list = 0, None  # e.g. RungeKutta4 (src/ode.pyx:260-263)
        if ncp != 0:
            output_list = np.linspace(t0, tfinal, ncp + 1)[1:]
        elif ncp_list is not None and backward:   # src/ode.pyx:277-281
            nl = np.array(ncp_list, dtype=float, ndmin=1)
            output_list = -np.sort(-nl[np.logical_and(nl < t0, nl >= tfinal)])
            if len(output_list) == 0 or output_list[-1] > tfinal:
                output_list = np.append(output_list, tfinal)
        elif ncp_list is not None:
            nl = np.array(ncp_list, dtype=float, ndmin=1)
            output_list = nl[np.logical_and(nl > t0, nl <= tfinal)]
            if len(output_list) == 0 or output_list[-1] < tfinal:
                output_list = np.append(output_list, tfinal)
        else:
            output_list = None

        self.problem.initialize(self)
        self.initialize()
        h = self._get_handle()
        order = self.options["order"]
        reverse = 0
        if isinstance(order, str):
            if order not in ("ascending", "descending", "two-ended"):
                raise AssimuloException("order must be None, 'ascending', 'descending', 'two-ended' or a permutation "
                                        "of range(N).")
            reverse = {"ascending": 0, "descending": 1, "two-ended": 2}[order]
            order = None
        h.set_fetch_reverse(reverse)
        h.set_opts(self._opts_dict())
        h.set_time_limit(float(self.options["time_limit"] or 0.0))
        h.set_lanes_per_member(int(self.options["lanes_per_member"]))
        h.set_backward(backward)
        h.set_host_mode(HOST_MODES[self.options["host_mode"]])
        pin = None
        if self.options["reuse_buffers"]:
            # page-locked result buffers allocated once and overwritten by every simulate() call: rows returned
            # by an earlier call alias them (documented trade: no ensemble-sized allocation or copy per call)
            if getattr(self, "_pin", None) is None or self._pin[0].shape[0] != self.N:
                c = _core()
                self._pin = (c.pinned_empty((self.N,), np.float64), c.pinned_empty((self.N, self._leny), np.float64),
                             c.pinned_empty((self.N,), np.intc))
            pin = self._pin
        self.problem.handle_result(self, t0, self.y)
        if self._dirty and output_list is None and order is None and not backward:
            # host -> device -> host in one chunk-pipelined call (aens_simulate_host): uploads, kernels and
            # downloads of consecutive member ranges overlap.  Only y comes back eagerly; the per-member end times
            # and status codes are fetched on first access -- the summary says whether any member did not complete.
            nch = self.options["host_chunks"] or max(1, min(32, self.N >> 19))
            h.set_output(None, int(self.options["event_slots"]))
            h.set_schedule(int(self.options["refill_threshold"]), None)
            zc_out = self.options["host_mode"] in ("zero-copy", "zero-copy-out")
            if pin is not None:
                y = pin[1]
            elif zc_out:   # the kernel writes the final rows itself: they must land in page-locked memory
                y = _core().pinned_empty((self.N, self._leny), np.float64)
            else:
                y = np.empty((self.N, self._leny))
            sw = None
            if self.problem.n_sw:
                if pin is not None or zc_out:
                    if getattr(self, "_pin_sw", None) is None or self._pin_sw.shape[0] != self.N:
                        self._pin_sw = _core().pinned_empty((self.N, self.problem.n_sw), np.uint8)
                    sw = self._pin_sw
                else:
                    sw = np.empty((self.N, self.problem.n_sw), dtype=np.uint8)
            # a sweep whose members share one initial state sends that row, not [N, n] (aens_set_shared_y0)
            y_in = self._y_row if self._y_row is not None else self.y
            sums, nbad = h.simulate_host(y_in, self.problem.p0 if self.problem.n_p else None,
                                         self._sw_in if self.problem.n_sw else None, self.t, tfinal,
                                         out_y=y, out_sw=sw, nchunks=int(nch), N=self.N)
            self._y_row = None
            self._dirty = False
            self._launched = True
            self.elapsed_kernel_ms = None
            self.y = y
            self._status, self._t_members, self._events = None, None, None
        else:
            if self._dirty:
                h.set_ensemble(self._y_row if self._y_row is not None else self.y,
                               self.problem.p0 if self.problem.n_p else None,
                               self._sw_in if self.problem.n_sw else None, self.t, N=self.N)
                self._dirty = False
            self._y_row = None
            reduce_rows = bool(self.options["reduce_rows"]) and output_list is not None
            if reduce_rows != getattr(self, "_reducing", False):
                h.set_output(None, int(self.options["event_slots"]))  # the mode is switched with no grid set
                h.set_row_reduction(reduce_rows)
                self._reducing = reduce_rows
            h.set_output(output_list, int(self.options["event_slots"]))
            h.set_schedule(int(self.options["refill_threshold"]), order)
            h.simulate(tfinal)
            self._launched = True
            self.elapsed_kernel_ms = h.last_kernel_ms()
            if pin is not None:
                tm, y, sw, status = h.get_final(out_t=pin[0], out_y=pin[1], out_status=pin[2])
            else:
                tm, y, sw, status = h.get_final()
            self._t_members, self.y, self._status, self._events = tm, y, status, None
            sums, nbad = h.get_summary()
        self._tfinal = tfinal
        self.n_not_complete = int(nbad)
        if sw is not None:
            self.sw = sw.astype(bool)
            self._sw_in = sw
        for k in STAT_NAMES:
            self.statistics[k] = int(sums[k])
        self.row_statistics = None
        if output_list is not None and self.options["reduce_rows"]:
            # the rows were reduced over the ensemble on the device: handle_result sees the final states
            # only, the per-row ensemble statistics are in self.row_statistics
            self.row_statistics = dict(h.get_row_reduction(), t=np.array(output_list))
            self.problem.handle_result(self, tfinal, self.y)
        elif output_list is not None:
            dense = h.get_dense()
            for k, tk in enumerate(output_list):
                self.problem.handle_result(self, float(tk), dense[k])
        else:
            self.problem.handle_result(self, tfinal, self.y)
        # the ensemble's current time: members that failed or were terminated stopped early (see status)
        self.t = tfinal if nbad == 0 else float(np.max(self.t_members) if backward else np.min(self.t_members))
        self.finalize()
        self.problem.finalize(self)
        if nbad and self.options["time_limit"] and (self.status == ST_TIME_LIMIT).any():
            # the reference raises from report_solution in the middle of the run (src/explicit_ode.pyx:290-292); here
            # every member that was started has finished and its result is kept (self.y, self.status, statistics)
            late = int((self.status == ST_TIME_LIMIT).sum())
            raise TimeLimitExceeded("The time limit was exceeded at integration time %.8E (%d of %d members not "
                                    "integrated)." % (self.t, late, self.N))
        if self.options["reuse_buffers"]:
            return self.t_sol, self.y_sol  # list of [N, n] rows, no ensemble-sized stacking copy
        return self.t_sol, np.array(self.y_sol)

    def ensemble_statistics(self):
        """mean / var / min / max of the current (final) states over the members that completed, reduced on the device
        (ensemble outputs without materialising ensemble-sized arrays on the host)."""
        if self._handle is None or not self._launched:
            raise AssimuloException("simulate() first.")
        return self._handle.reduce_final()

    def end_time_statistics(self, status=1):
        """mean / var / min / max / count of the members' end times over the members with the given status code,
        reduced on the device.  status=1 (terminated by handle_event) gives first-passage-time statistics of a model
        that stops at its first event."""
        if self._handle is None or not self._launched:
            raise AssimuloException("simulate() first.")
        return self._handle.reduce_end_times(int(status))

    @property
    def status(self):
        """Per-member status codes (include/assimulo_cuda.h AENS_ST_*); 0 = reached tfinal."""
        if self._status is None:
            if self._launched and self.n_not_complete:
                self._status = self._handle.get_final(want_y=False, want_sw=False, want_t=False)[3]
            else:
                self._status = np.zeros(self.N, dtype=np.intc)
        return self._status

    @property
    def t_members(self):
        """Per-member end times (differ from self.t only for members that failed or were terminated)."""
        if self._t_members is None:
            if self._launched and self.n_not_complete:
                self._t_members = self._handle.get_final(want_y=False, want_sw=False)[0]
            else:
                self._t_members = np.full(self.N, self.t)
        return self._t_members

    __call__ = simulate

    @property
    def event_log(self):
        """(count[N], t_event[N, slots], info[N, slots]): ring of the last `event_slots` state events of each
        member (state and time events).  info bit 2i: indicator i fired; bit 2i+1: event_info[i] == +1 (else -1), cf.
        src/explicit_ode.pyx:351-357; bit 30: time event."""
        if self._events is None:
            if self._launched and (self.problem.n_g > 0 or self.problem.time_events):
                self._events = self._handle.get_events()
            else:
                self._events = (np.zeros(self.N, dtype=np.intc), np.zeros((self.N, 0)),
                                np.zeros((self.N, 0), dtype=np.intc))
        return self._events

    def event_times(self, member):
        """Chronological event times of one member (at most `event_slots` most recent)."""
        cnt, te, _ = self.event_log
        c, S = int(cnt[member]), te.shape[1]
        if S == 0 or c == 0:
            return np.zeros(0)
        idx = [(e % S) for e in range(max(0, c - S), c)]
        return te[member, idx]

    def get_options(self):
        return self.options.copy()

    def print_statistics(self, verbose=30):
        print("Final Run Statistics: %s (%d members)\n" % (self.problem.name, self.N))
        for k in STAT_NAMES:
            print(" %-22s : %d" % (k, self.statistics[k]))
        print("\nSolver options:\n\n Solver                  : %s (batched, %s arithmetic)" %
              (self._solver_name, self.options["arith"]))

    # ---- the reference's general solver options (src/ode.pyx:343-507): same names, same validation ------
    def _set_verbosity(self, verb):
        try:
            self.options["verbosity"] = int(verb)
        except Exception:
            raise AssimuloException("Verbosity must be an integer.")
    verbosity = property(lambda self: self.options["verbosity"], _set_verbosity,
                         doc="QUIET 50 / WHISPER 40 / NORMAL 30 / LOUD 20 / SCREAM 10 (src/ode.pyx:343-370).")

    def _set_time_limit(self, time_limit):
        if time_limit < 0:
            raise AssimuloException("The time limit must be positive or zero.")
        self.options["time_limit"] = time_limit
    time_limit = property(lambda self: self.options["time_limit"], _set_time_limit,
                          doc="Upper bound in seconds for one simulate() call, 0 = none (src/ode.pyx:372-392).  The "
                              "device checks it whenever a warp asks for its next members: members not started by "
                              "then keep their state (status -9) and simulate() raises TimeLimitExceeded.")

    def _set_display_progress(self, v):
        self.options["display_progress"] = bool(v)
    display_progress = property(lambda self: self.options["display_progress"], _set_display_progress)

    def _set_report_continuously(self, v):
        self.options["report_continuously"] = bool(v)
    report_continuously = property(lambda self: self.options["report_continuously"], _set_report_continuously,
                                   doc="Accepted for compatibility: with an output grid the device always reports "
                                       "interpolated rows step by step (the reference's report_continuously=True "
                                       "semantics, SURVEY A.7); without one only initial and final rows exist.")

    def _set_number_threads(self, v):
        self.options["num_threads"] = int(v)
    num_threads = property(lambda self: self.options["num_threads"], _set_number_threads)

    def _set_store_event_points(self, v):
        self.options["store_event_points"] = bool(v)
    store_event_points = property(lambda self: self.options["store_event_points"], _set_store_event_points,
                                  doc="Event points are always available through event_log (src/ode.pyx:456-475).")

    def _set_clock_step(self, v):
        self.options["clock_step"] = v
    clock_step = property(lambda self: self.options["clock_step"], _set_clock_step)

    def _set_backward(self, v):
        if bool(v) and not self._supports_backward:
            raise AssimuloException("Backward integration is available with Dopri5, Radau5ODE and RodasODE: %s has no "
                                    "notion of direction in the reference either (its step size is never signed)."
                                    % self._solver_name)
        if bool(v) and (self.problem.time_events or self.options["reduce_rows"]):
            raise AssimuloException("Backward integration is not available with time events or reduce_rows.")
        self.options["backward"] = bool(v)
    backward = property(lambda self: self.options["backward"], _set_backward,
                        doc="Integrate from the current time DOWN to tfinal (src/ode.pyx:490-507).  The device "
                            "integrates the time-reversed problem forward, which reproduces the reference's POSNEG "
                            "integration value for value.")

    # ---- options shared by all batched solvers ------------------------------------------------------
    def _get_arith(self):
        """'fast' (FMA contraction, fp32 step-size-controller powers) or 'strict' (reference operation order,
        no contraction: reproduces the CPU reference step for step)."""
        return self.options["arith"]

    def _set_arith(self, v):
        if v not in ARITH:
            raise AssimuloException("arith must be 'fast' or 'strict'.")
        self.options["arith"] = v
    arith = property(_get_arith, _set_arith)

    def _get_refill(self):
        """A warp fetches new members once this many of its 32 lanes are idle (1..32)."""
        return self.options["refill_threshold"]

    def _set_refill(self, v):
        v = int(v)
        if not 1 <= v <= 32:
            raise AssimuloException("refill_threshold must be in 1..32.")
        self.options["refill_threshold"] = v
    refill_threshold = property(_get_refill, _set_refill)

    def _get_order(self):
        """The order in which members are handed to the warps: None / 'ascending' (by index), 'descending' (from the
        last member down: for ensembles sorted by increasing cost), 'two-ended' (32-member batches alternately from
        both ends: sorted ensembles whose inputs stream from host memory) or a permutation of range(N)."""
        return self.options["order"]

    def _set_order(self, v):
        if v is None or isinstance(v, str):
            self.options["order"] = v
        else:
            self.options["order"] = np.ascontiguousarray(v, dtype=np.intc)
    order = property(_get_order, _set_order)

    def _get_event_slots(self):
        return self.options["event_slots"]

    def _set_event_slots(self, v):
        self.options["event_slots"] = int(v)
    event_slots = property(_get_event_slots, _set_event_slots)

    def _set_host_mode(self, v):
        if v not in HOST_MODES:
            raise AssimuloException("host_mode must be one of %s." % sorted(HOST_MODES))
        self.options["host_mode"] = v
    host_mode = property(lambda self: self.options["host_mode"], _set_host_mode,
                         doc="How simulate() moves host arrays when inputs and results live in host memory: 'auto' / "
                             "'zero-copy' (one launch, the warps read and write page-locked host memory themselves), "
                             "'staged' (chunked copy-engine pipeline), 'zero-copy-in' / 'zero-copy-out' (one side each).")

    def _set_lanes(self, v):
        v = int(v)
        if not 0 <= v <= 32:
            raise AssimuloException("lanes_per_member must be 0 (automatic), 1 or the model's group size.")
        self.options["lanes_per_member"] = v
    lanes_per_member = property(lambda self: self.options["lanes_per_member"], _set_lanes,
                                doc="0 / 1: one thread per member (the default: it measured faster); G: the model's "
                                    "sub-warp-per-member Dopri5 kernel where it defines one (gyro: 4 lanes x 3 states).")

    def _get_reduce_rows(self):
        """True: the rows of the output grid (ncp / ncp_list) are reduced over the ensemble inside the integration
        kernel -- mean, variance, min, max per state component and the number of members that reached the row, in
        ``row_statistics`` after simulate() -- instead of being stored per member ([ncp, N, n] through HBM and PCIe)."""
        return self.options["reduce_rows"]

    def _set_reduce_rows(self, v):
        self.options["reduce_rows"] = bool(v)
    reduce_rows = property(_get_reduce_rows, _set_reduce_rows)

    # helpers for the option properties below
    def _float_opt(self, key, value, msg):
        try:
            self.options[key] = float(value)
        except (ValueError, TypeError):
            raise self._exc(msg)

    def _set_atol_common(self, atol):
        a = np.array(atol, dtype=float) if len(np.array(atol, dtype=float).shape) > 0 else np.array([atol], dtype=float)
        if len(a) == 1:
            a = a * np.ones(self._leny)
        elif len(a) != self._leny:
            raise self._exc("atol must be of length one or same as the dimension of the problem.")
        self.options["atol"] = a

    def _set_rtol_common(self, rtol):
        try:
            r = float(rtol)
        except (ValueError, TypeError):
            raise self._exc("Relative tolerance must be a (scalar) float.")
        if r <= 0.0:
            raise self._exc("Relative tolerance must be a positive (scalar) float.")
        self.options["rtol"] = r

    def _set_maxsteps_common(self, max_steps):
        try:
            max_steps = int(max_steps)
        except (TypeError, ValueError):
            raise self._exc("Maximum number of steps must be a positive integer.")
        self.options["maxsteps"] = max_steps


def _prop(key, setter=None, doc=None):
    def g(self):
        return self.options[key]

    def s(self, v):
        if setter is None:
            self.options[key] = v
        else:
            setter(self, v)
    return property(g, s, doc=doc)


class _FixedStep(Batched_Explicit_ODE):
    def _set_h(self, h):
        try:
            self.options["h"] = float(h)
        except Exception:
            raise AssimuloException("Step-size must be a (scalar) float.")
    h = _prop("h", _set_h, "Fixed step-size, default 0.01 (runge_kutta.py:784, euler.pyx:522).")


class ExplicitEuler(_FixedStep):
    """Explicit Euler with state events and linear dense output (src/solvers/euler.pyx:485-692)."""
    _solver_name = "ExplicitEuler"

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options["h"] = 0.01
        self.options["arith"] = "strict"  # fixed-step parity bar is <=1e-13: reference operation order, no FMA
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)


class ImplicitEuler(_FixedStep):
    """Implicit Euler, fixed step, Newton iteration with a lazily updated Jacobian and a per-member dense LU
    (src/solvers/euler.pyx:29-483).  Per-member failures: status -8 = 'Newton iteration failed'."""
    _solver_name = "ImplicitEuler"

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options.update(h=0.01, usejac=bool(problem.has_jac), newt=7, atol=1.0e-6 * np.ones(self._leny), rtol=1.0e-6)
        self.options["arith"] = "strict"  # see ExplicitEuler
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)

    def _set_newt(self, v):
        try:
            self.options["newt"] = int(v)
        except (ValueError, TypeError):
            raise AssimuloException("The newt must be an integer or float.")

    def _set_usejac(self, v):
        self.options["usejac"] = bool(v)

    def _set_rtol(self, rtol):
        try:
            r = float(rtol)
        except (ValueError, TypeError):
            raise AssimuloException("Relative tolerance must be a (scalar) float.")
        if r <= 0.0:
            raise AssimuloException("Relative tolerance must be a positive (scalar) float.")
        self.options["rtol"] = r

    def _set_atol(self, atol):
        a = np.array(atol, dtype=float) if len(np.array(atol, dtype=float).shape) > 0 else np.array([atol], dtype=float)
        if len(a) == 1:
            a = a * np.ones(self._leny)
        elif len(a) != self._leny:
            raise AssimuloException("atol must be of length one or same as the dimension of the problem.")
        self.options["atol"] = a

    newt = _prop("newt", _set_newt, "Maximal number of Newton iterations, default 7.")
    usejac = _prop("usejac", _set_usejac, "Use the model's analytic Jacobian instead of finite differences.")
    atol = _prop("atol", _set_atol, "Absolute tolerance(s) of the Newton iteration, default 1e-6.")
    rtol = _prop("rtol", _set_rtol, "Relative tolerance of the Newton iteration, default 1e-6.")


class RungeKutta4(_FixedStep):
    """Classical RK4, fixed step; no events, no interpolated output (src/solvers/runge_kutta.py:746-872)."""
    _solver_name = "RungeKutta4"

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options["h"] = 0.01
        self.options["arith"] = "strict"  # see ExplicitEuler


class RungeKutta34(Batched_Explicit_ODE):
    """Adaptive RK of order four, no step rejection (src/solvers/runge_kutta.py:430-743)."""
    _solver_name = "RungeKutta34"

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options.update(atol=1.0e-6 * np.ones(self._leny), rtol=1.0e-6, inith=0.01, maxsteps=10000)
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)

    def _set_inith(self, v):
        try:
            self.options["inith"] = float(v)
        except (ValueError, TypeError):
            raise Explicit_ODE_Exception("Initial step must be an integer or float.")

    def _set_atol(self, atol):
        try:
            a = np.array(atol, dtype=float)
        except (ValueError, TypeError):
            raise Explicit_ODE_Exception("Absolute tolerance must be a positive float or a float vector.")
        if a.ndim == 0:
            a = a * np.ones(self._leny)
        if len(a) != self._leny or (a <= 0.0).any():
            raise Explicit_ODE_Exception("Absolute tolerance must be a positive float or a float vector.")
        self.options["atol"] = a

    def _set_rtol(self, rtol):
        try:
            r = float(rtol)
        except (TypeError, ValueError):
            raise Explicit_ODE_Exception("Relative tolerance must be a float.")
        if r <= 0.0:
            raise Explicit_ODE_Exception("Relative tolerance must be a positive (scalar) float.")
        self.options["rtol"] = r

    inith = _prop("inith", _set_inith, "Initial step-size, default 0.01.")
    atol = _prop("atol", _set_atol, "Absolute tolerance(s), default 1e-6.")
    rtol = _prop("rtol", _set_rtol, "Relative tolerance, default 1e-6.")
    maxsteps = _prop("maxsteps", Batched_Explicit_ODE._set_maxsteps_common, "Maximum number of steps, default 10000.")


class Dopri5(Batched_Explicit_ODE):
    """Dormand-Prince 5(4) with Hairer's step-size control and dense output
    (src/solvers/runge_kutta.py:26-428 over thirdparty/hairer/dopri5.f)."""
    _solver_name = "Dopri5"
    _supports_backward = True
    _exc = Dopri5_Exception

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options.update(safe=0.9, fac1=0.2, fac2=10.0, beta=0.04, maxh=np.inf, inith=0.0,
                            atol=1.0e-6 * np.ones(self._leny), rtol=1.0e-6, maxsteps=100000)
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)

    atol = _prop("atol", Batched_Explicit_ODE._set_atol_common, "Absolute tolerance(s), default 1e-6.")
    rtol = _prop("rtol", Batched_Explicit_ODE._set_rtol_common, "Relative tolerance, default 1e-6.")
    maxsteps = _prop("maxsteps", Batched_Explicit_ODE._set_maxsteps_common, "Maximum number of steps, default 100000.")
    fac1 = _prop("fac1", lambda s, v: s._float_opt("fac1", v, "The fac1 must be an integer or float."),
                 "Lower bound of the step-size ratio, default 0.2.")
    fac2 = _prop("fac2", lambda s, v: s._float_opt("fac2", v, "The fac2 must be an integer or float."),
                 "Upper bound of the step-size ratio, default 10.")
    safe = _prop("safe", lambda s, v: s._float_opt("safe", v, "The safe must be an integer or float."),
                 "Safety factor, default 0.9.")
    beta = _prop("beta", lambda s, v: s._float_opt("beta", v, "The beta must be an integer or float."),
                 "Stabilisation of the step-size controller, default 0.04.")
    inith = _prop("inith", lambda s, v: s._float_opt("inith", v, "The initial step must be an integer or float."),
                  "Initial step-size; 0 lets HINIT choose (default).")

    def _set_maxh(self, v):
        try:
            self.options["maxh"] = float(v)
        except (ValueError, TypeError):
            raise Dopri5_Exception("Maximal stepsize must be a (scalar) float.")
        if self.options["maxh"] < 0:
            raise Dopri5_Exception("Maximal stepsize must be a positiv (scalar) float.")
    maxh = _prop("maxh", _set_maxh, "Maximal step-size, default inf.")


class Radau5ODE(Batched_Explicit_ODE):
    """Radau IIA(5), 3 stages, simplified Newton with per-member dense LU
    (src/solvers/radau5.py:69-430, src/lib/radau_core.py, core thirdparty/radau5/radau5.c)."""
    _solver_name = "Radau5ODE"
    _supports_backward = True
    _exc = Radau_Exception

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options.update(inith=0.01, newt=7, thet=1.e-3, fnewt=None, quot1=1.0, quot2=1.2, fac1=0.2, fac2=8.0,
                            maxh=None, safe=0.9, atol=1.0e-6 * np.ones(self._leny), rtol=1.0e-6,
                            usejac=bool(problem.has_jac), maxsteps=100000, linear_solver="DENSE")
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)

    def _set_newt(self, v):
        try:
            self.options["newt"] = int(v)
        except (ValueError, TypeError):
            raise Radau_Exception("The attribute 'newt' must be an integer or float.")

    def _set_maxh(self, v):
        try:
            self.options["maxh"] = float(v)
        except (ValueError, TypeError):
            raise Radau_Exception("Maximal stepsize must be a (scalar) float.")
        if self.options["maxh"] < 0:
            raise Radau_Exception("Maximal stepsize must be a positiv (scalar) float.")

    def _set_usejac(self, v):
        if v and not self.problem.has_jac:
            raise Radau_Exception("Use of an analytical Jacobian is enabled, but problem does contain a 'jac' function.")
        self.options["usejac"] = bool(v)

    def _set_linear_solver(self, v):
        try:
            up = v.upper()
        except Exception:
            raise Radau_Exception("'linear_solver' parameter needs to be the STRING 'DENSE' or 'SPARSE'. "
                                  "Set value: {}, type: {}".format(v, type(v)))
        if up not in ("DENSE", "SPARSE"):
            raise Radau_Exception("'linear_solver' parameter needs to be either 'DENSE' or 'SPARSE'. Set value: {}".format(v))
        if up == "SPARSE":
            raise Radau_Exception("The batched Radau5ODE only provides the dense per-member LU (SuperLU is out of scope).")
        self.options["linear_solver"] = up

    atol = _prop("atol", Batched_Explicit_ODE._set_atol_common, "Absolute tolerance(s), default 1e-6.")
    rtol = _prop("rtol", Batched_Explicit_ODE._set_rtol_common, "Relative tolerance, default 1e-6.")
    maxsteps = _prop("maxsteps", Batched_Explicit_ODE._set_maxsteps_common, "Maximum number of steps, default 100000.")
    newt = _prop("newt", _set_newt, "Maximal number of Newton iterations, default 7.")
    maxh = _prop("maxh", _set_maxh, "Maximal step-size; None = tfinal - t.")
    usejac = _prop("usejac", _set_usejac, "Use the model's analytic Jacobian instead of finite differences.")
    linear_solver = _prop("linear_solver", _set_linear_solver, "'DENSE' only.")
    inith = _prop("inith", lambda s, v: s._float_opt("inith", v, "The initial step must be an integer or float."),
                  "Initial step-size, default 0.01.")
    fnewt = _prop("fnewt", lambda s, v: s._float_opt("fnewt", v, "The attribute 'fnewt' must be an integer or float."),
                  "Stopping criterion of the Newton iteration; None/0 = automatic.")
    safe = _prop("safe", lambda s, v: s._float_opt("safe", v, "The attribute 'safe' must be an integer or float."),
                 "Safety factor, default 0.9.")
    thet = _prop("thet", lambda s, v: s._float_opt("thet", v, "The attribute 'thet' must be an integer or float."),
                 "Jacobian recomputation threshold, default 1e-3.")
    quot1 = _prop("quot1", lambda s, v: s._float_opt("quot1", v, "The attribute 'quot1' must be an integer or float."),
                  "If quot1 < hnew/hold < quot2 the step size is kept, default 1.0.")
    quot2 = _prop("quot2", lambda s, v: s._float_opt("quot2", v, "The attribute 'quot2' must be an integer or float."),
                  "See quot1, default 1.2.")
    fac1 = _prop("fac1", lambda s, v: s._float_opt("fac1", v, "The attribute 'fac1' must be an integer or float."),
                 "Lower bound of the step-size ratio, default 0.2.")
    fac2 = _prop("fac2", lambda s, v: s._float_opt("fac2", v, "The attribute 'fac2' must be an integer or float."),
                 "Upper bound of the step-size ratio, default 8.")


class RodasODE(Batched_Explicit_ODE):
    """Rosenbrock method of order (3)4 with step-size control and continuous output -- RODAS by Hairer & Wanner
    (src/solvers/rosenbrock.py:259-465 over thirdparty/hairer/rodas_decsol.f).  One LU decomposition per attempt,
    six stages, no Newton iteration; analytic or finite-difference Jacobian, df/dt by a forward difference."""
    _solver_name = "RodasODE"
    _supports_backward = True
    _exc = Rodas_Exception

    def __init__(self, problem, **kw):
        Batched_Explicit_ODE.__init__(self, problem, **kw)
        self.options.update(inith=0.01, fac1=0.2, fac2=6.0, maxh=np.inf, safe=0.9, atol=1.0e-6 * np.ones(self._leny),
                            rtol=1.0e-6, usejac=bool(problem.has_jac), maxsteps=10000)
        self.supports.update(report_continuously=True, interpolated_output=True, state_events=True)

    def _set_usejac(self, v):
        self.options["usejac"] = bool(v)

    def _set_maxh(self, v):
        try:
            self.options["maxh"] = float(v)
        except (ValueError, TypeError):
            raise Rodas_Exception("Maximal stepsize must be a (scalar) float.")
        if self.options["maxh"] < 0:
            raise Rodas_Exception("Maximal stepsize must be a positiv (scalar) float.")

    atol = _prop("atol", Batched_Explicit_ODE._set_atol_common, "Absolute tolerance(s), default 1e-6.")
    rtol = _prop("rtol", Batched_Explicit_ODE._set_rtol_common, "Relative tolerance, default 1e-6.")
    maxsteps = _prop("maxsteps", Batched_Explicit_ODE._set_maxsteps_common, "Maximum number of steps, default 10000.")
    usejac = _prop("usejac", _set_usejac, "Use the model's analytic Jacobian instead of finite differences.")
    maxh = _prop("maxh", _set_maxh, "Maximal step-size, default inf.")
    inith = _prop("inith", lambda s, v: s._float_opt("inith", v, "The initial step must be an integer or float."),
                  "Initial step-size, default 0.01.")
    fac1 = _prop("fac1", lambda s, v: s._float_opt("fac1", v, "The fac1 must be an integer or float."),
                 "Lower bound of the step-size ratio, default 0.2.")
    fac2 = _prop("fac2", lambda s, v: s._float_opt("fac2", v, "The fac2 must be an integer or float."),
                 "Upper bound of the step-size ratio, default 6.")
    safe = _prop("safe", lambda s, v: s._float_opt("safe", v, "The safe must be an integer or float."),
                 "Safety factor, default 0.9.")
