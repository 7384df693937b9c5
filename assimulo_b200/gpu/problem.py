"""Batched_Explicit_Problem -- the ensemble twin of ``assimulo.problem.Explicit_Problem``
(reference: src/problem.pyx:138-171, 403-488).

An Explicit_Problem holds ``rhs, y0, t0, p0, sw0, name`` and optional ``jac / state_events / handle_event`` as
Python callables for ONE trajectory.  A Batched_Explicit_Problem holds the same things for N independent
members; the callables are CUDA C device functions, either picked from the built-in model library
(``model="vdp"``) or supplied as a source string compiled by NVRTC (``source=...``), because a Python function
cannot run on the device.
"""
import numpy as np

from .exception import AssimuloException

#: built-in models: name -> (n, n_p, n_g, n_sw, has_jac, default parameter row)
BUILTIN_MODELS = {
    "vdp": (2, 1, 0, 0, True, [1.0]),
    "ball": (2, 2, 2, 2, False, [-9.81, 0.88]),
    "robertson": (3, 3, 0, 0, True, [0.04, 1e4, 3e7]),
    "gyro": (12, 3, 0, 0, False, [1000.0, 5000.0, 6000.0]),
    "linear": (1, 3, 0, 0, True, [-1.0, 0.0, 0.0]),
    "linear_ev": (1, 3, 1, 1, True, [-1.0, 0.0, 0.5]),
    "extended": (3, 0, 3, 3, False, []),
    "time_ev": (1, 6, 0, 0, True, [1.0, 2.0, 2.5, 3.0, 1.0, 1.0]),
}
#: built-in models that define time events (Explicit_Problem.time_events, src/explicit_ode.pyx:212-216)
TIME_EVENT_MODELS = ("time_ev",)


class Batched_Explicit_Problem(object):
    """N independent explicit ODEs  y_i' = f(t, y_i, sw_i; p_i),  y_i(t0) = y0_i.

    Parameters::

        model   name of a built-in model (see BUILTIN_MODELS), or None when `source` is given
        source  CUDA C source defining ``aens_rhs`` (and ``aens_jac``, ``aens_events``, ``aens_handle_event``,
                ``aens_time_events`` / ``aens_handle_time_event`` as announced by has_jac / n_g / time_events),
                see include/assimulo_cuda.h
        y0      [N, n] initial states (a single [n] row is broadcast when p or N fixes the ensemble size)
        p0      [N, n_p] parameters (Explicit_Problem.p0)
        sw0     [N, n_sw] switches (Explicit_Problem.sw0)
        t0      common start time
        N       ensemble size, only needed when neither y0 nor p0 is 2-D
        detect_shared_y0  notice [N, n] initial states whose rows are all equal (default): the solvers then send
                one row to the device instead of N
    """

    def __init__(self, model=None, y0=None, p0=None, sw0=None, t0=0.0, name="---", source=None, n=None, n_p=0,
                 n_g=0, n_sw=0, has_jac=False, N=None, time_events=False, fused_rhs=False, detect_shared_y0=True):
        if (model is None) == (source is None):
            raise AssimuloException("Give exactly one of model= (built-in) or source= (CUDA C string).")
        if model is not None:
            if model not in BUILTIN_MODELS:
                raise AssimuloException("Unknown built-in model '%s'. Available: %s" % (model, sorted(BUILTIN_MODELS)))
            n, n_p, n_g, n_sw, has_jac, pdef = BUILTIN_MODELS[model]
            time_events = model in TIME_EVENT_MODELS
        else:
            if n is None:
                raise AssimuloException("source= requires the state dimension n.")
            pdef = [0.0] * n_p
        self.model, self.source, self.name = model, source, name
        self.n, self.n_p, self.n_g, self.n_sw, self.has_jac = int(n), int(n_p), int(n_g), int(n_sw), bool(has_jac)
        #: the model defines aens_time_events / aens_handle_time_event (Explicit_Problem.time_events)
        self.time_events = bool(time_events)
        #: the source also defines aens_rhs_h (hk = h * f, AENS_MODEL_RHS_H); built-in models carry their own
        self.fused_rhs = bool(fused_rhs) and source is not None
        self.t0 = float(t0)
        if y0 is None:
            raise AssimuloException("y0 must be given.")
        y0 = np.asarray(y0, dtype=np.float64)
        p0 = np.asarray(pdef if p0 is None else p0, dtype=np.float64)
        sizes = [a.shape[0] for a in (y0, p0) if a.ndim == 2]
        if sw0 is not None and np.asarray(sw0).ndim == 2:
            sizes.append(np.asarray(sw0).shape[0])
        if N is not None:
            sizes.append(int(N))
        N = max(sizes) if sizes else 1
        if N < 1:
            raise AssimuloException("The ensemble is empty: y0 / p0 / sw0 need at least one member.")
        if y0.shape[-1] != self.n or y0.ndim > 2:
            raise AssimuloException("y0 must have shape [n] or [N, n] with n = %d." % self.n)
        self.N = N
        #: the one initial state all members share, when y0 was given as a single row (a pure parameter sweep): the
        #: solvers then pass this row instead of [N, n] to the device (aens_set_shared_y0)
        self.y0_row = y0.reshape(-1).copy() if (y0.size == self.n and N > 1) else None
        if detect_shared_y0 and self.y0_row is None and y0.ndim == 2 and N > 1 and y0.shape[0] == N and \
                (y0 == y0[0]).all():
            # the same sweep written the long way (N identical rows, as a loop over Explicit_Problem(rhs_i, y0) holds
            # them): detected once here so that the solvers still send ONE row instead of [N, n] per simulate()
            self.y0_row = y0[0].copy()
        self.y0 = np.ascontiguousarray(np.broadcast_to(y0.reshape(-1, self.n) if y0.ndim == 2 else y0[None, :],
                                                       (N, self.n)))
        if self.n_p:
            p0 = p0.reshape(-1, self.n_p) if p0.ndim == 2 else p0.reshape(1, self.n_p) if p0.size == self.n_p \
                else p0.reshape(-1, 1)
            self.p0 = np.ascontiguousarray(np.broadcast_to(p0, (N, self.n_p)))
        else:
            self.p0 = np.zeros((N, 0))
        if self.n_sw:
            if sw0 is None:
                sw0 = self._default_sw0()
            raw = np.asarray(sw0)
            if raw.dtype == np.uint8 and raw.shape == (N, self.n_sw) and raw.flags.c_contiguous and raw.max(initial=0) <= 1:
                self.sw0 = raw   # already in the device's form: kept as it is (it may be page-locked memory)
            else:
                sw0 = np.asarray(sw0, dtype=bool)
                self.sw0 = np.ascontiguousarray(np.broadcast_to(sw0.reshape(-1, self.n_sw), (N, self.n_sw))).astype(np.uint8)
        else:
            self.sw0 = np.zeros((N, 0), dtype=np.uint8)

    def _default_sw0(self):
        return {"ball": [True, False], "linear_ev": [True], "extended": [False, True, True]}.get(
            self.model, [True] * self.n_sw)

    # Hooks with the reference's names; they run on the host around the device launch.
    def initialize(self, solver):
        pass

    def finalize(self, solver):
        pass

    def handle_result(self, solver, t, y):
        """Called once per output row with y of shape [N, n] (the reference calls it per point with y[n],
        src/problem.pyx:145-158)."""
        solver.t_sol.append(t)
        solver.y_sol.append(y)
