# cython: language_level=3, boundscheck=True, wraparound=False
"""Cython host binding of libassimulo_cuda.so (include/assimulo_cuda.h).

Plays the role ``thirdparty/radau5/radau5ode.pyx`` (RadauMemory + radau5_py_solve) and the f2py module
``assimulo.lib.dopri5`` play in the reference: the only place where Python objects meet the native ABI.  The
difference is that no callback ever crosses back into Python, so the GIL is released around every call that
touches the device (the reference holds it throughout, SURVEY.md 8b).
"""
import numpy as np
cimport numpy as np

np.import_array()

cdef extern from "assimulo_cuda.h":
    int AENS_MAX_N
    int AENS_NSTATS
    ctypedef struct aens_opts_t:
        int solver
        int arith
        int maxsteps
        int newt
        int usejac
        int reserved
        double rtol
        double atol[16]
        double h
        double inith
        double maxh
        double safe
        double fac1
        double fac2
        double beta
        double thet
        double fnewt
        double quot1
        double quot2
    ctypedef struct aens_handle_t:
        pass
    int aens_default_opts(int solver, aens_opts_t *o) nogil
    int aens_create(int device, aens_handle_t **h) nogil
    int aens_destroy(aens_handle_t **h) nogil
    const char *aens_last_error(const aens_handle_t *h) nogil
    const char *aens_version() nogil
    int aens_nstats() nogil
    int aens_set_model_builtin(aens_handle_t *h, const char *name) nogil
    int aens_set_model_source(aens_handle_t *h, const char *src, int n, int n_p, int n_g, int n_sw, int has_jac) nogil
    int aens_model_dims(const aens_handle_t *h, int *n, int *n_p, int *n_g, int *n_sw, int *has_jac) nogil
    int aens_set_ensemble(aens_handle_t *h, long long N, const double *y0, const double *p,
                          const unsigned char *sw0, double t0) nogil
    int aens_set_opts(aens_handle_t *h, const aens_opts_t *o) nogil
    int aens_reset(aens_handle_t *h) nogil
    int aens_set_output(aens_handle_t *h, int n_out, const double *t_out, int event_slots) nogil
    int aens_set_schedule(aens_handle_t *h, int refill_threshold, const int *order) nogil
    int aens_simulate(aens_handle_t *h, double tfinal) nogil
    int aens_simulate_async(aens_handle_t *h, double tfinal) nogil
    int aens_sync(aens_handle_t *h) nogil
    int aens_last_kernel_ms(aens_handle_t *h, float *ms) nogil
    int aens_kernel_ms_history(aens_handle_t *h, float *ms, int n) nogil
    int aens_get_final(aens_handle_t *h, double *t, double *y, unsigned char *sw, int *status) nogil
    int aens_get_stats(aens_handle_t *h, int which, int *out) nogil
    int aens_get_stats_sum(aens_handle_t *h, long long *sums) nogil
    int aens_get_dense(aens_handle_t *h, int slot, double *y_out) nogil
    int aens_get_dense_all(aens_handle_t *h, double *y_out) nogil
    int aens_get_events(aens_handle_t *h, int *count, double *t_ev, int *info) nogil
    int aens_device_ptr(aens_handle_t *h, int which, void **ptr, long long *nbytes) nogil
    int aens_stream(aens_handle_t *h, void **stream) nogil
    int aens_measure_fp64_peak(int device, double seconds, double *tflops, double *mhz) nogil
    long long aens_nvrtc_check(const char *src, int n, int n_p, int n_g, int n_sw, int has_jac, int strict,
                               char *errbuf, int errbuf_len) nogil
    long long aens_launch_count(const aens_handle_t *h) nogil
    int aens_host_alloc(long long nbytes, void **ptr) nogil
    int aens_host_free(void *ptr) nogil
    int aens_set_fetch_reverse(aens_handle_t *h, int reverse) nogil
    int aens_simulate_host(aens_handle_t *h, long long N, const double *y0, const double *p,
                           const unsigned char *sw0, double t0, double tfinal, double *y_out, double *t_out,
                           unsigned char *sw_out, int *status_out, int nchunks, long long *sums,
                           long long *n_not_complete) nogil
    int aens_get_summary(aens_handle_t *h, long long *sums, long long *n_not_complete) nogil
    int aens_set_host_mode(aens_handle_t *h, int mode) nogil
    int aens_set_shared_y0(aens_handle_t *h, int on) nogil
    int aens_reduce_final(aens_handle_t *h, double *mean, double *var, double *mn, double *mx, long long *count) nogil
    int aens_reduce_end_times(aens_handle_t *h, int status, double *mean, double *var, double *mn, double *mx,
                              long long *count) nogil
    int aens_set_row_reduction(aens_handle_t *h, int on) nogil
    int aens_get_row_reduction(aens_handle_t *h, double *mean, double *var, double *mn, double *mx,
                               long long *count) nogil
    int aens_set_time_limit(aens_handle_t *h, double seconds) nogil
    int aens_set_backward(aens_handle_t *h, int on) nogil
    int aens_check_failures(aens_handle_t *h, long long *code) nogil
    int aens_set_lanes_per_member(aens_handle_t *h, int lanes) nogil
    int aens_last_lanes_per_member(const aens_handle_t *h) nogil
    int aens_comm_unique_id(void *id128) nogil
    int aens_comm_init(aens_handle_t *h, int rank, int nranks, const void *id128) nogil
    int aens_comm_destroy(aens_handle_t *h) nogil
    int aens_allgather_results(aens_handle_t *h, int what) nogil
    int aens_set_gather_in_simulate(aens_handle_t *h, int what) nogil
    int aens_last_gather_ms(aens_handle_t *h, float *ms) nogil
    int aens_gathered_ptr(aens_handle_t *h, void **ptr, long long *rec_bytes, int *nranks) nogil
    int aens_get_gathered(aens_handle_t *h, long long n_total, double *t, double *y, int *status, int *stats,
                          double *ev_t, int *ev_info) nogil
    int aens_measure_hostlink(int device, long long nbytes, int mode, double seconds, double *gbs) nogil

STAT_NAMES = ("nsteps", "nfcns", "nerrfails", "njacs", "nlus", "nstatefcns", "nstateevents", "nstepstotal",
              "ntimeevents", "nniters", "nnfails")
MODEL_HAS_JAC, MODEL_TIME_EVENTS, MODEL_RHS_H = 1, 2, 4
GATHER_T, GATHER_Y, GATHER_STATUS, GATHER_STATS, GATHER_EVENTS = 1, 2, 4, 8, 16
HOSTLINK_MODES = ("copy engine H2D", "copy engine D2H", "copy engine H2D+D2H", "SM reads of host memory",
                  "SM writes to host memory", "SM reads+writes")
SOLVER_IDS = {"ExplicitEuler": 0, "RungeKutta4": 1, "RungeKutta34": 2, "Dopri5": 3, "Radau5ODE": 4,
              "ImplicitEuler": 5, "RodasODE": 6}
MAX_N = 16


class AensError(RuntimeError):
    """Raised for a negative return code of the C ABI; carries ``code`` and the library's message."""
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libassimulo_cuda error %d: %s" % (code, msg))
        self.code = code


def version():
    return aens_version().decode()


if aens_nstats() != len(STAT_NAMES) or aens_nstats() != 11:
    # the fixed-size statistic buffers below (long long[11]) must match the library this module was linked against
    raise ImportError("libassimulo_cuda.so reports %d statistic slots, this binding was built for %d"
                      % (aens_nstats(), len(STAT_NAMES)))


def default_opts(solver):
    """Reference defaults of `solver` (name) as a dict."""
    cdef aens_opts_t o
    cdef int rc = aens_default_opts(SOLVER_IDS[solver], &o)
    if rc != 0:
        raise AensError(rc, "unknown solver")
    return _opts_to_dict(&o)


cdef dict _opts_to_dict(aens_opts_t *o):
    return dict(solver=o.solver, arith=o.arith, maxsteps=o.maxsteps, newt=o.newt, usejac=o.usejac, rtol=o.rtol,
                atol=[o.atol[i] for i in range(16)], h=o.h, inith=o.inith, maxh=o.maxh, safe=o.safe,
                fac1=o.fac1, fac2=o.fac2, beta=o.beta, thet=o.thet, fnewt=o.fnewt, quot1=o.quot1, quot2=o.quot2)


def nvrtc_check(src, int n, int n_p=0, int n_g=0, int n_sw=0, int has_jac=0, bint strict=False):
    """Compile-only check of a CUDA model source (no GPU needed).  Returns the cubin size."""
    cdef bytes b = src.encode() if isinstance(src, str) else src
    cdef char err[4096]
    cdef const char *cs = b
    cdef long long r
    err[0] = 0
    with nogil:
        r = aens_nvrtc_check(cs, n, n_p, n_g, n_sw, has_jac, strict, err, 4096)
    if r < 0:
        raise AensError(<int>r, err.decode(errors="replace"))
    return r


cdef class _PinnedBlock:
    """Owner of one cudaHostAlloc block; numpy arrays created by pinned_empty keep it alive via .base."""
    cdef void *ptr
    cdef long long nbytes

    def __cinit__(self):
        self.ptr = NULL
        self.nbytes = 0

    def __dealloc__(self):
        if self.ptr != NULL:
            aens_host_free(self.ptr)
            self.ptr = NULL


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (needs a CUDA device)."""
    dt = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    cdef long long nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    cdef _PinnedBlock blk = _PinnedBlock()
    cdef int rc = aens_host_alloc(nbytes, &blk.ptr)
    if rc != 0:
        raise AensError(rc, "cudaHostAlloc(%d bytes) failed" % nbytes)
    blk.nbytes = nbytes
    cdef np.npy_intp n = nbytes
    cdef np.ndarray raw = np.PyArray_SimpleNewFromData(1, &n, np.NPY_UINT8, blk.ptr)
    np.set_array_base(raw, blk)
    return raw.view(dt).reshape(shape)


def comm_unique_id():
    """128-byte NCCL id for aens_comm_init: rank 0 creates it, the host program hands it to the other ranks."""
    cdef char buf[128]
    cdef int rc = aens_comm_unique_id(buf)
    if rc != 0:
        raise AensError(rc, "aens_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf[:128])


def measure_hostlink(int device=0, long long nbytes=1 << 28, int mode=0, double seconds=0.25):
    """GB/s the host link of `device` sustains in `mode` (index into HOSTLINK_MODES)."""
    cdef double gbs = 0.0
    cdef int rc
    with nogil:
        rc = aens_measure_hostlink(device, nbytes, mode, seconds, &gbs)
    if rc != 0:
        raise AensError(rc, "host link measurement failed (no GPU?)")
    return gbs


cdef _need(np.ndarray a, dtype, long long size, name):
    """explicit argument checks (an `assert` disappears under python -O and a short buffer is a host memory overrun)"""
    if a.dtype != dtype or a.size != size or not a.flags.c_contiguous:
        raise ValueError("%s must be a C-contiguous %s array with %d elements" % (name, np.dtype(dtype).name, size))


def measure_fp64_peak(int device=0, double seconds=1.0):
    cdef double tf = 0.0, mhz = 0.0
    cdef int rc
    with nogil:
        rc = aens_measure_fp64_peak(device, seconds, &tf, &mhz)
    if rc != 0:
        raise AensError(rc, "fp64 peak measurement failed (no GPU?)")
    return tf, mhz


cdef class Handle:
    """One ensemble on one GPU (owns the device memory and a CUDA stream)."""
    cdef aens_handle_t *h
    cdef public int n, n_p, n_g, n_sw, has_jac
    cdef public long long N
    cdef public int n_out, event_slots

    def __cinit__(self, int device=0):
        self.h = NULL
        cdef int rc
        with nogil:
            rc = aens_create(device, &self.h)
        if rc != 0:
            raise AensError(rc, "aens_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)"
                            % device)
        self.N = 0
        self.n_out = 0
        self.event_slots = 16

    def __dealloc__(self):
        if self.h != NULL:
            aens_destroy(&self.h)

    cdef _check(self, int rc):
        if rc < 0:
            raise AensError(rc, aens_last_error(self.h).decode(errors="replace"))
        return rc

    def close(self):
        if self.h != NULL:
            aens_destroy(&self.h)

    def _dims(self):
        cdef int n, p, g, s, j
        self._check(aens_model_dims(self.h, &n, &p, &g, &s, &j))
        self.n, self.n_p, self.n_g, self.n_sw, self.has_jac = n, p, g, s, j

    def set_model_builtin(self, name):
        cdef bytes b = name.encode()
        self._check(aens_set_model_builtin(self.h, b))
        self._dims()

    def set_model_source(self, src, int n, int n_p=0, int n_g=0, int n_sw=0, int has_jac=0):
        # has_jac: 0/1 or a bit set of MODEL_HAS_JAC | MODEL_TIME_EVENTS
        cdef bytes b = src.encode() if isinstance(src, str) else src
        cdef const char *cs = b
        cdef int rc
        with nogil:
            rc = aens_set_model_source(self.h, cs, n, n_p, n_g, n_sw, has_jac)
        self._check(rc)
        self._dims()

    def _shared_row(self, y0, p, sw0, N):
        """y0 given as ONE row [n] shared by all members (aens_set_shared_y0): returns (y0 as [1, n], N) with N taken
        from p / sw0 / the N argument; (y0 as [N, n], N) otherwise."""
        y = np.ascontiguousarray(y0, dtype=np.float64)
        if y.ndim != 1:
            if y.ndim != 2:
                raise ValueError("y0 must have shape [N, n] or [n]")
            self._check(aens_set_shared_y0(self.h, 0))
            return y, len(y)
        if N is None:
            shp = np.shape(p) if (self.n_p and p is not None) else np.shape(sw0) if (self.n_sw and sw0 is not None) else ()
            if len(shp) != 2:
                raise ValueError("a shared y0 row needs a 2-D p or sw0, or N=, to fix the ensemble size")
            N = shp[0]
        self._check(aens_set_shared_y0(self.h, 1))
        return y.reshape(1, -1), int(N)

    def set_ensemble(self, y0, p=None, sw0=None, double t0=0.0, N=None):
        cdef np.ndarray[double, ndim=2, mode="c"] y
        cdef np.ndarray[double, ndim=2, mode="c"] pp
        cdef np.ndarray[np.uint8_t, ndim=2, mode="c"] ss
        cdef const double *pptr = NULL
        cdef const unsigned char *sptr = NULL
        cdef long long N_
        y, N_ = self._shared_row(y0, p, sw0, N)
        N = N_
        if y.shape[1] != self.n:
            raise ValueError("y0 must have shape [N, %d] or [%d]" % (self.n, self.n))
        if self.n_p:
            pp = np.ascontiguousarray(p, dtype=np.float64)
            if pp.shape[0] != N or pp.shape[1] != self.n_p:
                raise ValueError("p must have shape [N, %d]" % self.n_p)
            pptr = &pp[0, 0]
        if self.n_sw:
            ss = np.ascontiguousarray(sw0, dtype=np.uint8)
            if ss.shape[0] != N or ss.shape[1] != self.n_sw:
                raise ValueError("sw0 must have shape [N, %d]" % self.n_sw)
            sptr = &ss[0, 0]
        cdef const double *yptr = &y[0, 0]
        cdef int rc
        with nogil:
            rc = aens_set_ensemble(self.h, N_, yptr, pptr, sptr, t0)
        aens_set_shared_y0(self.h, 0)
        self._check(rc)
        self.N = N_
        self.n_out = 0

    def reset(self):
        self._check(aens_reset(self.h))

    def set_opts(self, dict d):
        cdef aens_opts_t o
        self._check(aens_default_opts(d["solver"], &o))
        o.arith = d.get("arith", o.arith)
        o.maxsteps = d.get("maxsteps", o.maxsteps)
        o.newt = d.get("newt", o.newt)
        o.usejac = d.get("usejac", o.usejac)
        o.rtol = d.get("rtol", o.rtol)
        atol = np.broadcast_to(np.asarray(d.get("atol", 1e-6), dtype=np.float64), (self.n,))
        for i in range(self.n):
            o.atol[i] = atol[i]
        o.h = d.get("h", o.h)
        o.inith = d.get("inith", o.inith)
        o.maxh = d.get("maxh", o.maxh)
        o.safe = d.get("safe", o.safe)
        o.fac1 = d.get("fac1", o.fac1)
        o.fac2 = d.get("fac2", o.fac2)
        o.beta = d.get("beta", o.beta)
        o.thet = d.get("thet", o.thet)
        o.fnewt = d.get("fnewt", o.fnewt)
        o.quot1 = d.get("quot1", o.quot1)
        o.quot2 = d.get("quot2", o.quot2)
        self._check(aens_set_opts(self.h, &o))

    def set_output(self, t_out=None, int event_slots=16):
        cdef np.ndarray[double, ndim=1, mode="c"] tg
        cdef const double *tp = NULL
        cdef int n_out = 0
        if t_out is not None and len(t_out) > 0:
            tg = np.ascontiguousarray(t_out, dtype=np.float64)
            tp = &tg[0]
            n_out = tg.shape[0]
        cdef int rc
        with nogil:
            rc = aens_set_output(self.h, n_out, tp, event_slots)
        self._check(rc)
        self.n_out = n_out
        self.event_slots = event_slots

    def set_schedule(self, int refill_threshold=32, order=None):
        cdef np.ndarray[int, ndim=1, mode="c"] od
        cdef const int *op = NULL
        if order is not None:
            od = np.ascontiguousarray(order, dtype=np.intc)
            if od.shape[0] != self.N or not np.array_equal(np.sort(od), np.arange(self.N, dtype=np.intc)):
                raise ValueError("order must be a permutation of range(N)")
            op = &od[0]
        cdef int rc
        with nogil:
            rc = aens_set_schedule(self.h, refill_threshold, op)
        self._check(rc)

    def simulate(self, double tfinal):
        cdef int rc
        with nogil:
            rc = aens_simulate(self.h, tfinal)
        self._check(rc)

    def set_lanes_per_member(self, int lanes):
        """0 automatic, 1 one thread per member, > 1 the model's sub-warp-per-member kernel or an error."""
        self._check(aens_set_lanes_per_member(self.h, lanes))

    def last_lanes_per_member(self):
        return aens_last_lanes_per_member(self.h)

    def check_failures(self):
        """first bounds violation recorded by an AENS_CHECKED build of the library (0 = none)"""
        cdef long long code = 0
        self._check(aens_check_failures(self.h, &code))
        return code

    def set_backward(self, on):
        """options["backward"]: the next simulate() integrates from the members' current time down to tfinal."""
        self._check(aens_set_backward(self.h, 1 if on else 0))

    def set_time_limit(self, double seconds):
        """ODE.time_limit: members a warp fetches later than `seconds` after the launch started are not integrated
        (status -9); 0 = no limit."""
        self._check(aens_set_time_limit(self.h, seconds))

    # ---- the final all-gather over NCCL (SURVEY 8e) ----
    def comm_init(self, int rank, int nranks, bytes uid):
        if len(uid) != 128:
            raise ValueError("the communicator id is 128 bytes (comm_unique_id)")
        cdef const char *p = uid
        cdef int rc
        with nogil:
            rc = aens_comm_init(self.h, rank, nranks, p)
        self._check(rc)

    def comm_destroy(self):
        self._check(aens_comm_destroy(self.h))

    def allgather_results(self, int what=0):
        """ONE ncclAllGather of this rank's packed results (asynchronous on the handle's stream)."""
        cdef int rc
        with nogil:
            rc = aens_allgather_results(self.h, what)
        self._check(rc)

    def set_gather_in_simulate(self, int what):
        """what != 0: every simulate ends with the all-gather of these fields, issued right behind the last kernel."""
        self._check(aens_set_gather_in_simulate(self.h, what))

    def last_gather_ms(self):
        cdef float ms = 0
        cdef int rc
        with nogil:
            rc = aens_last_gather_ms(self.h, &ms)
        self._check(rc)
        return ms

    def gathered_ptr(self):
        cdef void *p = NULL
        cdef long long b = 0
        cdef int r = 0
        self._check(aens_gathered_ptr(self.h, &p, &b, &r))
        return <size_t>p, b, r

    def get_gathered(self, long long n_total, bint want_t=True, bint want_y=True, bint want_status=True,
                     bint want_stats=True, bint want_events=False):
        """The gathered results of the first n_total members (rank order) as host arrays: dict with t[N], y[N, n],
        status[N], stats[11, N] and, for models with events, ev_t / ev_info [N, slots]."""
        cdef np.ndarray[double, ndim=1, mode="c"] t
        cdef np.ndarray[double, ndim=2, mode="c"] y
        cdef np.ndarray[int, ndim=1, mode="c"] st
        cdef np.ndarray[int, ndim=2, mode="c"] stats
        cdef np.ndarray[double, ndim=2, mode="c"] evt
        cdef np.ndarray[int, ndim=2, mode="c"] evi
        cdef double *tp = NULL
        cdef double *yp = NULL
        cdef int *sp = NULL
        cdef int *ssp = NULL
        cdef double *etp = NULL
        cdef int *eip = NULL
        out = {}
        if want_t:
            t = np.empty(n_total, dtype=np.float64)
            tp = &t[0]
            out["t"] = t
        if want_y:
            y = np.empty((n_total, self.n), dtype=np.float64)
            yp = &y[0, 0]
            out["y"] = y
        if want_status:
            st = np.empty(n_total, dtype=np.intc)
            sp = &st[0]
            out["status"] = st
        if want_stats:
            stats = np.empty((len(STAT_NAMES), n_total), dtype=np.intc)
            ssp = &stats[0, 0]
            out["stats"] = stats
        cdef int S = self.event_slots if ((self.n_g > 0 or (self.has_jac & 2)) and self.event_slots > 0) else 0
        if want_events and S > 0:
            evt = np.empty((n_total, S), dtype=np.float64)
            evi = np.empty((n_total, S), dtype=np.intc)
            etp = &evt[0, 0]
            eip = &evi[0, 0]
            out["ev_t"] = evt
            out["ev_info"] = evi
        cdef int rc
        with nogil:
            rc = aens_get_gathered(self.h, n_total, tp, yp, sp, ssp, etp, eip)
        self._check(rc)
        return out

    def set_host_mode(self, int mode):
        """0 auto, 1 chunked staging pipeline (copy engines both ways), 2 zero-copy (raises at simulate_host if the
        arrays are not mappable), 3 chunked with zero-copy inputs and copy-engine results, 4 the other way round."""
        self._check(aens_set_host_mode(self.h, mode))

    def set_fetch_reverse(self, int reverse):
        self._check(aens_set_fetch_reverse(self.h, reverse))

    def simulate_host(self, y0, p, sw0, double t0, double tfinal, out_y=None, out_t=None, out_sw=None,
                      out_status=None, int nchunks=8, N=None):
        """aens_simulate_host: upload, integrate and download in one chunk-pipelined call.  y0/p/sw0 as for
        set_ensemble; out_* are caller-owned C-contiguous arrays (ideally pinned_empty) or None to skip that
        result.  A 1-D y0 is ONE row shared by all members (only p / sw0 cross PCIe then).  Returns (statistic sums
        dict, number of members that did not complete)."""
        cdef np.ndarray[double, ndim=2, mode="c"] y
        cdef np.ndarray[double, ndim=2, mode="c"] pp
        cdef np.ndarray[np.uint8_t, ndim=2, mode="c"] ss
        cdef const double *pptr = NULL
        cdef const unsigned char *sptr = NULL
        cdef long long N_
        y, N_ = self._shared_row(y0, p, sw0, N)
        N = N_
        if y.shape[1] != self.n:
            raise ValueError("y0 must have shape [N, %d] or [%d]" % (self.n, self.n))
        if self.n_p:
            pp = np.ascontiguousarray(p, dtype=np.float64)
            if pp.shape[0] != N or pp.shape[1] != self.n_p:
                raise ValueError("p must have shape [N, %d]" % self.n_p)
            pptr = &pp[0, 0]
        if self.n_sw:
            ss = np.ascontiguousarray(sw0, dtype=np.uint8)
            if ss.shape[0] != N or ss.shape[1] != self.n_sw:
                raise ValueError("sw0 must have shape [N, %d]" % self.n_sw)
            sptr = &ss[0, 0]
        cdef double *yo = NULL
        cdef double *to = NULL
        cdef unsigned char *so = NULL
        cdef int *sto = NULL
        cdef np.ndarray a
        if out_y is not None:
            a = out_y
            _need(a, np.float64, N * self.n, "out_y")
            yo = <double*>np.PyArray_DATA(a)
        if out_t is not None:
            a = out_t
            _need(a, np.float64, N, "out_t")
            to = <double*>np.PyArray_DATA(a)
        if out_sw is not None and self.n_sw:
            a = out_sw
            _need(a, np.uint8, N * self.n_sw, "out_sw")
            so = <unsigned char*>np.PyArray_DATA(a)
        if out_status is not None:
            a = out_status
            _need(a, np.intc, N, "out_status")
            sto = <int*>np.PyArray_DATA(a)
        cdef const double *yptr = &y[0, 0]
        cdef long long sums[11]
        cdef long long nbad = 0
        cdef int rc
        with nogil:
            rc = aens_simulate_host(self.h, N_, yptr, pptr, sptr, t0, tfinal, yo, to, so, sto, nchunks, sums, &nbad)
        aens_set_shared_y0(self.h, 0)
        self._check(rc)
        self.N = N
        self.n_out = 0
        return {STAT_NAMES[i]: sums[i] for i in range(11)}, nbad

    def reduce_final(self):
        """Ensemble statistics of the final states, computed on the device: dict(mean, var, min, max [n], count [n])."""
        cdef np.ndarray[double, ndim=2, mode="c"] out = np.empty((4, self.n), dtype=np.float64)
        cdef np.ndarray[np.int64_t, ndim=1, mode="c"] cnt = np.empty(self.n, dtype=np.int64)
        cdef double *o = &out[0, 0]
        cdef long long *cp = <long long*>&cnt[0]
        cdef int n = self.n
        cdef int rc
        with nogil:
            rc = aens_reduce_final(self.h, o, o + n, o + 2 * n, o + 3 * n, cp)
        self._check(rc)
        return {"mean": out[0], "var": out[1], "min": out[2], "max": out[3], "count": cnt}

    def reduce_end_times(self, int status=1):
        """mean / var / min / max / count of the end times of the members with the given status, reduced on the
        device (status 1 = terminated by handle_event: first-passage times)."""
        cdef double o[4]
        cdef long long cnt = 0
        cdef int rc
        with nogil:
            rc = aens_reduce_end_times(self.h, status, &o[0], &o[1], &o[2], &o[3], &cnt)
        self._check(rc)
        return {"mean": o[0], "var": o[1], "min": o[2], "max": o[3], "count": cnt}

    def set_row_reduction(self, on):
        """Reduce the dense-output rows over the ensemble inside the kernel instead of storing them (call before
        set_output sets the grid)."""
        self._check(aens_set_row_reduction(self.h, 1 if on else 0))

    def get_row_reduction(self):
        """dict(mean, var, min, max [n_out, n], count [n_out]) of the reduced dense-output rows."""
        cdef int R = self.n_out, n = self.n
        cdef np.ndarray[double, ndim=3, mode="c"] out = np.empty((4, max(R, 1), n), dtype=np.float64)
        cdef np.ndarray[np.int64_t, ndim=1, mode="c"] cnt = np.empty(max(R, 1), dtype=np.int64)
        cdef double *o = &out[0, 0, 0]
        cdef long long *cp = <long long*>&cnt[0]
        cdef long long RN = <long long>max(R, 1) * n
        cdef int rc
        with nogil:
            rc = aens_get_row_reduction(self.h, o, o + RN, o + 2 * RN, o + 3 * RN, cp)
        self._check(rc)
        return {"mean": out[0], "var": out[1], "min": out[2], "max": out[3], "count": cnt}

    def get_summary(self):
        cdef long long s[11]
        cdef long long nbad = 0
        cdef int rc
        with nogil:
            rc = aens_get_summary(self.h, s, &nbad)
        self._check(rc)
        return {STAT_NAMES[i]: s[i] for i in range(11)}, nbad

    def simulate_async(self, double tfinal):
        cdef int rc
        with nogil:
            rc = aens_simulate_async(self.h, tfinal)
        self._check(rc)

    def sync(self):
        cdef int rc
        with nogil:
            rc = aens_sync(self.h)
        self._check(rc)

    def last_kernel_ms(self):
        cdef float ms = 0
        self._check(aens_last_kernel_ms(self.h, &ms))
        return ms

    def kernel_ms_history(self, int n=64):
        """durations (ms) of the most recent integration-kernel launches, newest first (at most 64)"""
        cdef float ms[64]
        cdef int k
        if n > 64:
            n = 64
        with nogil:
            k = aens_kernel_ms_history(self.h, ms, n)
        self._check(k)
        return [ms[i] for i in range(k)]

    def get_final(self, bint want_y=True, bint want_sw=True, bint want_t=True, out_t=None, out_y=None,
                  out_status=None):
        """Returns (t[N] or None, y[N,n] or None, sw[N,n_sw] or None, status[N]).  out_* may be caller-owned
        (e.g. pinned) C-contiguous arrays of the right shape/dtype that are filled instead of fresh ones."""
        cdef np.ndarray t_arr, st_arr, y_arr, sw_arr
        cdef double *tp = NULL
        cdef double *yp = NULL
        cdef unsigned char *sp = NULL
        cdef int *stp = NULL
        cdef object to = None
        cdef object yo = None
        cdef object so = None
        if want_t:
            t_arr = out_t if out_t is not None else np.empty(self.N, dtype=np.float64)
            _need(t_arr, np.float64, self.N, "out_t")
            tp = <double*>np.PyArray_DATA(t_arr)
            to = t_arr
        st_arr = out_status if out_status is not None else np.empty(self.N, dtype=np.intc)
        _need(st_arr, np.intc, self.N, "out_status")
        stp = <int*>np.PyArray_DATA(st_arr)
        if want_y:
            y_arr = out_y if out_y is not None else np.empty((self.N, self.n), dtype=np.float64)
            _need(y_arr, np.float64, self.N * self.n, "out_y")
            yp = <double*>np.PyArray_DATA(y_arr)
            yo = y_arr
        if want_sw and self.n_sw:
            sw_arr = np.empty((self.N, self.n_sw), dtype=np.uint8)
            sp = <unsigned char*>np.PyArray_DATA(sw_arr)
            so = sw_arr
        cdef int rc
        with nogil:
            rc = aens_get_final(self.h, tp, yp, sp, stp)
        self._check(rc)
        return to, yo, so, st_arr

    def get_stats(self, which):
        cdef int w = STAT_NAMES.index(which) if isinstance(which, str) else which
        cdef np.ndarray[int, ndim=1, mode="c"] out = np.empty(self.N, dtype=np.intc)
        cdef int *op = &out[0]
        cdef int rc
        with nogil:
            rc = aens_get_stats(self.h, w, op)
        self._check(rc)
        return out

    def get_stats_sum(self):
        cdef long long s[11]
        cdef int rc
        with nogil:
            rc = aens_get_stats_sum(self.h, s)
        self._check(rc)
        return {STAT_NAMES[i]: s[i] for i in range(11)}

    def get_dense(self, slot=None):
        """Row `slot` as [N,n], or all rows as [n_out,N,n] when slot is None."""
        cdef np.ndarray[double, ndim=3, mode="c"] a3
        cdef np.ndarray[double, ndim=2, mode="c"] a2
        cdef double *p
        cdef int rc, k
        if slot is None:
            a3 = np.empty((self.n_out, self.N, self.n), dtype=np.float64)
            if self.n_out == 0:
                return a3
            p = &a3[0, 0, 0]
            with nogil:
                rc = aens_get_dense_all(self.h, p)
            self._check(rc)
            return a3
        a2 = np.empty((self.N, self.n), dtype=np.float64)
        p = &a2[0, 0]
        k = slot
        with nogil:
            rc = aens_get_dense(self.h, k, p)
        self._check(rc)
        return a2

    def get_events(self):
        """Returns (count[N], t_ev[N,slots], info[N,slots]) -- ring order, see assimulo_cuda.h."""
        cdef int S = self.event_slots if ((self.n_g > 0 or (self.has_jac & 2)) and self.event_slots > 0) else 0
        cdef np.ndarray[int, ndim=1, mode="c"] cnt = np.zeros(self.N, dtype=np.intc)
        cdef np.ndarray[double, ndim=2, mode="c"] te = np.full((self.N, max(S, 1)), np.nan)
        cdef np.ndarray[int, ndim=2, mode="c"] info = np.zeros((self.N, max(S, 1)), dtype=np.intc)
        cdef int *cp = &cnt[0]
        cdef double *tp = &te[0, 0] if S > 0 else NULL
        cdef int *ip = &info[0, 0] if S > 0 else NULL
        cdef int rc
        with nogil:
            rc = aens_get_events(self.h, cp, tp, ip)
        self._check(rc)
        return cnt, te[:, :S], info[:, :S]

    def device_ptr(self, int which):
        cdef void *p = NULL
        cdef long long b = 0
        self._check(aens_device_ptr(self.h, which, &p, &b))
        return <size_t>p, b

    def stream(self):
        cdef void *s = NULL
        self._check(aens_stream(self.h, &s))
        return <size_t>s

    def launch_count(self):
        return aens_launch_count(self.h)
