/*
 * assimulo_cuda.h -- C ABI of libassimulo_cuda.so, the B200 (sm_100a) ensemble integrator behind Assimulo's
 * Explicit_Problem + solver.simulate() surface.
 *
 * The reference has no single native plugin ABI for this path; it has three, and this header
 * replaces all of them for the batched path.  Each entry point cites the reference interface it stands in for
 * (paths relative to the reference tree):
 *
 *   aens_create / aens_destroy          radau_setup_mem / radau_free_mem       thirdparty/radau5/radau5_io.h:7-9
 *   aens_set_model_* / aens_set_ensemble  the callbacks + y0 that Explicit_Problem carries: FP_CB_f, FP_CB_jac
 *                                        (thirdparty/radau5/radau5_impl.h:26-29), FP_event_func
 *                                        (src/ode_event_locator.h:8-9), fcn of thirdparty/hairer/dopri5.pyf:6-13
 *   aens_set_opts                       radau_set_* setters (radau5_io.h:14-34) and the WORK/IWORK packing of
 *                                        Dopri5.integrate (src/solvers/runge_kutta.py:146-161)
 *   aens_simulate                       radau5_solve (thirdparty/radau5/radau5.h:6-13), dopri5()
 *                                        (dopri5.pyf:31-50), f_event_locator (src/ode_event_locator.h:11-18) and
 *                                        the segment loop Explicit_ODE._simulate (src/explicit_ode.pyx:137-274),
 *                                        for ALL members in one launch
 *   aens_get_final / aens_get_stats     (x, y, idid, iwork[16..19]) of dopri5(); radau_get_stats (radau5_io.h:11)
 *   aens_get_dense                      contd5 (dopri5.pyf:51-61), radau_get_cont_output (radau5.h:15)
 *   aens_get_events                     event_info / t_event returned by Explicit_ODE.event_locator
 *                                        (src/explicit_ode.pyx:332-361)
 *   aens_last_error                     radau_get_err_msg (radau5_io.h:12)
 *
 * Conventions kept from the reference's C API: int return (0 ok, >0 informational, <0 error); an opaque handle
 * owns all device memory, the caller owns every host array; error text is retrievable from the handle; one handle =
 * one CUDA stream = not thread-safe.  New: no callback ever crosses the boundary -- right-hand sides, Jacobians,
 * event functions and handle_event are CUDA device functions (built-in model library or NVRTC-compiled source).
 * There is NO CPU fallback: every compute entry point fails with AENS_ERR_CUDA when no usable GPU is present.
 *
 * Host array layout: member-major ("AoS") y[N][n], p[N][n_p], sw[N][n_sw] exactly like a Python loop over
 * Explicit_Problem(rhs_i, y0_i) would hold them; the library transposes to its SoA device layout.
 */
#ifndef ASSIMULO_CUDA_H
#define ASSIMULO_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define AENS_MAX_N 16   /* state dimension limit of the register-resident kernels */
#define AENS_MAX_P 16
#define AENS_MAX_G 8
#define AENS_MAX_SW 8

/* return codes */
#define AENS_OK 0
#define AENS_ERR_ARG (-1)            /* inconsistent input (cf. RADAU_ERROR_INCONSISTENT_INPUT = -2) */
#define AENS_ERR_CUDA (-2)           /* CUDA runtime / driver failure or no GPU */
#define AENS_ERR_NVRTC (-3)          /* model source failed to compile; text in aens_last_error */
#define AENS_ERR_STATE (-4)          /* call order violated (e.g. simulate before set_ensemble) */
#define AENS_ERR_UNSUPPORTED (-5)    /* model/solver combination not available */

/* solver ids (the five Explicit_ODE subclasses of BASELINE.json + ImplicitEuler and RodasODE, SURVEY 8f-2,3) */
#define AENS_SOLVER_EULER 0          /* src/solvers/euler.pyx:485 ExplicitEuler */
#define AENS_SOLVER_RK4 1            /* src/solvers/runge_kutta.py:746 RungeKutta4 */
#define AENS_SOLVER_RK34 2           /* src/solvers/runge_kutta.py:430 RungeKutta34 */
#define AENS_SOLVER_DOPRI5 3         /* src/solvers/runge_kutta.py:26 Dopri5 */
#define AENS_SOLVER_RADAU5 4         /* src/solvers/radau5.py:69 Radau5ODE */
#define AENS_SOLVER_IMPLEULER 5      /* src/solvers/euler.pyx:29 ImplicitEuler */
#define AENS_SOLVER_RODAS 6          /* src/solvers/rosenbrock.py:259 RodasODE */
#define AENS_NSOLVERS 7

/* per-member status written by aens_simulate (aens_get_final) */
#define AENS_ST_COMPLETE 0           /* reached tfinal (ID_PY_COMPLETE) */
#define AENS_ST_TERMINATED 1         /* handle_event asked to stop (TerminateSimulation) */
#define AENS_ST_MAXSTEPS (-2)        /* DOPRI5 IDID=-2 / RK34 'Final time not reached' */
#define AENS_ST_STEP_TOO_SMALL (-3)  /* DOPRI5 IDID=-3 */
#define AENS_ST_RODAS_SINGULAR (-4)  /* RODAS IDID=-4: matrix repeatedly singular */
#define AENS_ST_RADAU_NMAX (-5)      /* RADAU_ERROR_NMAX_TOO_SMALL */
#define AENS_ST_RADAU_STEP_TOO_SMALL (-6)
#define AENS_ST_RADAU_SINGULAR (-7)   /* also: singular Newton matrix in ImplicitEuler (numpy LinAlgError) */
#define AENS_ST_NEWTON_FAILED (-8)    /* ImplicitEuler: 'Newton iteration failed' (euler.pyx:415,426) */
#define AENS_ST_TIME_LIMIT (-9)       /* aens_set_time_limit: the member was not started before the deadline
                                       * (TimeLimitExceeded, src/explicit_ode.pyx:290-292) */

/* statistics slots for aens_get_stats (names: src/ode.pyx:153-172) */
#define AENS_STAT_NSTEPS 0           /* accepted steps */
#define AENS_STAT_NFCNS 1
#define AENS_STAT_NERRFAILS 2
#define AENS_STAT_NJACS 3
#define AENS_STAT_NLUS 4
#define AENS_STAT_NSTATEFCNS 5
#define AENS_STAT_NSTATEEVENTS 6
#define AENS_STAT_NSTEPSTOTAL 7      /* step attempts (DOPRI5 NSTEP, Radau nsteps) */
#define AENS_STAT_NTIMEEVENTS 8      /* time events (src/explicit_ode.pyx:241-242) */
#define AENS_STAT_NNITERS 9          /* nonlinear iterations (ImplicitEuler) */
#define AENS_STAT_NNFAILS 10         /* nonlinear convergence failures (ImplicitEuler) */
#define AENS_NSTATS 11

/* arithmetic mode */
#define AENS_ARITH_STRICT 0  /* reference operation order, no FMA contraction, fp64 pow/sqrt/div everywhere */
#define AENS_ARITH_FAST 1    /* FMA contraction; step-size controller powers via fp32 MUFU; same accept test */

typedef struct {
    int solver;          /* AENS_SOLVER_* */
    int arith;           /* AENS_ARITH_* */
    int maxsteps;        /* Dopri5/Radau5ODE: 100000, RungeKutta34: 10000 */
    int newt;            /* Radau5ODE, ImplicitEuler: max Newton iterations (7) */
    int usejac;          /* Radau5ODE, ImplicitEuler: 1 = model's analytic Jacobian, 0 = finite differences */
    int reserved;
    double rtol;         /* scalar, broadcast to n (runge_kutta.py:172, radau5.py:374) */
    double atol[AENS_MAX_N];
    double h;            /* ExplicitEuler / RungeKutta4 fixed step (0.01) */
    double inith;        /* Dopri5: 0 => HINIT; RungeKutta34, Radau5ODE, RodasODE: 0.01 */
    double maxh;         /* 0 => unset (Dopri5: inf; Radau5ODE: tfinal - t) */
    double safe, fac1, fac2, beta;
    double thet, fnewt /* 0 => automatic */, quot1, quot2;
} aens_opts_t;

typedef struct aens_handle_s aens_handle_t;

/* fills *o with the reference's defaults for `solver` (runge_kutta.py:56-64,449-452,784; euler.pyx:522;
 * radau5.py:99-113) */
int aens_default_opts(int solver, aens_opts_t *o);

int aens_create(int device, aens_handle_t **h);
int aens_destroy(aens_handle_t **h);
const char *aens_last_error(const aens_handle_t *h);
const char *aens_version(void);
/* AENS_NSTATS of the library that is actually loaded: the length every `sums` / statistics array must have (for
 * callers that cannot include this header, e.g. ctypes) */
int aens_nstats(void);

/* Models.  Built-ins: "vdp" "ball" "robertson" "gyro" "linear" "linear_ev" "extended" "time_ev".
 * aens_set_model_source compiles `cuda_src` with NVRTC for sm_100a.  The source must define
 *   __device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd);
 * and, when the corresponding flag/count says so,
 *   __device__ void aens_jac(double t, const double* y, const unsigned char* sw, const double* p, double* J_colmajor);
 *   __device__ void aens_events(double t, const double* y, const unsigned char* sw, const double* p, double* g);
 *   __device__ int  aens_handle_event(double t, double* y, unsigned char* sw, const int* event_info, const double* p);
 * (handle_event returns non-zero to terminate the member), and, when `has_jac` carries the flag
 * AENS_MODEL_TIME_EVENTS, the time-event pair of Explicit_Problem.time_events / handle_event with
 * event_info = [[], True] (src/explicit_ode.pyx:212-216,239-259):
 *   __device__ double aens_time_events(double t, const double* y, const unsigned char* sw, const double* p);
 *                     next time event strictly after t, or AENS_NO_TIME_EVENT (the reference's None)
 *   __device__ int    aens_handle_time_event(double t, double* y, unsigned char* sw, const double* p);
 * Optional, flag AENS_MODEL_RHS_H: the fused form hk = h * f(t, y) the explicit Runge-Kutta stages of the fast
 * arithmetic build consume (a model whose f is a product with a parameter folds h into it once per step); the strict
 * build never calls it:
 *   __device__ void aens_rhs_h(double h, double t, const double* y, const unsigned char* sw, const double* p, double* hk); */
#define AENS_MODEL_HAS_JAC 1
#define AENS_MODEL_TIME_EVENTS 2
#define AENS_MODEL_RHS_H 4
#define AENS_NO_TIME_EVENT (1.0e300 * 1.0e300) /* +inf */
int aens_set_model_builtin(aens_handle_t *h, const char *name);
/* has_jac: 0/1 as before, or a bit set of AENS_MODEL_* flags */
int aens_set_model_source(aens_handle_t *h, const char *cuda_src, int n, int n_p, int n_g, int n_sw, int has_jac);
int aens_model_dims(const aens_handle_t *h, int *n, int *n_p, int *n_g, int *n_sw, int *has_jac);

/* Ensemble of N members: y0[N][n], p[N][n_p] (may be NULL if n_p==0), sw0[N][n_sw] (NULL if n_sw==0), common t0.
 * Resets every member to (t0, y0, sw0) and clears statistics and event logs. */
int aens_set_ensemble(aens_handle_t *h, long long N, const double *y0, const double *p,
                      const unsigned char *sw0, double t0);
/* Back to (t0, y0, sw0) of the last aens_set_ensemble, device to device, asynchronous on the handle's stream
 * (Explicit_ODE.reset(), src/explicit_ode.pyx:98-106). */
int aens_reset(aens_handle_t *h);
int aens_set_opts(aens_handle_t *h, const aens_opts_t *o);

/* Output: n_out grid times shared by all members (ODE.simulate's output_list, src/ode.pyx:273-287);
 * n_out = 0 => final state only.  event_slots = per-member ring size of the event log. */
int aens_set_output(aens_handle_t *h, int n_out, const double *t_out, int event_slots);

/* Scheduling of members onto persistent warps: a warp fetches new members when >= refill_threshold of its 32
 * lanes are idle (32 = warp-granular chunks, 1 = refill every finished lane at once); `order` (may be NULL) is a
 * permutation of 0..N-1 giving the fetch order (e.g. most expensive first). */
int aens_set_schedule(aens_handle_t *h, int refill_threshold, const int *order);

/* Lanes per member.  The kernels integrate one member per thread; models with a larger state that define the
 * cooperative form of their right-hand side (built-in: "gyro", 4 lanes x 3 components) also have a
 * sub-warp-per-member Dopri5 kernel that keeps a quarter of the state in each lane's registers (csrc/aens_split.cuh).
 * lanes = 0 (default): the library picks -- one thread per member, which measured faster on the gyroscope ensemble
 * (profiles/r2_split.md); 1: always one thread per member; > 1: that many lanes (FAST arithmetic, no row reduction,
 * not the one-call host modes) or AENS_ERR_UNSUPPORTED.  aens_last_lanes_per_member reports what the most recent
 * launch used. */
int aens_set_lanes_per_member(aens_handle_t *h, int lanes);
int aens_last_lanes_per_member(const aens_handle_t *h);

/* No fetch permutation: hand members out from the highest index downwards (reverse = 1) instead of upwards --
 * for ensembles whose cost grows with the member index (most expensive first, no permutation array to upload) -- or
 * in 32-member batches alternately from the top and the bottom of the range (reverse = 2): for such ensembles in
 * zero-copy host mode, where it evens out the rate at which inputs are pulled over PCIe. */
int aens_set_fetch_reverse(aens_handle_t *h, int reverse);

/* Backward integration (the reference's options["backward"], src/ode.pyx:490-507, src/explicit_ode.pyx:185,209,312):
 * with on != 0 aens_simulate integrates every member from its current time DOWN to tfinal (tfinal <= t) and the output
 * grid is descending.  Available with the three solvers whose reference cores carry the direction of integration --
 * Dopri5, Radau5ODE, RodasODE (POSNEG: dopri5.f:380, radau5.c:341, rodas_decsol.f:569) --, state events included; not
 * with time events, row reduction or aens_simulate_host.  The kernels integrate the time-reversed problem forward
 * (csrc/aens_models.cuh AensRev), which reproduces the reference's backward run value for value. */
int aens_set_backward(aens_handle_t *h, int on);

/* Integrates every member from its current time to tfinal (continuation semantics of ODE.simulate,
 * src/ode.pyx:195-196,222).  Synchronous.  Returns AENS_OK even if individual members failed: see status[]. */
int aens_simulate(aens_handle_t *h, double tfinal);
/* Same, asynchronous on the handle's stream; pair with aens_sync. */
int aens_simulate_async(aens_handle_t *h, double tfinal);
int aens_sync(aens_handle_t *h);
/* device time of the last aens_simulate[_async] launch in milliseconds (CUDA events on the handle's stream) */
int aens_last_kernel_ms(aens_handle_t *h, float *ms);
/* Durations of the most recent integration-kernel launches, newest first (each launch records its own CUDA event
 * pair on the launching stream; the ring keeps 64): lets a caller that queued launches back to back with
 * aens_simulate_async read every kernel's time afterwards.  Synchronises the stream; returns how many were written. */
int aens_kernel_ms_history(aens_handle_t *h, float *ms, int n);

/* aens_set_ensemble + aens_simulate + aens_get_final in ONE call for inputs and results in host memory -- what a
 * Python loop `for i: Solver(Explicit_Problem(rhs_i, y0_i)).simulate(tfinal)` does, end to end.  The members are
 * cut into `nchunks` contiguous ranges; host->device copies, kernels and device->host copies of consecutive
 * ranges overlap on separate CUDA streams.  Outputs (any may be NULL): y_out[N][n], t_out[N], sw_out[N][n_sw],
 * status_out[N]; sums[AENS_NSTATS] = statistic totals; *n_not_complete = members whose status != AENS_ST_COMPLETE.
 * Uses the options, event-slot count and fetch direction set on the handle; needs n_out == 0 and no fetch
 * permutation.  Host arrays should be page-locked (aens_host_alloc) for the copies to overlap. */
int aens_simulate_host(aens_handle_t *h, long long N, const double *y0, const double *p, const unsigned char *sw0,
                       double t0, double tfinal, double *y_out, double *t_out, unsigned char *sw_out,
                       int *status_out, int nchunks, long long *sums, long long *n_not_complete);
/* How aens_simulate_host moves data: 0 = automatic (zero-copy when every host array is page-locked, mapped and
 * 16-byte aligned -- ONE kernel launch whose warps read their members' rows from host memory over PCIe when they
 * fetch them and write the final rows back when they finish; otherwise the chunked staging pipeline),
 * 1 = always the chunked staging pipeline (copy engine both ways), 2 = zero-copy or fail,
 * 3 = chunked, inputs zero-copy, results by copy engine, 4 = chunked, inputs by copy engine, results zero-copy
 * (which side of the host link the SMs or the copy engines drive better depends on the machine and on how many GPUs
 * share it: profiles/r2_pcie.md). */
int aens_set_host_mode(aens_handle_t *h, int mode);
/* Parameter sweeps whose members share one initial state (every Explicit_Problem of the loop built with the same
 * y0): with on != 0 the y0 argument of aens_set_ensemble / aens_simulate_host is ONE row y0[n] instead of
 * y0[N][n]; in zero-copy mode the row travels in the kernel's argument block, so only p (and sw0) cross PCIe. */
int aens_set_shared_y0(aens_handle_t *h, int on);
/* statistic totals and the number of members that did not complete, after aens_simulate */
int aens_get_summary(aens_handle_t *h, long long *sums /*[AENS_NSTATS]*/, long long *n_not_complete);

/* Results (any pointer may be NULL): t[N], y[N][n], sw[N][n_sw], status[N]. */
int aens_get_final(aens_handle_t *h, double *t, double *y, unsigned char *sw, int *status);
int aens_get_stats(aens_handle_t *h, int which, int *out /*[N]*/);
int aens_get_stats_sum(aens_handle_t *h, long long *sums /*[AENS_NSTATS]*/);
/* Ensemble statistics of the final states computed on the device (nothing ensemble-sized crosses PCIe): per state
 * component, over the members whose status is AENS_ST_COMPLETE: mean[n], var[n] (population variance), min[n],
 * max[n], count[n].  Any pointer may be NULL. */
int aens_reduce_final(aens_handle_t *h, double *mean, double *var, double *min, double *max, long long *count);
/* The same reduction over the members' END TIMES, restricted to the members whose status equals `status`
 * (scalars: mean, var, min, max, count).  With status = AENS_ST_TERMINATED these are first-passage-time statistics:
 * a model whose handle_event returns "terminate" (the reference's TerminateSimulation, src/explicit_ode.pyx:259-262)
 * stops each member at its first event, and the end time of a terminated member is the located event time. */
int aens_reduce_end_times(aens_handle_t *h, int status, double *mean, double *var, double *min, double *max,
                          long long *count);
/* Row reduction mode -- the ensemble-level counterpart of handle_result() being called with every dense-output row
 * (src/explicit_ode.pyx:308-318, src/problem.pyx:145-158): with on != 0 the rows of the output grid are not stored
 * per member ([n_out][n][N] in HBM, N x n_out x n doubles to download) but reduced over the ensemble inside the
 * integration kernel.  Call before aens_set_output; aens_get_dense then fails.  After aens_simulate,
 * aens_get_row_reduction returns, per grid row and state component, over the members that reached that row:
 * mean[n_out][n], var[n_out][n] (population variance), min[n_out][n], max[n_out][n] and count[n_out]
 * (any pointer may be NULL; rows no member reached hold NaN and count 0).  Sums are accumulated in the order the
 * warps finish rows, so mean and var are reproducible to rounding only; min, max and count are exact. */
int aens_set_row_reduction(aens_handle_t *h, int on);
int aens_get_row_reduction(aens_handle_t *h, double *mean, double *var, double *min, double *max, long long *count);
/* dense output row `slot` (0..n_out-1) as y_out[N][n]; rows beyond a member's end time are NaN */
int aens_get_dense(aens_handle_t *h, int slot, double *y_out);
/* whole buffer, y_out[n_out][N][n] */
int aens_get_dense_all(aens_handle_t *h, double *y_out);
/* event log: count[N] = events seen; t_ev[N][event_slots], info[N][event_slots] hold the last
 * min(count,event_slots) events of each member in ring order (event e sits in slot e % event_slots);
 * info bit 2i = indicator i fired, bit 2i+1 = its event_info is +1 (else -1); bit 30 = time event.  count
 * includes time events. */
int aens_get_events(aens_handle_t *h, int *count, double *t_ev, int *info);

/* Raw device pointers (SoA, for zero-copy consumers such as an NCCL all-gather of final states):
 * which = 0: t[N]  1: y[n][N]  2: status[N] (int)  3: stats[AENS_NSTATS][N] (int)  4: ev_t[slots][N]
 *         5: ev_info[slots][N] (int)  6: dense[n_out][n][N]  7: p[n_p][N]  8: sw[n_sw][N] (uint8) */
int aens_device_ptr(aens_handle_t *h, int which, void **ptr, long long *bytes);
/* CUDA stream of the handle (cudaStream_t as void*) */
int aens_stream(aens_handle_t *h, void **stream);

/* Page-locked host memory, mapped into the device address space, for the caller's arrays: host<->device copies of
 * pinned buffers run at full PCIe rate and asynchronously, and aens_simulate_host can read/write them directly from
 * the kernel (the reference has no analogue; its arrays never leave the host). */
int aens_host_alloc(long long bytes, void **ptr);
int aens_host_free(void *ptr);

/* Time limit of one aens_simulate / aens_simulate_host call in seconds of device time, 0 = none (ODE.time_limit,
 * src/ode.pyx:372-392, checked per step in Explicit_ODE.report_solution, src/explicit_ode.pyx:290-292, where it raises
 * TimeLimitExceeded).  One launch integrates the whole ensemble, so the limit is checked where the device code has a
 * natural seam: when a warp asks for its next members.  Members fetched after the deadline are not integrated; they
 * keep their state and get status AENS_ST_TIME_LIMIT (members already in flight run to completion). */
int aens_set_time_limit(aens_handle_t *h, double seconds);

/* Multi-GPU: the ensemble shards by contiguous member blocks, one handle (process) per GPU, and the path's ONLY
 * collective is one all-gather of the final results over NCCL (NVLink 5 / NVSwitch) at the end of simulate()
 * (handle convention of thirdparty/radau5/radau5.h:6-13: opaque handle, int return, error text on the
 * handle).  Rank 0 obtains a 128-byte id with aens_comm_unique_id and hands it to the other ranks by any host-side
 * means (file, socket, MPI, torch.distributed ...); every rank then calls aens_comm_init on its handle.
 * aens_allgather_results packs the selected SoA result buffers of this rank's block into one record and issues ONE
 * ncclAllGather on the handle's stream (asynchronous; aens_sync waits); every rank must hold the same number of
 * members (pad the last block) and select the same fields.  what = 0 selects everything the model has.
 * aens_get_gathered unpacks the gathered records on the device and returns the first n_total members in rank order as
 * host arrays (any pointer may be NULL): t[n_total], y[n_total][n], status[n_total], stats[AENS_NSTATS][n_total],
 * ev_t[n_total][event_slots], ev_info[n_total][event_slots].  aens_gathered_ptr exposes the raw gathered buffer
 * ([nranks] records of rec_bytes each) to device-side consumers. */
#define AENS_COMM_ID_BYTES 128
#define AENS_GATHER_T 1
#define AENS_GATHER_Y 2
#define AENS_GATHER_STATUS 4
#define AENS_GATHER_STATS 8
#define AENS_GATHER_EVENTS 16
int aens_comm_unique_id(void *id128);
int aens_comm_init(aens_handle_t *h, int rank, int nranks, const void *id128);
int aens_comm_destroy(aens_handle_t *h);
int aens_allgather_results(aens_handle_t *h, int what);
/* what != 0: every aens_simulate / aens_simulate_async / aens_simulate_host on a handle with a communicator ends with
 * aens_allgather_results(h, what), issued by the library at the earliest point -- right behind the last integration
 * kernel.  In the chunked host modes the exchange then runs over NVLink while the copy engines are still moving the
 * last chunks' results to the host over PCIe (the two do not share a link).  0 switches it off (default). */
int aens_set_gather_in_simulate(aens_handle_t *h, int what);
int aens_last_gather_ms(aens_handle_t *h, float *ms);
int aens_gathered_ptr(aens_handle_t *h, void **ptr, long long *rec_bytes, int *nranks);
int aens_get_gathered(aens_handle_t *h, long long n_total, double *t, double *y, int *status, int *stats,
                      double *ev_t, int *ev_info);

/* Measurement helpers (not on the product path): sustained fp64 FMA throughput of the device measured with a
 * dependent-chain DFMA kernel for ~`seconds`; result in TFLOP/s (2 flops per DFMA). */
int aens_measure_fp64_peak(int device, double seconds, double *tflops, double *sm_clock_mhz_est);
/* What the host link of `device` sustains with `bytes`-sized page-locked buffers for ~`seconds`, in GB/s:
 * mode 0 copy engine host->device, 1 copy engine device->host, 2 both at once (sum of both directions),
 * 3 SM-initiated reads of mapped host memory (what the zero-copy host mode does), 4 SM-initiated writes, 5 both. */
int aens_measure_hostlink(int device, long long bytes, int mode, double seconds, double *gbs);
/* compile-only check of a model source with NVRTC for sm_100a (needs no GPU); returns the cubin size in bytes
 * (> 0) or a negative AENS_ERR_* with the compiler log in errbuf */
long long aens_nvrtc_check(const char *cuda_src, int n, int n_p, int n_g, int n_sw, int has_jac, int strict,
                           char *errbuf, int errbuf_len);
/* Debug aid: a library built with -DAENS_CHECKED (python assimulo_b200/build.py --variant checked -D AENS_CHECKED=1)
 * tests every index its kernels form into an ensemble-sized array; *code = the first violation recorded by any kernel
 * of this handle, 0 = none (always 0 with the shipped library, which carries no checks). */
int aens_check_failures(aens_handle_t *h, long long *code);
/* number of kernel launches issued by this handle since creation */
long long aens_launch_count(const aens_handle_t *h);

#ifdef __cplusplus
}
#endif
#endif /* ASSIMULO_CUDA_H */
