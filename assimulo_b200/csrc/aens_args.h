/* Kernel argument block shared by host code, AOT kernels and NVRTC-compiled kernels (plain C layout). */
#ifndef AENS_ARGS_H
#define AENS_ARGS_H

#include "assimulo_cuda.h"

/* Device (HBM) layout is SoA throughout: component-major, member index fastest, so that the 32 lanes of a
 * warp that hold consecutive members load/store 256 contiguous bytes per component (DESIGN.md "data layout"). */
typedef struct {
    long long N;              /* members on this device */
    double *y;                /* [n][N]      in: state at t, out: state at the member's end time */
    const double *p;          /* [n_p][N]    parameters */
    unsigned char *sw;        /* [n_sw][N]   switches, in/out */
    double *t;                /* [N]         in: current time, out: end time */
    int *status;              /* [N]         AENS_ST_* */
    int *stats;               /* [AENS_NSTATS][N] */
    double *aux;              /* [1][N]      solver carry-over between simulate() calls (ExplicitEuler._inith) */
    double *ev_t;             /* [ev_slots][N] event-time ring */
    int *ev_info;             /* [ev_slots][N] event-info ring */
    double *dense;            /* [n_grid][n][N] dense-output rows on the shared grid */
    const double *t_grid;     /* [n_grid] */
    int *grid_idx;            /* [N] next grid row of each member (continuation across segments) */
    unsigned long long *queue;/* work counter for the persistent warps */
    const int *order;         /* optional fetch permutation */
    int ev_slots;
    int n_grid;
    int refill_threshold;
    int reverse;              /* no `order`: fetch members from q_end-1 downwards (cost grows with the index) */
    long long q_begin, q_end; /* member range [q_begin, q_end) this launch integrates (chunked host pipeline) */
    /* Direct host I/O (aens_simulate_host, zero-copy mode).  When in_y != NULL a member is fetched from the
     * caller's member-major arrays -- page-locked host memory mapped into the device address space, read over
     * PCIe by the warp that is about to integrate it -- instead of from the SoA state; the SoA state, the
     * parameters and the pristine copy for aens_reset are written from the fetched values.  out_* (each may be
     * NULL) receive the final values, again member-major and straight into host memory.  A parameter sweep whose
     * members share one initial state passes that row by value (y0_shared / y0_row, aens_set_shared_y0). */
    const double *in_y;       /* [N][n] */
    const double *in_p;       /* [N][n_p] */
    const unsigned char *in_sw; /* [N][n_sw] */
    double t0;                /* common start time of the fetched members */
    double *p_store;          /* [n_p][N]   SoA copy of in_p */
    double *y0_store;         /* [n][N]     pristine state */
    unsigned char *sw0_store; /* [n_sw][N] */
    double *out_y;            /* [N][n] */
    double *out_t;            /* [N] */
    unsigned char *out_sw;    /* [N][n_sw] */
    int *out_status;          /* [N] */
    double tf;
    /* Row reduction mode (aens_set_row_reduction): when red != NULL the dense-output rows are not stored but added
     * to per-warp partial records red[warp][row][5 n + 1] (aens_reduce_grid in aens_device.cuh). */
    double *red;
    /* step-size-controller constants of the FAST Dopri5 build in the log2 domain, computed once on the host
     * (aens_abi.cu, fill_ctl) so that the kernel reads them as constant-bank operands instead of holding them in
     * registers: [0] (0.2-0.75 beta)/2  [1] beta/2  [2] -lg2 safe  [3] -lg2 fac1  [4] -lg2 fac2
     * [5] beta/2 * lg2(1e-8)  (FACOLD = 1e-4 at segment start)  [6] lg2(1e-8)  [7] reserved */
    float ctl[8];
    aens_opts_t o;
    /* (appended so that the offsets the hot loops read stay where they were) */
    int zc;                   /* 1: zero-copy mode (the in_* / out_* pointers above are in use) */
    int y0_shared;            /* zero-copy mode: every member starts from y0_row instead of a row of in_y */
    double y0_row[AENS_MAX_N];
    /* time limit of this simulate() call in nanoseconds of %globaltimer, 0 = none (ODE.time_limit, src/ode.pyx:372-392;
     * checked in Explicit_ODE.report_solution, src/explicit_ode.pyx:290-292).  *t_start holds the call's start time
     * (written by the first warp that fetches work); a warp that fetches members after the deadline stores them
     * unintegrated with status AENS_ST_TIME_LIMIT. */
    unsigned long long limit_ns;
    unsigned long long *t_start; /* start time of the simulate() call, shared by all launches of a chunked call */
    /* Radau5ODE: the tolerance transform of radau5_solve (radau5.c:192-216) depends on the options only, so the host
     * evaluates it once (with glibc's pow, like the reference) and the kernel reads it from the constant bank instead
     * of carrying it in registers: rtol' = 0.1 rtol^(2/3), atol'_i = rtol' atol_i / rtol, the automatic fnewt, and
     * cfac = safe (2 newt + 1) (radau5.c:348). */
    double rd_rtol1, rd_fnewt, rd_cfac;
    double rd_atol1[AENS_MAX_N];
    /* AENS_CHECKED builds (tools/checked_build.sh): every index the kernels form into an ensemble-sized array is tested
     * against its bound; the first violation's code lands here (aens_check_failures) instead of corrupting memory */
    unsigned long long *check;
    long long red_warps;      /* number of per-warp record sets behind `red` */
} aens_args_t;

#endif
