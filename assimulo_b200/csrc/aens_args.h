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
    int pad0;
    double tf;
    aens_opts_t o;
} aens_args_t;

#endif
