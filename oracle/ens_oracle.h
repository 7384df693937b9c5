/*
 * oracle/ens_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar, one trajectory at a
 * time) of the reference's hot path; see ens_oracle.c for the per-function reference citations.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef ENS_ORACLE_H
#define ENS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAX_N 16
#define ORA_MAX_G 8

enum { ORA_EULER = 0, ORA_RK4 = 1, ORA_RK34 = 2, ORA_DOPRI5 = 3, ORA_RADAU5 = 4, ORA_IMPLEULER = 5, ORA_RODAS = 6 };
enum { ORA_VDP = 0, ORA_BALL = 1, ORA_ROBERTSON = 2, ORA_GYRO = 3, ORA_LINEAR = 4, ORA_LINEAR_EV = 5,
       ORA_EXTENDED = 6, ORA_TIME_EV = 7 };

/* per-lane status */
#define ORA_ST_COMPLETE 0
#define ORA_ST_TERMINATED 1
#define ORA_ERR_MAXSTEPS (-2)
#define ORA_ERR_STEP_TOO_SMALL (-3)
#define ORA_ERR_SINGULAR (-7)
#define ORA_ERR_NEWTON (-8)

/* statistics slots (names follow src/ode.pyx:153-172) */
enum { ORA_NSTEPS = 0, ORA_NFCNS, ORA_NERRFAILS, ORA_NJACS, ORA_NLUS, ORA_NSTATEFCNS, ORA_NSTATEEVENTS,
       ORA_NSTEPSTOTAL, ORA_NTIMEEVENTS, ORA_NNITERS, ORA_NNFAILS, ORA_NSTATS };

typedef struct {
    int solver;
    int maxsteps;
    int newt;      /* Radau5: max Newton iterations */
    int usejac;    /* Radau5: analytic Jacobian */
    double rtol;
    double atol[ORA_MAX_N];
    double h;      /* fixed step (Euler, RK4) */
    double inith;  /* RK34 / Radau5 / Dopri5 (0 = HINIT) */
    double maxh;   /* 0 = unset (Dopri5: inf is passed as a huge number by the caller) */
    double safe, fac1, fac2, beta;
    double thet, fnewt /* 0 = auto */, quot1, quot2;
} ora_opts_t;

void ora_default_opts(int solver, int n, ora_opts_t *o);
int  ora_model_dims(int model, int *n, int *np, int *ng, int *nsw);

/* Simulate N independent members.  All arrays are member-major ("AoS"): y0[N][n], p[N][np], sw0[N][nsw].
 * Outputs (any may be NULL): y_out[N][n], t_out[N], sw_out[N][nsw], status[N], stats[N][ORA_NSTATS],
 * dense[N][n_grid][n] (rows for t_grid[k], only filled up to each member's end time),
 * ev_t[N][max_ev], ev_info[N][max_ev] (bit 2i = component i fired, bit 2i+1 = its sign is +1, bit 30 = time
 * event).
 * Returns 0. */
int ora_simulate(int model, const ora_opts_t *opts, long N, const double *y0, const double *p,
                 const unsigned char *sw0, double t0, double tf, int n_grid, const double *t_grid,
                 double *y_out, double *t_out, unsigned char *sw_out, int *status, int *stats,
                 double *dense, int max_ev, double *ev_t, int *ev_info);

#ifdef __cplusplus
}
#endif
#endif
