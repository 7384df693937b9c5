/*
 * Built-in model library: right-hand sides, Jacobians, event functions and handle_event as CUDA device
 * functions (a Python callable cannot run on the device; this replaces Explicit_Problem.rhs/jac/state_events/
 * handle_event, src/problem.pyx:403-488).  Each struct is the device twin of a model the reference ships:
 *
 *   Vdp        examples/radau5ode_vanderpol.py:42-48       (mu as parameter p[0])
 *   Ball       examples/lsodar_bouncing_ball.py:28-75      p = [g, restitution]
 *   Robertson  examples/radau5ode_with_jac_sparse.py:46-62 p = [k1, k2, k3]
 *   Gyro       examples/cvode_gyro.py:39-57 (12 states)    p = [I1, I2, I3]
 *   Linear     y' = p0*y + p1 (tests/solvers/test_rungekutta.py:40-47, examples/dopri5_basic.py)
 *   LinearEv   + g = y - p2, handle_event: sw[0] <- False  (tests/test_solvers.py:27-57)
 *   Extended   tests/conftest.py:26-114 Extended_Problem (event iteration inside handle_event)
 *   TimeEv     tests/solvers/test_rungekutta.py:49-82 (time events + handle_event with event_info[1] = True)
 *
 * Interface of a model M (all static, all __device__):
 *   constexpr int N, NP, NG, NSW; constexpr bool HAS_JAC;
 *   void rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd);
 *   void jac(t, y, sw, p, double* J)            column-major J[i + j*N] = d f_i / d y_j
 *   void events(t, y, sw, p, double* g);
 *   int  handle_event(double t, double* y, unsigned char* sw, const int* event_info, const double* p);
 *   void prep(double* p)   called once per member on its private copy of the parameters, FAST build only: a model
 *                          may replace a divisor by its reciprocal so that rhs multiplies (AensGyro)
 * The operation order inside each rhs is the order of the Python model, so that the STRICT build reproduces the
 * reference's floating-point results bit for bit.
 */
#ifndef AENS_MODELS_CUH
#define AENS_MODELS_CUH

#define AENS_DEV __device__ __forceinline__

/* what a model inherits unless it says otherwise: no parameter preparation, no time events */
struct AensModelDefaults {
    static constexpr bool HAS_TE = false;
    /* optional fused "step times right-hand side": rhs_h(h, t, y, sw, p, hk) returns hk = h * f(t, y).  The explicit
     * Runge-Kutta stages of the FAST Dopri5 build only ever need h * k_j; a model whose f is a product with a
     * parameter can fold h into that factor once per step (Van der Pol: (h mu) * w instead of h * (mu * w)). */
    static constexpr bool HAS_RHS_H = false;
    /* sub-warp-per-member variant (aens_split.cuh): G lanes share one member, each owning NL of its components;
     * comp(j, c) = state index of component c of lane j, rhs_g() = this lane's part of f, fetching what it needs from
     * the other lanes of its group with warp shuffles.  G = 1: the model has no such form. */
    static constexpr int G = 1, NL = 0;
    /* true only for AensRev<M>: the kernel's time axis is s = -t (backward integration) */
    static constexpr bool REV = false;
    static AENS_DEV void prep(double *) {}
    static AENS_DEV double time_events(double, const double *, const unsigned char *, const double *) {
        return AENS_NO_TIME_EVENT;
    }
    static AENS_DEV int handle_time_event(double, double *, unsigned char *, const double *) { return 0; }
};

struct AensVdp : AensModelDefaults {
    static constexpr int N = 2, NP = 1, NG = 0, NSW = 0;
    static constexpr bool HAS_JAC = true;
    static AENS_DEV void rhs(double, const double *y, const unsigned char *, const double *p, double *yd) {
        yd[0] = y[1];
        yd[1] = p[0] * ((1. - y[0] * y[0]) * y[1] - y[0]);
    }
    static constexpr bool HAS_RHS_H = true;
    static AENS_DEV void rhs_h(double h, double, const double *y, const unsigned char *, const double *p, double *hk) {
        const double hmu = h * p[0]; /* common to the stages of one attempt */
        hk[0] = h * y[1];
        hk[1] = hmu * ((1. - y[0] * y[0]) * y[1] - y[0]);
    }
    static AENS_DEV void jac(double, const double *y, const unsigned char *, const double *p, double *J) {
        J[0] = 0.;
        J[1] = p[0] * (-2. * y[0] * y[1] - 1.);
        J[2] = 1.;
        J[3] = p[0] * (1. - y[0] * y[0]);
    }
    static AENS_DEV void events(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV int handle_event(double, double *, unsigned char *, const int *, const double *) { return 0; }
};

struct AensBall : AensModelDefaults {
    static constexpr int N = 2, NP = 2, NG = 2, NSW = 2;
    static constexpr bool HAS_JAC = false;
    static AENS_DEV void rhs(double, const double *y, const unsigned char *, const double *p, double *yd) {
        yd[0] = y[1];
        yd[1] = p[0];
    }
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *, double *J) {
        J[0] = 0.; J[1] = 0.; J[2] = 1.; J[3] = 0.;
    }
    static AENS_DEV void events(double, const double *y, const unsigned char *sw, const double *, double *g) {
        g[0] = sw[0] ? y[0] : 5.;
        g[1] = sw[1] ? y[1] : 5.;
    }
    static AENS_DEV int handle_event(double, double *y, unsigned char *sw, const int *info, const double *p) {
        if (info[0] != 0) { sw[0] = 0; sw[1] = 1; y[1] = -p[1] * y[1]; }
        else              { sw[0] = 1; sw[1] = 0; }
        return 0;
    }
};

struct AensRobertson : AensModelDefaults {
    static constexpr int N = 3, NP = 3, NG = 0, NSW = 0;
    static constexpr bool HAS_JAC = true;
    static AENS_DEV void rhs(double, const double *y, const unsigned char *, const double *p, double *yd) {
        double yd0 = -p[0] * y[0] + p[1] * y[1] * y[2];
        double yd2 = p[2] * y[1] * y[1];
        yd[0] = yd0;
        yd[1] = -yd0 - yd2;
        yd[2] = yd2;
    }
    static AENS_DEV void jac(double, const double *y, const unsigned char *, const double *p, double *J) {
        J[0] = -p[0];        J[1] = p[0];                                J[2] = 0.;
        J[3] = p[1] * y[2];  J[4] = -p[1] * y[2] - 2. * p[2] * y[1];     J[5] = 2. * p[2] * y[1];
        J[6] = p[1] * y[1];  J[7] = -p[1] * y[1];                        J[8] = 0.;
    }
    static AENS_DEV void events(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV int handle_event(double, double *, unsigned char *, const int *, const double *) { return 0; }
};

struct AensGyro : AensModelDefaults {
    static constexpr int N = 12, NP = 3, NG = 0, NSW = 0;
    static constexpr bool HAS_JAC = false;
#if !(defined(AENS_STRICT) && AENS_STRICT)
    static AENS_DEV void prep(double *p) { p[0] = 1.0 / p[0]; p[1] = 1.0 / p[1]; p[2] = 1.0 / p[2]; }
#endif
    static AENS_DEV void rhs(double, const double *u, const unsigned char *, const double *p, double *ud) {
#if defined(AENS_STRICT) && AENS_STRICT
        const double om0 = u[0] / p[0], om1 = u[1] / p[1], om2 = u[2] / p[2]; /* omega = pi / I as in the reference */
#else
        const double om0 = u[0] * p[0], om1 = u[1] * p[1], om2 = u[2] * p[2]; /* p = 1/I after prep() */
#endif
        ud[0] = om2 * u[1] + (-om1) * u[2];
        ud[1] = (-om2) * u[0] + om0 * u[2];
        ud[2] = om1 * u[0] + (-om0) * u[1];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double q0 = u[3 + c], q1 = u[6 + c], q2 = u[9 + c];
            ud[3 + c] = om2 * q1 + (-om1) * q2;
            ud[6 + c] = (-om2) * q0 + om0 * q2;
            ud[9 + c] = om1 * q0 + (-om0) * q1;
        }
    }
#if !(defined(AENS_STRICT) && AENS_STRICT)
    /* Four lanes per member: lane 0 owns the body angular momentum pi = u[0..2], lane 1 + c owns column c of the
     * rotation matrix, (u[3+c], u[6+c], u[9+c]).  Every lane rotates its own 3-vector v: v' = curl(omega) v with
     * omega = pi / I fetched from lane 0 -- the same instructions in all four lanes (p = 1/I after prep()). */
    static constexpr int G = 4, NL = 3;
    static AENS_DEV int comp(int j, int c) { return j == 0 ? c : 3 + 3 * c + (j - 1); }
    static AENS_DEV void rhs_g(unsigned gm, int, double, const double *v, const double *p, double *vd) {
        const double om0 = __shfl_sync(gm, v[0], 0, 4) * p[0], om1 = __shfl_sync(gm, v[1], 0, 4) * p[1],
                     om2 = __shfl_sync(gm, v[2], 0, 4) * p[2];
        vd[0] = om2 * v[1] + (-om1) * v[2];
        vd[1] = (-om2) * v[0] + om0 * v[2];
        vd[2] = om1 * v[0] + (-om0) * v[1];
    }
#endif
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV void events(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV int handle_event(double, double *, unsigned char *, const int *, const double *) { return 0; }
};

struct AensLinear : AensModelDefaults {
    static constexpr int N = 1, NP = 3, NG = 0, NSW = 0;
    static constexpr bool HAS_JAC = true;
    static AENS_DEV void rhs(double, const double *y, const unsigned char *, const double *p, double *yd) {
        yd[0] = p[0] * y[0] + p[1];
    }
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *p, double *J) { J[0] = p[0]; }
    static AENS_DEV void events(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV int handle_event(double, double *, unsigned char *, const int *, const double *) { return 0; }
};

struct AensLinearEv : AensModelDefaults {
    static constexpr int N = 1, NP = 3, NG = 1, NSW = 1;
    static constexpr bool HAS_JAC = true;
    static AENS_DEV void rhs(double, const double *y, const unsigned char *, const double *p, double *yd) {
        yd[0] = p[0] * y[0] + p[1];
    }
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *p, double *J) { J[0] = p[0]; }
    static AENS_DEV void events(double, const double *y, const unsigned char *, const double *p, double *g) {
        g[0] = y[0] - p[2];
    }
    static AENS_DEV int handle_event(double, double *, unsigned char *sw, const int *, const double *) {
        sw[0] = 0;
        return 0;
    }
};

struct AensExtended : AensModelDefaults {
    static constexpr int N = 3, NP = 0, NG = 3, NSW = 3;
    static constexpr bool HAS_JAC = false;
    static AENS_DEV void rhs(double, const double *, const unsigned char *sw, const double *, double *yd) {
        yd[0] = sw[0] ? 1.0 : -1.0;
        yd[1] = 0.0;
        yd[2] = 0.0;
    }
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV void events(double t, const double *y, const unsigned char *, const double *, double *g) {
        g[0] = y[1] - 1.0;
        g[1] = -y[2] + 1.0;
        g[2] = -t + 1.0;
    }
    static AENS_DEV int handle_event(double t, double *y, unsigned char *sw, const int *info_in, const double *p) {
        int info[3] = {info_in[0], info_in[1], info_in[2]};
        for (int iter = 0; iter < 100; ++iter) { /* event iteration, conftest.py:64-80 */
            double b[3], a[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) if (info[i] != 0) sw[i] = !sw[i];
            events(t, y, sw, p, b);
            y[1] = sw[1] ? -1.0 : 3.0;
            y[2] = sw[2] ? 0.0 : 2.0;
            events(t, y, sw, p, a);
            int any = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                info[i] = ((b[i] < 0.0 && a[i] > 0.0) || (b[i] > 0.0 && a[i] < 0.0)) ? 1 : 0;
                any |= info[i];
            }
            if (!any) break;
        }
        return 0;
    }
};

/* tests/solvers/test_rungekutta.py:49-82 test_time_event: y' = p5, time events at p0 < p1 < p2 < p3, handle_event:
 * y += p4 */
struct AensTimeEv : AensModelDefaults {
    static constexpr int N = 1, NP = 6, NG = 0, NSW = 0;
    static constexpr bool HAS_JAC = true;
    static constexpr bool HAS_TE = true;
    static AENS_DEV void rhs(double, const double *, const unsigned char *, const double *p, double *yd) { yd[0] = p[5]; }
    static AENS_DEV void jac(double, const double *, const unsigned char *, const double *, double *J) { J[0] = 0.; }
    static AENS_DEV void events(double, const double *, const unsigned char *, const double *, double *) {}
    static AENS_DEV int handle_event(double, double *, unsigned char *, const int *, const double *) { return 0; }
    static AENS_DEV double time_events(double t, const double *, const unsigned char *, const double *p) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t < p[k]) return p[k];
        return AENS_NO_TIME_EVENT;
    }
    static AENS_DEV int handle_time_event(double, double *y, unsigned char *, const double *p) {
        y[0] += p[4];
        return 0;
    }
};

/* Backward integration (options["backward"], src/explicit_ode.pyx:185,209,312).  The reference's cores integrate
 * towards a smaller tfinal by carrying the sign of the direction through the step size (POSNEG: dopri5.f:380,
 * radau5.c:341, rodas_decsol.f:569).  Here the same integration runs FORWARD in s = -t on the time-reversed problem
 *     dy/ds = -f(-s, y),   g(-s, y),   J_s = -J
 * with the unchanged steppers.  IEEE arithmetic is symmetric under a sign change of its operands, so every intermediate
 * of the reversed run is the exact mirror image of the reference's (x + c h -> -(x + c h), h k -> h k), and the states,
 * step counts and event times come out identical; only RODAS's df/dt forward difference looks at the other side of x.
 * The host hands the kernel -tfinal and the negated output grid; Lane::load / store and the event log flip the sign of
 * the times they move between the member's registers and memory (M::REV). */
template <class M>
struct AensRev : M {
    static constexpr bool REV = true;
    static constexpr bool HAS_RHS_H = false;
    static constexpr int G = 1;
    static AENS_DEV void rhs(double s, const double *y, const unsigned char *sw, const double *p, double *yd) {
        M::rhs(-s, y, sw, p, yd);
#pragma unroll
        for (int i = 0; i < M::N; ++i) yd[i] = -yd[i];
    }
    static AENS_DEV void jac(double s, const double *y, const unsigned char *sw, const double *p, double *J) {
        M::jac(-s, y, sw, p, J);
#pragma unroll(M::N <= 4 ? 64 : 1)
        for (int i = 0; i < M::N * M::N; ++i) J[i] = -J[i];
    }
    static AENS_DEV void events(double s, const double *y, const unsigned char *sw, const double *p, double *g) {
        M::events(-s, y, sw, p, g);
    }
    static AENS_DEV int handle_event(double s, double *y, unsigned char *sw, const int *info, const double *p) {
        return M::handle_event(-s, y, sw, info, p);
    }
};

#ifdef AENS_USER_MODEL
/* NVRTC path: the user source defines the plain functions; dimensions arrive as -D macros. */
#ifndef AENS_USER_HAS_TE
#define AENS_USER_HAS_TE 0
#endif
struct AensUser : AensModelDefaults {
    static constexpr bool HAS_TE = (AENS_USER_HAS_TE != 0);
#if AENS_USER_HAS_TE
    static AENS_DEV double time_events(double t, const double *y, const unsigned char *sw, const double *p) {
        return aens_time_events(t, y, sw, p);
    }
    static AENS_DEV int handle_time_event(double t, double *y, unsigned char *sw, const double *p) {
        return aens_handle_time_event(t, y, sw, p);
    }
#endif
    static constexpr int N = AENS_USER_N, NP = AENS_USER_NP, NG = AENS_USER_NG, NSW = AENS_USER_NSW;
    static constexpr bool HAS_JAC = (AENS_USER_HAS_JAC != 0);
#if defined(AENS_USER_HAS_RHS_H) && AENS_USER_HAS_RHS_H
    static constexpr bool HAS_RHS_H = true;
    static AENS_DEV void rhs_h(double h, double t, const double *y, const unsigned char *sw, const double *p, double *hk) {
        aens_rhs_h(h, t, y, sw, p, hk);
    }
#endif
    static AENS_DEV void rhs(double t, const double *y, const unsigned char *sw, const double *p, double *yd) {
        aens_rhs(t, y, sw, p, yd);
    }
    static AENS_DEV void jac(double t, const double *y, const unsigned char *sw, const double *p, double *J) {
#if AENS_USER_HAS_JAC
        aens_jac(t, y, sw, p, J);
#endif
    }
    static AENS_DEV void events(double t, const double *y, const unsigned char *sw, const double *p, double *g) {
#if AENS_USER_NG > 0
        aens_events(t, y, sw, p, g);
#endif
    }
    static AENS_DEV int handle_event(double t, double *y, unsigned char *sw, const int *info, const double *p) {
#if AENS_USER_NG > 0
        return aens_handle_event(t, y, sw, info, p);
#else
        return 0;
#endif
    }
};
#endif

#endif /* AENS_MODELS_CUH */
