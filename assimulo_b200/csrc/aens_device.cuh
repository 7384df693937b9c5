/*
 * aens_device.cuh -- device side of the ensemble integrator (sm_100a; also compiled by NVRTC for user models).
 *
 * One trajectory ("member") per thread, the complete Assimulo driver in the kernel:
 *
 *   Explicit_ODE._simulate segment loop          src/explicit_ode.pyx:137-274   -> aens_ens_kernel (phase machine)
 *   Explicit_ODE.event_locator + f_event_locator src/explicit_ode.pyx:332-361,
 *                                                src/ode_event_locator.c:17-149 -> aens_locate
 *   report_solution grid interpolation           src/explicit_ode.pyx:308-318   -> aens_emit_grid
 *   ExplicitEuler.integrate/_step/interpolate    src/solvers/euler.pyx:570-650  -> AensEuler
 *   RungeKutta4._iter/_step                      src/solvers/runge_kutta.py:839-862 -> AensRK4
 *   RungeKutta34._iter/_step/adjust_stepsize     src/solvers/runge_kutta.py:618-723 -> AensRK34
 *   Dopri5.integrate/_solout + DOPCOR/HINIT/CONTD5 runge_kutta.py:78-188, thirdparty/hairer/dopri5.f:356-692
 *                                                                                -> AensDopri5
 *   (Radau5ODE lives in aens_radau5.cuh)
 *
 * Stage vectors are per-thread arrays with compile-time indices (fully unrolled) => registers.  Global memory
 * is touched only when a member is fetched (y0, p, sw: SoA, coalesced when a warp fetches 32 consecutive
 * members), when it finishes, for dense-output rows and for event-log entries.
 *
 * Two arithmetic builds of this same source (macro AENS_STRICT):
 *   STRICT=1  compiled with --fmad=false; every expression keeps the reference's association order; pow/sqrt/div
 *             are the fp64 library/IEEE versions.  Reproduces the CPU oracle step for step.
 *   STRICT=0  FMA contraction allowed; the step-size controllers evaluate their fractional powers with the fp32
 *             MUFU lg2/ex2 pair (they only steer h); the accept test and all state arithmetic stay fp64.
 */
#ifndef AENS_DEVICE_CUH
#define AENS_DEVICE_CUH

#include "aens_args.h"
#include "aens_models.cuh"

#ifndef AENS_STRICT
#define AENS_STRICT 0
#endif

#define AENS_BLOCK 128

/* Bounds checking of our own (compute-sanitizer is not available on every pool): compiled in with -DAENS_CHECKED, every
 * index into an ensemble-sized array is tested where it is formed; a violation records `code` (first one wins) in
 * *a.check and the access is skipped or clamped by the caller's own logic.  Empty in the shipped build. */
#ifdef AENS_CHECKED
#define AENS_CHK(a, cond, code)                                                                  \
    do {                                                                                         \
        if (!(cond) && (a).check) atomicCAS((a).check, 0ull, (unsigned long long)(code));        \
    } while (0)
#else
#define AENS_CHK(a, cond, code) ((void)0)
#endif

/* return codes of Stepper::attempt (values of the reference's ID_PY_* where they exist, constants.pxi:54-58) */
#define AENS_R_CONTINUE 0
#define AENS_R_EVENT 2
#define AENS_R_COMPLETE 3
#define AENS_R_ERROR (-1)

namespace aens {

AENS_DEV double dabs(double x) { return fabs(x); }
/* g > 0 for the event indicators, on the bit pattern: positive non-zero doubles are exactly the positive 64-bit
 * integers (keeps the locator's sign tests off the fp64 pipe; differs from `g > 0.` only for NaN) */
AENS_DEV bool is_pos(double g) { return __double_as_longlong(g) > 0; }
AENS_DEV double dmax(double a, double b) { return a >= b ? a : b; } /* Fortran MAX / radau_max / fmax on non-NaN */
AENS_DEV double dmin(double a, double b) { return a <= b ? a : b; }

#if !AENS_STRICT
/* fp32 MUFU power for step-size heuristics only: x^e = ex2(e * lg2(x)).  x=0 -> 0, x=inf -> inf (e>0). */
AENS_DEV float flg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
AENS_DEV float fex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
AENS_DEV float fpow(float x, float e) { return fex2(e * flg2(x)); }
/* the argument of larger magnitude, decided on the high words (sign, exponent, 20 mantissa bits): when they tie the
 * two magnitudes agree to 2^-20, which is all a tolerance scale needs; keeps the compare off the fp64 pipe */
AENS_DEV double absmax_hi(double a, double b) {
    /* the high words compared as fp32 magnitudes: for finite non-tiny doubles the float with the same bit pattern
     * orders like the 31-bit integer, and |.| is a free operand modifier (one FSETP instead of two masks + ISETP) */
    return (fabsf(__int_as_float(__double2hiint(a))) >= fabsf(__int_as_float(__double2hiint(b)))) ? a : b;
}
/* MUFU.RCP64H seed of 1/b (~2^-23 relative) for the error norms: a perturbation of the tolerance scale, far below
 * what an accept/reject decision can see; r_recip() adds one Newton step (2^-46) where a quotient feeds a state */
AENS_DEV double rcp_seed(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    /* the instruction defines the low word as zero, which costs a move per use; any low word is as good (2^-20 of the
     * seed's own 2^-23 error budget... i.e. 2^-20 relative at most), so the operand's is kept in place */
    return __hiloint2double(__double2hiint(r), __double2loint(b));
}
#endif

/* FAST build: divisions become multiplications with a MUFU.RCP64H seed refined by one Newton step (2^-46 relative).
 * Everything they touch is either a norm that steers the iteration / the step size, or the LU factors of the
 * SIMPLIFIED Newton iteration, whose accuracy only influences the convergence rate: the iteration's fixed point is
 * defined by the right-hand-side evaluations, which stay exact fp64.  Fractional powers go through fp32 MUFU. */
#if AENS_STRICT
AENS_DEV double r_recip(double b) { return 1. / b; }
AENS_DEV double r_div(double a, double b) { return a / b; }
AENS_DEV double r_pow(double x, double e) { return pow(x, e); }
AENS_DEV double r_sqrt(double x) { return sqrt(x); }
#else
AENS_DEV double r_recip(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return fma(r, fma(-b, r, 1.0), r);
}
AENS_DEV double r_div(double a, double b) { return a * r_recip(b); }
AENS_DEV double r_pow(double x, double e) { return (double)fpow((float)x, (float)e); }
AENS_DEV double r_sqrt(double x) { return (double)sqrtf((float)x); }
#endif


/* one member's row of a member-major array [N][K]; 16-byte accesses when the row stride allows (the host side only
 * selects this path for 16-byte aligned bases): a warp that handles 32 consecutive members then moves one contiguous
 * 32*K*8-byte block, which matters when the array lives in host memory behind PCIe */
template <int K>
AENS_DEV void load_row(const double *base, long long idx, double *out) {
    const double *r = base + idx * K;
    if constexpr (K % 2 == 0) {
#pragma unroll
        for (int k = 0; k < K / 2; ++k) {
            const double2 v = reinterpret_cast<const double2 *>(r)[k];
            out[2 * k] = v.x;
            out[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = r[k];
    }
}
template <int K>
AENS_DEV void store_row(double *base, long long idx, const double *v) {
    double *r = base + idx * K;
    if constexpr (K % 2 == 0) {
#pragma unroll
        for (int k = 0; k < K / 2; ++k) reinterpret_cast<double2 *>(r)[k] = make_double2(v[2 * k], v[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = v[k];
    }
}

/* ------------------------------------------------------------------------------------------------------- */
/* Per-member state common to all solvers                                                                   */
/* ------------------------------------------------------------------------------------------------------- */
template <class M>
struct Lane {
    static constexpr int N = M::N, NP = (M::NP > 0 ? M::NP : 1), NG = (M::NG > 0 ? M::NG : 1),
                         NSW = (M::NSW > 0 ? M::NSW : 1);
    double t;
    double y[N];
    double p[NP];
    double g_old[NG];
    unsigned char sw[NSW];
    int event_info[NG];
    int st[AENS_NSTATS];
    int gi;      /* next dense-output row */
    int status;
    double tnext; /* models with time events: end of the current segment = min(next time event, tfinal) */

    /* the time integrate() runs to (tevent of Explicit_ODE._simulate, src/explicit_ode.pyx:212-222) */
    AENS_DEV double tend(const aens_args_t &a) const {
        if constexpr (M::HAS_TE) return tnext;
        else return a.tf;
    }

    AENS_DEV void rhs(double tt, const double *yy, double *yd) { M::rhs(tt, yy, sw, p, yd); }
    /* hk = h * f(tt, yy): the model's fused form where it has one (FAST build only: it re-associates a product) */
    AENS_DEV void rhs_h(double h, double tt, const double *yy, double *hk) {
        if constexpr (M::HAS_RHS_H && !AENS_STRICT) {
            M::rhs_h(h, tt, yy, sw, p, hk);
        } else {
            double yd[M::N];
            M::rhs(tt, yy, sw, p, yd);
#pragma unroll
            for (int i = 0; i < M::N; ++i) hk[i] = h * yd[i];
        }
    }

    AENS_DEV void load(const aens_args_t &a, long long idx) {
        AENS_CHK(a, idx >= 0 && idx < a.N, 1);
        if (a.zc) { /* zero-copy fetch from the caller's member-major host arrays */
            if (a.y0_shared) {
#pragma unroll
                for (int i = 0; i < M::N; ++i) y[i] = a.y0_row[i];
            } else {
                load_row<M::N>(a.in_y, idx, y);
            }
            t = a.t0; /* (the host passes -t0 to the time-reversed kernels) */
#pragma unroll
            for (int i = 0; i < M::N; ++i) a.y0_store[(long long)i * a.N + idx] = y[i];
            if (M::NP > 0) {
                load_row<NP>(a.in_p, idx, p);
#pragma unroll
                for (int i = 0; i < M::NP; ++i) a.p_store[(long long)i * a.N + idx] = p[i];
            }
#pragma unroll
            for (int i = 0; i < M::NSW; ++i) {
                sw[i] = a.in_sw[idx * M::NSW + i];
                a.sw0_store[(long long)i * a.N + idx] = sw[i];
            }
            gi = 0;
        } else {
            t = M::REV ? -a.t[idx] : a.t[idx]; /* memory holds t, a time-reversed kernel works in s = -t */
#pragma unroll
            for (int i = 0; i < M::N; ++i) y[i] = a.y[(long long)i * a.N + idx];
#pragma unroll
            for (int i = 0; i < M::NP; ++i) p[i] = a.p[(long long)i * a.N + idx];
#pragma unroll
            for (int i = 0; i < M::NSW; ++i) sw[i] = a.sw[(long long)i * a.N + idx];
            gi = a.n_grid > 0 ? a.grid_idx[idx] : 0;
        }
#pragma unroll
        for (int i = 0; i < AENS_NSTATS; ++i) st[i] = 0;
        status = AENS_ST_COMPLETE;
#if !AENS_STRICT
        M::prep(p);
#endif
    }
    /* FILL_ROWS (kernels that store dense-output rows): the rows this member never reached -- it failed, was
     * terminated, or the grid extends beyond this call's tfinal -- are set to NaN here, by the member itself, instead
     * of pre-filling the whole [rows][n][N] buffer before every launch (40 GB for config 5) */
    template <bool FILL_ROWS>
    AENS_DEV void store(const aens_args_t &a, long long idx) {
        AENS_CHK(a, idx >= 0 && idx < a.N, 2);
        AENS_CHK(a, gi >= 0 && gi <= a.n_grid, 3);
        if constexpr (FILL_ROWS) {
            if (a.dense) {
                const double qnan = __longlong_as_double(0x7ff8000000000000ll);
                for (int k = gi; k < a.n_grid; ++k) {
#pragma unroll
                    for (int i = 0; i < M::N; ++i) a.dense[((long long)k * M::N + i) * a.N + idx] = qnan;
                }
            }
        }
        a.t[idx] = M::REV ? -t : t;
#pragma unroll
        for (int i = 0; i < M::N; ++i) a.y[(long long)i * a.N + idx] = y[i];
#pragma unroll
        for (int i = 0; i < M::NSW; ++i) a.sw[(long long)i * a.N + idx] = sw[i];
#pragma unroll
        for (int i = 0; i < AENS_NSTATS; ++i) a.stats[(long long)i * a.N + idx] = st[i];
        if (a.n_grid > 0) a.grid_idx[idx] = gi;
        a.status[idx] = status;
        if (a.out_y) store_row<M::N>(a.out_y, idx, y);
        if (a.out_t) a.out_t[idx] = M::REV ? -t : t;
        if (a.out_status) a.out_status[idx] = status;
        if (M::NSW > 0 && a.out_sw) {
#pragma unroll
            for (int i = 0; i < M::NSW; ++i) a.out_sw[idx * M::NSW + i] = sw[i];
        }
    }
};

/* set_problem_data() of every solver class (runge_kutta.py:78-97, euler.pyx:537-555, radau5.py:232-274):
 * evaluate g at the segment start */
template <class M>
AENS_DEV void seg_init_events(Lane<M> &L) {
    if (M::NG > 0) {
        M::events(L.t, L.y, L.sw, L.p, L.g_old);
        L.st[AENS_STAT_NSTATEFCNS] += 1;
    }
}

/* src/ode_event_locator.c:17-149 wrapped as in src/explicit_ode.pyx:332-361.  `S::interp(x, out)` is the dense
 * output of the step just taken.  On AENS_R_EVENT, t_high / y_high hold the event point. */
/* bit i = (g[i] > 0): the locator's sign tests on one small mask per bracket end instead of on the doubles */
template <int NG>
AENS_DEV unsigned pos_mask(const double *g) {
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < NG; ++i) m |= (is_pos(g[i]) ? 1u : 0u) << i;
    return m;
}

template <class M, class S>
AENS_DEV int aens_locate(S &s, Lane<M> &L, double t_low, double &t_high, double *y_high) {
    constexpr int NG = M::NG > 0 ? M::NG : 1;
    double g_low[NG], g_mid[NG], g_high[NG];
#pragma unroll
    for (int i = 0; i < M::NG; ++i) g_low[i] = L.g_old[i];
    const double TOL = dmax(dabs(t_low), dabs(t_high)) * 1.e-13;
    M::events(t_high, y_high, L.sw, L.p, g_high);
    L.st[AENS_STAT_NSTATEFCNS] += 1;
    /* detection (ode_event_locator.c:55): some (g_low > 0) != (g_high > 0) */
    unsigned pm_low = pos_mask<M::NG>(g_low), pm_high = pos_mask<M::NG>(g_high);
    const bool ev = (pm_low != pm_high);
    if (ev) {
        int side = 0, sideprev = -1;
        double alpha = 1.;
        const double half_tol = TOL / 2.;
        while (dabs(t_high - t_low) > TOL) {
            if (side == sideprev) {
                if (side == 2) alpha *= 2.; else alpha /= 2.;
            } else {
                alpha = 1.;
            }
            /* component with the largest |g_high/(g_low-g_high)| among the sign-changed ones (ties -> last).  With a
             * single sign change -- the usual case -- that component is the answer and the quotients are not needed
             * (they only rank the candidates); the STRICT build keeps the reference's loop as it stands. */
            const unsigned changed = pm_low ^ pm_high;
            double gh = g_high[0], gl = g_low[0];
            if (!AENS_STRICT && (changed & (changed - 1u)) == 0u) {
#pragma unroll
                for (int i = 1; i < M::NG; ++i)
                    if ((changed >> i) & 1u) { gh = g_high[i]; gl = g_low[i]; }
            } else {
                double maxfrac = 0.;
#pragma unroll
                for (int i = 0; i < M::NG; ++i) {
                    if ((changed >> i) & 1u) {
                        const double gfrac = dabs(r_div(g_high[i], g_low[i] - g_high[i]));
                        if (gfrac >= maxfrac) { maxfrac = gfrac; gh = g_high[i]; gl = g_low[i]; }
                    }
                }
            }
            double t_mid;
            if (gh == 0 || gl == 0) t_mid = (t_low + t_high) / 2.;
            else t_mid = t_high - r_div((t_high - t_low) * gh, gh - alpha * gl);
            if (dabs(t_mid - t_low) < half_tol) {
                const double fracint = r_div(dabs(t_low - t_high), TOL);
                t_mid = t_low + r_div(t_high - t_low, 2. * dmin(fracint, 5.));
            }
            if (dabs(t_mid - t_high) < half_tol) {
                const double fracint = r_div(dabs(t_low - t_high), TOL);
                t_mid = t_high - r_div(t_high - t_low, 2. * dmin(fracint, 5.));
            }
            s.interp(t_mid, y_high);
            M::events(t_mid, y_high, L.sw, L.p, g_mid);
            L.st[AENS_STAT_NSTATEFCNS] += 1;
            sideprev = side;
            const unsigned pm_mid = pos_mask<M::NG>(g_mid);
            if (pm_low != pm_mid) { /* an event in [t_low, t_mid] */
                t_high = t_mid;
#pragma unroll
                for (int i = 0; i < M::NG; ++i) g_high[i] = g_mid[i];
                pm_high = pm_mid;
                side = 1;
            } else {
                t_low = t_mid;
#pragma unroll
                for (int i = 0; i < M::NG; ++i) g_low[i] = g_mid[i];
                pm_low = pm_mid;
                side = 2;
            }
        }
        s.interp(t_high, y_high);
        const unsigned changed = pm_low ^ pm_high;
#pragma unroll
        for (int i = 0; i < M::NG; ++i)
            L.event_info[i] = ((changed >> i) & 1u) ? (((pm_high >> i) & 1u) ? 1 : -1) : 0;
        L.st[AENS_STAT_NSTATEEVENTS] += 1;
    }
#pragma unroll
    for (int i = 0; i < M::NG; ++i) L.g_old[i] = g_high[i];
    return ev ? AENS_R_EVENT : AENS_R_CONTINUE;
}

/* report_solution(): rows `while grid[k] <= t` from the interpolant of the last step (explicit_ode.pyx:308-318).
 * Row layout [k][component][member]: lanes of a warp that emit the same row store 256 contiguous bytes. */
/* Row reduction mode (ensemble outputs without materialising [rows][n][N] in HBM).  Instead of storing
 * its row values a lane hands them to its warp through a shared tile; the lanes of the warp that arrived here
 * together add them to the WARP's private partial record of that row,
 *     rec[row] = { pivot[n], sum(v - pivot)[n], sum((v - pivot)^2)[n], min[n], max[n], count },
 * with plain loads and stores (the records of the ~10 rows a warp is working on stay in L1/L2); aens_combine_rows_k
 * merges the partials of all warps after the launch.  The lanes of a warp reach this point in arbitrary subsets
 * (others are rejecting a step, locating an event, ...), and independent thread scheduling may interleave two such
 * subsets, so a per-warp lock in shared memory serialises them. */
AENS_DEV int *aens_red_lock() {
    __shared__ int lock[AENS_BLOCK / 32];
    return lock;
}

template <int ITEMS, int RS>
struct RedShared {
    double tile[AENS_BLOCK / 32][ITEMS][33];
    unsigned tmask[AENS_BLOCK / 32][RS];
};
template <int ITEMS, int RS>
AENS_DEV RedShared<ITEMS, RS> &aens_red_shared() {
    __shared__ RedShared<ITEMS, RS> sh;
    return sh;
}

template <class M, class S>
AENS_DEV void aens_reduce_grid(S &s, Lane<M> &L, const aens_args_t &a, double t) {
    constexpr int N = M::N, REC = 5 * N + 1;
    /* rows handled per pass: a step usually crosses several rows, and up to 32 (row, component) items give every
     * lane of the warp one column to accumulate in the loop below; 33-double pitch = conflict-free both ways */
    constexpr int RS = (32 / N) > 8 ? 8 : (32 / N), ITEMS = RS * N;
    /* (one shared tile per kernel whatever the interpolant type: see aens_red_shared) */
    RedShared<ITEMS, RS> &sh = aens_red_shared<ITEMS, RS>();
    double (*tile)[ITEMS][33] = sh.tile;
    unsigned (*tmask)[RS] = sh.tmask;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned mask = __activemask();
    bool pend = L.gi < a.n_grid && __ldg(a.t_grid + L.gi) <= t;
    if (!__any_sync(mask, pend)) return;
    const int nact = __popc(mask), rank = __popc(mask & ((1u << lane) - 1u));
    int *lk = aens_red_lock() + w;
    if (rank == 0) {
        while (atomicCAS(lk, 0, 1) != 0) __nanosleep(40);
        __threadfence_block();
    }
    __syncwarp(mask);
    AENS_CHK(a, a.red != nullptr && (long long)(blockIdx.x * (AENS_BLOCK / 32) + w) < a.red_warps, 5);
    double *base = a.red + (long long)(blockIdx.x * (AENS_BLOCK / 32) + w) * a.n_grid * REC;
    /* one column of the tile (the values the lanes in mm hold for one row and component) into (sm, q, mn, mx);
     * two interleaved partial sums halve the dependent fp64 chains, min and max share one compare per pair.
     * (Testing the values against the record's bounds first and updating the extrema in a rare second scan was
     * slower: parameter sweeps are sorted, so nearly every batch of members moves an extremum.) */
    auto accumulate = [](const double *col, unsigned mm, double K, double &sm, double &q, double &mn, double &mx) {
        if (mm == 0xffffffffu) {
            double s1 = 0.0, q1 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double v0 = col[j], v1 = col[j + 1], d0 = v0 - K, d1 = v1 - K;
                sm += d0; s1 += d1;
                q = fma(d0, d0, q); q1 = fma(d1, d1, q1);
                const double lo = v0 < v1 ? v0 : v1, hi = v0 < v1 ? v1 : v0;
                mn = lo < mn ? lo : mn;
                mx = hi > mx ? hi : mx;
            }
            sm += s1; q += q1;
        } else {
            for (unsigned b = mm; b; b &= b - 1u) {
                const double v = col[__ffs(b) - 1], d = v - K;
                sm += d;
                q = fma(d, d, q);
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
        }
    };
    for (;;) {
        const int r0 = __reduce_min_sync(mask, pend ? L.gi : 0x7fffffff);
        if (r0 == 0x7fffffff) break;
        /* enough lanes for one item each: the item's record is loaded before the interpolation hides the latency,
         * and the lane of component 0 also owns the row's count */
        const bool one_each = nact >= ITEMS;
        const int k1 = rank / N, c1 = rank - k1 * N;
        const bool own = one_each && rank < ITEMS && r0 + k1 < a.n_grid;
        AENS_CHK(a, r0 >= 0 && r0 < a.n_grid, 6);
        double K = 0.0, sm = 0.0, q = 0.0, mn = 0.0, mx = 0.0, cnt0 = 0.0;
        double *rec1 = base + (long long)(r0 + k1) * REC;
        if (own) {
            cnt0 = rec1[5 * N];
            K = rec1[c1]; sm = rec1[N + c1]; q = rec1[2 * N + c1]; mn = rec1[3 * N + c1]; mx = rec1[4 * N + c1];
        }
#pragma unroll
        for (int k = 0; k < RS; ++k) {
            const bool mine = pend && L.gi == r0 + k;
            const unsigned mm = __ballot_sync(mask, mine);
            if (rank == 0) tmask[w][k] = mm;
            if (mine) {
                double out[N];
                AENS_CHK(a, k * N + N <= ITEMS && L.gi < a.n_grid, 7);
                s.interp(__ldg(a.t_grid + L.gi), out);
#pragma unroll
                for (int i = 0; i < N; ++i) tile[w][k * N + i][lane] = out[i];
                L.gi += 1;
                pend = L.gi < a.n_grid && __ldg(a.t_grid + L.gi) <= t;
            }
        }
        __syncwarp(mask);
        if (one_each) {
            const unsigned mm = own ? tmask[w][k1] : 0u;
            if (mm) {
                const double *col = tile[w][rank];
                if (cnt0 == 0.0) {
                    K = col[__ffs(mm) - 1];
                    if (!(dabs(K) < 1.0e300)) K = 0.0; /* a non-finite first value must not become the pivot */
                    sm = 0.0; q = 0.0; mn = 1.0e300 * 1.0e300; mx = -1.0e300 * 1.0e300;
                }
                accumulate(col, mm, K, sm, q, mn, mx);
                rec1[c1] = K; rec1[N + c1] = sm; rec1[2 * N + c1] = q; rec1[3 * N + c1] = mn; rec1[4 * N + c1] = mx;
                if (c1 == 0) rec1[5 * N] = cnt0 + (double)__popc(mm);
            }
            __syncwarp(mask);
            continue;
        }
        for (int it = rank; it < ITEMS; it += nact) {
            const int k = it / N, c = it - k * N;
            const unsigned mm = tmask[w][k];
            if (mm == 0u) continue;
            AENS_CHK(a, r0 + k < a.n_grid, 8);
            double *rec = base + (long long)(r0 + k) * REC;
            const double *col = tile[w][it];
            if (rec[5 * N] == 0.0) {
                K = col[__ffs(mm) - 1];
                if (!(dabs(K) < 1.0e300)) K = 0.0;
                sm = 0.0; q = 0.0; mn = 1.0e300 * 1.0e300; mx = -1.0e300 * 1.0e300;
            } else {
                K = rec[c]; sm = rec[N + c]; q = rec[2 * N + c]; mn = rec[3 * N + c]; mx = rec[4 * N + c];
            }
            accumulate(col, mm, K, sm, q, mn, mx);
            rec[c] = K; rec[N + c] = sm; rec[2 * N + c] = q; rec[3 * N + c] = mn; rec[4 * N + c] = mx;
        }
        __syncwarp(mask); /* every item has read its row's count and the tile */
        for (int k = rank; k < RS; k += nact) {
            const unsigned mm = tmask[w][k];
            if (mm) base[(long long)(r0 + k) * REC + 5 * N] += (double)__popc(mm);
        }
        __syncwarp(mask);
    }
    if (rank == 0) {
        __threadfence_block();
        atomicExch(lk, 0);
    }
}

template <class M, class S>
AENS_DEV void aens_emit_grid(S &s, Lane<M> &L, const aens_args_t &a, long long idx, double t) {
    if constexpr (S::REDUCE) { /* the d2 kernels: rows are reduced over the ensemble, not stored */
        aens_reduce_grid<M, S>(s, L, a, t);
        return;
    }
    while (L.gi < a.n_grid && __ldg(a.t_grid + L.gi) <= t) {
        double out[M::N];
        AENS_CHK(a, a.dense != nullptr && L.gi >= 0 && idx >= 0 && idx < a.N, 4);
        s.interp(__ldg(a.t_grid + L.gi), out);
#pragma unroll
        for (int i = 0; i < M::N; ++i) a.dense[((long long)L.gi * M::N + i) * a.N + idx] = out[i];
        L.gi += 1;
    }
}

/* Grid rows at or before the start of a segment (the first SOLOUT call of dopri5 / radau5 / rodas with XOLD = X reports
 * them from the state itself, dopri5.f:399-404): the same emit path as every other row -- stored by the d1 kernels,
 * reduced over the ensemble by the d2 kernels (which have no row buffer at all). */
template <class M, bool RED>
struct HoldState {
    static constexpr bool REDUCE = RED;
    const double *y;
    AENS_DEV void interp(double, double *out) const {
#pragma unroll
        for (int i = 0; i < M::N; ++i) out[i] = y[i];
    }
};
template <class M, class S>
AENS_DEV void aens_emit_start(Lane<M> &L, const aens_args_t &a, long long idx) {
    if (L.gi < a.n_grid && __ldg(a.t_grid + L.gi) <= L.t) {
        HoldState<M, S::REDUCE> hs{L.y};
        aens_emit_grid<M, HoldState<M, S::REDUCE> >(hs, L, a, idx, L.t);
    }
}

/* After an accepted step [told, t] whose end state is in L.y: event location, then grid rows.
 * Mirrors the order of Dopri5._solout / Radau5ODE._solout / the loops of RungeKutta34 and ExplicitEuler:
 * locator first, then output up to the (possibly shortened) t. */
template <class M, class S>
AENS_DEV int aens_post_step(S &s, Lane<M> &L, const aens_args_t &a, long long idx, double told, double tnew) {
    int flag = AENS_R_CONTINUE;
    if (M::NG > 0) {
        double te = tnew;
        double ye[M::N];
#pragma unroll
        for (int i = 0; i < M::N; ++i) ye[i] = L.y[i];
        flag = aens_locate<M, S>(s, L, told, te, ye);
        if (flag == AENS_R_EVENT) {
            /* grid rows up to the event time come from the interpolant of the FULL step, which for the
             * one-step methods reads L.y as its right end point: emit before L.y becomes the event state */
            if (S::HAS_DENSE) aens_emit_grid<M, S>(s, L, a, idx, te);
#pragma unroll
            for (int i = 0; i < M::N; ++i) L.y[i] = ye[i];
            L.t = te;
            return flag;
        }
    }
    if (S::HAS_DENSE) aens_emit_grid<M, S>(s, L, a, idx, tnew);
    L.t = tnew;
    return flag;
}

/* ------------------------------------------------------------------------------------------------------- */
/* ExplicitEuler, src/solvers/euler.pyx:485-692, and ImplicitEuler, src/solvers/euler.pyx:29-483             */
/* (same integrate() loop, same linear dense output, same post-event partial step; they differ in _step)     */
/* ------------------------------------------------------------------------------------------------------- */
/* small dense LU of aens_radau5.cuh (Moler's DEC/SOL = LAPACK's partial pivoting) */
template <int N> AENS_DEV int lu_dec(double *a, int *ip);
template <int N, bool PIV = true> AENS_DEV void lu_sol(const double *a, double *b, const int *ip);

template <class M, int DENSE, bool IMPL = false>
struct Euler {
    static constexpr bool HAS_DENSE = DENSE != 0;
    static constexpr bool REDUCE = DENSE == 2;
    static constexpr int N = M::N;
    Lane<M> L;
    double h, inith;          /* inith = _inith: remainder of a grid step cut by an event */
    double yold[DENSE ? N : 1], told, hstep;
    bool pending;
    /* ImplicitEuler only (euler.pyx:87-92): Jacobian reuse state */
    double jac[IMPL ? N * N : 1];
    int since;                /* _steps_since_last_jac */
    bool needjac, curjac;

    AENS_DEV void seg_end() {}
    AENS_DEV void late_status(const aens_args_t &) {}
    AENS_DEV void load_aux(const aens_args_t &a, long long idx) {
        inith = a.aux[idx];
        /* a fresh solver object: _needjac = True (euler.pyx:89).  The reference keeps its Jacobian across
         * simulate() calls of one solver object; the batched solver re-evaluates it at the start of every call. */
        needjac = true; curjac = false; since = 0;
    }
    AENS_DEV void store_aux(const aens_args_t &a, long long idx) { a.aux[idx] = inith; }

    AENS_DEV void interp(double x, double *out) { /* euler.pyx:440-441, 649-650 */
        if (DENSE) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = yold[i] + (x - told) / hstep * (L.y[i] - yold[i]);
        }
    }
    AENS_DEV void seg_init(const aens_args_t &a, long long) {
        h = dmin(a.o.h, dabs(L.tend(a) - L.t));
        seg_init_events(L);
        pending = (inith != 0);
    }
    /* ExplicitEuler._step, euler.pyx:636-647 */
    AENS_DEV bool step_explicit(double hs) {
        double f[N];
        L.rhs(L.t, L.y, f);
        L.st[AENS_STAT_NFCNS] += 1;
        L.st[AENS_STAT_NSTEPS] += 1;
        L.st[AENS_STAT_NSTEPSTOTAL] += 1;
        told = L.t;
        hstep = hs;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (DENSE) yold[i] = L.y[i];
            L.y[i] = L.y[i] + hs * f[i];
        }
        return true;
    }
    /* ImplicitEuler._jacobian, euler.pyx:322-345 (J column-major) */
    AENS_DEV void jacobian(const aens_args_t &a, double t, const double *y) {
        curjac = true; needjac = false; since = 0;
        if (a.o.usejac && M::HAS_JAC) {
            M::jac(t, y, L.sw, L.p, jac);
        } else {
            double f0[N];
            L.rhs(t, y, f0);
#pragma unroll(N <= 4 ? 64 : 1)
            for (int i = 0; i < N; ++i) {
                const double delt = sqrt(2.220446049250313e-16 * dmax(dabs(y[i]), 1.e-5));
                double yp[N], fp[N];
#pragma unroll(N <= 4 ? 64 : 1)
                for (int j = 0; j < N; ++j) yp[j] = (j == i) ? y[j] + delt : y[j];
                L.rhs(t, yp, fp);
#pragma unroll(N <= 4 ? 64 : 1)
                for (int j = 0; j < N; ++j) jac[j + i * N] = (fp[j] - f0[j]) / delt;
            }
        }
        L.st[AENS_STAT_NJACS] += 1;
    }
    /* ImplicitEuler._step, euler.pyx:363-438: y_{n+1} = y_n + h f(t_{n+1}, y_{n+1}) by a Newton iteration with a lazily
     * updated Jacobian (refreshed after 20 steps or after a failed iteration) and np.linalg.solve = LU with partial
     * pivoting.  Returns false when the iteration fails with a current Jacobian ("Newton iteration failed"). */
    AENS_DEV bool step_implicit(const aens_args_t &a, double hs) {
        const double t = L.t, tn1 = t + hs;
        double yn[N], yn1[N], ynew[N], f[N];
        L.rhs(t, L.y, f);
        L.st[AENS_STAT_NFCNS] += 1;
        L.st[AENS_STAT_NSTEPSTOTAL] += 1;
#pragma unroll(N <= 4 ? 64 : 1)
        for (int i = 0; i < N; ++i) { yn[i] = L.y[i]; yn1[i] = L.y[i] + hs * f[i]; ynew[i] = yn1[i]; }
        bool conv = false;
        for (int j = 0; j < 2 && !conv; ++j) {
            bool fail = false;
            if (needjac) jacobian(a, tn1, yn1);
            double A[N * N];
            int ip[N];
#pragma unroll(N <= 4 ? 64 : 1)
            for (int c = 0; c < N; ++c) {
#pragma unroll(N <= 4 ? 64 : 1)
                for (int r = 0; r < N; ++r) A[r + c * N] = hs * jac[r + c * N] - (r == c ? 1.0 : 0.0);
            }
            if (lu_dec<N>(A, ip) != 0) { L.status = AENS_ST_RADAU_SINGULAR; return false; } /* LinAlgError */
            double old_norm = 0.0;
            int i = 0;
            for (; i < a.o.newt; ++i) {
                L.st[AENS_STAT_NNITERS] += 1;
                double b[N];
                L.rhs(tn1, yn1, f);
                L.st[AENS_STAT_NFCNS] += 1;
#pragma unroll(N <= 4 ? 64 : 1)
                for (int k = 0; k < N; ++k) b[k] = yn[k] - yn1[k] + hs * f[k];
                lu_sol<N>(A, b, ip);
                double sum = 0.0;
#pragma unroll(N <= 4 ? 64 : 1)
                for (int k = 0; k < N; ++k) {
                    ynew[k] = yn1[k] - b[k];
                    const double w = 1.0 / (a.o.rtol * dabs(yn1[k]) + a.o.atol[k]);
                    const double prod = (ynew[k] - yn1[k]) * w;
                    sum += prod * prod;
                }
                const double new_norm = sqrt(sum / N); /* WRMS, euler.pyx:349-361 */
                if (new_norm < 0.1) { conv = true; break; }
                if (i > 0 && new_norm / old_norm > 2) { fail = true; break; }
#pragma unroll(N <= 4 ? 64 : 1)
                for (int k = 0; k < N; ++k) yn1[k] = ynew[k];
                old_norm = new_norm;
            }
            if (!conv && i >= a.o.newt) fail = true;
            if (fail) {
                L.st[AENS_STAT_NNFAILS] += 1;
                if (curjac) { L.status = AENS_ST_NEWTON_FAILED; return false; }
                needjac = true;
            }
            if (conv) {
                since += 1;
                L.st[AENS_STAT_NSTEPS] += 1;
                if (since > 20) needjac = true;
            }
        }
        if (!conv) { L.status = AENS_ST_NEWTON_FAILED; return false; }
        curjac = false;
        told = t;
        hstep = hs;
#pragma unroll(N <= 4 ? 64 : 1)
        for (int i = 0; i < N; ++i) {
            if (DENSE) yold[i] = yn[i];
            L.y[i] = ynew[i];
        }
        return true;
    }
    AENS_DEV bool step(const aens_args_t &a, double hs) {
        if constexpr (IMPL) return step_implicit(a, hs);
        else return step_explicit(hs);
    }
    /* the 1e-15 cut-off exists only in ExplicitEuler (euler.pyx:631-632 vs :218-219) */
    AENS_DEV double trim(double v) const { return (IMPL || dabs(v) > 1e-15) ? v : 0.0; }

    AENS_DEV int attempt(const aens_args_t &a, long long idx) {
        if (pending) { /* euler.pyx:583-600 / 167-184 */
            const double hs = inith;
            if (!step(a, hs)) return AENS_R_ERROR;
            const double tn = told + hs;
            const int flag = aens_post_step<M, Euler>(*this, L, a, idx, tn - hs, tn);
            pending = false;
            if (flag == AENS_R_EVENT) inith = inith - (L.t - told);
            else inith = 0;
            h = dmin(h, dabs(L.tend(a) - L.t));
            if (flag == AENS_R_EVENT && inith == 0) inith = trim(h - (L.t - told)); /* euler.pyx:629-632 / 216-219 */
            return flag;
        }
        const bool inner = (L.t + h < L.tend(a));
        const double hs = h;
        if (!step(a, hs)) return AENS_R_ERROR;
        const double tn = told + hs;
        int flag = aens_post_step<M, Euler>(*this, L, a, idx, tn - hs, tn);
        if (inner) h = dmin(h, dabs(L.tend(a) - L.t));
        if (flag == AENS_R_EVENT) {
            if (inith == 0) inith = trim(h - (L.t - told)); /* euler.pyx:629-632 / 216-219 */
            return AENS_R_EVENT;
        }
        return inner ? AENS_R_CONTINUE : AENS_R_COMPLETE;
    }
};
template <class M, int DENSE>
using ImplEuler = Euler<M, DENSE, true>;

/* ------------------------------------------------------------------------------------------------------- */
/* RungeKutta4, src/solvers/runge_kutta.py:839-862 (no events, no dense output in the reference)            */
/* ------------------------------------------------------------------------------------------------------- */
template <class M, int DENSE>
struct RK4 {
    static constexpr bool HAS_DENSE = false;
    static constexpr bool REDUCE = false;
    static constexpr int N = M::N;
    Lane<M> L;
    double h;
    double h6;                /* h / 6. of the current h: the fp64 division runs only when h changes (last step), and
                               * yields the same double as recomputing it every step like the reference does */
    AENS_DEV void seg_end() {}
    AENS_DEV void late_status(const aens_args_t &) {}
    AENS_DEV void load_aux(const aens_args_t &, long long) {}
    AENS_DEV void store_aux(const aens_args_t &, long long) {}
    AENS_DEV void interp(double, double *) {}
    AENS_DEV void seg_init(const aens_args_t &a, long long) {
        h = dmin(a.o.h, dabs(L.tend(a) - L.t));
        h6 = h / 6.;
    }
    AENS_DEV int attempt(const aens_args_t &a, long long) {
        const bool inner = (L.t + h < L.tend(a));
        double Y1[N], Y2[N], Y3[N], Y4[N], arg[N];
        const double t = L.t;
        L.rhs(t, L.y, Y1);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y1[i] / 2.;
        L.rhs(t + h / 2., arg, Y2);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y2[i] / 2.;
        L.rhs(t + h / 2., arg, Y3);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y3[i];
        L.rhs(t + h, arg, Y4);
#pragma unroll
        for (int i = 0; i < N; ++i) L.y[i] = L.y[i] + h6 * (Y1[i] + 2. * Y2[i] + 2. * Y3[i] + Y4[i]);
        L.t = t + h;
        L.st[AENS_STAT_NFCNS] += 4;
        L.st[AENS_STAT_NSTEPS] += 1;
        L.st[AENS_STAT_NSTEPSTOTAL] += 1;
        if (inner) {
            const double hn = dmin(h, dabs(L.tend(a) - L.t));
            if (hn != h) { h = hn; h6 = hn / 6.; }
            return AENS_R_CONTINUE;
        }
        return AENS_R_COMPLETE;
    }
};

/* ------------------------------------------------------------------------------------------------------- */
/* RungeKutta34, src/solvers/runge_kutta.py:618-723                                                          */
/* ------------------------------------------------------------------------------------------------------- */
template <class M, int DENSE>
struct RK34 {
    static constexpr bool HAS_DENSE = DENSE != 0;
    static constexpr bool REDUCE = DENSE == 2;
    static constexpr int N = M::N;
    Lane<M> L;
    double h, Y1[N];
    double yold[DENSE ? N : 1], flow[DENSE ? N : 1], told, tnext, hstep;
    int it;

    AENS_DEV void seg_end() {}
    AENS_DEV void late_status(const aens_args_t &) {}
    AENS_DEV void load_aux(const aens_args_t &, long long) {}
    AENS_DEV void store_aux(const aens_args_t &, long long) {}
    AENS_DEV void interp(double time, double *out) { /* Hermite, runge_kutta.py:715-720; f_high = Y1 */
        if (DENSE) {
            const double th = (time - told) / (tnext - told);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double y = yold[i], yn = L.y[i];
                out[i] = (1 - th) * y + th * yn + th * (th - 1) *
                         ((1 - 2 * th) * (yn - y) + (th - 1) * hstep * flow[i] + th * hstep * Y1[i]);
            }
        }
    }
    AENS_DEV void seg_init(const aens_args_t &a, long long) {
        seg_init_events(L);
        h = dmin(a.o.inith, dabs(L.tend(a) - L.t));
        L.rhs(L.t, L.y, Y1);
        it = 0;
    }
    AENS_DEV int attempt(const aens_args_t &a, long long idx) {
        const bool inner = (L.t + h < L.tend(a));
        if (it >= a.o.maxsteps) { /* for-else: 'Final time not reached within maximum number of steps' */
            L.status = AENS_ST_MAXSTEPS;
            return AENS_R_ERROR;
        }
        double Y2[N], Y3[N], Z3[N], Y4[N], arg[N], sc[N];
        const double t = L.t;
        L.st[AENS_STAT_NFCNS] += 5;
#pragma unroll
        for (int i = 0; i < N; ++i) sc[i] = dabs(L.y[i]) * a.o.rtol + a.o.atol[i];
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y1[i] / 2.;
        L.rhs(t + h / 2., arg, Y2);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y2[i] / 2.;
        L.rhs(t + h / 2., arg, Y3);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] - h * Y1[i] + 2.0 * h * Y2[i];
        L.rhs(t + h, arg, Z3);
#pragma unroll
        for (int i = 0; i < N; ++i) arg[i] = L.y[i] + h * Y3[i];
        L.rhs(t + h, arg, Y4);
#if AENS_STRICT
        const double h6 = h / 6.0;
#else
        const double h6 = h * (1.0 / 6.0);
#endif
        double err2 = 0.;
#pragma unroll
        for (int i = 0; i < N; ++i) {
#if AENS_STRICT
            const double e = h6 * (2.0 * Y2[i] + Z3[i] - 2.0 * Y3[i] - Y4[i]) / sc[i];
            err2 = fma(e, e, err2); /* numpy x.dot(x) = OpenBLAS ddot accumulates fused (oracle/ens_oracle.c) */
#else
            const double e = h6 * (2.0 * Y2[i] + Z3[i] - 2.0 * Y3[i] - Y4[i]) * rcp_seed(sc[i]); /* 2^-23 on the tolerance */
            err2 = fma(e, e, err2);
#endif
        }
        const double tn = t + h;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (DENSE) { yold[i] = L.y[i]; flow[i] = Y1[i]; }
            L.y[i] = L.y[i] + h6 * (Y1[i] + 2.0 * Y2[i] + 2.0 * Y3[i] + Y4[i]);
        }
        L.rhs(tn, L.y, Y1);
        told = t; tnext = tn; hstep = h;
        L.st[AENS_STAT_NSTEPS] += 1;
        L.st[AENS_STAT_NSTEPSTOTAL] += 1;
        const int flag = aens_post_step<M, RK34>(*this, L, a, idx, tn - h, tn);
        if (inner) {
            /* adjust_stepsize, runge_kutta.py:682-692 */
            double fac;
#if AENS_STRICT
            const double error = sqrt(err2);
            if (error == 0.0) fac = 2.0;
            else fac = dmin(pow(1.0 / error, 1.0 / 4.0), 2.);
#else
            /* (1/err)^(1/4) = err2^(-1/8); err2 == 0 -> +inf -> clipped to 2 */
            fac = (double)fminf(fpow((float)err2, -0.125f), 2.0f);
            if (!(err2 > 0.0)) fac = 2.0;
#endif
            h = h * fac;
            h = dmin(h, dabs(L.tend(a) - L.t));
            it += 1;
        }
        if (flag == AENS_R_EVENT) return AENS_R_EVENT;
        return inner ? AENS_R_CONTINUE : AENS_R_COMPLETE;
    }
};

/* ------------------------------------------------------------------------------------------------------- */
/* Dopri5: DOPCOR dopri5.f:356-556, HINIT :558-619, CONTD5 :621-644, CDOPRI :646-692                         */
/* ------------------------------------------------------------------------------------------------------- */
namespace dp {
/* The tableau lives in __constant__ memory: a DFMA/DMUL then reads its coefficient straight from the constant
 * bank (c[3][..] operand).  As constexpr literals every coefficient costs two UMOVs per use -- 20 % of all issued
 * instructions in the first profile (profiles/r1_dopri5_vdp.md). */
struct Tab {
    double c2, c3, c4, c5;
    double a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, a71, a73, a74, a75, a76;
    double e1, e3, e4, e5, e6, e7;
    double d1, d3, d4, d5, d6, d7;
    double inv101; /* 1/1.01 of the end-point test (as a literal it costs two UMOVs per attempt) */
};
static __constant__ Tab T = {
    0.2, 0.3, 0.8, 8.0 / 9.0,
    0.2, 3.0 / 40.0, 9.0 / 40.0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 19372.0 / 6561.0, -25360.0 / 2187.0,
    64448.0 / 6561.0, -212.0 / 729.0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0,
    -5103.0 / 18656.0, 35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0,
    71.0 / 57600.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0,
    -12715105075.0 / 11282082432.0, 87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0,
    701980252875.0 / 199316789632.0, -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0,
    1.0 / 1.01};
constexpr double uround = 2.3e-16; /* dopri5.f:253-254 */
}  // namespace dp

/* lg2 of the state dimension (folds the 1/N of the RMS norm into the log-domain controller) */
__host__ __device__ constexpr float lg2_int(int n) {
    return n == 1 ? 0.f : n == 2 ? 1.f : n == 3 ? 1.5849625007f : n == 4 ? 2.f : n == 5 ? 2.3219280949f :
           n == 6 ? 2.5849625007f : n == 7 ? 2.8073549221f : n == 8 ? 3.f : n == 9 ? 3.1699250014f :
           n == 10 ? 3.3219280949f : n == 11 ? 3.4594316186f : n == 12 ? 3.5849625007f : n == 13 ? 3.7004397181f :
           n == 14 ? 3.8073549221f : n == 15 ? 3.9068905956f : 4.f;
}

template <class M, int DENSE>
struct Dopri5 {
    static constexpr bool HAS_DENSE = DENSE != 0;
    static constexpr bool REDUCE = DENSE == 2;
    static constexpr int N = M::N;
    /* FAST build, small systems: the stage arguments are accumulated incrementally (see attempt_fast) */
    static constexpr bool INCREMENTAL = (N <= 4);
    Lane<M> L;
    double k1[N];
    double h, hmax;
#if AENS_STRICT
    double facold;             /* FACOLD */
#else
    float lfacold;             /* beta/2 * lg2(max(err^2, 1e-8)): the controller works on lg2 of the squared norm */
    float lfmin;               /* lower bound of lg2 fac for the next accepted step: lg2 facc2, or 0 after a rejection
                                * ('a step after a rejection does not grow', dopri5.f:516) -- the REJECT flag as data */
#endif
    double cont[DENSE ? 5 * N : 1], told, hstep;
    int nstep, naccpt, nrej;   /* per dopri5() call, i.e. per segment; flushed into L.st when the segment ends */
    int last, reject;          /* flags (int: byte-sized bools cost PRMT packing instructions in the hot loop) */

    AENS_DEV void load_aux(const aens_args_t &, long long) {}
    AENS_DEV void store_aux(const aens_args_t &, long long) {}

    AENS_DEV void interp(double x, double *out) { /* CONTD5 */
        if (DENSE) {
#if AENS_STRICT
            const double th = (x - told) / hstep, th1 = 1.0 - th;
#else
            const double th = (x - told) * hstep, th1 = 1.0 - th; /* FAST: hstep holds 1/h (set once per accepted step) */
#endif
#pragma unroll
            for (int i = 0; i < N; ++i)
                out[i] = cont[i] + th * (cont[N + i] + th1 * (cont[2 * N + i] + th * (cont[3 * N + i] + th1 * cont[4 * N + i])));
        }
    }

    /* attempt_fast() leaves the hot loop with AENS_R_ERROR and no status: which of the two limits it was is decided
     * here, once, instead of by predicated instructions in every attempt (dopri5.f:405-410) */
    AENS_DEV void late_status(const aens_args_t &a) {
        if (L.status == AENS_ST_COMPLETE) L.status = (nstep > a.o.maxsteps) ? AENS_ST_MAXSTEPS : AENS_ST_STEP_TOO_SMALL;
    }
    /* The hot loop without events and dense output (FAST) counts in its accept block only: naccpt, and in nrej the
     * number of attempts that preceded the first accepted one (NREJCT leaves those out, dopri5.f:528); NREJCT then
     * follows from NSTEP when the call returns. */
    static constexpr bool COUNT_IN_ACCEPT_ONLY = !AENS_STRICT && M::NG == 0 && !DENSE;
    /* called by the kernel whenever attempt() returned anything but CONTINUE, i.e. when dopri5() returns:
     * NSTEP / NACCPT / NREJCT / NFCN of this call -> the member's statistics (runge_kutta.py:183-186) */
    AENS_DEV void seg_end() {
        L.st[AENS_STAT_NSTEPSTOTAL] += nstep;
        L.st[AENS_STAT_NSTEPS] += naccpt;
        L.st[AENS_STAT_NERRFAILS] += COUNT_IN_ACCEPT_ONLY ? (naccpt ? nstep - naccpt - nrej : 0) : nrej;
        L.st[AENS_STAT_NFCNS] += 6 * nstep;
    }

    AENS_DEV double hinit(const aens_args_t &a) { /* HINIT with IORD=5, ITOL=1, POSNEG=+1 */
#if AENS_STRICT
        double dnf = 0.0, dny = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = a.o.atol[i] + a.o.rtol * dabs(L.y[i]);
            const double q1 = k1[i] / sk, q2 = L.y[i] / sk;
            dnf = dnf + q1 * q1;
            dny = dny + q2 * q2;
        }
        double hh;
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
        else hh = sqrt(dny / dnf) * 0.01;
        hh = dmin(hh, hmax);
        double y1[N], f1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = L.y[i] + hh * k1[i];
        L.rhs(L.t + hh, y1, f1);
        double der2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = a.o.atol[i] + a.o.rtol * dabs(L.y[i]);
            const double q = (f1[i] - k1[i]) / sk;
            der2 = der2 + q * q;
        }
        der2 = sqrt(der2) / hh;
        const double der12 = dmax(dabs(der2), sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = dmax(1.0e-6, dabs(hh) * 1.0e-3);
        else h1 = pow(0.01 / der12, 1.0 / 5);
        return dmin(dmin(100 * dabs(hh), h1), hmax);
#else
        /* FAST: the result is only the controller's starting guess (and HINIT runs again after every state event):
         * reciprocal seeds instead of divisions, fp32 MUFU square roots and fifth root */
        double dnf = 0.0, dny = 0.0, rsk[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            rsk[i] = rcp_seed(fma(a.o.rtol, dabs(L.y[i]), a.o.atol[i]));
            const double q1 = k1[i] * rsk[i], q2 = L.y[i] * rsk[i];
            dnf = fma(q1, q1, dnf);
            dny = fma(q2, q2, dny);
        }
        double hh;
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
        else hh = (double)(sqrtf((float)(dny * rcp_seed(dnf))) * 0.01f);
        hh = dmin(hh, hmax);
        double y1[N], f1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = fma(hh, k1[i], L.y[i]);
        L.rhs(L.t + hh, y1, f1);
        double der2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double q = (f1[i] - k1[i]) * rsk[i];
            der2 = fma(q, q, der2);
        }
        const double der12 = dmax(r_sqrt(der2) * rcp_seed(hh), r_sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = dmax(1.0e-6, dabs(hh) * 1.0e-3);
        else h1 = (double)fpow((float)(0.01 * rcp_seed(der12)), 0.2f);
        return dmin(dmin(100 * dabs(hh), h1), hmax);
#endif
    }

#if !AENS_STRICT
    /* dopri5.f:411-414 for the NEXT attempt, evaluated as soon as its step size is known: if x + 1.01 h would pass
     * tfinal the step goes exactly to tfinal and is the last one.  The bounds depend only on the end point xn of
     * the step in flight (rem, rem/1.01, hmax: ready long before the error norm), so that one compare + select
     * follow the controller on the loop-carried critical path.  `use_hmax`: dopri5.f:515 clamps accepted steps to
     * HMAX (after which the end-point test is repeated with the clamped value).  A rejected step never needs the
     * clamp: its hnew < h and h already satisfied the bounds at the same x. */
    struct Bounds { double rem, b101; };
    AENS_DEV Bounds next_bounds(const aens_args_t &a, double xn) const {
        Bounds b;
        b.rem = L.tend(a) - xn;
        b.b101 = b.rem * dp::T.inv101;
        return b;
    }
    /* Both magnitude tests look at the high words only (positive doubles order like their bit patterns; a tie of
     * the high words means the two agree to 2^-20): a step size that exceeds rem/1.01 or HMAX by less than that is
     * let through, which costs at most one extra short step at the very end. */
    AENS_DEV void clamp_next(const Bounds &b, double hnew, bool accept, bool use_hmax) {
        /* dopri5.f:515 first (accepted steps only), then the end-point test on the clamped value; all selects */
        const bool big = use_hmax && accept && (__double2hiint(hnew) > __double2hiint(hmax));
        const double h1 = big ? hmax : hnew;
        const bool over = accept && (__double2hiint(h1) > __double2hiint(b.b101));
        h = over ? b.rem : h1;
        last = over;
    }
    /* dopri5.f:409 `0.1 |h| <= |x| uround` (uround = 2.3e-16, i.e. |h| <= |x| 2^-48.6), tested on the exponents only:
     * |h| 2^50 <= |x| (1 + 2^-20) implies 0.1 |h| <= |x| uround, so every member flagged here is also flagged by the
     * reference; one whose h sits within a factor 2.6 above that is flagged an attempt or two later, after the
     * controller has shrunk h once more.  Keeps the test off the fp64 pipe. */
    AENS_DEV bool step_too_small(double x) const {
        /* high words as integers: hi(h 2^50) = hi(h) + (50 << 20) exactly, and the high word is monotone in the
         * magnitude, so this is |h| 2^50 <= |x| decided on the top 52 bits (h > 0: forward integration only) */
        return __double2hiint(h) + (50 << 20) <= (__double2hiint(x) & 0x7fffffff);
    }
#endif

    AENS_DEV void seg_init(const aens_args_t &a, long long idx) {
        seg_init_events(L);
        hmax = dabs((a.o.maxh == 0.0) ? (L.tend(a) - L.t) : a.o.maxh); /* dopri5.f:301-305, :392 */
#if AENS_STRICT
        facold = 1.0e-4;
#else
        lfacold = a.ctl[5];
        lfmin = a.ctl[4];
#endif
        last = false;
        reject = false;
        nstep = 0;
        naccpt = 0;
        nrej = 0;
        L.rhs(L.t, L.y, k1);
        h = a.o.inith;
        if (h == 0.0) h = hinit(a);
#if !AENS_STRICT
        clamp_next(next_bounds(a, L.t), h, true, false);
#endif
        L.st[AENS_STAT_NFCNS] += 2;
        /* first SOLOUT call (XOLD = X, dopri5.f:399-404): Assimulo runs the locator on a zero-length interval --
         * one more g evaluation at the very same point, never an event -- and reports grid rows <= t */
        if (M::NG > 0) L.st[AENS_STAT_NSTATEFCNS] += 1;
        if (DENSE) aens_emit_start<M, Dopri5>(L, a, idx);
    }

    AENS_DEV int attempt(const aens_args_t &a, long long idx) {
#if AENS_STRICT
        return attempt_ref(a, idx);
#else
        return attempt_fast(a, idx);
#endif
    }

#if AENS_STRICT
    /* DOPCOR in the reference's operation order (compiled with --fmad=false) */
    AENS_DEV int attempt_ref(const aens_args_t &a, long long idx) {
        using dp::T;
        using dp::uround;
        if (nstep > a.o.maxsteps) { L.status = AENS_ST_MAXSTEPS; return AENS_R_ERROR; }
        const double x = L.t;
        if (0.1 * dabs(h) <= dabs(x) * uround) { L.status = AENS_ST_STEP_TOO_SMALL; return AENS_R_ERROR; }
        if ((x + 1.01 * h - L.tend(a)) > 0.0) { h = L.tend(a) - x; last = true; }
        nstep += 1;
        double k2[N], k3[N], k4[N], k5[N], k6[N], y1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = L.y[i] + h * T.a21 * k1[i];
        L.rhs(x + T.c2 * h, y1, k2);
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = L.y[i] + h * (T.a31 * k1[i] + T.a32 * k2[i]);
        L.rhs(x + T.c3 * h, y1, k3);
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = L.y[i] + h * (T.a41 * k1[i] + T.a42 * k2[i] + T.a43 * k3[i]);
        L.rhs(x + T.c4 * h, y1, k4);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y1[i] = L.y[i] + h * (T.a51 * k1[i] + T.a52 * k2[i] + T.a53 * k3[i] + T.a54 * k4[i]);
        L.rhs(x + T.c5 * h, y1, k5);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y1[i] = L.y[i] + h * (T.a61 * k1[i] + T.a62 * k2[i] + T.a63 * k3[i] + T.a64 * k4[i] + T.a65 * k5[i]);
        const double xph = x + h;
        L.rhs(xph, y1, k6);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y1[i] = L.y[i] + h * (T.a71 * k1[i] + T.a73 * k3[i] + T.a74 * k4[i] + T.a75 * k5[i] + T.a76 * k6[i]);
        L.rhs(xph, y1, k2); /* FSAL stage kept in K2 */
        double c4v[DENSE ? N : 1];
        double err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (DENSE)
                c4v[i] = h * (T.d1 * k1[i] + T.d3 * k3[i] + T.d4 * k4[i] + T.d5 * k5[i] + T.d6 * k6[i] + T.d7 * k2[i]);
            const double ei = (T.e1 * k1[i] + T.e3 * k3[i] + T.e4 * k4[i] + T.e5 * k5[i] + T.e6 * k6[i] + T.e7 * k2[i]) * h;
            const double sk = a.o.atol[i] + a.o.rtol * dmax(dabs(L.y[i]), dabs(y1[i]));
            const double q = ei / sk;
            err2 = err2 + q * q;
        }
        const double err = sqrt(err2 / N);
        const double expo1 = 0.2 - a.o.beta * 0.75;
        const double facc1 = 1.0 / a.o.fac1, facc2 = 1.0 / a.o.fac2;
        const double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, a.o.beta);
        fac = dmax(facc2, dmin(facc1, fac / a.o.safe));
        double hnew = h / fac;
        const bool accept = (err <= 1.0);
        if (accept) facold = dmax(err, 1.0e-4);
        else hnew = h / dmin(facc1, fac11 / a.o.safe);
        if (accept) {
            naccpt += 1;
            /* (stiffness detection dopri5.f:474-494 only prints under Assimulo's IPRINT: omitted) */
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (DENSE) {
                    const double ydiff = y1[i] - L.y[i];
                    const double bspl = h * k1[i] - ydiff;
                    cont[i] = L.y[i];
                    cont[N + i] = ydiff;
                    cont[2 * N + i] = bspl;
                    cont[3 * N + i] = -h * k2[i] + ydiff - bspl;
                    cont[4 * N + i] = c4v[i];
                }
                k1[i] = k2[i];
                L.y[i] = y1[i];
            }
            told = x;
            hstep = h;
            const int flag = aens_post_step<M, Dopri5>(*this, L, a, idx, x, xph);
            if (flag == AENS_R_EVENT) return AENS_R_EVENT;
            if (last) return AENS_R_COMPLETE;
            if (dabs(hnew) > hmax) hnew = hmax;
            if (reject) hnew = dmin(dabs(hnew), dabs(h));
            reject = false;
        } else {
            reject = true;
            if (naccpt >= 1) nrej += 1;
            last = false;
        }
        h = hnew;
        return AENS_R_CONTINUE;
    }
#else
    /* DOPCOR re-associated for the fp64 pipe of sm_100a (measured, profiles/r1_fp64_pipe.md: a DFMA with three
     * distinct register operands issues at 11/12 rate with a 12-cycle dependent latency, one with a constant-bank
     * operand at full rate with 8 cycles):
     *  - small systems accumulate every later stage argument as soon as h*k_j exists,
     *        A_s += a_sj * (h k_j)  for all s > j,   E += e_j (h k_j),   D += d_j (h k_j)
     *    -- the same multiply-add count as the row-by-row sums of dopri5.f:421-448, but each is a
     *    constant-bank-operand DFMA, they are mutually independent, and only ONE of them separates two
     *    consecutive right-hand-side evaluations on the critical path;
     *  - the error norm scales with the MUFU.RCP64H seed of 1/sk (2^-23 relative: a perturbation of the
     *    tolerance, not of the solution); the accept test err^2 <= 1 itself is evaluated in fp64;
     *  - Hairer's PI controller runs in the log2 domain in fp32 (one MUFU.LG2 + one MUFU.EX2 per attempt);
     *  - accept/reject commits by select, the end-point clamp of the next step is hoisted (clamp_next). */
    AENS_DEV int attempt_fast(const aens_args_t &a, long long idx) {
        using dp::T;
        using dp::uround;
        const double x = L.t;
        if ((nstep > a.o.maxsteps) | step_too_small(x)) return AENS_R_ERROR; /* status: late_status() */
        nstep += 1;
        /* the last step lands on the end point exactly: x + (xend - x) can miss xend by an ulp, which the
         * reference tolerates at tfinal (100 eps slack of the driver loop) but not at a time event, where the next
         * segment would then be one ulp long */
        const double xph = last ? L.tend(a) : x + h;
        const Bounds nb = next_bounds(a, xph);
        double y1[N], k2[N], E[N], hk1[N], hk7[DENSE ? N : 1], D[DENSE ? N : 1];
        if constexpr (INCREMENTAL) {
            double A3[N], A4[N], A5[N], A6[N], hk[N], kk[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                hk1[i] = h * k1[i];
                kk[i] = fma(T.a21, hk1[i], L.y[i]); /* stage-2 argument */
                A3[i] = fma(T.a31, hk1[i], L.y[i]);
                A4[i] = fma(T.a41, hk1[i], L.y[i]);
                A5[i] = fma(T.a51, hk1[i], L.y[i]);
                A6[i] = fma(T.a61, hk1[i], L.y[i]);
                y1[i] = fma(T.a71, hk1[i], L.y[i]);
                E[i] = T.e1 * hk1[i];
                if (DENSE) D[i] = T.d1 * hk1[i];
            }
            L.rhs_h(h, fma(T.c2, h, x), kk, hk);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                A3[i] = fma(T.a32, hk[i], A3[i]);
                A4[i] = fma(T.a42, hk[i], A4[i]);
                A5[i] = fma(T.a52, hk[i], A5[i]);
                A6[i] = fma(T.a62, hk[i], A6[i]);
            }
            L.rhs_h(h, fma(T.c3, h, x), A3, hk);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                A4[i] = fma(T.a43, hk[i], A4[i]);
                A5[i] = fma(T.a53, hk[i], A5[i]);
                A6[i] = fma(T.a63, hk[i], A6[i]);
                y1[i] = fma(T.a73, hk[i], y1[i]);
                E[i] = fma(T.e3, hk[i], E[i]);
                if (DENSE) D[i] = fma(T.d3, hk[i], D[i]);
            }
            L.rhs_h(h, fma(T.c4, h, x), A4, hk);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                A5[i] = fma(T.a54, hk[i], A5[i]);
                A6[i] = fma(T.a64, hk[i], A6[i]);
                y1[i] = fma(T.a74, hk[i], y1[i]);
                E[i] = fma(T.e4, hk[i], E[i]);
                if (DENSE) D[i] = fma(T.d4, hk[i], D[i]);
            }
            L.rhs_h(h, fma(T.c5, h, x), A5, hk);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                A6[i] = fma(T.a65, hk[i], A6[i]);
                y1[i] = fma(T.a75, hk[i], y1[i]);
                E[i] = fma(T.e5, hk[i], E[i]);
                if (DENSE) D[i] = fma(T.d5, hk[i], D[i]);
            }
            L.rhs_h(h, xph, A6, hk);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y1[i] = fma(T.a76, hk[i], y1[i]);
                E[i] = fma(T.e6, hk[i], E[i]);
                if (DENSE) D[i] = fma(T.d6, hk[i], D[i]);
            }
            L.rhs(xph, y1, k2); /* FSAL stage */
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double hk2 = h * k2[i];
                E[i] = fma(T.e7, hk2, E[i]);
                if (DENSE) { D[i] = fma(T.d7, hk2, D[i]); hk7[i] = hk2; }
            }
        } else {
            /* larger systems: N-wide instruction-level parallelism already; row sums keep the live set small */
            double k3[N], k4[N], k5[N], k6[N];
#pragma unroll
            for (int i = 0; i < N; ++i) y1[i] = fma(h * T.a21, k1[i], L.y[i]);
            L.rhs(fma(T.c2, h, x), y1, k2);
#pragma unroll
            for (int i = 0; i < N; ++i) y1[i] = fma(h, fma(T.a32, k2[i], T.a31 * k1[i]), L.y[i]);
            L.rhs(fma(T.c3, h, x), y1, k3);
#pragma unroll
            for (int i = 0; i < N; ++i) y1[i] = fma(h, fma(T.a43, k3[i], fma(T.a42, k2[i], T.a41 * k1[i])), L.y[i]);
            L.rhs(fma(T.c4, h, x), y1, k4);
#pragma unroll
            for (int i = 0; i < N; ++i)
                y1[i] = fma(h, fma(T.a54, k4[i], fma(T.a53, k3[i], fma(T.a52, k2[i], T.a51 * k1[i]))), L.y[i]);
            L.rhs(fma(T.c5, h, x), y1, k5);
#pragma unroll
            for (int i = 0; i < N; ++i)
                y1[i] = fma(h, fma(T.a65, k5[i], fma(T.a64, k4[i], fma(T.a63, k3[i], fma(T.a62, k2[i], T.a61 * k1[i])))),
                            L.y[i]);
            L.rhs(xph, y1, k6);
#pragma unroll
            for (int i = 0; i < N; ++i)
                y1[i] = fma(h, fma(T.a76, k6[i], fma(T.a75, k5[i], fma(T.a74, k4[i], fma(T.a73, k3[i], T.a71 * k1[i])))),
                            L.y[i]);
            L.rhs(xph, y1, k2); /* FSAL stage kept in K2 */
#pragma unroll
            for (int i = 0; i < N; ++i) {
                E[i] = h * fma(T.e7, k2[i], fma(T.e6, k6[i], fma(T.e5, k5[i], fma(T.e4, k4[i], fma(T.e3, k3[i], T.e1 * k1[i])))));
                if (DENSE) {
                    D[i] = h * fma(T.d7, k2[i], fma(T.d6, k6[i], fma(T.d5, k5[i], fma(T.d4, k4[i], fma(T.d3, k3[i], T.d1 * k1[i])))));
                    hk1[i] = h * k1[i];
                    hk7[i] = h * k2[i];
                }
            }
        }
        /* error norm dopri5.f:451-461 on the squared scale: err^2 = (1/N) sum (E_i / sk_i)^2 */
        double err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = fma(a.o.rtol, dabs(absmax_hi(L.y[i], y1[i])), a.o.atol[i]);
            const double q = E[i] * rcp_seed(sk);
            err2 = fma(q, q, err2);
        }
        const bool accept = (err2 <= (double)N);
        /* Hairer's PI controller in the log2 domain (fac = 2^lf, hnew = h 2^-lf), dopri5.f:463-468,527-529:
         *   accept:  lf = clamp(expo1 lg2 err - beta lg2 facold - lg2 safe, lg2 facc2, lg2 facc1)
         *   reject:  lf = min(lg2 facc1, expo1 lg2 err - lg2 safe)
         * and 'a step after a rejection does not grow' (:516) is lf >= 0. */
        constexpr float LG2N = lg2_int(N);
        const float le2 = flg2((float)err2) - LG2N;
        const float l11 = fmaf(a.ctl[0], le2, a.ctl[2]);
        float lf = fminf(a.ctl[3], accept ? l11 - lfacold : l11);
        if (accept) {
            lf = fmaxf(lf, lfmin);
            lfacold = a.ctl[1] * fmaxf(le2, a.ctl[6]); /* FACOLD = max(err, 1e-4) */
        }
        const double hnew = h * (double)fex2(-lf);
        const int was_last = last;
        const double h_step = h;
        clamp_next(nb, hnew, accept, true);
        if constexpr (M::NG == 0 && !DENSE) {
            /* final-state-only members: one short accept block (the compiler branches around the register moves of a
             * select-commit anyway), nothing on the reject side */
            if (accept) {
                nrej = naccpt ? nrej : nstep - 1; /* see COUNT_IN_ACCEPT_ONLY */
                naccpt += 1;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    k1[i] = k2[i];
                    L.y[i] = y1[i];
                }
                L.t = xph;
            }
            lfmin = accept ? a.ctl[4] : 0.0f;
            return (accept && was_last) ? AENS_R_COMPLETE : AENS_R_CONTINUE;
        } else {
            if (accept) {
                naccpt += 1;
                if (DENSE) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double ydiff = y1[i] - L.y[i];
                        const double bspl = hk1[i] - ydiff;
                        cont[i] = L.y[i];
                        cont[N + i] = ydiff;
                        cont[2 * N + i] = bspl;
                        cont[3 * N + i] = ydiff - hk7[i] - bspl;
                        cont[4 * N + i] = D[i];
                    }
                    told = x;
                    hstep = r_recip(h_step);
                }
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    k1[i] = k2[i];
                    L.y[i] = y1[i];
                }
                lfmin = a.ctl[4];
                const int flag = aens_post_step<M, Dopri5>(*this, L, a, idx, x, xph);
                if (flag == AENS_R_EVENT) return AENS_R_EVENT;
                if (was_last) return AENS_R_COMPLETE;
            } else {
                lfmin = 0.0f;
                if (naccpt >= 1) nrej += 1;
            }
            return AENS_R_CONTINUE;
        }
    }
#endif
};

}  // namespace aens

/* ------------------------------------------------------------------------------------------------------- */
/* The ensemble kernel: persistent warps pull members from an atomic queue; every lane runs the phase        */
/* machine  FETCH -> SEGINIT -> STEP* -> (EVENT -> handle_event -> SEGINIT | DONE -> STORE -> FETCH).          */
/* ------------------------------------------------------------------------------------------------------- */
template <class S, class M>
__device__ __forceinline__ void aens_run(const aens_args_t &a) {
    enum { PH_IDLE = 0, PH_SEGINIT = 1, PH_STEP = 2 };
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double eps = 2.220446049250313e-16 * 100; /* explicit_ode.pyx:184 */
    S s;
    if constexpr (S::REDUCE) { /* the per-warp locks of aens_reduce_grid */
        if (threadIdx.x < AENS_BLOCK / 32) aens::aens_red_lock()[threadIdx.x] = 0;
        __syncthreads();
    }
    long long idx = -1;
    int phase = PH_IDLE;
    bool te_at_end = false; /* models with time events: the current segment ends on a time event */
    bool pending = false; /* a finished member waits in this lane's registers for the warp's next common store point */
    bool drained = false; /* warp-uniform: the queue has no more members */
    for (;;) {
        const unsigned idle_mask = __ballot_sync(FULL, phase == PH_IDLE);
        if (idle_mask == FULL && drained) break;
        if (!drained && (idle_mask == FULL || __popc(idle_mask) >= a.refill_threshold)) {
            const int nidle = __popc(idle_mask);
            const int leader = __ffs(idle_mask) - 1;
            unsigned long long base = 0;
            int late = 0; /* warp-uniform: this batch is fetched after the call's time limit */
            if (lane == leader) {
                base = atomicAdd(a.queue, (unsigned long long)nidle);
                if (a.limit_ns) { /* ODE.time_limit: checked when a warp asks for work, i.e. between members */
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    unsigned long long t_start = atomicCAS(a.t_start, 0ull, now);
                    if (t_start == 0ull) t_start = now;
                    late = (now - t_start > a.limit_ns) ? 1 : 0;
                }
            }
            base = __shfl_sync(FULL, base, leader);
            late = __shfl_sync(FULL, late, leader);
            if (phase == PH_IDLE) {
                /* results leave the registers here, for all idle lanes of the warp at once: lanes that were fetched
                 * together hold consecutive members, so their rows form contiguous blocks (SoA state in HBM and, in
                 * zero-copy mode, the caller's member-major host arrays) */
                if (pending) {
                    s.L.template store<S::HAS_DENSE && !S::REDUCE>(a, idx);
                    s.store_aux(a, idx);
                    pending = false;
                }
                const long long q = a.q_begin + (long long)base + __popc(idle_mask & ((1u << lane) - 1u));
                if (q < a.q_end) {
                    if (a.order) {
                        idx = (long long)a.order[q];
                        AENS_CHK(a, idx >= 0 && idx < a.N, 9);
                    } else if (a.reverse == 2) {
                        /* two-ended: 32-member batches alternately from the top and from the bottom of the range, so
                         * that an ensemble sorted by cost is consumed at an even rate from both ends (evens out the
                         * PCIe demand of the zero-copy host mode); a bijection on the range for any size */
                        const long long qq = q - a.q_begin, b = qq >> 5, r = qq & 31, k = b >> 1;
                        idx = (b & 1) ? a.q_begin + (k << 5) + r : a.q_end - 1 - ((k << 5) + r);
                    } else {
                        idx = a.reverse ? a.q_end - 1 - (q - a.q_begin) : q;
                    }
                    s.L.load(a, idx);
                    s.load_aux(a, idx);
                    /* driver loop test, explicit_ode.pyx:209 */
                    if (s.L.t + eps * aens::dmax(aens::dabs(s.L.t), aens::dabs(a.tf)) < a.tf) {
                        if (late) { /* TimeLimitExceeded (explicit_ode.pyx:290-292): stored as fetched, not integrated */
                            s.L.status = AENS_ST_TIME_LIMIT;
                            pending = true;
                        } else {
                            phase = PH_SEGINIT;
                        }
                    } else {
                        pending = true;
                    }
                }
            }
            if (a.q_begin + (long long)base + nidle >= a.q_end) drained = true;
        }
        if (phase == PH_SEGINIT) {
            if constexpr (M::HAS_TE) { /* tevent = min(time_events(t, y, sw), tfinal), explicit_ode.pyx:212-216 */
                const double tret = M::time_events(s.L.t, s.L.y, s.L.sw, s.L.p);
                s.L.tnext = (tret < a.tf) ? tret : a.tf;
                te_at_end = (tret == s.L.tnext); /* a time event falls on the end of this segment */
            }
            s.seg_init(a, idx);
            phase = PH_STEP;
        }
        /* lanes that are stepping stay in a tight attempt loop until one of them leaves the STEP phase */
        const unsigned stepmask = __ballot_sync(FULL, phase == PH_STEP);
        if (phase == PH_STEP) {
            int r;
            do {
                r = s.attempt(a, idx);
            } while (__all_sync(stepmask, r == AENS_R_CONTINUE));
            if (r != AENS_R_CONTINUE) {
                if (r == AENS_R_ERROR) s.late_status(a);
                s.seg_end();
                bool done = true;
                if (r == AENS_R_EVENT) {
                    /* explicit_ode.pyx:233-264: log, handle_event, restart the solver (initialize=True) */
                    const int ne = s.L.st[AENS_STAT_NSTATEEVENTS] + s.L.st[AENS_STAT_NTIMEEVENTS] - 1;
                    if (a.ev_slots > 0) {
                        int mask = 0;
#pragma unroll
                        for (int i = 0; i < M::NG; ++i)
                            if (s.L.event_info[i] != 0)
                                mask |= (1 << (2 * i)) | ((s.L.event_info[i] > 0 ? 1 : 0) << (2 * i + 1));
                        const long long slot = (long long)(ne % a.ev_slots) * a.N + idx;
                        AENS_CHK(a, ne >= 0 && slot >= 0 && slot < (long long)a.ev_slots * a.N && a.ev_info != nullptr, 10);
                        a.ev_t[slot] = M::REV ? -s.L.t : s.L.t;
                        a.ev_info[slot] = mask;
                    }
                    const int term = M::handle_event(s.L.t, s.L.y, s.L.sw, s.L.event_info, s.L.p);
                    if (term) s.L.status = AENS_ST_TERMINATED;
                    else if (s.L.t != a.tf &&
                             s.L.t + eps * aens::dmax(aens::dabs(s.L.t), aens::dabs(a.tf)) < a.tf) {
                        phase = PH_SEGINIT;
                        done = false;
                    }
                } else if (r == AENS_R_COMPLETE) {
                    if constexpr (M::HAS_TE) {
                        /* explicit_ode.pyx:233-264 with flag == ID_COMPLETE: (tevent != tfinal) or (tret == tevent) */
                        if (s.L.tnext != a.tf || te_at_end) {
                            s.L.st[AENS_STAT_NTIMEEVENTS] += 1;
                            if (a.ev_slots > 0) {
                                const int ne = s.L.st[AENS_STAT_NSTATEEVENTS] + s.L.st[AENS_STAT_NTIMEEVENTS] - 1;
                                const long long slot = (long long)(ne % a.ev_slots) * a.N + idx;
                                a.ev_t[slot] = s.L.t;
                                a.ev_info[slot] = 1 << 30;
                            }
                            if (M::handle_time_event(s.L.t, s.L.y, s.L.sw, s.L.p)) s.L.status = AENS_ST_TERMINATED;
                        }
                    }
                    if (s.L.status != AENS_ST_TERMINATED && s.L.t != a.tf &&
                        s.L.t + eps * aens::dmax(aens::dabs(s.L.t), aens::dabs(a.tf)) < a.tf) {
                        phase = PH_SEGINIT; /* integrate() returned short of tfinal: the driver calls it again */
                        done = false;
                    }
                }
                if (done) {
                    pending = true;
                    phase = PH_IDLE;
                }
            }
        }
    }
    if (pending) { /* the warp's last members: every lane is idle here */
        s.L.template store<S::HAS_DENSE && !S::REDUCE>(a, idx);
        s.store_aux(a, idx);
    }
}

/* DENSE: 0 = no interpolation data kept, 1 = dense output (grid rows stored, event location), 2 = as 1 but the grid
 * rows are reduced over the ensemble (aens_reduce_grid) */
#define AENS_DEFINE_KERNEL(NAME, STEPPER, MODEL, DENSE, MINBLOCKS)                                   \
    extern "C" __global__ void __launch_bounds__(AENS_BLOCK, MINBLOCKS) NAME(const aens_args_t a) {  \
        aens_run<aens::STEPPER<MODEL, DENSE>, MODEL>(a);                                             \
    }

#endif /* AENS_DEVICE_CUH */
