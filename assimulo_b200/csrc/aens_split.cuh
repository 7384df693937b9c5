/*
 * aens_split.cuh -- sub-warp-per-trajectory variant for larger state dimensions (north_star: "one trajectory per thread
 * (or per warp for larger state dimension)").
 *
 * One thread per member keeps 14 n doubles of DOPRI5 state in registers (dopri5.f:367-369: Y, Y1, K1..K6, YSTI, CONT):
 * for the 12-state gyroscope that is ~240 registers per thread, two CTAs per SM, eight warps to hide the 8-12 cycle
 * latency of dependent DFMAs.  Here G = M::G consecutive lanes of a warp share ONE member, each owning NL = M::NL of
 * its components for the whole integration (register use falls by ~G, resident warps rise accordingly); the lanes of a
 * group exchange data only where the mathematics couples components:
 *   - the right-hand side: M::rhs_g() fetches what it needs from its group with warp shuffles (the gyroscope: the three
 *     angular velocities owned by lane 0; every lane then rotates its own 3-vector -- the same code in all four lanes);
 *   - the error norm and HINIT's norms: an xor-butterfly over the group (every lane ends up with the identical sum, so
 *     the step-size control is evaluated redundantly, and identically, by the G lanes -- no broadcast, no divergence).
 * Everything else -- stage arithmetic, dense-output coefficients, interpolation, stores -- is per component and
 * needs no communication.  The driver is the phase machine of aens_run with groups in the place of lanes.
 *
 * Scope: Dopri5 (dopri5.f:356-556 DOPCOR, :558-619 HINIT, :621-644 CONTD5), FAST arithmetic, models without state or
 * time events, final states (d0) or stored dense-output rows (d1).  The summation order of the norms differs from the
 * reference's (component blocks instead of 1..n), which is why the STRICT build, whose contract is step-for-step
 * identity with the oracle, keeps one thread per member.  Parity of this variant: tests/test_gpu_round2.py
 * (against the oracle within 10 x rtol, step totals within 2 %, and against the thread-per-member kernel).
 */
#ifndef AENS_SPLIT_CUH
#define AENS_SPLIT_CUH

#include "aens_device.cuh"

#if !AENS_STRICT

namespace aens {

/* sum over the G lanes of a group (G a power of two, groups aligned): identical in every lane of the group */
template <int G>
AENS_DEV double group_sum(unsigned mask, double v) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(mask, v, o, 32);
    return v;
}

template <class M, int DENSE>
struct Dopri5Split {
    static constexpr int G = M::G, NL = M::NL, N = M::N;
    static constexpr bool HAS_DENSE = DENSE != 0;
    int j;                 /* this lane's position in its group */
    double t, y[NL], p[M::NP > 0 ? M::NP : 1], k1[NL], atl[NL];
    double h, hmax;
    float lfacold, lfmin;
    double cont[DENSE ? 5 * NL : 1], told, rh;
    int nstep, naccpt, nrej, last;
    int st_nsteps, st_nfcns, st_nerrfails, st_total, gi, status;

    AENS_DEV void rhs(unsigned gm, double tt, const double *v, double *vd) { M::rhs_g(gm, j, tt, v, p, vd); }

    AENS_DEV void load(const aens_args_t &a, long long idx) {
        t = a.t[idx];
#pragma unroll
        for (int c = 0; c < NL; ++c) y[c] = a.y[(long long)M::comp(j, c) * a.N + idx];
#pragma unroll
        for (int i = 0; i < M::NP; ++i) p[i] = a.p[(long long)i * a.N + idx];
        M::prep(p);
#pragma unroll
        for (int c = 0; c < NL; ++c) atl[c] = a.o.atol[M::comp(j, c)]; /* (a register-indexed constant-bank read otherwise) */
        gi = a.n_grid > 0 ? a.grid_idx[idx] : 0;
        st_nsteps = st_nfcns = st_nerrfails = st_total = 0;
        status = AENS_ST_COMPLETE;
    }
    AENS_DEV void store(const aens_args_t &a, long long idx) {
        if (DENSE && a.dense) { /* rows this member never reached read NaN (see Lane::store) */
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            for (int k = gi; k < a.n_grid; ++k) {
#pragma unroll
                for (int c = 0; c < NL; ++c) a.dense[((long long)k * N + M::comp(j, c)) * a.N + idx] = qnan;
            }
        }
#pragma unroll
        for (int c = 0; c < NL; ++c) a.y[(long long)M::comp(j, c) * a.N + idx] = y[c];
        if (j == 0) { /* one lane of the group writes the member's scalars */
            a.t[idx] = t;
            a.status[idx] = status;
            if (a.n_grid > 0) a.grid_idx[idx] = gi;
#pragma unroll
            for (int i = 0; i < AENS_NSTATS; ++i) a.stats[(long long)i * a.N + idx] = 0;
            a.stats[(long long)AENS_STAT_NSTEPS * a.N + idx] = st_nsteps;
            a.stats[(long long)AENS_STAT_NFCNS * a.N + idx] = st_nfcns;
            a.stats[(long long)AENS_STAT_NERRFAILS * a.N + idx] = st_nerrfails;
            a.stats[(long long)AENS_STAT_NSTEPSTOTAL * a.N + idx] = st_total;
        }
    }

    AENS_DEV void interp(double x, double *out) const { /* CONTD5 */
        if (DENSE) {
            const double th = (x - told) * rh, th1 = 1.0 - th;
#pragma unroll
            for (int c = 0; c < NL; ++c)
                out[c] = cont[c] + th * (cont[NL + c] + th1 * (cont[2 * NL + c] + th * (cont[3 * NL + c] + th1 * cont[4 * NL + c])));
        }
    }
    AENS_DEV void emit(const aens_args_t &a, long long idx, double tt, bool from_state) {
        while (gi < a.n_grid && __ldg(a.t_grid + gi) <= tt) {
            double out[NL];
            if (from_state) {
#pragma unroll
                for (int c = 0; c < NL; ++c) out[c] = y[c];
            } else {
                interp(__ldg(a.t_grid + gi), out);
            }
#pragma unroll
            for (int c = 0; c < NL; ++c) a.dense[((long long)gi * N + M::comp(j, c)) * a.N + idx] = out[c];
            gi += 1;
        }
    }

    struct Bounds { double rem, b101; };
    AENS_DEV Bounds next_bounds(const aens_args_t &a, double xn) const {
        Bounds b;
        b.rem = a.tf - xn;
        b.b101 = b.rem * dp::T.inv101;
        return b;
    }
    AENS_DEV void clamp_next(const Bounds &b, double hnew, bool accept, bool use_hmax) { /* see Dopri5::clamp_next */
        const bool big = use_hmax && accept && (__double2hiint(hnew) > __double2hiint(hmax));
        const double h1 = big ? hmax : hnew;
        const bool over = accept && (__double2hiint(h1) > __double2hiint(b.b101));
        h = over ? b.rem : h1;
        last = over;
    }
    AENS_DEV bool step_too_small(double x) const {
        return __double2hiint(h) + (50 << 20) <= (__double2hiint(x) & 0x7fffffff);
    }

    /* HINIT (dopri5.f:558-619): the three norms are sums over all n components, i.e. over the group */
    AENS_DEV double hinit(unsigned gm, const aens_args_t &a) {
        double dnf = 0.0, dny = 0.0, rsk[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            rsk[c] = rcp_seed(fma(a.o.rtol, dabs(y[c]), atl[c]));
            const double q1 = k1[c] * rsk[c], q2 = y[c] * rsk[c];
            dnf = fma(q1, q1, dnf);
            dny = fma(q2, q2, dny);
        }
        dnf = group_sum<G>(gm, dnf);
        dny = group_sum<G>(gm, dny);
        double hh;
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) hh = 1.0e-6;
        else hh = (double)(sqrtf((float)(dny * rcp_seed(dnf))) * 0.01f);
        hh = dmin(hh, hmax);
        double y1[NL], f1[NL];
#pragma unroll
        for (int c = 0; c < NL; ++c) y1[c] = fma(hh, k1[c], y[c]);
        rhs(gm, t + hh, y1, f1);
        double der2 = 0.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            const double q = (f1[c] - k1[c]) * rsk[c];
            der2 = fma(q, q, der2);
        }
        der2 = group_sum<G>(gm, der2);
        const double der12 = dmax(r_sqrt(der2) * rcp_seed(hh), r_sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = dmax(1.0e-6, dabs(hh) * 1.0e-3);
        else h1 = (double)fpow((float)(0.01 * rcp_seed(der12)), 0.2f);
        return dmin(dmin(100 * dabs(hh), h1), hmax);
    }

    /* gm: the lanes that execute this call together (whole groups) */
    AENS_DEV void seg_init(unsigned gm, const aens_args_t &a, long long idx) {
        hmax = dabs((a.o.maxh == 0.0) ? (a.tf - t) : a.o.maxh);
        lfacold = a.ctl[5];
        lfmin = a.ctl[4];
        last = 0;
        nstep = naccpt = nrej = 0;
        rhs(gm, t, y, k1);
        h = a.o.inith;
        if (h == 0.0) h = hinit(gm, a); /* (uniform over the ensemble: no divergence around the shuffles) */
        clamp_next(next_bounds(a, t), h, true, false);
        st_nfcns += 2;
        if (DENSE) emit(a, idx, t, true);
    }
    AENS_DEV void seg_end() {
        st_total += nstep;
        st_nsteps += naccpt;
        st_nerrfails += nrej;
        st_nfcns += 6 * nstep;
    }
    AENS_DEV void late_status(const aens_args_t &a) {
        if (status == AENS_ST_COMPLETE) status = (nstep > a.o.maxsteps) ? AENS_ST_MAXSTEPS : AENS_ST_STEP_TOO_SMALL;
    }

    /* one DOPCOR attempt (Dopri5::attempt_fast, row-sum form) on this lane's NL components */
    AENS_DEV int attempt(unsigned gm, const aens_args_t &a, long long idx) {
        using dp::T;
        const double x = t;
        /* dopri5.f:405-410.  No early return: the other groups of the warp are about to shuffle with this one named in
         * their mask, so a failing group walks through the attempt (its result is discarded below) */
        const bool bad = (nstep > a.o.maxsteps) | step_too_small(x);
        nstep += bad ? 0 : 1;
        const double xph = last ? a.tf : x + h;
        const Bounds nb = next_bounds(a, xph);
        double y1[NL], k2[NL], k3[NL], k4[NL], k5[NL], k6[NL], E[NL], D[DENSE ? NL : 1];
#pragma unroll
        for (int c = 0; c < NL; ++c) y1[c] = fma(h * T.a21, k1[c], y[c]);
        rhs(gm, fma(T.c2, h, x), y1, k2);
#pragma unroll
        for (int c = 0; c < NL; ++c) y1[c] = fma(h, fma(T.a32, k2[c], T.a31 * k1[c]), y[c]);
        rhs(gm, fma(T.c3, h, x), y1, k3);
#pragma unroll
        for (int c = 0; c < NL; ++c) y1[c] = fma(h, fma(T.a43, k3[c], fma(T.a42, k2[c], T.a41 * k1[c])), y[c]);
        rhs(gm, fma(T.c4, h, x), y1, k4);
#pragma unroll
        for (int c = 0; c < NL; ++c)
            y1[c] = fma(h, fma(T.a54, k4[c], fma(T.a53, k3[c], fma(T.a52, k2[c], T.a51 * k1[c]))), y[c]);
        rhs(gm, fma(T.c5, h, x), y1, k5);
#pragma unroll
        for (int c = 0; c < NL; ++c)
            y1[c] = fma(h, fma(T.a65, k5[c], fma(T.a64, k4[c], fma(T.a63, k3[c], fma(T.a62, k2[c], T.a61 * k1[c])))), y[c]);
        rhs(gm, xph, y1, k6);
#pragma unroll
        for (int c = 0; c < NL; ++c)
            y1[c] = fma(h, fma(T.a76, k6[c], fma(T.a75, k5[c], fma(T.a74, k4[c], fma(T.a73, k3[c], T.a71 * k1[c])))), y[c]);
        rhs(gm, xph, y1, k2); /* FSAL stage kept in K2 */
        double err2 = 0.0;
#pragma unroll
        for (int c = 0; c < NL; ++c) {
            E[c] = h * fma(T.e7, k2[c], fma(T.e6, k6[c], fma(T.e5, k5[c], fma(T.e4, k4[c], fma(T.e3, k3[c], T.e1 * k1[c])))));
            if (DENSE)
                D[c] = h * fma(T.d7, k2[c], fma(T.d6, k6[c], fma(T.d5, k5[c], fma(T.d4, k4[c], fma(T.d3, k3[c], T.d1 * k1[c])))));
            const double sk = fma(a.o.rtol, dabs(absmax_hi(y[c], y1[c])), atl[c]);
            const double q = E[c] * rcp_seed(sk);
            err2 = fma(q, q, err2);
        }
        err2 = group_sum<G>(gm, err2); /* identical in the G lanes: everything below is evaluated redundantly */
        if (bad) return AENS_R_ERROR;
        const bool accept = (err2 <= (double)N);
        constexpr float LG2N = lg2_int(N);
        const float le2 = flg2((float)err2) - LG2N;
        const float l11 = fmaf(a.ctl[0], le2, a.ctl[2]);
        float lf = fminf(a.ctl[3], accept ? l11 - lfacold : l11);
        if (accept) {
            lf = fmaxf(lf, lfmin);
            lfacold = a.ctl[1] * fmaxf(le2, a.ctl[6]);
        }
        const double hnew = h * (double)fex2(-lf);
        const int was_last = last;
        const double h_step = h;
        clamp_next(nb, hnew, accept, true);
        if (accept) {
            naccpt += 1;
            if (DENSE) {
#pragma unroll
                for (int c = 0; c < NL; ++c) {
                    const double ydiff = y1[c] - y[c];
                    const double bspl = h_step * k1[c] - ydiff;
                    cont[c] = y[c];
                    cont[NL + c] = ydiff;
                    cont[2 * NL + c] = bspl;
                    cont[3 * NL + c] = ydiff - h_step * k2[c] - bspl;
                    cont[4 * NL + c] = D[c];
                }
                told = x;
                rh = r_recip(h_step);
            }
#pragma unroll
            for (int c = 0; c < NL; ++c) {
                k1[c] = k2[c];
                y[c] = y1[c];
            }
            lfmin = a.ctl[4];
            if (DENSE) emit(a, idx, xph, false);
            t = xph;
            if (was_last) return AENS_R_COMPLETE;
        } else {
            lfmin = 0.0f;
            if (naccpt >= 1) nrej += 1;
        }
        return AENS_R_CONTINUE;
    }
};

}  // namespace aens

/* The phase machine of aens_run with G-lane groups in the place of lanes: the lanes of a group always hold the same
 * member and take the same control path (every decision derives from group-uniform values). */
template <class S, class M>
__device__ __forceinline__ void aens_run_split(const aens_args_t &a) {
    enum { PH_IDLE = 0, PH_SEGINIT = 1, PH_STEP = 2 };
    constexpr int G = M::G;
    constexpr unsigned LEADERS = (G == 2) ? 0x55555555u : (G == 4) ? 0x11111111u : (G == 8) ? 0x01010101u : 0x00010001u;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double eps = 2.220446049250313e-16 * 100;
    S s;
    s.j = lane & (G - 1);
    long long idx = -1;
    int phase = PH_IDLE;
    bool pending = false, drained = false;
    for (;;) {
        const unsigned idle_mask = __ballot_sync(FULL, phase == PH_IDLE);
        if (idle_mask == FULL && drained) break;
        if (!drained && (idle_mask == FULL || __popc(idle_mask) >= a.refill_threshold)) {
            const unsigned idle_groups = idle_mask & LEADERS;
            const int nidle = __popc(idle_groups);
            const int leader = __ffs(idle_mask) - 1;
            unsigned long long base = 0;
            int late = 0;
            if (lane == leader) {
                base = atomicAdd(a.queue, (unsigned long long)nidle);
                if (a.limit_ns) {
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
                if (pending) {
                    s.store(a, idx);
                    pending = false;
                }
                const long long q = a.q_begin + (long long)base + __popc(idle_groups & ((1u << (lane & ~(G - 1))) - 1u));
                if (q < a.q_end) {
                    if (a.order) idx = (long long)a.order[q];
                    else idx = a.reverse ? a.q_end - 1 - (q - a.q_begin) : q;
                    s.load(a, idx);
                    if (s.t + eps * aens::dmax(aens::dabs(s.t), aens::dabs(a.tf)) < a.tf) {
                        if (late) { s.status = AENS_ST_TIME_LIMIT; pending = true; }
                        else phase = PH_SEGINIT;
                    } else {
                        pending = true;
                    }
                }
            }
            if (a.q_begin + (long long)base + nidle >= a.q_end) drained = true;
        }
        const unsigned initmask = __ballot_sync(FULL, phase == PH_SEGINIT);
        if (phase == PH_SEGINIT) {
            s.seg_init(initmask, a, idx);
            phase = PH_STEP;
        }
        const unsigned stepmask = __ballot_sync(FULL, phase == PH_STEP);
        if (phase == PH_STEP) {
            int r;
            do {
                r = s.attempt(stepmask, a, idx);
            } while (__all_sync(stepmask, r == AENS_R_CONTINUE));
            if (r != AENS_R_CONTINUE) {
                if (r == AENS_R_ERROR) s.late_status(a);
                s.seg_end();
                pending = true;
                phase = PH_IDLE;
            }
        }
    }
    if (pending) s.store(a, idx);
}

#define AENS_DEFINE_SPLIT_KERNEL(NAME, MODEL, DENSE, MINBLOCKS)                                       \
    extern "C" __global__ void __launch_bounds__(AENS_BLOCK, MINBLOCKS) NAME(const aens_args_t a) {   \
        aens_run_split<aens::Dopri5Split<MODEL, DENSE>, MODEL>(a);                                    \
    }

#endif /* !AENS_STRICT */
#endif /* AENS_SPLIT_CUH */
