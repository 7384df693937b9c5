/*
 * oracle/dopri5_c.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
 *
 * C restatement of Hairer's DOPRI5 as vendored by the reference in
 * /root/reference/thirdparty/hairer/dopri5.f.  The reference ships this solver as Fortran 77 built with
 * f2py; this image has no Fortran compiler, so the algorithm is restated here, operation for operation
 * (same association order, no FMA contraction: build with -O2 -ffp-contract=off), and exposed with the
 * argument list of thirdparty/hairer/dopri5.pyf:31-61 so that the reference's own
 * src/solvers/runge_kutta.py::Dopri5 runs unmodified on top of it (see oracle/dopri5_f2py_shim.py).
 *
 *   driver / parameter defaulting   dopri5.f:199-350   -> dopri5_c()
 *   core integrator DOPCOR          dopri5.f:356-556   -> dopcor()
 *   initial step HINIT              dopri5.f:558-619   -> hinit()
 *   dense output CONTD5             dopri5.f:621-644   -> contd5_c()
 *   tableau CDOPRI                  dopri5.f:646-692   -> constants below
 *
 * Parity pin: the reference's own known answers for Dopri5 (tests/solvers/test_rungekutta.py:40-47,
 * examples/dopri5_basic.py:58, examples/dopri5_with_disc.py:157-159, tests/test_solvers.py:52-57) run
 * through this file in tests/test_oracle_reference.py, plus SciPy's independent C translation of the same
 * Hairer code (step counts / f-evals).
 */
#include <math.h>
#include <stdlib.h>

typedef void (*dop_fcn_t)(int n, double x, const double *y, double *f, void *ext);
/* returns the new IRTRN (dopri5.pyf: irtrn is intent(in,out)) */
typedef int (*dop_solout_t)(int nr, double xold, double x, const double *y, int n,
                            const double *cont, const int *icomp, int nrd, int irtrn, void *ext);

/* COMMON /CONDO5/ XOLD,H  (dopri5.f:372,629): global on purpose -- the reference is not re-entrant. */
static double condo5_xold = 0.0, condo5_h = 0.0;

/* CDOPRI, dopri5.f:646-692 */
#define C2 (0.2)
#define C3 (0.3)
#define C4 (0.8)
#define C5 (8.0 / 9.0)
#define A21 (0.2)
#define A31 (3.0 / 40.0)
#define A32 (9.0 / 40.0)
#define A41 (44.0 / 45.0)
#define A42 (-56.0 / 15.0)
#define A43 (32.0 / 9.0)
#define A51 (19372.0 / 6561.0)
#define A52 (-25360.0 / 2187.0)
#define A53 (64448.0 / 6561.0)
#define A54 (-212.0 / 729.0)
#define A61 (9017.0 / 3168.0)
#define A62 (-355.0 / 33.0)
#define A63 (46732.0 / 5247.0)
#define A64 (49.0 / 176.0)
#define A65 (-5103.0 / 18656.0)
#define A71 (35.0 / 384.0)
#define A73 (500.0 / 1113.0)
#define A74 (125.0 / 192.0)
#define A75 (-2187.0 / 6784.0)
#define A76 (11.0 / 84.0)
#define E1 (71.0 / 57600.0)
#define E3 (-71.0 / 16695.0)
#define E4 (71.0 / 1920.0)
#define E5 (-17253.0 / 339200.0)
#define E6 (22.0 / 525.0)
#define E7 (-1.0 / 40.0)
#define D1 (-12715105075.0 / 11282082432.0)
#define D3 (87487479700.0 / 32700410799.0)
#define D4 (-10690763975.0 / 1880347072.0)
#define D5 (701980252875.0 / 199316789632.0)
#define D6 (-1453857185.0 / 822651844.0)
#define D7 (69997945.0 / 29380423.0)

static double dsign(double a, double b) { /* Fortran SIGN(a,b) */
    double m = fabs(a);
    return (b >= 0.0) ? m : -m;
}

/* HINIT, dopri5.f:558-619.  Note the norms are NOT divided by n and DER2 = sqrt(sum)/H. */
static double hinit(int n, dop_fcn_t fcn, void *ext, double x, const double *y, double posneg,
                    const double *f0, double *f1, double *y1, int iord, double hmax,
                    const double *atol, const double *rtol, int itol)
{
    double dnf = 0.0, dny = 0.0, sk, h, der2, der12, h1, q;
    int i;
    for (i = 0; i < n; ++i) {
        sk = (itol == 0) ? atol[0] + rtol[0] * fabs(y[i]) : atol[i] + rtol[i] * fabs(y[i]);
        q = f0[i] / sk;
        dnf = dnf + q * q;
        q = y[i] / sk;
        dny = dny + q * q;
    }
    if (dnf <= 1.0e-10 || dny <= 1.0e-10)
        h = 1.0e-6;
    else
        h = sqrt(dny / dnf) * 0.01;
    h = fmin(h, hmax);
    h = dsign(h, posneg);
    for (i = 0; i < n; ++i)
        y1[i] = y[i] + h * f0[i];
    fcn(n, x + h, y1, f1, ext);
    der2 = 0.0;
    for (i = 0; i < n; ++i) {
        sk = (itol == 0) ? atol[0] + rtol[0] * fabs(y[i]) : atol[i] + rtol[i] * fabs(y[i]);
        q = (f1[i] - f0[i]) / sk;
        der2 = der2 + q * q;
    }
    der2 = sqrt(der2) / h;
    der12 = fmax(fabs(der2), sqrt(dnf));
    if (der12 <= 1.0e-15)
        h1 = fmax(1.0e-6, fabs(h) * 1.0e-3);
    else
        h1 = pow(0.01 / der12, 1.0 / iord);
    h = fmin(fmin(100 * fabs(h), h1), hmax);
    return dsign(h, posneg);
}

/* DOPCOR, dopri5.f:356-556 */
static int dopcor(int n, dop_fcn_t fcn, void *fext, double *x_io, double *y, double xend, double hmax,
                  double *h_io, const double *rtol, const double *atol, int itol,
                  dop_solout_t solout, void *sext, int iout, long nmax, double uround, int nstiff,
                  double safe, double beta, double fac1, double fac2,
                  double *y1, double *k1, double *k2, double *k3, double *k4, double *k5, double *k6,
                  double *ysti, double *cont, const int *icomp, int nrd,
                  int *nfcn, int *nstep, int *naccpt, int *nrejct)
{
    double x = *x_io, h = *h_io;
    double facold = 1.0e-4;
    double expo1 = 0.2 - beta * 0.75;
    double facc1 = 1.0 / fac1;
    double facc2 = 1.0 / fac2;
    double posneg = dsign(1.0, xend - x);
    double hlamb = 0.0, xph, err, sk, fac11, fac, hnew, xold, stnum, stden, q;
    int last = 0, reject = 0, iasti = 0, nonsti = 0, irtrn, i, j;

    fcn(n, x, y, k1, fext);
    hmax = fabs(hmax);
    if (h == 0.0)
        h = hinit(n, fcn, fext, x, y, posneg, k1, k2, k3, 5, hmax, atol, rtol, itol);
    *nfcn += 2;
    xold = x;
    if (iout != 0) {
        irtrn = 1;
        condo5_xold = xold;
        condo5_h = h;
        irtrn = solout(*naccpt + 1, xold, x, y, n, cont, icomp, nrd, irtrn, sext);
        if (irtrn < 0) goto interrupted;
    } else {
        irtrn = 0;
    }
    for (;;) {
        if (*nstep > nmax) { *x_io = x; *h_io = h; return -2; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { *x_io = x; *h_io = h; return -3; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) {
            h = xend - x;
            last = 1;
        }
        *nstep += 1;
        if (irtrn >= 2) fcn(n, x, y, k1, fext);
        for (i = 0; i < n; ++i) y1[i] = y[i] + h * A21 * k1[i];
        fcn(n, x + C2 * h, y1, k2, fext);
        for (i = 0; i < n; ++i) y1[i] = y[i] + h * (A31 * k1[i] + A32 * k2[i]);
        fcn(n, x + C3 * h, y1, k3, fext);
        for (i = 0; i < n; ++i) y1[i] = y[i] + h * (A41 * k1[i] + A42 * k2[i] + A43 * k3[i]);
        fcn(n, x + C4 * h, y1, k4, fext);
        for (i = 0; i < n; ++i) y1[i] = y[i] + h * (A51 * k1[i] + A52 * k2[i] + A53 * k3[i] + A54 * k4[i]);
        fcn(n, x + C5 * h, y1, k5, fext);
        for (i = 0; i < n; ++i)
            ysti[i] = y[i] + h * (A61 * k1[i] + A62 * k2[i] + A63 * k3[i] + A64 * k4[i] + A65 * k5[i]);
        xph = x + h;
        fcn(n, xph, ysti, k6, fext);
        for (i = 0; i < n; ++i)
            y1[i] = y[i] + h * (A71 * k1[i] + A73 * k3[i] + A74 * k4[i] + A75 * k5[i] + A76 * k6[i]);
        fcn(n, xph, y1, k2, fext);
        if (iout >= 2) {
            for (j = 0; j < nrd; ++j) {
                i = icomp[j] - 1;
                cont[4 * nrd + j] = h * (D1 * k1[i] + D3 * k3[i] + D4 * k4[i] + D5 * k5[i] + D6 * k6[i] + D7 * k2[i]);
            }
        }
        for (i = 0; i < n; ++i)
            k4[i] = (E1 * k1[i] + E3 * k3[i] + E4 * k4[i] + E5 * k5[i] + E6 * k6[i] + E7 * k2[i]) * h;
        *nfcn += 6;
        /* error estimation, dopri5.f:451-461 */
        err = 0.0;
        for (i = 0; i < n; ++i) {
            if (itol == 0) sk = atol[0] + rtol[0] * fmax(fabs(y[i]), fabs(y1[i]));
            else           sk = atol[i] + rtol[i] * fmax(fabs(y[i]), fabs(y1[i]));
            q = k4[i] / sk;
            err = err + q * q;
        }
        err = sqrt(err / n);
        /* step size controller with Lund stabilisation, dopri5.f:463-468 */
        fac11 = pow(err, expo1);
        fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1.0e-4);
            *naccpt += 1;
            /* stiffness detection, dopri5.f:474-494 (IPRINT>0 in Assimulo: message only, never aborts) */
            if ((*naccpt % nstiff) == 0 || iasti > 0) {
                stnum = 0.0; stden = 0.0;
                for (i = 0; i < n; ++i) {
                    q = k2[i] - k6[i]; stnum = stnum + q * q;
                    q = y1[i] - ysti[i]; stden = stden + q * q;
                }
                if (stden > 0.0) hlamb = h * sqrt(stnum / stden);
                if (hlamb > 3.25) {
                    nonsti = 0;
                    iasti += 1;
                } else {
                    nonsti += 1;
                    if (nonsti == 6) iasti = 0;
                }
            }
            if (iout >= 2) {
                for (j = 0; j < nrd; ++j) {
                    double yd0, ydiff, bspl;
                    i = icomp[j] - 1;
                    yd0 = y[i];
                    ydiff = y1[i] - yd0;
                    bspl = h * k1[i] - ydiff;
                    cont[j] = y[i];
                    cont[nrd + j] = ydiff;
                    cont[2 * nrd + j] = bspl;
                    cont[3 * nrd + j] = -h * k2[i] + ydiff - bspl;
                }
            }
            for (i = 0; i < n; ++i) { k1[i] = k2[i]; y[i] = y1[i]; }
            xold = x;
            x = xph;
            if (iout != 0) {
                condo5_xold = xold;
                condo5_h = h;
                irtrn = solout(*naccpt + 1, xold, x, y, n, cont, icomp, nrd, irtrn, sext);
                if (irtrn < 0) goto interrupted;
            }
            if (last) {
                h = hnew;
                *x_io = x; *h_io = h;
                return 1;
            }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            hnew = h / fmin(facc1, fac11 / safe);
            reject = 1;
            if (*naccpt >= 1) *nrejct += 1;
            last = 0;
        }
        h = hnew;
    }
interrupted:
    *x_io = x; *h_io = h;
    return 2;
}

/* DOPRI5 driver, dopri5.f:199-350.  work/iwork are 0-based views of the Fortran WORK(1..)/IWORK(1..). */
int dopri5_c(int n, dop_fcn_t fcn, void *fext, double *x, double *y, double xend,
             const double *rtol, const double *atol, int itol,
             dop_solout_t solout, void *sext, int iout,
             double *work, int lwork, int *iwork, int liwork)
{
    int nfcn = 0, nstep = 0, naccpt = 0, nrejct = 0, arret = 0, idid, i;
    long nmax;
    int meth, nstiff, nrdens;
    double uround, safe, fac1, fac2, beta, hmax, h;

    if (iwork[0] == 0) nmax = 100000; else { nmax = iwork[0]; if (nmax <= 0) arret = 1; }
    if (iwork[1] == 0) meth = 1; else { meth = iwork[1]; if (meth <= 0 || meth >= 4) arret = 1; }
    nstiff = iwork[3];
    if (nstiff == 0) nstiff = 1000;
    if (nstiff < 0) nstiff = (int)nmax + 10;
    nrdens = iwork[4];
    if (nrdens < 0 || nrdens > n) arret = 1;
    else if (nrdens == n) for (i = 0; i < nrdens; ++i) iwork[20 + i] = i + 1;
    if (work[0] == 0.0) uround = 2.3e-16;
    else { uround = work[0]; if (uround <= 1.0e-35 || uround >= 1.0) arret = 1; }
    if (work[1] == 0.0) safe = 0.9;
    else { safe = work[1]; if (safe >= 1.0 || safe <= 1.0e-4) arret = 1; }
    fac1 = (work[2] == 0.0) ? 0.2 : work[2];
    fac2 = (work[3] == 0.0) ? 10.0 : work[3];
    if (work[4] == 0.0) beta = 0.04;
    else if (work[4] < 0.0) beta = 0.0;
    else { beta = work[4]; if (beta > 0.2) arret = 1; }
    hmax = (work[5] == 0.0) ? xend - *x : work[5];
    h = work[6];
    if (20 + 8 * n + 5 * nrdens > lwork) arret = 1;
    if (20 + nrdens > liwork) arret = 1;
    if (arret) return -1;
    {
        double *y1 = work + 20, *k1 = y1 + n, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n, *k5 = k4 + n,
               *k6 = k5 + n, *ys = k6 + n, *co = ys + n;
        idid = dopcor(n, fcn, fext, x, y, xend, hmax, &h, rtol, atol, itol, solout, sext, iout, nmax,
                      uround, nstiff, safe, beta, fac1, fac2, y1, k1, k2, k3, k4, k5, k6, ys, co,
                      iwork + 20, nrdens, &nfcn, &nstep, &naccpt, &nrejct);
    }
    work[6] = h;
    iwork[16] = nfcn;
    iwork[17] = nstep;
    iwork[18] = naccpt;
    iwork[19] = nrejct;
    return idid;
}

/* CONTD5, dopri5.f:621-644 (ii is 1-based as in Fortran) */
double contd5_c(int ii, double x, const double *con, const int *icomp, int nd)
{
    int i = 0, j;
    double theta, theta1;
    for (j = 1; j <= nd; ++j)
        if (icomp[j - 1] == ii) i = j;
    if (i == 0) return 0.0;
    theta = (x - condo5_xold) / condo5_h;
    theta1 = 1.0 - theta;
    return con[i - 1] + theta * (con[nd + i - 1] + theta1 * (con[2 * nd + i - 1] + theta *
           (con[3 * nd + i - 1] + theta1 * con[4 * nd + i - 1])));
}
