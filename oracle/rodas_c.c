/*
 * oracle/rodas_c.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
 *
 * C restatement of Hairer & Wanner's RODAS as vendored by the reference in
 * /root/reference/thirdparty/hairer/rodas_decsol.f, for the case Assimulo's RodasODE drives
 * (src/solvers/rosenbrock.py:376-421): explicit problem (IMAS = 0), full Jacobian (IJOB = 1), METH = 1,
 * non-autonomous (IFCN = 1) with df/dx by finite differences (IDFX = 0), Jacobian analytic (IJAC = 1) or by
 * finite differences (IJAC = 0), vector tolerances (ITOL = 1), Gustafsson predictive controller (PRED).
 * The reference ships it as Fortran 77 built with f2py; this image has no Fortran compiler, so the algorithm
 * is restated operation for operation (same association order; build with -O2 -ffp-contract=off) and exposed
 * with the argument list of thirdparty/hairer/rodas_decsol.pyf so that the reference's own
 * src/solvers/rosenbrock.py::RodasODE runs unmodified on top of it (oracle/rodas_f2py_shim.py).
 *
 *   driver / parameter defaulting   rodas_decsol.f:1-530     -> rodas_c()
 *   core integrator ROSCOR          rodas_decsol.f:532-872   -> roscor()
 *   dense output CONTRO             rodas_decsol.f:874-882   -> contro_c()
 *   coefficients ROCOE (METH = 1)   rodas_decsol.f:889-940   -> constants below
 *   DECOMR / SLVROD (IJOB = 1)      rodas_decsol.f:1051-1068, 3036-3060
 *   DEC / SOL                       rodas_decsol.f (Moler, CACM alg. 423)
 *
 * Parity pin: "parity unpinned" at the level of the Fortran itself -- it never ran here.  The restatement is
 * pinned by the reference's single known answer for RodasODE (tests/solvers/test_rosenbrock.py:91 and
 * examples/rodasode_vanderpol.py:88: Van der Pol mu = 1e6, y1(2) = 1.7061680350 +- 1e-4) and by agreement with
 * Radau5ODE (100 % reference code) to the requested tolerance on every golden case (tests/test_oracle_reference.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef void (*rod_fcn_t)(int n, double x, const double *y, double *f, void *ext);
typedef void (*rod_jac_t)(int n, double x, const double *y, double *fjac /* column-major [ld][n] */, int ld, void *ext);
/* returns the new IRTRN */
typedef int (*rod_solout_t)(int nr, double xold, double x, const double *y, const double *cont, int lrc, int n,
                            int irtrn, void *ext);

/* COMMON /CONROS/ XOLD,HOUT,NN (rodas_decsol.f:546): global on purpose -- the reference is not re-entrant */
static double conros_xold = 0.0, conros_h = 0.0;
static int conros_n = 0;

/* ROCOE, METH = 1 (rodas_decsol.f:896-940) */
static const double C2 = 0.386, C3 = 0.21, C4 = 0.63;
static const double D1 = 0.2500000000000000e+00, D2 = -0.1043000000000000e+00, D3 = 0.1035000000000000e+00,
                    D4 = -0.3620000000000023e-01;
static const double A21 = 0.1544000000000000e+01, A31 = 0.9466785280815826e+00, A32 = 0.2557011698983284e+00,
                    A41 = 0.3314825187068521e+01, A42 = 0.2896124015972201e+01, A43 = 0.9986419139977817e+00,
                    A51 = 0.1221224509226641e+01, A52 = 0.6019134481288629e+01, A53 = 0.1253708332932087e+02,
                    A54 = -0.6878860361058950e+00;
static const double C21 = -0.5668800000000000e+01, C31 = -0.2430093356833875e+01, C32 = -0.2063599157091915e+00,
                    C41 = -0.1073529058151375e+00, C42 = -0.9594562251023355e+01, C43 = -0.2047028614809616e+02,
                    C51 = 0.7496443313967647e+01, C52 = -0.1024680431464352e+02, C53 = -0.3399990352819905e+02,
                    C54 = 0.1170890893206160e+02, C61 = 0.8083246795921522e+01, C62 = -0.7981132988064893e+01,
                    C63 = -0.3152159432874371e+02, C64 = 0.1631930543123136e+02, C65 = -0.6058818238834054e+01;
static const double GAMMA = 0.2500000000000000e+00;
static const double D21 = 0.1012623508344586e+02, D22 = -0.7487995877610167e+01, D23 = -0.3480091861555747e+02,
                    D24 = -0.7992771707568823e+01, D25 = 0.1025137723295662e+01, D31 = -0.6762803392801253e+00,
                    D32 = 0.6087714651680015e+01, D33 = 0.1643084320892478e+02, D34 = 0.2476722511418386e+02,
                    D35 = -0.6594389125716872e+01;

#define A_(a, i, j) (a)[(i) + (size_t)(j) * n]

/* DEC: Gaussian elimination with partial pivoting; multipliers stored negated */
static int dec(int n, double *a, int *ip)
{
    int k, i, j, m;
    double t;
    ip[n - 1] = 1;
    for (k = 0; k < n - 1; ++k) {
        m = k;
        for (i = k + 1; i < n; ++i)
            if (fabs(A_(a, i, k)) > fabs(A_(a, m, k))) m = i;
        ip[k] = m;
        t = A_(a, m, k);
        if (m != k) {
            ip[n - 1] = -ip[n - 1];
            A_(a, m, k) = A_(a, k, k);
            A_(a, k, k) = t;
        }
        if (t == 0.0) { ip[n - 1] = 0; return k + 1; }
        t = 1.0 / t;
        for (i = k + 1; i < n; ++i) A_(a, i, k) = -A_(a, i, k) * t;
        for (j = k + 1; j < n; ++j) {
            t = A_(a, m, j);
            A_(a, m, j) = A_(a, k, j);
            A_(a, k, j) = t;
            if (t != 0.0)
                for (i = k + 1; i < n; ++i) A_(a, i, j) = A_(a, i, j) + A_(a, i, k) * t;
        }
    }
    if (A_(a, n - 1, n - 1) == 0.0) { ip[n - 1] = 0; return n; }
    return 0;
}

static void sol(int n, const double *a, double *b, const int *ip)
{
    int k, i, m, kb;
    double t;
    if (n > 1) {
        for (k = 0; k < n - 1; ++k) {
            m = ip[k];
            t = b[m];
            b[m] = b[k];
            b[k] = t;
            for (i = k + 1; i < n; ++i) b[i] = b[i] + A_(a, i, k) * t;
        }
        for (kb = 0; kb < n - 1; ++kb) {
            k = n - 1 - kb;
            b[k] = b[k] / A_(a, k, k);
            t = -b[k];
            for (i = 0; i < k; ++i) b[i] = b[i] + A_(a, i, k) * t;
        }
    }
    b[0] = b[0] / A_(a, 0, 0);
}

/* SLVROD, IJOB = 1 (rodas_decsol.f:3036-3060) */
static void slvrod(int n, const double *e, const int *ip, const double *dy, double *ak, const double *fx,
                   const double *ynew, double hd, int stage1)
{
    int i;
    if (hd == 0.0) for (i = 0; i < n; ++i) ak[i] = dy[i];
    else for (i = 0; i < n; ++i) ak[i] = dy[i] + hd * fx[i];
    if (stage1) for (i = 0; i < n; ++i) ak[i] = ak[i] + ynew[i];
    sol(n, e, ak, ip);
}

typedef struct {
    int nfcn, njac, nstep, naccpt, nrejct, ndec, nsol;
} rod_stats_t;

/* ROSCOR (rodas_decsol.f:532-872); returns IDID */
static int roscor(int n, rod_fcn_t fcn, void *fext, double *x, double *y, double xend, double hmax, double *h,
                  const double *rtol, const double *atol, int itol, rod_jac_t jac, void *jext, int ijac, int autnms,
                  rod_solout_t solout, void *sext, int iout, int nmax, double uround, double fac1, double fac2,
                  double safe, int pred, double *w, int *ip, double *cont, rod_stats_t *st)
{
    double *ynew = w, *dy1 = w + n, *dy = w + 2 * n, *ak1 = w + 3 * n, *ak2 = w + 4 * n, *ak3 = w + 5 * n,
           *ak4 = w + 6 * n, *ak5 = w + 7 * n, *ak6 = w + 8 * n, *fx = w + 9 * n, *fjac = w + 10 * n,
           *e = w + 10 * n + (size_t)n * n;
    const int lrc = 4 * n;
    double posneg, hmaxn, hopt = 0.0, fac, err, hnew, sk, facgus, hacc = 0.0, erracc = 0.0, delt, xdelt, ysafe;
    double hd1 = 0.0, hd2 = 0.0, hd3 = 0.0, hd4 = 0.0;
    double hc21, hc31, hc32, hc41, hc42, hc43, hc51, hc52, hc53, hc54, hc61, hc62, hc63, hc64, hc65;
    int reject = 0, last = 0, nsing = 0, irtrn, i, j, ier;
    conros_n = n;
    posneg = (xend - *x) >= 0.0 ? 1.0 : -1.0;
    hmaxn = fmin(fabs(hmax), fabs(xend - *x));
    if (fabs(*h) <= 10.0 * uround) *h = 1.0e-6;
    *h = fmin(fabs(*h), hmaxn);
    *h = copysign(*h, posneg);
    if (iout != 0) {
        conros_xold = *x;
        irtrn = 1;
        conros_h = *h;
        irtrn = solout(st->naccpt + 1, conros_xold, *x, y, cont, lrc, n, irtrn, sext);
        if (irtrn < 0) return 2;
    }
L1:
    if (st->nstep > nmax) return -2;
    if (0.1 * fabs(*h) <= fabs(*x) * uround) return -3;
    if (last) { *h = hopt; return 1; }
    hopt = *h;
    if ((*x + *h * 1.0001 - xend) * posneg >= 0.0) { *h = xend - *x; last = 1; }
    fcn(n, *x, y, dy1, fext);
    st->nfcn += 1;
    st->njac += 1;
    if (ijac == 0) {
        for (i = 0; i < n; ++i) {
            ysafe = y[i];
            delt = sqrt(uround * fmax(1.0e-5, fabs(ysafe)));
            y[i] = ysafe + delt;
            fcn(n, *x, y, ak1, fext);
            for (j = 0; j < n; ++j) A_(fjac, j, i) = (ak1[j] - dy1[j]) / delt;
            y[i] = ysafe;
        }
    } else {
        jac(n, *x, y, fjac, n, jext);
    }
    if (!autnms) { /* IDFX = 0 */
        delt = sqrt(uround * fmax(1.0e-5, fabs(*x)));
        xdelt = *x + delt;
        fcn(n, xdelt, y, ak1, fext);
        for (j = 0; j < n; ++j) fx[j] = (ak1[j] - dy1[j]) / delt;
    }
L2:
    fac = 1.0 / (*h * GAMMA);
    for (j = 0; j < n; ++j) { /* DECOMR, IJOB = 1 */
        for (i = 0; i < n; ++i) A_(e, i, j) = -A_(fjac, i, j);
        A_(e, j, j) = A_(e, j, j) + fac;
    }
    ier = dec(n, e, ip);
    if (ier != 0) goto L80;
    st->ndec += 1;
    hc21 = C21 / *h; hc31 = C31 / *h; hc32 = C32 / *h; hc41 = C41 / *h; hc42 = C42 / *h; hc43 = C43 / *h;
    hc51 = C51 / *h; hc52 = C52 / *h; hc53 = C53 / *h; hc54 = C54 / *h;
    hc61 = C61 / *h; hc62 = C62 / *h; hc63 = C63 / *h; hc64 = C64 / *h; hc65 = C65 / *h;
    if (!autnms) { hd1 = *h * D1; hd2 = *h * D2; hd3 = *h * D3; hd4 = *h * D4; }
    slvrod(n, e, ip, dy1, ak1, fx, ynew, hd1, 0);
    for (i = 0; i < n; ++i) ynew[i] = y[i] + A21 * ak1[i];
    fcn(n, *x + C2 * *h, ynew, dy, fext);
    for (i = 0; i < n; ++i) ynew[i] = hc21 * ak1[i];
    slvrod(n, e, ip, dy, ak2, fx, ynew, hd2, 1);
    for (i = 0; i < n; ++i) ynew[i] = y[i] + A31 * ak1[i] + A32 * ak2[i];
    fcn(n, *x + C3 * *h, ynew, dy, fext);
    for (i = 0; i < n; ++i) ynew[i] = hc31 * ak1[i] + hc32 * ak2[i];
    slvrod(n, e, ip, dy, ak3, fx, ynew, hd3, 1);
    for (i = 0; i < n; ++i) ynew[i] = y[i] + A41 * ak1[i] + A42 * ak2[i] + A43 * ak3[i];
    fcn(n, *x + C4 * *h, ynew, dy, fext);
    for (i = 0; i < n; ++i) ynew[i] = hc41 * ak1[i] + hc42 * ak2[i] + hc43 * ak3[i];
    slvrod(n, e, ip, dy, ak4, fx, ynew, hd4, 1);
    for (i = 0; i < n; ++i) ynew[i] = y[i] + A51 * ak1[i] + A52 * ak2[i] + A53 * ak3[i] + A54 * ak4[i];
    fcn(n, *x + *h, ynew, dy, fext);
    for (i = 0; i < n; ++i) ak6[i] = hc52 * ak2[i] + hc54 * ak4[i] + hc51 * ak1[i] + hc53 * ak3[i];
    slvrod(n, e, ip, dy, ak5, fx, ak6, 0.0, 1);
    for (i = 0; i < n; ++i) ynew[i] = ynew[i] + ak5[i];
    fcn(n, *x + *h, ynew, dy, fext);
    for (i = 0; i < n; ++i)
        cont[i] = hc61 * ak1[i] + hc62 * ak2[i] + hc65 * ak5[i] + hc64 * ak4[i] + hc63 * ak3[i];
    slvrod(n, e, ip, dy, ak6, fx, cont, 0.0, 1);
    for (i = 0; i < n; ++i) ynew[i] = ynew[i] + ak6[i];
    st->nsol += 6;
    st->nfcn += 5;
    if (iout != 0) {
        for (i = 0; i < n; ++i) {
            cont[i] = y[i];
            cont[i + 2 * n] = D21 * ak1[i] + D22 * ak2[i] + D23 * ak3[i] + D24 * ak4[i] + D25 * ak5[i];
            cont[i + 3 * n] = D31 * ak1[i] + D32 * ak2[i] + D33 * ak3[i] + D34 * ak4[i] + D35 * ak5[i];
        }
    }
    st->nstep += 1;
    err = 0.0;
    for (i = 0; i < n; ++i) {
        if (itol == 0) sk = atol[0] + rtol[0] * fmax(fabs(y[i]), fabs(ynew[i]));
        else sk = atol[i] + rtol[i] * fmax(fabs(y[i]), fabs(ynew[i]));
        err = err + (ak6[i] / sk) * (ak6[i] / sk);
    }
    err = sqrt(err / n);
    fac = fmax(fac2, fmin(fac1, pow(err, 0.25) / safe));
    hnew = *h / fac;
    if (err <= 1.0) {
        st->naccpt += 1;
        if (pred) {
            if (st->naccpt > 1) {
                facgus = (hacc / *h) * pow(err * err / erracc, 0.25) / safe;
                facgus = fmax(fac2, fmin(fac1, facgus));
                fac = fmax(fac, facgus);
                hnew = *h / fac;
            }
            hacc = *h;
            erracc = fmax(1.0e-2, err);
        }
        for (i = 0; i < n; ++i) y[i] = ynew[i];
        conros_xold = *x;
        *x = *x + *h;
        if (iout != 0) {
            for (i = 0; i < n; ++i) cont[n + i] = y[i];
            irtrn = 1;
            conros_h = *h;
            irtrn = solout(st->naccpt + 1, conros_xold, *x, y, cont, lrc, n, irtrn, sext);
            if (irtrn < 0) return 2;
        }
        if (fabs(hnew) > hmaxn) hnew = posneg * hmaxn;
        if (reject) hnew = posneg * fmin(fabs(hnew), fabs(*h));
        reject = 0;
        *h = hnew;
        goto L1;
    } else {
        reject = 1;
        last = 0;
        *h = hnew;
        if (st->naccpt >= 1) st->nrejct += 1;
        goto L2;
    }
L80:
    nsing += 1;
    if (nsing >= 5) return -4;
    *h = *h * 0.5;
    reject = 1;
    last = 0;
    goto L2;
}

/* RODAS driver (rodas_decsol.f:1-530) for IMAS = 0, full Jacobian.  work / iwork follow the Fortran layout
 * (1-based WORK(1..5), IWORK(1..3) read; IWORK(14..20) written); *h is in/out. */
int rodas_c(int n, rod_fcn_t fcn, void *fext, int ifcn, double *x, double *y, double xend, double *h,
            const double *rtol, const double *atol, int itol, rod_jac_t jac, void *jext, int ijac,
            rod_solout_t solout, void *sext, int iout, const double *work, int lwork, int *iwork, int liwork)
{
    int nmax, meth, pred, i, idid, *ip;
    double uround, hmax, fac1, fac2, safe, *w, *cont;
    rod_stats_t st;
    memset(&st, 0, sizeof(st));
    (void)lwork; (void)liwork;
    nmax = iwork[0] == 0 ? 100000 : iwork[0];
    if (nmax <= 0) return -1;
    meth = iwork[1] == 0 ? 1 : iwork[1];
    if (meth != 1) return -1; /* only METH = 1 is restated (the one Assimulo uses) */
    pred = iwork[2] <= 1;
    uround = work[0] == 0.0 ? 1.0e-16 : work[0];
    if (uround < 1.0e-16 || uround >= 1.0) return -1;
    hmax = work[1] == 0.0 ? xend - *x : work[1];
    fac1 = work[2] == 0.0 ? 5.0 : 1.0 / work[2];
    fac2 = work[3] == 0.0 ? 1.0 / 6.0 : 1.0 / work[3];
    if (fac1 < 1.0 || fac2 > 1.0) return -1;
    safe = work[4] == 0.0 ? 0.9 : work[4];
    if (safe <= 0.001 || safe >= 1.0) return -1;
    for (i = 0; i < (itol == 0 ? 1 : n); ++i)
        if (atol[i] <= 0.0 || rtol[i] <= 10.0 * uround) return -1;
    w = (double *)malloc(sizeof(double) * ((size_t)10 * n + 2 * (size_t)n * n + 4 * n));
    ip = (int *)malloc(sizeof(int) * n);
    if (!w || !ip) { free(w); free(ip); return -1; }
    cont = w + (size_t)10 * n + 2 * (size_t)n * n;
    memset(cont, 0, sizeof(double) * 4 * n);
    idid = roscor(n, fcn, fext, x, y, xend, hmax, h, rtol, atol, itol, jac, jext, ijac, ifcn == 0, solout, sext,
                  iout, nmax, uround, fac1, fac2, safe, pred, w, ip, cont, &st);
    iwork[13] = st.nfcn; iwork[14] = st.njac; iwork[15] = st.nstep; iwork[16] = st.naccpt;
    iwork[17] = st.nrejct; iwork[18] = st.ndec; iwork[19] = st.nsol;
    free(w);
    free(ip);
    return idid;
}

/* CONTRO (rodas_decsol.f:874-882), i is 1-based like the Fortran */
double contro_c(int i, double x, const double *cont, int lrc)
{
    const int n = conros_n;
    const double s = (x - conros_xold) / conros_h;
    (void)lrc;
    return cont[i - 1] * (1 - s) + s * (cont[i - 1 + n] + (1 - s) * (cont[i - 1 + n * 2] + s * cont[i - 1 + n * 3]));
}
