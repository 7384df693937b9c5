/*
 * One translation unit per (built-in model, arithmetic mode): compiled as
 *   nvcc -DAENS_KMODEL=AensVdp -DAENS_KNAME=vdp -DAENS_STRICT=1 --fmad=false ...   -> strict kernels
 *   nvcc -DAENS_KMODEL=AensVdp -DAENS_KNAME=vdp -DAENS_STRICT=0 ...                -> fast kernels
 * and exporting one lookup function  aens_kernels_<name>_<s|f>(solver, dense) -> kernel entry point.
 */
#include "aens_device.cuh"
#include "aens_radau5.cuh"
#include "aens_split.cuh"

#define AENS_CAT_(a, b) a##b
#define AENS_CAT(a, b) AENS_CAT_(a, b)
#if AENS_STRICT
#define AENS_SUF _s
#else
#define AENS_SUF _f
#endif
#define KN(base) AENS_CAT(AENS_CAT(AENS_CAT(base, AENS_KNAME), _), AENS_CAT(x, AENS_SUF))

/* minimum resident CTAs per SM requested from ptxas (__launch_bounds__), per solver; tuned in profiles/ */
#ifndef AENS_MB_EULER
#define AENS_MB_EULER 1
#endif
#ifndef AENS_MB_RK4
#define AENS_MB_RK4 1
#endif
#ifndef AENS_MB_RK34
#define AENS_MB_RK34 1
#endif
#ifndef AENS_MB_DOPRI5
#define AENS_MB_DOPRI5 1
#endif
/* Radau5ODE / RodasODE on small systems: three CTAs per SM (<= 168 registers; the Jacobian, the collocation
 * polynomial and f(x, y) live in shared memory, aens_radau5.cuh) -- 1.41 -> 1.18 ms on config 4, profiles/r2_radau.md */
#ifndef AENS_MB_RADAU5
#define AENS_MB_RADAU5 (AENS_KMODEL::N <= 4 ? 3 : 1)
#endif

typedef AENS_KMODEL KM;
constexpr bool kEvents = (KM::NG > 0);

AENS_DEFINE_KERNEL(KN(aens_euler_d1_), Euler, KM, true, AENS_MB_EULER)
AENS_DEFINE_KERNEL(KN(aens_rk4_d0_), RK4, KM, false, AENS_MB_RK4)
AENS_DEFINE_KERNEL(KN(aens_rk34_d1_), RK34, KM, true, AENS_MB_RK34)
AENS_DEFINE_KERNEL(KN(aens_dopri5_d1_), Dopri5, KM, true, AENS_MB_DOPRI5)
AENS_DEFINE_KERNEL(KN(aens_radau5_d1_), Radau5, KM, true, AENS_MB_RADAU5)
AENS_DEFINE_KERNEL(KN(aens_impleuler_d1_), ImplEuler, KM, true, AENS_MB_EULER)
AENS_DEFINE_KERNEL(KN(aens_rodas_d1_), Rodas, KM, true, AENS_MB_RADAU5)
/* row reduction mode (aens_set_row_reduction): the grid rows are reduced over the ensemble inside the kernel */
AENS_DEFINE_KERNEL(KN(aens_euler_d2_), Euler, KM, 2, AENS_MB_EULER)
AENS_DEFINE_KERNEL(KN(aens_rk34_d2_), RK34, KM, 2, AENS_MB_RK34)
AENS_DEFINE_KERNEL(KN(aens_dopri5_d2_), Dopri5, KM, 2, AENS_MB_DOPRI5)
AENS_DEFINE_KERNEL(KN(aens_radau5_d2_), Radau5, KM, 2, AENS_MB_RADAU5)
AENS_DEFINE_KERNEL(KN(aens_impleuler_d2_), ImplEuler, KM, 2, AENS_MB_EULER)
AENS_DEFINE_KERNEL(KN(aens_rodas_d2_), Rodas, KM, 2, AENS_MB_RADAU5)
/* models without events additionally get dense-free variants (no interpolation data kept in registers) */
AENS_DEFINE_KERNEL(KN(aens_euler_d0_), Euler, KM, kEvents, AENS_MB_EULER)
AENS_DEFINE_KERNEL(KN(aens_rk34_d0_), RK34, KM, kEvents, AENS_MB_RK34)
AENS_DEFINE_KERNEL(KN(aens_dopri5_d0_), Dopri5, KM, kEvents, AENS_MB_DOPRI5)
AENS_DEFINE_KERNEL(KN(aens_radau5_d0_), Radau5, KM, kEvents, AENS_MB_RADAU5)
AENS_DEFINE_KERNEL(KN(aens_impleuler_d0_), ImplEuler, KM, kEvents, AENS_MB_EULER)

/* backward integration: the three solvers whose reference cores carry POSNEG, on the time-reversed model */
typedef AensRev<KM> KMR;
AENS_DEFINE_KERNEL(KN(aens_dopri5_b1_), Dopri5, KMR, true, AENS_MB_DOPRI5)
AENS_DEFINE_KERNEL(KN(aens_radau5_b1_), Radau5, KMR, true, AENS_MB_RADAU5)
AENS_DEFINE_KERNEL(KN(aens_rodas_b1_), Rodas, KMR, true, AENS_MB_RADAU5)
extern "C" const void *AENS_CAT(AENS_CAT(aens_kernels_bwd_, AENS_KNAME), AENS_SUF)(int solver) {
    switch (solver) {
    case AENS_SOLVER_DOPRI5: return (const void *)KN(aens_dopri5_b1_);
    case AENS_SOLVER_RADAU5: return (const void *)KN(aens_radau5_b1_);
    case AENS_SOLVER_RODAS: return (const void *)KN(aens_rodas_b1_);
    }
    return nullptr;
}

/* sub-warp-per-member Dopri5 (aens_split.cuh) for the models that define rhs_g; FAST build only */
#ifndef AENS_MB_SPLIT
#define AENS_MB_SPLIT 1
#endif
#if defined(AENS_KSPLIT) && !AENS_STRICT
AENS_DEFINE_SPLIT_KERNEL(KN(aens_dopri5g_d0_), KM, 0, AENS_MB_SPLIT)
AENS_DEFINE_SPLIT_KERNEL(KN(aens_dopri5g_d1_), KM, 1, AENS_MB_SPLIT)
extern "C" const void *AENS_CAT(AENS_CAT(aens_kernels_split_, AENS_KNAME), AENS_SUF)(int solver, int dense, int *lanes) {
    if (lanes) *lanes = KM::G;
    if (solver != AENS_SOLVER_DOPRI5 || dense == 2) return nullptr;
    return dense ? (const void *)KN(aens_dopri5g_d1_) : (const void *)KN(aens_dopri5g_d0_);
}
#endif

extern "C" const void *AENS_CAT(AENS_CAT(aens_kernels_, AENS_KNAME), AENS_SUF)(int solver, int dense) {
#define AENS_PICK(base) (dense == 2 ? (const void *)KN(base##d2_) : dense ? (const void *)KN(base##d1_) : (const void *)KN(base##d0_))
    switch (solver) {
    case AENS_SOLVER_EULER: return AENS_PICK(aens_euler_);
    case AENS_SOLVER_RK4: return (const void *)KN(aens_rk4_d0_);
    case AENS_SOLVER_RK34: return AENS_PICK(aens_rk34_);
    case AENS_SOLVER_DOPRI5: return AENS_PICK(aens_dopri5_);
    case AENS_SOLVER_RADAU5: return AENS_PICK(aens_radau5_);
    case AENS_SOLVER_IMPLEULER: return AENS_PICK(aens_impleuler_);
    case AENS_SOLVER_RODAS: /* cont[] is always kept */
        return dense == 2 ? (const void *)KN(aens_rodas_d2_) : (const void *)KN(aens_rodas_d1_);
    }
    return nullptr;
}
