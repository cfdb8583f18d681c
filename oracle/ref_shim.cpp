// ref_shim.cpp -- extern "C" doorway into the compiled, unmodified reference
// (test infrastructure; see oracle/Makefile).  The reference exports C++-mangled
// free functions (gs.h:3-17, multigrid.cpp:17-21,97-99,124-126); this file only
// forwards to them so that Python (ctypes) and bench.py can call them.
#include <cstring>
#include "gs.h"   // from /root/reference (-I), unmodified

void mg_inner(double** u, double** rhs, double** v1, double** v2, double* tmp, double dx, int n,
              int lvl, int maxlvl, int shape, double dt, double nu);
void mg_outer(double** utow, double** v1tow, double** v2tow, double** rhstow, double* tmp,
              double nu, int maxlvl, int n, double dt, double dx, double tol, int shape);
void timestepper(double* uT, double* u0, double* v1, double* v2, double nu, int maxlvl, int n,
                 double dt, double T, double dx, double tol, int shape);

extern "C" {

void ref_compute_rhs(double* rhs, double* u, long n, double* v1, double* v2, double k, double nu, double h)
{ compute_rhs(rhs, u, n, v1, v2, k, nu, h); }
void ref_residual(double* res, double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ residual(res, u, rhs, n, v1, v2, k, nu, h); }
double ref_compute_norm(double* res, long n) { return compute_norm(res, n); }
void ref_gauss_seidel(double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ gauss_seidel(u, rhs, n, v1, v2, k, nu, h); }
void ref_gauss_seidel2(double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ gauss_seidel2(u, rhs, n, v1, v2, k, nu, h); }
void ref_prolongation(double* up, double* u, int n) { prolongation(up, u, n); }
void ref_restriction(double* u, double* up, int n) { restriction(u, up, n); }

void ref_mg_inner(double** u, double** rhs, double** v1, double** v2, double* tmp, double dx, int n,
                  int lvl, int maxlvl, int shape, double dt, double nu)
{ mg_inner(u, rhs, v1, v2, tmp, dx, n, lvl, maxlvl, shape, dt, nu); }

void ref_mg_outer(double** utow, double** v1tow, double** v2tow, double** rhstow, double* tmp,
                  double nu, int maxlvl, int n, double dt, double dx, double tol, int shape)
{ mg_outer(utow, v1tow, v2tow, rhstow, tmp, nu, maxlvl, n, dt, dx, tol, shape); }

void ref_timestepper(double* uT, double* u0, double* v1, double* v2, double nu, int maxlvl, int n,
                     double dt, double T, double dx, double tol, int shape)
{ timestepper(uT, u0, v1, v2, nu, maxlvl, n, dt, T, dx, tol, shape); }

// The reference's mg_outer returns nothing.  This re-runs ITS loop (multigrid.cpp:104-114)
// with ITS functions and records the norms: hist[0] = res0, hist[k] = norm after cycle k.
int ref_mg_outer_hist(double** utow, double** v1tow, double** v2tow, double** rhstow, double* tmp,
                      double nu, int maxlvl, int n, double dt, double dx, double tol, int shape,
                      double* hist)
{
    residual(tmp, utow[0], rhstow[0], n, v1tow[0], v2tow[0], dt, nu, dx);
    double res0 = compute_norm(tmp, n), res = res0;
    hist[0] = res0;
    int it = 0;
    for (; it < 50 && res / res0 > tol; ++it) {
        mg_inner(utow, rhstow, v1tow, v2tow, tmp, dx, n, 0, maxlvl, shape, dt, nu);
        residual(tmp, utow[0], rhstow[0], n, v1tow[0], v2tow[0], dt, nu, dx);
        res = compute_norm(tmp, n);
        hist[it + 1] = res;
    }
    return it;
}

// The reference parallelises by calling the SAME functions from inside
// `omp parallel` + `omp single` so that their orphaned tasks fan out to the team
// (multigrid.cpp:252-258).  These two wrappers do exactly that with a caller-chosen team.
void ref_timestepper_omp(int nthreads, double* uT, double* u0, double* v1, double* v2, double nu,
                         int maxlvl, int n, double dt, double T, double dx, double tol, int shape)
{
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        { timestepper(uT, u0, v1, v2, nu, maxlvl, n, dt, T, dx, tol, shape); }
    }
}

int ref_mg_outer_hist_omp(int nthreads, double** utow, double** v1tow, double** v2tow, double** rhstow,
                          double* tmp, double nu, int maxlvl, int n, double dt, double dx, double tol,
                          int shape, double* hist)
{
    int it = 0;
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        { it = ref_mg_outer_hist(utow, v1tow, v2tow, rhstow, tmp, nu, maxlvl, n, dt, dx, tol, shape, hist); }
    }
    return it;
}

// one V-cycle + convergence check (mg_inner + residual + compute_norm, multigrid.cpp:110-113),
// the unit bench.py times; returns the residual norm.
double ref_cycle_and_norm_omp(int nthreads, double** utow, double** v1tow, double** v2tow, double** rhstow,
                              double* tmp, double nu, int maxlvl, int n, double dt, double dx, int shape)
{
    double rn = 0.0;
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        {
            mg_inner(utow, rhstow, v1tow, v2tow, tmp, dx, n, 0, maxlvl, shape, dt, nu);
            residual(tmp, utow[0], rhstow[0], n, v1tow[0], v2tow[0], dt, nu, dx);
            rn = compute_norm(tmp, n);
        }
    }
    return rn;
}

}  // extern "C"
