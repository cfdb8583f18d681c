// include/compat/gs.h -- source-level stand-in for the reference's gs.h (gs.h:3-17): the CPU
// operator API on HOST pointers, served by the B200 library.
//
// With  g++ -fopenmp -I<this directory> multigrid.cpp -lmgb200  the reference's UNMODIFIED CPU
// driver (multigrid.cpp: mg_inner / mg_outer / timestepper / main) compiles and runs with every
// grid operator executed on the GPU: each call copies its operands to the device, runs the sm_100a
// kernel of include/mgb200.h and copies the result back (mgb200_host_*).  That is an operator-level
// drop-in for tests and for code that must keep host-resident towers; it moves every field over
// PCIe for every operator, so a time loop belongs on mgb200_timestepper_host or the handle API,
// which keep the towers in HBM.  Arithmetic: MGB200_ARITH_EXACT (bit-identical to gs.cpp).
// No CUDA header is needed: plain C++ over the C ABI.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "../mgb200.h"

namespace mgb200_gs_detail {
inline void check(int rc, const char* what)
{
    if (rc != MGB200_OK) { std::fprintf(stderr, "mgb200 %s: %s\n", what, mgb200_last_error()); std::exit(1); }
}
}  // namespace mgb200_gs_detail

// gs.h:3
static inline void residual(double* res, double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ mgb200_gs_detail::check(mgb200_host_residual(res, u, rhs, n, v1, v2, k, nu, h, MGB200_ARITH_EXACT), "residual"); }

// gs.h:5
static inline double compute_norm(double* res, long n)
{
    double out = 0.0;
    mgb200_gs_detail::check(mgb200_host_compute_norm(res, n, &out), "compute_norm");
    return out;
}

// gs.h:9
static inline void gauss_seidel(double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ mgb200_gs_detail::check(mgb200_host_gauss_seidel(u, rhs, n, v1, v2, k, nu, h, 1, MGB200_ARITH_EXACT), "gauss_seidel"); }

// gs.h:10 (the parallel-for variant of the same iteration; unused by the reference driver)
static inline void gauss_seidel2(double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ gauss_seidel(u, rhs, n, v1, v2, k, nu, h); }

// gs.h:13
static inline void compute_rhs(double* rhs, double* u, long n, double* v1, double* v2, double k, double nu, double h)
{ mgb200_gs_detail::check(mgb200_host_compute_rhs(rhs, u, n, v1, v2, k, nu, h, MGB200_ARITH_EXACT), "compute_rhs"); }

// gs.h:16 (n = coarse n; up holds (2n+1)^2 doubles)
static inline void prolongation(double* up, double* u, int n)
{ mgb200_gs_detail::check(mgb200_host_prolongation(up, u, n), "prolongation"); }

// gs.h:17 (n = fine n; u holds (n/2+1)^2 doubles)
static inline void restriction(double* u, double* up, int n)
{ mgb200_gs_detail::check(mgb200_host_restriction(u, up, n), "restriction"); }
