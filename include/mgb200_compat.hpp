// mgb200_compat.hpp -- the reference's own spellings on top of the C ABI (header-only C++).
//
// A maintainer of soniareilly/HPCClassMultigridProject who wants the B200 path replaces
//     #include "gscu.h"            (multigrid.cu:12)
// with
//     #include "mgb200_compat.hpp"
//     using namespace mgb200::compat;
// and links -lmgb200 instead of gscu.o.  The DRIVER functions keep the reference argument lists
// exactly (multigrid.cu:17-21, 101-103, 130-132); the OPERATORS keep the gs.h argument lists
// (gs.h:3-17) but take DEVICE pointers, like the host launchers of gscu.h:15-16.  The reference
// launches most operators as __global__ kernels with a caller-chosen grid (gscu.h:3-14); here the
// launch geometry is the library's business, so those become plain host calls:
//
//     residual<<<numBlocks,threadsPerBlock>>>(tmp,u,rhs,n,v1,v2,dt,nu,dx)   ->   residual(tmp,u,rhs,n,v1,v2,dt,nu,dx)
//
// (see INTEGRATION.md for the full mapping).  Error behaviour: the reference returns void and
// checks nothing (multigrid.cu:202-208 is defined but unused); these wrappers throw
// std::runtime_error with mgb200_last_error() so a failure cannot pass silently.
#pragma once
#include <stdexcept>
#include <string>

#include "mgb200.h"

namespace mgb200 {
namespace compat {

// arithmetic used by the wrappers; set to MGB200_ARITH_EXACT for results bit-identical to gs.cpp
inline int& arithmetic()
{
    static int a = MGB200_ARITH_FAST;
    return a;
}

inline void check(int rc, const char* what)
{
    if (rc != MGB200_OK) throw std::runtime_error(std::string(what) + ": " + mgb200_last_error());
}

// ---- operators, gs.h:3-17 argument lists, device pointers, ld = n+1 ------------------------
inline void residual(double* res, double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ check(mgb200_residual(res, u, rhs, n, n + 1, v1, v2, k, nu, h, arithmetic(), nullptr), "residual"); }

// gs.h:5 semantics (interior only, input preserved) -- not the destructive gs.cu:45-60
inline double compute_norm(double* res, long n)
{
    double out = 0.0;
    check(mgb200_compute_norm(res, n, n + 1, &out, nullptr), "compute_norm");
    return out;
}

inline void gauss_seidel(double* u, double* rhs, long n, double* v1, double* v2, double k, double nu, double h)
{ check(mgb200_gauss_seidel(u, rhs, n, n + 1, v1, v2, k, nu, h, 1, arithmetic(), nullptr), "gauss_seidel"); }

inline void compute_rhs(double* rhs, double* u, long n, double* v1, double* v2, double k, double nu, double h)
{ check(mgb200_compute_rhs(rhs, u, n, n + 1, v1, v2, k, nu, h, arithmetic(), nullptr), "compute_rhs"); }

// up: (2n+1)^2 output, u: (n+1)^2 input (gs.h:16)
inline void prolongation(double* up, double* u, int n)
{ check(mgb200_prolongation(up, 2L * n + 1, u, n + 1L, n, nullptr), "prolongation"); }

// u: (n/2+1)^2 output, up: (n+1)^2 input (gs.h:17)
inline void restriction(double* u, double* up, int n)
{ check(mgb200_restriction(u, n / 2 + 1L, up, n + 1L, n, nullptr), "restriction"); }

// a = b + c over n entries (gscu.h:3, gs.cu:7-11: the FIRST argument is the output).  n must be a
// whole field, (m+1)^2.
inline void vecadd(double* a, double* b, double* c, long n)
{
    long m = 0;
    while ((m + 1) * (m + 1) < n) ++m;
    if ((m + 1) * (m + 1) != n) throw std::runtime_error("vecadd: length is not a whole (m+1)^2 field");
    check(mgb200_vecadd(a, b, c, m, m + 1, nullptr), "vecadd");
}

// ---- drivers, reference argument lists (multigrid.cu:130-132: DEVICE pointers) --------------
inline void timestepper(double* uT, double* u0, double* v1, double* v2, double nu, int maxlvl, int n, double dt,
                        double T, double dx, double tol, int shape)
{
    mgb200_options o;
    mgb200_default_options(&o);
    o.arith = arithmetic();
    check(mgb200_timestepper_device(uT, u0, v1, v2, nu, maxlvl, n, dt, T, dx, tol, shape, &o, nullptr), "timestepper");
}

// multigrid.cpp:124-126 flavour: HOST pointers
inline void timestepper_host(double* uT, double* u0, double* v1, double* v2, double nu, int maxlvl, int n, double dt,
                             double T, double dx, double tol, int shape)
{
    mgb200_options o;
    mgb200_default_options(&o);
    o.arith = arithmetic();
    check(mgb200_timestepper_host(uT, u0, v1, v2, nu, maxlvl, n, dt, T, dx, tol, shape, &o, nullptr), "timestepper");
}

// mg_inner / mg_outer with the reference argument lists on CALLER-OWNED towers (multigrid.cu:17-21,
// 101-103: host arrays of device pointers, level l dense with stride n_l+1, one spare array tmp).
// The cycle runs operator by operator on those arrays, in the reference's order.
inline void mg_inner(double** u, double** rhs, double** v1, double** v2, double* tmp, double dx, int n, int lvl, int maxlvl,
                     int shape, double dt, double nu)
{
    mgb200_options o;
    mgb200_default_options(&o);
    o.arith = arithmetic();
    check(mgb200_mg_inner(u, rhs, v1, v2, tmp, dx, n, lvl, maxlvl, shape, dt, nu, &o, nullptr), "mg_inner");
}

inline void mg_outer(double** utow, double** v1tow, double** v2tow, double** rhstow, double* tmp, double nu, int maxlvl, int n,
                     double dt, double dx, double tol, int shape)
{
    mgb200_options o;
    mgb200_default_options(&o);
    o.arith = arithmetic();
    check(mgb200_mg_outer(utow, v1tow, v2tow, rhstow, tmp, nu, maxlvl, n, dt, dx, tol, shape, &o, nullptr, nullptr), "mg_outer");
}

// The FUSED solver owns its towers (split layout, no (N/2+1)^2 over-allocation, one CUDA graph per
// cycle), so its mg_inner / mg_outer are methods of a handle created from the level-0 fields:
class Multigrid {
public:
    Multigrid(int n, int maxlvl, double nu, double dt, double dx, double tol, int shape)
    {
        mgb200_options o;
        mgb200_default_options(&o);
        o.arith = arithmetic();
        o.shape = shape;
        check(mgb200_create(&h_, n, maxlvl, nu, dt, dx, tol, &o), "mgb200_create");
    }
    ~Multigrid() { mgb200_destroy(h_); }
    Multigrid(const Multigrid&) = delete;
    Multigrid& operator=(const Multigrid&) = delete;
    // timestepper prologue (multigrid.cu:143-170): device fields, ld = n+1
    void set_fields(double* u0, double* v1, double* v2, int n) { check(mgb200_set_fields_device(h_, u0, v1, v2, n + 1L), "set_fields"); }
    // compute_rhs (multigrid.cu:176)
    double compute_rhs() { double r = 0; check(mgb200_form_rhs(h_, &r), "form_rhs"); return r; }
    // mg_inner at level 0 + residual + compute_norm (multigrid.cu:116-119); returns the norm
    double mg_inner() { double r = 0; check(mgb200_cycle(h_, &r), "cycle"); return r; }
    // mg_outer (multigrid.cu:101-126); returns the number of cycles
    int mg_outer(mgb200_solve_info* info = nullptr)
    {
        mgb200_solve_info local;
        check(mgb200_solve(h_, info ? info : &local), "solve");
        return (info ? info : &local)->cycles;
    }
    void get_u(double* uT, int n) { check(mgb200_get_u_device(h_, uT, n + 1L), "get_u"); }
    mgb200_solver* handle() { return h_; }

private:
    mgb200_solver* h_ = nullptr;
};

}  // namespace compat
}  // namespace mgb200
