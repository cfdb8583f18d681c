// capi_drivers.cu -- the reference's driver functions with THEIR argument lists on CALLER-OWNED
// towers (mg_inner / mg_outer, multigrid.cpp:17-21,97-99 = multigrid.cu:17-21,101-103), and the
// gs.h operator flavour on HOST pointers (gs.h:3-17).
//
// mg_inner / mg_outer here run the reference's cycle operator by operator on the arrays the caller
// passes: dense row-major level arrays with the level's true stride n_l+1, towers of maxlvl device
// pointers, one spare array `tmp` of (n+1)^2 doubles re-used at every level with that level's
// stride -- exactly the data model of the reference (the caller may have built its coarse velocity
// towers the way multigrid.cpp:148-160 does; nothing here looks at how).  This is the drop-in for
// code that keeps its own towers; the fused, graph-captured solver (solver.cu) is reached through
// the handle API or mgb200_timestepper_*, which own their towers.
#include <cmath>
#include <vector>

#include "ops_basic.cuh"

using namespace mgb200;

namespace {

struct DriverOpt { int niter, coarse_maxit, max_cycle, arith; double coarse_tol; };

DriverOpt driver_opt(const mgb200_options* o)
{
    mgb200_options d;
    mgb200_default_options(&d);
    if (o && o->struct_size == (int)sizeof(mgb200_options)) d = *o;
    return DriverOpt{d.niter, d.coarse_maxit, d.max_cycle, d.arith, d.coarse_tol};
}

// residual ; compute_norm (gs.cpp:55-107) of one level -> host value
int residual_norm(double* tmp, const double* u, const double* rhs, const double* v1, const double* v2, long n, const Stencil& st,
                  int arith, double* ws, double* out, cudaStream_t s)
{
    const long cnt = residual_partials_count(n);
    MGB_TRY(launch_residual(tmp, u, rhs, v1, v2, n, natural_layout(n + 1), st, arith, ws, s));
    MGB_TRY(launch_reduce_partials(ws, cnt, ws + cnt, s));
    double h = 0.0;
    MGB_CUDA(cudaMemcpyAsync(&h, ws + cnt, sizeof(double), cudaMemcpyDeviceToHost, s));
    MGB_CUDA(cudaStreamSynchronize(s));
    *out = std::sqrt(h);
    return MGB200_OK;
}

int smooth(double* u, const double* rhs, const double* v1, const double* v2, long n, const Stencil& st, int iters, int arith,
           cudaStream_t s)
{
    const Layout L = natural_layout(n + 1);
    for (int it = 0; it < iters; ++it) {
        MGB_TRY(launch_gs_colour(u, rhs, v1, v2, n, L, st, 0, arith, s));
        MGB_TRY(launch_gs_colour(u, rhs, v1, v2, n, L, st, 1, arith, s));
    }
    return MGB200_OK;
}

// multigrid.cpp:17-92
int inner(double** u, double** rhs, double** v1, double** v2, double* tmp, double dx, long n, int lvl, int maxlvl, int shape,
          double dt, double nu, const DriverOpt& o, double* ws, cudaStream_t s)
{
    double *ui = u[lvl], *rhsi = rhs[lvl], *v1i = v1[lvl], *v2i = v2[lvl];
    const Stencil st = make_stencil(dt, nu, dx);
    const long nnew = n / 2;
    for (int sh = 0; sh < shape; ++sh) {                                          // :52
        if (lvl == maxlvl - 1) {
            // :55-65 -- the loop condition is tested before the first iteration with res = 1.0
            double res_exact = 1.0;
            for (int i = 0; i < o.coarse_maxit && res_exact > o.coarse_tol; ++i) {
                MGB_TRY(smooth(ui, rhsi, v1i, v2i, n, st, 1, o.arith, s));
                MGB_TRY(residual_norm(tmp, ui, rhsi, v1i, v2i, n, st, o.arith, ws, &res_exact, s));
            }
        } else {
            double *ui1 = u[lvl + 1], *rhsi1 = rhs[lvl + 1];
            MGB_TRY(smooth(ui, rhsi, v1i, v2i, n, st, o.niter, o.arith, s));      // :69-72
            MGB_TRY(launch_residual(tmp, ui, rhsi, v1i, v2i, n, natural_layout(n + 1), st, o.arith, nullptr, s));   // :73
            MGB_TRY(launch_restrict(rhsi1, natural_layout(nnew + 1), tmp, natural_layout(n + 1), n, s));            // :75
            MGB_CUDA(cudaMemsetAsync(ui1, 0, (size_t)(nnew + 1) * (nnew + 1) * sizeof(double), s));                 // :77
            MGB_TRY(inner(u, rhs, v1, v2, tmp, 2 * dx, nnew, lvl + 1, maxlvl, shape, dt, nu, o, ws, s));            // :79
            MGB_TRY(launch_prolong(ui, natural_layout(n + 1), ui1, natural_layout(nnew + 1), nnew, true, s));       // :81-83
            MGB_TRY(smooth(ui, rhsi, v1i, v2i, n, st, o.niter, o.arith, s));      // :85-88
        }
    }
    return MGB200_OK;
}

bool bad_tower(double** t, int maxlvl)
{
    if (!t) return true;
    for (int l = 0; l < maxlvl; ++l)
        if (!t[l]) return true;
    return false;
}

// gs.h flavour: operate on HOST arrays through temporary device copies
struct DevCopy {
    double* d = nullptr;
    size_t bytes = 0;
    cudaStream_t s;
    explicit DevCopy(cudaStream_t st) : s(st) {}
    int up(const double* h, size_t count)
    {
        bytes = count * sizeof(double);
        MGB_CUDA(cudaMallocAsync(&d, bytes, s));
        if (h) MGB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
        return MGB200_OK;
    }
    int down(double* h)
    {
        MGB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
        return MGB200_OK;
    }
    ~DevCopy() { if (d) cudaFreeAsync(d, s); }
};

}  // namespace

extern "C" {

int mgb200_mg_inner(double** u, double** rhs, double** v1, double** v2, double* tmp, double dx, int n, int lvl, int maxlvl,
                    int shape, double dt, double nu, const mgb200_options* opt, void* stream)
{
    if (maxlvl < 1 || lvl < 0 || lvl >= maxlvl || n < 4 || (n & 1) || shape < 1 || !tmp || bad_tower(u, maxlvl) || bad_tower(rhs, maxlvl) ||
        bad_tower(v1, maxlvl) || bad_tower(v2, maxlvl))
        return fail(MGB200_ERR_INVALID, "mg_inner: bad argument");
    if ((long)n >> (maxlvl - 1 - lvl) < 2) return fail(MGB200_ERR_INVALID, "mg_inner: too many levels for n");
    MGB_TRY(ops_basic_init());
    const DriverOpt o = driver_opt(opt);
    cudaStream_t s = (cudaStream_t)stream;
    double* ws = nullptr;
    MGB_CUDA(cudaMallocAsync(&ws, (size_t)(residual_partials_count(n) + 1) * sizeof(double), s));
    const int rc = inner(u, rhs, v1, v2, tmp, dx, n, lvl, maxlvl, shape, dt, nu, o, ws, s);
    cudaFreeAsync(ws, s);
    return rc;
}

int mgb200_mg_outer(double** utow, double** v1tow, double** v2tow, double** rhstow, double* tmp, double nu, int maxlvl, int n,
                    double dt, double dx, double tol, int shape, const mgb200_options* opt, void* stream, mgb200_solve_info* info)
{
    if (maxlvl < 1 || n < 4 || (n & 1) || shape < 1 || !tmp || bad_tower(utow, maxlvl) || bad_tower(rhstow, maxlvl) ||
        bad_tower(v1tow, maxlvl) || bad_tower(v2tow, maxlvl))
        return fail(MGB200_ERR_INVALID, "mg_outer: bad argument");
    MGB_TRY(ops_basic_init());
    const DriverOpt o = driver_opt(opt);
    cudaStream_t s = (cudaStream_t)stream;
    double* ws = nullptr;
    MGB_CUDA(cudaMallocAsync(&ws, (size_t)(residual_partials_count(n) + 1) * sizeof(double), s));
    const Stencil st = make_stencil(dt, nu, dx);
    mgb200_solve_info local{};
    double res = 0.0;
    int rc = residual_norm(tmp, utow[0], rhstow[0], v1tow[0], v2tow[0], n, st, o.arith, ws, &res, s);   // :104-105
    local.res0 = local.res = local.hist[0] = res;
    int iter = 0;
    // :108 (a NaN residual ends the loop exactly as in the reference: NaN > tol is false)
    for (; rc == MGB200_OK && iter < o.max_cycle && res / local.res0 > tol; ++iter) {
        rc = inner(utow, rhstow, v1tow, v2tow, tmp, dx, n, 0, maxlvl, shape, dt, nu, o, ws, s);          // :110
        if (rc == MGB200_OK) rc = residual_norm(tmp, utow[0], rhstow[0], v1tow[0], v2tow[0], n, st, o.arith, ws, &res, s);   // :112-113
        if (iter + 1 < 52) local.hist[iter + 1] = res;
    }
    cudaFreeAsync(ws, s);
    local.cycles = iter; local.res = res;
    local.converged = (res / local.res0 <= tol) ? 1 : 0;
    if (info) *info = local;
    if (rc == MGB200_OK) report_solve(iter, o.max_cycle, local.res0, res, tol);
    return rc;
}

// ---- gs.h:3-17 on HOST pointers --------------------------------------------------------------
int mgb200_host_residual(double* res, const double* u, const double* rhs, long n, const double* v1, const double* v2, double k,
                         double nu, double h, int arith)
{
    if (!res || !u || !rhs || !v1 || !v2 || n < 2) return fail(MGB200_ERR_INVALID, "host_residual: bad argument");
    const size_t m = (size_t)(n + 1) * (n + 1);
    cudaStream_t s = nullptr;
    DevCopy dr(s), du(s), df(s), d1(s), d2(s);
    MGB_TRY(dr.up(res, m)); MGB_TRY(du.up(u, m)); MGB_TRY(df.up(rhs, m)); MGB_TRY(d1.up(v1, m)); MGB_TRY(d2.up(v2, m));
    MGB_TRY(mgb200_residual(dr.d, du.d, df.d, n, n + 1, d1.d, d2.d, k, nu, h, arith, s));
    MGB_TRY(dr.down(res));
    MGB_CUDA(cudaStreamSynchronize(s));
    return MGB200_OK;
}

int mgb200_host_compute_norm(const double* res, long n, double* out)
{
    if (!res || !out || n < 2) return fail(MGB200_ERR_INVALID, "host_compute_norm: bad argument");
    cudaStream_t s = nullptr;
    DevCopy dr(s);
    MGB_TRY(dr.up(res, (size_t)(n + 1) * (n + 1)));
    return mgb200_compute_norm(dr.d, n, n + 1, out, s);
}

int mgb200_host_gauss_seidel(double* u, const double* rhs, long n, const double* v1, const double* v2, double k, double nu,
                             double h, int iters, int arith)
{
    if (!u || !rhs || !v1 || !v2 || n < 2) return fail(MGB200_ERR_INVALID, "host_gauss_seidel: bad argument");
    const size_t m = (size_t)(n + 1) * (n + 1);
    cudaStream_t s = nullptr;
    DevCopy du(s), df(s), d1(s), d2(s);
    MGB_TRY(du.up(u, m)); MGB_TRY(df.up(rhs, m)); MGB_TRY(d1.up(v1, m)); MGB_TRY(d2.up(v2, m));
    MGB_TRY(mgb200_gauss_seidel(du.d, df.d, n, n + 1, d1.d, d2.d, k, nu, h, iters, arith, s));
    MGB_TRY(du.down(u));
    MGB_CUDA(cudaStreamSynchronize(s));
    return MGB200_OK;
}

int mgb200_host_compute_rhs(double* rhs, const double* u, long n, const double* v1, const double* v2, double k, double nu, double h,
                            int arith)
{
    if (!rhs || !u || !v1 || !v2 || n < 2) return fail(MGB200_ERR_INVALID, "host_compute_rhs: bad argument");
    const size_t m = (size_t)(n + 1) * (n + 1);
    cudaStream_t s = nullptr;
    DevCopy df(s), du(s), d1(s), d2(s);
    MGB_TRY(df.up(rhs, m)); MGB_TRY(du.up(u, m)); MGB_TRY(d1.up(v1, m)); MGB_TRY(d2.up(v2, m));
    MGB_TRY(mgb200_compute_rhs(df.d, du.d, n, n + 1, d1.d, d2.d, k, nu, h, arith, s));
    MGB_TRY(df.down(rhs));
    MGB_CUDA(cudaStreamSynchronize(s));
    return MGB200_OK;
}

int mgb200_host_prolongation(double* up, const double* u, int n)
{
    if (!up || !u || n < 1) return fail(MGB200_ERR_INVALID, "host_prolongation: bad argument");
    cudaStream_t s = nullptr;
    DevCopy dup(s), du(s);
    MGB_TRY(dup.up(nullptr, (size_t)(2L * n + 1) * (2L * n + 1))); MGB_TRY(du.up(u, (size_t)(n + 1L) * (n + 1L)));
    MGB_TRY(mgb200_prolongation(dup.d, 2L * n + 1, du.d, n + 1L, n, s));
    MGB_TRY(dup.down(up));
    MGB_CUDA(cudaStreamSynchronize(s));
    return MGB200_OK;
}

int mgb200_host_restriction(double* u, const double* up, int n)
{
    if (!up || !u || n < 2 || (n & 1)) return fail(MGB200_ERR_INVALID, "host_restriction: bad argument");
    cudaStream_t s = nullptr;
    DevCopy du(s), dup(s);
    MGB_TRY(du.up(nullptr, (size_t)(n / 2 + 1L) * (n / 2 + 1L))); MGB_TRY(dup.up(up, (size_t)(n + 1L) * (n + 1L)));
    MGB_TRY(mgb200_restriction(du.d, n / 2 + 1L, dup.d, n + 1L, n, s));
    MGB_TRY(du.down(u));
    MGB_CUDA(cudaStreamSynchronize(s));
    return MGB200_OK;
}

}  // extern "C"
