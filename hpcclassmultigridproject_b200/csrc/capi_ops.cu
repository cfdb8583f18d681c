// capi_ops.cu -- C-ABI operator entry points on caller-owned DEVICE arrays in the reference's
// natural layout (element (i,j) at p[i*ld+j]).  They replace the gs.h / gscu.h operators one for
// one; see include/mgb200.h for the file:line of each reference interface.
#include <cmath>

#include "ops_basic.cuh"

using namespace mgb200;

namespace {

// scratch for block partial sums (+1 slot for the reduced value): allocated and released in stream order on
// the caller's stream (cudaMallocAsync), so concurrent calls on different streams or devices never share it
struct Scratch {
    double* buf = nullptr;
    cudaStream_t s;
    explicit Scratch(cudaStream_t st) : s(st) {}
    int get(long need)
    {
        MGB_CUDA(cudaMallocAsync(&buf, (size_t)need * sizeof(double), s));
        return MGB200_OK;
    }
    ~Scratch() { if (buf) cudaFreeAsync(buf, s); }
};

inline bool bad_grid(long n, long ld) { return n < 2 || ld < n + 1; }

}  // namespace

extern "C" {

int mgb200_gauss_seidel(double* u, const double* rhs, long n, long ld, const double* v1, const double* v2, double dt,
                        double nu, double dx, int iters, int arith, void* stream)
{
    if (!u || !rhs || !v1 || !v2 || bad_grid(n, ld) || iters < 0) return fail(MGB200_ERR_INVALID, "gauss_seidel: bad argument");
    const Stencil st = make_stencil(dt, nu, dx);
    const Layout L = natural_layout(ld);
    for (int it = 0; it < iters; ++it) {
        MGB_TRY(launch_gs_colour(u, rhs, v1, v2, n, L, st, 0, arith, (cudaStream_t)stream));   // red: (i+j) even
        MGB_TRY(launch_gs_colour(u, rhs, v1, v2, n, L, st, 1, arith, (cudaStream_t)stream));   // black
    }
    return MGB200_OK;
}

int mgb200_residual(double* res, const double* u, const double* rhs, long n, long ld, const double* v1, const double* v2,
                    double dt, double nu, double dx, int arith, void* stream)
{
    if (!res || !u || !rhs || !v1 || !v2 || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "residual: bad argument");
    return launch_residual(res, u, rhs, v1, v2, n, natural_layout(ld), make_stencil(dt, nu, dx), arith, nullptr,
                           (cudaStream_t)stream);
}

int mgb200_norm2_async(const double* a, long n, long ld, double* out_dev, void* stream)
{
    if (!a || !out_dev || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "norm2: bad argument");
    const long cnt = residual_partials_count(n);
    Scratch sc((cudaStream_t)stream);
    MGB_TRY(sc.get(cnt + 1));
    double* ws = sc.buf;
    MGB_TRY(launch_square_partials(a, n, natural_layout(ld), ws, (cudaStream_t)stream));
    return launch_reduce_partials(ws, cnt, out_dev, (cudaStream_t)stream);
}

int mgb200_compute_norm(const double* a, long n, long ld, double* out_host, void* stream)
{
    if (!a || !out_host || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "compute_norm: bad argument");
    const long cnt = residual_partials_count(n);
    Scratch sc((cudaStream_t)stream);
    MGB_TRY(sc.get(cnt + 1));
    double* ws = sc.buf;
    MGB_TRY(launch_square_partials(a, n, natural_layout(ld), ws, (cudaStream_t)stream));
    MGB_TRY(launch_reduce_partials(ws, cnt, ws + cnt, (cudaStream_t)stream));
    double h = 0.0;
    MGB_CUDA(cudaMemcpyAsync(&h, ws + cnt, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *out_host = std::sqrt(h);                                                       // gs.cpp:106
    return MGB200_OK;
}

int mgb200_residual_norm2_async(double* res, const double* u, const double* rhs, long n, long ld, const double* v1,
                                const double* v2, double dt, double nu, double dx, int arith, double* out_dev,
                                void* stream)
{
    if (!u || !rhs || !v1 || !v2 || !out_dev || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "residual_norm2: bad argument");
    const long cnt = residual_partials_count(n);
    Scratch sc((cudaStream_t)stream);
    MGB_TRY(sc.get(cnt + 1));
    double* ws = sc.buf;
    MGB_TRY(launch_residual(res, u, rhs, v1, v2, n, natural_layout(ld), make_stencil(dt, nu, dx), arith, ws,
                            (cudaStream_t)stream));
    return launch_reduce_partials(ws, cnt, out_dev, (cudaStream_t)stream);
}

int mgb200_compute_rhs(double* rhs, const double* u, long n, long ld, const double* v1, const double* v2, double dt,
                       double nu, double dx, int arith, void* stream)
{
    if (!rhs || !u || !v1 || !v2 || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "compute_rhs: bad argument");
    return launch_compute_rhs(rhs, u, v1, v2, n, natural_layout(ld), make_stencil(dt, nu, dx), arith, nullptr,
                              (cudaStream_t)stream);
}

int mgb200_restriction(double* coarse, long ldc, const double* fine, long ldf, long nf, void* stream)
{
    if (!coarse || !fine || nf < 2 || (nf & 1) || ldf < nf + 1 || ldc < nf / 2 + 1) return fail(MGB200_ERR_INVALID, "restriction: bad argument");
    return launch_restrict(coarse, natural_layout(ldc), fine, natural_layout(ldf), nf, (cudaStream_t)stream);
}

int mgb200_restriction_fw(double* coarse, long ldc, const double* fine, long ldf, long nf, void* stream)
{
    if (!coarse || !fine || nf < 2 || (nf & 1) || ldf < nf + 1 || ldc < nf / 2 + 1) return fail(MGB200_ERR_INVALID, "restriction_fw: bad argument");
    return launch_restrict_fw(coarse, natural_layout(ldc), fine, natural_layout(ldf), nf, false, (cudaStream_t)stream);
}

int mgb200_prolongation(double* fine, long ldf, const double* coarse, long ldc, long nc, void* stream)
{
    if (!coarse || !fine || nc < 1 || ldf < 2 * nc + 1 || ldc < nc + 1) return fail(MGB200_ERR_INVALID, "prolongation: bad argument");
    return launch_prolong(fine, natural_layout(ldf), coarse, natural_layout(ldc), nc, false, (cudaStream_t)stream);
}

int mgb200_prolong_correct(double* u_fine, long ldf, const double* coarse, long ldc, long nc, void* stream)
{
    if (!coarse || !u_fine || nc < 1 || ldf < 2 * nc + 1 || ldc < nc + 1) return fail(MGB200_ERR_INVALID, "prolong_correct: bad argument");
    return launch_prolong(u_fine, natural_layout(ldf), coarse, natural_layout(ldc), nc, true, (cudaStream_t)stream);
}

int mgb200_vecadd(double* c, const double* a, const double* b, long n, long ld, void* stream)
{
    if (!a || !b || !c || n < 0 || ld < n + 1) return fail(MGB200_ERR_INVALID, "vecadd: bad argument");
    return launch_vecadd(c, a, b, n, natural_layout(ld), (cudaStream_t)stream);
}

int mgb200_initial_conditions(double* u0, double* v1, double* v2, long n, long ld, double vscale, void* stream)
{
    if (!u0 || !v1 || !v2 || bad_grid(n, ld)) return fail(MGB200_ERR_INVALID, "initial_conditions: bad argument");
    return launch_initial_conditions(u0, v1, v2, n, natural_layout(ld), vscale, (cudaStream_t)stream);
}

int mgb200_initial_conditions_rows(double* u0, double* v1, double* v2, long n, long ld, double vscale, long row_lo, long row_hi,
                                   void* stream)
{
    if (!u0 || !v1 || !v2 || bad_grid(n, ld) || row_lo < 0 || row_hi > n || row_lo > row_hi)
        return fail(MGB200_ERR_INVALID, "initial_conditions_rows: bad argument");
    return launch_initial_conditions(u0, v1, v2, n, natural_layout(ld, row_lo), vscale, (cudaStream_t)stream, row_lo, row_hi);
}

}  // extern "C"
