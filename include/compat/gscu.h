// include/compat/gscu.h -- source-level stand-in for the reference's gscu.h (gscu.h:3-16).
//
// With  nvcc -I<this directory> multigrid.cu -lmgb200  the reference's UNMODIFIED CUDA driver
// (multigrid.cu: mg_inner / mg_outer / timestepper / main) compiles and runs on the B200 library:
//   * the host launchers  gauss_seidel(...)  and  compute_norm(...)  (gscu.h:15-16) call the C ABI
//     (include/mgb200.h) and therefore the sm_100a kernels of libmgb200;
//   * the __global__ operators the reference launches itself with <<<grid, block>>> keep their names
//     and argument lists.  They are geometry-agnostic shims: whatever 1-D/2-D launch shape the caller
//     picked, each thread takes the node with its LINEAR thread index, so accesses are coalesced
//     (the reference maps threadIdx.x to the row: stride-(N+1) accesses, gs.cu:313-314).  The
//     reference's launches always supply at least one thread per node the kernel has to write
//     (multigrid.cu:51-53,66,77,79,85,87,148-150,176,233-235).  Arithmetic follows gs.cpp's
//     expression order with unfused IEEE operations, like MGB200_ARITH_EXACT.
//   * compute_norm follows gs.cpp:86-107 (interior only, input preserved), not the destructive
//     gs.cu:45-60; gaussian_u0 produces the CPU initial condition (multigrid.cpp:219-233), not the
//     out-of-bounds variant of gs.cu:221-230.
// This path keeps the reference's one-operator-per-launch structure; the fused solver is reached
// through mgb200_timestepper_device / mgb200_compat.hpp instead.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "../mgb200.h"

namespace mgb200_gscu_detail {
__device__ __forceinline__ long linear_tid()
{
    const long block = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const long nthr = (long)blockDim.x * blockDim.y;
    return block * nthr + (long)threadIdx.y * blockDim.x + threadIdx.x;
}
__device__ __forceinline__ long total_threads() { return (long)gridDim.x * gridDim.y * blockDim.x * blockDim.y; }
// gs.cpp:9-20 in the reference's rounding order
__device__ __forceinline__ double cminus(double v, double nu, double h, double r) { return __dmul_rn(r, __dadd_rn(__dmul_rn(__dmul_rn(-v, h), 0.5), nu)); }
__device__ __forceinline__ double cplus(double v, double nu, double h, double r) { return __dmul_rn(r, __dadd_rn(__dmul_rn(__dmul_rn(v, h), 0.5), nu)); }
inline void check(int rc, const char* what)
{
    if (rc != MGB200_OK) { std::fprintf(stderr, "mgb200 %s: %s\n", what, mgb200_last_error()); std::exit(1); }
}
}  // namespace mgb200_gscu_detail

// a = b + c (gs.cu:7-11)
static __global__ void vecadd(double* a, double* b, double* c, long n)
{
    for (long i = mgb200_gscu_detail::linear_tid(); i < n; i += mgb200_gscu_detail::total_threads()) a[i] = __dadd_rn(b[i], c[i]);
}

// a[i] *= a[i] (gs.cu:13-17)
static __global__ void square_ker(double* a, long n)
{
    for (long i = mgb200_gscu_detail::linear_tid(); i < n; i += mgb200_gscu_detail::total_threads()) a[i] = __dmul_rn(a[i], a[i]);
}

// sum[b] = a[b*B] + ... + a[b*B + B - 1] for block b of B = blockDim.x*blockDim.y threads (gs.cu:25-43): the block
// layout IS the contract here, so this shim keeps it (pairwise tree like the reference, warp shuffles instead of
// shared-memory halving).  The reference calls it in place (sum == a, gs.cu:54), where block b's store to a[b] races
// with block 0's read of the same element; that hazard belongs to the call pattern, not to the kernel, and
// compute_norm() below does not use it.
static __global__ void reduction_kernel(double* sum, const double* a, long N)
{
    __shared__ double warp_sums[32];
    const long nthr = (long)blockDim.x * blockDim.y;
    const long tid = (long)threadIdx.y * blockDim.x + threadIdx.x;
    const long block = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const long idx = block * nthr + tid;
    double v = idx < N ? a[idx] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_sums[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        v = tid < (nthr + 31) / 32 ? warp_sums[tid] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (tid == 0) sum[block] = v;
    }
}

static __global__ void gpucopy(double* dest, const double* source, long n)
{
    for (long i = mgb200_gscu_detail::linear_tid(); i < n; i += mgb200_gscu_detail::total_threads()) dest[i] = source[i];
}

// bilinear interpolation, one thread per COARSE node writing its up-to-four fine children (gs.cu:63-81)
static __global__ void prolongation(double* up, double* u, int n)
{
    const long m = (long)(n + 1) * (n + 1), nn = 2L * n + 1;
    for (long t = mgb200_gscu_detail::linear_tid(); t < m; t += mgb200_gscu_detail::total_threads()) {
        const long i = t / (n + 1), j = t % (n + 1);
        const double c00 = u[i * (n + 1) + j];
        up[2 * i * nn + 2 * j] = c00;
        if (i < n) up[(2 * i + 1) * nn + 2 * j] = __dmul_rn(__dadd_rn(c00, u[(i + 1) * (n + 1) + j]), 0.5);
        if (j < n) up[2 * i * nn + 2 * j + 1] = __dmul_rn(__dadd_rn(c00, u[i * (n + 1) + j + 1]), 0.5);
        if (i < n && j < n)
            up[(2 * i + 1) * nn + 2 * j + 1] = __dmul_rn(
                __dadd_rn(__dadd_rn(__dadd_rn(c00, u[(i + 1) * (n + 1) + j]), u[i * (n + 1) + j + 1]), u[(i + 1) * (n + 1) + j + 1]), 0.25);
    }
}

// injection (gs.cu:83-92)
static __global__ void restriction(double* u, double* up, int n)
{
    const long nc = n / 2 + 1, m = nc * nc;
    for (long t = mgb200_gscu_detail::linear_tid(); t < m; t += mgb200_gscu_detail::total_threads()) {
        const long i = t / nc, j = t % nc;
        u[i * nc + j] = up[(2 * i) * (long)(n + 1) + 2 * j];
    }
}

// rhs = B u on the interior (gs.cu:94-155, formula gs.cpp:44)
static __global__ void compute_rhs(double* rhs, double* u, int N, double* v1, double* v2, double dt, double nu, double dx)
{
    using namespace mgb200_gscu_detail;
    const long ni = N - 1, m = ni * ni, ld = N + 1;
    const double r = 0.5 * dt / (dx * dx), diag = __dadd_rn(1.0, __dmul_rn(__dmul_rn(4.0, r), nu));
    for (long t = linear_tid(); t < m; t += total_threads()) {
        const long p = (1 + t / ni) * ld + 1 + t % ni;
        const double a = cminus(v2[p], nu, dx, r), b = cplus(v2[p], nu, dx, r), c = cminus(v1[p], nu, dx, r), d = cplus(v1[p], nu, dx, r);
        double s = __dmul_rn(diag, u[p]);
        s = __dsub_rn(s, __dmul_rn(c, u[p - ld])); s = __dsub_rn(s, __dmul_rn(a, u[p - 1]));
        s = __dsub_rn(s, __dmul_rn(d, u[p + ld])); s = __dsub_rn(s, __dmul_rn(b, u[p + 1]));
        rhs[p] = s;
    }
}

// r = rhs - A u on the interior (gs.cu:157-218, formula gs.cpp:75)
static __global__ void residual(double* res, double* u, double* rhs, int N, double* v1, double* v2, double dt, double nu, double dx)
{
    using namespace mgb200_gscu_detail;
    const long ni = N - 1, m = ni * ni, ld = N + 1;
    const double r = 0.5 * dt / (dx * dx), diag = __dsub_rn(1.0, __dmul_rn(__dmul_rn(4.0, r), nu));
    for (long t = linear_tid(); t < m; t += total_threads()) {
        const long p = (1 + t / ni) * ld + 1 + t % ni;
        const double a = cminus(v2[p], nu, dx, r), b = cplus(v2[p], nu, dx, r), c = cminus(v1[p], nu, dx, r), d = cplus(v1[p], nu, dx, r);
        double s = __dmul_rn(diag, u[p]);
        s = __dadd_rn(s, __dmul_rn(c, u[p - ld])); s = __dadd_rn(s, __dmul_rn(a, u[p - 1]));
        s = __dadd_rn(s, __dmul_rn(d, u[p + ld])); s = __dadd_rn(s, __dmul_rn(b, u[p + 1]));
        res[p] = __dsub_rn(rhs[p], s);
    }
}

// one colour of RB-GS (gs.cu:307-376, formula gs.cpp:130); rb = (i+j) % 2 of the nodes updated
static __global__ void gs_ker(double* u, double* rhs, long N, double* v1, double* v2, double dt, double nu, double dx, int rb)
{
    using namespace mgb200_gscu_detail;
    const long ni = N - 1, m = ni * ni, ld = N + 1;
    const double r = 0.5 * dt / (dx * dx), diag = __dsub_rn(1.0, __dmul_rn(__dmul_rn(4.0, r), nu));
    for (long t = linear_tid(); t < m; t += total_threads()) {
        const long i = 1 + t / ni, j = 1 + t % ni, p = i * ld + j;
        if (((i + j) & 1) != rb) continue;
        const double a = cminus(v2[p], nu, dx, r), b = cplus(v2[p], nu, dx, r), c = cminus(v1[p], nu, dx, r), d = cplus(v1[p], nu, dx, r);
        double s = __dsub_rn(rhs[p], __dmul_rn(c, u[p - ld]));
        s = __dsub_rn(s, __dmul_rn(a, u[p - 1])); s = __dsub_rn(s, __dmul_rn(d, u[p + ld])); s = __dsub_rn(s, __dmul_rn(b, u[p + 1]));
        u[p] = __ddiv_rn(s, diag);
    }
}

// CPU initial condition (multigrid.cpp:219, :227-233): Gaussian, boundary lines zeroed with i < N
static __global__ void gaussian_u0(double* u0, double x0, double y0, double sigma, int n, double dx)
{
    const long m = (long)(n + 1) * (n + 1);
    for (long t = mgb200_gscu_detail::linear_tid(); t < m; t += mgb200_gscu_detail::total_threads()) {
        const long i = t / (n + 1), j = t % (n + 1);
        const bool zeroed = (i == 0 && j < n) || (j == n && i < n) || (i == n && j >= 1) || (j == 0 && i < n);
        u0[t] = zeroed ? 0.0 : exp(-sigma * ((i * dx - x0) * (i * dx - x0) + (j * dx - y0) * (j * dx - y0)));
    }
}

static __global__ void rotating_v(double* v1, double* v2, double kx, double ky, int n, double dx)
{
    const long m = (long)(n + 1) * (n + 1);
    for (long t = mgb200_gscu_detail::linear_tid(); t < m; t += mgb200_gscu_detail::total_threads()) {
        const long i = t / (n + 1), j = t % (n + 1);
        v1[t] = -ky * sin(kx * i * dx) * cos(ky * j * dx);
        v2[t] = kx * cos(kx * i * dx) * sin(ky * j * dx);
    }
}

// host launchers (gscu.h:15-16) on the library's kernels, default stream like the reference
static inline double compute_norm(double* a, int N)
{
    double out = 0.0;
    mgb200_gscu_detail::check(mgb200_compute_norm(a, N, N + 1, &out, nullptr), "compute_norm");
    return out;
}

static inline void gauss_seidel(double* u, double* rhs, long N, double* v1, double* v2, double dt, double nu, double dx)
{
    mgb200_gscu_detail::check(mgb200_gauss_seidel(u, rhs, N, N + 1, v1, v2, dt, nu, dx, 1, MGB200_ARITH_EXACT, nullptr), "gauss_seidel");
}
