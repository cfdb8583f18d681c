// common.cuh -- shared device/host helpers: error handling, field layouts, stencil arithmetic.
//
// Stencil (reference gs.cpp:9-20, :44, :75, :130), with r = dt/(2 dx^2):
//   west  a = r(-v2 dx/2 + nu)   east  b = r(+v2 dx/2 + nu)
//   north c = r(-v1 dx/2 + nu)   south d = r(+v1 dx/2 + nu)
//   (A u)_ij = (1-4 r nu) u_ij + c u_{i-1,j} + a u_{i,j-1} + d u_{i+1,j} + b u_{i,j+1}
//   (B u)_ij = (1+4 r nu) u_ij - c u_{i-1,j} - a u_{i,j-1} - d u_{i+1,j} - b u_{i,j+1}
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>

#include "../../include/mgb200.h"

namespace mgb200 {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local text behind mgb200_last_error)
void set_error(const std::string& msg);
int  fail(int code, const std::string& msg);

#define MGB_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::mgb200::fail(MGB200_ERR_CUDA, std::string(#expr) + ": " +             \
                                  cudaGetErrorString(_e) + " (" + __FILE__ + ":" +         \
                                  std::to_string(__LINE__) + ")");                         \
    } while (0)

#define MGB_TRY(expr)                     \
    do {                                  \
        int _rc = (expr);                 \
        if (_rc != MGB200_OK) return _rc; \
    } while (0)

// every kernel launch goes through check_launch: error check + launch accounting
long& launch_counter();
static inline int check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MGB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    ++launch_counter();
    return MGB200_OK;
}

// end-of-solve diagnostics shared by the drivers: the reference's only failure report is the warning of
// multigrid.cpp:117-119 (which, by an off-by-one, fires on convergence in exactly MAX_CYCLE-1 cycles and not
// on hitting MAX_CYCLE); here the warning goes to stderr when the loop ended WITHOUT meeting the tolerance,
// and a non-finite residual norm (the reference then just leaves its loop: NaN > tol is false) is named.
// Returns false in those cases.  MGB200_QUIET=1 silences the text.
bool report_solve(int cycles, int max_cycle, double res0, double res, double tol);

static inline long round_up(long x, long m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// Field layouts.
//   natural: element (i,j) at i*pitch + j                         (the reference's, pitch = n+1)
//   split  : element (i,j) at i*pitch + (j&1)*odd + (j>>1)        (solver-internal)
// The split layout stores the even-column and odd-column nodes of a row in two contiguous
// runs, so that one colour of a red-black half-sweep is a unit-stride vector and a row segment
// of one parity is a single 16-byte-aligned bulk (TMA) copy.  pitch and odd are multiples of 16
// doubles (128 B).
struct Layout {
    long pitch;
    long odd;        // < 0 : natural layout
    long row0 = 0;   // global index of the first row held by the array (row slabs; 0 = whole field)
    __host__ __device__ __forceinline__ long at(long i, long j) const
    {
        return odd < 0 ? (i - row0) * pitch + j : (i - row0) * pitch + (j & 1) * odd + (j >> 1);
    }
    __host__ __device__ __forceinline__ bool split() const { return odd >= 0; }
};

static inline Layout natural_layout(long ld, long row0 = 0) { return Layout{ld, -1, row0}; }
// n even.  Even run: n/2+1 nodes, odd run: n/2 nodes; 32 doubles of slack after each run so that
// aligned, slightly over-long bulk copies of the streaming kernels stay inside the row.
static inline Layout split_layout(long n)
{
    long odd = round_up(n / 2 + 1, 16) + 32;
    return Layout{2 * odd, odd, 0};
}
static inline size_t layout_elems(const Layout& L, long n) { return (size_t)L.pitch * (size_t)(n + 1); }

// ---------------------------------------------------------------------------------------------
// Stencil constants of one level, prepared on the host in the reference's expression order.
struct Stencil {
    double r;         // 0.5*dt/(dx*dx)                 gs.cpp:10
    double nu, h;
    double diag;      // 1.0 - 4.0*r*nu                 gs.cpp:75,130
    double diag_rhs;  // 1.0 + 4.0*r*nu                 gs.cpp:44
    double inv_diag;  // 1/diag            (fast arithmetic only)
    double hr;        // r*h/2, rnu = r*nu (fast arithmetic only)
    double rnu;
};
Stencil make_stencil(double dt, double nu, double dx);

struct Coef4 { double a, b, c, d; };   // west, east, north, south

template <int ARITH>
struct Arith;

// EXACT: the reference's rounding sequence, one IEEE operation per source operation, no
// contraction (the _rn intrinsics are never fused by nvcc).  u is bit-identical to gs.cpp.
template <>
struct Arith<MGB200_ARITH_EXACT> {
    static __device__ __forceinline__ double minus(double v, const Stencil& s)   // gs.cpp:15
    {
        return __dmul_rn(s.r, __dadd_rn(__dmul_rn(__dmul_rn(-v, s.h), 0.5), s.nu));
    }
    static __device__ __forceinline__ double plus(double v, const Stencil& s)    // gs.cpp:19
    {
        return __dmul_rn(s.r, __dadd_rn(__dmul_rn(__dmul_rn(v, s.h), 0.5), s.nu));
    }
    static __device__ __forceinline__ Coef4 coef(double v1, double v2, const Stencil& s)
    {
        return Coef4{minus(v2, s), plus(v2, s), minus(v1, s), plus(v1, s)};      // gs.cpp:40-43
    }
    // gs.cpp:130   (rhs - c*up - a*lf - d*dn - b*rt) / diag
    static __device__ __forceinline__ double gs(double rhs, double up, double lf, double dn, double rt,
                                                const Coef4& k, const Stencil& s)
    {
        double t = __dsub_rn(rhs, __dmul_rn(k.c, up));
        t = __dsub_rn(t, __dmul_rn(k.a, lf));
        t = __dsub_rn(t, __dmul_rn(k.d, dn));
        t = __dsub_rn(t, __dmul_rn(k.b, rt));
        return __ddiv_rn(t, s.diag);
    }
    // the same update split in two: head = (rhs - c*up) - a*lf ; tail = ((head - d*dn) - b*rt) / diag
    static __device__ __forceinline__ double gs_head(double rhs, double up, double lf, const Coef4& k)
    {
        return __dsub_rn(__dsub_rn(rhs, __dmul_rn(k.c, up)), __dmul_rn(k.a, lf));
    }
    static __device__ __forceinline__ double gs_tail(double t, double dn, double rt, double kd, double kb, const Stencil& s)
    {
        t = __dsub_rn(t, __dmul_rn(kd, dn));
        t = __dsub_rn(t, __dmul_rn(kb, rt));
        return __ddiv_rn(t, s.diag);
    }
    // gs.cpp:75    rhs - (diag*u + c*up + a*lf + d*dn + b*rt)
    static __device__ __forceinline__ double residual(double rhs, double u, double up, double lf, double dn,
                                                      double rt, const Coef4& k, const Stencil& s)
    {
        double t = __dmul_rn(s.diag, u);
        t = __dadd_rn(t, __dmul_rn(k.c, up));
        t = __dadd_rn(t, __dmul_rn(k.a, lf));
        t = __dadd_rn(t, __dmul_rn(k.d, dn));
        t = __dadd_rn(t, __dmul_rn(k.b, rt));
        return __dsub_rn(rhs, t);
    }
    // gs.cpp:44    diag_rhs*u - c*up - a*lf - d*dn - b*rt
    static __device__ __forceinline__ double rhs(double u, double up, double lf, double dn, double rt,
                                                 const Coef4& k, const Stencil& s)
    {
        double t = __dmul_rn(s.diag_rhs, u);
        t = __dsub_rn(t, __dmul_rn(k.c, up));
        t = __dsub_rn(t, __dmul_rn(k.a, lf));
        t = __dsub_rn(t, __dmul_rn(k.d, dn));
        t = __dsub_rn(t, __dmul_rn(k.b, rt));
        return t;
    }
};

// FAST: same formulas contracted to FMAs, division replaced by the reciprocal of the diagonal.
template <>
struct Arith<MGB200_ARITH_FAST> {
    static __device__ __forceinline__ Coef4 coef(double v1, double v2, const Stencil& s)
    {
        return Coef4{fma(-v2, s.hr, s.rnu), fma(v2, s.hr, s.rnu), fma(-v1, s.hr, s.rnu), fma(v1, s.hr, s.rnu)};
    }
    static __device__ __forceinline__ double gs(double rhs, double up, double lf, double dn, double rt,
                                                const Coef4& k, const Stencil& s)
    {
        double t = fma(-k.c, up, rhs);
        t = fma(-k.a, lf, t);
        t = fma(-k.d, dn, t);
        t = fma(-k.b, rt, t);
        return t * s.inv_diag;
    }
    static __device__ __forceinline__ double gs_head(double rhs, double up, double lf, const Coef4& k)
    {
        return fma(-k.a, lf, fma(-k.c, up, rhs));
    }
    static __device__ __forceinline__ double gs_tail(double t, double dn, double rt, double kd, double kb, const Stencil& s)
    {
        t = fma(-kd, dn, t);
        t = fma(-kb, rt, t);
        return t * s.inv_diag;
    }
    static __device__ __forceinline__ double residual(double rhs, double u, double up, double lf, double dn,
                                                      double rt, const Coef4& k, const Stencil& s)
    {
        double t = s.diag * u;
        t = fma(k.c, up, t);
        t = fma(k.a, lf, t);
        t = fma(k.d, dn, t);
        t = fma(k.b, rt, t);
        return rhs - t;
    }
    static __device__ __forceinline__ double rhs(double u, double up, double lf, double dn, double rt,
                                                 const Coef4& k, const Stencil& s)
    {
        double t = s.diag_rhs * u;
        t = fma(-k.c, up, t);
        t = fma(-k.a, lf, t);
        t = fma(-k.d, dn, t);
        t = fma(-k.b, rt, t);
        return t;
    }
};

// ---------------------------------------------------------------------------------------------
// deterministic block sum (warp shuffle tree, then one warp over the per-warp sums)
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// all threads of the block must call; result valid in thread 0.  `scratch` >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* scratch)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = (lane < (nthr + 31) / 32) ? scratch[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

}  // namespace mgb200
