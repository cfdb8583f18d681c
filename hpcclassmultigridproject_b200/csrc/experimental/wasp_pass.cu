// wasp_pass.cu -- EXPERIMENTAL: CUDA side of the warp-autonomous streaming pass (wasp_body.cuh).
// NOT compiled into libmgb200.so: built on its own (libmgb200x.so) by the tests that exercise it
// (tests/test_wasp_kernel.py: host emulation of the same body on CPU, compile check for sm_100a;
// the GPU parity / timing test runs only when MGB200_TEST_EXPERIMENTAL=1).  It exists so that the
// next round can measure the scheme on a B200 without writing it first.
#include <algorithm>
#include <cstdlib>

#define WP_FN __device__ __forceinline__
#include "wasp_body.cuh"

namespace mgb200 {
// the error plumbing of common.cuh, local to this stand-alone library
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
int fail(int code, const std::string& m) { g_err = m; return code; }
long& launch_counter() { static long c = 0; return c; }
// same sequence as solver.cu's make_stencil (this library is linked on its own)
Stencil make_stencil(double dt, double nu, double dx)
{
    volatile double r = 0.5 * dt / (dx * dx);
    volatile double four_r = 4.0 * r;
    volatile double four_r_nu = four_r * nu;
    Stencil s;
    s.r = r; s.nu = nu; s.h = dx;
    s.diag = 1.0 - four_r_nu;
    s.diag_rhs = 1.0 + four_r_nu;
    s.inv_diag = 1.0 / s.diag;
    s.hr = r * dx * 0.5;
    s.rnu = r * nu;
    return s;
}

namespace wasp {

WP_FN int wp_lane() { return threadIdx.x & 31; }
WP_FN double wp_shfl_up(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
WP_FN double wp_shfl_down(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
WP_FN V2 wp_ld2(const double* p)
{
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    return V2{t.x, t.y};
}
WP_FN void wp_st2(double* p, V2 v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }
WP_FN double wp_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int ARITH>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA) k_wasp_pass(const Params p)
{
    const int tile = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (tile >= p.nstrips * p.nbands) return;            // whole warps leave together
    run_strip<ARITH>(p, tile);
}

// K = 3 passes: one image per flavour (prolongation role or not, epilogue kind)
template <int ARITH, int PRE, int POSTK>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA) k_wasp_pass3(const Params p)
{
    const int tile = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (tile >= p.nstrips * p.nbands) return;
    run_strip_flavour<ARITH, PRE, POSTK>(p, tile);
}

using WaspKernel = void (*)(const Params);
template <int ARITH>
static WaspKernel kernel_of(int K, bool pre, int post)
{
    if (K != KMAX) return k_wasp_pass<ARITH>;
    if (pre) return post == POST_NORM2 ? k_wasp_pass3<ARITH, 1, POST_NORM2> : post == POST_INJECT ? k_wasp_pass3<ARITH, 1, POST_INJECT>
                                                                                                  : k_wasp_pass3<ARITH, 1, POST_NONE>;
    return post == POST_NORM2 ? k_wasp_pass3<ARITH, 0, POST_NORM2> : post == POST_INJECT ? k_wasp_pass3<ARITH, 0, POST_INJECT>
                                                                                          : k_wasp_pass3<ARITH, 0, POST_NONE>;
}

}  // namespace wasp
}  // namespace mgb200

using namespace mgb200;

extern "C" {

const char* mgb200x_last_error(void) { return g_err.c_str(); }

// number of tiles (= partial sums of a POST_NORM2 pass) for rows_per_band (<= 0: default)
long mgb200x_wasp_tiles(long n, long rows_per_band)
{
    int ns, nb; long RB;
    wasp::plan(n, n + 1, rows_per_band > 0 ? rows_per_band : 512, ns, nb, RB);
    return (long)ns * nb;
}

// One pass over level n; all arrays are DEVICE arrays in the solver's split layout
// (pitch = 2*odd, odd = roundup(n/2+1,16)+32; coarse arrays likewise for n/2), whole level.
// u_in may be NULL (zero iterate), cu NULL (no prolongation); post: 0 none, 1 injection into crhs,
// 2 sum of squares into partials[0..tiles).  arith: MGB200_ARITH_*.
int mgb200x_wasp_pass(long n, const double* u_in, double* u_out, const double* rhs, const double* v1, const double* v2,
                      const double* cu, double* crhs, double* partials, int K, int post, int arith, double dt, double nu,
                      double dx, long rows_per_band, void* stream)
{
    if (n < 4 || (n & 1) || !u_out || !rhs || !v1 || !v2 || K < 0 || K > wasp::KMAX || u_out == u_in)
        return fail(MGB200_ERR_INVALID, "wasp_pass: bad argument");
    if ((post == POST_INJECT && !crhs) || (post == POST_NORM2 && !partials)) return fail(MGB200_ERR_INVALID, "wasp_pass: epilogue target missing");
    const Layout L = split_layout(n), Lc = split_layout(n / 2);
    wasp::Params p{};
    p.u_in = u_in; p.rhs = rhs; p.v1 = v1; p.v2 = v2; p.cu = cu; p.u_out = u_out; p.crhs = crhs; p.partials = partials;
    p.n = n; p.nhalf = n / 2; p.pitch = L.pitch; p.odd = L.odd; p.cpitch = Lc.pitch; p.codd = Lc.odd;
    p.K = K; p.post = post; p.pre = cu ? 1 : 0;
    p.own_lo = 0; p.own_hi = n; p.row0 = 0; p.crow0 = 0;          // whole level (slab windows: see wasp_body.cuh)
    wasp::plan(n, n + 1, rows_per_band > 0 ? rows_per_band : 512, p.nstrips, p.nbands, p.RB);
    p.st = make_stencil(dt, nu, dx);
    const int tiles = p.nstrips * p.nbands;
    const unsigned grid = (unsigned)((tiles + wasp::WARPS_PER_CTA - 1) / wasp::WARPS_PER_CTA);
    cudaStream_t s = (cudaStream_t)stream;
    if ((long)(n + 1) * L.pitch >= (1L << 31)) return fail(MGB200_ERR_INVALID, "wasp_pass: level too large for 32-bit row offsets");
    const wasp::WaspKernel kern = arith == MGB200_ARITH_EXACT ? wasp::kernel_of<MGB200_ARITH_EXACT>(K, p.pre != 0, post)
                                                              : wasp::kernel_of<MGB200_ARITH_FAST>(K, p.pre != 0, post);
    kern<<<grid, 32 * wasp::WARPS_PER_CTA, 0, s>>>(p);
    return check_launch("k_wasp_pass");
}

}  // extern "C"
