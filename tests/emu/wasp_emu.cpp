// wasp_emu.cpp -- TEST INFRASTRUCTURE: host emulation of the experimental warp-autonomous pass.
// The per-thread code of csrc/experimental/wasp_body.cuh is compiled for the host unchanged; the 32
// lanes of a warp run as 32 real threads and a shuffle is a rendezvous (write own slot, barrier,
// read the neighbour's, barrier).  A shuffle reached by only part of the warp therefore hangs the
// emulation, which is how divergence around shuffles would show (the test has a time limit).
// Never linked into libmgb200.so.
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __shfl_down_sync(unsigned, double v, int) { return v; }   // (common.cuh's block_sum: unused here)
static inline void __syncthreads() {}
#define WP_FN static inline
#include <cuda_runtime.h>
static uint3 threadIdx, blockDim;
#include "../../hpcclassmultigridproject_b200/csrc/experimental/wasp_body.cuh"

namespace mgb200 {
void set_error(const std::string&) {}
int fail(int code, const std::string&) { return code; }
long& launch_counter() { static long c = 0; return c; }
namespace wasp {

static thread_local int t_lane = 0;
static pthread_barrier_t g_bar;
static double g_slot[32];

WP_FN int wp_lane() { return t_lane; }
WP_FN double wp_shfl_up(double v)
{
    g_slot[t_lane] = v;
    pthread_barrier_wait(&g_bar);
    const double r = t_lane > 0 ? g_slot[t_lane - 1] : v;
    pthread_barrier_wait(&g_bar);
    return r;
}
WP_FN double wp_shfl_down(double v)
{
    g_slot[t_lane] = v;
    pthread_barrier_wait(&g_bar);
    const double r = t_lane < 31 ? g_slot[t_lane + 1] : v;
    pthread_barrier_wait(&g_bar);
    return r;
}
WP_FN V2 wp_ld2(const double* p) { return V2{p[0], p[1]}; }
WP_FN void wp_st2(double* p, V2 v) { p[0] = v.x; p[1] = v.y; }
WP_FN double wp_warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) {
        g_slot[t_lane] = v;
        pthread_barrier_wait(&g_bar);
        const double other = t_lane + o < 32 ? g_slot[t_lane + o] : v;
        pthread_barrier_wait(&g_bar);
        v += other;
    }
    return v;
}

}  // namespace wasp
}  // namespace mgb200

using namespace mgb200;

extern "C" {

// all arrays: HOST arrays in the split layout (pitch = 2*odd); returns the number of tiles
long wasp_emu_run(long n, long pitch, long odd, long cpitch, long codd, const double* u_in, double* u_out, const double* rhs,
                  const double* v1, const double* v2, const double* cu, double* crhs, double* partials, int K, int post,
                  int arith, double dt, double nu, double dx, long rows_per_band, long own_lo, long own_hi, long row0, long crow0)
{
    wasp::Params p{};
    p.u_in = u_in; p.rhs = rhs; p.v1 = v1; p.v2 = v2; p.cu = cu; p.u_out = u_out; p.crhs = crhs; p.partials = partials;
    p.n = n; p.nhalf = n / 2; p.pitch = pitch; p.odd = odd; p.cpitch = cpitch; p.codd = codd;
    p.K = K; p.post = post; p.pre = cu ? 1 : 0;
    // own_hi < 0: the whole level; otherwise a row slab (arrays hold global rows row0.. / crow0..)
    if (own_hi < 0) { own_lo = 0; own_hi = n; row0 = 0; crow0 = 0; }
    p.own_lo = own_lo; p.own_hi = own_hi; p.row0 = row0; p.crow0 = crow0;
    wasp::plan(n, own_hi - own_lo + 1, rows_per_band, p.nstrips, p.nbands, p.RB);
    // Stencil exactly as make_stencil (solver.cu)
    volatile double r = 0.5 * dt / (dx * dx);
    volatile double four_r = 4.0 * r;
    volatile double four_r_nu = four_r * nu;
    p.st.r = r; p.st.nu = nu; p.st.h = dx;
    p.st.diag = 1.0 - four_r_nu; p.st.diag_rhs = 1.0 + four_r_nu; p.st.inv_diag = 1.0 / p.st.diag;
    p.st.hr = r * dx * 0.5; p.st.rnu = r * nu;
    const int tiles = p.nstrips * p.nbands;
    pthread_barrier_init(&wasp::g_bar, nullptr, 32);
    for (int tile = 0; tile < tiles; ++tile) {
        std::vector<std::thread> lanes;
        for (int l = 0; l < 32; ++l)
            lanes.emplace_back([&, l] {
                wasp::t_lane = l;
                // K = 3: the compile-time flavours the CUDA dispatcher picks; otherwise the generic entry
                const bool fl = p.K == wasp::KMAX;
                if (arith == MGB200_ARITH_EXACT) {
                    if (!fl) wasp::run_strip<MGB200_ARITH_EXACT>(p, tile);
                    else if (p.pre && p.post == POST_NORM2) wasp::run_strip_flavour<MGB200_ARITH_EXACT, 1, POST_NORM2>(p, tile);
                    else if (p.pre && p.post == POST_INJECT) wasp::run_strip_flavour<MGB200_ARITH_EXACT, 1, POST_INJECT>(p, tile);
                    else if (p.pre) wasp::run_strip_flavour<MGB200_ARITH_EXACT, 1, POST_NONE>(p, tile);
                    else if (p.post == POST_NORM2) wasp::run_strip_flavour<MGB200_ARITH_EXACT, 0, POST_NORM2>(p, tile);
                    else if (p.post == POST_INJECT) wasp::run_strip_flavour<MGB200_ARITH_EXACT, 0, POST_INJECT>(p, tile);
                    else wasp::run_strip_flavour<MGB200_ARITH_EXACT, 0, POST_NONE>(p, tile);
                } else {
                    if (!fl) wasp::run_strip<MGB200_ARITH_FAST>(p, tile);
                    else if (p.pre && p.post == POST_NORM2) wasp::run_strip_flavour<MGB200_ARITH_FAST, 1, POST_NORM2>(p, tile);
                    else if (p.pre && p.post == POST_INJECT) wasp::run_strip_flavour<MGB200_ARITH_FAST, 1, POST_INJECT>(p, tile);
                    else if (p.pre) wasp::run_strip_flavour<MGB200_ARITH_FAST, 1, POST_NONE>(p, tile);
                    else if (p.post == POST_NORM2) wasp::run_strip_flavour<MGB200_ARITH_FAST, 0, POST_NORM2>(p, tile);
                    else if (p.post == POST_INJECT) wasp::run_strip_flavour<MGB200_ARITH_FAST, 0, POST_INJECT>(p, tile);
                    else wasp::run_strip_flavour<MGB200_ARITH_FAST, 0, POST_NONE>(p, tile);
                }
            });
        for (auto& t : lanes) t.join();
    }
    pthread_barrier_destroy(&wasp::g_bar);
    return tiles;
}

}  // extern "C"
