// sp_emu.cpp -- TEST INFRASTRUCTURE: a host emulation of the fused streaming kernel.
//
// There is no GPU in the build container, so the index logic of the streaming pass (ring slots,
// pipeline lags, halo validity, split-layout clipping) is exercised here by compiling the very
// same per-thread step code (csrc/stream_pass_body.cuh) for the host and running the threads of a
// block one after another inside each step.  Asynchronous primitives become immediate memcpy;
// shared memory starts as NaN so that any read of a cell that was never staged poisons the
// result.  Thread order inside a step is selectable (ascending / descending / shuffled): a result
// that depends on it reveals an intra-step race.  Never linked into libmgb200.so.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __shfl_down_sync(unsigned, double v, int) { return v; }
static inline void __syncthreads() {}
#define SP_FN static inline
#include <cuda_runtime.h>
static uint3 threadIdx, blockDim;
#include "../../hpcclassmultigridproject_b200/csrc/stream_pass_body.cuh"

namespace mgb200 {
void set_error(const std::string&) {}
int fail(int code, const std::string&) { return code; }
long& launch_counter() { static long c = 0; return c; }
namespace sp {
static int g_lane = 0;   // lane of the thread being emulated
SP_FN bool sp_elect() { return g_lane == 0; }
SP_FN void sp_bar_expect(const Smem&, int, unsigned) {}
// software model of the 3-D tensor copy: box {inner, 2, rows} at (x, 0, z), zeros outside the tensor
SP_FN void sp_tma_load(const Params& p, const Smem& sm, int which, unsigned soff, int x, int z, int)
{
    double* sdst = reinterpret_cast<double*>(sm.raw + soff);
    const double* src = which == FIELD_U ? p.u_in : which == FIELD_F ? p.rhs : which == FIELD_V1 ? p.v1 : which == FIELD_V2 ? p.v2 : p.cu;
    const bool coarse = which == FIELD_C;
    const long odd = coarse ? p.codd : p.odd, pitch = coarse ? p.cpitch : p.pitch;
    const long nrows = coarse ? p.crows_mem : p.rows_mem;
    const int inner = coarse ? p.CW : p.SWK, brows = coarse ? CROWS : GROUP;
    for (int zz = 0; zz < brows; ++zz)
        for (int par = 0; par < 2; ++par)
            for (int xx = 0; xx < inner; ++xx) {
                const long gx = (long)x + xx, gz = (long)z + zz;
                const bool in = gx >= 0 && gx < odd && gz >= 0 && gz < nrows;
                sdst[((long)zz * 2 + par) * inner + xx] = in ? src[gz * pitch + par * odd + gx] : 0.0;
            }
}
SP_FN void sp_tma_prefetch(const Params&, int, int, int) {}
SP_FN void sp_bar_wait(const Smem&, int, unsigned) {}
SP_FN void sp_bulk_store(const Smem& sm, double* gdst, unsigned soff, unsigned bytes) { std::memcpy(gdst, sm.raw + soff, bytes); }
SP_FN void sp_store_commit() {}
SP_FN void sp_store_wait_read2() {}
SP_FN void sp_fence_async() {}
SP_FN D2 sp_lds2(const Smem& sm, unsigned off) { return *reinterpret_cast<const D2*>(sm.raw + off); }
SP_FN double sp_lds1(const Smem& sm, unsigned off) { return *reinterpret_cast<const double*>(sm.raw + off); }
template <int DIR>
SP_FN double sp_side_neighbour(const Smem& sm, const D2&, unsigned off) { return sp_lds1(sm, off); }
SP_FN void sp_sts2(const Smem& sm, unsigned off, D2 v) { *reinterpret_cast<D2*>(sm.raw + off) = v; }
SP_FN void sp_sts1(const Smem& sm, unsigned off, double v) { *reinterpret_cast<double*>(sm.raw + off) = v; }
}  // namespace sp
}  // namespace mgb200

using namespace mgb200;
using namespace mgb200::sp;

extern "C" {

// Runs one pass over level n on HOST arrays in the split layout (rows_mem == 0: the whole level;
// otherwise a row slab: the arrays hold global rows row0.. and rows own_lo..own_hi are produced).  wk/nbands <= 0: use the planner.
// order: 0 ascending thread ids, 1 descending, 2 shuffled per step.  partials: >= ntiles doubles.
// Returns the number of tiles, or -1 on a bad argument.
long sp_emu_run(long n, long pitch, long odd, long cpitch, long codd, const double* u_in, double* u_out,
                const double* rhs, const double* v1, const double* v2, const double* cu, double* crhs,
                double* partials, int K, int post, int arith, double dt, double nu, double dx, int wk, int nbands,
                int order, long own_lo, long own_hi, long row0, long rows_mem, long crow0, long crows_mem)
{
    if (K < 0 || K > KMAX || n < 8 || (n & 3)) return -1;
    if (rows_mem == 0) { own_lo = 0; own_hi = n; row0 = 0; rows_mem = n + 1; crow0 = 0; crows_mem = n / 2 + 1; }
    Params p{};
    p.n = n; p.nhalf = n / 2; p.pitch = pitch; p.odd = odd; p.cpitch = cpitch; p.codd = codd;
    p.own_lo = own_lo; p.own_hi = own_hi; p.row0 = row0; p.rows_mem = rows_mem;
    p.mem_lo = row0; p.mem_hi = row0 + rows_mem - 1; p.crow0 = crow0; p.crows_mem = crows_mem;
    Plan pl = make_plan(n, own_hi - own_lo + 1, K, 148);
    if (wk > 0) {
        pl.WK = wk; pl.SWK = wk + 2 * HK; pl.nstrips = (int)((n / 2 + 1 + wk - 1) / wk);
    }
    if (nbands > 0) {
        const long nr = own_hi - own_lo + 1;
        pl.RBAND = (nr + nbands - 1) / nbands; pl.nbands = (int)((nr + pl.RBAND - 1) / pl.RBAND);
    }
    if (pl.SWK > SWK_MAX || (pl.SWK & 15)) return -1;
    p.RBAND = pl.RBAND; p.WK = pl.WK; p.SWK = pl.SWK; p.nstrips = pl.nstrips; p.nbands = pl.nbands;
    p.K = K; p.pre = cu ? 1 : 0; p.post = post; p.write_u = (K > 0 || cu) ? 1 : 0; p.u_is_zero = u_in ? 0 : 1;
    // Stencil exactly as make_stencil (solver.cu)
    {
        volatile double r = 0.5 * dt / (dx * dx);
        volatile double four_r = 4.0 * r;
        volatile double four_r_nu = four_r * nu;
        p.st.r = r; p.st.nu = nu; p.st.h = dx; p.st.diag = 1.0 - four_r_nu; p.st.diag_rhs = 1.0 + four_r_nu;
        p.st.inv_diag = 1.0 / p.st.diag; p.st.hr = r * dx * 0.5; p.st.rnu = r * nu;
    }
    p.CW = pl.SWK / 2 + 8;
    p.u_in = u_in; p.rhs = rhs; p.v1 = v1; p.v2 = v2; p.cu = cu;
    p.u_out = u_out; p.crhs = crhs; p.partials = partials;
    const long ntiles = (long)pl.nstrips * pl.nbands;
    std::vector<unsigned char> smem(SMEM_BYTES);
    std::vector<int> perm(THREADS);
    unsigned rng = 12345u;
    for (long tile = 0; tile < ntiles; ++tile) {
        double* sd = reinterpret_cast<double*>(smem.data());
        for (size_t q = 0; q < SMEM_BYTES / 8; ++q) sd[q] = std::numeric_limits<double>::quiet_NaN();
        Smem sm;
        carve(sm, smem.data(), p.SWK);
        const Tile tl = make_tile(p, tile);
        const Geo geo = make_geo(p);
        for (g_lane = 0; g_lane < 32; ++g_lane) producer_prologue(p, tl, geo, sm);
        std::vector<ThreadState> st(THREADS);
        for (int tid = 0; tid < THREADS; ++tid) st[tid] = init_thread(p, tl, geo, tid);
        wait_first_row(sm);
        const int t1 = last_step(p, tl);
        for (int t = first_step(tl); t <= t1; ++t) {
            std::iota(perm.begin(), perm.end(), 0);
            if (order == 1) std::reverse(perm.begin(), perm.end());
            if (order == 2)
                for (int a = THREADS - 1; a > 0; --a) {
                    rng = rng * 1664525u + 1013904223u;
                    std::swap(perm[a], perm[(rng >> 8) % (a + 1)]);
                }
            for (int q = 0; q < THREADS; ++q) {
                const int tid = perm[q];
                const int lane = tid & 31;
                g_lane = lane;
                if (arith == MGB200_ARITH_EXACT) role_step<MGB200_ARITH_EXACT>(p, tl, geo, sm, st[tid], t, lane);
                else role_step<MGB200_ARITH_FAST>(p, tl, geo, sm, st[tid], t, lane);
                end_step(tl, geo, sm, st[tid], t);
            }
        }
        if (post == POST_NORM2) {
            double s = 0.0;
            for (const ThreadState& a : st) s += a.acc;
            partials[tile] = s;
        }
    }
    return ntiles;
}

}  // extern "C"
