// stream_pass.cu -- device primitives (inline PTX: mbarrier + cp.async.bulk, i.e. the TMA engine
// in its 1-D bulk form), the kernel, the tile planner and the launcher of the fused streaming pass.
// The per-thread logic lives in stream_pass_body.cuh.
#include <algorithm>
#include <map>
#include <mutex>

#include "stream_pass_body.cuh"

namespace mgb200 {
namespace sp {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

SP_FN void sp_bar_expect(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

SP_FN void sp_bulk_load(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

SP_FN void sp_bar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned a = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}

SP_FN void sp_bulk_store(double* gdst, const double* ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

SP_FN void sp_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SP_FN void sp_store_wait_read2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
SP_FN void sp_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int ARITH>
__global__ void __launch_bounds__(THREADS, 1) k_stream_pass(const __grid_constant__ Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double scratch[32];
    Smem sm;
    carve(sm, smem_raw);
    const Tile tl = make_tile(p, blockIdx.x);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < RING; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.full[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sp_fence_async();
    }
    __syncthreads();
    if (tid == PRODUCER_WARP * 32) producer_prologue(p, tl, sm);
    ThreadState st = init_thread(p, tl, tid);
    wait_first_row(sm);
    const int t1 = last_step(p, tl);
    for (int t = first_step(tl); t <= t1; ++t) {
        thread_step<ARITH>(p, tl, sm, st, t, lane);
        __syncthreads();
    }
    if (tid == PRODUCER_WARP * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.post == POST_NORM2) {
        const double tot = block_sum(st.acc, scratch);
        if (tid == 0) p.partials[blockIdx.x] = tot;
    }
}

static int g_sms = 0;
static double* g_zero_row = nullptr;

static const Plan& plan_for(long n, int K)
{
    static std::map<std::pair<long, int>, Plan> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(n, K);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, make_plan(n, K, g_sms > 0 ? g_sms : 148)).first;
    return it->second;
}

}  // namespace sp

using namespace sp;

int stream_pass_init()
{
    static bool done = false;
    if (done) return MGB200_OK;
    int dev = 0;
    MGB_CUDA(cudaGetDevice(&dev));
    MGB_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    MGB_CUDA(cudaFuncSetAttribute(k_stream_pass<MGB200_ARITH_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MGB_CUDA(cudaFuncSetAttribute(k_stream_pass<MGB200_ARITH_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MGB_CUDA(cudaMalloc(&g_zero_row, 4 * SWK_MAX * sizeof(double)));
    MGB_CUDA(cudaMemset(g_zero_row, 0, 4 * SWK_MAX * sizeof(double)));
    done = true;
    return MGB200_OK;
}

long stream_pass_tiles(long n, int iters)
{
    if (iters >= 0) {
        const Plan& pl = plan_for(n, iters > KMAX ? KMAX : iters);
        return (long)pl.nstrips * pl.nbands;
    }
    long m = 0;
    for (int K = 0; K <= KMAX; ++K) {
        const Plan& pl = plan_for(n, K);
        m = std::max(m, (long)pl.nstrips * pl.nbands);
    }
    return m;
}

int stream_pass(const StreamPassArgs& a, cudaStream_t s)
{
    if (a.iters < 0 || a.iters > KMAX) return fail(MGB200_ERR_INVALID, "stream_pass: iters out of range");
    if (a.n < 8 || (a.n & 3)) return fail(MGB200_ERR_INVALID, "stream_pass: n must be a multiple of 4, >= 8");
    if (!a.L.split() || a.u_out == a.u_in) return fail(MGB200_ERR_INVALID, "stream_pass: needs the split layout and u_out != u_in");
    MGB_TRY(stream_pass_init());
    const Plan& pl = plan_for(a.n, a.iters);
    Params p{};
    p.n = a.n; p.nhalf = a.n / 2;
    p.pitch = a.L.pitch; p.odd = a.L.odd;
    p.cpitch = a.Lc.pitch; p.codd = a.Lc.odd;
    p.RBAND = pl.RBAND; p.WK = pl.WK; p.SWK = pl.SWK; p.nstrips = pl.nstrips; p.nbands = pl.nbands;
    p.K = a.iters;
    p.pre = a.coarse_u ? 1 : 0;
    p.post = a.post;
    p.write_u = (a.iters > 0 || a.coarse_u) ? 1 : 0;
    p.u_is_zero = a.u_in ? 0 : 1;
    p.st = a.st;
    p.u_in = a.u_in; p.rhs = a.rhs; p.v1 = a.v1; p.v2 = a.v2; p.cu = a.coarse_u; p.zero_row = g_zero_row;
    p.u_out = a.u_out; p.crhs = a.coarse_rhs; p.partials = a.partials;
    if (p.post == POST_INJECT && !p.crhs) return fail(MGB200_ERR_INVALID, "stream_pass: POST_INJECT without coarse_rhs");
    if (p.post == POST_NORM2 && !p.partials) return fail(MGB200_ERR_INVALID, "stream_pass: POST_NORM2 without partials");
    const unsigned grid = (unsigned)(pl.nstrips * pl.nbands);
    if (a.arith == MGB200_ARITH_EXACT) k_stream_pass<MGB200_ARITH_EXACT><<<grid, THREADS, SMEM_BYTES, s>>>(p);
    else k_stream_pass<MGB200_ARITH_FAST><<<grid, THREADS, SMEM_BYTES, s>>>(p);
    return check_launch("k_stream_pass");
}

}  // namespace mgb200
