// comm.cu -- NCCL through dlopen/dlsym (see comm.cuh).  Only the handful of entry points the
// sharded solver needs are bound.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "comm.cuh"

namespace mgb200 {

namespace {

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl g_nccl;
std::mutex g_mu;

int load_nccl()
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_nccl.lib) return MGB200_OK;
    // RTLD_NOLOAD first: reuse the copy torch has already mapped (same SONAME), else load the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(MGB200_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define BIND(field, sym)                                                                     \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym));                  \
    if (!g_nccl.field) return fail(MGB200_ERR_NCCL, std::string("libnccl lacks ") + sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(AllReduce, "ncclAllReduce");
    BIND(AllGather, "ncclAllGather");
    BIND(Broadcast, "ncclBroadcast");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    g_nccl.lib = h;
    return MGB200_OK;
}

#define MGB_NCCL(expr)                                                                                      \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) return fail(MGB200_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

}  // namespace

struct Comm {
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
};

int comm_unique_id(unsigned char out[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    MGB_TRY(load_nccl());
    ncclUniqueId id;
    MGB_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(out, &id, 128);
    return MGB200_OK;
}

int comm_create(Comm** out, int rank, int nranks, const unsigned char idbytes[128])
{
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MGB200_ERR_INVALID, "comm_create: bad rank / nranks");
    Comm* c = new Comm();
    c->rank = rank; c->nranks = nranks;
    if (nranks > 1) {
        int rc = load_nccl();
        if (rc != MGB200_OK) { delete c; return rc; }
        ncclUniqueId id;
        std::memcpy(&id, idbytes, 128);
        ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
        if (r != ncclSuccess) { delete c; return fail(MGB200_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    }
    *out = c;
    return MGB200_OK;
}

void comm_destroy(Comm* c)
{
    if (!c) return;
    if (c->nccl) g_nccl.CommDestroy(c->nccl);
    delete c;
}

int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_size(const Comm* c) { return c ? c->nranks : 1; }

int comm_p2p(Comm* c, const P2P* ops, int nops, cudaStream_t s)
{
    if (!c || c->nranks == 1 || nops == 0) return MGB200_OK;
    MGB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < nops; ++k) {
        const P2P& o = ops[k];
        if (o.count == 0) continue;
        if (o.send) MGB_NCCL(g_nccl.Send(o.buf, o.count, ncclFloat64, o.peer, c->nccl, s));
        else MGB_NCCL(g_nccl.Recv(o.buf, o.count, ncclFloat64, o.peer, c->nccl, s));
    }
    MGB_NCCL(g_nccl.GroupEnd());
    return MGB200_OK;
}

int comm_broadcast(Comm* c, double* buf, size_t count, int root, cudaStream_t s)
{
    if (!c || c->nranks == 1 || count == 0) return MGB200_OK;
    MGB_NCCL(g_nccl.Broadcast(buf, buf, count, ncclFloat64, root, c->nccl, s));
    return MGB200_OK;
}

int comm_allreduce_sum(Comm* c, double* buf, size_t count, cudaStream_t s)
{
    if (!c || c->nranks == 1) return MGB200_OK;
    MGB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->nccl, s));
    return MGB200_OK;
}

namespace {
constexpr int PUSH_CTAS = 32, PUSH_TPB = 256;

__global__ void __launch_bounds__(PUSH_TPB) k_peer_push(PeerPush a)
{
    const long tid = (long)blockIdx.x * PUSH_TPB + threadIdx.x, nth = (long)gridDim.x * PUSH_TPB;
    for (int k = 0; k < a.nseg; ++k) {
        const double2* __restrict__ src = reinterpret_cast<const double2*>(a.seg[k].src);
        double2* __restrict__ dst = reinterpret_cast<double2*>(a.seg[k].dst);
        const long cnt = a.seg[k].count >> 1;
        for (long i = tid; i < cnt; i += nth) dst[i] = src[i];
    }
    __threadfence_system();                       // this thread's peer stores before the arrival below
    __syncthreads();
    if (threadIdx.x != 0) return;
    unsigned* arrive = reinterpret_cast<unsigned*>(a.sync + 16);
    if (atomicInc(arrive, gridDim.x - 1) != gridDim.x - 1) return;   // wraps to 0 for the next launch
    __threadfence_system();                       // every CTA's stores before the counters
    for (int k = 0; k < a.nraise; ++k) atomicAdd_system(a.raise[k], 1);
    for (int k = 0; k < a.nwait; ++k) {
        const int slot = a.wait_slot[k];
        const int want = a.sync[8 + slot] + a.wait_count[k];
        int seen;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(a.sync + slot) : "memory");
            if (seen - want < 0 && clock64() - t0 > SYNC_SPIN_BUDGET) { a.sync[SYNC_ABORT] = 1; break; }   // the peer is gone
        } while (seen - want < 0);
        a.sync[8 + slot] = want;
    }
}
}  // namespace

namespace {
__device__ __forceinline__ void wait_arrivals(int* sync, int slot, int count)
{
    const int want = sync[8 + slot] + count;
    int seen;
    const long long t0 = clock64();
    do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(sync + slot) : "memory");
        if (seen - want < 0 && clock64() - t0 > SYNC_SPIN_BUDGET) { sync[SYNC_ABORT] = 1; break; }          // the peer is gone
    } while (seen - want < 0);
    sync[8 + slot] = want;
}

__global__ void k_norm_allreduce(NormReduce a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (a.rank != 0) {
        *reinterpret_cast<volatile double*>(a.peer_red[0] + a.rank) = a.value[0];
        __threadfence_system();
        atomicAdd_system(a.peer_sync[0] + SYNC_NORM_UP, 1);
        wait_arrivals(a.sync, SYNC_NORM_DOWN, 1);
        a.value[0] = *reinterpret_cast<volatile double*>(a.red);
    } else {
        wait_arrivals(a.sync, SYNC_NORM_UP, a.nranks - 1);
        double sum = a.value[0];
        for (int r = 1; r < a.nranks; ++r) sum = __dadd_rn(sum, *reinterpret_cast<volatile double*>(a.red + r));
        a.value[0] = sum;
        for (int r = 1; r < a.nranks; ++r) *reinterpret_cast<volatile double*>(a.peer_red[r]) = sum;
        __threadfence_system();
        for (int r = 1; r < a.nranks; ++r) atomicAdd_system(a.peer_sync[r] + SYNC_NORM_DOWN, 1);
    }
}
}  // namespace

int launch_norm_allreduce(const NormReduce& a, cudaStream_t s)
{
    k_norm_allreduce<<<1, 32, 0, s>>>(a);
    return check_launch("k_norm_allreduce");
}

int launch_peer_push(const PeerPush& a, cudaStream_t s)
{
    const bool copies = a.nseg > 0;
    k_peer_push<<<copies ? PUSH_CTAS : 1, copies ? PUSH_TPB : 32, 0, s>>>(a);
    return check_launch("k_peer_push");
}

int comm_allgather_bytes(Comm* c, const void* send, void* recv, size_t bytes, cudaStream_t s)
{
    if (!c || c->nranks == 1) {
        MGB_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return MGB200_OK;
    }
    MGB_NCCL(g_nccl.AllGather(send, recv, bytes, ncclChar, c->nccl, s));
    return MGB200_OK;
}

}  // namespace mgb200
