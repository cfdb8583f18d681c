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

int comm_allreduce_sum(Comm* c, double* buf, size_t count, cudaStream_t s)
{
    if (!c || c->nranks == 1) return MGB200_OK;
    MGB_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->nccl, s));
    return MGB200_OK;
}

}  // namespace mgb200
