// comm.cuh -- one-process-per-GPU communication for the row-slab sharded solver: a thin layer over
// NCCL (loaded at run time from the process: torch.distributed has already mapped libnccl.so.2, a
// single-GPU user never needs it).  Point-to-point rows go between slab neighbours, one double is
// all-reduced per convergence check.
#pragma once
#include "common.cuh"

namespace mgb200 {

struct Comm;

// rank 0 creates the 128-byte NCCL unique id; the host (bench.py / tests) broadcasts it
int comm_unique_id(unsigned char out[128]);
int comm_create(Comm** out, int rank, int nranks, const unsigned char id[128]);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_size(const Comm* c);

// a batch of sends/receives issued as one NCCL group on `s`
struct P2P { int peer; double* buf; size_t count; bool send; };
int comm_p2p(Comm* c, const P2P* ops, int nops, cudaStream_t s);
// in-place sum of `count` doubles on every rank
int comm_allreduce_sum(Comm* c, double* buf, size_t count, cudaStream_t s);

}  // namespace mgb200
