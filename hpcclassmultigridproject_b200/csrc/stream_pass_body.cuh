// stream_pass_body.cuh -- the fused streaming pass: geometry and per-thread step logic.
//
// One tile = a column strip x a row band of one level, handled by one thread block that STREAMS
// down the rows.  Row t lands in a shared-memory ring (bulk async copies, split layout: one run
// per column parity); behind the load front a software pipeline of half-sweeps follows, stage s
// (colour s&1) working on row t-2-2s.  With in-place updates a lag of TWO rows per half-sweep
// makes all stages of one step mutually independent (stage s on row i needs stage s-1 finished on
// row i+1 -- done one step earlier -- and stage s+1 not yet started on row i-1 -- one step
// later), so a step costs a single block barrier.  For K fused RB iterations (2K stages):
//
//   step t:  [prolong+correct row t] | stage s: row t-2-2s | store row t-4K-1 | residual row t-4K-2
//
// HBM traffic per node: u,rhs,v1,v2 read once, u written once (+ 1/4-size coarse array), whatever K.
// Redundant work: HK halo pairs per strip side, 2K+1 halo rows per band side (recomputed, never
// stored); tiles write to u_out != u_in, so no tile ever sees another tile's results.
//
// Warp roles (23 warps): 12 stage warps (two per half-sweep, 2 nodes per lane as one 16-byte
// vector), 2 prolongation warps, 8 residual-epilogue warps (one node per lane), 1 producer warp (one lane drives the
// TMA engine).  A warp's step is an almost serial dependency chain (address -> LDS -> 6 dependent
// FP64 operations -> STS), so the block is wide (many short chains) rather than deep.  Per-thread state that advances by one row per step (ring slots, rows) is kept
// incrementally, so a step has no division and almost no address arithmetic.
//
// Loads are 3-D TMA tensor copies: one instruction brings GROUP rows x both parity runs x SWK pairs
// of one field ({SWK, 2, GROUP} box of the {run, parity, row} view of the split layout) and
// zero-fills whatever lies outside the array; five instructions per GROUP rows feed a tile.
//
// This header holds no CUDA-specific instruction: everything asynchronous goes through the sp_*
// primitives declared below, defined with inline PTX in stream_pass.cu (and as plain memcpy in the
// host emulation used by tests/test_stream_pass_emu.py).
#pragma once
#include "common.cuh"
#include "stream_pass.cuh"

#ifndef SP_FN
#define SP_FN __device__ __forceinline__
#endif

namespace mgb200 {
namespace sp {

constexpr int SWK_MAX = 128;   // pairs (= 2 columns) per shared-memory row
constexpr int HK = 4;          // halo pairs per strip side (8 columns >= 2K+1 for K <= 3)
constexpr int GROUP = 4;       // rows per TMA box
constexpr int RING = 24;       // fine-row ring slots: rows t-4K-3 .. t in use + one group in flight
constexpr int NGROUP = RING / GROUP;
#ifndef MGB200_SP_PREFETCH
#define MGB200_SP_PREFETCH 1
#endif
constexpr int LEAD = 5;        // a group is requested LEAD steps before its first row is consumed
constexpr int CROWS = 3;       // coarse rows travelling with a group of fine rows
constexpr int CW_MAX = SWK_MAX / 2 + 8;   // doubles per coarse parity run in smem
constexpr int KMAX = 3;
constexpr int NSTW = 2;        // warps per half-sweep stage (64 pairs each)
constexpr int NSTAGE = 2 * KMAX * NSTW;
constexpr int NPRE = 2;        // prolongation warps (64 pairs each)
constexpr int NPOST = 8;       // residual-epilogue warps: 4 x 32 even-column nodes, 4 x 32 odd-column nodes
constexpr int WARPS = NSTAGE + NPRE + NPOST + 1;
constexpr int THREADS = WARPS * 32;
constexpr int PRODUCER_WARP = WARPS - 1;

// opaque storage for a CUtensorMap (128 bytes, 64-byte aligned); filled by the host launcher
struct alignas(64) TensorMapStorage { unsigned long long q[16]; };

struct Params {
    TensorMapStorage maps[5];  // Field order: u_in, rhs, v1, v2, coarse u
    long n, nhalf;             // level size, n/2
    long pitch, odd;           // split layout of this level
    long cpitch, codd;         // split layout of the next coarser level
    long RBAND;                // owned rows per band
    // row window (row-slab sharding; a single GPU owns and holds rows 0..n):
    long own_lo, own_hi;       // rows this launch must produce
    long mem_lo, mem_hi;       // rows present in the arrays (own rows + halo rows received from neighbours)
    long row0, crow0;          // global row index of memory row 0 of the fine / coarse arrays
    long rows_mem, crows_mem;  // rows held by the fine / coarse arrays
    int WK, SWK;               // owned pairs per strip, pairs per smem row (WK + 2 HK)
    int nstrips, nbands;
    int K;                     // fused RB iterations, 0..3
    int pre;                   // 1: u += P(coarse u) before smoothing
    int post;                  // StreamPost
    int write_u;               // 0 for a pure residual pass
    int u_is_zero;             // u_in == 0 everywhere: rows are zero-filled by an out-of-bounds box
    int CW;                    // SWK/2 + 8: coarse pairs per smem run
    Stencil st;
    const double* u_in;
    const double* rhs;
    const double* v1;
    const double* v2;
    const double* cu;          // coarse u (pre)
    double* u_out;
    double* crhs;              // coarse rhs (POST_INJECT)
    double* partials;          // POST_NORM2
};

struct Tile {
    int R0, R1;                // rows staged: [R0, R1]
    int rb0, rb1;              // rows owned:  [rb0, rb1]
    int k0;                    // global pair index of smem column 0 (may be -HK)
    int kb;                    // first owned pair
};

struct Smem {
    unsigned char* raw;        // start of the dynamic shared memory (generic address; host emulation)
    unsigned base32;           // its shared-space address (device only)
    unsigned bar_off;          // byte offset of the NGROUP load-completion barriers
};
// byte offsets from `raw`: U ring at 0, rhs / v1 / v2 rings at ringb, 2 ringb, 3 ringb, the coarse
// rows [NGROUP][CROWS][2][CW] at 4 ringb, then the barriers

enum Field { FIELD_U = 0, FIELD_F = 1, FIELD_V1 = 2, FIELD_V2 = 3, FIELD_C = 4 };

// dynamic shared memory of a tile with SWK pairs per row (all sub-arrays 128-byte aligned)
constexpr size_t smem_bytes(int swk)
{
    return (size_t)4 * RING * 2 * swk * 8 + (size_t)NGROUP * CROWS * 2 * (swk / 2 + 8) * 8 + NGROUP * 8 + 128;
}
constexpr size_t SMEM_BYTES = smem_bytes(SWK_MAX);

struct alignas(16) D2 { double x, y; };    // 16-byte vector; aligned accesses only (even pair index)


SP_FN void carve(Smem& sm, unsigned char* base, int swk)
{
    sm.raw = base; sm.base32 = 0;
    sm.bar_off = (unsigned)(4 * RING * 2 * swk * 8 + NGROUP * CROWS * 2 * (swk / 2 + 8) * 8);
}

SP_FN Tile make_tile(const Params& p, long tile)
{
    Tile tl;
    const long strip = tile % p.nstrips, band = tile / p.nstrips;
    const long HR = 2 * p.K + 1;
    long rb0 = p.own_lo + band * p.RBAND, rb1 = rb0 + p.RBAND - 1;
    if (rb1 > p.own_hi) rb1 = p.own_hi;
    long R0 = rb0 - HR, R1 = rb1 + HR;
    if (R0 < p.mem_lo) R0 = p.mem_lo;
    if (R1 > p.mem_hi) R1 = p.mem_hi;
    tl.kb = (int)(strip * p.WK);
    tl.k0 = tl.kb - HK;
    tl.rb0 = (int)rb0; tl.rb1 = (int)rb1; tl.R0 = (int)R0; tl.R1 = (int)R1;
    return tl;
}

SP_FN int ring_slot(const Tile& tl, int row) { return (row - tl.R0) % RING; }


// ------------------------------------------------------------------------------------------
// primitives (inline PTX in stream_pass.cu, plain C++ in the host emulation).  Shared memory is
// addressed by byte offsets from the start of the U ring, barriers by index.
SP_FN bool sp_elect();                                   // true in exactly one lane of the converged warp
SP_FN void sp_bar_expect(const Smem& sm, int bar, unsigned bytes);
// 3-D tensor load of field `which`: box {SWK (CW for FIELD_C), 2, GROUP (CROWS)} whose first element
// is (pair x, parity 0, memory row z) of the field; out-of-bounds elements arrive as zeros
SP_FN void sp_tma_load(const Params& p, const Smem& sm, int which, unsigned soff, int x, int z, int bar);
// same box, pulled into L2 only (no shared-memory destination)
SP_FN void sp_tma_prefetch(const Params& p, int which, int x, int z);
SP_FN void sp_bar_wait(const Smem& sm, int bar, unsigned parity);
SP_FN void sp_bulk_store(const Smem& sm, double* gdst, unsigned soff, unsigned bytes);
SP_FN void sp_store_commit();
SP_FN void sp_store_wait_read2();
SP_FN void sp_fence_async();
// 16-byte vectors need 16-byte aligned offsets; on the device these are ld/st.shared with
// register+immediate addresses
SP_FN D2 sp_lds2(const Smem& sm, unsigned off);
SP_FN double sp_lds1(const Smem& sm, unsigned off);
SP_FN void sp_sts2(const Smem& sm, unsigned off, D2 v);
SP_FN void sp_sts1(const Smem& sm, unsigned off, double v);
// The stage warps' third horizontal neighbour: the node of the other run just left (DIR = -1:
// mid.y of the previous lane) or just right (DIR = +1: mid.x of the next lane) of the lane's
// vector `mid`; the warp's edge lane reads it from shared memory at `off`.  (The host emulation
// always reads shared memory: the value is the same.)
template <int DIR>
SP_FN double sp_side_neighbour(const Smem& sm, const D2& mid, unsigned off);

// ------------------------------------------------------------------------------------------
// per-thread state, advanced by one row per step.  Rows are addressed by BYTE OFFSETS into the U
// ring that walk the ring incrementally; the rhs / v1 / v2 rings sit at +ringb, +2 ringb, +3 ringb.
enum Role { ROLE_STAGE = 0, ROLE_PRE = 1, ROLE_POST = 2, ROLE_PRODUCER = 3 };

struct Geo {
    unsigned swkb;    // bytes of one parity run: SWK * 8
    unsigned rowb;    // bytes of one ring row (both runs)
    unsigned ringb;   // bytes of one field's ring
};
SP_FN Geo make_geo(const Params& p)
{
    Geo g;
    g.swkb = (unsigned)p.SWK * 8u; g.rowb = 2u * g.swkb; g.ringb = (unsigned)RING * g.rowb;
    return g;
}

struct ThreadState {
    int role;
    int idx;            // stage number / chunk number
    int kk;             // first local pair handled (stage: also kk + 64)
    int row;            // the role's row at the current step
    int lo, hi;         // rows on which the role acts: [lo, hi]
    unsigned a_prev, a_cur, a_next;   // offsets of even-run element kk of rows row-1, row, row+1
    unsigned lim;       // ringb + kk*8: wrap limit of those offsets
    unsigned ok_cur;    // pre / post: validity bits of the lane's nodes
    unsigned okp[2];    // stage: validity bits of the lane's two targets for column parity 0 / 1
    int par;            // stage: column parity of the colour in `row`
    int wphase;         // (row awaited at the end of the step - R0) mod GROUP
    int wgroup;         // its group slot
    unsigned wpar;      // and that barrier's phase parity
    double acc;         // POST_NORM2 accumulator
    // stage: operands carried in registers from step to step (shared memory bandwidth is what
    // bounds this kernel): the other-parity nodes of the current row become the next row's upper
    // neighbours, the nodes loaded from the row below become the next row's horizontal neighbours
    D2 c_up, c_mid;
    // last stage of a POST_NORM2 pass: the residual of the nodes it has just updated costs no loads
    // (all operands are in registers), so it accumulates their squares itself and the epilogue warps
    // only visit the other colour.  nmask[par]: owned targets, nlo..nhi: owned interior rows
    unsigned nmask[2];
    int nlo, nhi;
    // producer: next row to store and next group to request, all advanced incrementally
    double* gst;        // global address of row `row`'s even run at the first owned pair
    unsigned nE8, nO8;  // bytes of the owned parts of the even / odd run
    int gnext;          // next group to request
    int gz;             // its first row (memory row index)
    int gslot;          // its group slot
};

// validity of a GS / residual target at local pair kk of parity par
SP_FN bool target_ok(const Params& p, const Tile& tl, int kk, int par)
{
    const long kg = (long)tl.k0 + kk;
    if (kk >= p.SWK) return false;
    if (par == 0) return kk >= 1 && kg >= 1 && kg <= p.nhalf - 1;          // even column 2kg: neighbours O[kk-1], O[kk]
    return kk <= p.SWK - 2 && kg >= 0 && kg <= p.nhalf - 1;                 // odd column 2kg+1: neighbours E[kk], E[kk+1]
}

SP_FN unsigned ring_adv(unsigned a, const Geo& g, unsigned lim)
{
    a += g.rowb;
    return a >= lim ? a - g.ringb : a;
}

SP_FN ThreadState init_thread(const Params& p, const Tile& tl, const Geo& geo, int tid)
{
    ThreadState s;
    const int warp = tid >> 5, lane = tid & 31;
    s.acc = 0.0; s.c_up = D2{0.0, 0.0}; s.c_mid = D2{0.0, 0.0}; s.nmask[0] = s.nmask[1] = 0; s.nlo = 1; s.nhi = 0; s.ok_cur = 0; s.okp[0] = s.okp[1] = 0; s.idx = 0; s.kk = 0; s.par = 0;
    s.lo = 1; s.hi = 0;                                  // empty range
    int off = 0;                                         // role row = t - off
    if (warp < NSTAGE) {
        s.role = ROLE_STAGE; s.idx = warp / NSTW; s.kk = 64 * (warp % NSTW) + 2 * lane; off = 2 + 2 * s.idx;
        if (s.idx < 2 * p.K && s.kk < p.SWK) { s.lo = tl.R0 + 1; s.hi = tl.R1 - 1; }   // rows row-1, row+1 must be staged
    } else if (warp < NSTAGE + NPRE) {
        s.role = ROLE_PRE; s.idx = warp - NSTAGE; s.kk = 64 * s.idx + 2 * lane; off = 0;
        for (int e = 0; e < 2; ++e) {
            const long kg = (long)tl.k0 + s.kk + e;
            if (s.kk + e < p.SWK && kg >= 0 && kg <= p.nhalf - 1) s.ok_cur |= 1u << e;          // pair holds an interior odd column
            if (s.kk + e < p.SWK && kg >= 1 && kg <= p.nhalf - 1) s.ok_cur |= 1u << (2 + e);    // ... and an interior even column
        }
        if (p.pre && (s.ok_cur & 3u)) { s.lo = tl.R0 < 1 ? 1 : tl.R0; s.hi = tl.R1 > p.n - 1 ? (int)p.n - 1 : tl.R1; }
    } else if (warp < NSTAGE + NPRE + NPOST) {
        // warps 0-3: the even-column node of pair kk, warps 4-7: the odd-column node (norm only)
        s.role = ROLE_POST; s.idx = warp - NSTAGE - NPRE; s.kk = HK + 32 * (s.idx & 3) + lane; off = 4 * p.K + 2;
        s.par = s.idx >> 2;
        if (s.kk < HK + p.WK && (s.par == 0 || p.post == POST_NORM2) && target_ok(p, tl, s.kk, s.par)) s.ok_cur = 1u;   // owned nodes only
        if (p.post != POST_NONE && s.ok_cur) { s.lo = tl.rb0 < 1 ? 1 : tl.rb0; s.hi = tl.rb1 > p.n - 1 ? (int)p.n - 1 : tl.rb1; }
    } else {
        // the producer follows the row that is stored in the current step: finished one step earlier
        s.role = ROLE_PRODUCER; s.kk = HK; off = 4 * p.K + 1;
        if (p.write_u) { s.lo = tl.rb0; s.hi = tl.rb1; }
    }
    s.gst = nullptr; s.nE8 = s.nO8 = 0; s.gnext = 2; s.gz = tl.R0 + 2 * GROUP - (int)p.row0; s.gslot = 2 % NGROUP;
    if (s.role == ROLE_PRODUCER) {
        long eE = (long)tl.kb + p.WK, eO = eE;
        if (eE > p.nhalf + 1) eE = p.nhalf + 1;
        if (eO > p.nhalf) eO = p.nhalf;
        s.nE8 = (unsigned)(((eE - tl.kb + 1) & ~1L) * 8);    // whole 16-byte units (the layout has slack)
        s.nO8 = (unsigned)((eO - tl.kb) * 8);
        s.gst = p.u_out + ((long)(tl.R0 - off) - p.row0) * p.pitch + tl.kb;
    }
    s.row = tl.R0 - off;
    const int slot = ((RING - off) % RING + RING) % RING;               // row R0 sits in slot 0
    s.lim = geo.ringb + (unsigned)s.kk * 8u;
    s.a_cur = (unsigned)slot * geo.rowb + (unsigned)s.kk * 8u;
    s.a_next = ring_adv(s.a_cur, geo, s.lim);
    s.a_prev = (slot == 0 ? (unsigned)(RING - 1) : (unsigned)(slot - 1)) * geo.rowb + (unsigned)s.kk * 8u;
    if (s.role == ROLE_STAGE) {
        s.par = (s.idx + s.row) & 1;
        for (int e = 0; e < 2; ++e)
            for (int par = 0; par < 2; ++par)
                if (target_ok(p, tl, s.kk + e, par)) {
                    s.okp[par] |= 1u << e;
                    if (s.kk + e >= HK && s.kk + e < HK + p.WK) s.nmask[par] |= 1u << e;      // owned pair
                }
        if (p.post == POST_NORM2 && p.K > 0 && s.idx == 2 * p.K - 1) {
            s.nlo = tl.rb0 < 1 ? 1 : tl.rb0; s.nhi = tl.rb1 > p.n - 1 ? (int)p.n - 1 : tl.rb1;
        }
    }
    s.wphase = 0; s.wgroup = 0; s.wpar = 0;
    return s;
}

// ------------------------------------------------------------------------------------------
// One half-sweep stage on row st.row (colour = stage & 1), one 16-byte vector per lane, PAR = the
// column parity of the colour in this row.  Branch-free: operands are always loaded (every offset
// stays inside the ring) and only the two stores are predicated, because a warp's step is a serial
// dependency chain in which every taken or not-taken branch costs more than the arithmetic.
// With pair index kk even, the horizontal neighbours of targets (kk, kk+1) are three consecutive
// nodes of the OTHER run: from kk-1 (even columns: O[kk-1], O[kk], O[kk+1]) or from kk (odd
// columns: E[kk], E[kk+1], E[kk+2]).
template <int ARITH, int PAR, bool NORM = false>
SP_FN void stage_compute(const Params& p, const Geo& geo, const Smem& sm, ThreadState& st, const D2 f, const D2 w1, const D2 w2)
{
    const unsigned po = PAR ? geo.swkb : 0u, oo = PAR ? 0u : geo.swkb;
    const unsigned c = st.a_cur + po;
    // up = run PAR of row-1 = what was `m` one step ago; m = run 1-PAR of row = what was `dn` one
    // step ago (neither is written by any stage in between: see the lag argument at the top)
    const D2 up = st.c_up, m = st.c_mid;
    const D2 dn = sp_lds2(sm, st.a_next + po);
    const double x = sp_side_neighbour<PAR ? 1 : -1>(sm, m, st.a_cur + oo + (PAR ? 16u : 0u) - (PAR ? 0u : 8u));
    const double n0 = PAR ? m.x : x, n1 = PAR ? m.y : m.x, n2 = PAR ? x : m.y;
    const Coef4 c0 = Arith<ARITH>::coef(w1.x, w2.x, p.st);
    const Coef4 c1 = Arith<ARITH>::coef(w1.y, w2.y, p.st);
    const double o0 = Arith<ARITH>::gs(f.x, up.x, n0, dn.x, n1, c0, p.st);
    const double o1 = Arith<ARITH>::gs(f.y, up.y, n1, dn.y, n2, c1, p.st);
    const bool act = st.row >= st.lo && st.row <= st.hi;
    // one 16-byte store in the (warp-uniform but for strip edges) common case: two 8-byte stores
    // at a 16-byte lane stride cost twice the shared-memory wavefronts
    const unsigned okb = act ? st.okp[PAR] : 0u;
    if (okb == 3u) sp_sts2(sm, c, D2{o0, o1});
    else {
        if (okb & 1u) sp_sts1(sm, c, o0);
        if (okb & 2u) sp_sts1(sm, c + 8u, o1);
    }
    if (NORM && st.row >= st.nlo && st.row <= st.nhi) {   // last stage of a norm pass: residual of the new values
        const double r0 = Arith<ARITH>::residual(f.x, o0, up.x, n0, dn.x, n1, c0, p.st);
        const double r1 = Arith<ARITH>::residual(f.y, o1, up.y, n1, dn.y, n2, c1, p.st);
        if (st.nmask[PAR] & 1u) st.acc += r0 * r0;
        if (st.nmask[PAR] & 2u) st.acc += r1 * r1;
    }
    st.c_up = m; st.c_mid = dn;
}

// rhs, v1, v2 of the nodes a step updates: read-only ring rows, independent of what the other
// stages store in the same step (so they may be fetched before the step barrier)
template <int PAR>
SP_FN void stage_fetch(const Geo& geo, const Smem& sm, const ThreadState& st, D2& f, D2& w1, D2& w2)
{
    const unsigned c = st.a_cur + (PAR ? geo.swkb : 0u);
    f = sp_lds2(sm, c + geo.ringb); w1 = sp_lds2(sm, c + 2u * geo.ringb); w2 = sp_lds2(sm, c + 3u * geo.ringb);
}

template <int ARITH, int PAR, bool NORM = false>
SP_FN void stage_step(const Params& p, const Geo& geo, const Smem& sm, ThreadState& st)
{
    D2 f, w1, w2;
    stage_fetch<PAR>(geo, sm, st, f, w1, w2);
    stage_compute<ARITH, PAR, NORM>(p, geo, sm, st, f, w1, w2);
}

// prolongation + correction of row t (gs.cpp:238-241 fused with multigrid.cpp:83), interior nodes.
// The coarse rows of a group start at coarse row (first fine row of the group) >> 1; coarse column
// (k0 + kk) sits at local index kk: even kk in the E run at kk/2, odd kk in the O run at kk/2.
SP_FN void pre_step(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, ThreadState& st)
{
    const int t = st.row;
    if (t < st.lo || t > st.hi) return;
    const int g = (t - tl.R0) / GROUP;
    const int crow = (t >> 1) - ((tl.R0 + GROUP * g) >> 1);            // 0..2; an odd fine row also uses crow+1
    const unsigned cwb = (unsigned)p.CW * 8u;
    const unsigned c0 = 4u * geo.ringb + (unsigned)((g % NGROUP) * CROWS + crow) * 2u * cwb + (unsigned)(st.kk >> 1) * 8u;
    const unsigned c1 = c0 + 2u * cwb;
    // coarse columns kk, kk+1, kk+2 of coarse row I (and I+1 for an odd fine row)
    const double a0 = sp_lds1(sm, c0), a1 = sp_lds1(sm, c0 + cwb), a2 = sp_lds1(sm, c0 + 8u);
    const D2 ue = sp_lds2(sm, st.a_cur), uo = sp_lds2(sm, st.a_cur + geo.swkb);
    double e0, e1, o0, o1;
    if ((t & 1) == 0) {
        e0 = a0; e1 = a1;                                                                   // gs.cpp:238
        o0 = __dmul_rn(__dadd_rn(a0, a1), 0.5); o1 = __dmul_rn(__dadd_rn(a1, a2), 0.5);     // gs.cpp:240
    } else {
        const double b0 = sp_lds1(sm, c1), b1 = sp_lds1(sm, c1 + cwb), b2 = sp_lds1(sm, c1 + 8u);
        e0 = __dmul_rn(__dadd_rn(a0, b0), 0.5); e1 = __dmul_rn(__dadd_rn(a1, b1), 0.5);     // gs.cpp:239
        o0 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a0, b0), a1), b1), 0.25);              // gs.cpp:241
        o1 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a1, b1), a2), b2), 0.25);
    }
    // multigrid.cpp:83.  16-byte stores in the common case (8-byte stores at a 16-byte lane stride
    // cost twice the shared-memory wavefronts)
    const double ne0 = __dadd_rn(ue.x, e0), ne1 = __dadd_rn(ue.y, e1), no0 = __dadd_rn(uo.x, o0), no1 = __dadd_rn(uo.y, o1);
    if ((st.ok_cur & 12u) == 12u) sp_sts2(sm, st.a_cur, D2{ne0, ne1});
    else {
        if (st.ok_cur & 4u) sp_sts1(sm, st.a_cur, ne0);
        if (st.ok_cur & 8u) sp_sts1(sm, st.a_cur + 8u, ne1);
    }
    if ((st.ok_cur & 3u) == 3u) sp_sts2(sm, st.a_cur + geo.swkb, D2{no0, no1});
    else {
        if (st.ok_cur & 1u) sp_sts1(sm, st.a_cur + geo.swkb, no0);
        if (st.ok_cur & 2u) sp_sts1(sm, st.a_cur + geo.swkb + 8u, no1);
    }
}

// residual epilogue on finished row q, one node per lane (PAR = its column parity): injection into
// the coarse rhs (even columns of even rows) or sum of squares
// POSTK: the pass's epilogue kind if known at compile time (the kernel is specialised per flavour), -1: p.post
template <int ARITH, int PAR, int POSTK = -1>
SP_FN void post_step(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, ThreadState& st)
{
    const int post = POSTK >= 0 ? POSTK : p.post;
    const int q = st.row;
    if (q < st.lo || q > st.hi) return;
    if (post == POST_INJECT && (q & 1)) return;
    // norm after smoothing: the last half-sweep's colour ((i+j) odd) is summed by the last stage itself
    if (post == POST_NORM2 && p.K > 0 && ((q + PAR) & 1)) return;
    const unsigned po = PAR ? geo.swkb : 0u;
    const unsigned c = st.a_cur + po;                     // the node itself
    const unsigned o = st.a_cur + (geo.swkb - po);        // the other run at the same pair index
    const double u = sp_lds1(sm, c), up = sp_lds1(sm, st.a_prev + po), dn = sp_lds1(sm, st.a_next + po);
    // even column: left = O[kk-1], right = O[kk]; odd column: left = E[kk], right = E[kk+1]
    const double lf = sp_lds1(sm, PAR ? o : o - 8u), rt = sp_lds1(sm, PAR ? o + 8u : o);
    const Coef4 k = Arith<ARITH>::coef(sp_lds1(sm, c + 2u * geo.ringb), sp_lds1(sm, c + 3u * geo.ringb), p.st);
    const double rv = Arith<ARITH>::residual(sp_lds1(sm, c + geo.ringb), u, up, lf, dn, rt, k, p.st);
    if (post == POST_INJECT) {
        const long kg = (long)tl.k0 + st.kk;
        p.crhs[((long)(q >> 1) - p.crow0) * p.cpitch + (kg & 1) * p.codd + (kg >> 1)] = rv;       // gs.cpp:283
    } else {
        st.acc += rv * rv;
    }
}

// ------------------------------------------------------------------------------------------
SP_FN int first_step(const Tile& tl) { return tl.R0; }
SP_FN int last_step(const Params& p, const Tile& tl) { return tl.rb1 + 4 * p.K + 2; }
SP_FN int num_groups(const Tile& tl) { return (tl.R1 - tl.R0) / GROUP + 1; }

// producer: request the group starting at memory row z (global row z + row0) into group slot gs:
// the four fields plus the three coarse rows its prolongation needs, all completing on barrier gs;
// then start the HBM fetch of the group after next, whose shared-memory request will hit L2.
// Executed by the whole (converged) producer warp with warp-uniform operands; one elected lane
// issues the instructions.
SP_FN void issue_group_loads(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, int z, int gs)
{
    if (!sp_elect()) return;
    const unsigned fine = (unsigned)GROUP * geo.rowb, coarse = (unsigned)(CROWS * 2 * p.CW * 8);
    sp_bar_expect(sm, gs, 4u * fine + (p.pre ? coarse : 0u));
    const unsigned so = (unsigned)gs * fine;
    // a zero iterate is produced by a box that lies entirely below the last memory row
    sp_tma_load(p, sm, FIELD_U, so, tl.k0, p.u_is_zero ? (int)p.rows_mem + 64 : z, gs);
    sp_tma_load(p, sm, FIELD_F, so + geo.ringb, tl.k0, z, gs);
    sp_tma_load(p, sm, FIELD_V1, so + 2u * geo.ringb, tl.k0, z, gs);
    sp_tma_load(p, sm, FIELD_V2, so + 3u * geo.ringb, tl.k0, z, gs);
    if (p.pre)
        sp_tma_load(p, sm, FIELD_C, 4u * geo.ringb + (unsigned)gs * coarse, tl.k0 / 2, ((z + (int)p.row0) >> 1) - (int)p.crow0, gs);
    const int zp = z + 2 * GROUP;
    if (MGB200_SP_PREFETCH && zp + (int)p.row0 <= tl.R1) {
        if (!p.u_is_zero) sp_tma_prefetch(p, FIELD_U, tl.k0, zp);
        sp_tma_prefetch(p, FIELD_F, tl.k0, zp);
        sp_tma_prefetch(p, FIELD_V1, tl.k0, zp);
        sp_tma_prefetch(p, FIELD_V2, tl.k0, zp);
    }
}

// producer prologue (whole warp): the first two groups
SP_FN void producer_prologue(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm)
{
    const int ng = num_groups(tl);
    for (int g = 0; g < 2 && g < ng; ++g) issue_group_loads(p, tl, geo, sm, tl.R0 + GROUP * g - (int)p.row0, g);
}

// producer step t (whole warp): bulk-store the owned part of the row finished by the previous
// step, and LEAD steps before a group's first row is consumed request it: its ring slots were last
// read (row t-4K-3 and older) in the previous step.
SP_FN void producer_step(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, ThreadState& st, int t)
{
    if (st.row >= st.lo && st.row <= st.hi && sp_elect()) {
        if (st.nE8) sp_bulk_store(sm, st.gst, st.a_cur, st.nE8);
        if (st.nO8) sp_bulk_store(sm, st.gst + p.odd, st.a_cur + geo.swkb, st.nO8);
        sp_store_commit();
    }
    st.gst += p.pitch;
    if (((t + LEAD - tl.R0) & (GROUP - 1)) == 0 && st.gnext < num_groups(tl)) {
        if (sp_elect()) sp_store_wait_read2();          // the slots' previous rows have left shared memory
        issue_group_loads(p, tl, geo, sm, st.gz, st.gslot);
        st.gnext += 1; st.gz += GROUP;
        st.gslot = st.gslot + 1 == NGROUP ? 0 : st.gslot + 1;
    }
}

// wait until the first group has landed (every thread, before the first step)
SP_FN void wait_first_row(const Smem& sm) { sp_bar_wait(sm, 0, 0u); }

// after a stage step: the next row has the other column parity
SP_FN void stage_flip(ThreadState& st) { st.par ^= 1; }

// the role's work of step t (on entry row t has landed: awaited at the end of the previous step)
template <int ARITH>
SP_FN void role_step(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, ThreadState& st, int t, int lane)
{
    if (st.role == ROLE_STAGE) {
        if (st.nhi >= st.nlo) {
            if (st.par) stage_step<ARITH, 1, true>(p, geo, sm, st);
            else stage_step<ARITH, 0, true>(p, geo, sm, st);
        } else {
            if (st.par) stage_step<ARITH, 1>(p, geo, sm, st);
            else stage_step<ARITH, 0>(p, geo, sm, st);
        }
        // the last stage's rows go to the bulk-store engine next step: generic -> async proxy
        if (st.idx == 2 * p.K - 1) sp_fence_async();
        stage_flip(st);
    } else if (st.role == ROLE_PRE) {
        pre_step(p, tl, geo, sm, st);
        if (p.K == 0) sp_fence_async();
    } else if (st.role == ROLE_POST) {
        if (st.par) post_step<ARITH, 1>(p, tl, geo, sm, st);
        else post_step<ARITH, 0>(p, tl, geo, sm, st);
    } else {
        producer_step(p, tl, geo, sm, st, t);
    }
}

// end of step t: advance the row pointers; row t+1 must have landed before anyone touches it in
// step t+1, so wait whenever it opens a new group.  (A block barrier follows.)
SP_FN void advance_row(const Geo& geo, ThreadState& st)
{
    st.row += 1;
    st.a_prev = st.a_cur; st.a_cur = st.a_next; st.a_next = ring_adv(st.a_next, geo, st.lim);
}

// called after the LAST step of a group (the step whose successor opens group `next_group`)
SP_FN void wait_group(const Tile& tl, const Smem& sm, ThreadState& st, int t)
{
    st.wgroup += 1;
    if (st.wgroup == NGROUP) { st.wgroup = 0; st.wpar ^= 1u; }
    if (t + 1 <= tl.R1) sp_bar_wait(sm, st.wgroup, st.wpar);
}

SP_FN void end_step(const Tile& tl, const Geo& geo, const Smem& sm, ThreadState& st, int t)
{
    advance_row(geo, st);
    st.wphase += 1;
    if (st.wphase == GROUP) {
        st.wphase = 0;
        st.wgroup += 1;
        if (st.wgroup == NGROUP) { st.wgroup = 0; st.wpar ^= 1u; }
        if (t + 1 <= tl.R1) sp_bar_wait(sm, st.wgroup, st.wpar);
    }
}

// ------------------------------------------------------------------------------------------
// Tile planner.  Cost model (relative): a tile takes (rows + fill) steps, a step costs
// c0 + SWK (barrier latency + work proportional to the strip width); tiles run one per SM in
// waves of `sms`.  Search strip width (SWK a multiple of 16: TMA boxes land 128-byte aligned) and
// band count for the cheapest plan.
struct Plan { int WK, SWK, nstrips, nbands; long RBAND; };

// resident tiles per SM: one (736 threads x up to 88 registers fill the register file; 217.5 KB of shared memory)
inline int tiles_per_sm(int) { return 1; }

// n: level size (columns); nrows: rows this launch produces (n+1 on a single GPU, the slab otherwise)
inline Plan make_plan(long n, long nrows, int K, int sms, int force_swk = 0)
{
    const long npairs = n / 2 + 1;
    Plan best{};
    double best_cost = 1e300;
    for (int SWK = 32; SWK <= SWK_MAX; SWK += 16) {
        if (force_swk && SWK != force_swk) continue;
        const int WK = SWK - 2 * HK;
        const int slots = sms * tiles_per_sm(SWK);
        const int nstrips = (int)((npairs + WK - 1) / WK);
        for (int nb = 1; nb <= 4096; nb = nb < 16 ? nb + 1 : nb * 2) {
            const long RB = (nrows + nb - 1) / nb;
            if (nb > 1 && RB < 8) break;
            const int nbands = (int)((nrows + RB - 1) / RB);
            const long tiles = (long)nstrips * nbands;
            const long waves = (tiles + slots - 1) / slots;
            const double steps = (double)RB + 2.0 * (2 * K + 1) + 4.0 * K + 2.0 + 2 * GROUP;
            // a step costs a fixed latency plus work proportional to the strip width; two
            // co-resident tiles hide each other's latency but share the SM's throughput
            const double per_step = tiles_per_sm(SWK) == 2 ? 48.0 + 2.0 * SWK : 96.0 + SWK;
            const double cost = (double)waves * steps * per_step;
            if (cost < best_cost) { best_cost = cost; best = Plan{WK, SWK, nstrips, nbands, RB}; }
        }
    }
    return best;
}

}  // namespace sp
}  // namespace mgb200
