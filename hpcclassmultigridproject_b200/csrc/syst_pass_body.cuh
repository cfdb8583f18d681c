// syst_pass_body.cuh -- the fused streaming pass as a SYSTOLIC pipeline of warps: geometry and the
// per-thread step logic (no CUDA-specific instruction in this file: everything asynchronous goes
// through the sy_* primitives, inline PTX in syst_pass.cu, a thread-per-lane host model in
// tests/emu/syst_emu.cpp).
//
// One pass = [u += P(coarse u)] -> K red-black Gauss-Seidel iterations -> [residual -> injection |
// residual -> sum of squares] on one level (reference: multigrid.cpp:69-88 around gs.cpp:109-189,
// :55-83, :268-292, :228-266), u / rhs / v1 / v2 read once from HBM and u written once.
//
// One tile = a column strip x a row band, one thread block, streaming DOWN the rows.  TMA boxes
// bring GROUP rows of the four fields (plus the coarse rows their prolongation needs) into a
// shared-memory ring; 2K half-sweep stages follow the load front, stage s (colour s&1) two rows
// behind stage s-1, updating the ring in place; finished rows leave through bulk stores.
//
// What is different from a barrier-per-row pipeline: the warps are coupled ONLY by the data
// dependencies of the half-sweeps.  Every stage warp publishes "step k done" on its own ring of
// mbarriers and a warp waits for exactly the warps whose results it reads:
//
//     stage s, step k   needs   stage s-1 (both column halves), step k-1          [row below + strip-half edge]
//     stage 0           needs   the TMA group of its lower row                     [full barriers]
//     producer, step k  needs   the last stage, step k                             [row to store, ring slots to refill]
//
// The reverse (write-after-read) hazards are implied by these: stage s+1 cannot touch a row before
// stage s has finished the row below it, and by then stage s has long read everything it needs of
// it.  A waiting warp sleeps in mbarrier.try_wait; there is no block barrier in the row loop, so a
// slow step of one warp is absorbed by the slack of the ring instead of stalling all the others.
//
// The prologue and the epilogue of the pass are folded into the first and the last stage:
//   * prolongation + correction: a Gauss-Seidel update never reads its own old value, so only the
//     nodes of the SECOND colour need the correction before the first half-sweep, and stage 0 reads
//     each of them exactly once (as its lower neighbours): it adds P(coarse u) on the fly and stores
//     the corrected values back for the strip-half edges;
//   * residual: after its update of row i the last stage holds the final values of every
//     neighbour of the first-colour nodes of row i-1 in registers (its targets of the last three
//     steps), so their residual costs three vector loads (rhs, v1, v2); the residual of its own
//     targets costs none.  Injection needs the even columns of even rows only: every other step.
// Shared-memory traffic per node and pass: about 160 B (down leg) / 180 B (up leg).
//
// Warps: 2 per stage (64 column pairs each, one 16-byte vector = 2 nodes per lane) + 1 producer.
#pragma once
#include "common.cuh"
#include "stream_pass.cuh"

#ifndef SY_FN
#define SY_FN __device__ __forceinline__
#endif

namespace mgb200 {
namespace sy {

constexpr int SWK_MAX = 128;   // pairs (= 2 columns) per shared-memory row
constexpr int HK = 4;          // halo pairs per strip side (8 columns >= 2K+1 for K <= 3)
constexpr int GROUP = 4;       // rows per TMA box
constexpr int RING = 24;       // fine-row ring slots
constexpr int NGROUP = RING / GROUP;
constexpr int CROWS = 3;       // coarse rows travelling with a group of fine rows
constexpr int KMAX = 3;
constexpr int NSTW = 2;        // warps per half-sweep stage (64 pairs each)
constexpr int NSTAGE = 2 * KMAX * NSTW;
constexpr int WARPS = NSTAGE + 1;
constexpr int THREADS = WARPS * 32;
constexpr int PRODUCER_WARP = NSTAGE;
// progress barriers per stage warp: step k is published on slot k % PBSLOTS.  How far a warp can run
// ahead of one that waits for it is bounded by the ring (a stage cannot pass the load front, and the load
// front cannot pass the last stage by more than RING rows) plus the steps after the last staged row
// (at most 4 KMAX, none of which waits for data): with PBSLOTS beyond that a slot cannot complete
// twice before its waiter has looked at it.
constexpr int PBSLOTS = 64;
static_assert(PBSLOTS > RING + 4 * KMAX && (PBSLOTS & (PBSLOTS - 1)) == 0, "progress barrier ring too short");
// the last stage has read ring row r-1 for the last time once it has completed its step on row r
// (that step still needs rhs / v1 / v2 of row r-1 for the residual)
constexpr int REFILL_ROW = RING - GROUP;   // group g may be requested once the last stage is done with row (first row of g) - REFILL_ROW

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
    int K;                     // fused RB iterations, 1..3
    int pre;                   // 1: u += P(coarse u) before smoothing
    int post;                  // StreamPost
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
    unsigned full_off;         // byte offset of the NGROUP load-completion barriers
    unsigned pb_off;           // byte offset of the NSTAGE x PBSLOTS progress barriers
};
// byte offsets from `raw`: U ring at 0, rhs / v1 / v2 rings at ringb, 2 ringb, 3 ringb, the coarse
// rows [NGROUP][CROWS][2][CW] at 4 ringb, then the barriers

enum Field { FIELD_U = 0, FIELD_F = 1, FIELD_V1 = 2, FIELD_V2 = 3, FIELD_C = 4 };

constexpr size_t smem_bytes(int swk)
{
    return (size_t)4 * RING * 2 * swk * 8 + (size_t)NGROUP * CROWS * 2 * (swk / 2 + 8) * 8 + NGROUP * 8 + NSTAGE * PBSLOTS * 8 + 128;
}
constexpr size_t SMEM_BYTES = smem_bytes(SWK_MAX);

struct alignas(16) D2 { double x, y; };    // 16-byte vector; aligned accesses only (even pair index)

SY_FN void carve(Smem& sm, unsigned char* base, int swk)
{
    sm.raw = base; sm.base32 = 0;
    sm.full_off = (unsigned)(4 * RING * 2 * swk * 8 + NGROUP * CROWS * 2 * (swk / 2 + 8) * 8);
    sm.pb_off = sm.full_off + NGROUP * 8;
}

SY_FN Tile make_tile(const Params& p, long tile)
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

// ------------------------------------------------------------------------------------------
// primitives (inline PTX in syst_pass.cu, plain C++ / atomics in the host emulation).  Shared
// memory is addressed by byte offsets from the start of the U ring.
SY_FN int  sy_lane();
SY_FN bool sy_elect();                                   // true in exactly one lane of the converged warp
SY_FN void sy_syncwarp();                                // execution + memory barrier among the warp's lanes
SY_FN void sy_full_expect(const Smem& sm, int g, unsigned bytes);
// 3-D tensor load of field `which`: box {SWK (CW for FIELD_C), 2, GROUP (CROWS)} whose first element
// is (pair x, parity 0, memory row z) of the field; out-of-bounds elements arrive as zeros
SY_FN void sy_tma_load(const Params& p, const Smem& sm, int which, unsigned soff, int x, int z, int g);
SY_FN void sy_tma_prefetch(const Params& p, int which, int x, int z);   // same box, into L2 only
SY_FN void sy_full_wait(const Smem& sm, int g, unsigned parity);
// progress barriers: one arrival (an elected lane, after sy_syncwarp) completes a phase; the wait
// has acquire semantics for the whole warp's earlier shared-memory writes
SY_FN void sy_pb_arrive(const Smem& sm, int warp, int slot);
SY_FN void sy_pb_wait(const Smem& sm, int warp, int slot, unsigned parity);
SY_FN void sy_bulk_store(const Smem& sm, double* gdst, unsigned soff, unsigned bytes);
SY_FN void sy_store_commit();
SY_FN void sy_store_wait_read0();                        // every committed store has read its shared memory
SY_FN void sy_store_wait_read1();                        // ... all but the most recent one
SY_FN void sy_store_wait_all();
SY_FN void sy_fence_async();                             // generic -> async proxy (rows handed to the bulk-store engine)
SY_FN D2 sy_lds2(const Smem& sm, unsigned off);
SY_FN double sy_lds1(const Smem& sm, unsigned off);
SY_FN void sy_sts2(const Smem& sm, unsigned off, D2 v);
SY_FN void sy_sts1(const Smem& sm, unsigned off, double v);
// The third horizontal neighbour of a lane's vector `mid` (run R of some row): the node of run R
// just left of it (dir < 0: mid.y of the previous lane) or just right (dir > 0: mid.x of the next
// lane).  Lanes 0 / 31 read it from shared memory at `off` (it belongs to the other warp of the
// stage); `outer`: the lane sits on the strip's outer edge, where the value is never used: 0.
// (The host emulation reads shared memory in every lane: the value is the same.)
SY_FN double sy_side(const Smem& sm, const D2& mid, unsigned off, int dir, bool outer);

// ------------------------------------------------------------------------------------------
struct Geo {
    unsigned swkb;    // bytes of one parity run: SWK * 8
    unsigned rowb;    // bytes of one ring row (both runs)
    unsigned ringb;   // bytes of one field's ring
    unsigned cwb;     // bytes of one coarse parity run
    unsigned cgrpb;   // bytes of the coarse rows of one group
    int nh;           // warps per stage that hold live pairs (1 for strips of <= 64 pairs)
};
SY_FN Geo make_geo(const Params& p)
{
    Geo g;
    g.swkb = (unsigned)p.SWK * 8u; g.rowb = 2u * g.swkb; g.ringb = (unsigned)RING * g.rowb;
    g.cwb = (unsigned)p.CW * 8u; g.cgrpb = (unsigned)CROWS * 2u * g.cwb;
    g.nh = p.SWK > 64 ? 2 : 1;
    return g;
}

SY_FN unsigned ring_adv(unsigned a, const Geo& g, unsigned lim)
{
    a += g.rowb;
    return a >= lim ? a - g.ringb : a;
}

SY_FN int first_step(const Tile& tl) { return tl.R0; }
// the last stage works on row t - 4K + 1 in step t and must still visit row rb1 + 1 (residual of row rb1)
SY_FN int last_step(const Params& p, const Tile& tl) { return tl.rb1 + 4 * p.K; }
SY_FN int num_groups(const Tile& tl) { return (tl.R1 - tl.R0) / GROUP + 1; }

// per-thread state of a stage warp, advanced by one row per step
struct Stage {
    int s, h;           // stage, column half
    int kk;             // first local pair of the lane's vector (even)
    int row;            // the stage's row at the current step: t - 1 - 2s
    int lo, hi;         // rows the stage updates: [lo, hi]
    unsigned a_prev, a_cur, a_next;   // ring offsets of even-run element kk of rows row-1, row, row+1
    unsigned lim;       // ringb + kk*8: wrap limit of those offsets
    unsigned okp0, okp1;   // update masks of the lane's two targets for column parity 0 / 1 (scalars: a dynamically
                           // indexed array would live in local memory)
    bool outer_l, outer_r;            // the lane's left / right side neighbour lies outside the strip
    bool live;          // the lane's vector lies inside the strip (strips narrower than the stage's 64 x NSTW pairs)
    // operands carried in registers from step to step: the other-parity nodes of the current row become
    // the next row's upper neighbours, the nodes loaded from the row below become the next row's
    // horizontal neighbours (no other stage writes them in between: stage s+1 is two rows behind)
    D2 c_up, c_mid;
    // last stage, residual epilogue: its (effective) targets of the previous two steps
    D2 o1, o2;
    unsigned rmask0, rmask1;   // owned nodes with an interior column, per column parity
    int elo, ehi;       // owned interior rows
    double acc;         // POST_NORM2 accumulator
    // stage 0 of a pass with prolongation
    unsigned cmask0, cmask1;   // interior-column bits per column parity
};

SY_FN Stage init_stage(const Params& p, const Tile& tl, const Geo& geo, int warp, int lane)
{
    Stage st;
    st.s = warp / NSTW; st.h = warp % NSTW;
    int kk = 64 * st.h + 2 * lane;
    if (kk > p.SWK - 2) kk = p.SWK - 2;               // lanes beyond the strip shadow its last vector (all their masks are 0)
    const bool live = 64 * st.h + 2 * lane < p.SWK;
    st.kk = kk;
    st.row = tl.R0 - 1 - 2 * st.s;
    st.lo = tl.R0 + 1; st.hi = tl.R1 - 1;             // rows row-1 and row+1 must be staged
    st.c_up = D2{0.0, 0.0}; st.c_mid = D2{0.0, 0.0}; st.o1 = D2{0.0, 0.0}; st.o2 = D2{0.0, 0.0};
    st.acc = 0.0;
    st.elo = tl.rb0 < 1 ? 1 : tl.rb0; st.ehi = tl.rb1 > p.n - 1 ? (int)p.n - 1 : tl.rb1;
    st.okp0 = st.okp1 = st.rmask0 = st.rmask1 = st.cmask0 = st.cmask1 = 0;
    for (int e = 0; e < 2 && live; ++e) {
        const int k = kk + e;
        const long kg = (long)tl.k0 + k;
        const bool owned = k >= HK && k < HK + p.WK;
        // even column 2kg: interior for 1 <= kg <= nhalf-1, neighbours O[k-1], O[k]
        const bool int0 = kg >= 1 && kg <= p.nhalf - 1;
        // odd column 2kg+1: interior for 0 <= kg <= nhalf-1, neighbours E[k], E[k+1]
        const bool int1 = kg >= 0 && kg <= p.nhalf - 1;
        if (int0 && k >= 1) st.okp0 |= 1u << e;
        if (int1 && k <= p.SWK - 2) st.okp1 |= 1u << e;
        if (int0 && owned) st.rmask0 |= 1u << e;
        if (int1 && owned) st.rmask1 |= 1u << e;
        if (int0) st.cmask0 |= 1u << e;
        if (int1) st.cmask1 |= 1u << e;
    }
    st.live = live;
    st.outer_l = kk == 0;
    st.outer_r = kk + 2 >= p.SWK;
    const int off = 1 + 2 * st.s;                                    // row = t - off
    const int slot = ((RING - off) % RING + RING) % RING;            // row R0 sits in slot 0
    st.lim = geo.ringb + (unsigned)kk * 8u;
    st.a_cur = (unsigned)slot * geo.rowb + (unsigned)kk * 8u;
    st.a_next = ring_adv(st.a_cur, geo, st.lim);
    st.a_prev = (slot == 0 ? (unsigned)(RING - 1) : (unsigned)(slot - 1)) * geo.rowb + (unsigned)kk * 8u;
    return st;
}

SY_FN void advance_row(const Geo& geo, Stage& st)
{
    st.row += 1;
    st.a_prev = st.a_cur; st.a_cur = st.a_next; st.a_next = ring_adv(st.a_next, geo, st.lim);
}

SY_FN void pb_wait_step(const Smem& sm, int warp, int k)
{
    sy_pb_wait(sm, warp, k & (PBSLOTS - 1), (unsigned)(k / PBSLOTS) & 1u);
}

// bilinear prolongation (gs.cpp:238-240) of the coarse iterate at the lane's two nodes of column
// parity PAR in fine row `t` (whose second-colour nodes have exactly that parity: even columns of odd
// rows, odd columns of even rows).  g: the TMA group of row t, G0: its first row.
// Coarse column (k0 + k) sits at local index k: even k in the E run at k/2, odd k in the O run at k/2.
SY_FN D2 prolong_pair(const Params& p, const Geo& geo, const Smem& sm, const Stage& st, int PAR, int t, int gslot, int G0)
{
    const int crow = (t >> 1) - (G0 >> 1);                           // 0..2; an odd fine row also uses crow+1
    const unsigned c0 = 4u * geo.ringb + (unsigned)gslot * geo.cgrpb + (unsigned)crow * 2u * geo.cwb + (unsigned)(st.kk >> 1) * 8u;
    const double a0 = sy_lds1(sm, c0), a1 = sy_lds1(sm, c0 + geo.cwb);
    D2 e;
    if (PAR) {                                                       // even row, odd columns (gs.cpp:240)
        const double a2 = sy_lds1(sm, c0 + 8u);
        e.x = __dmul_rn(__dadd_rn(a0, a1), 0.5); e.y = __dmul_rn(__dadd_rn(a1, a2), 0.5);
    } else {                                                         // odd row, even columns (gs.cpp:239)
        const unsigned c1 = c0 + 2u * geo.cwb;
        const double b0 = sy_lds1(sm, c1), b1 = sy_lds1(sm, c1 + geo.cwb);
        e.x = __dmul_rn(__dadd_rn(a0, b0), 0.5); e.y = __dmul_rn(__dadd_rn(a1, b1), 0.5);
    }
    return e;
}

// ------------------------------------------------------------------------------------------
// One step of a stage warp: the half-sweep on row st.row (colour = stage & 1), one 16-byte vector
// per lane, PAR = the column parity of that colour in this row.  With pair index kk even, the
// horizontal neighbours of targets (kk, kk+1) are three consecutive nodes of the OTHER run: from
// kk-1 (even columns: O[kk-1], O[kk], O[kk+1]) or from kk (odd columns: E[kk], E[kk+1], E[kk+2]).
//   FIRST   stage 0: waits for the TMA group of the row below instead of an upstream stage
//   PRE     (with FIRST) adds the prolongated coarse iterate to the nodes it loads from the row below
//   LAST    the warp's rows go to the bulk-store engine; EPI: residual epilogue kind
//   CHECKED the step may touch rows above R0 (first steps of a tile): loads of unstaged rows are skipped
template <int ARITH, bool FIRST, bool PRE, bool LAST, int EPI>
SY_FN void stage_step(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, Stage& st, const int k, const int PAR,
                      const bool CHECKED)
{
    const unsigned po = PAR ? geo.swkb : 0u, oo = PAR ? 0u : geo.swkb;
    const unsigned c = st.a_cur + po;
    const int row = st.row;
    // is the row itself / the row below staged?  (first and last steps of a tile only; a lane outside the strip
    // shadows the last live vector and must not load what that lane is storing)
    const bool own_ok = !CHECKED || (row >= tl.R0 && row <= tl.R1);
    const bool dn_ok = !CHECKED || (row + 1 >= tl.R0 && row + 1 <= tl.R1);
    const int w = st.s * NSTW + st.h;
    // ---- before the wait: everything that does not depend on the upstream stage's last step
    D2 f{0.0, 0.0}, w1{0.0, 0.0}, w2{0.0, 0.0};
    double x = 0.0;
    const D2 up = st.c_up, m = st.c_mid;
    if (FIRST && PRE && geo.nh == 2 && k >= 1) pb_wait_step(sm, w ^ 1, k - 1);   // the other half's corrected edge node
    if (own_ok) {
        if (st.live) { f = sy_lds2(sm, c + geo.ringb); w1 = sy_lds2(sm, c + 2u * geo.ringb); w2 = sy_lds2(sm, c + 3u * geo.ringb); }
        x = sy_side(sm, m, st.a_cur + oo + (PAR ? 16u : 0u) - (PAR ? 0u : 8u), PAR ? 1 : -1, PAR ? st.outer_r : st.outer_l);
    }
    const double n0 = PAR ? m.x : x, n1 = PAR ? m.y : m.x, n2 = PAR ? x : m.y;
    const Coef4 c0 = Arith<ARITH>::coef(w1.x, w2.x, p.st);
    const Coef4 c1 = Arith<ARITH>::coef(w1.y, w2.y, p.st);
    const double h0 = Arith<ARITH>::gs_head(f.x, up.x, n0, c0);
    const double h1 = Arith<ARITH>::gs_head(f.y, up.y, n1, c1);
    // ---- wait for the data of the row below (and, last stage: for the other half's previous targets)
    const int t = tl.R0 + k;                             // = row + 1 + 2s: the row below stage 0's
    if (FIRST) {
        if ((k & (GROUP - 1)) == 0 && t <= tl.R1) {
            const int g = k / GROUP;
            sy_full_wait(sm, g % NGROUP, (unsigned)(g / NGROUP) & 1u);
        }
    } else if (k >= 1) {
        pb_wait_step(sm, w - NSTW, k - 1);
        if (geo.nh == 2) pb_wait_step(sm, (w - NSTW) ^ 1, k - 1);
    }
    if (EPI != POST_NONE && geo.nh == 2 && k >= 1) pb_wait_step(sm, w ^ 1, k - 1);
    // ---- after the wait
    D2 dn{0.0, 0.0};
    if (dn_ok && st.live) dn = sy_lds2(sm, st.a_next + po);
    if (FIRST && PRE && dn_ok && t >= 1 && t <= (int)p.n - 1) {
        // multigrid.cpp:83 on the second-colour nodes of row t (the first colour is overwritten unread)
        const int g = k / GROUP;
        const D2 e = prolong_pair(p, geo, sm, st, PAR, t, g % NGROUP, tl.R0 + g * GROUP);
        const unsigned cm = PAR ? st.cmask1 : st.cmask0;
        if (cm & 1u) dn.x = __dadd_rn(dn.x, e.x);
        if (cm & 2u) dn.y = __dadd_rn(dn.y, e.y);
        if (cm == 3u) sy_sts2(sm, st.a_next + po, dn);
        else {
            if (cm & 1u) sy_sts1(sm, st.a_next + po, dn.x);
            if (cm & 2u) sy_sts1(sm, st.a_next + po + 8u, dn.y);
        }
    }
    const double o0 = Arith<ARITH>::gs_tail(h0, dn.x, n1, c0.d, c0.b, p.st);
    const double o1 = Arith<ARITH>::gs_tail(h1, dn.y, n2, c1.d, c1.b, p.st);
    const bool act = row >= st.lo && row <= st.hi;
    const unsigned okb = act ? (PAR ? st.okp1 : st.okp0) : 0u;
    const unsigned rm = PAR ? st.rmask1 : st.rmask0;
    // one 16-byte store in the common case: two 8-byte stores at a 16-byte lane stride cost twice the wavefronts
    if (okb == 3u) sy_sts2(sm, c, D2{o0, o1});
    else {
        if (okb & 1u) sy_sts1(sm, c, o0);
        if (okb & 2u) sy_sts1(sm, c + 8u, o1);
    }
    if (EPI != POST_NONE) {
        // what the ring now holds at the lane's targets: the new value, or the old one where masked
        D2 oe{o0, o1};
        if (okb != 3u) {
            D2 old{0.0, 0.0};
            if (own_ok && st.live) old = sy_lds2(sm, c);
            if (!(okb & 1u)) oe.x = old.x;
            if (!(okb & 2u)) oe.y = old.y;
        }
        // residual (gs.cpp:75) of the FIRST-colour nodes of row q = row-1 (run PAR): own value = up,
        // upper neighbour = the targets of two steps ago, lower = this step's, horizontal = last step's
        const int q = row - 1;
        if ((EPI == POST_NORM2 || PAR == 0) && q >= st.elo && q <= st.ehi) {
            const unsigned cq = st.a_prev + po;
            D2 fq{0.0, 0.0}, a{0.0, 0.0}, b{0.0, 0.0};
            if (st.live) { fq = sy_lds2(sm, cq + geo.ringb); a = sy_lds2(sm, cq + 2u * geo.ringb); b = sy_lds2(sm, cq + 3u * geo.ringb); }
            const double xq = sy_side(sm, st.o1, st.a_prev + oo + (PAR ? 16u : 0u) - (PAR ? 0u : 8u), PAR ? 1 : -1,
                                      PAR ? st.outer_r : st.outer_l);
            const double m0 = PAR ? st.o1.x : xq, m1 = PAR ? st.o1.y : st.o1.x, m2 = PAR ? xq : st.o1.y;
            const Coef4 q0 = Arith<ARITH>::coef(a.x, b.x, p.st), q1 = Arith<ARITH>::coef(a.y, b.y, p.st);
            const double r0 = Arith<ARITH>::residual(fq.x, up.x, st.o2.x, m0, oe.x, m1, q0, p.st);
            const double r1 = Arith<ARITH>::residual(fq.y, up.y, st.o2.y, m1, oe.y, m2, q1, p.st);
            if (EPI == POST_INJECT) {
                // gs.cpp:283: coarse node (q/2, kg) for fine even column 2kg; kg even -> E run, kg+1 -> O run
                const long kg = (long)tl.k0 + st.kk;
                double* crow = p.crhs + ((long)(q >> 1) - p.crow0) * p.cpitch + (kg >> 1);
                if (st.rmask0 & 1u) crow[0] = r0;
                if (st.rmask0 & 2u) crow[p.codd] = r1;
            } else {
                if (rm & 1u) st.acc += r0 * r0;
                if (rm & 2u) st.acc += r1 * r1;
            }
        }
        if (EPI == POST_NORM2 && row >= st.elo && row <= st.ehi) {
            // the second colour: the nodes just updated, every operand in registers
            const double r0 = Arith<ARITH>::residual(f.x, o0, up.x, n0, dn.x, n1, c0, p.st);
            const double r1 = Arith<ARITH>::residual(f.y, o1, up.y, n1, dn.y, n2, c1, p.st);
            if (rm & 1u) st.acc += r0 * r0;
            if (rm & 2u) st.acc += r1 * r1;
        }
        st.o2 = st.o1; st.o1 = oe;
    }
    if (LAST) sy_fence_async();
    st.c_up = m; st.c_mid = dn;
    // ---- publish the step
    sy_syncwarp();
    if (sy_elect()) sy_pb_arrive(sm, w, k & (PBSLOTS - 1));
    advance_row(geo, st);
}

template <int ARITH, bool FIRST, bool PRE, bool LAST, int EPI>
SY_FN void stage_loop(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, Stage& st)
{
    const int klast = last_step(p, tl) - first_step(tl);
    // column parity of the stage's colour in its row at step k: (s + row) & 1 with row = R0 + k - 1 - 2s
    int par = (st.s + st.row) & 1;
    int k = 0;
    // first steps: rows above R0 (checked); then align the unrolled loop on parity 0; last steps: rows below R1
    // (checked: the slots of unstaged rows may still be in use by the stages behind)
    const int kchk = 2 * st.s + 1;                       // first step with row >= R0
    int kmain = tl.R1 - tl.R0 + 2 * st.s;                // last step with row + 1 <= R1
    if (kmain > klast) kmain = klast;
    for (; k <= klast && (k < kchk || par != 0); ++k, par ^= 1)
        stage_step<ARITH, FIRST, PRE, LAST, EPI>(p, tl, geo, sm, st, k, par, true);
    for (; k + 1 <= kmain; k += 2) {
        stage_step<ARITH, FIRST, PRE, LAST, EPI>(p, tl, geo, sm, st, k, 0, false);
        stage_step<ARITH, FIRST, PRE, LAST, EPI>(p, tl, geo, sm, st, k + 1, 1, false);
    }
    for (; k <= klast; ++k, par ^= 1)
        stage_step<ARITH, FIRST, PRE, LAST, EPI>(p, tl, geo, sm, st, k, par, true);
}

// ------------------------------------------------------------------------------------------
// producer: request group g (first memory row z) into its ring slot: the four fields plus the three
// coarse rows its prolongation needs, all completing on the slot's full barrier; then start the HBM
// fetch of the group after next, whose shared-memory request will hit L2.  Whole (converged) warp,
// warp-uniform operands; one elected lane issues.
SY_FN void issue_group_loads(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, int g)
{
    if (!sy_elect()) return;
    const int gs = g % NGROUP;
    const int z = tl.R0 + GROUP * g - (int)p.row0;
    const unsigned fine = (unsigned)GROUP * geo.rowb;
    sy_full_expect(sm, gs, 4u * fine + (p.pre ? geo.cgrpb : 0u));
    const unsigned so = (unsigned)gs * fine;
    // a zero iterate is produced by a box that lies entirely below the last memory row
    sy_tma_load(p, sm, FIELD_U, so, tl.k0, p.u_is_zero ? (int)p.rows_mem + 64 : z, gs);
    sy_tma_load(p, sm, FIELD_F, so + geo.ringb, tl.k0, z, gs);
    sy_tma_load(p, sm, FIELD_V1, so + 2u * geo.ringb, tl.k0, z, gs);
    sy_tma_load(p, sm, FIELD_V2, so + 3u * geo.ringb, tl.k0, z, gs);
    if (p.pre)
        sy_tma_load(p, sm, FIELD_C, 4u * geo.ringb + (unsigned)gs * geo.cgrpb, tl.k0 / 2, ((z + (int)p.row0) >> 1) - (int)p.crow0, gs);
    const int zp = z + 2 * GROUP;
    if (g + 2 >= NGROUP && zp + (int)p.row0 <= tl.R1) {               // (the first NGROUP groups are requested at once)
        if (!p.u_is_zero) sy_tma_prefetch(p, FIELD_U, tl.k0, zp);
        sy_tma_prefetch(p, FIELD_F, tl.k0, zp);
        sy_tma_prefetch(p, FIELD_V1, tl.k0, zp);
        sy_tma_prefetch(p, FIELD_V2, tl.k0, zp);
    }
}

// The producer follows the last stage: in step k that stage finishes row r = t - 4K + 1.  The row is
// stored as soon as both of its warps have published the step; group g is requested once the last
// stage has completed row (first row of g) - REFILL_ROW, its last use of the slot's previous rows
// (rhs / v1 / v2 of the row above for the residual).
SY_FN void producer_loop(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm)
{
    const int ng = num_groups(tl);
    int gnext = 0;
    for (; gnext < ng && gnext < NGROUP; ++gnext) issue_group_loads(p, tl, geo, sm, gnext);   // fresh slots
    const int klast = last_step(p, tl) - first_step(tl);
    const int wl = (2 * p.K - 1) * NSTW;                 // the last stage's first warp
    long eE = (long)tl.kb + p.WK, eO = eE;
    if (eE > p.nhalf + 1) eE = p.nhalf + 1;
    if (eO > p.nhalf) eO = p.nhalf;
    const unsigned nE8 = (unsigned)(((eE - tl.kb + 1) & ~1L) * 8);   // whole 16-byte units (the layout has slack)
    const unsigned nO8 = (unsigned)((eO - tl.kb) * 8);
    int r = tl.R0 - 4 * p.K + 1;                                     // the last stage's row in step 0
    double* gst = p.u_out + ((long)r - p.row0) * p.pitch + tl.kb;
    unsigned a = (unsigned)((((r - tl.R0) % RING) + RING) % RING) * geo.rowb + (unsigned)HK * 8u;
    for (int k = 0; k <= klast; ++k, ++r) {
        const bool store = r >= tl.rb0 && r <= tl.rb1;
        const bool load = gnext < ng && r >= tl.R0 + GROUP * gnext - REFILL_ROW;
        if (store || load) {
            pb_wait_step(sm, wl, k);
            if (geo.nh == 2) pb_wait_step(sm, wl + 1, k);
        }
        if (store && sy_elect()) {
            if (nE8) sy_bulk_store(sm, gst, a, nE8);
            if (nO8) sy_bulk_store(sm, gst + p.odd, a + geo.swkb, nO8);
            sy_store_commit();
        }
        if (load) {
            if (sy_elect()) { if (store) sy_store_wait_read1(); else sy_store_wait_read0(); }
            issue_group_loads(p, tl, geo, sm, gnext);
            ++gnext;
        }
        gst += p.pitch;
        a += geo.rowb;
        if (a >= geo.ringb + (unsigned)HK * 8u) a -= geo.ringb;
    }
    if (sy_elect()) sy_store_wait_all();
}

// the work of one warp on one tile; returns the thread's share of the POST_NORM2 sum
template <int ARITH, bool PRE, int POSTK>
SY_FN double run_warp(const Params& p, const Tile& tl, const Geo& geo, const Smem& sm, int warp, int lane)
{
    if (warp == PRODUCER_WARP) {
        // the producer's lanes run in lockstep on the device (warp-uniform code, one elected lane issues); the host
        // model runs lanes as free threads, so it lets only the issuing lane walk the loop
#ifdef SY_HOST_MODEL
        if (lane != 0) return 0.0;
#endif
        producer_loop(p, tl, geo, sm);
        return 0.0;
    }
    const int s = warp / NSTW, h = warp % NSTW;
    if (s >= 2 * p.K || h >= geo.nh) return 0.0;
    Stage st = init_stage(p, tl, geo, warp, lane);
    if (s == 0) stage_loop<ARITH, true, PRE, false, POST_NONE>(p, tl, geo, sm, st);
    else if (s == 2 * p.K - 1) stage_loop<ARITH, false, false, true, POSTK>(p, tl, geo, sm, st);
    else stage_loop<ARITH, false, false, false, POST_NONE>(p, tl, geo, sm, st);
    return st.acc;
}

// ------------------------------------------------------------------------------------------
// Tile planner.  Cost model (relative): a tile takes (rows + fill) steps, a step costs
// c0 + SWK (latency + work proportional to the strip width); tiles run one per SM in waves of
// `sms`.  Search strip width (SWK a multiple of 16: TMA boxes land 128-byte aligned) and band
// count for the cheapest plan.
struct Plan { int WK, SWK, nstrips, nbands; long RBAND; };

// n: level size (columns); nrows: rows this launch produces (n+1 on a single GPU, the slab otherwise)
inline Plan make_plan(long n, long nrows, int K, int sms, int force_swk = 0)
{
    const long npairs = n / 2 + 1;
    Plan best{};
    double best_cost = 1e300;
    for (int SWK = 32; SWK <= SWK_MAX; SWK += 16) {
        if (force_swk && SWK != force_swk) continue;
        const int WK = SWK - 2 * HK;
        const int nstrips = (int)((npairs + WK - 1) / WK);
        for (int nb = 1; nb <= 4096; nb = nb < 16 ? nb + 1 : nb * 2) {
            const long RB = (nrows + nb - 1) / nb;
            if (nb > 1 && RB < 8) break;
            const int nbands = (int)((nrows + RB - 1) / RB);
            const long tiles = (long)nstrips * nbands;
            const long waves = (tiles + sms - 1) / sms;
            const double steps = (double)RB + 2.0 * (2 * K + 1) + 4.0 * K + 2.0 * GROUP;
            const double per_step = 96.0 + SWK;
            const double cost = (double)waves * steps * per_step;
            if (cost < best_cost) { best_cost = cost; best = Plan{WK, SWK, nstrips, nbands, RB}; }
        }
    }
    return best;
}

}  // namespace sy
}  // namespace mgb200
