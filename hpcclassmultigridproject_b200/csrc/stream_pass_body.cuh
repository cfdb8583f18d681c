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
// Warp roles (13 warps): 6 stage warps (one half-sweep each, 4 nodes per lane as two 16-byte
// vectors), 2 prolongation warps, 4 residual-epilogue warps, 1 producer warp (one lane drives the
// TMA engine).  Per-thread state that advances by one row per step (ring slots, rows) is kept
// incrementally, so a step has no division and almost no address arithmetic.
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
constexpr int RING = 24;       // fine-row ring slots  >= DEPTH + 4K + 4
constexpr int DEPTH = 8;       // rows in flight ahead of the compute front
constexpr int CRING = 8;       // coarse-row ring slots
constexpr int CW = 72;         // doubles per coarse parity run in smem (>= SWK_MAX/2 + 2)
constexpr int KMAX = 3;
constexpr int NSTAGE = 2 * KMAX;
constexpr int NPRE = 2;        // prolongation warps (64 pairs each)
constexpr int NPOST = 4;       // residual-epilogue warps (32 pairs each)
constexpr int WARPS = NSTAGE + NPRE + NPOST + 1;
constexpr int THREADS = WARPS * 32;
constexpr int PRODUCER_WARP = WARPS - 1;

struct Params {
    long n, nhalf;             // level size, n/2
    long pitch, odd;           // split layout of this level
    long cpitch, codd;         // split layout of the next coarser level
    long RBAND;                // owned rows per band
    int WK, SWK;               // owned pairs per strip, pairs per smem row (WK + 2 HK)
    int nstrips, nbands;
    int K;                     // fused RB iterations, 0..3
    int pre;                   // 1: u += P(coarse u) before smoothing
    int post;                  // StreamPost
    int write_u;               // 0 for a pure residual pass
    int u_is_zero;             // u_in == 0 everywhere: rows come from zero_row
    Stencil st;
    const double* u_in;
    const double* rhs;
    const double* v1;
    const double* v2;
    const double* cu;          // coarse u (pre)
    const double* zero_row;    // >= 2*SWK_MAX zeros
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
    double* U;                 // [RING][2][SWK_MAX]
    double* F;                 // rhs
    double* V1;
    double* V2;
    double* C;                 // [CRING][2][CW]
    unsigned long long* full;  // [RING] load-completion barriers
};

constexpr size_t SMEM_BYTES = (size_t)4 * RING * 2 * SWK_MAX * 8 + (size_t)CRING * 2 * CW * 8 + RING * 8 + 128;

struct alignas(16) D2 { double x, y; };    // 16-byte vector; aligned accesses only (even pair index)

SP_FN D2 ld2(const double* p) { return *reinterpret_cast<const D2*>(p); }
SP_FN void st2(double* p, D2 v) { *reinterpret_cast<D2*>(p) = v; }

SP_FN void carve(Smem& sm, unsigned char* base)
{
    const size_t row = (size_t)RING * 2 * SWK_MAX;
    double* d = reinterpret_cast<double*>(base);
    sm.U = d; sm.F = d + row; sm.V1 = d + 2 * row; sm.V2 = d + 3 * row;
    sm.C = d + 4 * row;
    sm.full = reinterpret_cast<unsigned long long*>(sm.C + (size_t)CRING * 2 * CW);
}

SP_FN Tile make_tile(const Params& p, long tile)
{
    Tile tl;
    const long strip = tile % p.nstrips, band = tile / p.nstrips;
    const long HR = 2 * p.K + 1;
    long rb0 = band * p.RBAND, rb1 = rb0 + p.RBAND - 1;
    if (rb1 > p.n) rb1 = p.n;
    long R0 = rb0 - HR, R1 = rb1 + HR;
    if (R0 < 0) R0 = 0;
    if (R1 > p.n) R1 = p.n;
    tl.kb = (int)(strip * p.WK);
    tl.k0 = tl.kb - HK;
    tl.rb0 = (int)rb0; tl.rb1 = (int)rb1; tl.R0 = (int)R0; tl.R1 = (int)R1;
    return tl;
}

SP_FN int wrap_slot(int s) { return s >= RING ? s - RING : (s < 0 ? s + RING : s); }
SP_FN int ring_slot(const Tile& tl, int row) { return (row - tl.R0) % RING; }
SP_FN int rowix(int slot, int par) { return (slot * 2 + par) * SWK_MAX; }

// ------------------------------------------------------------------------------------------
// asynchronous primitives
SP_FN void sp_bar_expect(unsigned long long* bar, unsigned bytes);
SP_FN void sp_bulk_load(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar);
SP_FN void sp_bar_wait(unsigned long long* bar, unsigned parity);
SP_FN void sp_bulk_store(double* gdst, const double* ssrc, unsigned bytes);
SP_FN void sp_store_commit();
SP_FN void sp_store_wait_read2();
SP_FN void sp_fence_async();

// ------------------------------------------------------------------------------------------
// producer: bulk loads of fine row r (4 fields x 2 parity runs) and of the coarse rows that
// travel with it, all completing on full[slot(r)]
SP_FN void issue_row_loads(const Params& p, const Tile& tl, const Smem& sm, int r)
{
    const int slot = ring_slot(tl, r);
    unsigned long long* bar = &sm.full[slot];
    // even run: pairs [k0, k0+SWK) clipped to [0, nhalf+1 rounded up to even); odd run: to [0, nhalf)
    const long s0 = tl.k0 < 0 ? 0 : tl.k0;
    long eE = (long)tl.k0 + p.SWK, eO = eE;
    const long capE = (p.nhalf + 2) & ~1L, capO = p.nhalf;
    if (eE > capE) eE = capE;
    if (eO > capO) eO = capO;
    const long nE = eE > s0 ? eE - s0 : 0, nO = eO > s0 ? eO - s0 : 0;
    // coarse rows loaded with this fine row: I is first needed by fine row 2I-1
    long cI[2]; int ncI = 0;
    if (p.pre) {
        if (r == tl.R0) {
            cI[ncI++] = r >> 1;
            if (r & 1) cI[ncI++] = (r + 1) >> 1;
        } else if (r & 1) {
            cI[ncI++] = (r + 1) >> 1;
        }
    }
    const long m0 = tl.k0 / 2;                          // k0 is a multiple of 4 (or -4)
    const long ms = m0 < 0 ? 0 : m0;
    long ceE = m0 + p.SWK / 2 + 2, ceO = m0 + p.SWK / 2;
    const long ccapE = (p.nhalf / 2 + 2) & ~1L, ccapO = p.nhalf / 2;
    if (ceE > ccapE) ceE = ccapE;
    if (ceO > ccapO) ceO = ccapO;
    const long cnE = ceE > ms ? ceE - ms : 0, cnO = ceO > ms ? ceO - ms : 0;

    const unsigned total = (unsigned)(4 * (nE + nO) * 8 + ncI * (cnE + cnO) * 8);
    sp_bar_expect(bar, total);
    const long goff = (long)r * p.pitch + s0;
    const long soff = s0 - tl.k0;
    const double* usrc = p.u_is_zero ? p.zero_row : p.u_in + goff;
    const double* usrcO = p.u_is_zero ? p.zero_row : p.u_in + goff + p.odd;
    if (nE > 0) {
        sp_bulk_load(sm.U + rowix(slot, 0) + soff, usrc, (unsigned)(nE * 8), bar);
        sp_bulk_load(sm.F + rowix(slot, 0) + soff, p.rhs + goff, (unsigned)(nE * 8), bar);
        sp_bulk_load(sm.V1 + rowix(slot, 0) + soff, p.v1 + goff, (unsigned)(nE * 8), bar);
        sp_bulk_load(sm.V2 + rowix(slot, 0) + soff, p.v2 + goff, (unsigned)(nE * 8), bar);
    }
    if (nO > 0) {
        sp_bulk_load(sm.U + rowix(slot, 1) + soff, usrcO, (unsigned)(nO * 8), bar);
        sp_bulk_load(sm.F + rowix(slot, 1) + soff, p.rhs + goff + p.odd, (unsigned)(nO * 8), bar);
        sp_bulk_load(sm.V1 + rowix(slot, 1) + soff, p.v1 + goff + p.odd, (unsigned)(nO * 8), bar);
        sp_bulk_load(sm.V2 + rowix(slot, 1) + soff, p.v2 + goff + p.odd, (unsigned)(nO * 8), bar);
    }
    for (int q = 0; q < ncI; ++q) {
        const long I = cI[q];
        const int cs = (int)(I % CRING);
        const long cg = I * p.cpitch + ms;
        if (cnE > 0) sp_bulk_load(sm.C + (long)(cs * 2 + 0) * CW + (ms - m0), p.cu + cg, (unsigned)(cnE * 8), bar);
        if (cnO > 0) sp_bulk_load(sm.C + (long)(cs * 2 + 1) * CW + (ms - m0), p.cu + cg + p.codd, (unsigned)(cnO * 8), bar);
    }
}

// producer: bulk store of the owned part of finished row q
SP_FN void issue_row_store(const Params& p, const Tile& tl, const Smem& sm, int q)
{
    const int slot = ring_slot(tl, q);
    const int lo = tl.kb - tl.k0;                       // = HK
    long eE = (long)tl.kb + p.WK, eO = eE;
    if (eE > p.nhalf + 1) eE = p.nhalf + 1;
    if (eO > p.nhalf) eO = p.nhalf;
    const long nE = (eE - tl.kb + 1) & ~1L;             // rounded up to a whole 16-byte unit (layout slack)
    const long nO = eO - tl.kb;
    const long goff = (long)q * p.pitch + tl.kb;
    if (nE > 0) sp_bulk_store(p.u_out + goff, sm.U + rowix(slot, 0) + lo, (unsigned)(nE * 8));
    if (nO > 0) sp_bulk_store(p.u_out + goff + p.odd, sm.U + rowix(slot, 1) + lo, (unsigned)(nO * 8));
    sp_store_commit();
}

// ------------------------------------------------------------------------------------------
// per-thread state, advanced by one row per step
enum Role { ROLE_STAGE = 0, ROLE_PRE = 1, ROLE_POST = 2, ROLE_PRODUCER = 3 };

struct ThreadState {
    int role;
    int idx;       // stage number / chunk number
    int kk;        // first local pair handled (stage: also kk + 64)
    int row;       // the role's row at the current step
    int slot;      // ring slot of that row
    unsigned ok;   // validity bits (role specific)
    int wslot;     // ring slot of the row whose arrival is awaited at the end of the step
    unsigned wpar; // and its phase parity
    double acc;    // POST_NORM2 accumulator
};

// validity of a GS / residual target at local pair kk of parity par
SP_FN bool target_ok(const Params& p, const Tile& tl, int kk, int par)
{
    const long kg = (long)tl.k0 + kk;
    if (kk >= p.SWK) return false;
    if (par == 0) return kk >= 1 && kg >= 1 && kg <= p.nhalf - 1;          // even column 2kg: neighbours O[kk-1], O[kk]
    return kk <= p.SWK - 2 && kg >= 0 && kg <= p.nhalf - 1;                 // odd column 2kg+1: neighbours E[kk], E[kk+1]
}

SP_FN ThreadState init_thread(const Params& p, const Tile& tl, int tid)
{
    ThreadState s;
    const int warp = tid >> 5, lane = tid & 31;
    s.acc = 0.0; s.ok = 0; s.idx = 0; s.kk = 0;
    int off;                                             // role row = t - off
    if (warp < NSTAGE) {
        s.role = ROLE_STAGE; s.idx = warp; s.kk = 2 * lane; off = 2 + 2 * warp;
        for (int g = 0; g < 2; ++g)
            for (int e = 0; e < 2; ++e)
                for (int par = 0; par < 2; ++par)
                    if (target_ok(p, tl, s.kk + 64 * g + e, par)) s.ok |= 1u << (par * 4 + g * 2 + e);
    } else if (warp < NSTAGE + NPRE) {
        s.role = ROLE_PRE; s.idx = warp - NSTAGE; s.kk = 64 * s.idx + 2 * lane; off = 0;
        for (int e = 0; e < 2; ++e) {
            const long kg = (long)tl.k0 + s.kk + e;
            if (s.kk + e < p.SWK && kg >= 0 && kg <= p.nhalf - 1) s.ok |= 1u << e;          // pair holds an interior odd column
            if (s.kk + e < p.SWK && kg >= 1 && kg <= p.nhalf - 1) s.ok |= 1u << (2 + e);    // ... and an interior even column
        }
    } else if (warp < NSTAGE + NPRE + NPOST) {
        s.role = ROLE_POST; s.idx = warp - NSTAGE - NPRE; s.kk = HK + 32 * s.idx + lane; off = 4 * p.K + 2;
        if (s.kk < HK + p.WK) {                          // owned pairs only
            if (target_ok(p, tl, s.kk, 0)) s.ok |= 1u;
            if (target_ok(p, tl, s.kk, 1)) s.ok |= 2u;
        }
    } else {
        s.role = ROLE_PRODUCER; off = 0;
    }
    s.row = tl.R0 - off;
    s.slot = ((RING - off) % RING + RING) % RING;        // slot of row R0 is 0
    s.wslot = 0; s.wpar = 0;
    return s;
}

// one half-sweep stage on row st.row: colour = stage & 1; two 16-byte vectors per lane
template <int ARITH>
SP_FN void stage_row(const Params& p, const Tile& tl, const Smem& sm, const ThreadState& st)
{
    const int i = st.row;
    if (st.idx >= 2 * p.K || i <= tl.R0 || i >= tl.R1) return;      // rows i-1 and i+1 must be staged
    const int par = (st.idx + i) & 1;                                 // column parity of this colour in row i
    const unsigned ok = (st.ok >> (par * 4)) & 15u;
    const int sc = st.slot, su = wrap_slot(sc - 1), sd = wrap_slot(sc + 1);
    const int bc = rowix(sc, par), bo = rowix(sc, par ^ 1), bu = rowix(su, par), bd = rowix(sd, par);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const unsigned okg = (ok >> (2 * g)) & 3u;
        if (okg == 0) continue;
        const int kk = st.kk + 64 * g;
        const D2 up = ld2(sm.U + bu + kk), dn = ld2(sm.U + bd + kk);
        const D2 f = ld2(sm.F + bc + kk), w1 = ld2(sm.V1 + bc + kk), w2 = ld2(sm.V2 + bc + kk);
        const D2 mid = ld2(sm.U + bo + kk);                           // other parity, same pair indices
        double lf0, rt0, lf1, rt1;
        if (par == 0) {            // even columns: left = O[kk-1], right = O[kk]
            const double prev = (okg & 1u) ? sm.U[bo + kk - 1] : 0.0;
            lf0 = prev; rt0 = mid.x; lf1 = mid.x; rt1 = mid.y;
        } else {                   // odd columns: left = E[kk], right = E[kk+1]
            const double next = (okg & 2u) ? sm.U[bo + kk + 2] : 0.0;
            lf0 = mid.x; rt0 = mid.y; lf1 = mid.y; rt1 = next;
        }
        const Coef4 c0 = Arith<ARITH>::coef(w1.x, w2.x, p.st);
        const Coef4 c1 = Arith<ARITH>::coef(w1.y, w2.y, p.st);
        D2 out;
        out.x = Arith<ARITH>::gs(f.x, up.x, lf0, dn.x, rt0, c0, p.st);
        out.y = Arith<ARITH>::gs(f.y, up.y, lf1, dn.y, rt1, c1, p.st);
        if (okg == 3u) st2(sm.U + bc + kk, out);
        else if (okg == 1u) sm.U[bc + kk] = out.x;
        else sm.U[bc + kk + 1] = out.y;
    }
}

// prolongation + correction of row t (gs.cpp:238-241 fused with multigrid.cpp:83), interior nodes.
// Coarse column (k0 + kk) sits at local index kk of the coarse ring row: even kk in the E run at
// kk/2, odd kk in the O run at kk/2.
SP_FN void prolong_row(const Params& p, const Tile& tl, const Smem& sm, const ThreadState& st)
{
    const int t = st.row;
    if (!p.pre || (st.ok & 3u) == 0 || t < 1 || t > p.n - 1 || t < tl.R0 || t > tl.R1) return;
    const int kk = st.kk, a = kk >> 1;                  // kk even
    const int I = t >> 1;
    const int c0 = (I % CRING) * 2 * CW, c1 = ((I + 1) % CRING) * 2 * CW;
    const int be = rowix(st.slot, 0), bo = rowix(st.slot, 1);
    // coarse columns kk, kk+1, kk+2 of coarse row I (and I+1 for an odd fine row)
    const double a0 = sm.C[c0 + a], a1 = sm.C[c0 + CW + a], a2 = sm.C[c0 + a + 1];
    D2 ue = ld2(sm.U + be + kk), uo = ld2(sm.U + bo + kk);
    double e0, e1, o0, o1;
    if ((t & 1) == 0) {
        e0 = a0; e1 = a1;                                                                   // gs.cpp:238
        o0 = __dmul_rn(__dadd_rn(a0, a1), 0.5); o1 = __dmul_rn(__dadd_rn(a1, a2), 0.5);     // gs.cpp:240
    } else {
        const double b0 = sm.C[c1 + a], b1 = sm.C[c1 + CW + a], b2 = sm.C[c1 + a + 1];
        e0 = __dmul_rn(__dadd_rn(a0, b0), 0.5); e1 = __dmul_rn(__dadd_rn(a1, b1), 0.5);     // gs.cpp:239
        o0 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a0, b0), a1), b1), 0.25);              // gs.cpp:241
        o1 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a1, b1), a2), b2), 0.25);
    }
    if (st.ok & 4u) sm.U[be + kk] = __dadd_rn(ue.x, e0);                                    // multigrid.cpp:83
    if (st.ok & 8u) sm.U[be + kk + 1] = __dadd_rn(ue.y, e1);
    if (st.ok & 1u) sm.U[bo + kk] = __dadd_rn(uo.x, o0);
    if (st.ok & 2u) sm.U[bo + kk + 1] = __dadd_rn(uo.y, o1);
}

// residual epilogue on finished row q: injection into the coarse rhs or sum of squares
template <int ARITH>
SP_FN void post_row(const Params& p, const Tile& tl, const Smem& sm, ThreadState& st)
{
    const int q = st.row;
    const int lo = tl.rb0 < 1 ? 1 : tl.rb0, hi = tl.rb1 > p.n - 1 ? (int)p.n - 1 : tl.rb1;
    if (p.post == POST_NONE || st.ok == 0 || q < lo || q > hi) return;
    if (p.post == POST_INJECT && (q & 1)) return;
    const int kk = st.kk;
    const int sc = st.slot, su = wrap_slot(sc - 1), sd = wrap_slot(sc + 1);
    const int e = rowix(sc, 0) + kk, o = rowix(sc, 1) + kk;
    const double ue = sm.U[e], uo = sm.U[o];
    if (st.ok & 1u) {                                    // even column 2kg
        const Coef4 c = Arith<ARITH>::coef(sm.V1[e], sm.V2[e], p.st);
        const double rv = Arith<ARITH>::residual(sm.F[e], ue, sm.U[rowix(su, 0) + kk], sm.U[o - 1],
                                                 sm.U[rowix(sd, 0) + kk], uo, c, p.st);
        if (p.post == POST_INJECT) {
            const long kg = (long)tl.k0 + kk;
            p.crhs[(long)(q >> 1) * p.cpitch + (kg & 1) * p.codd + (kg >> 1)] = rv;         // gs.cpp:283
        } else {
            st.acc += rv * rv;
        }
    }
    if (p.post == POST_NORM2 && (st.ok & 2u)) {          // odd column 2kg+1
        const Coef4 c = Arith<ARITH>::coef(sm.V1[o], sm.V2[o], p.st);
        const double rv = Arith<ARITH>::residual(sm.F[o], uo, sm.U[rowix(su, 1) + kk], ue,
                                                 sm.U[rowix(sd, 1) + kk], sm.U[e + 1], c, p.st);
        st.acc += rv * rv;
    }
}

// ------------------------------------------------------------------------------------------
SP_FN int first_step(const Tile& tl) { return tl.R0; }
SP_FN int last_step(const Params& p, const Tile& tl) { return tl.rb1 + 4 * p.K + 2; }

// producer prologue: the first DEPTH rows
SP_FN void producer_prologue(const Params& p, const Tile& tl, const Smem& sm)
{
    for (int r = tl.R0; r < tl.R0 + DEPTH && r <= tl.R1; ++r) issue_row_loads(p, tl, sm, r);
}

// wait until row R0 has landed (every thread, before the first step)
SP_FN void wait_first_row(const Smem& sm) { sp_bar_wait(&sm.full[0], 0u); }

// everything a thread does in step t; a block barrier separates consecutive steps.  On entry row
// t has landed (awaited at the end of the previous step).
template <int ARITH>
SP_FN void thread_step(const Params& p, const Tile& tl, const Smem& sm, ThreadState& st, int t, int lane)
{
    if (st.role == ROLE_PRODUCER) {
        if (lane == 0) {
            const int q = t - 4 * p.K - 1;              // finished by the previous step
            if (p.write_u && q >= tl.rb0 && q <= tl.rb1) issue_row_store(p, tl, sm, q);
            const int r = t + DEPTH;
            if (r <= tl.R1) {
                sp_store_wait_read2();                  // the slot's previous row has left shared memory
                issue_row_loads(p, tl, sm, r);
            }
        }
    } else if (st.role == ROLE_STAGE) {
        stage_row<ARITH>(p, tl, sm, st);
        // the last stage's rows go to the bulk-store engine next step: generic -> async proxy
        if (st.idx == 2 * p.K - 1) sp_fence_async();
    } else if (st.role == ROLE_PRE) {
        prolong_row(p, tl, sm, st);
        if (p.K == 0) sp_fence_async();
    } else {
        post_row<ARITH>(p, tl, sm, st);
    }
    st.row += 1;
    st.slot = wrap_slot(st.slot + 1);
    // row t+1 must have landed before anyone touches it in step t+1
    st.wslot += 1;
    if (st.wslot == RING) { st.wslot = 0; st.wpar ^= 1u; }
    if (t + 1 <= tl.R1) sp_bar_wait(&sm.full[st.wslot], st.wpar);
}

// ------------------------------------------------------------------------------------------
// Tile planner.  Cost model (relative): a tile takes (rows + fill) steps, a step costs
// c0 + SWK (barrier latency + work proportional to the strip width); tiles run one per SM in
// waves of `sms`.  Search strip width and band count for the cheapest plan.
struct Plan { int WK, SWK, nstrips, nbands; long RBAND; };

inline Plan make_plan(long n, int K, int sms)
{
    const long npairs = n / 2 + 1, nrows = n + 1;
    Plan best{};
    double best_cost = 1e300;
    for (int WK = 16; WK <= SWK_MAX - 2 * HK; WK += 4) {
        const int nstrips = (int)((npairs + WK - 1) / WK);
        for (int nb = 1; nb <= 4096; nb = nb < 16 ? nb + 1 : nb * 2) {
            const long RB = (nrows + nb - 1) / nb;
            if (nb > 1 && RB < 32) break;
            const int nbands = (int)((nrows + RB - 1) / RB);
            const long tiles = (long)nstrips * nbands;
            const long waves = (tiles + sms - 1) / sms;
            const double steps = (double)RB + 2.0 * (2 * K + 1) + 4.0 * K + 2.0 + DEPTH;
            const double cost = (double)waves * steps * (40.0 + WK + 2 * HK);
            if (cost < best_cost) { best_cost = cost; best = Plan{WK, WK + 2 * HK, nstrips, nbands, RB}; }
        }
    }
    return best;
}

}  // namespace sp
}  // namespace mgb200
