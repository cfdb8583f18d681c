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
// This header holds no CUDA-specific instruction: everything asynchronous goes through the sp_*
// primitives declared at the top, defined with inline PTX in stream_pass.cu.
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
constexpr int NSTW = 2;        // warps per half-sweep stage
constexpr int NAUX = 2;        // warps for prolongation / residual epilogue
constexpr int WARPS = 2 * KMAX * NSTW + NAUX + 1;
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
    long R0, R1;               // rows staged: [R0, R1]
    long rb0, rb1;             // rows owned:  [rb0, rb1]
    long k0;                   // global pair index of smem column 0 (may be -HK)
    long kb;                   // first owned pair
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
    tl.kb = strip * p.WK;
    tl.k0 = tl.kb - HK;
    tl.rb0 = band * p.RBAND;
    tl.rb1 = tl.rb0 + p.RBAND - 1;
    if (tl.rb1 > p.n) tl.rb1 = p.n;
    tl.R0 = tl.rb0 - HR; if (tl.R0 < 0) tl.R0 = 0;
    tl.R1 = tl.rb1 + HR; if (tl.R1 > p.n) tl.R1 = p.n;
    return tl;
}

SP_FN int ring_slot(const Tile& tl, long row) { return (int)(row - tl.R0) % RING; }
SP_FN long rowix(int slot, int par) { return (long)(slot * 2 + par) * SWK_MAX; }

// ------------------------------------------------------------------------------------------
// asynchronous primitives (PTX in stream_pass.cu)
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
SP_FN void issue_row_loads(const Params& p, const Tile& tl, const Smem& sm, long r)
{
    const int slot = ring_slot(tl, r);
    unsigned long long* bar = &sm.full[slot];
    // even run: pairs [k0, k0+SWK) clipped to [0, nhalf+1 rounded up to even); odd run: to [0, nhalf)
    const long s0 = tl.k0 < 0 ? 0 : tl.k0;
    long eE = tl.k0 + p.SWK, eO = eE;
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
    const long goff = r * p.pitch + s0;
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
SP_FN void issue_row_store(const Params& p, const Tile& tl, const Smem& sm, long q)
{
    const int slot = ring_slot(tl, q);
    const long lo = tl.kb - tl.k0;                      // = HK
    long eE = tl.kb + p.WK, eO = eE;
    if (eE > p.nhalf + 1) eE = p.nhalf + 1;
    if (eO > p.nhalf) eO = p.nhalf;
    const long nE = (eE - tl.kb + 1) & ~1L;             // rounded up to a whole 16-byte unit (layout slack)
    const long nO = eO - tl.kb;
    const long goff = q * p.pitch + tl.kb;
    if (nE > 0) sp_bulk_store(p.u_out + goff, sm.U + rowix(slot, 0) + lo, (unsigned)(nE * 8));
    if (nO > 0) sp_bulk_store(p.u_out + goff + p.odd, sm.U + rowix(slot, 1) + lo, (unsigned)(nO * 8));
    sp_store_commit();
}

// ------------------------------------------------------------------------------------------
// one half-sweep stage on row i: colour = stage & 1
template <int ARITH>
SP_FN void stage_row(const Params& p, const Tile& tl, const Smem& sm, int stage, int part, int lane, long i)
{
    if (i <= tl.R0 || i >= tl.R1) return;               // rows i-1 and i+1 must be staged
    const int par = (int)((stage + i) & 1);             // column parity of this colour in row i
    const int su = ring_slot(tl, i - 1), sc = ring_slot(tl, i), sd = ring_slot(tl, i + 1);
    const double* Uu = sm.U + rowix(su, par);
    const double* Ud = sm.U + rowix(sd, par);
    double* Uc = sm.U + rowix(sc, par);
    const double* Uo = sm.U + rowix(sc, par ^ 1);       // the other parity of row i (left/right neighbours)
    const double* Fc = sm.F + rowix(sc, par);
    const double* V1c = sm.V1 + rowix(sc, par);
    const double* V2c = sm.V2 + rowix(sc, par);
    const int chunk = SWK_MAX / NSTW;
    const int kend = (part + 1) * chunk < p.SWK ? (part + 1) * chunk : p.SWK;
    for (int kk = part * chunk + lane; kk < kend; kk += 32) {
        const long kg = tl.k0 + kk;
        double lf, rt;
        bool ok;
        if (par == 0) {           // even column 2kg: interior 1..nhalf-1, neighbours O[kk-1], O[kk]
            ok = kk >= 1 && kg >= 1 && kg <= p.nhalf - 1;
            lf = ok ? Uo[kk - 1] : 0.0;
            rt = ok ? Uo[kk] : 0.0;
        } else {                  // odd column 2kg+1: 0..nhalf-1, neighbours E[kk], E[kk+1]
            ok = kk <= p.SWK - 2 && kg >= 0 && kg <= p.nhalf - 1;
            lf = ok ? Uo[kk] : 0.0;
            rt = ok ? Uo[kk + 1] : 0.0;
        }
        if (ok) {
            const Coef4 c = Arith<ARITH>::coef(V1c[kk], V2c[kk], p.st);
            Uc[kk] = Arith<ARITH>::gs(Fc[kk], Uu[kk], lf, Ud[kk], rt, c, p.st);
        }
    }
}

// prolongation + correction of row t (gs.cpp:238-241 fused with multigrid.cpp:83), interior nodes
SP_FN void prolong_row(const Params& p, const Tile& tl, const Smem& sm, int part, int lane, long t)
{
    if (t < 1 || t > p.n - 1 || t < tl.R0 || t > tl.R1) return;
    const int sc = ring_slot(tl, t);
    double* UE = sm.U + rowix(sc, 0);
    double* UO = sm.U + rowix(sc, 1);
    const long I = t >> 1;
    const double* C0E = sm.C + (long)((I % CRING) * 2 + 0) * CW;
    const double* C0O = sm.C + (long)((I % CRING) * 2 + 1) * CW;
    const double* C1E = sm.C + (long)(((I + 1) % CRING) * 2 + 0) * CW;
    const double* C1O = sm.C + (long)(((I + 1) % CRING) * 2 + 1) * CW;
    const bool oddrow = (t & 1) != 0;
    const int chunk = SWK_MAX / NAUX;
    const int kend = (part + 1) * chunk < p.SWK ? (part + 1) * chunk : p.SWK;
    for (int kk = part * chunk + lane; kk < kend; kk += 32) {
        const long kg = tl.k0 + kk;
        if (kg < 0 || kg > p.nhalf - 1) continue;       // pairs holding at least one interior column
        // coarse column kg lives at local index kk; kk+1 is its right neighbour
        const int a = kk >> 1, b = (kk + 1) >> 1;
        const double c00 = (kk & 1) ? C0O[a] : C0E[a];
        const double c01 = ((kk + 1) & 1) ? C0O[b] : C0E[b];
        if (!oddrow) {
            if (kg >= 1) UE[kk] = __dadd_rn(UE[kk], c00);                                         // gs.cpp:238
            UO[kk] = __dadd_rn(UO[kk], __dmul_rn(__dadd_rn(c00, c01), 0.5));                      // gs.cpp:240
        } else {
            const double c10 = (kk & 1) ? C1O[a] : C1E[a];
            const double c11 = ((kk + 1) & 1) ? C1O[b] : C1E[b];
            if (kg >= 1) UE[kk] = __dadd_rn(UE[kk], __dmul_rn(__dadd_rn(c00, c10), 0.5));         // gs.cpp:239
            double s = __dadd_rn(c00, c10);                                                        // gs.cpp:241
            s = __dadd_rn(s, c01);
            s = __dadd_rn(s, c11);
            UO[kk] = __dadd_rn(UO[kk], __dmul_rn(s, 0.25));
        }
    }
}

// residual epilogue on finished row q: injection into the coarse rhs or sum of squares
template <int ARITH>
SP_FN void post_row(const Params& p, const Tile& tl, const Smem& sm, int part, int lane, long q, double& acc)
{
    long lo = tl.rb0 < 1 ? 1 : tl.rb0, hi = tl.rb1 > p.n - 1 ? p.n - 1 : tl.rb1;
    if (q < lo || q > hi) return;
    if (p.post == POST_INJECT && (q & 1)) return;
    const int su = ring_slot(tl, q - 1), sc = ring_slot(tl, q), sd = ring_slot(tl, q + 1);
    const int klo = (int)(tl.kb - tl.k0), khi = klo + p.WK;   // owned local pairs
    const int chunk = (p.WK + NAUX - 1) / NAUX;
    const int kbeg = klo + part * chunk;
    const int kend = kbeg + chunk < khi ? kbeg + chunk : khi;
    for (int kk = kbeg + lane; kk < kend; kk += 32) {
        const long kg = tl.k0 + kk;
        // even column 2kg
        if (kg >= 1 && kg <= p.nhalf - 1) {
            const long e = rowix(sc, 0) + kk, o = rowix(sc, 1) + kk;
            const Coef4 c = Arith<ARITH>::coef(sm.V1[e], sm.V2[e], p.st);
            const double rv = Arith<ARITH>::residual(sm.F[e], sm.U[e], sm.U[rowix(su, 0) + kk], sm.U[o - 1],
                                                     sm.U[rowix(sd, 0) + kk], sm.U[o], c, p.st);
            if (p.post == POST_INJECT)
                p.crhs[(q >> 1) * p.cpitch + (kg & 1) * p.codd + (kg >> 1)] = rv;                 // gs.cpp:283
            else
                acc += rv * rv;
        }
        // odd column 2kg+1 (norm only)
        if (p.post == POST_NORM2 && kg >= 0 && kg <= p.nhalf - 1) {
            const long e = rowix(sc, 0) + kk, o = rowix(sc, 1) + kk;
            const Coef4 c = Arith<ARITH>::coef(sm.V1[o], sm.V2[o], p.st);
            const double rv = Arith<ARITH>::residual(sm.F[o], sm.U[o], sm.U[rowix(su, 1) + kk], sm.U[e],
                                                     sm.U[rowix(sd, 1) + kk], sm.U[e + 1], c, p.st);
            acc += rv * rv;
        }
    }
}

// ------------------------------------------------------------------------------------------
SP_FN long first_step(const Tile& tl) { return tl.R0; }
SP_FN long last_step(const Params& p, const Tile& tl) { return tl.rb1 + 4 * p.K + 2; }

// producer prologue: the first DEPTH rows
SP_FN void producer_prologue(const Params& p, const Tile& tl, const Smem& sm)
{
    for (long r = tl.R0; r < tl.R0 + DEPTH && r <= tl.R1; ++r) issue_row_loads(p, tl, sm, r);
}

// everything thread `tid` does in step t (a block barrier separates consecutive steps)
template <int ARITH>
SP_FN void thread_step(const Params& p, const Tile& tl, const Smem& sm, long t, int tid, double& acc)
{
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            const long q = t - 4 * p.K - 1;             // finished by the previous step
            if (p.write_u && q >= tl.rb0 && q <= tl.rb1) issue_row_store(p, tl, sm, q);
            const long r = t + DEPTH;
            if (r <= tl.R1) {
                sp_store_wait_read2();                  // the slot's previous row has left shared memory
                issue_row_loads(p, tl, sm, r);
            }
        }
        return;
    }
    if (t <= tl.R1) {                                   // row t must have landed
        const int u = (int)(t - tl.R0);
        sp_bar_wait(&sm.full[u % RING], (unsigned)((u / RING) & 1));
    }
    if (warp < 2 * KMAX * NSTW) {
        const int stage = warp / NSTW, part = warp % NSTW;
        if (stage < 2 * p.K) stage_row<ARITH>(p, tl, sm, stage, part, lane, t - 2 - 2 * stage);
    } else {
        const int part = warp - 2 * KMAX * NSTW;
        if (p.pre) prolong_row(p, tl, sm, part, lane, t);
        if (p.post != POST_NONE) post_row<ARITH>(p, tl, sm, part, lane, t - 4 * p.K - 2, acc);
    }
    sp_fence_async();                                   // my smem writes -> visible to the bulk-store engine
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
