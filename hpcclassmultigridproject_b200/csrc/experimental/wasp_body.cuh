// wasp_body.cuh -- EXPERIMENTAL (round-2 candidate, not on any default path): the "warp-autonomous"
// streaming pass of DESIGN.md section 9.  Same contract as the production pass (stream_pass.cuh:
// [u += P(coarse u)] -> K RB-GS iterations -> [residual -> injection | sum of squares], split
// layout, out of place), different execution scheme:
//
//   * one WARP owns a strip of 64 column pairs (lane l: pairs 2l, 2l+1 as one 16-byte vector per
//     parity run; 4 halo pairs per side are recomputed, 56 are owned) and a band of rows (2K+1
//     halo rows per side recomputed);
//   * the warp keeps a window of 2K+3 rows of u IN REGISTERS and, per step t, takes row t in and
//     runs ALL 2K half-sweeps itself, in order, stage s on row t-1-s (run in order by one owner the
//     stages need a lag of one row only; every stage of a step works on the same column parity,
//     (t-1)&1, which is therefore a compile-time constant of the step's code);
//   * horizontal neighbours across lanes come from shuffles, rhs / v1 / v2 are read from global
//     memory where they are needed (three reads per node and pass, the later ones from cache),
//     finished rows go to HBM with 16-byte stores from registers;
//   * no shared memory, no block barrier, no TMA: warps never communicate.
//
// The index logic is the one pinned on the CPU by tests/emu/wasp_proto.cpp.  This header holds no
// CUDA-specific instruction: lane id, shuffles and the vector accesses go through the wp_*
// primitives (inline in wasp_pass.cu; thread + barrier emulation in tests/emu/wasp_emu.cpp).
#pragma once
#include "../common.cuh"
#include "../stream_pass.cuh"

#ifndef WP_FN
#define WP_FN __device__ __forceinline__
#endif

namespace mgb200 {
namespace wasp {

constexpr int KMAX = 3;
constexpr int HP = 4;                   // halo pairs per strip side (8 columns >= 2K+1)
constexpr int PAIRS = 64;               // pairs per strip (2 per lane)
constexpr int OWN = PAIRS - 2 * HP;     // owned pairs per strip
constexpr int WARPS_PER_CTA = 4;

struct alignas(16) V2 { double x, y; };

struct Params {
    const double *u_in, *rhs, *v1, *v2, *cu;
    double *u_out, *crhs, *partials;
    long n, nhalf;                      // nhalf = n/2: even run k = 0..nhalf, odd run k = 0..nhalf-1
    long pitch, odd, cpitch, codd;
    int K, post, pre;
    int nstrips, nbands;
    long RB;                            // owned rows per band (bands tile rows own_lo..own_hi)
    // row window (row-slab sharding, same meaning as StreamPassArgs): rows own_lo..own_hi are
    // produced; the fine arrays hold global rows row0.., the coarse arrays global rows crow0..
    // A whole level is own_lo = 0, own_hi = n, row0 = crow0 = 0.
    long own_lo, own_hi, row0, crow0;
    Stencil st;
};

// ---- primitives supplied by the includer ------------------------------------------------------
WP_FN int wp_lane();
WP_FN double wp_shfl_up(double v);      // value of lane-1 (lane 0 gets its own)
WP_FN double wp_shfl_down(double v);    // value of lane+1 (lane 31 gets its own)
WP_FN V2 wp_ld2(const double* p);       // 16-byte read-only load
WP_FN void wp_st2(double* p, V2 v);
WP_FN double wp_warp_sum(double v);     // fixed tree

struct Strip {
    long k;                             // the lane's first pair (even; may lie outside the level)
    bool live;                          // the lane's two pairs are addressable
    unsigned ok[2];                     // update masks per column parity: bit 0 = .x, bit 1 = .y
    unsigned own[2];                    // owned (stored / summed) nodes per column parity, interior or not
    unsigned oint[2];                   // owned AND interior column
    long b0, b1, R0, R1;                // owned rows, staged rows
    long ulo, uhi;                      // rows the half-sweeps may update: interior rows strictly inside the staged ones
    long elo, ehi;                      // rows of the epilogue: owned interior rows
};

WP_FN Strip make_strip(const Params& p, int strip, int band, int lane)
{
    Strip s;
    const long kown0 = (long)strip * OWN;
    s.k = kown0 - HP + 2 * lane;
    s.live = s.k >= 0 && s.k <= p.nhalf;          // (the layout's slack covers pair k+1)
    s.ok[0] = s.ok[1] = s.own[0] = s.own[1] = s.oint[0] = s.oint[1] = 0;
    for (int e = 0; e < 2; ++e) {
        const long kk = s.k + e;
        if (!s.live) break;
        const bool owned = kk >= kown0 && kk < kown0 + OWN;
        // even column 2kk: exists for kk <= nhalf, interior for 1 <= kk <= nhalf-1; its left
        // neighbour O[kk-1] is the previous lane's (no such lane for the very first element)
        if (kk >= 1 && kk <= p.nhalf - 1 && !(lane == 0 && e == 0)) s.ok[0] |= 1u << e;
        if (owned && kk <= p.nhalf) s.own[0] |= 1u << e;
        if (owned && kk >= 1 && kk <= p.nhalf - 1) s.oint[0] |= 1u << e;
        // odd column 2kk+1: exists and is interior for 0 <= kk <= nhalf-1; its right neighbour
        // E[kk+1] is the next lane's for the very last element
        if (kk >= 0 && kk <= p.nhalf - 1 && !(lane == 31 && e == 1)) s.ok[1] |= 1u << e;
        if (owned && kk <= p.nhalf - 1) { s.own[1] |= 1u << e; s.oint[1] |= 1u << e; }
    }
    const long hrow = 2 * p.K + 1;
    s.b0 = p.own_lo + (long)band * p.RB;
    s.b1 = s.b0 + p.RB - 1 < p.own_hi ? s.b0 + p.RB - 1 : p.own_hi;
    s.R0 = s.b0 - hrow < 0 ? 0 : s.b0 - hrow;          // (a slab holds at least 2K+1 halo rows per side)
    s.R1 = s.b1 + hrow > p.n ? p.n : s.b1 + hrow;
    s.ulo = s.R0 + 1 > 1 ? s.R0 + 1 : 1;
    s.uhi = s.R1 - 1 < p.n - 1 ? s.R1 - 1 : p.n - 1;
    s.elo = s.b0 > 1 ? s.b0 : 1;
    s.ehi = s.b1 < p.n - 1 ? s.b1 : p.n - 1;
    return s;
}

WP_FN void st_masked(double* q, V2 v, unsigned m)
{
    if (m == 3u) wp_st2(q, v);
    else { if (m & 1u) q[0] = v.x; if (m & 2u) q[1] = v.y; }
}

// per-lane addressing: array bases with the lane's first pair folded in (a lane outside the level
// -- first / last strip only -- is parked on pair 0: it reads legal memory, and whatever it computes
// is masked (ok / own are 0) and never reaches a valid node of a live lane), row offsets as 32-bit
// element counts relative to the arrays' first held row (the launcher checks that they fit)
struct Lane {
    const double *u, *rhs, *v1, *v2, *cu;
    double* out;
    int pitch, odd, cpitch, codd;
};

WP_FN V2 ld_at(const double* base, int off) { return wp_ld2(base + off); }

// prolongation + correction of the freshly loaded row t (gs.cpp:238-241 + multigrid.cpp:83), interior
// nodes.  The lane's fine columns 2k .. 2k+3 lie between coarse columns k, k+1, k+2 (k even: even run
// index k/2, odd run index k/2, even run index k/2+1).  No bounds checks: a coarse value beyond the
// level (slack of the layout) only ever feeds a node that is masked below.
WP_FN void prolong_row(const Params& p, const Strip& s, const Lane& ln, long t, V2& E, V2& O)
{
    if (t < 1 || t > p.n - 1) return;
    const int cro = (int)(((t >> 1) - p.crow0) * ln.cpitch);
    const long k = s.k;
    const double a0 = ln.cu[cro], a1 = ln.cu[cro + ln.codd], a2 = ln.cu[cro + 1];
    double e0, e1, o0, o1;
    if ((t & 1) == 0) {
        e0 = a0; e1 = a1;                                                                   // gs.cpp:238
        o0 = __dmul_rn(__dadd_rn(a0, a1), 0.5); o1 = __dmul_rn(__dadd_rn(a1, a2), 0.5);     // gs.cpp:240
    } else {
        const int c1 = cro + ln.cpitch;
        const double b0 = ln.cu[c1], b1 = ln.cu[c1 + ln.codd], b2 = ln.cu[c1 + 1];
        e0 = __dmul_rn(__dadd_rn(a0, b0), 0.5); e1 = __dmul_rn(__dadd_rn(a1, b1), 0.5);     // gs.cpp:239
        o0 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a0, b0), a1), b1), 0.25);              // gs.cpp:241
        o1 = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a1, b1), a2), b2), 0.25);
    }
    // interior columns only: even column 2kk for 1 <= kk <= nhalf-1, odd column 2kk+1 for kk <= nhalf-1
    if (s.live) {
        if (k >= 1 && k <= p.nhalf - 1) E.x = __dadd_rn(E.x, e0);
        if (k + 1 <= p.nhalf - 1) E.y = __dadd_rn(E.y, e1);
        if (k <= p.nhalf - 1) O.x = __dadd_rn(O.x, o0);
        if (k + 1 <= p.nhalf - 1) O.y = __dadd_rn(O.y, o1);
    }
}

// operands of the nodes of column parity PAR in a row: `m` is the other run of the same row
template <int PAR>
WP_FN void side(const V2& m, double& n0, double& n1, double& n2)
{
    // odd columns: neighbours E[kk], E[kk+1], E[kk+2] (the last from the next lane);
    // even columns: O[kk-1] (from the previous lane), O[kk], O[kk+1]
    const double x = PAR ? wp_shfl_down(m.x) : wp_shfl_up(m.y);
    n0 = PAR ? m.x : x; n1 = PAR ? m.y : m.x; n2 = PAR ? x : m.y;
}

// One step of the strip.  Window index j holds row (t - j) once the new row has come in.
// ro: element offset of row t.  FULL: every part of the step is active (steady state of a band):
// no range checks.  Row e = t-2K-2 of the epilogue has the parity of t, so it is even iff PAR == 1.
// PRE / POSTK: the pass flavour when known at compile time (-1: read p.pre / p.post).
template <int ARITH, int K, int PAR, bool FULL, int PRE, int POSTK>
WP_FN void step(const Params& p, const Strip& s, const Lane& ln, long t, int ro, V2 (&wE)[2 * K + 3], V2 (&wO)[2 * K + 3], double& acc)
{
    constexpr int NW = 2 * K + 3;
    const int post = POSTK >= 0 ? POSTK : p.post;
    const bool pre = PRE >= 0 ? (PRE != 0) : (p.pre != 0);
    // ---- epilogue of row e = t-2K-2: indices are still relative to t-1 here (row t-1-j at j)
    {
        const long e = t - 2 * K - 2;
        constexpr int c = 2 * K + 1;
        const int eo = ro - (2 * K + 2) * ln.pitch;
        if (post != POST_NONE && (FULL || (e >= s.elo && e <= s.ehi))) {
            // Sum of squares after smoothing (K > 0): the last half-sweep's colour ((i+j) odd) is summed
            // by the last stage itself, the epilogue visits the other colour only: the even columns of an
            // even row (PAR == 1), the odd columns of an odd row.  Injection: even columns of even rows.
            if ((post == POST_NORM2 && K == 0) || PAR == 1) {
                // even columns
                double n0, n1, n2;
                side<0>(wO[c], n0, n1, n2);
                const V2 f = ld_at(ln.rhs, eo), a = ld_at(ln.v1, eo), b = ld_at(ln.v2, eo);
                const Coef4 c0 = Arith<ARITH>::coef(a.x, b.x, p.st), c1 = Arith<ARITH>::coef(a.y, b.y, p.st);
                const double r0 = Arith<ARITH>::residual(f.x, wE[c].x, wE[c + 1].x, n0, wE[c - 1].x, n1, c0, p.st);
                const double r1 = Arith<ARITH>::residual(f.y, wE[c].y, wE[c + 1].y, n1, wE[c - 1].y, n2, c1, p.st);
                if (post == POST_INJECT) {
                    // coarse node (e/2, kk) for even column 2kk: kk = k is even (coarse even run), k+1 odd
                    double* row = p.crhs + ((e >> 1) - p.crow0) * p.cpitch;
                    if (s.oint[0] & 1u) row[s.k >> 1] = r0;                                   // gs.cpp:283
                    if (s.oint[0] & 2u) row[p.codd + (s.k >> 1)] = r1;
                } else {
                    if (s.oint[0] & 1u) acc += r0 * r0;
                    if (s.oint[0] & 2u) acc += r1 * r1;
                }
            }
            if (post == POST_NORM2 && (K == 0 || PAR == 0)) {
                double n0, n1, n2;
                side<1>(wE[c], n0, n1, n2);
                const V2 f = ld_at(ln.rhs, eo + ln.odd), a = ld_at(ln.v1, eo + ln.odd), b = ld_at(ln.v2, eo + ln.odd);
                const Coef4 c0 = Arith<ARITH>::coef(a.x, b.x, p.st), c1 = Arith<ARITH>::coef(a.y, b.y, p.st);
                const double r0 = Arith<ARITH>::residual(f.x, wO[c].x, wO[c + 1].x, n0, wO[c - 1].x, n1, c0, p.st);
                const double r1 = Arith<ARITH>::residual(f.y, wO[c].y, wO[c + 1].y, n1, wO[c - 1].y, n2, c1, p.st);
                if (s.oint[1] & 1u) acc += r0 * r0;
                if (s.oint[1] & 2u) acc += r1 * r1;
            }
        }
    }
    // ---- the window moves down one row; row t comes in (+ prolongation and correction)
#pragma unroll
    for (int j = NW - 1; j >= 1; --j) { wE[j] = wE[j - 1]; wO[j] = wO[j - 1]; }
    if (FULL || t <= s.R1) {
        if (p.u_in) { wE[0] = ld_at(ln.u, ro); wO[0] = ld_at(ln.u, ro + ln.odd); }
        else { wE[0] = V2{0.0, 0.0}; wO[0] = V2{0.0, 0.0}; }
        if (pre) prolong_row(p, s, ln, t, wE[0], wO[0]);
    }
    // ---- all half-sweeps in order: stage q on row t-1-q = window index 1+q, all on column parity PAR
#pragma unroll
    for (int q = 0; q < 2 * K; ++q) {
        const long i = t - 1 - q;
        const int a = 1 + q;
        if (FULL || (i >= s.ulo && i <= s.uhi)) {                             // warp-uniform
            V2& tgt = PAR ? wO[a] : wE[a];
            const V2 up = PAR ? wO[a + 1] : wE[a + 1], dn = PAR ? wO[a - 1] : wE[a - 1];
            double n0, n1, n2;
            side<PAR>(PAR ? wE[a] : wO[a], n0, n1, n2);
            const int so = ro - (1 + q) * ln.pitch + (PAR ? ln.odd : 0);
            const V2 f = ld_at(ln.rhs, so), va = ld_at(ln.v1, so), vb = ld_at(ln.v2, so);
            const Coef4 c0 = Arith<ARITH>::coef(va.x, vb.x, p.st), c1 = Arith<ARITH>::coef(va.y, vb.y, p.st);
            const double o0 = Arith<ARITH>::gs(f.x, up.x, n0, dn.x, n1, c0, p.st);
            const double o1 = Arith<ARITH>::gs(f.y, up.y, n1, dn.y, n2, c1, p.st);
            if (s.ok[PAR] & 1u) tgt.x = o0;
            if (s.ok[PAR] & 2u) tgt.y = o1;
            if (q == 2 * K - 1 && post == POST_NORM2 && i >= s.elo && i <= s.ehi) {
                // residual of the nodes just updated: all operands are in registers (owned nodes are never masked)
                const double r0 = Arith<ARITH>::residual(f.x, o0, up.x, n0, dn.x, n1, c0, p.st);
                const double r1 = Arith<ARITH>::residual(f.y, o1, up.y, n1, dn.y, n2, c1, p.st);
                if (s.oint[PAR] & 1u) acc += r0 * r0;
                if (s.oint[PAR] & 2u) acc += r1 * r1;
            }
        }
    }
    // ---- row t-2K-1 (window index 2K+1) is final: store the owned part
    {
        const long f = t - 2 * K - 1;
        if ((FULL || (f >= s.b0 && f <= s.b1)) && (K > 0 || pre)) {
            double* row = ln.out + (ro - (2 * K + 1) * ln.pitch);
            if (s.own[0]) st_masked(row, wE[2 * K + 1], s.own[0]);
            if (s.own[1]) st_masked(row + ln.odd, wO[2 * K + 1], s.own[1]);
        }
    }
}

template <int ARITH, int K, int PRE = -1, int POSTK = -1>
WP_FN void run_strip_k(const Params& p, int strip, int band, int tile)
{
    const int lane = wp_lane();
    const Strip s = make_strip(p, strip, band, lane);
    const long ks = s.live ? s.k : 0;
    Lane ln;
    ln.u = p.u_in ? p.u_in + ks : nullptr; ln.rhs = p.rhs + ks; ln.v1 = p.v1 + ks; ln.v2 = p.v2 + ks; ln.out = p.u_out + ks;
    ln.cu = p.cu ? p.cu + (ks >> 1) : nullptr;
    ln.pitch = (int)p.pitch; ln.odd = (int)p.odd; ln.cpitch = (int)p.cpitch; ln.codd = (int)p.codd;
    V2 wE[2 * K + 3], wO[2 * K + 3];
#pragma unroll
    for (int j = 0; j < 2 * K + 3; ++j) { wE[j] = V2{0.0, 0.0}; wO[j] = V2{0.0, 0.0}; }
    double acc = 0.0;
    // steps R0 .. R1+2K+2.  Steady state (every stage, the store and the epilogue active, a row to
    // load): tf0 .. tf1, run without range checks.  The column parity (t-1)&1 picks the code.
    const long tend = s.R1 + 2 * K + 2;
    long tf0 = s.ulo + 2 * K, tf1 = s.uhi + 1;
    if (s.elo + 2 * K + 2 > tf0) tf0 = s.elo + 2 * K + 2;
    if (s.b0 + 2 * K + 1 > tf0) tf0 = s.b0 + 2 * K + 1;
    if (s.ehi + 2 * K + 2 < tf1) tf1 = s.ehi + 2 * K + 2;
    if (s.b1 + 2 * K + 1 < tf1) tf1 = s.b1 + 2 * K + 1;
    if (s.R1 < tf1) tf1 = s.R1;
    int ro = (int)((s.R0 - p.row0) * p.pitch);
    for (long t = s.R0; t <= tend; ++t, ro += ln.pitch) {
        const bool full = t >= tf0 && t <= tf1;
        if ((t - 1) & 1) {
            if (full) step<ARITH, K, 1, true, PRE, POSTK>(p, s, ln, t, ro, wE, wO, acc);
            else step<ARITH, K, 1, false, PRE, POSTK>(p, s, ln, t, ro, wE, wO, acc);
        } else {
            if (full) step<ARITH, K, 0, true, PRE, POSTK>(p, s, ln, t, ro, wE, wO, acc);
            else step<ARITH, K, 0, false, PRE, POSTK>(p, s, ln, t, ro, wE, wO, acc);
        }
    }
    if ((POSTK >= 0 ? POSTK : p.post) == POST_NORM2) {
        const double tot = wp_warp_sum(acc);
        if (lane == 0) p.partials[tile] = tot;
    }
}

// generic entry: K and the flavour read from the parameters (passes with K < KMAX)
template <int ARITH>
WP_FN void run_strip(const Params& p, int tile)
{
    const int strip = tile % p.nstrips, band = tile / p.nstrips;
    switch (p.K) {
        case 0: run_strip_k<ARITH, 0>(p, strip, band, tile); break;
        case 1: run_strip_k<ARITH, 1>(p, strip, band, tile); break;
        case 2: run_strip_k<ARITH, 2>(p, strip, band, tile); break;
        default: run_strip_k<ARITH, 3>(p, strip, band, tile); break;
    }
}

// K = KMAX with the flavour fixed at compile time: the passes that matter (one kernel image each)
template <int ARITH, int PRE, int POSTK>
WP_FN void run_strip_flavour(const Params& p, int tile)
{
    run_strip_k<ARITH, KMAX, PRE, POSTK>(p, tile % p.nstrips, tile / p.nstrips, tile);
}

// geometry: strips of OWN pairs over pairs 0..nhalf, bands of RB rows over the nrows produced rows
inline void plan(long n, long nrows, long rows_per_band, int& nstrips, int& nbands, long& RB)
{
    const long npairs = n / 2 + 1;
    nstrips = (int)((npairs + OWN - 1) / OWN);
    RB = rows_per_band < 1 ? 1 : rows_per_band;
    nbands = (int)((nrows + RB - 1) / RB);
}

}  // namespace wasp
}  // namespace mgb200
