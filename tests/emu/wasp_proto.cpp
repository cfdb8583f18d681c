// wasp_proto.cpp -- CPU prototype of the "warp-autonomous" streaming pass sketched in DESIGN.md
// section 9 (round-2 candidate; NOT part of the product, no CUDA here).  It pins down the index
// logic that a CUDA version has to reproduce and is checked bit for bit against the oracle by
// tests/test_wasp_prototype.py:
//
//   * one "warp" owns a strip of WP column pairs (HP halo pairs per side, recomputed) and a band of
//     rows (2K+1 halo rows per side, recomputed) and keeps a WINDOW of 2K+3 rows of u privately
//     (registers in the CUDA version: here a small array per strip);
//   * per step t it takes row t in, then runs ALL 2K half-sweeps itself, IN ORDER, stage s on row
//     t-1-s.  Executed in order by one owner the stages need a lag of ONE row: stage s-1 has just
//     finished row t-s (= row+1) in this very step, finished row and row-1 in the two steps before,
//     and stage s+1 touches row-1 only later in this step;
//   * row t-2K-1 is final after step t and is stored; the epilogue (residual -> injection / sum of
//     squares) of a row needs the final row below it, so row t-2K-2 is handled at the START of
//     step t, just before row t takes the window slot of row t-2K-3;
//   * no shared iterate, no barrier between the strips: strips only read u_in and write u_out.
//
// Arithmetic: the reference's expressions (gs.cpp:15,19,44,75,130,238-241) evaluated as written
// (compile with -ffp-contract=off), so the comparison with the oracle is exact.
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Ctx {
    long n, ld;
    double nu, dx, r, diag;
    const double *rhs, *v1, *v2;
};

inline double cminus(double v, const Ctx& c) { return c.r * (-v * c.dx / 2.0 + c.nu); }   // gs.cpp:15
inline double cplus(double v, const Ctx& c) { return c.r * (v * c.dx / 2.0 + c.nu); }     // gs.cpp:19

// window row storage: columns [c0, c1] of one row
struct Row {
    std::vector<double> v;
    long row = -1000000;
};

}  // namespace

extern "C" {

// One pass over level n (natural layout, ld = n+1): [u += P(cu)] ; K RB-GS iterations ;
// [post = 1: crhs = injected residual | post = 2: *sumsq = sum of squared residuals].
// wp: owned pairs per strip, hp: halo pairs per side, rb: owned rows per band.
// Returns 0, or a negative code if the geometry is unusable.
int wasp_pass(long n, const double* u_in, double* u_out, const double* rhs, const double* v1, const double* v2,
              const double* cu, double* crhs, double* sumsq, int K, int post, double dt, double nu, double dx,
              int wp, int hp, long rb)
{
    if (K < 0 || wp < 1 || rb < 1 || 2 * hp < 2 * K + 1) return -1;     // halo columns must cover 2K half-sweeps (+1 for the epilogue)
    Ctx c{n, n + 1, nu, dx, 0.5 * dt / (dx * dx), 0.0, rhs, v1, v2};
    c.diag = 1.0 - 4.0 * c.r * nu;
    const long ld = c.ld, nc = n / 2, ldc = nc + 1;
    const int NW = 2 * K + 3;                                           // window rows
    const long hrow = 2 * K + 1;                                        // halo rows per band side
    double total = 0.0;
    if (u_out != u_in) {                                                // boundary lines are never produced by a strip
        for (long j = 0; j <= n; ++j) { u_out[j] = u_in[j]; u_out[n * ld + j] = u_in[n * ld + j]; }
        for (long i = 0; i <= n; ++i) { u_out[i * ld] = u_in[i * ld]; u_out[i * ld + n] = u_in[i * ld + n]; }
    }
    for (long b0 = 1; b0 <= n - 1; b0 += rb) {                          // bands of owned interior rows
        const long b1 = b0 + rb - 1 < n - 1 ? b0 + rb - 1 : n - 1;
        const long R0 = b0 - hrow < 0 ? 0 : b0 - hrow, R1 = b1 + hrow > n ? n : b1 + hrow;
        for (long p0 = 0; 2 * p0 <= n; p0 += wp) {                      // strips of owned pairs (columns 2p, 2p+1)
            const long o0 = 2 * p0, o1 = (2 * (p0 + wp) - 1 < n ? 2 * (p0 + wp) - 1 : n);     // owned columns
            const long c0 = o0 - 2 * hp < 0 ? 0 : o0 - 2 * hp, c1 = o1 + 2 * hp > n ? n : o1 + 2 * hp;
            const long w = c1 - c0 + 1;
            std::vector<Row> win(NW);
            for (auto& r : win) r.v.assign(w, std::nan(""));
            auto slot = [&](long row) -> Row& { return win[((row % NW) + NW) % NW]; };
            auto at = [&](long row, long col) -> double& {
                Row& r = slot(row);
                return r.v[col - c0];
            };
            auto held = [&](long row) { return slot(row).row == row; };
            for (long t = R0; t <= R1 + 2 * K + 2; ++t) {
                // ---- epilogue on row t-2K-2: the row below it became final in the previous step, and the row
                // above it is still in the window (its slot is the one row t is about to take)
                const long e = t - 2 * K - 2;
                if (post && e >= b0 && e <= b1) {
                    if (!held(e - 1) || !held(e) || !held(e + 1)) return -4;
                    for (long j = (o0 < 1 ? 1 : o0); j <= (o1 > n - 1 ? n - 1 : o1); ++j) {
                        if (post == 1 && ((e & 1) || (j & 1))) continue;              // injection reads even rows / even columns only
                        const long q = e * ld + j;
                        const double west = cminus(v2[q], c), east = cplus(v2[q], c), north = cminus(v1[q], c), south = cplus(v1[q], c);
                        const double rv = rhs[q] - (c.diag * at(e, j) + north * at(e - 1, j) + west * at(e, j - 1) + south * at(e + 1, j) + east * at(e, j + 1));   // gs.cpp:75
                        if (post == 1) crhs[(e >> 1) * ldc + (j >> 1)] = rv;          // gs.cpp:283 on the coarse interior
                        else total += rv * rv;
                    }
                }
                // ---- row t enters the window (+ prolongation and correction, multigrid.cpp:81-83)
                if (t <= R1) {
                    Row& r = slot(t);
                    r.row = t;
                    for (long j = c0; j <= c1; ++j) {
                        double x = u_in ? u_in[t * ld + j] : 0.0;
                        if (cu && t >= 1 && t <= n - 1 && j >= 1 && j <= n - 1) {
                            const long I = t >> 1, J = j >> 1;
                            double pv;
                            if (!(t & 1) && !(j & 1)) pv = cu[I * ldc + J];                                        // gs.cpp:238
                            else if ((t & 1) && !(j & 1)) pv = (cu[I * ldc + J] + cu[(I + 1) * ldc + J]) / 2;        // :239
                            else if (!(t & 1) && (j & 1)) pv = (cu[I * ldc + J] + cu[I * ldc + J + 1]) / 2;          // :240
                            else pv = (cu[I * ldc + J] + cu[(I + 1) * ldc + J] + cu[I * ldc + J + 1] + cu[(I + 1) * ldc + J + 1]) / 4;   // :241
                            x = x + pv;
                        }
                        r.v[j - c0] = x;
                    }
                }
                // ---- all half-sweeps, in order: stage s on row t-1-s
                for (int s = 0; s < 2 * K; ++s) {
                    const long i = t - 1 - s;
                    if (i <= R0 || i >= R1 || i < 1 || i > n - 1) continue;          // rows i-1, i+1 must be held; interior only
                    if (!held(i - 1) || !held(i) || !held(i + 1)) return -2;         // the window is deep enough (checked, not assumed)
                    const int colour = s & 1;
                    for (long j = c0 + 1; j <= c1 - 1; ++j) {                        // columns j-1, j+1 must be held
                        if (j < 1 || j > n - 1 || ((i + j) & 1) != colour) continue;
                        const long q = i * ld + j;
                        const double west = cminus(v2[q], c), east = cplus(v2[q], c), north = cminus(v1[q], c), south = cplus(v1[q], c);
                        at(i, j) = (rhs[q] - north * at(i - 1, j) - west * at(i, j - 1) - south * at(i + 1, j) - east * at(i, j + 1)) / c.diag;   // gs.cpp:130
                    }
                }
                // ---- row t-2K-1 is final: store its owned part
                const long f = t - 2 * K - 1;
                if (f >= b0 && f <= b1) {
                    if (!held(f)) return -3;
                    for (long j = (o0 < 1 ? 1 : o0); j <= (o1 > n - 1 ? n - 1 : o1); ++j) u_out[f * ld + j] = at(f, j);
                }
            }
        }
    }
    if (post == 2 && sumsq) *sumsq = total;
    return 0;
}

}  // extern "C"
