// syst_tsan_main.cpp -- TEST INFRASTRUCTURE: driver of the ThreadSanitizer build of the host model
// (syst_emu.cpp).  Runs the pass flavours on seeded random fields; the sanitizer reports any two
// conflicting shared-memory accesses that the mbarrier protocol of csrc/syst_pass_body.cuh leaves
// unordered.  Results are not checked here (tests/test_syst_pass_emu.py checks them bit for bit with
// the plain build of the same model).
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" long syst_emu_run(long n, long pitch, long odd, long cpitch, long codd, const double* u_in, double* u_out,
                             const double* rhs, const double* v1, const double* v2, const double* cu, double* crhs,
                             double* partials, int K, int post, int arith, double dt, double nu, double dx, int wk, int nbands,
                             int jitter_level, long own_lo, long own_hi, long row0, long rows_mem, long crow0, long crows_mem);

static long lay_odd(long n) { return (n / 2 + 1 + 15) / 16 * 16 + 32; }

static std::vector<double> field(long rows, long pitch, unsigned seed, bool zero_edge = false, long n = 0, long odd = 0)
{
    std::vector<double> a((size_t)rows * pitch);
    unsigned r = seed * 2654435761u + 12345u;
    for (auto& x : a) { r = r * 1664525u + 1013904223u; x = (double)(r >> 8) / (1 << 24) - 0.5; }
    if (zero_edge) {                                       // a coarse correction has a zero boundary
        for (long j = 0; j < pitch; ++j) { a[j] = 0.0; a[(size_t)(rows - 1) * pitch + j] = 0.0; }
        for (long i = 0; i < rows; ++i) { a[(size_t)i * pitch] = 0.0; a[(size_t)i * pitch + n / 2] = 0.0; (void)odd; }
    }
    return a;
}

int main()
{
    struct Case { long n; int K, post, pre, wk, nbands, jitter; };
    const Case cases[] = {
        {64, 3, 1, 0, 24, 2, 4},      // down leg, one warp per stage
        {64, 3, 2, 1, 24, 1, 8},      // up leg with the norm
        {96, 1, 1, 0, 0, 0, 2},       // K = 1: stage 0 feeds the last stage directly
        {96, 2, 2, 1, 56, 2, 4},      // K = 2
        {288, 3, 1, 1, 120, 2, 4},    // two warps per stage: strip-half edges, prolongation + injection
        {288, 3, 2, 1, 120, 1, 0},    // ... with the norm
        {288, 3, 0, 0, 88, 3, 8},     // a partially filled second half
        {512, 3, 1, 1, 120, 4, 2},    // interior strips: the mask-free step
    };
    for (const Case& c : cases) {
        const long n = c.n, odd = lay_odd(n), pitch = 2 * odd, codd = lay_odd(n / 2), cpitch = 2 * codd;
        auto u = field(n + 1, pitch, 1), rhs = field(n + 1, pitch, 2), v1 = field(n + 1, pitch, 3), v2 = field(n + 1, pitch, 4);
        auto cu = field(n / 2 + 1, cpitch, 5, true, n / 2, codd);
        std::vector<double> out((size_t)(n + 1) * pitch, 0.0), crhs((size_t)(n / 2 + 1) * cpitch, 0.0), parts(4096, 0.0);
        const double dx = 1.0 / n, dt = dx / 10;
        const long nt = syst_emu_run(n, pitch, odd, cpitch, codd, u.data(), out.data(), rhs.data(), v1.data(), v2.data(),
                                     c.pre ? cu.data() : nullptr, crhs.data(), parts.data(), c.K, c.post, 1, dt, -4e-4, dx, c.wk,
                                     c.nbands, c.jitter, 0, 0, 0, 0, 0, 0);
        std::printf("n=%ld K=%d post=%d pre=%d wk=%d bands=%d: %ld tiles\n", n, c.K, c.post, c.pre, c.wk, c.nbands, nt);
        std::fflush(stdout);
        if (nt <= 0) return 2;
    }
    std::printf("all passes done\n");
    return 0;
}
