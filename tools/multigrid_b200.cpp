// multigrid_b200 -- command-line front end equivalent to the reference's `multigrid` / `multigrid_cu`
// executables (multigrid.cpp:188-293, multigrid.cu:210-273), whose parameters are hard-coded
// literals there and flags here.  Same initial conditions, same solver parameters, same output
// format: uT.txt with one "%d\t%d\t%f\n" line per node, row-major (multigrid.cpp:269-275), plus an
// optional raw float64 dump.  Host side only: every compute call goes through the C ABI.
//
//   multigrid_b200 [--N 256] [--steps 100] [--nu -4e-4] [--vscale 1] [--tol 1e-6] [--shape 1]
//                  [--niter 3] [--exact] [--unfused] [--out uT.txt] [--bin uT.f64] [--quiet]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/mgb200.h"

static void die(const char* what)
{
    std::fprintf(stderr, "multigrid_b200: %s: %s\n", what, mgb200_last_error());
    std::exit(1);
}

int main(int argc, char** argv)
{
    long N = 256;                       // multigrid.cpp:192
    int steps = 100, shape = 1;         // :239, :241
    double nu = -4e-4, vscale = 1.0, tol = 1e-6;   // :235, :240
    std::string out = "uT.txt", bin;
    bool quiet = false;
    mgb200_options opt;
    mgb200_default_options(&opt);
    for (int a = 1; a < argc; ++a) {
        const std::string k = argv[a];
        auto val = [&]() -> const char* { if (a + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", k.c_str()); std::exit(2); } return argv[++a]; };
        if (k == "--N") N = std::atol(val());
        else if (k == "--steps") steps = std::atoi(val());
        else if (k == "--nu") nu = std::atof(val());
        else if (k == "--vscale") vscale = std::atof(val());
        else if (k == "--tol") tol = std::atof(val());
        else if (k == "--shape") shape = std::atoi(val());
        else if (k == "--niter") opt.niter = std::atoi(val());
        else if (k == "--exact") opt.arith = MGB200_ARITH_EXACT;
        else if (k == "--unfused") opt.plan = MGB200_PLAN_UNFUSED;
        else if (k == "--out") out = val();
        else if (k == "--bin") bin = val();
        else if (k == "--quiet") quiet = true;
        else { std::fprintf(stderr, "unknown flag %s\n", k.c_str()); return 2; }
    }
    opt.shape = shape;
    const int maxlvl = int(std::log2((double)N)) - 4;      // multigrid.cpp:193
    const double dx = 1.0 / N, dt = dx / 10;               // :194, :238
    mgb200_solver* s = nullptr;
    if (mgb200_create(&s, N, maxlvl, nu, dt, dx, tol, &opt) != MGB200_OK) die("create");
    if (mgb200_set_fields_reference_ic(s, vscale) != MGB200_OK) die("initial conditions");
    std::vector<mgb200_solve_info> infos(steps > 0 ? steps : 1);
    const auto t0 = std::chrono::steady_clock::now();
    if (mgb200_timestep(s, steps, infos.data()) != MGB200_OK) die("timestep");
    if (mgb200_synchronize(s) != MGB200_OK) die("synchronize");
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long cycles = 0;
    for (int k = 0; k < steps; ++k) cycles += infos[k].cycles;
    std::vector<double> uT((size_t)(N + 1) * (N + 1));
    if (mgb200_get_u_host(s, uT.data()) != MGB200_OK) die("get_u");
    if (!quiet)
        std::printf("\nB200 time, N = %ld: %f s  (%d steps, %ld cycles, last ||r||/||r0|| = %.3e)\n%g\n", N, sec, steps, cycles,
                    steps > 0 ? infos[steps - 1].res / infos[steps - 1].res0 : 0.0, uT[(N / 2) * (N + 1) + N / 2]);
    if (!out.empty() && out != "-") {
        FILE* f = std::fopen(out.c_str(), "w");
        if (!f) { std::perror(out.c_str()); return 1; }
        for (long i = 0; i <= N; ++i)
            for (long j = 0; j <= N; ++j) std::fprintf(f, "%ld\t%ld\t%f\n", i, j, uT[i * (N + 1) + j]);   // multigrid.cpp:272
        std::fclose(f);
    }
    if (!bin.empty()) {
        FILE* f = std::fopen(bin.c_str(), "wb");
        if (!f) { std::perror(bin.c_str()); return 1; }
        std::fwrite(uT.data(), sizeof(double), uT.size(), f);
        std::fclose(f);
    }
    mgb200_destroy(s);
    return 0;
}
