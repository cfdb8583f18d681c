// ref_fresh_main.cpp -- test infrastructure.  A minimal fresh-process driver around the
// UNMODIFIED reference timestepper (multigrid.cpp:124), linked WITHOUT ref_prelude.h.
// It allocates u0/v1/v2/uT exactly like the reference main (multigrid.cpp:198-203), calls
// timestepper once and dumps uT as raw doubles.  tests/test_oracle_vs_ref.py uses it to check
// that the zero-filled towers of the prelude build are what a real reference process computes.
//   usage: ref_fresh N steps nu vscale tol shape out.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "mg_oracle.h"

void timestepper(double* uT, double* u0, double* v1, double* v2, double nu, int maxlvl, int n,
                 double dt, double T, double dx, double tol, int shape);

int main(int argc, char** argv)
{
    if (argc != 8) { std::fprintf(stderr, "usage: %s N steps nu vscale tol shape out.bin\n", argv[0]); return 2; }
    const int N = std::atoi(argv[1]), steps = std::atoi(argv[2]), shape = std::atoi(argv[6]);
    const double nu = std::atof(argv[3]), vscale = std::atof(argv[4]), tol = std::atof(argv[5]);
    const int maxlvl = int(std::log2(N)) - 4;
    const double dx = 1.0 / N, dt = dx / 10;
    const size_t m = size_t(N + 1) * (N + 1);
    double* uT = (double*)std::malloc(sizeof(double) * m);
    double* u0 = (double*)std::malloc(sizeof(double) * m);
    double* v1 = (double*)std::malloc(sizeof(double) * m);
    double* v2 = (double*)std::malloc(sizeof(double) * m);
    orc_initial_conditions(u0, v1, v2, N, vscale);
    timestepper(uT, u0, v1, v2, nu, maxlvl, N, dt, steps * dt, dx, tol, shape);
    FILE* f = std::fopen(argv[7], "wb");
    std::fwrite(uT, sizeof(double), m, f);
    std::fclose(f);
    return 0;
}
