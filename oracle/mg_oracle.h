/*
 * mg_oracle.h -- CPU oracle for the geometric-multigrid hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library under
 * hpcclassmultigridproject_b200/csrc) may link, call or load this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and only as the checker.
 *
 * It is a plain-C restatement of the reference's serial CPU algorithm
 * (/root/reference/gs.cpp + multigrid.cpp).  Parity status: PINNED -- every
 * function here is checked bit-for-bit against the compiled, unmodified reference
 * (oracle/_ref/libmgref_O0.so, built by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, against golden vectors generated from that build
 * (tests/golden/, generator tests/golden/make_golden.py) and against the
 * known-answer material of the reference's own print tests
 * (prolrestest.cpp:80-118, resnormtest.cpp:222-283).
 *
 * All fields are dense row-major FP64, (n+1) x (n+1) nodes, idx = i*(n+1)+j.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- grid operators (reference gs.h:3-17) ------------------------------- */
void   orc_compute_rhs (double *rhs, const double *u, long n, const double *v1,
                        const double *v2, double dt, double nu, double dx);
void   orc_residual    (double *res, const double *u, const double *rhs, long n,
                        const double *v1, const double *v2, double dt, double nu, double dx);
double orc_norm        (const double *res, long n);
void   orc_gauss_seidel(double *u, const double *rhs, long n, const double *v1,
                        const double *v2, double dt, double nu, double dx);
/* out is (2nc+1)^2, in is (nc+1)^2 */
void   orc_prolongation(double *fine, const double *coarse, long nc);
/* out is (nf/2+1)^2, in is (nf+1)^2 */
void   orc_restriction (double *coarse, const double *fine, long nf);
void   orc_restriction_fw(double *coarse, const double *fine, long nf);   /* gs.cpp:277-280 (commented out there) */

/* ---- driver (reference multigrid.cpp:17-186) ----------------------------- */
typedef struct orc_solver orc_solver;

/* Allocates the level towers exactly as timestepper does (multigrid.cpp:131-162),
 * including the coarse-velocity tower quirk on
 * zero-filled buffers.  u0/v1/v2 are copied. */
orc_solver *orc_create(long n, int maxlvl, const double *u0, const double *v1,
                       const double *v2, double nu, double dt, double dx,
                       double tol, int shape);
void        orc_destroy(orc_solver *s);

/* one V/W cycle from level `lvl` downwards (mg_inner, multigrid.cpp:17-92) */
void orc_cycle(orc_solver *s, int lvl);
/* rhs_0 = B u_0 (compute_rhs call at multigrid.cpp:167) */
void orc_form_rhs(orc_solver *s);
/* residual norm of level 0 (multigrid.cpp:104-105 / 112-113) */
double orc_residual_norm(orc_solver *s);
/* mg_outer (multigrid.cpp:97-120).  Returns the number of cycles; if hist != NULL
 * it receives res0 followed by the norm after each cycle (<= 51 doubles). */
int  orc_solve(orc_solver *s, double *hist);
/* nsteps Crank-Nicolson steps; cycles[k] (may be NULL) = cycles used by step k */
void orc_advance(orc_solver *s, int nsteps, int *cycles);
/* access to the level arrays (for tests) */
double *orc_level_u  (orc_solver *s, int lvl);
double *orc_level_rhs(orc_solver *s, int lvl);
double *orc_level_v1 (orc_solver *s, int lvl);
double *orc_level_v2 (orc_solver *s, int lvl);
double *orc_tmp      (orc_solver *s);

/* timestepper (multigrid.cpp:124-186): T/dt steps, result copied into uT */
void orc_timestepper(double *uT, const double *u0, const double *v1, const double *v2,
                     double nu, int maxlvl, int n, double dt, double T, double dx,
                     double tol, int shape);

/* initial conditions of the reference main (multigrid.cpp:206-233), with the
 * velocity multiplied by vscale */
void orc_initial_conditions(double *u0, double *v1, double *v2, long n, double vscale);

#ifdef __cplusplus
}
#endif
#endif
