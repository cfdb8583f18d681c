/*
 * mg_oracle.c -- CPU oracle (test infrastructure, see mg_oracle.h).
 *
 * Restates, in plain C, the arithmetic of the reference's serial path.  The
 * expression ORDER of every floating-point formula is the reference's, because
 * the parity bar for the field u is bit-for-bit; compile with
 * -ffp-contract=off (no FMA) -- see oracle/Makefile.
 *
 * Reference citations are file:line into /root/reference.
 */
#include "mg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* stencil coefficients: gs.cpp:9-20 (r, a, b) as used at gs.cpp:40-43         */
typedef struct {
    double west;   /* "aa": multiplies u[i][j-1], from v2 */
    double east;   /* "bb": multiplies u[i][j+1], from v2 */
    double north;  /* "cc": multiplies u[i-1][j], from v1 */
    double south;  /* "dd": multiplies u[i+1][j], from v1 */
} stencil_t;

static inline double cn_ratio(double dx, double dt)
{
    return 0.5 * dt / (dx * dx);                 /* gs.cpp:10 */
}

static inline double coef_minus(double v, double nu, double dx, double r)
{
    return r * (-v * dx / 2.0 + nu);             /* gs.cpp:15 */
}

static inline double coef_plus(double v, double nu, double dx, double r)
{
    return r * (v * dx / 2.0 + nu);              /* gs.cpp:19 */
}

static inline stencil_t stencil_at(const double *v1, const double *v2, long p,
                                   double nu, double dx, double r)
{
    stencil_t s;
    s.west  = coef_minus(v2[p], nu, dx, r);      /* gs.cpp:40 */
    s.east  = coef_plus (v2[p], nu, dx, r);      /* gs.cpp:41 */
    s.north = coef_minus(v1[p], nu, dx, r);      /* gs.cpp:42 */
    s.south = coef_plus (v1[p], nu, dx, r);      /* gs.cpp:43 */
    return s;
}

/* ------------------------------------------------------------------------- */
/* rhs = B u on the interior (gs.cpp:24-53, formula :44)                       */
void orc_compute_rhs(double *rhs, const double *u, long n, const double *v1,
                     const double *v2, double dt, double nu, double dx)
{
    const long ld = n + 1;
    const double r = cn_ratio(dx, dt);
    for (long i = 1; i < n; ++i) {
        for (long j = 1; j < n; ++j) {
            const long p = i * ld + j;
            const stencil_t s = stencil_at(v1, v2, p, nu, dx, r);
            rhs[p] = (1.0 + 4.0 * r * nu) * u[p] - s.north * u[p - ld] - s.west * u[p - 1]
                     - s.south * u[p + ld] - s.east * u[p + 1];
        }
    }
}

/* res = rhs - A u on the interior (gs.cpp:55-83, formula :75)                 */
void orc_residual(double *res, const double *u, const double *rhs, long n,
                  const double *v1, const double *v2, double dt, double nu, double dx)
{
    const long ld = n + 1;
    const double r = cn_ratio(dx, dt);
    for (long i = 1; i < n; ++i) {
        for (long j = 1; j < n; ++j) {
            const long p = i * ld + j;
            const stencil_t s = stencil_at(v1, v2, p, nu, dx, r);
            res[p] = rhs[p] - ((1.0 - 4.0 * r * nu) * u[p] + s.north * u[p - ld]
                               + s.west * u[p - 1] + s.south * u[p + ld] + s.east * u[p + 1]);
        }
    }
}

/* sqrt of the serial row-major sum of squares over the interior (gs.cpp:86-107) */
double orc_norm(const double *res, long n)
{
    const long ld = n + 1;
    double acc = 0.0;
    for (long i = 1; i < n; ++i)
        for (long j = 1; j < n; ++j)
            acc += res[i * ld + j] * res[i * ld + j];
    return sqrt(acc);
}

/* one colour of red-black Gauss-Seidel: nodes with (i+j) % 2 == colour.
 * Update formula gs.cpp:130.  Within a colour the order is irrelevant (a node
 * only reads the other colour), so the four row/column loops of gs.cpp:121-184
 * collapse into one loop per colour with identical results. */
static void rb_half_sweep(double *u, const double *rhs, long n, const double *v1,
                          const double *v2, double nu, double dx, double r, int colour)
{
    const long ld = n + 1;
    const double diag = 1.0 - 4.0 * r * nu;
    for (long i = 1; i < n; ++i) {
        long j0 = ((i + colour) & 1) ? 1 : 2;    /* first interior j with (i+j)%2==colour */
        for (long j = j0; j < n; j += 2) {
            const long p = i * ld + j;
            const stencil_t s = stencil_at(v1, v2, p, nu, dx, r);
            u[p] = (rhs[p] - s.north * u[p - ld] - s.west * u[p - 1] - s.south * u[p + ld]
                    - s.east * u[p + 1]) / diag;
        }
    }
}

/* one RB-GS iteration: "red" = (i+j) even first (gs.cpp:121-151: (odd,odd) and
 * (even,even) nodes), then "black" = (i+j) odd (gs.cpp:156-184). */
void orc_gauss_seidel(double *u, const double *rhs, long n, const double *v1,
                      const double *v2, double dt, double nu, double dx)
{
    const double r = cn_ratio(dx, dt);
    rb_half_sweep(u, rhs, n, v1, v2, nu, dx, r, 0);
    rb_half_sweep(u, rhs, n, v1, v2, nu, dx, r, 1);
}

/* bilinear interpolation coarse (nc+1)^2 -> fine (2nc+1)^2, every fine node
 * written (gs.cpp:228-266).  Summation order of the 4-point average follows
 * gs.cpp:241: ((c00 + c10) + c01) + c11 with c10 = next ROW. */
void orc_prolongation(double *fine, const double *coarse, long nc)
{
    const long ldc = nc + 1, ldf = 2 * nc + 1;
    for (long I = 0; I <= nc; ++I) {
        for (long J = 0; J <= nc; ++J) {
            const double c00 = coarse[I * ldc + J];
            fine[2 * I * ldf + 2 * J] = c00;                                    /* :238,:254,:257,:265 */
            if (I < nc)
                fine[(2 * I + 1) * ldf + 2 * J] = (c00 + coarse[(I + 1) * ldc + J]) / 2;      /* :239,:255 */
            if (J < nc)
                fine[2 * I * ldf + 2 * J + 1] = (c00 + coarse[I * ldc + J + 1]) / 2;          /* :240,:258 */
            if (I < nc && J < nc)
                fine[(2 * I + 1) * ldf + 2 * J + 1] =
                    (c00 + coarse[(I + 1) * ldc + J] + coarse[I * ldc + J + 1]
                     + coarse[(I + 1) * ldc + J + 1]) / 4;                                     /* :241 */
        }
    }
}

/* injection fine (nf+1)^2 -> coarse (nf/2+1)^2, boundary included (gs.cpp:268-292,
 * active formula :283; the full-weighting lines :278-280 are commented out there). */
void orc_restriction(double *coarse, const double *fine, long nf)
{
    const long nc = nf / 2, ldc = nc + 1, ldf = nf + 1;
    for (long I = 0; I <= nc; ++I)
        for (long J = 0; J <= nc; ++J)
            coarse[I * ldc + J] = fine[2 * I * ldf + 2 * J];
}

/* Full-weighting restriction exactly as sketched in the reference's commented-out lines
 * gs.cpp:277-280 (never active there, "get this working if time permits"): interior coarse nodes
 * take the [1 2 1; 2 4 2; 1 2 1]/16 stencil, grouped by fine rows as written there; boundary nodes
 * (where the sketch would index out of bounds) are injected.  Opt-in only: the reference's active
 * path is orc_restriction. */
void orc_restriction_fw(double *coarse, const double *fine, long nf)
{
    const long nc = nf / 2, ldc = nc + 1, ldf = nf + 1;
    for (long I = 0; I <= nc; ++I)
        for (long J = 0; J <= nc; ++J) {
            if (I == 0 || J == 0 || I == nc || J == nc) { coarse[I * ldc + J] = fine[2 * I * ldf + 2 * J]; continue; }
            const double *a = fine + (2 * I - 1) * ldf + 2 * J, *b = a + ldf, *c = b + ldf;
            double t = (a[-1] + 2 * a[0] + a[1]) / 16;                                     /* :278 */
            t += (2 * b[-1] + 4 * b[0] + 2 * b[1]) / 16;                                   /* :279 */
            t += (c[-1] + 2 * c[0] + c[1]) / 16;                                           /* :280 */
            coarse[I * ldc + J] = t;
        }
}

/* ------------------------------------------------------------------------- */
#define ORC_MAXLVL 32
#define ORC_NITER 3            /* multigrid.cpp:41  */
#define ORC_COARSE_MAXIT 1000  /* multigrid.cpp:60  */
#define ORC_COARSE_TOL 1e-5    /* multigrid.cpp:60  */
#define ORC_MAX_CYCLE 50       /* multigrid.cpp:94  */

struct orc_solver {
    long n;
    int maxlvl, shape;
    int coarse_exact;          /* opt-in: direct solve of the coarsest level (orc_coarse_lu_solve) */
    double nu, dt, dx, tol;
    double *u[ORC_MAXLVL], *rhs[ORC_MAXLVL], *v1[ORC_MAXLVL], *v2[ORC_MAXLVL];
    double *tmp;
};

orc_solver *orc_create(long n, int maxlvl, const double *u0, const double *v1,
                       const double *v2, double nu, double dt, double dx,
                       double tol, int shape)
{
    if (maxlvl < 1 || maxlvl > ORC_MAXLVL) return NULL;
    orc_solver *s = (orc_solver *)calloc(1, sizeof *s);
    const size_t m0 = (size_t)(n + 1) * (size_t)(n + 1);
    s->n = n; s->maxlvl = maxlvl; s->shape = shape;
    s->nu = nu; s->dt = dt; s->dx = dx; s->tol = tol;
    /* level 0: private copies (multigrid.cpp:138-145) */
    s->u[0]   = (double *)calloc(m0, sizeof(double));
    s->rhs[0] = (double *)calloc(m0, sizeof(double));
    s->v1[0]  = (double *)calloc(m0, sizeof(double));
    s->v2[0]  = (double *)calloc(m0, sizeof(double));
    memcpy(s->u[0],  u0, m0 * sizeof(double));
    memcpy(s->v1[0], v1, m0 * sizeof(double));
    memcpy(s->v2[0], v2, m0 * sizeof(double));
    /* coarse levels (multigrid.cpp:148-160).  The reference never halves n in
     * this loop, so EVERY coarse array has (n/2+1)^2 entries and EVERY velocity
     * level is produced by restriction(dst, src, n/2), i.e. with strides
     * (n/4+1) <- (n/2+1) regardless of the level.  Only the first (n/4+1)^2
     * flat entries are written; the remainder is fresh-malloc memory, zero in
     * practice, zero here by construction. */
    const long nh = n >> 1;
    const size_t mh = (size_t)(nh + 1) * (size_t)(nh + 1);
    for (int l = 1; l < maxlvl; ++l) {
        s->u[l]   = (double *)calloc(mh, sizeof(double));
        s->rhs[l] = (double *)calloc(mh, sizeof(double));
        s->v1[l]  = (double *)calloc(mh, sizeof(double));
        s->v2[l]  = (double *)calloc(mh, sizeof(double));
        orc_restriction(s->v1[l], s->v1[l - 1], nh);   /* multigrid.cpp:155 */
        orc_restriction(s->v2[l], s->v2[l - 1], nh);   /* multigrid.cpp:157 */
    }
    s->tmp = (double *)calloc(m0, sizeof(double));      /* multigrid.cpp:162 */
    return s;
}

/* OPT-IN, no reference counterpart: replace the coarse velocity towers by TRUE injection,
 * v_l[i][j] = v_{l-1}[2i][2j] with each level's own stride (what multigrid.cpp:148-160 presumably meant:
 * the reference never halves n there, see above).  Pinned only to this formula -- which is plain
 * subsampling, v_l = v_0[::2^l, ::2^l], checked as such in the tests. */
void orc_correct_towers(orc_solver *s)
{
    for (int l = 1; l < s->maxlvl; ++l) {
        const long nf = s->n >> (l - 1);
        const size_t mh = (size_t)((s->n >> 1) + 1) * (size_t)((s->n >> 1) + 1);
        memset(s->v1[l], 0, mh * sizeof(double));
        memset(s->v2[l], 0, mh * sizeof(double));
        orc_restriction(s->v1[l], s->v1[l - 1], nf);
        orc_restriction(s->v2[l], s->v2[l - 1], nf);
    }
}

void orc_destroy(orc_solver *s)
{
    if (!s) return;
    for (int l = 0; l < s->maxlvl; ++l) {
        free(s->u[l]); free(s->rhs[l]); free(s->v1[l]); free(s->v2[l]);
    }
    free(s->tmp);
    free(s);
}

double *orc_level_u  (orc_solver *s, int l) { return s->u[l]; }
double *orc_level_rhs(orc_solver *s, int l) { return s->rhs[l]; }
double *orc_level_v1 (orc_solver *s, int l) { return s->v1[l]; }
double *orc_level_v2 (orc_solver *s, int l) { return s->v2[l]; }
double *orc_tmp      (orc_solver *s)        { return s->tmp; }

/* OPT-IN, no compiled-reference twin: exact_solve.cpp:1-55 of the reference (banded LU of the coarsest level through
 * LAPACKE dgbtrf/dgbtrs) was never finished -- it does not compile and nothing calls it.  This is the direct solve it
 * set out to do, written out so that the GPU kernel can be checked bit for bit: interior unknowns p = (i-1)(n-1)+(j-1),
 * half bandwidth w = n-1, band storage ab[p][w+q-p], LU WITHOUT pivoting (the matrix is strictly diagonally dominant:
 * partial pivoting would not exchange rows), k-outer elimination, then forward and back substitution.  Dirichlet
 * neighbours move to the right-hand side in the order of gs.cpp:130 (north, west, south, east). */
void orc_coarse_lu_solve(double *u, const double *f, long n, const double *v1, const double *v2,
                         double dt, double nu, double dx)
{
    const long ni = n - 1, m = ni * ni, w = ni, bw = 2 * w + 1, ld = n + 1;
    const double r = cn_ratio(dx, dt);
    const double diag = 1.0 - 4.0 * r * nu;                    /* gs.cpp:75,130 */
    double *ab = (double *)calloc((size_t)m * bw, sizeof(double));
    double *y = (double *)calloc((size_t)m, sizeof(double));
    for (long p = 0; p < m; ++p) {
        const long i = 1 + p / ni, j = 1 + p % ni, g = i * ld + j;
        const stencil_t c = stencil_at(v1, v2, g, nu, dx, r);
        double *row = ab + p * bw + w;
        row[0] = diag;
        if (i > 1) row[-ni] = c.north;
        if (i < n - 1) row[ni] = c.south;
        if (j > 1) row[-1] = c.west;
        if (j < n - 1) row[1] = c.east;
        double b = f[g];
        if (i == 1) b = b - c.north * u[(i - 1) * ld + j];
        if (j == 1) b = b - c.west * u[i * ld + j - 1];
        if (i == n - 1) b = b - c.south * u[(i + 1) * ld + j];
        if (j == n - 1) b = b - c.east * u[i * ld + j + 1];
        y[p] = b;
    }
    for (long k = 0; k < m - 1; ++k) {
        const long rmax = w < m - 1 - k ? w : m - 1 - k;
        const double pivot = ab[k * bw + w];
        for (long rr = 1; rr <= rmax; ++rr) {
            const double l = ab[(k + rr) * bw + w - rr] / pivot;
            for (long c = 1; c <= rmax; ++c)
                ab[(k + rr) * bw + w - rr + c] = ab[(k + rr) * bw + w - rr + c] - l * ab[k * bw + w + c];
            ab[(k + rr) * bw + w - rr] = l;
        }
    }
    for (long k = 0; k < m - 1; ++k) {
        const long rmax = w < m - 1 - k ? w : m - 1 - k;
        for (long rr = 1; rr <= rmax; ++rr) y[k + rr] = y[k + rr] - ab[(k + rr) * bw + w - rr] * y[k];
    }
    for (long k = m - 1; k >= 0; --k) {
        const double xk = y[k] / ab[k * bw + w];
        const long rmax = w < k ? w : k;
        for (long rr = 1; rr <= rmax; ++rr) y[k - rr] = y[k - rr] - ab[(k - rr) * bw + w + rr] * xk;
        y[k] = xk;
    }
    for (long p = 0; p < m; ++p) u[(1 + p / ni) * ld + 1 + p % ni] = y[p];
    free(ab); free(y);
}

void orc_set_coarse_exact(orc_solver *s, int on) { s->coarse_exact = on; }

/* mg_inner (multigrid.cpp:17-92).  Level l works on n_l = n >> l nodes per side
 * with spacing dx * 2^l; tmp is the single level-0 scratch re-read with the
 * stride of whichever level uses it. */
void orc_cycle(orc_solver *s, int l)
{
    const long nl = s->n >> l;
    const double h = s->dx * (double)(1L << l);   /* dx2 = 2*dx per level, exact (multigrid.cpp:49) */
    double *u = s->u[l], *f = s->rhs[l], *v1 = s->v1[l], *v2 = s->v2[l];

    for (int rep = 0; rep < s->shape; ++rep) {                       /* :52 */
        if (l == s->maxlvl - 1 && s->coarse_exact) {
            orc_coarse_lu_solve(u, f, nl, v1, v2, s->dt, s->nu, h);   /* opt-in, see above */
        } else if (l == s->maxlvl - 1) {
            /* coarsest level: smooth until the ABSOLUTE residual norm drops (:58-65) */
            double rn = 1.0;
            for (int it = 0; it < ORC_COARSE_MAXIT && rn > ORC_COARSE_TOL; ++it) {
                orc_gauss_seidel(u, f, nl, v1, v2, s->dt, s->nu, h);
                orc_residual(s->tmp, u, f, nl, v1, v2, s->dt, s->nu, h);
                rn = orc_norm(s->tmp, nl);
            }
        } else {
            const long nc = nl / 2;
            for (int it = 0; it < ORC_NITER; ++it)                    /* :69-72 */
                orc_gauss_seidel(u, f, nl, v1, v2, s->dt, s->nu, h);
            orc_residual(s->tmp, u, f, nl, v1, v2, s->dt, s->nu, h); /* :73 */
            orc_restriction(s->rhs[l + 1], s->tmp, nl);              /* :75 */
            memset(s->u[l + 1], 0, (size_t)(nc + 1) * (nc + 1) * sizeof(double)); /* :77 */
            orc_cycle(s, l + 1);                                      /* :79 */
            orc_prolongation(s->tmp, s->u[l + 1], nc);               /* :81 */
            const long m = (nl + 1) * (nl + 1);
            for (long p = 0; p < m; ++p) u[p] += s->tmp[p];          /* :83 */
            for (int it = 0; it < ORC_NITER; ++it)                    /* :85-88 */
                orc_gauss_seidel(u, f, nl, v1, v2, s->dt, s->nu, h);
        }
    }
}

void orc_form_rhs(orc_solver *s)
{
    orc_compute_rhs(s->rhs[0], s->u[0], s->n, s->v1[0], s->v2[0], s->dt, s->nu, s->dx);
}

double orc_residual_norm(orc_solver *s)
{
    orc_residual(s->tmp, s->u[0], s->rhs[0], s->n, s->v1[0], s->v2[0], s->dt, s->nu, s->dx);
    return orc_norm(s->tmp, s->n);
}

/* mg_outer (multigrid.cpp:97-120) */
int orc_solve(orc_solver *s, double *hist)
{
    double res0 = orc_residual_norm(s), res = res0;                  /* :104-105 */
    int it = 0;
    if (hist) hist[0] = res0;
    for (; it < ORC_MAX_CYCLE && res / res0 > s->tol; ++it) {        /* :108 */
        orc_cycle(s, 0);                                              /* :110 */
        res = orc_residual_norm(s);                                   /* :112-113 */
        if (hist) hist[it + 1] = res;
    }
    return it;
}

void orc_advance(orc_solver *s, int nsteps, int *cycles)
{
    for (int k = 0; k < nsteps; ++k) {                                /* :165 */
        orc_form_rhs(s);                                              /* :167 */
        int c = orc_solve(s, NULL);                                   /* :169 */
        if (cycles) cycles[k] = c;
    }
}

/* timestepper (multigrid.cpp:124-186) */
void orc_timestepper(double *uT, const double *u0, const double *v1, const double *v2,
                     double nu, int maxlvl, int n, double dt, double T, double dx,
                     double tol, int shape)
{
    orc_solver *s = orc_create(n, maxlvl, u0, v1, v2, nu, dt, dx, tol, shape);
    orc_advance(s, (int)(T / dt), NULL);                              /* :165 */
    memcpy(uT, s->u[0], (size_t)(n + 1) * (n + 1) * sizeof(double)); /* :175 */
    orc_destroy(s);
}

/* initial conditions of main (multigrid.cpp:206-233) */
void orc_initial_conditions(double *u0, double *v1, double *v2, long n, double vscale)
{
    const double PI = 3.1415926535897932;                             /* :14 */
    const double x0 = 0.2, y0 = 0.4, sigma = 100.0;                   /* :206-207 */
    const double kx = 1.0 * PI, ky = 1.0 * PI;                        /* :208-209 */
    const double dx = 1.0 / n;
    const long ld = n + 1;
    for (long i = 0; i <= n; ++i) {
        for (long j = 0; j <= n; ++j) {
            u0[i * ld + j] = exp(-sigma * ((i * dx - x0) * (i * dx - x0)
                                           + (j * dx - y0) * (j * dx - y0)));     /* :219 */
            v1[i * ld + j] = (-ky * sin(kx * i * dx) * cos(ky * j * dx)) * vscale; /* :222 */
            v2[i * ld + j] = (kx * cos(kx * i * dx) * sin(ky * j * dx)) * vscale;  /* :223 */
        }
    }
    /* boundary lines zeroed with i < n: node (n,0) keeps its Gaussian value (:227-233) */
    for (long i = 0; i < n; ++i) {
        u0[i] = 0.0;
        u0[i * ld + n] = 0.0;
        u0[n * ld + i + 1] = 0.0;
        u0[i * ld] = 0.0;
    }
}
