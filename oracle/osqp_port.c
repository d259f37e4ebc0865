/*
 * oracle/osqp_port.c — CPU restatement of the OSQP algorithm.   TEST INFRASTRUCTURE ONLY.
 *
 * The reference hands every FTOCP to the third-party `osqp` package
 * (src/fnc/controller/PredictiveControllers.py:259-283: new OSQP(), setup(P,q,A,l,u,
 * verbose=False, polish=True), solve(), cold start, every step).  `osqp` is an unpinned,
 * un-vendored pip dependency (README.md:18) that is not installable in this image, so its
 * PUBLISHED algorithm is restated here from:
 *   B. Stellato, G. Banjac, P. Goulart, A. Bemporad, S. Boyd,
 *   "OSQP: an operator splitting solver for quadratic programs", Math. Prog. Comp. 12 (2020):
 *     Alg. 1 (ADMM iteration, alpha = 1.6, sigma = 1e-6, rho = 0.1, 1e3*rho on equality rows),
 *     Sec. 3.4 (termination, eps_abs = eps_rel = 1e-3, checked every 25 iterations),
 *     Sec. 5.1 (Ruiz equilibration, 10 passes, with cost scaling),
 *     Sec. 5.2 (adaptive rho, re-factor when rho changes by more than 5x),
 *     Sec. 4   (polish: active-set guess, delta = 1e-6 regularised reduced KKT,
 *               3 iterative-refinement steps),
 * and the quasi-definite KKT system is factorised with an up-looking sparse LDL^T
 *   (T. Davis, "Algorithm 849: a concise sparse Cholesky factorization package", ACM TOMS 2005)
 * after a minimum-degree ordering.  Deviations from the real package (stated, deterministic):
 *   - the adaptive-rho interval is fixed at 25 iterations (the package derives it from wall
 *     clock timings of setup vs. iteration);
 *   - no primal/dual infeasibility certificates (the FTOCP is always feasible: soft lanes);
 *   - minimum-degree ordering instead of AMD.
 * PARITY UNPINNED against the real osqp binary; pinned by oracle/kkt.py instead.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_FACTOR 1e3

typedef struct {
    double rho, sigma, alpha, eps_abs, eps_rel, delta;
    int max_iter, check_every, scaling_iters, adaptive_rho, adaptive_interval;
    double adaptive_tol;
    int polish, polish_refine;
    int polish_strict; /* 0 = the package's acceptance rule; 1 = oracle mode: equality rows always active,
                          accept only if the worse of the two residuals improves */
} osqp_settings;

typedef struct {
    int iters, status, polished, rho_updates; /* status 1 = solved, 2 = max_iter, <0 = error */
    double pri_res, dua_res, obj, rho_final;
} osqp_info;

void osqp_port_default_settings(osqp_settings *s) {
    s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
    s->delta = 1e-6; s->max_iter = 4000; s->check_every = 25; s->scaling_iters = 10;
    s->adaptive_rho = 1; s->adaptive_interval = 25; s->adaptive_tol = 5.0;
    s->polish = 1; s->polish_refine = 3; s->polish_strict = 0;
}

/* ------------------------------------------------------------------ sparse LDL^T ---- */
typedef struct {
    int n;
    int *Lp, *Li, *parent, *lnz, *flag, *pattern, *perm, *iperm;
    double *Lx, *D, *Dinv, *y, *tmp;
    int lcap;
} ldl_t;

static void ldl_free(ldl_t *f) {
    free(f->Lp); free(f->Li); free(f->parent); free(f->lnz); free(f->flag); free(f->pattern);
    free(f->perm); free(f->iperm); free(f->Lx); free(f->D); free(f->Dinv); free(f->y); free(f->tmp);
    memset(f, 0, sizeof(*f));
}

/* minimum-degree ordering on the graph of an upper-CSC symmetric matrix (bitset elimination) */
static void min_degree_order(int n, const int *Ap, const int *Ai, int *perm) {
    int words = (n + 63) / 64;
    uint64_t *adj = (uint64_t *)calloc((size_t)n * words, sizeof(uint64_t));
    char *gone = (char *)calloc(n, 1);
    int *deg = (int *)malloc(n * sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p];
            if (i == j) continue;
            adj[(size_t)i * words + (j >> 6)] |= 1ull << (j & 63);
            adj[(size_t)j * words + (i >> 6)] |= 1ull << (i & 63);
        }
    for (int i = 0; i < n; i++) {
        int c = 0;
        for (int w = 0; w < words; w++) c += __builtin_popcountll(adj[(size_t)i * words + w]);
        deg[i] = c;
    }
    for (int k = 0; k < n; k++) {
        int best = -1, bd = n + 1;
        for (int i = 0; i < n; i++) if (!gone[i] && deg[i] < bd) { bd = deg[i]; best = i; }
        perm[k] = best; gone[best] = 1;
        uint64_t *ab = adj + (size_t)best * words;
        for (int w = 0; w < words; w++) {
            uint64_t bits = ab[w];
            while (bits) {
                int i = (w << 6) + __builtin_ctzll(bits);
                bits &= bits - 1;
                uint64_t *ai = adj + (size_t)i * words;
                int c = 0;
                for (int v = 0; v < words; v++) { ai[v] |= ab[v]; }
                ai[best >> 6] &= ~(1ull << (best & 63));
                ai[i >> 6] &= ~(1ull << (i & 63));
                for (int v = 0; v < words; v++) c += __builtin_popcountll(ai[v]);
                deg[i] = c;
            }
        }
    }
    free(adj); free(gone); free(deg);
}

/* permute symmetric upper-CSC A into upper-CSC C = P A P^T ; map[p] = position in C of A entry p */
static void sym_permute(int n, const int *Ap, const int *Ai, const int *iperm, int *Cp, int *Ci, int *map) {
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i2 = iperm[Ai[p]], j2 = iperm[j];
            cnt[(i2 > j2 ? i2 : j2) + 1]++;
        }
    Cp[0] = 0;
    for (int j = 0; j < n; j++) Cp[j + 1] = Cp[j] + cnt[j + 1];
    memcpy(cnt, Cp, n * sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i2 = iperm[Ai[p]], j2 = iperm[j];
            int c = i2 > j2 ? i2 : j2, r = i2 > j2 ? j2 : i2;
            int q = cnt[c]++;
            Ci[q] = r; map[p] = q;
        }
    free(cnt);
}

static int ldl_symbolic(ldl_t *f, int n, const int *Ap, const int *Ai) {
    f->n = n;
    f->Lp = (int *)malloc((n + 1) * sizeof(int));
    f->parent = (int *)malloc(n * sizeof(int));
    f->lnz = (int *)malloc(n * sizeof(int));
    f->flag = (int *)malloc(n * sizeof(int));
    f->pattern = (int *)malloc(n * sizeof(int));
    f->D = (double *)malloc(n * sizeof(double));
    f->Dinv = (double *)malloc(n * sizeof(double));
    f->y = (double *)calloc(n, sizeof(double));
    f->tmp = (double *)malloc(n * sizeof(double));
    for (int k = 0; k < n; k++) {
        f->parent[k] = -1; f->flag[k] = k; f->lnz[k] = 0;
        for (int p = Ap[k]; p < Ap[k + 1]; p++) {
            int i = Ai[p];
            if (i < k)
                for (; f->flag[i] != k; i = f->parent[i]) {
                    if (f->parent[i] == -1) f->parent[i] = k;
                    f->lnz[i]++; f->flag[i] = k;
                }
        }
    }
    f->Lp[0] = 0;
    for (int k = 0; k < n; k++) f->Lp[k + 1] = f->Lp[k] + f->lnz[k];
    f->lcap = f->Lp[n];
    f->Li = (int *)malloc((f->lcap + 1) * sizeof(int));
    f->Lx = (double *)malloc((f->lcap + 1) * sizeof(double));
    return 0;
}

static int ldl_numeric(ldl_t *f, const int *Ap, const int *Ai, const double *Ax) {
    int n = f->n;
    double *Y = f->y;
    for (int k = 0; k < n; k++) {
        int top = n;
        Y[k] = 0.0; f->flag[k] = k; f->lnz[k] = 0;
        for (int p = Ap[k]; p < Ap[k + 1]; p++) {
            int i = Ai[p];
            if (i <= k) {
                Y[i] += Ax[p];
                int len = 0;
                for (; f->flag[i] != k; i = f->parent[i]) { f->pattern[len++] = i; f->flag[i] = k; }
                while (len > 0) f->pattern[--top] = f->pattern[--len];
            }
        }
        f->D[k] = Y[k]; Y[k] = 0.0;
        for (; top < n; top++) {
            int i = f->pattern[top];
            double yi = Y[i];
            Y[i] = 0.0;
            int p2 = f->Lp[i] + f->lnz[i];
            for (int p = f->Lp[i]; p < p2; p++) Y[f->Li[p]] -= f->Lx[p] * yi;
            double lki = yi * f->Dinv[i];
            f->D[k] -= lki * yi;
            f->Li[p2] = k; f->Lx[p2] = lki; f->lnz[i]++;
        }
        if (f->D[k] == 0.0) return -1;
        f->Dinv[k] = 1.0 / f->D[k];
    }
    return 0;
}

/* solve (P K P^T) with permutation: x <- K^{-1} x */
static void ldl_solve(const ldl_t *f, double *x) {
    int n = f->n;
    double *t = f->tmp;
    for (int i = 0; i < n; i++) t[i] = x[f->perm[i]];
    for (int j = 0; j < n; j++) {
        double tj = t[j];
        for (int p = f->Lp[j]; p < f->Lp[j + 1]; p++) t[f->Li[p]] -= f->Lx[p] * tj;
    }
    for (int j = 0; j < n; j++) t[j] *= f->Dinv[j];
    for (int j = n - 1; j >= 0; j--) {
        double tj = t[j];
        for (int p = f->Lp[j]; p < f->Lp[j + 1]; p++) tj -= f->Lx[p] * t[f->Li[p]];
        t[j] = tj;
    }
    for (int i = 0; i < n; i++) x[f->perm[i]] = t[i];
}

/* ------------------------------------------------------------------ helpers -------- */
static double vinf(const double *v, int n) { double m = 0; for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > m) m = a; } return m; }
/* y = P x, P symmetric stored upper CSC */
static void sym_mv(int n, const int *Pp, const int *Pi, const double *Px, const double *x, double *y) {
    for (int i = 0; i < n; i++) y[i] = 0;
    for (int j = 0; j < n; j++)
        for (int p = Pp[j]; p < Pp[j + 1]; p++) {
            int i = Pi[p];
            y[i] += Px[p] * x[j];
            if (i != j) y[j] += Px[p] * x[i];
        }
}
static void a_mv(int m, int n, const int *Ap, const int *Ai, const double *Ax, const double *x, double *y) {
    for (int i = 0; i < m; i++) y[i] = 0;
    for (int j = 0; j < n; j++) { double xj = x[j]; for (int p = Ap[j]; p < Ap[j + 1]; p++) y[Ai[p]] += Ax[p] * xj; }
}
static void at_mv(int m, int n, const int *Ap, const int *Ai, const double *Ax, const double *y, double *x) {
    (void)m;
    for (int j = 0; j < n; j++) { double s = 0; for (int p = Ap[j]; p < Ap[j + 1]; p++) s += Ax[p] * y[Ai[p]]; x[j] = s; }
}
static double limit_scale(double v) { if (v < MIN_SCALING) v = 1.0; if (v > MAX_SCALING) v = MAX_SCALING; return v; }

/* ------------------------------------------------------------------ workspace ------ */
typedef struct {
    int n, m, nk;
    const int *Pp, *Pi, *Ap, *Ai;
    double *Px, *Ax, *q, *l, *u;          /* scaled copies */
    double *Dv, *Ev, *Dinv, *Einv, c, cinv;
    double *rho, *rhoinv; int *is_eq;
    int *Kp, *Ki; double *Kx;             /* KKT upper CSC (original order) */
    int *KPp, *KPi; double *KPx; int *kmap; /* permuted */
    int *pdiag_idx, *rho_idx, *P_idx, *A_idx; /* positions in Kx */
    ldl_t F;
    double *x, *z, *y, *xt, *zt, *rhs, *xprev, *zprev, *Ax_, *Px_, *Aty, *w1, *w2;
} work_t;

static void *xm(size_t b) { void *p = malloc(b ? b : 1); return p; }

static void build_kkt_pattern(work_t *w) {
    int n = w->n, m = w->m;
    int nnzP = w->Pp[n], nnzA = w->Ap[n];
    int nk = n + m;
    /* column j<n: P upper entries of col j (+ diag if missing); column n+i: A row i entries + diag.
       Build by counting: A^T part goes in columns n..n+m-1 rows j (upper since j < n+i). */
    int *cnt = (int *)calloc(nk + 1, sizeof(int));
    char *hasdiag = (char *)calloc(n, 1);
    for (int j = 0; j < n; j++) for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) { if (w->Pi[p] == j) hasdiag[j] = 1; cnt[j + 1]++; }
    for (int j = 0; j < n; j++) if (!hasdiag[j]) cnt[j + 1]++;
    for (int j = 0; j < n; j++) for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) cnt[n + w->Ai[p] + 1]++;
    for (int i = 0; i < m; i++) cnt[n + i + 1]++;
    w->Kp = (int *)xm((nk + 1) * sizeof(int));
    w->Kp[0] = 0;
    for (int j = 0; j < nk; j++) w->Kp[j + 1] = w->Kp[j] + cnt[j + 1];
    int nnzK = w->Kp[nk];
    w->Ki = (int *)xm(nnzK * sizeof(int));
    w->Kx = (double *)xm(nnzK * sizeof(double));
    w->P_idx = (int *)xm(nnzP * sizeof(int));
    w->A_idx = (int *)xm(nnzA * sizeof(int));
    w->pdiag_idx = (int *)xm(n * sizeof(int));
    w->rho_idx = (int *)xm(m * sizeof(int));
    int *pos = (int *)xm(nk * sizeof(int));
    memcpy(pos, w->Kp, nk * sizeof(int));
    for (int j = 0; j < n; j++) {
        for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
            int q = pos[j]++; w->Ki[q] = w->Pi[p]; w->P_idx[p] = q;
            if (w->Pi[p] == j) w->pdiag_idx[j] = q;
        }
        if (!hasdiag[j]) { int q = pos[j]++; w->Ki[q] = j; w->pdiag_idx[j] = q; }
    }
    for (int j = 0; j < n; j++)
        for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) { int c = n + w->Ai[p]; int q = pos[c]++; w->Ki[q] = j; w->A_idx[p] = q; }
    for (int i = 0; i < m; i++) { int q = pos[n + i]++; w->Ki[q] = n + i; w->rho_idx[i] = q; }
    free(cnt); free(hasdiag); free(pos);
    w->nk = nk;
}

static void fill_kkt(work_t *w, double sigma) {
    int n = w->n, m = w->m;
    int nnzK = w->Kp[w->nk];
    for (int p = 0; p < nnzK; p++) w->Kx[p] = 0.0;
    for (int p = 0; p < w->Pp[n]; p++) w->Kx[w->P_idx[p]] += w->Px[p];
    for (int j = 0; j < n; j++) w->Kx[w->pdiag_idx[j]] += sigma;
    for (int p = 0; p < w->Ap[n]; p++) w->Kx[w->A_idx[p]] = w->Ax[p];
    for (int i = 0; i < m; i++) w->Kx[w->rho_idx[i]] = -w->rhoinv[i];
}

static int factor_kkt(work_t *w, int first) {
    int nk = w->nk, nnzK = w->Kp[nk];
    if (first) {
        w->F.perm = (int *)xm(nk * sizeof(int));
        w->F.iperm = (int *)xm(nk * sizeof(int));
        min_degree_order(nk, w->Kp, w->Ki, w->F.perm);
        for (int i = 0; i < nk; i++) w->F.iperm[w->F.perm[i]] = i;
        w->KPp = (int *)xm((nk + 1) * sizeof(int));
        w->KPi = (int *)xm(nnzK * sizeof(int));
        w->KPx = (double *)xm(nnzK * sizeof(double));
        w->kmap = (int *)xm(nnzK * sizeof(int));
        sym_permute(nk, w->Kp, w->Ki, w->F.iperm, w->KPp, w->KPi, w->kmap);
        ldl_symbolic(&w->F, nk, w->KPp, w->KPi);
    }
    for (int p = 0; p < nnzK; p++) w->KPx[w->kmap[p]] = w->Kx[p];
    return ldl_numeric(&w->F, w->KPp, w->KPi, w->KPx);
}

static void set_rho_vec(work_t *w, double rho) {
    for (int i = 0; i < w->m; i++) {
        double r;
        if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) { r = RHO_MIN; w->is_eq[i] = -1; }
        else if (w->u[i] - w->l[i] < RHO_TOL) { r = RHO_EQ_FACTOR * rho; w->is_eq[i] = 1; }
        else { r = rho; w->is_eq[i] = 0; }
        w->rho[i] = r; w->rhoinv[i] = 1.0 / r;
    }
}

static void ruiz_scale(work_t *w, int iters) {
    int n = w->n, m = w->m;
    double *dn = (double *)xm((n + m) * sizeof(double));
    for (int i = 0; i < n; i++) w->Dv[i] = 1.0;
    for (int i = 0; i < m; i++) w->Ev[i] = 1.0;
    w->c = 1.0;
    for (int it = 0; it < iters; it++) {
        for (int i = 0; i < n + m; i++) dn[i] = 0;
        /* column inf-norms of [P A'; A 0] */
        for (int j = 0; j < n; j++)
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
                double a = fabs(w->Px[p]); int i = w->Pi[p];
                if (a > dn[j]) dn[j] = a;
                if (a > dn[i]) dn[i] = a;
            }
        for (int j = 0; j < n; j++)
            for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
                double a = fabs(w->Ax[p]); int i = w->Ai[p];
                if (a > dn[j]) dn[j] = a;
                if (a > dn[n + i]) dn[n + i] = a;
            }
        for (int i = 0; i < n + m; i++) dn[i] = 1.0 / sqrt(limit_scale(dn[i]));
        for (int j = 0; j < n; j++)
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) w->Px[p] *= dn[j] * dn[w->Pi[p]];
        for (int j = 0; j < n; j++)
            for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) w->Ax[p] *= dn[j] * dn[n + w->Ai[p]];
        for (int j = 0; j < n; j++) { w->q[j] *= dn[j]; w->Dv[j] *= dn[j]; }
        for (int i = 0; i < m; i++) w->Ev[i] *= dn[n + i];
        /* cost scaling */
        double *cn = dn; /* reuse: column norms of P */
        for (int j = 0; j < n; j++) cn[j] = 0;
        for (int j = 0; j < n; j++)
            for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
                double a = fabs(w->Px[p]); int i = w->Pi[p];
                if (a > cn[j]) cn[j] = a;
                if (a > cn[i]) cn[i] = a;
            }
        double mean = 0; for (int j = 0; j < n; j++) mean += cn[j]; mean /= (n > 0 ? n : 1);
        double qn = limit_scale(vinf(w->q, n));
        double ct = mean > qn ? mean : qn;
        ct = 1.0 / limit_scale(ct);
        for (int p = 0; p < w->Pp[n]; p++) w->Px[p] *= ct;
        for (int j = 0; j < n; j++) w->q[j] *= ct;
        w->c *= ct;
    }
    for (int j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->Dv[j];
    for (int i = 0; i < m; i++) w->Einv[i] = 1.0 / w->Ev[i];
    w->cinv = 1.0 / w->c;
    for (int i = 0; i < m; i++) {
        w->l[i] = (w->l[i] <= -OSQP_INFTY) ? -OSQP_INFTY : w->l[i] * w->Ev[i];
        w->u[i] = (w->u[i] >= OSQP_INFTY) ? OSQP_INFTY : w->u[i] * w->Ev[i];
    }
    free(dn);
}

/* unscaled residuals + tolerances for the current (scaled) iterate */
static void residuals(work_t *w, const double *x, const double *z, const double *y,
                      double *pri, double *dua, double *npri, double *ndua,
                      double *spri, double *sdua, double *snpri, double *sndua) {
    int n = w->n, m = w->m;
    a_mv(m, n, w->Ap, w->Ai, w->Ax, x, w->Ax_);
    sym_mv(n, w->Pp, w->Pi, w->Px, x, w->Px_);
    at_mv(m, n, w->Ap, w->Ai, w->Ax, y, w->Aty);
    double rp = 0, na = 0, nz = 0, srp = 0, sna = 0, snz = 0;
    for (int i = 0; i < m; i++) {
        double r = w->Ax_[i] - z[i];
        double a;
        a = fabs(r) * w->Einv[i]; if (a > rp) rp = a;
        a = fabs(w->Ax_[i]) * w->Einv[i]; if (a > na) na = a;
        a = fabs(z[i]) * w->Einv[i]; if (a > nz) nz = a;
        a = fabs(r); if (a > srp) srp = a;
        a = fabs(w->Ax_[i]); if (a > sna) sna = a;
        a = fabs(z[i]); if (a > snz) snz = a;
    }
    double rd = 0, np_ = 0, nat = 0, nq = 0, srd = 0, snp = 0, snat = 0, snq = 0;
    for (int j = 0; j < n; j++) {
        double r = w->Px_[j] + w->q[j] + w->Aty[j];
        double a;
        a = fabs(r) * w->Dinv[j]; if (a > rd) rd = a;
        a = fabs(w->Px_[j]) * w->Dinv[j]; if (a > np_) np_ = a;
        a = fabs(w->Aty[j]) * w->Dinv[j]; if (a > nat) nat = a;
        a = fabs(w->q[j]) * w->Dinv[j]; if (a > nq) nq = a;
        a = fabs(r); if (a > srd) srd = a;
        a = fabs(w->Px_[j]); if (a > snp) snp = a;
        a = fabs(w->Aty[j]); if (a > snat) snat = a;
        a = fabs(w->q[j]); if (a > snq) snq = a;
    }
    *pri = rp; *dua = rd * w->cinv;
    *npri = na > nz ? na : nz;
    double t = np_ > nat ? np_ : nat; t = t > nq ? t : nq; *ndua = t * w->cinv;
    *spri = srp; *sdua = srd; *snpri = sna > snz ? sna : snz;
    t = snp > snat ? snp : snat; *sndua = t > snq ? t : snq;
}

static double objective_unscaled(work_t *w, const double *x) {
    int n = w->n;
    sym_mv(n, w->Pp, w->Pi, w->Px, x, w->Px_);
    double o = 0; for (int j = 0; j < n; j++) o += x[j] * (0.5 * w->Px_[j] + w->q[j]);
    return o * w->cinv;
}

/* ------------------------------------------------------------------ polish ---------- */
static int polish(work_t *w, const osqp_settings *s, double *x, double *z, double *y, osqp_info *info) {
    int n = w->n, m = w->m;
    int *act = (int *)xm(m * sizeof(int));   /* index into reduced rows or -1 */
    int *side = (int *)xm(m * sizeof(int));  /* -1 lower, +1 upper */
    int mr = 0;
    for (int i = 0; i < m; i++) {
        if (s->polish_strict && w->is_eq[i] == 1) { act[i] = mr++; side[i] = -1; }
        else if (z[i] - w->l[i] < -y[i]) { act[i] = mr++; side[i] = -1; }
        else if (w->u[i] - z[i] < y[i]) { act[i] = mr++; side[i] = +1; }
        else { act[i] = -1; side[i] = 0; }
    }
    int nk = n + mr;
    /* reduced KKT upper CSC: [P + delta I, Ared'; Ared, -delta I] */
    int nnzP = w->Pp[n];
    int *Kp = (int *)calloc(nk + 1, sizeof(int));
    char *hasdiag = (char *)calloc(n, 1);
    for (int j = 0; j < n; j++) for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) { if (w->Pi[p] == j) hasdiag[j] = 1; Kp[j + 1]++; }
    for (int j = 0; j < n; j++) if (!hasdiag[j]) Kp[j + 1]++;
    for (int j = 0; j < n; j++) for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) if (act[w->Ai[p]] >= 0) Kp[n + act[w->Ai[p]] + 1]++;
    for (int i = 0; i < mr; i++) Kp[n + i + 1]++;
    for (int j = 0; j < nk; j++) Kp[j + 1] += Kp[j];
    int nnzK = Kp[nk];
    int *Ki = (int *)xm(nnzK * sizeof(int));
    double *Kx = (double *)xm(nnzK * sizeof(double));
    int *pos = (int *)xm(nk * sizeof(int));
    memcpy(pos, Kp, nk * sizeof(int));
    (void)nnzP;
    for (int j = 0; j < n; j++) {
        int saw = 0;
        for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
            int q = pos[j]++; Ki[q] = w->Pi[p]; Kx[q] = w->Px[p];
            if (w->Pi[p] == j) { Kx[q] += s->delta; saw = 1; }
        }
        if (!saw) { int q = pos[j]++; Ki[q] = j; Kx[q] = s->delta; }
    }
    for (int j = 0; j < n; j++)
        for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
            int r = act[w->Ai[p]];
            if (r >= 0) { int q = pos[n + r]++; Ki[q] = j; Kx[q] = w->Ax[p]; }
        }
    for (int i = 0; i < mr; i++) { int q = pos[n + i]++; Ki[q] = n + i; Kx[q] = -s->delta; }
    ldl_t F; memset(&F, 0, sizeof(F));
    F.perm = (int *)xm(nk * sizeof(int)); F.iperm = (int *)xm(nk * sizeof(int));
    min_degree_order(nk, Kp, Ki, F.perm);
    for (int i = 0; i < nk; i++) F.iperm[F.perm[i]] = i;
    int *KPp = (int *)xm((nk + 1) * sizeof(int)), *KPi = (int *)xm(nnzK * sizeof(int)), *kmap = (int *)xm(nnzK * sizeof(int));
    double *KPx = (double *)xm(nnzK * sizeof(double));
    sym_permute(nk, Kp, Ki, F.iperm, KPp, KPi, kmap);
    for (int p = 0; p < nnzK; p++) KPx[kmap[p]] = Kx[p];
    ldl_symbolic(&F, nk, KPp, KPi);
    int ok = ldl_numeric(&F, KPp, KPi, KPx) == 0;
    int success = 0;
    if (ok) {
        double *rhs = (double *)xm(nk * sizeof(double)), *sol = (double *)xm(nk * sizeof(double));
        double *res = (double *)xm(nk * sizeof(double));
        double *yr = (double *)xm((mr + 1) * sizeof(double));
        for (int j = 0; j < n; j++) rhs[j] = -w->q[j];
        for (int i = 0; i < m; i++) if (act[i] >= 0) rhs[n + act[i]] = side[i] < 0 ? w->l[i] : w->u[i];
        memcpy(sol, rhs, nk * sizeof(double));
        ldl_solve(&F, sol);
        for (int it = 0; it < s->polish_refine; it++) {
            /* res = rhs - K0 sol, K0 = [P Ared'; Ared 0] */
            sym_mv(n, w->Pp, w->Pi, w->Px, sol, res);
            for (int j = 0; j < n; j++) res[j] = rhs[j] - res[j];
            for (int i = 0; i < mr; i++) res[n + i] = rhs[n + i];
            for (int j = 0; j < n; j++)
                for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
                    int r = act[w->Ai[p]];
                    if (r >= 0) { res[j] -= w->Ax[p] * sol[n + r]; res[n + r] -= w->Ax[p] * sol[j]; }
                }
            ldl_solve(&F, res);
            for (int i = 0; i < nk; i++) sol[i] += res[i];
        }
        double *xp = sol, *zp = (double *)xm(m * sizeof(double)), *yp = (double *)xm(m * sizeof(double));
        a_mv(m, n, w->Ap, w->Ai, w->Ax, xp, zp);
        for (int i = 0; i < m; i++) yp[i] = act[i] >= 0 ? sol[n + act[i]] : 0.0;
        /* polished residuals (unscaled): primal = violation of [l,u] by A x */
        double pr = 0;
        for (int i = 0; i < m; i++) {
            double v = 0;
            if (zp[i] < w->l[i]) v = w->l[i] - zp[i]; else if (zp[i] > w->u[i]) v = zp[i] - w->u[i];
            v *= w->Einv[i]; if (v > pr) pr = v;
        }
        sym_mv(n, w->Pp, w->Pi, w->Px, xp, w->Px_);
        at_mv(m, n, w->Ap, w->Ai, w->Ax, yp, w->Aty);
        double dr = 0;
        for (int j = 0; j < n; j++) { double a = fabs(w->Px_[j] + w->q[j] + w->Aty[j]) * w->Dinv[j]; if (a > dr) dr = a; }
        dr *= w->cinv;
        if (s->polish_strict) {
            double before = info->pri_res > info->dua_res ? info->pri_res : info->dua_res;
            double after = pr > dr ? pr : dr;
            success = after < before;
        } else
        success = (pr < info->pri_res && dr < info->dua_res) ||
                  (pr < info->pri_res && info->dua_res < 1e-10) ||
                  (dr < info->dua_res && info->pri_res < 1e-10);
        if (success) {
            memcpy(x, xp, n * sizeof(double));
            memcpy(z, zp, m * sizeof(double));
            memcpy(y, yp, m * sizeof(double));
            info->pri_res = pr; info->dua_res = dr;
        }
        free(rhs); free(sol); free(res); free(yr); free(zp); free(yp);
    }
    ldl_free(&F);
    free(act); free(side); free(Kp); free(hasdiag); free(Ki); free(Kx); free(pos); free(KPp); free(KPi); free(kmap); free(KPx);
    return success;
}

/* ------------------------------------------------------------------ driver ---------- */
/* P upper-triangular CSC (n x n), A CSC (m x n).  x_out[n], y_out[m] (unscaled). */
int osqp_port_solve(int n, int m, const int *Pp, const int *Pi, const double *Px_in, const double *q_in,
                    const int *Ap, const int *Ai, const double *Ax_in, const double *l_in, const double *u_in,
                    const osqp_settings *s, double *x_out, double *y_out, osqp_info *info) {
    work_t W; memset(&W, 0, sizeof(W));
    work_t *w = &W;
    w->n = n; w->m = m; w->Pp = Pp; w->Pi = Pi; w->Ap = Ap; w->Ai = Ai;
    int nnzP = Pp[n], nnzA = Ap[n];
    w->Px = (double *)xm(nnzP * sizeof(double)); memcpy(w->Px, Px_in, nnzP * sizeof(double));
    w->Ax = (double *)xm(nnzA * sizeof(double)); memcpy(w->Ax, Ax_in, nnzA * sizeof(double));
    w->q = (double *)xm(n * sizeof(double)); memcpy(w->q, q_in, n * sizeof(double));
    w->l = (double *)xm(m * sizeof(double)); w->u = (double *)xm(m * sizeof(double));
    for (int i = 0; i < m; i++) {
        w->l[i] = l_in[i] < -OSQP_INFTY ? -OSQP_INFTY : l_in[i];
        w->u[i] = u_in[i] > OSQP_INFTY ? OSQP_INFTY : u_in[i];
    }
    w->Dv = (double *)xm(n * sizeof(double)); w->Dinv = (double *)xm(n * sizeof(double));
    w->Ev = (double *)xm(m * sizeof(double)); w->Einv = (double *)xm(m * sizeof(double));
    w->rho = (double *)xm(m * sizeof(double)); w->rhoinv = (double *)xm(m * sizeof(double));
    w->is_eq = (int *)xm(m * sizeof(int));
    int nk = n + m;
    w->x = (double *)calloc(n, sizeof(double)); w->z = (double *)calloc(m, sizeof(double)); w->y = (double *)calloc(m, sizeof(double));
    w->xt = (double *)xm(n * sizeof(double)); w->zt = (double *)xm(m * sizeof(double));
    w->rhs = (double *)xm(nk * sizeof(double));
    w->xprev = (double *)xm(n * sizeof(double)); w->zprev = (double *)xm(m * sizeof(double));
    w->Ax_ = (double *)xm(m * sizeof(double)); w->Px_ = (double *)xm(n * sizeof(double)); w->Aty = (double *)xm(n * sizeof(double));

    /* setup: scaling, rho, KKT, ordering, factorisation (PC.py:269-275 does this every call) */
    if (s->scaling_iters > 0) ruiz_scale(w, s->scaling_iters);
    else {
        for (int j = 0; j < n; j++) { w->Dv[j] = w->Dinv[j] = 1; }
        for (int i = 0; i < m; i++) { w->Ev[i] = w->Einv[i] = 1; }
        w->c = w->cinv = 1;
    }
    double rho = s->rho;
    set_rho_vec(w, rho);
    build_kkt_pattern(w);
    fill_kkt(w, s->sigma);
    int rc = factor_kkt(w, 1);
    memset(info, 0, sizeof(*info));
    if (rc != 0) { info->status = -1; goto done; }

    {
        double pri = 0, dua = 0, npri = 0, ndua = 0, spri, sdua, snpri, sndua;
        int iter, status = 2;
        for (iter = 1; iter <= s->max_iter; iter++) {
            memcpy(w->xprev, w->x, n * sizeof(double));
            memcpy(w->zprev, w->z, m * sizeof(double));
            for (int j = 0; j < n; j++) w->rhs[j] = s->sigma * w->xprev[j] - w->q[j];
            for (int i = 0; i < m; i++) w->rhs[n + i] = w->zprev[i] - w->rhoinv[i] * w->y[i];
            ldl_solve(&w->F, w->rhs);
            for (int j = 0; j < n; j++) w->xt[j] = w->rhs[j];
            for (int i = 0; i < m; i++) w->zt[i] = w->zprev[i] + w->rhoinv[i] * (w->rhs[n + i] - w->y[i]);
            for (int j = 0; j < n; j++) w->x[j] = s->alpha * w->xt[j] + (1.0 - s->alpha) * w->xprev[j];
            for (int i = 0; i < m; i++) {
                double zr = s->alpha * w->zt[i] + (1.0 - s->alpha) * w->zprev[i];
                double v = zr + w->rhoinv[i] * w->y[i];
                if (v < w->l[i]) v = w->l[i];
                if (v > w->u[i]) v = w->u[i];
                w->z[i] = v;
                w->y[i] = w->y[i] + w->rho[i] * (zr - v);
            }
            int check = (s->check_every > 0 && iter % s->check_every == 0) || iter == s->max_iter;
            int adapt = s->adaptive_rho && s->adaptive_interval > 0 && iter % s->adaptive_interval == 0;
            if (check || adapt) {
                residuals(w, w->x, w->z, w->y, &pri, &dua, &npri, &ndua, &spri, &sdua, &snpri, &sndua);
                if (check) {
                    double ep = s->eps_abs + s->eps_rel * npri;
                    double ed = s->eps_abs + s->eps_rel * ndua;
                    if (pri <= ep && dua <= ed) { status = 1; break; }
                }
                if (adapt) {
                    double a = spri / (snpri + 1e-10), b = sdua / (sndua + 1e-10);
                    double rn = rho * sqrt(a / (b + 1e-10));
                    if (rn < RHO_MIN) rn = RHO_MIN;
                    if (rn > RHO_MAX) rn = RHO_MAX;
                    if (rn > rho * s->adaptive_tol || rn < rho / s->adaptive_tol) {
                        rho = rn;
                        for (int i = 0; i < m; i++) {
                            double r = w->is_eq[i] == 1 ? RHO_EQ_FACTOR * rho : (w->is_eq[i] == -1 ? RHO_MIN : rho);
                            w->rho[i] = r; w->rhoinv[i] = 1.0 / r;
                        }
                        fill_kkt(w, s->sigma);
                        if (factor_kkt(w, 0) != 0) { status = -1; break; }
                        info->rho_updates++;
                    }
                }
            }
        }
        if (iter > s->max_iter) iter = s->max_iter;
        residuals(w, w->x, w->z, w->y, &pri, &dua, &npri, &ndua, &spri, &sdua, &snpri, &sndua);
        info->iters = iter; info->status = status; info->pri_res = pri; info->dua_res = dua; info->rho_final = rho;
        if (s->polish && status == 1) info->polished = polish(w, s, w->x, w->z, w->y, info);
        info->obj = objective_unscaled(w, w->x);
        for (int j = 0; j < n; j++) x_out[j] = w->x[j] * w->Dv[j];
        if (y_out) for (int i = 0; i < m; i++) y_out[i] = w->y[i] * w->Ev[i] * w->cinv;
    }
done:
    ldl_free(&w->F);
    free(w->Px); free(w->Ax); free(w->q); free(w->l); free(w->u); free(w->Dv); free(w->Dinv); free(w->Ev); free(w->Einv);
    free(w->rho); free(w->rhoinv); free(w->is_eq); free(w->Kp); free(w->Ki); free(w->Kx); free(w->KPp); free(w->KPi); free(w->KPx);
    free(w->kmap); free(w->pdiag_idx); free(w->rho_idx); free(w->P_idx); free(w->A_idx);
    free(w->x); free(w->z); free(w->y); free(w->xt); free(w->zt); free(w->rhs); free(w->xprev); free(w->zprev);
    free(w->Ax_); free(w->Px_); free(w->Aty);
    return info->status;
}

/* Batch of QPs sharing one sparsity pattern; values differ per problem.  Threads = OpenMP. */
int osqp_port_solve_batch(int nprob, int n, int m, const int *Pp, const int *Pi, const double *Px /*[nprob][nnzP]*/,
                          const double *q /*[nprob][n]*/, const int *Ap, const int *Ai, const double *Ax /*[nprob][nnzA]*/,
                          const double *l, const double *u, const osqp_settings *s, int nthreads,
                          double *x_out /*[nprob][n]*/, double *y_out /*[nprob][m] or NULL*/, osqp_info *infos) {
    int nnzP = Pp[n], nnzA = Ap[n];
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < nprob; b++) {
        osqp_port_solve(n, m, Pp, Pi, Px + (size_t)b * nnzP, q + (size_t)b * n, Ap, Ai, Ax + (size_t)b * nnzA,
                        l + (size_t)b * m, u + (size_t)b * m, s, x_out + (size_t)b * n,
                        y_out ? y_out + (size_t)b * m : NULL, infos + b);
    }
    return 0;
}

int osqp_port_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
