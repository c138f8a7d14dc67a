/* orc_lcp.c -- CPU restatement (TEST INFRASTRUCTURE, see orc.h) of the dense LCP solver behind dWorldStep.
 *
 * The reference selects it with fast_step == false (src/WorldPhysics.cpp:367-370: dWorldStep instead of
 * dWorldQuickStep).  The arithmetic is ODE's (un-vendored dependency simulation/ode-16, manifest.xml:15): step.cpp
 * builds A = J invM J^T + diag(cfm/h) and rhs = c/h - J (v/h + invM fe) per island and hands them to dSolveLCP
 * (lcp.cpp), Dantzig's principal pivoting method as described by Baraff ("Fast contact force computation for
 * nonpenetrating rigid bodies", SIGGRAPH 94) with box bounds:
 *   - the unbounded rows (lo = -inf, hi = +inf) go into the clamped set C first and are solved at once;
 *   - rows whose bounds follow another row's multiplier (findex >= 0: friction) are processed last, and when the
 *     first of them is reached ALL of them get lo = -|hi * x[findex]|, hi = +|hi * x[findex]| from the solution of
 *     the friction-less problem found so far (0 / 0 if that multiplier is 0);
 *   - every further row i is the "driving index": x[i] is pushed in the direction that reduces |w[i]| while
 *     A(C,C) dx(C) = -dir A(C,i) keeps w(C) = 0, up to the largest step s that keeps all processed rows valid;
 *     whichever row limits the step changes set (cases 1..6 below), until i itself is in C or at a bound.
 * ODE keeps an LDL^T factorisation of A(C,C) and updates it row by row; here it is a Cholesky factor that grows by
 * one row when an index enters C and is rebuilt when one leaves -- the same solves up to rounding.  With cfm > 0 the
 * matrix is positive definite and the solution of each of the two stages is unique, so the row order only decides
 * ties.  PARITY UNPINNED like the rest of the oracle: restated from the published algorithm, no libode to run against.
 * Computed in double in both builds of the oracle. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc_internal.h"

typedef struct {
    int n, nC;
    const double *A; /* n x n, row-major, symmetric */
    int *C;          /* indexes in C by position */
    double *L;       /* n x n: Cholesky factor of A(C,C) in position space, row-major lower triangle */
    double *t;       /* n scratch */
    int failed;
} chol_t;

static void chol_append(chol_t *S, int idx)
{
    const int p = S->nC, n = S->n;
    double *row = S->L + (size_t)p * n;
    double d = S->A[(size_t)idx * n + idx];
    for (int q = 0; q < p; q++) {
        double s = S->A[(size_t)idx * n + S->C[q]];
        const double *lq = S->L + (size_t)q * n;
        for (int r = 0; r < q; r++) s -= lq[r] * row[r];
        row[q] = s / lq[q];
        d -= row[q] * row[q];
    }
    if (!(d > 0)) {
        S->failed = 1;
        d = 1e-300;
    }
    row[p] = sqrt(d);
    S->C[p] = idx;
    S->nC = p + 1;
}

static void chol_remove(chol_t *S, int idx)
{
    int nC = S->nC, k = 0;
    int *old = (int *)malloc(sizeof(int) * (size_t)(nC > 0 ? nC : 1));
    memcpy(old, S->C, sizeof(int) * (size_t)nC);
    S->nC = 0;
    for (int q = 0; q < nC; q++)
        if (old[q] != idx) chol_append(S, old[q]), k++;
    (void)k;
    free(old);
}

/* out(C) = A(C,C)^-1 rhs, rhs and out by position */
static void chol_solve(const chol_t *S, const double *rhs, double *out)
{
    const int p = S->nC, n = S->n;
    for (int q = 0; q < p; q++) {
        double s = rhs[q];
        const double *lq = S->L + (size_t)q * n;
        for (int r = 0; r < q; r++) s -= lq[r] * out[r];
        out[q] = s / lq[q];
    }
    for (int q = p - 1; q >= 0; q--) {
        double s = out[q];
        for (int r = q + 1; r < p; r++) s -= S->L[(size_t)r * n + q] * out[r];
        out[q] = s / S->L[(size_t)q * n + q];
    }
}

/* Solves  w = A x - b,  lo <= x <= hi,  w = 0 inside the bounds, w >= 0 at lo, w <= 0 at hi.  lo / hi are
 * modified for the rows with findex >= 0 (their effective bounds are returned).  Returns 0, or 1 if the solver gave
 * up (step <= 0 or the iteration cap: the rows not yet processed keep x = 0 as in ODE). */
int orc_solve_lcp(int n, const double *A, double *x, const double *b, double *w, double *lo, double *hi,
                  const int *findex)
{
    if (n <= 0) return 0;
    chol_t S;
    S.n = n, S.nC = 0, S.A = A, S.failed = 0;
    S.C = (int *)malloc(sizeof(int) * (size_t)n);
    S.L = (double *)calloc((size_t)n * n, sizeof(double));
    S.t = (double *)malloc(sizeof(double) * (size_t)n);
    int *ord = (int *)malloc(sizeof(int) * (size_t)n);
    char *set = (char *)calloc((size_t)n, 1); /* 0 not processed, 1 in C, 2 in N at lo, 3 in N at hi */
    double *dx = (double *)calloc((size_t)n, sizeof(double));
    double *dw = (double *)calloc((size_t)n, sizeof(double));
    double *rp = (double *)malloc(sizeof(double) * (size_t)n), *sp = (double *)malloc(sizeof(double) * (size_t)n);
    int err = 0, no = 0, nub = 0;
    for (int i = 0; i < n; i++) x[i] = 0, w[i] = 0;
    /* processing order: unbounded rows, then the bounded ones, then the rows with findex >= 0 */
    for (int i = 0; i < n; i++)
        if (findex[i] < 0 && lo[i] == -INFINITY && hi[i] == INFINITY) ord[no++] = i;
    nub = no;
    for (int i = 0; i < n; i++)
        if (findex[i] < 0 && !(lo[i] == -INFINITY && hi[i] == INFINITY)) ord[no++] = i;
    for (int i = 0; i < n; i++)
        if (findex[i] >= 0) ord[no++] = i;

    /* the unbounded rows at once */
    for (int k = 0; k < nub; k++) {
        chol_append(&S, ord[k]);
        set[ord[k]] = 1;
    }
    if (nub > 0) {
        for (int q = 0; q < nub; q++) rp[q] = b[S.C[q]];
        chol_solve(&S, rp, sp);
        for (int q = 0; q < nub; q++) x[S.C[q]] = sp[q];
    }

    int hit_first_friction_index = 0;
    long guard = 0;
    const long guard_max = 50L * n + 100;
    for (int k = nub; k < n && !err; k++) {
        const int i = ord[k];
        if (!hit_first_friction_index && findex[i] >= 0) {
            for (int kk = k; kk < n; kk++) {
                const int r = ord[kk];
                const double wfk = x[findex[r]];
                if (wfk == 0) {
                    hi[r] = 0, lo[r] = 0;
                } else {
                    hi[r] = fabs(hi[r] * wfk), lo[r] = -hi[r];
                }
            }
            hit_first_friction_index = 1;
        }
        /* w[i] for the driving index (the rows behind it are "don't care") */
        {
            double s = -b[i];
            for (int j = 0; j < n; j++)
                if (set[j]) s += A[(size_t)i * n + j] * x[j];
            w[i] = s;
        }
        if (lo[i] == 0 && w[i] >= 0) {
            set[i] = 2;
            continue;
        }
        if (hi[i] == 0 && w[i] <= 0) {
            set[i] = 3;
            continue;
        }
        if (w[i] == 0) {
            chol_append(&S, i);
            set[i] = 1;
            continue;
        }
        for (;;) {
            if (++guard > guard_max) {
                err = 1;
                break;
            }
            const double dirf = w[i] <= 0 ? 1.0 : -1.0;
            /* dx(C) = -dir A(C,C)^-1 A(C,i) */
            for (int q = 0; q < S.nC; q++) rp[q] = -dirf * A[(size_t)S.C[q] * n + i];
            chol_solve(&S, rp, sp);
            for (int q = 0; q < S.nC; q++) dx[S.C[q]] = sp[q];
            /* dw = A dx for the rows in N and for i */
            for (int j = 0; j < n; j++) {
                if (!(set[j] >= 2 || j == i)) continue;
                double s = A[(size_t)j * n + i] * dirf;
                for (int q = 0; q < S.nC; q++) s += A[(size_t)j * n + S.C[q]] * sp[q];
                dw[j] = s;
            }
            int cmd = 1, si = 0;
            double s = -w[i] / dw[i];
            if (dirf > 0) {
                if (hi[i] < INFINITY) {
                    const double s2 = (hi[i] - x[i]) * dirf;
                    if (s2 < s) s = s2, cmd = 3;
                }
            } else {
                if (lo[i] > -INFINITY) {
                    const double s2 = (lo[i] - x[i]) * dirf;
                    if (s2 < s) s = s2, cmd = 2;
                }
            }
            for (int j = 0; j < n; j++) {
                if (set[j] < 2) continue;
                if ((set[j] == 2 && dw[j] < 0) || (set[j] == 3 && dw[j] > 0)) {
                    if (lo[j] == 0 && hi[j] == 0) continue;
                    const double s2 = -w[j] / dw[j];
                    if (s2 < s) s = s2, cmd = 4, si = j;
                }
            }
            for (int q = nub; q < S.nC; q++) {
                const int j = S.C[q];
                if (dx[j] < 0 && lo[j] > -INFINITY) {
                    const double s2 = (lo[j] - x[j]) / dx[j];
                    if (s2 < s) s = s2, cmd = 5, si = j;
                }
                if (dx[j] > 0 && hi[j] < INFINITY) {
                    const double s2 = (hi[j] - x[j]) / dx[j];
                    if (s2 < s) s = s2, cmd = 6, si = j;
                }
            }
            if (!(s > 0)) {
                /* ODE: "LCP internal error, s <= 0": give up with the current solution */
                for (int kk = k; kk < n; kk++) x[ord[kk]] = 0, w[ord[kk]] = 0;
                err = 1;
                break;
            }
            for (int q = 0; q < S.nC; q++) x[S.C[q]] += s * dx[S.C[q]];
            x[i] += s * dirf;
            for (int j = 0; j < n; j++)
                if (set[j] >= 2) w[j] += s * dw[j];
            w[i] += s * dw[i];
            switch (cmd) {
            case 1: w[i] = 0; chol_append(&S, i); set[i] = 1; break;
            case 2: x[i] = lo[i]; set[i] = 2; break;
            case 3: x[i] = hi[i]; set[i] = 3; break;
            case 4: w[si] = 0; chol_append(&S, si); set[si] = 1; break;
            case 5: x[si] = lo[si]; set[si] = 2; chol_remove(&S, si); break;
            case 6: x[si] = hi[si]; set[si] = 3; chol_remove(&S, si); break;
            }
            if (cmd <= 3) break;
        }
    }
    /* (the unbounded rows never leave C: the bound checks of the clamped set start at position nub) */
    if (S.failed) err = 1;
    free(S.C), free(S.L), free(S.t), free(ord), free(set), free(dx), free(dw), free(rp), free(sp);
    return err;
}
