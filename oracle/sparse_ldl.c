/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
 *
 * sparse_ldl.c: sparse LDL^T factorisation of a symmetric quasi-definite matrix given as
 * the UPPER triangle in CSC form, no pivoting — the role QDLDL plays inside OSQP
 * (OSQP's default linsys_solver; the reference reaches it through
 * OsqpEigen::Solver::initSolver/solve, modules/Walking_module/src/MPCSolver.cpp:295-309).
 * QDLDL is a third-party dependency that is absent from /root/reference; this is a
 * restatement of its published algorithm (up-looking LDL^T with an elimination tree,
 * T. Davis, "Algorithm 849: a concise sparse Cholesky factorization package").
 */
#include "oracle_internal.h"

#include <stdlib.h>
#include <string.h>

/* Elimination tree + column counts of L.  Returns total nnz(L) or -1 on bad input. */
static int ldl_etree(int n, const int* Ap, const int* Ai, int* work, int* Lnz, int* etree)
{
    int i, j, p, sum = 0;
    for (i = 0; i < n; i++) {
        work[i] = 0;
        Lnz[i] = 0;
        etree[i] = -1;
        if (Ap[i] == Ap[i + 1]) return -1; /* empty column: no diagonal */
    }
    for (j = 0; j < n; j++) {
        work[j] = j;
        for (p = Ap[j]; p < Ap[j + 1]; p++) {
            i = Ai[p];
            if (i > j) return -1; /* not upper triangular */
            while (work[i] != j) {
                if (etree[i] == -1) etree[i] = j;
                Lnz[i]++;
                work[i] = j;
                i = etree[i];
            }
        }
    }
    for (i = 0; i < n; i++) sum += Lnz[i];
    return sum;
}

wco_ldl* wco_ldl_symbolic(int n, const int* Ap, const int* Ai)
{
    wco_ldl* f = (wco_ldl*)calloc(1, sizeof(wco_ldl));
    int* work = (int*)malloc(sizeof(int) * (size_t)n);
    f->n = n;
    f->etree = (int*)malloc(sizeof(int) * (size_t)n);
    f->Lnz = (int*)malloc(sizeof(int) * (size_t)n);
    f->Lp = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int nnz = ldl_etree(n, Ap, Ai, work, f->Lnz, f->etree);
    free(work);
    if (nnz < 0) {
        wco_ldl_free(f);
        return NULL;
    }
    f->Li = (int*)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    f->Lx = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    f->D = (double*)malloc(sizeof(double) * (size_t)n);
    f->Dinv = (double*)malloc(sizeof(double) * (size_t)n);
    f->iwork = (int*)malloc(sizeof(int) * (size_t)(3 * n));
    f->bwork = (unsigned char*)malloc((size_t)n);
    f->fwork = (double*)malloc(sizeof(double) * (size_t)n);
    return f;
}

void wco_ldl_free(wco_ldl* f)
{
    if (!f) return;
    free(f->etree); free(f->Lnz); free(f->Lp); free(f->Li); free(f->Lx);
    free(f->D); free(f->Dinv); free(f->iwork); free(f->bwork); free(f->fwork);
    free(f);
}

/* Numeric factorisation.  Returns the number of positive entries of D, or -1 if a zero
 * pivot is met. */
int wco_ldl_numeric(wco_ldl* f, const int* Ap, const int* Ai, const double* Ax)
{
    const int n = f->n;
    int* Lp = f->Lp; int* Li = f->Li; double* Lx = f->Lx;
    double* D = f->D; double* Dinv = f->Dinv;
    const int* etree = f->etree; const int* Lnz = f->Lnz;
    int* yIdx = f->iwork;
    int* elimBuffer = f->iwork + n;
    int* LNextSpaceInCol = f->iwork + 2 * n;
    unsigned char* yMarkers = f->bwork;
    double* yVals = f->fwork;
    int i, j, k, nnzY, bidx, cidx, nextIdx, nnzE, tmpIdx, positive = 0;
    double yVals_cidx;

    Lp[0] = 0;
    for (i = 0; i < n; i++) {
        Lp[i + 1] = Lp[i] + Lnz[i];
        yMarkers[i] = 0;
        yVals[i] = 0.0;
        D[i] = 0.0;
        LNextSpaceInCol[i] = Lp[i];
    }
    /* first pivot */
    D[0] = Ax[0];
    if (D[0] == 0.0) return -1;
    if (D[0] > 0.0) positive++;
    Dinv[0] = 1.0 / D[0];

    for (k = 1; k < n; k++) {
        nnzY = 0;
        tmpIdx = Ap[k + 1];
        for (i = Ap[k]; i < tmpIdx; i++) {
            bidx = Ai[i];
            if (bidx == k) {
                D[k] = Ax[i];
                continue;
            }
            yVals[bidx] = Ax[i];
            nextIdx = bidx;
            if (!yMarkers[nextIdx]) {
                yMarkers[nextIdx] = 1;
                elimBuffer[0] = nextIdx;
                nnzE = 1;
                nextIdx = etree[bidx];
                while (nextIdx != -1 && nextIdx < k) {
                    if (yMarkers[nextIdx]) break;
                    yMarkers[nextIdx] = 1;
                    elimBuffer[nnzE++] = nextIdx;
                    nextIdx = etree[nextIdx];
                }
                while (nnzE) yIdx[nnzY++] = elimBuffer[--nnzE];
            }
        }
        for (i = nnzY - 1; i >= 0; i--) {
            cidx = yIdx[i];
            tmpIdx = LNextSpaceInCol[cidx];
            yVals_cidx = yVals[cidx];
            for (j = Lp[cidx]; j < tmpIdx; j++) yVals[Li[j]] -= Lx[j] * yVals_cidx;
            Li[tmpIdx] = k;
            Lx[tmpIdx] = yVals_cidx * Dinv[cidx];
            D[k] -= yVals_cidx * Lx[tmpIdx];
            LNextSpaceInCol[cidx]++;
            yVals[cidx] = 0.0;
            yMarkers[cidx] = 0;
        }
        if (D[k] == 0.0) return -1;
        if (D[k] > 0.0) positive++;
        Dinv[k] = 1.0 / D[k];
    }
    return positive;
}

/* x <- (L D L^T)^{-1} x */
void wco_ldl_solve(const wco_ldl* f, double* x)
{
    const int n = f->n;
    const int* Lp = f->Lp; const int* Li = f->Li; const double* Lx = f->Lx;
    int i, j;
    for (i = 0; i < n; i++) {
        const double xi = x[i];
        for (j = Lp[i]; j < Lp[i + 1]; j++) x[Li[j]] -= Lx[j] * xi;
    }
    for (i = 0; i < n; i++) x[i] *= f->Dinv[i];
    for (i = n - 1; i >= 0; i--) {
        double xi = x[i];
        for (j = Lp[i]; j < Lp[i + 1]; j++) xi -= Lx[j] * x[Li[j]];
        x[i] = xi;
    }
}
