/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Dense-input convenience wrapper around the OSQP
 * restatement (osqp_ref.c) so that Python tests can cross-check it on arbitrary small QPs.
 */
#include "oracle_internal.h"
#include "wc_oracle.h"

#include <stdlib.h>

int wco_osqp_solve_dense(int n, int m, const double* P, const double* A, const double* q,
                         const double* l, const double* u, const wc_osqp_settings* settings,
                         int linsys_kind, double* x_out, double* y_out, double* info_out)
{
    int cap = n * n + m * n + 1, nt = 0, i, j;
    int* ti = (int*)malloc(sizeof(int) * (size_t)cap);
    int* tj = (int*)malloc(sizeof(int) * (size_t)cap);
    double* tx = (double*)malloc(sizeof(double) * (size_t)cap);
    for (j = 0; j < n; j++)
        for (i = 0; i <= j; i++)
            if (P[(size_t)i * n + j] != 0.0) { ti[nt] = i; tj[nt] = j; tx[nt] = P[(size_t)i * n + j]; nt++; }
    wco_csc* Pc = wco_csc_from_triplets(n, n, nt, ti, tj, tx);
    nt = 0;
    for (j = 0; j < n; j++)
        for (i = 0; i < m; i++)
            if (A[(size_t)i * n + j] != 0.0) { ti[nt] = i; tj[nt] = j; tx[nt] = A[(size_t)i * n + j]; nt++; }
    wco_csc* Ac = wco_csc_from_triplets(m, n, nt, ti, tj, tx);
    free(ti); free(tj); free(tx);
    wco_osqp* w = wco_osqp_setup(Pc, Ac, q, l, u, settings, NULL);
    wco_csc_free(Pc); wco_csc_free(Ac);
    if (!w) return -100;
    if (linsys_kind) wco_osqp_set_linsys(w, linsys_kind);
    int st = wco_osqp_solve(w);
    const double* sx = wco_osqp_solution_x(w);
    const double* sy = wco_osqp_solution_y(w);
    for (i = 0; i < n; i++) x_out[i] = sx[i];
    for (i = 0; i < m; i++) y_out[i] = sy[i];
    if (info_out) {
        info_out[0] = wco_osqp_iter(w);
        info_out[1] = wco_osqp_rho_updates(w);
        info_out[2] = wco_osqp_rho(w);
        info_out[3] = wco_osqp_pri_res(w);
        info_out[4] = wco_osqp_dua_res(w);
    }
    wco_osqp_free(w);
    return st;
}
