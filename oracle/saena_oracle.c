/*
 * saena_oracle.c -- CPU restatement of the reference's AMG solve-phase hot path.
 * TEST INFRASTRUCTURE (see saena_oracle.h for the rules and the parity status).
 *
 * Every function names the reference lines it follows; paths are relative to
 * /root/reference/.  The arithmetic order is the reference's: row-sequential
 * local sums accumulated into a temporary, remote contributions added
 * afterwards sender by sender (the reference takes them in MPI_Waitany arrival
 * order; the oracle fixes ascending sender rank), sequential dot products.
 */
#include "saena_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static double *dalloc(long n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

/* ------------------------------------------------------------------ matvec
 * src/saena_matrix_matvec.cpp:9-113 (double halo) and :448-550 (float halo);
 * prolong_matrix.cpp:489-758 and restrict_matrix.cpp:612-871 are the same loops. */
/* ------------------------------------------------------------------ dense matvec
 * src/saena_matrix_dense.cpp:181-260 (matvec_dense) and :262-340 (matvec_dense_float), selected by
 * saena_matrix::matvec when use_dense is set (include/saena_matrix.tpp:5-7).  The dense copy holds the same
 * entries (convert_saena_matrix, :763-793) and the product is taken block by block, the owner of the
 * block going round the ring from the rank itself (k = rank .. rank + nprocs - 1, :214-256): per block a
 * row sum over ascending columns into `tmp`, then w[i] += tmp.  The explicit zeros of the dense rows add
 * nothing, so the sums below run over the stored entries in the same order.  The float variant casts
 * the WHOLE vector to float before anything is multiplied (:281-282), the rank's own block included --
 * unlike matvec_sparse_float, which only casts what travels. */
static void so_matvec_dense(const so_operator *const *ops, int nranks, const double *const *v, double *const *w) {
    for (int r = 0; r < nranks; ++r) {
        const so_operator *A = ops[r];
        const int dbl = A->use_double;
        double *tmp = dalloc(A->M);
        for (int i = 0; i < A->M; ++i) w[r][i] = 0.0;
        for (int step = 0; step < nranks; ++step) {
            const int owner = (r + step) % nranks;
            if (owner == r) {
                const double *v_p = v[r] - A->col_offset;
                long iter = 0;
                for (int i = 0; i < A->M; ++i) {
                    const int jend = A->nnzPerRow_local[i];
                    double t = 0.0;
                    for (int j = 0; j < jend; ++j) {
                        const double x = v_p[A->col_local[iter + j]];
                        t += A->val_local[iter + j] * (dbl ? x : (double)(float)x);
                    }
                    w[r][i] += t;
                    iter += jend;
                }
                continue;
            }
            int k = -1;
            for (int q = 0; q < A->numRecvProc; ++q)
                if (A->recvProcRank[q] == owner) k = q;
            if (k < 0) continue;   /* nothing stored in that block: a block of zeros */
            for (int i = 0; i < A->M; ++i) tmp[i] = 0.0;
            long it = A->nnzPerProcScan[owner];
            const int *npc = A->nnzPerCol_remote + A->rdispls[owner];
            const int *ids = ops[owner]->vIndex + ops[owner]->vdispls[r];   /* the owner's local ids of my ghost columns */
            for (int j = 0; j < A->recvProcCount[k]; ++j) {
                const double x = v[owner][ids[j]];
                const double xr = dbl ? x : (double)(float)x;
                for (int i = 0; i < npc[j]; ++i) tmp[A->row_remote[it + i]] += A->val_remote[it + i] * xr;
                it += npc[j];
            }
            for (int i = 0; i < A->M; ++i) w[r][i] += tmp[i];
        }
        free(tmp);
    }
}

void so_matvec(const so_operator *const *ops, int nranks, const double *const *v, double *const *w) {
    for (int r = 0; r < nranks; ++r)
        if (ops[r]->use_dense) { so_matvec_dense(ops, nranks, v, w); return; }
    /* :25-26 / :463-464  pack vSend (as float when !use_double) */
    double **vSend = (double **)calloc((size_t)nranks, sizeof(double *));
    for (int r = 0; r < nranks; ++r) {
        const so_operator *A = ops[r];
        vSend[r] = dalloc(A->vIndexSize);
        for (int i = 0; i < A->vIndexSize; ++i) {
            double x = v[r][A->vIndex[i]];
            vSend[r][i] = A->use_double ? x : (double)(float)x;
        }
    }
    for (int r = 0; r < nranks; ++r) {
        const so_operator *A = ops[r];
        /* :32-41  the exchange: receiver r gets, from each sender s, the slice of s's vSend that
         * starts at s.vdispls[r]; it lands at r.rdispls[s] in vecValues. */
        double *vecValues = dalloc(A->col_remote_size);
        for (int k = 0; k < A->numRecvProc; ++k) {
            const int s = A->recvProcRank[k];
            memcpy(vecValues + A->rdispls[s], vSend[s] + ops[s]->vdispls[r],
                   sizeof(double) * (size_t)A->recvProcCount[k]);
        }
        /* :44  w = 0 */
        for (int i = 0; i < A->M; ++i) w[r][i] = 0.0;
        /* :55-80  local loop; v_p = v - split[rank], col_local holds global ids */
        const double *v_p = v[r] - A->col_offset;
        long iter = 0;
        for (int i = 0; i < A->M; ++i) {
            const int jend = A->nnzPerRow_local[i];
            double tmp = 0.0;
            for (int j = 0; j < jend; ++j) tmp += A->val_local[iter + j] * v_p[A->col_local[iter + j]];
            w[r][i] += tmp;
            iter += jend;
        }
        /* :87-110  remote loop, one sender at a time, column by column */
        for (int k = 0; k < A->numRecvProc; ++k) {
            const int s = A->recvProcRank[k];
            long it = A->nnzPerProcScan[s];
            const double *vv = vecValues + A->rdispls[s];
            const int *npc = A->nnzPerCol_remote + A->rdispls[s];
            for (int j = 0; j < A->recvProcCount[k]; ++j) {
                const double vrem = vv[j];
                for (int i = 0; i < npc[j]; ++i) w[r][A->row_remote[it + i]] += A->val_remote[it + i] * vrem;
                it += npc[j];
            }
        }
        free(vecValues);
    }
    for (int r = 0; r < nranks; ++r) free(vSend[r]);
    free(vSend);
}

/* include/saena_matrix.tpp:16-23 */
void so_residual(const so_operator *const *ops, int nranks, const double *const *u, const double *const *rhs,
                 double *const *res) {
    so_matvec(ops, nranks, u, res);
    for (int r = 0; r < nranks; ++r)
        for (int i = 0; i < ops[r]->M; ++i) res[r][i] -= rhs[r][i];
}

/* include/saena_matrix.tpp:35-43: res = c * w .* (rhs - A u) */
static void residual_multiply(const so_level *const *lv, int nranks, const double *const *u,
                              const double *const *rhs, double *const *res, double c) {
    const so_operator **ops = (const so_operator **)calloc((size_t)nranks, sizeof(void *));
    for (int r = 0; r < nranks; ++r) ops[r] = &lv[r]->A;
    so_matvec(ops, nranks, u, res);
    for (int r = 0; r < nranks; ++r)
        for (int i = 0; i < ops[r]->M; ++i) res[r][i] = c * lv[r]->inv_diag[i] * (rhs[r][i] - res[r][i]);
    free(ops);
}

/* src/saena_matrix.cpp:1044-1071; omega is float(2.0/3) promoted (saena_matrix.h:182) */
void so_jacobi(const so_level *const *lv, int nranks, int iter, double *const *u, const double *const *rhs) {
    const float omega = (float)(2.0 / 3);
    const so_operator **ops = (const so_operator **)calloc((size_t)nranks, sizeof(void *));
    double **temp1 = (double **)calloc((size_t)nranks, sizeof(double *));
    for (int r = 0; r < nranks; ++r) { ops[r] = &lv[r]->A; temp1[r] = dalloc(ops[r]->M); }
    for (int j = 0; j < iter; ++j) {
        so_matvec(ops, nranks, (const double *const *)u, temp1);
        for (int r = 0; r < nranks; ++r)
            for (int i = 0; i < ops[r]->M; ++i) {
                temp1[r][i] -= rhs[r][i];
                temp1[r][i] *= lv[r]->inv_diag[i] * omega;
                u[r][i] -= temp1[r][i];
            }
    }
    for (int r = 0; r < nranks; ++r) free(temp1[r]);
    free(temp1);
    free(ops);
}

/* src/saena_matrix.cpp:1074-1131 */
void so_chebyshev(const so_level *const *lv, int nranks, int iter, double *const *u, const double *const *rhs) {
    const double eig = lv[0]->eig_max;
    const double alpha = 0.13 * eig;
    const double beta = eig;
    const double delta = (beta - alpha) / 2.0;
    const double theta = (beta + alpha) / 2.0;
    const double s1 = theta / delta;
    const double twos1 = 2.0 * s1;
    double rhok = 1.0 / s1;
    double rhokp1, two_rhokp1, d1, d2;

    double **res = (double **)calloc((size_t)nranks, sizeof(double *));
    double **d = (double **)calloc((size_t)nranks, sizeof(double *));
    for (int r = 0; r < nranks; ++r) { res[r] = dalloc(lv[r]->A.M); d[r] = dalloc(lv[r]->A.M); }

    /* :1099-1109 first sweep */
    residual_multiply(lv, nranks, (const double *const *)u, rhs, d, 1.0 / theta);
    for (int r = 0; r < nranks; ++r)
        for (int i = 0; i < lv[r]->A.M; ++i) u[r][i] += d[r][i];

    /* :1111-1130 */
    for (int k = 1; k < iter; ++k) {
        rhokp1 = 1.0 / (twos1 - rhok);
        two_rhokp1 = 2.0 * rhokp1;
        d1 = rhokp1 * rhok;
        d2 = two_rhokp1 / delta;
        rhok = rhokp1;
        residual_multiply(lv, nranks, (const double *const *)u, rhs, res, d2);
        for (int r = 0; r < nranks; ++r)
            for (int j = 0; j < lv[r]->A.M; ++j) {
                d[r][j] = (d1 * d[r][j]) + res[r][j];
                u[r][j] += d[r][j];
            }
    }
    for (int r = 0; r < nranks; ++r) { free(res[r]); free(d[r]); }
    free(res);
    free(d);
}

/* include/aux_functions.h:116-123 */
double so_dot(int nranks, const int *M, const double *const *a, const double *const *b) {
    double dot = 0.0;
    for (int r = 0; r < nranks; ++r) {
        double dot_l = 0.0;
        for (int i = 0; i < M[r]; ++i) dot_l += a[r][i] * b[r][i];
        dot += dot_l; /* MPI_Allreduce(SUM): rank order here */
    }
    return dot;
}

/* ------------------------------------------------------------------ coarsest solve
 * src/saena_object_solve.cpp:793-958 hands the system to SuperLU_DIST pdgssvx (LU with
 * equilibration, MC64 row permutation, NATURAL column order and double iterative refinement,
 * external/SuperLU_DIST_5.4.0/SRC/util.c:321-341).  Restated as dense LU with partial pivoting
 * plus refinement steps until the correction stalls; both give the exact solution of a <=~100
 * row system to rounding. */
void so_coarsest_solve(const so_hierarchy *h, const double *rhs, double *u) {
    const int n = h->coarse_n;
    double *LU = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    int *piv = (int *)malloc(sizeof(int) * (size_t)n);
    memcpy(LU, h->coarse_dense, sizeof(double) * (size_t)n * (size_t)n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(LU[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(LU[(size_t)i * n + k]) > best) { best = fabs(LU[(size_t)i * n + k]); p = i; }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double t = LU[(size_t)k * n + j];
                LU[(size_t)k * n + j] = LU[(size_t)p * n + j];
                LU[(size_t)p * n + j] = t;
            }
        const double dkk = LU[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const double l = LU[(size_t)i * n + k] / dkk;
            LU[(size_t)i * n + k] = l;
            for (int j = k + 1; j < n; ++j) LU[(size_t)i * n + j] -= l * LU[(size_t)k * n + j];
        }
    }
    double *x = dalloc(n), *r = dalloc(n), *c = dalloc(n);
    for (int pass = 0; pass < 4; ++pass) {
        /* residual of the current iterate (x = 0 on the first pass) */
        for (int i = 0; i < n; ++i) {
            double s = rhs[i];
            for (int j = 0; j < n; ++j) s -= h->coarse_dense[(size_t)i * n + j] * x[j];
            r[i] = s;
        }
        memcpy(c, r, sizeof(double) * (size_t)n);
        for (int k = 0; k < n; ++k)
            if (piv[k] != k) { double t = c[k]; c[k] = c[piv[k]]; c[piv[k]] = t; }
        for (int i = 0; i < n; ++i) {
            double s = c[i];
            for (int j = 0; j < i; ++j) s -= LU[(size_t)i * n + j] * c[j];
            c[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = c[i];
            for (int j = i + 1; j < n; ++j) s -= LU[(size_t)i * n + j] * c[j];
            c[i] = s / LU[(size_t)i * n + i];
        }
        for (int i = 0; i < n; ++i) x[i] += c[i];
    }
    memcpy(u, x, sizeof(double) * (size_t)n);
    free(x); free(r); free(c); free(LU); free(piv);
}

/* src/saena_object_solve.cpp:14-114 (the coarsest level lives on rank 0) */
void so_coarsest_cg(const so_hierarchy *h, const double *rhs, double *u) {
    const double CG_coarsest_tol = 1e-12; /* saena_object.h:156 */
    const int CG_coarsest_max_iter = 150; /* saena_object.h:155 */
    const int n = h->nranks;
    const so_operator *A = &h->level[(h->nlevels - 1) * n].A;
    const so_operator *ops[1] = {A};
    const int sz = A->M;
    double *res = dalloc(sz), *dir = dalloc(sz), *mv = dalloc(sz);
    memcpy(res, rhs, sizeof(double) * (size_t)sz);
    const double *rp[1] = {res};
    const int M1[1] = {sz};
    const double initial_dot = so_dot(1, M1, rp, rp);
    const double thres = initial_dot * CG_coarsest_tol * CG_coarsest_tol;
    double dot = initial_dot, dot_prev, factor;
    int max_iter = CG_coarsest_max_iter;
    if (dot < CG_coarsest_tol * CG_coarsest_tol) max_iter = 0;
    memcpy(dir, res, sizeof(double) * (size_t)sz);
    int i = 1;
    while (i < max_iter) {
        const double *dp[1] = {dir};
        double *mp[1] = {mv};
        so_matvec(ops, 1, dp, mp);
        const double *mcp[1] = {mv};
        factor = so_dot(1, M1, dp, mcp);
        factor = dot / factor;
        for (int j = 0; j < sz; ++j) {
            u[j] += factor * dir[j];
            res[j] -= factor * mv[j];
        }
        dot_prev = dot;
        dot = so_dot(1, M1, rp, rp);
        if (dot < thres) break;
        factor = dot / dot_prev;
        for (int j = 0; j < sz; ++j) dir[j] = res[j] + factor * dir[j];
        i++;
    }
    free(res); free(dir); free(mv);
}

/* ------------------------------------------------------------------ repartition
 * src/grid.cpp:99-130 (repart_u): blocks of the old-partition vector go to the ranks owning
 * them in the new partition.  `back` runs the plan in reverse (grid.cpp:132-163). */
static void repart(const so_hierarchy *h, int l, int back, double *const *src, double *const *dst) {
    const int n = h->nranks;
    for (int r = 0; r < n; ++r) {
        const so_level *lv = &h->level[l * n + r];
        const so_block *send = back ? lv->repart_recv : lv->repart_send;
        const int ns = back ? lv->n_repart_recv : lv->n_repart_send;
        for (int k = 0; k < ns; ++k) {
            const int peer = send[k].peer;
            const so_level *pl = &h->level[l * n + peer];
            const so_block *precv = back ? pl->repart_send : pl->repart_recv;
            const int nr = back ? pl->n_repart_send : pl->n_repart_recv;
            for (int q = 0; q < nr; ++q)
                if (precv[q].peer == r) {
                    memcpy(dst[peer] + precv[q].offset, src[r] + send[k].offset,
                           sizeof(double) * (size_t)send[k].count);
                    break;
                }
        }
    }
}

static int has_repart(const so_hierarchy *h, int l) {
    for (int r = 0; r < h->nranks; ++r)
        if (h->level[l * h->nranks + r].n_repart_send || h->level[l * h->nranks + r].n_repart_recv) return 1;
    return 0;
}

/* ------------------------------------------------------------------ V-cycle
 * src/saena_object_solve.cpp:961-1431 */
static void smooth(const so_hierarchy *h, int l, int smoother, int iters, double *const *u,
                   const double *const *rhs) {
    const int n = h->nranks;
    const so_level **lv = (const so_level **)calloc((size_t)n, sizeof(void *));
    for (int r = 0; r < n; ++r) lv[r] = &h->level[l * n + r];
    /* include/saena_object.tpp:85-96 */
    if (smoother == SO_CHEBYSHEV) so_chebyshev(lv, n, iters, u, rhs);
    else so_jacobi(lv, n, iters, u, rhs);
    free(lv);
}

void so_vcycle(const so_hierarchy *h, int l, int smoother, int pre, int post, double *const *u,
               double *const *rhs) {
    const int n = h->nranks;
    /* :991-1057 coarsest level: direct solve (lives on rank 0) */
    if (l == h->nlevels - 1) {
        if (h->coarsest_cg) so_coarsest_cg(h, rhs[0], u[0]);
        else so_coarsest_solve(h, rhs[0], u[0]);
        return;
    }
    const so_operator **A = (const so_operator **)calloc((size_t)n, sizeof(void *));
    const so_operator **P = (const so_operator **)calloc((size_t)n, sizeof(void *));
    const so_operator **R = (const so_operator **)calloc((size_t)n, sizeof(void *));
    double **res = (double **)calloc((size_t)n, sizeof(double *));
    double **uCorr = (double **)calloc((size_t)n, sizeof(double *));
    double **res_coarse = (double **)calloc((size_t)n, sizeof(double *));
    double **res_coarse_new = (double **)calloc((size_t)n, sizeof(double *));
    double **uCorrCoarse = (double **)calloc((size_t)n, sizeof(double *));
    double **uCorrCoarse_old = (double **)calloc((size_t)n, sizeof(double *));
    for (int r = 0; r < n; ++r) {
        const so_level *lv = &h->level[l * n + r];
        A[r] = &lv->A; P[r] = &lv->P; R[r] = &lv->R;
        res[r] = dalloc(lv->A.M);
        uCorr[r] = dalloc(lv->A.M);
        res_coarse[r] = dalloc(lv->M_coarse_old);
        res_coarse_new[r] = dalloc(lv->M_coarse);
        uCorrCoarse[r] = dalloc(lv->M_coarse);
        uCorrCoarse_old[r] = dalloc(lv->M_coarse_old);
    }
    /* :1105-1107 pre-smooth */
    if (pre) smooth(h, l, smoother, pre, u, (const double *const *)rhs);
    /* :1140 residual */
    so_residual(A, n, (const double *const *)u, (const double *const *)rhs, res);
    /* :1175 restrict */
    so_matvec(R, n, (const double *const *)res, res_coarse);
    /* :1201-1203 repart_u, :1249 uCorrCoarse = 0, :1256 recurse */
    const int rp = has_repart(h, l);
    if (rp) repart(h, l, 0, res_coarse, res_coarse_new);
    double **rc = rp ? res_coarse_new : res_coarse;
    /* :1245-1247 scale_vector(res_coarse, coarse inv_sq_diag_orig) */
    if (h->scale)
        for (int r = 0; r < n; ++r) {
            const so_level *cl = &h->level[(l + 1) * n + r];
            for (int i = 0; i < cl->A.M; ++i) rc[r][i] *= cl->inv_sq_diag[i];
        }
    so_vcycle(h, l + 1, smoother, pre, post, uCorrCoarse, rc);
    /* :1264-1266 */
    if (h->scale)
        for (int r = 0; r < n; ++r) {
            const so_level *cl = &h->level[(l + 1) * n + r];
            for (int i = 0; i < cl->A.M; ++i) uCorrCoarse[r][i] *= cl->inv_sq_diag[i];
        }
    /* :1301-1303 repart_back_u */
    if (rp) repart(h, l, 1, uCorrCoarse, uCorrCoarse_old);
    /* :1325 prolong, :1360-1361 correct */
    so_matvec(P, n, (const double *const *)(rp ? uCorrCoarse_old : uCorrCoarse), uCorr);
    for (int r = 0; r < n; ++r)
        for (int i = 0; i < A[r]->M; ++i) u[r][i] -= uCorr[r][i];
    /* :1397-1399 post-smooth */
    if (post) smooth(h, l, smoother, post, u, (const double *const *)rhs);

    for (int r = 0; r < n; ++r) {
        free(res[r]); free(uCorr[r]); free(res_coarse[r]); free(res_coarse_new[r]);
        free(uCorrCoarse[r]); free(uCorrCoarse_old[r]);
    }
    free(res); free(uCorr); free(res_coarse); free(res_coarse_new); free(uCorrCoarse); free(uCorrCoarse_old);
    free(A); free(P); free(R);
}

/* ------------------------------------------------------------------ PCG
 * src/saena_object_solve.cpp:2389-2801 */
int so_solve_pcg(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                 int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len) {
    const int n = h->nranks;
    const so_operator **A = (const so_operator **)calloc((size_t)n, sizeof(void *));
    int *M = (int *)calloc((size_t)n, sizeof(int));
    double **r = (double **)calloc((size_t)n, sizeof(double *));
    double **rho = (double **)calloc((size_t)n, sizeof(double *));
    double **p = (double **)calloc((size_t)n, sizeof(double *));
    double **hh = (double **)calloc((size_t)n, sizeof(double *));
    for (int k = 0; k < n; ++k) {
        A[k] = &h->level[k].A;
        M[k] = A[k]->M;
        r[k] = dalloc(M[k]); rho[k] = dalloc(M[k]); p[k] = dalloc(M[k]); hh[k] = dalloc(M[k]);
        for (int i = 0; i < M[k]; ++i) u[k][i] = 0.0; /* :2482 */
    }
    int nh = 0;
    /* :2496-2501 */
    so_residual(A, n, (const double *const *)u, rhs, r);
    const double init_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
    double current_dot = init_dot;
    if (nh < hist_cap) hist[nh] = sqrt(init_dot);
    ++nh;
    int i = 0;
    if (h->nlevels == 1) {
        /* :2507-2521 max_level == 0: direct solver only */
        so_vcycle(h, 0, smoother, pre, post, u, (double *const *)rhs);
        so_residual(A, n, (const double *const *)u, rhs, r);
        current_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
        if (nh < hist_cap) hist[nh] = sqrt(current_dot);
        ++nh;
        i = 0;
        goto done;
    }
    /* :2535-2537 rho = 0; vcycle(rho, r) */
    so_vcycle(h, 0, smoother, pre, post, rho, r);
    /* :2554 p = rho */
    for (int k = 0; k < n; ++k) memcpy(p[k], rho[k], sizeof(double) * (size_t)M[k]);
    const double THRSHLD = init_dot * tol * tol; /* :2558 */
    for (i = 0; i < max_iter; i++) {
        so_matvec(A, n, (const double *const *)p, hh);                                      /* :2571 */
        const double rho_res = so_dot(n, M, (const double *const *)r, (const double *const *)rho); /* :2580 */
        const double pdoth = so_dot(n, M, (const double *const *)p, (const double *const *)hh);    /* :2581 */
        const double alpha = rho_res / pdoth;                                               /* :2588 */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < M[k]; ++j) {                                                /* :2593-2596 */
                u[k][j] -= alpha * p[k][j];
                r[k][j] -= alpha * hh[k][j];
            }
        current_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);     /* :2603 */
        if (nh < hist_cap) hist[nh] = sqrt(current_dot);
        ++nh;
        if (current_dot < THRSHLD) break;                                                   /* :2620 */
        for (int k = 0; k < n; ++k) memset(rho[k], 0, sizeof(double) * (size_t)M[k]);       /* :2640 */
        so_vcycle(h, 0, smoother, pre, post, rho, r);                                       /* :2641 */
        double beta = so_dot(n, M, (const double *const *)r, (const double *const *)rho);   /* :2655 */
        beta /= rho_res;                                                                    /* :2662 */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < M[k]; ++j) p[k][j] = rho[k][j] + beta * p[k][j];            /* :2665-2667 */
    }
    if (i == max_iter) i--; /* :2673-2674 */
done:
    if (h->scale) /* :2709-2711 */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < M[k]; ++j) u[k][j] *= h->level[k].inv_sq_diag[j];
    *hist_len = nh;
    for (int k = 0; k < n; ++k) { free(r[k]); free(rho[k]); free(p[k]); free(hh[k]); }
    free(r); free(rho); free(p); free(hh); free(M); free(A);
    return i + 1; /* :2678-2682 prints i+1 */
}

/* src/saena_object_solve.cpp:1883-2014 */
int so_solve_vcycle(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                    int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len) {
    const int n = h->nranks;
    const so_operator **A = (const so_operator **)calloc((size_t)n, sizeof(void *));
    int *M = (int *)calloc((size_t)n, sizeof(int));
    double **r = (double **)calloc((size_t)n, sizeof(double *));
    for (int k = 0; k < n; ++k) {
        A[k] = &h->level[k].A;
        M[k] = A[k]->M;
        r[k] = dalloc(M[k]);
        for (int i = 0; i < M[k]; ++i) u[k][i] = 0.0;
    }
    int nh = 0;
    so_residual(A, n, (const double *const *)u, rhs, r);
    const double init_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
    double current_dot = init_dot;
    if (nh < hist_cap) hist[nh] = sqrt(init_dot);
    ++nh;
    const double THRSHLD = init_dot * tol * tol;
    int i = 0;
    for (; i < max_iter; ++i) {
        so_vcycle(h, 0, smoother, pre, post, u, (double *const *)rhs);
        so_residual(A, n, (const double *const *)u, rhs, r);
        current_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
        if (nh < hist_cap) hist[nh] = sqrt(current_dot);
        ++nh;
        if (current_dot < THRSHLD) break;
    }
    if (i == max_iter) --i;
    if (h->scale) /* :2000-2002 */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < M[k]; ++j) u[k][j] *= h->level[k].inv_sq_diag[j];
    *hist_len = nh;
    for (int k = 0; k < n; ++k) free(r[k]);
    free(r); free(M); free(A);
    return i + 1;
}

/* src/saena_object_solve.cpp:2017-2117 -- saena_object::solve_smoother: the smoother alone as a stationary iteration,
 * `pre` sweeps per iteration (:2074), residual and <r,r> after each (:2075-2076), same stop rule (:2080) */
int so_solve_smoother(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                      int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len) {
    (void)post;
    const int n = h->nranks;
    const so_operator **A = (const so_operator **)calloc((size_t)n, sizeof(void *));
    const so_level **lv = (const so_level **)calloc((size_t)n, sizeof(void *));
    int *M = (int *)calloc((size_t)n, sizeof(int));
    double **r = (double **)calloc((size_t)n, sizeof(double *));
    for (int k = 0; k < n; ++k) {
        lv[k] = &h->level[k];
        A[k] = &h->level[k].A;
        M[k] = A[k]->M;
        r[k] = dalloc(M[k]);
        for (int i = 0; i < M[k]; ++i) u[k][i] = 0.0;                                        /* :2051 */
    }
    int nh = 0;
    so_residual(A, n, (const double *const *)u, rhs, r);                                    /* :2061 */
    const double init_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
    double current_dot = init_dot;
    if (nh < hist_cap) hist[nh] = sqrt(init_dot);
    ++nh;
    const double THRSHLD = init_dot * tol * tol;                                            /* :2069 */
    int i = 0;
    for (; i < max_iter; ++i) {
        if (smoother) so_chebyshev(lv, n, pre, u, rhs);                                     /* :2074, saena_object.tpp:85-96 */
        else so_jacobi(lv, n, pre, u, rhs);
        so_residual(A, n, (const double *const *)u, rhs, r);
        current_dot = so_dot(n, M, (const double *const *)r, (const double *const *)r);
        if (nh < hist_cap) hist[nh] = sqrt(current_dot);
        ++nh;
        if (current_dot < THRSHLD) break;                                                   /* :2080 */
    }
    if (i == max_iter) --i;                                                                 /* :2086-2087 */
    if (h->scale)                                                                           /* :2100-2102 */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < M[k]; ++j) u[k][j] *= h->level[k].inv_sq_diag[j];
    *hist_len = nh;
    for (int k = 0; k < n; ++k) free(r[k]);
    free(r); free(M); free(A); free(lv);
    return i + 1;
}
