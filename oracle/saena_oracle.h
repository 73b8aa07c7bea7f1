/*
 * saena_oracle.h -- CPU restatement of the reference's AMG solve-phase hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as
 * the checker -- the product (saena_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement is checked in tests/test_oracle_vs_reference.py
 * against the UNMODIFIED reference compiled by oracle/Makefile (oracle/_ref/libsaena_ref.so)
 * on the same hierarchy in the same process, and against the golden vectors that run wrote
 * to tests/golden/ (generator: tests/golden/make_golden.py).  The reference itself ships no
 * golden vectors or asserting tests.
 *
 * All ranks of a row-partitioned run are emulated inside one process: every
 * distributed function takes the per-rank objects of all `nranks` ranks and
 * plays the MPI exchange with plain copies, in the reference's order.
 */
#ifndef SAENA_ORACLE_H
#define SAENA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* One rank's share of A, P or R in the reference layout
 * (saena_matrix.h:105-149, prolong_matrix.h:17-81, restrict_matrix.h:15-66). */
typedef struct so_operator {
    int M;                       /* local rows */
    int n_local_cols;            /* length of the local input vector */
    int col_offset;              /* split[rank] of the column partition */
    long nnz_local;
    const int *nnzPerRow_local;  /* [M] */
    const int *col_local;        /* [nnz_local] GLOBAL column ids */
    const double *val_local;     /* [nnz_local] */
    long nnz_remote;
    int col_remote_size;
    const int *row_remote;       /* [nnz_remote] local row ids, column-major by owner */
    const double *val_remote;    /* [nnz_remote] */
    const int *nnzPerCol_remote; /* [col_remote_size] */
    const long *nnzPerProcScan;  /* [nprocs+1] */
    int vIndexSize;
    const int *vIndex;           /* [vIndexSize] local ids to send */
    const int *vdispls;          /* [nprocs] */
    const int *rdispls;          /* [nprocs] */
    int numSendProc;
    const int *sendProcRank, *sendProcCount;
    int numRecvProc;
    const int *recvProcRank, *recvProcCount;
    int use_double;              /* 0: ghost values are cast to float (matvec_sparse_float) */
    int use_dense;               /* saena_matrix::use_dense: the product goes through saena_matrix_dense */
} so_operator;

/* (peer, offset, count) block of Grid::repart_u / repart_back_u (grid.cpp:99-163) */
typedef struct so_block {
    int peer, offset, count;
} so_block;

/* grids[l] on one rank (grid.h:11-78) */
typedef struct so_level {
    so_operator A, P, R;     /* P and R unused on the coarsest level */
    const double *inv_diag;  /* [A.M] */
    const double *inv_sq_diag; /* [A.M] inv_sq_diag_orig, NULL unless the hierarchy is scaled */
    double eig_max;          /* eig_max_of_invdiagXA */
    int M_coarse_old;        /* Ac.M_old */
    int M_coarse;            /* Ac.M */
    int n_repart_send, n_repart_recv;
    const so_block *repart_send, *repart_recv;
} so_level;

/* The hierarchy of all ranks: level[l * nranks + r]. */
typedef struct so_hierarchy {
    int nranks;
    int nlevels;             /* max_level + 1 */
    const so_level *level;
    int coarse_n;            /* rows of the coarsest operator (lives on rank 0) */
    const double *coarse_dense; /* [coarse_n * coarse_n] row-major dense copy of it */
    int scale;               /* saena_object::scale */
    int coarsest_cg;         /* saena_object::direct_solver == "CG" (default 0: "SuperLU") */
} so_hierarchy;

enum { SO_JACOBI = 0, SO_CHEBYSHEV = 1 };

/* w = Op v, distributed.  saena_matrix_matvec.cpp:9-113 / :448-550, prolong_matrix.cpp:489-758,
 * restrict_matrix.cpp:612-871.  ops[r], v[r], w[r] belong to rank r. */
void so_matvec(const so_operator *const *ops, int nranks, const double *const *v, double *const *w);

/* saena_matrix.tpp:16-23: res = A u - rhs */
void so_residual(const so_operator *const *ops, int nranks, const double *const *u, const double *const *rhs,
                 double *const *res);

/* saena_matrix.cpp:1044-1071 */
void so_jacobi(const so_level *const *lv, int nranks, int iter, double *const *u, const double *const *rhs);
/* saena_matrix.cpp:1074-1131 */
void so_chebyshev(const so_level *const *lv, int nranks, int iter, double *const *u, const double *const *rhs);

/* aux_functions.h:116-123: per-rank sequential sums, then summed over ranks in rank order */
double so_dot(int nranks, const int *M, const double *const *a, const double *const *b);

/* Dense LU with partial pivoting + one step of iterative refinement on the coarsest operator
 * (stands for SuperLU_DIST pdgssvx, saena_object_solve.cpp:793-958). */
void so_coarsest_solve(const so_hierarchy *h, const double *rhs, double *u);

/* saena_object_solve.cpp:14-114: CG on the coarsest level, <=150 iterations to 1e-12 (saena_object.h:155-156).
 * u is the initial guess on entry (the V-cycle passes 0). */
void so_coarsest_cg(const so_hierarchy *h, const double *rhs, double *u);

/* saena_object_solve.cpp:961-1431 starting at grid `l`; u[r], rhs[r] have length level[l].A.M */
void so_vcycle(const so_hierarchy *h, int l, int smoother, int pre, int post, double *const *u,
               double *const *rhs);

/* saena_object_solve.cpp:2389-2801.  u[r] (length level[0].A.M) receives the solution;
 * hist[0] = sqrt(<r0,r0>), hist[k] = sqrt(<r,r>) after iteration k.  Returns the iteration
 * count the reference prints (i+1). */
int so_solve_pcg(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                 int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len);

/* saena_object_solve.cpp:1883-2014: stationary V-cycle iteration, same stop rule */
int so_solve_vcycle(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                    int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len);

/* saena_object_solve.cpp:2017-2117: the smoother alone as a stationary iteration (`pre` sweeps per iteration; post unused) */
int so_solve_smoother(const so_hierarchy *h, const double *const *rhs, double *const *u, int max_iter, double tol,
                      int smoother, int pre, int post, double *hist, int hist_cap, int *hist_len);

#ifdef __cplusplus
}
#endif
#endif
