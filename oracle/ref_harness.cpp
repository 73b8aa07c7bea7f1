/*
 * ref_harness.cpp -- C entry points around the UNMODIFIED reference library
 * (TEST INFRASTRUCTURE; linked into oracle/_ref/libsaena_ref.so by
 * oracle/Makefile, never into the product).
 *
 * It drives the reference exactly as experiments/Poisson.cpp does
 * (/root/reference/experiments/Poisson.cpp:81-246: laplacian3D -> assemble ->
 * rhs -> amg::set_matrix -> amg::set_rhs -> amg::solve_pCG) on one MPI rank,
 * and exposes
 *   - the finished hierarchy (the arrays saena_matrix::set_off_on_diagonal,
 *     prolong_matrix::findLocalRemote and restrict_matrix::transposeP built),
 *     so tests can hand the same hierarchy to the CUDA path and to the C oracle;
 *   - the reference's own hot-path functions (matvec, chebyshev, jacobi,
 *     residual, R/P matvec, vcycle, solve_pCG with a residual history taken by
 *     ref_shim/ref_hooks.h) as the parity oracle;
 *   - the reference CPU solve as the bench's cpu_baseline ("kind": "reference").
 */
#define SAENA_REF_HARNESS_TU
#include "ref_shim/ref_hooks.h"

#include "saena.hpp"
#include "saena_object.h"
#include "saena_matrix.h"
#include "grid.h"
#include "aux_functions2.h"
#include "lambda_lanczos.hpp"   // the reference's own Lanczos engine (external/lambda_lanczos), templates only

#include <cstdio>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#include <vector>

// ---------------------------------------------------------------- residual history
static std::vector<double> g_rr;
static int g_rr_len = -1;  // only <r,r> of vectors of this length are recorded (level-0 size)
extern "C" void saena_ref_record_rr(double rr, int sz) { if (g_rr_len < 0 || sz == g_rr_len) g_rr.push_back(rr); }

namespace {

struct Handle {
    saena::matrix *A = nullptr;
    saena::vector *rhs = nullptr;
    saena::options *opts = nullptr;
    saena::amg *solver = nullptr;
    bool vcycle_mem = false;
};

// The reference prints its setup report unconditionally; keep test logs readable.
struct QuietStdout {
    int saved = -1;
    explicit QuietStdout(bool on) {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        dup2(devnull, 1);
        close(devnull);
    }
    ~QuietStdout() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

void ensure_mpi() {
    int inited = 0;
    MPI_Initialized(&inited);
    if (!inited) MPI_Init(nullptr, nullptr);
}

saena_object *obj(Handle *h) { return h->solver->get_object(); }

void finish_setup(Handle *h, saena::vector *rhs, bool quiet, bool scale = false) {
    QuietStdout q(quiet);
    h->solver = new saena::amg();
    h->solver->set_scale(scale);  // Poisson.cpp:41,193 run with scale = false
    h->solver->set_matrix(h->A, h->opts);
    h->solver->set_rhs(*rhs);
}

}  // namespace

extern "C" {

struct sref_opts {
    int max_iter;
    double tol;
    const char *smoother;
    int pre, post;
    const char *psmoother;
    float conn_str;
    int dynamic_levels, max_level, float_level;
    double filter_thre, filter_max;
    int filter_start, filter_rate;
};

// switch_to_dense / dense_thre / dense_sz_thre (data/options006_poisson.xml: 0 / 0.1 / 5000) come from the
// environment so that the options struct the Python side mirrors keeps its layout:
// SREF_SWITCH_TO_DENSE=1 [SREF_DENSE_THRE=0.1] [SREF_DENSE_SZ_THRE=5000]
static saena::options *make_opts(const sref_opts *o) {
    const char *sd = getenv("SREF_SWITCH_TO_DENSE"), *dt = getenv("SREF_DENSE_THRE"), *ds = getenv("SREF_DENSE_SZ_THRE");
    return new saena::options(o->max_iter, o->tol, o->smoother, o->pre, o->post, o->psmoother, o->conn_str,
                              o->dynamic_levels != 0, o->max_level, o->float_level, o->filter_thre,
                              o->filter_max, o->filter_start, o->filter_rate, sd && atoi(sd) != 0,
                              dt ? (float)atof(dt) : 0.1f, ds ? atoi(ds) : 5000);
}

void *sref_poisson_new_scaled(int mx, const sref_opts *o, int quiet, int scale);

// 3D 7-point Poisson, exactly the driver sequence of experiments/Poisson.cpp.
void *sref_poisson_new(int mx, const sref_opts *o, int quiet) { return sref_poisson_new_scaled(mx, o, quiet, 0); }

// ... with the `scale` switch of Poisson.cpp:41 exposed (A.assemble(scale), solver.set_scale(scale))
void *sref_poisson_new_scaled(int mx, const sref_opts *o, int quiet, int scale) {
    ensure_mpi();
    MPI_Comm comm = MPI_COMM_WORLD;
    Handle *h = new Handle();
    {
        QuietStdout q(quiet != 0);
        h->A = new saena::matrix(comm);
        saena::laplacian3D(h->A, mx, mx, mx);
        h->A->set_remove_boundary(true);
        h->A->assemble(scale != 0);
    }
    value_t *rhs_std = nullptr;
    index_t orig_sz = saena::laplacian3D_set_rhs(rhs_std, mx, mx, mx, comm);
    index_t my_split = 0;
    saena::find_split(orig_sz, my_split, comm);
    h->rhs = new saena::vector(comm);
    h->rhs->set(&rhs_std[0], orig_sz, my_split);
    h->rhs->assemble();
    h->opts = make_opts(o);
    finish_setup(h, h->rhs, quiet != 0, scale != 0);
    saena_free(rhs_std);
    return h;
}

// Generic square matrix from COO triplets (global ids) + dense rhs; boundary removal off.
void *sref_coo_new(int n, long nnz, const int *row, const int *col, const double *val, const double *rhs,
                   const sref_opts *o, int quiet) {
    ensure_mpi();
    MPI_Comm comm = MPI_COMM_WORLD;
    Handle *h = new Handle();
    {
        QuietStdout q(quiet != 0);
        h->A = new saena::matrix(comm);
        for (long i = 0; i < nnz; ++i) h->A->set(row[i], col[i], val[i]);
        h->A->set_remove_boundary(false);
        h->A->assemble(false);
    }
    h->rhs = new saena::vector(comm);
    h->rhs->set(rhs, n, 0);
    h->rhs->assemble();
    h->opts = make_opts(o);
    finish_setup(h, h->rhs, quiet != 0);
    return h;
}

// What a lazy update does to the solver object -- grids[0].A replaced by a matrix of the same pattern and other
// values, the coarse operators kept (saena_object::update1, src/saena_object_lazy.cpp:7-35).  In this version of
// the reference the bodies of update1/2/3 are compiled out (#if 0), so the replacement is done here directly;
// the point is the drop-in's reaction (tests/test_adaptor_multirank.py).  The new matrix is kept alive.
int sref_replace_A0_scaled_poisson(void *hv, int mx, double factor) {
    Handle *h = (Handle *)hv;
    MPI_Comm comm = MPI_COMM_WORLD;
    QuietStdout q(true);
    saena::matrix *A2 = new saena::matrix(comm);
    saena::laplacian3D(A2, mx, mx, mx);
    A2->set_remove_boundary(true);
    A2->assemble(false);
    saena_matrix *m = A2->get_internal_matrix();
    for (nnz_t i = 0; i < m->nnz_l_local; ++i) m->val_local[i] *= factor;
    for (nnz_t i = 0; i < m->nnz_l_remote; ++i) m->val_remote[i] *= factor;
    for (index_t i = 0; i < m->M; ++i) m->inv_diag[i] /= factor;
    for (auto &e : m->entry) e.val *= factor;
    saena_object *o = obj(h);
    m->eig_max_of_invdiagXA = o->grids[0].A->eig_max_of_invdiagXA;   // as update1 does
    m->use_double = o->grids[0].A->use_double;
    o->grids[0].A = m;
    return 0;
}

// multi-rank runs (oracle/mprun.py): who am I, and a clean shutdown of the MPI stand-in
int sref_rank() { ensure_mpi(); int r = 0; MPI_Comm_rank(MPI_COMM_WORLD, &r); return r; }
int sref_size() { ensure_mpi(); int s = 1; MPI_Comm_size(MPI_COMM_WORLD, &s); return s; }
void sref_barrier() { ensure_mpi(); MPI_Barrier(MPI_COMM_WORLD); }
double sref_wtime() { return MPI_Wtime(); }
void sref_finalize() { int inited = 0; MPI_Initialized(&inited); if (inited) MPI_Finalize(); }

// Several ranks: every rank contributes its own entries (global ids) and a contiguous block of the rhs
// starting at rhs_offset (saena::vector::set(values, size, offset)); assemble() repartitions.
void *sref_coo_new_part(long nnz, const int *row, const int *col, const double *val, int rhs_n, int rhs_offset,
                        const double *rhs, const sref_opts *o, int quiet) {
    ensure_mpi();
    MPI_Comm comm = MPI_COMM_WORLD;
    Handle *h = new Handle();
    {
        QuietStdout q(quiet != 0);
        h->A = new saena::matrix(comm);
        for (long i = 0; i < nnz; ++i) h->A->set(row[i], col[i], val[i]);
        h->A->set_remove_boundary(false);
        h->A->assemble(false);
    }
    h->rhs = new saena::vector(comm);
    h->rhs->set(rhs, rhs_n, rhs_offset);
    h->rhs->assemble();
    h->opts = make_opts(o);
    finish_setup(h, h->rhs, quiet != 0);
    return h;
}

void sref_free(void *hv) {
    Handle *h = (Handle *)hv;
    if (!h) return;
    if (h->vcycle_mem) obj(h)->free_vcycle_memory();
    h->solver->destroy();
    h->A->destroy();
    delete h->solver;
    delete h->rhs;
    delete h->opts;
    delete h->A;
    delete h;
}

// number of grids = max_level + 1; operators P/R/Ac exist on levels < max_level
int sref_max_level(void *hv) { return obj((Handle *)hv)->max_level; }
int sref_scale(void *hv) { return obj((Handle *)hv)->scale ? 1 : 0; }

struct sref_level_info {
    int M, Mbig, Nbig;
    long nnz_l, nnz_local, nnz_remote;
    int col_remote_size, vIndexSize, recvSize, numRecvProc, numSendProc;
    int use_double, active;
    double eig_max;
    int use_dense;   // saena_matrix::use_dense: matvec goes through saena_matrix_dense (kind 0 only)
};

static saena_matrix *level_A(Handle *h, int l) { return obj(h)->grids[l].A; }

// kind: 0 = A_l, 1 = P_l (fine rows x coarse cols), 2 = R_l (coarse rows x fine cols)
int sref_level_info_get(void *hv, int l, int kind, sref_level_info *out) {
    Handle *h = (Handle *)hv;
    memset(out, 0, sizeof(*out));
    if (l < 0 || l > obj(h)->max_level) return 1;
    Grid &g = obj(h)->grids[l];
    // a rank that a shrink left out of this level's communicator owns nothing of it
    if (!g.A || !g.A->active || (kind != 0 && !g.active)) { out->use_double = 1; return 0; }
    if (kind == 0) {
        saena_matrix *A = g.A;
        out->M = A->M; out->Mbig = A->Mbig; out->Nbig = A->Mbig;
        out->nnz_l = A->nnz_l; out->nnz_local = A->nnz_l_local; out->nnz_remote = A->nnz_l_remote;
        out->col_remote_size = A->col_remote_size; out->vIndexSize = A->vIndexSize; out->recvSize = A->recvSize;
        out->numRecvProc = A->numRecvProc; out->numSendProc = A->numSendProc;
        out->use_double = A->use_double; out->active = A->active; out->eig_max = A->eig_max_of_invdiagXA;
        out->use_dense = A->use_dense ? 1 : 0;
    } else if (l >= obj(h)->max_level) {
        return 1;
    } else if (kind == 1) {
        prolong_matrix &P = g.P;
        out->M = P.M; out->Mbig = P.Mbig; out->Nbig = P.Nbig;
        out->nnz_l = P.nnz_l; out->nnz_local = P.nnz_l_local; out->nnz_remote = P.nnz_l_remote;
        out->col_remote_size = P.col_remote_size; out->vIndexSize = P.vIndexSize; out->recvSize = P.recvSize;
        out->numRecvProc = P.numRecvProc; out->numSendProc = P.numSendProc;
        out->use_double = P.use_double; out->active = g.active;
    } else {
        restrict_matrix &R = g.R;
        out->M = R.M; out->Mbig = R.Mbig; out->Nbig = R.Nbig;
        out->nnz_l = R.nnz_l; out->nnz_local = R.nnz_l_local; out->nnz_remote = R.nnz_l_remote;
        out->col_remote_size = R.col_remote_size; out->vIndexSize = R.vIndexSize; out->recvSize = R.recvSize;
        out->numRecvProc = R.numRecvProc; out->numSendProc = R.numSendProc;
        out->use_double = R.use_double; out->active = g.active;
    }
    return 0;
}

// field ids for sref_array
enum { F_NNZ_PER_ROW_LOCAL = 0, F_COL_LOCAL = 1, F_VAL_LOCAL = 2, F_INV_DIAG = 3, F_SPLIT = 4, F_SPLIT_NEW = 5,
       F_ROW_REMOTE = 6, F_VAL_REMOTE = 7, F_NNZ_PER_COL_REMOTE = 8, F_ENTRY_ROW = 9, F_ENTRY_COL = 10,
       F_ENTRY_VAL = 11, F_INV_SQ_DIAG_ORIG = 12,
       // the halo plan (ranks are ranks of the LEVEL's communicator: see sref_level_comm)
       F_VINDEX = 13, F_SEND_PROC_RANK = 14, F_SEND_PROC_COUNT = 15, F_VDISPLS = 16, F_RECV_PROC_RANK = 17,
       F_RECV_PROC_COUNT = 18, F_RDISPLS = 19 };

// Copies one layout array of an operator into `dst` (if non-null) and returns its element count.
long sref_array(void *hv, int l, int kind, int field, void *dst) {
    Handle *h = (Handle *)hv;
    Grid &g = obj(h)->grids[l];
    if (!g.A || !g.A->active || (kind != 0 && !g.active)) return 0;
#define COPY_VEC(v) do { if (dst && !(v).empty()) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0])); \
                         return (long)(v).size(); } while (0)
#define COPY_PTR(p, n) do { if (dst && (n) > 0) memcpy(dst, (p), (size_t)(n) * sizeof((p)[0])); return (long)(n); } while (0)
#define COPY_ENTRY(vec, member, T) do { if (dst) { T *d = (T *)dst; for (size_t i = 0; i < (vec).size(); ++i) \
                         d[i] = (T)(vec)[i].member; } return (long)(vec).size(); } while (0)
    if (kind == 0) {
        saena_matrix *A = g.A;
        switch (field) {
            case F_NNZ_PER_ROW_LOCAL: COPY_VEC(A->nnzPerRow_local);
            case F_COL_LOCAL: COPY_PTR(A->col_local, A->nnz_l_local);
            case F_VAL_LOCAL: COPY_PTR(A->val_local, A->nnz_l_local);
            case F_INV_DIAG: COPY_PTR(A->inv_diag, A->M);
            case F_INV_SQ_DIAG_ORIG: COPY_VEC(A->inv_sq_diag_orig);
            case F_SPLIT: COPY_VEC(A->split);
            case F_ROW_REMOTE: COPY_PTR(A->row_remote, A->nnz_l_remote);
            case F_VAL_REMOTE: COPY_PTR(A->val_remote, A->nnz_l_remote);
            case F_NNZ_PER_COL_REMOTE: COPY_VEC(A->nnzPerCol_remote);
            case F_VINDEX: COPY_VEC(A->vIndex);
            case F_SEND_PROC_RANK: COPY_VEC(A->sendProcRank);
            case F_SEND_PROC_COUNT: COPY_VEC(A->sendProcCount);
            case F_VDISPLS: COPY_VEC(A->vdispls);
            case F_RECV_PROC_RANK: COPY_VEC(A->recvProcRank);
            case F_RECV_PROC_COUNT: COPY_VEC(A->recvProcCount);
            case F_RDISPLS: COPY_VEC(A->rdispls);
            case F_ENTRY_ROW: COPY_ENTRY(A->entry, row, int);
            case F_ENTRY_COL: COPY_ENTRY(A->entry, col, int);
            case F_ENTRY_VAL: COPY_ENTRY(A->entry, val, double);
            default: return -1;
        }
    } else if (kind == 1) {
        prolong_matrix &P = g.P;
        switch (field) {
            case F_NNZ_PER_ROW_LOCAL: COPY_VEC(P.nnzPerRow_local);
            case F_COL_LOCAL: COPY_VEC(P.col_local);
            case F_VAL_LOCAL: COPY_VEC(P.val_local);
            case F_SPLIT: COPY_VEC(P.split);
            case F_SPLIT_NEW: COPY_VEC(P.splitNew);
            case F_ROW_REMOTE: COPY_VEC(P.row_remote);
            case F_VAL_REMOTE: COPY_VEC(P.val_remote);
            case F_NNZ_PER_COL_REMOTE: COPY_VEC(P.nnzPerCol_remote);
            case F_VINDEX: COPY_VEC(P.vIndex);
            case F_SEND_PROC_RANK: COPY_VEC(P.sendProcRank);
            case F_SEND_PROC_COUNT: COPY_VEC(P.sendProcCount);
            case F_VDISPLS: COPY_VEC(P.vdispls);
            case F_RECV_PROC_RANK: COPY_VEC(P.recvProcRank);
            case F_RECV_PROC_COUNT: COPY_VEC(P.recvProcCount);
            case F_RDISPLS: COPY_VEC(P.rdispls);
            case F_ENTRY_ROW: COPY_ENTRY(P.entry, row, int);
            case F_ENTRY_COL: COPY_ENTRY(P.entry, col, int);
            case F_ENTRY_VAL: COPY_ENTRY(P.entry, val, double);
            default: return -1;
        }
    } else {
        restrict_matrix &R = g.R;
        switch (field) {
            case F_NNZ_PER_ROW_LOCAL: COPY_VEC(R.nnzPerRow_local);
            case F_COL_LOCAL: COPY_VEC(R.col_local);
            case F_VAL_LOCAL: COPY_VEC(R.val_local);
            case F_SPLIT: COPY_VEC(R.split);
            case F_SPLIT_NEW: COPY_VEC(R.splitNew);
            case F_ROW_REMOTE: COPY_VEC(R.row_remote);
            case F_VAL_REMOTE: COPY_VEC(R.val_remote);
            case F_NNZ_PER_COL_REMOTE: COPY_VEC(R.nnzPerCol_remote);
            case F_VINDEX: COPY_VEC(R.vIndex);
            case F_SEND_PROC_RANK: COPY_VEC(R.sendProcRank);
            case F_SEND_PROC_COUNT: COPY_VEC(R.sendProcCount);
            case F_VDISPLS: COPY_VEC(R.vdispls);
            case F_RECV_PROC_RANK: COPY_VEC(R.recvProcRank);
            case F_RECV_PROC_COUNT: COPY_VEC(R.recvProcCount);
            case F_RDISPLS: COPY_VEC(R.rdispls);
            case F_ENTRY_ROW: COPY_ENTRY(R.entry, row, int);
            case F_ENTRY_COL: COPY_ENTRY(R.entry, col, int);
            case F_ENTRY_VAL: COPY_ENTRY(R.entry, val, double);
            default: return -1;
        }
    }
#undef COPY_VEC
#undef COPY_PTR
#undef COPY_ENTRY
}

// The communicator of level l (it shrinks as the levels get small, saena_matrix_shrink.cpp): the WORLD
// ranks of its members in its own rank order; every rank number in that level's halo plans and
// splits refers to this order.  Returns the size (0 on a rank that is not a member).  Collective
// over the level's communicator.
int sref_level_comm(void *hv, int l, int *world_ranks) {
    Handle *h = (Handle *)hv;
    Grid &g = obj(h)->grids[l];
    if (!g.A || !g.A->active) return 0;
    int n = 0, me = 0;
    MPI_Comm_size(g.A->comm, &n);
    MPI_Comm_rank(MPI_COMM_WORLD, &me);
    MPI_Allgather(&me, 1, MPI_INT, world_ranks, 1, MPI_INT, g.A->comm);
    return n;
}

// Grid::repart_u's plan of level l (grid.cpp:3-97; ranks of the level's communicator): which = 0 the
// sends (scount3 / sproc_id / sdispls2), 1 the receives; and the coarse sizes around it
int sref_repart_plan(void *hv, int l, int which, int *peer, int *offset, int *count) {
    Handle *h = (Handle *)hv;
    Grid &g = obj(h)->grids[l];
    if (!g.A || !g.A->active || !g.active) return 0;
    const std::vector<int> &cnt = which ? g.rcount3 : g.scount3, &id = which ? g.rproc_id : g.sproc_id,
                           &dsp = which ? g.rdispls2 : g.sdispls2;
    if (peer)
        for (size_t i = 0; i < cnt.size(); ++i) { peer[i] = id[i]; offset[i] = dsp[id[i]]; count[i] = cnt[i]; }
    return (int)cnt.size();
}
void sref_coarse_sizes(void *hv, int l, int *M_old, int *M_new) {
    Handle *h = (Handle *)hv;
    Grid &g = obj(h)->grids[l];
    *M_old = *M_new = 0;
    if (!g.A || !g.A->active || !g.active) return;
    *M_old = (int)g.Ac.M_old;
    *M_new = g.Ac.active ? (int)g.Ac.M : 0;
}

// solver parameters the reference ended up with
void sref_params(void *hv, int *pre, int *post, int *max_iter, double *tol, int *smoother_is_cheb) {
    saena_object *o = obj((Handle *)hv);
    *pre = o->preSmooth; *post = o->postSmooth; *max_iter = o->solver_max_iter; *tol = o->solver_tol;
    *smoother_is_cheb = (o->smoother == "chebyshev");
}

// grids[0].rhs (repartitioned, boundary rows removed), length grids[0].A->M
void sref_rhs(void *hv, double *out) {
    saena_object *o = obj((Handle *)hv);
    memcpy(out, o->grids[0].rhs, sizeof(double) * (size_t)o->grids[0].A->M);
}

// ---- the reference's own hot-path functions ----
void sref_matvec(void *hv, int l, int kind, const double *v, double *w) {
    Grid &g = obj((Handle *)hv)->grids[l];
    if (kind == 0) g.A->matvec(v, w);
    else if (kind == 1) g.P.matvec(v, w);
    else g.R.matvec(v, w);
}

void sref_residual(void *hv, int l, const double *u, const double *rhs, double *res) {
    obj((Handle *)hv)->grids[l].A->residual(u, rhs, res);
}

// smoother: 0 = jacobi, 1 = chebyshev (saena_object.tpp:85-96 dispatch)
void sref_smooth(void *hv, int l, int smoother, int iters, double *u, const double *rhs) {
    saena_matrix *A = level_A((Handle *)hv, l);
    if (smoother == 1) A->chebyshev(iters, u, rhs);
    else A->jacobi(iters, u, rhs);
}

double sref_dot(void *hv, const double *a, const double *b, int n) {
    double d = 0;
    dotProduct(a, b, n, &d, level_A((Handle *)hv, 0)->comm);
    return d;
}

// One V-cycle starting at grid `l` with the solver's current smoother settings.
void sref_vcycle(void *hv, int l, int pre, int post, int smoother, double *u, double *rhs) {
    Handle *h = (Handle *)hv;
    saena_object *o = obj(h);
    if (!h->vcycle_mem) { o->alloc_vcycle_memory(); h->vcycle_mem = true; }
    o->preSmooth = pre; o->postSmooth = post; o->smoother = smoother ? "chebyshev" : "jacobi";
    o->vcycle(&o->grids[l], u, rhs);
}

// saena_object::find_eig (saena_object.cpp:572-590) + find_eig_lamlan (lamlan_saena.h:13-79) on level l,
// with the Lanczos start vector given by the caller instead of std::random_device -- the only
// change, through the engine's own `init_vector` hook -- so that the result is reproducible.
// Returns what find_eig stores: 1.0001 * eigenvalue.  scale_matrix / scale_back_matrix round-trip
// the values in place, as in the reference's setup: use a solver object of its own for this call.
double sref_find_eig_start(void *hv, int l, const double *start, int *iters) {
    saena_matrix *A = level_A((Handle *)hv, l);
    A->scale_matrix(false);
    auto mv_mul = [&](const std::vector<value_t> &in, std::vector<value_t> &out) { A->matvec(&in[0], &out[0]); };
    lambda_lanczos::LambdaLanczos<value_t> engine(mv_mul, A->M, true, A->comm);
    engine.init_vector = [&](std::vector<value_t> &v) {
        for (size_t i = 0; i < v.size(); ++i) v[i] = start[i];
    };
    value_t eigenvalue = 0.0;
    std::vector<value_t> eigenvector;
    const int itern = engine.run(eigenvalue, eigenvector);
    A->scale_back_matrix(false);
    if (iters) *iters = itern;
    return 1.0001 * eigenvalue;
}

// saena_object::direct_solver: "SuperLU" (default, saena_object.h:165) or "CG"
void sref_set_direct_solver(void *hv, int use_cg) { obj((Handle *)hv)->direct_solver = use_cg ? "CG" : "SuperLU"; }

// solve_coarsest_CG (saena_object_solve.cpp:14-114) on the coarsest grid
void sref_coarsest_cg(void *hv, double *u, double *rhs) {
    saena_object *o = obj((Handle *)hv);
    o->solve_coarsest_CG(o->grids[o->max_level].A, u, rhs);
}

// Coarsest-level direct solve as the reference does it (SuperLU_DIST pdgssvx).
void sref_coarsest_solve(void *hv, double *u, double *rhs) {
    saena_object *o = obj((Handle *)hv);
    o->solve_coarsest_SuperLU(o->grids[o->max_level].A, u, rhs);
}

// saena::amg::solve_pCG; returns the reported iteration count (i+1) and the
// history sqrt(<r,r>) [0] = initial, [k] = after iteration k.
int sref_solve_pcg(void *hv, int max_iter, double tol, int smoother, int pre, int post, double *u_out,
                   double *hist, int hist_cap, int *hist_len, int quiet) {
    Handle *h = (Handle *)hv;
    if (h->vcycle_mem) { obj(h)->free_vcycle_memory(); h->vcycle_mem = false; }
    h->opts->set_solve_params(max_iter, tol, smoother ? "chebyshev" : "jacobi", pre, post);
    g_rr.clear();
    g_rr_len = obj(h)->grids[0].A->M;
    value_t *u = nullptr;
    {
        QuietStdout q(quiet != 0);
        h->solver->solve_pCG(u, h->opts, false);
    }
    g_rr_len = -1;
    const int M = obj(h)->grids[0].A->M;
    if (u_out) memcpy(u_out, u, sizeof(double) * (size_t)M);
    saena_free(u);
    int n = (int)g_rr.size();
    *hist_len = n;
    for (int i = 0; i < n && i < hist_cap; ++i) hist[i] = sqrt(g_rr[i]);
    return n - 1;  // one <r,r> before the loop, one per executed iteration
}

// saena_object::solve (which = 1: stationary V-cycles, src/saena_object_solve.cpp:1883-2014) and
// saena_object::solve_smoother (which = 2: the smoother alone, :2017-2117), called on the saena_object itself -- the
// public forwarders (src/saena.cpp:751-768) only add return_vec, a permutation back to the caller's rhs ordering.
// Same outputs as sref_solve_pcg: u in the matrix's row partition, history of sqrt(<r,r>), iteration count.
int sref_solve_stationary(void *hv, int which, int max_iter, double tol, int smoother, int pre, int post, double *u_out,
                          double *hist, int hist_cap, int *hist_len) {
    Handle *h = (Handle *)hv;
    if (h->vcycle_mem) { obj(h)->free_vcycle_memory(); h->vcycle_mem = false; }
    obj(h)->set_solve_params(max_iter, tol, smoother ? "chebyshev" : "jacobi", pre, post);
    g_rr.clear();
    g_rr_len = obj(h)->grids[0].A->M;
    value_t *u = nullptr;
    {
        QuietStdout q(true);
        if (which == 1) obj(h)->solve(u);   // both allocate u when it is null and start from zero (:1920-1924, :2047-2051)
        else obj(h)->solve_smoother(u);
    }
    g_rr_len = -1;
    const int M = obj(h)->grids[0].A->M;
    if (u_out) memcpy(u_out, u, sizeof(double) * (size_t)M);
    saena_free(u);
    int n = (int)g_rr.size();
    *hist_len = n;
    for (int i = 0; i < n && i < hist_cap; ++i) hist[i] = sqrt(g_rr[i]);
    return n - 1;
}

// Wall-clock seconds of `reps` reference solve_pCG calls (cpu_baseline).
double sref_time_solve_pcg(void *hv, int reps) {
    Handle *h = (Handle *)hv;
    if (h->vcycle_mem) { obj(h)->free_vcycle_memory(); h->vcycle_mem = false; }
    QuietStdout q(true);
    double t0 = MPI_Wtime();
    for (int i = 0; i < reps; ++i) {
        value_t *u = nullptr;
        h->solver->solve_pCG(u, h->opts, false);
        saena_free(u);
    }
    return MPI_Wtime() - t0;
}

// Seconds per application of A_level, the way saena_object::profile_matvecs times it (src/saena_object.cpp:618-638:
// v = 1, a barrier, `reps` timed matvecs with v and w swapped in between) -- returned instead of printed.
double sref_time_matvec(void *hv, int level, int reps) {
    Handle *h = (Handle *)hv;
    if (level < 0 || level > obj(h)->max_level) return -1.0;
    Grid &g = obj(h)->grids[level];
    if (!g.active || !g.A || !g.A->active) return 0.0;
    std::vector<value_t> v(g.A->M, 1), w(g.A->M);
    MPI_Barrier(g.A->comm);
    double t = 0;
    for (int i = 0; i < reps; ++i) {
        const double t1 = MPI_Wtime();
        g.A->matvec(&v[0], &w[0]);
        t += MPI_Wtime() - t1;
        std::swap(v, w);
    }
    return t / (reps > 0 ? reps : 1);
}

// saena::amg::profile_matvecs through the public API (experiments/Poisson.cpp:262): the reference's host loop in
// the reference builds, the adaptor's device timings in the drop-in builds.  Prints "matvec level l" lines.
void sref_profile_matvecs(void *hv) {
    Handle *h = (Handle *)hv;
    h->solver->profile_matvecs();
    fflush(stdout);
}

}  // extern "C"
