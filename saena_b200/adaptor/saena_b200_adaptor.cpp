// saena_b200_adaptor.cpp -- the drop-in: strong definitions of the five saena.hpp methods that
// make up the solve path, routed to the GPU through the C ABI of include/saena_b200.h.
//
// How it is used: build the reference as usual, weaken the five symbols of its
// forwarding TU src/saena.cpp (objcopy --weaken-symbol, no source edit), and link this file +
// libsaena_b200.so.  Every other saena.hpp call (matrix assembly, amg::set_matrix = the whole AMG
// setup, amg::set_rhs, options, destroy...) keeps running the reference's host code unchanged;
// include/saena.hpp itself is untouched, so the class layout user code sees is identical.
//
//   replaced                                         reference definition
//   saena::amg::solve_pCG(u, opts, print_info)       src/saena.cpp:786-799 -> saena_object::solve_pCG
//   saena::amg::solve(u, opts)                       src/saena.cpp:760-767 -> saena_object::solve
//   saena::amg::solve_CG(u, opts)                    src/saena.cpp:770-777 -> saena_object::solve_CG
//   saena::amg::solve_smoother(u, opts)              src/saena.cpp:751-758 -> saena_object::solve_smoother
//   saena::matrix::matvec(std::vector&, std::vector&) src/saena.cpp:226-228 -> saena_matrix::matvec
//
// The hierarchy saena_object::setup left in `grids` is uploaded once, lazily on the first solve
// (side table keyed by the saena_object*, because the header's class layout must not change) and
// released by saena_b200_adaptor_release() / at exit.  The reference's own CPU path stays
// callable in the same process as `solver.get_object()->solve_pCG(u)` -- that is how
// tests/test_public_api_dropin.py checks parity on the very same hierarchy object.
//
// Error convention: the reference prints and terminates; a non-zero status from
// the C ABI is mapped to exactly that.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <map>
#include <vector>

#include "saena.hpp"
#include "saena_object.h"
#include "saena_matrix.h"
#include "saena_vector.h"
#include "grid.h"

#include "saena_b200.h"

namespace {

struct DeviceSide {
    saena_b200_ctx *ctx = nullptr;
    int rank = 0, nprocs = 1;
    // what the upload was made from: a second set_matrix, or a lazy update (saena_object::update1/2/3,
    // src/saena_object_lazy.cpp:7-209 -- bodies compiled out in this version of the reference), replaces
    // grids[0].A; the device copy is then stale and is rebuilt on the next solve
    const saena_matrix *A0 = nullptr;
    int max_level = -1;
    MPI_Comm comm = MPI_COMM_NULL;   // the communicator the context was created on (grids[0].A->comm)
    // The side tables are keyed by object addresses, and an address can be handed out again after a destroy():
    // what the upload was made from is remembered too (array address, sizes) and compared before every use.
    const void *vals = nullptr;
    long nnz = -1;
    int M = -1;
    void remember(const saena_matrix *A) { vals = A->val_local; nnz = (long)A->nnz_l; M = (int)A->M; }
    bool made_from(const saena_matrix *A) const { return vals == A->val_local && nnz == (long)A->nnz_l && M == (int)A->M; }
};

std::map<const saena_object *, DeviceSide> g_solvers;
std::map<const saena_matrix *, DeviceSide> g_matrices;  // stand-alone saena::matrix::matvec
int g_verbose = 0;
int g_last_iterations = 0;
std::vector<double> g_last_history;

[[noreturn]] void die(saena_b200_ctx *ctx, const char *what) {
    std::printf("Error: saena_b200 %s: %s\n", what, saena_b200_last_error(ctx));
    std::fflush(stdout);
    int inited = 0;
    MPI_Initialized(&inited);
    if (inited) MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);
    std::exit(EXIT_FAILURE);
}

#define CK(ctx, call, what) do { if (call) die(ctx, what); } while (0)

// Collective, like every call of the reference's API: on several ranks nobody unmaps its peer-memory arena while a
// neighbour's last kernel may still be raising flags in it.
void release(DeviceSide &ds) {
    if (!ds.ctx) return;
    if (ds.nprocs > 1 && ds.comm != MPI_COMM_NULL) MPI_Barrier(ds.comm);
    saena_b200_destroy(ds.ctx);
    ds.ctx = nullptr;
}

saena_b200_ctx *new_context(MPI_Comm comm, int &rank, int &nprocs) {
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &nprocs);
    unsigned char id[SAENA_B200_NCCL_ID_BYTES] = {0};
    if (nprocs > 1) {
        if (rank == 0 && saena_b200_nccl_unique_id(id)) die(nullptr, "nccl id");
        MPI_Bcast(id, SAENA_B200_NCCL_ID_BYTES, MPI_BYTE, 0, comm);
    }
    // one rank per GPU: the local device index is the rank among the ranks of this node
    const char *lr = std::getenv("OMPI_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = std::getenv("MV2_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = std::getenv("LOCAL_RANK");
    if (!lr) lr = std::getenv("SBMPI_RANK");   // oracle/mprun.py (the tests' multi-process MPI stand-in, one node)
    const int device = lr ? std::atoi(lr) : 0;
    saena_b200_ctx *ctx = nullptr;
    CK(nullptr, saena_b200_init(&ctx, device, rank, nprocs, nprocs > 1 ? id : nullptr), "init");
    return ctx;
}

// The communicator of a level shrinks as the levels get small (saena_matrix_shrink.cpp): every rank
// number in that level's halo plans, displacement tables and Grid::repart_u plans is a rank of the
// LEVEL's communicator.  The device context is built once on grids[0].A->comm, so everything is
// translated to ranks of that communicator here.  (The same translation, on the same arrays, is what
// oracle/ref.py does for the multi-rank oracle -- checked against the reference on 2..8 ranks in
// tests/test_multirank_reference.py.)
struct LevelComm {
    std::vector<int> world;  // world[rank in the level's communicator] = rank in grids[0].A->comm; empty: not a member
    int me = -1;             // my rank in the level's communicator
};

LevelComm level_comm(saena_matrix *A, int world_rank, int world_size) {
    LevelComm lc;
    if (!A || !A->active) return lc;
    if (world_size == 1) { lc.world.assign(1, 0); lc.me = 0; return lc; }
    int n = 0;
    MPI_Comm_size(A->comm, &n);
    MPI_Comm_rank(A->comm, &lc.me);
    lc.world.resize(n);
    MPI_Allgather(&world_rank, 1, MPI_INT, lc.world.data(), 1, MPI_INT, A->comm);
    return lc;
}

// halo plan of one operator with world ranks: owns the translated tables the descriptor points to
struct WorldPlan {
    std::vector<int32_t> sendRank, recvRank, vdispls, rdispls;
};

template <class Op>
void fill_plan(saena_b200_operator_desc &d, Op &op, const LevelComm &lc, int world_size, WorldPlan &wp) {
    d.nnz_remote = op.nnz_l_remote;
    d.col_remote_size = (int32_t)op.col_remote_size;
    d.row_remote = op.nnz_l_remote ? &op.row_remote[0] : nullptr;
    d.val_remote = op.nnz_l_remote ? &op.val_remote[0] : nullptr;
    d.nnzPerCol_remote = op.nnzPerCol_remote.empty() ? nullptr : op.nnzPerCol_remote.data();
    d.vIndexSize = (int32_t)op.vIndexSize;
    d.vIndex = op.vIndex.empty() ? nullptr : op.vIndex.data();
    wp.vdispls.assign(world_size, 0);
    wp.rdispls.assign(world_size, 0);
    for (size_t r = 0; r < lc.world.size(); ++r) {   // a one-rank communicator builds no tables at all
        if (r < op.vdispls.size()) wp.vdispls[lc.world[r]] = op.vdispls[r];
        if (r < op.rdispls.size()) wp.rdispls[lc.world[r]] = op.rdispls[r];
    }
    wp.sendRank.clear();
    wp.recvRank.clear();
    for (int i = 0; i < (int)op.numSendProc; ++i) wp.sendRank.push_back(lc.world[op.sendProcRank[i]]);
    for (int i = 0; i < (int)op.numRecvProc; ++i) wp.recvRank.push_back(lc.world[op.recvProcRank[i]]);
    d.numSendProc = (int32_t)op.numSendProc;
    d.sendProcRank = wp.sendRank.data();
    d.sendProcCount = op.sendProcCount.data();
    d.vdispls = wp.vdispls.data();
    d.numRecvProc = (int32_t)op.numRecvProc;
    d.recvProcRank = wp.recvRank.data();
    d.recvProcCount = op.recvProcCount.data();
    d.rdispls = wp.rdispls.data();
    d.use_double = op.use_double ? 1 : 0;
}

// a level this rank is not a member of: operators with no rows, no columns, no plan
void upload_empty(saena_b200_ctx *ctx, int kind, int level, int world_size) {
    static const int32_t none = 0;
    std::vector<int32_t> zeros(world_size, 0);
    saena_b200_operator_desc d{};
    d.kind = kind;
    d.level = level;
    d.nnzPerRow_local = &none;
    d.vdispls = zeros.data();
    d.rdispls = zeros.data();
    d.use_double = 1;
    CK(ctx, saena_b200_upload_operator(ctx, &d), "upload empty operator");
}

// saena_matrix keeps its local/remote arrays as raw pointers (saena_matrix.h:107-113)
void upload_A(saena_b200_ctx *ctx, saena_matrix *A, int level, const LevelComm &lc, int world_size) {
    saena_b200_operator_desc d{};
    WorldPlan wp;
    d.kind = SAENA_B200_KIND_A;
    d.level = level;
    d.M = (int32_t)A->M;
    d.n_local_cols = (int32_t)A->M;
    d.col_offset = (int32_t)A->split[lc.me];
    d.nnz_local = A->nnz_l_local;
    d.nnzPerRow_local = A->nnzPerRow_local.data();
    d.col_local = A->col_local;
    d.val_local = A->val_local;
    fill_plan(d, *A, lc, world_size, wp);
    d.row_remote = A->row_remote;
    d.val_remote = A->val_remote;
    CK(ctx, saena_b200_upload_operator(ctx, &d), "upload A");
}

template <class Op>
void upload_PR(saena_b200_ctx *ctx, Op &op, int kind, int level, int n_rows, int col_offset, int n_local_cols,
               const LevelComm &lc, int world_size) {
    saena_b200_operator_desc d{};
    WorldPlan wp;
    d.kind = kind;
    d.level = level;
    d.M = n_rows;
    d.n_local_cols = n_local_cols;
    d.col_offset = col_offset;
    d.nnz_local = op.nnz_l_local;
    d.nnzPerRow_local = op.nnzPerRow_local.data();
    d.col_local = op.col_local.data();
    d.val_local = op.val_local.data();
    fill_plan(d, op, lc, world_size, wp);
    CK(ctx, saena_b200_upload_operator(ctx, &d), kind == SAENA_B200_KIND_P ? "upload P" : "upload R");
}

// Walk saena_object::grids (include/saena_object.h:189, include/grid.h) and upload once.
DeviceSide &device_side(saena_object *obj) {
    auto it = g_solvers.find(obj);
    if (it != g_solvers.end()) {
        if (it->second.A0 == obj->grids[0].A && it->second.max_level == obj->max_level &&
            it->second.made_from(obj->grids[0].A))
            return it->second;
        // the hierarchy changed under the solver object (update1/2/3, set_matrix again): upload it anew
        if (g_verbose && it->second.rank == 0) std::printf("saena_b200: hierarchy changed, uploading again\n");
        release(it->second);
        g_solvers.erase(it);
    }
    DeviceSide ds;
    saena_matrix *A0 = obj->grids[0].A;
    ds.A0 = A0;
    ds.max_level = obj->max_level;
    ds.ctx = new_context(A0->comm, ds.rank, ds.nprocs);
    ds.comm = A0->comm;
    ds.remember(A0);
    const int L = obj->max_level;
    const std::vector<double> no_diag(1, 0.0);
    for (int l = 0; l <= L; ++l) {
        Grid &g = obj->grids[l];
        saena_matrix *A = g.A;
        const LevelComm lc = level_comm(A, ds.rank, ds.nprocs);   // collective over the level's communicator
        if (lc.world.empty()) {
            // a shrink left this rank out of level l: it owns nothing of it
            upload_empty(ds.ctx, SAENA_B200_KIND_A, l, ds.nprocs);
            if (l < L) {
                upload_empty(ds.ctx, SAENA_B200_KIND_P, l, ds.nprocs);
                upload_empty(ds.ctx, SAENA_B200_KIND_R, l, ds.nprocs);
            }
            CK(ds.ctx, saena_b200_upload_level_aux(ds.ctx, l, no_diag.data(), 1.0, 0, 0, 0, nullptr, 0, nullptr),
               "upload level aux");
            continue;
        }
        upload_A(ds.ctx, A, l, lc, ds.nprocs);
        // switch_to_dense: the reference applies this level through saena_matrix_dense (include/saena_matrix.tpp:5-7).
        // The sparse arrays uploaded above hold the same matrix (generate_dense_matrix leaves them in place,
        // src/saena_matrix_setup.cpp:1638-1646); the flag makes the device reproduce the dense path's float cast of
        // the whole input vector when the level runs in float precision (src/saena_matrix_dense.cpp:281-282)
        if (A->use_dense) CK(ds.ctx, saena_b200_set_operator_dense(ds.ctx, l, SAENA_B200_KIND_A, 1), "set operator dense");
        std::vector<saena_b200_block> send, recv;
        int M_old = 0, M_new = 0;
        if (l < L) {
            prolong_matrix &P = g.P;
            restrict_matrix &R = g.R;
            // fine side: A's split; coarse side: splitNew, the partition R writes into (before Grid::repart_u)
            upload_PR(ds.ctx, P, SAENA_B200_KIND_P, l, (int)P.M, (int)P.splitNew[lc.me],
                      (int)(P.splitNew[lc.me + 1] - P.splitNew[lc.me]), lc, ds.nprocs);
            upload_PR(ds.ctx, R, SAENA_B200_KIND_R, l, (int)R.M, (int)A->split[lc.me],
                      (int)(A->split[lc.me + 1] - A->split[lc.me]), lc, ds.nprocs);
            M_old = (int)g.Ac.M_old;
            M_new = g.Ac.active ? (int)g.Ac.M : 0;
            // Grid::repart_u plan (grid.cpp:99-130, on Ac.comm_old = this level's communicator): receive
            // rcount3[i] values from rproc_id[i] at rdispls2[rproc_id[i]]; send scount3[i] values to
            // sproc_id[i] from sdispls2[sproc_id[i]]
            for (size_t i = 0; i < g.scount3.size(); ++i)
                send.push_back({lc.world[g.sproc_id[i]], g.sdispls2[g.sproc_id[i]], g.scount3[i]});
            for (size_t i = 0; i < g.rcount3.size(); ++i)
                recv.push_back({lc.world[g.rproc_id[i]], g.rdispls2[g.rproc_id[i]], g.rcount3[i]});
            // a one-rank plan that only copies the vector onto itself is the identity
            if (ds.nprocs == 1) { send.clear(); recv.clear(); }
        }
        CK(ds.ctx, saena_b200_upload_level_aux(ds.ctx, l, A->inv_diag, A->eig_max_of_invdiagXA, M_old, M_new,
                                               (int)send.size(), send.data(), (int)recv.size(), recv.data()),
           "upload level aux");
        if (obj->scale)  // D^-1/2 hooks of the V-cycle (saena_object_solve.cpp:1245-1247,1264-1266,2709-2711)
            CK(ds.ctx, saena_b200_upload_level_scale(ds.ctx, l, A->inv_sq_diag_orig.data()), "upload level scale");
    }
    // coarsest operator: the COO entries setup_SuperLU passes on (saena_object_solve.cpp:282-308)
    saena_matrix *Ac = obj->grids[L].A;
    if (!Ac || !Ac->active) {
        CK(ds.ctx, saena_b200_upload_coarsest(ds.ctx, 0, 0, nullptr, nullptr, nullptr), "upload coarsest");
    } else if (Ac->M == Ac->Mbig) {
        std::vector<int32_t> r(Ac->entry.size()), c(Ac->entry.size());
        std::vector<double> v(Ac->entry.size());
        for (size_t i = 0; i < Ac->entry.size(); ++i) {
            r[i] = Ac->entry[i].row; c[i] = Ac->entry[i].col; v[i] = Ac->entry[i].val;
        }
        CK(ds.ctx, saena_b200_upload_coarsest(ds.ctx, (int)Ac->Mbig, (int64_t)v.size(), r.data(), c.data(), v.data()),
           "upload coarsest");
    } else if (Ac->M == 0) {
        CK(ds.ctx, saena_b200_upload_coarsest(ds.ctx, 0, 0, nullptr, nullptr, nullptr), "upload coarsest");
    } else {
        std::printf("Error: saena_b200: the coarsest level must live on one rank (enable_shrink_c)\n");
        std::exit(EXIT_FAILURE);
    }
    CK(ds.ctx, saena_b200_finalize(ds.ctx), "finalize");
    if (ds.nprocs > 1 && !std::getenv("SAENA_B200_HALO_NCCL")) {
        // peer-memory halo: all-gather the ranks' export blobs over MPI and import them
        int64_t n = 0;
        CK(ds.ctx, saena_b200_p2p_export(ds.ctx, nullptr, 0, &n), "p2p export");
        std::vector<char> mine((size_t)n), all((size_t)n * ds.nprocs);
        CK(ds.ctx, saena_b200_p2p_export(ds.ctx, mine.data(), n, &n), "p2p export");
        MPI_Allgather(mine.data(), (int)n, MPI_BYTE, all.data(), (int)n, MPI_BYTE, A0->comm);
        CK(ds.ctx, saena_b200_p2p_import(ds.ctx, all.data(), n), "p2p import");
        MPI_Barrier(A0->comm);
    }
    if (!std::getenv("SAENA_B200_NO_AUTOTUNE")) {
        // setup-time choices by measurement (collective, untimed by the drivers: experiments/Poisson.cpp:191-214 times
        // the solves only): every operator's row mapping, then -- several ranks -- per operator the fused halo kernel
        // or the separate launches, whichever is faster here
        CK(ds.ctx, saena_b200_autotune_mapping(ds.ctx, 10, 0.03, nullptr), "mapping autotune");
        if (ds.nprocs > 1 && !std::getenv("SAENA_B200_HALO_NCCL"))
            CK(ds.ctx, saena_b200_autotune_halo(ds.ctx, 10), "halo autotune");
    }
    if (g_verbose && ds.rank == 0) std::printf("saena_b200: hierarchy of %d levels uploaded\n", L + 1);
    return g_solvers[obj] = ds;
}

int smoother_id(const std::string &s) {
    if (s == "chebyshev") return SAENA_B200_CHEBYSHEV;
    if (s == "jacobi") return SAENA_B200_JACOBI;
    std::printf("Error: Unknown smoother");  // saena_object.tpp:92-94
    std::exit(EXIT_FAILURE);
}

enum Which { PCG, VCYCLE, CG, SMOOTHER };

int run_solver(saena_object *obj, Which which, value_t *&u, saena::options *opts, bool print_info) {
    obj->set_solve_params(opts->get_max_iter(), opts->get_tol(), opts->get_smoother(), opts->get_preSmooth(),
                          opts->get_postSmooth());
    DeviceSide &ds = device_side(obj);
    if (obj->direct_solver != "SuperLU" && obj->direct_solver != "CG") {
        if (!ds.rank) std::printf("Error: Unknown direct solver! \n");  // saena_object_solve.cpp:1011-1013
        std::exit(EXIT_FAILURE);
    }
    CK(ds.ctx, saena_b200_set_coarsest_solver(ds.ctx, obj->direct_solver == "CG"), "set coarsest solver");
    saena_matrix *A = obj->grids[0].A;
    const index_t sz = A->M;
    if (u == nullptr) u = saena_aligned_alloc<value_t>(sz);  // saena_object_solve.cpp:2478-2480
    const int cap = obj->solver_max_iter + 2;
    g_last_history.assign(cap, 0.0);
    int iters = 0, nh = 0;
    const int sm = smoother_id(obj->smoother);
    int rc = 0;
    if (which == PCG)
        rc = saena_b200_solve_pcg(ds.ctx, obj->grids[0].rhs, u, obj->solver_max_iter, obj->solver_tol, sm,
                                  obj->preSmooth, obj->postSmooth, &iters, g_last_history.data(), cap, &nh);
    else if (which == VCYCLE)
        rc = saena_b200_solve_vcycle(ds.ctx, obj->grids[0].rhs, u, obj->solver_max_iter, obj->solver_tol, sm,
                                     obj->preSmooth, obj->postSmooth, &iters, g_last_history.data(), cap, &nh);
    else if (which == SMOOTHER)
        rc = saena_b200_solve_smoother(ds.ctx, obj->grids[0].rhs, u, obj->solver_max_iter, obj->solver_tol, sm,
                                       obj->preSmooth, obj->postSmooth, &iters, g_last_history.data(), cap, &nh);
    else
        rc = saena_b200_solve_cg(ds.ctx, obj->grids[0].rhs, u, obj->solver_max_iter, obj->solver_tol, &iters,
                                 g_last_history.data(), cap, &nh);
    if (rc) die(ds.ctx, "solve");
    g_last_history.resize(nh < cap ? nh : cap);
    g_last_iterations = iters;
    if (print_info && ds.rank == 0 && !g_last_history.empty()) {
        // the summary saena_object::solve_pCG prints (saena_object_solve.cpp:2501-2503, :2678-2682)
        const double r0 = g_last_history.front(), r1 = g_last_history.back();
        std::printf("\ninitial residual        = %e \n", r0);
        std::printf("stopped at iteration    = %d \nfinal absolute residual = %e"
                    "\nrelative residual       = %e \n", iters, r1, r1 / r0);
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// the replaced saena.hpp methods (same signatures, include/saena.hpp:211-224, :59-60)
// ---------------------------------------------------------------------------------------------
int saena::amg::solve_pCG(value_t *&u, saena::options *opts, const bool print_info) {
    // saena.cpp:786-799: no return_vec on this path
    return run_solver(m_pImpl, PCG, u, opts, print_info);
}

int saena::amg::solve(value_t *&u, saena::options *opts) {
    run_solver(m_pImpl, VCYCLE, u, opts, true);
    m_pImpl->grids[0].rhs_orig->return_vec(u);  // host post-step kept (saena.cpp:764-765)
    return 0;
}

int saena::amg::solve_CG(value_t *&u, saena::options *opts) {
    run_solver(m_pImpl, CG, u, opts, true);
    m_pImpl->grids[0].rhs_orig->return_vec(u);  // saena.cpp:774-775
    return 0;
}

int saena::amg::solve_smoother(value_t *&u, saena::options *opts) {
    // src/saena.cpp:751-758 -> saena_object::solve_smoother (src/saena_object_solve.cpp:2017-2117): the smoother
    // alone as a stationary iteration, on the device like the other solvers
    run_solver(m_pImpl, SMOOTHER, u, opts, true);
    m_pImpl->grids[0].rhs_orig->return_vec(u);  // host post-step kept (saena.cpp:755-756)
    return 0;
}

void saena::amg::destroy() {
    // src/saena.cpp:882-884 forwards to saena_object::destroy; the device copy of the hierarchy goes first
    auto it = g_solvers.find(m_pImpl);
    if (it != g_solvers.end()) {
        release(it->second);
        g_solvers.erase(it);
    }
    m_pImpl->destroy();
}

void saena::matrix::destroy() {
    // src/saena.cpp:246-248 forwards to saena_matrix::destroy; a device copy made for stand-alone matvec goes first
    auto it = g_matrices.find(m_pImpl);
    if (it != g_matrices.end()) {
        release(it->second);
        g_matrices.erase(it);
    }
    m_pImpl->destroy();
}

void saena::amg::profile_matvecs() {
    // saena_object::profile_matvecs (src/saena_object.cpp:618-638; called by experiments/Poisson.cpp:262 and
    // profile_file.cpp:235): 5 timed A_l matvecs per level, the average printed through print_time_all over the
    // level's communicator.  Here the device's applications of the same operators, each launch timed with CUDA
    // events on the library's stream (saena_b200_time_matvec), printed through the same function: the driver's
    // "matvec level l" lines become the GPU's.  Every rank calls for every level (a rank a shrink left out of a
    // level holds an empty operator there and prints nothing, as in the reference).
    saena_object *obj = m_pImpl;
    DeviceSide &ds = device_side(obj);
    const int iter = 5;
    for (int l = 0; l <= obj->max_level; ++l) {
        float ms = 0.f;
        CK(ds.ctx, saena_b200_time_matvec(ds.ctx, l, SAENA_B200_KIND_A, iter, 0, &ms), "time matvec");
        if (obj->grids[l].active && obj->grids[l].A && obj->grids[l].A->active)
            print_time_all(1e-3 * ms, "matvec level " + std::to_string(l), obj->grids[l].A->comm);
    }
}

void saena::matrix::matvec(std::vector<value_t> &v, std::vector<value_t> &w) {
    saena_matrix *A = m_pImpl;
    auto it = g_matrices.find(A);
    if (it != g_matrices.end() && !it->second.made_from(A)) {   // another matrix behind the same address
        release(it->second);
        g_matrices.erase(it);
        it = g_matrices.end();
    }
    if (it == g_matrices.end()) {
        DeviceSide ds;
        ds.ctx = new_context(A->comm, ds.rank, ds.nprocs);
        ds.comm = A->comm;
        ds.remember(A);
        upload_A(ds.ctx, A, 0, level_comm(A, ds.rank, ds.nprocs), ds.nprocs);
        CK(ds.ctx, saena_b200_upload_level_aux(ds.ctx, 0, A->inv_diag, A->eig_max_of_invdiagXA, 0, 0, 0, nullptr, 0,
                                               nullptr), "upload level aux");
        // a lone operator: one level, no coarsest factor needed for matvec
        CK(ds.ctx, saena_b200_upload_coarsest(ds.ctx, 0, 0, nullptr, nullptr, nullptr), "upload coarsest");
        CK(ds.ctx, saena_b200_finalize(ds.ctx), "finalize");
        it = g_matrices.emplace(A, ds).first;
    }
    if (w.size() < (size_t)A->M) w.resize(A->M);
    CK(it->second.ctx, saena_b200_matvec(it->second.ctx, 0, SAENA_B200_KIND_A, v.data(), w.data()), "matvec");
}

// ---------------------------------------------------------------------------------------------
// small C surface for drivers and tests
// ---------------------------------------------------------------------------------------------
extern "C" {

void saena_b200_adaptor_set_verbose(int v) { g_verbose = v; }

// Free the device side of one solver (saena::amg::destroy() does it too), or of everything (NULL) -- the
// stand-alone saena::matrix::matvec contexts included, which no reference call releases.  Collective.
void saena_b200_adaptor_release(saena::amg *solver) {
    if (solver) {
        auto it = g_solvers.find(solver->get_object());
        if (it != g_solvers.end()) {
            release(it->second);
            g_solvers.erase(it);
        }
        return;
    }
    for (auto &kv : g_solvers) release(kv.second);
    for (auto &kv : g_matrices) release(kv.second);
    g_solvers.clear();
    g_matrices.clear();
}

// iteration count / residual history of the last device solve (the reference only prints them)
int saena_b200_adaptor_last_iterations(void) { return g_last_iterations; }
int saena_b200_adaptor_last_history(double *out, int cap) {
    const int n = (int)g_last_history.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = g_last_history[i];
    return n;
}

}  // extern "C"
