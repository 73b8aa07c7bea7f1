// abi_recorder.cpp -- TEST INFRASTRUCTURE: a stand-in for libsaena_b200.so that RECORDS what the drop-in
// adaptor (saena_b200/adaptor/saena_b200_adaptor.cpp) uploads through the C ABI of include/saena_b200.h
// instead of putting it on a GPU.  Linked with the reference (multi-process MPI build) and the adaptor
// into oracle/_ref/libsaena_dropin_rec_mp.so, it lets the adaptor's multi-rank walk over
// saena_object::grids -- per-level communicators, rank translation, ranks a shrink left out -- run in
// this image, N processes and no GPU, and be compared array by array with oracle/ref.py's extraction
// (which is pinned against the reference's numbers).  It computes nothing: the solve entry points return
// a zero vector.  Never linked into the product.
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "saena_b200.h"

namespace {
struct RecOp {
    bool present = false;
    saena_b200_operator_desc d{};
    std::vector<int32_t> nnzPerRow_local, col_local, row_remote, nnzPerCol_remote, vIndex, sendProcRank, sendProcCount,
        vdispls, recvProcRank, recvProcCount, rdispls;
    std::vector<double> val_local, val_remote;
};
struct RecLevel {
    RecOp op[3];
    bool aux = false;
    std::vector<double> inv_diag;
    double eig = 0;
    int M_old = 0, M_new = 0;
    std::vector<int32_t> send, recv;  // flattened (peer, offset, count)
};
}  // namespace

struct saena_b200_ctx {
    int rank = 0, nranks = 1;
    std::vector<RecLevel> levels;
    int coarse_n = 0;
    std::vector<int32_t> crow, ccol;
    std::vector<double> cval;
    bool finalized = false;
    int solves = 0;
    std::string error;
};

static saena_b200_ctx *g_last = nullptr;
static int g_inits = 0;
static std::vector<int> g_dense_levels;   // levels the adaptor flagged as saena_matrix::use_dense (last upload)

template <class T>
static void copy_in(std::vector<T> &dst, const T *src, size_t n) {
    dst.clear();
    if (src && n) dst.assign(src, src + n);
}

extern "C" {

int saena_b200_nccl_unique_id(void *out128) { memset(out128, 0x5a, SAENA_B200_NCCL_ID_BYTES); return 0; }
int saena_b200_init(saena_b200_ctx **ctx_out, int device_id, int rank, int nranks, const void *nccl_id) {
    (void)device_id; (void)nccl_id;
    saena_b200_ctx *c = new saena_b200_ctx();
    c->rank = rank; c->nranks = nranks;
    *ctx_out = c;
    g_last = c;
    ++g_inits;
    g_dense_levels.clear();
    return 0;
}
static int g_destroys = 0;
int saena_b200_destroy(saena_b200_ctx *ctx) { if (g_last == ctx) g_last = nullptr; delete ctx; ++g_destroys; return 0; }
extern "C" int rec_destroys() { return g_destroys; }
const char *saena_b200_last_error(const saena_b200_ctx *ctx) { return ctx ? ctx->error.c_str() : "recorder"; }

int saena_b200_upload_operator(saena_b200_ctx *ctx, const saena_b200_operator_desc *d) {
    if (d->level < 0 || d->kind < 0 || d->kind > 2) { ctx->error = "bad level / kind"; return 1; }
    if ((int)ctx->levels.size() <= d->level) ctx->levels.resize(d->level + 1);
    RecOp &o = ctx->levels[d->level].op[d->kind];
    o.present = true;
    o.d = *d;
    copy_in(o.nnzPerRow_local, d->nnzPerRow_local, (size_t)d->M);
    copy_in(o.col_local, d->col_local, (size_t)d->nnz_local);
    copy_in(o.val_local, d->val_local, (size_t)d->nnz_local);
    copy_in(o.row_remote, d->row_remote, (size_t)d->nnz_remote);
    copy_in(o.val_remote, d->val_remote, (size_t)d->nnz_remote);
    copy_in(o.nnzPerCol_remote, d->nnzPerCol_remote, (size_t)d->col_remote_size);
    copy_in(o.vIndex, d->vIndex, (size_t)d->vIndexSize);
    copy_in(o.sendProcRank, d->sendProcRank, (size_t)d->numSendProc);
    copy_in(o.sendProcCount, d->sendProcCount, (size_t)d->numSendProc);
    copy_in(o.vdispls, d->vdispls, (size_t)ctx->nranks);
    copy_in(o.recvProcRank, d->recvProcRank, (size_t)d->numRecvProc);
    copy_in(o.recvProcCount, d->recvProcCount, (size_t)d->numRecvProc);
    copy_in(o.rdispls, d->rdispls, (size_t)ctx->nranks);
    return 0;
}
int saena_b200_upload_level_aux(saena_b200_ctx *ctx, int level, const double *inv_diag, double eig_max, int M_coarse_old,
                                int M_coarse, int n_send, const saena_b200_block *send, int n_recv,
                                const saena_b200_block *recv) {
    if (level < 0 || level >= (int)ctx->levels.size() || !ctx->levels[level].op[0].present) { ctx->error = "aux before A"; return 1; }
    RecLevel &lv = ctx->levels[level];
    lv.aux = true;
    copy_in(lv.inv_diag, inv_diag, (size_t)lv.op[0].d.M);
    lv.eig = eig_max; lv.M_old = M_coarse_old; lv.M_new = M_coarse;
    lv.send.clear(); lv.recv.clear();
    for (int i = 0; i < n_send; ++i) { lv.send.push_back(send[i].peer); lv.send.push_back(send[i].offset); lv.send.push_back(send[i].count); }
    for (int i = 0; i < n_recv; ++i) { lv.recv.push_back(recv[i].peer); lv.recv.push_back(recv[i].offset); lv.recv.push_back(recv[i].count); }
    return 0;
}
int saena_b200_upload_level_scale(saena_b200_ctx *, int, const double *) { return 0; }
int saena_b200_upload_coarsest(saena_b200_ctx *ctx, int n, int64_t nnz, const int32_t *row, const int32_t *col, const double *val) {
    ctx->coarse_n = n;
    copy_in(ctx->crow, row, (size_t)nnz); copy_in(ctx->ccol, col, (size_t)nnz); copy_in(ctx->cval, val, (size_t)nnz);
    return 0;
}
int saena_b200_set_coarsest_solver(saena_b200_ctx *, int) { return 0; }
int saena_b200_set_operator_dense(saena_b200_ctx *, int level, int kind, int use_dense) {
    if (kind == 0 && use_dense) g_dense_levels.push_back(level);
    return 0;
}
static std::vector<int> g_timed_levels;    // saena_b200_time_matvec calls (saena::amg::profile_matvecs through the adaptor)
int saena_b200_time_matvec(saena_b200_ctx *, int level, int kind, int reps, int, float *ms_out) {
    if (kind == 0 && reps == 5) g_timed_levels.push_back(level);
    *ms_out = 0.25f * (float)(level + 1);
    return 0;
}
extern "C" int rec_timed_levels(int *out, int cap) {
    int n = 0;
    for (int l : g_timed_levels) if (n < cap) out[n++] = l;
    return (int)g_timed_levels.size();
}
extern "C" int rec_dense_levels(int *out, int cap) {
    int n = 0;
    for (int l : g_dense_levels) if (n < cap) out[n++] = l;
    return (int)g_dense_levels.size();
}
int saena_b200_finalize(saena_b200_ctx *ctx) {
    for (RecLevel &lv : ctx->levels) if (!lv.op[0].present || !lv.aux) { ctx->error = "finalize: a level lacks A or its aux data"; return 1; }
    ctx->finalized = true;
    return 0;
}
int saena_b200_p2p_export(saena_b200_ctx *, void *buf, int64_t cap, int64_t *size_out) {
    *size_out = 16;
    if (buf && cap >= 16) memset(buf, 0, 16);
    return 0;
}
int saena_b200_p2p_import(saena_b200_ctx *, const void *, int64_t) { return 0; }
int saena_b200_autotune_halo(saena_b200_ctx *, int) { return 0; }

static int fake_solve(saena_b200_ctx *ctx, double *u, int *iters, double *hist, int hist_cap, int *hist_len) {
    if (!ctx->finalized) { ctx->error = "solve before finalize"; return 1; }
    const int n = ctx->levels.empty() ? 0 : ctx->levels[0].op[0].d.M;
    for (int i = 0; i < n; ++i) u[i] = 0.0;
    *iters = 1;
    if (hist_cap > 0) hist[0] = 1.0;
    *hist_len = hist_cap > 0 ? 1 : 0;
    ++ctx->solves;
    return 0;
}
int saena_b200_solve_pcg(saena_b200_ctx *ctx, const double *, double *u, int, double, int, int, int, int *iters, double *hist,
                         int hist_cap, int *hist_len) { return fake_solve(ctx, u, iters, hist, hist_cap, hist_len); }
int saena_b200_solve_vcycle(saena_b200_ctx *ctx, const double *, double *u, int, double, int, int, int, int *iters,
                            double *hist, int hist_cap, int *hist_len) { return fake_solve(ctx, u, iters, hist, hist_cap, hist_len); }
int saena_b200_solve_smoother(saena_b200_ctx *ctx, const double *, double *u, int, double, int, int, int, int *iters,
                            double *hist, int hist_cap, int *hist_len) { return fake_solve(ctx, u, iters, hist, hist_cap, hist_len); }
int saena_b200_solve_cg(saena_b200_ctx *ctx, const double *, double *u, int, double, int *iters, double *hist, int hist_cap,
                        int *hist_len) { return fake_solve(ctx, u, iters, hist, hist_cap, hist_len); }
int saena_b200_autotune_mapping(saena_b200_ctx *, int, double, int *changed) { if (changed) *changed = 0; return 0; }
int saena_b200_matvec(saena_b200_ctx *ctx, int, int, const double *, double *w) {
    const int n = ctx->levels.empty() ? 0 : ctx->levels[0].op[0].d.M;
    for (int i = 0; i < n; ++i) w[i] = 0.0;
    return 0;
}

// ---- what was recorded (read by oracle/mp_worker.py in the same process) ----
int rec_info(int *rank, int *nranks, int *levels, int *coarse_n, int *solves) {
    if (!g_last) return 1;
    *rank = g_last->rank; *nranks = g_last->nranks; *levels = (int)g_last->levels.size();
    *coarse_n = g_last->coarse_n; *solves = g_last->solves;
    return 0;
}
int rec_inits() { return g_inits; }
// out[8]: present, M, n_local_cols, col_offset, use_double, nnz_local, nnz_remote, col_remote_size
int rec_op_scalars(int level, int kind, long *out) {
    const RecOp &o = g_last->levels[level].op[kind];
    out[0] = o.present; out[1] = o.d.M; out[2] = o.d.n_local_cols; out[3] = o.d.col_offset; out[4] = o.d.use_double;
    out[5] = (long)o.d.nnz_local; out[6] = (long)o.d.nnz_remote; out[7] = o.d.col_remote_size;
    return 0;
}
#define REC_COPY(v) do { if (dst && !(v).empty()) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0])); return (long)(v).size(); } while (0)
long rec_op_array(int level, int kind, int field, void *dst) {
    const RecOp &o = g_last->levels[level].op[kind];
    switch (field) {
        case 0: REC_COPY(o.nnzPerRow_local); case 1: REC_COPY(o.col_local); case 2: REC_COPY(o.val_local);
        case 3: REC_COPY(o.row_remote); case 4: REC_COPY(o.val_remote); case 5: REC_COPY(o.nnzPerCol_remote);
        case 6: REC_COPY(o.vIndex); case 7: REC_COPY(o.sendProcRank); case 8: REC_COPY(o.sendProcCount);
        case 9: REC_COPY(o.vdispls); case 10: REC_COPY(o.recvProcRank); case 11: REC_COPY(o.recvProcCount);
        case 12: REC_COPY(o.rdispls);
    }
    return -1;
}
int rec_level_aux(int level, double *eig, int *M_old, int *M_new) {
    const RecLevel &lv = g_last->levels[level];
    *eig = lv.eig; *M_old = lv.M_old; *M_new = lv.M_new;
    return lv.aux ? 0 : 1;
}
long rec_level_array(int level, int field, void *dst) {
    const RecLevel &lv = g_last->levels[level];
    switch (field) { case 0: REC_COPY(lv.inv_diag); case 1: REC_COPY(lv.send); case 2: REC_COPY(lv.recv); }
    return -1;
}
long rec_coarsest(int field, void *dst) {
    switch (field) { case 0: REC_COPY(g_last->crow); case 1: REC_COPY(g_last->ccol); case 2: REC_COPY(g_last->cval); }
    return -1;
}
#undef REC_COPY

}  // extern "C"
