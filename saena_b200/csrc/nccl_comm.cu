// nccl_comm.cu -- the communication the reference does with MPI, over NCCL / NVLink.
//
//   halo exchange      MPI_Irecv/MPI_Isend of the packed ghost values, tag 1
//                      (/root/reference/src/saena_matrix_matvec.cpp:32-41, :470-478)
//                      -> one ncclGroup of ncclSend/ncclRecv on the comm stream, overlapped with
//                         the interior-row SpMV on the compute stream
//   dot products       MPI_Allreduce(1 double, SUM) (include/aux_functions.h:121)
//                      -> ncclAllReduce on the device scalar, in place
//   Grid::repart_u     MPI_Isend/Irecv of coarse-vector blocks, tag 0 (src/grid.cpp:99-163)
//                      -> ncclSend/ncclRecv per plan entry (local blocks are device copies)
//
// libnccl is bound at run time (dlopen) so that the single-GPU library has no NCCL dependency;
// in a torchrun process this resolves to the libnccl.so.2 torch already loaded.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "common.h"

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string &err) {
    if (g_nccl.handle) return true;
    // SAENA_B200_NCCL_LIB names the library explicitly; otherwise the loader's search path decides (in a torchrun
    // process libnccl.so.2 resolves to the copy torch already mapped)
    const char *names[] = {getenv("SAENA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        err = std::string("cannot load libnccl: ") + dlerror();
        return false;
    }
#define SB_SYM(field, name)                                                     \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                       \
    if (!g_nccl.field) {                                                        \
        err = std::string("libnccl lacks ") + name;                             \
        return false;                                                           \
    }
    SB_SYM(GetUniqueId, "ncclGetUniqueId")
    SB_SYM(CommInitRank, "ncclCommInitRank")
    SB_SYM(CommDestroy, "ncclCommDestroy")
    SB_SYM(GroupStart, "ncclGroupStart")
    SB_SYM(GroupEnd, "ncclGroupEnd")
    SB_SYM(Send, "ncclSend")
    SB_SYM(Recv, "ncclRecv")
    SB_SYM(AllReduce, "ncclAllReduce")
    SB_SYM(GetErrorString, "ncclGetErrorString")
#undef SB_SYM
    g_nccl.handle = h;
    return true;
}
}  // namespace

#define SB_NCCL(call)                                                                              \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            ctx->error = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                    \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

static_assert(sizeof(ncclUniqueId) == SAENA_B200_NCCL_ID_BYTES, "NCCL id size");

int sb_nccl_unique_id(void *out, std::string &err) {
    if (!load_nccl(err)) return 1;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) {
        err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
        return 1;
    }
    memcpy(out, &id, sizeof(id));
    return 0;
}

int sb_nccl_init(saena_b200_ctx *ctx, const void *id_bytes) {
    if (ctx->nranks == 1) return 0;
    if (!id_bytes) SB_FAIL("init: nranks > 1 needs the NCCL unique id");
    if (!load_nccl(ctx->error)) return 1;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm;
    SB_NCCL(g_nccl.CommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->nccl_comm = comm;
    return 0;
}

void sb_nccl_destroy(saena_b200_ctx *ctx) {
    if (ctx->nccl_comm) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

int sb_halo_exchange(saena_b200_ctx *ctx, DevOperator &op, cudaStream_t s) {
    if (op.sends.empty() && op.recvs.empty()) return 0;
    if (!ctx->nccl_comm) SB_FAIL("halo exchange without a communicator");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const ncclDataType_t dt = op.use_double ? ncclDouble : ncclFloat;
    const size_t esz = op.use_double ? sizeof(double) : sizeof(float);
    SB_NCCL(g_nccl.GroupStart());
    for (const HaloPeer &r : op.recvs)
        SB_NCCL(g_nccl.Recv((char *)op.ghost_buf + (size_t)r.offset * esz, (size_t)r.count, dt, r.peer, comm, s));
    for (const HaloPeer &p : op.sends)
        SB_NCCL(g_nccl.Send((const char *)op.send_buf + (size_t)p.offset * esz, (size_t)p.count, dt, p.peer, comm, s));
    SB_NCCL(g_nccl.GroupEnd());
    return 0;
}

int sb_allreduce_sum(saena_b200_ctx *ctx, double *dev_vals, int count, cudaStream_t s) {
    if (ctx->nranks == 1) return 0;
    if (!ctx->nccl_comm) SB_FAIL("all-reduce without a communicator (a detached context has no peers)");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    SB_NCCL(g_nccl.AllReduce(dev_vals, dev_vals, (size_t)count, ncclDouble, ncclSum, comm, s));
    return 0;
}

int sb_allreduce_max(saena_b200_ctx *ctx, double *dev_vals, int count, cudaStream_t s) {
    if (ctx->nranks == 1) return 0;
    if (!ctx->nccl_comm) SB_FAIL("all-reduce without a communicator (a detached context has no peers)");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    SB_NCCL(g_nccl.AllReduce(dev_vals, dev_vals, (size_t)count, ncclDouble, ncclMax, comm, s));
    return 0;
}

int sb_repart(saena_b200_ctx *ctx, const RepartPlan &plan, bool backward, const double *src, double *dst,
              cudaStream_t s) {
    // forward : src is in the old partition (send offsets), dst in the new one (recv offsets)
    // backward: src is in the new partition (recv offsets), dst in the old one (send offsets)
    const std::vector<saena_b200_block> &out = backward ? plan.recv : plan.send;
    const std::vector<saena_b200_block> &in = backward ? plan.send : plan.recv;
    bool remote = false;
    for (const auto &b : out) remote |= (b.peer != ctx->rank);
    for (const auto &b : in) remote |= (b.peer != ctx->rank);
    // the block a rank keeps is a device copy
    for (const auto &o : out)
        if (o.peer == ctx->rank)
            for (const auto &i : in)
                if (i.peer == ctx->rank)
                    SB_CUDA(cudaMemcpyAsync(dst + i.offset, src + o.offset, sizeof(double) * (size_t)o.count,
                                            cudaMemcpyDeviceToDevice, s));
    if (!remote) return 0;
    if (!ctx->nccl_comm) SB_FAIL("repartition without a communicator");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    SB_NCCL(g_nccl.GroupStart());
    for (const auto &i : in)
        if (i.peer != ctx->rank) SB_NCCL(g_nccl.Recv(dst + i.offset, (size_t)i.count, ncclDouble, i.peer, comm, s));
    for (const auto &o : out)
        if (o.peer != ctx->rank) SB_NCCL(g_nccl.Send(src + o.offset, (size_t)o.count, ncclDouble, o.peer, comm, s));
    SB_NCCL(g_nccl.GroupEnd());
    return 0;
}
