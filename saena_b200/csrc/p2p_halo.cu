// p2p_halo.cu -- the ghost-value exchange of the distributed SpMV over NVLink peer memory.
//
// What it replaces: the MPI_Isend / MPI_Irecv pair of saena_matrix::matvec_sparse
// (/root/reference/src/saena_matrix_matvec.cpp:25-41, :463-478).  With NCCL (nccl_comm.cu) one
// exchange costs 22-34 us stand-alone on a B200 box -- pack kernel, ncclSend/ncclRecv launch and
// rendezvous -- a floor that dominates the coarse levels when 256^3 is spread over 8 GPUs.
// Here the pack kernel IS the exchange: it gathers v[vIndex[i]] and stores each value straight
// into the receiving rank's ghost buffer (an IPC-mapped peer allocation; NVSwitch gives every
// peer the same bandwidth), then the last CTA raises an "arrived" flag in the receiver's memory.
//
//   sender, comm stream  : wait consumed[op][receiver] == 1, then reset it to 0   (my arena)
//                          pack kernel: peer stores + __threadfence_system + arrived[op][me] = 1
//   receiver, compute    : interior rows kernel (overlaps the above)
//                          wait arrived[op][sender] == 1, then reset it to 0     (my arena)
//                          boundary / merged kernel reads the ghost values
//                          signal kernel: consumed[op][me] = 1 in every sender's arena
//
// The flags are binary and every wait is followed by its own reset, so one application leaves the
// flags exactly as it found them (arrived 0, consumed 1): the sequence carries no counter and can
// be captured once into a CUDA graph and replayed (solve.cu).  A sender cannot raise `arrived`
// again before it has seen `consumed`, which the receiver raises only after its reset and its
// read of the ghost values -- no reset can swallow a later signal.
//
// No kernel ever spins: wait + reset are one cuStreamBatchMemOp (stream memory operations, in
// order) on the waiting rank's OWN memory, the writers are ordinary kernels.  One cudaMalloc'd arena per rank holds all flags and all ghost
// areas, so a single IPC handle per rank is exchanged (saena_b200_p2p_export / _import, carried
// by whatever bootstrap channel the host has: torch.distributed in bench.py, MPI in the adaptor).
#include <cuda.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.h"

namespace {

typedef CUresult (*BatchMemOpFn)(CUstream, unsigned int, CUstreamBatchMemOpParams *, unsigned int);
BatchMemOpFn g_batch_mem_op = nullptr;

bool load_driver_entry(std::string &err) {
    if (g_batch_mem_op) return true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamBatchMemOp", &fn, cudaEnableDefault, &qr);
    if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        err = "cuStreamBatchMemOp is not available from this driver";
        return false;
    }
    g_batch_mem_op = (BatchMemOpFn)fn;
    return true;
}

// one batch on stream s: wait until every flag is 1, then reset every flag to 0 (in this order)
int wait_and_reset(saena_b200_ctx *ctx, const std::vector<unsigned long long *> &flags, cudaStream_t s,
                   const char *what) {
    if (flags.empty()) return 0;
    std::vector<CUstreamBatchMemOpParams> ops(2 * flags.size());
    memset(ops.data(), 0, sizeof(CUstreamBatchMemOpParams) * ops.size());
    for (size_t i = 0; i < flags.size(); ++i) {
        ops[i].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_64;
        ops[i].waitValue.address = (CUdeviceptr)flags[i];
        ops[i].waitValue.value64 = 1;
        ops[i].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
        CUstreamBatchMemOpParams &w = ops[flags.size() + i];
        w.writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_64;
        w.writeValue.address = (CUdeviceptr)flags[i];
        w.writeValue.value64 = 0;
        w.writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
    }
    const CUresult r = g_batch_mem_op((CUstream)s, (unsigned int)ops.size(), ops.data(), 0);
    if (r != CUDA_SUCCESS) SB_FAIL(std::string("cuStreamBatchMemOp(") + what + ") failed, CUresult " + std::to_string((int)r));
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// operators in a fixed order every rank agrees on: index = level * 3 + kind
std::vector<DevOperator *> all_ops(saena_b200_ctx *ctx) {
    std::vector<DevOperator *> v;
    for (DevLevel &lv : ctx->levels) {
        v.push_back(&lv.A);
        v.push_back(&lv.P);
        v.push_back(&lv.R);
    }
    return v;
}

unsigned long long *flag_ptr(char *arena_base, int nranks, size_t n_ops,
                             int which /*0 arrived, 1 consumed (binary, stream memory ops); 2 arrived, 3 consumed (epochs, fused kernel)*/,
                             size_t op_index, int peer) {
    return (unsigned long long *)arena_base + ((size_t)which * n_ops + op_index) * nranks + peer;
}

// export blob: [int64 magic, rank, nranks, n_ops, arena_bytes][64-byte IPC handle]
//              then per operator: [present, ghost_off, esz, ghost_d_off, recv_elem_off[nranks] (-1: none), recv_count[nranks]]
const int64_t P2P_MAGIC = 0x5342323030503250LL;

}  // namespace

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
void sb_arena_free(saena_b200_ctx *ctx) {
    for (size_t p = 0; p < ctx->peer_arena.size(); ++p)
        if (ctx->peer_arena[p]) cudaIpcCloseMemHandle(ctx->peer_arena[p]);
    ctx->peer_arena.clear();
    ctx->p2p_ready = false;
    for (DevOperator *op : all_ops(ctx)) {
        op->p2p = false;
        op->fused = false;
        op->ghost_buf = nullptr;
        op->ghost_d = nullptr;
        op->x_ext = nullptr;
        cudaFree(op->fh.epoch); cudaFree(op->fh.tickets); cudaFree(op->fh.segs);
        cudaFree(op->fh.wait_consumed); cudaFree(op->fh.wait_arrived); cudaFree(op->fh.signal_consumed);
        op->fh = FusedHaloDev();
        cudaFree(op->p2p_segs); cudaFree(op->p2p_ticket); cudaFree(op->p2p_signal_consumed);
        op->p2p_segs = nullptr; op->p2p_ticket = nullptr; op->p2p_signal_consumed = nullptr;
        op->p2p_wait_arrived.clear();
        op->p2p_wait_consumed.clear();
    }
    cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
}

int sb_arena_build(saena_b200_ctx *ctx) {
    sb_arena_free(ctx);
    std::vector<DevOperator *> ops = all_ops(ctx);
    const size_t n_ops = ops.size();
    size_t off = align_up(4 * n_ops * (size_t)ctx->nranks * sizeof(unsigned long long), 256);
    bool any = false;
    for (DevOperator *op : ops) {
        if (!op->present) continue;
        const size_t esz = op->use_double ? sizeof(double) : sizeof(float);
        if (op->merged) {
            // [x_ext = n_local_cols + recvSize doubles][float ghosts when the halo is f32]
            op->ghost_arena_off = off;
            off = align_up(off + sizeof(double) * ((size_t)op->n_local_cols + op->recvSize), 256);
            if (!op->use_double) off = align_up(off + sizeof(float) * (size_t)op->recvSize, 256);
            any = true;
        } else if (op->recvSize) {
            op->ghost_arena_off = off;
            off = align_up(off + esz * (size_t)op->recvSize, 256);
            any = true;
        }
        if (op->recvSize) {  // landing area of the fused kernel: always doubles
            op->ghost_d_off = off;
            off = align_up(off + sizeof(double) * (size_t)op->recvSize, 256);
        }
    }
    if (!any && ctx->nranks == 1) return 0;
    SB_CUDA(cudaMalloc((void **)&ctx->arena, off));
    SB_CUDA(cudaMemset(ctx->arena, 0, off));
    {
        // rest state of the flags: arrived 0, consumed 1 ("the receiver is done with what I sent last")
        std::vector<unsigned long long> ones(n_ops * (size_t)ctx->nranks, 1ull);
        SB_CUDA(cudaMemcpy(flag_ptr(ctx->arena, ctx->nranks, n_ops, 1, 0, 0), ones.data(),
                           sizeof(unsigned long long) * ones.size(), cudaMemcpyHostToDevice));
    }
    ctx->arena_bytes = off;
    for (DevOperator *op : ops) {
        if (!op->present) continue;
        if (op->merged) {
            op->x_ext = (double *)(ctx->arena + op->ghost_arena_off);
            if (op->use_double) {
                op->ghost_buf = op->x_ext + op->n_local_cols;  // received in place
            } else {
                const size_t f = align_up(op->ghost_arena_off + sizeof(double) * ((size_t)op->n_local_cols + op->recvSize), 256);
                op->ghost_buf = ctx->arena + f;
            }
        } else if (op->recvSize) {
            op->ghost_buf = ctx->arena + op->ghost_arena_off;
        }
        if (op->recvSize) op->ghost_d = (double *)(ctx->arena + op->ghost_d_off);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
p2p_pack_kernel(int n, const int *__restrict__ vIndex, const double *__restrict__ v, const P2PSegment *__restrict__ segs,
                int n_segs, unsigned int *ticket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int s = 0;
        while (s + 1 < n_segs && i >= segs[s + 1].start) ++s;  // a handful of receivers
        ((T *)segs[s].dst)[i - segs[s].start] = (T)v[vIndex[i]];
    }
    // publish: every CTA's stores are fenced system-wide before its ticket; the last CTA raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int s = 0; s < n_segs; ++s) *(volatile unsigned long long *)segs[s].arrived = 1ull;
            __threadfence_system();
            *ticket = 0u;
        }
    }
}

__global__ void p2p_signal_kernel(unsigned long long *const *flags, int n) {
    const int i = threadIdx.x;
    if (i < n) {
        *(volatile unsigned long long *)flags[i] = 1ull;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------
// per-application steps
// ---------------------------------------------------------------------------------------------
int sb_p2p_pack_and_signal(saena_b200_ctx *ctx, DevOperator &op, const double *x, cudaStream_t s) {
    // my previous values must have been consumed by every receiver before I overwrite them
    SB_TRY(wait_and_reset(ctx, op.p2p_wait_consumed, s, "consumed"));
    if (op.vIndexSize) {
        ++ctx->launches;
        const int blocks = (op.vIndexSize + 255) / 256;
        if (op.use_double)
            p2p_pack_kernel<double><<<blocks, 256, 0, s>>>(op.vIndexSize, op.vIndex, x, op.p2p_segs, (int)op.sends.size(),
                                                          op.p2p_ticket);
        else
            p2p_pack_kernel<float><<<blocks, 256, 0, s>>>(op.vIndexSize, op.vIndex, x, op.p2p_segs, (int)op.sends.size(),
                                                         op.p2p_ticket);
        SB_CUDA(cudaGetLastError());
    }
    return 0;
}

int sb_p2p_wait_arrived(saena_b200_ctx *ctx, DevOperator &op, cudaStream_t s) {
    return wait_and_reset(ctx, op.p2p_wait_arrived, s, "arrived");
}

int sb_p2p_signal_consumed(saena_b200_ctx *ctx, DevOperator &op, cudaStream_t s) {
    if (op.recvs.empty()) return 0;
    ++ctx->launches;
    p2p_signal_kernel<<<1, 32, 0, s>>>(op.p2p_signal_consumed, (int)op.recvs.size());
    SB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// export / import (C ABI)
// ---------------------------------------------------------------------------------------------
extern "C" {

int saena_b200_p2p_export(saena_b200_ctx *ctx, void *buf, int64_t cap, int64_t *size_out) {
    if (!ctx) return 1;
    SB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->finalized) SB_FAIL("p2p_export: call saena_b200_finalize first");
    if (ctx->nranks == 1 || !ctx->arena) SB_FAIL("p2p_export: nothing to export (one rank or no halo)");
    std::vector<DevOperator *> ops = all_ops(ctx);
    const size_t n_ops = ops.size();
    const size_t per_op = 4 + 2 * (size_t)ctx->nranks;
    const size_t n64 = 5 + 8 + n_ops * per_op;  // header, 64-byte handle, table
    *size_out = (int64_t)(n64 * sizeof(int64_t));
    if (!buf || cap < *size_out) return 0;  // size query
    std::vector<int64_t> out(n64, 0);
    out[0] = P2P_MAGIC;
    out[1] = ctx->rank;
    out[2] = ctx->nranks;
    out[3] = (int64_t)n_ops;
    out[4] = (int64_t)ctx->arena_bytes;
    cudaIpcMemHandle_t h;
    SB_CUDA(cudaIpcGetMemHandle(&h, ctx->arena));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(&out[5], &h, 64);
    for (size_t k = 0; k < n_ops; ++k) {
        int64_t *e = &out[13 + k * per_op];
        DevOperator *op = ops[k];
        for (int p = 0; p < ctx->nranks; ++p) e[4 + p] = -1;
        if (!op->present || op->recvs.empty()) continue;
        e[0] = 1;
        e[1] = (int64_t)((char *)op->ghost_buf - ctx->arena);
        e[2] = op->use_double ? 8 : 4;
        e[3] = (int64_t)op->ghost_d_off;
        for (const HaloPeer &r : op->recvs) {
            e[4 + r.peer] = r.offset;
            e[4 + ctx->nranks + r.peer] = r.count;
        }
    }
    memcpy(buf, out.data(), (size_t)*size_out);
    return 0;
}

// blobs: the export of every rank, concatenated in rank order, each `blob_bytes` long
int saena_b200_p2p_import(saena_b200_ctx *ctx, const void *blobs, int64_t blob_bytes) {
    if (!ctx) return 1;
    SB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->arena) SB_FAIL("p2p_import: no halo arena (finalize first)");
    if (!load_driver_entry(ctx->error)) return 1;
    std::vector<DevOperator *> ops = all_ops(ctx);
    const size_t n_ops = ops.size();
    const int N = ctx->nranks;
    const size_t per_op = 4 + 2 * (size_t)N;
    auto blob = [&](int p) { return (const int64_t *)((const char *)blobs + (size_t)p * (size_t)blob_bytes); };
    for (int p = 0; p < N; ++p) {
        const int64_t *b = blob(p);
        if (b[0] != P2P_MAGIC || b[1] != p || b[2] != N || (size_t)b[3] != n_ops)
            SB_FAIL("p2p_import: blob of a peer does not match this hierarchy");
    }
    // open the arenas of the ranks I send to
    ctx->peer_arena.assign(N, nullptr);
    for (DevOperator *op : ops)
        if (op->present)
            for (const HaloPeer &s : op->sends)
                if (!ctx->peer_arena[s.peer]) {
                    cudaIpcMemHandle_t h;
                    memcpy(&h, &blob(s.peer)[5], 64);
                    void *ptr = nullptr;
                    SB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
                    ctx->peer_arena[s.peer] = ptr;
                }
    // receivers of my "consumed" signals are the ranks I receive from: their arenas too
    for (DevOperator *op : ops)
        if (op->present)
            for (const HaloPeer &r : op->recvs)
                if (!ctx->peer_arena[r.peer]) {
                    cudaIpcMemHandle_t h;
                    memcpy(&h, &blob(r.peer)[5], 64);
                    void *ptr = nullptr;
                    SB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
                    ctx->peer_arena[r.peer] = ptr;
                }
    for (size_t k = 0; k < n_ops; ++k) {
        DevOperator *op = ops[k];
        if (!op->present || (op->sends.empty() && op->recvs.empty())) continue;
        std::vector<P2PSegment> segs;
        std::vector<FusedSeg> fsegs;
        std::vector<unsigned long long *> f_wait_consumed, f_wait_arrived, f_signal;
        op->p2p_wait_consumed.clear();
        op->p2p_wait_arrived.clear();
        const size_t esz = op->use_double ? 8 : 4;
        long long expect_start = 0;
        for (const HaloPeer &s : op->sends) {
            const int64_t *e = &blob(s.peer)[13 + k * per_op];
            if (!e[0] || e[4 + ctx->rank] < 0 || e[4 + N + ctx->rank] != s.count || (size_t)e[2] != esz)
                SB_FAIL("p2p_import: a peer's receive plan does not match my send plan");
            if (s.offset != expect_start) SB_FAIL("p2p_import: send slices are not contiguous in rank order");
            expect_start += s.count;
            char *pa = (char *)ctx->peer_arena[s.peer];
            P2PSegment g;
            g.start = s.offset;
            g.count = s.count;
            g.dst = pa + e[1] + (size_t)e[4 + ctx->rank] * esz;
            g.arrived = flag_ptr(pa, N, n_ops, 0, k, ctx->rank);
            segs.push_back(g);
            op->p2p_wait_consumed.push_back(flag_ptr(ctx->arena, N, n_ops, 1, k, s.peer));
            FusedSeg fg;
            fg.start = s.offset;
            fg.count = s.count;
            fg.dst = (double *)(pa + e[3]) + e[4 + ctx->rank];
            fg.arrived = flag_ptr(pa, N, n_ops, 2, k, ctx->rank);
            fsegs.push_back(fg);
            f_wait_consumed.push_back(flag_ptr(ctx->arena, N, n_ops, 3, k, s.peer));
        }
        std::vector<unsigned long long *> signal;
        for (const HaloPeer &r : op->recvs) {
            op->p2p_wait_arrived.push_back(flag_ptr(ctx->arena, N, n_ops, 0, k, r.peer));
            signal.push_back(flag_ptr((char *)ctx->peer_arena[r.peer], N, n_ops, 1, k, ctx->rank));
            f_wait_arrived.push_back(flag_ptr(ctx->arena, N, n_ops, 2, k, r.peer));
            f_signal.push_back(flag_ptr((char *)ctx->peer_arena[r.peer], N, n_ops, 3, k, ctx->rank));
        }
        cudaFree(op->p2p_segs); cudaFree(op->p2p_ticket); cudaFree(op->p2p_signal_consumed);
        op->p2p_segs = nullptr; op->p2p_ticket = nullptr; op->p2p_signal_consumed = nullptr;
        SB_CUDA(cudaMalloc((void **)&op->p2p_segs, sizeof(P2PSegment) * std::max<size_t>(segs.size(), 1)));
        if (!segs.empty())
            SB_CUDA(cudaMemcpy(op->p2p_segs, segs.data(), sizeof(P2PSegment) * segs.size(), cudaMemcpyHostToDevice));
        SB_CUDA(cudaMalloc((void **)&op->p2p_ticket, sizeof(unsigned int)));
        SB_CUDA(cudaMemset(op->p2p_ticket, 0, sizeof(unsigned int)));
        SB_CUDA(cudaMalloc((void **)&op->p2p_signal_consumed, sizeof(void *) * std::max<size_t>(signal.size(), 1)));
        if (!signal.empty())
            SB_CUDA(cudaMemcpy(op->p2p_signal_consumed, signal.data(), sizeof(void *) * signal.size(),
                               cudaMemcpyHostToDevice));
        op->p2p = true;
        // fused kernel: its own epoch flags, landing area and device tables
        {
            FusedHaloDev &fh = op->fh;
            cudaFree(fh.epoch); cudaFree(fh.tickets); cudaFree(fh.segs);
            cudaFree(fh.wait_consumed); cudaFree(fh.wait_arrived); cudaFree(fh.signal_consumed);
            fh = FusedHaloDev();
            SB_CUDA(cudaMalloc((void **)&fh.epoch, sizeof(unsigned long long)));
            SB_CUDA(cudaMemset(fh.epoch, 0, sizeof(unsigned long long)));
            SB_CUDA(cudaMalloc((void **)&fh.tickets, 2 * sizeof(unsigned int)));
            SB_CUDA(cudaMemset(fh.tickets, 0, 2 * sizeof(unsigned int)));
            auto up = [&](auto **dst, const auto &v) -> int {
                typedef typename std::remove_reference<decltype(v[0])>::type T;
                SB_CUDA(cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(v.size(), 1)));
                if (!v.empty()) SB_CUDA(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
                return 0;
            };
            SB_TRY(up(&fh.segs, fsegs));
            SB_TRY(up(&fh.wait_consumed, f_wait_consumed));
            SB_TRY(up(&fh.wait_arrived, f_wait_arrived));
            SB_TRY(up(&fh.signal_consumed, f_signal));
            op->fused = ctx->fused_default;
        }
    }
    ctx->p2p_ready = true;
    sb_invalidate_graphs(ctx);  // a captured V-cycle holds the other transport's nodes
    return 0;
}

// 0: back to ncclSend/ncclRecv for the halo (the imported mappings stay open); 1: peer stores with
// separate launches (pack kernel, stream memory-op flags, boundary kernel); 2: the fused kernel
int saena_b200_p2p_enable(saena_b200_ctx *ctx, int on) {
    if (!ctx) return 1;
    if (on && !ctx->p2p_ready) SB_FAIL("p2p_enable: import the peers' exports first");
    SB_CUDA(cudaSetDevice(ctx->device));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    for (DevOperator *op : all_ops(ctx))
        if (op->present && op->p2p_segs) {
            op->p2p = on != 0;
            op->fused = on == 2;
        }
    sb_invalidate_graphs(ctx);
    return 0;
}

}  // extern "C"
