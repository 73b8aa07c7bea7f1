// p2p_halo.cu -- the ghost-value exchange of the distributed SpMV over NVLink peer memory: the arena every rank
// exports, the import that wires the senders to it, and the SEPARATE-LAUNCH form of the exchange.
//
// What it replaces: the MPI_Isend / MPI_Irecv pair of saena_matrix::matvec_sparse
// (/root/reference/src/saena_matrix_matvec.cpp:25-41, :463-478).  With NCCL (nccl_comm.cu) one
// exchange costs 22-34 us stand-alone on a B200 box -- pack kernel, ncclSend/ncclRecv launch and
// rendezvous -- a floor that dominates the coarse levels when 256^3 is spread over 8 GPUs.
// Here the pack IS the exchange: it gathers v[vIndex[i]] and stores each value straight into the
// receiving rank's landing area (an IPC-mapped peer allocation; NVSwitch gives every peer the same
// bandwidth), then its last CTA raises an "arrived" counter in the receiver's memory.
//
// Two forms of one hand-shake (halo_sync.cuh: monotonic counters, two landing buffers, bounded waits):
//   fused             fused_halo.cu -- pack, interior rows and ghost rows in ONE kernel
//   separate launches this file:
//       comm stream    : pack kernel (waits for `consumed` of two applications ago, peer stores, raises `arrived`)
//       compute stream : interior-row kernel (overlaps the above), join the comm stream, wait kernel (1 CTA spins on
//                        `arrived` in this rank's OWN memory), boundary-row kernel on the current landing buffer,
//                        release kernel (raises `consumed` at the senders, advances this rank's counter)
//     The compute stream joins the comm stream before it touches the ghost values: the caller may overwrite x as
//     soon as the application returns, and no later launch of this rank (a fused kernel whose waiting CTAs fill the
//     chip, say) can be resident while this rank's pack kernel still waits for an SM -- the cycle behind round 1's
//     4-GPU hang.
// One cudaMalloc'd arena per rank holds all flags and all landing areas, so a single IPC handle per rank is
// exchanged (saena_b200_p2p_export / _import, carried by whatever bootstrap channel the host has:
// torch.distributed in bench.py, MPI in the adaptor).
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "halo_sync.cuh"

namespace {

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
                             int which /*0 arrived, 1 consumed*/,
                             size_t op_index, int peer) {
    return (unsigned long long *)arena_base + ((size_t)which * n_ops + op_index) * nranks + peer;
}

// export blob: [int64 magic, rank, nranks, n_ops, arena_bytes][64-byte IPC handle]
//              then per operator: [present, recvSize, spare, ghost_d_off, recv_elem_off[nranks] (-1: none), recv_count[nranks]]
const int64_t P2P_MAGIC = 0x5342323030503251LL;

}  // namespace

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
static void free_halo_sync(HaloSyncDev &hs) {
    cudaFree(hs.epoch); cudaFree(hs.tickets); cudaFree(hs.segs);
    cudaFree(hs.wait_consumed); cudaFree(hs.wait_arrived); cudaFree(hs.signal_consumed);
    hs = HaloSyncDev();
}

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
        free_halo_sync(op->hs);
    }
    cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
}

int sb_arena_build(saena_b200_ctx *ctx) {
    sb_arena_free(ctx);
    std::vector<DevOperator *> ops = all_ops(ctx);
    const size_t n_ops = ops.size();
    size_t off = align_up(2 * n_ops * (size_t)ctx->nranks * sizeof(unsigned long long), 256);
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
        if (op->recvSize) {  // landing area of the peer-memory exchange: always doubles, two buffers
            op->ghost_d_off = off;
            off = align_up(off + 2 * sizeof(double) * (size_t)op->recvSize, 256);
        }
    }
    if (!any && ctx->nranks == 1) return 0;
    SB_CUDA(cudaMalloc((void **)&ctx->arena, off));
    SB_CUDA(cudaMemset(ctx->arena, 0, off));   // every counter of the hand-shake starts at 0
    ctx->arena_bytes = off;
    for (size_t k = 0; k < n_ops; ++k) ops[k]->op_id = (int)k;
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
// kernels of the separate-launch form (the roles of fused_halo_spmv_kernel, one launch each)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
p2p_pack_kernel(const __grid_constant__ HaloSync h, const double *__restrict__ x, const int *__restrict__ vIndex,
                int vIndexSize, int round_float) {
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long *)&h.epoch[0];
    __syncthreads();
    sb_halo_pack_cta(h, blockIdx.x, gridDim.x, s_epoch, x, vIndex, vIndexSize, round_float != 0);
}

// one CTA: a thread per sender spins (bounded) on this rank's own memory
__global__ void __launch_bounds__(256)
p2p_wait_kernel(const __grid_constant__ HaloSync h) {
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long *)&h.epoch[1];
    __syncthreads();
    sb_halo_wait_cta(h, s_epoch);
}

// merged operators applied by the ordinary kernels: the current landing buffer -> the tail of x_ext
__global__ void __launch_bounds__(256)
p2p_gather_ghosts_kernel(const unsigned long long *__restrict__ epoch, const double *__restrict__ ghost, int n,
                         double *__restrict__ dst) {
    const double *src = ghost + (size_t)(epoch[1] & 1ull) * (size_t)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
p2p_release_kernel(const __grid_constant__ HaloSync h) {
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long *)&h.epoch[1];
    __syncthreads();
    sb_halo_release_cta(h, 1, s_epoch);
}

// ---------------------------------------------------------------------------------------------
// per-application steps
// ---------------------------------------------------------------------------------------------
int sb_p2p_pack(saena_b200_ctx *ctx, DevOperator &op, const double *x, cudaStream_t s) {
    if (op.sends.empty()) return 0;
    if (op.sends.size() > 256) SB_FAIL("peer-memory halo: more than 256 receivers (one spinning thread each)");
    ++ctx->launches;
    const int blocks = std::max(1, (op.vIndexSize + SB_PACK_PER_CTA - 1) / SB_PACK_PER_CTA);
    p2p_pack_kernel<<<blocks, 256, 0, s>>>(sb_halo_sync_args(ctx, op), x, op.vIndex, op.vIndexSize, !op.use_double);
    SB_CUDA(cudaGetLastError());
    return 0;
}

int sb_p2p_wait_arrived(saena_b200_ctx *ctx, DevOperator &op, cudaStream_t s) {
    if (op.recvs.empty()) return 0;
    if (op.recvs.size() > 256) SB_FAIL("peer-memory halo: more than 256 senders (one spinning thread each)");
    ++ctx->launches;
    p2p_wait_kernel<<<1, 256, 0, s>>>(sb_halo_sync_args(ctx, op));
    SB_CUDA(cudaGetLastError());
    return 0;
}

int sb_p2p_gather_ghosts(saena_b200_ctx *ctx, DevOperator &op, double *dst, cudaStream_t s) {
    if (op.recvSize == 0) return 0;
    ++ctx->launches;
    const int blocks = std::min((op.recvSize + 255) / 256, 8 * ctx->sm_count);
    p2p_gather_ghosts_kernel<<<blocks, 256, 0, s>>>(op.hs.epoch, op.ghost_d, op.recvSize, dst);
    SB_CUDA(cudaGetLastError());
    return 0;
}

int sb_p2p_release(saena_b200_ctx *ctx, DevOperator &op, cudaStream_t s) {
    if (op.recvs.empty()) return 0;
    ++ctx->launches;
    p2p_release_kernel<<<1, 32, 0, s>>>(sb_halo_sync_args(ctx, op));
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
        e[1] = op->recvSize;   // distance between this rank's two landing buffers, in doubles
        e[3] = (int64_t)op->ghost_d_off;
        for (const HaloPeer &r : op->recvs) {
            e[4 + r.peer] = r.offset;
            e[4 + ctx->nranks + r.peer] = r.count;
        }
    }
    memcpy(buf, out.data(), (size_t)*size_out);
    return 0;
}

// blobs: the export of every rank, concatenated in rank order, each `blob_bytes` long.
// Collective in the sense that no rank may apply an operator before every rank has returned from its import (the
// callers put a barrier after it): the counters of the hand-shake are reset here.
int saena_b200_p2p_import(saena_b200_ctx *ctx, const void *blobs, int64_t blob_bytes) {
    if (!ctx) return 1;
    SB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->arena) SB_FAIL("p2p_import: no halo arena (finalize first)");
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
    // open the arenas of the ranks I send to, and of the ranks I receive from (my `consumed` signals land there)
    if ((int)ctx->peer_arena.size() != N) ctx->peer_arena.assign(N, nullptr);
    auto open_peer = [&](int peer) -> int {
        if (ctx->peer_arena[peer]) return 0;
        cudaIpcMemHandle_t h;
        memcpy(&h, &blob(peer)[5], 64);
        void *ptr = nullptr;
        SB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_arena[peer] = ptr;
        return 0;
    };
    for (DevOperator *op : ops)
        if (op->present) {
            for (const HaloPeer &s : op->sends) SB_TRY(open_peer(s.peer));
            for (const HaloPeer &r : op->recvs) SB_TRY(open_peer(r.peer));
        }
    // my own counters start from zero again (a second import after a fault: every rank re-imports, then a barrier)
    SB_CUDA(cudaMemset(ctx->arena, 0, align_up(2 * n_ops * (size_t)N * sizeof(unsigned long long), 256)));
    SB_CUDA(cudaMemset(ctx->fault_dev, 0, sizeof(unsigned long long) * S_FAULT_WORDS));
    ctx->faulted = false;
    for (size_t k = 0; k < n_ops; ++k) {
        DevOperator *op = ops[k];
        if (!op->present || (op->sends.empty() && op->recvs.empty())) continue;
        std::vector<HaloSeg> segs;
        std::vector<unsigned long long *> wait_consumed, wait_arrived, signal;
        long long expect_start = 0;
        for (const HaloPeer &s : op->sends) {
            const int64_t *e = &blob(s.peer)[13 + k * per_op];
            if (!e[0] || e[4 + ctx->rank] < 0 || e[4 + N + ctx->rank] != s.count)
                SB_FAIL("p2p_import: a peer's receive plan does not match my send plan");
            if (s.offset != expect_start) SB_FAIL("p2p_import: send slices are not contiguous in rank order");
            expect_start += s.count;
            char *pa = (char *)ctx->peer_arena[s.peer];
            HaloSeg g;
            g.start = s.offset;
            g.count = s.count;
            g.dst = (double *)(pa + e[3]) + e[4 + ctx->rank];
            g.dst_stride = e[1];
            g.arrived = flag_ptr(pa, N, n_ops, 0, k, ctx->rank);
            segs.push_back(g);
            wait_consumed.push_back(flag_ptr(ctx->arena, N, n_ops, 1, k, s.peer));
        }
        if (expect_start != op->vIndexSize) SB_FAIL("p2p_import: send slices do not cover vIndex");
        for (const HaloPeer &r : op->recvs) {
            wait_arrived.push_back(flag_ptr(ctx->arena, N, n_ops, 0, k, r.peer));
            signal.push_back(flag_ptr((char *)ctx->peer_arena[r.peer], N, n_ops, 1, k, ctx->rank));
        }
        HaloSyncDev &hs = op->hs;
        free_halo_sync(hs);
        SB_CUDA(cudaMalloc((void **)&hs.epoch, 2 * sizeof(unsigned long long)));
        SB_CUDA(cudaMemset(hs.epoch, 0, 2 * sizeof(unsigned long long)));
        SB_CUDA(cudaMalloc((void **)&hs.tickets, 2 * sizeof(unsigned int)));
        SB_CUDA(cudaMemset(hs.tickets, 0, 2 * sizeof(unsigned int)));
        auto up = [&](auto **dst, const auto &v) -> int {
            typedef typename std::remove_reference<decltype(v[0])>::type T;
            SB_CUDA(cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(v.size(), 1)));
            if (!v.empty()) SB_CUDA(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
            return 0;
        };
        SB_TRY(up(&hs.segs, segs));
        SB_TRY(up(&hs.wait_consumed, wait_consumed));
        SB_TRY(up(&hs.wait_arrived, wait_arrived));
        SB_TRY(up(&hs.signal_consumed, signal));
        op->p2p = true;
        op->fused = ctx->fused_default;
    }
    SB_CUDA(cudaDeviceSynchronize());
    ctx->p2p_ready = true;
    sb_invalidate_graphs(ctx);  // a captured V-cycle holds the other transport's nodes
    return 0;
}

// 0: back to ncclSend/ncclRecv for the halo (the imported mappings stay open); 1: peer stores with
// separate launches (pack kernel, wait kernel, boundary kernel, release kernel); 2: the fused kernel.
// Collective only between 0 and non-zero: 1 and 2 are the same protocol on the wire.
int saena_b200_p2p_enable(saena_b200_ctx *ctx, int on) {
    if (!ctx) return 1;
    if (on && !ctx->p2p_ready) SB_FAIL("p2p_enable: import the peers' exports first");
    SB_CUDA(cudaSetDevice(ctx->device));
    SB_TRY(sb_sync_stream(ctx, ctx->stream));
    SB_TRY(sb_sync_stream(ctx, ctx->comm_stream));
    for (DevOperator *op : all_ops(ctx))
        if (op->present && op->hs.epoch) {
            op->p2p = on != 0;
            op->fused = on == 2;
        }
    sb_invalidate_graphs(ctx);
    return 0;
}

}  // extern "C"
