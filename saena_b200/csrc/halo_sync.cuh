// halo_sync.cuh -- the device-side hand-shake of the ghost-value exchange over NVLink peer memory.
//
// What it replaces: MPI_Isend / MPI_Irecv / MPI_Waitany of saena_matrix::matvec_sparse
// (/root/reference/src/saena_matrix_matvec.cpp:32-41, :87-110; float variant :470-478) and of the P / R products
// (src/prolong_matrix.cpp:489-624, src/restrict_matrix.cpp:612-744).
//
// One protocol, used by the fused kernel (fused_halo.cu: pack, interior rows and ghost rows in one launch) and by the
// separate launches (p2p_halo.cu).  Per operator and per (sender S, receiver R) pair:
//
//   landing area     2 x recvSize doubles in R's arena; application k (0-based) of the operator lands in buffer k & 1
//   arrived[S->R]    in R's arena, written by S's last pack CTA after a system-wide fence: k + 1 once the values of
//                    application k are visible.  R's receiving role waits for arrived >= k + 1.
//   consumed[R->S]   in S's arena, written by R's last receiving CTA: k + 1 once R no longer reads buffer k & 1 of
//                    application k.  S may overwrite that buffer in application k + 2, so its pack role waits for
//                    consumed >= k - 1 -- a signal sent a whole application earlier: with two buffers the sender
//                    never waits on the receiver inside one application, and an application costs ONE NVLink
//                    crossing (values + flag) instead of the two of a single-buffered exchange.
//   epoch[2]         this rank's own application counters (pack role / receiving role), device memory, advanced by the
//                    last CTA of the role.  Everything a launch needs is read on the device: graph-replayable.
//
// All counters are monotonic.  Any mix of fused / separate launches on the two sides of a pair is legal, because both
// forms move the same counters by the same rules.
//
// Bounded waits.  A spin that outlives `timeout_ns` (globaltimer) records {what, operator, peer slot, wanted, seen}
// in the context's fault words and returns; once a fault is recorded every later wait returns at once, so the
// launch sequence of the running solve drains (with meaningless numbers) instead of hanging, each rank on its own
// clock.  The host reads the fault words with the Krylov scalars and the ABI call returns non-zero.
#pragma once

#include "common.h"

struct HaloSync {
    unsigned long long *epoch;    // [0] pack role, [1] receiving role
    unsigned int *tickets;        // [0] pack CTAs, [1] receiving CTAs
    const HaloSeg *segs;
    unsigned long long *const *wait_consumed;    // [n_segs]
    unsigned long long *const *wait_arrived;     // [n_recv]
    unsigned long long *const *signal_consumed;  // [n_recv]
    unsigned long long *fault;    // [S_FAULT_WORDS] in this rank's memory
    unsigned long long timeout_ns;
    int n_segs, n_recv;
    int op_id;                    // level * 3 + kind
};

enum { SB_FAULT_CONSUMED = 1, SB_FAULT_ARRIVED = 2 };

__device__ __forceinline__ unsigned long long sb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *f >= want; false when the wait was abandoned (timeout here, or a fault recorded earlier)
static __device__ __noinline__ bool sb_spin_until(const volatile unsigned long long *f, unsigned long long want,
                                           const HaloSync &h, int what, int slot) {
    if (*f >= want) return true;
    volatile unsigned long long *fault = h.fault;
    if (fault[0] != 0ull) return false;
    const unsigned long long t0 = sb_globaltimer();
    unsigned int n = 0;
    while (*f < want) {
        __nanosleep(32);
        if ((++n & 127u) == 0u) {
            if (fault[0] != 0ull) return false;
            if (sb_globaltimer() - t0 > h.timeout_ns) {
                const unsigned long long code = (unsigned long long)what | ((unsigned long long)(unsigned int)h.op_id << 8) |
                                                ((unsigned long long)(unsigned int)slot << 32);
                if (atomicCAS(h.fault, 0ull, code) == 0ull) {
                    fault[1] = want;
                    fault[2] = *f;
                    __threadfence();
                }
                return false;
            }
        }
    }
    return true;
}

// ---- pack role: CTA `b` of `n_pack` (256 threads, SB_PACK_PER_CTA packed values each).  `done` = applications this
//      role has completed (epoch[0]).  Publication costs ONE system-scope fence per CTA: all threads store, the CTA
//      barrier orders those stores before thread 0, whose fence (cumulative) orders them before its ticket; the CTA that
//      takes the last ticket fences once more (acquire side of the tickets) and raises the receivers' counters.  Round 2
//      measurement behind it: with a fence in every thread (262 144 of them for 512^3's level-0 halo on 8 GPUs) the
//      exchange added 24 us to a 345 us SpMV; a system-scope fence costs microseconds, and a CTA holds its SM slot
//      until its slowest thread is through.
constexpr int SB_PACK_PER_CTA = 1024;

__device__ __forceinline__ void sb_halo_pack_cta(const HaloSync &h, int b, int n_pack, unsigned long long done,
                                                 const double *__restrict__ x, const int *__restrict__ vIndex,
                                                 int vIndexSize, bool round_float) {
    const int tid = threadIdx.x;
    // buffer done & 1 was last read by the receivers in application done - 2
    if (done >= 2ull && tid < h.n_segs) sb_spin_until(h.wait_consumed[tid], done - 1ull, h, SB_FAULT_CONSUMED, tid);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SB_PACK_PER_CTA / 256; ++q) {
        const int i = b * SB_PACK_PER_CTA + q * 256 + tid;
        if (i < vIndexSize) {
            int s = 0;
            while (s + 1 < h.n_segs && i >= h.segs[s + 1].start) ++s;  // a handful of receivers
            double v = x[vIndex[i]];
            if (round_float) v = (double)(float)v;  // matvec_sparse_float (:463-464, :538): the value the receiver would widen
            h.segs[s].dst[(long long)(done & 1ull) * h.segs[s].dst_stride + (i - h.segs[s].start)] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(&h.tickets[0], 1u);
        if (t == (unsigned int)n_pack - 1u) {
            __threadfence_system();
            for (int s = 0; s < h.n_segs; ++s) *(volatile unsigned long long *)h.segs[s].arrived = done + 1ull;
            h.tickets[0] = 0u;
            *(volatile unsigned long long *)&h.epoch[0] = done + 1ull;
        }
    }
}

// ---- receiving role, entry: wait until every sender's values of application `done` have landed (whole CTA)
__device__ __forceinline__ void sb_halo_wait_cta(const HaloSync &h, unsigned long long done) {
    if ((int)threadIdx.x < h.n_recv)
        sb_spin_until(h.wait_arrived[threadIdx.x], done + 1ull, h, SB_FAULT_ARRIVED, (int)threadIdx.x);
    __threadfence();
    __syncthreads();
}

// ---- receiving role, exit: the last of `n_cta` CTAs tells the senders and advances the epoch (call from every CTA)
__device__ __forceinline__ void sb_halo_release_cta(const HaloSync &h, int n_cta, unsigned long long done) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&h.tickets[1], 1u);
        if (t == (unsigned int)n_cta - 1u) {
            for (int r = 0; r < h.n_recv; ++r) *(volatile unsigned long long *)h.signal_consumed[r] = done + 1ull;
            h.tickets[1] = 0u;
            *(volatile unsigned long long *)&h.epoch[1] = done + 1ull;
            // (no fence behind these stores: nothing of this launch depends on when they land, and a system-scope fence
            //  here is microseconds at the very end of every operator application)
        }
    }
}

// host side: the argument block of one operator (fused_halo.cu, p2p_halo.cu)
static inline HaloSync sb_halo_sync_args(const saena_b200_ctx *ctx, const DevOperator &op) {
    HaloSync h{};
    h.epoch = op.hs.epoch;
    h.tickets = op.hs.tickets;
    h.segs = op.hs.segs;
    h.wait_consumed = op.hs.wait_consumed;
    h.wait_arrived = op.hs.wait_arrived;
    h.signal_consumed = op.hs.signal_consumed;
    h.fault = ctx->fault_dev;
    h.timeout_ns = ctx->halo_timeout_ns;
    h.n_segs = (int)op.sends.size();
    h.n_recv = (int)op.recvs.size();
    h.op_id = op.op_id;
    return h;
}
