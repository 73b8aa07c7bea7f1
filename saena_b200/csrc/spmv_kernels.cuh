// spmv_kernels.cuh -- FP64 / int32-column CSR SpMV kernels for sm_100a with fused epilogues.
//
// Replaces the scalar loops of saena_matrix::matvec_sparse (/root/reference/src/
// saena_matrix_matvec.cpp:68-80 local part, :87-110 remote part), the residual variants of
// include/saena_matrix.tpp:16-43 and the update loops of saena_matrix::chebyshev / jacobi
// (src/saena_matrix.cpp:1044-1131).  All of it is HBM-bound (<= 0.17 flop/byte): no tensor
// cores; what matters is that every byte of val/col/vectors is fetched once, in full 128-byte
// lines, with enough loads in flight.
//
// Three row mappings, chosen per operator from nnz/row and row-length regularity (sb_choose_mapping):
//   spmv_sell          32-row slices stored column-major (padded per slice): lane = row, every load
//                      coalesced and independent.  Short regular rows (the 7-point level 0).
//   spmv_rowgroup<TPR> 32..256 threads per row: the deep coarse levels (thousands of nnz per row)
//   spmv_vec<LANES>    LANES (1..16) lanes cooperate on a row, a warp owns 32 consecutive rows
//                      and results are transposed so lane j finishes row j: the epilogue's
//                      vector streams (rhs, inv_diag, d, u) are read and written fully coalesced.
//   spmv_stream<LPR>   a CTA owns a block of consecutive rows holding <= STREAM_TILE non-zeros;
//                      val/col are streamed in perfectly coalesced order, products go to shared
//                      memory, then LPR lanes per row reduce their segment.  Best for short rows
//                      (7-point level 0) where a per-row mapping wastes lanes.
#pragma once

#include "common.h"

template <int EPI>
__device__ __forceinline__ void sb_epilogue(int i, double ax, const EpiArgs &e) {
    if (EPI == EPI_PLAIN) {
        e.out[i] = ax;
    } else if (EPI == EPI_RESIDUAL) {
        e.out[i] = ax - e.rhs[i];
    } else if (EPI == EPI_CHEB_FIRST) {
        const double d = e.c1 * e.inv_diag[i] * (e.rhs[i] - ax);
        e.d_out[i] = d;
        e.out[i] = e.u_in[i] + d;
    } else if (EPI == EPI_CHEB_NEXT) {
        const double res = e.c2 * e.inv_diag[i] * (e.rhs[i] - ax);
        const double d = (e.c1 * e.d_in[i]) + res;
        e.d_out[i] = d;
        e.out[i] = e.u_in[i] + d;
    } else if (EPI == EPI_JACOBI) {
        const double t = (ax - e.rhs[i]) * (e.inv_diag[i] * e.c1);
        e.out[i] = e.u_in[i] - t;
    } else if (EPI == EPI_SUB) {
        e.out[i] = e.u_in[i] - ax;
    }
}

// (Fetching the epilogue's vector inputs before the row loop was tried in round 1: +3 % on the
// 7-point level, -12 %/-20 % on the 68 and 267 nnz/row levels -- the extra live registers cost the
// loop its load batching at 32 registers/thread.  Not adopted.)
__device__ __forceinline__ bool sb_row_skipped(const uint32_t *__restrict__ mask, int row) {
    return mask != nullptr && ((mask[row >> 5] >> (row & 31)) & 1u);
}

// streaming loads of the matrix arrays: read once, do not pollute L1
__device__ __forceinline__ double sb_ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int sb_ld_stream(const int *p) { return __ldcs(p); }

// Where the gathered operand comes from.  XLocal: the rank's own vector.  XGhost (merged mode of the
// fused halo kernel, csrc/fused_halo.cu): columns >= n_local are ghost values the neighbours stored
// into this rank's landing area over NVLink.  Plain (L1-cached) loads: the gather reuses ghost
// values across rows as much as local ones (deep levels: thousands of entries per row), and no
// line of the landing area can be stale in L1 -- L1 is invalidated at every launch and within a
// launch nothing reads the area before the `arrived` flag has been observed.
struct XLocal {
    const double *x;
    __device__ __forceinline__ double ld(int c) const { return __ldg(x + c); }
};
struct XGhost {
    const double *x;
    const double *ghost;
    int n_local;
    __device__ __forceinline__ double ld(int c) const {
        const double *p = c < n_local ? x + c : ghost + (c - n_local);  // one load from a selected address: no branch
        return *p;
    }
};

// Every mapping is a __device__ body taking the CTA's index as an argument (`vb`), wrapped by a
// plain __global__ kernel (one operator application on one rank) and reused by the fused halo
// kernel, where a CTA's role -- pack, interior rows, rows waiting for ghost values -- depends on
// its index.
// ---------------------------------------------------------------------------------------------
// spmv_vec: LANES lanes per row, warp = 32 consecutive rows, transposed epilogue
// ---------------------------------------------------------------------------------------------
constexpr int VEC_UNROLL = 4;

template <int LANES, int EPI, typename OffT, typename XS>
__device__ __forceinline__ void
spmv_vec_body(int vb, int row_begin, int M, const OffT *__restrict__ rowptr, const int *__restrict__ col,
              const double *__restrict__ val, const XS xs, const EpiArgs &e,
              const uint32_t *__restrict__ skip_mask) {
    constexpr int G = 32 / LANES;  // rows in flight per warp per step
    const int lane = threadIdx.x & 31;
    const int warp = (vb * blockDim.x + threadIdx.x) >> 5;
    const int row0 = row_begin + warp * 32;  // rows [row_begin, M)
    if (row0 >= M) return;
    const int my_row = row0 + lane;
    // lane j fetches the extent of row j (coalesced), shuffled to the cooperating lanes below
    OffT my_start = 0, my_end = 0;
    if (my_row < M) {
        my_start = rowptr[my_row];
        my_end = rowptr[my_row + 1];
    }
    double mine = 0.0;
    const int g = lane / LANES;   // which row of the step this lane works on
    const int sub = lane % LANES;
#pragma unroll
    for (int t = 0; t < LANES; ++t) {
        const int src = t * G + g;  // lane that owns the row this group works on in step t
        const OffT start = __shfl_sync(0xffffffffu, my_start, src);
        const OffT end = __shfl_sync(0xffffffffu, my_end, src);
        double sum = 0.0;
        for (OffT k = start + sub; k < end; k += LANES * VEC_UNROLL) {
            // issue all loads of the trip before the first use: VEC_UNROLL independent chains
            int c[VEC_UNROLL];
            double a[VEC_UNROLL];
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) {
                const OffT kk = k + q * LANES;
                const bool in = kk < end;
                c[q] = in ? sb_ld_stream(col + kk) : 0;
                a[q] = in ? sb_ld_stream(val + kk) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) sum += a[q] * xs.ld(c[q]);
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        // lane j = t*G + g' wants the sum of group g' = j % G
        const double v = __shfl_sync(0xffffffffu, sum, (lane % G) * LANES);
        if (lane / G == t) mine = v;
    }
    if (my_row < M && !sb_row_skipped(skip_mask, my_row)) sb_epilogue<EPI>(my_row, mine, e);
}

template <int LANES, int EPI, typename OffT>
__global__ void __launch_bounds__(256)
spmv_vec_kernel(int row_begin, int M, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                const double *__restrict__ val, const double *__restrict__ x, EpiArgs e,
                const uint32_t *__restrict__ skip_mask) {
    spmv_vec_body<LANES, EPI, OffT>(blockIdx.x, row_begin, M, rowptr, col, val, XLocal{x}, e, skip_mask);
}

// ---------------------------------------------------------------------------------------------
// spmv_rowgroup: TPR = 32..256 threads per row, 256/TPR rows per CTA.  For the deep coarse levels
// of a smoothed-aggregation hierarchy: a few thousand rows with thousands of non-zeros each
// (256^3 Poisson: level 4 has 21 466 rows x 3025 nnz/row).  One row per warp-group keeps all SMs
// busy where the 32-rows-per-warp mapping above would leave most of the chip idle.
// ---------------------------------------------------------------------------------------------
template <int TPR, int EPI, typename OffT, typename XS>
__device__ __forceinline__ void
spmv_rowgroup_body(int vb, int row_begin, int M, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                   const double *__restrict__ val, const XS xs, const EpiArgs &e,
                   const uint32_t *__restrict__ skip_mask) {
    constexpr int ROWS = 256 / TPR;
    constexpr int WPR = TPR / 32;  // warps per row
    __shared__ double s_part[8];
    const int g = threadIdx.x / TPR, sub = threadIdx.x % TPR;
    const int row = row_begin + vb * ROWS + g;  // rows [row_begin, M)
    double sum = 0.0;
    if (row < M) {
        const OffT start = rowptr[row], end = rowptr[row + 1];
        for (OffT k = start + sub; k < end; k += TPR * VEC_UNROLL) {
            int c[VEC_UNROLL];
            double a[VEC_UNROLL];
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) {
                const OffT kk = k + q * TPR;
                const bool in = kk < end;
                c[q] = in ? sb_ld_stream(col + kk) : 0;
                a[q] = in ? sb_ld_stream(val + kk) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) sum += a[q] * xs.ld(c[q]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (WPR == 1) {
        if (sub == 0 && row < M && !sb_row_skipped(skip_mask, row)) sb_epilogue<EPI>(row, sum, e);
    } else {
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (sub == 0 && row < M && !sb_row_skipped(skip_mask, row)) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < WPR; ++w) tot += s_part[g * WPR + w];
            sb_epilogue<EPI>(row, tot, e);
        }
    }
}

template <int TPR, int EPI, typename OffT>
__global__ void __launch_bounds__(256)
spmv_rowgroup_kernel(int row_begin, int M, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, EpiArgs e,
                     const uint32_t *__restrict__ skip_mask) {
    spmv_rowgroup_body<TPR, EPI, OffT>(blockIdx.x, row_begin, M, rowptr, col, val, XLocal{x}, e, skip_mask);
}

// ---------------------------------------------------------------------------------------------
// spmv_stream: CTA = block of consecutive rows with <= STREAM_TILE nnz
// ---------------------------------------------------------------------------------------------
template <int LPR, int EPI, typename OffT>
__global__ void __launch_bounds__(STREAM_THREADS)
spmv_stream_kernel(int M, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                   const double *__restrict__ val, const double *__restrict__ x, EpiArgs e,
                   const uint32_t *__restrict__ skip_mask, const int *__restrict__ blk_row) {
    constexpr int ROWS = STREAM_THREADS / LPR;  // max rows per block
    __shared__ double s_prod[STREAM_TILE];
    __shared__ OffT s_rp[STREAM_THREADS + 1];
    __shared__ double s_red[STREAM_THREADS / 32];
    const int tid = threadIdx.x;
    const int r0 = blk_row[blockIdx.x];
    const int r1 = blk_row[blockIdx.x + 1];
    const int nrows = r1 - r0;  // <= ROWS
    for (int i = tid; i <= nrows; i += STREAM_THREADS) s_rp[i] = rowptr[r0 + i];
    __syncthreads();
    const OffT base = s_rp[0];
    const OffT nnzb = s_rp[nrows] - base;

    if (nnzb > STREAM_TILE) {
        // a single row longer than the tile (the host puts such a row alone in its block)
        double sum = 0.0;
        for (OffT k = tid; k < nnzb; k += STREAM_THREADS)
            sum += sb_ld_stream(val + base + k) * __ldg(x + sb_ld_stream(col + base + k));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < STREAM_THREADS / 32; ++w) tot += s_red[w];
            if (!sb_row_skipped(skip_mask, r0)) sb_epilogue<EPI>(r0, tot, e);
        }
        return;
    }

    // phase 1: products in storage order -- every load is a full coalesced line
    const int n = (int)nnzb;
    {
        constexpr int U = STREAM_TILE / STREAM_THREADS;  // every thread's share of a full tile
        int c[U];
        double a[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {  // all loads of the tile in flight before the first use
            const int k = tid + q * STREAM_THREADS;
            const bool in = k < n;
            c[q] = in ? sb_ld_stream(col + base + k) : 0;
            a[q] = in ? sb_ld_stream(val + base + k) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int k = tid + q * STREAM_THREADS;
            if (k < n) s_prod[k] = a[q] * __ldg(x + c[q]);
        }
    }
    __syncthreads();

    // phase 2: LPR lanes reduce one row's segment
    if (LPR == 1) {
        if (tid < nrows) {
            const int a = (int)(s_rp[tid] - base), b = (int)(s_rp[tid + 1] - base);
            double sum = 0.0;
            for (int k = a; k < b; ++k) sum += s_prod[k];
            const int row = r0 + tid;
            if (!sb_row_skipped(skip_mask, row)) sb_epilogue<EPI>(row, sum, e);
        }
    } else {
        __shared__ double s_rowsum[ROWS];
        const int rr = tid / LPR, sub = tid % LPR;
        double sum = 0.0;
        if (rr < nrows) {
            const int a = (int)(s_rp[rr] - base), b = (int)(s_rp[rr + 1] - base);
            for (int k = a + sub; k < b; k += LPR) sum += s_prod[k];
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (sub == 0 && rr < nrows) s_rowsum[rr] = sum;
        __syncthreads();
        if (tid < nrows) {
            const int row = r0 + tid;
            if (!sb_row_skipped(skip_mask, row)) sb_epilogue<EPI>(row, s_rowsum[tid], e);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// spmv_sell: sliced layout for short, regular rows (the 7-point level 0).  At upload the CSR
// rows are regrouped in slices of 32 consecutive rows, each slice stored column-major and padded
// to its longest row (val 0.0, col = the row's own last column), still FP64 values / int32
// columns.  Lane t owns row t of the slice: the j-th entries of the 32 rows are one 256-byte and
// one 128-byte line, every load of the loop is independent and fully coalesced, and for a
// stencil the x gather is coalesced too (neighbour j of 32 consecutive rows = 32 consecutive x).
// No shuffles, no shared memory; the epilogue streams are coalesced by construction.
// ---------------------------------------------------------------------------------------------
template <int EPI, typename XS>
__device__ __forceinline__ void
spmv_sell_body(int vb, int row_begin, int M, const long long *__restrict__ slice_ptr, const int *__restrict__ col,
               const double *__restrict__ val, const XS xs, const EpiArgs &e,
               const uint32_t *__restrict__ skip_mask) {
    const int row = row_begin + vb * blockDim.x + threadIdx.x;  // rows [row_begin, M), row_begin % 32 == 0
    const int lane = threadIdx.x & 31;
    const int slice = row >> 5;
    if ((slice << 5) >= M) return;
    const long long base = slice_ptr[slice];
    const int len = (int)((slice_ptr[slice + 1] - base) >> 5);
    const int *cp = col + base + lane;
    const double *vp = val + base + lane;
    double sum = 0.0;
    int j = 0;
    for (; j + 4 <= len; j += 4) {
        int c[4];
        double a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c[q] = sb_ld_stream(cp + (j + q) * 32);
            a[q] = sb_ld_stream(vp + (j + q) * 32);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sum += a[q] * xs.ld(c[q]);
    }
    for (; j < len; ++j) sum += sb_ld_stream(vp + j * 32) * xs.ld(sb_ld_stream(cp + j * 32));
    if (row < M && !sb_row_skipped(skip_mask, row)) sb_epilogue<EPI>(row, sum, e);
}

template <int EPI>
__global__ void __launch_bounds__(256)
spmv_sell_kernel(int row_begin, int M, const long long *__restrict__ slice_ptr, const int *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ x, EpiArgs e,
                 const uint32_t *__restrict__ skip_mask) {
    spmv_sell_body<EPI>(blockIdx.x, row_begin, M, slice_ptr, col, val, XLocal{x}, e, skip_mask);
}

// ---------------------------------------------------------------------------------------------
// spmv_sellp: the sliced layout over a row permutation (SELL-C-sigma with C = 32, sigma = 256 = one CTA).  Slot t of
// the grid works on row perm[t]; the rows of a CTA's 256 slots are the 256 consecutive rows of its window, sorted by
// length, so every slice is nearly rectangular whatever the row-length distribution; the row sums go back to row
// order through shared memory, so the epilogue's vector streams are as coalesced as in spmv_sell.  Each row is
// summed by one lane in column order, exactly as in spmv_sell: same result, bit for bit.
// ---------------------------------------------------------------------------------------------
template <int EPI, typename XS>
__device__ __forceinline__ void
spmv_sellp_body(int vb, int M, const long long *__restrict__ slice_ptr, const int *__restrict__ perm,
                const int *__restrict__ col, const double *__restrict__ val, const XS xs, const EpiArgs &e) {
    // blockDim.x == 256 == the window; the grid holds exactly ceil(M / 256) CTAs, so every slot exists
    __shared__ double s_sum[256];
    const int slot = vb * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int slice = slot >> 5;
    const long long base = slice_ptr[slice];
    const int len = (int)((slice_ptr[slice + 1] - base) >> 5);
    const int *cp = col + base + lane;
    const double *vp = val + base + lane;
    double sum = 0.0;
    int j = 0;
    for (; j + 4 <= len; j += 4) {
        int c[4];
        double a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c[q] = sb_ld_stream(cp + (j + q) * 32);
            a[q] = sb_ld_stream(vp + (j + q) * 32);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sum += a[q] * xs.ld(c[q]);
    }
    for (; j < len; ++j) sum += sb_ld_stream(vp + j * 32) * xs.ld(sb_ld_stream(cp + j * 32));
    // back to row order through shared memory: the window's rows are the 256 consecutive rows from vb * 256, each in
    // exactly one slot, so thread t finishes row vb * 256 + t and the epilogue's vector streams stay coalesced
    const int wbase = vb * 256;
    const int row = perm[slot];
    if (row >= 0) s_sum[row - wbase] = sum;
    __syncthreads();
    const int mine = wbase + threadIdx.x;
    if (mine < M) sb_epilogue<EPI>(mine, s_sum[threadIdx.x], e);
}

template <int EPI>
__global__ void __launch_bounds__(256)
spmv_sellp_kernel(int M, const long long *__restrict__ slice_ptr, const int *__restrict__ perm,
                  const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x, EpiArgs e) {
    spmv_sellp_body<EPI>(blockIdx.x, M, slice_ptr, perm, col, val, XLocal{x}, e);
}

// ---------------------------------------------------------------------------------------------
// boundary rows: rows with entries in other ranks' columns.  8 or 32 lanes per row: local segment
// (recomputed -- these rows are a few percent of the block) + remote segment read from the ghost
// buffer the halo exchange filled (float when the operator's use_double is false:
// matvec_sparse_float, saena_matrix_matvec.cpp:531-538 widens on use).
// ---------------------------------------------------------------------------------------------
template <int LANES, int EPI, typename OffT, typename GhostT>
__device__ __forceinline__ void
spmv_boundary_body(int vb, int n_brows, const int *__restrict__ brow, const OffT *__restrict__ rowptr,
                   const int *__restrict__ col, const double *__restrict__ val,
                   const int *__restrict__ brow_ptr, const int *__restrict__ bcol,
                   const double *__restrict__ bval, const double *__restrict__ x,
                   const GhostT *__restrict__ ghost, const EpiArgs &e) {
    const int gid = vb * blockDim.x + threadIdx.x;
    const int b = gid / LANES, sub = gid % LANES;
    double sum = 0.0;
    int row = -1;
    if (b < n_brows) {
        row = brow[b];
        const OffT s = rowptr[row], t = rowptr[row + 1];
        for (OffT k = s + sub; k < t; k += LANES * VEC_UNROLL) {
            int c[VEC_UNROLL];
            double a[VEC_UNROLL];
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) {
                const OffT kk = k + q * LANES;
                const bool in = kk < t;
                c[q] = in ? col[kk] : 0;
                a[q] = in ? val[kk] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < VEC_UNROLL; ++q) sum += a[q] * __ldg(x + c[q]);
        }
        const int rs = brow_ptr[b], rt = brow_ptr[b + 1];
        for (int k = rs + sub; k < rt; k += LANES) sum += bval[k] * (double)ghost[bcol[k]];
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (sub == 0 && row >= 0) sb_epilogue<EPI>(row, sum, e);
}

template <int LANES, int EPI, typename OffT, typename GhostT>
__global__ void __launch_bounds__(256)
spmv_boundary_kernel(int n_brows, const int *__restrict__ brow, const OffT *__restrict__ rowptr,
                     const int *__restrict__ col, const double *__restrict__ val,
                     const int *__restrict__ brow_ptr, const int *__restrict__ bcol,
                     const double *__restrict__ bval, const double *__restrict__ x,
                     const GhostT *__restrict__ ghost, const unsigned long long *__restrict__ epoch, int ghost_stride,
                     EpiArgs e) {
    // peer-memory exchange with separate launches (p2p_halo.cu): two landing buffers, the current one follows from
    // the receiving role's application counter; NCCL path: epoch == nullptr, one typed buffer
    if (epoch) ghost += (size_t)(epoch[1] & 1ull) * (size_t)ghost_stride;
    spmv_boundary_body<LANES, EPI, OffT, GhostT>(blockIdx.x, n_brows, brow, rowptr, col, val, brow_ptr, bcol, bval, x,
                                                 ghost, e);
}

// merged mode: ghost values that travelled as float are widened into the tail of x_ext
static __global__ void widen_ghost_kernel(int n, const float *__restrict__ in, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

// halo pack: vSend[i] = v[vIndex[i]] (saena_matrix_matvec.cpp:25-26; :463-464 casts to float)
template <typename SendT>
__global__ void halo_pack_kernel(int n, const int *__restrict__ vIndex, const double *__restrict__ v,
                                 SendT *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (SendT)v[vIndex[i]];
}
