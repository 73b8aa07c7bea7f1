// spgemm.cu -- C = A B for CSR operands on the device: the Galerkin triple product of the AMG setup.
//
// What it stands for in the reference: saena_object::triple_mat_mult (/root/reference/src/saena_object_setup2.cpp:361,
// Ac = R (A P)) and the matmat machinery under it (src/saena_object_setup_matmat.cpp:27-1160), whose innermost
// product is MKL's mkl_dcsrmultcsr (:214-218).  Any exact CSR x CSR product is equivalent up to the summation order
// of each output entry, so parity is "same pattern, values within 1e-13" against the reference's own
// coarse operators (tests/golden/*.npz hold them; tests/test_zzz_spgemm_gpu.py).
//
// Row-wise Gustavson with per-row accumulators, two passes (hand-written, no cuSPARSE / CUB / Thrust):
//   symbolic  ub_i = sum over a_ik of nnz(B_k) bounds row i; rows are binned by min(ub_i, N) and each bin counts the
//             distinct columns of its rows with the cheapest structure that can hold them: a per-warp hash set
//             (<= 32, <= 128 and <= 512 candidates), a per-CTA hash set in shared memory (<= 8192), or a bitmask over
//             all N columns (shared memory up to 1.5 M columns, else a slab in global memory).  An exclusive scan of
//             the counts gives C's row offsets.
//   numeric   rows are binned again, by their exact nnz: per-warp hash tables (<= 32, <= 128, <= 512 entries), per-CTA
//             hash tables in shared memory (<= 4096 and <= 8192 entries: 96 / 192 KB), and for rows denser than that a
//             dense accumulator over all N columns (shared memory up to 24 576 columns, else a slab in global memory
//             per resident CTA).  Hash tables are sorted in place (bitonic, empty slots last) so that every output
//             row has ascending columns, as the reference's row-major operators do; the dense accumulators are read
//             out in column order.
// The deep levels of a smoothed-aggregation hierarchy are where this matters: R (A P) on level 2 of the 256^3
// Poisson hierarchy is ~1e11 scalar products landing on 1 395 entries per row -- expand / sort / compress moves every
// product through global memory several times, a shared-memory accumulator touches it once.
// Accumulation uses shared / global atomics: the summation order inside an entry is not fixed (as with MKL's
// threaded product); values agree to rounding.
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "saena_b200.h"

extern thread_local std::string g_sb_init_error;

namespace {

#define SG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_sb_init_error = std::string("spgemm: ") + #call + ": " + cudaGetErrorString(e_);     \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

constexpr int EMPTY = 0x7fffffff;  // sorts behind every column

__device__ __forceinline__ unsigned sg_hash(int c, unsigned mask) { return ((unsigned)c * 2654435761u >> 7) & mask; }

// ---------------------------------------------------------------------------------------------
// upper bound of the products of each row (warp per row)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sg_ub_kernel(int M, const int64_t *__restrict__ a_rp, const int *__restrict__ a_col, const int64_t *__restrict__ b_rp,
             long long *__restrict__ ub) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    long long s = 0;
    for (int64_t k = a_rp[row] + lane; k < a_rp[row + 1]; k += 32) {
        const int c = a_col[k];
        s += b_rp[c + 1] - b_rp[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) ub[row] = s;
}

// ---------------------------------------------------------------------------------------------
// binning: bin of a row from a size (symbolic: min(ub, N); numeric: nnz of the row)
// ---------------------------------------------------------------------------------------------
constexpr int N_BINS = 7;
struct BinLimits { long long hi[N_BINS]; };   // row goes to the first bin with size <= hi[b]; bin 0: size == 0

__device__ __forceinline__ int sg_bin_of(long long sz, const BinLimits &lim) {
    int b = 0;
    while (b < N_BINS - 1 && sz > lim.hi[b]) ++b;
    return b;
}

template <typename T>
__global__ void __launch_bounds__(256)
sg_bin_count_kernel(int M, const T *__restrict__ size, long long cap, BinLimits lim, unsigned int *__restrict__ counts) {
    __shared__ unsigned int s_c[N_BINS];
    if (threadIdx.x < N_BINS) s_c[threadIdx.x] = 0u;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        long long sz = (long long)size[i];
        if (sz > cap) sz = cap;
        atomicAdd(&s_c[sg_bin_of(sz, lim)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < N_BINS && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_c[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(256)
sg_bin_fill_kernel(int M, const T *__restrict__ size, long long cap, BinLimits lim, const unsigned int *__restrict__ start,
                   unsigned int *__restrict__ cursor, int *__restrict__ rows) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        long long sz = (long long)size[i];
        if (sz > cap) sz = cap;
        const int b = sg_bin_of(sz, lim);
        rows[start[b] + atomicAdd(&cursor[b], 1u)] = i;
    }
}

// ---------------------------------------------------------------------------------------------
// symbolic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sg_set_insert(int *tbl, unsigned mask, int c) {
    unsigned h = sg_hash(c, mask);
    while (true) {
        const int old = atomicCAS(&tbl[h], EMPTY, c);
        if (old == EMPTY || old == c) return;
        h = (h + 1u) & mask;
    }
}

// a warp per row, hash set of TBL columns per warp
template <int TBL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
sg_sym_warp_kernel(int n_rows, const int *__restrict__ rows, const int64_t *__restrict__ a_rp, const int *__restrict__ a_col,
                   const int64_t *__restrict__ b_rp, const int *__restrict__ b_col, int *__restrict__ cnt) {
    __shared__ int s_tbl[WARPS][TBL];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + w;
    if (r >= n_rows) return;
    const int row = rows[r];
    int *tbl = s_tbl[w];
    for (int i = lane; i < TBL; i += 32) tbl[i] = EMPTY;
    __syncwarp();
    for (int64_t k = a_rp[row]; k < a_rp[row + 1]; ++k) {
        const int ac = a_col[k];
        for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32) sg_set_insert(tbl, TBL - 1, b_col[j]);
    }
    __syncwarp();
    int n = 0;
    for (int i = lane; i < TBL; i += 32) n += tbl[i] != EMPTY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cnt[row] = n;
}

__device__ __forceinline__ int sg_block_sum(int v, int *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;  // every thread
}

// a CTA per row, hash set of `tbl_size` columns in dynamic shared memory
__global__ void __launch_bounds__(256)
sg_sym_cta_kernel(int n_rows, const int *__restrict__ rows, int tbl_size, const int64_t *__restrict__ a_rp,
                  const int *__restrict__ a_col, const int64_t *__restrict__ b_rp, const int *__restrict__ b_col,
                  int *__restrict__ cnt) {
    extern __shared__ int s_dyn[];
    __shared__ int s_red[8];
    const int row = rows[blockIdx.x];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < tbl_size; i += 256) s_dyn[i] = EMPTY;
    __syncthreads();
    for (int64_t k = a_rp[row] + w; k < a_rp[row + 1]; k += 8) {
        const int ac = a_col[k];
        for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32) sg_set_insert(s_dyn, (unsigned)tbl_size - 1u, b_col[j]);
    }
    __syncthreads();
    int n = 0;
    for (int i = threadIdx.x; i < tbl_size; i += 256) n += s_dyn[i] != EMPTY;
    n = sg_block_sum(n, s_red);
    if (threadIdx.x == 0) cnt[row] = n;
    (void)n_rows;
}

// a CTA per row, one bit per column of B: in dynamic shared memory (slab == nullptr) or in this CTA's slab in global
// memory (grid-stride over the rows of the bin, one slab per CTA)
__global__ void __launch_bounds__(256)
sg_sym_bitmask_kernel(int n_rows, const int *__restrict__ rows, int n_words, unsigned int *slab,
                      const int64_t *__restrict__ a_rp, const int *__restrict__ a_col, const int64_t *__restrict__ b_rp,
                      const int *__restrict__ b_col, int *__restrict__ cnt) {
    extern __shared__ int s_dyn[];
    __shared__ int s_red[8];
    unsigned int *bits = slab ? slab + (size_t)blockIdx.x * (size_t)n_words : (unsigned int *)s_dyn;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int row = rows[r];
        for (int i = threadIdx.x; i < n_words; i += 256) bits[i] = 0u;
        __syncthreads();
        for (int64_t k = a_rp[row] + w; k < a_rp[row + 1]; k += 8) {
            const int ac = a_col[k];
            for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32) {
                const int c = b_col[j];
                atomicOr(&bits[c >> 5], 1u << (c & 31));
            }
        }
        __syncthreads();
        int n = 0;
        for (int i = threadIdx.x; i < n_words; i += 256) n += __popc(bits[i]);
        n = sg_block_sum(n, s_red);
        if (threadIdx.x == 0) cnt[row] = n;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of int counts into int64 offsets (three launches: per-block totals, their scan, the offsets)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;  // elements per CTA (256 threads x 4)

__global__ void __launch_bounds__(256)
sg_scan_totals_kernel(int M, const int *__restrict__ cnt, long long *__restrict__ block_tot) {
    __shared__ long long s_w[8];
    const int base = blockIdx.x * SCAN_BLOCK;
    long long s = 0;
    for (int q = 0; q < 4; ++q) {
        const int i = base + q * 256 + threadIdx.x;
        if (i < M) s += cnt[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        block_tot[blockIdx.x] = t;
    }
}

__global__ void sg_scan_blocks_kernel(int n_blocks, long long *block_tot) {   // one thread: n_blocks <= a few 10^4
    long long run = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const long long t = block_tot[b];
        block_tot[b] = run;
        run += t;
    }
    block_tot[n_blocks] = run;
}

__global__ void __launch_bounds__(256)
sg_scan_offsets_kernel(int M, const int *__restrict__ cnt, const long long *__restrict__ block_tot, int64_t *__restrict__ rp) {
    // thread t owns 4 consecutive elements of the CTA's 1024
    __shared__ long long s_t[256];
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    int v[4];
    long long mine = 0;
    for (int q = 0; q < 4; ++q) { v[q] = base + q < M ? cnt[base + q] : 0; mine += v[q]; }
    s_t[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {   // Hillis-Steele inclusive scan of the 256 thread totals
        const long long add = threadIdx.x >= o ? s_t[threadIdx.x - o] : 0;
        __syncthreads();
        s_t[threadIdx.x] += add;
        __syncthreads();
    }
    long long off = block_tot[blockIdx.x] + s_t[threadIdx.x] - mine;
    for (int q = 0; q < 4; ++q) {
        if (base + q < M) rp[base + q] = off;
        off += v[q];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) rp[M] = block_tot[gridDim.x];
}

// ---------------------------------------------------------------------------------------------
// numeric
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sg_map_add(int *keys, double *vals, unsigned mask, int c, double v) {
    unsigned h = sg_hash(c, mask);
    while (true) {
        const int old = atomicCAS(&keys[h], EMPTY, c);
        if (old == EMPTY || old == c) { atomicAdd(&vals[h], v); return; }
        h = (h + 1u) & mask;
    }
}

// in-place bitonic sort of (key, value) pairs, ascending keys; n is a power of two; `nthreads` threads cooperate and
// `sync` separates the stages (__syncwarp for a warp's table, __syncthreads for a CTA's)
template <bool WARP>
__device__ __forceinline__ void sg_bitonic_sort(int *keys, double *vals, int n, int tid, int nthreads) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += nthreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const int a = keys[i], b = keys[ixj];
                    if ((a > b) == up) {
                        keys[i] = b; keys[ixj] = a;
                        const double t = vals[i]; vals[i] = vals[ixj]; vals[ixj] = t;
                    }
                }
            }
            if (WARP) __syncwarp(); else __syncthreads();
        }
    }
}

template <int TBL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
sg_num_warp_kernel(int n_rows, const int *__restrict__ rows, const int64_t *__restrict__ a_rp, const int *__restrict__ a_col,
                   const double *__restrict__ a_val, const int64_t *__restrict__ b_rp, const int *__restrict__ b_col,
                   const double *__restrict__ b_val, const int64_t *__restrict__ c_rp, int *__restrict__ c_col,
                   double *__restrict__ c_val) {
    extern __shared__ int s_dyn[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + w;
    if (r >= n_rows) return;
    const int row = rows[r];
    double *vals = (double *)s_dyn + (size_t)w * TBL;                    // [WARPS][TBL] doubles, then [WARPS][TBL] ints
    int *keys = (int *)((double *)s_dyn + (size_t)WARPS * TBL) + (size_t)w * TBL;
    for (int i = lane; i < TBL; i += 32) { keys[i] = EMPTY; vals[i] = 0.0; }
    __syncwarp();
    for (int64_t k = a_rp[row]; k < a_rp[row + 1]; ++k) {
        const int ac = a_col[k];
        const double av = a_val[k];
        for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32) sg_map_add(keys, vals, TBL - 1, b_col[j], av * b_val[j]);
    }
    __syncwarp();
    sg_bitonic_sort<true>(keys, vals, TBL, lane, 32);
    const int64_t o = c_rp[row];
    const int n = (int)(c_rp[row + 1] - o);
    for (int i = lane; i < n; i += 32) { c_col[o + i] = keys[i]; c_val[o + i] = vals[i]; }
}

__global__ void __launch_bounds__(256)
sg_num_cta_kernel(int n_rows, const int *__restrict__ rows, int tbl_size, const int64_t *__restrict__ a_rp,
                  const int *__restrict__ a_col, const double *__restrict__ a_val, const int64_t *__restrict__ b_rp,
                  const int *__restrict__ b_col, const double *__restrict__ b_val, const int64_t *__restrict__ c_rp,
                  int *__restrict__ c_col, double *__restrict__ c_val) {
    extern __shared__ int s_dyn[];
    const int row = rows[blockIdx.x];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *vals = (double *)s_dyn;
    int *keys = (int *)(vals + tbl_size);
    for (int i = threadIdx.x; i < tbl_size; i += 256) { keys[i] = EMPTY; vals[i] = 0.0; }
    __syncthreads();
    for (int64_t k = a_rp[row] + w; k < a_rp[row + 1]; k += 8) {
        const int ac = a_col[k];
        const double av = a_val[k];
        for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32)
            sg_map_add(keys, vals, (unsigned)tbl_size - 1u, b_col[j], av * b_val[j]);
    }
    __syncthreads();
    sg_bitonic_sort<false>(keys, vals, tbl_size, threadIdx.x, 256);
    const int64_t o = c_rp[row];
    const int n = (int)(c_rp[row + 1] - o);
    for (int i = threadIdx.x; i < n; i += 256) { c_col[o + i] = keys[i]; c_val[o + i] = vals[i]; }
    (void)n_rows;
}

// dense accumulator over all N columns + one flag bit per column; in dynamic shared memory (slab == nullptr) or in
// this CTA's slab in global memory.  Read out in column order: every thread owns a contiguous range of flag words.
__global__ void __launch_bounds__(256)
sg_num_dense_kernel(int n_rows, const int *__restrict__ rows, int N, double *slab, const int64_t *__restrict__ a_rp,
                    const int *__restrict__ a_col, const double *__restrict__ a_val, const int64_t *__restrict__ b_rp,
                    const int *__restrict__ b_col, const double *__restrict__ b_val, const int64_t *__restrict__ c_rp,
                    int *__restrict__ c_col, double *__restrict__ c_val) {
    extern __shared__ int s_dyn[];
    __shared__ int s_scan[256];
    const int n_words = (N + 31) >> 5;
    const size_t slab_doubles = (size_t)N + (size_t)((n_words + 1) >> 1);   // N accumulators + the flag words
    double *acc = slab ? slab + (size_t)blockIdx.x * slab_doubles : (double *)s_dyn;
    unsigned int *bits = (unsigned int *)(acc + N);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words_per_thread = (n_words + 255) / 256;
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int row = rows[r];
        for (int i = threadIdx.x; i < N; i += 256) acc[i] = 0.0;
        for (int i = threadIdx.x; i < n_words; i += 256) bits[i] = 0u;
        __syncthreads();
        for (int64_t k = a_rp[row] + w; k < a_rp[row + 1]; k += 8) {
            const int ac = a_col[k];
            const double av = a_val[k];
            for (int64_t j = b_rp[ac] + lane; j < b_rp[ac + 1]; j += 32) {
                const int c = b_col[j];
                atomicAdd(&acc[c], av * b_val[j]);
                atomicOr(&bits[c >> 5], 1u << (c & 31));
            }
        }
        __syncthreads();
        const int w0 = threadIdx.x * words_per_thread, w1 = min(n_words, w0 + words_per_thread);
        int mine = 0;
        for (int i = w0; i < w1; ++i) mine += __popc(bits[i]);
        s_scan[threadIdx.x] = mine;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int add = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += add;
            __syncthreads();
        }
        int64_t out = c_rp[row] + (s_scan[threadIdx.x] - mine);
        for (int i = w0; i < w1; ++i) {
            unsigned int m = bits[i];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                const int c = (i << 5) + b;
                c_col[out] = c;
                c_val[out] = acc[c];
                ++out;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
sg_row_nnz_kernel(int M, const int64_t *__restrict__ rp, int *__restrict__ cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) cnt[i] = (int)(rp[i + 1] - rp[i]);
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> int alloc(T **p, size_t n) {
        *p = nullptr;
        cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) { g_sb_init_error = std::string("spgemm: cudaMalloc: ") + cudaGetErrorString(e); return 1; }
        ptrs.push_back(*p);
        return 0;
    }
};

// rows of every bin, grouped: rows[start[b] .. start[b+1])
template <typename T>
int bin_rows(Scratch &sc, int M, const T *size_dev, long long cap, const BinLimits &lim, int **rows_out, unsigned int start[N_BINS + 1]) {
    unsigned int *d_counts = nullptr, *d_start = nullptr, *d_cursor = nullptr;
    if (sc.alloc(&d_counts, N_BINS) || sc.alloc(&d_start, N_BINS) || sc.alloc(&d_cursor, N_BINS) || sc.alloc(rows_out, (size_t)M)) return 1;
    SG_CUDA(cudaMemset(d_counts, 0, sizeof(unsigned int) * N_BINS));
    SG_CUDA(cudaMemset(d_cursor, 0, sizeof(unsigned int) * N_BINS));
    const int blocks = std::max(1, std::min((M + 255) / 256, 1184));
    sg_bin_count_kernel<T><<<blocks, 256>>>(M, size_dev, cap, lim, d_counts);
    unsigned int h[N_BINS];
    SG_CUDA(cudaMemcpy(h, d_counts, sizeof(h), cudaMemcpyDeviceToHost));
    start[0] = 0;
    for (int b = 0; b < N_BINS; ++b) start[b + 1] = start[b] + h[b];
    SG_CUDA(cudaMemcpy(d_start, start, sizeof(unsigned int) * N_BINS, cudaMemcpyHostToDevice));
    sg_bin_fill_kernel<T><<<blocks, 256>>>(M, size_dev, cap, lim, d_start, d_cursor, *rows_out);
    SG_CUDA(cudaGetLastError());
    return 0;
}

int max_dyn_smem() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

int sm_count() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
}

// slabs in global memory for the rows no shared-memory structure can hold: as many CTAs as a budget allows
const size_t SLAB_BUDGET_BYTES = (size_t)2 << 30;

}  // namespace

extern "C" {

int saena_b200_spgemm_symbolic(int M, int K, int N, const int64_t *a_rowptr, const int32_t *a_col,
                               const int64_t *b_rowptr, const int32_t *b_col, int64_t *c_rowptr, int64_t *nnz_c) {
    (void)K;
    if (M < 0 || N < 0 || !c_rowptr || !nnz_c) { g_sb_init_error = "spgemm_symbolic: bad arguments"; return 1; }
    Scratch sc;
    long long *ub = nullptr;
    int *cnt = nullptr, *rows = nullptr;
    if (sc.alloc(&ub, (size_t)M) || sc.alloc(&cnt, (size_t)M)) return 1;
    SG_CUDA(cudaMemset(cnt, 0, sizeof(int) * std::max(M, 1)));
    if (M > 0) {
        sg_ub_kernel<<<(int)(((long long)M * 32 + 255) / 256), 256>>>(M, a_rowptr, a_col, b_rowptr, ub);
        SG_CUDA(cudaGetLastError());
        // bins by min(ub, N): 0 | <= 32 | <= 128 | <= 512 | <= 8192 | bitmask in shared memory | bitmask in global memory
        const int smem = max_dyn_smem();
        const int n_words = (N + 31) / 32;
        BinLimits lim;
        lim.hi[0] = 0; lim.hi[1] = 32; lim.hi[2] = 128; lim.hi[3] = 512; lim.hi[4] = 8192;
        lim.hi[5] = ((size_t)n_words * 4 + 64 <= (size_t)smem) ? (long long)N : 8192;   // no shared bitmask: bin 5 stays empty
        lim.hi[6] = (long long)N;
        unsigned int st[N_BINS + 1];
        if (bin_rows<long long>(sc, M, ub, (long long)N, lim, &rows, st)) return 1;
        auto nb = [&](int b) { return (int)(st[b + 1] - st[b]); };
        if (nb(1)) sg_sym_warp_kernel<64, 8><<<(nb(1) + 7) / 8, 256>>>(nb(1), rows + st[1], a_rowptr, a_col, b_rowptr, b_col, cnt);
        if (nb(2)) sg_sym_warp_kernel<256, 8><<<(nb(2) + 7) / 8, 256>>>(nb(2), rows + st[2], a_rowptr, a_col, b_rowptr, b_col, cnt);
        if (nb(3)) sg_sym_warp_kernel<1024, 8><<<(nb(3) + 7) / 8, 256>>>(nb(3), rows + st[3], a_rowptr, a_col, b_rowptr, b_col, cnt);
        if (nb(4)) {
            const int tbl = 16384;
            SG_CUDA(cudaFuncSetAttribute(sg_sym_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tbl * 4));
            sg_sym_cta_kernel<<<nb(4), 256, tbl * 4>>>(nb(4), rows + st[4], tbl, a_rowptr, a_col, b_rowptr, b_col, cnt);
        }
        if (nb(5)) {
            SG_CUDA(cudaFuncSetAttribute(sg_sym_bitmask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_words * 4));
            const int grid = std::min(nb(5), 8 * sm_count());
            sg_sym_bitmask_kernel<<<grid, 256, n_words * 4>>>(nb(5), rows + st[5], n_words, nullptr, a_rowptr, a_col, b_rowptr, b_col, cnt);
        }
        if (nb(6)) {
            const int grid = (int)std::max<size_t>(1, std::min<size_t>({(size_t)nb(6), (size_t)4 * sm_count(), SLAB_BUDGET_BYTES / ((size_t)n_words * 4)}));
            unsigned int *slab = nullptr;
            if (sc.alloc(&slab, (size_t)grid * n_words)) return 1;
            sg_sym_bitmask_kernel<<<grid, 256, 0>>>(nb(6), rows + st[6], n_words, slab, a_rowptr, a_col, b_rowptr, b_col, cnt);
        }
        SG_CUDA(cudaGetLastError());
    }
    // row offsets
    const int n_blocks = std::max(1, (M + SCAN_BLOCK - 1) / SCAN_BLOCK);
    long long *tot = nullptr;
    if (sc.alloc(&tot, (size_t)n_blocks + 1)) return 1;
    sg_scan_totals_kernel<<<n_blocks, 256>>>(M, cnt, tot);
    sg_scan_blocks_kernel<<<1, 1>>>(n_blocks, tot);
    sg_scan_offsets_kernel<<<n_blocks, 256>>>(M, cnt, tot, c_rowptr);
    SG_CUDA(cudaGetLastError());
    long long total = 0;
    SG_CUDA(cudaMemcpy(&total, tot + n_blocks, sizeof(long long), cudaMemcpyDeviceToHost));
    *nnz_c = (int64_t)total;
    SG_CUDA(cudaDeviceSynchronize());
    return 0;
}

int saena_b200_spgemm_numeric(int M, int K, int N, const int64_t *a_rowptr, const int32_t *a_col, const double *a_val,
                              const int64_t *b_rowptr, const int32_t *b_col, const double *b_val,
                              const int64_t *c_rowptr, int32_t *c_col, double *c_val) {
    (void)K;
    if (M <= 0) return 0;
    Scratch sc;
    int *cnt = nullptr, *rows = nullptr;
    if (sc.alloc(&cnt, (size_t)M)) return 1;
    sg_row_nnz_kernel<<<std::max(1, std::min((M + 255) / 256, 4736)), 256>>>(M, c_rowptr, cnt);
    const int smem = max_dyn_smem();
    // bins by nnz: 0 | <= 32 | <= 128 | <= 512 | <= 4096 (96 KB table) | <= 8192 (192 KB table) | dense accumulator
    BinLimits lim;
    lim.hi[0] = 0; lim.hi[1] = 32; lim.hi[2] = 128; lim.hi[3] = 512;
    lim.hi[4] = (8192 * 12 <= smem) ? 4096 : 512;
    lim.hi[5] = (16384 * 12 <= smem) ? 8192 : lim.hi[4];
    lim.hi[6] = (long long)N;
    unsigned int st[N_BINS + 1];
    if (bin_rows<int>(sc, M, cnt, (long long)N, lim, &rows, st)) return 1;
    auto nb = [&](int b) { return (int)(st[b + 1] - st[b]); };
    if (nb(1))
        sg_num_warp_kernel<64, 8><<<(nb(1) + 7) / 8, 256, 8 * 64 * 12>>>(nb(1), rows + st[1], a_rowptr, a_col, a_val, b_rowptr,
                                                                          b_col, b_val, c_rowptr, c_col, c_val);
    if (nb(2))
        sg_num_warp_kernel<256, 8><<<(nb(2) + 7) / 8, 256, 8 * 256 * 12>>>(nb(2), rows + st[2], a_rowptr, a_col, a_val, b_rowptr,
                                                                            b_col, b_val, c_rowptr, c_col, c_val);
    if (nb(3)) {
        SG_CUDA(cudaFuncSetAttribute(sg_num_warp_kernel<1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 1024 * 12));
        sg_num_warp_kernel<1024, 4><<<(nb(3) + 3) / 4, 128, 4 * 1024 * 12>>>(nb(3), rows + st[3], a_rowptr, a_col, a_val,
                                                                              b_rowptr, b_col, b_val, c_rowptr, c_col, c_val);
    }
    for (int b = 4; b <= 5; ++b)
        if (nb(b)) {
            const int tbl = b == 4 ? 8192 : 16384;
            SG_CUDA(cudaFuncSetAttribute(sg_num_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tbl * 12));
            sg_num_cta_kernel<<<nb(b), 256, tbl * 12>>>(nb(b), rows + st[b], tbl, a_rowptr, a_col, a_val, b_rowptr, b_col, b_val,
                                                        c_rowptr, c_col, c_val);
        }
    if (nb(6)) {
        const int n_words = (N + 31) / 32;
        const size_t slab_doubles = (size_t)N + (size_t)((n_words + 1) / 2);
        const size_t bytes = slab_doubles * 8;
        if (bytes + 2048 <= (size_t)smem) {
            SG_CUDA(cudaFuncSetAttribute(sg_num_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            const int grid = std::min(nb(6), sm_count());
            sg_num_dense_kernel<<<grid, 256, bytes>>>(nb(6), rows + st[6], N, nullptr, a_rowptr, a_col, a_val, b_rowptr, b_col,
                                                      b_val, c_rowptr, c_col, c_val);
        } else {
            const int grid = (int)std::max<size_t>(1, std::min<size_t>({(size_t)nb(6), (size_t)4 * sm_count(), SLAB_BUDGET_BYTES / bytes}));
            double *slab = nullptr;
            if (sc.alloc(&slab, (size_t)grid * slab_doubles)) return 1;
            sg_num_dense_kernel<<<grid, 256, 0>>>(nb(6), rows + st[6], N, slab, a_rowptr, a_col, a_val, b_rowptr, b_col, b_val,
                                                  c_rowptr, c_col, c_val);
        }
    }
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
