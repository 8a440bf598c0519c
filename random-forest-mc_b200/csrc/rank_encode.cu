// process_dataset on the device: lossless rank coding of raw training columns (sm_100a).
//
// Reference boundary: RandomForestMC.process_dataset (src/random_forest_mc/model.py:162-195) keeps
// `dataset_numpy = {col: ndarray}`; every later step only compares values or takes order statistics
// of them, so a column is replaced by `code = 1 + rank among its sorted distinct values` (0 is the
// missing/unseen sentinel) plus the float64 table of those distinct values.
//
// Per numeric column, on one of two alternating streams:
//   raw column (H2D) -> order-preserving 64-bit keys -> bitonic sort (tiles of 4096 keys sorted and
//   merged in shared memory, larger strides as coalesced global passes) -> distinct keys compacted
//   in order (flag / block scan / scatter) = the value table -> code of every row by a binary search
//   of that table -> one uint32 code column.
// Categorical columns arrive as host-ranked int32 codes (strings never leave the host).
// When every column is coded the narrowest code type is known and one tiled transpose writes the
// row-major `codes[n_rows][pitch]` matrix both kernels read.
//
// All of it is byte/integer streaming work: the bound is HBM (L2 for columns that fit it).
#include "common.cuh"

namespace rfmc {

constexpr int SORT_TILE = 4096;     // keys per CTA in the shared-memory phases (32 KB)
constexpr int SORT_THREADS = 1024;
constexpr int SCAN_THREADS = 1024;
constexpr uint64_t KEY_PAD = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ double raw_as_double(const uint8_t* p, int dtype, int64_t i) {
    switch (dtype) {
        case RFMC_DT_F64: return reinterpret_cast<const double*>(p)[i];
        case RFMC_DT_F32: return (double)reinterpret_cast<const float*>(p)[i];
        case RFMC_DT_I64: return (double)reinterpret_cast<const long long*>(p)[i];
        case RFMC_DT_I32: return (double)reinterpret_cast<const int*>(p)[i];
        case RFMC_DT_I16: return (double)reinterpret_cast<const short*>(p)[i];
        case RFMC_DT_I8: return (double)reinterpret_cast<const signed char*>(p)[i];
        case RFMC_DT_U8: return (double)p[i];
        case RFMC_DT_U16: return (double)reinterpret_cast<const unsigned short*>(p)[i];
        case RFMC_DT_U32: return (double)reinterpret_cast<const unsigned int*>(p)[i];
        case RFMC_DT_U64: return (double)reinterpret_cast<const unsigned long long*>(p)[i];
    }
    return 0.0;
}

// Order-preserving map double -> uint64 (unsigned compare == numeric compare).  -0.0 and +0.0 are
// one value (one code; the table always holds +0.0), NaN is "missing" and sorts with the padding.
__device__ __forceinline__ uint64_t key_of(double x) {
    if (x != x) return KEY_PAD;
    if (x == 0.0) x = 0.0;
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) make_keys_kernel(const uint8_t* raw, int dtype, int64_t n, int64_t padded,
                                                        uint64_t* keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    keys[i] = i < n ? key_of(raw_as_double(raw, dtype, i)) : KEY_PAD;
}

__device__ __forceinline__ void compare_exchange(uint64_t& a, uint64_t& b, bool ascending) {
    if ((a > b) == ascending) {
        const uint64_t t = a;
        a = b;
        b = t;
    }
}

// Stages k = 2 .. SORT_TILE of the bitonic network, one tile per CTA, in shared memory.
__global__ void __launch_bounds__(SORT_THREADS) bitonic_tile_sort_kernel(uint64_t* keys) {
    __shared__ uint64_t s[SORT_TILE];
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = keys[base + i];
    __syncthreads();
    for (int k = 2; k <= SORT_TILE; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < SORT_TILE / 2; t += SORT_THREADS) {
                const int i = 2 * t - (t & (j - 1));
                compare_exchange(s[i], s[i + j], ((base + i) & k) == 0);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) keys[base + i] = s[i];
}

// One compare-exchange pass of stage k at a stride j >= SORT_TILE (both loads coalesced).
__global__ void __launch_bounds__(256) bitonic_global_pass_kernel(uint64_t* keys, int64_t j, int64_t k, int64_t pairs) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pairs) return;
    const int64_t i = 2 * t - (t & (j - 1));
    uint64_t a = keys[i], b = keys[i + j];
    if ((a > b) == ((i & k) == 0)) {
        keys[i] = b;
        keys[i + j] = a;
    }
}

// The strides below SORT_TILE of stage k, one tile per CTA, in shared memory.
__global__ void __launch_bounds__(SORT_THREADS) bitonic_tile_merge_kernel(uint64_t* keys, int64_t k) {
    __shared__ uint64_t s[SORT_TILE];
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = keys[base + i];
    __syncthreads();
    const bool ascending = (base & k) == 0;  // k >= 2*SORT_TILE: one direction per tile
    for (int j = SORT_TILE >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < SORT_TILE / 2; t += SORT_THREADS) {
            const int i = 2 * t - (t & (j - 1));
            compare_exchange(s[i], s[i + j], ascending);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) keys[base + i] = s[i];
}

__device__ __forceinline__ bool starts_run(const uint64_t* keys, int64_t i, int64_t padded) {
    if (i >= padded) return false;
    const uint64_t k = keys[i];
    return k != KEY_PAD && (i == 0 || keys[i - 1] != k);
}

__global__ void __launch_bounds__(SCAN_THREADS) distinct_count_kernel(const uint64_t* keys, int64_t padded,
                                                                      uint32_t* block_count) {
    const int64_t i = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const int c = __syncthreads_count(starts_run(keys, i, padded));
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}

// Exclusive scan of the per-block counts by one CTA (at most padded / 1024 entries); total -> *n_distinct.
__global__ void __launch_bounds__(SCAN_THREADS) distinct_offsets_kernel(uint32_t* block_count, int n_blocks,
                                                                        uint32_t* n_distinct) {
    __shared__ uint32_t warp_sum[SCAN_THREADS / 32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_count[i] : 0u;
        uint32_t x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sum[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, w, d);
                if (lane >= d) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (warp ? warp_sum[warp - 1] : 0u) + (x - v);
        if (i < n_blocks) block_count[i] = before;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_distinct = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) distinct_scatter_kernel(const uint64_t* keys, int64_t padded,
                                                                        const uint32_t* block_offset, double* table) {
    __shared__ uint32_t warp_sum[SCAN_THREADS / 32];
    const int64_t i = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool flag = starts_run(keys, i, padded);
    const unsigned ballot = __ballot_sync(FULL, flag);
    if (lane == 0) warp_sum[warp] = (uint32_t)__popc(ballot);
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, w, d);
            if (lane >= d) w += y;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    if (flag) {
        const uint32_t pos = block_offset[blockIdx.x] + (warp ? warp_sum[warp - 1] : 0u) +
                             (uint32_t)__popc(ballot & ((1u << lane) - 1u));
        table[pos] = value_of(keys[i]);
    }
}

// code = 1 + rank = number of table entries <= x (x is in the table unless it is NaN -> 0).
__global__ void __launch_bounds__(256) rank_kernel(const uint8_t* raw, int dtype, int64_t n, const double* table,
                                                   const uint32_t* n_distinct, uint32_t* col_codes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = raw_as_double(raw, dtype, i);
    uint32_t lo = 0, hi = *n_distinct;
    if (x != x) hi = 0;
    while (hi > lo) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(&table[mid]) <= x)
            lo = mid + 1;
        else
            hi = mid;
    }
    col_codes[i] = lo;
}

// [feature][row] uint32 code columns -> row-major codes[n_rows][pitch] in the final code type.
template <typename CodeT>
__global__ void __launch_bounds__(256) interleave_kernel(const uint32_t* col_codes, int64_t n_rows, int32_t n_features,
                                                         int64_t stride, int tile_rows, CodeT* codes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CodeT* tile = reinterpret_cast<CodeT*>(smem_raw);
    const int64_t row0 = (int64_t)blockIdx.x * tile_rows;
    const int rows = (int)((n_rows - row0) < tile_rows ? (n_rows - row0) : tile_rows);
    const int words = (int)((size_t)tile_rows * stride * sizeof(CodeT) / 4);
    for (int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0u;
    __syncthreads();
    const int total = n_features * tile_rows;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int f = i / tile_rows, r = i - f * tile_rows;
        if (r < rows) tile[(size_t)r * stride + f] = (CodeT)col_codes[(size_t)f * n_rows + row0 + r];
    }
    __syncthreads();
    const int out_words = (int)((size_t)rows * stride * sizeof(CodeT) / 4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(codes + row0 * stride);
    for (int i = threadIdx.x; i < out_words; i += blockDim.x) dst[i] = reinterpret_cast<const uint32_t*>(tile)[i];
}

int64_t rank_padded_rows(int64_t n_rows) {
    int64_t p = SORT_TILE;
    while (p < n_rows) p <<= 1;
    return p;
}

int rank_scan_blocks(int64_t padded) { return (int)((padded + SCAN_THREADS - 1) / SCAN_THREADS); }

// Everything for one numeric column, asynchronous on `stream`.  d_raw holds the column (already
// copied), d_keys[padded] and d_block[rank_scan_blocks(padded)] are scratch; d_table[n_rows]
// receives the sorted distinct values, *d_n_distinct their number, d_col_codes[n_rows] the codes.
int launch_rank_column(const void* d_raw, int dtype, int64_t n_rows, int64_t padded, uint64_t* d_keys, uint32_t* d_block,
                       double* d_table, uint32_t* d_n_distinct, uint32_t* d_col_codes, cudaStream_t stream) {
    const unsigned b256 = (unsigned)((padded + 255) / 256);
    make_keys_kernel<<<b256, 256, 0, stream>>>((const uint8_t*)d_raw, dtype, n_rows, padded, d_keys);
    const unsigned tiles = (unsigned)(padded / SORT_TILE);
    bitonic_tile_sort_kernel<<<tiles, SORT_THREADS, 0, stream>>>(d_keys);
    const int64_t pairs = padded / 2;
    const unsigned pb = (unsigned)((pairs + 255) / 256);
    for (int64_t k = 2 * (int64_t)SORT_TILE; k <= padded; k <<= 1) {
        for (int64_t j = k >> 1; j >= SORT_TILE; j >>= 1)
            bitonic_global_pass_kernel<<<pb, 256, 0, stream>>>(d_keys, j, k, pairs);
        bitonic_tile_merge_kernel<<<tiles, SORT_THREADS, 0, stream>>>(d_keys, k);
    }
    const int nb = rank_scan_blocks(padded);
    distinct_count_kernel<<<nb, SCAN_THREADS, 0, stream>>>(d_keys, padded, d_block);
    distinct_offsets_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_block, nb, d_n_distinct);
    distinct_scatter_kernel<<<nb, SCAN_THREADS, 0, stream>>>(d_keys, padded, d_block, d_table);
    rank_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>((const uint8_t*)d_raw, dtype, n_rows, d_table,
                                                                      d_n_distinct, d_col_codes);
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

int launch_interleave(const uint32_t* d_col_codes, int64_t n_rows, int32_t n_features, int64_t stride, int code_width,
                      void* d_codes, cudaStream_t stream) {
    const size_t row_bytes = (size_t)stride * code_width;
    int tile_rows = (int)((192 * 1024) / row_bytes);
    if (tile_rows > 128) tile_rows = 128;
    if (tile_rows >= 32) tile_rows &= ~31;
    if (tile_rows < 1) {
        set_error("dataset encode: too many feature columns for the shared-memory tile");
        return 1;
    }
    const size_t smem = (size_t)tile_rows * row_bytes;
    const unsigned blocks = (unsigned)((n_rows + tile_rows - 1) / tile_rows);
    switch (code_width) {
        case 1:
            RFMC_CUDA(cudaFuncSetAttribute(interleave_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            interleave_kernel<uint8_t><<<blocks, 256, smem, stream>>>(d_col_codes, n_rows, n_features, stride, tile_rows, (uint8_t*)d_codes);
            break;
        case 2:
            RFMC_CUDA(cudaFuncSetAttribute(interleave_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            interleave_kernel<uint16_t><<<blocks, 256, smem, stream>>>(d_col_codes, n_rows, n_features, stride, tile_rows, (uint16_t*)d_codes);
            break;
        case 4:
            RFMC_CUDA(cudaFuncSetAttribute(interleave_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            interleave_kernel<uint32_t><<<blocks, 256, smem, stream>>>(d_col_codes, n_rows, n_features, stride, tile_rows, (uint32_t*)d_codes);
            break;
        default: set_error("unsupported code width"); return 1;
    }
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rfmc
