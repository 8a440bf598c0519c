// Host-side, bit-exact restatement of the draw behind split_train_val (model.py:198-210):
//
//     np.random.choice(class_indices, N, replace=False)
//
// on numpy's GLOBAL LEGACY generator.  The algorithm lives in numpy (a dependency of the
// reference, not vendored in it; numpy 2.3.5 here: numpy/random/mtrand.pyx `choice` ->
// `permutation` -> `shuffle` -> `_shuffle_raw`, numpy/random/src/distributions/distributions.c
// `random_interval`, numpy/random/src/mt19937/mt19937.c):
//
//     idx = permutation(n)[:N]        permutation(n) = shuffle(arange(n))
//     shuffle: for i = n-1 .. 1:  j_i = random_interval(i);  swap(x[i], x[j_i])
//     random_interval(max): mask = smallest 2^k-1 >= max; draw 32-bit MT19937 words until
//                           (word & mask) <= max            (max <= 0xffffffff here)
//
// This is host code: the draw is an inherently sequential walk over the MT19937 stream and it is
// the Amdahl floor of fit() once the grower runs on the GPU (numpy needs ~5 ms
// per 250k-row class).  Same stream, same indices, same end state, computed in three passes:
//
//   1. words   - MT19937 blocks of 624 are regenerated and tempered whole (auto-vectorised);
//   2. reject  - branch-free masked rejection (8 words per step with AVX2: compress the accepted
//                lanes; a vector with a lane whose fate depends on its neighbours goes scalar);
//   3. resolve - only the first N entries of the final permutation are needed.  Steps i >= N only
//                change prefix position p when j_i == p, and what lands there is whatever sat at
//                position i when step i ran, i.e. the end of the chain i -> (smallest i' > i with
//                j_i' == i) -> ...  One ascending scan over J with a "current chain head" bitmap
//                resolves all N chains; steps i < N are then simulated on the N-element prefix.
//
// State hand-off is np.random.get_state() / set_state(): key[624] + pos.  Parity:
// tests/test_host_rng.py compares indices AND end state with numpy itself.
#include <stdint.h>
#include <string.h>

#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include "../../include/rfmc_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

#if defined(__x86_64__) && defined(__GNUC__)
#define RFMC_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define RFMC_CLONES
#endif

RFMC_CLONES void mt_regen(uint32_t* k) {
    int i = 0;
    for (; i < MT_N - MT_M; ++i) {
        uint32_t y = (k[i] & UPPER_MASK) | (k[i + 1] & LOWER_MASK);
        k[i] = k[i + MT_M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    for (; i < MT_N - 1; ++i) {
        uint32_t y = (k[i] & UPPER_MASK) | (k[i + 1] & LOWER_MASK);
        k[i] = k[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
    }
    uint32_t y = (k[MT_N - 1] & UPPER_MASK) | (k[0] & LOWER_MASK);
    k[MT_N - 1] = k[MT_M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
}

RFMC_CLONES void mt_temper(const uint32_t* k, uint32_t* out) {
    for (int i = 0; i < MT_N; ++i) {
        uint32_t y = k[i];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        out[i] = y;
    }
}

struct Scratch {
    std::vector<uint32_t> A;  // accepted draws in stream order: A[k] = j_i for i = n-1-k
    std::vector<uint32_t> live;
    std::vector<int32_t> owner;
    std::vector<int32_t> head;
};
thread_local Scratch g_s;

struct Stream {  // numpy's MT19937 state + the tempered copy of the current block
    uint32_t* key;
    int pos;
    uint32_t tw[MT_N];
    void refill() {
        mt_regen(key);
        mt_temper(key, tw);
        pos = 0;
    }
};

// Branch-free scalar rejection over one mask level: consume words until i reaches lo.
inline void reject_scalar(Stream& st, uint32_t mask, int64_t lo, int64_t& i, uint32_t* A, int64_t& k, int max_words) {
    while (i > lo && max_words > 0) {
        if (st.pos == MT_N) st.refill();
        int avail = MT_N - st.pos;
        if (avail > max_words) avail = max_words;
        const uint32_t* w = st.tw + st.pos;
        int t = 0;
        while (t < avail && i > lo) {
            const uint32_t j = w[t++] & mask;
            A[k] = j;
            const int64_t acc = (int64_t)(j <= (uint32_t)i);
            k += acc;
            i -= acc;
        }
        st.pos += t;
        max_words -= t;
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
#define RFMC_HAVE_AVX2 1
alignas(32) uint32_t g_compress_lut[256][8];
bool g_lut_ready = false;
void build_lut() {
    for (int m = 0; m < 256; ++m) {
        int c = 0;
        for (int b = 0; b < 8; ++b)
            if (m & (1 << b)) g_compress_lut[m][c++] = (uint32_t)b;
        for (; c < 8; ++c) g_compress_lut[m][c] = 0;
    }
    g_lut_ready = true;
}

// 8 words at a time.  Lane l sees i - (accepted before l) >= i - 7, so u <= i-7 is accepted and
// u > i is rejected whatever the other lanes do; a vector with a lane in between goes scalar.
__attribute__((target("avx2"))) void reject_avx2(Stream& st, uint32_t mask, int64_t lo, int64_t& i, uint32_t* A,
                                                   int64_t& k) {
    const __m256i vmask = _mm256_set1_epi32((int)mask);
    while (i > lo) {
        if (st.pos == MT_N) st.refill();
        if (MT_N - st.pos < 8 || i - 8 < lo) {
            reject_scalar(st, mask, lo, i, A, k, 8);
            continue;
        }
        const __m256i u = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(st.tw + st.pos)), vmask);
        const __m256i gt_i = _mm256_cmpgt_epi32(u, _mm256_set1_epi32((int)i));
        const __m256i gt_safe = _mm256_cmpgt_epi32(u, _mm256_set1_epi32((int)(i - 7)));
        if (_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_xor_si256(gt_i, gt_safe)))) {
            reject_scalar(st, mask, lo, i, A, k, 8);
            continue;
        }
        const int m = (~_mm256_movemask_ps(_mm256_castsi256_ps(gt_i))) & 0xFF;
        const __m256i perm = _mm256_load_si256((const __m256i*)g_compress_lut[m]);
        _mm256_storeu_si256((__m256i*)(A + k), _mm256_permutevar8x32_epi32(u, perm));
        const int c = __builtin_popcount((unsigned)m);
        k += c;
        i -= c;
        st.pos += 8;
    }
}

// Ascending scan over steps q >= size (descending k): 8 membership tests per gather.
__attribute__((target("avx2"))) void resolve_avx2(const uint32_t* A, int64_t n, int64_t size, uint32_t* live,
                                                    int32_t* owner, int32_t* head) {
    int64_t k = n - 1 - size;  // q = n-1-k
    const __m256i one = _mm256_set1_epi32(1), m31 = _mm256_set1_epi32(31);
    while (k >= 0) {
        if (k >= 7) {
            const __m256i j = _mm256_loadu_si256((const __m256i*)(A + k - 7));
            const __m256i g = _mm256_i32gather_epi32((const int*)live, _mm256_srli_epi32(j, 5), 4);
            const __m256i bit = _mm256_and_si256(_mm256_srlv_epi32(g, _mm256_and_si256(j, m31)), one);
            if (_mm256_testz_si256(bit, bit)) {
                k -= 8;
                continue;
            }
        }
        const int64_t stop = k >= 7 ? k - 8 : -1;
        for (; k > stop; --k) {
            const uint32_t j = A[k];
            if ((live[j >> 5] >> (j & 31)) & 1u) {
                const int64_t q = n - 1 - k;
                const int32_t p = owner[j];
                live[j >> 5] &= ~(1u << (j & 31));
                live[q >> 5] |= 1u << (q & 31);
                owner[q] = p;
                head[p] = (int32_t)q;
            }
        }
    }
}
#endif

void resolve_scalar(const uint32_t* A, int64_t n, int64_t size, uint32_t* live, int32_t* owner, int32_t* head) {
    for (int64_t k = n - 1 - size; k >= 0; --k) {
        const uint32_t j = A[k];
        if ((live[j >> 5] >> (j & 31)) & 1u) {
            const int64_t q = n - 1 - k;
            const int32_t p = owner[j];
            live[j >> 5] &= ~(1u << (j & 31));
            live[q >> 5] |= 1u << (q & 31);
            owner[q] = p;
            head[p] = (int32_t)q;
        }
    }
}

}  // namespace

extern "C" int rfmc_legacy_permutation_prefix(uint32_t* key, int32_t* pos_io, int64_t n, int64_t size, int64_t* out) {
    if (!key || !pos_io || !out || n < 0 || size < 0 || size > n || n > 0x7fffffffll || *pos_io < 0 || *pos_io > MT_N)
        return 1;
    if (n == 0) return 0;
    if (n == 1) {
        if (size) out[0] = 0;
        return 0;
    }
    Scratch& s = g_s;
    if ((int64_t)s.A.size() < n + 16) {
        s.A.resize((size_t)n + 16);
        s.owner.resize((size_t)n);
    }
    uint32_t* A = s.A.data();
    Stream st;
    st.key = key;
    st.pos = *pos_io;
    if (st.pos < MT_N) mt_temper(key, st.tw);
    bool avx2 = false;
#ifdef RFMC_HAVE_AVX2
    avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && !g_lut_ready) build_lut();
#endif

    // ---- passes 1+2: words and rejection, i = n-1 .. 1; A[k] = j_(n-1-k) ----
    uint32_t mask = (uint32_t)(n - 1);
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    int64_t i = n - 1, k = 0;
    while (i >= 1) {
        const int64_t lo = (int64_t)(mask >> 1);  // this mask serves i in (lo, mask]
#ifdef RFMC_HAVE_AVX2
        if (avx2 && mask >= 1023)
            reject_avx2(st, mask, lo, i, A, k);
        else
#endif
            reject_scalar(st, mask, lo, i, A, k, 0x7fffffff);
        mask >>= 1;
    }

    // ---- pass 3: resolve the first `size` entries ----
    if ((int64_t)s.head.size() < size + 1) s.head.resize((size_t)size + 1);
    int32_t* head = s.head.data();
    int32_t* owner = s.owner.data();
    const size_t words = (size_t)((n + 31) / 32);
    if (s.live.size() < words) s.live.resize(words);
    uint32_t* live = s.live.data();
    memset(live, 0, words * sizeof(uint32_t));
    for (int64_t p = 0; p < size; ++p) {
        head[p] = (int32_t)p;
        owner[p] = (int32_t)p;
        live[p >> 5] |= 1u << (p & 31);
    }
#ifdef RFMC_HAVE_AVX2
    if (avx2)
        resolve_avx2(A, n, size, live, owner, head);
    else
#endif
        resolve_scalar(A, n, size, live, owner, head);
    for (int64_t q = size - 1; q >= 1; --q) {  // steps i < size act inside the prefix only
        const uint32_t j = A[n - 1 - q];
        const int32_t t = head[q];
        head[q] = head[j];
        head[j] = t;
    }
    for (int64_t p = 0; p < size; ++p) out[p] = head[p];
    *pos_io = st.pos;
    return 0;
}
