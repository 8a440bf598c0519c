// Host-side, bit-exact restatement of the numpy draws behind split_train_val / sampleFeats
// (model.py:198-219):
//
//     np.random.choice(class_indices, N, replace=False)       once per class per tree slot
//     np.random.randint(min_feature, max_feature + 1)         once per candidate
//
// on numpy's GLOBAL LEGACY generator.  The algorithms live in numpy (a dependency of the
// reference, not vendored in it; numpy 2.3.5 here: numpy/random/mtrand.pyx `choice` ->
// `permutation` -> `shuffle` -> `_shuffle_raw`, `randint` -> `_bounded_integers._rand_int64`,
// numpy/random/src/distributions/distributions.c `random_interval` /
// `buffered_bounded_masked_uint32`, numpy/random/src/mt19937/mt19937.c):
//
//     idx = permutation(n)[:N]        permutation(n) = shuffle(arange(n))
//     shuffle: for i = n-1 .. 1:  j_i = random_interval(i);  swap(x[i], x[j_i])
//     random_interval(max): mask = smallest 2^k-1 >= max; draw 32-bit MT19937 words until
//                           (word & mask) <= max            (max <= 0xffffffff here)
//     randint(lo, hi):      lo + the same masked rejection with max = hi-1-lo (no draw if 0)
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
//                j_i' == i) -> ...  One ascending scan over the draws with a "current chain head"
//                bitmap resolves all N chains; steps i < N are then simulated on the prefix.
//
// Passes 1+2 are the sequential part (they decide where the next draw starts in the stream);
// pass 3 only reads that draw's accepted words, so rfmc_legacy_draw_plan hands it to worker
// threads while the main thread already walks the stream for the next class / slot.
//
// State hand-off is np.random.get_state() / set_state(): key[624] + pos.  Parity:
// tests/test_host_rng.py compares indices, randint values AND end state with numpy itself.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define RFMC_HAVE_AVX2 1
#endif

#include "../../include/rfmc_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

#if defined(__x86_64__) && defined(__GNUC__)
#define RFMC_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
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

#ifdef RFMC_HAVE_AVX2
// AVX-512: next block and its tempered copy in one pass (the new words are tempered in registers).
__attribute__((target("avx512f"))) void mt_regen_temper_avx512(uint32_t* k, uint32_t* out) {
    const __m512i up = _mm512_set1_epi32((int)UPPER_MASK), one = _mm512_set1_epi32(1), va = _mm512_set1_epi32((int)MATRIX_A);
    const __m512i c1 = _mm512_set1_epi32((int)0x9d2c5680u), c2 = _mm512_set1_epi32((int)0xefc60000u);
#define RFMC_MT_STEP(i, from)                                                                              \
    {                                                                                                      \
        const __m512i a = _mm512_loadu_si512(k + (i)), b = _mm512_loadu_si512(k + (i) + 1);                 \
        const __m512i m = _mm512_loadu_si512(k + (from));                                                  \
        const __m512i y = _mm512_ternarylogic_epi32(up, a, b, 0xCA); /* (a & UPPER) | (b & LOWER) */        \
        __m512i t = _mm512_xor_si512(m, _mm512_srli_epi32(y, 1));                                          \
        t = _mm512_mask_xor_epi32(t, _mm512_test_epi32_mask(y, one), t, va);                               \
        _mm512_storeu_si512(k + (i), t);                                                                   \
        t = _mm512_xor_si512(t, _mm512_srli_epi32(t, 11));                                                 \
        t = _mm512_ternarylogic_epi32(t, _mm512_slli_epi32(t, 7), c1, 0x78); /* t ^ ((t << 7) & c1) */     \
        t = _mm512_ternarylogic_epi32(t, _mm512_slli_epi32(t, 15), c2, 0x78);                              \
        t = _mm512_xor_si512(t, _mm512_srli_epi32(t, 18));                                                 \
        _mm512_storeu_si512(out + (i), t);                                                                 \
    }
#define RFMC_MT_SCALAR(i, from, next)                                  \
    {                                                                  \
        uint32_t y = (k[(i)] & UPPER_MASK) | (k[(next)] & LOWER_MASK); \
        y = k[(from)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);       \
        k[(i)] = y;                                                    \
        y ^= (y >> 11);                                                \
        y ^= (y << 7) & 0x9d2c5680u;                                   \
        y ^= (y << 15) & 0xefc60000u;                                  \
        y ^= (y >> 18);                                                \
        out[(i)] = y;                                                  \
    }
    int i = 0;
    for (; i + 16 <= MT_N - MT_M; i += 16) RFMC_MT_STEP(i, i + MT_M)
    for (; i < MT_N - MT_M; ++i) RFMC_MT_SCALAR(i, i + MT_M, i + 1)
    for (; i + 16 <= MT_N - 1; i += 16) RFMC_MT_STEP(i, i + (MT_M - MT_N))  // reads words written >= 227 - 15 steps ago
    for (; i < MT_N - 1; ++i) RFMC_MT_SCALAR(i, i + (MT_M - MT_N), i + 1)
    RFMC_MT_SCALAR(MT_N - 1, MT_M - 1, 0)
#undef RFMC_MT_STEP
#undef RFMC_MT_SCALAR
}
#endif

inline void mt_next_block(uint32_t* k, uint32_t* out) {
#ifdef RFMC_HAVE_AVX2
    static const bool avx512 = __builtin_cpu_supports("avx512f");
    if (avx512) {
        mt_regen_temper_avx512(k, out);
        return;
    }
#endif
    mt_regen(k);
    mt_temper(k, out);
}

// A ring of MT19937 blocks filled by a producer thread: block b+1 = regen(block b), tempered.  The
// stream itself does not depend on how it is consumed, so generation runs ahead of the rejection
// walk; the consumer's end state is simply (untempered key of its current block, position).
struct BlockRing {
    static constexpr int K = 64;
    uint32_t key[K][MT_N];
    uint32_t tw[K][MT_N];
    std::atomic<int64_t> produced{1};  // blocks 0 .. produced-1 are ready
    std::atomic<int64_t> current{0};   // block the consumer is reading
    std::atomic<bool> stop{false};
    std::thread worker;
    void start(const uint32_t* key0) {
        memcpy(key[0], key0, sizeof(key[0]));
        mt_temper(key[0], tw[0]);
        worker = std::thread([this] {
            for (int64_t b = 1; !stop.load(std::memory_order_relaxed); ++b) {
                while (b - current.load(std::memory_order_acquire) >= K) {
                    if (stop.load(std::memory_order_relaxed)) return;
                    std::this_thread::yield();
                }
                uint32_t* k = key[b % K];
                memcpy(k, key[(b - 1) % K], sizeof(key[0]));
                mt_next_block(k, tw[b % K]);
                produced.store(b + 1, std::memory_order_release);
            }
        });
    }
    void finish() {
        stop.store(true);
        if (worker.joinable()) worker.join();
    }
};

struct Stream {  // numpy's MT19937 state + the tempered copy of the current block
    uint32_t* key;
    const uint32_t* tw;
    int pos;
    BlockRing* ring = nullptr;
    int64_t block = 0;
    uint32_t own_tw[MT_N];
    void attach(uint32_t* k, int p) {
        key = k;
        pos = p;
        tw = own_tw;
        if (pos < MT_N) mt_temper(key, own_tw);
    }
    void attach_ring(BlockRing* r, int p) {
        ring = r;
        block = 0;
        key = r->key[0];
        tw = r->tw[0];
        pos = p;
    }
    void refill() {
        if (ring) {
            ++block;
            ring->current.store(block, std::memory_order_release);
            while (ring->produced.load(std::memory_order_acquire) <= block) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            key = ring->key[block % BlockRing::K];
            tw = ring->tw[block % BlockRing::K];
        } else {
            mt_next_block(key, own_tw);
        }
        pos = 0;
    }
    uint32_t next() {
        if (pos == MT_N) refill();
        return tw[pos++];
    }
};

inline uint32_t all_ones_at_least(uint32_t v) {
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    return v;
}

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

#ifdef RFMC_HAVE_AVX2
alignas(32) uint32_t g_compress_lut[256][8];
std::once_flag g_lut_once;
void build_lut() {
    for (int m = 0; m < 256; ++m) {
        int c = 0;
        for (int b = 0; b < 8; ++b)
            if (m & (1 << b)) g_compress_lut[m][c++] = (uint32_t)b;
        for (; c < 8; ++c) g_compress_lut[m][c] = 0;
    }
}

// 8 words at a time.  Lane l sees i - (accepted before l) >= i - 7, so u <= i-7 is accepted and
// u > i is rejected whatever the other lanes do; a vector with a lane in between goes scalar.
__attribute__((target("avx2"))) void reject_avx2(Stream& st, uint32_t mask, int64_t lo, int64_t& i, uint32_t* A,
                                                   int64_t& k) {
    const __m256i vmask = _mm256_set1_epi32((int)mask);
    const bool wide = mask >= 32767;  // 32 words per step while a neighbour-dependent lane stays rare
    while (i > lo) {
        if (st.pos == MT_N) st.refill();
        if (wide && MT_N - st.pos >= 32 && i - 32 >= lo) {
            // same argument with a window of 32: u <= i-31 is accepted, u > i rejected, whatever the
            // other 31 lanes do -> one loop-carried update of i per 32 words
            const __m256i* w = (const __m256i*)(st.tw + st.pos);
            const __m256i vi = _mm256_set1_epi32((int)i), vs = _mm256_set1_epi32((int)(i - 31));
            const __m256i u0 = _mm256_and_si256(_mm256_loadu_si256(w + 0), vmask);
            const __m256i u1 = _mm256_and_si256(_mm256_loadu_si256(w + 1), vmask);
            const __m256i u2 = _mm256_and_si256(_mm256_loadu_si256(w + 2), vmask);
            const __m256i u3 = _mm256_and_si256(_mm256_loadu_si256(w + 3), vmask);
            const __m256i g0 = _mm256_cmpgt_epi32(u0, vi), g1 = _mm256_cmpgt_epi32(u1, vi);
            const __m256i g2 = _mm256_cmpgt_epi32(u2, vi), g3 = _mm256_cmpgt_epi32(u3, vi);
            const __m256i amb = _mm256_or_si256(
                _mm256_or_si256(_mm256_xor_si256(g0, _mm256_cmpgt_epi32(u0, vs)), _mm256_xor_si256(g1, _mm256_cmpgt_epi32(u1, vs))),
                _mm256_or_si256(_mm256_xor_si256(g2, _mm256_cmpgt_epi32(u2, vs)), _mm256_xor_si256(g3, _mm256_cmpgt_epi32(u3, vs))));
            if (_mm256_testz_si256(amb, amb)) {
                const int m0 = (~_mm256_movemask_ps(_mm256_castsi256_ps(g0))) & 0xFF;
                const int m1 = (~_mm256_movemask_ps(_mm256_castsi256_ps(g1))) & 0xFF;
                const int m2 = (~_mm256_movemask_ps(_mm256_castsi256_ps(g2))) & 0xFF;
                const int m3 = (~_mm256_movemask_ps(_mm256_castsi256_ps(g3))) & 0xFF;
                const int c0 = __builtin_popcount((unsigned)m0), c1 = __builtin_popcount((unsigned)m1);
                const int c2 = __builtin_popcount((unsigned)m2), c3 = __builtin_popcount((unsigned)m3);
                uint32_t* dst = A + k;
                _mm256_storeu_si256((__m256i*)dst,
                                    _mm256_permutevar8x32_epi32(u0, _mm256_load_si256((const __m256i*)g_compress_lut[m0])));
                _mm256_storeu_si256((__m256i*)(dst + c0),
                                    _mm256_permutevar8x32_epi32(u1, _mm256_load_si256((const __m256i*)g_compress_lut[m1])));
                _mm256_storeu_si256((__m256i*)(dst + c0 + c1),
                                    _mm256_permutevar8x32_epi32(u2, _mm256_load_si256((const __m256i*)g_compress_lut[m2])));
                _mm256_storeu_si256((__m256i*)(dst + c0 + c1 + c2),
                                    _mm256_permutevar8x32_epi32(u3, _mm256_load_si256((const __m256i*)g_compress_lut[m3])));
                const int c = c0 + c1 + c2 + c3;
                k += c;
                i -= c;
                st.pos += 32;
                continue;
            }
        }
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

// AVX-512: up to 64 words per loop-carried update of i (a block of 624 words is 9 groups of four vectors and one of
// three); accepted lanes leave through vpcompressd.  Lane l of a group of W words sees i - (accepts before l) >= i - (W-1):
// a word u <= i - (W-1) is accepted and a word u > i rejected whatever the other lanes do.  The few lanes in between
// (about W^2 / (mask + 1) per group) are settled exactly, in lane order, against the accepts counted so far -- so the
// window stays at 64 words down to the small masks instead of shrinking with them.  The last draws of a level (fewer
// than W steps left) are left to the paths below.
__attribute__((target("avx512f,avx2,bmi,bmi2,popcnt,lzcnt"))) void reject_avx512(Stream& st, uint32_t mask, int64_t lo, int64_t& i,
                                                                                 uint32_t* A, int64_t& k) {
    const __m512i vmask = _mm512_set1_epi32((int)mask);
    while (i > lo) {
        if (st.pos == MT_N) st.refill();
        const int avail = MT_N - st.pos;
        if (avail < 16) {
            reject_scalar(st, mask, lo, i, A, k, avail);
            continue;
        }
        int nv = avail >> 4;
        if (nv > 4) nv = 4;
        const int W = nv << 4;
        if (i - W < lo) break;
        const uint32_t* w = st.tw + st.pos;
        const __m512i vi = _mm512_set1_epi32((int)i), vs = _mm512_set1_epi32((int)(i - (W - 1)));
        __m512i u[4];
        uint64_t rej = 0, notsure = 0;
#pragma GCC unroll 4
        for (int v = 0; v < 4; ++v) {
            if (v < nv) {
                u[v] = _mm512_and_si512(_mm512_loadu_si512(w + 16 * v), vmask);
                rej |= (uint64_t)_mm512_cmpgt_epi32_mask(u[v], vi) << (16 * v);
                notsure |= (uint64_t)_mm512_cmpgt_epi32_mask(u[v], vs) << (16 * v);
            }
        }
        const uint64_t all = W == 64 ? ~0ull : ((1ull << W) - 1);
        uint64_t acc = ~notsure & all;  // accepted whatever the other lanes do
        uint64_t am = notsure & ~rej;   // i - (W-1) < u <= i: depends on the accepts before it
        while (am) {
            const int l = __builtin_ctzll(am);
            am &= am - 1;
            const int before = __builtin_popcountll(acc & ((1ull << l) - 1));
            if ((w[l] & mask) <= (uint32_t)(i - before)) acc |= 1ull << l;
        }
        uint32_t* dst = A + k;
        int c = 0;
#pragma GCC unroll 4
        for (int v = 0; v < 4; ++v) {
            if (v < nv) {
                const __mmask16 m = (__mmask16)(acc >> (16 * v));
                _mm512_storeu_si512(dst + c, _mm512_maskz_compress_epi32(m, u[v]));
                c += __builtin_popcount((unsigned)m);
            }
        }
        k += c;
        i -= c;
        st.pos += W;
    }
}

// Ascending scan over steps q >= size (descending k).  About size * ln(n / size) steps hit a live
// position, and a hit rewrites the bitmap the next membership test reads: testing and updating in
// one loop costs a pipeline flush per hit (a gather cannot take its data from the store buffer).
// So the scan runs in chunks: the vector pass tests a whole chunk against the bitmap as it was when
// the chunk began and only appends suspects -- bit set then, or j inside the chunk's own step
// range (the only positions that can become live meanwhile) -- to a list; the scalar pass then
// applies the suspects in order against the current bitmap.  Chunks grow with q so that the second
// kind of suspect stays rare.
inline void apply_hit(const uint32_t* A, int64_t n, int64_t k, uint32_t* live, int32_t* owner, int32_t* head) {
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

inline int64_t chunk_len(int64_t q0) {
    int64_t c = q0 / 64;
    c = c < 256 ? 256 : (c > 8192 ? 8192 : c);
    return c & ~(int64_t)15;
}

__attribute__((target("avx512f"))) void resolve_avx512(const uint32_t* A, int64_t n, int64_t size, uint32_t* live,
                                                        int32_t* owner, int32_t* head, int32_t* suspects) {
    int64_t k = n - 1 - size;  // q = n-1-k
    const __m512i m31 = _mm512_set1_epi32(31), one = _mm512_set1_epi32(1);
    const __m512i rev = _mm512_setr_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    while (k >= 15) {
        const int64_t q0 = n - 1 - k;
        int64_t len = chunk_len(q0);
        if (len > k + 1) len = (k + 1) & ~(int64_t)15;
        if (len == 0) break;
        const __m512i vq0 = _mm512_set1_epi32((int)q0);
        int cnt = 0;
        for (int64_t kk = k; kk > k - len; kk -= 16) {
            // lane 0 = A[kk] (first in scan order)
            const __m512i j = _mm512_permutexvar_epi32(rev, _mm512_loadu_si512((const void*)(A + kk - 15)));
            const __m512i g = _mm512_i32gather_epi32(_mm512_srli_epi32(j, 5), (const int*)live, 4);
            const __mmask16 hit = _mm512_test_epi32_mask(_mm512_srlv_epi32(g, _mm512_and_si512(j, m31)), one);
            const __mmask16 late = _mm512_cmpge_epu32_mask(j, vq0);
            const __mmask16 m = hit | late;
            const __m512i ks = _mm512_sub_epi32(_mm512_set1_epi32((int)kk), _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15));
            _mm512_mask_compressstoreu_epi32(suspects + cnt, m, ks);
            cnt += __builtin_popcount((unsigned)m);
        }
        for (int t = 0; t < cnt; ++t) apply_hit(A, n, suspects[t], live, owner, head);
        k -= len;
    }
    for (; k >= 0; --k) apply_hit(A, n, k, live, owner, head);
}

__attribute__((target("avx2"))) void resolve_avx2(const uint32_t* A, int64_t n, int64_t size, uint32_t* live,
                                                    int32_t* owner, int32_t* head, int32_t* suspects) {
    int64_t k = n - 1 - size;
    const __m256i m31 = _mm256_set1_epi32(31);
    while (k >= 15) {
        const int64_t q0 = n - 1 - k;
        int64_t len = chunk_len(q0);
        if (len > k + 1) len = (k + 1) & ~(int64_t)15;
        if (len == 0) break;
        const __m256i vq0 = _mm256_set1_epi32((int)q0 - 1);
        int cnt = 0;
        for (int64_t kk = k; kk > k - len; kk -= 8) {
            const __m256i j = _mm256_loadu_si256((const __m256i*)(A + kk - 7));  // lane 7 = A[kk]
            const __m256i g = _mm256_i32gather_epi32((const int*)live, _mm256_srli_epi32(j, 5), 4);
            const __m256i sh = _mm256_sllv_epi32(g, _mm256_sub_epi32(m31, _mm256_and_si256(j, m31)));
            const __m256i late = _mm256_cmpgt_epi32(j, vq0);  // j < 2^31 here
            unsigned m = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_or_si256(sh, late)));
            while (m) {  // highest lane first = scan order
                const int L = 31 - __builtin_clz(m);
                suspects[cnt++] = (int32_t)(kk - 7 + L);
                m &= ~(1u << L);
            }
        }
        for (int t = 0; t < cnt; ++t) apply_hit(A, n, suspects[t], live, owner, head);
        k -= len;
    }
    for (; k >= 0; --k) apply_hit(A, n, k, live, owner, head);
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

bool have_avx2() {
#ifdef RFMC_HAVE_AVX2
    static const bool ok = __builtin_cpu_supports("avx2");
    if (ok) std::call_once(g_lut_once, build_lut);
    return ok;
#else
    return false;
#endif
}

// passes 1+2: A[k] = j_(n-1-k), k = 0 .. n-2.  A needs n + 64 entries (vector stores overshoot).
void draw_accepted(Stream& st, int64_t n, uint32_t* A) {
    if (n < 2) return;
    const bool avx2 = have_avx2();
#ifdef RFMC_HAVE_AVX2
    static const bool avx512 = __builtin_cpu_supports("avx512f") && avx2;
#endif
    uint32_t mask = all_ones_at_least((uint32_t)(n - 1));
    int64_t i = n - 1, k = 0;
    while (i >= 1) {
        const int64_t lo = (int64_t)(mask >> 1);  // this mask serves i in (lo, mask]
#ifdef RFMC_HAVE_AVX2
        if (avx512 && mask >= 1023) reject_avx512(st, mask, lo, i, A, k);  // leaves the end of the level to the paths below
        if (avx2 && mask >= 1023)
            reject_avx2(st, mask, lo, i, A, k);
        else
#endif
            reject_scalar(st, mask, lo, i, A, k, 0x7fffffff);
        mask >>= 1;
    }
}

struct Scratch {
    std::vector<uint32_t> A, live;
    std::vector<int32_t> owner, head, suspects;
};
thread_local Scratch g_s;

// pass 3: head[0..size) = permutation(n)[:size] given the accepted draws.
const int32_t* resolve_prefix(const uint32_t* A, int64_t n, int64_t size) {
    Scratch& s = g_s;
    if ((int64_t)s.head.size() < size + 1) s.head.resize((size_t)size + 1);
    if ((int64_t)s.owner.size() < n) s.owner.resize((size_t)n);
    const size_t words = (size_t)((n + 31) / 32);
    if (s.live.size() < words) s.live.resize(words);
    int32_t* head = s.head.data();
    int32_t* owner = s.owner.data();
    uint32_t* live = s.live.data();
    memset(live, 0, words * sizeof(uint32_t));
    for (int64_t p = 0; p < size; ++p) {
        head[p] = (int32_t)p;
        owner[p] = (int32_t)p;
        live[p >> 5] |= 1u << (p & 31);
    }
    if (n >= 2) {
#ifdef RFMC_HAVE_AVX2
        if (have_avx2()) {
            if (s.suspects.size() < 8192 + 64) s.suspects.resize(8192 + 64);
            static const bool avx512 = __builtin_cpu_supports("avx512f");
            if (avx512)
                resolve_avx512(A, n, size, live, owner, head, s.suspects.data());
            else
                resolve_avx2(A, n, size, live, owner, head, s.suspects.data());
        } else
#endif
            resolve_scalar(A, n, size, live, owner, head);
        for (int64_t q = size - 1; q >= 1; --q) {  // steps i < size act inside the prefix only
            const uint32_t j = A[n - 1 - q];
            const int32_t t = head[q];
            head[q] = head[j];
            head[j] = t;
        }
    }
    return head;
}

struct Task {
    uint32_t* A = nullptr;
    int64_t n = 0, size = 0, n_train = 0;
    const int64_t* rows = nullptr;
    int32_t* out_train = nullptr;
    int32_t* out_val = nullptr;
};

void run_task(const Task& t) {
    const int32_t* head = resolve_prefix(t.A, t.n, t.size);
    for (int64_t p = 0; p < t.size; ++p) {
        const int32_t row = (int32_t)t.rows[head[p]];
        if (p < t.n_train)
            t.out_train[p] = row;
        else
            t.out_val[p - t.n_train] = row;
    }
}

struct Pool {  // bounded hand-off of resolve tasks to worker threads
    std::mutex mu;
    std::condition_variable cv_task, cv_buf;
    std::deque<Task> tasks;
    std::vector<uint32_t*> free_bufs;
    bool closing = false;
    void worker() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_task.wait(lk, [&] { return closing || !tasks.empty(); });
                if (tasks.empty()) return;
                t = tasks.front();
                tasks.pop_front();
            }
            run_task(t);
            {
                std::lock_guard<std::mutex> lk(mu);
                free_bufs.push_back(t.A);
            }
            cv_buf.notify_one();
        }
    }
};

}  // namespace

extern "C" int rfmc_legacy_permutation_prefix(uint32_t* key, int32_t* pos_io, int64_t n, int64_t size, int64_t* out) {
    if (!key || !pos_io || !out || n < 0 || size < 0 || size > n || n > 0x7fffffffll || *pos_io < 0 || *pos_io > MT_N)
        return 1;
    if (n == 0) return 0;
    Scratch& s = g_s;
    if ((int64_t)s.A.size() < n + 128) s.A.resize((size_t)n + 128);
    Stream st;
    st.attach(key, *pos_io);
    draw_accepted(st, n, s.A.data());
    const int32_t* head = resolve_prefix(s.A.data(), n, size);
    for (int64_t p = 0; p < size; ++p) out[p] = head[p];
    *pos_io = st.pos;
    return 0;
}

extern "C" int rfmc_legacy_draw_plan(uint32_t* key, int32_t* pos_io, int32_t n_slots, int32_t n_classes,
                                     const int64_t* class_rows, const int64_t* class_off, int64_t per_class,
                                     int64_t n_train_pclass, int32_t n_cands, int64_t randint_low, int64_t randint_high,
                                     int32_t* idx_train, int32_t* idx_val, int32_t* feat_counts, uint32_t* key_snap,
                                     int32_t* pos_snap, int32_t n_threads) {
    if (!key || !pos_io || !class_rows || !class_off || n_slots < 0 || n_classes < 1 || per_class < 0 ||
        n_train_pclass < 0 || n_cands < 0 || randint_high <= randint_low || *pos_io < 0 || *pos_io > MT_N)
        return 1;
    int64_t max_n = 0;
    for (int c = 0; c < n_classes; ++c) {
        const int64_t n = class_off[c + 1] - class_off[c];
        if (n < per_class || n > 0x7fffffffll) return 1;
        if (n > max_n) max_n = n;
    }
    const int64_t n_tr = per_class < n_train_pclass ? per_class : n_train_pclass;  // selected[:batch_train]
    const int64_t n_va = per_class - n_tr;                                          // selected[batch_train:]
    const uint64_t rng = (uint64_t)(randint_high - 1 - randint_low);
    if (rng > 0xFFFFFFFEull) return 1;
    const uint32_t rmask = all_ones_at_least((uint32_t)rng);
    const bool use_ring = n_threads >= 0x100;  // bit 8: MT19937 blocks from a producer thread
    n_threads &= 0xFF;
    if (n_threads > 32) n_threads = 32;

    Pool pool;
    std::vector<std::vector<uint32_t>> bufs((size_t)(n_threads ? 2 * n_threads + 2 : 1));
    for (auto& b : bufs) {
        b.resize((size_t)max_n + 128);
        pool.free_bufs.push_back(b.data());
    }
    std::vector<std::thread> workers;
    for (int t = 0; t < n_threads; ++t) workers.emplace_back([&pool] { pool.worker(); });

    Stream st;
    std::unique_ptr<BlockRing> ring;
    if (use_ring) {  // MT19937 blocks are generated ahead of the walk by a producer thread
        ring.reset(new BlockRing());
        ring->start(key);
        st.attach_ring(ring.get(), *pos_io);
    } else {
        st.attach(key, *pos_io);
    }
    for (int32_t s = 0; s < n_slots; ++s) {
        if (key_snap) memcpy(key_snap + (size_t)s * MT_N, st.key, MT_N * sizeof(uint32_t));
        if (pos_snap) pos_snap[s] = st.pos;
        for (int c = 0; c < n_classes; ++c) {
            Task t;
            t.n = class_off[c + 1] - class_off[c];
            t.size = per_class;
            t.n_train = n_tr;
            t.rows = class_rows + class_off[c];
            t.out_train = idx_train + ((size_t)s * n_classes + c) * n_tr;
            t.out_val = idx_val + ((size_t)s * n_classes + c) * n_va;
            if (n_threads) {
                std::unique_lock<std::mutex> lk(pool.mu);
                pool.cv_buf.wait(lk, [&] { return !pool.free_bufs.empty(); });
                t.A = pool.free_bufs.back();
                pool.free_bufs.pop_back();
            } else {
                t.A = bufs[0].data();
            }
            draw_accepted(st, t.n, t.A);
            if (n_threads) {
                {
                    std::lock_guard<std::mutex> lk(pool.mu);
                    pool.tasks.push_back(t);
                }
                pool.cv_task.notify_one();
            } else {
                run_task(t);
            }
        }
        for (int32_t k = 0; k < n_cands; ++k) {  // np.random.randint(low, high), one per candidate
            uint32_t v = 0;
            if (rng != 0) {
                do {
                    v = st.next() & rmask;
                } while (v > (uint32_t)rng);
            }
            feat_counts[(size_t)s * n_cands + k] = (int32_t)(randint_low + (int64_t)v);
        }
    }
    {
        std::lock_guard<std::mutex> lk(pool.mu);
        pool.closing = true;
    }
    pool.cv_task.notify_all();
    for (auto& w : workers) w.join();
    if (ring) {
        ring->finish();
        memcpy(key, st.key, MT_N * sizeof(uint32_t));
    }
    *pos_io = st.pos;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// CPython's random.sample(population, k) (model.py:216, `random_sample(feature_cols, ...)`) for a
// whole plan of candidates: the INDICES it would pick, on CPython's own MT19937 stream
// (random.getstate()[1] = 624 state words + index; same generator and tempering as numpy's).
// Restates CPython 3.12 Lib/random.py:
//   _randbelow_with_getrandbits(n): k = n.bit_length(); r = getrandbits(k) until r < n
//   getrandbits(k <= 32)          : genrand_uint32() >> (32 - k)      (Modules/_randommodule.c)
//   sample(): setsize = 21 (+ 4 ** ceil(log(3k, 4)) when k > 5);
//             n <= setsize: pool selection  (j = randbelow(n - i); take pool[j]; pool[j] = pool[n-i-1])
//             else        : rejection on a set of already selected indices
// Parity: tests/test_host_rng.py compares with random.sample itself, end state included.
// ---------------------------------------------------------------------------------------------
#include <math.h>

namespace {
inline uint32_t py_randbelow(Stream& st, uint32_t n) {
    const int k = 32 - __builtin_clz(n);  // n.bit_length(), n >= 1
    uint32_t r;
    do {
        r = st.next() >> (32 - k);
    } while (r >= n);
    return r;
}
}  // namespace

extern "C" int rfmc_pyrandom_sample_plan(uint32_t* key, int32_t* pos_io, int32_t n_pop, int32_t n_calls, const int32_t* ks,
                                         int32_t* out) {
    if (!key || !pos_io || !ks || !out || n_pop < 1 || n_calls < 0 || *pos_io < 0 || *pos_io > MT_N) return 1;
    Stream st;
    st.attach(key, *pos_io);
    std::vector<int32_t> pool((size_t)n_pop);
    std::vector<uint8_t> taken((size_t)n_pop);
    int64_t at = 0;
    for (int32_t c = 0; c < n_calls; ++c) {
        const int32_t k = ks[c];
        if (k < 0 || k > n_pop) return 1;
        long long setsize = 21;
        if (k > 5) setsize += (long long)llround(pow(4.0, ceil(log((double)k * 3.0) / log(4.0))));
        if ((long long)n_pop <= setsize) {
            for (int32_t i = 0; i < n_pop; ++i) pool[i] = i;
            for (int32_t i = 0; i < k; ++i) {
                const uint32_t j = py_randbelow(st, (uint32_t)(n_pop - i));
                out[at++] = pool[j];
                pool[j] = pool[n_pop - i - 1];
            }
        } else {
            memset(taken.data(), 0, taken.size());
            for (int32_t i = 0; i < k; ++i) {
                uint32_t j = py_randbelow(st, (uint32_t)n_pop);
                while (taken[j]) j = py_randbelow(st, (uint32_t)n_pop);
                taken[j] = 1;
                out[at++] = (int32_t)j;
            }
        }
    }
    *pos_io = st.pos;
    return 0;
}
