// Host-side, bit-exact restatement of the draw behind split_train_val (model.py:198-210):
//
//     np.random.choice(class_indices, N, replace=False)
//
// on numpy's GLOBAL LEGACY generator.  The algorithm lives in numpy (a dependency of the
// reference, not vendored in it; numpy 2.3.5 here: numpy/random/mtrand.pyx `choice` ->
// `permutation` -> `shuffle` -> `_shuffle_raw`, and numpy/random/src/distributions/
// distributions.c `random_interval`, numpy/random/src/mt19937/mt19937.c):
//
//     idx = permutation(n)[:N]        permutation(n) = shuffle(arange(n))
//     shuffle: for i = n-1 .. 1:  j = random_interval(i);  swap(x[i], x[j])
//     random_interval(max): mask = smallest 2^k-1 >= max; draw 32-bit MT19937 words until
//                           (word & mask) <= max            (max <= 0xffffffff here)
//
// This is host code (the draw is an inherently sequential walk over the MT19937 stream); it exists
// because numpy's generic-itemsize memcpy shuffle costs ~19 ms per tree slot at 1M rows and is the
// Amdahl floor of fit() once the grower runs on the GPU.  State hand-off is
// np.random.get_state() / set_state(): key[624] + pos, so the stream continues exactly where
// numpy would have left it.  Parity: tests/test_host_rng.py compares indices AND end state with
// numpy itself.
#include <stdint.h>

#include <vector>

#include "../../include/rfmc_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

struct MT {
    uint32_t* key;
    int pos;
};

inline void mt_regen(MT& s) {
    uint32_t* k = s.key;
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
    s.pos = 0;
}

inline uint32_t mt_next(MT& s) {
    if (s.pos == MT_N) mt_regen(s);
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

thread_local std::vector<int32_t> g_perm;

}  // namespace

extern "C" int rfmc_legacy_permutation_prefix(uint32_t* key, int32_t* pos, int64_t n, int64_t size, int64_t* out) {
    if (!key || !pos || !out || n < 0 || size < 0 || size > n || n > 0x7fffffffll || *pos < 0 || *pos > MT_N) return 1;
    MT s{key, *pos};
    if ((int64_t)g_perm.size() < n) g_perm.resize((size_t)n);
    int32_t* x = g_perm.data();
    for (int64_t i = 0; i < n; ++i) x[i] = (int32_t)i;
    uint32_t mask = 0;
    if (n > 1) {
        mask = (uint32_t)(n - 1);
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
    }
    for (uint32_t i = (uint32_t)(n > 0 ? n - 1 : 0); i >= 1; --i) {
        while ((mask >> 1) >= i) mask >>= 1;  // smallest 2^k - 1 >= i
        uint32_t j;
        do {
            j = mt_next(s) & mask;
        } while (j > i);
        int32_t t = x[j];
        x[j] = x[i];
        x[i] = t;
    }
    for (int64_t k = 0; k < size; ++k) out[k] = x[k];
    *pos = s.pos;
    return 0;
}
