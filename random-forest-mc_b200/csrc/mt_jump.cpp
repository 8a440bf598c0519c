// MT19937 jump-ahead polynomials over GF(2) -- host arithmetic behind the device walk of numpy's
// legacy stream (split_train_val / sampleFeats, model.py:198-219).
//
// The device generates one stream's words in parallel: chunk c starts from the 624-word window
// x[cJ .. cJ+623] of the untempered sequence.  MT19937 is a linear recurrence over GF(2) whose state
// (top bit of x[k], x[k+1..k+623]) has the primitive characteristic polynomial phi of degree 19937,
// so with g(t) = t^(Q-1) mod phi(t):
//
//         x[Q-1+t] = XOR over the set bits i of g of x[i+t]            for every t >= 1
//
// (every later word is a linear function of the state).  The window at Q is t = 1..624 of that sum
// over the first 19936 + 624 words after the known window -- a correlation the GPU evaluates
// (csrc/rng_walk.cu, jump_kernel).  This file only produces g: polynomial products with carry-less
// multiplies and reduction by phi, which has 135 terms (exponents below; derived from the generator's
// own output by Berlekamp-Massey, tools/mt19937_charpoly.py, and checked by tests/test_mt_jump_cpu.py
// against plain sequential generation).
#include <stdint.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define RFMC_HAVE_PCLMUL 1
#endif

#include "../../include/rfmc_b200.h"

namespace {

constexpr int DEG = 19937;
constexpr int PW = 312;  // 64-bit words of a reduced polynomial (19968 bits)

const int PHI_LOW[] = {  // exponents of phi below 19937
    0,     1189,  1416,  1585,  1643,  1870,  2493,  2773,  3000,  3227,  3454,  3681,  3908,  4135,  4362,  4753,  5661,
    6337,  6569,  7129,  7477,  7525,  7583,  7752,  7979,  8206,  9505,  9901,  9969,  10128, 10693, 10761, 10920, 11089,
    11147, 11157, 11215, 11321, 11374, 11384, 11485, 11611, 11712, 11717, 11838, 11881, 11944, 11997, 12277, 12335, 12393,
    12504, 12509, 12620, 12673, 12731, 12736, 12789, 12905, 12958, 12963, 13137, 13185, 13190, 13243, 13301, 13412, 13528,
    13533, 13639, 13697, 13760, 13813, 13866, 14093, 14151, 14209, 14320, 14325, 14436, 14547, 14552, 14605, 14721, 14774,
    14779, 14953, 15001, 15006, 15059, 15117, 15228, 15344, 15349, 15455, 15513, 15576, 15629, 15682, 15909, 15967, 16025,
    16136, 16141, 16252, 16363, 16368, 16421, 16537, 16590, 16595, 16817, 16822, 16875, 16933, 17044, 17160, 17271, 17329,
    17445, 17498, 17725, 17783, 17841, 17952, 18068, 18179, 18237, 18406, 18633, 18691, 18860, 19087, 19314};
constexpr int N_LOW = sizeof(PHI_LOW) / sizeof(PHI_LOW[0]);

struct Poly {
    uint64_t w[PW];
};

inline void xor_at(uint64_t* r, uint64_t h, int64_t bit) {  // r ^= h << bit
    const int64_t q = bit >> 6;
    const int s = (int)(bit & 63);
    r[q] ^= h << s;
    if (s) r[q + 1] ^= h >> (64 - s);
}

// r: 2*PW words, reduced in place to degree < DEG (the next term of phi lies 623 below the leading
// one, so a 64-bit slab folds strictly into lower words and one top-down pass is enough).
void reduce(uint64_t* r) {
    for (int w = 2 * PW - 1; w >= PW; --w) {
        const uint64_t h = r[w];
        if (!h) continue;
        r[w] = 0;
        const int64_t base = (int64_t)64 * w - DEG;
        for (int t = 0; t < N_LOW; ++t) xor_at(r, h, base + PHI_LOW[t]);
    }
    constexpr int TOP = DEG & 63;  // 33: bits >= 33 of word 311 are >= t^19937
    const uint64_t h = r[PW - 1] >> TOP;
    if (h) {
        r[PW - 1] &= (1ull << TOP) - 1;
        for (int t = 0; t < N_LOW; ++t) xor_at(r, h, PHI_LOW[t]);
    }
}

void mul_portable(const Poly& a, const Poly& b, uint64_t* r) {
    memset(r, 0, sizeof(uint64_t) * 2 * PW);
    for (int i = 0; i < PW; ++i) {
        uint64_t aw = a.w[i];
        while (aw) {
            const int bit = __builtin_ctzll(aw);
            aw &= aw - 1;
            const int64_t at = (int64_t)64 * i + bit;
            const int64_t q = at >> 6;
            const int s = (int)(at & 63);
            if (s == 0) {
                for (int j = 0; j < PW; ++j) r[q + j] ^= b.w[j];
            } else {
                uint64_t carry = 0;
                for (int j = 0; j < PW; ++j) {
                    r[q + j] ^= (b.w[j] << s) | carry;
                    carry = b.w[j] >> (64 - s);
                }
                r[q + PW] ^= carry;
            }
        }
    }
}

#ifdef RFMC_HAVE_PCLMUL
__attribute__((target("pclmul,sse2"))) void mul_clmul(const Poly& a, const Poly& b, uint64_t* r) {
    memset(r, 0, sizeof(uint64_t) * 2 * PW);
    for (int i = 0; i < PW; ++i) {
        if (!a.w[i]) continue;
        const __m128i va = _mm_set_epi64x(0, (long long)a.w[i]);
        for (int j = 0; j < PW; ++j) {
            const __m128i p = _mm_clmulepi64_si128(va, _mm_set_epi64x(0, (long long)b.w[j]), 0);
            __m128i cur = _mm_loadu_si128((const __m128i*)(r + i + j));
            _mm_storeu_si128((__m128i*)(r + i + j), _mm_xor_si128(cur, p));
        }
    }
}
#endif

Poly mulmod(const Poly& a, const Poly& b) {
    std::vector<uint64_t> r(2 * PW + 2);
#ifdef RFMC_HAVE_PCLMUL
    static const bool clmul = __builtin_cpu_supports("pclmul");
    if (clmul)
        mul_clmul(a, b, r.data());
    else
#endif
        mul_portable(a, b, r.data());
    reduce(r.data());
    Poly out;
    memcpy(out.w, r.data(), sizeof(out.w));
    return out;
}

Poly monomial(int64_t e) {  // t^e, e < DEG
    Poly p;
    memset(p.w, 0, sizeof(p.w));
    p.w[e >> 6] = 1ull << (e & 63);
    return p;
}

std::mutex g_mu;
std::vector<Poly> g_sq;            // g_sq[k] = t^(2^k) mod phi
std::map<int64_t, Poly> g_cache;   // t^e mod phi

const Poly& square_k(int k) {  // caller holds g_mu
    if (g_sq.empty()) g_sq.push_back(monomial(1));
    while ((int)g_sq.size() <= k) g_sq.push_back(mulmod(g_sq.back(), g_sq.back()));
    return g_sq[(size_t)k];
}

Poly xpow(int64_t e) {  // caller holds g_mu
    auto it = g_cache.find(e);
    if (it != g_cache.end()) return it->second;
    Poly acc;
    bool have = false;
    if (e < DEG) {
        acc = monomial(e);
        have = true;
    } else {
        // reuse the closest cached smaller exponent: t^e = t^e0 * t^(e - e0)
        int64_t rest = e;
        auto lo = g_cache.upper_bound(e);
        if (lo != g_cache.begin()) {
            --lo;
            if (lo->first > 0 && __builtin_popcountll((unsigned long long)(e - lo->first)) < __builtin_popcountll((unsigned long long)e)) {
                acc = lo->second;
                have = true;
                rest = e - lo->first;
            }
        }
        for (int k = 0; rest; ++k, rest >>= 1) {
            if (!(rest & 1)) continue;
            const Poly& s = square_k(k);
            acc = have ? mulmod(acc, s) : s;
            have = true;
        }
    }
    if (g_cache.size() > 4096) g_cache.clear();
    g_cache.emplace(e, acc);
    return acc;
}

// plain MT19937 regeneration (numpy/random/src/mt19937/mt19937.c: mt19937_gen), for the host check
void mt_next_words(std::vector<uint32_t>& x, size_t upto) {  // extend x (>= 624 words) to `upto` words
    while (x.size() < upto) {
        const size_t m = x.size();
        const uint32_t y = (x[m - 624] & 0x80000000u) | (x[m - 623] & 0x7fffffffu);
        x.push_back(x[m - 227] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu));
    }
}

}  // namespace

// t^(offset-1) mod phi as 624 32-bit words (bit i = coefficient of t^i), offset >= 1.
extern "C" int rfmc_mt_jump_poly(int64_t offset, uint32_t* out) {
    if (!out || offset < 1) return 1;
    Poly p;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        p = xpow(offset - 1);
    }
    for (int i = 0; i < PW; ++i) {
        out[2 * i] = (uint32_t)p.w[i];
        out[2 * i + 1] = (uint32_t)(p.w[i] >> 32);
    }
    return 0;
}

// Host evaluation of the jump (the checker of the device kernel and of the polynomials): the window
// x[offset .. offset+623] of the untempered sequence whose window at 0 is `key`.
extern "C" int rfmc_mt_jump_window_host(const uint32_t* key, int64_t offset, uint32_t* out) {
    if (!key || !out || offset < 0) return 1;
    if (offset == 0) {
        memcpy(out, key, 624 * sizeof(uint32_t));
        return 0;
    }
    uint32_t g[624];
    if (rfmc_mt_jump_poly(offset, g)) return 1;
    std::vector<uint32_t> x(key, key + 624);
    mt_next_words(x, (size_t)DEG + 624 + 1);
    memset(out, 0, 624 * sizeof(uint32_t));
    for (int i = 0; i < DEG; ++i) {
        if (!((g[i >> 5] >> (i & 31)) & 1u)) continue;
        const uint32_t* src = x.data() + i + 1;
        for (int t = 0; t < 624; ++t) out[t] ^= src[t];
    }
    return 0;
}
