// Device walk of numpy's global legacy MT19937 stream for split_train_val / sampleFeats
// (model.py:198-219): per tree slot one np.random.choice(class_rows, N, replace=False) per class,
// then one np.random.randint(min_feature, max_feature + 1) per candidate.  Same stream, same row
// ids, same feature counts and same end state as numpy (and as the host restatement in
// host_rng.cpp, which documents the numpy call chain); tests/test_gpu_rng.py compares all three.
//
//   choice(a, N, replace=False) = a[permutation(n)[:N]];  permutation = shuffle(arange(n))
//   shuffle: for i = n-1 .. 1: j_i = random_interval(i); swap(x[i], x[j_i])
//   random_interval(i): mask = smallest 2^k - 1 >= i; draw 32-bit words until (word & mask) <= i
//
// The kernels of one plan of slots, all on one CUDA stream:
//
//   mt_fill_kernel      the stream's tempered words W[0 .. L) into HBM.  One CTA per chunk of
//                       1680 blocks of 624 words; chunk c jumps straight to word c*J: the window
//                       x[cJ .. cJ+623] is a GF(2) correlation of the first 19936+624 words of the
//                       sequence with the polynomial t^(cJ-1) mod phi (mt_jump.cpp), then blocks are
//                       regenerated in three dependent phases of 227 / 227 / 170 lanes.
//   classify / chain / final   the masked rejection walk: which word feeds which Fisher-Yates step.  A
//                       word is accepted iff (w & mask) <= i, and i falls by one per accept: sequential
//                       only through the running count of accepts.  All per-word work runs GPU-wide in
//                       chunks of 4,096 words decided against a band around the EXPECTED count; one warp
//                       per stream settles the few undecided words exactly, in order; a last pass with
//                       exact counts at both ends of every chunk writes the draws A[k] = j_(n-1-k) and
//                       verifies every decision (see the section comment below): exact, never
//                       statistical.
//   resolve_scan_kernel one CTA per (slot, class): only the first N entries of the permutation are
//                       needed.  Steps i >= N touch prefix position p only through chains
//                       p -> (first i with j_i == p) -> ...; one ascending scan over the draws with the
//                       live chain heads in a shared-memory hash set resolves all N chains (tiles of
//                       up to 8,192 steps tested in parallel, winners by atomicMin, the rare hits on
//                       heads born inside the same tile applied in order).
//   resolve_prefix_kernel  steps i < N act inside the prefix (one warp per draw), then
//                       idx_train / idx_val = class_rows[head[p]] straight into the grower's inputs.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace rfmc {
namespace {

constexpr int MT_N = 624;
constexpr int JUMP_DEG = 19937;
constexpr int JUMP_SEQ = JUMP_DEG + MT_N;  // x[0 .. 20560]

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t m) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// ------------------------------------------------------------------------------------------------
// mt_fill_kernel
// ------------------------------------------------------------------------------------------------
constexpr int FILL_THREADS = 256;

__global__ void __launch_bounds__(FILL_THREADS) mt_fill_kernel(const uint32_t* __restrict__ key,
                                                               const uint32_t* __restrict__ polys, int blocks_per_chunk,
                                                               int64_t n_blocks, uint32_t* __restrict__ W) {
    extern __shared__ __align__(16) uint32_t sm[];  // JUMP_SEQ + 1 words; the fill reuses the first 2 * 624
    const int tid = threadIdx.x;
    const int64_t c = blockIdx.x;
    uint32_t win[3];
    if (c == 0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int t = tid + r * FILL_THREADS;
            win[r] = t < MT_N ? key[t] : 0u;
        }
    } else {
        uint32_t* seq = sm;
        for (int t = tid; t < MT_N; t += FILL_THREADS) seq[t] = key[t];
        __syncthreads();
        // x[m] = x[m-227] ^ mix(x[m-624], x[m-623]): 227 new words per step depend on older ones only
        for (int m0 = MT_N; m0 <= JUMP_SEQ; m0 += 227) {
            const int m = m0 + tid;
            if (tid < 227 && m <= JUMP_SEQ) seq[m] = mt_mix(seq[m - 624], seq[m - 623], seq[m - 227]);
            __syncthreads();
        }
        const uint32_t* g = polys + (size_t)(c - 1) * MT_N;
        win[0] = win[1] = win[2] = 0u;
        const int t0 = tid + 1, t1 = tid + 1 + FILL_THREADS;
        const int t2 = tid + 1 + 2 * FILL_THREADS <= MT_N ? tid + 1 + 2 * FILL_THREADS : 0;  // 0: idle third output
        for (int wd = 0; wd < MT_N; ++wd) {
            uint32_t gw = __ldg(&g[wd]);
            const uint32_t* base = seq + wd * 32;
            while (gw) {
                const int b = __ffs(gw) - 1;
                gw &= gw - 1;
                win[0] ^= base[b + t0];
                win[1] ^= base[b + t1];
                win[2] ^= base[b + t2];
            }
        }
        __syncthreads();
    }
    uint32_t* cur = sm;
    uint32_t* nxt = sm + MT_N;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int t = tid + r * FILL_THREADS;
        if (t < MT_N) cur[t] = win[r];
    }
    __syncthreads();
    const int64_t b0 = c * blocks_per_chunk;
    int64_t b1 = b0 + blocks_per_chunk;
    if (b1 > n_blocks) b1 = n_blocks;
    for (int64_t b = b0; b < b1; ++b) {
        uint32_t* out = W + b * MT_N;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int t = tid + r * FILL_THREADS;
            if (t < MT_N) out[t] = mt_temper(cur[t]);
        }
        if (b + 1 == b1) break;
        if (tid < 227) nxt[tid] = mt_mix(cur[tid], cur[tid + 1], cur[tid + 397]);
        __syncthreads();
        if (tid < 227) nxt[tid + 227] = mt_mix(cur[tid + 227], cur[tid + 228], nxt[tid]);
        __syncthreads();
        if (tid < 169) nxt[tid + 454] = mt_mix(cur[tid + 454], cur[tid + 455], nxt[tid + 227]);
        if (tid == 169) nxt[623] = mt_mix(cur[623], nxt[0], nxt[396]);
        __syncthreads();
        uint32_t* t = cur;
        cur = nxt;
        nxt = t;
    }
}

// ------------------------------------------------------------------------------------------------
// The rejection walk: which word of the stream feeds which draw.
//
// Every accepted word gets an ORDINAL k (0, 1, 2, ... over the whole plan); what an ordinal needs is
// fixed in advance (the schedule below: per slot the classes' shuffle steps i = n-1 .. 1, then the
// randint draws), so the walk is the recurrence  k(p+1) = k(p) + [word p is accepted at ordinal k(p)].
// It is sequential only through k.  The stream is cut into chunks of 4,096 words and worked in rounds:
//
//   classify_kernel  one CTA per chunk, a whole round of chunks at once.  The chain (below) predicts
//                    the ordinal of every chunk boundary from the schedule (expected acceptance rates)
//                    with a 7.5 sigma band.  A word whose verdict is the same for every ordinal of its
//                    band is CERTAIN (counted with ballots); the rest are PENDING and leave as
//                    (word, position, certain accepts before it).
//   chain_kernel     one warp per stream: walks the round's chunks in order and settles the pending
//                    words exactly (32 at a time, ballot fix-point over the exact ordinals) -> the exact
//                    ordinal at every chunk boundary; then predicts the next round's boundaries.
//   final_kernel     one CTA per chunk, all chunks at once: with the exact ordinal at both ends of its
//                    chunk it decides every word again against a tight band, settles its own pending
//                    words, checks EVERY word's exact ordinal against the band it was decided with (a
//                    miss redoes the chunk with all words pending), writes the accepted draws
//                    A[k] = word & mask and the stream position of every slot boundary, and compares its
//                    count with the chain's.  A chunk whose count differs from the chain's fails the
//                    plan loudly (status), so a band that was too narrow in classify_kernel can cost
//                    time but never a wrong draw.
// ------------------------------------------------------------------------------------------------
constexpr int CH_WORDS = 4096;   // words per chunk
constexpr int CH_THREADS = 512;  // 8 words per thread
constexpr int CH_ROWS = CH_WORDS / CH_THREADS;
constexpr int CH_SEGS = CH_WORDS / 32;
constexpr int SCHED_MAX_ENT = RFMC_MAX_CLASSES + 1;

struct Sched {
    int32_t n_ent;                   // draws-with-steps per slot: classes with n >= 2, then the randint block
    int32_t steps;                   // ordinals per slot
    int32_t ks[SCHED_MAX_ENT + 1];   // first ordinal of entity e inside a slot; ks[n_ent] = steps
    int32_t nn[SCHED_MAX_ENT];       // class size n of a shuffle entity; 0 = the randint block
    uint32_t rmask, rrange;
    int64_t k_total;                 // n_slots * steps
};

__device__ __forceinline__ int ent_of(const Sched& sc, int32_t kr, int e) {
    while (kr >= sc.ks[e + 1]) ++e;
    while (kr < sc.ks[e]) --e;
    return e;
}

__device__ __forceinline__ int32_t wrap_slot(const Sched& sc, int32_t kr) {
    while (kr >= sc.steps) kr -= sc.steps;
    while (kr < 0) kr += sc.steps;
    return kr;
}

// Is word w accepted at slot-relative ordinal kr?  u = the draw it would deliver.
__device__ __forceinline__ bool accept_at(const Sched& sc, uint32_t w, int32_t kr, int& e, uint32_t& u) {
    e = ent_of(sc, kr, e);
    const int32_t n = sc.nn[e];
    if (n == 0) {
        u = w & sc.rmask;
        return u <= sc.rrange;
    }
    const int32_t i = n - 1 - (kr - sc.ks[e]);  // >= 1
    const uint32_t M = 0xFFFFFFFFu >> __clz(i);
    u = w & M;
    return u <= (uint32_t)i;
}

// Verdict of word w over every ordinal of [klo, klo + span] (slot-relative, periodic in the slot):
// bit 0 = accepted at all of them, bit 1 = rejected at all of them.  Inside one mask level the test
// u <= i is monotone in the ordinal, so a level segment is settled by its two ends.
__device__ __forceinline__ int band_verdict(const Sched& sc, uint32_t w, int32_t klo, int32_t span, int e_hint) {
    int32_t k = wrap_slot(sc, klo);
    int e = ent_of(sc, k, e_hint);
    int32_t rem = span + 1;
    bool all_acc = true, all_rej = true;
    for (int it = 0; it < 40; ++it) {
        const int32_t n = sc.nn[e];
        int32_t cnt;
        if (n == 0) {
            cnt = sc.ks[e + 1] - k;
            const bool acc = (w & sc.rmask) <= sc.rrange;
            all_acc &= acc;
            all_rej &= !acc;
        } else {
            const int32_t i_hi = n - 1 - (k - sc.ks[e]);
            const uint32_t M = 0xFFFFFFFFu >> __clz(i_hi);
            cnt = i_hi - (int32_t)(M >> 1);  // ordinals left in this level
            cnt = cnt < rem ? cnt : rem;
            const uint32_t u = w & M;
            all_acc &= u <= (uint32_t)(i_hi - (cnt - 1));
            all_rej &= u > (uint32_t)i_hi;
        }
        cnt = cnt < rem ? cnt : rem;
        rem -= cnt;
        if (rem <= 0) return (all_acc ? 1 : 0) | (all_rej ? 2 : 0);
        if (!(all_acc || all_rej)) return 0;
        k += cnt;
        if (k >= sc.ks[e + 1]) {
            ++e;
            if (e == sc.n_ent) {
                e = 0;
                k = 0;
            }
        }
    }
    return 0;  // a band across more than 40 level segments: leave the word pending
}

struct ChunkHdr {
    int32_t n_sure, n_pend;
};

struct WalkBuffers {
    const uint32_t* W;     // tempered words; chunk c = W[pos0 + c * 4096 ...]
    int64_t n_words;
    int32_t pos0;
    int32_t chunks_per_round;
    int64_t n_chunks;      // chunks of the whole plan
    int64_t* k0;           // [n_chunks + 1] exact ordinal at every chunk boundary (chain_kernel)
    int64_t* kpred;        // [chunks_per_round + 1] predicted boundaries of the round being classified
    int32_t* eband;        // [chunks_per_round] half-width of the band
    ChunkHdr* hdr;         // [chunks_per_round]
    uint2* pend;           // [chunks_per_round][4096] {word, position | certain accepts before << 16}
    uint32_t* A;           // accepted draws by ordinal
    int64_t* slot_pos;     // [n_slots + 1]
    int32_t* status;       // [3] != 0: a chunk's count differed from the chain's; [1] chunks redone all-pending
};

// Decide the CTA's 4,096 words against bands [kr0 + dk(q) - E(q), .. + E(q)], dk linear from 0 to dk_end.
// Certain accepts are counted, pending words compacted in stream order.  Leaves per-thread ballots and
// the segments' exclusive prefixes in shared memory for the caller.
struct ChunkShared {
    uint32_t seg[CH_SEGS];
    uint32_t tot;
};

template <bool TIGHT>
__device__ __forceinline__ void classify_words(const Sched& sc, ChunkShared& sh, const uint32_t (&w)[CH_ROWS], int32_t kr0,
                                               int32_t dk_end, int32_t e_band, bool all_pending, int e_hint,
                                               uint32_t (&bs)[CH_ROWS], uint32_t (&ba)[CH_ROWS], int32_t (&klo)[CH_ROWS],
                                               int32_t (&khi)[CH_ROWS], uint32_t& level_mask) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The usual chunk: every band stays inside ONE mask level of ONE class's shuffle.  Then the step number is
    // i_base - ordinal offset and the verdict two compares; the schedule walk (band_verdict) is for the chunks
    // that touch a level, class or slot boundary.
    bool one_level = false;
    int32_t i_base = 0;
    level_mask = 0u;
    {
        const int32_t n = sc.nn[e_hint];
        const int32_t kmin = kr0 - (TIGHT ? 0 : e_band), kmax = kr0 + dk_end + (TIGHT ? 0 : e_band);
        if (!all_pending && n > 0 && dk_end >= 0 && kmin >= sc.ks[e_hint] && kmax < sc.ks[e_hint + 1]) {
            i_base = n - 1 - (kr0 - sc.ks[e_hint]);
            const int32_t i_top = i_base + (kr0 - kmin), i_bot = i_base - (kmax - kr0);  // >= 1: kmax is inside the class
            if (__clz(i_top) == __clz(i_bot)) {
                one_level = true;
                level_mask = 0xFFFFFFFFu >> __clz(i_top);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < CH_ROWS; ++v) {
        const int q = v * CH_THREADS + tid;
        const int32_t dk = (int32_t)(((int64_t)dk_end * q) >> 12);
        int32_t E = e_band;
        if (TIGHT) {  // both ends of the chunk are exact: the walk is a bridge, widest in the middle
            const int m = q < CH_WORDS - q ? q : CH_WORDS - q;
            E += (int32_t)(3.75f * __fsqrt_rn((float)m)) + 1;
        }
        int32_t lo_k = dk - E, hi_k = dk + E;
        if (TIGHT) {  // exact start: between none and q accepts before word q, at most dk_end in the whole chunk
            lo_k = lo_k < 0 ? 0 : lo_k;
            hi_k = hi_k > q ? q : hi_k;
            hi_k = hi_k > dk_end ? dk_end : hi_k;
            const int32_t floor_k = dk_end - (CH_WORDS - q);
            lo_k = lo_k < floor_k ? floor_k : lo_k;
        }
        klo[v] = lo_k;
        khi[v] = hi_k;
        int verdict = 0;
        if (!all_pending && hi_k >= lo_k) {
            if (one_level) {
                const uint32_t u = w[v] & level_mask;
                verdict = u <= (uint32_t)(i_base - hi_k) ? 1 : (u > (uint32_t)(i_base - lo_k) ? 2 : 0);
            } else {
                verdict = band_verdict(sc, w[v], kr0 + lo_k, hi_k - lo_k, e_hint);
            }
        }
        bs[v] = __ballot_sync(FULL, verdict == 1);
        ba[v] = __ballot_sync(FULL, verdict == 0);
        if (lane == 0) sh.seg[v * (CH_THREADS / 32) + warp] = (uint32_t)__popc(bs[v]) | ((uint32_t)__popc(ba[v]) << 16);
    }
    if (!one_level) level_mask = 0u;
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 128 segment counts (both 16-bit fields at once)
        constexpr int PER = CH_SEGS / 32;
        uint32_t x[PER], s = 0;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            x[e] = sh.seg[lane * PER + e];
            s += x[e];
        }
        uint32_t inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += t;
        }
        uint32_t run = inc - s;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            sh.seg[lane * PER + e] = run;
            run += x[e];
        }
        if (lane == 31) sh.tot = inc;
    }
    __syncthreads();
}

__device__ __forceinline__ void load_chunk(const WalkBuffers& b, int64_t chunk, uint32_t (&w)[CH_ROWS]) {
    const int64_t base = b.pos0 + chunk * CH_WORDS;
#pragma unroll
    for (int v = 0; v < CH_ROWS; ++v) {
        const int64_t p = base + v * CH_THREADS + threadIdx.x;
        w[v] = p < b.n_words ? __ldg(&b.W[p]) : 0u;
    }
}

__global__ void __launch_bounds__(CH_THREADS) classify_kernel(const __grid_constant__ Sched sc, WalkBuffers b, int64_t chunk0) {
    __shared__ ChunkShared sh;
    const int g = blockIdx.x;
    const int64_t chunk = chunk0 + g;
    if (chunk >= b.n_chunks) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    uint32_t w[CH_ROWS], bs[CH_ROWS], ba[CH_ROWS];
    int32_t klo[CH_ROWS], khi[CH_ROWS];
    load_chunk(b, chunk, w);
    const int64_t ks = b.kpred[g], ke = b.kpred[g + 1];
    const int32_t kr0 = (int32_t)(ks % sc.steps);
    const int e_hint = ent_of(sc, kr0, 0);
    uint32_t level_mask;
    classify_words<false>(sc, sh, w, kr0, (int32_t)(ke - ks), b.eband[g], false, e_hint, bs, ba, klo, khi, level_mask);
    const uint32_t tot = sh.tot;
    if (tid == 0) b.hdr[g] = ChunkHdr{(int32_t)(tot & 0xFFFFu), (int32_t)(tot >> 16)};
    uint2* out = b.pend + (size_t)g * CH_WORDS;
#pragma unroll
    for (int v = 0; v < CH_ROWS; ++v) {
        if ((ba[v] >> lane) & 1u) {
            const uint32_t pre = sh.seg[v * (CH_THREADS / 32) + warp];
            const uint32_t e = (pre >> 16) + (uint32_t)__popc(ba[v] & lt);
            const uint32_t sure_before = (pre & 0xFFFFu) + (uint32_t)__popc(bs[v] & lt);
            // first guess for the chain: the verdict at the predicted ordinal itself
            const int q = v * CH_THREADS + tid;
            int ent = e_hint;
            uint32_t u;
            const uint32_t guess = accept_at(sc, w[v], wrap_slot(sc, kr0 + (int32_t)(((int64_t)(ke - ks) * q) >> 12)), ent, u) ? 1u : 0u;
            out[e] = make_uint2(w[v], (uint32_t)q | (sure_before << 16) | (guess << 31));
        }
    }
}

// Settle a list of pending words in stream order, one warp, 32 at a time.  fetch(e) -> {word, certain
// accepts before it}; kr0 = slot-relative ordinal at the chunk's start.  Returns the accepts among them;
// res(e, accepted pending before e, own verdict) is told every word's outcome.
template <typename Fetch, typename Store>
__device__ __forceinline__ int settle_pending(const Sched& sc, int n_pend, int32_t kr0, int e_hint, Fetch fetch, Store store) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    int D = 0;
    uint2 nxt = lane < n_pend ? fetch(lane) : make_uint2(0u, 0u);
    for (int base = 0; base < n_pend; base += 32) {
        const int e = base + lane;
        const bool valid = e < n_pend;
        const uint2 cur = nxt;
        if (e + 32 < n_pend) nxt = fetch(e + 32);
        const int32_t k_base = kr0 + (int32_t)cur.y + D;  // ordinal if no pending word of this group before it is accepted
        int ent = e_hint;
        uint32_t u;
        uint32_t bits = __ballot_sync(FULL, valid && accept_at(sc, cur.x, wrap_slot(sc, k_base), ent, u));
        for (;;) {  // lane l is final once the lanes below it are: at most 32 rounds, usually 2
            const int c = __popc(bits & lt);
            const uint32_t nb = __ballot_sync(FULL, valid && accept_at(sc, cur.x, wrap_slot(sc, k_base + c), ent, u));
            if (nb == bits) break;
            bits = nb;
        }
        if (valid) store(e, D + __popc(bits & lt), (bits >> lane) & 1u);
        D += __popc(bits);
    }
    return D;
}

// Expected walk from slot-relative ordinal kr0 over `words` words (the words may run through several
// mask levels, classes and slots): the expected number of accepts, and the variance of that count.
// Single precision is enough: the band built from it only decides how many words stay pending.
__device__ void predict_span(const Sched& sc, int32_t kr0, float words, double& dk_out, float& var_out) {
    double kr = (double)kr0, dk = 0.0;
    float var = 0.0f;
    int e = ent_of(sc, kr0, 0);
    for (int guard = 0; guard < 4096 && words > 0.0f; ++guard) {
        if (kr >= (double)sc.steps) {
            kr -= (double)sc.steps;
            e = 0;
        }
        const int32_t ki = (int32_t)kr;
        e = ent_of(sc, ki, e);
        const int32_t n = sc.nn[e];
        if (n == 0) {
            const float p = ((float)sc.rrange + 1.0f) / ((float)sc.rmask + 1.0f);
            const float left = (float)((double)sc.ks[e + 1] - kr);
            const float need = left / p;
            const bool stays = words < need;
            const float use = stays ? words : need;
            const double adv = stays ? (double)(use * p) : (double)sc.ks[e + 1] - kr;
            kr += adv;
            dk += adv;
            var += use * p * (1.0f - p);
            words -= use;
            continue;
        }
        int32_t i_int = n - 1 - (ki - sc.ks[e]);
        i_int = i_int < 1 ? 1 : i_int;
        const uint32_t M = 0xFFFFFFFFu >> __clz(i_int);
        const float lo = (float)(M >> 1), Mp = (float)M + 1.0f;
        const double level_end = (double)sc.ks[e] + (double)(n - 1) - (double)(M >> 1);  // ordinal where i reaches lo
        const float i_f = (float)((double)(n - 1) - (kr - (double)sc.ks[e]));
        const float togo = (float)(level_end - kr);
        if (togo < 0.5f) {
            dk += level_end - kr;
            kr = level_end;
            continue;
        }
        const float r_here = fminf(1.0f, (i_f + 1.0f) / Mp), r_end = (lo + 1.0f) / Mp;
        const float need = Mp * __logf((i_f + 1.0f) / (lo + 1.0f));  // expected words to the end of this level
        if (words < need) {
            const float d = (i_f + 1.0f) * (1.0f - __expf(-words / Mp));
            const float r_mid = 0.5f * (r_here + fminf(1.0f, (i_f + 1.0f - d) / Mp));
            kr += (double)d;
            dk += (double)d;
            var += words * r_mid * (1.0f - r_mid);
            words = 0.0f;
        } else {
            const float r_mid = 0.5f * (r_here + r_end);
            dk += level_end - kr;
            kr = level_end;
            var += need * r_mid * (1.0f - r_mid);
            words -= need;
        }
    }
    dk_out = dk;
    var_out = var;
}

// Acceptance rate at slot-relative ordinal kr, and an id of its (entity, mask level).
__device__ __forceinline__ float rate_at(const Sched& sc, int32_t kr, int& level_id) {
    kr = wrap_slot(sc, kr);
    const int e = ent_of(sc, kr, 0);
    const int32_t n = sc.nn[e];
    if (n == 0) {
        level_id = e << 6;
        return ((float)sc.rrange + 1.0f) / ((float)sc.rmask + 1.0f);
    }
    const int32_t i = n - 1 - (kr - sc.ks[e]);
    const int lz = __clz(i);
    level_id = (e << 6) | lz;
    return ((float)i + 1.0f) / ((float)(0xFFFFFFFFu >> lz) + 1.0f);
}

// How far the expected walk of one chunk may leave the chord between its ends: inside one level the rate
// falls smoothly (words * dr / 8); across a level or class boundary it jumps (words * dr / 4).
__device__ __forceinline__ int32_t chord_slack(const Sched& sc, int32_t kr_a, int32_t kr_b) {
    int la, lb;
    const float ra = rate_at(sc, kr_a, la), rb = rate_at(sc, kr_b, lb);
    if (la == lb) return 4 + (int32_t)((float)CH_WORDS * 0.125f * fabsf(ra - rb));
    const float r_lo = fminf(0.5f, fminf(ra, rb));
    return 4 + (int32_t)((float)CH_WORDS * 0.25f * (1.0f - r_lo));
}

constexpr int CN_THREADS = 1024;  // one pending word per thread
constexpr int CN_ROUND_MAX = 256; // chunks per round the chain's tables hold

struct ChainShared {
    int32_t ent0[CN_ROUND_MAX + 1];   // pending words before chunk g of the round
    int32_t sure0[CN_ROUND_MAX + 1];  // certain accepts before chunk g of the round
    uint32_t bal[32];                 // accept ballots of the unit's 32 warps
    uint32_t pre[32];                 // exclusive prefix of their counts
    int32_t tot;
};

// The chain: the round's pending words in stream order, 1,024 at a time, one per thread.  Jacobi fix-point:
// every word is decided at the ordinal that the PREVIOUS iteration's verdicts put it at; the counts are
// re-derived from the new verdicts by a ballot scan; an iteration that changes no verdict is the exact walk
// (the recurrence is causal, so its fixed point is unique).  The first guess is the classifier's verdict at
// the predicted ordinal; 3-4 iterations of ~300 cycles settle 1,024 words.
__global__ void __launch_bounds__(CN_THREADS, 1) chain_kernel(const __grid_constant__ Sched sc, WalkBuffers b, int64_t chunk0,
                                                              int n_here, int n_next) {
    __shared__ ChainShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    int64_t k = chunk0 == 0 ? 0 : b.k0[chunk0];
    if (chunk0 == 0 && tid == 0) b.k0[0] = 0;
    // prefix sums of the round's chunk headers
    if (warp == 0) {
        int run_e = 0, run_s = 0;
        for (int g0 = 0; g0 < n_here; g0 += 32) {
            const int g = g0 + lane;
            const ChunkHdr h = g < n_here ? b.hdr[g] : ChunkHdr{0, 0};
            int ie = h.n_pend, is = h.n_sure;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int te = __shfl_up_sync(FULL, ie, d), ts = __shfl_up_sync(FULL, is, d);
                if (lane >= d) {
                    ie += te;
                    is += ts;
                }
            }
            if (g < n_here) {
                sh.ent0[g + 1] = run_e + ie;
                sh.sure0[g + 1] = run_s + is;
            }
            run_e += __shfl_sync(FULL, ie, 31);
            run_s += __shfl_sync(FULL, is, 31);
        }
        if (lane == 0) sh.ent0[0] = sh.sure0[0] = 0;
    }
    __syncthreads();
    const int total = n_here > 0 ? sh.ent0[n_here] : 0;
    const int32_t kr0 = (int32_t)(k % sc.steps);
    const int e_hint = ent_of(sc, kr0, 0);
    // the usual round lies inside one class's shuffle: the step number is then i0 - (accepts before the word)
    const bool one_class = n_here > 0 && sc.nn[e_hint] > 0 && kr0 + sh.sure0[n_here] + total < sc.ks[e_hint + 1];
    const int32_t i0 = sc.nn[e_hint] - 1 - (kr0 - sc.ks[e_hint]);
    // boundaries that have no pending word before them
    for (int g = tid; g < n_here; g += CN_THREADS)
        if (sh.ent0[g + 1] == 0) b.k0[chunk0 + g + 1] = k + sh.sure0[g + 1];

    auto fetch = [&](int E, uint2& x) {  // entry E of the round: {word, certain accepts before it in the round | guess << 31}
        x = make_uint2(0u, 0u);
        if (E >= total) return;
        int lo = 0, hi = n_here;  // chunk g with ent0[g] <= E < ent0[g + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sh.ent0[mid] <= E) lo = mid; else hi = mid;
        }
        const uint2 raw = __ldg(&b.pend[(size_t)lo * CH_WORDS + (E - sh.ent0[lo])]);
        x = make_uint2(raw.x, (uint32_t)(sh.sure0[lo] + (int)((raw.y >> 16) & 0x7FFFu)) | (raw.y & 0x80000000u));
    };

    int D_unit = 0;  // accepted pending words before the current unit
    uint2 nxt;
    fetch(tid, nxt);
    for (int u0 = 0; u0 < total; u0 += CN_THREADS) {
        const uint2 cur = nxt;
        fetch(u0 + CN_THREADS + tid, nxt);  // in flight while this unit iterates
        const int E = u0 + tid;
        const bool valid = E < total;
        const int32_t before = (int32_t)(cur.y & 0x7FFFFFFFu);
        uint32_t bits = __ballot_sync(FULL, valid && (cur.y >> 31));
        for (;;) {
            if (lane == 0) {
                sh.bal[warp] = bits;
                sh.pre[warp] = (uint32_t)__popc(bits);
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t c = sh.pre[lane];
                uint32_t inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += t;
                }
                sh.pre[lane] = inc - c;
                if (lane == 31) sh.tot = (int)inc;
            }
            __syncthreads();
            const int32_t D = D_unit + (int32_t)sh.pre[warp] + __popc(bits & lt);
            bool acc = false;
            if (valid) {
                if (one_class) {
                    const int32_t i = i0 - before - D;
                    acc = (cur.x & (0xFFFFFFFFu >> __clz(i))) <= (uint32_t)i;
                } else {
                    int en = e_hint;
                    uint32_t uu;
                    acc = accept_at(sc, cur.x, wrap_slot(sc, kr0 + before + D), en, uu);
                }
            }
            const uint32_t nb = __ballot_sync(FULL, acc);
            const bool changed = nb != bits;
            bits = nb;
            if (!__syncthreads_or(changed)) break;
        }
        // sh.bal / sh.pre / sh.tot describe the converged verdicts: exact ordinals at the chunk boundaries that
        // fall inside this unit (boundary B = pending words before chunk g + 1)
        const int u1 = u0 + CN_THREADS < total ? u0 + CN_THREADS : total;
        for (int g = tid; g < n_here; g += CN_THREADS) {
            const int B = sh.ent0[g + 1];
            if (B > u0 && B <= u1) {
                const int r = B - u0;  // entries of the unit before the boundary
                const int D = r >= CN_THREADS ? sh.tot : (int)sh.pre[r >> 5] + __popc(sh.bal[r >> 5] & ((1u << (r & 31)) - 1u));
                b.k0[chunk0 + g + 1] = k + sh.sure0[g + 1] + D_unit + D;
            }
        }
        D_unit += sh.tot;
        __syncthreads();
    }
    if (n_here > 0) k += sh.sure0[n_here] + D_unit;
    // boundaries of the next round, from the exact ordinal reached: one thread per chunk
    if (n_next > 0) {
        const int32_t kr1 = (int32_t)(k % sc.steps);
        if (tid == 0) b.kpred[0] = k;
        float var = 0.0f;
        if (tid < n_next) {
            double dk;
            predict_span(sc, kr1, (float)(tid + 1) * (float)CH_WORDS, dk, var);
            b.kpred[tid + 1] = k + (int64_t)(dk + 0.5);
        }
        __syncthreads();
        if (tid < n_next) {
            const int32_t ka = (int32_t)((b.kpred[tid] - k + kr1) % sc.steps), kb = (int32_t)((b.kpred[tid + 1] - k + kr1) % sc.steps);
            b.eband[tid] = 8 + (int32_t)(7.5f * sqrtf(var + 1.0f)) + chord_slack(sc, ka, kb);
        }
    }
    if (tid == 0) atomicAdd(&b.status[4], total >> 10);  // thousands of pending words the chain settled
}

struct FinalShared {
    ChunkShared cs;
    uint32_t pend_w[CH_WORDS];
    uint16_t pend_s[CH_WORDS];
    uint16_t pend_res[CH_WORDS];  // accepted pending words before it | own verdict << 15
    int32_t pend_acc;
};

__global__ void __launch_bounds__(CH_THREADS, 2) final_kernel(const __grid_constant__ Sched sc, WalkBuffers b) {
    __shared__ FinalShared sh;
    const int64_t chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const int64_t ka = b.k0[chunk], kb = b.k0[chunk + 1];
    if (ka >= sc.k_total) return;  // the plan ended before this chunk
    const int64_t slot_a = ka / sc.steps;
    const int32_t kr0 = (int32_t)(ka - slot_a * sc.steps);
    const int32_t dk_end = (int32_t)(kb - ka);
    const int e_hint = ent_of(sc, kr0, 0);
    const int64_t next_slot_k = (slot_a + 1) * sc.steps;  // first slot boundary at or after this chunk's start
    uint32_t w[CH_ROWS], bs[CH_ROWS], ba[CH_ROWS];
    int32_t klo[CH_ROWS], khi[CH_ROWS];
    load_chunk(b, chunk, w);
    // how far the expected walk may bend away from the chord between the exact ends
    const int32_t e_band = chord_slack(sc, kr0, kr0 + (dk_end > 0 && dk_end <= CH_WORDS ? dk_end : 0));
    const bool sane = dk_end >= 0 && dk_end <= CH_WORDS;
    bool counted = false;
    for (int attempt = sane ? 0 : 1; attempt < 2 && !counted; ++attempt) {
        const bool all_pending = attempt == 1;
        uint32_t level_mask;
        classify_words<true>(sc, sh.cs, w, kr0, sane ? dk_end : 0, e_band, all_pending, e_hint, bs, ba, klo, khi, level_mask);
        const uint32_t tot = sh.cs.tot;
        const int n_pend = (int)(tot >> 16), n_sure = (int)(tot & 0xFFFFu);
#pragma unroll
        for (int v = 0; v < CH_ROWS; ++v) {
            if ((ba[v] >> lane) & 1u) {
                const uint32_t pre = sh.cs.seg[v * (CH_THREADS / 32) + warp];
                const uint32_t e = (pre >> 16) + (uint32_t)__popc(ba[v] & lt);
                sh.pend_w[e] = w[v];
                sh.pend_s[e] = (uint16_t)((pre & 0xFFFFu) + (uint32_t)__popc(bs[v] & lt));
            }
        }
        __syncthreads();
        if (warp == 0) {
            const int D = settle_pending(
                sc, n_pend, kr0, e_hint, [&](int e) { return make_uint2(sh.pend_w[e], (uint32_t)sh.pend_s[e]); },
                [&](int e, int before, uint32_t acc) { sh.pend_res[e] = (uint16_t)((uint32_t)before | (acc << 15)); });
            if (lane == 0) sh.pend_acc = D;
        }
        __syncthreads();
        const int pend_acc = sh.pend_acc;
        bool viol = false;
#pragma unroll
        for (int v = 0; v < CH_ROWS; ++v) {
            const uint32_t pre = sh.cs.seg[v * (CH_THREADS / 32) + warp];
            const int a_idx = (int)(pre >> 16) + __popc(ba[v] & lt);
            const uint32_t res = a_idx < n_pend ? sh.pend_res[a_idx] : (uint32_t)pend_acc;
            const int k = (int)(pre & 0xFFFFu) + __popc(bs[v] & lt) + (int)(res & 0x7FFFu);  // exact accepts before this word
            if (!all_pending && (k < klo[v] || k > khi[v])) viol = true;
        }
        if (__syncthreads_or(viol)) {
            if (tid == 0) atomicAdd(&b.status[1], 1);
            continue;  // a word left its band: decide the chunk again with every word pending (exact by itself)
        }
        counted = true;
        if (n_sure + pend_acc != dk_end) {
            if (tid == 0) atomicExch(&b.status[3], 1);  // the chain counted this chunk differently: the plan is void
            return;
        }
#pragma unroll
        for (int v = 0; v < CH_ROWS; ++v) {
            const uint32_t pre = sh.cs.seg[v * (CH_THREADS / 32) + warp];
            const int a_idx = (int)(pre >> 16) + __popc(ba[v] & lt);
            const uint32_t res = a_idx < n_pend ? sh.pend_res[a_idx] : (uint32_t)pend_acc;
            const int k = (int)(pre & 0xFFFFu) + __popc(bs[v] & lt) + (int)(res & 0x7FFFu);
            const bool is_pend = (ba[v] >> lane) & 1u;
            const bool acc = is_pend ? ((res >> 15) & 1u) : ((bs[v] >> lane) & 1u);
            const int64_t kg = ka + k;
            if (acc && kg < sc.k_total) {
                uint32_t u;
                if (level_mask) {
                    u = w[v] & level_mask;
                } else {
                    int ent = e_hint;
                    accept_at(sc, w[v], wrap_slot(sc, kr0 + k), ent, u);
                }
                b.A[kg] = u;
                const int64_t after = kg + 1;  // a slot ends with its last draw
                if (after >= next_slot_k && (after - next_slot_k) % sc.steps == 0)
                    b.slot_pos[after / sc.steps] = b.pos0 + chunk * CH_WORDS + v * CH_THREADS + tid + 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// resolve_scan_kernel
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 512;
constexpr int RS_PER = 16;  // steps per thread in the widest tile (8,192 steps)
constexpr int RS_LATE_MAX = 1024;
constexpr int RS_CAND_MAX = 2048;
constexpr uint32_t H_EMPTY = 0xFFFFFFFFu, H_TOMB = 0xFFFFFFFEu;

struct ResolveArgs {
    const uint32_t* A;
    const int32_t* class_n;
    const int64_t* class_aoff;
    int32_t n_classes;
    int64_t slot_stride;  // draws per slot in A (all classes' shuffle steps + the randint block)
    int32_t size;      // entries of the permutation wanted per class
    int32_t cap_log2;  // hash capacity
    uint32_t* heads;   // [n_slots * n_classes][size]
    int32_t* status;   // status[2] != 0: late list overflow
};

struct Hash {
    uint32_t* keys;
    uint32_t* win;
    uint16_t* own;
    uint32_t mask;
    int shift;
    bool full = false;
    __device__ __forceinline__ uint32_t slot_of(uint32_t key) const { return (key * 0x9E3779B1u) >> shift; }
    __device__ __forceinline__ int find(uint32_t key) const {
        uint32_t s = slot_of(key);
        for (uint32_t probes = 0; probes <= mask; ++probes) {
            const uint32_t k = keys[s];
            if (k == key) return (int)s;
            if (k == H_EMPTY) return -1;
            s = (s + 1) & mask;
        }
        return -1;
    }
    // returns 1 when a never-used slot was taken (the table got fuller)
    __device__ __forceinline__ int insert(uint32_t key, uint16_t owner) {
        uint32_t s = slot_of(key);
        for (uint32_t probes = 0; probes <= 2 * mask + 1; ++probes) {
            const uint32_t k = keys[s];
            if (k == H_EMPTY || k == H_TOMB) {
                if (atomicCAS(&keys[s], k, key) == k) {
                    own[s] = owner;
                    return k == H_EMPTY;
                }
                continue;  // somebody else took it: look at the slot again
            }
            s = (s + 1) & mask;
        }
        full = true;  // cannot happen while the table is rebuilt at half load; reported, never silent
        return 0;
    }
};

__global__ void __launch_bounds__(RS_THREADS) resolve_scan_kernel(ResolveArgs a) {
    extern __shared__ __align__(16) unsigned char rs_dyn[];
    __shared__ uint32_t s_late[RS_LATE_MAX];   // q of steps whose j lies inside the tile
    __shared__ uint32_t s_found[RS_LATE_MAX];  // late steps whose j is a live head right now
    __shared__ uint32_t s_cand_q[RS_CAND_MAX];  // steps that found their j among the live heads, and the slot
    __shared__ uint32_t s_cand_s[RS_CAND_MAX];
    __shared__ int s_n_late, s_n_found, s_used, s_n_cand;
    const int tid = threadIdx.x;
    const int draw = blockIdx.x;
    const int c = draw % a.n_classes, s = draw / a.n_classes;
    const int32_t n = a.class_n[c];
    const int32_t size = a.size;
    const uint32_t cap = 1u << a.cap_log2;
    Hash h;
    h.keys = reinterpret_cast<uint32_t*>(rs_dyn);
    h.win = h.keys + cap;
    uint32_t* head = h.win + cap;
    h.own = reinterpret_cast<uint16_t*>(head + size);
    h.mask = cap - 1;
    h.shift = 32 - a.cap_log2;
    const uint32_t* __restrict__ A = a.A + (size_t)s * a.slot_stride + a.class_aoff[c];
    uint32_t* heads_out = a.heads + (size_t)draw * size;

    for (uint32_t t = tid; t < cap; t += RS_THREADS) {
        h.keys[t] = H_EMPTY;
        h.win[t] = 0xFFFFFFFFu;
    }
    if (tid == 0) s_used = 0;
    __syncthreads();
    int fresh = 0;
    for (int p = tid; p < size; p += RS_THREADS) {
        head[p] = (uint32_t)p;
        fresh += h.insert((uint32_t)p, (uint16_t)p);
    }
    if (fresh) atomicAdd(&s_used, fresh);
    __syncthreads();

    int32_t q0 = size;
    while (q0 < n) {
        int T = (q0 / 8) & ~(RS_THREADS - 1);
        T = T < RS_THREADS ? RS_THREADS : (T > RS_THREADS * RS_PER ? RS_THREADS * RS_PER : T);
        if (T > n - q0) T = n - q0;
        const int per = (T + RS_THREADS - 1) / RS_THREADS;
        // a table that is mostly tombstones makes every miss a long probe: rebuild it from the live heads
        if ((uint32_t)s_used * 2u > cap) {
            __syncthreads();
            uint32_t* tmp_k = h.win;  // win is all-ones between tiles: borrow it
            uint32_t* tmp_o = h.win + cap / 2;
            for (int p = tid; p < size; p += RS_THREADS) {
                tmp_k[p] = head[p];
                tmp_o[p] = (uint32_t)p;
            }
            for (uint32_t t = tid; t < cap; t += RS_THREADS) h.keys[t] = H_EMPTY;
            __syncthreads();
            for (int p = tid; p < size; p += RS_THREADS) h.insert(tmp_k[p], (uint16_t)tmp_o[p]);
            __syncthreads();
            for (uint32_t t = tid; t < cap; t += RS_THREADS) h.win[t] = 0xFFFFFFFFu;
            if (tid == 0) s_used = size;
            __syncthreads();
        }
        if (tid == 0) {
            s_n_late = 0;
            s_n_found = 0;
            s_n_cand = 0;
        }
        __syncthreads();
        // phase 1: every step of the tile looks its j up among the live heads
        for (int r = 0; r < per; ++r) {
            const int32_t q = q0 + r * RS_THREADS + tid;
            if (q < q0 + T) {
                const uint32_t j = __ldg(&A[n - 1 - q]);
                if (j >= (uint32_t)q0) {
                    if (j < (uint32_t)q) {
                        const int at = atomicAdd(&s_n_late, 1);
                        if (at < RS_LATE_MAX) s_late[at] = (uint32_t)q;
                    }
                } else {
                    const int sl = h.find(j);
                    if (sl >= 0) {
                        atomicMin(&h.win[sl], (uint32_t)q);  // several steps may hit one head: the first one kills it
                        const int at = atomicAdd(&s_n_cand, 1);
                        if (at < RS_CAND_MAX) {
                            s_cand_q[at] = (uint32_t)q;
                            s_cand_s[at] = (uint32_t)sl;
                        }
                    }
                }
            }
        }
        __syncthreads();
        int n_cand = s_n_cand;
        if (n_cand > RS_CAND_MAX) {
            if (tid == 0) a.status[2] = 3;
            n_cand = RS_CAND_MAX;
        }
        // phase 2a: the winner of every hit head
        for (int e = tid; e < n_cand; e += RS_THREADS)
            if (h.win[s_cand_s[e]] == s_cand_q[e]) s_cand_s[e] |= 0x80000000u;
        __syncthreads();
        // phase 2b: winners move their chain's head; every touched slot gives its ticket back
        for (int e = tid; e < n_cand; e += RS_THREADS) {
            const uint32_t sv = s_cand_s[e];
            const uint32_t sl = sv & 0x7FFFFFFFu;
            h.win[sl] = 0xFFFFFFFFu;
            if (sv >> 31) {
                const uint16_t o = h.own[sl];
                h.keys[sl] = H_TOMB;
                head[o] = s_cand_q[e];
                s_cand_s[e] = 0x80000000u | o;
            }
        }
        __syncthreads();
        fresh = 0;
        for (int e = tid; e < n_cand; e += RS_THREADS) {
            const uint32_t sv = s_cand_s[e];
            if (sv >> 31) fresh += h.insert(s_cand_q[e], (uint16_t)(sv & 0xFFFFu));
        }
        if (fresh) atomicAdd(&s_used, fresh);
        __syncthreads();
        // Steps whose j is a position of this very tile: j is live only if step j was itself a hit.  Rounds:
        // every pending late step looks j up in parallel; the (few) found ones are applied in step order;
        // a hit creates a head that a later round may find.  Different rounds touch different heads.
        int n_late = s_n_late;
        if (n_late > RS_LATE_MAX) {
            if (tid == 0) a.status[2] = 1;
            n_late = RS_LATE_MAX;
        }
        while (n_late > 0) {
            for (int e = tid; e < n_late; e += RS_THREADS) {
                const uint32_t q = s_late[e];
                if (q != 0xFFFFFFFFu && h.find(__ldg(&A[n - 1 - (int32_t)q])) >= 0) {
                    s_found[atomicAdd(&s_n_found, 1)] = q;
                    s_late[e] = 0xFFFFFFFFu;
                }
            }
            __syncthreads();
            const int n_found = s_n_found;
            __syncthreads();
            if (n_found == 0) break;
            if (tid == 0) {
                for (int x = 0; x < n_found; ++x) {  // ascending q
                    int best = x;
                    for (int y = x + 1; y < n_found; ++y)
                        if (s_found[y] < s_found[best]) best = y;
                    const uint32_t q = s_found[best];
                    s_found[best] = s_found[x];
                    const int sl = h.find(__ldg(&A[n - 1 - (int32_t)q]));
                    if (sl >= 0) {
                        const uint16_t o = h.own[sl];
                        h.keys[sl] = H_TOMB;
                        head[o] = q;
                        s_used += h.insert(q, o);
                    }
                }
                s_n_found = 0;
            }
            __syncthreads();
        }
        q0 += T;
    }
    __syncthreads();
    if (h.full) a.status[2] = 2;
    for (int p = tid; p < size; p += RS_THREADS) heads_out[p] = head[p];
}

// ------------------------------------------------------------------------------------------------
// resolve_prefix_kernel: steps i < size swap inside the prefix; then the row ids.
// ------------------------------------------------------------------------------------------------
struct PrefixArgs {
    const uint32_t* A;
    const int32_t* class_n;
    const int64_t* class_aoff;
    const int32_t* class_rows;  // every class's ascending row ids back to back
    const int64_t* class_roff;  // [n_classes + 1]
    int32_t n_classes, n_draws;
    int64_t slot_stride;
    int32_t size, n_tr, n_va;
    const uint32_t* heads;
    int32_t* idx_train;  // [n_slots][n_classes * n_tr]
    int32_t* idx_val;    // [n_slots][n_classes * n_va]
};

__global__ void resolve_prefix_kernel(PrefixArgs a) {
    extern __shared__ __align__(16) uint32_t pf_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int draw = blockIdx.x * (blockDim.x >> 5) + warp;
    if (draw >= a.n_draws) return;
    const int c = draw % a.n_classes, s = draw / a.n_classes;
    const int32_t n = a.class_n[c];
    const int32_t size = a.size;
    uint32_t* head = pf_dyn + (size_t)warp * 2 * size;
    uint32_t* js = head + size;
    const uint32_t* __restrict__ A = a.A + (size_t)s * a.slot_stride + a.class_aoff[c];
    const uint32_t* hin = a.heads + (size_t)draw * size;
    for (int p = lane; p < size; p += 32) {
        head[p] = hin[p];
        js[p] = p >= 1 ? __ldg(&A[n - 1 - p]) : 0u;  // j of step q = p (q = 0 is no step)
    }
    __syncwarp();
    if (lane == 0) {
        for (int q = size - 1; q >= 1; --q) {
            const uint32_t j = js[q];
            const uint32_t t = head[q];
            head[q] = head[j];
            head[j] = t;
        }
    }
    __syncwarp();
    const int32_t* rows = a.class_rows + a.class_roff[c];
    for (int p = lane; p < size; p += 32) {
        const int32_t row = __ldg(&rows[head[p]]);
        if (p < a.n_tr)
            a.idx_train[((size_t)s * a.n_classes + c) * a.n_tr + p] = row;
        else
            a.idx_val[((size_t)s * a.n_classes + c) * a.n_va + (p - a.n_tr)] = row;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side of the plan
// ------------------------------------------------------------------------------------------------
constexpr int RNG_BLOCKS_PER_CHUNK = 1680;  // 1,048,320 words per fill CTA

// Expected number of stream words one shuffle of n elements consumes: sum over steps i of (mask_i + 1) / (i + 1).
static double expected_words(int64_t n) {
    double total = 0.0;
    if (n < 2) return 0.0;
    uint64_t M = (uint64_t)(n - 1);
    M |= M >> 1, M |= M >> 2, M |= M >> 4, M |= M >> 8, M |= M >> 16, M |= M >> 32;
    int64_t hi = n - 1;
    while (hi >= 1) {
        const int64_t lo = (int64_t)(M >> 1);  // this mask serves i in (lo, hi]
        double h = 0.0;
        if (hi - lo <= 64) {
            for (int64_t i = lo + 1; i <= hi; ++i) h += 1.0 / (double)(i + 1);
        } else {
            auto H = [](double x) { return log(x) + 0.5772156649015329 + 1.0 / (2.0 * x) - 1.0 / (12.0 * x * x); };
            h = H((double)(hi + 1)) - H((double)(lo + 1));
        }
        total += (double)(M + 1) * h;
        hi = lo;
        M >>= 1;
    }
    return total;
}

static void build_schedule(const rfmc_drawplan* pl, Sched& sc) {
    const rfmc_dataset* ds = pl->ds;
    sc.n_ent = 0;
    int32_t k = 0;
    for (int c = 0; c < pl->n_classes; ++c) {
        const int64_t n = ds->h_class_n[c];
        if (n < 2) continue;
        sc.ks[sc.n_ent] = k;
        sc.nn[sc.n_ent] = (int32_t)n;
        k += (int32_t)(n - 1);
        ++sc.n_ent;
    }
    if (pl->n_cands > 0 && pl->rrange != 0u) {
        sc.ks[sc.n_ent] = k;
        sc.nn[sc.n_ent] = 0;
        k += pl->n_cands;
        ++sc.n_ent;
    }
    sc.ks[sc.n_ent] = k;
    sc.steps = k;
    sc.rmask = pl->rmask;
    sc.rrange = pl->rrange;
    sc.k_total = (int64_t)pl->n_slots * k;
}

__global__ void init_slot_pos_kernel(int64_t* slot_pos, int n, int64_t first, int64_t rest) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_pos[i] = i == 0 ? first : rest;
}

int drawplan_launch(rfmc_drawplan* pl) {
    rfmc_dataset* ds = pl->ds;
    cudaStream_t st = pl->stream;
    const int C = pl->n_classes;
    auto mark = [&](int i) {
        if (pl->timing) cudaEventRecord(pl->ev[i], st);
    };
    Sched sc;
    build_schedule(pl, sc);
    mark(0);
    int launches = 0;
    {
        const size_t smem = (size_t)(JUMP_SEQ + 1) * sizeof(uint32_t);
        RFMC_CUDA(cudaFuncSetAttribute(mt_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mt_fill_kernel<<<(unsigned)pl->n_chunks, FILL_THREADS, smem, st>>>(pl->d_key, pl->d_polys, pl->blocks_per_chunk, pl->n_blocks,
                                                                           pl->d_W);
        RFMC_CUDA(cudaGetLastError());
        ++launches;
    }
    mark(1);
    // a slot boundary that is never reached stays -1: the host then knows the words ran out
    init_slot_pos_kernel<<<(pl->n_slots + 1 + 255) / 256, 256, 0, st>>>(pl->d_slot_pos, pl->n_slots + 1, (int64_t)pl->pos0,
                                                                     sc.steps > 0 ? (int64_t)-1 : (int64_t)pl->pos0);
    RFMC_CUDA(cudaGetLastError());
    ++launches;
    if (sc.steps > 0) {
        WalkBuffers b;
        b.W = pl->d_W;
        b.n_words = pl->n_blocks * MT_N;
        b.pos0 = pl->pos0;
        b.chunks_per_round = pl->chunks_per_round;
        b.n_chunks = pl->n_walk_chunks;
        b.k0 = pl->d_k0;
        b.kpred = pl->d_kpred;
        b.eband = pl->d_eband;
        b.hdr = (ChunkHdr*)pl->d_hdr;
        b.pend = (uint2*)pl->d_pend;
        b.A = pl->d_A;
        b.slot_pos = pl->d_slot_pos;
        b.status = pl->d_status;
        const int G = pl->chunks_per_round;
        const int64_t N = pl->n_walk_chunks;
        chain_kernel<<<1, CN_THREADS, 0, st>>>(sc, b, 0, 0, (int)(N < G ? N : G));
        ++launches;
        for (int64_t c0 = 0; c0 < N; c0 += G) {
            const int here = (int)(N - c0 < G ? N - c0 : G);
            const int64_t left = N - c0 - here;
            classify_kernel<<<here, CH_THREADS, 0, st>>>(sc, b, c0);
            chain_kernel<<<1, CN_THREADS, 0, st>>>(sc, b, c0, here, (int)(left < G ? left : G));
            launches += 2;
        }
        RFMC_CUDA(cudaGetLastError());
        mark(2);
        final_kernel<<<(unsigned)N, CH_THREADS, 0, st>>>(sc, b);
        RFMC_CUDA(cudaGetLastError());
        ++launches;
    } else {
        mark(2);
    }
    mark(3);
    if (pl->size > 0) {
        ResolveArgs r;
        r.A = pl->d_A;
        r.class_n = ds->d_class_n;
        r.class_aoff = ds->d_class_aoff;
        r.n_classes = C;
        r.slot_stride = sc.steps;
        r.size = (int32_t)pl->size;
        r.cap_log2 = pl->cap_log2;
        r.heads = pl->d_heads;
        r.status = pl->d_status;
        const size_t cap = (size_t)1 << pl->cap_log2;
        const size_t smem = cap * 8 + (size_t)pl->size * 4 + cap * 2;
        RFMC_CUDA(cudaFuncSetAttribute(resolve_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resolve_scan_kernel<<<(unsigned)(pl->n_slots * C), RS_THREADS, smem, st>>>(r);
        RFMC_CUDA(cudaGetLastError());
        PrefixArgs x;
        x.A = pl->d_A;
        x.class_n = ds->d_class_n;
        x.class_aoff = ds->d_class_aoff;
        x.class_rows = ds->d_class_rows;
        x.class_roff = ds->d_class_roff;
        x.n_classes = C;
        x.n_draws = pl->n_slots * C;
        x.slot_stride = sc.steps;
        x.size = (int32_t)pl->size;
        x.n_tr = (int32_t)pl->n_tr;
        x.n_va = (int32_t)pl->n_va;
        x.heads = pl->d_heads;
        x.idx_train = pl->d_idx_train;
        x.idx_val = pl->d_idx_val;
        int wpb = 4;
        while (wpb > 1 && (size_t)wpb * 2 * pl->size * 4 > 160 * 1024) wpb >>= 1;
        const size_t psm = (size_t)wpb * 2 * pl->size * 4;
        RFMC_CUDA(cudaFuncSetAttribute(resolve_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        resolve_prefix_kernel<<<(unsigned)((x.n_draws + wpb - 1) / wpb), wpb * 32, psm, st>>>(x);
        RFMC_CUDA(cudaGetLastError());
        launches += 2;
    }
    mark(4);
    pl->launches = launches;
    pl->steps_per_slot = sc.steps;
    return 0;
}

int64_t drawplan_words_needed(const std::vector<int64_t>& class_n, int n_slots, int n_cands, uint32_t rmask, uint32_t rrange,
                              int attempt) {
    double per_slot = 0.0, steps = 0.0;
    for (int64_t n : class_n) {
        per_slot += expected_words(n);
        steps += (double)n;
    }
    if (rrange) per_slot += (double)n_cands * ((double)rmask + 1.0) / ((double)rrange + 1.0);
    const double mean = per_slot * n_slots;
    const double sigma = sqrt(2.0 * (steps + n_cands) * n_slots);
    return (int64_t)(mean + (8.0 + 8.0 * attempt) * sigma) + 16384 + (int64_t)attempt * (1 << 20);
}

int drawplan_geometry(rfmc_drawplan* pl, int64_t words) {
    const int64_t walk_chunks = (words + CH_WORDS - 1) / CH_WORDS + 1;
    pl->n_walk_chunks = walk_chunks;
    pl->chunks_per_round = 128;
    if (const char* e = getenv("RFMC_DRAW_ROUND_CHUNKS")) {  // experiment knob: chunks classified per round (<= 256)
        const int want = atoi(e);
        if (want >= 8 && want <= 256) pl->chunks_per_round = want;
    }
    pl->n_blocks = (pl->pos0 + walk_chunks * CH_WORDS + MT_N - 1) / MT_N + 1;
    // fill CTAs: about two per SM; the chunk length is one of five sizes so that the jump polynomials are shared
    int bpc = RNG_BLOCKS_PER_CHUNK;
    while (bpc > 105 && pl->n_blocks / bpc < 296) bpc /= 2;
    pl->blocks_per_chunk = pl->n_blocks < bpc ? (int)pl->n_blocks : bpc;
    pl->n_chunks = (pl->n_blocks + pl->blocks_per_chunk - 1) / pl->blocks_per_chunk;
    return 0;
}

int64_t drawplan_chunk_words(const rfmc_drawplan* pl) { return (int64_t)pl->blocks_per_chunk * MT_N; }

size_t resolve_smem_bytes(int64_t size, int cap_log2) {
    const size_t cap = (size_t)1 << cap_log2;
    return cap * 10 + (size_t)size * 4 + 3 * 1024 + RS_LATE_MAX * 8 + RS_CAND_MAX * 8;
}

}  // namespace rfmc
