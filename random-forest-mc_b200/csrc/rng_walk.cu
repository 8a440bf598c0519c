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
// Four kernels per plan of slots, all on one CUDA stream:
//
//   mt_fill_kernel      the stream's tempered words W[0 .. L) into HBM.  One CTA per chunk of
//                       1680 blocks of 624 words; chunk c jumps straight to word c*J: the window
//                       x[cJ .. cJ+623] is a GF(2) correlation of the first 19936+624 words of the
//                       sequence with the polynomial t^(cJ-1) mod phi (mt_jump.cpp), then blocks are
//                       regenerated in three dependent phases of 227 / 227 / 170 lanes.
//   walk_kernel         the masked rejection walk: which word feeds which Fisher-Yates step.  This is
//                       the sequential chain of the stream (a word is accepted iff (w & mask) <= i, and
//                       i falls by one per accept), executed by ONE CTA in chunks of up to 8,192 words:
//                       every thread decides its words against the EXPECTED number of earlier accepts
//                       +- a 7.5 sigma band; words outside the band are certain, the few inside
//                       (~0.2 % at the top level) are resolved exactly by one warp with a ballot
//                       fix-point over the exact prefix counts; every certain decision is then verified
//                       against the exact count (a violated band reruns the chunk on the exact warp
//                       path), so the result is exact, never statistical.  Accepted draws are
//                       written compacted: A[k] = j_(n-1-k).
//   resolve_scan_kernel one CTA per (slot, class): only the first N entries of the permutation are
//                       needed.  Steps i >= N touch prefix position p only through chains
//                       p -> (first i with j_i == p) -> ...; one ascending scan over the draws with the
//                       live chain heads in a shared-memory hash set resolves all N chains (tiles of
//                       up to 8,192 steps tested in parallel, winners by atomicMin, the rare hits on
//                       heads born inside the same tile applied in order).
//   resolve_prefix_kernel  steps i < N act inside the prefix (one warp per draw), then
//                       idx_train / idx_val = class_rows[head[p]] straight into the grower's inputs.
#include <cmath>

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
// walk_kernel
// ------------------------------------------------------------------------------------------------
constexpr int WK_THREADS = 1024;
constexpr int WK_ROWS = 8;        // words per thread of the widest chunk
constexpr int WK_AMB_MAX = 3072;  // undecided words one chunk may hold before it falls to the warp path

struct WalkArgs {
    const uint32_t* W;
    int64_t n_words;
    int32_t pos0;
    int32_t n_slots, n_classes, n_cands;
    const int32_t* class_n;     // [n_classes]
    const int64_t* class_aoff;  // [n_classes + 1] offsets of each class's draws inside a slot's segment of A
    uint32_t rmask, rrange;     // randint: accept iff (w & rmask) <= rrange; rrange == 0: no draw
    uint32_t* A;
    int32_t* randv;     // [n_slots][n_cands] accepted randint offsets
    int64_t* slot_pos;  // [n_slots + 1] stream position before each slot; [n_slots] = after the plan
    int32_t* status;    // 0 ok, 1 words exhausted; status[1] = chunks that fell back to the warp path
};

struct WalkShared {
    uint32_t seg[32 * WK_ROWS];  // per 32-word segment: certain accepts | undecided << 16, then their exclusive prefix
    uint32_t amb_u[WK_AMB_MAX];
    uint16_t amb_s[WK_AMB_MAX];    // certain accepts before the undecided word
    uint16_t amb_res[WK_AMB_MAX];  // undecided accepts before it | its own decision << 15
    uint32_t tot, amb_acc;
    int32_t cut;
    // state hand-off of the warp path
    int64_t p;
    int32_t i;
    int32_t done;
};

// Exact warp path: 32 words per round.  thr falls by `dec` per accept (1: shuffle step, 0: randint).
// Stops when `need` draws are accepted, `max_words` are consumed or the words run out.
__device__ void warp_walk(const uint32_t* __restrict__ W, int64_t n_words, int64_t& p, uint32_t M, int32_t& thr, int dec,
                          int32_t& need, uint32_t* __restrict__ out, int64_t max_words, bool& exhausted) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    while (need > 0 && max_words > 0) {
        if (p + 32 > n_words) {
            exhausted = true;
            return;
        }
        const uint32_t u = __ldg(&W[p + lane]) & M;
        uint32_t b = __ballot_sync(FULL, (int32_t)u <= thr);
        if (dec) {
            for (;;) {  // lane l is exact once lanes < l are: at most 32 rounds, usually 2
                const int c = __popc(b & lt);
                const uint32_t nb = __ballot_sync(FULL, (int32_t)u <= thr - c);
                if (nb == b) break;
                b = nb;
            }
        }
        int consumed = 32;
        if (max_words < 32) {
            consumed = (int)max_words;
            b &= (1u << consumed) - 1u;
        }
        int n_acc = __popc(b);
        if (n_acc >= need) {
            const int cutl = (int)__fns(b, 0, need);  // lane of the need-th accept
            b &= (2u << cutl) - 1u;
            consumed = cutl + 1;
            n_acc = need;
        }
        if ((b >> lane) & 1u) out[__popc(b & lt)] = u;
        out += n_acc;
        p += consumed;
        max_words -= consumed;
        thr -= dec * n_acc;
        need -= n_acc;
    }
}

// One chunk of V * 1024 words inside one mask level.  Returns false when the chunk has to be redone
// on the warp path (band violated or too many undecided words); then nothing may be committed.
template <int V>
__device__ __forceinline__ bool chunk_step(WalkShared& sh, const uint32_t (&w)[WK_ROWS], uint32_t M, int32_t i, int32_t need,
                                           uint32_t* __restrict__ out, int& accepted, int& consumed) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const float invM = 1.0f / ((float)M + 1.0f);
    const float ip1 = (float)i + 1.0f;
    float re = (ip1 - (float)(V * WK_THREADS)) * invM;  // acceptance probability at the chunk's end (>= 1/2 inside a level)
    re = fminf(fmaxf(re, 0.5f), 1.0f);
    const float vr = re * (1.0f - re);
    uint32_t u[V], bs[V], ba[V];
    int klo[V], khi[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int q = v * WK_THREADS + tid;
        u[v] = w[v] & M;
        const float kp = ip1 * (-expm1f(-(float)q * invM));  // expected accepts before word q
        const float E = 6.0f + 7.5f * sqrtf((float)q * vr + 1.0f);
        int lo_k = (int)floorf(kp - E), hi_k = (int)ceilf(kp + E);
        lo_k = lo_k < 0 ? 0 : lo_k;
        hi_k = hi_k > q ? q : hi_k;
        klo[v] = lo_k;
        khi[v] = hi_k;
        const bool sure_acc = (int32_t)u[v] <= i - hi_k;
        const bool sure_rej = (int32_t)u[v] > i - lo_k;
        bs[v] = __ballot_sync(FULL, sure_acc);
        ba[v] = __ballot_sync(FULL, !sure_acc && !sure_rej);
        if (lane == 0) sh.seg[v * 32 + warp] = (uint32_t)__popc(bs[v]) | ((uint32_t)__popc(ba[v]) << 16);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 32 * V segment counts (both 16-bit fields at once)
        uint32_t x[V], s = 0;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            x[e] = sh.seg[lane * V + e];
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
        for (int e = 0; e < V; ++e) {
            sh.seg[lane * V + e] = run;
            run += x[e];
        }
        if (lane == 31) sh.tot = inc;
    }
    __syncthreads();
    const uint32_t tot = sh.tot;
    const int n_amb = (int)(tot >> 16), n_sure = (int)(tot & 0xFFFFu);
    const bool overflow = n_amb > WK_AMB_MAX;
    if (!overflow) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if ((ba[v] >> lane) & 1u) {
                const uint32_t pre = sh.seg[v * 32 + warp];
                const int e = (int)(pre >> 16) + __popc(ba[v] & lt);
                sh.amb_u[e] = u[v];
                sh.amb_s[e] = (uint16_t)((pre & 0xFFFFu) + (uint32_t)__popc(bs[v] & lt));
            }
        }
    }
    __syncthreads();
    if (warp == 0 && !overflow) {  // the undecided words, in stream order, against the exact counts
        int D = 0;
        for (int base = 0; base < n_amb; base += 32) {
            const int e = base + lane;
            const bool valid = e < n_amb;
            const int32_t uu = valid ? (int32_t)sh.amb_u[e] : 0;
            const int32_t ss = valid ? (int32_t)sh.amb_s[e] + D : 0;
            uint32_t b = __ballot_sync(FULL, valid && uu <= i - ss);
            for (;;) {
                const int c = __popc(b & lt);
                const uint32_t nb = __ballot_sync(FULL, valid && uu <= i - ss - c);
                if (nb == b) break;
                b = nb;
            }
            if (valid) sh.amb_res[e] = (uint16_t)((uint32_t)(D + __popc(b & lt)) | (((b >> lane) & 1u) << 15));
            D += __popc(b);
        }
        if (lane == 0) sh.amb_acc = (uint32_t)D;
    }
    __syncthreads();
    bool viol = overflow;
    const int amb_acc = (int)sh.amb_acc;
    if (!overflow) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t pre = sh.seg[v * 32 + warp];
            const int a_idx = (int)(pre >> 16) + __popc(ba[v] & lt);
            const uint32_t res = a_idx < n_amb ? sh.amb_res[a_idx] : (uint32_t)amb_acc;
            const int k = (int)(pre & 0xFFFFu) + __popc(bs[v] & lt) + (int)(res & 0x7FFFu);  // exact accepts before this word
            const bool is_amb = (ba[v] >> lane) & 1u;
            const bool acc = is_amb ? ((res >> 15) & 1u) : ((bs[v] >> lane) & 1u);
            if (k < need) {
                if (k < klo[v] || k > khi[v]) viol = true;
                if (acc) {
                    out[k] = u[v];
                    if (k == need - 1) sh.cut = v * WK_THREADS + tid;
                }
            }
        }
    }
    if (__syncthreads_or(viol)) return false;
    const int total = n_sure + amb_acc;
    if (total >= need) {
        accepted = need;
        consumed = sh.cut + 1;
    } else {
        accepted = total;
        consumed = V * WK_THREADS;
    }
    return true;
}

__global__ void __launch_bounds__(WK_THREADS, 1) walk_kernel(WalkArgs a) {
    __shared__ WalkShared sh;
    const int tid = threadIdx.x;
    const uint32_t* __restrict__ W = a.W;
    int64_t p = a.pos0;
    int fallbacks = 0;
    // one chunk of prefetched words: pf[v] = W[pf_p + v * 1024 + tid]
    uint32_t pf[WK_ROWS];
    int64_t pf_p = -1;
    auto prefetch = [&](int64_t at) {
        pf_p = at;
#pragma unroll
        for (int v = 0; v < WK_ROWS; ++v) {
            const int64_t q = at + v * WK_THREADS + tid;
            pf[v] = q < a.n_words ? __ldg(&W[q]) : 0u;
        }
    };
    const int64_t slot_stride = a.class_aoff[a.n_classes];
    for (int s = 0; s < a.n_slots; ++s) {
        if (tid == 0) a.slot_pos[s] = p;
        for (int c = 0; c < a.n_classes; ++c) {
            const int32_t n = a.class_n[c];
            if (n < 2) continue;
            uint32_t* out = a.A + (size_t)s * slot_stride + a.class_aoff[c];
            uint32_t M = (uint32_t)(n - 1);
            M |= M >> 1;
            M |= M >> 2;
            M |= M >> 4;
            M |= M >> 8;
            M |= M >> 16;
            int32_t i = n - 1;
            while (i >= 1) {
                const int32_t lo = (int32_t)(M >> 1);
                while (i > lo) {
                    const int need = i - lo;
                    const int V = M >= 32767u ? 8 : (M >= 16383u ? 4 : (M >= 8191u ? 2 : (M >= 1023u ? 1 : 0)));
                    bool done = false;
                    if (V > 0 && p + V * WK_THREADS <= a.n_words) {
                        uint32_t w[WK_ROWS];
                        if (pf_p != p) prefetch(p);
#pragma unroll
                        for (int v = 0; v < WK_ROWS; ++v) w[v] = pf[v];
                        prefetch(p + V * WK_THREADS);
                        int accepted = 0, consumed = 0;
                        bool ok;
                        if (V == 8)
                            ok = chunk_step<8>(sh, w, M, i, need, out, accepted, consumed);
                        else if (V == 4)
                            ok = chunk_step<4>(sh, w, M, i, need, out, accepted, consumed);
                        else if (V == 2)
                            ok = chunk_step<2>(sh, w, M, i, need, out, accepted, consumed);
                        else
                            ok = chunk_step<1>(sh, w, M, i, need, out, accepted, consumed);
                        if (ok) {
                            p += consumed;
                            i -= accepted;
                            out += accepted;
                            done = true;
                        } else {
                            ++fallbacks;
                        }
                    }
                    if (!done) {
                        // exact warp path: the low levels, the tail of the words, and chunks whose band failed
                        if (tid < 32) {
                            int64_t pp = p;
                            int32_t thr = i, nd = need;
                            bool ex = false;
                            warp_walk(W, a.n_words, pp, M, thr, 1, nd, out, V > 0 ? (int64_t)V * WK_THREADS : (int64_t)1 << 40, ex);
                            if (tid == 0) {
                                sh.p = pp;
                                sh.i = thr;
                                sh.done = ex ? 1 : 0;
                            }
                        }
                        __syncthreads();
                        const int64_t np_ = sh.p;
                        const int32_t ni = sh.i;
                        const int ex = sh.done;
                        __syncthreads();
                        out += i - ni;
                        p = np_;
                        i = ni;
                        if (ex) {
                            if (tid == 0) a.status[0] = 1;
                            return;
                        }
                    }
                }
                M >>= 1;
            }
        }
        if (a.n_cands > 0) {
            if (a.rrange == 0u) {
                for (int k = tid; k < a.n_cands; k += WK_THREADS) a.randv[(size_t)s * a.n_cands + k] = 0;
            } else {
                if (tid < 32) {
                    int64_t pp = p;
                    int32_t thr = (int32_t)a.rrange, nd = a.n_cands;
                    bool ex = false;
                    warp_walk(W, a.n_words, pp, a.rmask, thr, 0, nd, (uint32_t*)(a.randv + (size_t)s * a.n_cands), (int64_t)1 << 40, ex);
                    if (tid == 0) {
                        sh.p = pp;
                        sh.done = ex ? 1 : 0;
                    }
                }
                __syncthreads();
                p = sh.p;
                const int ex = sh.done;
                __syncthreads();
                if (ex) {
                    if (tid == 0) a.status[0] = 1;
                    return;
                }
            }
        }
    }
    if (tid == 0) {
        a.slot_pos[a.n_slots] = p;
        a.status[1] = fallbacks;
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
    const uint32_t* __restrict__ A = a.A + (size_t)s * a.class_aoff[a.n_classes] + a.class_aoff[c];
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
    const uint32_t* __restrict__ A = a.A + (size_t)s * a.class_aoff[a.n_classes] + a.class_aoff[c];
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

int drawplan_launch(rfmc_drawplan* pl) {
    rfmc_dataset* ds = pl->ds;
    cudaStream_t st = pl->stream;
    const int C = pl->n_classes;
    auto mark = [&](int i) {
        if (pl->timing) cudaEventRecord(pl->ev[i], st);
    };
    mark(0);
    // fill
    {
        const size_t smem = (size_t)(JUMP_SEQ + 1) * sizeof(uint32_t);
        RFMC_CUDA(cudaFuncSetAttribute(mt_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mt_fill_kernel<<<(unsigned)pl->n_chunks, FILL_THREADS, smem, st>>>(pl->d_key, pl->d_polys, pl->blocks_per_chunk, pl->n_blocks,
                                                                           pl->d_W);
        RFMC_CUDA(cudaGetLastError());
    }
    mark(1);
    // walk
    {
        WalkArgs w;
        w.W = pl->d_W;
        w.n_words = pl->n_blocks * MT_N;
        w.pos0 = pl->pos0;
        w.n_slots = pl->n_slots;
        w.n_classes = C;
        w.n_cands = pl->n_cands;
        w.class_n = ds->d_class_n;
        w.class_aoff = ds->d_class_aoff;
        w.rmask = pl->rmask;
        w.rrange = pl->rrange;
        w.A = pl->d_A;
        w.randv = pl->d_randv;
        w.slot_pos = pl->d_slot_pos;
        w.status = pl->d_status;
        walk_kernel<<<1, WK_THREADS, 0, st>>>(w);
        RFMC_CUDA(cudaGetLastError());
    }
    mark(2);
    if (pl->size > 0) {
        ResolveArgs r;
        r.A = pl->d_A;
        r.class_n = ds->d_class_n;
        r.class_aoff = ds->d_class_aoff;
        r.n_classes = C;
        r.size = (int32_t)pl->size;
        r.cap_log2 = pl->cap_log2;
        r.heads = pl->d_heads;
        r.status = pl->d_status;
        const size_t cap = (size_t)1 << pl->cap_log2;
        const size_t smem = cap * 8 + (size_t)pl->size * 4 + cap * 2;
        RFMC_CUDA(cudaFuncSetAttribute(resolve_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resolve_scan_kernel<<<(unsigned)(pl->n_slots * C), RS_THREADS, smem, st>>>(r);
        RFMC_CUDA(cudaGetLastError());
        mark(3);
        PrefixArgs x;
        x.A = pl->d_A;
        x.class_n = ds->d_class_n;
        x.class_aoff = ds->d_class_aoff;
        x.class_rows = ds->d_class_rows;
        x.class_roff = ds->d_class_roff;
        x.n_classes = C;
        x.n_draws = pl->n_slots * C;
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
    }
    else
        mark(3);
    mark(4);
    pl->launches = pl->size > 0 ? 4 : 2;
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
    // the walk reads whole chunks of 8,192 words ahead of what it consumes
    return (int64_t)(mean + (8.0 + 8.0 * attempt) * sigma) + 65536 + (int64_t)attempt * (1 << 20);
}

int drawplan_geometry(rfmc_drawplan* pl, int64_t words) {
    pl->n_blocks = (pl->pos0 + words + MT_N - 1) / MT_N + 1;
    pl->blocks_per_chunk = pl->n_blocks < RNG_BLOCKS_PER_CHUNK ? (int)pl->n_blocks : RNG_BLOCKS_PER_CHUNK;
    pl->n_chunks = (pl->n_blocks + pl->blocks_per_chunk - 1) / pl->blocks_per_chunk;
    return 0;
}

int64_t drawplan_chunk_words(const rfmc_drawplan* pl) { return (int64_t)pl->blocks_per_chunk * MT_N; }

size_t resolve_smem_bytes(int64_t size, int cap_log2) {
    const size_t cap = (size_t)1 << cap_log2;
    return cap * 10 + (size_t)size * 4 + 3 * 1024 + RS_LATE_MAX * 8 + RS_CAND_MAX * 8;
}

}  // namespace rfmc
