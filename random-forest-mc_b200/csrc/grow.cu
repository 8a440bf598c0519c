// Candidate-tree grower + hold-out scorer for sm_100a.
//
// Replaces plantTree/growTree/splitData/genLeaf (model.py:222-293) and validationTree
// (model.py:296-314) of the reference for a whole batch of candidates in one launch.
//
// Mapping: one persistent CTA (256 threads = 8 warps) per candidate; the candidate's tree is grown
// breadth-first.  Within a level every warp takes nodes from a shared ticket counter and does, for
// its node, exactly what one growTree() call does in the reference:
//   class histogram -> purity / depth test -> for each feature of the rotated list:
//   order statistics of the node's codes (median) or their mode -> split value in the column's
//   dtype arithmetic -> code threshold -> side sizes -> accept or rotate -> stable partition.
// Node numbering is breadth-first and deterministic (a block-wide scan assigns child ids after
// every level), so the result is comparable array-for-array with oracle/code_model.py.
//
// Memory: the candidate's bootstrap rows are gathered ONCE from the row-major HBM code matrix
// (one coalesced row read per warp) into a compact per-candidate [feature][row] matrix that stays
// L2-resident; all later accesses index that compact matrix by local row position.
//
// On-chip: per-warp shared-memory histograms (warp-aggregated atomics via match.any), ballots and
// shuffles for rank counting, prefix sums and arg-max.
#include "common.cuh"

namespace rfmc {

constexpr int GROW_THREADS = 256;
constexpr int GROW_WARPS = GROW_THREADS / 32;
constexpr int HIST_WORDS = 512;  // per warp

template <typename CodeT>
struct GrowArgs {
    // dataset
    const CodeT* codes;
    int64_t stride;
    const uint8_t* labels;
    int32_t n_classes;
    const uint8_t* col_kind;
    const int32_t* col_tag;
    const int32_t* col_nuniq;
    const int64_t* tab_off;
    const double* tables;
    // batch
    int32_t n_train, n_val, n_cand, max_depth, mss, nt_pad, max_nodes, fmax;
    const int32_t* idx_train;
    const int32_t* idx_val;
    const int32_t* cand_slot;
    const int32_t* feat_off;
    const uint16_t* feat_list;
    // scratch (per CTA)
    CodeT* cc;
    uint8_t* lab;
    uint32_t* pos;
    WorkItem* work;
    uint32_t* resna;
    // outputs (per candidate)
    rfmc_node* nodes;
    int32_t* counts;
    uint8_t* votes;
    int32_t* nnodes;
    int32_t* correct;
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Warp-aggregated histogram increment: lanes with equal bins elect one leader that adds the
// group's population.  Must be called by all 32 lanes.
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool active, int lane) {
    uint32_t key = active ? bin : 0xFFFFFFFFu;
    uint32_t m = __match_any_sync(FULL, key);
    if (active && (__ffs(m) - 1) == lane) atomicAdd(&hist[bin], (uint32_t)__popc(m));
}

// Locate the bin of hist[0..256) that contains 0-based rank r; r becomes the rank inside that bin.
__device__ __forceinline__ uint32_t find_bin256(const uint32_t* hist, uint32_t& r, int lane) {
    uint32_t loc[8];
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        loc[q] = hist[lane * 8 + q];
        s += loc[q];
    }
    uint32_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    uint32_t exc = inc - s;
    bool mine = (r >= exc) && (r < inc);
    uint32_t m = __ballot_sync(FULL, mine);
    int L = m ? (__ffs(m) - 1) : 31;
    uint32_t digit = 0, rr = 0;
    if (lane == L) {
        uint32_t acc = exc;
        digit = lane * 8 + 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (r < acc + loc[q]) {
                digit = lane * 8 + q;
                rr = r - acc;
                break;
            }
            acc += loc[q];
        }
    }
    digit = __shfl_sync(FULL, digit, L);
    r = __shfl_sync(FULL, rr, L);
    return digit;
}

// The two middle order statistics (0-based ranks k0 and k0+1) of a node with more than 32 rows:
// MSD radix select over 8-bit digits; both ranks share one histogram until their prefixes diverge.
template <typename CodeT>
__device__ void select_two(const CodeT* col, const uint32_t* P, uint32_t n, int npass, uint32_t k0, uint32_t* hist,
                           int lane, uint32_t& ca, uint32_t& cb) {
    uint32_t prefA = 0, prefB = 0, rA = k0, rB = k0 + 1;
    bool same = true;
    for (int pass = npass - 1; pass >= 0; --pass) {
        const int shift = pass * 8;
        const uint32_t himask = (pass == npass - 1) ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = lane; i < (same ? 256 : 512); i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t base = 0; base < n; base += 32) {
            uint32_t i = base + lane;
            bool act = i < n;
            uint32_t c = act ? (uint32_t)col[P[i]] : 0u;
            uint32_t d = (c >> shift) & 255u;
            hist_add(hist, d, act && ((c & himask) == (prefA & himask)), lane);
            if (!same) hist_add(hist + 256, d, act && ((c & himask) == (prefB & himask)), lane);
        }
        __syncwarp();
        uint32_t dA = find_bin256(hist, rA, lane);
        uint32_t dB;
        if (same) {
            dB = find_bin256(hist, rB, lane);
            if (dB != dA) same = false;
        } else {
            dB = find_bin256(hist + 256, rB, lane);
        }
        prefA |= dA << shift;
        prefB |= dB << shift;
        __syncwarp();
    }
    ca = prefA;
    cb = prefB;
}

// numpy's q=0.5 linear interpolation between adjacent order statistics, in the column's dtype
// (see encode.median_value for the derivation).  No FMA contraction: explicit _rn intrinsics.
__device__ __forceinline__ double median_value(double va, double vb, bool odd, int tag) {
    const int kind = tag & 0xFF;
    if (kind == RFMC_TAG_F32) {
        float a = (float)va, b = (float)vb;
        float d = __fsub_rn(b, a);
        float r = odd ? __fadd_rn(a, __fmul_rn(d, 0.0f)) : __fsub_rn(b, __fmul_rn(d, 0.5f));
        return (double)r;
    }
    double d = __dsub_rn(vb, va);
    if (kind == RFMC_TAG_INT && ((tag >> 16) & 1)) {
        int bits = (tag >> 8) & 0xFF;
        double half = exp2((double)(bits - 1));
        if (d >= half) d = __dsub_rn(d, 2.0 * half);
    }
    return odd ? __dadd_rn(va, __dmul_rn(d, 0.0)) : __dsub_rn(vb, __dmul_rn(d, 0.5));
}

// Number of entries of the sorted table tab[0..K) that are < sv, given that every entry below lo
// is < sv and every entry at or above hi is >= sv.  32-ary search by the whole warp.
__device__ uint32_t count_less(const double* __restrict__ tab, uint32_t lo, uint32_t hi, double sv, int lane) {
    while (hi > lo) {
        uint32_t len = hi - lo;
        if (len <= 32) {
            bool p = (lane < (int)len) && (__ldg(&tab[lo + lane]) < sv);
            lo += __popc(__ballot_sync(FULL, p));
            break;
        }
        uint32_t chunk = (len + 31) / 32;
        uint32_t idx = lo + (lane + 1) * chunk - 1;
        bool p = (idx < hi) && (__ldg(&tab[idx]) < sv);
        uint32_t cnt = __popc(__ballot_sync(FULL, p));
        uint32_t nlo = lo + cnt * chunk;
        if (nlo > hi) nlo = hi;
        uint32_t nhi = nlo + chunk - 1;
        if (nhi > hi) nhi = hi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(FULL, v, d);
        v = o > v ? o : v;
    }
    return v;
}

template <typename CodeT>
struct CandCtx {
    const GrowArgs<CodeT>* a;
    CodeT* cc;
    uint8_t* lab;
    uint32_t* pos;
    uint32_t* resna;
    rfmc_node* nodes;
    int32_t* counts;
    uint8_t* votes;
    const uint16_t* flist;
    int nf;
};

// One growTree() call (model.py:266-291) for work item j of the current level, by one warp.
template <typename CodeT>
__device__ void process_node(const CandCtx<CodeT>& c, WorkItem* cur, uint32_t j, int level, uint32_t* hist, int lane) {
    const GrowArgs<CodeT>& a = *c.a;
    const WorkItem w = cur[j];
    const uint32_t n = w.hi - w.lo;
    const uint32_t* P = c.pos + (size_t)(level & 1) * a.nt_pad + w.lo;
    uint32_t* Pn = c.pos + (size_t)((level + 1) & 1) * a.nt_pad + w.lo;
    const int C = a.n_classes;
    const uint32_t lt = lanemask_lt();

    // ---- class histogram of the node (np.unique(target_vals), model.py:267-268) ----
    for (int i = lane; i < RFMC_MAX_CLASSES; i += 32) hist[i] = 0;
    __syncwarp();
    for (uint32_t base = 0; base < n; base += 32) {
        uint32_t i = base + lane;
        bool act = i < n;
        uint32_t l = act ? (uint32_t)c.lab[P[i]] : 0u;
        hist_add(hist, l, act, lane);
    }
    __syncwarp();
    const uint32_t cls0 = hist[lane];
    const uint32_t cls1 = hist[lane + 32];
    __syncwarp();
    const int present = __popc(__ballot_sync(FULL, cls0 > 0)) + __popc(__ballot_sync(FULL, cls1 > 0));
    const int depth = level + 1;
    bool leaf = (depth >= a.max_depth) || (present == 1);

    // outputs shared by leaf and split: class counts and the leaf vote (first max, sorted labels)
    if (lane < C) c.counts[(size_t)w.node * C + lane] = (int32_t)cls0;
    if (lane + 32 < C) c.counts[(size_t)w.node * C + lane + 32] = (int32_t)cls1;
    {
        unsigned long long k0 = ((unsigned long long)cls0 << 8) | (unsigned long long)(255 - lane);
        unsigned long long k1 = ((unsigned long long)cls1 << 8) | (unsigned long long)(255 - (lane + 32));
        if (lane >= C) k0 = 0;
        if (lane + 32 >= C) k1 = 0;
        unsigned long long k = warp_max_u64(k0 > k1 ? k0 : k1);
        if (lane == 0) c.votes[w.node] = (uint8_t)(255 - (int)(k & 255));
    }

    bool accepted = false;
    uint32_t thr_store = 0, thr_part = 0, na = 0, flags = 0, f_acc = 0, off_next = 0;
    bool cat_acc = false;
    double sv_acc = 0.0;

    if (!leaf) {
        for (int k = 0; k < c.nf; ++k) {  // model.py:273-280: at most one full rotation
            int fl = (int)w.off + k;
            if (fl >= c.nf) fl -= c.nf;
            const uint32_t f = c.flist[fl];
            const bool cat = a.col_kind[f] == RFMC_KIND_CATEGORICAL;
            const CodeT* col = c.cc + (size_t)fl * a.nt_pad;
            uint32_t t_store = 0, t_part = 0, cnt_a = 0, fl_flags = cat ? RFMC_FLAG_CATEGORICAL : 0u;
            double sv = 0.0;

            if (n == 2) {  // model.py:258-262
                uint32_t c0 = col[P[0]], c1 = col[P[1]];
                t_store = t_part = c0 > c1 ? c0 : c1;
                fl_flags |= RFMC_FLAG_TWO_ROW;
                cnt_a = 1;
                if (a.mss <= 1) {
                    accepted = true;
                    if (lane == 0) {
                        bool second_hi = c0 <= c1;
                        Pn[0] = second_hi ? P[1] : P[0];
                        Pn[1] = second_hi ? P[0] : P[1];
                    }
                }
            } else if (n <= 32) {
                const bool act = lane < (int)n;
                const uint32_t my = act ? (uint32_t)col[P[lane]] : 0xFFFFFFFFu;
                if (cat) {  // mode, ties -> lowest code (model.py:249-250)
                    uint32_t grp = __match_any_sync(FULL, my);
                    unsigned long long key =
                        act ? (((unsigned long long)__popc(grp) << 32) | (unsigned long long)(0xFFFFFFFFu - my)) : 0ull;
                    key = warp_max_u64(key);
                    t_store = t_part = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                } else {  // median (model.py:237)
                    uint32_t rank = 0;
                    for (uint32_t jj = 0; jj < n; ++jj) {
                        uint32_t o = __shfl_sync(FULL, my, jj);
                        rank += (o < my) || (o == my && (int)jj < lane);
                    }
                    const uint32_t k0 = (n - 1) >> 1;
                    uint32_t mA = __ballot_sync(FULL, act && rank == k0);
                    uint32_t mB = __ballot_sync(FULL, act && rank == k0 + 1);
                    uint32_t ca = __shfl_sync(FULL, my, __ffs(mA) - 1);
                    uint32_t cb = __shfl_sync(FULL, my, __ffs(mB) - 1);
                    const double* tab = a.tables + a.tab_off[f];
                    const uint32_t K = (uint32_t)a.col_nuniq[f];
                    double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                    sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                    if (sv != sv) {
                        t_store = t_part = 0xFFFFFFFFu;  // every comparison with NaN is False
                    } else {
                        uint32_t less;
                        if (sv == va) less = ca - 1;
                        else if (sv == vb) less = cb - 1;
                        else if (va < sv && sv < vb) less = count_less(tab, ca, cb - 1, sv, lane);
                        else less = count_less(tab, 0, K, sv, lane);
                        t_store = t_part = less + 1;
                    }
                }
                uint32_t ma = __ballot_sync(FULL, act && (cat ? my == t_part : my >= t_part));
                cnt_a = __popc(ma);
                if (!cat && (cnt_a == 0 || cnt_a == n) && t_store != 0xFFFFFFFFu) {  // model.py:242-245
                    const double* tab = a.tables + a.tab_off[f];
                    const uint32_t K = (uint32_t)a.col_nuniq[f];
                    uint32_t i0 = t_store - 1;
                    t_part = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
                    ma = __ballot_sync(FULL, act && my >= t_part);
                    cnt_a = __popc(ma);
                }
                if (cnt_a >= (uint32_t)a.mss && n - cnt_a >= (uint32_t)a.mss) {
                    accepted = true;
                    if (act) {
                        bool go = (ma >> lane) & 1u;
                        uint32_t mb = ~ma & ((n == 32) ? FULL : ((1u << n) - 1u));
                        uint32_t dst = go ? __popc(ma & lt) : cnt_a + __popc(mb & lt);
                        Pn[dst] = P[lane];
                    }
                }
            } else {
                const uint32_t K = (uint32_t)a.col_nuniq[f];
                if (cat) {
                    if (K + 1 <= (uint32_t)HIST_WORDS) {
                        for (uint32_t i = lane; i <= K; i += 32) hist[i] = 0;
                        __syncwarp();
                        for (uint32_t base = 0; base < n; base += 32) {
                            uint32_t i = base + lane;
                            bool act = i < n;
                            uint32_t cv = act ? (uint32_t)col[P[i]] : 0u;
                            hist_add(hist, cv, act, lane);
                        }
                        __syncwarp();
                        unsigned long long key = 0;
                        for (uint32_t i = lane; i <= K; i += 32) {
                            unsigned long long kk = ((unsigned long long)hist[i] << 32) | (unsigned long long)(0xFFFFFFFFu - i);
                            key = kk > key ? kk : key;
                        }
                        key = warp_max_u64(key);
                        t_store = t_part = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                        __syncwarp();
                    } else {  // high-cardinality categorical: all-pairs counting (rare path)
                        unsigned long long key = 0;
                        for (uint32_t base = 0; base < n; base += 32) {
                            bool act = base + lane < n;
                            uint32_t my = act ? (uint32_t)col[P[base + lane]] : 0xFFFFFFFFu;
                            uint32_t cnt = 0;
                            for (uint32_t b2 = 0; b2 < n; b2 += 32) {
                                uint32_t other = (b2 + lane < n) ? (uint32_t)col[P[b2 + lane]] : 0xFFFFFFFEu;
#pragma unroll 8
                                for (int s = 0; s < 32; ++s) cnt += (__shfl_sync(FULL, other, s) == my);
                            }
                            unsigned long long kk =
                                act ? (((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - my)) : 0ull;
                            key = kk > key ? kk : key;
                        }
                        key = warp_max_u64(key);
                        t_store = t_part = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                    }
                } else {
                    int npass = (K < 0x100u) ? 1 : ((K < 0x10000u) ? 2 : ((K < 0x1000000u) ? 3 : 4));
                    uint32_t ca, cb;
                    select_two<CodeT>(col, P, n, npass, (n - 1) >> 1, hist, lane, ca, cb);
                    const double* tab = a.tables + a.tab_off[f];
                    double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                    sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                    if (sv != sv) {
                        t_store = t_part = 0xFFFFFFFFu;
                    } else {
                        uint32_t less;
                        if (sv == va) less = ca - 1;
                        else if (sv == vb) less = cb - 1;
                        else if (va < sv && sv < vb) less = count_less(tab, ca, cb - 1, sv, lane);
                        else less = count_less(tab, 0, K, sv, lane);
                        t_store = t_part = less + 1;
                    }
                }
                // side sizes
                for (int attempt = 0; attempt < 2; ++attempt) {
                    cnt_a = 0;
                    for (uint32_t base = 0; base < n; base += 32) {
                        uint32_t i = base + lane;
                        bool act = i < n;
                        uint32_t cv = act ? (uint32_t)col[P[i]] : 0u;
                        bool go = act && (cat ? cv == t_part : cv >= t_part);
                        cnt_a += __popc(__ballot_sync(FULL, go));
                    }
                    if (cat || attempt == 1 || !(cnt_a == 0 || cnt_a == n) || t_store == 0xFFFFFFFFu) break;
                    const double* tab = a.tables + a.tab_off[f];
                    uint32_t i0 = t_store - 1;
                    uint32_t t2 = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
                    if (t2 == t_part) break;
                    t_part = t2;
                }
                if (cnt_a >= (uint32_t)a.mss && n - cnt_a >= (uint32_t)a.mss) {
                    accepted = true;
                    uint32_t oa = 0, ob = cnt_a;
                    for (uint32_t base = 0; base < n; base += 32) {  // stable partition
                        uint32_t i = base + lane;
                        bool act = i < n;
                        uint32_t pi = act ? P[i] : 0u;
                        uint32_t cv = act ? (uint32_t)col[pi] : 0u;
                        bool go = act && (cat ? cv == t_part : cv >= t_part);
                        uint32_t ma = __ballot_sync(FULL, go);
                        uint32_t mb = __ballot_sync(FULL, act && !go);
                        if (act) Pn[go ? oa + __popc(ma & lt) : ob + __popc(mb & lt)] = pi;
                        oa += __popc(ma);
                        ob += __popc(mb);
                    }
                }
            }
            if (accepted) {
                thr_store = t_store;
                thr_part = t_part;
                na = cnt_a;
                flags = fl_flags;
                f_acc = f;
                cat_acc = cat;
                sv_acc = (cat || (fl_flags & RFMC_FLAG_TWO_ROW)) ? 0.0 : sv;
                off_next = (uint32_t)((fl + 1 == c.nf) ? 0 : fl + 1);
                break;
            }
        }
    }
    (void)thr_part;
    (void)cat_acc;
    if (lane == 0) {
        rfmc_node nd;
        if (accepted) {
            nd.feat = (int32_t)f_acc;
            nd.thr = thr_store;
            nd.child = -1;  // assigned by the level scan
            nd.flags = flags;
            nd.sval = sv_acc;
            c.resna[j] = na;
            cur[j].off = off_next;
        } else {
            nd.feat = -1;
            nd.thr = 0;
            nd.child = -1;
            nd.flags = 0;
            nd.sval = 0.0;
            c.resna[j] = 0;
        }
        c.nodes[w.node] = nd;
    }
}

template <typename CodeT>
__global__ void __launch_bounds__(GROW_THREADS) grow_kernel(GrowArgs<CodeT> a) {
    __shared__ uint32_t s_hist[GROW_WARPS][HIST_WORDS];
    __shared__ uint32_t s_wtot[GROW_WARPS];
    __shared__ uint32_t s_ticket;
    __shared__ int s_correct;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();

    for (int cand = blockIdx.x; cand < a.n_cand; cand += gridDim.x) {
        CandCtx<CodeT> c;
        c.a = &a;
        c.cc = a.cc + (size_t)blockIdx.x * a.fmax * a.nt_pad;
        c.lab = a.lab + (size_t)blockIdx.x * a.nt_pad;
        c.pos = a.pos + (size_t)blockIdx.x * 2 * a.nt_pad;
        c.resna = a.resna + (size_t)blockIdx.x * a.nt_pad;
        c.nodes = a.nodes + (size_t)cand * a.max_nodes;
        c.counts = a.counts + (size_t)cand * a.max_nodes * a.n_classes;
        c.votes = a.votes + (size_t)cand * a.max_nodes;
        const int fo = a.feat_off[cand];
        c.flist = a.feat_list + fo;
        c.nf = a.feat_off[cand + 1] - fo;
        WorkItem* work = a.work + (size_t)blockIdx.x * 2 * a.nt_pad;
        const int slot = a.cand_slot[cand];
        const int32_t* idxT = a.idx_train + (size_t)slot * a.n_train;
        const int32_t* idxV = a.idx_val + (size_t)slot * a.n_val;

        // ---- gather the bootstrap rows once: coalesced row read -> compact [feature][row] ----
        for (int i = warp; i < a.n_train; i += GROW_WARPS) {
            const int64_t row = idxT[i];
            const CodeT* src = a.codes + row * a.stride;
            for (int fl = lane; fl < c.nf; fl += 32) c.cc[(size_t)fl * a.nt_pad + i] = __ldg(&src[c.flist[fl]]);
            if (lane == 0) {
                c.lab[i] = __ldg(&a.labels[row]);
                c.pos[i] = (uint32_t)i;
            }
        }
        if (tid == 0) {
            work[0] = WorkItem{0u, (uint32_t)a.n_train, 0u, 0u};
            s_correct = 0;
        }
        uint32_t count = 1, n_nodes = 1;
        int level = 0;
        __syncthreads();

        while (count > 0) {
            WorkItem* cur = work + (size_t)(level & 1) * a.nt_pad;
            WorkItem* nxt = work + (size_t)((level + 1) & 1) * a.nt_pad;
            if (tid == 0) s_ticket = 0;
            __syncthreads();
            // ---- phase 1: one warp per node ----
            while (true) {
                uint32_t j = 0;
                if (lane == 0) j = atomicAdd(&s_ticket, 1u);
                j = __shfl_sync(FULL, j, 0);
                if (j >= count) break;
                process_node<CodeT>(c, cur, j, level, s_hist[warp], lane);
            }
            __syncthreads();
            // ---- phase 2: deterministic child numbering (block-wide scan over split flags) ----
            uint32_t running = 0;
            for (uint32_t tile = 0; tile < count; tile += GROW_THREADS) {
                uint32_t j = tile + tid;
                uint32_t na = (j < count) ? c.resna[j] : 0u;
                bool split = na > 0;
                uint32_t bal = __ballot_sync(FULL, split);
                if (lane == 0) s_wtot[warp] = __popc(bal);
                __syncthreads();
                uint32_t woff = 0, tot = 0;
#pragma unroll
                for (int w8 = 0; w8 < GROW_WARPS; ++w8) {
                    uint32_t t = s_wtot[w8];
                    if (w8 < warp) woff += t;
                    tot += t;
                }
                if (split) {
                    uint32_t s = running + woff + __popc(bal & lt);
                    WorkItem w = cur[j];
                    uint32_t child = n_nodes + 2 * s;
                    c.nodes[w.node].child = (int32_t)child;
                    nxt[2 * s] = WorkItem{w.lo, w.lo + na, child, w.off};
                    nxt[2 * s + 1] = WorkItem{w.lo + na, w.hi, child + 1, w.off};
                }
                running += tot;
                __syncthreads();
            }
            n_nodes += 2 * running;
            count = 2 * running;
            ++level;
        }

        // ---- validationTree (model.py:296-314): hold-out rows through the finished tree ----
        int good = 0;
        for (int i = tid; i < a.n_val; i += GROW_THREADS) {
            const int64_t row = idxV[i];
            const CodeT* src = a.codes + row * a.stride;
            uint32_t node = 0;
            while (true) {
                const rfmc_node nd = c.nodes[node];
                if (nd.feat < 0) break;
                uint32_t cv = __ldg(&src[nd.feat]);
                bool go = (nd.flags & RFMC_FLAG_CATEGORICAL) ? (cv == nd.thr) : (cv >= nd.thr);
                node = (uint32_t)nd.child + (go ? 0u : 1u);
            }
            good += (c.votes[node] == __ldg(&a.labels[row]));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) good += __shfl_xor_sync(FULL, good, d);
        if (lane == 0 && good) atomicAdd(&s_correct, good);
        __syncthreads();
        if (tid == 0) {
            a.nnodes[cand] = (int32_t)n_nodes;
            a.correct[cand] = s_correct;
        }
        __syncthreads();
    }
}

template <typename CodeT>
static int launch_grow_t(rfmc_batch* b, cudaStream_t stream) {
    rfmc_dataset* ds = b->ds;
    GrowArgs<CodeT> a;
    a.codes = (const CodeT*)ds->d_codes;
    a.stride = ds->stride;
    a.labels = ds->d_labels;
    a.n_classes = ds->n_classes;
    a.col_kind = ds->d_kind;
    a.col_tag = ds->d_tag;
    a.col_nuniq = ds->d_nuniq;
    a.tab_off = ds->d_taboff;
    a.tables = ds->d_tables;
    a.n_train = b->n_train;
    a.n_val = b->n_val;
    a.n_cand = b->n_cand;
    a.max_depth = b->max_depth;
    a.mss = b->mss;
    a.nt_pad = b->nt_pad;
    a.max_nodes = b->max_nodes;
    a.fmax = b->fmax;
    a.idx_train = b->d_idx_train;
    a.idx_val = b->d_idx_val;
    a.cand_slot = b->d_cand_slot;
    a.feat_off = b->d_feat_off;
    a.feat_list = b->d_feat_list;
    a.cc = (CodeT*)b->d_cc;
    a.lab = b->d_lab;
    a.pos = b->d_pos;
    a.work = b->d_work;
    a.resna = b->d_resna;
    a.nodes = b->d_nodes;
    a.counts = b->d_counts;
    a.votes = b->d_votes;
    a.nnodes = b->d_nnodes;
    a.correct = b->d_correct;
    grow_kernel<CodeT><<<b->grid, GROW_THREADS, 0, stream>>>(a);
    RFMC_CUDA(cudaGetLastError());
    b->launches = 1;
    return 0;
}

int launch_grow(rfmc_batch* b, cudaStream_t stream) {
    switch (b->ds->code_width) {
        case 1: return launch_grow_t<uint8_t>(b, stream);
        case 2: return launch_grow_t<uint16_t>(b, stream);
        case 4: return launch_grow_t<uint32_t>(b, stream);
    }
    set_error("unsupported code width");
    return 1;
}

int grow_grid_size(const rfmc_dataset* ds, int n_cand) {
    int per_sm = 0;
    cudaError_t e;
    switch (ds->code_width) {
        case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grow_kernel<uint8_t>, GROW_THREADS, 0); break;
        case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grow_kernel<uint16_t>, GROW_THREADS, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grow_kernel<uint32_t>, GROW_THREADS, 0); break;
    }
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    long long g = (long long)ds->sm_count * per_sm;
    if (g > n_cand) g = n_cand;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- pack selected candidates' node tables back to back (survivor export) -------------------
__global__ void pack_kernel(const rfmc_node* nodes, const int32_t* counts, const int32_t* nnodes, int max_nodes, int C,
                            const int32_t* cand_ids, const int64_t* dst_off, rfmc_node* nodes_out, int32_t* counts_out) {
    const int sel = blockIdx.y;
    const int cand = cand_ids[sel];
    const int n = nnodes[cand];
    const rfmc_node* src = nodes + (size_t)cand * max_nodes;
    const int32_t* csrc = counts + (size_t)cand * max_nodes * C;
    rfmc_node* dst = nodes_out + dst_off[sel];
    int32_t* cdst = counts_out + dst_off[sel] * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += gridDim.x * blockDim.x) cdst[i] = csrc[i];
}

int launch_pack(rfmc_batch* b, const int32_t* d_cand_ids, const int64_t* d_dst_off, int32_t n_sel, rfmc_node* nodes_out,
                int32_t* counts_out, cudaStream_t stream) {
    if (n_sel <= 0) return 0;
    int bx = (b->max_nodes + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n_sel);
    pack_kernel<<<grid, 256, 0, stream>>>(b->d_nodes, b->d_counts, b->d_nnodes, b->max_nodes, b->ds->n_classes, d_cand_ids,
                                          d_dst_off, nodes_out, counts_out);
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rfmc
