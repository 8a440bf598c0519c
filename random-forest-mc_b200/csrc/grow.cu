// Candidate-tree grower + hold-out scorer for sm_100a.
//
// Replaces plantTree/growTree/splitData/genLeaf (model.py:222-293) and validationTree
// (model.py:296-314) of the reference for a whole batch of candidates in one launch.
//
// Mapping: one persistent CTA (256 threads = 8 warps) per candidate; the candidate's tree is grown
// breadth-first.  A node does exactly what one growTree() call does in the reference:
//   class histogram -> purity / depth test -> for each feature of the rotated list:
//   order statistics of the node's codes (median) or their mode -> split value in the column's
//   dtype arithmetic -> code threshold -> side sizes -> accept or rotate -> stable partition.
//
// A tree grown to purity on 8,000 rows has ~14,000 nodes; 59 % of its split nodes hold 2-4 rows,
// 93 % hold <= 32, and every leaf is pure (88 % are single rows).  So the level's work list is kept
// in three size classes, each with its own execution shape:
//   tiny  (<= 8 rows)  one THREAD per node (32 nodes in flight per warp, 19-comparator sorting network);
//   small (9..32 rows) one WARP per node, one row per lane (ballots, shuffles, match.any);
//   large (> 32 rows)  one WARP per node, 128 rows per iteration (4 independent gathers in flight
//                      per lane), MSD radix select over per-warp shared-memory histograms.
// A split's child that is provably a leaf already (a single row, or all rows of one class) is
// written by the parent and never becomes a work item: that removes ~half of all node visits.
//
// Node numbering is breadth-first and deterministic: every node carries its position in its level;
// after a level, a block-wide scan over the split flags (by level position) assigns child ids, so
// the output equals oracle/code_model.py array-for-array whatever order the warps ran in.
//
// Memory: the candidate's bootstrap rows are gathered ONCE from the row-major HBM code matrix
// (one coalesced row read per warp) into a compact per-candidate [feature][row] matrix that stays
// L2-resident; the row-position array and the labels live in shared memory (24 KB at 8,000
// rows), so a node's dependent-load chain is: work item -> (smem) positions -> codes -> value
// table.  Nodes of <= 32 rows hold their rows in registers and partition their segment in place;
// larger nodes partition into a global scratch segment and copy back.  On-chip: per-warp shared-memory histograms with warp-aggregated atomics (match.any),
// ballots and shuffles for rank counting, prefix sums and arg-max.
#include <cstdlib>

#include "common.cuh"

namespace rfmc {

constexpr int GROW_THREADS = 256;
#ifndef RFMC_GROW_MIN_CTAS
#define RFMC_GROW_MIN_CTAS 4  // 64 registers per thread; the kernel is latency-bound, see profiles/
#endif
constexpr int GROW_WARPS = GROW_THREADS / 32;
#ifdef RFMC_GROW_TIMING
// debug build only: per-level cycle counts of CTA 0's first candidate (see profiles/grow_levels.py)
__device__ unsigned long long g_dbg[8192];
#define RFMC_TICK(slot) if (blockIdx.x == 0 && cand == 0 && threadIdx.x == 0 && (slot) < 8192) g_dbg[(slot)] = clock64()
#else
#define RFMC_TICK(slot)
#endif
constexpr int HIST_WORDS = 512;  // per warp
constexpr uint32_t NOT_FINAL = 0xFFu;
constexpr int TINY_MAX = 8, SMALL_MAX = 32, MEDIUM_MAX = 128, HUGE_MIN = 1024;  // the sorting network in process_tiny is for 8

template <typename CodeT, typename PosT>
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
    // per tree slot (written by slot_gather_kernel): compact [feature][row] codes and labels of the slot's
    // bootstrap rows -- shared by every candidate of the slot
    const CodeT* slot_codes;
    const uint8_t* slot_lab;
    int32_t n_features;
    // scratch (per CTA)
    PosT* pos_g;
    WorkItem* work;   // 2 buffers x list capacity
    uint32_t* resna;  // 2 buffers x nt_pad, indexed by level position
    uint32_t* rank;   // nt_pad
    uint16_t* resfin; // nt_pad
    int32_t use_smem;
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

// Digit histograms of a radix pass: 256 bins, a warp's 32 digits rarely coincide, so the election of
// hist_add (MATCH.ANY + leader) costs more than it saves -- one shared-memory atomic per lane.
__device__ __forceinline__ void hist_add_sparse(uint32_t* hist, uint32_t bin, bool active) {
    if (active) atomicAdd(&hist[bin], 1u);
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
// less_a / less_b: rows of the node with a code below ca / cb; eq_a / eq_b: rows with exactly that code --
// the side sizes of any split at ca, ca + 1, cb or cb + 1 follow from them without another pass over the rows.
template <typename CodeT, typename PosT>
__device__ void select_two(const CodeT* col, const PosT* P, uint32_t n, int npass, uint32_t k0, uint32_t* hist, int lane,
                           uint32_t& ca, uint32_t& cb, uint32_t& less_a, uint32_t& eq_a, uint32_t& less_b, uint32_t& eq_b) {
    uint32_t prefA = 0, prefB = 0, rA = k0, rB = k0 + 1;
    bool same = true;
    for (int pass = npass - 1; pass >= 0; --pass) {
        const int shift = pass * 8;
        const uint32_t himask = (pass == npass - 1) ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = lane; i < (same ? 256 : 512); i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t base = 0; base < n; base += 128) {
            uint32_t c[4];
            bool act[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * 32 + lane;
                act[u] = i < n;
                c[u] = act[u] ? (uint32_t)P[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) c[u] = act[u] ? (uint32_t)col[c[u]] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (base + u * 32 >= n) break;
                const uint32_t d = (c[u] >> shift) & 255u;
                hist_add_sparse(hist, d, act[u] && ((c[u] & himask) == (prefA & himask)));
                if (!same) hist_add_sparse(hist + 256, d, act[u] && ((c[u] & himask) == (prefB & himask)));
            }
        }
        __syncwarp();
        uint32_t dA = find_bin256(hist, rA, lane);
        uint32_t dB;
        if (same) {
            dB = find_bin256(hist, rB, lane);
            eq_b = hist[dB];
            if (dB != dA) same = false;
        } else {
            dB = find_bin256(hist + 256, rB, lane);
            eq_b = hist[256 + dB];
        }
        eq_a = hist[dA];  // after the last pass: rows equal to the selected code
        prefA |= dA << shift;
        prefB |= dB << shift;
        __syncwarp();
    }
    ca = prefA;
    cb = prefB;
    less_a = k0 - rA;  // the rank inside the final bin is what is left of the rank
    less_b = k0 + 1 - rB;
}

// Rows of the node with code >= t, for the thresholds a median split can take (numpy's `>=`, then `>`
// when a side is empty, model.py:238-245); false when t is none of them and the rows have to be counted.
__device__ __forceinline__ bool side_from_select(uint32_t t, uint32_t n, uint32_t ca, uint32_t cb, uint32_t less_a, uint32_t eq_a,
                                                 uint32_t less_b, uint32_t eq_b, uint32_t& cnt_a) {
    if (t == ca) cnt_a = n - less_a;
    else if (cb == 0u) return false;  // categorical mode: only ca is set
    else if (t == cb) cnt_a = n - less_b;
    else if (t == ca + 1u) cnt_a = n - less_a - eq_a;
    else if (t == cb + 1u) cnt_a = n - less_b - eq_b;
    else return false;
    return true;
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
// is < sv and every entry at or above hi is >= sv (binary search, one thread).
__device__ uint32_t count_less_scalar(const double* __restrict__ tab, uint32_t lo, uint32_t hi, double sv) {
    while (hi > lo) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(&tab[mid]) < sv)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// A split stores lower_bound(table, sv) + 1 as its code threshold.  That search is the longest
// dependent-load chain of a node, and the node itself does not need its result: no row OF THE NODE has a code strictly between the two middle order statistics, so when
// va < sv < vb every threshold in (ca, cb] partitions the node's rows identically.  The node is split
// with cb, the exact threshold is marked pending (thr = ca, the search's lower bound) and resolved
// for all nodes of the tree at once, thread per node, before the hold-out rows are scored.
constexpr uint32_t FLAG_PENDING = 0x80000000u;  // internal: never leaves the kernel
__device__ __forceinline__ uint32_t threshold_deferred(const double* tab, uint32_t K, double sv, double va, double vb,
                                                       uint32_t ca, uint32_t cb, uint32_t& t_store, uint32_t& pending) {
    pending = 0;
    if (sv != sv) return t_store = 0xFFFFFFFFu;  // every comparison with NaN is False
    if (sv == va) return t_store = ca;
    if (sv == vb) return t_store = cb;
    if (va < sv && sv < vb) {
        pending = FLAG_PENDING;
        t_store = ca;
        return cb;
    }
    return t_store = count_less_scalar(tab, 0, K, sv) + 1;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(FULL, v, d);
        v = o > v ? o : v;
    }
    return v;
}

__device__ __forceinline__ int size_class(int n) {
    return n <= TINY_MAX ? 0 : (n <= SMALL_MAX ? 1 : (n <= MEDIUM_MAX ? 2 : (n <= HUGE_MIN ? 3 : 4)));
}

// dynamic shared memory of a grower CTA: labels | positions | split bitmap | bitmap prefix
template <typename PosT>
__host__ __device__ __forceinline__ size_t smem_bits_offset(int nt_pad) {
    const size_t lab = (((size_t)nt_pad + 15) / 16) * 16;
    return (lab + (size_t)nt_pad * sizeof(PosT) + 15) / 16 * 16;
}

template <typename CodeT, typename PosT>
struct CandCtx {
    const GrowArgs<CodeT, PosT>* a;
    const CodeT* cc;  // this candidate's slot matrix: column of feature f at cc + f * nt_pad
    const uint8_t* lab;
    PosT* pos;  // nt_pad: every node owns a segment; small/tiny nodes partition it in place
    PosT* tmp;  // nt_pad (global): partition target of > 32-row nodes, copied back
    uint32_t* resna;  // this level's buffer
    uint32_t* sbits;  // shared-memory bitmap of the level positions that split (nullptr: ranks come from resna)
    uint16_t* resfin;
    rfmc_node* nodes;
    int32_t* counts;
    uint8_t* votes;
    const uint16_t* flist;
    int nf;
    uint32_t base;  // node id of level position 0
    int level;
};

// A node that split marks its level position: the child numbering (phase 2 of a level) ranks the set bits.
__device__ __forceinline__ void mark_split(uint32_t* sbits, uint32_t lpos, uint32_t na) {
    if (sbits != nullptr && na > 0) atomicOr(&sbits[lpos >> 5], 1u << (lpos & 31));
}

__device__ __forceinline__ void write_leaf(rfmc_node* nd) {
    nd->feat = -1;
    nd->thr = 0;
    nd->child = -1;
    nd->flags = 0;
    nd->sval = 0.0;
}

// ---------------------------------------------------------------------------------------------
// tiny nodes: one THREAD per node, n in 1..8
// ---------------------------------------------------------------------------------------------
template <typename CodeT, typename PosT>
__device__ void process_tiny(const CandCtx<CodeT, PosT>& c, WorkItem* item) {
    const GrowArgs<CodeT, PosT>& a = *c.a;
    const WorkItem w = *item;
    const int n = (int)(w.hi - w.lo);
    PosT* P = c.pos + w.lo;  // rows are in registers: the partition goes back in place
    const int C = a.n_classes;
    const uint32_t node = c.base + w.lpos;

    uint32_t p[TINY_MAX];
    uint32_t l[TINY_MAX];
#pragma unroll
    for (int q = 0; q < TINY_MAX; ++q) {
        p[q] = (q < n) ? (uint32_t)P[q] : 0u;
        l[q] = (q < n) ? (uint32_t)c.lab[p[q]] : 0xFFFFu;
    }
    // class counts + vote (first max in sorted-label order)
    int32_t* crow = c.counts + (size_t)node * C;
    uint32_t best_c = 0, best_n = 0;
    for (int k = 0; k < C; ++k) {
        uint32_t cnt = 0;
#pragma unroll
        for (int q = 0; q < TINY_MAX; ++q) cnt += (l[q] == (uint32_t)k);
        crow[k] = (int32_t)cnt;
        if (cnt > best_n) {
            best_n = cnt;
            best_c = (uint32_t)k;
        }
    }
    c.votes[node] = (uint8_t)best_c;
    const bool pure = (int)best_n == n;
    const int depth = c.level + 1;
    rfmc_node nd;
    write_leaf(&nd);
    uint32_t na_out = 0, fin_out = NOT_FINAL | (NOT_FINAL << 8);
    if (!(depth >= a.max_depth || pure)) {
        for (int k = 0; k < c.nf; ++k) {  // model.py:273-280
            int fl = (int)w.off + k;
            if (fl >= c.nf) fl -= c.nf;
            const uint32_t f = c.flist[fl];
            const bool cat = a.col_kind[f] == RFMC_KIND_CATEGORICAL;
            const CodeT* col = c.cc + (size_t)f * a.nt_pad;
            uint32_t cv[TINY_MAX];
#pragma unroll
            for (int q = 0; q < TINY_MAX; ++q) cv[q] = (q < n) ? (uint32_t)col[p[q]] : 0xFFFFFFFFu;
            uint32_t t_store, t_part, go = 0, pend = 0, fl_flags = cat ? RFMC_FLAG_CATEGORICAL : 0u;
            double sv = 0.0;
            if (n == 2) {  // model.py:258-262
                t_store = t_part = cv[0] > cv[1] ? cv[0] : cv[1];
                fl_flags |= RFMC_FLAG_TWO_ROW;
                go = (cv[0] <= cv[1]) ? 2u : 1u;
            } else if (cat) {  // mode, ties -> lowest code
                uint32_t bc = 0, bn = 0;
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) {
                    if (q < n) {
                        uint32_t cnt = 0;
#pragma unroll
                        for (int r = 0; r < TINY_MAX; ++r) cnt += (cv[r] == cv[q]);
                        if (cnt > bn || (cnt == bn && cv[q] < bc)) {
                            bn = cnt;
                            bc = cv[q];
                        }
                    }
                }
                t_store = t_part = bc;
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) go |= ((q < n) && cv[q] == bc) ? (1u << q) : 0u;
            } else {  // median: ranks (n-1)/2 and (n-1)/2 + 1 of the sorted codes (absent rows sort last)
                uint32_t s[TINY_MAX];
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) s[q] = cv[q];
                uint32_t t;
#define RFMC_CSWAP(i, j) if (s[i] > s[j]) { t = s[i]; s[i] = s[j]; s[j] = t; }
                RFMC_CSWAP(0, 1) RFMC_CSWAP(2, 3) RFMC_CSWAP(4, 5) RFMC_CSWAP(6, 7)
                RFMC_CSWAP(0, 2) RFMC_CSWAP(1, 3) RFMC_CSWAP(4, 6) RFMC_CSWAP(5, 7)
                RFMC_CSWAP(1, 2) RFMC_CSWAP(5, 6)
                RFMC_CSWAP(0, 4) RFMC_CSWAP(1, 5) RFMC_CSWAP(2, 6) RFMC_CSWAP(3, 7)
                RFMC_CSWAP(2, 4) RFMC_CSWAP(3, 5)
                RFMC_CSWAP(1, 2) RFMC_CSWAP(3, 4) RFMC_CSWAP(5, 6)
#undef RFMC_CSWAP
                const int k0 = (n - 1) >> 1;
                uint32_t ca = 0, cb = 0;
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) {
                    if (q == k0) ca = s[q];
                    if (q == k0 + 1) cb = s[q];
                }
                const double* tab = a.tables + a.tab_off[f];
                const uint32_t K = (uint32_t)a.col_nuniq[f];
                const double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                t_part = threshold_deferred(tab, K, sv, va, vb, ca, cb, t_store, pend);
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) go |= ((q < n) && cv[q] >= t_part) ? (1u << q) : 0u;
                const int cnt = __popc(go);
                if ((cnt == 0 || cnt == n) && t_store != 0xFFFFFFFFu && !pend) {  // model.py:242-245
                    const uint32_t i0 = t_store - 1;
                    t_part = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
                    go = 0;
#pragma unroll
                    for (int q = 0; q < TINY_MAX; ++q) go |= ((q < n) && cv[q] >= t_part) ? (1u << q) : 0u;
                }
            }
            const int cnt_a = __popc(go);
            if (cnt_a >= a.mss && n - cnt_a >= a.mss) {
                // stable partition + purity of both sides
                uint32_t oa = 0, ob = (uint32_t)cnt_a, la = 0xFFFFu, lb = 0xFFFFu;
                bool pa = true, pb = true;
#pragma unroll
                for (int q = 0; q < TINY_MAX; ++q) {
                    if (q < n) {
                        if ((go >> q) & 1u) {
                            P[oa++] = (PosT)p[q];
                            if (la == 0xFFFFu) la = l[q];
                            pa = pa && (l[q] == la);
                        } else {
                            P[ob++] = (PosT)p[q];
                            if (lb == 0xFFFFu) lb = l[q];
                            pb = pb && (l[q] == lb);
                        }
                    }
                }
                nd.feat = (int32_t)f;
                nd.thr = t_store;
                nd.flags = fl_flags | pend;
                nd.sval = (cat || (fl_flags & RFMC_FLAG_TWO_ROW)) ? 0.0 : sv;
                na_out = (uint32_t)cnt_a;
                fin_out = (pa ? la : NOT_FINAL) | ((pb ? lb : NOT_FINAL) << 8);
                item->off = (uint32_t)((fl + 1 == c.nf) ? 0 : fl + 1);
                break;
            }
        }
    }
    c.nodes[node] = nd;
    c.resna[w.lpos] = na_out;
    mark_split(c.sbits, w.lpos, na_out);
    c.resfin[w.lpos] = (uint16_t)fin_out;
}

// ---------------------------------------------------------------------------------------------
// small (9..32 rows) and medium (33..128 rows) nodes: a GROUP of G lanes per node, 4 rows per lane,
// everything about the node in registers.  G = 8 runs four nodes per warp (four independent
// dependent-load chains in flight instead of one), G = 32 one node per warp.  The rows' codes are
// read once per attempted feature; the order statistics come from all-pairs rank counting against
// the group's keys in shared memory (broadcast reads, no collectives inside the counting loop);
// the partition goes back in place.  Warp collectives (ballots, shuffles) are only issued from
// warp-uniform code: groups that are done or empty carry dummy values through them.
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ uint32_t group_ballot(bool p, int gi) {
    const uint32_t b = __ballot_sync(FULL, p);
    if (G == 32) return b;
    return (b >> (gi * G)) & ((1u << (G & 31)) - 1u);
}
template <int G>
__device__ __forceinline__ uint32_t group_sum(uint32_t v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_min(uint32_t v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ unsigned long long group_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, v, d);
        v = o > v ? o : v;
    }
    return v;
}

template <typename CodeT, typename PosT, int G>
__device__ void process_group(const CandCtx<CodeT, PosT>& c, WorkItem* items, uint32_t n_valid, uint32_t* wsm, int lane) {
    constexpr int NG = 32 / G, R = 4, CAP = R * G, GW = HIST_WORDS / NG, HW = GW - CAP - 2;
    constexpr bool PACKED = sizeof(CodeT) <= 2;  // key = code << 7 | row: one compare per pair
    static_assert(HW >= RFMC_MAX_CLASSES, "group scratch too small for the class histogram");
    const GrowArgs<CodeT, PosT>& a = *c.a;
    const int gi = lane / G, gl = lane % G;
    const bool valid = (uint32_t)gi < n_valid;
    uint32_t* s_keys = wsm + gi * GW;  // CAP words
    uint32_t* s_sel = s_keys + CAP;    // 2 words: codes of the two middle order statistics
    uint32_t* s_h = s_sel + 2;         // HW words: class histogram, then categorical histograms
    WorkItem w = WorkItem{0u, 0u, 0u, 0u};
    if (valid) w = items[gi];
    const uint32_t n = w.hi - w.lo;  // 0 for an empty group
    PosT* P = c.pos + w.lo;
    const int C = a.n_classes;
    const uint32_t node = c.base + w.lpos;
    const uint32_t ltg = (1u << gl) - 1u;
    const uint32_t gmask = (G == 32) ? FULL : ((1u << (G & 31)) - 1u);

    uint32_t p[R], l[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const uint32_t r = (uint32_t)(q * G + gl);
        p[q] = (r < n) ? (uint32_t)P[r] : 0u;
        l[q] = (r < n) ? (uint32_t)c.lab[p[q]] : 0xFFu;
    }
    // ---- class histogram (np.unique(target_vals), model.py:267-268) ----
    for (int b = gl; b < C; b += G) s_h[b] = 0;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < R; ++q)
        if ((uint32_t)(q * G + gl) < n) atomicAdd(&s_h[l[q]], 1u);
    __syncwarp();
    uint32_t present = 0;
    unsigned long long vk = 0;
    for (int b = gl; b < C; b += G) {
        const uint32_t cnt = s_h[b];
        if (valid) c.counts[(size_t)node * C + b] = (int32_t)cnt;
        present += cnt > 0;
        const unsigned long long kk = ((unsigned long long)cnt << 8) | (unsigned long long)(255 - b);
        vk = kk > vk ? kk : vk;
    }
    present = group_sum<G>(present);
    vk = group_max_u64<G>(vk);
    if (valid && gl == 0) c.votes[node] = (uint8_t)(255 - (int)(vk & 255));
    __syncwarp();  // s_h is reused by categorical attempts

    const int depth = c.level + 1;
    bool searching = valid && c.nf > 0 && !(depth >= a.max_depth || present == 1);
    bool accepted = false;
    uint32_t thr_store = 0, na = 0, flags = 0, f_acc = 0, off_next = 0, fin = NOT_FINAL | (NOT_FINAL << 8);
    double sv_acc = 0.0;
    int k = 0;

    while (__any_sync(FULL, searching)) {  // model.py:273-280: at most one full rotation per node
        int fl = (int)w.off + k;
        if (fl >= c.nf) fl -= c.nf;
        if (!searching) fl = 0;
        const uint32_t f = c.flist[fl];
        const bool cat = a.col_kind[f] == RFMC_KIND_CATEGORICAL;
        const CodeT* col = c.cc + (size_t)f * a.nt_pad;
        const uint32_t K = (uint32_t)a.col_nuniq[f];
        const bool use_hist = cat && (K + 1 <= (uint32_t)HW);
        uint32_t cv[R], key[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const uint32_t r = (uint32_t)(q * G + gl);
            cv[q] = (searching && r < n) ? (uint32_t)col[p[q]] : 0xFFFFFFFFu;
            key[q] = PACKED ? ((cv[q] << 7) | r) : cv[q];
        }
        if (searching) {
            if (use_hist) {
                for (uint32_t b = gl; b <= K; b += G) s_h[b] = 0;
            } else if (cat || G != 32 || !PACKED) {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if ((uint32_t)(q * G + gl) < n) s_keys[q * G + gl] = key[q];
            }
        }
        __syncwarp();
        unsigned long long mk = 0;
        if (searching) {
            if (!cat && G == 32 && PACKED) {
                // median (model.py:237), whole warp on one node: bitwise radix select of the key of rank
                // (n-1)/2 over the registers (4 compares + 4 ballots per bit; the branch is warp-uniform),
                // then the smallest key above it.  3x fewer instructions than all-pairs at 128 rows.
                const uint32_t k0 = (n - 1) >> 1;
                uint32_t res = 0;
                for (int bit = 31 - __clz((int)((K << 7) | 127u)); bit >= 0; --bit) {
                    const uint32_t trial = res | (1u << bit);
                    uint32_t below = 0;
#pragma unroll
                    for (int q = 0; q < R; ++q) below += __popc(__ballot_sync(FULL, key[q] < trial));
                    if (below <= k0) res = trial;
                }
                uint32_t nx = 0xFFFFFFFFu;
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (key[q] > res) nx = min(nx, key[q]);
                nx = group_min<32>(nx);
                if (lane == 0) {
                    s_sel[0] = res >> 7;
                    s_sel[1] = nx >> 7;
                }
            } else if (!cat) {  // median by all-pairs rank counting against the group's keys
                uint32_t rk[R];
#pragma unroll
                for (int q = 0; q < R; ++q) rk[q] = 0;
                for (uint32_t jj = 0; jj < n; ++jj) {
                    const uint32_t o = s_keys[jj];
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        if (PACKED)
                            rk[q] += (o < key[q]);
                        else
                            rk[q] += (o < key[q]) || (o == key[q] && jj < (uint32_t)(q * G + gl));
                    }
                }
                const uint32_t k0 = (n - 1) >> 1;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if ((uint32_t)(q * G + gl) < n) {
                        if (rk[q] == k0) s_sel[0] = cv[q];
                        if (rk[q] == k0 + 1) s_sel[1] = cv[q];
                    }
                }
            } else if (use_hist) {  // mode (model.py:249-250)
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if ((uint32_t)(q * G + gl) < n) atomicAdd(&s_h[cv[q]], 1u);
            } else {  // high-cardinality categorical: all-pairs equality counting (rare path)
                uint32_t eq[R];
#pragma unroll
                for (int q = 0; q < R; ++q) eq[q] = 0;
                for (uint32_t jj = 0; jj < n; ++jj) {
                    const uint32_t o = PACKED ? (s_keys[jj] >> 7) : s_keys[jj];
#pragma unroll
                    for (int q = 0; q < R; ++q) eq[q] += (o == cv[q]);
                }
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if ((uint32_t)(q * G + gl) < n) {
                        const unsigned long long kk = ((unsigned long long)eq[q] << 32) | (unsigned long long)(0xFFFFFFFFu - cv[q]);
                        mk = kk > mk ? kk : mk;
                    }
                }
            }
        }
        __syncwarp();
        uint32_t t_store = 0, t_part = 0, pend = 0;
        double sv = 0.0;
        const double* tab = a.tables + a.tab_off[f];
        if (searching) {
            if (!cat) {
                const uint32_t ca = s_sel[0], cb = s_sel[1];
                const double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                t_part = threshold_deferred(tab, K, sv, va, vb, ca, cb, t_store, pend);
            } else if (use_hist) {
                for (uint32_t b = gl; b <= K; b += G) {
                    const unsigned long long kk = ((unsigned long long)s_h[b] << 32) | (unsigned long long)(0xFFFFFFFFu - b);
                    mk = kk > mk ? kk : mk;
                }
            }
        }
        mk = group_max_u64<G>(mk);
        if (searching && cat) t_store = t_part = 0xFFFFFFFFu - (uint32_t)(mk & 0xFFFFFFFFull);
        // ---- side sizes ----
        uint32_t ba[R], cnt_a = 0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const bool go = searching && (uint32_t)(q * G + gl) < n && (cat ? cv[q] == t_part : cv[q] >= t_part);
            ba[q] = group_ballot<G>(go, gi);
            cnt_a += __popc(ba[q]);
        }
        const bool degenerate = searching && !cat && (cnt_a == 0 || cnt_a == n) && t_store != 0xFFFFFFFFu && !pend;
        if (__any_sync(FULL, degenerate)) {  // model.py:242-245: retry with '>'
            if (degenerate) {
                const uint32_t i0 = t_store - 1;
                t_part = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
            }
            cnt_a = 0;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const bool go = searching && (uint32_t)(q * G + gl) < n && (cat ? cv[q] == t_part : cv[q] >= t_part);
                ba[q] = group_ballot<G>(go, gi);
                cnt_a += __popc(ba[q]);
            }
        }
        const bool acc_now = searching && cnt_a >= (uint32_t)a.mss && n - cnt_a >= (uint32_t)a.mss;
        if (__any_sync(FULL, acc_now)) {
            // a side whose rows all carry one label is a finished leaf (model.py:268)
            uint32_t ka = 0xFFFFFFFFu, kb = 0xFFFFFFFFu;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const uint32_t r = (uint32_t)(q * G + gl);
                if (r < n) {
                    const uint32_t kk = (r << 8) | l[q];
                    if ((ba[q] >> gl) & 1u)
                        ka = min(ka, kk);
                    else
                        kb = min(kb, kk);
                }
            }
            ka = group_min<G>(ka);
            kb = group_min<G>(kb);
            const uint32_t la = ka & 0xFFu, lb = kb & 0xFFu;
            bool bad_a = false, bad_b = false;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                if ((uint32_t)(q * G + gl) < n) {
                    if ((ba[q] >> gl) & 1u)
                        bad_a |= l[q] != la;
                    else
                        bad_b |= l[q] != lb;
                }
            }
            const bool pa = group_ballot<G>(bad_a, gi) == 0u, pb = group_ballot<G>(bad_b, gi) == 0u;
            if (acc_now) {  // stable partition, in place (every row of the node is in a register)
                uint32_t before_a = 0, before_b = 0;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const uint32_t r0 = (uint32_t)(q * G);
                    const uint32_t actm = (n >= r0 + G) ? gmask : (n > r0 ? ((1u << (n - r0)) - 1u) : 0u);
                    const uint32_t bb = ~ba[q] & actm;
                    if (r0 + gl < n) {
                        const bool go = (ba[q] >> gl) & 1u;
                        const uint32_t dst = go ? before_a + __popc(ba[q] & ltg) : cnt_a + before_b + __popc(bb & ltg);
                        P[dst] = (PosT)p[q];
                    }
                    before_a += __popc(ba[q]);
                    before_b += __popc(bb);
                }
                accepted = true;
                searching = false;
                thr_store = t_store;
                na = cnt_a;
                flags = (cat ? RFMC_FLAG_CATEGORICAL : 0u) | pend;
                f_acc = f;
                sv_acc = cat ? 0.0 : sv;
                off_next = (uint32_t)((fl + 1 == c.nf) ? 0 : fl + 1);
                fin = (pa ? la : NOT_FINAL) | ((pb ? lb : NOT_FINAL) << 8);
            }
        }
        if (searching) {
            ++k;
            if (k >= c.nf) searching = false;
        }
        __syncwarp();  // scratch is rewritten by the next attempt
    }
    if (valid && gl == 0) {
        rfmc_node nd;
        write_leaf(&nd);
        if (accepted) {
            nd.feat = (int32_t)f_acc;
            nd.thr = thr_store;
            nd.flags = flags;
            nd.sval = sv_acc;
            items[gi].off = off_next;
        }
        c.nodes[node] = nd;
        c.resna[w.lpos] = accepted ? na : 0u;
        mark_split(c.sbits, w.lpos, accepted ? na : 0u);
        c.resfin[w.lpos] = (uint16_t)fin;
    }
}

// ---------------------------------------------------------------------------------------------
// large (129..1024 rows) nodes: one WARP per node
// ---------------------------------------------------------------------------------------------
template <typename CodeT, typename PosT>
__device__ void process_warp(const CandCtx<CodeT, PosT>& c, WorkItem* item, uint32_t* hist, int lane) {
    const GrowArgs<CodeT, PosT>& a = *c.a;
    const WorkItem w = *item;
    const uint32_t n = w.hi - w.lo;
    PosT* P = c.pos + w.lo;
    PosT* Pn = c.tmp + w.lo;  // partition target, copied back
    const int C = a.n_classes;
    const uint32_t lt = lanemask_lt();
    const uint32_t node = c.base + w.lpos;

    // ---- class histogram of the node (np.unique(target_vals), model.py:267-268) ----
    for (int i = lane; i < RFMC_MAX_CLASSES; i += 32) hist[i] = 0;
    __syncwarp();
    {
        for (uint32_t base = 0; base < n; base += 128) {
            uint32_t l[4];
            bool act[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = base + u * 32 + lane;
                act[u] = i < n;
                l[u] = act[u] ? (uint32_t)P[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) l[u] = act[u] ? (uint32_t)c.lab[l[u]] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (base + u * 32 >= n) break;
                hist_add(hist, l[u], act[u], lane);
            }
        }
    }
    __syncwarp();
    const uint32_t cls0 = hist[lane];
    const uint32_t cls1 = hist[lane + 32];
    __syncwarp();
    const int present = __popc(__ballot_sync(FULL, cls0 > 0)) + __popc(__ballot_sync(FULL, cls1 > 0));
    const int depth = c.level + 1;
    const bool leaf = (depth >= a.max_depth) || (present == 1);

    if (lane < C) c.counts[(size_t)node * C + lane] = (int32_t)cls0;
    if (lane + 32 < C) c.counts[(size_t)node * C + lane + 32] = (int32_t)cls1;
    {
        unsigned long long k0 = ((unsigned long long)cls0 << 8) | (unsigned long long)(255 - lane);
        unsigned long long k1 = ((unsigned long long)cls1 << 8) | (unsigned long long)(255 - (lane + 32));
        if (lane >= C) k0 = 0;
        if (lane + 32 >= C) k1 = 0;
        unsigned long long k = warp_max_u64(k0 > k1 ? k0 : k1);
        if (lane == 0) c.votes[node] = (uint8_t)(255 - (int)(k & 255));
    }

    bool accepted = false;
    uint32_t thr_store = 0, na = 0, flags = 0, f_acc = 0, off_next = 0, fin = NOT_FINAL | (NOT_FINAL << 8);
    double sv_acc = 0.0;

    if (!leaf) {
        for (int k = 0; k < c.nf; ++k) {  // model.py:273-280: at most one full rotation
            int fl = (int)w.off + k;
            if (fl >= c.nf) fl -= c.nf;
            const uint32_t f = c.flist[fl];
            const bool cat = a.col_kind[f] == RFMC_KIND_CATEGORICAL;
            const CodeT* col = c.cc + (size_t)f * a.nt_pad;
            uint32_t t_store = 0, t_part = 0, cnt_a = 0, pend = 0;
            uint32_t sel_ca = 0, sel_cb = 0, less_a = 0, eq_a = 0, less_b = 0, eq_b = 0;
            bool have_sel = false;
            const uint32_t fl_flags = cat ? RFMC_FLAG_CATEGORICAL : 0u;
            double sv = 0.0;

            {
                const uint32_t K = (uint32_t)a.col_nuniq[f];
                if (cat) {
                    if (K + 1 <= (uint32_t)HIST_WORDS) {
                        for (uint32_t i = lane; i <= K; i += 32) hist[i] = 0;
                        __syncwarp();
                        for (uint32_t base = 0; base < n; base += 128) {
                            uint32_t cv[4];
                            bool act[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint32_t i = base + u * 32 + lane;
                                act[u] = i < n;
                                cv[u] = act[u] ? (uint32_t)P[i] : 0u;
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[cv[u]] : 0u;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (base + u * 32 >= n) break;
                                hist_add(hist, cv[u], act[u], lane);
                            }
                        }
                        __syncwarp();
                        unsigned long long key = 0;
                        for (uint32_t i = lane; i <= K; i += 32) {
                            unsigned long long kk = ((unsigned long long)hist[i] << 32) | (unsigned long long)(0xFFFFFFFFu - i);
                            key = kk > key ? kk : key;
                        }
                        key = warp_max_u64(key);
                        t_store = t_part = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                        sel_ca = t_part;  // rows equal to the mode: its histogram count
                        less_a = n - (uint32_t)(key >> 32);
                        have_sel = true;
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
                    select_two<CodeT, PosT>(col, P, n, npass, (n - 1) >> 1, hist, lane, sel_ca, sel_cb, less_a, eq_a, less_b, eq_b);
                    have_sel = true;
                    const uint32_t ca = sel_ca, cb = sel_cb;
                    const double* tab = a.tables + a.tab_off[f];
                    double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                    sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                    t_part = threshold_deferred(tab, K, sv, va, vb, ca, cb, t_store, pend);
                }
                // side sizes: from the select's rank bookkeeping when the threshold is one of its codes, else by counting
                for (int attempt = 0; attempt < 2; ++attempt) {
                    cnt_a = 0;
                    const bool known = have_sel && side_from_select(t_part, n, sel_ca, sel_cb, less_a, eq_a, less_b, eq_b, cnt_a);
                    for (uint32_t base = 0; !known && base < n; base += 128) {
                        uint32_t cv[4];
                        bool act[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t i = base + u * 32 + lane;
                            act[u] = i < n;
                            cv[u] = act[u] ? (uint32_t)P[i] : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[cv[u]] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool go = act[u] && (cat ? cv[u] == t_part : cv[u] >= t_part);
                            cnt_a += __popc(__ballot_sync(FULL, go));
                        }
                    }
                    if (cat || attempt == 1 || !(cnt_a == 0 || cnt_a == n) || t_store == 0xFFFFFFFFu || pend) break;
                    const double* tab = a.tables + a.tab_off[f];
                    uint32_t i0 = t_store - 1;
                    uint32_t t2 = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
                    if (t2 == t_part) break;
                    t_part = t2;
                }
                if (cnt_a >= (uint32_t)a.mss && n - cnt_a >= (uint32_t)a.mss) {
                    accepted = true;
                    uint32_t oa = 0, ob = cnt_a;
                    for (uint32_t base = 0; base < n; base += 128) {  // stable partition, four gathers in flight per lane
                        uint32_t pi[4], cv[4];
                        bool act[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t i = base + u * 32 + lane;
                            act[u] = i < n;
                            pi[u] = act[u] ? (uint32_t)P[i] : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[pi[u]] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (base + u * 32 >= n) break;
                            const bool go = act[u] && (cat ? cv[u] == t_part : cv[u] >= t_part);
                            const uint32_t ma = __ballot_sync(FULL, go);
                            const uint32_t mb = __ballot_sync(FULL, act[u] && !go);
                            if (act[u]) Pn[go ? oa + __popc(ma & lt) : ob + __popc(mb & lt)] = (PosT)pi[u];
                            oa += __popc(ma);
                            ob += __popc(mb);
                        }
                    }
                    __syncwarp();
                    for (uint32_t i = lane; i < n; i += 32) P[i] = Pn[i];
                }
            }
            if (accepted) {
                thr_store = t_store;
                na = cnt_a;
                flags = fl_flags | pend;
                f_acc = f;
                sv_acc = cat ? 0.0 : sv;
                off_next = (uint32_t)((fl + 1 == c.nf) ? 0 : fl + 1);
                break;
            }
        }
    }
    if (lane == 0) {
        rfmc_node nd;
        write_leaf(&nd);
        if (accepted) {
            nd.feat = (int32_t)f_acc;
            nd.thr = thr_store;
            nd.flags = flags;
            nd.sval = sv_acc;
            item->off = off_next;
        }
        c.nodes[node] = nd;
        c.resna[w.lpos] = accepted ? na : 0u;
        mark_split(c.sbits, w.lpos, accepted ? na : 0u);
        c.resfin[w.lpos] = (uint16_t)fin;
    }
}

// ---------------------------------------------------------------------------------------------
// huge nodes (> HUGE_MIN rows): the whole CTA works on ONE node.  The top levels of a tree hold one
// to a few such nodes; giving each to a single warp leaves seven warps of the CTA waiting at the
// level barrier.  Every warp takes a contiguous slice of the node's rows, accumulates into its own
// shared-memory histogram, the 8 histograms are summed into s_comb, and every warp then derives the
// same digits / mode / counts from s_comb, so all control flow stays CTA-uniform.  The stable
// partition uses per-warp side counts (slices are contiguous, so slice order = row order).
// ---------------------------------------------------------------------------------------------
template <typename CodeT, typename PosT>
__device__ void process_cta(const CandCtx<CodeT, PosT>& c, WorkItem* item, uint32_t (*s_hist)[HIST_WORDS], uint32_t* s_comb,
                            uint32_t* s_part, int tid) {
    const GrowArgs<CodeT, PosT>& a = *c.a;
    const int lane = tid & 31, warp = tid >> 5;
    const WorkItem w = *item;
    const uint32_t n = w.hi - w.lo;
    PosT* P = c.pos + w.lo;
    PosT* T = c.tmp + w.lo;
    const int C = a.n_classes;
    const uint32_t lt = lanemask_lt();
    const uint32_t node = c.base + w.lpos;
    uint32_t* hist = s_hist[warp];
    const uint32_t per = ((n + GROW_WARPS * 32 - 1) / (GROW_WARPS * 32)) * 32;
    const uint32_t r0 = (uint32_t)warp * per < n ? (uint32_t)warp * per : n;
    const uint32_t r1 = r0 + per < n ? r0 + per : n;

    // ---- class histogram ----
    for (int i = lane; i < RFMC_MAX_CLASSES; i += 32) hist[i] = 0;
    __syncwarp();
    for (uint32_t base = r0; base < r1; base += 128) {
        uint32_t l[4];
        bool act[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = base + u * 32 + lane;
            act[u] = i < r1;
            l[u] = act[u] ? (uint32_t)P[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) l[u] = act[u] ? (uint32_t)c.lab[l[u]] : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) hist_add(hist, l[u], act[u], lane);
    }
    __syncthreads();
    if (tid < RFMC_MAX_CLASSES) {
        uint32_t sum = 0;
#pragma unroll
        for (int w8 = 0; w8 < GROW_WARPS; ++w8) sum += s_hist[w8][tid];
        s_comb[tid] = sum;
    }
    __syncthreads();
    const uint32_t cls0 = s_comb[lane];
    const uint32_t cls1 = s_comb[lane + 32];
    const int present = __popc(__ballot_sync(FULL, cls0 > 0)) + __popc(__ballot_sync(FULL, cls1 > 0));
    const int depth = c.level + 1;
    const bool leaf = (depth >= a.max_depth) || (present == 1);
    if (warp == 0) {
        if (lane < C) c.counts[(size_t)node * C + lane] = (int32_t)cls0;
        if (lane + 32 < C) c.counts[(size_t)node * C + lane + 32] = (int32_t)cls1;
        unsigned long long k0 = ((unsigned long long)cls0 << 8) | (unsigned long long)(255 - lane);
        unsigned long long k1 = ((unsigned long long)cls1 << 8) | (unsigned long long)(255 - (lane + 32));
        if (lane >= C) k0 = 0;
        if (lane + 32 >= C) k1 = 0;
        unsigned long long k = warp_max_u64(k0 > k1 ? k0 : k1);
        if (lane == 0) c.votes[node] = (uint8_t)(255 - (int)(k & 255));
    }
    __syncthreads();  // s_comb is reused below

    bool accepted = false;
    uint32_t thr_store = 0, na = 0, flags = 0, f_acc = 0, off_next = 0;
    double sv_acc = 0.0;
    if (!leaf) {
        for (int k = 0; k < c.nf; ++k) {
            int fl = (int)w.off + k;
            if (fl >= c.nf) fl -= c.nf;
            const uint32_t f = c.flist[fl];
            const bool cat = a.col_kind[f] == RFMC_KIND_CATEGORICAL;
            const CodeT* col = c.cc + (size_t)f * a.nt_pad;
            const uint32_t K = (uint32_t)a.col_nuniq[f];
            uint32_t t_store = 0, t_part = 0, cnt_a = 0, pend = 0;
            const uint32_t fl_flags = cat ? RFMC_FLAG_CATEGORICAL : 0u;
            double sv = 0.0;
            if (cat) {
                if (K + 1 <= (uint32_t)HIST_WORDS) {
                    for (uint32_t i = lane; i <= K; i += 32) hist[i] = 0;
                    __syncwarp();
                    for (uint32_t base = r0; base < r1; base += 128) {
                        uint32_t cv[4];
                        bool act[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t i = base + u * 32 + lane;
                            act[u] = i < r1;
                            cv[u] = act[u] ? (uint32_t)P[i] : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[cv[u]] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u) hist_add(hist, cv[u], act[u], lane);
                    }
                    __syncthreads();
                    for (uint32_t b = tid; b <= K; b += GROW_THREADS) {
                        uint32_t sum = 0;
#pragma unroll
                        for (int w8 = 0; w8 < GROW_WARPS; ++w8) sum += s_hist[w8][b];
                        s_comb[b] = sum;
                    }
                    __syncthreads();
                    unsigned long long key = 0;
                    for (uint32_t i = lane; i <= K; i += 32) {
                        unsigned long long kk = ((unsigned long long)s_comb[i] << 32) | (unsigned long long)(0xFFFFFFFFu - i);
                        key = kk > key ? kk : key;
                    }
                    key = warp_max_u64(key);
                    t_store = t_part = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                } else {  // high-cardinality categorical: warp 0 counts all pairs (rare path)
                    if (warp == 0) {
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
                        if (lane == 0) s_comb[0] = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
                    }
                    __syncthreads();
                    t_store = t_part = s_comb[0];
                }
            } else {
                const int npass = (K < 0x100u) ? 1 : ((K < 0x10000u) ? 2 : ((K < 0x1000000u) ? 3 : 4));
                uint32_t prefA = 0, prefB = 0, rA = (n - 1) >> 1, rB = ((n - 1) >> 1) + 1;
                bool same = true;
                for (int pass = npass - 1; pass >= 0; --pass) {
                    const int shift = pass * 8;
                    const uint32_t himask = (pass == npass - 1) ? 0u : (0xFFFFFFFFu << (shift + 8));
                    for (int i = lane; i < (same ? 256 : 512); i += 32) hist[i] = 0;
                    __syncwarp();
                    for (uint32_t base = r0; base < r1; base += 128) {
                        uint32_t cv[4];
                        bool act[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t i = base + u * 32 + lane;
                            act[u] = i < r1;
                            cv[u] = act[u] ? (uint32_t)P[i] : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[cv[u]] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t d = (cv[u] >> shift) & 255u;
                            hist_add_sparse(hist, d, act[u] && ((cv[u] & himask) == (prefA & himask)));
                            if (!same) hist_add_sparse(hist + 256, d, act[u] && ((cv[u] & himask) == (prefB & himask)));
                        }
                    }
                    __syncthreads();
                    {
                        uint32_t sa = 0, sb = 0;
#pragma unroll
                        for (int w8 = 0; w8 < GROW_WARPS; ++w8) {
                            sa += s_hist[w8][tid];
                            if (!same) sb += s_hist[w8][256 + tid];
                        }
                        s_comb[tid] = sa;
                        if (!same) s_comb[256 + tid] = sb;
                    }
                    __syncthreads();
                    const uint32_t dA = find_bin256(s_comb, rA, lane);
                    uint32_t dB;
                    if (same) {
                        dB = find_bin256(s_comb, rB, lane);
                        if (dB != dA) same = false;
                    } else {
                        dB = find_bin256(s_comb + 256, rB, lane);
                    }
                    prefA |= dA << shift;
                    prefB |= dB << shift;
                }
                const uint32_t ca = prefA, cb = prefB;
                const double* tab = a.tables + a.tab_off[f];
                const double va = __ldg(&tab[ca - 1]), vb = __ldg(&tab[cb - 1]);
                sv = median_value(va, vb, (n & 1) != 0, a.col_tag[f]);
                t_part = threshold_deferred(tab, K, sv, va, vb, ca, cb, t_store, pend);
            }
            // ---- side sizes (per-warp slices, summed through shared memory) ----
            uint32_t mine_a = 0;
            for (int attempt = 0; attempt < 2; ++attempt) {
                mine_a = 0;
                for (uint32_t base = r0; base < r1; base += 128) {
                    uint32_t cv[4];
                    bool act[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = base + u * 32 + lane;
                        act[u] = i < r1;
                        cv[u] = act[u] ? (uint32_t)P[i] : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[cv[u]] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool go = act[u] && (cat ? cv[u] == t_part : cv[u] >= t_part);
                        mine_a += __popc(__ballot_sync(FULL, go));
                    }
                }
                __syncthreads();  // previous readers of s_part are done
                if (lane == 0) s_part[warp] = mine_a;
                __syncthreads();
                cnt_a = 0;
#pragma unroll
                for (int w8 = 0; w8 < GROW_WARPS; ++w8) cnt_a += s_part[w8];
                if (cat || attempt == 1 || !(cnt_a == 0 || cnt_a == n) || t_store == 0xFFFFFFFFu || pend) break;
                const double* tab = a.tables + a.tab_off[f];
                const uint32_t i0 = t_store - 1;
                const uint32_t t2 = t_store + ((i0 < K && __ldg(&tab[i0]) == sv) ? 1u : 0u);
                if (t2 == t_part) break;
                t_part = t2;
            }
            if (cnt_a >= (uint32_t)a.mss && n - cnt_a >= (uint32_t)a.mss) {
                accepted = true;
                uint32_t oa = 0, ob = cnt_a;  // slices are contiguous: slice order is row order
#pragma unroll
                for (int w8 = 0; w8 < GROW_WARPS; ++w8) {
                    if (w8 < warp) {
                        const uint32_t s0 = (uint32_t)w8 * per < n ? (uint32_t)w8 * per : n;
                        const uint32_t s1 = s0 + per < n ? s0 + per : n;
                        oa += s_part[w8];
                        ob += (s1 - s0) - s_part[w8];
                    }
                }
                for (uint32_t base = r0; base < r1; base += 128) {  // four gathers in flight per lane
                    uint32_t pi[4], cv[4];
                    bool act[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = base + u * 32 + lane;
                        act[u] = i < r1;
                        pi[u] = act[u] ? (uint32_t)P[i] : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) cv[u] = act[u] ? (uint32_t)col[pi[u]] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (base + u * 32 >= r1) break;
                        const bool go = act[u] && (cat ? cv[u] == t_part : cv[u] >= t_part);
                        const uint32_t ma = __ballot_sync(FULL, go);
                        const uint32_t mb = __ballot_sync(FULL, act[u] && !go);
                        if (act[u]) T[go ? oa + __popc(ma & lt) : ob + __popc(mb & lt)] = (PosT)pi[u];
                        oa += __popc(ma);
                        ob += __popc(mb);
                    }
                }
                __syncthreads();
                for (uint32_t i = tid; i < n; i += GROW_THREADS) P[i] = T[i];
                __syncthreads();
                thr_store = t_store;
                na = cnt_a;
                flags = fl_flags | pend;
                f_acc = f;
                sv_acc = cat ? 0.0 : sv;
                off_next = (uint32_t)((fl + 1 == c.nf) ? 0 : fl + 1);
                break;
            }
        }
    }
    if (tid == 0) {
        rfmc_node nd;
        write_leaf(&nd);
        if (accepted) {
            nd.feat = (int32_t)f_acc;
            nd.thr = thr_store;
            nd.flags = flags;
            nd.sval = sv_acc;
            item->off = off_next;
        }
        c.nodes[node] = nd;
        c.resna[w.lpos] = accepted ? na : 0u;
        mark_split(c.sbits, w.lpos, accepted ? na : 0u);
        c.resfin[w.lpos] = (uint16_t)(NOT_FINAL | (NOT_FINAL << 8));
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Slot gather: the bootstrap rows of every tree slot, read once from the row-major code matrix and
// written as a compact [slot][feature][row] matrix (+ the rows' labels).  Every candidate of a slot
// grows from the same rows (model.py:324-342 draws them once per survivedTree call), so this replaces
// one gather per CANDIDATE inside the grower by one per SLOT, spread over the whole GPU.
// A warp owns 32 rows: it copies a chunk of words of each row verbatim into shared memory with
// coalesced loads (consecutive lanes = consecutive words of a row), then reads the tile back
// lane-per-row -- conflict-free, the tile pitch is an odd number of words -- and stores 32
// consecutive codes of one feature per instruction.  HBM-bound.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_WARPS = 8, GATHER_CHUNK_WORDS = 65;

template <typename CodeT>
__global__ void __launch_bounds__(GATHER_WARPS * 32) slot_gather_kernel(const CodeT* __restrict__ codes, int64_t stride,
                                                                         const uint8_t* __restrict__ labels, int n_features,
                                                                         const int32_t* __restrict__ idx_train, int n_train,
                                                                         int nt_pad, CodeT* __restrict__ slot_codes,
                                                                         uint8_t* __restrict__ slot_lab) {
    extern __shared__ __align__(16) uint32_t g_tile[];
    constexpr int PER = (int)(sizeof(uint32_t) / sizeof(CodeT));  // codes per word
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const int pitch_w = (int)((stride * (int64_t)sizeof(CodeT)) >> 2);
    const int cw_max = pitch_w < GATHER_CHUNK_WORDS ? pitch_w : GATHER_CHUNK_WORDS;
    const int tp = cw_max | 1;
    uint32_t* tile = g_tile + (size_t)warp * 32 * tp;
    const CodeT* tile_c = reinterpret_cast<const CodeT*>(tile);
    const uint32_t* codes_w = reinterpret_cast<const uint32_t*>(codes);
    const int i0 = (blockIdx.x * GATHER_WARPS + warp) * 32;
    if (i0 >= n_train) return;
    const int rows = n_train - i0 < 32 ? n_train - i0 : 32;
    int64_t my_off = 0;
    if (lane < rows) {
        const int64_t row = (int64_t)__ldg(&idx_train[(size_t)slot * n_train + i0 + lane]);
        my_off = row * pitch_w;
        slot_lab[(size_t)slot * nt_pad + i0 + lane] = __ldg(&labels[row]);
    }
    CodeT* out = slot_codes + (size_t)slot * n_features * nt_pad + i0 + lane;
    for (int w0 = 0; w0 < pitch_w; w0 += cw_max) {
        const int cw = pitch_w - w0 < cw_max ? pitch_w - w0 : cw_max;
        const int total = rows * cw;
        const int dq = 32 / cw, dr = 32 % cw;
        int t = lane / cw, r = lane % cw;
        for (int e = lane; e - lane < total; e += 32) {
            const int64_t off = __shfl_sync(FULL, my_off, t < rows ? t : 0);
            if (e < total) tile[t * tp + r] = __ldg(&codes_w[off + w0 + r]);
            t += dq;
            r += dr;
            if (r >= cw) {
                r -= cw;
                ++t;
            }
        }
        __syncwarp();
        const int f_end = (w0 + cw) * PER < n_features ? (w0 + cw) * PER : n_features;
        if (lane < rows) {
            const CodeT* mine = tile_c + (size_t)lane * tp * PER - (size_t)w0 * PER;
#pragma unroll 4
            for (int f = w0 * PER; f < f_end; ++f) out[(size_t)f * nt_pad] = mine[f];
        }
        __syncwarp();
    }
}

template <typename CodeT, typename PosT>
__global__ void __launch_bounds__(GROW_THREADS, RFMC_GROW_MIN_CTAS) grow_kernel(GrowArgs<CodeT, PosT> a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(16) uint32_t s_scratch[(GROW_WARPS + 1) * HIST_WORDS];
    uint32_t(*s_hist)[HIST_WORDS] = reinterpret_cast<uint32_t(*)[HIST_WORDS]>(s_scratch);  // per-warp histograms
    uint32_t* s_comb = s_scratch + GROW_WARPS * HIST_WORDS;                                 // the CTA's combined one
    __shared__ uint32_t s_wtot[GROW_WARPS];
    __shared__ uint32_t s_ticket[4];
    __shared__ uint32_t s_cnt[2][5];  // items per size class (tiny, small, medium, large, huge), this level / next level
    __shared__ uint32_t s_part[GROW_WARPS];
    __shared__ int s_correct;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const int nt = a.n_train;
    // list regions inside one work buffer: tiny items hold >= 2 rows, small >= 9, medium >= 33, large >= 129
    const uint32_t cap_t = (uint32_t)nt / 2 + 1, cap_s = (uint32_t)nt / (TINY_MAX + 1) + 1, cap_m = (uint32_t)nt / (SMALL_MAX + 1) + 1;
    const uint32_t cap_l = (uint32_t)nt / (MEDIUM_MAX + 1) + 1, cap_h = (uint32_t)nt / (HUGE_MIN + 1) + 1;
    const uint32_t reg0[5] = {0u, cap_t, cap_t + cap_s, cap_t + cap_s + cap_m, cap_t + cap_s + cap_m + cap_l};
    const uint32_t cap_all = cap_t + cap_s + cap_m + cap_l + cap_h;

    uint8_t* lab_s = s_dyn;  // shared-memory copy of the slot's labels (when it fits)
    PosT* pos = a.use_smem ? reinterpret_cast<PosT*>(s_dyn + (((size_t)a.nt_pad + 15) / 16) * 16)
                           : a.pos_g + (size_t)blockIdx.x * 2 * a.nt_pad;
    PosT* tmp = a.pos_g + (size_t)blockIdx.x * 2 * a.nt_pad + a.nt_pad;
    // split bitmap of the level + exclusive popcount prefix per word (only with the shared-memory layout)
    const uint32_t bit_words = (uint32_t)a.nt_pad / 32 + 1;
    uint32_t* s_bits = a.use_smem ? reinterpret_cast<uint32_t*>(s_dyn + smem_bits_offset<PosT>(a.nt_pad)) : nullptr;
    uint32_t* s_pref = a.use_smem ? s_bits + bit_words : nullptr;

    for (int cand = blockIdx.x; cand < a.n_cand; cand += gridDim.x) {
        CandCtx<CodeT, PosT> c;
        c.a = &a;
        c.pos = pos;
        c.tmp = tmp;
        c.resfin = a.resfin + (size_t)blockIdx.x * a.nt_pad;
        c.sbits = s_bits;
        c.nodes = a.nodes + (size_t)cand * a.max_nodes;
        c.counts = a.counts + (size_t)cand * a.max_nodes * a.n_classes;
        c.votes = a.votes + (size_t)cand * a.max_nodes;
        const int fo = a.feat_off[cand];
        c.flist = a.feat_list + fo;
        c.nf = a.feat_off[cand + 1] - fo;
        WorkItem* work = a.work + (size_t)blockIdx.x * 2 * cap_all;
        uint32_t* resna = a.resna + (size_t)blockIdx.x * 2 * a.nt_pad;
        uint32_t* rank = a.rank + (size_t)blockIdx.x * a.nt_pad;
        const int slot = a.cand_slot[cand];
        const int32_t* idxV = a.idx_val + (size_t)slot * a.n_val;
        const int C = a.n_classes;

        // ---- the slot's rows were gathered by slot_gather_kernel: only positions and labels to set up ----
        RFMC_TICK(0);
        c.cc = a.slot_codes + (size_t)slot * a.n_features * a.nt_pad;
        {
            const uint8_t* lab_g = a.slot_lab + (size_t)slot * a.nt_pad;
            c.lab = a.use_smem ? lab_s : lab_g;
            for (int i = tid; i < nt; i += GROW_THREADS) {
                if (a.use_smem) lab_s[i] = lab_g[i];
                pos[i] = (PosT)i;
            }
        }
        if (tid == 0) {
            const int cls = size_class(nt);
            s_cnt[0][0] = s_cnt[0][1] = s_cnt[0][2] = s_cnt[0][3] = s_cnt[0][4] = 0;
            s_cnt[0][cls] = 1;
            work[reg0[cls]] = WorkItem{0u, (uint32_t)nt, 0u, 0u};
            resna[0] = 0;
            s_correct = 0;
        }
        uint32_t width = 1, base = 0;
        int level = 0;
        __syncthreads();
        RFMC_TICK(1);

        while (width > 0) {
            const int cb = level & 1, nb = cb ^ 1;
            WorkItem* cur = work + (size_t)cb * cap_all;
            WorkItem* nxt = work + (size_t)nb * cap_all;
            uint32_t* resna_cur = resna + (size_t)cb * a.nt_pad;
            uint32_t* resna_nxt = resna + (size_t)nb * a.nt_pad;
            c.resna = resna_cur;
            c.base = base;
            c.level = level;
            const uint32_t n_tiny = s_cnt[cb][0], n_small = s_cnt[cb][1], n_medium = s_cnt[cb][2], n_large = s_cnt[cb][3],
                           n_huge = s_cnt[cb][4];
            if (tid < 5) {
                if (tid < 4) s_ticket[tid] = 0;
                s_cnt[nb][tid] = 0;
            }
            const uint32_t lvl_words = (width + 31) / 32;
            if (s_bits != nullptr)
                for (uint32_t i = tid; i < lvl_words; i += GROW_THREADS) s_bits[i] = 0u;
            __syncthreads();
#ifdef RFMC_GROW_TIMING
            if (blockIdx.x == 0 && cand == 0 && tid == 0 && 16 + 8 * level + 7 < 8192) {
                g_dbg[16 + 8 * level + 0] = clock64();
                g_dbg[16 + 8 * level + 4] = n_tiny;
                g_dbg[16 + 8 * level + 5] = n_small;
                g_dbg[16 + 8 * level + 6] = n_medium;
                g_dbg[16 + 8 * level + 7] = n_large + (n_huge << 16);
            }
#endif
            // ---- phase 1: huge nodes by the whole CTA, large and medium nodes by warps, small nodes by
            //      quarter warps, tiny nodes by threads ----
            for (uint32_t j = 0; j < n_huge; ++j)
                process_cta<CodeT, PosT>(c, cur + reg0[4] + j, s_hist, s_comb, s_part, tid);
            while (true) {
                uint32_t j = 0;
                if (lane == 0) j = atomicAdd(&s_ticket[3], 1u);
                j = __shfl_sync(FULL, j, 0);
                if (j >= n_large) break;
                process_warp<CodeT, PosT>(c, cur + reg0[3] + j, s_hist[warp], lane);
            }
            while (true) {
                uint32_t j = 0;
                if (lane == 0) j = atomicAdd(&s_ticket[2], 1u);
                j = __shfl_sync(FULL, j, 0);
                if (j >= n_medium) break;
                process_group<CodeT, PosT, 32>(c, cur + reg0[2] + j, 1u, s_hist[warp], lane);
            }
            while (true) {
                uint32_t j = 0;
                if (lane == 0) j = atomicAdd(&s_ticket[1], 4u);
                j = __shfl_sync(FULL, j, 0);
                if (j >= n_small) break;
                process_group<CodeT, PosT, 8>(c, cur + reg0[1] + j, n_small - j, s_hist[warp], lane);
            }
            while (true) {
                uint32_t j = 0;
                if (lane == 0) j = atomicAdd(&s_ticket[0], 32u);
                j = __shfl_sync(FULL, j, 0);
                if (j >= n_tiny) break;
                if (j + lane < n_tiny) process_tiny<CodeT, PosT>(c, cur + reg0[0] + j + lane);
            }
            __syncthreads();
            RFMC_TICK(16 + 8 * level + 1);
            // ---- phase 2a: split rank of every level position (block-wide exclusive scan) ----
            // each warp owns a contiguous range of positions and walks it 32 at a time (coalesced);
            // pass 1 counts, pass 2 re-reads the flags and writes the exclusive ranks
            uint32_t splits = 0;
            if (s_bits != nullptr) {
                // the nodes that split set their bit while they were processed: rank = set bits below the position
                if (warp == 0) {
                    uint32_t run = 0;
                    for (uint32_t w0 = 0; w0 < lvl_words; w0 += 32) {
                        const uint32_t v = (w0 + lane < lvl_words) ? (uint32_t)__popc(s_bits[w0 + lane]) : 0u;
                        uint32_t inc = v;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(FULL, inc, d);
                            if (lane >= d) inc += t;
                        }
                        if (w0 + lane < lvl_words) s_pref[w0 + lane] = run + inc - v;
                        run += __shfl_sync(FULL, inc, 31);
                    }
                    if (lane == 0) s_wtot[0] = run;
                }
                __syncthreads();
                splits = s_wtot[0];
            } else {
                const uint32_t per = ((width + GROW_WARPS * 32 - 1) / (GROW_WARPS * 32)) * 32;
                const uint32_t p0 = (uint32_t)warp * per, p1 = (p0 + per < width) ? p0 + per : width;
                uint32_t mine = 0;
                for (uint32_t p = p0; p < p1; p += 32)
                    mine += __popc(__ballot_sync(FULL, (p + lane < p1) && resna_cur[p + lane] > 0));
                if (lane == 0) s_wtot[warp] = mine;
                __syncthreads();
                uint32_t run = 0;
#pragma unroll
                for (int w8 = 0; w8 < GROW_WARPS; ++w8) {
                    const uint32_t t = s_wtot[w8];
                    if (w8 < warp) run += t;
                    splits += t;
                }
                for (uint32_t p = p0; p < p1; p += 32) {
                    const bool in = p + lane < p1;
                    const uint32_t bal = __ballot_sync(FULL, in && resna_cur[p + lane] > 0);
                    if (in) rank[p + lane] = run + __popc(bal & lt);
                    run += __popc(bal);
                }
                __syncthreads();
            }
            RFMC_TICK(16 + 8 * level + 2);
            // ---- phase 2b: child ids, finished leaves, next level's work lists ----
            const uint32_t base_next = base + width;
            const uint32_t n_items = n_tiny + n_small + n_medium + n_large + n_huge;
            for (uint32_t q = tid; q < n_items; q += GROW_THREADS) {
                uint32_t qq = q;
                int rc = 0;
                if (qq >= n_tiny) { qq -= n_tiny; rc = 1;
                    if (qq >= n_small) { qq -= n_small; rc = 2;
                        if (qq >= n_medium) { qq -= n_medium; rc = 3;
                            if (qq >= n_large) { qq -= n_large; rc = 4; } } } }
                const WorkItem w = cur[reg0[rc] + qq];
                uint32_t s;
                if (s_bits != nullptr) {
                    const uint32_t bw = s_bits[w.lpos >> 5], bit = 1u << (w.lpos & 31);
                    if (!(bw & bit)) continue;
                    s = s_pref[w.lpos >> 5] + (uint32_t)__popc(bw & (bit - 1u));
                } else {
                    if (resna_cur[w.lpos] == 0) continue;
                    s = rank[w.lpos];
                }
                const uint32_t na = resna_cur[w.lpos];
                const uint32_t fin = c.resfin[w.lpos];
                c.nodes[base + w.lpos].child = (int32_t)(base_next + 2 * s);
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const uint32_t lo = side ? w.lo + na : w.lo, hi = side ? w.hi : w.lo + na;
                    const uint32_t sz = hi - lo, cpos = 2 * s + side, f8 = (fin >> (8 * side)) & 0xFFu;
                    if (s_bits == nullptr) resna_nxt[cpos] = 0;  // positions that never become work items must rank as leaves
                    if (f8 != NOT_FINAL) {  // finished leaf: all rows of one class
                        const uint32_t cn = base_next + cpos;
                        write_leaf(&c.nodes[cn]);
                        for (int k = 0; k < C; ++k) c.counts[(size_t)cn * C + k] = (k == (int)f8) ? (int32_t)sz : 0;
                        c.votes[cn] = (uint8_t)f8;
                    } else {
                        const int cls = size_class((int)sz);
                        const uint32_t at = atomicAdd(&s_cnt[nb][cls], 1u);
                        nxt[reg0[cls] + at] = WorkItem{lo, hi, cpos, w.off};
                    }
                }
            }
            base = base_next;
            width = 2 * splits;
            __syncthreads();
            RFMC_TICK(16 + 8 * level + 3);
            ++level;
        }
        RFMC_TICK(2);
        const uint32_t n_nodes = base;

        // ---- exact code thresholds of the splits that deferred their table search (thread per node) ----
        for (uint32_t i = tid; i < n_nodes; i += GROW_THREADS) {
            const uint32_t fl = c.nodes[i].flags;
            if (fl & FLAG_PENDING) {
                const int32_t f = c.nodes[i].feat;
                const double* tab = a.tables + a.tab_off[f];
                c.nodes[i].thr = count_less_scalar(tab, c.nodes[i].thr, (uint32_t)a.col_nuniq[f], c.nodes[i].sval) + 1u;
                c.nodes[i].flags = fl & ~FLAG_PENDING;
            }
        }
        __syncthreads();
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
        RFMC_TICK(3);
#ifdef RFMC_GROW_TIMING
        if (blockIdx.x == 0 && cand == 0 && tid == 0) g_dbg[4] = (unsigned long long)level;
#endif
        if (tid == 0) {
            a.nnodes[cand] = (int32_t)n_nodes;
            a.correct[cand] = s_correct;
        }
        __syncthreads();
    }
    (void)lt;
}

template <typename CodeT, typename PosT>
static size_t smem_bytes(int nt_pad) {
    return smem_bits_offset<PosT>(nt_pad) + 2 * ((size_t)nt_pad / 32 + 1) * sizeof(uint32_t);
}
constexpr size_t SMEM_BUDGET = 100 * 1024;  // per CTA, keeps >= 2 CTAs per SM

template <typename CodeT, typename PosT>
static int launch_grow_t(rfmc_batch* b, cudaStream_t stream) {
    rfmc_dataset* ds = b->ds;
    GrowArgs<CodeT, PosT> a;
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
    a.slot_codes = (const CodeT*)b->d_cc;
    a.slot_lab = b->d_lab;
    a.n_features = ds->n_features;
    a.pos_g = (PosT*)b->d_pos;
    a.work = b->d_work;
    a.resna = b->d_resna;
    a.rank = b->d_rank;
    a.resfin = b->d_resfin;
    const size_t dyn = smem_bytes<CodeT, PosT>(b->nt_pad);
    a.use_smem = dyn <= SMEM_BUDGET;
    a.nodes = b->d_nodes;
    a.counts = b->d_counts;
    a.votes = b->d_votes;
    a.nnodes = b->d_nnodes;
    a.correct = b->d_correct;
    {   // one pass over every slot's bootstrap rows: [row][feature] in HBM -> [slot][feature][row]
        const int pitch_w = (int)((ds->stride * (int64_t)sizeof(CodeT)) >> 2);
        const int cw = pitch_w < GATHER_CHUNK_WORDS ? pitch_w : GATHER_CHUNK_WORDS;
        const size_t gsm = (size_t)GATHER_WARPS * 32 * (cw | 1) * sizeof(uint32_t);
        auto gk = slot_gather_kernel<CodeT>;
        RFMC_CUDA(cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
        dim3 grid((unsigned)((b->n_train + GATHER_WARPS * 32 - 1) / (GATHER_WARPS * 32)), (unsigned)b->n_slots);
        if (b->timing) RFMC_CUDA(cudaEventRecord(b->ev[0], stream));
        gk<<<grid, GATHER_WARPS * 32, gsm, stream>>>((const CodeT*)ds->d_codes, ds->stride, ds->d_labels, ds->n_features,
                                                      b->d_idx_train, b->n_train, b->nt_pad, (CodeT*)b->d_cc, b->d_lab);
        RFMC_CUDA(cudaGetLastError());
    }
    auto k = grow_kernel<CodeT, PosT>;
    const size_t smem = a.use_smem ? dyn : 0;
    if (smem > 0) RFMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
    if (b->timing) RFMC_CUDA(cudaEventRecord(b->ev[1], stream));
    k<<<b->grid, GROW_THREADS, smem, stream>>>(a);
    RFMC_CUDA(cudaGetLastError());
    b->launches = 2;
    if (b->timing) RFMC_CUDA(cudaEventRecord(b->ev[2], stream));
    return 0;
}

int launch_grow(rfmc_batch* b, cudaStream_t stream) {
    const bool p16 = b->n_train <= 65535;
    switch (b->ds->code_width) {
        case 1: return p16 ? launch_grow_t<uint8_t, uint16_t>(b, stream) : launch_grow_t<uint8_t, uint32_t>(b, stream);
        case 2: return p16 ? launch_grow_t<uint16_t, uint16_t>(b, stream) : launch_grow_t<uint16_t, uint32_t>(b, stream);
        case 4: return p16 ? launch_grow_t<uint32_t, uint16_t>(b, stream) : launch_grow_t<uint32_t, uint32_t>(b, stream);
    }
    set_error("unsupported code width");
    return 1;
}

template <typename CodeT, typename PosT>
static int occupancy_t(int nt_pad) {
    int per_sm = 0;
    const size_t dyn = smem_bytes<CodeT, PosT>(nt_pad);
    const size_t smem = dyn <= SMEM_BUDGET ? dyn : 0;
    auto k = grow_kernel<CodeT, PosT>;
    if (smem > 0) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, GROW_THREADS, smem) != cudaSuccess) per_sm = 1;
    return per_sm < 1 ? 1 : per_sm;
}

int debug_read(unsigned long long* out, int n) {
#ifdef RFMC_GROW_TIMING
    if (n > 8192) n = 8192;
    return cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : 1;
#else
    (void)out;
    (void)n;
    return 1;
#endif
}

int grow_grid_size(const rfmc_dataset* ds, int n_cand, int n_train) {
    const int nt_pad = (n_train + 15) / 16 * 16;
    const bool p16 = n_train <= 65535;
    int per_sm = 1;
    switch (ds->code_width) {
        case 1: per_sm = p16 ? occupancy_t<uint8_t, uint16_t>(nt_pad) : occupancy_t<uint8_t, uint32_t>(nt_pad); break;
        case 2: per_sm = p16 ? occupancy_t<uint16_t, uint16_t>(nt_pad) : occupancy_t<uint16_t, uint32_t>(nt_pad); break;
        default: per_sm = p16 ? occupancy_t<uint32_t, uint16_t>(nt_pad) : occupancy_t<uint32_t, uint32_t>(nt_pad); break;
    }
    if (const char* e = getenv("RFMC_GROW_CTAS_PER_SM")) {  // experiment knob: fewer resident candidates
        const int want = atoi(e);
        if (want >= 1 && want < per_sm) per_sm = want;
    }
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
