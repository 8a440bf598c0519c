// Forest traversal + voting kernel for sm_100a.
//
// Replaces DecisionTreeMC._useTree/useTree (tree.py:155-212), BaseRandomForestMC.useForest
// (forest.py:204-233), maxProbClas (forest.py:200-202) and the row loops of testForest /
// testForestProbs (forest.py:235-236, 257-258).
//
// Mapping: one thread per row, 128 rows per CTA.  The CTA's rows are copied verbatim from the
// row-major HBM code matrix into shared memory (row pitch = odd number of 32-bit words, so the
// per-thread row reads are bank-conflict-free whatever column each thread asks for).  Every thread
// walks the trees IN FOREST ORDER and accumulates in fp64 with separate multiply and add, which is
// the reference's summation order (forest.py:209-211, 216-218, 226-227): the probabilities come
// out bit-identical, not merely within tolerance.  Node records are 16-byte loads; all threads of
// a CTA are inside the same tree at the same time, so node traffic is L1/L2 broadcast traffic.
//
// Missing values (tree.py:166-181): a NaN / unseen cell is code 0 and fails both tests -> '<'.
// A column absent from the frame takes '>=' and switches the leaf payload to the renormalised
// variant the host precomputed (tree.py:201-210).
#include <cstdlib>

#include "common.cuh"

namespace rfmc {

constexpr int PRED_THREADS = 128;
#ifndef RFMC_PRED_ILP
#define RFMC_PRED_ILP 4
#endif
constexpr int PRED_ILP = RFMC_PRED_ILP;

struct PredictArgs {
    const void* codes;
    int64_t n_rows;
    int64_t stride;  // elements
    int32_t n_features;
    const rfmc_pnode* nodes;
    const int64_t* root;
    const uint4* blocks;     // blocked layout (nullptr: walk `nodes`)
    const uint32_t* root_b;  // block of every tree's root
    const uint2* tops;       // every tree's top `top_levels` levels as an implicit heap of 8-byte records
    int32_t top_levels;      // even, >= 2 (0: no tops)
    // compact layout of a forest of small trees (nullptr: not eligible), see build_compact in abi.cu
    const uint2* crec;
    const uint32_t* crec4;  // 4-byte records (nullptr: not eligible)
    const uint32_t* cpay;
    const uint32_t* ctree;
    const uint8_t* cdepth;
    const uint32_t* cstage;
    int32_t n_stages;
    int32_t n_trees;
    const uint8_t* vote;
    const uint8_t* vote_absent;
    const double* probs;
    const double* probs_absent;
    const double* scores;
    double score_fsum;
    int32_t n_classes;
    const uint8_t* absent;
    double* probs_out;
    int32_t* labels_out;
};

// One 32-byte block in ONE request (LDG.E.256 on sm_100a): the vote kernel sits at the L1TEX
// request-throughput limit (97 % in ncu), so a block fetched as two 16-byte loads costs twice.
__device__ __forceinline__ void ldg256(const uint4* p, uint4& lo, uint4& hi) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
        : "l"(p));
}

// MODE: 0 hard, 1 hard+weighted, 2 soft, 3 soft+weighted
template <typename CodeT, int CMAX, int MODE, bool HAS_ABSENT, bool TILE>
__global__ void __launch_bounds__(PRED_THREADS) predict_kernel(PredictArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CodeT* tile = reinterpret_cast<CodeT*>(smem_raw);
    uint8_t* s_absent = smem_raw + (TILE ? (size_t)PRED_THREADS * a.stride * sizeof(CodeT) : 0);

    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * PRED_THREADS;
    const int rows = (int)((a.n_rows - row0) < PRED_THREADS ? (a.n_rows - row0) : PRED_THREADS);
    const CodeT* gcodes = reinterpret_cast<const CodeT*>(a.codes) + row0 * a.stride;

    if (TILE) {
        // rows*stride*sizeof(CodeT) is a multiple of 4 by the layout contract
        const uint32_t* src = reinterpret_cast<const uint32_t*>(gcodes);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tile);
        const int words = (int)((size_t)rows * a.stride * sizeof(CodeT) / 4);
        for (int i = tid; i < words; i += PRED_THREADS) dst[i] = __ldg(&src[i]);
    }
    if (HAS_ABSENT) {
        for (int i = tid; i < a.n_features; i += PRED_THREADS) s_absent[i] = a.absent[i];
    }
    __syncthreads();
    if (tid >= rows) return;

    const CodeT* myrow = TILE ? (tile + (size_t)tid * a.stride) : (gcodes + (size_t)tid * a.stride);
    const int C = a.n_classes;

    double acc[CMAX];
    uint32_t cnt[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        acc[c] = 0.0;
        cnt[c] = 0;
    }

    // PRED_ILP trees are walked concurrently (their node loads are independent, so PRED_ILP loads
    // are in flight per thread instead of one); the leaves are then accumulated in forest order.
    //
    // Blocked layout (when the forest's codes fit 16 bits): a 32-byte block holds a node and both
    // its children as 8-byte records {feat | cat<<14 | leaf<<15, thr (u16), next (u32)}, so ONE
    // dependent 32-byte L2 sector serves two tree levels; a child's `next` is the block of its own
    // '>=' child ('<' child = next + 1).  The kernel is bound by those dependent sector fetches
    // (L1 hit rate ~3 %, L2 hit rate 99 %), so halving them is the lever.
    for (int t0 = 0; t0 < a.n_trees; t0 += PRED_ILP) {
        uint32_t node[PRED_ILP], pay[PRED_ILP];
        bool done[PRED_ILP], seen[PRED_ILP];
#pragma unroll
        for (int k = 0; k < PRED_ILP; ++k) {
            const bool valid = t0 + k < a.n_trees;
            node[k] = valid ? (a.blocks ? __ldg(&a.root_b[t0 + k]) : (uint32_t)__ldg(&a.root[t0 + k])) : 0u;
            done[k] = !valid;
            seen[k] = false;
            pay[k] = 0u;
        }
        bool any = true;
        if (a.blocks != nullptr) {
            uint4 b0[PRED_ILP], b1[PRED_ILP];
            while (any) {
                any = false;
#pragma unroll
                for (int k = 0; k < PRED_ILP; ++k) {
                    if (!done[k]) ldg256(a.blocks + 2 * (size_t)node[k], b0[k], b1[k]);
                }
#pragma unroll
                for (int k = 0; k < PRED_ILP; ++k) {
                    if (done[k]) continue;
                    const uint32_t p_lo = b0[k].x;
                    if (p_lo & 0x8000u) {
                        pay[k] = b0[k].y;
                        done[k] = true;
                        continue;
                    }
                    uint32_t f = p_lo & 0x3FFFu;
                    uint32_t cv = (uint32_t)myrow[f];
                    bool go = (p_lo & 0x4000u) ? (cv == (p_lo >> 16)) : (cv >= (p_lo >> 16));
                    if (HAS_ABSENT) {
                        const bool ab = s_absent[f] != 0;
                        go = go || ab;
                        seen[k] = seen[k] || ab;
                    }
                    const uint32_t x_lo = go ? b0[k].z : b1[k].x, x_hi = go ? b0[k].w : b1[k].y;
                    if (x_lo & 0x8000u) {
                        pay[k] = x_hi;
                        done[k] = true;
                        continue;
                    }
                    f = x_lo & 0x3FFFu;
                    cv = (uint32_t)myrow[f];
                    go = (x_lo & 0x4000u) ? (cv == (x_lo >> 16)) : (cv >= (x_lo >> 16));
                    if (HAS_ABSENT) {
                        const bool ab = s_absent[f] != 0;
                        go = go || ab;
                        seen[k] = seen[k] || ab;
                    }
                    node[k] = x_hi + (go ? 0u : 1u);
                    any = true;
                }
            }
        } else {
            uint4 nd[PRED_ILP];
            while (any) {
                any = false;
#pragma unroll
                for (int k = 0; k < PRED_ILP; ++k)
                    if (!done[k]) nd[k] = __ldg(reinterpret_cast<const uint4*>(a.nodes) + node[k]);
#pragma unroll
                for (int k = 0; k < PRED_ILP; ++k) {
                    if (done[k]) continue;
                    if (nd[k].x & RFMC_PNODE_LEAF) {
                        pay[k] = nd[k].y;
                        done[k] = true;
                        continue;
                    }
                    const uint32_t f = nd[k].x & 0xFFFFu;
                    const uint32_t cv = (uint32_t)myrow[f];
                    bool go = (nd[k].x & RFMC_PNODE_CATEGORICAL) ? (cv == nd[k].y) : (cv >= nd[k].y);
                    if (HAS_ABSENT) {
                        const bool ab = s_absent[f] != 0;
                        go = go || ab;
                        seen[k] = seen[k] || ab;
                    }
                    node[k] = nd[k].z + (go ? 0u : 1u);
                    any = true;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PRED_ILP; ++k) {
            const int t = t0 + k;
            if (t >= a.n_trees) break;
            const uint32_t payload = pay[k];
            if (MODE == 0 || MODE == 1) {
                const int v = (HAS_ABSENT && seen[k]) ? a.vote_absent[payload] : a.vote[payload];
                if (MODE == 0) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) cnt[c] += (c == v) ? 1u : 0u;
                } else {
                    const double s = __ldg(&a.scores[t]);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) acc[c] = __dadd_rn(acc[c], (c == v) ? s : 0.0);
                }
            } else {
                const double* p = ((HAS_ABSENT && seen[k]) ? a.probs_absent : a.probs) + (size_t)payload * C;
                if (MODE == 2) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c)
                        if (c < C) acc[c] = __dadd_rn(acc[c], __ldg(&p[c]));
                } else {
                    const double s = __ldg(&a.scores[t]);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c)
                        if (c < C) acc[c] = __dadd_rn(acc[c], __dmul_rn(__ldg(&p[c]), s));
                }
            }
        }
    }

    const double denom = (MODE == 0 || MODE == 2) ? (double)a.n_trees : a.score_fsum;
    const int64_t row = row0 + tid;
    int best = 0;
    double bestv = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const double num = (MODE == 0) ? (double)cnt[c] : acc[c];
            const double v = __ddiv_rn(num, denom);
            if (a.probs_out) a.probs_out[row * C + c] = v;
            if (c == 0 || v > bestv) {  // first maximum in class_vals order (forest.py:200-202)
                bestv = v;
                best = c;
            }
        }
    }
    if (a.labels_out) a.labels_out[row] = best;
}

// ---------------------------------------------------------------------------------------------
// Vote kernel with the tree tops in shared memory (EXPERIMENT, opt-in with RFMC_PRED_TOPS=1: measured
// 0.64x / 0.99x of the kernel above on the C2b / C5 forests, see profiles/r2_vote_staged_tops.md -- the walk
// is bound by the two dependent loads per level and the issue slots around them, not by where the node
// records live).
//
// The kernel above asks L2 (through the L1 tags) for one 32-byte sector per lane every two tree levels:
// that request stream is its limit (ncu: l1tex throughput 91 %), not HBM.  Here the first `top_levels`
// levels of every tree are laid out by the host as an implicit heap of 8-byte records (same record as the
// blocked layout: {feat | cat << 14 | leaf << 15 | thr << 16, payload or next block}); a CTA of 256 rows
// stages the heaps of four trees at a time with 16-byte cp.async copies (double-buffered: the copy of the
// next four overlaps the walk of these four), every thread walks the top levels of its four trees out of
// shared memory -- children of heap slot i are 2i+1 / 2i+2, no pointer to chase -- and only what lies below
// the staged levels is fetched from global memory, entering the blocked layout at the record's `next`.  A
// tree of the default configuration (~70 nodes, depth 7-8) is walked without a single global request; a
// tree grown to purity on 8,000 rows (depth 16-17) keeps its last ~4 levels global.  Staging costs
// heap bytes / 256 rows per (row, tree): 32 B for ten levels, against ~7 x 32 B of sector requests.
// Accumulation is unchanged: leaves are consumed in forest order, fp64, multiply and add separate.
// ---------------------------------------------------------------------------------------------
constexpr int TOP_THREADS = 256;
constexpr int TOP_GROUP = 4;  // trees staged and walked together

template <typename CodeT, int CMAX, int MODE, bool HAS_ABSENT>
__global__ void __launch_bounds__(TOP_THREADS) predict_top_kernel(PredictArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int heap_n = (1 << a.top_levels) - 1;
    const int heap_pad = (heap_n + 2) & ~1;  // records per staged tree, 16-byte multiple
    uint2* stage = reinterpret_cast<uint2*>(smem_raw);  // [2][TOP_GROUP][heap_pad]
    CodeT* tile = reinterpret_cast<CodeT*>(smem_raw + (size_t)2 * TOP_GROUP * heap_pad * sizeof(uint2));
    uint8_t* s_absent = reinterpret_cast<uint8_t*>(tile) + (size_t)TOP_THREADS * a.stride * sizeof(CodeT);

    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * TOP_THREADS;
    const int rows = (int)((a.n_rows - row0) < TOP_THREADS ? (a.n_rows - row0) : TOP_THREADS);
    const CodeT* gcodes = reinterpret_cast<const CodeT*>(a.codes) + row0 * a.stride;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(gcodes);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tile);
        const int words = (int)((size_t)rows * a.stride * sizeof(CodeT) / 4);
        for (int i = tid; i < words; i += TOP_THREADS) dst[i] = __ldg(&src[i]);
    }
    if (HAS_ABSENT)
        for (int i = tid; i < a.n_features; i += TOP_THREADS) s_absent[i] = a.absent[i];

    auto stage_group = [&](int t0, int buf) {  // heaps of trees t0 .. t0+3 -> stage[buf]
        const int n_tr = a.n_trees - t0 < TOP_GROUP ? a.n_trees - t0 : TOP_GROUP;
        const int per_tree = heap_pad / 2;  // 16-byte pieces
        const uint4* src = reinterpret_cast<const uint4*>(a.tops + (size_t)t0 * heap_pad);
        uint4* dst = reinterpret_cast<uint4*>(stage + (size_t)buf * TOP_GROUP * heap_pad);
        for (int i = tid; i < n_tr * per_tree; i += TOP_THREADS) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + i));
        }
        asm volatile("cp.async.commit_group;");
    };

    const bool live = tid < rows;
    const CodeT* myrow = tile + (size_t)(live ? tid : 0) * a.stride;
    const int C = a.n_classes;
    double acc[CMAX];
    uint32_t cnt[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        acc[c] = 0.0;
        cnt[c] = 0;
    }
    stage_group(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < a.n_trees; t0 += TOP_GROUP, buf ^= 1) {
        if (t0 + TOP_GROUP < a.n_trees) {
            stage_group(t0 + TOP_GROUP, buf ^ 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();  // this group's heaps (and, the first time, the row tile) are in shared memory
        const uint2* heaps = stage + (size_t)buf * TOP_GROUP * heap_pad;
        uint32_t node[TOP_GROUP], pay[TOP_GROUP], idx[TOP_GROUP];
        bool done[TOP_GROUP], seen[TOP_GROUP];
#pragma unroll
        for (int k = 0; k < TOP_GROUP; ++k) {
            done[k] = !(t0 + k < a.n_trees);
            seen[k] = false;
            pay[k] = 0u;
            idx[k] = 0u;
            node[k] = 0u;
        }
        // the staged levels: four independent chains of shared-memory loads per thread
        for (int d = 0; d < a.top_levels; ++d) {
            const bool last = d + 1 == a.top_levels;
#pragma unroll
            for (int k = 0; k < TOP_GROUP; ++k) {
                if (done[k]) continue;
                const uint2 rec = heaps[(size_t)k * heap_pad + idx[k]];
                if (rec.x & 0x8000u) {
                    pay[k] = rec.y;
                    done[k] = true;
                    continue;
                }
                const uint32_t f = rec.x & 0x3FFFu;
                const uint32_t cv = (uint32_t)myrow[f];
                bool go = (rec.x & 0x4000u) ? (cv == (rec.x >> 16)) : (cv >= (rec.x >> 16));
                if (HAS_ABSENT) {
                    const bool ab = s_absent[f] != 0;
                    go = go || ab;
                    seen[k] = seen[k] || ab;
                }
                if (last)
                    node[k] = rec.y + (go ? 0u : 1u);  // below the staged levels: the child's block in global memory
                else
                    idx[k] = 2u * idx[k] + (go ? 1u : 2u);
            }
        }
        // what is left of the deep trees: the blocked layout, two levels per 32-byte request
        bool any = false;
#pragma unroll
        for (int k = 0; k < TOP_GROUP; ++k) any = any || !done[k];
        uint4 b0[TOP_GROUP], b1[TOP_GROUP];
        while (any) {
            any = false;
#pragma unroll
            for (int k = 0; k < TOP_GROUP; ++k)
                if (!done[k]) ldg256(a.blocks + 2 * (size_t)node[k], b0[k], b1[k]);
#pragma unroll
            for (int k = 0; k < TOP_GROUP; ++k) {
                if (done[k]) continue;
                const uint32_t p_lo = b0[k].x;
                if (p_lo & 0x8000u) {
                    pay[k] = b0[k].y;
                    done[k] = true;
                    continue;
                }
                uint32_t f = p_lo & 0x3FFFu;
                uint32_t cv = (uint32_t)myrow[f];
                bool go = (p_lo & 0x4000u) ? (cv == (p_lo >> 16)) : (cv >= (p_lo >> 16));
                if (HAS_ABSENT) {
                    const bool ab = s_absent[f] != 0;
                    go = go || ab;
                    seen[k] = seen[k] || ab;
                }
                const uint32_t x_lo = go ? b0[k].z : b1[k].x, x_hi = go ? b0[k].w : b1[k].y;
                if (x_lo & 0x8000u) {
                    pay[k] = x_hi;
                    done[k] = true;
                    continue;
                }
                f = x_lo & 0x3FFFu;
                cv = (uint32_t)myrow[f];
                go = (x_lo & 0x4000u) ? (cv == (x_lo >> 16)) : (cv >= (x_lo >> 16));
                if (HAS_ABSENT) {
                    const bool ab = s_absent[f] != 0;
                    go = go || ab;
                    seen[k] = seen[k] || ab;
                }
                node[k] = x_hi + (go ? 0u : 1u);
                any = true;
            }
        }
#pragma unroll
        for (int k = 0; k < TOP_GROUP; ++k) {
            const int t = t0 + k;
            if (t >= a.n_trees) break;
            const uint32_t payload = pay[k];
            if (MODE == 0 || MODE == 1) {
                const int v = (HAS_ABSENT && seen[k]) ? a.vote_absent[payload] : a.vote[payload];
                if (MODE == 0) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) cnt[c] += (c == v) ? 1u : 0u;
                } else {
                    const double sc = __ldg(&a.scores[t]);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) acc[c] = __dadd_rn(acc[c], (c == v) ? sc : 0.0);
                }
            } else {
                const double* p = ((HAS_ABSENT && seen[k]) ? a.probs_absent : a.probs) + (size_t)payload * C;
                if (MODE == 2) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c)
                        if (c < C) acc[c] = __dadd_rn(acc[c], __ldg(&p[c]));
                } else {
                    const double sc = __ldg(&a.scores[t]);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c)
                        if (c < C) acc[c] = __dadd_rn(acc[c], __dmul_rn(__ldg(&p[c]), sc));
                }
            }
        }
        __syncthreads();  // everybody is done with stage[buf] before the copy after next overwrites it
    }
    if (!live) return;
    const double denom = (MODE == 0 || MODE == 2) ? (double)a.n_trees : a.score_fsum;
    const int64_t row = row0 + tid;
    int best = 0;
    double bestv = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const double num = (MODE == 0) ? (double)cnt[c] : acc[c];
            const double v = __ddiv_rn(num, denom);
            if (a.probs_out) a.probs_out[row * C + c] = v;
            if (c == 0 || v > bestv) {  // first maximum in class_vals order (forest.py:200-202)
                bestv = v;
                best = c;
            }
        }
    }
    if (a.labels_out) a.labels_out[row] = best;
}

// ---------------------------------------------------------------------------------------------
// Vote kernel for forests of SMALL trees (the reference's default configuration: 10 + 10 rows per class grow
// trees of a few dozen nodes; BASELINE config 5 votes 4,096 of them).  The generic kernel above spends ~27
// issue slots per tree level on leaf tests, node-kind selects and per-walk bookkeeping.  Here the forest is
// laid out so that a level is the same handful of instructions for every node kind (abi.cu, build_compact):
//     rec = stage[idx];  go = (row[rec.feat] - rec.lo) <= rec.span;  idx = rec.next + !go
// a leaf points at itself, so every row of the CTA walks a tree for exactly `depth[tree]` steps: no leaf test,
// no divergence.  Whole trees (up to 1,024 records per stage) are copied to shared memory by the CTA and
// walked from there, SMALL_ILP trees in flight per thread; the leaves are consumed in forest order with the
// reference's fp64 multiply-then-add, exactly as above.
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 256;
constexpr int SMALL_STAGE = 1024;  // records per stage (SMALL_STAGE_NODES in abi.cu)
#ifndef RFMC_SMALL_ILP
#define RFMC_SMALL_ILP 2  // trees in flight per thread: measured 1 / 2 / 4 / 8 -> 24.3 / 22.5 / 23.3 / 23.6 ms (C5, weighted soft)
#endif

__device__ __forceinline__ uint32_t lds_code(uint32_t addr, uint8_t) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_code(uint32_t addr, uint16_t) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t lds_code(uint32_t addr, uint32_t) {  // instantiated, never launched (codes <= 16 bits only)
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <typename CodeT, int CMAX, int MODE, bool REC4, int SMALL_ILP = RFMC_SMALL_ILP>
__global__ void __launch_bounds__(SMALL_THREADS) predict_small_kernel(PredictArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t REC_BYTES = REC4 ? 4u : 8u;
    uint32_t* s_pay = reinterpret_cast<uint32_t*>(smem_raw + (size_t)SMALL_STAGE * REC_BYTES);
    CodeT* tile = reinterpret_cast<CodeT*>(smem_raw + (size_t)SMALL_STAGE * (REC_BYTES + sizeof(uint32_t)));
    const uint32_t rec_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * SMALL_THREADS;
    const int rows = (int)((a.n_rows - row0) < SMALL_THREADS ? (a.n_rows - row0) : SMALL_THREADS);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(reinterpret_cast<const CodeT*>(a.codes) + row0 * a.stride);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tile);
        const int words = (int)((size_t)rows * a.stride * sizeof(CodeT) / 4);
        for (int i = tid; i < words; i += SMALL_THREADS) dst[i] = __ldg(&src[i]);
    }
    const bool live = tid < rows;
    const uint32_t rowaddr = (uint32_t)__cvta_generic_to_shared(tile + (size_t)(live ? tid : 0) * a.stride);
    const int C = a.n_classes;
    double acc[CMAX];
    uint32_t cnt[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        acc[c] = 0.0;
        cnt[c] = 0;
    }

    for (int s = 0; s < a.n_stages; ++s) {
        const int t0 = (int)__ldg(&a.cstage[s]), t1 = (int)__ldg(&a.cstage[s + 1]);
        const uint32_t n0 = __ldg(&a.ctree[t0]), n1 = __ldg(&a.ctree[t1]);
        __syncthreads();  // the previous stage's walks (and, the first time, the row tile's stores) are done
        for (uint32_t i = tid; i < n1 - n0; i += SMALL_THREADS) {
            if (REC4)
                reinterpret_cast<uint32_t*>(smem_raw)[i] = __ldg(&a.crec4[n0 + i]);
            else
                reinterpret_cast<uint2*>(smem_raw)[i] = __ldg(&a.crec[n0 + i]);
            const uint32_t payload = __ldg(&a.cpay[n0 + i]);
            // hard votes: the leaf's class is looked up once per staged record instead of once per walk
            s_pay[i] = (MODE == 0 || MODE == 1) ? (uint32_t)__ldg(&a.vote[payload]) : payload;
        }
        __syncthreads();
        for (int tg = t0; tg < t1; tg += SMALL_ILP) {
            uint32_t tb[SMALL_ILP], at[SMALL_ILP];  // shared-memory address of the tree's first record / of the current one
            int levels = 0;
#pragma unroll
            for (int k = 0; k < SMALL_ILP; ++k) {
                const int t = tg + k < t1 ? tg + k : t1 - 1;  // a short last group walks its last tree again (result unused)
                tb[k] = rec_base + (__ldg(&a.ctree[t]) - n0) * REC_BYTES;
                at[k] = tb[k];
                const int d = (int)__ldg(&a.cdepth[t]);
                levels = d > levels ? d : levels;
            }
#pragma unroll 1
            for (int l = 0; l < levels; ++l) {
#pragma unroll
                for (int k = 0; k < SMALL_ILP; ++k) {
                    if (REC4) {
                        uint32_t r;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(at[k]));
                        const uint32_t cv = lds_code(rowaddr + (r & 0xFFu) * (uint32_t)sizeof(CodeT), CodeT());
                        at[k] = at[k] + ((r >> 8) & 0x7Fu) * 4u + ((cv - (r >> 16)) > ((r & 0x8000u) << 1) ? 4u : 0u);
                    } else {
                        uint32_t x, y;
                        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(at[k]));
                        const uint32_t cv = lds_code(rowaddr + (x & 0xFFFFu) * (uint32_t)sizeof(CodeT), CodeT());
                        at[k] = tb[k] + (y >> 16) * REC_BYTES + ((cv - (x >> 16)) > (y & 0xFFFFu) ? REC_BYTES : 0u);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < SMALL_ILP; ++k) {
                const int t = tg + k;
                if (t >= t1) break;
                const uint32_t payload = s_pay[(at[k] - rec_base) / REC_BYTES];
                if (MODE == 0 || MODE == 1) {
                    const int v = (int)payload;  // staged as the class already
                    if (MODE == 0) {
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) cnt[c] += (c == v) ? 1u : 0u;
                    } else {
                        const double sc = __ldg(&a.scores[t]);
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) acc[c] = __dadd_rn(acc[c], (c == v) ? sc : 0.0);
                    }
                } else {
                    const double* p = a.probs + (size_t)payload * C;
                    double pv[CMAX];
                    if (CMAX == 4 && C == 4) {  // a payload's four probabilities are one aligned 32-byte row
                        const double2 lo2 = __ldg(reinterpret_cast<const double2*>(p)), hi2 = __ldg(reinterpret_cast<const double2*>(p) + 1);
                        pv[0] = lo2.x, pv[1] = lo2.y, pv[2 % CMAX] = hi2.x, pv[3 % CMAX] = hi2.y;
                    } else {
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) pv[c] = c < C ? __ldg(&p[c]) : 0.0;
                    }
                    if (MODE == 2) {
#pragma unroll
                        for (int c = 0; c < CMAX; ++c)
                            if (c < C) acc[c] = __dadd_rn(acc[c], pv[c]);
                    } else {
                        const double sc = __ldg(&a.scores[t]);
#pragma unroll
                        for (int c = 0; c < CMAX; ++c)
                            if (c < C) acc[c] = __dadd_rn(acc[c], __dmul_rn(pv[c], sc));
                    }
                }
            }
        }
    }
    if (!live) return;
    const double denom = (MODE == 0 || MODE == 2) ? (double)a.n_trees : a.score_fsum;
    const int64_t row = row0 + tid;
    int best = 0;
    double bestv = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const double num = (MODE == 0) ? (double)cnt[c] : acc[c];
            const double v = __ddiv_rn(num, denom);
            if (a.probs_out) a.probs_out[row * C + c] = v;
            if (c == 0 || v > bestv) {  // first maximum in class_vals order (forest.py:200-202)
                bestv = v;
                best = c;
            }
        }
    }
    if (a.labels_out) a.labels_out[row] = best;
}

template <typename CodeT, int CMAX, int MODE, bool HAS_ABSENT>
static int launch_tile(const PredictArgs& a, cudaStream_t stream) {
    if (a.n_rows == 0) return 0;
    if (!HAS_ABSENT && a.crec != nullptr && sizeof(CodeT) <= 2) {
        const bool rec4 = a.crec4 != nullptr;
        const size_t smem = (size_t)SMALL_STAGE * ((rec4 ? 4 : 8) + sizeof(uint32_t)) + (size_t)SMALL_THREADS * a.stride * sizeof(CodeT);
        if (smem <= 100 * 1024) {  // two or more CTAs per SM
            const unsigned grid = (unsigned)((a.n_rows + SMALL_THREADS - 1) / SMALL_THREADS);
            auto run = [&](auto k) {
                if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
                k<<<grid, SMALL_THREADS, smem, stream>>>(a);
            };
            if (rec4) run(predict_small_kernel<CodeT, CMAX, MODE, true>);
            else run(predict_small_kernel<CodeT, CMAX, MODE, false>);
            RFMC_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const int64_t blocks = (a.n_rows + PRED_THREADS - 1) / PRED_THREADS;
    size_t tile_bytes = (size_t)PRED_THREADS * a.stride * sizeof(CodeT);
    size_t extra = HAS_ABSENT ? (size_t)((a.n_features + 15) / 16 * 16) : 0;
    if (tile_bytes + extra <= 200 * 1024) {
        auto k = predict_kernel<CodeT, CMAX, MODE, HAS_ABSENT, true>;
        size_t smem = tile_bytes + extra;
        if (smem > 48 * 1024) RFMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(unsigned)blocks, PRED_THREADS, smem, stream>>>(a);
    } else {
        auto k = predict_kernel<CodeT, CMAX, MODE, HAS_ABSENT, false>;
        if (extra > 48 * 1024) RFMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)extra));
        k<<<(unsigned)blocks, PRED_THREADS, extra, stream>>>(a);
    }
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

template <typename CodeT, int CMAX, int MODE>
static int launch_absent(const PredictArgs& a, bool has_absent, cudaStream_t stream) {
    return has_absent ? launch_tile<CodeT, CMAX, MODE, true>(a, stream) : launch_tile<CodeT, CMAX, MODE, false>(a, stream);
}

template <typename CodeT, int CMAX>
static int launch_mode(const PredictArgs& a, int mode, bool has_absent, cudaStream_t stream) {
    switch (mode) {
        case 0: return launch_absent<CodeT, CMAX, 0>(a, has_absent, stream);
        case 1: return launch_absent<CodeT, CMAX, 1>(a, has_absent, stream);
        case 2: return launch_absent<CodeT, CMAX, 2>(a, has_absent, stream);
        default: return launch_absent<CodeT, CMAX, 3>(a, has_absent, stream);
    }
}

template <typename CodeT>
static int launch_classes(const PredictArgs& a, int mode, bool has_absent, cudaStream_t stream) {
    if (a.n_classes <= 4) return launch_mode<CodeT, 4>(a, mode, has_absent, stream);
    if (a.n_classes <= 16) return launch_mode<CodeT, 16>(a, mode, has_absent, stream);
    return launch_mode<CodeT, RFMC_MAX_CLASSES>(a, mode, has_absent, stream);
}

// Per-(row, tree) leaf export: payload index, bit 31 set when the walk crossed an absent column.
// Serves the single-tree API (Tree(row), tree.py:112-113) and per-tree votes (model.py:406-407).
template <typename CodeT>
__global__ void __launch_bounds__(PRED_THREADS) leaves_kernel(PredictArgs a, int32_t* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * PRED_THREADS + threadIdx.x;
    if (row >= a.n_rows) return;
    const CodeT* myrow = reinterpret_cast<const CodeT*>(a.codes) + row * a.stride;
    for (int t = 0; t < a.n_trees; ++t) {
        uint32_t node = (uint32_t)__ldg(&a.root[t]);
        bool seen = false;
        uint4 nd;
        while (true) {
            nd = __ldg(reinterpret_cast<const uint4*>(a.nodes) + node);
            if (nd.x & RFMC_PNODE_LEAF) break;
            const uint32_t f = nd.x & 0xFFFFu;
            const uint32_t cv = (uint32_t)__ldg(&myrow[f]);
            bool go = (nd.x & RFMC_PNODE_CATEGORICAL) ? (cv == nd.y) : (cv >= nd.y);
            if (a.absent != nullptr && a.absent[f]) {
                go = true;
                seen = true;
            }
            node = nd.z + (go ? 0u : 1u);
        }
        out[row * a.n_trees + t] = (int32_t)(nd.y | (seen ? 0x80000000u : 0u));
    }
}

int launch_leaves(rfmc_forest* f, const void* d_codes, int code_width, int64_t n_rows, int64_t row_stride,
                  int32_t n_features, const uint8_t* d_absent, int32_t* d_out, cudaStream_t stream) {
    PredictArgs a;
    a.codes = d_codes;
    a.n_rows = n_rows;
    a.stride = row_stride;
    a.n_features = n_features;
    a.nodes = f->d_nodes;
    a.root = f->d_root;
    a.blocks = nullptr;
    a.root_b = nullptr;
    a.tops = nullptr;
    a.top_levels = 0;
    a.crec = nullptr;
    a.crec4 = nullptr;
    a.cpay = nullptr;
    a.ctree = nullptr;
    a.cdepth = nullptr;
    a.cstage = nullptr;
    a.n_stages = 0;
    a.n_trees = f->n_trees;
    a.vote = f->d_vote;
    a.vote_absent = f->d_vote_absent;
    a.probs = f->d_probs;
    a.probs_absent = f->d_probs_absent;
    a.scores = f->d_scores;
    a.score_fsum = f->score_fsum;
    a.n_classes = f->n_classes;
    a.absent = d_absent;
    a.probs_out = nullptr;
    a.labels_out = nullptr;
    const unsigned blocks = (unsigned)((n_rows + PRED_THREADS - 1) / PRED_THREADS);
    if (blocks == 0) return 0;
    switch (code_width) {
        case 1: leaves_kernel<uint8_t><<<blocks, PRED_THREADS, 0, stream>>>(a, d_out); break;
        case 2: leaves_kernel<uint16_t><<<blocks, PRED_THREADS, 0, stream>>>(a, d_out); break;
        case 4: leaves_kernel<uint32_t><<<blocks, PRED_THREADS, 0, stream>>>(a, d_out); break;
        default: set_error("unsupported code width"); return 1;
    }
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

int launch_predict(rfmc_forest* f, const void* d_codes, int code_width, int64_t n_rows, int64_t row_stride,
                   int32_t n_features, const uint8_t* d_absent, bool has_absent, int soft, int weighted, double* d_probs,
                   int32_t* d_labels, cudaStream_t stream) {
    PredictArgs a;
    a.codes = d_codes;
    a.n_rows = n_rows;
    a.stride = row_stride;
    a.n_features = n_features;
    a.nodes = f->d_nodes;
    a.root = f->d_root;
    a.blocks = (code_width <= 2) ? f->d_blocks : nullptr;
    a.root_b = f->d_root_b;
    // measured slower than the all-global blocked walk on B200 (profiles/r2_vote_staged_tops.md): opt-in only
    a.tops = (a.blocks != nullptr && getenv("RFMC_PRED_TOPS")) ? f->d_tops : nullptr;
    a.top_levels = f->top_levels;
    const char* small_off = getenv("RFMC_PRED_SMALL");  // experiment knob: 0 = always the generic kernel
    const bool small = f->d_crec != nullptr && code_width <= 2 && !(small_off && atoi(small_off) == 0);
    a.crec = small ? f->d_crec : nullptr;
    a.crec4 = f->d_crec4;
    if (const char* e = getenv("RFMC_PRED_SMALL_REC4")) {  // experiment knob: 0 = 8-byte records even when eligible
        if (atoi(e) == 0) a.crec4 = nullptr;
    }
    a.cpay = f->d_cpay;
    a.ctree = f->d_ctree;
    a.cdepth = f->d_cdepth;
    a.cstage = f->d_cstage;
    a.n_stages = f->n_stages;
    a.n_trees = f->n_trees;
    a.vote = f->d_vote;
    a.vote_absent = f->d_vote_absent;
    a.probs = f->d_probs;
    a.probs_absent = f->d_probs_absent;
    a.scores = f->d_scores;
    a.score_fsum = f->score_fsum;
    a.n_classes = f->n_classes;
    a.absent = d_absent;
    a.probs_out = d_probs;
    a.labels_out = d_labels;
    const int mode = (soft ? 2 : 0) + (weighted ? 1 : 0);
    switch (code_width) {
        case 1: return launch_classes<uint8_t>(a, mode, has_absent, stream);
        case 2: return launch_classes<uint16_t>(a, mode, has_absent, stream);
        case 4: return launch_classes<uint32_t>(a, mode, has_absent, stream);
    }
    set_error("unsupported code width");
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Frame encoding: raw column values -> the forest's code space (flat.encode_frame on the GPU).
// The comparisons of tree.py:177-180 only need code(x) = #{split values t of that feature with
// t <= x} (numeric; NaN -> 0) or the id of the category (categorical, mapped on the host).  One
// thread per row walks the frame's columns (each column read is coalesced across the CTA's rows),
// fills its row of a shared-memory tile and the tile leaves as whole coalesced rows.  This is the
// one HBM-streaming kernel of the path: 8 bytes in, 2 bytes out per numeric cell.
// ---------------------------------------------------------------------------------------------
struct EncodeArgs {
    const uint8_t* raw;        // staged chunk: column c starts at raw + col_off[c], chunk_rows elements
    const int64_t* col_off;    // [n_cols]
    const int32_t* col_feat;   // [n_cols] destination feature index
    const int32_t* col_dtype;  // [n_cols] RFMC_DT_*
    int32_t n_cols;
    int64_t n_rows;            // rows in this chunk
    int64_t stride;            // elements per output row
    const double* thresholds;  // concatenated per-feature sorted split values
    const int64_t* thr_off;    // [n_features+1]
    void* codes;               // [n_rows][stride]
};

__device__ __forceinline__ double load_as_double(const uint8_t* p, int dtype, int64_t i) {
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

template <typename CodeT>
__global__ void __launch_bounds__(PRED_THREADS) encode_kernel(EncodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CodeT* tile = reinterpret_cast<CodeT*>(smem_raw);
    const int tid = threadIdx.x;
    const int T = (int)blockDim.x;  // rows per tile: 128, or fewer when the rows are wide
    const int64_t row0 = (int64_t)blockIdx.x * T;
    const int rows = (int)((a.n_rows - row0) < T ? (a.n_rows - row0) : T);
    const int words = (int)((size_t)T * a.stride * sizeof(CodeT) / 4);
    for (int i = tid; i < words; i += T) reinterpret_cast<uint32_t*>(tile)[i] = 0u;
    __syncthreads();
    if (tid < rows) {
        CodeT* myrow = tile + (size_t)tid * a.stride;
        const int64_t r = row0 + tid;
        for (int c = 0; c < a.n_cols; ++c) {
            const int dt = a.col_dtype[c];
            const int f = a.col_feat[c];
            const uint8_t* base = a.raw + a.col_off[c];
            uint32_t code;
            if (dt == RFMC_DT_CODE) {
                code = (uint32_t)reinterpret_cast<const int*>(base)[r];
            } else {
                const double x = load_as_double(base, dt, r);
                const double* t = a.thresholds + a.thr_off[f];
                uint32_t lo = 0, hi = (uint32_t)(a.thr_off[f + 1] - a.thr_off[f]);
                if (x != x) hi = 0;  // NaN fails every comparison -> code 0
                while (hi > lo) {  // upper_bound: number of split values <= x
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if (__ldg(&t[mid]) <= x)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                code = lo;
            }
            myrow[f] = (CodeT)code;
        }
    }
    __syncthreads();
    const int out_words = (int)((size_t)rows * a.stride * sizeof(CodeT) / 4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<CodeT*>(a.codes) + row0 * a.stride);
    for (int i = tid; i < out_words; i += T) dst[i] = reinterpret_cast<const uint32_t*>(tile)[i];
}

int launch_encode(const void* d_raw, const int64_t* d_col_off, const int32_t* d_col_feat, const int32_t* d_col_dtype,
                  int32_t n_cols, int64_t n_rows, int64_t row_stride, int code_width, const double* d_thresholds,
                  const int64_t* d_thr_off, void* d_codes, cudaStream_t stream) {
    EncodeArgs a;
    a.raw = (const uint8_t*)d_raw;
    a.col_off = d_col_off;
    a.col_feat = d_col_feat;
    a.col_dtype = d_col_dtype;
    a.n_cols = n_cols;
    a.n_rows = n_rows;
    a.stride = row_stride;
    a.thresholds = d_thresholds;
    a.thr_off = d_thr_off;
    a.codes = d_codes;
    if (n_rows == 0) return 0;
    int T = PRED_THREADS;  // rows per tile; wide frames get smaller tiles so the tile still fits shared memory
    while (T > 32 && (size_t)T * row_stride * code_width > 200 * 1024) T >>= 1;
    const size_t smem = (size_t)T * row_stride * code_width;
    if (smem > 200 * 1024) {
        set_error("frame encode: too many feature columns for the shared-memory tile");
        return 1;
    }
    const unsigned blocks = (unsigned)((n_rows + T - 1) / T);
    switch (code_width) {
        case 1:
            if (smem > 48 * 1024) RFMC_CUDA(cudaFuncSetAttribute(encode_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            encode_kernel<uint8_t><<<blocks, T, smem, stream>>>(a);
            break;
        case 2:
            if (smem > 48 * 1024) RFMC_CUDA(cudaFuncSetAttribute(encode_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            encode_kernel<uint16_t><<<blocks, T, smem, stream>>>(a);
            break;
        case 4:
            if (smem > 48 * 1024) RFMC_CUDA(cudaFuncSetAttribute(encode_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            encode_kernel<uint32_t><<<blocks, T, smem, stream>>>(a);
            break;
        default: set_error("unsupported code width"); return 1;
    }
    RFMC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rfmc
