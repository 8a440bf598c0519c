// extern "C" entry points of the device draw plan (include/rfmc_b200.h, section "split_train_val /
// sampleFeats on the device"); kernels in rng_walk.cu, jump polynomials in mt_jump.cpp.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"

using namespace rfmc;

namespace {

constexpr int MT_N = 624;

// Jump polynomials t^(c*J - 1) mod phi for the standard chunk length, resident per device and grown
// on demand: they are constants of the generator, not of a stream.
struct PolyTable {
    uint32_t* d = nullptr;
    int64_t have = 0, cap = 0;
};
std::atomic<bool> g_plan_timing{false};
std::mutex g_poly_mu;
std::map<std::pair<int, int64_t>, PolyTable> g_polys;

int poly_table(int device, int64_t chunk_words, int64_t n_chunks, cudaStream_t st, const uint32_t** out) {
    *out = nullptr;
    if (n_chunks <= 1) return 0;
    const int64_t want = n_chunks - 1;
    std::lock_guard<std::mutex> lk(g_poly_mu);
    PolyTable& t = g_polys[{device, chunk_words}];
    if (t.have < want) {
        if (t.cap < want) {
            // Plans in flight on other streams hold the table's address: a table that has to grow is replaced
            // by a larger one and LEFT ALLOCATED (a few MB, once per chunk length), never freed under them.
            int64_t cap = t.cap ? t.cap : 1024;
            while (cap < want) cap *= 2;
            uint32_t* nd = nullptr;
            RFMC_CUDA(cudaMalloc((void**)&nd, (size_t)cap * MT_N * 4));
            if (t.have) RFMC_CUDA(cudaMemcpy(nd, t.d, (size_t)t.have * MT_N * 4, cudaMemcpyDeviceToDevice));
            t.d = nd;
            t.cap = cap;
        }
        std::vector<uint32_t> h((size_t)(want - t.have) * MT_N);
        for (int64_t c = t.have; c < want; ++c)
            RFMC_REQUIRE(rfmc_mt_jump_poly((c + 1) * chunk_words, h.data() + (size_t)(c - t.have) * MT_N) == 0, "jump polynomial failed");
        RFMC_CUDA(cudaMemcpy(t.d + (size_t)t.have * MT_N, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        t.have = want;
    }
    (void)st;
    *out = t.d;
    return 0;
}

inline uint32_t all_ones_at_least(uint32_t v) {
    v |= v >> 1, v |= v >> 2, v |= v >> 4, v |= v >> 8, v |= v >> 16;
    return v;
}

inline uint32_t untemper(uint32_t y) {
    y ^= y >> 18;
    y ^= (y << 15) & 0xefc60000u;
    uint32_t x = y;
    for (int k = 0; k < 5; ++k) x = y ^ ((x << 7) & 0x9d2c5680u);
    y = x;
    for (int k = 0; k < 3; ++k) x = y ^ (x >> 11);
    return x;
}

void plan_free_buffers(rfmc_drawplan* pl) {
    release(pl->d_W);
    release(pl->d_A);
    release(pl->d_heads);
    release(pl->d_idx_train);
    release(pl->d_idx_val);
    release(pl->d_randv);
    release(pl->d_slot_pos);
    release(pl->d_status);
    release(pl->d_key);
    release(pl->d_k0);
    release(pl->d_kpred);
    release(pl->d_eband);
    release(pl->d_hdr);
    release(pl->d_pend);
    pl->d_k0 = pl->d_kpred = nullptr;
    pl->d_eband = nullptr;
    pl->d_hdr = pl->d_pend = nullptr;
    pl->d_W = pl->d_A = pl->d_heads = nullptr;
    pl->d_idx_train = pl->d_idx_val = pl->d_randv = nullptr;
    pl->d_slot_pos = nullptr;
    pl->d_status = nullptr;
    pl->d_key = nullptr;
}

int plan_start(rfmc_drawplan* pl) {
    rfmc_dataset* ds = pl->ds;
    const int C = pl->n_classes;
    const int64_t words = drawplan_words_needed(ds->h_class_n, pl->n_slots, pl->n_cands, pl->rmask, pl->rrange, pl->attempt);
    drawplan_geometry(pl, words);
    RFMC_REQUIRE(poly_table(ds->device, drawplan_chunk_words(pl), pl->n_chunks, pl->stream, &pl->d_polys) == 0, "jump table failed");
    int64_t a_per_slot = 0;  // ordinals per slot: every class's shuffle steps, then the randint block
    for (int c = 0; c < C; ++c) a_per_slot += ds->h_class_n[c] > 1 ? ds->h_class_n[c] - 1 : 0;
    if (pl->n_cands > 0 && pl->rrange != 0u) a_per_slot += pl->n_cands;
    RFMC_REQUIRE(a_per_slot < 0x7fffffffll, "too many draws per slot");
    const size_t S = (size_t)pl->n_slots;
    const size_t G = (size_t)pl->chunks_per_round;
    RFMC_CUDA(dev_alloc((void**)&pl->d_k0, ((size_t)pl->n_walk_chunks + 1) * 8));
    RFMC_CUDA(dev_alloc((void**)&pl->d_kpred, (G + 1) * 8));
    RFMC_CUDA(dev_alloc((void**)&pl->d_eband, G * 4));
    RFMC_CUDA(dev_alloc(&pl->d_hdr, G * 8));
    RFMC_CUDA(dev_alloc(&pl->d_pend, G * 4096 * 8));
    RFMC_CUDA(dev_alloc((void**)&pl->d_key, MT_N * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_W, (size_t)pl->n_blocks * MT_N * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_A, (S * (size_t)a_per_slot + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_heads, (S * C * (size_t)pl->size + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_idx_train, (S * C * (size_t)pl->n_tr + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_idx_val, (S * C * (size_t)pl->n_va + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_randv, (S * (size_t)pl->n_cands + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&pl->d_slot_pos, (S + 1) * 8));
    RFMC_CUDA(dev_alloc((void**)&pl->d_status, 64 * 4));
    RFMC_CUDA(cudaMemcpyAsync(pl->d_key, pl->h_key, MT_N * 4, cudaMemcpyHostToDevice, pl->stream));
    RFMC_CUDA(cudaMemsetAsync(pl->d_status, 0, 256, pl->stream));
    return drawplan_launch(pl);
}

}  // namespace

extern "C" {

int rfmc_dataset_set_class_rows(rfmc_dataset* ds, int32_t n_classes, const int64_t* class_rows, const int64_t* class_off) {
    RFMC_REQUIRE(ds != nullptr && class_rows != nullptr && class_off != nullptr, "NULL argument");
    RFMC_REQUIRE(n_classes >= 1 && n_classes <= RFMC_MAX_CLASSES, "n_classes must be in 1..64");
    RFMC_CUDA(cudaSetDevice(ds->device));
    const int64_t total = class_off[n_classes];
    RFMC_REQUIRE(total >= 0 && total <= ds->n_rows, "class rows exceed the dataset");
    std::vector<int32_t> rows32((size_t)total + 1), cn((size_t)n_classes);
    std::vector<int64_t> aoff((size_t)n_classes + 1, 0), roff(class_off, class_off + n_classes + 1);
    for (int64_t i = 0; i < total; ++i) {
        RFMC_REQUIRE(class_rows[i] >= 0 && class_rows[i] < ds->n_rows, "row id out of range");
        rows32[(size_t)i] = (int32_t)class_rows[i];
    }
    ds->h_class_n.assign((size_t)n_classes, 0);
    for (int c = 0; c < n_classes; ++c) {
        const int64_t n = class_off[c + 1] - class_off[c];
        RFMC_REQUIRE(n >= 0 && n < 0x7fffffffll, "bad class size");
        ds->h_class_n[c] = n;
        cn[c] = (int32_t)n;
        aoff[c + 1] = aoff[c] + (n > 1 ? n - 1 : 0);
    }
    release(ds->d_class_n);
    release(ds->d_class_aoff);
    release(ds->d_class_rows);
    release(ds->d_class_roff);
    ds->d_class_n = nullptr, ds->d_class_aoff = nullptr, ds->d_class_rows = nullptr, ds->d_class_roff = nullptr;
    RFMC_CUDA(dev_alloc((void**)&ds->d_class_n, (size_t)n_classes * 4));
    RFMC_CUDA(dev_alloc((void**)&ds->d_class_aoff, ((size_t)n_classes + 1) * 8));
    RFMC_CUDA(dev_alloc((void**)&ds->d_class_rows, ((size_t)total + 1) * 4));
    RFMC_CUDA(dev_alloc((void**)&ds->d_class_roff, ((size_t)n_classes + 1) * 8));
    RFMC_CUDA(cudaMemcpy(ds->d_class_n, cn.data(), (size_t)n_classes * 4, cudaMemcpyHostToDevice));
    RFMC_CUDA(cudaMemcpy(ds->d_class_aoff, aoff.data(), ((size_t)n_classes + 1) * 8, cudaMemcpyHostToDevice));
    RFMC_CUDA(cudaMemcpy(ds->d_class_rows, rows32.data(), ((size_t)total + 1) * 4, cudaMemcpyHostToDevice));
    RFMC_CUDA(cudaMemcpy(ds->d_class_roff, roff.data(), ((size_t)n_classes + 1) * 8, cudaMemcpyHostToDevice));
    ds->n_draw_classes = n_classes;
    return 0;
}

int rfmc_draw_plan_device(rfmc_dataset* ds, const uint32_t* key, int32_t pos, int32_t n_slots, int64_t per_class,
                          int64_t n_train_pclass, int32_t n_cands, int64_t randint_low, int64_t randint_high, void* stream,
                          rfmc_drawplan** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(ds != nullptr && key != nullptr, "NULL argument");
    RFMC_REQUIRE(ds->n_draw_classes > 0, "rfmc_dataset_set_class_rows was not called");
    RFMC_REQUIRE(pos >= 0 && pos <= MT_N, "pos must be in 0..624");
    RFMC_REQUIRE(n_slots >= 1 && per_class >= 0 && n_train_pclass >= 0 && n_cands >= 0, "bad plan shape");
    RFMC_REQUIRE(randint_high > randint_low, "low >= high");
    const uint64_t rng = (uint64_t)(randint_high - 1 - randint_low);
    RFMC_REQUIRE(rng <= 0x7FFFFFFEull, "randint range too wide");
    RFMC_REQUIRE(per_class <= RFMC_DRAW_MAX_PER_CLASS, "per_class exceeds the device resolve (use rfmc_legacy_draw_plan)");
    for (int64_t n : ds->h_class_n) RFMC_REQUIRE(n >= per_class, "a class is smaller than per_class");
    RFMC_CUDA(cudaSetDevice(ds->device));
    rfmc_drawplan* pl = new rfmc_drawplan();
    pl->ds = ds;
    pl->stream = (cudaStream_t)stream;
    pl->n_slots = n_slots;
    pl->n_classes = ds->n_draw_classes;
    pl->n_cands = n_cands;
    pl->pos0 = pos;
    pl->size = per_class;
    pl->n_tr = per_class < n_train_pclass ? per_class : n_train_pclass;
    pl->n_va = per_class - pl->n_tr;
    pl->randint_low = randint_low;
    pl->rrange = (uint32_t)rng;
    pl->rmask = all_ones_at_least((uint32_t)rng);
    int cl = 10;
    while (((int64_t)1 << cl) < 4 * per_class) ++cl;
    pl->cap_log2 = cl;
    memcpy(pl->h_key, key, sizeof(pl->h_key));
    pl->timing = g_plan_timing.load();
    if (pl->timing)
        for (int i = 0; i < 5; ++i)
            if (cudaEventCreate(&pl->ev[i]) != cudaSuccess) pl->timing = false;
    if (plan_start(pl)) {
        rfmc_drawplan_destroy(pl);
        return 1;
    }
    *out = pl;
    return 0;
}

int rfmc_drawplan_finish(rfmc_drawplan* pl, int32_t* feat_counts, uint32_t* key_end, int32_t* pos_end) {
    RFMC_REQUIRE(pl != nullptr, "plan is NULL");
    RFMC_CUDA(cudaSetDevice(pl->ds->device));
    for (;;) {
        RFMC_CUDA(cudaMemcpyAsync(pl->h_status, pl->d_status, 32, cudaMemcpyDeviceToHost, pl->stream));
        pl->h_slot_pos.resize((size_t)pl->n_slots + 1);
        RFMC_CUDA(cudaMemcpyAsync(pl->h_slot_pos.data(), pl->d_slot_pos, ((size_t)pl->n_slots + 1) * 8, cudaMemcpyDeviceToHost, pl->stream));
        RFMC_CUDA(cudaStreamSynchronize(pl->stream));
        RFMC_REQUIRE(pl->h_status[2] == 0, "device draw: resolve scratch overflow");
        RFMC_REQUIRE(pl->h_status[3] == 0, "device draw: the final pass and the chain disagree on a chunk's accept count");
        if (pl->h_slot_pos[(size_t)pl->n_slots] >= 0) break;
        // the generated stretch of the stream was too short (beyond 8 sigma): redo the plan with a longer one
        RFMC_REQUIRE(pl->attempt < 3, "device draw: stream words exhausted");
        pl->attempt += 1;
        plan_free_buffers(pl);
        if (plan_start(pl)) return 1;
    }
    pl->finished = true;
    if (feat_counts && pl->n_cands > 0) {
        const size_t n = (size_t)pl->n_slots * pl->n_cands;
        if (pl->rrange == 0u) {
            for (size_t i = 0; i < n; ++i) feat_counts[i] = (int32_t)pl->randint_low;  // randint over one value draws nothing
        } else {
            // the randint block closes every slot's stretch of A
            const size_t stride = (size_t)pl->steps_per_slot;
            RFMC_CUDA(cudaMemcpy2D(feat_counts, (size_t)pl->n_cands * 4, pl->d_A + (stride - (size_t)pl->n_cands), stride * 4,
                                   (size_t)pl->n_cands * 4, (size_t)pl->n_slots, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) feat_counts[i] = (int32_t)(pl->randint_low + (int64_t)feat_counts[i]);
        }
    }
    if (key_end && pos_end) return rfmc_drawplan_state_at(pl, pl->n_slots, key_end, pos_end);
    return 0;
}

int rfmc_drawplan_state_at(rfmc_drawplan* pl, int32_t slot, uint32_t* key, int32_t* pos) {
    RFMC_REQUIRE(pl != nullptr && key != nullptr && pos != nullptr, "NULL argument");
    RFMC_REQUIRE(pl->finished, "rfmc_drawplan_finish has not been called");
    RFMC_REQUIRE(slot >= 0 && slot <= pl->n_slots, "slot out of range");
    RFMC_CUDA(cudaSetDevice(pl->ds->device));
    const int64_t P = pl->h_slot_pos[(size_t)slot];
    // numpy regenerates a block only when it needs its first word: after the last word of block b the
    // state is (block b, pos 624), not (block b+1, pos 0)
    int64_t blk = P / MT_N;
    int32_t p = (int32_t)(P % MT_N);
    if (p == 0 && blk > 0) {
        blk -= 1;
        p = MT_N;
    }
    if (blk == 0) {
        memcpy(key, pl->h_key, sizeof(pl->h_key));
    } else {
        RFMC_CUDA(cudaMemcpy(key, pl->d_W + (size_t)blk * MT_N, MT_N * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < MT_N; ++i) key[i] = untemper(key[i]);
    }
    *pos = p;
    return 0;
}

int rfmc_drawplan_indices(rfmc_drawplan* pl, int32_t slot0, int32_t n_slots, int32_t* idx_train, int32_t* idx_val) {
    RFMC_REQUIRE(pl != nullptr, "plan is NULL");
    RFMC_REQUIRE(pl->finished, "rfmc_drawplan_finish has not been called");
    RFMC_REQUIRE(slot0 >= 0 && n_slots >= 0 && slot0 + n_slots <= pl->n_slots, "slot range out of bounds");
    RFMC_CUDA(cudaSetDevice(pl->ds->device));
    const size_t tr = (size_t)pl->n_classes * pl->n_tr, va = (size_t)pl->n_classes * pl->n_va;
    if (idx_train && tr && n_slots)
        RFMC_CUDA(cudaMemcpy(idx_train, pl->d_idx_train + (size_t)slot0 * tr, (size_t)n_slots * tr * 4, cudaMemcpyDeviceToHost));
    if (idx_val && va && n_slots)
        RFMC_CUDA(cudaMemcpy(idx_val, pl->d_idx_val + (size_t)slot0 * va, (size_t)n_slots * va * 4, cudaMemcpyDeviceToHost));
    return 0;
}

/* test hook (not in the header): the accepted draws j_(n-1-k), k = 0 .. n-2, of one class of one slot */
int rfmc_drawplan_draws(rfmc_drawplan* pl, int32_t slot, int32_t cls, uint32_t* out) {
    RFMC_REQUIRE(pl != nullptr && pl->finished && out != nullptr, "plan is NULL or not finished");
    RFMC_REQUIRE(slot >= 0 && slot < pl->n_slots && cls >= 0 && cls < pl->n_classes, "out of range");
    RFMC_CUDA(cudaSetDevice(pl->ds->device));
    int64_t off = 0;
    const int64_t per_slot = pl->steps_per_slot;
    for (int c = 0; c < cls; ++c) off += pl->ds->h_class_n[c] > 1 ? pl->ds->h_class_n[c] - 1 : 0;
    const int64_t m = pl->ds->h_class_n[cls] > 1 ? pl->ds->h_class_n[cls] - 1 : 0;
    if (m) RFMC_CUDA(cudaMemcpy(out, pl->d_A + (size_t)slot * per_slot + off, (size_t)m * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int rfmc_drawplan_info(rfmc_drawplan* pl, int64_t* words_generated, int64_t* words_consumed, int64_t* words_pending,
                       int32_t* chunk_fallbacks, int32_t* launches) {
    // chunk_fallbacks: chunks the final pass redid with every word pending (its tight band missed)
    RFMC_REQUIRE(pl != nullptr && pl->finished, "plan is NULL or not finished");
    if (words_generated) *words_generated = pl->n_blocks * MT_N;
    if (words_consumed) *words_consumed = pl->h_slot_pos[(size_t)pl->n_slots] - pl->pos0;
    if (words_pending) *words_pending = (int64_t)pl->h_status[4] << 10;
    if (chunk_fallbacks) *chunk_fallbacks = pl->h_status[1];
    if (launches) *launches = pl->launches * (pl->attempt + 1);
    return 0;
}

int rfmc_drawplan_destroy(rfmc_drawplan* pl) {
    if (!pl) return 0;
    cudaSetDevice(pl->ds->device);
    cudaStreamSynchronize(pl->stream);
    plan_free_buffers(pl);
    for (int i = 0; i < 5; ++i)
        if (pl->ev[i]) cudaEventDestroy(pl->ev[i]);
    delete pl;
    return 0;
}

int rfmc_drawplan_timing(int enabled) {
    g_plan_timing.store(enabled != 0);
    return 0;
}

int rfmc_drawplan_kernel_ms(rfmc_drawplan* pl, float* ms4) {
    RFMC_REQUIRE(pl != nullptr && ms4 != nullptr, "NULL argument");
    RFMC_REQUIRE(pl->timing && pl->ev[0], "plan was created without rfmc_drawplan_timing(1)");
    RFMC_CUDA(cudaSetDevice(pl->ds->device));
    RFMC_CUDA(cudaEventSynchronize(pl->ev[4]));
    for (int i = 0; i < 4; ++i) RFMC_CUDA(cudaEventElapsedTime(&ms4[i], pl->ev[i], pl->ev[i + 1]));
    return 0;
}

}  // extern "C"
