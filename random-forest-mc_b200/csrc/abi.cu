// extern "C" entry points of librfmc_b200.so (declared in include/rfmc_b200.h).
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace rfmc {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

// Device memory comes from a small caching allocator: a fit makes one batch per plan with the
// same scratch and output buffers, and cudaMalloc/cudaFree cost milliseconds each -- cudaFree also
// synchronises the whole device, which serialises host threads that feed the GPU through their own
// streams.  Every block is kept per device on release and handed out again to a request of similar
// size (small requests are rounded to a power of two so they always find a match).  Every release
// in this file happens after the work that used the block has been synchronised.
struct DevCache {
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;
    size_t cached = 0;
};
static std::mutex g_mem_mu;
static std::map<int, DevCache> g_mem;
constexpr size_t CACHE_SMALL = 1u << 20;  // below this: power-of-two size classes
constexpr size_t CACHE_CAP = 48ull << 30;

static size_t size_class(size_t bytes) {
    if (bytes >= CACHE_SMALL) return (bytes + 511) & ~(size_t)511;
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
}

static void flush_cache_locked(DevCache& dc) {
    for (auto& kv : dc.free_blocks) cudaFree(kv.second);
    dc.free_blocks.clear();
    dc.cached = 0;
}

cudaError_t dev_alloc(void** p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    bytes = size_class(bytes);
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        DevCache& dc = g_mem[dev];
        auto it = dc.free_blocks.lower_bound(bytes);
        if (it != dc.free_blocks.end() && it->first <= bytes + bytes / 2) {
            *p = it->second;
            dc.live[*p] = it->first;
            dc.cached -= it->first;
            dc.free_blocks.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);  // outside the lock: other host threads keep allocating from the cache
    std::lock_guard<std::mutex> lk(g_mem_mu);
    DevCache& dc = g_mem[dev];
    if (e != cudaSuccess) {
        cudaGetLastError();
        flush_cache_locked(dc);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) dc.live[*p] = bytes;
    return e;
}

void release(void* p) {
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        DevCache& dc = g_mem[dev];
        auto it = dc.live.find(p);
        if (it != dc.live.end()) {
            const size_t bytes = it->second;
            dc.live.erase(it);
            if (dc.cached + bytes <= CACHE_CAP) {
                dc.free_blocks.emplace(bytes, p);
                dc.cached += bytes;
                return;
            }
        }
    }
    cudaFree(p);
}

template <typename T>
static int upload(T** dst, const T* src, size_t count, cudaStream_t stream) {
    *dst = nullptr;
    if (count == 0) return 0;
    RFMC_CUDA(dev_alloc((void**)dst, count * sizeof(T)));
    RFMC_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    return 0;
}
template <typename T>
static int reserve(T** dst, size_t count) {
    *dst = nullptr;
    if (count == 0) count = 1;
    RFMC_CUDA(dev_alloc((void**)dst, count * sizeof(T)));
    return 0;
}

// Pinned staging buffers for device -> host survivor export come from a process-wide pool: the host
// threads of a multi-stream fit are short-lived, and page-locking tens of megabytes per thread per
// fit costs more than the copy it speeds up.
static std::mutex g_pin_mu;
static std::multimap<size_t, void*> g_pin_free;
static size_t g_pin_cached = 0;
constexpr size_t PIN_CAP = 8ull << 30;

struct PinnedLease {
    void* p = nullptr;
    size_t cap = 0;
    explicit PinnedLease(size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(g_pin_mu);
            auto it = g_pin_free.lower_bound(bytes);
            if (it != g_pin_free.end() && it->first <= 2 * bytes + (1u << 20)) {
                p = it->second;
                cap = it->first;
                g_pin_cached -= cap;
                g_pin_free.erase(it);
                return;
            }
        }
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return;
        }
        cap = want;
    }
    ~PinnedLease() {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lk(g_pin_mu);
            if (g_pin_cached + cap <= PIN_CAP) {
                g_pin_free.emplace(cap, p);
                g_pin_cached += cap;
                return;
            }
        }
        cudaFreeHost(p);
    }
    PinnedLease(const PinnedLease&) = delete;
    PinnedLease& operator=(const PinnedLease&) = delete;
};
}  // namespace rfmc

using namespace rfmc;

struct PredictPipe {
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    uint8_t* pin_in[2] = {nullptr, nullptr};
    uint8_t* d_in[2] = {nullptr, nullptr};
    double* pin_probs[2] = {nullptr, nullptr};
    double* d_probs[2] = {nullptr, nullptr};
    int32_t* pin_labels[2] = {nullptr, nullptr};
    int32_t* d_labels[2] = {nullptr, nullptr};
    uint8_t* pin_raw[2] = {nullptr, nullptr};
    uint8_t* d_raw[2] = {nullptr, nullptr};
    size_t cap_in = 0, cap_probs = 0, cap_labels = 0, cap_raw = 0;
    long launches = 0;
};

// One pipe (streams, events, pinned staging) per forest: host-buffer voting calls on the same forest
// from several threads take turns (ctypes releases the GIL); device-pointer calls do not use the pipe.
struct PipeLease {
    rfmc_forest* f;
    explicit PipeLease(rfmc_forest* forest) : f(forest) {
        f->pipe_mu.lock();
        if (!f->pipe) f->pipe = new PredictPipe();
    }
    ~PipeLease() { f->pipe_mu.unlock(); }
    PredictPipe* get() const { return f->pipe; }
};

static int pipe_reserve(PredictPipe* p, size_t in_bytes, size_t probs_bytes, size_t labels_bytes) {
    for (int i = 0; i < 2; ++i) {
        if (!p->st[i]) RFMC_CUDA(cudaStreamCreateWithFlags(&p->st[i], cudaStreamNonBlocking));
        if (!p->done[i]) RFMC_CUDA(cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming));
    }
    if (in_bytes > p->cap_in) {
        for (int i = 0; i < 2; ++i) {
            if (p->pin_in[i]) cudaFreeHost(p->pin_in[i]);
            release(p->d_in[i]);
            RFMC_CUDA(cudaHostAlloc((void**)&p->pin_in[i], in_bytes, cudaHostAllocDefault));
            RFMC_CUDA(cudaMalloc((void**)&p->d_in[i], in_bytes));
        }
        p->cap_in = in_bytes;
    }
    if (probs_bytes > p->cap_probs) {
        for (int i = 0; i < 2; ++i) {
            if (p->pin_probs[i]) cudaFreeHost(p->pin_probs[i]);
            release(p->d_probs[i]);
            RFMC_CUDA(cudaHostAlloc((void**)&p->pin_probs[i], probs_bytes, cudaHostAllocDefault));
            RFMC_CUDA(cudaMalloc((void**)&p->d_probs[i], probs_bytes));
        }
        p->cap_probs = probs_bytes;
    }
    if (labels_bytes > p->cap_labels) {
        for (int i = 0; i < 2; ++i) {
            if (p->pin_labels[i]) cudaFreeHost(p->pin_labels[i]);
            release(p->d_labels[i]);
            RFMC_CUDA(cudaHostAlloc((void**)&p->pin_labels[i], labels_bytes, cudaHostAllocDefault));
            RFMC_CUDA(cudaMalloc((void**)&p->d_labels[i], labels_bytes));
        }
        p->cap_labels = labels_bytes;
    }
    return 0;
}

static int pipe_reserve_raw(PredictPipe* p, size_t raw_bytes) {
    if (raw_bytes > p->cap_raw) {
        for (int i = 0; i < 2; ++i) {
            if (p->pin_raw[i]) cudaFreeHost(p->pin_raw[i]);
            release(p->d_raw[i]);
            RFMC_CUDA(cudaHostAlloc((void**)&p->pin_raw[i], raw_bytes, cudaHostAllocDefault));
            RFMC_CUDA(cudaMalloc((void**)&p->d_raw[i], raw_bytes));
        }
        p->cap_raw = raw_bytes;
    }
    return 0;
}

static void pipe_destroy(PredictPipe* p) {
    if (!p) return;
    for (int i = 0; i < 2; ++i) {
        if (p->pin_raw[i]) cudaFreeHost(p->pin_raw[i]);
        release(p->d_raw[i]);
        if (p->pin_in[i]) cudaFreeHost(p->pin_in[i]);
        if (p->pin_probs[i]) cudaFreeHost(p->pin_probs[i]);
        if (p->pin_labels[i]) cudaFreeHost(p->pin_labels[i]);
        release(p->d_in[i]);
        release(p->d_probs[i]);
        release(p->d_labels[i]);
        if (p->done[i]) cudaEventDestroy(p->done[i]);
        if (p->st[i]) cudaStreamDestroy(p->st[i]);
    }
    delete p;
}

// Blocked layout for the vote kernel: every block = 4 records of 8 bytes {lo = feat | cat<<14 |
// leaf<<15 | thr<<16, hi = next}: slot 0 a node, slots 1/2 its '>=' / '<' children; a split child's
// `next` is the block of ITS '>=' child (the '<' child is the next block).  Returns false when the
// forest is not eligible (codes or feature ids too wide), in which case the 16-byte nodes are used.
static bool build_blocks(const rfmc_pnode* nodes, int64_t n_nodes, const int64_t* roots, int32_t n_trees,
                         std::vector<uint32_t>& out, std::vector<uint32_t>& root_b, std::vector<uint32_t>& block_of) {
    block_of.assign((size_t)n_nodes, 0xFFFFFFFFu);  // node -> the block that holds it in slot 0 (even depths only)
    for (int64_t i = 0; i < n_nodes; ++i) {
        const rfmc_pnode& n = nodes[i];
        if (n.feat_flags & RFMC_PNODE_LEAF) continue;
        if ((n.feat_flags & 0xFFFFu) >= 0x4000u) return false;
        if (n.thr != 0xFFFFFFFFu && n.thr > 65534u) return false;
    }
    auto rec_lo = [](const rfmc_pnode& n) -> uint32_t {
        if (n.feat_flags & RFMC_PNODE_LEAF) return 0x8000u;
        const uint32_t thr = n.thr == 0xFFFFFFFFu ? 0xFFFFu : n.thr;
        return (n.feat_flags & 0x3FFFu) | ((n.feat_flags & RFMC_PNODE_CATEGORICAL) ? 0x4000u : 0u) | (thr << 16);
    };
    out.clear();
    out.reserve((size_t)n_nodes * 4);
    root_b.assign((size_t)n_trees, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> queue;  // (node, block)
    auto new_blocks = [&](int count) -> uint32_t {
        const uint32_t first = (uint32_t)(out.size() / 8);
        out.resize(out.size() + (size_t)count * 8, 0u);
        return first;
    };
    for (int32_t t = 0; t < n_trees; ++t) {
        queue.clear();
        root_b[t] = new_blocks(1);
        queue.emplace_back((uint32_t)roots[t], root_b[t]);
        for (size_t q = 0; q < queue.size(); ++q) {
            const uint32_t ni = queue[q].first, bi = queue[q].second;
            const rfmc_pnode& p = nodes[ni];
            block_of[ni] = bi;
            out[(size_t)bi * 8 + 0] = rec_lo(p);
            if (p.feat_flags & RFMC_PNODE_LEAF) {
                out[(size_t)bi * 8 + 1] = p.thr;  // payload
                continue;
            }
            for (int side = 0; side < 2; ++side) {
                const uint32_t ci = p.child + (uint32_t)side;
                const rfmc_pnode& x = nodes[ci];
                const size_t at = (size_t)bi * 8 + 2 + 2 * (size_t)side;
                out[at] = rec_lo(x);
                if (x.feat_flags & RFMC_PNODE_LEAF) {
                    out[at + 1] = x.thr;
                } else {
                    const uint32_t nb = new_blocks(2);
                    out[at + 1] = nb;
                    queue.emplace_back(x.child, nb);
                    queue.emplace_back(x.child + 1, nb + 1);
                }
            }
        }
    }
    return true;
}

// The top `levels` levels (even) of every tree as an implicit heap of 8-byte records for
// predict_top_kernel: slot 0 the root, children of slot i in 2i+1 ('>=') and 2i+2 ('<').  A leaf carries
// its payload; a split of the LAST staged level carries the block of its '>=' child in the blocked
// layout (the '<' child's block follows it), where the walk continues.
static void build_tops(const rfmc_pnode* nodes, const int64_t* roots, int32_t n_trees, int levels,
                       const std::vector<uint32_t>& block_of, std::vector<uint32_t>& out) {
    const size_t heap_n = ((size_t)1 << levels) - 1, heap_pad = (heap_n + 2) & ~(size_t)1;
    out.assign((size_t)n_trees * heap_pad * 2, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> todo;  // (node, heap slot)
    for (int32_t t = 0; t < n_trees; ++t) {
        uint32_t* heap = out.data() + (size_t)t * heap_pad * 2;
        todo.clear();
        todo.emplace_back((uint32_t)roots[t], 0u);
        for (size_t q = 0; q < todo.size(); ++q) {
            const uint32_t ni = todo[q].first, hi = todo[q].second;
            const rfmc_pnode& p = nodes[ni];
            if (p.feat_flags & RFMC_PNODE_LEAF) {
                heap[2 * hi] = 0x8000u;
                heap[2 * hi + 1] = p.thr;
                continue;
            }
            const uint32_t thr = p.thr == 0xFFFFFFFFu ? 0xFFFFu : p.thr;
            heap[2 * hi] = (p.feat_flags & 0x3FFFu) | ((p.feat_flags & RFMC_PNODE_CATEGORICAL) ? 0x4000u : 0u) | (thr << 16);
            if (2 * (size_t)hi + 2 < heap_n) {
                todo.emplace_back(p.child, 2 * hi + 1);
                todo.emplace_back(p.child + 1, 2 * hi + 2);
            } else {
                heap[2 * hi + 1] = block_of[p.child];
            }
        }
    }
}

// Compact layout for forests of SMALL trees (the reference's default configuration grows trees of a few dozen
// nodes): 8-byte records {feat | lo << 16, span | next << 16} in per-tree BFS order, children adjacent ('>=' at
// next, '<' at next + 1, tree-relative).  One unsigned compare serves every node kind:
//     go '>=' child  <=>  (code - lo) <= span      numeric: lo = thr, span = 0xFFFF - thr;  categorical: span = 0;
//                                                  leaf: lo = 0, span = 0xFFFF, next = itself (the walk stays put)
// so a walk is a fixed number of identical steps (the tree's depth) without a leaf test.  Trees are grouped into
// stages of at most `stage_nodes` records that the vote kernel copies to shared memory whole.  Returns false
// when the forest is not eligible (a tree larger than a stage, codes or feature ids too wide).
constexpr uint32_t SMALL_STAGE_NODES = 1024, SMALL_STAGE_TREES = 64;
// rec4: the same forest as 4-byte records {feat (8 bits) | delta << 8 (7 bits) | numeric << 15 | lo << 16} when every
// feature id is below 256 and every '>=' child lies at most 127 records behind its parent (BFS order: the width of a
// level); left empty otherwise.  next = own index + delta, span = numeric ? 0x10000 : 0, a leaf = {numeric, lo 0,
// delta 0}.  Half the bytes per record halve the shared-memory bank conflicts of 32 lanes in 32 different records.
static bool build_compact(const rfmc_pnode* nodes, const int64_t* roots, int32_t n_trees, std::vector<uint32_t>& rec,
                          std::vector<uint32_t>& rec4, std::vector<uint32_t>& pay, std::vector<uint32_t>& tree_at,
                          std::vector<uint8_t>& depth, std::vector<uint32_t>& stage) {
    rec.clear();
    rec4.clear();
    bool narrow = true;
    pay.clear();
    tree_at.assign(1, 0u);
    depth.clear();
    stage.assign(1, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> queue;  // (node, depth)
    uint32_t in_stage = 0;
    for (int32_t t = 0; t < n_trees; ++t) {
        const uint32_t base = (uint32_t)pay.size();
        queue.clear();
        queue.emplace_back((uint32_t)roots[t], 0u);
        uint32_t used = 1, deepest = 0;  // records handed out so far (tree-relative)
        for (size_t q = 0; q < queue.size(); ++q) {
            if (queue.size() > SMALL_STAGE_NODES) return false;
            const rfmc_pnode& p = nodes[queue[q].first];
            const uint32_t d = queue[q].second;
            if (p.feat_flags & RFMC_PNODE_LEAF) {
                rec.push_back(0u);
                rec.push_back(0xFFFFu | ((uint32_t)q << 16));
                rec4.push_back(0x8000u);
                pay.push_back(p.thr);
                deepest = d > deepest ? d : deepest;
                continue;
            }
            const uint32_t f = p.feat_flags & 0xFFFFu;
            if (f >= 0x8000u) return false;
            if (p.thr != 0xFFFFFFFFu && p.thr > 65534u) return false;
            const uint32_t lo = p.thr == 0xFFFFFFFFu ? 0xFFFFu : p.thr;
            const uint32_t span = (p.feat_flags & RFMC_PNODE_CATEGORICAL) ? 0u : (0xFFFFu - lo);
            rec.push_back(f | (lo << 16));
            rec.push_back(span | (used << 16));
            const uint32_t delta = used - (uint32_t)q;
            if (f >= 256u || delta > 127u) narrow = false;
            rec4.push_back((f & 0xFFu) | ((delta & 0x7Fu) << 8) | ((p.feat_flags & RFMC_PNODE_CATEGORICAL) ? 0u : 0x8000u) | (lo << 16));
            pay.push_back(0u);
            queue.emplace_back(p.child, d + 1);      // '>=' child at `used`
            queue.emplace_back(p.child + 1, d + 1);  // '<' child at `used + 1`
            used += 2;
        }
        if (deepest > 255u) return false;
        const uint32_t n = (uint32_t)queue.size();
        if (in_stage + n > SMALL_STAGE_NODES || (uint32_t)t - stage.back() >= SMALL_STAGE_TREES) {
            stage.push_back((uint32_t)t);
            in_stage = 0;
        }
        in_stage += n;
        depth.push_back((uint8_t)deepest);
        tree_at.push_back(base + n);
    }
    stage.push_back((uint32_t)n_trees);
    if (!narrow) rec4.clear();
    return true;
}

extern "C" {

int rfmc_abi_version(void) { return RFMC_ABI_VERSION; }

const char* rfmc_last_error(void) { return g_error.c_str(); }

int rfmc_stream_create(int device, void** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    RFMC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *out = (void*)st;
    return 0;
}

int rfmc_stream_destroy(int device, void* stream) {
    if (!stream) return 0;
    RFMC_CUDA(cudaSetDevice(device));
    RFMC_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return 0;
}

/* debug builds (-DRFMC_GROW_TIMING) only: per-level cycle counters of the grower; not part of the ABI header */
int rfmc_debug_read(unsigned long long* out, int n) { return debug_read(out, n); }

int rfmc_device_count(int* count) {
    RFMC_REQUIRE(count != nullptr, "count is NULL");
    *count = 0;
    RFMC_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int rfmc_dataset_create(int device, const void* codes, int code_width, int64_t n_rows, int32_t n_features,
                        int64_t row_stride, const uint8_t* labels, int32_t n_classes, const uint8_t* col_kind,
                        const int32_t* col_tag, const int32_t* col_nunique, const int64_t* table_off,
                        const double* tables, rfmc_dataset** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(code_width == 1 || code_width == 2 || code_width == 4, "code_width must be 1, 2 or 4");
    RFMC_REQUIRE(n_rows > 0 && n_features > 0 && n_features <= 65535, "bad matrix shape");
    RFMC_REQUIRE(row_stride >= n_features && (row_stride * code_width) % 4 == 0, "row pitch must be a multiple of 4 bytes");
    RFMC_REQUIRE(n_classes >= 1 && n_classes <= RFMC_MAX_CLASSES, "n_classes must be in 1..64");
    RFMC_CUDA(cudaSetDevice(device));
    rfmc_dataset* ds = new rfmc_dataset();
    ds->device = device;
    ds->code_width = code_width;
    ds->n_rows = n_rows;
    ds->n_features = n_features;
    ds->stride = row_stride;
    ds->n_classes = n_classes;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        delete ds;
        return 1;
    }
    ds->sm_count = prop.multiProcessorCount;
    int rc = 0;
    rc |= upload((uint8_t**)&ds->d_codes, (const uint8_t*)codes, (size_t)n_rows * row_stride * code_width, 0);
    rc |= upload(&ds->d_labels, labels, (size_t)n_rows, 0);
    rc |= upload(&ds->d_kind, col_kind, (size_t)n_features, 0);
    rc |= upload(&ds->d_tag, col_tag, (size_t)n_features, 0);
    rc |= upload(&ds->d_nuniq, col_nunique, (size_t)n_features, 0);
    rc |= upload(&ds->d_taboff, table_off, (size_t)n_features + 1, 0);
    size_t ntab = (size_t)table_off[n_features];
    if (ntab == 0) {
        rc |= reserve(&ds->d_tables, 1);
    } else {
        rc |= upload(&ds->d_tables, tables, ntab, 0);
    }
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) {
        set_error("dataset upload failed");
        rc = 1;
    }
    if (rc) {
        rfmc_dataset_destroy(ds);
        return 1;
    }
    *out = ds;
    return 0;
}

static size_t dtype_size(int dt) {
    switch (dt) {
        case RFMC_DT_F64: case RFMC_DT_I64: case RFMC_DT_U64: return 8;
        case RFMC_DT_F32: case RFMC_DT_I32: case RFMC_DT_U32: case RFMC_DT_CODE: return 4;
        case RFMC_DT_I16: case RFMC_DT_U16: return 2;
        case RFMC_DT_I8: case RFMC_DT_U8: return 1;
    }
    return 0;
}

namespace {
// scratch of rfmc_dataset_encode, released on every exit path
struct EncodeScratch {
    cudaStream_t stream[2] = {nullptr, nullptr};
    void* raw[2] = {nullptr, nullptr};
    uint64_t* keys[2] = {nullptr, nullptr};
    uint32_t* block[2] = {nullptr, nullptr};
    double* table[2] = {nullptr, nullptr};
    uint32_t* n_distinct[2] = {nullptr, nullptr};
    uint32_t* col_codes = nullptr;
    ~EncodeScratch() {
        for (int s = 0; s < 2; ++s) {
            if (stream[s]) {
                cudaStreamSynchronize(stream[s]);
                cudaStreamDestroy(stream[s]);
            }
            release(raw[s]);
            release(keys[s]);
            release(block[s]);
            release(table[s]);
            release(n_distinct[s]);
        }
        release(col_codes);
    }
};
}  // namespace

int rfmc_dataset_encode(int device, int64_t n_rows, int32_t n_features, const void* const* col_ptr,
                        const int32_t* col_dtype, const int32_t* col_tag, const int32_t* cat_nunique,
                        const uint8_t* labels, int32_t n_classes, rfmc_dataset** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && n_features > 0 && n_features <= 65535, "bad matrix shape");
    RFMC_REQUIRE(n_classes >= 1 && n_classes <= RFMC_MAX_CLASSES, "n_classes must be in 1..64");
    RFMC_REQUIRE(col_ptr && col_dtype && col_tag && cat_nunique && labels, "NULL argument");
    for (int c = 0; c < n_features; ++c) RFMC_REQUIRE(dtype_size(col_dtype[c]) != 0 && col_ptr[c], "bad column dtype / pointer");
    RFMC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RFMC_CUDA(cudaGetDeviceProperties(&prop, device));
    const int64_t padded = rank_padded_rows(n_rows);
    const int nb = rank_scan_blocks(padded);
    EncodeScratch sc;
    for (int s = 0; s < 2; ++s) {
        RFMC_CUDA(cudaStreamCreateWithFlags(&sc.stream[s], cudaStreamNonBlocking));
        RFMC_CUDA(dev_alloc(&sc.raw[s], (size_t)n_rows * 8));
        RFMC_CUDA(dev_alloc((void**)&sc.keys[s], (size_t)padded * 8));
        RFMC_CUDA(dev_alloc((void**)&sc.block[s], (size_t)nb * 4));
        RFMC_CUDA(dev_alloc((void**)&sc.table[s], (size_t)n_rows * 8));
        RFMC_CUDA(dev_alloc((void**)&sc.n_distinct[s], 4));
    }
    RFMC_CUDA(dev_alloc((void**)&sc.col_codes, (size_t)n_features * n_rows * 4));

    std::vector<std::vector<double>> tables(n_features);
    std::vector<int32_t> nuniq(n_features, 0);
    int pending[2] = {-1, -1};
    auto harvest = [&](int s) -> int {  // wait for the column in flight on stream s and fetch its value table
        const int c = pending[s];
        if (c < 0) return 0;
        pending[s] = -1;
        RFMC_CUDA(cudaStreamSynchronize(sc.stream[s]));
        uint32_t m = 0;
        RFMC_CUDA(cudaMemcpy(&m, sc.n_distinct[s], 4, cudaMemcpyDeviceToHost));
        tables[c].resize(m);
        if (m) RFMC_CUDA(cudaMemcpy(tables[c].data(), sc.table[s], (size_t)m * 8, cudaMemcpyDeviceToHost));
        nuniq[c] = (int32_t)m;
        return 0;
    };
    int turn = 0;
    for (int c = 0; c < n_features; ++c) {
        if (col_dtype[c] == RFMC_DT_CODE) continue;
        const int s = turn;
        turn ^= 1;
        if (harvest(s)) return 1;
        RFMC_CUDA(cudaMemcpyAsync(sc.raw[s], col_ptr[c], (size_t)n_rows * dtype_size(col_dtype[c]), cudaMemcpyHostToDevice,
                                  sc.stream[s]));
        if (launch_rank_column(sc.raw[s], col_dtype[c], n_rows, padded, sc.keys[s], sc.block[s], sc.table[s],
                               sc.n_distinct[s], sc.col_codes + (size_t)c * n_rows, sc.stream[s]))
            return 1;
        pending[s] = c;
    }
    for (int c = 0; c < n_features; ++c) {  // categorical columns: final codes ranked on the host (strings stay there)
        if (col_dtype[c] != RFMC_DT_CODE) continue;
        RFMC_REQUIRE(cat_nunique[c] >= 0, "cat_nunique must be >= 0");
        nuniq[c] = cat_nunique[c];
        RFMC_CUDA(cudaMemcpyAsync(sc.col_codes + (size_t)c * n_rows, col_ptr[c], (size_t)n_rows * 4, cudaMemcpyHostToDevice,
                                  sc.stream[turn]));
    }
    if (harvest(0) || harvest(1)) return 1;
    RFMC_CUDA(cudaStreamSynchronize(sc.stream[0]));
    RFMC_CUDA(cudaStreamSynchronize(sc.stream[1]));

    int32_t top = 1;
    for (int c = 0; c < n_features; ++c) top = nuniq[c] > top ? nuniq[c] : top;
    const int width = top <= 0xFF ? 1 : top <= 0xFFFF ? 2 : 4;
    int64_t words = ((int64_t)n_features * width + 3) / 4;
    if (words % 2 == 0) ++words;  // odd number of 32-bit words per row: conflict-free shared-memory tiles
    const int64_t stride = words * 4 / width;

    rfmc_dataset* ds = new rfmc_dataset();
    ds->device = device;
    ds->code_width = width;
    ds->n_rows = n_rows;
    ds->n_features = n_features;
    ds->stride = stride;
    ds->n_classes = n_classes;
    ds->sm_count = prop.multiProcessorCount;
    ds->h_nuniq = nuniq;
    ds->h_taboff.assign(n_features + 1, 0);
    std::vector<uint8_t> kind(n_features);
    for (int c = 0; c < n_features; ++c) {
        kind[c] = col_dtype[c] == RFMC_DT_CODE ? RFMC_KIND_CATEGORICAL : RFMC_KIND_NUMERIC;
        ds->h_taboff[c + 1] = ds->h_taboff[c] + (int64_t)tables[c].size();
    }
    ds->h_tables.reserve((size_t)ds->h_taboff[n_features] + 1);
    for (int c = 0; c < n_features; ++c) ds->h_tables.insert(ds->h_tables.end(), tables[c].begin(), tables[c].end());
    int rc = 0;
    rc |= reserve((uint8_t**)&ds->d_codes, (size_t)n_rows * stride * width);
    rc |= upload(&ds->d_labels, labels, (size_t)n_rows, sc.stream[0]);
    rc |= upload(&ds->d_kind, kind.data(), (size_t)n_features, sc.stream[0]);
    rc |= upload(&ds->d_tag, col_tag, (size_t)n_features, sc.stream[0]);
    rc |= upload(&ds->d_nuniq, nuniq.data(), (size_t)n_features, sc.stream[0]);
    rc |= upload(&ds->d_taboff, ds->h_taboff.data(), (size_t)n_features + 1, sc.stream[0]);
    if (ds->h_tables.empty())
        rc |= reserve(&ds->d_tables, 1);
    else
        rc |= upload(&ds->d_tables, ds->h_tables.data(), ds->h_tables.size(), sc.stream[0]);
    if (rc == 0) rc = launch_interleave(sc.col_codes, n_rows, n_features, stride, width, ds->d_codes, sc.stream[0]);
    if (rc == 0 && cudaStreamSynchronize(sc.stream[0]) != cudaSuccess) {
        set_error(std::string("dataset encode failed: ") + cudaGetErrorString(cudaGetLastError()));
        rc = 1;
    }
    if (rc) {
        rfmc_dataset_destroy(ds);
        return 1;
    }
    *out = ds;
    return 0;
}

int rfmc_dataset_describe(rfmc_dataset* ds, int32_t* code_width, int64_t* row_stride, int32_t* col_nunique,
                          int64_t* table_off) {
    RFMC_REQUIRE(ds != nullptr, "dataset is NULL");
    RFMC_REQUIRE((int)ds->h_nuniq.size() == ds->n_features, "dataset was not built by rfmc_dataset_encode");
    if (code_width) *code_width = ds->code_width;
    if (row_stride) *row_stride = ds->stride;
    if (col_nunique) memcpy(col_nunique, ds->h_nuniq.data(), sizeof(int32_t) * ds->n_features);
    if (table_off) memcpy(table_off, ds->h_taboff.data(), sizeof(int64_t) * (ds->n_features + 1));
    return 0;
}

int rfmc_dataset_tables(rfmc_dataset* ds, double* tables_out) {
    RFMC_REQUIRE(ds != nullptr && tables_out != nullptr, "NULL argument");
    RFMC_REQUIRE((int)ds->h_nuniq.size() == ds->n_features, "dataset was not built by rfmc_dataset_encode");
    if (!ds->h_tables.empty()) memcpy(tables_out, ds->h_tables.data(), sizeof(double) * ds->h_tables.size());
    return 0;
}

int rfmc_dataset_codes(rfmc_dataset* ds, void* codes_out) {
    RFMC_REQUIRE(ds != nullptr && codes_out != nullptr, "NULL argument");
    RFMC_CUDA(cudaSetDevice(ds->device));
    RFMC_CUDA(cudaMemcpy(codes_out, ds->d_codes, (size_t)ds->n_rows * ds->stride * ds->code_width, cudaMemcpyDeviceToHost));
    return 0;
}

int rfmc_dataset_destroy(rfmc_dataset* ds) {
    if (!ds) return 0;
    cudaSetDevice(ds->device);
    release(ds->d_codes);
    release(ds->d_labels);
    release(ds->d_kind);
    release(ds->d_tag);
    release(ds->d_nuniq);
    release(ds->d_taboff);
    release(ds->d_tables);
    release(ds->d_class_n);
    release(ds->d_class_aoff);
    release(ds->d_class_rows);
    release(ds->d_class_roff);
    delete ds;
    return 0;
}

static int batch_create_impl(rfmc_dataset* ds, int32_t n_slots, int32_t n_train, int32_t n_val, const int32_t* idx_train,
                             const int32_t* idx_val, bool idx_on_device, int32_t n_cand, const int32_t* cand_slot,
                             const int32_t* feat_off, const uint16_t* feat_list, int32_t max_depth, int32_t min_samples_split,
                             void* stream, rfmc_batch** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(ds != nullptr, "dataset is NULL");
    RFMC_REQUIRE(n_slots > 0 && n_train > 0 && n_val >= 0 && n_cand > 0, "bad batch shape");
    RFMC_REQUIRE(min_samples_split >= 1, "min_samples_split must be >= 1");
    RFMC_REQUIRE(max_depth >= 1, "max_depth must be >= 1");
    int32_t fmax = 0;
    for (int i = 0; i < n_cand; ++i) {
        int32_t nf = feat_off[i + 1] - feat_off[i];
        RFMC_REQUIRE(nf >= 1, "every candidate needs at least one feature");
        RFMC_REQUIRE(cand_slot[i] >= 0 && cand_slot[i] < n_slots, "cand_slot out of range");
        if (nf > fmax) fmax = nf;
    }
    for (int64_t i = 0; i < feat_off[n_cand]; ++i)
        RFMC_REQUIRE(feat_list[i] < ds->n_features, "feature index out of range");
    RFMC_CUDA(cudaSetDevice(ds->device));
    cudaStream_t st = (cudaStream_t)stream;
    rfmc_batch* b = new rfmc_batch();
    b->ds = ds;
    b->n_slots = n_slots;
    b->n_train = n_train;
    b->n_val = n_val;
    b->n_cand = n_cand;
    b->max_depth = max_depth;
    b->mss = min_samples_split;
    b->nt_pad = (n_train + 15) / 16 * 16;
    b->max_nodes = 2 * n_train;
    b->fmax = fmax;
    b->grid = grow_grid_size(ds, n_cand, n_train);
    b->stream = st;
    const size_t g = (size_t)b->grid, ntp = (size_t)b->nt_pad, C = (size_t)ds->n_classes;
    int rc = 0;
    {
        // All inputs leave through ONE pinned staging buffer and one copy.  A cudaMemcpyAsync from pageable memory
        // is not asynchronous: the driver stages it itself, under a lock shared by every host thread of the fit, and
        // the call waits for whatever the stream is still running (the previous plan's kernels) before it returns.
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t b_tr = idx_on_device ? 0 : al((size_t)n_slots * n_train * 4), b_va = idx_on_device ? 0 : al((size_t)n_slots * n_val * 4);
        const size_t b_cs = al((size_t)n_cand * 4), b_fo = al(((size_t)n_cand + 1) * 4), b_fl = al((size_t)feat_off[n_cand] * 2);
        const size_t total = b_tr + b_va + b_cs + b_fo + b_fl;
        PinnedLease* lease = new PinnedLease(total);
        uint8_t* d_in = nullptr;
        if (lease->p != nullptr && dev_alloc((void**)&d_in, total) == cudaSuccess) {
            uint8_t* h = (uint8_t*)lease->p;
            size_t at = 0;
            if (!idx_on_device) {
                std::memcpy(h + at, idx_train, (size_t)n_slots * n_train * 4);
                b->d_idx_train = (int32_t*)(d_in + at);
                at += b_tr;
                if (n_val) std::memcpy(h + at, idx_val, (size_t)n_slots * n_val * 4);
                b->d_idx_val = (int32_t*)(d_in + at);
                at += b_va;
            }
            std::memcpy(h + at, cand_slot, (size_t)n_cand * 4);
            b->d_cand_slot = (int32_t*)(d_in + at);
            at += b_cs;
            std::memcpy(h + at, feat_off, ((size_t)n_cand + 1) * 4);
            b->d_feat_off = (int32_t*)(d_in + at);
            at += b_fo;
            std::memcpy(h + at, feat_list, (size_t)feat_off[n_cand] * 2);
            b->d_feat_list = (uint16_t*)(d_in + at);
            b->d_inputs = d_in;
            b->up_lease = lease;
            if (cudaMemcpyAsync(d_in, h, total, cudaMemcpyHostToDevice, st) != cudaSuccess) {
                set_error(std::string("batch upload failed: ") + cudaGetErrorString(cudaGetLastError()));
                rc = 1;
            }
            if (idx_on_device) {
                b->d_idx_train = const_cast<int32_t*>(idx_train);
                b->d_idx_val = const_cast<int32_t*>(idx_val);
            }
            b->own_idx = false;  // slices of d_inputs, or the plan's arrays
        } else {
            delete lease;
            cudaGetLastError();
            if (idx_on_device) {
                b->own_idx = false;
                b->d_idx_train = const_cast<int32_t*>(idx_train);
                b->d_idx_val = const_cast<int32_t*>(idx_val);
            } else {
                rc |= upload(&b->d_idx_train, idx_train, (size_t)n_slots * n_train, st);
                rc |= upload(&b->d_idx_val, idx_val, (size_t)n_slots * n_val, st);
            }
            rc |= upload(&b->d_cand_slot, cand_slot, (size_t)n_cand, st);
            rc |= upload(&b->d_feat_off, feat_off, (size_t)n_cand + 1, st);
            rc |= upload(&b->d_feat_list, feat_list, (size_t)feat_off[n_cand], st);
        }
    }
    rc |= reserve((uint8_t**)&b->d_cc, (size_t)n_slots * ds->n_features * ntp * ds->code_width);  // [slot][feature][row]
    rc |= reserve(&b->d_lab, (size_t)n_slots * ntp);
    rc |= reserve(&b->d_pos, g * 2 * ntp);
    rc |= reserve(&b->d_work, g * 2 * ntp);
    rc |= reserve(&b->d_resna, g * 2 * ntp);
    rc |= reserve(&b->d_rank, g * ntp);
    rc |= reserve(&b->d_resfin, g * ntp);
    rc |= reserve(&b->d_nodes, (size_t)n_cand * b->max_nodes);
    rc |= reserve(&b->d_counts, (size_t)n_cand * b->max_nodes * C);
    rc |= reserve(&b->d_votes, (size_t)n_cand * b->max_nodes);
    rc |= reserve(&b->d_nnodes, (size_t)n_cand);
    rc |= reserve(&b->d_correct, (size_t)n_cand);
    if (rc) {
        rfmc_batch_destroy(b);
        return 1;
    }
    *out = b;
    return 0;
}

int rfmc_batch_create(rfmc_dataset* ds, int32_t n_slots, int32_t n_train, int32_t n_val, const int32_t* idx_train,
                      const int32_t* idx_val, int32_t n_cand, const int32_t* cand_slot, const int32_t* feat_off,
                      const uint16_t* feat_list, int32_t max_depth, int32_t min_samples_split, void* stream,
                      rfmc_batch** out) {
    return batch_create_impl(ds, n_slots, n_train, n_val, idx_train, idx_val, false, n_cand, cand_slot, feat_off, feat_list,
                             max_depth, min_samples_split, stream, out);
}

int rfmc_batch_create_from_plan(rfmc_dataset* ds, rfmc_drawplan* pl, int32_t n_cand, const int32_t* cand_slot,
                                const int32_t* feat_off, const uint16_t* feat_list, int32_t max_depth,
                                int32_t min_samples_split, void* stream, rfmc_batch** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(pl != nullptr && pl->ds == ds, "plan is NULL or belongs to another dataset");
    RFMC_REQUIRE((cudaStream_t)stream == pl->stream, "the batch must run on the plan's stream");
    return batch_create_impl(ds, pl->n_slots, (int32_t)(pl->n_classes * pl->n_tr), (int32_t)(pl->n_classes * pl->n_va),
                             pl->d_idx_train, pl->d_idx_val, true, n_cand, cand_slot, feat_off, feat_list, max_depth,
                             min_samples_split, stream, out);
}

int rfmc_batch_run(rfmc_batch* b, void* stream) {
    RFMC_REQUIRE(b != nullptr, "batch is NULL");
    RFMC_CUDA(cudaSetDevice(b->ds->device));
    b->stream = (cudaStream_t)stream;
    b->have_results = false;
    return launch_grow(b, b->stream);
}

int rfmc_batch_launches(rfmc_batch* b) { return b ? b->launches : 0; }

int rfmc_batch_set_timing(rfmc_batch* b, int enabled) {
    RFMC_REQUIRE(b != nullptr, "batch is NULL");
    RFMC_CUDA(cudaSetDevice(b->ds->device));
    if (enabled && !b->ev[0])
        for (int i = 0; i < 3; ++i) RFMC_CUDA(cudaEventCreate(&b->ev[i]));
    b->timing = enabled != 0;
    return 0;
}

int rfmc_batch_kernel_ms(rfmc_batch* b, float* ms_gather, float* ms_grow) {
    RFMC_REQUIRE(b != nullptr && ms_gather != nullptr && ms_grow != nullptr, "NULL argument");
    RFMC_REQUIRE(b->timing && b->ev[0], "timing was not enabled before the run");
    RFMC_CUDA(cudaSetDevice(b->ds->device));
    RFMC_CUDA(cudaEventSynchronize(b->ev[2]));
    RFMC_CUDA(cudaEventElapsedTime(ms_gather, b->ev[0], b->ev[1]));
    RFMC_CUDA(cudaEventElapsedTime(ms_grow, b->ev[1], b->ev[2]));
    return 0;
}

int rfmc_batch_results(rfmc_batch* b, int32_t* correct, int32_t* n_nodes) {
    RFMC_REQUIRE(b != nullptr, "batch is NULL");
    RFMC_CUDA(cudaSetDevice(b->ds->device));
    b->h_nnodes.resize(b->n_cand);
    RFMC_CUDA(cudaMemcpyAsync(b->h_nnodes.data(), b->d_nnodes, sizeof(int32_t) * b->n_cand, cudaMemcpyDeviceToHost, b->stream));
    if (correct)
        RFMC_CUDA(cudaMemcpyAsync(correct, b->d_correct, sizeof(int32_t) * b->n_cand, cudaMemcpyDeviceToHost, b->stream));
    RFMC_CUDA(cudaStreamSynchronize(b->stream));
    if (n_nodes) std::memcpy(n_nodes, b->h_nnodes.data(), sizeof(int32_t) * b->n_cand);
    b->have_results = true;
    return 0;
}

int rfmc_batch_pack(rfmc_batch* b, const int32_t* cand_ids, int32_t n_sel, rfmc_node* nodes_out, int32_t* counts_out,
                    int dst_on_device, void* stream) {
    RFMC_REQUIRE(b != nullptr, "batch is NULL");
    if (n_sel <= 0) return 0;
    RFMC_CUDA(cudaSetDevice(b->ds->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!b->have_results) {
        if (rfmc_batch_results(b, nullptr, nullptr)) return 1;
    }
    std::vector<int64_t> off(n_sel + 1, 0);
    for (int i = 0; i < n_sel; ++i) {
        RFMC_REQUIRE(cand_ids[i] >= 0 && cand_ids[i] < b->n_cand, "candidate id out of range");
        off[i + 1] = off[i] + b->h_nnodes[cand_ids[i]];
    }
    const size_t total = (size_t)off[n_sel], C = (size_t)b->ds->n_classes;
    int32_t* d_ids = nullptr;
    int64_t* d_off = nullptr;
    rfmc_node* d_nodes = nodes_out;
    int32_t* d_counts = counts_out;
    int rc = 0;
    rc |= upload(&d_ids, cand_ids, (size_t)n_sel, st);
    rc |= upload(&d_off, off.data(), (size_t)n_sel, st);
    if (!dst_on_device) {
        rc |= reserve(&d_nodes, total);
        rc |= reserve(&d_counts, total * C);
    }
    if (!rc) rc = launch_pack(b, d_ids, d_off, n_sel, d_nodes, d_counts, st);
    if (!rc && !dst_on_device) {
        // device -> pinned staging (full-rate DMA) -> caller's pageable buffers
        const size_t nb = total * sizeof(rfmc_node), cb = total * C * sizeof(int32_t);
        PinnedLease lease(nb + cb);
        uint8_t* pin = (uint8_t*)lease.p;
        void* h_nodes = pin ? (void*)pin : (void*)nodes_out;
        void* h_counts = pin ? (void*)(pin + nb) : (void*)counts_out;
        if (cudaMemcpyAsync(h_nodes, d_nodes, nb, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_counts, d_counts, cb, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            set_error("pack: device to host copy failed");
            rc = 1;
        }
        if (cudaStreamSynchronize(st) != cudaSuccess && !rc) {
            set_error(std::string("pack failed: ") + cudaGetErrorString(cudaGetLastError()));
            rc = 1;
        }
        if (!rc && pin) {
            std::memcpy(nodes_out, h_nodes, nb);
            std::memcpy(counts_out, h_counts, cb);
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) {
        set_error(std::string("pack failed: ") + cudaGetErrorString(cudaGetLastError()));
        rc = 1;
    }
    release(d_ids);
    release(d_off);
    if (!dst_on_device) {
        release(d_nodes);
        release(d_counts);
    }
    b->launches += 1;
    return rc;
}

int rfmc_batch_destroy(rfmc_batch* b) {
    if (!b) return 0;
    cudaSetDevice(b->ds->device);
    // blocks go back to a process-wide pool: nothing of this batch may still be in flight (error paths,
    // a close() before results())
    cudaStreamSynchronize(b->stream);
    if (b->own_idx) {
        release(b->d_idx_train);
        release(b->d_idx_val);
    }
    if (b->d_inputs) {
        release(b->d_inputs);
        delete (PinnedLease*)b->up_lease;
    } else {
        release(b->d_cand_slot);
        release(b->d_feat_off);
        release(b->d_feat_list);
    }
    release(b->d_cc);
    release(b->d_lab);
    release(b->d_pos);
    release(b->d_work);
    release(b->d_resna);
    release(b->d_rank);
    release(b->d_resfin);
    release(b->d_nodes);
    release(b->d_counts);
    release(b->d_votes);
    release(b->d_nnodes);
    release(b->d_correct);
    for (int i = 0; i < 3; ++i)
        if (b->ev[i]) cudaEventDestroy(b->ev[i]);
    delete b;
    return 0;
}

int rfmc_forest_create(int device, const rfmc_pnode* nodes, int64_t n_nodes, const int64_t* tree_root, int32_t n_trees,
                       int64_t n_payloads, const uint8_t* vote, const uint8_t* vote_absent, const double* probs,
                       const double* probs_absent, int32_t n_classes, const double* scores, double score_fsum,
                       rfmc_forest** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(n_nodes > 0 && n_trees > 0 && n_payloads > 0, "empty forest");
    RFMC_REQUIRE(n_nodes < 0xFFFFFFFFll, "forest too large for 32-bit node indices");
    RFMC_REQUIRE(n_classes >= 1 && n_classes <= RFMC_MAX_CLASSES, "n_classes must be in 1..64");
    RFMC_CUDA(cudaSetDevice(device));
    rfmc_forest* f = new rfmc_forest();
    f->device = device;
    f->n_nodes = n_nodes;
    f->n_trees = n_trees;
    f->n_payloads = n_payloads;
    f->n_classes = n_classes;
    f->score_fsum = score_fsum;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) f->sm_count = prop.multiProcessorCount;
    int rc = 0;
    rc |= upload(&f->d_nodes, nodes, (size_t)n_nodes, 0);
    rc |= upload(&f->d_root, tree_root, (size_t)n_trees, 0);
    rc |= upload(&f->d_vote, vote, (size_t)n_payloads, 0);
    rc |= upload(&f->d_vote_absent, vote_absent, (size_t)n_payloads, 0);
    rc |= upload(&f->d_probs, probs, (size_t)n_payloads * n_classes, 0);
    rc |= upload(&f->d_probs_absent, probs_absent, (size_t)n_payloads * n_classes, 0);
    rc |= upload(&f->d_scores, scores, (size_t)n_trees, 0);
    std::vector<uint32_t> blocks, root_b, block_of, tops;  // must outlive the asynchronous uploads below
    if (!rc && build_blocks(nodes, n_nodes, tree_root, n_trees, blocks, root_b, block_of)) {
        f->n_blocks = (int64_t)(blocks.size() / 8);
        rc |= upload((uint32_t**)&f->d_blocks, blocks.data(), blocks.size(), 0);
        rc |= upload(&f->d_root_b, root_b.data(), root_b.size(), 0);
        // staged levels: enough for the average tree, at most ten (8 KB per tree), always even
        int levels = 2;
        while (levels < 10 && ((int64_t)1 << levels) < 2 * (n_nodes / n_trees + 1)) levels += 2;
        if (const char* e = getenv("RFMC_PRED_TOP_LEVELS")) {
            const int want = atoi(e) & ~1;
            if (want >= 2 && want <= 12) levels = want;
        }
        build_tops(nodes, tree_root, n_trees, levels, block_of, tops);
        f->top_levels = levels;
        rc |= upload((uint32_t**)&f->d_tops, tops.data(), tops.size(), 0);
    }
    std::vector<uint32_t> crec, crec4, cpay, ctree, cstage;
    std::vector<uint8_t> cdepth;
    if (!rc && build_compact(nodes, tree_root, n_trees, crec, crec4, cpay, ctree, cdepth, cstage)) {
        rc |= upload((uint32_t**)&f->d_crec, crec.data(), crec.size(), 0);
        if (!crec4.empty()) rc |= upload(&f->d_crec4, crec4.data(), crec4.size(), 0);
        rc |= upload(&f->d_cpay, cpay.data(), cpay.size(), 0);
        rc |= upload(&f->d_ctree, ctree.data(), ctree.size(), 0);
        rc |= upload(&f->d_cdepth, cdepth.data(), cdepth.size(), 0);
        rc |= upload(&f->d_cstage, cstage.data(), cstage.size(), 0);
        f->n_stages = (int32_t)cstage.size() - 1;
    }
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) {
        set_error("forest upload failed");
        rc = 1;
    }
    if (rc) {
        rfmc_forest_destroy(f);
        return 1;
    }
    *out = f;
    return 0;
}

/* test hook (not in the header; host code only, no device needed): the compact small-tree layout of a forest
   (build_compact above).  Pass NULL buffers to query the sizes; returns 1 when the forest is not eligible. */
int rfmc_debug_compact_layout(const rfmc_pnode* nodes, const int64_t* tree_root, int32_t n_trees, int64_t* n_records,
                              int32_t* n_stages, uint32_t* rec_xy, uint32_t* pay, uint32_t* tree_at, uint8_t* depth,
                              uint32_t* stage, uint32_t* rec4, int32_t* has_rec4) {
    std::vector<uint32_t> crec, crec4, cpay, ctree, cstage;
    std::vector<uint8_t> cdepth;
    if (!nodes || !tree_root || n_trees <= 0) return 1;
    if (!build_compact(nodes, tree_root, n_trees, crec, crec4, cpay, ctree, cdepth, cstage)) return 1;
    if (n_records) *n_records = (int64_t)cpay.size();
    if (n_stages) *n_stages = (int32_t)cstage.size() - 1;
    if (has_rec4) *has_rec4 = crec4.empty() ? 0 : 1;
    if (rec4 && !crec4.empty()) std::memcpy(rec4, crec4.data(), crec4.size() * 4);
    if (rec_xy) std::memcpy(rec_xy, crec.data(), crec.size() * 4);
    if (pay) std::memcpy(pay, cpay.data(), cpay.size() * 4);
    if (tree_at) std::memcpy(tree_at, ctree.data(), ctree.size() * 4);
    if (depth) std::memcpy(depth, cdepth.data(), cdepth.size());
    if (stage) std::memcpy(stage, cstage.data(), cstage.size() * 4);
    return 0;
}

int rfmc_cache_trim(int device, int64_t* freed_bytes) {
    RFMC_CUDA(cudaSetDevice(device));
    size_t freed = 0;
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        DevCache& dc = g_mem[device];
        freed += dc.cached;
        flush_cache_locked(dc);
    }
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        for (auto& kv : g_pin_free) {
            freed += kv.first;
            cudaFreeHost(kv.second);
        }
        g_pin_free.clear();
        g_pin_cached = 0;
    }
    if (freed_bytes) *freed_bytes = (int64_t)freed;
    return 0;
}

int rfmc_forest_layout(rfmc_forest* f, int32_t* small_tree_stages) {
    RFMC_REQUIRE(f != nullptr && small_tree_stages != nullptr, "NULL argument");
    *small_tree_stages = f->d_crec != nullptr ? f->n_stages : 0;
    return 0;
}

int rfmc_forest_destroy(rfmc_forest* f) {
    if (!f) return 0;
    cudaSetDevice(f->device);
    release(f->d_nodes);
    release(f->d_root);
    release(f->d_vote);
    release(f->d_vote_absent);
    release(f->d_probs);
    release(f->d_probs_absent);
    release(f->d_scores);
    pipe_destroy(f->pipe);
    release(f->d_blocks);
    release(f->d_root_b);
    release(f->d_tops);
    release(f->d_crec);
    release(f->d_crec4);
    release(f->d_cpay);
    release(f->d_ctree);
    release(f->d_cdepth);
    release(f->d_cstage);
    release(f->d_thresholds);
    release(f->d_thr_off);
    delete f;
    return 0;
}

int rfmc_forest_predict(rfmc_forest* f, const void* codes, int code_width, int64_t n_rows, int64_t row_stride,
                        int32_t n_features, const uint8_t* absent, int soft, int weighted, double* probs_out,
                        int32_t* labels_out, int on_device, void* stream) {
    RFMC_REQUIRE(f != nullptr, "forest is NULL");
    RFMC_REQUIRE(code_width == 1 || code_width == 2 || code_width == 4, "code_width must be 1, 2 or 4");
    RFMC_REQUIRE(n_rows >= 0 && n_features > 0 && row_stride >= n_features, "bad matrix shape");
    RFMC_REQUIRE((row_stride * code_width) % 4 == 0, "row pitch must be a multiple of 4 bytes");
    if (n_rows == 0) return 0;
    RFMC_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    bool has_absent = false;
    if (absent)
        for (int i = 0; i < n_features; ++i) has_absent |= absent[i] != 0;
    uint8_t* d_absent = nullptr;
    int rc = 0;
    if (has_absent) rc |= upload(&d_absent, absent, (size_t)n_features, st);
    const size_t C = (size_t)f->n_classes;
    if (on_device) {
        if (!rc)
            rc = launch_predict(f, codes, code_width, n_rows, row_stride, n_features, d_absent, has_absent, soft, weighted,
                                probs_out, labels_out, st);
        if (d_absent) {
            cudaStreamSynchronize(st);
            release(d_absent);
        }
        return rc;
    }
    // Host buffers: chunked, double-buffered pipeline.  While chunk c votes on one stream, chunk
    // c+1 is staged through pinned memory and copied in on the other; results leave through
    // pinned memory as well, so no copy in the timed path is a blocking pageable transfer.
    if (rc) {
        release(d_absent);
        return rc;
    }
    cudaStreamSynchronize(st);
    PipeLease lease(f);
    PredictPipe* pp = lease.get();
    const size_t row_bytes = (size_t)row_stride * code_width;
    const int64_t chunk_rows = 32768;
    if (pipe_reserve(pp, chunk_rows * row_bytes, probs_out ? chunk_rows * C * sizeof(double) : 0,
                     labels_out ? chunk_rows * sizeof(int32_t) : 0)) {
        release(d_absent);
        return 1;
    }
    const int64_t n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    auto drain = [&](int64_t c) {  // copy chunk c's results from pinned staging to the caller
        const int sl = (int)(c & 1);
        const int64_t r0 = c * chunk_rows, nr = (n_rows - r0) < chunk_rows ? (n_rows - r0) : chunk_rows;
        if (cudaEventSynchronize(pp->done[sl]) != cudaSuccess) rc = 1;
        if (probs_out) memcpy(probs_out + r0 * C, pp->pin_probs[sl], (size_t)nr * C * sizeof(double));
        if (labels_out) memcpy(labels_out + r0, pp->pin_labels[sl], (size_t)nr * sizeof(int32_t));
    };
    for (int64_t c = 0; c < n_chunks && !rc; ++c) {
        const int sl = (int)(c & 1);
        if (c >= 2) drain(c - 2);
        const int64_t r0 = c * chunk_rows, nr = (n_rows - r0) < chunk_rows ? (n_rows - r0) : chunk_rows;
        memcpy(pp->pin_in[sl], (const uint8_t*)codes + (size_t)r0 * row_bytes, (size_t)nr * row_bytes);
        cudaStream_t cs = pp->st[sl];
        if (cudaMemcpyAsync(pp->d_in[sl], pp->pin_in[sl], (size_t)nr * row_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess) rc = 1;
        if (!rc)
            rc = launch_predict(f, pp->d_in[sl], code_width, nr, row_stride, n_features, d_absent, has_absent, soft, weighted,
                                probs_out ? pp->d_probs[sl] : nullptr, labels_out ? pp->d_labels[sl] : nullptr, cs);
        if (!rc && probs_out &&
            cudaMemcpyAsync(pp->pin_probs[sl], pp->d_probs[sl], (size_t)nr * C * sizeof(double), cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = 1;
        if (!rc && labels_out &&
            cudaMemcpyAsync(pp->pin_labels[sl], pp->d_labels[sl], (size_t)nr * sizeof(int32_t), cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = 1;
        if (!rc && cudaEventRecord(pp->done[sl], cs) != cudaSuccess) rc = 1;
        pp->launches += 1;
    }
    if (!rc) {
        if (n_chunks >= 2) drain(n_chunks - 2);
        drain(n_chunks - 1);
    }
    cudaError_t e0 = cudaStreamSynchronize(pp->st[0]), e1 = cudaStreamSynchronize(pp->st[1]);
    if (e0 != cudaSuccess || e1 != cudaSuccess) {
        set_error(std::string("predict failed: ") + cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
        rc = 1;
    } else if (rc && g_error.empty()) {
        set_error("predict failed");
    }
    release(d_absent);
    return rc;
}

int rfmc_forest_set_codebook(rfmc_forest* f, int32_t n_features, const int64_t* thr_off, const double* thresholds) {
    RFMC_REQUIRE(f != nullptr && thr_off != nullptr, "forest or offsets are NULL");
    RFMC_REQUIRE(n_features > 0, "bad feature count");
    RFMC_CUDA(cudaSetDevice(f->device));
    release(f->d_thresholds);
    release(f->d_thr_off);
    f->d_thresholds = nullptr;
    f->d_thr_off = nullptr;
    int rc = 0;
    const size_t nt = (size_t)thr_off[n_features];
    rc |= upload(&f->d_thr_off, thr_off, (size_t)n_features + 1, 0);
    if (nt)
        rc |= upload(&f->d_thresholds, thresholds, nt, 0);
    else
        rc |= reserve(&f->d_thresholds, 1);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) {
        set_error("codebook upload failed");
        rc = 1;
    }
    f->n_features_book = n_features;
    return rc;
}

int rfmc_forest_predict_frame(rfmc_forest* f, int64_t n_rows, int32_t n_features, int64_t row_stride, int code_width,
                              int32_t n_cols, const void* const* col_ptr, const int32_t* col_feat,
                              const int32_t* col_dtype, const uint8_t* absent, int soft, int weighted,
                              double* probs_out, int32_t* labels_out) {
    RFMC_REQUIRE(f != nullptr, "forest is NULL");
    RFMC_REQUIRE(f->d_thr_off != nullptr && f->n_features_book == n_features, "no code book set for this forest");
    RFMC_REQUIRE(code_width == 1 || code_width == 2 || code_width == 4, "code_width must be 1, 2 or 4");
    RFMC_REQUIRE(n_rows >= 0 && n_features > 0 && row_stride >= n_features && n_cols >= 0, "bad frame shape");
    RFMC_REQUIRE((row_stride * code_width) % 4 == 0, "row pitch must be a multiple of 4 bytes");
    if (n_rows == 0) return 0;
    RFMC_CUDA(cudaSetDevice(f->device));
    const int64_t chunk_rows = 32768;
    std::vector<int64_t> col_off((size_t)n_cols > 0 ? n_cols : 1, 0);
    size_t raw_bytes = 0;
    for (int c = 0; c < n_cols; ++c) {
        const size_t es = dtype_size(col_dtype[c]);
        RFMC_REQUIRE(es != 0, "unknown column dtype");
        RFMC_REQUIRE(col_feat[c] >= 0 && col_feat[c] < n_features, "column feature index out of range");
        col_off[c] = (int64_t)raw_bytes;
        raw_bytes += ((size_t)chunk_rows * es + 15) / 16 * 16;
    }
    if (raw_bytes == 0) raw_bytes = 16;
    bool has_absent = false;
    if (absent)
        for (int i = 0; i < n_features; ++i) has_absent |= absent[i] != 0;
    const size_t C = (size_t)f->n_classes;
    const size_t row_bytes = (size_t)row_stride * code_width;
    PipeLease lease(f);
    PredictPipe* pp = lease.get();
    if (pipe_reserve(pp, chunk_rows * row_bytes, probs_out ? chunk_rows * C * sizeof(double) : 0,
                     labels_out ? chunk_rows * sizeof(int32_t) : 0) ||
        pipe_reserve_raw(pp, raw_bytes))
        return 1;
    uint8_t* d_absent = nullptr;
    int64_t* d_col_off = nullptr;
    int32_t* d_col_feat = nullptr;
    int32_t* d_col_dtype = nullptr;
    int rc = 0;
    if (has_absent) rc |= upload(&d_absent, absent, (size_t)n_features, 0);
    rc |= upload(&d_col_off, col_off.data(), col_off.size(), 0);
    if (n_cols) {
        rc |= upload(&d_col_feat, col_feat, (size_t)n_cols, 0);
        rc |= upload(&d_col_dtype, col_dtype, (size_t)n_cols, 0);
    }
    cudaStreamSynchronize(0);
    const int64_t n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    auto drain = [&](int64_t c) {
        const int sl = (int)(c & 1);
        const int64_t r0 = c * chunk_rows, nr = (n_rows - r0) < chunk_rows ? (n_rows - r0) : chunk_rows;
        if (cudaEventSynchronize(pp->done[sl]) != cudaSuccess) rc = 1;
        if (probs_out) memcpy(probs_out + r0 * C, pp->pin_probs[sl], (size_t)nr * C * sizeof(double));
        if (labels_out) memcpy(labels_out + r0, pp->pin_labels[sl], (size_t)nr * sizeof(int32_t));
    };
    for (int64_t c = 0; c < n_chunks && !rc; ++c) {
        const int sl = (int)(c & 1);
        if (c >= 2) drain(c - 2);
        const int64_t r0 = c * chunk_rows, nr = (n_rows - r0) < chunk_rows ? (n_rows - r0) : chunk_rows;
        for (int k = 0; k < n_cols; ++k) {
            const size_t es = dtype_size(col_dtype[k]);
            memcpy(pp->pin_raw[sl] + col_off[k], (const uint8_t*)col_ptr[k] + (size_t)r0 * es, (size_t)nr * es);
        }
        cudaStream_t cs = pp->st[sl];
        if (cudaMemcpyAsync(pp->d_raw[sl], pp->pin_raw[sl], raw_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess) rc = 1;
        if (!rc)
            rc = launch_encode(pp->d_raw[sl], d_col_off, d_col_feat, d_col_dtype, n_cols, nr, row_stride, code_width,
                               f->d_thresholds, f->d_thr_off, pp->d_in[sl], cs);
        if (!rc)
            rc = launch_predict(f, pp->d_in[sl], code_width, nr, row_stride, n_features, d_absent, has_absent, soft, weighted,
                                probs_out ? pp->d_probs[sl] : nullptr, labels_out ? pp->d_labels[sl] : nullptr, cs);
        if (!rc && probs_out &&
            cudaMemcpyAsync(pp->pin_probs[sl], pp->d_probs[sl], (size_t)nr * C * sizeof(double), cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = 1;
        if (!rc && labels_out &&
            cudaMemcpyAsync(pp->pin_labels[sl], pp->d_labels[sl], (size_t)nr * sizeof(int32_t), cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = 1;
        if (!rc && cudaEventRecord(pp->done[sl], cs) != cudaSuccess) rc = 1;
        pp->launches += 2;
    }
    if (!rc) {
        if (n_chunks >= 2) drain(n_chunks - 2);
        drain(n_chunks - 1);
    }
    cudaError_t e0 = cudaStreamSynchronize(pp->st[0]), e1 = cudaStreamSynchronize(pp->st[1]);
    if (e0 != cudaSuccess || e1 != cudaSuccess) {
        set_error(std::string("predict_frame failed: ") + cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
        rc = 1;
    } else if (rc && g_error.empty()) {
        set_error("predict_frame failed");
    }
    release(d_absent);
    release(d_col_off);
    release(d_col_feat);
    release(d_col_dtype);
    return rc;
}

int rfmc_forest_leaves(rfmc_forest* f, const void* codes, int code_width, int64_t n_rows, int64_t row_stride,
                       int32_t n_features, const uint8_t* absent, int32_t* leaves_out, void* stream) {
    RFMC_REQUIRE(f != nullptr && leaves_out != nullptr, "forest or output is NULL");
    RFMC_REQUIRE(code_width == 1 || code_width == 2 || code_width == 4, "code_width must be 1, 2 or 4");
    RFMC_REQUIRE(n_rows >= 0 && n_features > 0 && row_stride >= n_features, "bad matrix shape");
    if (n_rows == 0) return 0;
    RFMC_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    bool has_absent = false;
    if (absent)
        for (int i = 0; i < n_features; ++i) has_absent |= absent[i] != 0;
    uint8_t* d_absent = nullptr;
    uint8_t* d_codes = nullptr;
    int32_t* d_out = nullptr;
    int rc = 0;
    if (has_absent) rc |= upload(&d_absent, absent, (size_t)n_features, st);
    rc |= upload(&d_codes, (const uint8_t*)codes, (size_t)n_rows * row_stride * code_width, st);
    rc |= reserve(&d_out, (size_t)n_rows * f->n_trees);
    if (!rc) rc = launch_leaves(f, d_codes, code_width, n_rows, row_stride, n_features, d_absent, d_out, st);
    if (!rc && cudaMemcpyAsync(leaves_out, d_out, (size_t)n_rows * f->n_trees * sizeof(int32_t), cudaMemcpyDeviceToHost,
                               st) != cudaSuccess)
        rc = 1;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error(std::string("leaves failed: ") + cudaGetErrorString(e));
        rc = 1;
    }
    release(d_absent);
    release(d_codes);
    release(d_out);
    return rc;
}

}  // extern "C"
