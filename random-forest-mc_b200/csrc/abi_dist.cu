// Survivor exchange of the rank-sharded fit through the C ABI (include/rfmc_b200.h, section
// "multi-GPU"): what `self.data.extend(Tree_list)` (model.py:403) is when the tree slots were grown by
// several processes, one per GPU.  Only the surviving trees' node tables travel -- one NCCL
// all-gather of sizes, one of the small per-tree records, one each of the node and count tables,
// device to device over NVLink.  NCCL is not linked: its entry points are taken from the library the
// process already loaded (the one that created the communicator), or dlopen'ed by soname.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

using namespace rfmc;

struct rfmc_gathered {
    int device = 0, world = 0, rank = 0;
    int32_t n_classes = 0, mask_bytes = 0;
    int64_t max_trees = 0, max_nodes = 0;
    std::vector<int64_t> trees_per_rank, nodes_per_rank;
    std::vector<uint8_t> small;  // [world][max_trees * (8 + 8 + mask_bytes)]: sizes | scores | masks
    rfmc_node* d_nodes = nullptr;  // [world][max_nodes]
    int32_t* d_counts = nullptr;   // [world][max_nodes][n_classes]
};

namespace {

struct Nccl {
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the instance that owns the caller's communicator
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        n.AllGather = (decltype(n.AllGather))dlsym(h, "ncclAllGather");
        n.CommCount = (decltype(n.CommCount))dlsym(h, "ncclCommCount");
        n.CommUserRank = (decltype(n.CommUserRank))dlsym(h, "ncclCommUserRank");
        n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
        n.ok = n.AllGather && n.CommCount && n.CommUserRank && n.GetErrorString;
    });
    return n;
}

#define RFMC_NCCL(call)                                                                          \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess) {                                                                 \
            rfmc::set_error(std::string(#call) + " failed: " + nccl().GetErrorString(r_));       \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

struct Scratch {  // device scratch of one exchange, released on every exit path
    std::vector<void*> blocks;
    cudaStream_t st;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() {
        cudaStreamSynchronize(st);
        for (void* p : blocks) release(p);
    }
    template <typename T>
    int get(T** out, size_t bytes) {
        *out = nullptr;
        RFMC_CUDA(dev_alloc((void**)out, bytes ? bytes : 1));
        blocks.push_back(*out);
        return 0;
    }
};

}  // namespace

extern "C" {

int rfmc_allgather_survivors(void* nccl_comm, int device, const rfmc_node* d_nodes, const int32_t* d_counts, int64_t n_nodes,
                             int32_t n_trees, const int64_t* tree_sizes, const double* scores, const uint8_t* masks,
                             int32_t mask_bytes, int32_t n_classes, void* stream, rfmc_gathered** out) {
    RFMC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RFMC_REQUIRE(nccl_comm != nullptr, "communicator is NULL");
    RFMC_REQUIRE(n_nodes >= 0 && n_trees >= 0 && mask_bytes >= 0, "bad sizes");
    RFMC_REQUIRE(n_classes >= 1 && n_classes <= RFMC_MAX_CLASSES, "n_classes must be in 1..64");
    RFMC_REQUIRE(n_trees == 0 || (tree_sizes && scores), "tree_sizes / scores are NULL");
    RFMC_REQUIRE(n_nodes == 0 || (d_nodes && d_counts), "node tables are NULL");
    RFMC_REQUIRE(mask_bytes == 0 || n_trees == 0 || masks, "masks are NULL");
    int64_t sum = 0;
    for (int i = 0; i < n_trees; ++i) sum += tree_sizes[i];
    RFMC_REQUIRE(sum == n_nodes, "tree_sizes do not add up to n_nodes");
    RFMC_REQUIRE(nccl().ok, "NCCL is not available in this process (libnccl.so.2)");
    RFMC_CUDA(cudaSetDevice(device));
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    cudaStream_t st = (cudaStream_t)stream;
    int world = 0, rank = 0;
    RFMC_NCCL(nccl().CommCount(comm, &world));
    RFMC_NCCL(nccl().CommUserRank(comm, &rank));
    Scratch sc(st);
    // 1. sizes first
    int64_t meta[2] = {n_trees, n_nodes};
    int64_t *d_meta = nullptr, *d_metas = nullptr;
    if (sc.get(&d_meta, sizeof(meta)) || sc.get(&d_metas, sizeof(meta) * world)) return 1;
    RFMC_CUDA(cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, st));
    RFMC_NCCL(nccl().AllGather(d_meta, d_metas, 2, ncclInt64, comm, st));
    std::vector<int64_t> metas((size_t)2 * world);
    RFMC_CUDA(cudaMemcpyAsync(metas.data(), d_metas, sizeof(meta) * world, cudaMemcpyDeviceToHost, st));
    RFMC_CUDA(cudaStreamSynchronize(st));
    rfmc_gathered* g = new rfmc_gathered();
    g->device = device;
    g->world = world;
    g->rank = rank;
    g->n_classes = n_classes;
    g->mask_bytes = mask_bytes;
    for (int r = 0; r < world; ++r) {
        g->trees_per_rank.push_back(metas[2 * r]);
        g->nodes_per_rank.push_back(metas[2 * r + 1]);
        g->max_trees = metas[2 * r] > g->max_trees ? metas[2 * r] : g->max_trees;
        g->max_nodes = metas[2 * r + 1] > g->max_nodes ? metas[2 * r + 1] : g->max_nodes;
    }
    auto fail = [&]() {
        rfmc_gathered_destroy(g);
        return 1;
    };
    // 2. the small per-tree records, each field padded to the longest rank
    const size_t rec = (size_t)g->max_trees * (16 + (size_t)mask_bytes);
    std::vector<uint8_t> mine(rec ? rec : 1, 0);
    if (n_trees) {
        memcpy(mine.data(), tree_sizes, (size_t)n_trees * 8);
        memcpy(mine.data() + (size_t)g->max_trees * 8, scores, (size_t)n_trees * 8);
        if (mask_bytes) memcpy(mine.data() + (size_t)g->max_trees * 16, masks, (size_t)n_trees * mask_bytes);
    }
    g->small.assign((rec ? rec : 1) * world, 0);
    uint8_t *d_small = nullptr, *d_smalls = nullptr;
    if (sc.get(&d_small, mine.size()) || sc.get(&d_smalls, mine.size() * world)) return fail();
    if (cudaMemcpyAsync(d_small, mine.data(), mine.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail();
    if (nccl().AllGather(d_small, d_smalls, mine.size(), ncclUint8, comm, st) != ncclSuccess) {
        set_error("ncclAllGather of the per-tree records failed");
        return fail();
    }
    if (cudaMemcpyAsync(g->small.data(), d_smalls, mine.size() * world, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail();
    // 3. the node tables and class counts, padded to the largest rank, device to device
    const size_t nb = (size_t)g->max_nodes * sizeof(rfmc_node), cb = (size_t)g->max_nodes * n_classes * sizeof(int32_t);
    if (dev_alloc((void**)&g->d_nodes, (nb ? nb : 1) * world) != cudaSuccess || dev_alloc((void**)&g->d_counts, (cb ? cb : 1) * world) != cudaSuccess) {
        set_error("allocating the gathered tables failed");
        return fail();
    }
    if (g->max_nodes > 0) {
        const void *send_n = d_nodes, *send_c = d_counts;
        if (n_nodes < g->max_nodes) {  // NCCL sends equal counts: stage this rank's tables in padded buffers
            uint8_t *pn = nullptr, *pc = nullptr;
            if (sc.get(&pn, nb) || sc.get(&pc, cb)) return fail();
            cudaMemsetAsync(pn, 0, nb, st);
            cudaMemsetAsync(pc, 0, cb, st);
            if (n_nodes) {
                cudaMemcpyAsync(pn, d_nodes, (size_t)n_nodes * sizeof(rfmc_node), cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync(pc, d_counts, (size_t)n_nodes * n_classes * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
            }
            send_n = pn;
            send_c = pc;
        }
        if (nccl().AllGather(send_n, g->d_nodes, nb, ncclUint8, comm, st) != ncclSuccess ||
            nccl().AllGather(send_c, g->d_counts, cb, ncclUint8, comm, st) != ncclSuccess) {
            set_error("ncclAllGather of the node tables failed");
            return fail();
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(std::string("survivor exchange failed: ") + cudaGetErrorString(cudaGetLastError()));
        return fail();
    }
    *out = g;
    return 0;
}

int rfmc_gathered_describe(rfmc_gathered* g, int32_t* world, int32_t* rank, int64_t* max_trees, int64_t* max_nodes,
                           int64_t* trees_per_rank, int64_t* nodes_per_rank) {
    RFMC_REQUIRE(g != nullptr, "handle is NULL");
    if (world) *world = g->world;
    if (rank) *rank = g->rank;
    if (max_trees) *max_trees = g->max_trees;
    if (max_nodes) *max_nodes = g->max_nodes;
    if (trees_per_rank) memcpy(trees_per_rank, g->trees_per_rank.data(), sizeof(int64_t) * g->world);
    if (nodes_per_rank) memcpy(nodes_per_rank, g->nodes_per_rank.data(), sizeof(int64_t) * g->world);
    return 0;
}

int rfmc_gathered_meta(rfmc_gathered* g, int32_t r, int64_t* tree_sizes, double* scores, uint8_t* masks) {
    RFMC_REQUIRE(g != nullptr && r >= 0 && r < g->world, "handle is NULL or rank out of range");
    const size_t rec = (size_t)g->max_trees * (16 + (size_t)g->mask_bytes);
    const uint8_t* row = g->small.data() + (rec ? rec : 1) * (size_t)r;
    const size_t nt = (size_t)g->trees_per_rank[(size_t)r];
    if (tree_sizes) memcpy(tree_sizes, row, nt * 8);
    if (scores) memcpy(scores, row + (size_t)g->max_trees * 8, nt * 8);
    if (masks && g->mask_bytes) memcpy(masks, row + (size_t)g->max_trees * 16, nt * g->mask_bytes);
    return 0;
}

int rfmc_gathered_fetch(rfmc_gathered* g, int32_t r, rfmc_node* nodes_out, int32_t* counts_out) {
    RFMC_REQUIRE(g != nullptr && r >= 0 && r < g->world, "handle is NULL or rank out of range");
    RFMC_CUDA(cudaSetDevice(g->device));
    const size_t n = (size_t)g->nodes_per_rank[(size_t)r];
    if (n == 0) return 0;
    if (nodes_out)
        RFMC_CUDA(cudaMemcpy(nodes_out, g->d_nodes + (size_t)r * g->max_nodes, n * sizeof(rfmc_node), cudaMemcpyDeviceToHost));
    if (counts_out)
        RFMC_CUDA(cudaMemcpy(counts_out, g->d_counts + (size_t)r * g->max_nodes * g->n_classes, n * g->n_classes * sizeof(int32_t),
                             cudaMemcpyDeviceToHost));
    return 0;
}

/* a 64-bit FNV-1a of the gathered tables of every rank (sizes, scores and node bytes): equal on every
 * rank after a correct exchange -- bench.py all-reduces it as its in-run parity check */
int rfmc_gathered_checksum(rfmc_gathered* g, uint64_t* out) {
    RFMC_REQUIRE(g != nullptr && out != nullptr, "NULL argument");
    RFMC_CUDA(cudaSetDevice(g->device));
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const uint8_t* b = (const uint8_t*)p;
        for (size_t i = 0; i < n; ++i) {
            h ^= b[i];
            h *= 1099511628211ull;
        }
    };
    std::vector<uint8_t> buf;
    for (int r = 0; r < g->world; ++r) {
        const size_t nt = (size_t)g->trees_per_rank[(size_t)r], n = (size_t)g->nodes_per_rank[(size_t)r];
        const size_t rec = (size_t)g->max_trees * (16 + (size_t)g->mask_bytes);
        const uint8_t* row = g->small.data() + (rec ? rec : 1) * (size_t)r;
        mix(row, nt * 8);
        mix(row + (size_t)g->max_trees * 8, nt * 8);
        buf.resize(n * sizeof(rfmc_node) + 1);
        if (n) {
            RFMC_CUDA(cudaMemcpy(buf.data(), g->d_nodes + (size_t)r * g->max_nodes, n * sizeof(rfmc_node), cudaMemcpyDeviceToHost));
            mix(buf.data(), n * sizeof(rfmc_node));
        }
    }
    *out = h;
    return 0;
}

int rfmc_gathered_destroy(rfmc_gathered* g) {
    if (!g) return 0;
    cudaSetDevice(g->device);
    release(g->d_nodes);
    release(g->d_counts);
    delete g;
    return 0;
}

}  // extern "C"
