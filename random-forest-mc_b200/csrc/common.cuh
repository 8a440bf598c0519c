// Shared declarations for the rfmc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/rfmc_b200.h"

namespace rfmc {

void set_error(const std::string& msg);

#define RFMC_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            rfmc::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_));       \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define RFMC_REQUIRE(cond, msg)            \
    do {                                   \
        if (!(cond)) {                     \
            rfmc::set_error(msg);          \
            return 1;                      \
        }                                  \
    } while (0)

constexpr unsigned FULL = 0xFFFFFFFFu;

struct WorkItem {  // one node of the level being processed
    uint32_t lo, hi;  // segment of the candidate's position array
    uint32_t lpos;    // position inside its level (node id = level base + lpos)
    uint32_t off;     // rotation offset into the candidate's feature list
};

}  // namespace rfmc

struct rfmc_dataset {
    int device = 0;
    int code_width = 0;
    int64_t n_rows = 0;
    int32_t n_features = 0;
    int64_t stride = 0;
    int32_t n_classes = 0;
    int sm_count = 0;
    void* d_codes = nullptr;
    uint8_t* d_labels = nullptr;
    uint8_t* d_kind = nullptr;
    int32_t* d_tag = nullptr;
    int32_t* d_nuniq = nullptr;
    int64_t* d_taboff = nullptr;
    double* d_tables = nullptr;
    // host copies of what rfmc_dataset_encode derived on the device (rfmc_dataset_describe / _tables)
    std::vector<int32_t> h_nuniq;
    std::vector<int64_t> h_taboff;
    std::vector<double> h_tables;
    // per-class row lists for the device draws (rfmc_dataset_set_class_rows)
    int32_t n_draw_classes = 0;
    int32_t* d_class_n = nullptr;     // [C] rows per class
    int64_t* d_class_aoff = nullptr;  // [C+1] offsets of each class's n-1 shuffle draws inside one slot
    int32_t* d_class_rows = nullptr;  // ascending row ids of every class back to back
    int64_t* d_class_roff = nullptr;  // [C+1]
    std::vector<int64_t> h_class_n;
};

struct rfmc_batch {
    rfmc_dataset* ds = nullptr;
    int32_t n_slots = 0, n_train = 0, n_val = 0, n_cand = 0, max_depth = 0, mss = 1;
    int32_t nt_pad = 0, max_nodes = 0, fmax = 0, grid = 0;
    int launches = 0;
    bool own_idx = true;  // false: idx_train / idx_val belong to a draw plan (rfmc_batch_create_from_plan)
    // inputs
    int32_t* d_idx_train = nullptr;
    int32_t* d_idx_val = nullptr;
    int32_t* d_cand_slot = nullptr;
    int32_t* d_feat_off = nullptr;
    uint16_t* d_feat_list = nullptr;
    void* d_inputs = nullptr;  // non-null: the five input arrays are slices of this one block (one pinned staging copy)
    void* up_lease = nullptr;  // the pinned staging buffer of that copy, held until the batch is destroyed
    // per tree slot: compact [feature][row] codes and labels of the bootstrap rows (slot_gather_kernel)
    void* d_cc = nullptr;
    uint8_t* d_lab = nullptr;
    // per-CTA scratch
    uint32_t* d_pos = nullptr;
    rfmc::WorkItem* d_work = nullptr;
    uint32_t* d_resna = nullptr;
    uint32_t* d_rank = nullptr;
    uint16_t* d_resfin = nullptr;
    // per-candidate outputs
    rfmc_node* d_nodes = nullptr;
    int32_t* d_counts = nullptr;
    uint8_t* d_votes = nullptr;
    int32_t* d_nnodes = nullptr;
    int32_t* d_correct = nullptr;
    std::vector<int32_t> h_nnodes;
    bool have_results = false;
    cudaStream_t stream = nullptr;
    bool timing = false;  // measurement hook: events around the two kernels of a run
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

// One plan of tree slots drawn on the device (csrc/rng_walk.cu): numpy's legacy stream walked in HBM.
struct rfmc_drawplan {
    rfmc_dataset* ds = nullptr;
    cudaStream_t stream = nullptr;
    int32_t n_slots = 0, n_classes = 0, n_cands = 0, pos0 = 0;
    int64_t size = 0, n_tr = 0, n_va = 0;  // per class: drawn, training part, hold-out part
    int64_t randint_low = 0;
    uint32_t rmask = 0, rrange = 0;
    int cap_log2 = 10;
    int attempt = 0;
    int64_t n_blocks = 0, n_chunks = 0;  // MT19937 blocks generated, fill CTAs
    int blocks_per_chunk = 0;
    int64_t n_walk_chunks = 0;           // 4,096-word chunks of the walk
    int chunks_per_round = 128;
    int64_t steps_per_slot = 0;
    int launches = 0;
    bool finished = false;
    uint32_t h_key[624];
    uint32_t* d_key = nullptr;
    const uint32_t* d_polys = nullptr;  // borrowed from the per-device table
    uint32_t* d_W = nullptr;
    uint32_t* d_A = nullptr;
    uint32_t* d_heads = nullptr;
    int32_t* d_idx_train = nullptr;
    int32_t* d_idx_val = nullptr;
    int32_t* d_randv = nullptr;
    int64_t* d_k0 = nullptr;
    int64_t* d_kpred = nullptr;
    int32_t* d_eband = nullptr;
    void* d_hdr = nullptr;
    void* d_pend = nullptr;
    int64_t* d_slot_pos = nullptr;
    int32_t* d_status = nullptr;
    std::vector<int64_t> h_slot_pos;
    int32_t h_status[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // measurement hook
    bool timing = false;
};

struct PredictPipe;

struct rfmc_forest {
    int device = 0;
    int64_t n_nodes = 0, n_payloads = 0;
    int32_t n_trees = 0, n_classes = 0;
    double score_fsum = 0.0;
    rfmc_pnode* d_nodes = nullptr;
    int64_t* d_root = nullptr;
    uint8_t* d_vote = nullptr;
    uint8_t* d_vote_absent = nullptr;
    double* d_probs = nullptr;
    double* d_probs_absent = nullptr;
    double* d_scores = nullptr;
    int sm_count = 0;
    PredictPipe* pipe = nullptr;  // pinned staging + streams of the host-buffer voting path
    std::mutex pipe_mu;           // one host-buffer voting call at a time per forest
    uint4* d_blocks = nullptr;       // blocked 8-byte-node layout for the vote kernel (nullptr: not eligible)
    uint32_t* d_root_b = nullptr;
    int64_t n_blocks = 0;
    uint2* d_tops = nullptr;  // per tree: the top `top_levels` levels as an implicit heap of 8-byte records
    int32_t top_levels = 0;
    // compact layout of a forest of small trees (predict_small_kernel): whole trees are staged in shared memory
    uint2* d_crec = nullptr;        // per node {feat | lo << 16, span | next << 16}, tree-relative child index
    uint32_t* d_crec4 = nullptr;    // the same as 4-byte records (nullptr: feature ids or level widths too large)
    uint32_t* d_cpay = nullptr;     // per node: payload index of a leaf
    uint32_t* d_ctree = nullptr;    // [n_trees + 1] first record of every tree
    uint8_t* d_cdepth = nullptr;    // [n_trees] levels to walk (depth of the deepest leaf)
    uint32_t* d_cstage = nullptr;   // [n_stages + 1] first tree of every shared-memory stage
    int32_t n_stages = 0;
    double* d_thresholds = nullptr;  // code book of the frame encoder (rfmc_forest_set_codebook)
    int64_t* d_thr_off = nullptr;
    int32_t n_features_book = 0;
};

namespace rfmc {
cudaError_t dev_alloc(void** p, size_t bytes);
void release(void* p);
int drawplan_launch(rfmc_drawplan* pl);
int64_t drawplan_words_needed(const std::vector<int64_t>& class_n, int n_slots, int n_cands, uint32_t rmask, uint32_t rrange,
                              int attempt);
int drawplan_geometry(rfmc_drawplan* pl, int64_t words);
int64_t drawplan_chunk_words(const rfmc_drawplan* pl);
int launch_grow(rfmc_batch* b, cudaStream_t stream);
int grow_grid_size(const rfmc_dataset* ds, int n_cand, int n_train);
int debug_read(unsigned long long* out, int n);
int launch_pack(rfmc_batch* b, const int32_t* d_cand_ids, const int64_t* d_dst_off, int32_t n_sel, rfmc_node* nodes_out,
                int32_t* counts_out, cudaStream_t stream);
int launch_predict(rfmc_forest* f, const void* d_codes, int code_width, int64_t n_rows, int64_t row_stride,
                   int32_t n_features, const uint8_t* d_absent, bool has_absent, int soft, int weighted, double* d_probs,
                   int32_t* d_labels, cudaStream_t stream);
int launch_leaves(rfmc_forest* f, const void* d_codes, int code_width, int64_t n_rows, int64_t row_stride,
                  int32_t n_features, const uint8_t* d_absent, int32_t* d_out, cudaStream_t stream);
int launch_encode(const void* d_raw, const int64_t* d_col_off, const int32_t* d_col_feat, const int32_t* d_col_dtype,
                  int32_t n_cols, int64_t n_rows, int64_t row_stride, int code_width, const double* d_thresholds,
                  const int64_t* d_thr_off, void* d_codes, cudaStream_t stream);
int64_t rank_padded_rows(int64_t n_rows);
int rank_scan_blocks(int64_t padded);
int launch_rank_column(const void* d_raw, int dtype, int64_t n_rows, int64_t padded, uint64_t* d_keys, uint32_t* d_block,
                       double* d_table, uint32_t* d_n_distinct, uint32_t* d_col_codes, cudaStream_t stream);
int launch_interleave(const uint32_t* d_col_codes, int64_t n_rows, int32_t n_features, int64_t stride, int code_width,
                      void* d_codes, cudaStream_t stream);
}  // namespace rfmc
