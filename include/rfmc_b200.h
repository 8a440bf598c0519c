/*
 * rfmc_b200.h -- C ABI of the B200-native engine for the RandomForestMC hot path.
 *
 * The reference (ysraell/random-forest-mc v1.4.0) is pure Python and has no FFI; its only seam is
 * the method surface of RandomForestMC / BaseRandomForestMC.  Each entry point below replaces the
 * arithmetic of one group of those methods (cited as file:line relative to the reference root) and
 * is what a maintainer's ctypes stub would bind (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; rfmc_last_error() returns the
 *     message of the last failure on the calling thread; no C++ exception crosses the boundary;
 *   - the caller owns host buffers (contiguous, little-endian); the library owns device memory
 *     behind opaque handles; a handle is bound to one CUDA device;
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails;
 *   - threads: handles may be used from any thread.  Host-buffer voting calls on ONE rfmc_forest
 *     (rfmc_forest_predict with on_device = 0, rfmc_forest_predict_frame) share its pinned staging and
 *     take turns under a per-forest lock; device-pointer calls run on the caller's stream and need
 *     no lock.  A batch / draw plan belongs to the thread (and stream) that drives it; create and
 *     destroy of different handles are safe concurrently (device blocks come from a locked pool).
 *
 * Encoding contract (random-forest-mc_b200/encode.py): feature values are dense rank codes, code 0 =
 * missing/unseen; numeric columns carry a sorted float64 value table; class labels are coded by
 * their position in the lexicographically sorted label list.
 */
#ifndef RFMC_B200_H
#define RFMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFMC_ABI_VERSION 3 /* 2: device draw plans, survivor exchange, jump polynomials; 3: rfmc_forest_layout, rfmc_cache_trim */

/* column kinds and arithmetic tags (tag | bits<<8 | signed<<16) */
#define RFMC_KIND_NUMERIC 0
#define RFMC_KIND_CATEGORICAL 1
#define RFMC_TAG_F64 0
#define RFMC_TAG_F32 1
#define RFMC_TAG_INT 2

/* node flags */
#define RFMC_FLAG_CATEGORICAL 1u /* test is code == thr, else code >= thr (tree.py:177-180) */
#define RFMC_FLAG_TWO_ROW 2u     /* split value is the native element uniq[thr-1] (model.py:258-262) */

#define RFMC_MAX_CLASSES 64

/* One node of a grown tree (24 bytes).  Breadth-first numbering, root = 0.
 * Split node: feat >= 0, '>=' child = child, '<' child = child + 1.  Leaf: feat = -1, child = -1. */
typedef struct rfmc_node {
    int32_t feat;   /* column index into the dataset's feature list */
    uint32_t thr;   /* code threshold (numeric) or category code (categorical) */
    int32_t child;  /* index of the '>=' child */
    uint32_t flags; /* RFMC_FLAG_* */
    double sval;    /* numeric split value of a >2-row node (model.py:237), else 0 */
} rfmc_node;

/* One node of a forest flattened for voting (16 bytes, absolute indices into the forest arrays). */
typedef struct rfmc_pnode {
    uint32_t feat_flags; /* bits 0..15 column, bit 30 categorical, bit 31 leaf */
    uint32_t thr;        /* split: code threshold / category code; leaf: payload index */
    uint32_t child;      /* absolute index of the '>=' child ('<' child = child + 1) */
    uint32_t reserved;
} rfmc_pnode;
#define RFMC_PNODE_LEAF 0x80000000u
#define RFMC_PNODE_CATEGORICAL 0x40000000u

typedef struct rfmc_dataset rfmc_dataset;
typedef struct rfmc_batch rfmc_batch;
typedef struct rfmc_forest rfmc_forest;
typedef struct rfmc_drawplan rfmc_drawplan;
typedef struct rfmc_gathered rfmc_gathered;

int rfmc_abi_version(void);
/* A non-blocking CUDA stream on `device` as an opaque pointer for the `stream` arguments below
 * (so that several host threads can keep independent batches in flight). */
int rfmc_stream_create(int device, void** out);
int rfmc_stream_destroy(int device, void* stream);
const char* rfmc_last_error(void);
int rfmc_device_count(int* count);
/* The library keeps freed device blocks (up to 48 GB per device) and pinned staging buffers (up to 8 GB) in
 * process-wide pools so that no cudaMalloc / cudaFree / cudaHostAlloc -- each a device-wide synchronisation --
 * happens between the batches of a fit.  rfmc_cache_trim gives everything that is cached and not in use back to
 * the driver (for a process that shares the GPU with another allocator, e.g. torch); *freed_bytes may be NULL. */
int rfmc_cache_trim(int device, int64_t* freed_bytes);

/* ---- process_dataset (model.py:162-195): upload the encoded training matrix once ------------
 * codes      [n_rows][row_stride] unsigned codes of `code_width` bytes (1, 2 or 4), row-major,
 *            row_stride*code_width an odd multiple of 4 bytes
 * labels     [n_rows] class code (sorted-label order), n_classes <= RFMC_MAX_CLASSES
 * col_kind   [n_features] RFMC_KIND_*;  col_tag [n_features] packed arithmetic tag
 * table_off  [n_features+1] offsets into `tables` (float64 sorted distinct values per numeric
 *            column; categorical columns have an empty range); col_nunique[n_features]          */
int rfmc_dataset_create(int device, const void* codes, int code_width, int64_t n_rows, int32_t n_features,
                        int64_t row_stride, const uint8_t* labels, int32_t n_classes, const uint8_t* col_kind,
                        const int32_t* col_tag, const int32_t* col_nunique, const int64_t* table_off,
                        const double* tables, rfmc_dataset** out);
int rfmc_dataset_destroy(rfmc_dataset* ds);

/* ---- process_dataset on the device (model.py:162-195) --------------------
 * Replaces the construction of `dataset_numpy` (model.py:194) by a rank-coded matrix built in HBM.
 * col_ptr[c]   host column of n_rows contiguous elements, rows with missing values already dropped
 *              (model.py:181); col_dtype[c] = RFMC_DT_* (below).  Numeric columns go to the GPU raw:
 *              sorted distinct values (the value table) by a bitonic sort + ordered compaction, then
 *              code = 1 + rank by binary search.  RFMC_DT_CODE = int32 codes 1..cat_nunique[c] of a
 *              categorical column, ranked on the host in np.unique order (strings stay on the host).
 * col_tag[c]   packed arithmetic tag of numeric columns (RFMC_TAG_* | bits<<8 | signed<<16)
 * labels       [n_rows] class code in sorted-label order
 * The code width (1/2/4 bytes) and the row pitch (odd number of 32-bit words) are chosen by the
 * library; read them back with rfmc_dataset_describe.  table_off [n_features+1], tables_out
 * [table_off[n_features]] float64.  rfmc_dataset_codes copies the matrix to the host
 * ([n_rows][row_stride] of code_width bytes) for inspection and tests.                           */
int rfmc_dataset_encode(int device, int64_t n_rows, int32_t n_features, const void* const* col_ptr,
                        const int32_t* col_dtype, const int32_t* col_tag, const int32_t* cat_nunique,
                        const uint8_t* labels, int32_t n_classes, rfmc_dataset** out);
int rfmc_dataset_describe(rfmc_dataset* ds, int32_t* code_width, int64_t* row_stride, int32_t* col_nunique,
                          int64_t* table_off);
int rfmc_dataset_tables(rfmc_dataset* ds, double* tables_out);
int rfmc_dataset_codes(rfmc_dataset* ds, void* codes_out);

/* ---- plantTree + validationTree for a batch of candidates (model.py:222-314) ----------------
 * The host draws rows and features with the reference's own RNGs (model.py:198-219) and hands
 * them over; the GPU grows every candidate and scores it on its slot's hold-out rows.
 * idx_train  [n_slots][n_train] row ids in the reference's order (class-major, draw order)
 * idx_val    [n_slots][n_val]
 * cand_slot  [n_cand] slot of each candidate; feat_off [n_cand+1] offsets into feat_list
 * feat_list  ordered feature indices of every candidate (the sampleFeats output)
 * Step 1 uploads (host -> device) and reserves outputs, step 2 launches on `stream`
 * (asynchronous), step 3 waits and returns per-candidate results.                               */
int rfmc_batch_create(rfmc_dataset* ds, int32_t n_slots, int32_t n_train, int32_t n_val, const int32_t* idx_train,
                      const int32_t* idx_val, int32_t n_cand, const int32_t* cand_slot, const int32_t* feat_off,
                      const uint16_t* feat_list, int32_t max_depth, int32_t min_samples_split, void* stream,
                      rfmc_batch** out);
int rfmc_batch_run(rfmc_batch* b, void* stream);
/* correct[n_cand]: hold-out rows classified correctly (th_val = correct / n_val, model.py:314);
 * n_nodes[n_cand]: size of each node table. */
int rfmc_batch_results(rfmc_batch* b, int32_t* correct, int32_t* n_nodes);
/* Copy the node tables of selected candidates, packed back to back, to `nodes_out` /
 * `counts_out` ([sum n_nodes][n_classes] class counts per node).  dst_on_device != 0: the
 * destinations are device pointers (e.g. an NCCL all-gather send buffer). */
int rfmc_batch_pack(rfmc_batch* b, const int32_t* cand_ids, int32_t n_sel, rfmc_node* nodes_out, int32_t* counts_out,
                    int dst_on_device, void* stream);
int rfmc_batch_destroy(rfmc_batch* b);
/* launches made by the last rfmc_batch_run on this batch (for bench.py's gpu_launches) */
int rfmc_batch_launches(rfmc_batch* b);
/* Measurement hook (bench.py's roofline): with timing enabled, rfmc_batch_run brackets its two
 * kernels with CUDA events on the launching stream; rfmc_batch_kernel_ms waits for the last run and
 * returns the device time of the slot gather and of the grower in milliseconds.                  */
int rfmc_batch_set_timing(rfmc_batch* b, int enabled);
int rfmc_batch_kernel_ms(rfmc_batch* b, float* ms_gather, float* ms_grow);

/* ---- useForest / testForest / testForestProbs (forest.py:204-258, tree.py:155-212) -----------
 * A forest flattened on the host: per leaf payload a vote and class probabilities, plus the
 * variants used when the row lacks a split column (tree.py:171-175,201-210).
 * probs arrays are [n_payloads][n_classes] float64 in class_vals order.                         */
int rfmc_forest_create(int device, const rfmc_pnode* nodes, int64_t n_nodes, const int64_t* tree_root, int32_t n_trees,
                       int64_t n_payloads, const uint8_t* vote, const uint8_t* vote_absent, const double* probs,
                       const double* probs_absent, int32_t n_classes, const double* scores, double score_fsum,
                       rfmc_forest** out);
int rfmc_forest_destroy(rfmc_forest* f);
/* codes [n_rows][row_stride] (same layout rule as the dataset); absent[n_features] != 0 marks a
 * column missing from the frame (NULL = none).  soft / weighted as setSoftVoting / setWeightedTrees
 * (forest.py:151-155).  probs_out [n_rows][n_classes] float64 and/or labels_out [n_rows] int32
 * (class_vals index; NULL to skip).  on_device != 0: codes and outputs are device pointers and
 * the call is asynchronous on `stream`; otherwise host buffers, copies included, synchronous.   */
int rfmc_forest_predict(rfmc_forest* f, const void* codes, int code_width, int64_t n_rows, int64_t row_stride,
                        int32_t n_features, const uint8_t* absent, int soft, int weighted, double* probs_out,
                        int32_t* labels_out, int on_device, void* stream);

/* CPython's random.sample(population, k) for n_calls candidates in a row (model.py:216): the
 * population INDICES it picks, on CPython's own MT19937 stream (key[624] / *pos from
 * random.getstate()[1], advanced in place).  out holds sum(ks) indices back to back. */
int rfmc_pyrandom_sample_plan(uint32_t* key, int32_t* pos, int32_t n_pop, int32_t n_calls, const int32_t* ks,
                              int32_t* out);

/* ---- voting straight from a raw frame: the encode of predict-time values fused in front ---------
 * rfmc_forest_set_codebook: per feature the sorted distinct numeric split values of the forest
 * (thr_off [n_features+1] into thresholds); code(x) = #{t <= x}, NaN -> 0, so that the test
 * `x >= split_val` of tree.py:177-180 becomes `code >= k+1`.
 * rfmc_forest_predict_frame: n_cols host columns of n_rows elements each; col_feat[c] = feature
 * index it feeds, col_dtype[c] = RFMC_DT_* (RFMC_DT_CODE: int32 codes already final, used for
 * categorical columns whose values the host mapped to ids).  Columns are staged through pinned
 * memory chunk by chunk, encoded on the GPU into the row-major code tile and voted; the frame is
 * never materialised as a code matrix on the host. */
#define RFMC_DT_F64 0
#define RFMC_DT_F32 1
#define RFMC_DT_I64 2
#define RFMC_DT_I32 3
#define RFMC_DT_I16 4
#define RFMC_DT_I8 5
#define RFMC_DT_U8 6
#define RFMC_DT_U16 7
#define RFMC_DT_U32 8
#define RFMC_DT_U64 9
#define RFMC_DT_CODE 10
int rfmc_forest_set_codebook(rfmc_forest* f, int32_t n_features, const int64_t* thr_off, const double* thresholds);
int rfmc_forest_predict_frame(rfmc_forest* f, int64_t n_rows, int32_t n_features, int64_t row_stride, int code_width,
                              int32_t n_cols, const void* const* col_ptr, const int32_t* col_feat,
                              const int32_t* col_dtype, const uint8_t* absent, int soft, int weighted,
                              double* probs_out, int32_t* labels_out);

/* Per-(row, tree) leaf export for DecisionTreeMC.__call__ (tree.py:112-113) and per-tree votes
 * (sampleClass2trees, model.py:406-407): leaves_out [n_rows][n_trees] = payload index, bit 31
 * set when the walk crossed an absent column.  Host buffers, synchronous. */
int rfmc_forest_leaves(rfmc_forest* f, const void* codes, int code_width, int64_t n_rows, int64_t row_stride,
                       int32_t n_features, const uint8_t* absent, int32_t* leaves_out, void* stream);

/* Which traversal kernel a forest votes with: *small_tree_stages > 0 when every tree fits a shared-memory stage
 * (<= 1,024 nodes) and the forest is held in the compact layout of predict_small_kernel (whole trees staged in
 * shared memory, branch-free fixed-depth walks; frames without absent columns, codes of at most 16 bits);
 * 0: the generic blocked walk (predict_kernel).  Both replace tree.py:155-212 / forest.py:204-233 and are
 * bit-equal; this is a measurement / diagnostics query. */
int rfmc_forest_layout(rfmc_forest* f, int32_t* small_tree_stages);

/* ---- split_train_val (model.py:198-210): the draw behind np.random.choice(a, size, replace=False)
 * on numpy's global legacy MT19937 stream, restated on the host: out[0..size) =
 * permutation(n)[:size].  key[624] / *pos are np.random.get_state()[1:3] and are advanced in
 * place exactly as numpy would advance them. */
int rfmc_legacy_permutation_prefix(uint32_t* key, int32_t* pos, int64_t n, int64_t size, int64_t* out);

/* The numpy-stream draws of a whole plan of tree slots in one call, in the reference's order
 * (model.py:198-219 as driven by survivedTree model.py:318-326): per slot, one
 * choice(class_rows[c], per_class, replace=False) per class, then n_cands times
 * randint(randint_low, randint_high).  class_rows holds every class's ascending row ids back to
 * back (class c = [class_off[c], class_off[c+1])).  Outputs: idx_train [n_slots][n_classes *
 * min(per_class, n_train_pclass)] (class-major, draw order), idx_val [n_slots][the rest],
 * feat_counts [n_slots][n_cands]; key_snap [n_slots][624] / pos_snap [n_slots] = the stream state
 * BEFORE each slot (for the exact rewind of a mis-speculated plan; NULL to skip).  The permutation
 * resolves run on (n_threads & 0xFF) worker threads while the calling thread walks the stream;
 * n_threads | 0x100 additionally generates the MT19937 blocks on a producer thread. */
int rfmc_legacy_draw_plan(uint32_t* key, int32_t* pos, int32_t n_slots, int32_t n_classes, const int64_t* class_rows,
                          const int64_t* class_off, int64_t per_class, int64_t n_train_pclass, int32_t n_cands,
                          int64_t randint_low, int64_t randint_high, int32_t* idx_train, int32_t* idx_val,
                          int32_t* feat_counts, uint32_t* key_snap, int32_t* pos_snap, int32_t n_threads);

/* ---- split_train_val / sampleFeats on the device (model.py:198-219) -----------------
 * The same draws as rfmc_legacy_draw_plan, walked on the GPU: numpy's legacy MT19937 stream is
 * generated in HBM (GF(2) jump-ahead, one CTA per chunk of the stream), the masked rejection of
 * random_interval is decided in chunks of 4,096 words (GPU-wide against a band around the expected
 * accept count, the undecided words settled exactly by one CTA per stream, every decision verified by
 * a final pass), the first per_class entries of every class's permutation are resolved, and idx_train
 * / idx_val stay in HBM as the inputs of rfmc_batch_create_from_plan.  Same row ids, same feature
 * counts, same end state as numpy.
 *
 * rfmc_dataset_set_class_rows  once per dataset: class_rows = every class's ascending row ids back to
 *                       back (class c = [class_off[c], class_off[c+1])), in class_vals order.
 * rfmc_draw_plan_device launches the plan on `stream` (asynchronous).  key[624] / pos are
 *                       np.random.get_state()[1:3]; per_class <= RFMC_DRAW_MAX_PER_CLASS.
 * rfmc_drawplan_finish  waits; feat_counts [n_slots][n_cands] = the randint values; key_end / pos_end
 *                       = the stream state after the plan (what np.random.set_state takes).
 * rfmc_drawplan_state_at  the stream state BEFORE slot `slot` (slot == n_slots: after the plan), for the
 *                       exact rewind of a mis-speculated plan.
 * rfmc_drawplan_indices device -> host copy of the row ids of slots [slot0, slot0 + n_slots).
 * rfmc_drawplan_info    words generated / consumed / left pending for the chain (thousands, rounded
 *                       down), chunks the final pass redid with every word pending, kernel launches.
 *                       rfmc_drawplan_timing(1) makes every later plan bracket its kernels with CUDA
 *                       events; rfmc_drawplan_kernel_ms returns {fill, classify + chain rounds, final
 *                       pass, resolve} in milliseconds (measurement hooks of bench.py).               */
#define RFMC_DRAW_MAX_PER_CLASS 4096
/* t^(offset-1) mod phi (the characteristic polynomial of MT19937) as 624 words, bit i = coefficient
 * of t^i; and its host evaluation: the untempered window x[offset .. offset+623] of the stream whose
 * window at 0 is key[624] (tests; the device evaluates the same correlation in mt_fill_kernel).   */
int rfmc_mt_jump_poly(int64_t offset, uint32_t* out624);
int rfmc_mt_jump_window_host(const uint32_t* key, int64_t offset, uint32_t* out624);
int rfmc_dataset_set_class_rows(rfmc_dataset* ds, int32_t n_classes, const int64_t* class_rows, const int64_t* class_off);
int rfmc_draw_plan_device(rfmc_dataset* ds, const uint32_t* key, int32_t pos, int32_t n_slots, int64_t per_class,
                          int64_t n_train_pclass, int32_t n_cands, int64_t randint_low, int64_t randint_high, void* stream,
                          rfmc_drawplan** out);
int rfmc_drawplan_finish(rfmc_drawplan* pl, int32_t* feat_counts, uint32_t* key_end, int32_t* pos_end);
int rfmc_drawplan_state_at(rfmc_drawplan* pl, int32_t slot, uint32_t* key, int32_t* pos);
int rfmc_drawplan_indices(rfmc_drawplan* pl, int32_t slot0, int32_t n_slots, int32_t* idx_train, int32_t* idx_val);
int rfmc_drawplan_info(rfmc_drawplan* pl, int64_t* words_generated, int64_t* words_consumed, int64_t* words_pending,
                       int32_t* chunk_fallbacks, int32_t* launches);
int rfmc_drawplan_timing(int enabled);
int rfmc_drawplan_kernel_ms(rfmc_drawplan* pl, float* ms4);
int rfmc_drawplan_destroy(rfmc_drawplan* pl);
/* rfmc_batch_create with the row ids of a finished or in-flight draw plan (same stream): nothing but
 * the feature lists crosses the bus.  The plan must outlive the batch. */
int rfmc_batch_create_from_plan(rfmc_dataset* ds, rfmc_drawplan* pl, int32_t n_cand, const int32_t* cand_slot,
                                const int32_t* feat_off, const uint16_t* feat_list, int32_t max_depth,
                                int32_t min_samples_split, void* stream, rfmc_batch** out);

/* ---- multi-GPU: the survivor exchange of the rank-sharded fit ---------------------------
 * `self.data.extend(Tree_list)` (model.py:403) when the tree slots were grown by one process per GPU:
 * every rank contributes the node tables of its surviving trees (packed back to back on its device by
 * rfmc_batch_pack, dst_on_device = 1) and receives everybody's, in rank order.  nccl_comm is an
 * ncclComm_t passed as void* (torch: ProcessGroupNCCL._comm_ptr()); NCCL is resolved at run time from
 * the library that created it.  Sizes travel first, then the per-tree records (size, survived score,
 * used-feature mask of mask_bytes bytes), then the node and count tables padded to the largest rank:
 * four all-gathers, device to device.  tree_sizes / scores / masks are host arrays of this rank's
 * n_trees trees; d_nodes / d_counts device arrays of n_nodes = sum(tree_sizes) entries.
 * rfmc_gathered_describe  world, rank, per-rank tree and node counts ([world] each).
 * rfmc_gathered_meta      rank r's tree sizes, scores, masks (host).
 * rfmc_gathered_fetch     device -> host copy of rank r's node and count tables (they stay in HBM
 *                         until somebody asks).
 * rfmc_gathered_checksum  FNV-1a over every rank's sizes, scores and node bytes: equal on all ranks
 *                         after a correct exchange.                                                  */
int rfmc_allgather_survivors(void* nccl_comm, int device, const rfmc_node* d_nodes, const int32_t* d_counts, int64_t n_nodes,
                             int32_t n_trees, const int64_t* tree_sizes, const double* scores, const uint8_t* masks,
                             int32_t mask_bytes, int32_t n_classes, void* stream, rfmc_gathered** out);
int rfmc_gathered_describe(rfmc_gathered* g, int32_t* world, int32_t* rank, int64_t* max_trees, int64_t* max_nodes,
                           int64_t* trees_per_rank, int64_t* nodes_per_rank);
int rfmc_gathered_meta(rfmc_gathered* g, int32_t r, int64_t* tree_sizes, double* scores, uint8_t* masks);
int rfmc_gathered_fetch(rfmc_gathered* g, int32_t r, rfmc_node* nodes_out, int32_t* counts_out);
int rfmc_gathered_checksum(rfmc_gathered* g, uint64_t* out);
int rfmc_gathered_destroy(rfmc_gathered* g);

#ifdef __cplusplus
}
#endif
#endif /* RFMC_B200_H */
