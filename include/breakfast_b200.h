/*
 * breakfast_b200.h — C ABI of libbreakfast_b200.so
 *
 * B200-native (sm_100a) replacement for the distance-and-clustering hot path of
 * rki-mf1/breakfast.  The reference has no FFI of its own; the seam this ABI
 * replaces is the Python call chain
 *
 *     breakfast.cluster_features                 src/breakfast/breakfast.py:279-340
 *       -> sparse_feature_matrix row sums        src/breakfast/breakfast.py:285-291
 *       -> get_neighbours_batch (per cardinality) src/breakfast/breakfast.py:223-276
 *            -> sklearn pairwise_distances_chunked(metric="manhattan")
 *               + _reduce_func (d <= max_dist)   src/breakfast/breakfast.py:226-228,261-267
 *       -> _to_graph + connected_components      src/breakfast/breakfast.py:93-113,325-326
 *
 * Everything crossing this boundary is a plain pointer + size.  Host buffers are
 * caller-owned, contiguous, never retained.  Every function returns BF_OK (0) or
 * a negative bf_status; bf_last_error() gives the message for the calling
 * thread.  No exceptions cross the boundary.  There is no CPU fallback: without
 * a CUDA device every compute entry point fails with BF_ERR_NO_DEVICE.
 *
 * Row format: strictly binary CSR.  Row i owns indices[indptr[i]..indptr[i+1]),
 * sorted ascending, unique, each in [0, n_cols).  Repeated tokens of one profile
 * (reference: breakfast.py:210-212 appends them twice, sklearn then takes L1 on
 * counts) are thermometer-coded into extra columns by the host before the call,
 * so |A xor B| on these rows equals the reference's manhattan distance.
 */
#ifndef BREAKFAST_B200_H
#define BREAKFAST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_ABI_VERSION 1

typedef enum bf_status {
    BF_OK = 0,
    BF_ERR_INVALID = -1,      /* bad argument / malformed CSR                         */
    BF_ERR_NO_DEVICE = -2,    /* no usable CUDA device (there is no CPU fallback)     */
    BF_ERR_CUDA = -3,         /* a CUDA runtime call failed; see bf_last_error()      */
    BF_ERR_OOM = -4,          /* device or host allocation failed                     */
    BF_ERR_OVERFLOW = -5,     /* a bounded device buffer was too small (async API only): bf_sync raised its capacity, run again */
    BF_ERR_STATE = -6         /* call order violated (e.g. run before upload)         */
} bf_status;

/* Engines for the pairwise phase (replaces sklearn _sparse_manhattan,
 * sklearn/metrics/_pairwise_fast.pyx:34-107, as called at breakfast.py:261-267).
 *   SKETCH: m-bit XOR-fold of every row (popc(fold(A)^fold(B)) <= |A xor B|, so
 *           "> max_dist" rejects with no false negatives); the tiled XOR/POPC
 *           kernel runs on the folded bitsets and survivors are verified exactly
 *           on the CSR rows.  Exact.  Default.
 *   FULL:   all n_cols columns as dense bitsets, tiled XOR/POPC over the whole
 *           width, threshold is exact, no verify stage.  Exact.
 *   HASHJOIN (max_dist <= 2 only, BF_ERR_INVALID beyond): no pair tiles at all.  Small distances have a closed
 *           form (distance 1: one row is the other minus a column; distance 2: minus two columns, or equal size
 *           and equal after deleting one column each), and with an additive 64-bit row hash each case is an
 *           equi-join against a hash table of the rows / of their one-deletion keys; every match is verified
 *           exactly on the CSR rows.  Exact.  O(nnz) probes at distance 1 instead of O(N^2 / pruning) pairs
 *           (no counterpart in the reference).
 */
typedef enum bf_engine { BF_ENGINE_SKETCH = 0, BF_ENGINE_FULL = 1, BF_ENGINE_HASHJOIN = 2 } bf_engine;

typedef struct bf_stats {
    int64_t n_rows;           /* rows on the B side (all profiles)                     */
    int64_t n_query;          /* rows on the A side (== n_rows for a full run)         */
    int64_t n_cols;
    int64_t nnz;
    int64_t bits_per_row;     /* bitset width the pair kernel contracted over          */
    int64_t pairs_total;      /* N(N-1)/2, or |Q|*|all| for a rectangle                */
    int64_t pairs_band;       /* candidate pairs: ||A|-|B|| <= max_dist    */
    int64_t pairs_evaluated;  /* pairs the tile kernel evaluated on this rank (incl. tile padding) */
    int64_t tiles_total;      /* tile pairs of the whole (upper-triangular) tile space */
    int64_t tiles_band;       /* tile pairs left by the band pruning on (|A|, |A n H|) */
    int64_t tiles_rank;       /* of those, processed by this rank                      */
    int64_t n_candidates;     /* sketch survivors handed to the exact verify kernel    */
    int64_t n_edges;          /* verified edges (d <= max_dist), this rank             */
    int64_t n_components;     /* after the last union-find/merge                       */
    /* device time of the phases of the last run.  128/256-bit sketches: ms_sort = sketch + sort-key pass over the
       columns and the radix sort, ms_pack = sorted tile layouts; other forms: ms_sort = keys + sort, ms_pack = bit-pack */
    double ms_h2d, ms_sort, ms_pack, ms_pairs, ms_verify, ms_cc, ms_merge, ms_d2h, ms_total;
    /* accumulated over every bf_run since the previous bf_sync (at most the last 256 runs): */
    int64_t runs_since_sync;  /* how many bf_run calls these sums cover                        */
    int64_t kernel_launches;  /* kernels this library launched in those runs (incl. merges)    */
    double ms_pairs_sum;      /* summed device time of the pair kernel (CUDA events, own stream) */
    double ms_total_sum;      /* summed device time of whole passes                            */
    /* pair kernel work of the last run on this rank */
    int64_t l2_warp_items;    /* units level 1 handed to level 2: 32-pair thread units (two-kernel forms; with the
                                 tensor-core level 1 only the pairs passing the 32-bit test are read at full width)
                                 or 1024-pair warp units (single-kernel form)                                     */
    int64_t popc32_executed;  /* POPC32 lane-ops the pair kernel executed: level 1 + level 2       */
    double ms_l1_sum;         /* summed device time of the level-1 kernel (two-kernel path), like ms_pairs_sum */
} bf_stats;

typedef struct bf_ctx bf_ctx;

/* ---- library / device --------------------------------------------------- */
int bf_abi_version(void);
const char* bf_last_error(void);
int bf_device_count(int* n_out);

/* ---- context ------------------------------------------------------------ */
/* `stream` is a cudaStream_t (or NULL: the context creates its own). All work of
 * the context is enqueued on it, so a caller timing with events on that stream
 * sees every kernel. */
int bf_ctx_create(int device, void* stream, bf_ctx** ctx_out);
void bf_ctx_destroy(bf_ctx* ctx);
/* Options: "engine" (bf_engine), "sketch_bits" (power of two, 128..2048),
 * "want_edges" (0/1), "cand_capacity" (entries), "blocks_per_sm",
 * "two_level" (0/1, default 1: 32-bit first-level fold inside the pair kernel for
 * single-chunk sketches; exact either way), "level1" (0 = level 1 on the integer
 * pipes, 1 = on the tensor cores: int8 mma.sync on +-1 expanded folds, 2 (default) = the same with two
 * column rows per accumulator; 128/256-bit sketches), "l1_ctas" (CTAs per SM of the level-1 kernel of
 * "level1" = 2: 1 or 2, 0 = default), "l2_sub" (CTAs of the level-2 kernel per queue segment, 0 = default),
 * "resident_csr16" (default 1: bf_upload_csr also keeps the compact 16-bit form of the matrix on the device and the
 * sketch pass streams it), "verify_csr16" (default 0: 1 = the verification reads that form too; measured slower),
 * "shard_pack_from" (rank count from which a multi-GPU pass shards its sketch pass; default 16 = replicated on one box), "items_capacity" / "units_capacity" (entries of the expanded work list and
 * of the level-2 queue, 0 = automatic), "merge_capacity" (entries per rank of the compact label exchange of a
 * multi-GPU pass, 0 = automatic). */
int bf_ctx_set_option(bf_ctx* ctx, const char* key, int64_t value);

/* ---- async, device-resident API (used by bench.py and the multi-rank host) */
/* H2D of the CSR (replaces the scipy csr_matrix construction, breakfast.py:214).
 * `query_rows` (host, ascending, unique) selects the A side for the incremental
 * path (reference: select_ind, breakfast.py:236-245,300); NULL = all rows. */
int bf_upload_csr(bf_ctx* ctx, const int64_t* indptr, const int32_t* indices,
                  int64_t n_rows, int32_t n_cols,
                  const int32_t* query_rows, int64_t n_query);
/* Pipelined variant for callers that stream batches: the host buffers must be page-locked
 * (bf_pinned_alloc) and stay valid until the next bf_sync/bf_download_labels.  The copy runs on the
 * context's own copy stream into the idle one of two device slots, so it overlaps the pass that is
 * still running; the next bf_run waits for it and switches slots. */
int bf_upload_csr_async(bf_ctx* ctx, const int64_t* indptr, const int32_t* indices,
                        int64_t n_rows, int32_t n_cols);
/* Compact host form of a binary CSR ("CSR16"), half the bytes of the plain form on the host link: 32-bit row offsets
 * `indptr32[n_rows + 1]`, the low 16 bits of every column `lo[nnz]` and, when n_cols > 65536, `split[n_rows]` = how many
 * of a row's (ascending) columns are below 65536.  Representable when n_cols <= 131072, nnz < 2^32 and no row has more than
 * 65535 columns - SARS-CoV-2 profiles (about 88 000 distinct mutations) are.  bf_csr16_encode fills caller-allocated
 * arrays from the plain form (it validates the rows like bf_upload_csr; BF_ERR_INVALID when not representable);
 * bf_upload_csr16_async is bf_upload_csr_async for this form: page-locked buffers, copy on the context's copy stream into
 * the idle slot, decoded on the device (k_csr16_decode) behind the pass that is still running.  Same seam as
 * bf_upload_csr: the scipy csr_matrix construction, breakfast.py:214. */
int bf_csr16_encode(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols,
                    uint32_t* indptr32_out, uint16_t* split_out /* may be NULL when n_cols <= 65536 */, uint16_t* lo_out);
int bf_upload_csr16_async(bf_ctx* ctx, const uint32_t* indptr32, const uint16_t* split, const uint16_t* lo,
                          int64_t n_rows, int32_t n_cols);
/* Use a CSR that already lives in device memory owned by the caller (e.g. row shards that the ranks
 * uploaded in parallel and all-gathered over NVLink with torch.distributed).  Not copied; the caller
 * keeps it alive and orders its producers before bf_run on the context's stream. */
int bf_adopt_csr_device(bf_ctx* ctx, const void* indptr_device, const void* indices_device,
                        int64_t n_rows, int32_t n_cols, int64_t nnz);
/* Enqueue one pass of the hot path on the uploaded rows for this rank's share of
 * the band tiles: cardinality sort -> bit-pack -> tile schedule -> XOR/POPC pair
 * kernel -> exact verify -> union-find -> labels (device).  Nothing is copied to
 * the host; call bf_sync to wait and read counters.  world >= 1, 0 <= rank < world. */
int bf_run(bf_ctx* ctx, int32_t max_dist, int32_t rank, int32_t world);
/* ---- multi-GPU: the library's own communicator (NCCL over NVLink / NVSwitch) ------------------------------------
 * With a communicator, bf_run(ctx, d, rank, world) does the exchange steps of a multi-GPU pass itself, on the
 * context's stream: (1) every rank streams only its share of the rows for the sketch + sort-key pass and the shares
 * are all-gathered (24 bytes per row at 128 bits); (2) after verify + hook every rank publishes the rows of its
 * union-find that are not their own root as (row, root) pairs, the lists are all-gathered and re-united, so that
 * every rank ends with the same canonical labels.  Replaces the reference's single-process loop over cardinalities
 * (breakfast.py:314-318) + networkx components (breakfast.py:325-326) for N GPUs; bf_labels_to_device /
 * bf_merge_labels_* remain for callers that bring their own collective.  All ranks must call bf_run together.
 * NCCL is bound at run time (dlopen of libnccl.so.2, or the path in BREAKFAST_B200_NCCL): BF_ERR_STATE if absent.
 *   multi-process (one rank per process): rank 0 calls bf_comm_unique_id and hands the 128 bytes to the others by any
 *     means, then every rank calls bf_ctx_comm_init_rank;
 *   single process, one context per distinct device: bf_comm_init_all(ctxs, n) - rank i = ctxs[i]; the contexts are
 *     then driven from one host thread each. */
#define BF_COMM_ID_BYTES 128
int bf_comm_unique_id(void* id_out /* BF_COMM_ID_BYTES */);
int bf_ctx_comm_init_rank(bf_ctx* ctx, const void* id, int32_t rank, int32_t world);
int bf_comm_init_all(bf_ctx** ctxs, int32_t n);
int bf_ctx_comm_destroy(bf_ctx* ctx);

/* D2D: copy this rank's labels (int32[n_rows], label = smallest row index of the
 * row's component as seen by this rank) into caller device memory, e.g. a torch
 * tensor that torch.distributed all-gathers over NCCL. */
int bf_labels_to_device(bf_ctx* ctx, void* dst_device);
/* Merge `world` gathered label arrays (device, int32[world][n_rows]) into the
 * context's labels: union(i, labels_r[i]) for every r, then pointer jumping. */
int bf_merge_labels_device(bf_ctx* ctx, const void* gathered_device, int32_t world);
/* Same, but the gathered labels are on the host (single-process multi-GPU use). */
int bf_merge_labels_host(bf_ctx* ctx, const int32_t* gathered_host, int32_t world);
/* Extra hyper-edges: every list is a set of rows that must end up in one
 * component (cached / ghost neighbour lists, reference cache.py:51-71 and
 * breakfast.py:304,93-113).  CSR of lists on the host. Unions into current labels. */
int bf_union_lists(bf_ctx* ctx, const int64_t* list_indptr, const int32_t* list_members,
                   int64_t n_lists);
/* Wait for the stream, read counters/timers.  BF_ERR_OVERFLOW: a bounded device buffer (work list, level-2
 * queue, candidate/edge buffer) was too small for this input; its capacity has been raised - call bf_run again
 * (a later buffer of the chain may overflow in turn; the one-shot entry points below retry by themselves). */
int bf_sync(bf_ctx* ctx, bf_stats* stats_out);
int bf_download_labels(bf_ctx* ctx, int32_t* labels_out /* [n_rows] */);
/* Edge list of the last run (needs option want_edges=1): original row indices. */
int bf_edge_count(bf_ctx* ctx, int64_t* n_edges_out);
int bf_download_edges(bf_ctx* ctx, int32_t* src_out, int32_t* dst_out);

/* ---- one-shot host-buffer API (what the Python host calls through ctypes) -- */
/* Full run on one GPU: labels_out[i] = smallest row index in row i's component
 * of the graph {d(a,b) <= max_dist}.  Replaces breakfast.py:314-318 + 325-326. */
int bf_cluster_csr(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols,
                   int32_t max_dist, int32_t device, int32_t engine,
                   int32_t* labels_out, bf_stats* stats_out);
/* Radius-neighbour edges between `query_rows` (NULL = all) and all rows; used for
 * the cache export and the incremental path (breakfast.py:236-254: X = new rows,
 * Y = all rows).  Each unordered pair is reported once. */
int bf_neighbours_csr(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols,
                      const int32_t* query_rows, int64_t n_query,
                      int32_t max_dist, int32_t device, int32_t engine,
                      void** edges_handle_out, int64_t* n_edges_out, bf_stats* stats_out);
int bf_edges_copy(void* edges_handle, int32_t* src_out, int32_t* dst_out);
void bf_edges_free(void* edges_handle);
/* Connected components on the GPU over explicit edges plus member lists
 * (either may be empty).  Replaces _to_graph + networkx connected_components,
 * breakfast.py:93-113,325-326. */
int bf_components(int64_t n_rows, const int32_t* src, const int32_t* dst, int64_t n_edges,
                  const int64_t* list_indptr, const int32_t* list_members, int64_t n_lists,
                  int32_t device, int32_t* labels_out, int64_t* n_components_out);

/* ---- helpers --------------------------------------------------------------- */
/* Page-locked host memory for callers that want DMA-speed H2D/D2H. */
int bf_pinned_alloc(int64_t bytes, void** ptr_out);
void bf_pinned_free(void* ptr);
/* Register-resident pipe microbenchmarks on `device`; result in 1e9 lane-ops/s.
 * name: "popc32", "lop3", "iadd3", "xor_popc_add", "imma_s8" (mma.sync m16n8k32 int8, result in
 * 1e9 int8 MACs/s) or "umma_i8" (tcgen05.mma kind::i8 M128 N256 K32 with TMEM accumulators, 1e9 int8 MACs/s).
 * Roofline denominators. */
int bf_measure_peak(int32_t device, const char* name, double* gops_out);

#ifdef __cplusplus
}
#endif
#endif /* BREAKFAST_B200_H */
