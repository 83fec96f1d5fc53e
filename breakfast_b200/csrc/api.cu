// api.cu — host side of libbreakfast_b200.so: the C ABI declared in include/breakfast_b200.h.
// Owns device memory, the stream, kernel launches and timers.  No torch, no CPU fallback.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/breakfast_b200.h"
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time: the library does not link libnccl, so it loads (and the single-GPU path works) on a box
// without it; the first communicator call dlopen()s libnccl.so.2 - the copy already in the process if the caller
// (torch, say) brought one, else the system's.  Only the handful of entry points the exchange steps need.
// ------------------------------------------------------------------------------------------------
namespace bfnccl {
typedef struct ncclComm* comm_t;
struct unique_id { char internal[128]; };
enum { kUint8 = 1 };
struct Api {
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*CommSplit)(comm_t, int, int, comm_t*, void*) = nullptr;   // optional (NCCL >= 2.18)
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool ok = false;
};
inline Api& api() {
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {getenv("BREAKFAST_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names)
        if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) {
        a.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        return a;
    }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.CommSplit = (decltype(a.CommSplit))sym("ncclCommSplit");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd && a.GetErrorString;
    if (!a.ok) a.error = "libnccl.so.2 lacks an expected symbol";
    return a;
}
}  // namespace bfnccl

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at api.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    g_err = buf;
    (void)cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return BF_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return BF_ERR_NO_DEVICE;
    return BF_ERR_CUDA;
}

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__);   \
    } while (0)
#define CKL() CK(cudaGetLastError())
// BF_TRACE_KERNELS=1 (diagnostics): an event after every launch of a pass; bf_sync prints the in-situ timeline of the
// last pass (api.cu line of the launch, microseconds since the previous mark) on stderr
#define CKLC(c) do { CK(cudaGetLastError()); ++(c)->launches_since_sync; trace_mark((c), __LINE__); } while (0)
#define TRY(expr)                    \
    do {                             \
        int rc_ = (expr);            \
        if (rc_ != BF_OK) return rc_; \
    } while (0)

#define NCK(call)                                                                                   \
    do {                                                                                            \
        int r_ = (call);                                                                            \
        if (r_ != 0) {                                                                              \
            char buf_[400];                                                                         \
            snprintf(buf_, sizeof buf_, "NCCL error %d (%s) at api.cu:%d: %s", r_,                  \
                     bfnccl::api().GetErrorString ? bfnccl::api().GetErrorString(r_) : "?", __LINE__, #call); \
            return fail(BF_ERR_CUDA, buf_);                                                         \
        }                                                                                           \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return BF_OK;
        if (p) {
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = cudaMalloc(&p, bytes);  // retry without slack
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            return cuda_fail(e, "cudaMalloc", __LINE__);
        }
        cap = want;
        return BF_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline unsigned grid_for(int64_t n, int block) { return (unsigned)std::max<int64_t>(1, ceil_div(n, block)); }

}  // namespace

struct bf_ctx;
static void trace_mark(bf_ctx* c, int line);

struct bf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;

    // options
    int engine = BF_ENGINE_SKETCH;
    int sketch_bits = 256;
    int want_edges = 0;
    int64_t cand_capacity = 0;  // 0 = auto
    int blocks_per_sm = 0;      // 0 = occupancy
    int two_level = 1;
    int level1 = 2;             // 2 = tensor cores, two column rows per accumulator (k_pairs_l1_imma2, default),
                                // 1 = tensor cores (k_pairs_l1_imma), 0 = integer pipes (k_pairs_l1)
    int l1_ctas = 0;            // CTAs per SM of k_pairs_l1_imma2 (1 or 2; 0 = the default, kL1CtasDefault)
    int l2_sub = 0;             // option "l2_sub": CTAs of k_pairs_l2_unit per queue segment (0 = default)
    int l1_segs = 0;            // segments of the level-2 queue of the last pass (= the grid of the tensor-core level 1)
    int64_t items_capacity = 0;  // 0 = auto
    int64_t units_capacity = 0;  // 0 = auto (level-2 queue)

    // communicator of a multi-GPU job (NCCL); with it bf_run does its own exchange steps
    bfnccl::comm_t comm = nullptr;
    bfnccl::comm_t comm_copy = nullptr;   // a duplicate of `comm` for collectives on the copy stream (sharded uploads)
    int comm_rank = 0, comm_world = 1;
    int64_t merge_capacity = 0;   // entries per rank of the compact label exchange, 0 = automatic
    int64_t merge_auto = 0;       // automatic capacity learnt from earlier passes (what the fullest list needed, + 25 %)
    unsigned long long merge_cap_used = 0;
    DevBuf xchg;

    // problem
    bool uploaded = false, ran = false, has_query = false;
    int64_t n_rows = 0, n_query = 0, nnz = 0;
    int32_t n_cols = 0;

    // last run
    int max_dist = 0, rank = 0, world = 1;
    int K4 = 1, n_chunks = 1;
    int64_t bits_per_row = 0, tilesA = 0, tilesB = 0;
    unsigned long long cand_cap_used = 0, items_cap_used = 0;
    int sched_ranges = 1;   // schedule entries per row tile of the last run (2 max_dist + 1)
    bool ran_two_level = false, ran_two_kernel = false;
    unsigned long long queue_cap_used = 0;
    float ms_h2d = 0, ms_merge = 0, ms_d2h = 0;

    // CSR on the device: two owned slots (async uploads fill the idle one while a pass runs on the
    // other) or caller-owned memory (bf_adopt_csr_device); d_indptr/d_indices is what kernels read
    DevBuf indptr[2], indices[2], query_rows, is_query;
    // the compact resident form ("CSR16": uint32 offsets, uint16 low column halves, uint16 split per row), one set per
    // slot: uploaded as such (bf_upload_csr16_async) or derived on the device once per upload (bf_upload_csr).  The staged
    // sketch pass (the one kernel that streams the whole matrix) reads it instead of the plain CSR - half the bytes.  The
    // plain form is always there as well: the verification and the other engines walk single rows of it.
    DevBuf c16_indptr[2], c16_split[2], c16_lo[2], c16_bad;
    bool c16_valid[2] = {false, false}, c16_has_split[2] = {false, false};
    int resident16 = 1;    // option "resident_csr16": 0 = the sketch pass streams the plain CSR
    int pack16_variant = 1;   // one lane per row, ATOMS (fastest of the four, profiles/)
    int active_slot = -1;  // slot the current matrix lives in, -1 = caller-owned memory (bf_adopt_csr_device) or none
    bool use16 = false;    // this pass's sketch kernel reads the compact form
    int shard_pack_from = 16;   // option "shard_pack_from"
    int verify16 = 0;      // option "verify_csr16" = 1: the exact verification reads its rows from the compact form too
                           // (measured: 217 us against 170 us on the plain CSR at 10^6 profiles - half the bytes, but 16-bit
                           // loads and the rebuilt 17th bit cost more than the bytes save; kept as an option)
    DevBuf hj_hash, hj_t1, hj_t2, hj_t2_rows;        // hash-join engine: row hashes, row table, one-deletion table
    const int64_t* d_indptr = nullptr;
    const int32_t* d_indices = nullptr;
    int cur = 0, pending = -1;
    int64_t pend_rows = 0, pend_nnz = 0;
    int32_t pend_cols = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_upload_done = nullptr, ev_upload_start = nullptr, ev_slot_free[2] = {};
    bool slot_used[2] = {false, false};
    DevBuf keysB[3], valsB[3], keysA[3], valsA[3], sort_counts, sort_max, sk_rows, sched_table;
    int sched_table_dist = -1, sched_table_n = 0;
    // the candidate-pair statistic (pairs_band) depends on the rows and max_dist only: counted once per uploaded matrix
    DevBuf band_cache;
    bool band_valid = false;
    int band_dist = -1;
    DevBuf bitsA, bitsB, foldsA[2], foldsB[2], fold8A[2], fold8B[2], jlo, jend, wprefix, nwork, items, queue, segcnt, cand, edges, parent, labels, counters, scratch, scratch2;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_aux[2] = {};
    // per-run event ring so that bf_sync can report sums over all runs since the last sync
    static constexpr int kRing = 128;
    cudaEvent_t ring[kRing][5] = {};  // pass start, pairs start, pairs end, pass end, level-1 end
    int64_t runs_since_sync = 0;
    int64_t launches_since_sync = 0;
    // BF_TRACE_KERNELS diagnostics
    int trace = -1;   // -1 = environment not read yet
    std::vector<cudaEvent_t> trace_ev;
    std::vector<int> trace_line;
    size_t trace_n = 0;
};

static void trace_mark(bf_ctx* c, int line) {
    if (c->trace < 0) {
        const char* e = getenv("BF_TRACE_KERNELS");
        c->trace = (e && *e && *e != '0') ? 1 : 0;
    }
    if (!c->trace) return;
    if (c->trace_n == c->trace_ev.size()) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        c->trace_ev.push_back(ev);
        c->trace_line.push_back(0);
    }
    c->trace_line[c->trace_n] = line;
    cudaEventRecord(c->trace_ev[c->trace_n++], c->stream);
}
static void trace_dump(bf_ctx* c) {
    if (c->trace != 1 || c->trace_n < 2) return;
    float total = 0;
    for (size_t i = 1; i < c->trace_n; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->trace_ev[i - 1], c->trace_ev[i]) == cudaSuccess) {
            fprintf(stderr, "[bf trace] api.cu:%-5d %9.1f us\n", c->trace_line[i], ms * 1e3f);
            total += ms;
        }
    }
    fprintf(stderr, "[bf trace] total %.1f us over %zu marks\n", total * 1e3f, c->trace_n);
}

namespace {

using namespace bf;

int set_device(bf_ctx* c) {
    CK(cudaSetDevice(c->device));
    return BF_OK;
}

// Stable LSD radix sort of (key, row) by the 64-bit key c << 32 | s << 16 | t (kernels.cuh, K2) over the bits that
// occur in the data, eight per pass; keys[0] / vals[0] hold the result whatever the number of passes (three buffers).
// keys_ready: keys[0] / vals[0] and the OR word of `side` were already written (k_pack_sketch_rows, k_gather_keys).
int sort_by_card(bf_ctx* c, const int32_t* rows_dev, int64_t n, DevBuf keys[3], DevBuf vals[3], int side, bool keys_ready) {
    if (n == 0) return BF_OK;
    sortkey_t* or_key = c->sort_max.as<sortkey_t>() + side;
    if (!keys_ready) {
        CK(cudaMemsetAsync(or_key, 0, sizeof(sortkey_t), c->stream));
        k_card_keys<<<grid_for(n, 256), 256, 0, c->stream>>>(c->d_indptr, rows_dev, n,
                                                              keys[0].as<sortkey_t>(), vals[0].as<int32_t>(), or_key);
        CKLC(c);
    }
    SortBufs b;
    for (int i = 0; i < 3; ++i) {
        b.k[i] = keys[i].as<sortkey_t>();
        b.v[i] = vals[i].as<int32_t>();
    }
    // one cooperative launch: a block per SM, all passes inside (grid-wide barriers between the phases)
    const int grid = c->num_sms;
    TRY(c->sort_counts.ensure((size_t)256 * grid * sizeof(uint32_t)));
    uint32_t* counts = c->sort_counts.as<uint32_t>();
    int64_t n_arg = n;
    void* args[] = {&b, &n_arg, &counts, &or_key};
    CK(cudaLaunchCooperativeKernel((const void*)k_radix_sort, dim3((unsigned)grid), dim3(SORT_THREADS), args, 0, c->stream));
    CKLC(c);
    return BF_OK;
}

int ensure_sort_buffers(bf_ctx* c, int64_t n, DevBuf keys[3], DevBuf vals[3]) {
    for (int i = 0; i < 3; ++i) {
        TRY(keys[i].ensure(std::max<int64_t>(n, 1) * sizeof(sortkey_t)));
        TRY(vals[i].ensure(std::max<int64_t>(n, 1) * sizeof(int32_t)));
    }
    TRY(c->sort_max.ensure(2 * sizeof(sortkey_t)));
    return BF_OK;
}

// (D, Ds) combinations a partner at distance <= d can have, in ascending order, each with the feasible interval of Dt
// (kernels.cuh K2b): exists u with |u| + |Ds - u| + |Dt - u| + |D - Ds - Dt + u| <= d
std::vector<SchedRange> make_sched_table(int d) {
    std::vector<SchedRange> t;
    for (int D = -d; D <= d; ++D)
        for (int Ds = -((d - D) / 2); Ds <= (D + d) / 2; ++Ds) {
            int lo = 1, hi = 0;
            for (int Dt = -d; Dt <= d; ++Dt) {
                bool ok = false;
                for (int u = -d; u <= d && !ok; ++u)
                    ok = std::abs(u) + std::abs(Ds - u) + std::abs(Dt - u) + std::abs(D - Ds - Dt + u) <= d;
                if (ok) {
                    if (lo > hi) lo = Dt;
                    hi = Dt;
                }
            }
            if (lo <= hi) t.push_back(SchedRange{(int8_t)D, (int8_t)Ds, (int8_t)lo, (int8_t)hi});
        }
    return t;
}

constexpr int kThreeKeyMaxDist = 8;   // beyond this the (D, Ds) table grows quadratically: two keys

// 128/256-bit sketches (kernels.cuh K1): do the rows of a side go through the staged two-step pack?
inline bool staged_pack(const bf_ctx* c) {
    return c->engine == BF_ENGINE_SKETCH && (c->sketch_bits == 128 || c->sketch_bits == 256);
}

// plain CSR of `slot` -> compact resident form of the same slot (k_csr16_encode); c16_bad tells afterwards whether the
// matrix was representable
int encode_c16(bf_ctx* c, int slot, int64_t n_rows, int64_t nnz, cudaStream_t stream) {
    TRY(c->c16_indptr[slot].ensure((size_t)(n_rows + 1) * sizeof(uint32_t)));
    TRY(c->c16_split[slot].ensure((size_t)n_rows * sizeof(uint16_t)));
    TRY(c->c16_lo[slot].ensure((size_t)(nnz + 8) * sizeof(uint16_t)));   // + 16 bytes: bulk copies end on a 16-byte boundary
    TRY(c->c16_bad.ensure(sizeof(int)));
    CK(cudaMemsetAsync(c->c16_bad.p, 0, sizeof(int), stream));
    k_csr16_encode<<<c->num_sms * 8, 256, 0, stream>>>(c->indptr[slot].as<int64_t>(), c->indices[slot].as<int32_t>(), n_rows,
                                                       c->c16_indptr[slot].as<uint32_t>(), c->c16_split[slot].as<uint16_t>(),
                                                       c->c16_lo[slot].as<uint16_t>(), c->c16_bad.as<int>());
    CKLC(c);
    return BF_OK;
}

// step 1 of the staged pack: sketches + sort keys of ALL rows of the matrix in storage order (the B side is the
// whole matrix; a query subset reads its rows' sketches and keys from the same staging arrays)
// rows the staging arrays must hold: with a communicator every rank owns an equal, block-aligned share (the last ones
// possibly past the end), so that the shares can be all-gathered in place
inline bool dist_run(const bf_ctx* c) { return c->comm != nullptr && c->world > 1; }
// Sharding the sketch + key pass over the ranks costs an all-gather of 24 bytes per row (about 0.08 ms on NVLink whatever
// the rank count), the pass itself 0.074 ms / world since it streams the compact form: measured on 8 B200s the sketch +
// sort phase takes 0.243 ms sharded against 0.20 ms replicated (profiles/r02_bench_n{2,8}_final.json), so the pass is
// replicated on one box; option "shard_pack_from" = the rank count from which it is sharded (default 16: never on 8 GPUs)
inline bool shard_pack(const bf_ctx* c) { return dist_run(c) && c->world >= c->shard_pack_from; }
inline int64_t staged_rows(const bf_ctx* c) {
    const int64_t blocks = ceil_div(c->n_rows, TILE);
    return shard_pack(c) ? ceil_div(blocks, c->world) * c->world * TILE : blocks * TILE;
}

int pack_stage_all_rows(bf_ctx* c) {
    const int64_t n = c->n_rows;
    if (n == 0) return BF_OK;
    int log2m = 0;
    while ((1 << log2m) < c->sketch_bits) ++log2m;
    const int words = c->sketch_bits / 32;
    TRY(c->sk_rows.ensure((size_t)staged_rows(c) * words * sizeof(uint32_t)));
    sortkey_t* or_key = c->sort_max.as<sortkey_t>();
    CK(cudaMemsetAsync(or_key, 0, 2 * sizeof(sortkey_t), c->stream));
    const int64_t blocks = ceil_div(n, TILE);
    int64_t block0 = 0, my_blocks = blocks, share = blocks;
    if (shard_pack(c)) {   // this rank streams only its share of the rows; the shares are exchanged below
        share = ceil_div(blocks, c->world);
        block0 = share * c->rank;
        my_blocks = std::max<int64_t>(0, std::min(share, blocks - block0));
    }
    if (my_blocks > 0 && c->use16) {
        const uint32_t* ip = c->c16_indptr[c->cur].as<uint32_t>();
        const uint16_t* sp = c->c16_has_split[c->cur] ? c->c16_split[c->cur].as<uint16_t>() : nullptr;
        const uint16_t* lo = c->c16_lo[c->cur].as<uint16_t>();
        // variants (option pack16_variant = lanes per row + 4 * [plain read-modify-write instead of ATOMS]): measured
        // in profiles/, the default is PACK16_DEFAULT
        const int v = c->pack16_variant;
        auto pack16 = k_pack_sketch_rows16<4, 2, true>;
        int lanes = 2;
#define BF_PACK16(W)                                                                                     \
        switch (v) {                                                                                     \
            case 1: pack16 = k_pack_sketch_rows16<W, 1, true>; lanes = 1; break;                         \
            case 5: pack16 = k_pack_sketch_rows16<W, 1, false>; lanes = 1; break;                        \
            case 6: pack16 = k_pack_sketch_rows16<W, 2, false>; lanes = 2; break;                        \
            default: pack16 = k_pack_sketch_rows16<W, 2, true>; lanes = 2; break;                        \
        }
        if (words == 4) { BF_PACK16(4) } else { BF_PACK16(8) }
#undef BF_PACK16
        pack16<<<(unsigned)my_blocks, TILE * lanes, 0, c->stream>>>(ip, sp, lo, n, c->sk_rows.as<uint32_t>(), c->keysB[0].as<sortkey_t>(),
                                                                     c->valsB[0].as<int32_t>(), or_key, block0);
        CKLC(c);
    } else if (my_blocks > 0) {
        if (words == 4)
            k_pack_sketch_rows<4><<<(unsigned)my_blocks, 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, log2m, c->sk_rows.as<uint32_t>(),
                                                                             c->keysB[0].as<sortkey_t>(), c->valsB[0].as<int32_t>(), or_key, block0);
        else
            k_pack_sketch_rows<8><<<(unsigned)my_blocks, 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, log2m, c->sk_rows.as<uint32_t>(),
                                                                             c->keysB[0].as<sortkey_t>(), c->valsB[0].as<int32_t>(), or_key, block0);
        CKLC(c);
    }
    if (shard_pack(c)) {
        // exchange step 1: all-gather of the sketch and key shares (16 + 8 bytes per row at 128 bits) over NVLink,
        // in place; then row numbers and the OR word over all keys
        bfnccl::Api& nc = bfnccl::api();
        const size_t rows_share = (size_t)share * TILE;
        const size_t sk_bytes = rows_share * words * sizeof(uint32_t), key_bytes = rows_share * sizeof(sortkey_t);
        char* sk = c->sk_rows.as<char>();
        char* keys = c->keysB[0].as<char>();
        NCK(nc.GroupStart());
        NCK(nc.AllGather(sk + (size_t)c->rank * sk_bytes, sk, sk_bytes, bfnccl::kUint8, c->comm, c->stream));
        NCK(nc.AllGather(keys + (size_t)c->rank * key_bytes, keys, key_bytes, bfnccl::kUint8, c->comm, c->stream));
        NCK(nc.GroupEnd());
        CK(cudaMemsetAsync(or_key, 0, sizeof(sortkey_t), c->stream));
        k_keys_finalize<<<grid_for(n, 256), 256, 0, c->stream>>>(c->keysB[0].as<sortkey_t>(), n, c->valsB[0].as<int32_t>(), or_key);
        CKLC(c);
    }
    return BF_OK;
}

// exchange step 2 (after verify + hook): every rank publishes the rows of its union-find that are not their own root
// as (row, root) pairs, the lists are all-gathered and re-united on every rank -> identical forests everywhere
int exchange_labels(bf_ctx* c) {
    const int64_t n = c->n_rows;
    if (n == 0) return BF_OK;
    bfnccl::Api& nc = bfnccl::api();
    // a rank's list has about (edges / world) entries; first guess 1.25 N / world, afterwards what the data needed
    const unsigned long long cap = c->merge_capacity > 0 ? (unsigned long long)c->merge_capacity
                                   : c->merge_auto > 0   ? (unsigned long long)c->merge_auto
                                                         : (unsigned long long)std::max<int64_t>(4096, n / c->world + n / (4 * c->world));
    c->merge_cap_used = cap;
    const size_t words_rank = (size_t)cap + 1;
    TRY(c->xchg.ensure(words_rank * c->world * sizeof(unsigned long long)));
    unsigned long long* mine = c->xchg.as<unsigned long long>() + words_rank * c->rank;
    CK(cudaEventRecord(c->ev_aux[0], c->stream));
    CK(cudaMemsetAsync(mine, 0, sizeof(unsigned long long), c->stream));
    k_uf_compact<<<grid_for(n, 256), 256, 0, c->stream>>>(c->parent.as<int>(), n, mine, cap);
    CKLC(c);
    NCK(nc.AllGather(mine, c->xchg.p, words_rank * sizeof(unsigned long long), bfnccl::kUint8, c->comm, c->stream));
    k_uf_merge_pairs<<<grid_for((int64_t)cap * c->world, 256), 256, 0, c->stream>>>(
        c->parent.as<int>(), c->xchg.as<unsigned long long>(), c->world, c->rank, cap, &c->counters.as<DevCounters>()->merge_fullest);
    CKLC(c);
    CK(cudaEventRecord(c->ev_aux[1], c->stream));
    c->ms_merge = -1.f;  // resolved in bf_sync
    return BF_OK;
}

int pack_rows(bf_ctx* c, const int32_t* perm_dev, int64_t n, DevBuf& bits, DevBuf folds[2], DevBuf fold8[2], int* parent_init) {
    if (n == 0) return BF_OK;
    const int64_t tiles = ceil_div(n, TILE);
    const size_t bytes = (size_t)tiles * c->n_chunks * c->K4 * TILE * 16;
    TRY(bits.ensure(bytes));
    if (c->engine == BF_ENGINE_SKETCH) {
        int log2m = 0;
        while ((1 << log2m) < c->sketch_bits) ++log2m;
        if (staged_pack(c)) {   // step 2: staged sketches -> sorted tile layouts
            for (int f = 0; f < 2; ++f) TRY(folds[f].ensure((size_t)tiles * TILE * sizeof(uint32_t)));
            uint32_t* f8a = nullptr;   // expanded row operand (fragment order) and column operand (row-major)
            uint4* f8b = nullptr;
            if (c->level1 >= 1) {
                for (int f = 0; f < 2; ++f) TRY(fold8[f].ensure((size_t)(tiles + IMMA_GROUP) * TILE * 32));  // + slack: a bulk copy never crosses the end
                f8a = fold8[0].as<uint32_t>();
                f8b = fold8[1].as<uint4>();
            }
            uint32_t *b32 = bits.as<uint32_t>(), *fo0 = folds[0].as<uint32_t>(), *fo1 = folds[1].as<uint32_t>();
            if (c->sketch_bits == 128)
                k_permute_store<4><<<(unsigned)tiles, TILE, 0, c->stream>>>(c->sk_rows.as<uint32_t>(), perm_dev, n, b32, fo0, fo1, f8a, f8b, c->level1 == 2, parent_init);
            else
                k_permute_store<8><<<(unsigned)tiles, TILE, 0, c->stream>>>(c->sk_rows.as<uint32_t>(), perm_dev, n, b32, fo0, fo1, f8a, f8b, c->level1 == 2, parent_init);
        } else {
            const size_t smem = (size_t)c->n_chunks * c->K4 * TILE * 16;
            k_pack_sketch<<<(unsigned)tiles, 256, smem, c->stream>>>(c->d_indptr, c->d_indices, perm_dev, n, log2m,
                                                                     c->n_chunks, c->K4, bits.as<uint32_t>());
        }
        CKLC(c);
    } else {
        CK(cudaMemsetAsync(bits.p, 0, bytes, c->stream));
        k_pack_full<<<grid_for(n * 32, 256), 256, 0, c->stream>>>(c->d_indptr,
                                                                  c->d_indices, perm_dev, n,
                                                                  c->n_chunks, c->K4, bits.as<uint32_t>());
        CKLC(c);
    }
    return BF_OK;
}

template <int K4, int STAGES, bool TWO_LEVEL>
int launch_pairs(bf_ctx* c, const uint4* A, const uint4* B, int64_t nA, int64_t nB, int triangular) {
    using L = PairSmem<K4, STAGES>;
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        CK(cudaFuncSetAttribute(k_pairs<K4, STAGES, TWO_LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes));
        attr_set[c->device & 63] = true;
    }
    int bps = c->blocks_per_sm;
    if (bps <= 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_pairs<K4, STAGES, TWO_LEVEL>, PAIR_THREADS, L::kTotalBytes));
        bps = std::max(1, std::min(bps, 4));
    }
    const unsigned grid = (unsigned)(c->num_sms * bps);
    k_pairs<K4, STAGES, TWO_LEVEL><<<grid, PAIR_THREADS, L::kTotalBytes, c->stream>>>(
        A, B, c->n_chunks, nA, nB, c->items.as<int2>(), c->items_cap_used, c->wprefix.as<unsigned long long>(),
        c->jlo.as<int32_t>(), c->jend.as<int32_t>(), c->tilesA, c->sched_ranges, c->nwork.as<unsigned long long>(), c->max_dist, triangular, c->rank,
        c->world, c->cand.as<uint2>(), c->cand_cap_used, c->counters.as<DevCounters>());
    CKLC(c);
    return BF_OK;
}

#ifndef BF_STAGES_SMALL
#define BF_STAGES_SMALL 16
#endif
int dispatch_pairs(bf_ctx* c, const uint4* A, const uint4* B, int64_t nA, int64_t nB, int tri) {
    const bool two = c->ran_two_level;
    if (c->K4 == 1) return two ? launch_pairs<1, BF_STAGES_SMALL, true>(c, A, B, nA, nB, tri) : launch_pairs<1, BF_STAGES_SMALL, false>(c, A, B, nA, nB, tri);
    if (c->K4 == 2) return two ? launch_pairs<2, BF_STAGES_SMALL, true>(c, A, B, nA, nB, tri) : launch_pairs<2, BF_STAGES_SMALL, false>(c, A, B, nA, nB, tri);
    return two ? launch_pairs<4, 4, true>(c, A, B, nA, nB, tri) : launch_pairs<4, 4, false>(c, A, B, nA, nB, tri);
}

// level 1 on the tensor cores (mma.sync int8 on the +-1 expanded folds) + level 2 on the queue
constexpr int kL1CtasDefault = 2;
template <int MINB>
int launch_imma2(bf_ctx* c, const uint32_t* f8a, const uint4* f8p, int64_t nA, int64_t nB, int tri) {
    using Cfg = Imma2Cfg<MINB>;
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        CK(cudaFuncSetAttribute(k_pairs_l1_imma2<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set[c->device & 63] = true;
    }
    k_pairs_l1_imma2<MINB><<<(unsigned)c->l1_segs, PAIR_THREADS, Cfg::kSmemBytes, c->stream>>>(
        f8a, f8p, nA, nB, c->items.as<int2>(), c->items_cap_used, c->nwork.as<unsigned long long>(), c->max_dist, tri,
        c->rank, c->world, 66560, 63 * 66560, c->queue.as<int2>(), c->queue_cap_used, c->segcnt.as<unsigned>(), c->counters.as<DevCounters>());
    CKLC(c);
    return BF_OK;
}

int launch_two_kernel_imma(bf_ctx* c, const uint4* A, const uint4* B, int64_t nA, int64_t nB, int tri) {
    const uint32_t* f8a = (c->has_query ? c->fold8A[0] : c->fold8B[0]).as<uint32_t>();
    const uint4* f8b = c->fold8B[1].as<uint4>();
    const bool packed = c->level1 == 2;
    const int ctas = packed ? (c->l1_ctas > 0 ? c->l1_ctas : kL1CtasDefault) : 1;
    c->l1_segs = c->num_sms * ctas;
    TRY(c->segcnt.ensure((size_t)c->l1_segs * sizeof(unsigned)));
    if (packed) {
        if (ctas == 2) TRY(launch_imma2<2>(c, f8a, f8b, nA, nB, tri));
        else TRY(launch_imma2<1>(c, f8a, f8b, nA, nB, tri));
    } else {
        static bool attr_set[64] = {};
        if (!attr_set[c->device & 63]) {
            CK(cudaFuncSetAttribute(k_pairs_l1_imma, cudaFuncAttributeMaxDynamicSharedMemorySize, IMMA_SMEM_BYTES));
            attr_set[c->device & 63] = true;
        }
        k_pairs_l1_imma<<<(unsigned)c->l1_segs, PAIR_THREADS, IMMA_SMEM_BYTES, c->stream>>>(
            f8a, f8b, nA, nB, c->items.as<int2>(), c->items_cap_used, c->nwork.as<unsigned long long>(), c->max_dist, tri,
            c->rank, c->world, c->queue.as<int2>(), c->queue_cap_used, c->segcnt.as<unsigned>(), c->counters.as<DevCounters>());
        CKLC(c);
    }
    CK(cudaEventRecord(c->ring[c->runs_since_sync % bf_ctx::kRing][4], c->stream));
    const uint4* fa = (c->has_query ? c->foldsA[0] : c->foldsB[0]).as<uint4>();
    const uint2* fb = c->foldsB[1].as<uint2>();
    auto l2 = packed ? (c->K4 == 1 ? k_pairs_l2_unit<1, true> : k_pairs_l2_unit<2, true>)
                     : (c->K4 == 1 ? k_pairs_l2_unit<1, false> : k_pairs_l2_unit<2, false>);
    // CTAs of level 2 per queue segment (measured at 10^6 profiles with two level-1 CTAs per SM: 16 -> 42 us, 32 -> 44 us,
    // 64 -> 50 us; a segment then holds about 6 000 units)
    const int n_sub = c->l2_sub > 0 ? c->l2_sub : (packed ? std::max(1, L2_SUB / (2 * ctas)) : L2_SUB);
    l2<<<c->l1_segs * n_sub, 256, 0, c->stream>>>(A, B, fa, fb, nA, nB, c->queue.as<int2>(), c->queue_cap_used, c->segcnt.as<unsigned>(),
                                                  c->l1_segs, n_sub, c->max_dist, tri, c->cand.as<uint2>(), c->cand_cap_used,
                                                  c->counters.as<DevCounters>());
    CKLC(c);
    return BF_OK;
}

// level 1 on the fold planes (persistent, one CTA per SM) + level 2 on the queue
int launch_two_kernel(bf_ctx* c, const uint4* A, const uint4* B, int64_t nA, int64_t nB, int tri) {
    if (c->level1 >= 1) return launch_two_kernel_imma(c, A, B, nA, nB, tri);
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        CK(cudaFuncSetAttribute(k_pairs_l1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_pairs_l1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_pairs_l1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L1_SMEM_BYTES));
        attr_set[c->device & 63] = true;
    }
    const uint32_t* fa = (c->has_query ? c->foldsA[0] : c->foldsB[0]).as<uint32_t>();
    const uint32_t* fb = c->foldsB[1].as<uint32_t>();
    auto l1 = k_pairs_l1<0>;
    if (c->max_dist == 1) l1 = k_pairs_l1<1>;
    else if (c->max_dist == 2) l1 = k_pairs_l1<2>;
    int bps = c->blocks_per_sm;
    if (bps <= 0) {  // two CTAs per SM when the register file allows it: more warps to hide POPC/min latency
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, l1, PAIR_THREADS, L1_SMEM_BYTES));
        bps = std::max(1, std::min(bps, 2));
    }
    l1<<<(unsigned)(c->num_sms * bps), PAIR_THREADS, L1_SMEM_BYTES, c->stream>>>(
        fa, fb, c->items.as<int2>(), c->items_cap_used, c->nwork.as<unsigned long long>(), c->max_dist, c->rank,
        c->world, 1u, c->queue.as<int2>(), c->queue_cap_used, c->counters.as<DevCounters>());
    CKLC(c);
    CK(cudaEventRecord(c->ring[c->runs_since_sync % bf_ctx::kRing][4], c->stream));
    if (c->K4 == 1)
        k_pairs_l2<1><<<c->num_sms * 8, 256, 0, c->stream>>>(A, B, nA, nB, c->queue.as<int2>(), c->queue_cap_used, c->max_dist,
                                                            tri, c->cand.as<uint2>(), c->cand_cap_used, c->counters.as<DevCounters>());
    else
        k_pairs_l2<2><<<c->num_sms * 8, 256, 0, c->stream>>>(A, B, nA, nB, c->queue.as<int2>(), c->queue_cap_used, c->max_dist,
                                                            tri, c->cand.as<uint2>(), c->cand_cap_used, c->counters.as<DevCounters>());
    CKLC(c);
    return BF_OK;
}

// candidate pairs of the metric (||A| - |B|| <= max_dist): a statistic of the uploaded matrix, not of a pass - counted by
// the first pass after an upload and restored from a 16-byte device cache by the following ones
int band_statistic(bf_ctx* c, const sortkey_t* keysA, int64_t nA, const sortkey_t* keysB, int64_t nB, int max_dist) {
    DevCounters* dc = c->counters.as<DevCounters>();
    TRY(c->band_cache.ensure(2 * sizeof(unsigned long long)));
    if (c->band_valid && c->band_dist == max_dist) {
        CK(cudaMemcpyAsync(&dc->band_ab, c->band_cache.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        return BF_OK;
    }
    k_band_count<<<grid_for(nA, 256), 256, 0, c->stream>>>(keysA, nA, keysB, nB, max_dist, &dc->band_ab);
    CKLC(c);
    if (c->has_query) {
        k_band_count<<<grid_for(nA, 256), 256, 0, c->stream>>>(keysA, nA, keysA, nA, max_dist, &dc->band_aa);
        CKLC(c);
    }
    CK(cudaMemcpyAsync(c->band_cache.p, &dc->band_ab, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    c->band_valid = true;
    c->band_dist = max_dist;
    return BF_OK;
}

inline unsigned long long hj_capacity(int64_t want) {
    unsigned long long cap = 1024;
    while ((int64_t)cap < want) cap <<= 1;
    return cap;
}

// K7: the joins of the hash-join engine -> candidates (row numbers) for k_verify_unite.  This rank probes the rows
// r = rank (mod world); the tables hold all rows.
int run_hashjoin_probes(bf_ctx* c) {
    const int64_t n = c->n_rows;
    const int d = c->max_dist;
    const unsigned long long cap1 = hj_capacity(2 * n);
    const unsigned long long* H = c->hj_hash.as<unsigned long long>();
    const unsigned long long* t1 = c->hj_t1.as<unsigned long long>();
    const unsigned char* isq = c->has_query ? c->is_query.as<unsigned char>() : nullptr;
    uint2* cand = c->cand.as<uint2>();
    DevCounters* dc = c->counters.as<DevCounters>();
    const int64_t my_rows = ceil_div(n, c->world);
    k_hj_probe_equal<<<grid_for(my_rows, 256), 256, 0, c->stream>>>(c->d_indptr, n, H, t1, cap1 - 1, c->rank, c->world, isq, cand, c->cand_cap_used, dc);
    CKLC(c);
    if (d >= 1) {
        k_hj_probe_deletions<1><<<grid_for(my_rows * 8, 256), 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, H, t1, nullptr, cap1 - 1, c->rank,
                                                                                   c->world, isq, cand, c->cand_cap_used, dc);
        CKLC(c);
    }
    if (d >= 2) {
        k_hj_probe_pairs<<<grid_for(my_rows * 32, 256), 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, H, t1, cap1 - 1, c->rank, c->world, isq, cand,
                                                                             c->cand_cap_used, dc);
        CKLC(c);
        const unsigned long long cap2 = hj_capacity(2 * std::max<int64_t>(c->nnz, 1));
        TRY(c->hj_t2.ensure((size_t)cap2 * sizeof(unsigned long long)));
        TRY(c->hj_t2_rows.ensure((size_t)cap2 * sizeof(int32_t)));
        CK(cudaMemsetAsync(c->hj_t2.p, 0xff, (size_t)cap2 * sizeof(unsigned long long), c->stream));
        k_hj_build_deletions<<<grid_for(n * 8, 256), 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, H, c->hj_t2.as<unsigned long long>(),
                                                                          c->hj_t2_rows.as<int32_t>(), cap2 - 1);
        CKLC(c);
        k_hj_probe_deletions<2><<<grid_for(my_rows * 8, 256), 256, 0, c->stream>>>(c->d_indptr, c->d_indices, n, H, c->hj_t2.as<unsigned long long>(),
                                                                                   c->hj_t2_rows.as<int32_t>(), cap2 - 1, c->rank, c->world, isq, cand,
                                                                                   c->cand_cap_used, dc);
        CKLC(c);
    }
    return BF_OK;
}

int finish_labels(bf_ctx* c) {
    const int64_t n = c->n_rows;
    if (n == 0) return BF_OK;
    CK(cudaMemsetAsync(&c->counters.as<DevCounters>()->n_comp, 0, sizeof(unsigned int), c->stream));
    k_uf_labels<<<grid_for(n, 256), 256, 0, c->stream>>>(c->parent.as<int>(), n, c->labels.as<int32_t>(),
                                                         &c->counters.as<DevCounters>()->n_comp);
    CKLC(c);
    return BF_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int bf_abi_version(void) { return BF_ABI_VERSION; }
const char* bf_last_error(void) { return g_err.c_str(); }

int bf_device_count(int* n_out) {
    if (!n_out) return fail(BF_ERR_INVALID, "n_out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *n_out = 0;
        return fail(BF_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    *n_out = n;
    return BF_OK;
}

int bf_ctx_create(int device, void* stream, bf_ctx** ctx_out) {
    if (!ctx_out) return fail(BF_ERR_INVALID, "ctx_out is null");
    *ctx_out = nullptr;
    int n = 0;
    TRY(bf_device_count(&n));
    if (n <= 0) return fail(BF_ERR_NO_DEVICE, "no CUDA device visible; breakfast_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(BF_ERR_INVALID, "device index out of range");
    bf_ctx* c = new (std::nothrow) bf_ctx();
    if (!c) return fail(BF_ERR_OOM, "host allocation failed");
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess && prop.major < 9) {
        delete c;
        return fail(BF_ERR_NO_DEVICE, "device is older than sm_90: bulk-async copies/mbarrier unavailable (built for sm_100a)");
    }
    if (e == cudaSuccess) {
        c->num_sms = prop.multiProcessorCount;
        if (stream) {
            c->stream = (cudaStream_t)stream;
        } else {
            e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
            c->own_stream = true;
        }
    }
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev_aux[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_upload_start);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_upload_done);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_slot_free[i], cudaEventDisableTiming);
    for (int i = 0; i < bf_ctx::kRing && e == cudaSuccess; ++i)
        for (int j = 0; j < 5 && e == cudaSuccess; ++j) e = cudaEventCreate(&c->ring[i][j]);
    if (e == cudaSuccess) {
        int rc = c->counters.ensure(sizeof(DevCounters));
        if (rc == BF_OK) rc = c->nwork.ensure(sizeof(unsigned long long));
        if (rc != BF_OK) {
            bf_ctx_destroy(c);
            return rc;
        }
    }
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "bf_ctx_create", __LINE__);
        bf_ctx_destroy(c);
        return rc;
    }
    *ctx_out = c;
    return BF_OK;
}

void bf_ctx_destroy(bf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_copy) {
        if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
        bfnccl::api().CommDestroy(c->comm_copy);
        c->comm_copy = nullptr;
    }
    if (c->comm) {
        bfnccl::api().CommDestroy(c->comm);
        c->comm = nullptr;
    }
    DevBuf* bufs[] = {&c->indptr[0], &c->indptr[1], &c->indices[0], &c->indices[1], &c->query_rows, &c->is_query,
                      &c->hj_hash, &c->hj_t1, &c->hj_t2, &c->hj_t2_rows, &c->c16_indptr[0], &c->c16_indptr[1], &c->c16_split[0], &c->c16_split[1], &c->c16_lo[0], &c->c16_lo[1], &c->c16_bad, &c->keysB[0], &c->keysB[1], &c->keysB[2],
                      &c->valsB[0], &c->valsB[1], &c->valsB[2], &c->keysA[0], &c->keysA[1], &c->keysA[2], &c->valsA[0], &c->valsA[1], &c->valsA[2], &c->sched_table,
                      &c->sort_counts, &c->sort_max, &c->sk_rows, &c->xchg, &c->band_cache, &c->bitsA, &c->bitsB, &c->foldsA[0], &c->foldsA[1], &c->foldsB[0], &c->foldsB[1], &c->fold8A[0], &c->fold8A[1], &c->fold8B[0], &c->fold8B[1], &c->jlo, &c->jend, &c->queue, &c->segcnt, &c->wprefix, &c->nwork, &c->items, &c->cand,
                      &c->edges, &c->parent, &c->labels, &c->counters, &c->scratch, &c->scratch2};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_aux) if (e) cudaEventDestroy(e);
    for (auto& e : c->trace_ev) cudaEventDestroy(e);
    for (auto& r : c->ring) for (auto& e : r) if (e) cudaEventDestroy(e);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_upload_start) cudaEventDestroy(c->ev_upload_start);
    if (c->ev_upload_done) cudaEventDestroy(c->ev_upload_done);
    for (auto& e : c->ev_slot_free) if (e) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    (void)cudaGetLastError();
    delete c;
}

int bf_ctx_set_option(bf_ctx* c, const char* key, int64_t value) {
    if (!c || !key) return fail(BF_ERR_INVALID, "null argument");
    std::string k(key);
    if (k == "engine") {
        if (value != BF_ENGINE_SKETCH && value != BF_ENGINE_FULL && value != BF_ENGINE_HASHJOIN)
            return fail(BF_ERR_INVALID, "engine must be 0 (sketch), 1 (full) or 2 (hash join)");
        c->engine = (int)value;
    } else if (k == "sketch_bits") {
        if (value < 128 || value > 2048 || (value & (value - 1))) return fail(BF_ERR_INVALID, "sketch_bits must be a power of two in [128, 2048]");
        c->sketch_bits = (int)value;
    } else if (k == "want_edges") {
        c->want_edges = value ? 1 : 0;
    } else if (k == "cand_capacity") {
        if (value < 0) return fail(BF_ERR_INVALID, "cand_capacity must be >= 0");
        c->cand_capacity = value;
    } else if (k == "two_level") {
        c->two_level = value ? 1 : 0;
    } else if (k == "items_capacity") {
        if (value < 0) return fail(BF_ERR_INVALID, "items_capacity must be >= 0");
        c->items_capacity = value;
    } else if (k == "level1") {
        if (value < 0 || value > 2) return fail(BF_ERR_INVALID, "level1 must be 0 (integer pipes), 1 (tensor cores) or 2 (tensor cores, two column rows per accumulator)");
        c->level1 = (int)value;
    } else if (k == "l2_sub") {
        if (value < 0 || value > 1024) return fail(BF_ERR_INVALID, "l2_sub must be in [0, 1024]");
        c->l2_sub = (int)value;
    } else if (k == "l1_ctas") {
        if (value < 0 || value > 2) return fail(BF_ERR_INVALID, "l1_ctas must be 0 (default), 1 or 2");
        c->l1_ctas = (int)value;
    } else if (k == "merge_capacity") {
        if (value < 0) return fail(BF_ERR_INVALID, "merge_capacity must be >= 0");
        c->merge_capacity = value;
    } else if (k == "units_capacity") {
        if (value < 0) return fail(BF_ERR_INVALID, "units_capacity must be >= 0");
        c->units_capacity = value;
    } else if (k == "pack16_variant") {
        if (value != 1 && value != 2 && value != 5 && value != 6) return fail(BF_ERR_INVALID, "pack16_variant must be 1, 2, 5 or 6");
        c->pack16_variant = (int)value;
    } else if (k == "resident_csr16") {
        c->resident16 = value ? 1 : 0;   // 0: the sketch pass streams the plain CSR
    } else if (k == "shard_pack_from") {
        if (value < 2) return fail(BF_ERR_INVALID, "shard_pack_from must be >= 2");
        c->shard_pack_from = (int)value;
    } else if (k == "verify_csr16") {
        c->verify16 = value ? 1 : 0;
    } else if (k == "blocks_per_sm") {
        if (value < 0 || value > 8) return fail(BF_ERR_INVALID, "blocks_per_sm must be in [0, 8]");
        c->blocks_per_sm = (int)value;
    } else {
        return fail(BF_ERR_INVALID, "unknown option: " + k);
    }
    return BF_OK;
}

int bf_upload_csr(bf_ctx* c, const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols,
                  const int32_t* query_rows, int64_t n_query) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (n_rows < 0 || n_cols < 0 || n_rows > (int64_t)2147483647 - 2 * TILE) return fail(BF_ERR_INVALID, "n_rows/n_cols out of range");
    if (n_rows > 0 && !indptr) return fail(BF_ERR_INVALID, "indptr is null");
    if (query_rows == nullptr && n_query != 0 && n_query != n_rows) return fail(BF_ERR_INVALID, "n_query given without query_rows");
    if (query_rows && (n_query < 0 || n_query > n_rows)) return fail(BF_ERR_INVALID, "n_query out of range");
    int64_t nnz = 0;
    if (n_rows > 0) {
        if (indptr[0] != 0) return fail(BF_ERR_INVALID, "indptr[0] must be 0");
        for (int64_t i = 0; i < n_rows; ++i)
            if (indptr[i + 1] < indptr[i]) return fail(BF_ERR_INVALID, "indptr must be non-decreasing");
        nnz = indptr[n_rows];
        if (nnz > 0 && !indices) return fail(BF_ERR_INVALID, "indices is null");
    }
    if (query_rows) {
        for (int64_t q = 0; q < n_query; ++q) {
            if (query_rows[q] < 0 || query_rows[q] >= n_rows) return fail(BF_ERR_INVALID, "query row out of range");
            if (q && query_rows[q] <= query_rows[q - 1]) return fail(BF_ERR_INVALID, "query_rows must be ascending and unique");
        }
    }
    TRY(set_device(c));
    c->uploaded = false;
    c->ran = false;
    CK(cudaEventRecord(c->ev_aux[0], c->stream));
    if (c->pending >= 0) {  // an async upload is still in flight: let it land before reusing state
        CK(cudaStreamSynchronize(c->copy_stream));
        c->pending = -1;
    }
    DevBuf& dip = c->indptr[c->cur];
    DevBuf& dix = c->indices[c->cur];
    TRY(dip.ensure((size_t)(n_rows + 1) * sizeof(int64_t)));
    TRY(dix.ensure((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    if (n_rows > 0) {
        CK(cudaMemcpyAsync(dip.p, indptr, (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        if (nnz > 0) CK(cudaMemcpyAsync(dix.p, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    } else {
        int64_t zero = 0;
        CK(cudaMemcpyAsync(dip.p, &zero, sizeof zero, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    c->d_indptr = dip.as<int64_t>();
    c->d_indices = dix.as<int32_t>();
    c->active_slot = c->cur;
    c->c16_valid[c->cur] = false;
    bool encoded = false;
    if (c->resident16 && n_rows > 0 && nnz > 0 && n_cols <= 131072 && nnz < ((int64_t)1 << 32)) {
        TRY(encode_c16(c, c->cur, n_rows, nnz, c->stream));   // compact resident form, checked on the device
        encoded = true;
    }
    c->has_query = query_rows != nullptr;
    c->n_query = c->has_query ? n_query : n_rows;
    if (c->has_query) {
        TRY(c->query_rows.ensure((size_t)std::max<int64_t>(n_query, 1) * sizeof(int32_t)));
        TRY(c->is_query.ensure((size_t)std::max<int64_t>(n_rows, 1)));
        if (n_query > 0) CK(cudaMemcpyAsync(c->query_rows.p, query_rows, (size_t)n_query * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(c->is_query.p, 0, (size_t)std::max<int64_t>(n_rows, 1), c->stream));
        if (n_query > 0) {
            k_mark_rows<<<grid_for(n_query, 256), 256, 0, c->stream>>>(c->query_rows.as<int32_t>(), n_query, c->is_query.as<unsigned char>());
            CKLC(c);
        }
    }
    CK(cudaEventRecord(c->ev_aux[1], c->stream));
    // the caller's buffers may be pageable and are not retained: wait for the copies
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->ms_h2d, c->ev_aux[0], c->ev_aux[1]));
    if (encoded) {
        int bad = 1;
        CK(cudaMemcpy(&bad, c->c16_bad.p, sizeof bad, cudaMemcpyDeviceToHost));
        c->c16_valid[c->cur] = bad == 0;   // not representable after all: the plain form stays in charge
        c->c16_has_split[c->cur] = true;
    }
    c->n_rows = n_rows;
    c->n_cols = n_cols;
    c->nnz = nnz;
    c->uploaded = true;
    c->band_valid = false;
    return BF_OK;
}

int bf_upload_csr_async(bf_ctx* c, const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (n_rows <= 0 || n_cols < 0 || n_rows > (int64_t)2147483647 - 2 * TILE) return fail(BF_ERR_INVALID, "n_rows/n_cols out of range");
    if (!indptr) return fail(BF_ERR_INVALID, "indptr is null");
    if (indptr[0] != 0) return fail(BF_ERR_INVALID, "indptr[0] must be 0");
    for (int64_t i = 0; i < n_rows; ++i)
        if (indptr[i + 1] < indptr[i]) return fail(BF_ERR_INVALID, "indptr must be non-decreasing");
    const int64_t nnz = indptr[n_rows];
    if (nnz > 0 && !indices) return fail(BF_ERR_INVALID, "indices is null");
    TRY(set_device(c));
    if (c->pending >= 0) return fail(BF_ERR_STATE, "an async upload is already pending; call bf_run first");
    const int slot = c->active_slot >= 0 ? c->active_slot ^ 1 : c->cur;
    TRY(c->indptr[slot].ensure((size_t)(n_rows + 1) * sizeof(int64_t)));
    TRY(c->indices[slot].ensure((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    if (c->slot_used[slot]) CK(cudaStreamWaitEvent(c->copy_stream, c->ev_slot_free[slot], 0));
    c->c16_valid[slot] = false;   // (the plain form only: nothing checks the columns on this path)
    CK(cudaEventRecord(c->ev_upload_start, c->copy_stream));
    CK(cudaMemcpyAsync(c->indptr[slot].p, indptr, (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->copy_stream));
    if (nnz > 0) CK(cudaMemcpyAsync(c->indices[slot].p, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(c->ev_upload_done, c->copy_stream));
    c->pending = slot;
    c->pend_rows = n_rows;
    c->pend_cols = n_cols;
    c->pend_nnz = nnz;
    c->uploaded = true;
    c->band_valid = false;
    return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// compact host form ("CSR16")
// ------------------------------------------------------------------------------------------------
int bf_csr16_encode(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols, uint32_t* indptr32_out,
                    uint16_t* split_out, uint16_t* lo_out) {
    if (n_rows < 0 || n_cols < 0) return fail(BF_ERR_INVALID, "negative size");
    if (n_cols > 131072) return fail(BF_ERR_INVALID, "CSR16 needs n_cols <= 131072");
    if (n_rows > 0 && (!indptr || !indptr32_out)) return fail(BF_ERR_INVALID, "null argument");
    if (n_rows == 0) {
        if (indptr32_out) indptr32_out[0] = 0;
        return BF_OK;
    }
    const int64_t nnz = indptr[n_rows];
    if (indptr[0] != 0 || nnz < 0 || nnz > (int64_t)0xffffffffll) return fail(BF_ERR_INVALID, "CSR16 needs 0 <= nnz < 2^32 and indptr[0] == 0");
    if (nnz > 0 && (!indices || !lo_out)) return fail(BF_ERR_INVALID, "null argument");
    if (n_cols > 65536 && !split_out) return fail(BF_ERR_INVALID, "split_out is needed when n_cols > 65536");
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        if (e < b) return fail(BF_ERR_INVALID, "indptr must be non-decreasing");
        if (e - b > 65535) return fail(BF_ERR_INVALID, "CSR16 needs rows of at most 65535 columns");
        indptr32_out[r] = (uint32_t)b;
        uint32_t low = 0;
        for (int64_t k = b; k < e; ++k) {
            const uint32_t col = (uint32_t)indices[k];
            if (col >= (uint32_t)n_cols) return fail(BF_ERR_INVALID, "column index out of range");
            if (k > b && indices[k] <= indices[k - 1]) return fail(BF_ERR_INVALID, "columns of a row must be ascending and unique");
            low += col < 65536u ? 1u : 0u;
            lo_out[k] = (uint16_t)(col & 0xffffu);
        }
        if (split_out) split_out[r] = (uint16_t)low;
    }
    indptr32_out[n_rows] = (uint32_t)nnz;
    return BF_OK;
}

int bf_upload_csr16_async(bf_ctx* c, const uint32_t* indptr32, const uint16_t* split, const uint16_t* lo, int64_t n_rows,
                          int32_t n_cols) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (n_rows <= 0 || n_cols < 0 || n_cols > 131072 || n_rows > (int64_t)2147483647 - 2 * TILE) return fail(BF_ERR_INVALID, "n_rows/n_cols out of range for CSR16");
    if (!indptr32) return fail(BF_ERR_INVALID, "indptr32 is null");
    if (indptr32[0] != 0) return fail(BF_ERR_INVALID, "indptr32[0] must be 0");
    for (int64_t i = 0; i < n_rows; ++i)
        if (indptr32[i + 1] < indptr32[i]) return fail(BF_ERR_INVALID, "indptr32 must be non-decreasing");
    const int64_t nnz = indptr32[n_rows];
    if (nnz > 0 && !lo) return fail(BF_ERR_INVALID, "lo is null");
    if (n_cols > 65536 && !split) return fail(BF_ERR_INVALID, "split is needed when n_cols > 65536");
    TRY(set_device(c));
    if (c->pending >= 0) return fail(BF_ERR_STATE, "an async upload is already pending; call bf_run first");
    const int slot = c->active_slot >= 0 ? c->active_slot ^ 1 : c->cur;
    TRY(c->indptr[slot].ensure((size_t)(n_rows + 1) * sizeof(int64_t)));
    TRY(c->indices[slot].ensure((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    TRY(c->c16_indptr[slot].ensure((size_t)(n_rows + 1) * sizeof(uint32_t)));
    TRY(c->c16_lo[slot].ensure((size_t)(nnz + 8) * sizeof(uint16_t)));   // + 16 bytes: bulk copies end on a 16-byte boundary
    if (split) TRY(c->c16_split[slot].ensure((size_t)n_rows * sizeof(uint16_t)));
    if (c->slot_used[slot]) CK(cudaStreamWaitEvent(c->copy_stream, c->ev_slot_free[slot], 0));
    CK(cudaEventRecord(c->ev_upload_start, c->copy_stream));
    CK(cudaMemcpyAsync(c->c16_indptr[slot].p, indptr32, (size_t)(n_rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
    if (split) CK(cudaMemcpyAsync(c->c16_split[slot].p, split, (size_t)n_rows * sizeof(uint16_t), cudaMemcpyHostToDevice, c->copy_stream));
    if (nnz > 0) {
        if (c->comm_copy && c->comm_world > 1) {
            // multi-GPU job (every rank makes this call with the same matrix): this rank copies only its share of the
            // column array over its own host link, the shares are all-gathered over NVLink on the copy stream
            const int W = c->comm_world;
            const size_t share = (size_t)((ceil_div(nnz, W) + 7) & ~(int64_t)7);   // entries per rank, 16-byte aligned
            TRY(c->c16_lo[slot].ensure((share * W + 8) * sizeof(uint16_t)));
            const size_t lo0 = std::min<size_t>((size_t)nnz, share * c->comm_rank), lo1 = std::min<size_t>((size_t)nnz, lo0 + share);
            char* base = c->c16_lo[slot].as<char>();
            if (lo1 > lo0)
                CK(cudaMemcpyAsync(base + lo0 * sizeof(uint16_t), lo + lo0, (lo1 - lo0) * sizeof(uint16_t), cudaMemcpyHostToDevice, c->copy_stream));
            NCK(bfnccl::api().AllGather(base + share * c->comm_rank * sizeof(uint16_t), base, share * sizeof(uint16_t), bfnccl::kUint8,
                                        c->comm_copy, c->copy_stream));
        } else {
            CK(cudaMemcpyAsync(c->c16_lo[slot].p, lo, (size_t)nnz * sizeof(uint16_t), cudaMemcpyHostToDevice, c->copy_stream));
        }
    }
    // decode on the copy stream as well (it overlaps the pass that is still running on the other slot): the
    // verification and the other engines walk rows of the plain CSR; the compact form stays resident for the sketch pass
    k_csr16_decode<<<c->num_sms * 8, 256, 0, c->copy_stream>>>(c->c16_indptr[slot].as<uint32_t>(), split ? c->c16_split[slot].as<uint16_t>() : nullptr,
                                                               c->c16_lo[slot].as<uint16_t>(), n_rows, c->indptr[slot].as<int64_t>(),
                                                               c->indices[slot].as<int32_t>());
    CKLC(c);
    c->c16_valid[slot] = true;
    c->c16_has_split[slot] = split != nullptr;
    CK(cudaEventRecord(c->ev_upload_done, c->copy_stream));
    c->pending = slot;
    c->pend_rows = n_rows;
    c->pend_cols = n_cols;
    c->pend_nnz = nnz;
    c->uploaded = true;
    c->band_valid = false;
    return BF_OK;
}

int bf_adopt_csr_device(bf_ctx* c, const void* indptr_device, const void* indices_device, int64_t n_rows,
                        int32_t n_cols, int64_t nnz) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (n_rows < 0 || n_cols < 0 || nnz < 0 || n_rows > (int64_t)2147483647 - 2 * TILE) return fail(BF_ERR_INVALID, "size out of range");
    if (!indptr_device || (nnz > 0 && !indices_device)) return fail(BF_ERR_INVALID, "null device pointer");
    if (c->pending >= 0) return fail(BF_ERR_STATE, "an async upload is pending");
    c->d_indptr = static_cast<const int64_t*>(indptr_device);
    c->d_indices = static_cast<const int32_t*>(indices_device);
    c->active_slot = -1;
    c->n_rows = c->n_query = n_rows;
    c->n_cols = n_cols;
    c->nnz = nnz;
    c->has_query = false;
    c->uploaded = true;
    c->band_valid = false;
    c->ran = false;
    c->ms_h2d = 0;
    return BF_OK;
}

int bf_run(bf_ctx* c, int32_t max_dist, int32_t rank, int32_t world) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->uploaded) return fail(BF_ERR_STATE, "bf_run before bf_upload_csr");
    if (max_dist < 0) return fail(BF_ERR_INVALID, "max_dist must be >= 0");
    if (world < 1 || rank < 0 || rank >= world) return fail(BF_ERR_INVALID, "need 0 <= rank < world");
    if (c->comm && world > 1 && (rank != c->comm_rank || world != c->comm_world))
        return fail(BF_ERR_INVALID, "rank/world differ from the communicator of this context");
    if (c->comm && world > 1 && !bfnccl::api().ok) return fail(BF_ERR_STATE, bfnccl::api().error);
    if (c->engine == BF_ENGINE_HASHJOIN && max_dist > 2) return fail(BF_ERR_INVALID, "the hash-join engine covers max_dist 0, 1 and 2");
    TRY(set_device(c));
    if (c->pending >= 0) {
        // the CSR of this pass was uploaded asynchronously into the idle slot: order after the copy
        CK(cudaStreamWaitEvent(c->stream, c->ev_upload_done, 0));
        c->cur = c->pending;
        c->pending = -1;
        c->active_slot = c->cur;
        c->d_indptr = c->indptr[c->cur].as<int64_t>();
        c->d_indices = c->indices[c->cur].as<int32_t>();
        c->n_rows = c->n_query = c->pend_rows;
        c->n_cols = c->pend_cols;
        c->nnz = c->pend_nnz;
        c->has_query = false;
        c->ms_h2d = -1.f;  // resolved in bf_sync from the copy-stream events
    }
    c->max_dist = max_dist;
    c->rank = rank;
    c->world = world;
    c->ran = false;
    c->ms_merge = 0;
    const int64_t nB = c->n_rows, nA = c->n_query;
    // the staged sketch pass streams the compact form of the matrix where the slot holds it
    c->use16 = c->resident16 && staged_pack(c) && c->active_slot >= 0 && c->c16_valid[c->cur];

    const bool hashjoin = c->engine == BF_ENGINE_HASHJOIN;
    if (hashjoin) {
        c->K4 = 1;
        c->n_chunks = 1;
        c->bits_per_row = 64;   // the additive row hash
    } else if (c->engine == BF_ENGINE_SKETCH) {
        const int words = c->sketch_bits / 32;
        c->K4 = std::min(4, words / 4);
        c->n_chunks = words / (4 * c->K4);
        c->bits_per_row = c->sketch_bits;
    } else {
        c->K4 = 4;
        c->n_chunks = (int)std::max<int64_t>(1, ceil_div(c->n_cols, 512));
        c->bits_per_row = (int64_t)c->n_chunks * 512;
    }
    c->tilesA = ceil_div(nA, TILE);
    c->tilesB = ceil_div(nB, TILE);
    c->ran_two_kernel = c->engine == BF_ENGINE_SKETCH && c->two_level && (c->sketch_bits == 128 || c->sketch_bits == 256);
    c->ran_two_level = !c->ran_two_kernel && c->engine == BF_ENGINE_SKETCH && c->n_chunks == 1 && c->two_level;

    cudaEvent_t* ring = c->ring[c->runs_since_sync % bf_ctx::kRing];
    c->trace_n = 0;
    trace_mark(c, __LINE__);
    CK(cudaEventRecord(c->ev[0], c->stream));
    CK(cudaEventRecord(ring[0], c->stream));
    CK(cudaMemsetAsync(c->counters.p, 0, sizeof(DevCounters), c->stream));
    CK(cudaMemsetAsync(c->nwork.p, 0, sizeof(unsigned long long), c->stream));
    TRY(c->parent.ensure((size_t)std::max<int64_t>(nB, 1) * sizeof(int)));
    TRY(c->labels.ensure((size_t)std::max<int64_t>(nB, 1) * sizeof(int32_t)));
    const bool active = nA > 0 && nB > 0;
    // the staged pack of the B side (k_permute_store, one thread per row) initialises the union-find as well
    const bool init_in_pack = active && c->engine == BF_ENGINE_SKETCH && staged_pack(c);
    if (nB > 0 && !init_in_pack) {
        k_uf_init<<<grid_for(nB, 256), 256, 0, c->stream>>>(c->parent.as<int>(), nB);
        CKLC(c);
    }

    DevBuf* keysA = c->has_query ? c->keysA : c->keysB;
    DevBuf* valsA = c->has_query ? c->valsA : c->valsB;
    if (active) {
        // ---- K2: sort keys (+ the staged sketches, which come out of the same pass over the columns) and sort
        TRY(ensure_sort_buffers(c, std::max(nB, staged_pack(c) ? staged_rows(c) : nB), c->keysB, c->valsB));
        if (c->has_query) TRY(ensure_sort_buffers(c, nA, c->keysA, c->valsA));
        const bool staged = staged_pack(c);
        if (hashjoin) {   // rows are only sorted by cardinality here, for the candidate-pair statistic of the metric
            TRY(sort_by_card(c, nullptr, nB, c->keysB, c->valsB, 0, false));
            if (c->has_query) TRY(sort_by_card(c, c->query_rows.as<int32_t>(), nA, c->keysA, c->valsA, 1, false));
            TRY(band_statistic(c, keysA[0].as<sortkey_t>(), nA, c->keysB[0].as<sortkey_t>(), nB, max_dist));
        } else if (staged) {
            TRY(pack_stage_all_rows(c));
            if (c->has_query) {   // before the sort of the B side reuses keysB[0]
                k_gather_keys<<<grid_for(nA, 256), 256, 0, c->stream>>>(c->keysB[0].as<sortkey_t>(), c->query_rows.as<int32_t>(), nA,
                                                                        c->keysA[0].as<sortkey_t>(), c->valsA[0].as<int32_t>(),
                                                                        c->sort_max.as<sortkey_t>() + 1);
                CKLC(c);
            }
        }
        if (!hashjoin) {
            TRY(sort_by_card(c, nullptr, nB, c->keysB, c->valsB, 0, staged));
            if (c->has_query) TRY(sort_by_card(c, c->query_rows.as<int32_t>(), nA, c->keysA, c->valsA, 1, staged));
        }
    }
    CK(cudaEventRecord(c->ev[1], c->stream));
    if (active && hashjoin) {
        // ---- K7: row hashes + the row table
        const unsigned long long cap1 = hj_capacity(2 * nB);
        TRY(c->hj_hash.ensure((size_t)nB * sizeof(unsigned long long)));
        TRY(c->hj_t1.ensure((size_t)cap1 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(c->hj_t1.p, 0xff, (size_t)cap1 * sizeof(unsigned long long), c->stream));
        k_hj_hash_rows<<<grid_for(nB * 8, 256), 256, 0, c->stream>>>(c->d_indptr, c->d_indices, nB, c->hj_hash.as<unsigned long long>(),
                                                                     c->hj_t1.as<unsigned long long>(), cap1 - 1);
        CKLC(c);
    }
    if (active && !hashjoin) {
        // ---- K1: bit-pack
        TRY(pack_rows(c, c->valsB[0].as<int32_t>(), nB, c->bitsB, c->foldsB, c->fold8B, init_in_pack ? c->parent.as<int>() : nullptr));
        if (c->has_query) TRY(pack_rows(c, c->valsA[0].as<int32_t>(), nA, c->bitsA, c->foldsA, c->fold8A, nullptr));
    }
    CK(cudaEventRecord(c->ev[2], c->stream));
    if (active && hashjoin) {
        unsigned long long cap = c->cand_capacity > 0 ? (unsigned long long)c->cand_capacity
                                                      : (unsigned long long)std::max<int64_t>((int64_t)1 << 22, 8 * nB);
        TRY(c->cand.ensure((size_t)cap * sizeof(uint2)));
        c->cand_cap_used = cap;
        if (c->want_edges) TRY(c->edges.ensure((size_t)cap * sizeof(uint2)));
    }
    if (active && !hashjoin) {
        // ---- K2b: schedule
        // three sort keys where the sketch pass produced them and the (D, Ds) table stays small, else the plain band
        const int n_keys = staged_pack(c) ? (max_dist <= kThreeKeyMaxDist ? 3 : 2) : 1;
        if (n_keys == 3 && c->sched_table_dist != max_dist) {
            const std::vector<SchedRange> table = make_sched_table(max_dist);
            TRY(c->sched_table.ensure(std::max<size_t>(table.size(), 1) * sizeof(SchedRange)));
            CK(cudaMemcpyAsync(c->sched_table.p, table.data(), table.size() * sizeof(SchedRange), cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));   // `table` is pageable and goes out of scope (once per max_dist)
            c->sched_table_dist = max_dist;
            c->sched_table_n = (int)table.size();
        }
        const int n_ranges = n_keys == 3 ? std::max(c->sched_table_n, 2 * max_dist + 1) : (n_keys == 2 ? 2 * max_dist + 1 : 1);
        const int64_t n_entries = c->tilesA * n_ranges;
        c->sched_ranges = n_ranges;
        TRY(c->jlo.ensure((size_t)n_entries * sizeof(int32_t)));
        TRY(c->wprefix.ensure((size_t)(c->tilesA + 1) * sizeof(unsigned long long)));
        TRY(c->jend.ensure((size_t)n_entries * sizeof(int32_t)));
        const int group = c->ran_two_kernel ? (c->level1 >= 1 ? IMMA_GROUP : L1_GROUP) : 1;
        k_schedule<<<grid_for(c->tilesA, std::max(1, SCHED_THREADS / n_ranges)), SCHED_THREADS, 0, c->stream>>>(
            keysA[0].as<sortkey_t>(), nA, c->keysB[0].as<sortkey_t>(), nB, max_dist, c->has_query ? 0 : 1, group, n_keys, n_ranges,
            c->sched_table.as<SchedRange>(), n_keys == 3 ? c->sched_table_n : 0,
            c->jlo.as<int32_t>(), c->jend.as<int32_t>(), c->wprefix.as<unsigned long long>(),
            &c->counters.as<DevCounters>()->n_tilepairs, &c->counters.as<DevCounters>()->sched_done, c->nwork.as<unsigned long long>());
        CKLC(c);   // (its last block also scans the item counts: wprefix and the item total are ready)
        TRY(band_statistic(c, keysA[0].as<sortkey_t>(), nA, c->keysB[0].as<sortkey_t>(), nB, max_dist));
        // explicit work list for the producer (bounded; items beyond it fall back to a binary search)
        {
            const unsigned long long worst = c->has_query ? (unsigned long long)c->tilesA * c->tilesB
                                                          : (unsigned long long)c->tilesB * (c->tilesB + 1) / 2;
            unsigned long long icap = c->items_capacity > 0 ? (unsigned long long)c->items_capacity
                                                            : std::min<unsigned long long>(worst, 1ull << 24);
            icap = std::max<unsigned long long>(icap, 1);
            TRY(c->items.ensure((size_t)icap * sizeof(int2)));
            c->items_cap_used = icap;
            k_expand_items<<<c->num_sms * 8, 256, 0, c->stream>>>(c->wprefix.as<unsigned long long>(), c->jlo.as<int32_t>(),
                                                                  c->tilesA, n_ranges, c->nwork.as<unsigned long long>(), icap,
                                                                  c->items.as<int2>(), group, c->jend.as<int32_t>());
            CKLC(c);
        }
        // candidate buffer
        unsigned long long cap = c->cand_capacity > 0 ? (unsigned long long)c->cand_capacity
                                                      : (unsigned long long)std::max<int64_t>((int64_t)1 << 22, 8 * nB);
        TRY(c->cand.ensure((size_t)cap * sizeof(uint2)));
        c->cand_cap_used = cap;
        if (c->want_edges) TRY(c->edges.ensure((size_t)cap * sizeof(uint2)));
        if (c->ran_two_kernel) {
            unsigned long long qcap = c->units_capacity > 0 ? (unsigned long long)c->units_capacity : (1ull << 23);
            TRY(c->queue.ensure((size_t)qcap * sizeof(int2)));
            TRY(c->segcnt.ensure((size_t)c->num_sms * sizeof(unsigned)));
            c->queue_cap_used = qcap;
        }
    }
    CK(cudaEventRecord(c->ev[3], c->stream));
    CK(cudaEventRecord(ring[1], c->stream));
    if (active && hashjoin) TRY(run_hashjoin_probes(c));
    if (active && !hashjoin) {
        // ---- K3: pairs
        const uint4* A = (c->has_query ? c->bitsA : c->bitsB).as<uint4>();
        const uint4* B = c->bitsB.as<uint4>();
        const int tri = c->has_query ? 0 : 1;
        if (c->ran_two_kernel) {
            TRY(launch_two_kernel(c, A, B, nA, nB, tri));
        } else {
            TRY(dispatch_pairs(c, A, B, nA, nB, tri));
        }
    }
    if (!(active && c->ran_two_kernel)) CK(cudaEventRecord(ring[4], c->stream));  // no separate level-1 kernel ran
    CK(cudaEventRecord(c->ev[4], c->stream));
    CK(cudaEventRecord(ring[2], c->stream));
    if (active) {
        // ---- K3b: verify + hook
        const int32_t* pA = hashjoin ? nullptr : valsA[0].as<int32_t>();
        const int32_t* pB = hashjoin ? nullptr : c->valsB[0].as<int32_t>();
        const unsigned char* isq = (c->has_query && !hashjoin) ? c->is_query.as<unsigned char>() : nullptr;
        // rows from the compact resident form where the slot holds it (half the bytes per column), else the plain CSR
        const bool rows16 = c->verify16 && c->resident16 && c->active_slot >= 0 && c->c16_valid[c->cur];
        auto launch_verify = [&](auto kernel, auto rows) -> int {
            int vbps = 0;   // grid-stride kernel: launch exactly the blocks that are resident at once
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&vbps, kernel, 256, 0));
            kernel<<<c->num_sms * std::max(1, vbps), 256, 0, c->stream>>>(
                c->cand.as<uint2>(), c->cand_cap_used, pA, pB, rows, max_dist, c->engine == BF_ENGINE_FULL ? 1 : 0, isq,
                c->parent.as<int>(), c->want_edges ? c->edges.as<uint2>() : nullptr, c->cand_cap_used, c->counters.as<DevCounters>());
            return BF_OK;
        };
        // window = max_dist: a wider compile-time window would still be exact
        if (rows16) {
            const RowStore16 rows{c->c16_indptr[c->cur].as<uint32_t>(),
                                  c->c16_has_split[c->cur] ? c->c16_split[c->cur].as<uint16_t>() : nullptr, c->c16_lo[c->cur].as<uint16_t>()};
            if (max_dist == 1) TRY(launch_verify(k_verify_unite<1, RowStore16>, rows));
            else if (max_dist == 2) TRY(launch_verify(k_verify_unite<2, RowStore16>, rows));
            else if (max_dist == 3) TRY(launch_verify(k_verify_unite<3, RowStore16>, rows));
            else TRY(launch_verify(k_verify_unite<0, RowStore16>, rows));
        } else {
            const RowStore rows{c->d_indptr, c->d_indices};
            if (max_dist == 1) TRY(launch_verify(k_verify_unite<1, RowStore>, rows));
            else if (max_dist == 2) TRY(launch_verify(k_verify_unite<2, RowStore>, rows));
            else if (max_dist == 3) TRY(launch_verify(k_verify_unite<3, RowStore>, rows));
            else TRY(launch_verify(k_verify_unite<0, RowStore>, rows));
        }
        CKLC(c);
    }
    CK(cudaEventRecord(c->ev[5], c->stream));
    // ---- exchange step 2 (multi-GPU with the library's own communicator): merge the ranks' forests
    if (dist_run(c)) TRY(exchange_labels(c));
    // ---- K4: labels
    TRY(finish_labels(c));
    CK(cudaEventRecord(c->ev[6], c->stream));
    CK(cudaEventRecord(ring[3], c->stream));
    if (c->active_slot >= 0) {
        CK(cudaEventRecord(c->ev_slot_free[c->cur], c->stream));  // the idle slot may be refilled after this
        c->slot_used[c->cur] = true;
    }
    ++c->runs_since_sync;
    c->ran = true;
    return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// communicators
// ------------------------------------------------------------------------------------------------
int bf_comm_unique_id(void* id_out) {
    if (!id_out) return fail(BF_ERR_INVALID, "id_out is null");
    bfnccl::Api& nc = bfnccl::api();
    if (!nc.ok) return fail(BF_ERR_STATE, nc.error);
    bfnccl::unique_id id;
    NCK(nc.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return BF_OK;
}

int bf_ctx_comm_init_rank(bf_ctx* c, const void* id, int32_t rank, int32_t world) {
    if (!c || !id) return fail(BF_ERR_INVALID, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(BF_ERR_INVALID, "need 0 <= rank < world");
    if (c->comm) return fail(BF_ERR_STATE, "the context already has a communicator");
    bfnccl::Api& nc = bfnccl::api();
    if (!nc.ok) return fail(BF_ERR_STATE, nc.error);
    TRY(set_device(c));
    bfnccl::unique_id uid;
    memcpy(&uid, id, sizeof uid);
    NCK(nc.CommInitRank(&c->comm, world, uid, rank));
    c->comm_rank = rank;
    c->comm_world = world;
    if (nc.CommSplit && world > 1) NCK(nc.CommSplit(c->comm, 0, rank, &c->comm_copy, nullptr));   // collective: every rank does it
    return BF_OK;
}

int bf_comm_init_all(bf_ctx** ctxs, int32_t n) {
    if (!ctxs || n < 1 || n > 64) return fail(BF_ERR_INVALID, "need 1 .. 64 contexts");
    int devs[64];
    bfnccl::comm_t comms[64];
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return fail(BF_ERR_INVALID, "null context");
        if (ctxs[i]->comm) return fail(BF_ERR_STATE, "a context already has a communicator");
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(BF_ERR_INVALID, "NCCL needs one context per distinct device");
    }
    bfnccl::Api& nc = bfnccl::api();
    if (!nc.ok) return fail(BF_ERR_STATE, nc.error);
    NCK(nc.CommInitAll(comms, n, devs));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->comm_rank = i;
        ctxs[i]->comm_world = n;
    }
    if (nc.CommSplit && n > 1) {   // duplicates for the copy streams
        NCK(nc.GroupStart());
        for (int i = 0; i < n; ++i) NCK(nc.CommSplit(ctxs[i]->comm, 0, i, &ctxs[i]->comm_copy, nullptr));
        NCK(nc.GroupEnd());
    }
    return BF_OK;
}

int bf_ctx_comm_destroy(bf_ctx* c) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->comm) return BF_OK;
    TRY(set_device(c));
    CK(cudaStreamSynchronize(c->stream));
    if (c->comm_copy) {
        CK(cudaStreamSynchronize(c->copy_stream));
        NCK(bfnccl::api().CommDestroy(c->comm_copy));
        c->comm_copy = nullptr;
    }
    NCK(bfnccl::api().CommDestroy(c->comm));
    c->comm = nullptr;
    c->comm_rank = 0;
    c->comm_world = 1;
    return BF_OK;
}

int bf_labels_to_device(bf_ctx* c, void* dst_device) {
    if (!c || !dst_device) return fail(BF_ERR_INVALID, "null argument");
    if (!c->ran) return fail(BF_ERR_STATE, "no run to take labels from");
    TRY(set_device(c));
    if (c->n_rows > 0) CK(cudaMemcpyAsync(dst_device, c->labels.p, (size_t)c->n_rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    return BF_OK;
}

int bf_merge_labels_device(bf_ctx* c, const void* gathered_device, int32_t world) {
    if (!c || !gathered_device) return fail(BF_ERR_INVALID, "null argument");
    if (!c->ran) return fail(BF_ERR_STATE, "merge before run");
    if (world < 1) return fail(BF_ERR_INVALID, "world must be >= 1");
    TRY(set_device(c));
    CK(cudaEventRecord(c->ev_aux[0], c->stream));
    if (c->n_rows > 0) {
        k_uf_merge_labels<<<grid_for(c->n_rows * world, 256), 256, 0, c->stream>>>(
            c->parent.as<int>(), static_cast<const int32_t*>(gathered_device), c->n_rows, world);
        CKLC(c);
    }
    TRY(finish_labels(c));
    CK(cudaEventRecord(c->ev_aux[1], c->stream));
    c->ms_merge = -1.f;  // resolved in bf_sync
    return BF_OK;
}

int bf_merge_labels_host(bf_ctx* c, const int32_t* gathered_host, int32_t world) {
    if (!c || !gathered_host) return fail(BF_ERR_INVALID, "null argument");
    if (!c->ran) return fail(BF_ERR_STATE, "merge before run");
    if (world < 1) return fail(BF_ERR_INVALID, "world must be >= 1");
    TRY(set_device(c));
    const size_t bytes = (size_t)c->n_rows * world * sizeof(int32_t);
    TRY(c->scratch.ensure(std::max<size_t>(bytes, 4)));
    if (bytes) CK(cudaMemcpyAsync(c->scratch.p, gathered_host, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = bf_merge_labels_device(c, c->scratch.p, world);
    if (rc != BF_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_union_lists(bf_ctx* c, const int64_t* list_indptr, const int32_t* list_members, int64_t n_lists) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->ran) return fail(BF_ERR_STATE, "bf_union_lists before run");
    if (n_lists < 0) return fail(BF_ERR_INVALID, "n_lists < 0");
    if (n_lists == 0) return BF_OK;
    if (!list_indptr) return fail(BF_ERR_INVALID, "list_indptr is null");
    const int64_t n_members = list_indptr[n_lists];
    if (list_indptr[0] != 0 || n_members < 0) return fail(BF_ERR_INVALID, "bad list_indptr");
    for (int64_t l = 0; l < n_lists; ++l)
        if (list_indptr[l + 1] < list_indptr[l]) return fail(BF_ERR_INVALID, "list_indptr must be non-decreasing");
    if (n_members == 0) return BF_OK;
    if (!list_members) return fail(BF_ERR_INVALID, "list_members is null");
    for (int64_t e = 0; e < n_members; ++e)
        if (list_members[e] < 0 || list_members[e] >= c->n_rows) return fail(BF_ERR_INVALID, "list member out of range");
    TRY(set_device(c));
    TRY(c->scratch.ensure((size_t)(n_lists + 1) * sizeof(int64_t)));
    TRY(c->scratch2.ensure((size_t)n_members * sizeof(int32_t)));
    CK(cudaMemcpyAsync(c->scratch.p, list_indptr, (size_t)(n_lists + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->scratch2.p, list_members, (size_t)n_members * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    k_uf_lists<<<grid_for(n_members, 256), 256, 0, c->stream>>>(c->parent.as<int>(), c->scratch.as<int64_t>(), c->scratch2.as<int32_t>(), n_lists, n_members);
    CKLC(c);
    TRY(finish_labels(c));
    CK(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_sync(bf_ctx* c, bf_stats* st) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->ran) return fail(BF_ERR_STATE, "bf_sync before bf_run");
    TRY(set_device(c));
    CK(cudaStreamSynchronize(c->stream));
    trace_dump(c);
    DevCounters h;
    unsigned long long nwork = 0;
    CK(cudaMemcpy(&h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&nwork, c->nwork.p, sizeof nwork, cudaMemcpyDeviceToHost));
    const bool overflow = h.n_cand > c->cand_cap_used && c->n_query > 0 && c->n_rows > 0;
    if (st) {
        memset(st, 0, sizeof *st);
        const int64_t N = c->n_rows, Q = c->n_query;
        st->n_rows = N;
        st->n_query = Q;
        st->n_cols = c->n_cols;
        st->nnz = c->nnz;
        st->bits_per_row = c->bits_per_row;
        if (!c->has_query) {
            st->pairs_total = N * (N - 1) / 2;
            st->pairs_band = ((int64_t)h.band_ab - N) / 2;
            st->tiles_total = c->tilesB * (c->tilesB + 1) / 2;
        } else {
            st->pairs_total = Q * N - Q - Q * (Q - 1) / 2;
            st->pairs_band = (int64_t)h.band_ab - Q - ((int64_t)h.band_aa - Q) / 2;
            st->tiles_total = c->tilesA * c->tilesB;
        }
        st->tiles_band = (int64_t)h.n_tilepairs;
        if (c->ran_two_kernel) st->tiles_rank = (int64_t)h.tilepairs_rank;
        else st->tiles_rank = (int64_t)nwork > c->rank ? ((int64_t)nwork - c->rank + c->world - 1) / c->world : 0;
        st->pairs_evaluated = st->tiles_rank * TILE * TILE;
        st->n_candidates = (int64_t)h.n_cand;
        st->n_edges = (int64_t)h.n_edges;
        st->n_components = h.n_comp;
        {
            const int64_t words = c->bits_per_row / 32;
            if (c->ran_two_kernel) {
                // level 1: one 32-bit test per pair, half of them by POPC when max_dist is 1 or 2 (the other
                // half runs POPC-free on the ALU/FMA pipes); level 2: `words` POPC per pair of every queued 32-pair unit
                st->l2_warp_items = (int64_t)std::min<unsigned long long>(h.n_units, c->queue_cap_used);
                if (c->level1 >= 1) {  // tensor-core level 1: no POPC there; level 2 = the exact 32-bit test of a unit's
                                       // 32 pairs + `words` POPC for every pair that passes it (counted on the device)
                    st->popc32_executed = st->l2_warp_items * 32 + (int64_t)h.l2_warp_items * words;
                } else {
                    const int64_t l1 = (c->max_dist == 1 || c->max_dist == 2) ? st->pairs_evaluated / 2 : st->pairs_evaluated;
                    st->popc32_executed = l1 + st->l2_warp_items * 32 * words;
                }
            } else {
                st->l2_warp_items = (int64_t)h.l2_warp_items;
                st->popc32_executed = c->ran_two_level ? st->pairs_evaluated + st->l2_warp_items * 1024 * words
                                                       : st->pairs_evaluated * words;
            }
        }
        float ms = 0;
        if (c->ms_h2d < 0) {
            // copy-stream events may already belong to the next (still running) async upload
            c->ms_h2d = 0;
            if (cudaEventQuery(c->ev_upload_done) == cudaSuccess &&
                cudaEventElapsedTime(&ms, c->ev_upload_start, c->ev_upload_done) == cudaSuccess)
                c->ms_h2d = ms;
            (void)cudaGetLastError();
        }
        st->ms_h2d = c->ms_h2d;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); st->ms_sort = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); st->ms_pack = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); st->ms_sort += ms;
        CK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); st->ms_pairs = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[4], c->ev[5])); st->ms_verify = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6])); st->ms_cc = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[6])); st->ms_total = ms;
        if (c->ms_merge < 0) {
            CK(cudaEventElapsedTime(&ms, c->ev_aux[0], c->ev_aux[1]));
            st->ms_merge = ms;
        }
        st->ms_d2h = c->ms_d2h;
        st->runs_since_sync = c->runs_since_sync;
        st->kernel_launches = c->launches_since_sync;
        const int64_t covered = std::min<int64_t>(c->runs_since_sync, bf_ctx::kRing);
        for (int64_t k = 0; k < covered; ++k) {
            cudaEvent_t* r = c->ring[(c->runs_since_sync - 1 - k) % bf_ctx::kRing];
            CK(cudaEventElapsedTime(&ms, r[1], r[2])); st->ms_pairs_sum += ms;
            if (c->ran_two_kernel) { CK(cudaEventElapsedTime(&ms, r[1], r[4])); st->ms_l1_sum += ms; }
            CK(cudaEventElapsedTime(&ms, r[0], r[3])); st->ms_total_sum += ms;
        }
    }
    c->runs_since_sync = 0;
    c->launches_since_sync = 0;
    // Every bounded buffer is checked here; whatever overflowed is given the capacity this pass asked for (a later
    // buffer of the chain may have seen only part of its input, so another round can follow) and the caller reruns.
    {
        std::string what;
        char buf[200];
        if (dist_run(c) && c->n_rows > 0) {
            if (h.merge_fullest > c->merge_cap_used) {
                snprintf(buf, sizeof buf, "label exchange: %u entries > capacity %llu; ", h.merge_fullest, c->merge_cap_used);
                what += buf;
                if (c->merge_capacity > 0) c->merge_capacity = (int64_t)h.merge_fullest + h.merge_fullest / 4 + 1024;
            }
            c->merge_auto = (int64_t)h.merge_fullest + h.merge_fullest / 4 + 1024;   // every rank sees every count: same value everywhere
        }
        if (c->ran_two_kernel && c->n_query > 0 && c->n_rows > 0) {
            if (nwork > c->items_cap_used) {
                c->items_capacity = (int64_t)(nwork + 1024);
                snprintf(buf, sizeof buf, "work list: %llu items > capacity %llu; ", (unsigned long long)nwork, c->items_cap_used);
                what += buf;
            }
            // the tensor-core level 1 cuts the queue into one segment per CTA: the fullest segment decides
            const unsigned long long seg_need = c->level1 >= 1 ? (unsigned long long)h.seg_max * (unsigned long long)std::max(c->l1_segs, 1) : 0;
            if (h.n_units > c->queue_cap_used || seg_need > c->queue_cap_used) {
                const unsigned long long need = std::max<unsigned long long>(h.n_units, seg_need);
                c->units_capacity = (int64_t)(need + need / 4 + 1024 + c->num_sms);
                snprintf(buf, sizeof buf, "level-2 queue: %llu units > capacity %llu; ", (unsigned long long)h.n_units, c->queue_cap_used);
                what += buf;
            }
        }
        if (overflow) {
            c->cand_capacity = std::max<int64_t>(c->cand_capacity, (int64_t)(h.n_cand + h.n_cand / 4 + 1024));
            snprintf(buf, sizeof buf, "candidate buffer: %llu candidates > capacity %llu; ", (unsigned long long)h.n_cand, c->cand_cap_used);
            what += buf;
        }
        if (!what.empty()) return fail(BF_ERR_OVERFLOW, "overflow - " + what + "capacities raised, run again");
    }
    return BF_OK;
}

int bf_download_labels(bf_ctx* c, int32_t* labels_out) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->ran) return fail(BF_ERR_STATE, "no labels: run first");
    if (c->n_rows == 0) return BF_OK;
    if (!labels_out) return fail(BF_ERR_INVALID, "labels_out is null");
    TRY(set_device(c));
    CK(cudaEventRecord(c->ev_aux[0], c->stream));
    CK(cudaMemcpyAsync(labels_out, c->labels.p, (size_t)c->n_rows * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev_aux[1], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->ms_d2h, c->ev_aux[0], c->ev_aux[1]));
    return BF_OK;
}

int bf_edge_count(bf_ctx* c, int64_t* n_edges_out) {
    if (!c || !n_edges_out) return fail(BF_ERR_INVALID, "null argument");
    if (!c->ran) return fail(BF_ERR_STATE, "no run");
    TRY(set_device(c));
    CK(cudaStreamSynchronize(c->stream));
    DevCounters h;
    CK(cudaMemcpy(&h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost));
    *n_edges_out = (int64_t)h.n_edges;
    return BF_OK;
}

int bf_download_edges(bf_ctx* c, int32_t* src_out, int32_t* dst_out) {
    if (!c) return fail(BF_ERR_INVALID, "ctx is null");
    if (!c->ran || !c->want_edges) return fail(BF_ERR_STATE, "edges need option want_edges=1 before bf_run");
    int64_t n = 0;
    TRY(bf_edge_count(c, &n));
    if (n == 0) return BF_OK;
    if ((unsigned long long)n > c->cand_cap_used) return fail(BF_ERR_OVERFLOW, "edge buffer overflow");
    if (!src_out || !dst_out) return fail(BF_ERR_INVALID, "null output");
    std::vector<uint2> tmp((size_t)n);
    CK(cudaMemcpy(tmp.data(), c->edges.p, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
    // canonical order so the export is deterministic whatever the atomics did
    std::sort(tmp.begin(), tmp.end(), [](const uint2& a, const uint2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    for (int64_t i = 0; i < n; ++i) {
        src_out[i] = (int32_t)tmp[(size_t)i].x;
        dst_out[i] = (int32_t)tmp[(size_t)i].y;
    }
    return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// one-shot host-buffer API
// ------------------------------------------------------------------------------------------------
static int run_with_retry(bf_ctx* c, int32_t max_dist, bf_stats* st) {
    bf_stats local;
    for (int attempt = 0; attempt < 6; ++attempt) {   // work list -> queue -> candidates can overflow one after the other
        TRY(bf_run(c, max_dist, 0, 1));
        int rc = bf_sync(c, &local);
        if (rc == BF_ERR_OVERFLOW) continue;   // bf_sync raised the capacities
        if (rc != BF_OK) return rc;
        if (st) *st = local;
        return BF_OK;
    }
    return fail(BF_ERR_OVERFLOW, "buffer overflow persisted after retries");
}

int bf_cluster_csr(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols, int32_t max_dist,
                   int32_t device, int32_t engine, int32_t* labels_out, bf_stats* stats_out) {
    bf_ctx* c = nullptr;
    TRY(bf_ctx_create(device, nullptr, &c));
    int rc = bf_ctx_set_option(c, "engine", engine);
    if (rc == BF_OK) rc = bf_upload_csr(c, indptr, indices, n_rows, n_cols, nullptr, 0);
    bf_stats st;
    if (rc == BF_OK) rc = run_with_retry(c, max_dist, &st);
    if (rc == BF_OK) rc = bf_download_labels(c, labels_out);
    if (rc == BF_OK && stats_out) {
        st.ms_d2h = c->ms_d2h;
        *stats_out = st;
    }
    std::string keep = g_err;
    bf_ctx_destroy(c);
    g_err = keep;
    return rc;
}

struct bf_edges_handle {
    std::vector<int32_t> src, dst;
};

int bf_neighbours_csr(const int64_t* indptr, const int32_t* indices, int64_t n_rows, int32_t n_cols,
                      const int32_t* query_rows, int64_t n_query, int32_t max_dist, int32_t device, int32_t engine,
                      void** edges_handle_out, int64_t* n_edges_out, bf_stats* stats_out) {
    if (!edges_handle_out || !n_edges_out) return fail(BF_ERR_INVALID, "null output");
    *edges_handle_out = nullptr;
    *n_edges_out = 0;
    bf_ctx* c = nullptr;
    TRY(bf_ctx_create(device, nullptr, &c));
    int rc = bf_ctx_set_option(c, "engine", engine);
    if (rc == BF_OK) rc = bf_ctx_set_option(c, "want_edges", 1);
    if (rc == BF_OK) rc = bf_upload_csr(c, indptr, indices, n_rows, n_cols, query_rows, query_rows ? n_query : 0);
    bf_stats st;
    if (rc == BF_OK) rc = run_with_retry(c, max_dist, &st);
    bf_edges_handle* h = nullptr;
    if (rc == BF_OK) {
        h = new (std::nothrow) bf_edges_handle();
        if (!h) rc = fail(BF_ERR_OOM, "host allocation failed");
    }
    if (rc == BF_OK) {
        try {
            h->src.resize((size_t)st.n_edges);
            h->dst.resize((size_t)st.n_edges);
        } catch (...) {
            rc = fail(BF_ERR_OOM, "host allocation failed");
        }
    }
    if (rc == BF_OK) rc = bf_download_edges(c, h->src.data(), h->dst.data());
    if (rc == BF_OK) {
        *edges_handle_out = h;
        *n_edges_out = st.n_edges;
        if (stats_out) *stats_out = st;
    } else {
        delete h;
    }
    std::string keep = g_err;
    bf_ctx_destroy(c);
    g_err = keep;
    return rc;
}

int bf_edges_copy(void* edges_handle, int32_t* src_out, int32_t* dst_out) {
    if (!edges_handle) return fail(BF_ERR_INVALID, "null handle");
    auto* h = static_cast<bf_edges_handle*>(edges_handle);
    if (h->src.empty()) return BF_OK;
    if (!src_out || !dst_out) return fail(BF_ERR_INVALID, "null output");
    memcpy(src_out, h->src.data(), h->src.size() * sizeof(int32_t));
    memcpy(dst_out, h->dst.data(), h->dst.size() * sizeof(int32_t));
    return BF_OK;
}

void bf_edges_free(void* edges_handle) { delete static_cast<bf_edges_handle*>(edges_handle); }

int bf_components(int64_t n_rows, const int32_t* src, const int32_t* dst, int64_t n_edges,
                  const int64_t* list_indptr, const int32_t* list_members, int64_t n_lists, int32_t device,
                  int32_t* labels_out, int64_t* n_components_out) {
    if (n_rows < 0 || n_edges < 0 || n_lists < 0) return fail(BF_ERR_INVALID, "negative size");
    if (n_edges > 0 && (!src || !dst)) return fail(BF_ERR_INVALID, "null edge arrays");
    for (int64_t e = 0; e < n_edges; ++e)
        if (src[e] < 0 || src[e] >= n_rows || dst[e] < 0 || dst[e] >= n_rows) return fail(BF_ERR_INVALID, "edge endpoint out of range");
    bf_ctx* c = nullptr;
    TRY(bf_ctx_create(device, nullptr, &c));
    auto body = [&]() -> int {
        c->n_rows = n_rows;
        c->uploaded = true;
    c->band_valid = false;
        TRY(c->parent.ensure((size_t)std::max<int64_t>(n_rows, 1) * sizeof(int)));
        TRY(c->labels.ensure((size_t)std::max<int64_t>(n_rows, 1) * sizeof(int32_t)));
        CK(cudaMemsetAsync(c->counters.p, 0, sizeof(DevCounters), c->stream));
        if (n_rows > 0) {
            k_uf_init<<<grid_for(n_rows, 256), 256, 0, c->stream>>>(c->parent.as<int>(), n_rows);
            CKLC(c);
        }
        if (n_edges > 0) {
            TRY(c->scratch.ensure((size_t)n_edges * sizeof(int32_t)));
            TRY(c->scratch2.ensure((size_t)n_edges * sizeof(int32_t)));
            CK(cudaMemcpyAsync(c->scratch.p, src, (size_t)n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->scratch2.p, dst, (size_t)n_edges * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            k_uf_edges<<<grid_for(n_edges, 256), 256, 0, c->stream>>>(c->parent.as<int>(), c->scratch.as<int32_t>(), c->scratch2.as<int32_t>(), n_edges);
            CKLC(c);
            CK(cudaStreamSynchronize(c->stream));
        }
        c->ran = true;
        TRY(finish_labels(c));
        TRY(bf_union_lists(c, list_indptr, list_members, n_lists));
        TRY(bf_download_labels(c, labels_out));
        if (n_components_out) {
            DevCounters h;
            CK(cudaMemcpy(&h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost));
            *n_components_out = h.n_comp;
        }
        return BF_OK;
    };
    int rc = body();
    std::string keep = g_err;
    bf_ctx_destroy(c);
    g_err = keep;
    return rc;
}

int bf_pinned_alloc(int64_t bytes, void** ptr_out) {
    if (!ptr_out || bytes < 0) return fail(BF_ERR_INVALID, "bad argument");
    *ptr_out = nullptr;
    CK(cudaHostAlloc(ptr_out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return BF_OK;
}

void bf_pinned_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
