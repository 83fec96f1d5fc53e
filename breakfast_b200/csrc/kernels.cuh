// kernels.cuh — sm_100a device code for the breakfast distance-and-clustering hot path.
//
// Pipeline (one pass = bf_run), default form (128/256-bit sketches, tensor-core level 1), nine launches:
//   K1a k_pack_sketch_rows16 (k_pack_sketch_rows on the plain CSR)  matrix in storage order -> staged sketches + sort keys
//                                                               (replaces csr_matrix + row sums, breakfast.py:214,287)
//   K2  k_radix_sort (one cooperative kernel, all passes)        rows sorted by (cardinality, two hash-half cardinalities)
//                                                               (replaces the np.isclose band, breakfast.py:250-254)
//   K1b k_permute_store                                          staged sketches -> tile-blocked bitsets, fold planes,
//                                                               int8 operands of level 1; union-find init
//   K2b k_schedule (its last block scans the item counts) + k_expand_items   three-key band-pruned tile-pair work list
//   K3a k_pairs_l1_imma2                                         level 1: int8 mma.sync on the 32-bit folds, two column rows
//                                                               per accumulator, dynamic item scheduler, survivors queued
//   K3b k_pairs_l2_unit                                          level 2: exact 32-bit test, full-width XOR/POPC, candidates
//                                                               (K3a+K3b replace sklearn _sparse_manhattan + _reduce_func,
//                                                                sklearn/metrics/_pairwise_fast.pyx:76-107, breakfast.py:226-228)
//   K3c k_verify_unite                                           exact |A xor B| on CSR rows of the candidates + union
//   K4  k_uf_labels (hook / path halving inside K3c)             replaces _to_graph + networkx connected_components
//                                                               (breakfast.py:93-113,325-326)
// Other forms: k_card_keys + k_pack_sketch / k_pack_full (wider sketches, FULL engine), k_pairs<K4> (single-kernel
// tiled XOR/POPC + threshold + compaction), k_pairs_l1_imma (one column row per accumulator), k_pairs_l1 + k_pairs_l2
// (level 1 on the integer pipes), k_hj_* (hash-join engine).
//
// Bitset layout in HBM ("tile-blocked"): rows are in sort-key order (K2), grouped in tiles of
// TILE=128 rows.  A row's bitset is cut into chunks of 4*K4 32-bit words (K4 = 16-byte groups per
// chunk, K4 in {1,2,4}).  One (tile, chunk) block is contiguous:
//       uint4 blk[K4][128]   blk[k4][row] = words 4*k4 .. 4*k4+3 of that row's chunk
// so (a) a whole operand block is ONE 1-D bulk-async (TMA) copy into shared memory, and (b) the
// 32 lanes of a warp that read 32 consecutive rows at the same k4 read 512 contiguous bytes:
// conflict-free 128-bit shared loads.
#pragma once
#include <climits>
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace bf {

constexpr int TILE = 128;              // rows per tile (both operands)
constexpr uint32_t KEY_CLAMP = 65535;  // cardinality sort key is clamped (1-Lipschitz, band test stays sound)
constexpr int PAIR_CONSUMER_WARPS = 16;
constexpr int PAIR_THREADS = (PAIR_CONSUMER_WARPS + 1) * 32;  // + 1 producer warp

struct DevCounters {
    unsigned long long n_cand;      // candidate cursor (may exceed capacity -> overflow)
    unsigned long long n_edges;     // verified edges
    unsigned long long band_ab;     // ordered in-band (a,b) count, A side vs B side (incl. self)
    unsigned long long band_aa;     // ordered in-band (a,a') count inside the A side (rectangle runs)
    unsigned long long l2_warp_items;  // (warp, tile pair) units that needed the full-width pass
    unsigned long long n_tilepairs;    // band tile pairs of the whole job (when work items group several column tiles)
    unsigned long long tilepairs_rank; // tile pairs this rank's level-1 kernel evaluated
    unsigned long long n_units;        // level-2 queue cursor (may exceed capacity -> overflow)
    unsigned int n_comp;
    unsigned int seg_max;              // fullest per-CTA segment of the level-2 pair queue (overflow check)
    unsigned int merge_fullest;        // longest compact label list of any rank in the exchange step (overflow check)
    unsigned int sched_done;           // blocks of k_schedule that have finished (the last one scans the item counts)
    unsigned int l1_ticket;            // next work item of this rank's level-1 kernel (dynamic scheduler of k_pairs_l1_imma2)
};

// ------------------------------------------------------------------------------------------
// K2: sort keys + stable LSD radix sort (8-bit digits over the bits that occur) — deterministic permutation
//
// key = c << 32 | s << 16 | t with c = |A| (the row's cardinality), s = |A n H1| and t = |A n H2|, H1 / H2 = the columns
// whose multiplicative hash has bit 31 / bit 30 set (s = t = 0 for the engines that do not stream the columns before
// the sort: H = {} is as valid as any H).  H1 and H2 cut the column space into four disjoint quadrants, and
// |A xor B| >= the sum over the quadrants of | |A n Q| - |B n Q| |.  Sorting by (c, s, t) therefore turns the partners of
// a row into a few contiguous runs, one per feasible (c' - c, s' - s), each bounded in t (see k_schedule) - about eight
// times fewer tile pairs than the plain cardinality band on SARS-CoV-2-shaped profiles (two keys: about four times).
// c is clamped to 16 bits (s, t <= c); k_schedule drops the s / t bounds wherever the clamped group is involved, so the
// band test stays sound for any input.
// ------------------------------------------------------------------------------------------
typedef unsigned long long sortkey_t;

__device__ __forceinline__ sortkey_t sort_key(int64_t card, uint32_t s, uint32_t t) {
    if (card >= (int64_t)KEY_CLAMP) return (sortkey_t)KEY_CLAMP << 32;   // clamped group: s and t carry no information
    return ((sortkey_t)card << 32) | ((sortkey_t)s << 16) | (sortkey_t)t;   // s, t <= card < 65535
}
__device__ __forceinline__ uint32_t key_card(sortkey_t k) { return (uint32_t)(k >> 32); }

// OR of all keys -> one atomic per warp; the radix passes run over the set bits of this word only
// (a warp first looks at the word: once the early blocks have set the bits that occur, nobody sends an atomic any more -
// tens of thousands of same-address atomics would otherwise queue up behind each other at about one per nanosecond)
__device__ __forceinline__ void or_reduce_key(sortkey_t k, sortkey_t* __restrict__ or_key) {
    const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)k), hi = __reduce_or_sync(0xffffffffu, (uint32_t)(k >> 32));
    if ((threadIdx.x & 31) == 0 && (lo | hi)) {
        const sortkey_t mine = ((sortkey_t)hi << 32) | lo;
        if (mine & ~__ldcg(or_key)) atomicOr(or_key, mine);
    }
}

__global__ void k_card_keys(const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                            int64_t n, sortkey_t* __restrict__ keys, int32_t* __restrict__ vals,
                            sortkey_t* __restrict__ or_key) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    sortkey_t k = 0;
    if (q < n) {
        int32_t r = rows ? rows[q] : (int32_t)q;
        k = sort_key(indptr[r + 1] - indptr[r], 0u, 0u);
        keys[q] = k;
        vals[q] = r;
    }
    or_reduce_key(k, or_key);
}

// keys of a row subset from the per-row keys of the whole matrix (written by k_pack_sketch_rows)
__global__ void k_gather_keys(const sortkey_t* __restrict__ row_keys, const int32_t* __restrict__ rows, int64_t n,
                              sortkey_t* __restrict__ keys, int32_t* __restrict__ vals, sortkey_t* __restrict__ or_key) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    sortkey_t k = 0;
    if (q < n) {
        const int32_t r = rows[q];
        k = row_keys[r];
        keys[q] = k;
        vals[q] = r;
    }
    or_reduce_key(k, or_key);
}

// The sort runs over the COMPRESSED key: the bits that are set in at least one key (or_key), eight per pass, so real
// data (cardinalities below 256, half cardinalities below 128: about 22 bits) needs three passes instead of six;
// R = ceil(popc(or_key) / 8) is decided on the device.  Three buffers make the result land in buffer 0 for any R:
//   R even: 0 -> 1 -> 0 ...      R odd >= 3: 0 -> 2 -> 1 -> 0 -> 1 -> 0 ...      R = 1: 0 -> 1, then a copy 1 -> 0.
struct SortBufs {
    sortkey_t* k[3];
    int32_t* v[3];
};
struct SortPlan {
    bool active, copy_only;
    int in, out;
    uint32_t pos[8];   // bit positions of this pass's digit bits (unused ones point at bit 63, which no key has)
};
__device__ __forceinline__ SortPlan sort_plan(sortkey_t or_key, int p) {
    SortPlan pl;
    const int R = (__popcll(or_key) + 7) >> 3;
    pl.active = p < R || (R == 1 && p == 1);
    pl.copy_only = R == 1 && p == 1;
    if ((R & 1) == 0) { pl.in = p & 1; pl.out = pl.in ^ 1; }
    else if (R == 1) { pl.in = p; pl.out = p ^ 1; }
    else if (p == 0) { pl.in = 0; pl.out = 2; }
    else if (p == 1) { pl.in = 2; pl.out = 1; }
    else { pl.in = (p + 1) & 1; pl.out = pl.in ^ 1; }
    sortkey_t m = or_key;
    for (int i = 0; i < 8 * p && m; ++i) m &= m - 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pl.pos[i] = m ? (uint32_t)(__ffsll((long long)m) - 1) : 63u;
        m &= m - 1;
    }
    return pl;
}
__device__ __forceinline__ uint32_t sort_digit(sortkey_t key, const uint32_t (&pos)[8]) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= (uint32_t)((key >> pos[i]) & 1ull) << i;
    return d;
}

// The whole sort is ONE cooperative kernel (k_radix_sort): a persistent grid of one 1024-thread block per SM, a block
// owns a contiguous slice of the rows, a warp a contiguous part of that slice.  Per pass: (1) every warp counts its
// digits into a private 256-bin histogram in shared memory (no atomics: __match_any_sync groups equal digits, the group
// leader adds the group size - reproducible ranks, stable pass); the block writes its 256 counts; grid-wide barrier;
// (2) every block turns the counts of all blocks into its own digit offsets (256 threads, one digit each), the warps
// re-read their keys and scatter; grid-wide barrier.  1 M rows: about 10 us per pass instead of three launches and
// 36 us (round 1).  counts[block * 256 + digit].
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_BATCH = 8;   // 32-row steps whose loads are in flight together

__global__ void __launch_bounds__(SORT_THREADS, 1) k_radix_sort(SortBufs b, int64_t n, uint32_t* __restrict__ counts,
                                                                const sortkey_t* __restrict__ or_key) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t whist[SORT_WARPS][256];
    __shared__ uint32_t part_total[4][256], part_before[4][256];
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = gridDim.x, blk = blockIdx.x;
    const sortkey_t ok = *or_key;
    const int R = (__popcll(ok) + 7) >> 3;
    const int n_slots = R == 1 ? 2 : R;
    // slices: 32-aligned, so that a warp's 32-row steps never straddle two warps' parts
    const int64_t per_block = ((n + G - 1) / G + 31) & ~(int64_t)31;
    const int64_t b0 = min(n, per_block * blk), b1 = min(n, b0 + per_block);
    const int64_t per_warp = (((b1 - b0) + SORT_WARPS - 1) / SORT_WARPS + 31) & ~(int64_t)31;
    const int64_t w0 = min(b1, b0 + per_warp * warp), w1 = min(b1, w0 + per_warp);
    for (int p = 0; p < n_slots; ++p) {
        const SortPlan pl = sort_plan(ok, p);
        const sortkey_t* __restrict__ keys = b.k[pl.in];
        const int32_t* __restrict__ vals = b.v[pl.in];
        sortkey_t* __restrict__ keys_out = b.k[pl.out];
        int32_t* __restrict__ vals_out = b.v[pl.out];
        if (pl.copy_only) {   // a single real pass left the result in buffer 1
            for (int64_t i = b0 + threadIdx.x; i < b1; i += SORT_THREADS) {
                keys_out[i] = keys[i];
                vals_out[i] = vals[i];
            }
            break;
        }
        for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&whist[0][0])[i] = 0;
        __syncthreads();
        // SORT_BATCH steps of 32 rows at a time: their loads are issued together (one memory round trip per batch
        // instead of one per step - a warp's part is a few hundred rows, so latency, not bandwidth, is what this costs)
        for (int64_t i0 = w0; i0 < w1; i0 += 32 * SORT_BATCH) {
            sortkey_t kb[SORT_BATCH];
#pragma unroll
            for (int u = 0; u < SORT_BATCH; ++u) {
                const int64_t i = i0 + 32 * u + lane;
                kb[u] = i < w1 ? __ldcg(&keys[i]) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < SORT_BATCH; ++u) {
                if (i0 + 32 * u >= w1) break;   // warp-uniform
                const bool valid = i0 + 32 * u + lane < w1;
                const uint32_t d = valid ? sort_digit(kb[u], pl.pos) : 256u + lane;  // invalid lanes match nobody
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                if (valid && (peers & ((1u << lane) - 1u)) == 0) whist[warp][d] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        if (threadIdx.x < 256) {   // digit = threadIdx.x: exclusive prefix over the warps, the block's count to global memory
            uint32_t run = 0;
#pragma unroll 8
            for (int w = 0; w < SORT_WARPS; ++w) {
                const uint32_t c = whist[w][threadIdx.x];
                whist[w][threadIdx.x] = run;
                run += c;
            }
            counts[(size_t)blk * 256 + threadIdx.x] = run;
        }
        __threadfence();
        grid.sync();
        {   // rows of every digit in earlier blocks and in all blocks: four threads per digit, loads coalesced over the digits
            const int q = threadIdx.x >> 8, d = threadIdx.x & 255;
            uint32_t before = 0, total = 0;
#pragma unroll 8
            for (int g = q; g < G; g += 4) {
                const uint32_t c = __ldcg(&counts[(size_t)g * 256 + d]);
                total += c;
                before += g < blk ? c : 0u;
            }
            part_total[q][d] = total;
            part_before[q][d] = before;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            const uint32_t total = part_total[0][threadIdx.x] + part_total[1][threadIdx.x] + part_total[2][threadIdx.x] + part_total[3][threadIdx.x];
            const uint32_t before = part_before[0][threadIdx.x] + part_before[1][threadIdx.x] + part_before[2][threadIdx.x] + part_before[3][threadIdx.x];
            // exclusive scan of the digit totals over the 256 digits (8 warps)
            uint32_t x = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) wsum[warp] = x;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            uint32_t base = x - total + before;
#pragma unroll
            for (int w = 0; w < 8; ++w) base += w < warp ? wsum[w] : 0u;
#pragma unroll 8
            for (int w = 0; w < SORT_WARPS; ++w) whist[w][threadIdx.x] += base;
        }
        __syncthreads();
        uint32_t* wpos = whist[warp];
        for (int64_t i0 = w0; i0 < w1; i0 += 32 * SORT_BATCH) {
            sortkey_t kb[SORT_BATCH];
            int32_t vb[SORT_BATCH];
#pragma unroll
            for (int u = 0; u < SORT_BATCH; ++u) {
                const int64_t i = i0 + 32 * u + lane;
                kb[u] = i < w1 ? __ldcg(&keys[i]) : 0ull;
                vb[u] = i < w1 ? __ldcg(&vals[i]) : 0;
            }
#pragma unroll
            for (int u = 0; u < SORT_BATCH; ++u) {
                if (i0 + 32 * u >= w1) break;   // warp-uniform
                const bool valid = i0 + 32 * u + lane < w1;
                const uint32_t d = valid ? sort_digit(kb[u], pl.pos) : 256u + lane;
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                const unsigned below = peers & ((1u << lane) - 1u);
                uint32_t pos = 0;
                if (valid) pos = wpos[d] + __popc(below);
                __syncwarp();
                if (valid && below == 0) wpos[d] += __popc(peers);
                __syncwarp();
                if (valid) {
                    keys_out[pos] = kb[u];
                    vals_out[pos] = vb[u];
                }
            }
        }
        __threadfence();
        grid.sync();
    }
}

// Exclusive scan (in place) of `data` by one block of THREADS threads inside another kernel (k_schedule's last block);
// total -> *total_out.  `data` was written by other blocks before a __threadfence, so it is read around L1.
template <int THREADS>
__device__ __forceinline__ void block_exclusive_scan_u64(unsigned long long* data, int64_t n, unsigned long long* total_out) {
    constexpr int PER = 8, WARPS = THREADS / 32;
    __shared__ unsigned long long warp_sums[WARPS];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += THREADS * PER) {
        const int64_t i0 = base + (int64_t)threadIdx.x * PER;
        unsigned long long v[PER], local = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            v[k] = (i0 + k < n) ? __ldcg(&data[i0 + k]) : 0ull;
            local += v[k];
        }
        unsigned long long x = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        unsigned long long before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const unsigned long long t = warp_sums[w];
            before += w < warp ? t : 0ull;
            all += t;
        }
        unsigned long long run = carry_s + before + (x - local);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (i0 + k < n) data[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += all;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// ------------------------------------------------------------------------------------------
// K2b: band-pruned tile schedule.  keys are sorted ascending, so a tile's min/max are its ends.
// Row tile I gets n_ranges schedule entries e = I * n_ranges + r, each one contiguous range [jlo[e], jend[e]) of B tiles.
// With D = c' - c, Ds = s' - s, Dt = t' - t for a partner (c', s', t') of a row (c, s, t) and u = the (unknown)
// difference on the quadrant H1 n H2, the four quadrant differences are u, Ds - u, Dt - u, D - Ds - Dt + u, so
//     |A xor B| >= min over u of |u| + |Ds - u| + |Dt - u| + |D - Ds - Dt + u|.
//   * n_keys = 3 and all rows of the tile share (c, s), c unclamped: one entry per feasible (D, Ds) - the host lists
//     them in ascending order with the feasible interval [dt_min, dt_max] of Dt (SchedRange table) - covering the keys
//     (c + D, s + Ds, tmin + dt_min) .. (c + D, s + Ds, tmax + dt_max); no s / t bound against the clamped group;
//   * else if all rows share one unclamped c (n_keys >= 2): entry per D = -d .. d with
//     s' - s in [-(d - D) / 2, (D + d) / 2] (integer divisions of non-negative numbers), any t;
//   * otherwise (tile across a cardinality boundary, clamped rows, n_keys = 1): entry 0 = the plain cardinality band
//     cmin - d .. cmax + d, the other entries empty.
// Ranges are clipped so that no B tile is listed twice for a row tile (and to J >= I in the triangular case).
// ------------------------------------------------------------------------------------------
// `group` column tiles form one work item (1 for the single-kernel path, L1_GROUP for the two-kernel
// path); an item never spans two entries; count[I] = number of work items of row tile I (all its entries),
// *n_tilepairs += number of tile pairs.
struct SchedRange {
    int8_t D, Ds, dt_min, dt_max;
};

// One thread per schedule entry (row tile, range): its two binary searches over the tile ends of B are independent of
// the other entries', so a block of SCHED_THREADS threads works on SCHED_THREADS / n_ranges row tiles at once and a
// tile's entries cost one search latency instead of n_ranges of them (a thread per row tile: 27 us at 1 M rows and
// max_dist 1; this form: see profiles/).  The clipping against the previous entry of the same tile (no B tile listed
// twice) is a short sequential pass of the entry-0 thread over shared memory.
constexpr int SCHED_THREADS = 256;
__global__ void __launch_bounds__(SCHED_THREADS)
k_schedule(const sortkey_t* __restrict__ keysA, int64_t nA, const sortkey_t* __restrict__ keysB,
                           int64_t nB, int max_dist, int triangular, int group, int n_keys, int n_ranges,
                           const SchedRange* __restrict__ table3, int n_table3, int32_t* __restrict__ jlo,
                           int32_t* __restrict__ jend, unsigned long long* __restrict__ count,
                           unsigned long long* __restrict__ n_tilepairs, unsigned int* __restrict__ done,
                           unsigned long long* __restrict__ n_work) {
    // count[I] = work items of row tile I; the block that finishes last turns count[0 .. tA] into the exclusive prefix
    // (count[tA] = *n_work = all items) - the scan used to be a launch of its own
    __shared__ int32_t raw_first[SCHED_THREADS], raw_end[SCHED_THREADS];
    __shared__ unsigned char raw_have[SCHED_THREADS];   // 0 = empty, 1 = range, 2 = the clamped group (listed once per tile)
    const int64_t tA = (nA + TILE - 1) / TILE, tB = (nB + TILE - 1) / TILE;
    const int tiles_per_block = max(1, SCHED_THREADS / n_ranges);   // n_ranges > SCHED_THREADS: a thread walks several entries
    const int tl = threadIdx.x / n_ranges, r0 = threadIdx.x % n_ranges;
    const int64_t I = (int64_t)blockIdx.x * tiles_per_block + tl;
    const bool active = tl < tiles_per_block && I < tA;
    sortkey_t kMin = 0, kMax = 0;
    if (active) {
        kMin = keysA[I * TILE];
        kMax = keysA[min(nA, (I + 1) * TILE) - 1];
    }
    const int64_t cMin = (int64_t)(kMin >> 32), cMax = (int64_t)(kMax >> 32);
    const int64_t sMin = (int64_t)((kMin >> 16) & 0xffffu), sMax = (int64_t)((kMax >> 16) & 0xffffu);
    const int64_t tMin = (int64_t)(kMin & 0xffffu), tMax = (int64_t)(kMax & 0xffffu);
    const bool single_c = n_keys >= 2 && cMin == cMax && cMax < (int64_t)KEY_CLAMP;
    const bool single_cs = n_keys >= 3 && single_c && sMin == sMax;
    const int n_gen = single_cs ? n_table3 : (single_c ? 2 * max_dist + 1 : 1);
    // with n_ranges <= SCHED_THREADS a thread has exactly one entry (r = r0); otherwise the one tile of the block is
    // walked by all threads in strides (its raw results do not fit shared memory: entry 0's thread redoes the chain
    // below from global memory, see the second pass)
    for (int r = r0; active && r < n_ranges; r += SCHED_THREADS) {
        int have = 0;
        sortkey_t lo_key = 0, hi_key = 0;
        if (r < n_gen) {
            if (single_cs) {
                const SchedRange e = table3[r];
                const int64_t c2 = cMin + e.D, s2 = sMin + e.Ds;
                if (c2 >= 0 && c2 <= (int64_t)KEY_CLAMP) {
                    if (c2 == (int64_t)KEY_CLAMP) {   // the clamped group carries no s / t: listed once, as a whole
                        have = 2;
                        lo_key = (sortkey_t)c2 << 32;
                        hi_key = ((sortkey_t)c2 << 32) | 0xffffffffull;
                    } else if (s2 >= 0) {
                        const int64_t t_lo = max((int64_t)0, tMin + e.dt_min), t_hi = min((int64_t)0xffff, tMax + e.dt_max);
                        if (t_lo <= t_hi) {
                            have = 1;
                            lo_key = ((sortkey_t)c2 << 32) | ((sortkey_t)s2 << 16) | (sortkey_t)t_lo;
                            hi_key = ((sortkey_t)c2 << 32) | ((sortkey_t)s2 << 16) | (sortkey_t)t_hi;
                        }
                    }
                }
            } else if (single_c) {
                const int64_t D = (int64_t)r - max_dist, c2 = cMin + D;
                if (c2 >= 0 && c2 <= (int64_t)KEY_CLAMP) {
                    int64_t s_lo = 0, s_hi = 0xffff;
                    if (c2 < (int64_t)KEY_CLAMP) {
                        s_lo = max((int64_t)0, sMin - (max_dist - D) / 2);
                        s_hi = min((int64_t)0xffff, sMax + (D + max_dist) / 2);
                    }
                    have = 1;
                    lo_key = ((sortkey_t)c2 << 32) | ((sortkey_t)s_lo << 16);
                    hi_key = ((sortkey_t)c2 << 32) | ((sortkey_t)s_hi << 16) | 0xffffull;
                }
            } else {
                have = 1;
                lo_key = (sortkey_t)max((int64_t)0, cMin - max_dist) << 32;
                hi_key = ((sortkey_t)min((int64_t)KEY_CLAMP, cMax + max_dist) << 32) | 0xffffffffull;
            }
        }
        int64_t first = 0, end = 0;
        if (have) {
            // first J with bMax[J] >= lo_key / first J with bMin[J] > hi_key: the two searches step together
            int64_t l1 = 0, r1 = tB, l2 = 0, r2 = tB;
            while (l1 < r1 || l2 < r2) {
                const int64_t m1 = (l1 + r1) >> 1, m2 = (l2 + r2) >> 1;
                const sortkey_t bMax = l1 < r1 ? keysB[min(nB, (m1 + 1) * TILE) - 1] : 0;
                const sortkey_t bMin = l2 < r2 ? keysB[m2 * TILE] : 0;
                if (l1 < r1) { if (bMax >= lo_key) r1 = m1; else l1 = m1 + 1; }
                if (l2 < r2) { if (bMin > hi_key) r2 = m2; else l2 = m2 + 1; }
            }
            first = l1;
            end = l2;
        }
        if (n_ranges <= SCHED_THREADS) {
            raw_first[threadIdx.x] = (int32_t)first;
            raw_end[threadIdx.x] = (int32_t)end;
            raw_have[threadIdx.x] = (unsigned char)have;
        } else {   // parked in the output arrays, clipped below
            jlo[I * n_ranges + r] = have ? (int32_t)first : -1 - have;
            jend[I * n_ranges + r] = (int32_t)end | (have == 2 ? (int32_t)0x40000000 : 0);
        }
    }
    __syncthreads();
    if (active && r0 == 0) {
    // second pass, entry 0's thread: clip every range against the end of the previous one
    int64_t prev_end = triangular ? I : 0;
    unsigned long long tp_sum = 0, items_sum = 0;
    bool clamp_done = false;
    for (int r = 0; r < n_ranges; ++r) {
        int have;
        int64_t first, end;
        const int64_t e = I * n_ranges + r;
        if (n_ranges <= SCHED_THREADS) {
            have = raw_have[threadIdx.x + r];
            first = raw_first[threadIdx.x + r];
            end = raw_end[threadIdx.x + r];
        } else {
            const int32_t f = jlo[e], en = jend[e];
            have = f < 0 ? 0 : ((en & 0x40000000) ? 2 : 1);
            first = f;
            end = en & 0x3fffffff;
        }
        if (have == 2) {
            if (clamp_done) have = 0;
            clamp_done = true;
        }
        if (have) {
            first = max(first, prev_end);
            end = max(end, first);
            prev_end = end;
        } else {
            first = end = prev_end;
        }
        jlo[e] = (int32_t)first;
        jend[e] = (int32_t)end;
        const unsigned long long tp = (unsigned long long)(end - first);
        items_sum += (tp + group - 1) / group;
        tp_sum += tp;
    }
    count[I] = items_sum;
    if (tp_sum) atomicAdd(n_tilepairs, tp_sum);
    }
    // last block done: exclusive scan of the counts
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) count[tA] = 0;
    __syncthreads();
    block_exclusive_scan_u64<SCHED_THREADS>(count, tA + 1, n_work);
}

// ordered in-band count: for every x in X, #{y in Y : ||x| - |y|| <= d}.  The metric's candidate pairs are defined
// on the cardinalities alone (upper key halves).  X is sorted, so only the first row of every run of equal
// cardinality searches (three binary searches) and counts for its whole run.
__global__ void __launch_bounds__(256) k_band_count(const sortkey_t* __restrict__ keysX, int64_t nX,
                                                    const sortkey_t* __restrict__ keysY, int64_t nY, int max_dist,
                                                    unsigned long long* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (i < nX) {
        const uint32_t k = key_card(keysX[i]);
        if (i == 0 || key_card(keysX[i - 1]) != k) {
            const sortkey_t lo_t = (sortkey_t)(k > (uint32_t)max_dist ? k - (uint32_t)max_dist : 0u) << 32;
            const sortkey_t hi_t = ((sortkey_t)min(k + (uint32_t)max_dist, KEY_CLAMP) << 32) | 0xffffffffull;
            const sortkey_t run_t = ((sortkey_t)k << 32) | 0xffffffffull;
            int64_t l = i, r = nX;   // end of this run in X
            while (l < r) { int64_t m = (l + r) >> 1; if (keysX[m] > run_t) r = m; else l = m + 1; }
            const int64_t run = l - i;
            l = 0; r = nY;
            while (l < r) { int64_t m = (l + r) >> 1; if (keysY[m] >= lo_t) r = m; else l = m + 1; }
            const int64_t a = l;
            r = nY;
            while (l < r) { int64_t m = (l + r) >> 1; if (keysY[m] > hi_t) r = m; else l = m + 1; }
            c = (unsigned long long)run * (unsigned long long)(l - a);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += ws[w];
        if (s) atomicAdd(out, s);
    }
}

// ------------------------------------------------------------------------------------------
// K1: bit-pack.  word `wd` of sorted row p lives at word_offset(...) (see layout above).
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t word_offset(int64_t tile, int n_chunks, int K4, int wd, int row) {
    const int wpc = 4 * K4;
    const int chunk = wd / wpc, r = wd - chunk * wpc;
    return ((((size_t)tile * n_chunks + chunk) * K4 + (r >> 2)) * TILE + row) * 4 + (r & 3);
}

// positions inside a 128-row fold tile: row-operand order (row = ty + 16 i  ->  ty*8 + i) and
// column-operand order (row = lane + 32 j  ->  lane*4 + j)
__host__ __device__ inline int fold_pos_a(int row) { return (row & 15) * 8 + (row >> 4); }
__host__ __device__ inline int fold_pos_b(int row) { return (row & 31) * 4 + (row >> 5); }
// level-2 unit order of the tensor-core level 1: lane l of a warp holds rows (l/4) + 8 s, s < 16, of the tile
__host__ __device__ inline int imma_unit_pos(int row) { return (row & 7) * 16 + (row >> 3); }

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-async-copy (TMA, SASS UBLKCP) wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16 bits -> 16 int8 values (+1 for a clear bit, -1 for a set bit)
__device__ __forceinline__ uint4 expand_pm1(uint32_t bits16) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t spread = (((bits16 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;  // bit i -> byte i
        w[q] = 0x01010101u ^ (spread * 0xfeu);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// the same for TWO rows at once: byte k = e_a[k] + 64 e_b[k] (e = +1 for a clear bit, -1 for a set bit), i.e. 65, 63,
// -63 or -65 - still an int8.  The int8 dot product of a +-1 row with this packed row is d_a + 64 d_b with
// d = 32 - 2 popc(f xor f'): one MMA accumulator carries the tests of two column rows (see k_pairs_l1_imma2).
__device__ __forceinline__ uint4 expand_pm1_pair(uint32_t a16, uint32_t b16) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t sa = (((a16 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;  // bit i -> byte i
        const uint32_t sb = (((b16 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
        w[q] = 0x41414141u ^ (sa * 0x7eu) ^ (sb * 0x80u);   // 0x41 = 65, ^0x7e -> 63, ^0x80 -> -63, both -> -65
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t fold_hash(uint32_t col, int log2m) {
    return (col * 2654435761u) >> (32 - log2m);  // multiplicative hash -> [0, m)
}

// SKETCH: one block per tile; the folded tile is assembled in shared memory (XOR-toggle per feature)
// and written out as one coalesced block.  HBM traffic: reads 4*nnz + 12*N bytes, writes N*m/8 bytes.
// Each warp owns 16 rows; lanes 0-15 fetch perm/indptr of all 16 rows at once (one dependent-load
// chain per warp, not per row), then the warp streams the rows' column lists.
__global__ void __launch_bounds__(256) k_pack_sketch(const int64_t* __restrict__ indptr,
                                                     const int32_t* __restrict__ indices,
                                                     const int32_t* __restrict__ perm, int64_t n, int log2m,
                                                     int n_chunks, int K4, uint32_t* __restrict__ bits) {
    extern __shared__ uint32_t tile_words[];  // n_chunks*K4*TILE*4 words
    const int words_per_tile = n_chunks * K4 * TILE * 4;
    for (int i = threadIdx.x; i < words_per_tile; i += 256) tile_words[i] = 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    int64_t my_b = 0, my_e = 0;
    if (lane < TILE / 8) {
        const int64_t p = tile * TILE + warp + 8 * lane;
        if (p < n) {
            const int32_t r = perm[p];
            my_b = indptr[r];
            my_e = indptr[r + 1];
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < TILE / 8; ++k) {
        const int64_t b = __shfl_sync(0xffffffffu, my_b, k), e = __shfl_sync(0xffffffffu, my_e, k);
        const int row = warp + 8 * k;
        for (int64_t q = b + lane; q < e; q += 32) {
            const uint32_t h = fold_hash((uint32_t)__ldg(&indices[q]), log2m);
            atomicXor(&tile_words[word_offset(0, n_chunks, K4, (int)(h >> 5), row)], 1u << (h & 31));
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(bits + (size_t)tile * words_per_tile);
    const uint4* src = reinterpret_cast<const uint4*>(tile_words);
    for (int i = threadIdx.x; i < words_per_tile / 4; i += 256) dst[i] = src[i];
}

// Stores of one packed row (128/256-bit sketches) at slot `row` of sorted tile `tile`: the sketch itself, its
// 32-bit fold in the two level-1/level-2 plane orders and, for the tensor-core level 1, the +-1 expanded fold.
// Called by all 128 threads of a tile's block (thread = row): with `packed8b` the int8 column operand holds TWO rows
// per 32 bytes (rows 2p and 2p + 1, expand_pm1_pair; the odd row's fold comes from the neighbouring lane).
template <int WORDS>
__device__ __forceinline__ void pack_store_row(const uint32_t (&w)[WORDS], int64_t tile, int row, uint32_t* __restrict__ bits,
                                               uint32_t* __restrict__ foldA, uint32_t* __restrict__ foldB,
                                               uint32_t* __restrict__ fold8a, uint4* __restrict__ fold8b, bool packed8b) {
    constexpr int K4 = WORDS / 4;
    uint4* dst = reinterpret_cast<uint4*>(bits) + (size_t)tile * (K4 * TILE);
#pragma unroll
    for (int g = 0; g < K4; ++g) dst[g * TILE + row] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
    // level-1 operand: the sketch folded once more to 32 bits, stored twice so that both the
    // row-operand (8 rows of a warp contiguous) and the column-operand (4 rows of a lane
    // contiguous) of k_pairs_l1 are single 128-bit shared loads
    uint32_t f = 0;
#pragma unroll
    for (int t = 0; t < WORDS; ++t) f ^= w[t];
    // (with the tensor-core level 1 the 32-bit planes feed level 2 only: rows in the order of a level-2
    // unit = the 16 rows of one accumulator-fragment lane contiguous, columns row-major)
    foldA[tile * TILE + (fold8b ? imma_unit_pos(row) : fold_pos_a(row))] = f;
    foldB[tile * TILE + (fold8b ? row : fold_pos_b(row))] = f;
    if (fold8b) {
        // tensor-core operands: bit k of the fold -> int8 (+1 if clear, -1 if set), so that the int8 dot
        // product of two rows is 32 - 2 popc(fa xor fb).
        //  column operand: plain row-major, 32 bytes per row (a lane's B fragment = 8 contiguous bytes)
        //  row operand: m16n8k32 A-fragment order - the 16 bytes {row r: k lo, row r+8: k lo, row r: k hi,
        //  row r+8: k hi} of lane (r%8)*4 + s4 of m-tile r/16 are contiguous, so a fragment is one LDS.128
        const uint4 e0 = expand_pm1(f & 0xffffu), e1 = expand_pm1(f >> 16);
        const uint32_t f_next = __shfl_down_sync(0xffffffffu, f, 1);   // fold of row + 1 (rows 2p, 2p + 1 share a warp)
        if (packed8b) {
            if ((row & 1) == 0) {
                uint4* tb = fold8b + (size_t)tile * TILE + (row >> 1) * 2;   // 64 packed rows x 32 bytes per tile
                tb[0] = expand_pm1_pair(f & 0xffffu, f_next & 0xffffu);
                tb[1] = expand_pm1_pair(f >> 16, f_next >> 16);
            }
        } else {
            uint4* tb = fold8b + (size_t)tile * (TILE * 2);
            tb[row * 2 + 0] = e0;
            tb[row * 2 + 1] = e1;
        }
        // (staging the tile in shared memory and writing it out with coalesced 16-byte stores was measured: no gain, 33.8 us
        // either way at 10^6 rows - the scattered words of a tile merge in L2)
        uint32_t* ta = fold8a + (size_t)tile * (TILE * 8);
        const int m = row >> 4, h = (row >> 3) & 1, fr = row & 7;
        const uint32_t words[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};   // bytes 4q .. 4q+3
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            uint32_t* frag = ta + ((m * 8 + fr) * 4 + s4) * 4;
            frag[h] = words[2 * s4];           // k lo of lane (fr * 4 + s4)'s slice (bytes 8 s4 .. 8 s4 + 3)
            frag[2 + h] = words[2 * s4 + 1];   // k hi (bytes 8 s4 + 4 .. 8 s4 + 7)
        }
    }
}

// SKETCH, m = 32*WORDS <= 256 bits, step 1 of 2 (before the sort): one pass over the CSR in storage order gives
// every row its sketch (row-major staging, 4*WORDS bytes per row) AND its sort key c << 16 | s (see K2).
// A block takes 128 consecutive rows - one contiguous stretch of `indices`.  Rows are short (about 90 columns), so a
// warp folds FOUR rows at a time, eight lanes each: every lane XORs the bits of its columns straight into the row's
// sketch in shared memory (one ATOMS.XOR per column - no per-word selects, no warp reduction; the columns of H are
// counted in a register), six loads per lane in flight.  Reads 4*nnz + 8*N bytes, writes N*(m/8 + 8) bytes.
// (Measured alternatives at 1 M rows of ~89 columns: a warp per row with per-lane word selects + redux.sync and
// lane-0 stores: 0.34 ms, in sorted (gather) or storage order alike; a thread per row: 0.69 ms, its 32-sector
// loads are L1-wavefront bound; the shared-memory form: 0.13 ms.)
template <int WORDS>
__global__ void __launch_bounds__(256) k_pack_sketch_rows(const int64_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices, int64_t n, int log2m,
                                                          uint32_t* __restrict__ sk_rows, sortkey_t* __restrict__ keys,
                                                          int32_t* __restrict__ vals, sortkey_t* __restrict__ or_key,
                                                          int64_t block0) {
    // block0: first 128-row block of this launch (a rank of a multi-GPU job takes a contiguous share of the blocks)
    __shared__ uint32_t sk[TILE][WORDS];
    __shared__ uint32_t sub[TILE];
    __shared__ int64_t row_b[TILE], row_e[TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (block0 + (int64_t)blockIdx.x) * TILE;
    if (threadIdx.x < TILE) {
        const int64_t r = row0 + threadIdx.x;
        int64_t b = 0, e = 0;
        if (r < n) {
            b = indptr[r];
            e = indptr[r + 1];
        }
        row_b[threadIdx.x] = b;
        row_e[threadIdx.x] = e;
        sub[threadIdx.x] = 0u;
#pragma unroll
        for (int t = 0; t < WORDS; ++t) sk[threadIdx.x][t] = 0u;
    }
    __syncthreads();
    constexpr int PACK_LOADS = 6;   // independent loads per lane before the first use (48 columns per row and round:
                                    // two rounds cover 96, the kernel is bound by issued instructions, ncu 76 %)
    const int sg = lane >> 3, l8 = lane & 7;
#pragma unroll 1
    for (int k = 0; k < TILE / 32; ++k) {
        const int row = warp * (TILE / 8) + 4 * k + sg;
        const int32_t* src = indices + row_b[row];
        const int len = (int)(row_e[row] - row_b[row]);   // a row has fewer than 2^31 columns
        uint32_t* dst = sk[row];
        uint32_t in_h = 0;   // this lane's columns in H1 (top hash bit set; low half) and in H2 (next bit; high half)
        for (int q = l8; q < len; q += 8 * PACK_LOADS) {
            int32_t col[PACK_LOADS];
#pragma unroll
            for (int t = 0; t < PACK_LOADS; ++t) col[t] = q + 8 * t < len ? __ldg(src + q + 8 * t) : -1;
#pragma unroll
            for (int t = 0; t < PACK_LOADS; ++t) {
                if (col[t] >= 0) {   // column ids are non-negative
                    const uint32_t h = fold_hash((uint32_t)col[t], log2m);
                    atomicXor(&dst[h >> 5], 1u << (h & 31));
                    in_h += (h >> (log2m - 1)) + ((h << (18 - log2m)) & 0x10000u);
                }
            }
        }
        if (in_h) atomicAdd(&sub[row], in_h);
    }
    __syncthreads();
    sortkey_t key = 0;
    if (threadIdx.x < TILE && row0 + threadIdx.x < n) {
        const int64_t r = row0 + threadIdx.x;
        uint32_t* out = sk_rows + (size_t)r * WORDS;
#pragma unroll
        for (int g = 0; g < WORDS / 4; ++g)
            reinterpret_cast<uint4*>(out)[g] = make_uint4(sk[threadIdx.x][4 * g], sk[threadIdx.x][4 * g + 1],
                                                          sk[threadIdx.x][4 * g + 2], sk[threadIdx.x][4 * g + 3]);
        key = sort_key(row_e[threadIdx.x] - row_b[threadIdx.x], sub[threadIdx.x] & 0xffffu, sub[threadIdx.x] >> 16);
        keys[r] = key;
        vals[r] = (int32_t)r;
    }
    or_reduce_key(key, or_key);
}

// The same step on the COMPACT RESIDENT FORM of the matrix ("CSR16", see k_csr16_encode / include/breakfast_b200.h:
// 32-bit row offsets, the low 16 bits of every column, per row the number of columns below 65536) - half the bytes of
// the plain CSR.  A block takes 128 consecutive rows.  Their columns are one contiguous stretch of `lo`: thread 0
// fetches it with 1-D bulk-async copies (TMA, completion on an mbarrier) in chunks of PACK16_CHUNK entries - one chunk
// for rows of up to 128 columns on average.  LPR lanes then walk every row out of shared memory (lane j takes columns
// j, j + LPR, ...).  A lane toggles the bits of its columns in a sketch of its own in shared memory (word w of thread t
// at priv[w][t]: bank = t mod 32, conflict-free, no other thread touches it), so the word select costs one address
// computation instead of one compare + XOR per word; the two quadrant counters ride in registers (2|H1| + |H2| and
// |H1|).  Row lengths differ a lot (40 ... 145 columns in SARS-CoV-2 profiles), and a warp takes as long as its longest
// row: the rows of a block are therefore dealt to the threads in order of length (a 64-bin counting sort in shared
// memory), which keeps the lanes of a warp within a few columns of each other.
// Reads 2 nnz + 6 N bytes, writes N (m/8 + 12) bytes; the sketch does not depend on which thread folded which row.
constexpr int PACK16_CHUNK = 16384;   // entries (32 KB of shared memory)
constexpr int PACK16_LANES = 2;       // lanes per row
template <int WORDS, int LPR, bool ATOM>
__global__ void __launch_bounds__(TILE * LPR) k_pack_sketch_rows16(const uint32_t* __restrict__ indptr32,
                                                                   const uint16_t* __restrict__ split,
                                                                   const uint16_t* __restrict__ lo, int64_t n,
                                                                   uint32_t* __restrict__ sk_rows, sortkey_t* __restrict__ keys,
                                                                   int32_t* __restrict__ vals, sortkey_t* __restrict__ or_key,
                                                                   int64_t block0) {
    static_assert(WORDS == 4 || WORDS == 8, "128- or 256-bit sketches");
    static_assert(LPR == 1 || LPR == 2, "one or two lanes per row");
    constexpr int LOG2M = WORDS == 4 ? 7 : 8;
    constexpr int NT = TILE * LPR;
    __shared__ __align__(16) uint16_t cols[PACK16_CHUNK + 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t priv[WORDS][NT];
    __shared__ uint32_t row_b[TILE + 1];
    __shared__ uint16_t row_s[TILE];
    __shared__ int bins[64];
    __shared__ unsigned char order[TILE];
    const int64_t row0 = (block0 + (int64_t)blockIdx.x) * TILE;
    const int tid = threadIdx.x;
    if (tid < 64) bins[tid] = 0;
#pragma unroll
    for (int t = 0; t < WORDS; ++t) priv[t][tid] = 0u;
    for (int i = tid; i <= TILE; i += NT) {
        row_b[i] = __ldg(&indptr32[min(row0 + i, n)]);
        if (i < TILE) row_s[i] = (split && row0 + i < n) ? __ldg(&split[row0 + i]) : (uint16_t)0xffffu;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // rows in order of length: counting sort on length / 4 (clamped), ties in any order
    int my_bin = 0;
    if (tid < TILE) {
        my_bin = (int)min((row_b[tid + 1] - row_b[tid]) >> 2, 63u);
        atomicAdd(&bins[my_bin], 1);
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the 64 bins by one warp
        const int a = bins[2 * tid], b2 = bins[2 * tid + 1];
        int incl = a + b2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        bins[2 * tid] = incl - a - b2;
        bins[2 * tid + 1] = incl - b2;
    }
    __syncthreads();
    if (tid < TILE) order[atomicAdd(&bins[my_bin], 1)] = (unsigned char)tid;
    __syncthreads();
    const int rt = order[tid / LPR], sub = tid % LPR;   // row of the block, lane of the row
    const int64_t r = row0 + rt;
    const uint32_t b = row_b[rt], e = row_b[rt + 1];
    const uint32_t sp = row_s[rt] == 0xffffu && !split ? 0xffffffffu : (uint32_t)row_s[rt];
    const uint32_t blk_b = row_b[0] & ~7u, blk_e = row_b[TILE];   // chunk starts stay 16-byte aligned
    uint32_t quad = 0, in_h1 = 0, phase = 0;
    const uint32_t hi_at = b + min(sp, e - b);   // first position of the row whose column is >= 65536
    uint32_t* mine = &priv[0][tid];
    for (uint32_t cb = blk_b; cb < blk_e; cb += PACK16_CHUNK) {
        const uint32_t ce = min(cb + (uint32_t)PACK16_CHUNK, blk_e);
        if (tid == 0) {
            const uint32_t bytes = ((ce - cb) * 2u + 15u) & ~15u;   // the array is allocated with 16 bytes of slack
            mbar_arrive_expect_tx(&bar, bytes);
            bulk_g2s(cols, lo + cb, bytes, &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        const uint32_t kb = max(b, cb), ke = min(e, ce);
        // lane `sub` takes the row's columns at positions = sub (mod LPR)
        uint32_t k = kb + ((sub + LPR - ((kb - b) % LPR)) % LPR);
        // fold_hash of column lo + 65536 [k >= hi_at]: h = h32 >> (32 - LOG2M); the multiplication distributes
        auto fold = [&](uint32_t lo16, uint32_t kk) {
            const uint32_t h32 = lo16 * 2654435761u + (kk >= hi_at ? 2654435761u << 16 : 0u);
            const uint32_t bit = 1u << ((h32 >> (32 - LOG2M)) & 31u);
            uint32_t* dst = mine + (h32 >> (37 - LOG2M)) * NT;   // word h >> 5 of this thread's sketch
            if (ATOM) atomicXor(dst, bit);
            else *dst ^= bit;
            quad += h32 >> 30;    // 2 |row n H1| + |row n H2| ...
            in_h1 += h32 >> 31;   // ... and |row n H1| (H1 / H2: hash bit 31 / 30 set)
        };
        for (; k + 3 * LPR < ke; k += 4 * LPR) {   // four columns per round: the loads first, all independent
            uint32_t c4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) c4[u] = cols[k + u * LPR - cb];
#pragma unroll
            for (int u = 0; u < 4; ++u) fold(c4[u], k + u * LPR);
        }
        for (; k < ke; k += LPR) fold(cols[k - cb], k);
        __syncthreads();   // everyone is done with this chunk before the next copy lands
    }
    uint32_t w[WORDS];
#pragma unroll
    for (int t = 0; t < WORDS; ++t) w[t] = priv[t][tid];
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) {
#pragma unroll
        for (int t = 0; t < WORDS; ++t) w[t] ^= __shfl_xor_sync(0xffffffffu, w[t], o);
        quad += __shfl_xor_sync(0xffffffffu, quad, o);
        in_h1 += __shfl_xor_sync(0xffffffffu, in_h1, o);
    }
    sortkey_t key = 0;
    if (sub == 0 && r < n) {
        uint4* out = reinterpret_cast<uint4*>(sk_rows + (size_t)r * WORDS);
#pragma unroll
        for (int g = 0; g < WORDS / 4; ++g) out[g] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
        key = sort_key((int64_t)(e - b), in_h1, quad - 2u * in_h1);
        keys[r] = key;
        vals[r] = (int32_t)r;
    }
    or_reduce_key(key, or_key);
}

// step 2 of 2 (after the sort): thread p of block `tile` fetches the staged sketch of the row at sorted position
// tile * 128 + p and writes all derived layouts of that slot (pack_store_row); pad slots are zero-filled.
template <int WORDS>
__global__ void __launch_bounds__(TILE) k_permute_store(const uint32_t* __restrict__ sk_rows,
                                                        const int32_t* __restrict__ perm, int64_t n,
                                                        uint32_t* __restrict__ bits, uint32_t* __restrict__ foldA,
                                                        uint32_t* __restrict__ foldB, uint32_t* __restrict__ fold8a,
                                                        uint4* __restrict__ fold8b, int packed8b, int* __restrict__ parent_init) {
    const int64_t p = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (parent_init && p < n) parent_init[p] = (int)p;   // the union-find starts here too (one launch less per pass)
    uint32_t w[WORDS];
#pragma unroll
    for (int t = 0; t < WORDS; ++t) w[t] = 0u;   // rows past n stay all-zero (never emitted: index check in the pair kernels)
    if (p < n) {
        const uint4* src = reinterpret_cast<const uint4*>(sk_rows + (size_t)perm[p] * WORDS);
#pragma unroll
        for (int g = 0; g < WORDS / 4; ++g) {
            const uint4 v = __ldg(&src[g]);
            w[4 * g] = v.x; w[4 * g + 1] = v.y; w[4 * g + 2] = v.z; w[4 * g + 3] = v.w;
        }
    }
    pack_store_row<WORDS>(w, blockIdx.x, threadIdx.x, bits, foldA, foldB, fold8a, fold8b, packed8b != 0);
}

// FULL: bit matrix pre-zeroed by the host (cudaMemsetAsync); one warp per row sets its bits.
__global__ void __launch_bounds__(256) k_pack_full(const int64_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices,
                                                   const int32_t* __restrict__ perm, int64_t n, int n_chunks,
                                                   int K4, uint32_t* __restrict__ bits) {
    const int lane = threadIdx.x & 31;
    int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= n) return;
    int32_t r = perm[p];
    int64_t b = indptr[r], e = indptr[r + 1];
    const int64_t tile = p / TILE;
    const int row = (int)(p - tile * TILE);
    for (int64_t k = b + lane; k < e; k += 32) {
        uint32_t c = (uint32_t)indices[k];
        atomicOr(&bits[word_offset(tile, n_chunks, K4, (int)(c >> 5), row)], 1u << (c & 31));
    }
}


// ------------------------------------------------------------------------------------------
// K2c: explicit work list.  items[w] = (I, J) for every band tile pair, so that the pair kernel's
// producer needs one load per item instead of a binary search.
// ------------------------------------------------------------------------------------------
// With group > 1 an item covers column tiles J0 .. J0+cnt-1 (cnt <= group <= 8) and is stored as
// (I, J0 | (cnt-1) << 29); jcount[I] (tile pairs of row tile I) is recomputed from the next prefix.
// work item w of row tile I (wprefix[I] <= w < wprefix[I + 1]) -> (first column tile, number of column tiles):
// walk the tile's n_ranges schedule entries
__device__ __forceinline__ int2 item_of_tile(int64_t I, unsigned long long off, int n_ranges, int group,
                                             const int32_t* __restrict__ jlo, const int32_t* __restrict__ jend) {
    int j0 = 0, cnt = 1;
    for (int r = 0; r < n_ranges; ++r) {
        const int lo = __ldg(&jlo[I * n_ranges + r]), en = __ldg(&jend[I * n_ranges + r]);
        const unsigned long long n_r = (unsigned long long)(en - lo + group - 1) / group;
        if (off < n_r) {
            j0 = lo + (int)off * group;
            cnt = min(group, en - j0);
            break;
        }
        off -= n_r;
    }
    return make_int2(j0, cnt);
}

__global__ void k_expand_items(const unsigned long long* __restrict__ wprefix, const int32_t* __restrict__ jlo,
                               int64_t tilesA, int n_ranges, const unsigned long long* __restrict__ n_work,
                               unsigned long long cap, int2* __restrict__ items, int group,
                               const int32_t* __restrict__ jend) {
    const unsigned long long W = min(*n_work, cap);
    for (unsigned long long w = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; w < W;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = tilesA - 1;  // largest I with wprefix[I] <= w (never a tile without items)
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (__ldg(&wprefix[mid]) <= w) lo = mid; else hi = mid - 1;
        }
        const int2 jc = item_of_tile(lo, w - __ldg(&wprefix[lo]), n_ranges, group, jlo, jend);
        items[w] = make_int2((int)lo, group == 1 ? jc.x : (jc.x | ((jc.y - 1) << 29)));
    }
}

// ------------------------------------------------------------------------------------------
// K3: tiled XOR/POPC pair kernel.
//   persistent grid; work item = one band tile pair (I, J); item w of the global list belongs to
//   rank (w % world); block b takes this rank's items b, b+grid, ...
//   warp 16 = producer: 32 lanes fetch 32 work items at once, lane 0 issues two bulk-async (TMA)
//   copies per chunk into a STAGES-deep shared-memory ring, completion on an mbarrier (expect_tx).
//   warps 0-15 = consumers: thread (warp, lane) owns the 8x4 pairs (warp+16i, lane+32j); per
//   16-byte k-group: 4 LDS.128 of B kept in registers, 8 broadcast LDS.128 of A, 128
//   LOP3(xor)+POPC+IADD.  Threshold in the epilogue; the (rare) hits go to a global candidate list.
//   Algorithmic work per evaluated pair: bits_per_row/32 POPC32 (+ as many XOR).
//
//   TWO_LEVEL (single-chunk sketches): every row's sketch is first XOR-folded once more to 32 bits
//   (fold of a fold is still a lower bound of |A xor B|), so level 1 costs ONE popc per pair and
//   keeps only a running minimum; a warp whose 1024 pairs all exceed the threshold is done.  Only
//   warps with a level-1 survivor run the full-width pass above (counted in l2_warp_items).
// ------------------------------------------------------------------------------------------
template <int K4, int STAGES>
struct PairSmem {
    static constexpr int kOperandBytes = K4 * TILE * 16;
    static constexpr int kStageBytes = 2 * kOperandBytes;
    static constexpr int kRingBytes = STAGES * kStageBytes;
    static constexpr int kTotalBytes = kRingBytes + STAGES * (8 + 8 + 8);
};

template <int K4, int STAGES, bool TWO_LEVEL>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
k_pairs(const uint4* __restrict__ bitsA, const uint4* __restrict__ bitsB, int n_chunks, int64_t nA, int64_t nB,
        const int2* __restrict__ items, unsigned long long items_cap,
        const unsigned long long* __restrict__ wprefix, const int32_t* __restrict__ jlo,
        const int32_t* __restrict__ jend, int64_t tilesA, int n_ranges,
        const unsigned long long* __restrict__ n_work, int threshold, int triangular, int rank, int world,
        uint2* __restrict__ cand, unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    using L = PairSmem<K4, STAGES>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kRingBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    int2* meta = reinterpret_cast<int2*>(empty_bar + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned long long W = *n_work;
    const unsigned long long first = (unsigned long long)rank + (unsigned long long)world * blockIdx.x;
    const unsigned long long stride = (unsigned long long)world * gridDim.x;

    if (warp == PAIR_CONSUMER_WARPS) {
        // ------------------------------- producer -------------------------------
        uint32_t it = 0;
        for (unsigned long long k0 = 0;; k0 += 32) {
            if (first + k0 * stride >= W) break;
            const unsigned long long w = first + (k0 + lane) * stride;
            int2 mine = make_int2(0, 0);
            if (w < W) {
                if (w < items_cap) {
                    mine = __ldg(&items[w]);
                } else {  // beyond the expanded table (huge bands): map w -> (I, J) by binary search
                    int64_t lo = 0, hi = tilesA - 1;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi + 1) >> 1;
                        if (__ldg(&wprefix[mid]) <= w) lo = mid; else hi = mid - 1;
                    }
                    mine = make_int2((int)lo, item_of_tile(lo, w - __ldg(&wprefix[lo]), n_ranges, 1, jlo, jend).x);
                }
            }
            for (int l = 0; l < 32; ++l) {
                if (first + (k0 + l) * stride >= W) break;
                const int Il = __shfl_sync(0xffffffffu, mine.x, l), Jl = __shfl_sync(0xffffffffu, mine.y, l);
                if (lane == 0) {
                    const uint4* gA = bitsA + (size_t)Il * n_chunks * (K4 * TILE);
                    const uint4* gB = bitsB + (size_t)Jl * n_chunks * (K4 * TILE);
                    for (int c = 0; c < n_chunks; ++c, ++it) {
                        const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(&empty_bar[stage], ph ^ 1u);
                        meta[stage] = make_int2(Il, Jl);
                        unsigned char* sa = smem + stage * L::kStageBytes;
                        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                        bulk_g2s(sa, gA + (size_t)c * (K4 * TILE), L::kOperandBytes, &full_bar[stage]);
                        bulk_g2s(sa + L::kOperandBytes, gB + (size_t)c * (K4 * TILE), L::kOperandBytes,
                                 &full_bar[stage]);
                    }
                }
                __syncwarp();
            }
        }
        return;
    }

    // --------------------------------- consumers ---------------------------------
    // 16 warps; thread (ty = warp, tx = lane) owns the 8 x 4 pairs (ty + 16 i, tx + 32 j).  A rows are
    // warp-uniform (broadcast LDS.128), B rows are 32 consecutive 16-byte groups (conflict-free).
    const int tx = lane, ty = warp;
    uint32_t it = 0;
    unsigned int l2_count = 0;
    for (unsigned long long w = first; w < W; w += stride) {
        int acc[8][4];
        int2 ij = make_int2(0, 0);
        bool full_pass = true;
        if constexpr (TWO_LEVEL) {
            // ---- level 1: one popc per pair on the 32-bit fold, running minimum only (n_chunks == 1)
            const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1u;
            mbar_wait(&full_bar[stage], ph);
            const uint4* sA = reinterpret_cast<const uint4*>(smem + stage * L::kStageBytes);
            const uint4* sB = sA + K4 * TILE;
            uint32_t fb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t f = 0;
#pragma unroll
                for (int k4 = 0; k4 < K4; ++k4) {
                    const uint4 v = sB[k4 * TILE + tx + 32 * j];
                    f ^= v.x ^ v.y ^ v.z ^ v.w;
                }
                fb[j] = f;
            }
            int mn = 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t fa = 0;
#pragma unroll
                for (int k4 = 0; k4 < K4; ++k4) {
                    const uint4 v = sA[k4 * TILE + ty + 16 * i];
                    fa ^= v.x ^ v.y ^ v.z ^ v.w;
                }
                mn = min(mn, min(min(__popc(fa ^ fb[0]), __popc(fa ^ fb[1])),
                                 min(__popc(fa ^ fb[2]), __popc(fa ^ fb[3]))));
            }
            full_pass = __any_sync(0xffffffffu, mn <= threshold);
            if (!full_pass) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                ++it;
                continue;
            }
            ++l2_count;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0;
        for (int c = 0; c < n_chunks; ++c, ++it) {
            const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1u;
            if (!TWO_LEVEL) mbar_wait(&full_bar[stage], ph);
            const uint4* sA = reinterpret_cast<const uint4*>(smem + stage * L::kStageBytes);
            const uint4* sB = sA + K4 * TILE;
#pragma unroll
            for (int k4 = 0; k4 < K4; ++k4) {
                uint4 b[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sB[k4 * TILE + tx + 32 * j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 a = sA[k4 * TILE + ty + 16 * i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j] += __popc(a.x ^ b[j].x) + __popc(a.y ^ b[j].y) + __popc(a.z ^ b[j].z) +
                                     __popc(a.w ^ b[j].w);
                    }
                }
            }
            if (c == n_chunks - 1) ij = meta[stage];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
        }
        // epilogue: threshold.  3-input min tree first (~0.5 op per pair); only a thread that owns a
        // hit builds a hit mask (straight-line, no branches) and walks its set bits, so the emit code
        // exists once and the instruction footprint of the hot loop stays small (an unrolled
        // per-accumulator branch here cost 38 % stall_no_inst in the first ncu capture).
        int mn = acc[0][0];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) mn = min(mn, acc[i][j]);
        if (mn <= threshold) {
            uint32_t mask = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mask |= (acc[i][j] <= threshold ? 1u : 0u) << (i * 4 + j);
            const int64_t gi0 = (int64_t)ij.x * TILE + ty, gj0 = (int64_t)ij.y * TILE + tx;
            while (mask) {
                const int bit = __ffs((int)mask) - 1;
                mask &= mask - 1;
                const int64_t gi = gi0 + 16 * (bit >> 2), gj = gj0 + 32 * (bit & 3);
                if (gi < nA && gj < nB && (!triangular || gi < gj)) {
                    unsigned long long pos = atomicAdd(&counters->n_cand, 1ull);
                    if (pos < cand_cap) cand[pos] = make_uint2((uint32_t)gi, (uint32_t)gj);
                }
            }
        }
    }
    if (TWO_LEVEL && lane == 0 && l2_count) atomicAdd(&counters->l2_warp_items, (unsigned long long)l2_count);
}

// ------------------------------------------------------------------------------------------
// K3 (two-kernel form, single-chunk sketches): level 1 on the 32-bit fold planes, level 2 on a queue.
//
// k_pairs_l1: persistent, warp 16 = producer (one bulk copy of the row tile's folds + one of up to
//   L1_GROUP consecutive column tiles' folds per work item), warps 0-15 = consumers.  Thread
//   (warp, lane) tests the 8 x 4 pairs (warp+16i, lane+32j) of every column tile of the item with ONE
//   32-bit XOR per pair and either POPC (columns j = 0,1) or, for max_dist 1 and 2, the POPC-free
//   "clear the lowest set bit max_dist times, then compare with 0" on the ALU pipe (columns j = 2,3),
//   so that the XU (POPC), ALU and FMA pipes share the work.  Only running minima are kept.  A thread
//   whose 32 pairs of a tile pair hold a survivor appends the unit (I, J, warp, lane) to the queue.
// k_pairs_l2: one thread per queued unit, operands straight from L2 (the sketches are L2-resident),
//   full-width XOR/POPC on the unit's 32 pairs, threshold, candidate emission.  Perfectly balanced,
//   no ring, no inter-warp coupling.
// ------------------------------------------------------------------------------------------
constexpr int L1_GROUP = 4;
constexpr int L1_STAGES = 8;
constexpr int L1_STAGE_BYTES = (1 + L1_GROUP) * TILE * 4;
constexpr int L1_SMEM_BYTES = L1_STAGES * L1_STAGE_BYTES + L1_STAGES * (8 + 8 + 8);

template <int T>  // T = max_dist if it is 1 or 2 (hybrid XU/ALU test), 0 = POPC only with a runtime threshold
__global__ void __launch_bounds__(PAIR_THREADS, 1)
k_pairs_l1(const uint32_t* __restrict__ foldA, const uint32_t* __restrict__ foldB, const int2* __restrict__ items,
           unsigned long long items_cap, const unsigned long long* __restrict__ n_work, int threshold, int rank,
           int world, uint32_t one, int2* __restrict__ queue, unsigned long long queue_cap,
           DevCounters* __restrict__ counters) {
    // `one` is 1, passed as an argument so that y * one - 1 stays an IMAD: the decrement of the POPC-free
    // test then runs on the FMA pipe and the ALU pipe (XOR, AND, min) stops being the bottleneck
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L1_STAGES * L1_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + L1_STAGES;
    int2* meta = reinterpret_cast<int2*>(empty_bar + L1_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < L1_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long W = min(*n_work, items_cap);
    const unsigned long long first = (unsigned long long)rank + (unsigned long long)world * blockIdx.x;
    const unsigned long long stride = (unsigned long long)world * gridDim.x;

    if (warp == PAIR_CONSUMER_WARPS) {
        uint32_t it = 0;
        unsigned long long tp = 0;
        for (unsigned long long k0 = 0;; k0 += 32) {
            if (first + k0 * stride >= W) break;
            const unsigned long long w = first + (k0 + lane) * stride;
            int2 mine = make_int2(0, 0);
            if (w < W) mine = __ldg(&items[w]);
            for (int l = 0; l < 32; ++l) {
                if (first + (k0 + l) * stride >= W) break;
                const int Il = __shfl_sync(0xffffffffu, mine.x, l), Jp = __shfl_sync(0xffffffffu, mine.y, l);
                if (lane == 0) {
                    const int J0 = Jp & 0x1fffffff, cnt = ((unsigned)Jp >> 29) + 1;
                    const uint32_t stage = it % L1_STAGES, ph = (it / L1_STAGES) & 1u;
                    mbar_wait(&empty_bar[stage], ph ^ 1u);
                    meta[stage] = make_int2(Il, Jp);
                    unsigned char* sa = smem + stage * L1_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(1 + cnt) * TILE * 4);
                    bulk_g2s(sa, foldA + (size_t)Il * TILE, TILE * 4, &full_bar[stage]);
                    bulk_g2s(sa + TILE * 4, foldB + (size_t)J0 * TILE, (uint32_t)cnt * TILE * 4, &full_bar[stage]);
                    ++it;
                    tp += cnt;
                }
                __syncwarp();
            }
        }
        if (lane == 0 && tp) atomicAdd(&counters->tilepairs_rank, tp);
        return;
    }

    uint32_t it = 0;
    for (unsigned long long w = first; w < W; w += stride, ++it) {
        const uint32_t stage = it % L1_STAGES, ph = (it / L1_STAGES) & 1u;
        mbar_wait(&full_bar[stage], ph);
        const uint4* sA = reinterpret_cast<const uint4*>(smem + stage * L1_STAGE_BYTES);
        const uint4* sB = sA + TILE / 4;
        const int2 ij = meta[stage];
        const int cnt = ((unsigned)ij.y >> 29) + 1;
        const uint4 a0 = sA[warp * 2], a1 = sA[warp * 2 + 1];  // folds of rows warp + 16 i, i = 0..7
        const uint32_t fa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        for (int jt = 0; jt < cnt; ++jt) {
            const uint4 b = sB[jt * (TILE / 4) + lane];        // folds of rows lane + 32 j, j = 0..3
            bool mine;
            if constexpr (T == 0) {
                int mn = 64;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    mn = min(mn, min(__popc(fa[i] ^ b.x), __popc(fa[i] ^ b.y)));
                    mn = min(mn, min(__popc(fa[i] ^ b.z), __popc(fa[i] ^ b.w)));
                }
                mine = mn <= threshold;
            } else {
                int mn = 64;
                uint32_t mz = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    mn = min(mn, min(__popc(fa[i] ^ b.x), __popc(fa[i] ^ b.y)));
                    uint32_t y2 = fa[i] ^ b.z, y3 = fa[i] ^ b.w;
#pragma unroll
                    for (int t = 0; t < T; ++t) {  // popc(y) <= T  <=>  y with its T lowest set bits cleared == 0
                        y2 &= y2 * one - 1u;
                        y3 &= y3 * one - 1u;
                    }
                    mz = min(mz, min(y2, y3));
                }
                mine = (mn <= T) || (mz == 0u);
            }
            if (mine) {  // rare: this thread's 32 pairs of tile pair (I, J0 + jt) go to level 2
                const unsigned long long pos = atomicAdd(&counters->n_units, 1ull);
                if (pos < queue_cap) queue[pos] = make_int2(ij.x | (warp << 24), ((ij.y & 0x1fffffff) + jt) | (lane << 24));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }
}

// ------------------------------------------------------------------------------------------
// K3a'': level 1 on the tensor cores through mma.sync (SASS IMMA.16832.S8, register accumulators).
// The pack kernel expands the 32-bit folds to +-1 int8 rows (32 B per row), so an m16n8k32 MMA gives
// acc = 32 - 2 popc(fa xor fb) for 128 pairs, and popc <= max_dist  <=>  acc >= 32 - 2 max_dist.
//   warp 16 = producer (same ring as k_pairs_l1, operands 4 KB per 128-row tile).
//   warps 0-15: warp w takes the 8 column rows 8w .. 8w+7 of every column tile against all 128 rows of the
//   row tile: the 8 A fragments (16 rows each) are loaded once per item and stay in registers, per column
//   tile one 64-bit shared load gives the B fragment, 8 IMMAs produce 32 accumulators per thread, a 3-input
//   max tree decides.  LDS.64 addresses row*32 + 8 (lane%4): conflict-free per half-warp.  The K order
//   inside a row is permuted identically for both operands (bytes 8w..8w+3 <-> MMA k = 4w.., bytes
//   8w+4..8w+7 <-> k = 16+4w..), which a dot product does not see.
// Measured IMMA rate on the box: 1949 MAC/clk/SM = 60.9 pairs/clk/SM (tools/experiments/imma_rate.cu)
// against ~27 pairs/clk/SM of the integer-pipe level 1; a tcgen05/TMEM variant was bit-exact but slower
// (accumulator round-trip latency with K = 32, see tools/experiments/l1_tcgen05_kernel.cuh.txt).
// ------------------------------------------------------------------------------------------
constexpr int IMMA_STAGES = 5;
constexpr int IMMA_GROUP = 8;    // column tiles per work item of the tensor-core level 1 (runs are six tiles long on average)
constexpr int L2_CBUF = 1024;   // candidates a CTA of k_pairs_l2_unit collects per round before one cursor update
constexpr int L2_SUB = 64;      // CTAs of k_pairs_l2_unit per queue segment when level 1 ran one CTA per SM (one unit per thread
                                // for segments up to 16 K units); halved when it ran two (half as many units per segment)
constexpr int IMMA_TILE_BYTES = TILE * 32;
constexpr int IMMA_STAGE_BYTES = (1 + IMMA_GROUP) * IMMA_TILE_BYTES;
constexpr int IMMA_SMEM_BYTES = IMMA_STAGES * IMMA_STAGE_BYTES + IMMA_STAGES * (8 + 8 + 8);

__device__ __forceinline__ void imma_16832(int (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    // not volatile: a pure function of its operands, so the scheduler may overlap it with the max trees
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "r"(0));
}

__global__ void __launch_bounds__(PAIR_THREADS, 1)
k_pairs_l1_imma(const uint32_t* __restrict__ fold8A, const uint4* __restrict__ fold8B, int64_t nA, int64_t nB,
                const int2* __restrict__ items, unsigned long long items_cap,
                const unsigned long long* __restrict__ n_work, int max_dist, int triangular, int rank, int world,
                int2* __restrict__ queue, unsigned long long queue_cap, unsigned* __restrict__ seg_counts,
                DevCounters* __restrict__ counters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + IMMA_STAGES * IMMA_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + IMMA_STAGES;
    int2* meta = reinterpret_cast<int2*>(empty_bar + IMMA_STAGES);
    __shared__ unsigned seg_cursor;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the level-2 queue is cut into one segment per CTA
    const unsigned seg_cap = (unsigned)min(queue_cap / gridDim.x, 0xffffffffull);
    int2* seg = queue + (size_t)blockIdx.x * seg_cap;
    if (threadIdx.x == 0) {
        seg_cursor = 0;
        for (int s = 0; s < IMMA_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long W = min(*n_work, items_cap);
    const unsigned long long first = (unsigned long long)rank + (unsigned long long)world * blockIdx.x;
    const unsigned long long stride = (unsigned long long)world * gridDim.x;

    if (warp == PAIR_CONSUMER_WARPS) {
        uint32_t it = 0;
        unsigned long long tp = 0;
        for (unsigned long long k0 = 0;; k0 += 32) {
            if (first + k0 * stride >= W) break;
            const unsigned long long w = first + (k0 + lane) * stride;
            int2 mine = make_int2(0, 0);
            if (w < W) mine = __ldg(&items[w]);
            for (int l = 0; l < 32; ++l) {
                if (first + (k0 + l) * stride >= W) break;
                const int Il = __shfl_sync(0xffffffffu, mine.x, l), Jp = __shfl_sync(0xffffffffu, mine.y, l);
                if (lane == 0) {
                    const int J0 = Jp & 0x1fffffff, cnt = ((unsigned)Jp >> 29) + 1;
                    const uint32_t stage = it % IMMA_STAGES, ph = (it / IMMA_STAGES) & 1u;
                    mbar_wait(&empty_bar[stage], ph ^ 1u);
                    meta[stage] = make_int2(Il, Jp);
                    unsigned char* sa = smem + stage * IMMA_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(1 + cnt) * IMMA_TILE_BYTES);
                    bulk_g2s(sa, fold8A + (size_t)Il * (TILE * 8), IMMA_TILE_BYTES, &full_bar[stage]);
                    bulk_g2s(sa + IMMA_TILE_BYTES, fold8B + (size_t)J0 * (TILE * 2), (uint32_t)cnt * IMMA_TILE_BYTES, &full_bar[stage]);
                    ++it;
                    tp += cnt;
                }
                __syncwarp();
            }
        }
        if (lane == 0 && tp) atomicAdd(&counters->tilepairs_rank, tp);
        return;
    }

    const int thr = 32 - 2 * max_dist;
    const int frow = lane >> 2, fk = (lane & 3) * 8;   // B fragment: row inside the warp's 8 rows, byte offset of this lane's k slice
    uint32_t stage = 0, ph = 0;
    const uint32_t W32 = (uint32_t)W, stride32 = (uint32_t)stride;   // the work list holds at most 2^24 items
    for (uint32_t w = (uint32_t)first; w < W32; w += stride32) {
        mbar_wait(&full_bar[stage], ph);
        const unsigned char* sA = smem + stage * IMMA_STAGE_BYTES;
        const unsigned char* sB = sA + IMMA_TILE_BYTES + (8 * warp + frow) * 32 + fk;
        const int2 ij = meta[stage];
        const int cnt = ((unsigned)ij.y >> 29) + 1;
        // A fragments of the 8 m-tiles, stored in fragment order by the pack kernel: one LDS.128 each
        uint4 a[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) a[m] = reinterpret_cast<const uint4*>(sA)[m * 32 + lane];
        // The column tiles of an item are unrolled (missing ones skipped, warp-uniformly): per tile one LDS.64 and the
        // MMAs in two groups of four m-tiles, so that the max tree of one group overlaps the tensor-pipe time of the next.
        int mxj[IMMA_GROUP];
#pragma unroll
        for (int jt = 0; jt < IMMA_GROUP; ++jt) {
            mxj[jt] = INT_MIN;   // below any threshold
            if (jt >= cnt) continue;   // a partly filled item does not pay for its missing tiles
            const uint2 b = *reinterpret_cast<const uint2*>(sB + jt * IMMA_TILE_BYTES);
            int mh[2];
#pragma unroll
            for (int hgrp = 0; hgrp < 2; ++hgrp) {
                int c[4][4];
#pragma unroll
                for (int m = 0; m < 4; ++m) imma_16832(c[m], a[4 * hgrp + m], b.x, b.y);
                int m0 = max(max(c[0][0], c[0][1]), max(c[0][2], c[0][3]));
                int m1 = max(max(c[1][0], c[1][1]), max(c[1][2], c[1][3]));
                m0 = max(m0, max(c[2][0], c[2][1]));
                m1 = max(m1, max(c[2][2], c[2][3]));
                m0 = max(m0, max(c[3][0], c[3][1]));
                m1 = max(m1, max(c[3][2], c[3][3]));
                mh[hgrp] = max(m0, m1);
            }
            mxj[jt] = max(mh[0], mh[1]);
        }
        // rare: one of this lane's 16 x 2 pairs of tile pair (I, J0 + jt) may be within max_dist -> the unit goes to this
        // CTA's segment of the level-2 queue (shared-memory cursor: no contended global atomic); one test for the item first
        int any = mxj[0];
#pragma unroll
        for (int jt = 1; jt < IMMA_GROUP; ++jt) any = max(any, mxj[jt]);
        if (any >= thr) {
#pragma unroll
            for (int jt = 0; jt < IMMA_GROUP; ++jt) {
                if (mxj[jt] >= thr) {
                    const unsigned pos = atomicAdd(&seg_cursor, 1u);
                    if (pos < seg_cap) seg[pos] = make_int2(ij.x | (warp << 24), ((ij.y & 0x1fffffff) + jt) | (lane << 24));
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == IMMA_STAGES) {
            stage = 0;
            ph ^= 1u;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PAIR_CONSUMER_WARPS * 32) : "memory");   // consumer warps only
    if (threadIdx.x == 0) {
        const unsigned n = seg_cursor;
        seg_counts[blockIdx.x] = n;
        if (n) {
            atomicAdd(&counters->n_units, (unsigned long long)n);
            atomicMax(&counters->seg_max, n);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K3a-2: the same level 1 with TWO column rows per accumulator (default).  The column operand holds rows 2p and 2p + 1
// of a tile as one int8 row e(2p) + 64 e(2p+1) (expand_pm1_pair: 65, 63, -63, -65), so one m16n8k32 MMA against the
// +-1 row operand gives acc = d1 + 64 d2 with d = 32 - 2 popc(fa xor fb) of the two pairs - half the MMAs, half the
// accumulators and half the operand bytes per evaluated pair.  Both tests come out of ONE multiply and a packed max:
// with x = acc + 63 and w = x * 66560 = (x << 16) + (x << 10) (one IMAD: acc * 66560 + 63 * 66560),
//     high half of w = x + floor(x / 64), strictly increasing in acc, and d2 >= thr  <=>  acc >= 64 thr - 32 (|d1| <= 32),
//     low half of w  = ((d1 + 63) mod 64) << 10, as a signed 16-bit number 1024 (d1 - 1) for every even d1 in [-30, 32]
//                      (d1 = -32, the complement fold, aliases d1 = 32: a false positive that level 2 rejects),
// so a 3-input signed 16x2 max tree (VIMNMX3.S16x2, half an ALU op per accumulator; the multiply is an IMAD on the FMA
// pipe) over a thread's accumulators decides both column rows at once: survivor  <=>  high >= 65 thr + 31 or
// low >= 1024 (thr - 1), thr = 32 - 2 max_dist (max_dist >= 32: every pair survives, as it must).
//   warp w: m-tiles 4 (w / 8) .. + 3 (64 rows) against the packed rows 8 (w % 8) .. + 7 (16 column rows) of every
//   column tile: 4 A fragments in registers per item, per column tile one LDS.64 + 4 IMMAs + 16 IMADs + 8 packed maxes.
//   A level-2 unit = the 8 x 4 pairs of one thread: rows (lane / 4) + 64 (w / 8) + 8 s, s < 8, columns
//   16 (w % 8) + 4 (lane % 4) + {0, 1, 2, 3}.
// ------------------------------------------------------------------------------------------
constexpr int IMMA2_TILEB_BYTES = (TILE / 2) * 32;
constexpr int IMMA2_STAGE_BYTES = IMMA_TILE_BYTES + IMMA_GROUP * IMMA2_TILEB_BYTES;
template <int MINB> struct Imma2Cfg {
    static constexpr int kStages = MINB == 1 ? 8 : 5;   // 160 KB for one CTA per SM, 2 x 100 KB for two
    static constexpr int kSmemBytes = kStages * IMMA2_STAGE_BYTES + kStages * (8 + 8 + 8);
};

template <int MINB>
__global__ void __launch_bounds__(PAIR_THREADS, MINB)
k_pairs_l1_imma2(const uint32_t* __restrict__ fold8A, const uint4* __restrict__ fold8P, int64_t nA, int64_t nB,
                 const int2* __restrict__ items, unsigned long long items_cap,
                 const unsigned long long* __restrict__ n_work, int max_dist, int triangular, int rank, int world,
                 int pack_mul, int pack_add, int2* __restrict__ queue, unsigned long long queue_cap, unsigned* __restrict__ seg_counts,
                 DevCounters* __restrict__ counters) {
    // pack_mul = 66560 and pack_add = 63 * 66560 are passed as arguments so that acc * pack_mul + pack_add stays ONE
    // IMAD (FMA pipe) instead of shifts and adds on the ALU pipe
    constexpr int STAGES = Imma2Cfg<MINB>::kStages;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * IMMA2_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    int2* meta = reinterpret_cast<int2*>(empty_bar + STAGES);
    __shared__ unsigned seg_cursor;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned seg_cap = (unsigned)min(queue_cap / gridDim.x, 0xffffffffull);   // one queue segment per CTA
    int2* seg = queue + (size_t)blockIdx.x * seg_cap;
    if (threadIdx.x == 0) {
        seg_cursor = 0;
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long W = min(*n_work, items_cap);
    // Work items are handed out dynamically: this rank's items are w = rank + world * t, t = 0, 1, ...; the producer
    // draws the next t from one device-wide ticket counter (items cost between one and eight column tiles plus their
    // survivors, and a static deal left the CTAs 10 % apart at 10^6 profiles).  The ticket and the item of the NEXT
    // stage are requested before the wait for a free slot, so their latency hides behind it.  A stage with row tile -1
    // tells the consumers that the list is exhausted.
    const unsigned long long n_mine = W > (unsigned long long)rank ? (W - rank + world - 1) / world : 0ull;

    if (warp == PAIR_CONSUMER_WARPS) {
        if (lane == 0) {
            uint32_t it = 0;
            unsigned long long tp = 0;
            unsigned long long t = atomicAdd(&counters->l1_ticket, 1u);
            int2 cur = t < n_mine ? __ldg(&items[(unsigned long long)rank + (unsigned long long)world * t]) : make_int2(-1, 0);
            while (true) {
                const bool last = cur.x < 0;
                unsigned long long t_next = 0;
                int2 nxt = make_int2(-1, 0);
                if (!last) {
                    t_next = atomicAdd(&counters->l1_ticket, 1u);
                    if (t_next < n_mine) nxt = __ldg(&items[(unsigned long long)rank + (unsigned long long)world * t_next]);
                }
                const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&empty_bar[stage], ph ^ 1u);
                meta[stage] = cur;
                if (last) {
                    mbar_arrive(&full_bar[stage]);
                    break;
                }
                const int J0 = cur.y & 0x1fffffff, cnt = ((unsigned)cur.y >> 29) + 1;
                unsigned char* sa = smem + stage * IMMA2_STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)IMMA_TILE_BYTES + (uint32_t)cnt * IMMA2_TILEB_BYTES);
                bulk_g2s(sa, fold8A + (size_t)cur.x * (TILE * 8), IMMA_TILE_BYTES, &full_bar[stage]);
                bulk_g2s(sa + IMMA_TILE_BYTES, fold8P + (size_t)J0 * TILE, (uint32_t)cnt * IMMA2_TILEB_BYTES, &full_bar[stage]);
                ++it;
                tp += cnt;
                cur = nxt;
            }
            if (tp) atomicAdd(&counters->tilepairs_rank, tp);
        }
        return;
    }

    const int thr = 32 - 2 * max_dist;
    // packed thresholds (high half: the odd column row, low half: the even one); max_dist >= 32 passes everything
    const int hi_thr = max_dist >= 32 ? -32768 : 65 * thr + 31, lo_thr = max_dist >= 32 ? -32768 : 1024 * (thr - 1);
    const bool all_pass = max_dist >= 32;
    const uint32_t thr_m1 = ((uint32_t)(hi_thr - 1) << 16) | ((uint32_t)(lo_thr - 1) & 0xffffu);   // (unused when all_pass)
    constexpr uint32_t kNone = 0x80008000u;   // below any threshold in both halves
    const int mhalf = warp >> 3, ngrp = warp & 7;
    const int frow = lane >> 2, fk = (lane & 3) * 8;   // B fragment: packed row inside the warp's 8, byte offset of this lane's k slice
    uint32_t stage = 0, ph = 0;
    while (true) {
        mbar_wait(&full_bar[stage], ph);
        const unsigned char* sA = smem + stage * IMMA2_STAGE_BYTES;
        const unsigned char* sB = sA + IMMA_TILE_BYTES + (8 * ngrp + frow) * 32 + fk;
        const int2 ij = meta[stage];
        if (ij.x < 0) break;   // the producer's end mark
        const int cnt = ((unsigned)ij.y >> 29) + 1;
        uint4 a[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] = reinterpret_cast<const uint4*>(sA)[(4 * mhalf + m) * 32 + lane];
        uint32_t mxj[IMMA_GROUP];
#pragma unroll
        for (int jt = 0; jt < IMMA_GROUP; ++jt) {
            mxj[jt] = kNone;
            if (jt >= cnt) continue;   // a partly filled item does not pay for its missing tiles
            const uint2 b = *reinterpret_cast<const uint2*>(sB + jt * IMMA2_TILEB_BYTES);
            int c[4][4];
#pragma unroll
            for (int m = 0; m < 4; ++m) imma_16832(c[m], a[m], b.x, b.y);
            uint32_t p[16];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) p[4 * m + q] = (uint32_t)(c[m][q] * pack_mul + pack_add);   // one IMAD
            uint32_t m0 = __vimax3_s16x2(p[0], p[1], p[2]), m1 = __vimax3_s16x2(p[3], p[4], p[5]);
            m0 = __vimax3_s16x2(m0, p[6], p[7]);
            m1 = __vimax3_s16x2(m1, p[8], p[9]);
            m0 = __vimax3_s16x2(m0, p[10], p[11]);
            m1 = __vimax3_s16x2(m1, p[12], p[13]);
            m0 = __vimax3_s16x2(m0, p[14], p[15]);
            mxj[jt] = __vmaxs2(m0, m1);
        }
        // One of this thread's 8 x 4 pairs of tile pair (I, J0 + jt) may be within max_dist -> the unit goes to this CTA's
        // segment of the level-2 queue.  Rare per thread (2.6 % of the units at 10^6 profiles), but two warp-items in three
        // have one: the test for the whole item is one more max tree + a packed compare ("some half above its threshold"
        // <=> the packed max with thresholds - 1 is not thresholds - 1), and the warp appends its units with ONE update of
        // the shared-memory cursor (prefix sum over the lanes' counts).
        static_assert(IMMA_GROUP == 8, "the item test below is written for eight column tiles");
        uint32_t any = __vimax3_s16x2(mxj[0], mxj[1], mxj[2]);
        any = __vimax3_s16x2(any, mxj[3], mxj[4]);
        any = __vimax3_s16x2(any, mxj[5], mxj[6]);
        any = __vmaxs2(any, mxj[7]);
        const bool mine = all_pass || __vmaxs2(any, thr_m1) != thr_m1;
        if (__any_sync(0xffffffffu, mine)) {
            uint32_t mask = 0;
            if (mine) {
#pragma unroll
                for (int jt = 0; jt < IMMA_GROUP; ++jt)
                    mask |= (jt < cnt && (all_pass || __vmaxs2(mxj[jt], thr_m1) != thr_m1) ? 1u : 0u) << jt;
            }
            const int n_mine = __popc(mask);
            int incl = n_mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            unsigned base = 0;
            if (lane == 31) base = atomicAdd(&seg_cursor, (unsigned)incl);
            unsigned pos = __shfl_sync(0xffffffffu, base, 31) + (unsigned)(incl - n_mine);
            while (mask) {
                const int jt = __ffs((int)mask) - 1;
                mask &= mask - 1;
                if (pos < seg_cap) seg[pos] = make_int2(ij.x | (warp << 24), ((ij.y & 0x1fffffff) + jt) | (lane << 24));
                ++pos;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) {
            stage = 0;
            ph ^= 1u;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PAIR_CONSUMER_WARPS * 32) : "memory");   // consumer warps only
    if (threadIdx.x == 0) {
        const unsigned n = seg_cursor;
        seg_counts[blockIdx.x] = n;
        if (n) {
            atomicAdd(&counters->n_units, (unsigned long long)n);
            atomicMax(&counters->seg_max, n);
        }
    }
}

// level 2 of the tensor-core level 1: one THREAD per queued unit = the 16 x 2 pairs (rows (tx/4) + 8 s of tile I,
// columns 8 ty + 2 (tx%4) + {0, 1} of tile J) held by one accumulator-fragment lane.  The unit's 16 + 2 32-bit
// folds are three contiguous loads (unit-order planes written by the pack kernel); only the pairs that pass the
// exact 32-bit test (about one in sixty) fetch the two full sketches.
// PACKED (units of k_pairs_l1_imma2): rows (tx/4) + 64 (ty/8) + 8 s, s < 8, columns 16 (ty%8) + 4 (tx%4) + {0..3} - the
// unit's 8 + 4 folds are again three contiguous loads of the same planes.
template <int K4, bool PACKED>
__global__ void __launch_bounds__(256)
k_pairs_l2_unit(const uint4* __restrict__ bitsA, const uint4* __restrict__ bitsB, const uint4* __restrict__ foldA,
                const uint2* __restrict__ foldB, int64_t nA, int64_t nB, const int2* __restrict__ queue,
                unsigned long long queue_cap, const unsigned* __restrict__ seg_counts, int n_segs, int n_sub, int threshold,
                int triangular, uint2* __restrict__ cand, unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    // blockIdx.x = segment * n_sub + slice: the CTAs of one segment stride over its units
    const unsigned seg_cap = (unsigned)min(queue_cap / (unsigned)n_segs, 0xffffffffull);
    const int sgm = blockIdx.x / n_sub, sub = blockIdx.x % n_sub;
    const unsigned n = min(__ldg(&seg_counts[sgm]), seg_cap);
    const int2* seg = queue + (size_t)sgm * seg_cap;
    // candidates are collected per CTA and appended with ONE global cursor update per round: a contended
    // same-address atomic per candidate (about one per nanosecond device-wide) was most of this kernel's time
    __shared__ uint2 cbuf[L2_CBUF];
    __shared__ unsigned cbuf_n, checks_n;
    __shared__ unsigned long long cbuf_base;
    if (threadIdx.x == 0) cbuf_n = checks_n = 0;
    __syncthreads();
    unsigned full_checks = 0;
    for (unsigned u0 = sub * blockDim.x; u0 < n; u0 += n_sub * blockDim.x) {   // uniform over the CTA
      const unsigned u = u0 + threadIdx.x;
      if (u < n) {
        const int2 unit = __ldg(&seg[u]);
        const int I = unit.x & 0x00ffffff, ty = (unit.x >> 24) & 15, J = unit.y & 0x00ffffff, tx = (unit.y >> 24) & 31;
        // NC columns per unit: bit NC s + j of `hits` = pair (row r0 + 8 s, column c0 + j) passes the 32-bit test
        constexpr int NC = PACKED ? 4 : 2;
        const int r0 = PACKED ? (tx >> 2) + 64 * (ty >> 3) : tx >> 2;
        const int c0 = PACKED ? 16 * (ty & 7) + 4 * (tx & 3) : 8 * ty + 2 * (tx & 3);
        uint32_t fb[NC];
        if constexpr (PACKED) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(foldB) + (((size_t)J * TILE + c0) >> 2));
            fb[0] = v.x; fb[1] = v.y; fb[2] = v.z; fb[3] = v.w;
        } else {
            const uint2 v = __ldg(&foldB[((size_t)J * TILE + c0) >> 1]);
            fb[0] = v.x; fb[1] = v.y;
        }
        // unit-order plane: rows (r % 8) + 8 s of a tile are contiguous over s (imma_unit_pos)
        const uint4* fa4 = foldA + ((size_t)I * TILE + imma_unit_pos(r0)) / 4;
        const int64_t gi0 = (int64_t)I * TILE + r0, gj0 = (int64_t)J * TILE + c0;
        // pass 1, no loads beyond the folds
        uint32_t hits = 0;
#pragma unroll
        for (int q = 0; q < 32 / NC / 4; ++q) {
            const uint4 fa = __ldg(&fa4[q]);
            const uint32_t f[4] = {fa.x, fa.y, fa.z, fa.w};
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int j = 0; j < NC; ++j) hits |= (__popc(f[t] ^ fb[j]) <= threshold ? 1u : 0u) << (NC * (4 * q + t) + j);
        }
        // pass 2: the lanes of a warp walk their own survivors together (one or two rounds, whatever the slots)
        while (hits) {
            const int bit = __ffs(hits) - 1;
            hits &= hits - 1;
            const int s = bit / NC, j = bit % NC;
            const int64_t gi = gi0 + 8 * s, gj = gj0 + j;
            if (gi >= nA || gj >= nB || (triangular && gi >= gj)) continue;
            ++full_checks;
            const uint4* a = bitsA + (size_t)I * (K4 * TILE) + (r0 + 8 * s);
            const uint4* b = bitsB + (size_t)J * (K4 * TILE) + (c0 + j);
            int d = 0;
#pragma unroll
            for (int k4 = 0; k4 < K4; ++k4) {
                const uint4 x = __ldg(&a[k4 * TILE]), y = __ldg(&b[k4 * TILE]);
                d += __popc(x.x ^ y.x) + __popc(x.y ^ y.y) + __popc(x.z ^ y.z) + __popc(x.w ^ y.w);
            }
            if (d <= threshold) {
                const unsigned slot = atomicAdd(&cbuf_n, 1u);
                if (slot < L2_CBUF) {
                    cbuf[slot] = make_uint2((uint32_t)gi, (uint32_t)gj);
                } else {   // buffer full (a dense cluster): append directly
                    const unsigned long long pos = atomicAdd(&counters->n_cand, 1ull);
                    if (pos < cand_cap) cand[pos] = make_uint2((uint32_t)gi, (uint32_t)gj);
                }
            }
        }
      }
      __syncthreads();
      const unsigned m = min(cbuf_n, (unsigned)L2_CBUF);
      if (threadIdx.x == 0 && m) cbuf_base = atomicAdd(&counters->n_cand, (unsigned long long)m);
      __syncthreads();
      for (unsigned i = threadIdx.x; i < m; i += blockDim.x)
          if (cbuf_base + i < cand_cap) cand[cbuf_base + i] = cbuf[i];
      __syncthreads();
      if (threadIdx.x == 0) cbuf_n = 0;
      __syncthreads();
    }
    if (full_checks) atomicAdd(&checks_n, full_checks);
    __syncthreads();
    if (threadIdx.x == 0 && checks_n) atomicAdd(&counters->l2_warp_items, (unsigned long long)checks_n);
}

// level 2 of the integer-pipe level 1: unit = the 8 x 4 pairs (ty + 16 i, tx + 32 j) of a k_pairs_l1 thread
template <int K4>
__global__ void __launch_bounds__(256)
k_pairs_l2(const uint4* __restrict__ bitsA, const uint4* __restrict__ bitsB, int64_t nA, int64_t nB,
           const int2* __restrict__ queue, unsigned long long queue_cap, int threshold, int triangular,
           uint2* __restrict__ cand, unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    // one THREAD per queued unit = the 32 pairs of tile pair (I, J) that a level-1 thread could not reject;
    // operands straight from L2
    const unsigned long long n_units = min(counters->n_units, queue_cap);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long u = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u < n_units; u += stride) {
        const int2 unit = __ldg(&queue[u]);
        const int I = unit.x & 0x00ffffff, ty = (unit.x >> 24) & 15, J = unit.y & 0x00ffffff, tx = (unit.y >> 24) & 31;
        const uint4* gA = bitsA + (size_t)I * (K4 * TILE);
        const uint4* gB = bitsB + (size_t)J * (K4 * TILE);
        constexpr int NI = 8, NJ = 4;
        auto row_of = [&](int i) { return ty + 16 * i; };
        auto col_of = [&](int j) { return tx + 32 * j; };
        int acc[NI][NJ];
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = 0;
#pragma unroll
        for (int k4 = 0; k4 < K4; ++k4) {
            uint4 b[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = __ldg(&gB[k4 * TILE + col_of(j)]);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const uint4 a = __ldg(&gA[k4 * TILE + row_of(i)]);
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    acc[i][j] += __popc(a.x ^ b[j].x) + __popc(a.y ^ b[j].y) + __popc(a.z ^ b[j].z) + __popc(a.w ^ b[j].w);
            }
        }
        const int64_t gi0 = (int64_t)I * TILE, gj0 = (int64_t)J * TILE;
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (acc[i][j] <= threshold) {
                    const int64_t gi = gi0 + row_of(i), gj = gj0 + col_of(j);
                    if (gi < nA && gj < nB && (!triangular || gi < gj)) {
                        unsigned long long pos = atomicAdd(&counters->n_cand, 1ull);
                        if (pos < cand_cap) cand[pos] = make_uint2((uint32_t)gi, (uint32_t)gj);
                    }
                }
            }
    }
}

// ------------------------------------------------------------------------------------------
// K4: lock-free union-find (hook larger root under smaller root, path halving).  parent[v] <= v
// always holds, so the final root of a component is its smallest row index — a canonical label
// independent of edge order, rank count and scheduling.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int v) {
    volatile int* P = parent;
    int p = P[v];
    while (p != v) {
        int g = P[p];
        if (g != p) P[v] = g;  // halving; g is an ancestor of v, so this never breaks the forest
        v = p;
        p = g;
    }
    return v;
}

__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

__global__ void k_uf_init(int* __restrict__ parent, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (int)i;
}

// labels[i] = root of i (= smallest row of its component); *n_roots += rows that are their own root
__global__ void __launch_bounds__(256) k_uf_labels(int* __restrict__ parent, int64_t n, int32_t* __restrict__ labels,
                                                   unsigned int* __restrict__ n_roots) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool is_root = false;
    if (i < n) {
        const int r = uf_find(parent, (int)i);
        labels[i] = r;
        is_root = r == (int)i;
    }
    // one atomic per block: a same-address atomic per warp (31 250 of them at 10^6 rows) was most of this kernel's time
    __shared__ unsigned int block_roots;
    if (threadIdx.x == 0) block_roots = 0;
    __syncthreads();
    const unsigned int m = __ballot_sync(0xffffffffu, is_root);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&block_roots, (unsigned int)__popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && block_roots) atomicAdd(n_roots, block_roots);
}

__global__ void k_uf_edges(int* __restrict__ parent, const int32_t* __restrict__ src,
                           const int32_t* __restrict__ dst, int64_t n_edges) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n_edges) uf_unite(parent, src[e], dst[e]);
}

// gathered[r][i] = label of row i on rank r  ->  union(i, gathered[r][i])
__global__ void k_uf_merge_labels(int* __restrict__ parent, const int32_t* __restrict__ gathered, int64_t n,
                                  int world) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * world) return;
    int i = (int)(idx % n);
    int l = gathered[idx];
    if (l != i) uf_unite(parent, i, l);
}

// ---- exchange steps of a multi-GPU pass (the collectives themselves are NCCL all-gathers issued by api.cu) ----
// after the all-gather of the per-rank shares of the staged keys: row numbers and the OR word over ALL keys
__global__ void __launch_bounds__(256) k_keys_finalize(const sortkey_t* __restrict__ keys, int64_t n, int32_t* __restrict__ vals,
                                                       sortkey_t* __restrict__ or_key) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    sortkey_t k = 0;
    if (i < n) {
        k = keys[i];
        vals[i] = (int32_t)i;
    }
    or_reduce_key(k, or_key);
}

// this rank's union-find as a compact list: one entry (row | root << 32) per row that is not its own root
// (slot 0 of `mine` = the number of such rows; entries beyond `cap` are dropped and reported by the merge kernel)
__global__ void __launch_bounds__(256) k_uf_compact(int* __restrict__ parent, int64_t n, unsigned long long* __restrict__ mine,
                                                    unsigned long long cap) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int r = 0;
    bool changed = false;
    if (i < n) {
        r = uf_find(parent, (int)i);
        changed = r != (int)i;
    }
    // one cursor update per block (a same-address atomic per warp is 31 250 atomics at 10^6 rows)
    __shared__ unsigned int warp_n[8];
    __shared__ unsigned long long block_base;
    const unsigned int m = __ballot_sync(0xffffffffu, changed);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_n[warp] = (unsigned int)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned int c = warp_n[w];
            warp_n[w] = total;   // exclusive prefix over the warps
            total += c;
        }
        block_base = total ? atomicAdd(&mine[0], (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    if (changed) {
        const unsigned long long pos = block_base + warp_n[warp] + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) mine[1 + pos] = (unsigned long long)(uint32_t)i | ((unsigned long long)(uint32_t)r << 32);
    }
}

// gathered[r] = the compact list of rank r ((cap + 1) words each): union(row, root) for every entry of every rank
// (this rank's own list is already in its forest: only its length is looked at)
__global__ void __launch_bounds__(256) k_uf_merge_pairs(int* __restrict__ parent, const unsigned long long* __restrict__ gathered,
                                                        int world, int rank, unsigned long long cap, unsigned int* __restrict__ fullest) {
    const unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const int r = (int)(idx / cap);
    if (r >= world) return;
    const unsigned long long j = idx - (unsigned long long)r * cap;
    const unsigned long long* lst = gathered + (size_t)r * (cap + 1);
    const unsigned long long cnt = lst[0];
    if (j == 0) atomicMax(fullest, (unsigned int)min(cnt, 0xffffffffull));
    if (r != rank && j < min(cnt, cap)) {
        const unsigned long long e = lst[1 + j];
        uf_unite(parent, (int)(uint32_t)e, (int)(uint32_t)(e >> 32));
    }
}

// member lists (CSR): every member is united with the first member of its list
__global__ void k_uf_lists(int* __restrict__ parent, const int64_t* __restrict__ list_indptr,
                           const int32_t* __restrict__ members, int64_t n_lists, int64_t n_members) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_members) return;
    int64_t l = 0, r = n_lists;  // last list with indptr[l] <= e
    while (l < r) {
        int64_t m = (l + r) >> 1;
        if (list_indptr[m + 1] <= e) l = m + 1; else r = m;
    }
    int head = members[list_indptr[l]];
    int me = members[e];
    if (me != head) uf_unite(parent, head, me);
}

// ------------------------------------------------------------------------------------------
// K3b: exact verification of the candidates on the CSR rows and hook.  A warp takes a batch of 32
// candidates (one per lane), then works through them one at a time cooperatively: the lanes stride
// over the shorter row and match each column against a +-max_dist window of the longer row (exact for
// the <= max_dist decision, see below), a warp sum gives |A n B| and d = |A| + |B| - 2|A n B| — the
// quantity sklearn's two-pointer merge accumulates (_pairwise_fast.pyx:83-105), on integers.  Hooks and the edge append are then done per
// lane with one aggregated cursor update per batch.
// ------------------------------------------------------------------------------------------
// The plain CSR as the kernels that walk single rows see it
struct RowStore {
    const int64_t* indptr;
    const int32_t* indices;
};
__device__ __forceinline__ int row_col(const RowStore& m, int64_t base, int k) { return __ldg(m.indices + base + k); }
__device__ __forceinline__ void row_extent(const RowStore& m, int r, int64_t& base, int& len) {
    base = __ldg(&m.indptr[r]);
    len = (int)(__ldg(&m.indptr[r + 1]) - base);   // a row has fewer than 2^31 columns
}
// The compact resident form (CSR16, see k_csr16_encode): half the bytes per column.  A row's position is handed around
// as base | split << 40 (offsets are below 2^32, split = how many of the row's ascending columns are below 65536, at most
// 65535), so that row_col can rebuild the 17-bit column without another lookup.
struct RowStore16 {
    const uint32_t* indptr32;
    const uint16_t* split;   // null: every column is below 65536
    const uint16_t* lo;
};
__device__ __forceinline__ int row_col(const RowStore16& m, int64_t base, int k) {
    const uint32_t b = (uint32_t)base;
    const int sp = (int)(base >> 40);
    return (int)__ldg(m.lo + b + k) | (k >= sp ? 0x10000 : 0);
}
__device__ __forceinline__ void row_extent(const RowStore16& m, int r, int64_t& base, int& len) {
    const uint32_t b = __ldg(&m.indptr32[r]);
    len = (int)(__ldg(&m.indptr32[r + 1]) - b);
    const int64_t sp = m.split ? (int64_t)__ldg(&m.split[r]) : (int64_t)0xffffff;
    base = (int64_t)b | (sp << 40);
}

// One chunk of 128 columns of the shorter row A and the longer row B of a candidate, lane t holding positions
// base + t + 32 s (s < 4); lane t < NW also holds B[base - NW + t] and B[base + 128 + t] (the halo).
template <int NW, typename Rows>
__device__ __forceinline__ void verify_load_chunk(const Rows& m, int lane, int base, int64_t ia, int la, int64_t ib, int lb,
                                                  int (&av)[4], int (&bv)[4], int& halo_lo, int& halo_hi) {
#pragma unroll
    for (int sw = 0; sw < 4; ++sw) {
        const int k = base + lane + 32 * sw;
        av[sw] = k < la ? row_col(m, ia, k) : -2;   // column ids are >= 0: the fillers match nothing
        bv[sw] = k < lb ? row_col(m, ib, k) : -1;
    }
    halo_lo = -1;
    halo_hi = -1;
    if (lane < NW) {
        if (NW / 2 > 0 && base - NW + lane >= 0 && base - NW + lane < lb) halo_lo = row_col(m, ib, base - NW + lane);
        if (base + 128 + lane < lb) halo_hi = row_col(m, ib, base + 128 + lane);
    }
}
// How many of this lane's columns of A occur in B at the offsets -ND ... +NW: the neighbours of B come from lane
// rotations.  A is the shorter row: with a = |A \ B|, b = |B \ A| (b - a = |B| - |A| >= 0, a + b <= d) a common column
// sits at positions i in A and j in B with j - i in [-a, b], and a <= d / 2 - so the window below needs only half the
// reach of the window above (none at all at max_dist 1).  A match outside the exact bounds is still a true common
// column (the columns of a row are distinct), so the static window is sound.
template <int NW>
__device__ __forceinline__ int verify_match_chunk(int lane, const int (&av)[4], const int (&bv)[4], int halo_lo, int halo_hi) {
    constexpr int ND = NW / 2;
    constexpr int NDA = ND > 0 ? ND : 1;
    int dn[4][NDA], up[4][NW];   // bv[sw] rotated by -o / +o lanes
#pragma unroll
    for (int sw = 0; sw < 4; ++sw) {
#pragma unroll
        for (int o = 1; o <= ND; ++o) dn[sw][o - 1] = __shfl_sync(0xffffffffu, bv[sw], (lane - o) & 31);
#pragma unroll
        for (int o = 1; o <= NW; ++o) up[sw][o - 1] = __shfl_sync(0xffffffffu, bv[sw], (lane + o) & 31);
    }
    // lane < o of the first sweep needs B[base + lane - o] = halo_lo of lane NW + lane - o;
    // lane >= 32 - o of the last sweep needs B[base + 128 + lane + o - 32] = halo_hi of that lane
    int hlo[NDA], hhi[NW];
#pragma unroll
    for (int o = 1; o <= ND; ++o) hlo[o - 1] = __shfl_sync(0xffffffffu, halo_lo, (NW + lane - o) & 31);
#pragma unroll
    for (int o = 1; o <= NW; ++o) hhi[o - 1] = __shfl_sync(0xffffffffu, halo_hi, (lane + o) & 31);
    int hits = 0;
#pragma unroll
    for (int sw = 0; sw < 4; ++sw) {
        bool hit = av[sw] == bv[sw];
#pragma unroll
        for (int o = 1; o <= ND; ++o) {
            const int below = lane >= o ? dn[sw][o - 1] : (sw > 0 ? dn[sw - 1][o - 1] : hlo[o - 1]);        // B[k - o]
            hit |= av[sw] == below;
        }
#pragma unroll
        for (int o = 1; o <= NW; ++o) {
            const int above = lane + o < 32 ? up[sw][o - 1] : (sw < 3 ? up[sw + 1][o - 1] : hhi[o - 1]);   // B[k + o]
            hit |= av[sw] == above;
        }
        hits += hit ? 1 : 0;
    }
    return hits;
}

// Generic form of the windowed match for one candidate, whole warp (any max_dist, rows of any length): every lane
// strides over A and looks B up in the +-max_dist window.  Returns this lane's share of |A n B|.
template <typename Rows>
__device__ __forceinline__ int verify_pair_windowed(const Rows& m, int lane, int64_t ia, int la, int64_t ib, int lb, int max_dist) {
    int inter = 0;
    for (int k = lane; k < la; k += 32) {
        const int x = row_col(m, ia, k);
        bool hit = false;
        for (int o = -max_dist; o <= max_dist; ++o) {
            const int j = k + o;
            if (j >= 0 && j < lb) hit |= (row_col(m, ib, j) == x);
        }
        inter += hit ? 1 : 0;
    }
    return inter;
}

// hooks the verified candidates of a warp's batch into the union-find and appends them to the edge list
__device__ __forceinline__ void verify_emit(unsigned int edge_mask, int lane, int ra, int rb, int* __restrict__ parent,
                                            uint2* __restrict__ edges, unsigned long long edge_cap, DevCounters* __restrict__ counters) {
    const bool is_edge = (edge_mask >> lane) & 1u;
    if (is_edge) uf_unite(parent, ra, rb);
    if (edge_mask) {
        unsigned long long pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(&counters->n_edges, (unsigned long long)__popc(edge_mask));
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (is_edge && edges) {
            const unsigned long long pos = pos0 + __popc(edge_mask & ((1u << lane) - 1u));
            if (pos < edge_cap) edges[pos] = make_uint2((uint32_t)min(ra, rb), (uint32_t)max(ra, rb));
        }
    }
}

// (Measured at 10^6 profiles, max-dist 1, 750 k candidates: 170 us with 44 registers / 5 resident blocks per SM; a
// register budget for 6 or 8 blocks - 40 / 32 registers, a few spills - 180 / 183 us, so more warps in flight do not
// help: the kernel moves 541 MB of scattered 360-byte rows at 3.2 TB/s, which is what random sector-granular reads reach.)
template <int DWIN, typename Rows>
__global__ void __launch_bounds__(256)
k_verify_unite(const uint2* __restrict__ cand, unsigned long long cand_cap, const int32_t* __restrict__ permA,
               const int32_t* __restrict__ permB, const Rows m, int max_dist, int already_exact,
               const unsigned char* __restrict__ is_query, int* __restrict__ parent, uint2* __restrict__ edges,
               unsigned long long edge_cap, DevCounters* __restrict__ counters) {
    const unsigned long long n = min(counters->n_cand, cand_cap);
    const int lane = threadIdx.x & 31;
    const unsigned long long warp_id = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    // a warp's candidates are worked through one after the other, so a lone batch of 32 is the kernel's minimum time:
    // when there are fewer candidates than 32 per warp (multi-GPU shares, small inputs) the batches shrink so that
    // every warp gets some.  (Equalising larger loads the same way was measured slower, see below.)
    // (Handing the batches out through a ticket counter instead: 190 us against 172 us at 10^6 profiles - measured, dropped.)
    const unsigned long long per_warp = (n + n_warps - 1) / n_warps;
    const unsigned long long batch = per_warp < 32 ? (per_warp ? per_warp : 1) : 32;
    for (unsigned long long base = warp_id * batch; base < n; base += n_warps * batch) {
        const unsigned long long c = base + lane;
        int ra = 0, rb = 0;
        bool keep = false;
        if ((unsigned long long)lane < batch && c < n) {
            const uint2 pr = cand[c];
            ra = permA ? permA[pr.x] : (int)pr.x;   // the hash-join engine emits row numbers, the tile engines sorted positions
            rb = permB ? permB[pr.y] : (int)pr.y;
            keep = true;
            if (is_query) keep = (ra != rb) && !(is_query[rb] && ra > rb);
        }
        unsigned int edge_mask = __ballot_sync(0xffffffffu, keep);
        if (!already_exact) {
            // row extents of all 32 candidates at once (one latency for the batch, not one per candidate)
            int64_t my_ia = 0, my_ib = 0;
            int my_la = 0, my_lb = 0;
            if (keep) {
                row_extent(m, ra, my_ia, my_la);
                row_extent(m, rb, my_ib, my_lb);
                if (my_la > my_lb) {   // A = the shorter row
                    const int64_t t = my_ia; my_ia = my_ib; my_ib = t;
                    const int tl = my_la; my_la = my_lb; my_lb = tl;
                }
            }
            // a pair whose lengths differ by more than max_dist is further apart than that already
            unsigned int todo = __ballot_sync(0xffffffffu, keep && my_lb - my_la <= max_dist);
            edge_mask = 0;
            // Both rows ascend.  If |A xor B| <= max_dist, a common column sits at positions that differ by at most
            // max_dist in the two rows (at most that many one-sided columns precede it), so matching A[k] against
            // B[k-d..k+d] finds every common column; if the distance is larger the count can only be too small, i.e. the
            // pair is still rejected.  No data-dependent addressing.
            while (todo) {
                const int l = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const int64_t ia = __shfl_sync(0xffffffffu, my_ia, l), ib = __shfl_sync(0xffffffffu, my_ib, l);
                const int la = __shfl_sync(0xffffffffu, my_la, l), lb = __shfl_sync(0xffffffffu, my_lb, l);
                int inter = 0;
                if (DWIN > 0) {
                    // The rows are worked through in chunks of 128 columns (one chunk for nearly all rows): each lane loads
                    // its <= 4 columns of either row once (all loads of a chunk independent and issued back to back, so a
                    // chunk costs about one memory round trip) and gets the +-DWIN neighbours of B from lane rotations
                    // instead of more loads.  (Measured and dropped: equalising the batches over the warps, 289 vs 238 us;
                    // requesting the next candidate's chunk before this one's is consumed, 208 vs 192 us - the extra
                    // registers cost a resident block per SM.)
                    constexpr int NW = DWIN > 0 ? DWIN : 1;
                    int av[4], bv[4], halo_lo, halo_hi;
                    for (int cbase = 0; cbase < la; cbase += 128) {
                        verify_load_chunk<NW, Rows>(m, lane, cbase, ia, la, ib, lb, av, bv, halo_lo, halo_hi);
                        inter += verify_match_chunk<NW>(lane, av, bv, halo_lo, halo_hi);
                    }
                } else {
                    inter = verify_pair_windowed(m, lane, ia, la, ib, lb, max_dist);
                }
                inter = __reduce_add_sync(0xffffffffu, inter);
                if ((int64_t)la + lb - 2 * (int64_t)inter <= (int64_t)max_dist) edge_mask |= 1u << l;
            }
        }
        verify_emit(edge_mask, lane, ra, rb, parent, edges, edge_cap, counters);
    }
}

// ------------------------------------------------------------------------------------------
// Compact host form of the CSR ("CSR16", include/breakfast_b200.h): 32-bit row offsets, the low 16 bits of every column
// and, per row, how many of its (ascending) columns are below 65536.  k_csr16_decode rebuilds the int64 / int32 CSR the
// other kernels read: a warp per row, 128-bit stores where the row start allows.  HBM: reads 2 nnz + 6 N bytes, writes
// 4 nnz + 8 N bytes (the host link carries half of what the plain form needs).
// ------------------------------------------------------------------------------------------
// The opposite direction, once per uploaded matrix (bf_upload_csr): plain CSR -> compact resident form, so that every
// pass over the matrix streams 2 bytes per column instead of 4.  A warp per row.  `bad` is raised if the matrix is not
// representable after all (a column outside [0, 131072), a row that is not strictly ascending or has more than 65535
// columns); the caller then keeps using the plain form.
__global__ void __launch_bounds__(256) k_csr16_encode(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                      int64_t n, uint32_t* __restrict__ indptr32, uint16_t* __restrict__ split,
                                                      uint16_t* __restrict__ lo, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp0 * 32; r0 <= n; r0 += n_warps * 32) {
        const int64_t r = r0 + lane;
        int64_t my_b = 0, my_e = 0;
        if (r <= n) {
            my_b = __ldg(&indptr[r]);
            indptr32[r] = (uint32_t)my_b;
            my_e = r < n ? __ldg(&indptr[r + 1]) : my_b;
        }
        bool wrong = my_e - my_b > 65535;
        for (int l = 0; l < 32 && r0 + l < n; ++l) {
            const int64_t b = __shfl_sync(0xffffffffu, my_b, l), e = __shfl_sync(0xffffffffu, my_e, l);
            uint32_t low = 0;
            for (int64_t k = b + lane; k < e; k += 32) {
                const int32_t col = __ldg(&indices[k]);
                wrong |= (uint32_t)col >= 131072u || (k > b && __ldg(&indices[k - 1]) >= col);
                low += (uint32_t)col < 65536u ? 1u : 0u;
                lo[k] = (uint16_t)((uint32_t)col & 0xffffu);
            }
            low = __reduce_add_sync(0xffffffffu, low);
            if (lane == 0) split[r0 + l] = (uint16_t)min(low, 65535u);
        }
        if (wrong) atomicOr(bad, 1);
    }
}

__global__ void __launch_bounds__(256) k_csr16_decode(const uint32_t* __restrict__ indptr32, const uint16_t* __restrict__ split,
                                                      const uint16_t* __restrict__ lo, int64_t n, int64_t* __restrict__ indptr,
                                                      int32_t* __restrict__ indices) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp0 * 32; r0 <= n; r0 += n_warps * 32) {
        // 32 rows per round: lane l fetches the extent of row r0 + l once, the warp then walks the rows
        const int64_t r = r0 + lane;
        uint32_t my_b = 0, my_e = 0, my_s = 0xffffffffu;
        if (r <= n) {
            my_b = __ldg(&indptr32[r]);
            indptr[r] = (int64_t)my_b;
            if (r < n) {
                my_e = __ldg(&indptr32[r + 1]);
                if (split) my_s = __ldg(&split[r]);
            }
        }
        for (int l = 0; l < 32 && r0 + l < n; ++l) {
            const uint32_t b = __shfl_sync(0xffffffffu, my_b, l), e = __shfl_sync(0xffffffffu, my_e, l);
            const uint32_t sp = __shfl_sync(0xffffffffu, my_s, l);
            for (uint32_t k = b + lane; k < e; k += 32)
                indices[k] = (int32_t)((uint32_t)__ldg(&lo[k]) | (k - b >= sp ? 0x10000u : 0u));
        }
    }
}

// ------------------------------------------------------------------------------------------
// K7: hash-join engine (max_dist <= 2) - exact candidate generation in O(nnz * m^(d-1)) probes instead of tile pairs.
// Two different rows A, B (ascending unique columns) are at distance
//     1  iff  one is the other minus one column,
//     2  iff  one is the other minus two columns ("superset"), or |A| = |B| and A minus some x equals B minus some y ("swap").
// With the additive row hash H(A) = sum of g(col) mod 2^64 (g = splitmix64) every case is an equi-join:
//     H(A) - g(x) = H(B), |B| = |A| - 1;     H(A) - g(x) - g(y) = H(B), |B| = |A| - 2;     H(A) - g(x) = H(B) - g(y), |A| = |B|.
// A true edge always satisfies its join, so none is lost; every match goes through k_verify_unite (exact distance on the
// rows), so a hash or tag collision cannot add one.  Tables: open addressing, one 64-bit word per entry
// (tag = upper key half, row), linear probing; T1 = the rows (L2-resident at 10^6 rows), T2 = every one-deletion key
// (swap join).  Same definition of an edge as the tile engines (breakfast.py:223-276 + sklearn _pairwise_fast.pyx:83-107);
// no counterpart in the reference.
// ------------------------------------------------------------------------------------------
constexpr unsigned long long HJ_EMPTY = ~0ull;
constexpr int HJ_CBUF = 1024;   // candidates a block collects before one cursor update

__device__ __forceinline__ unsigned long long hj_g(uint32_t col) {
    unsigned long long x = (unsigned long long)col + 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long hj_mix(unsigned long long k) {   // slot hash: sums of g are uniform, but cheap insurance
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    return k ^ (k >> 29);
}
__device__ __forceinline__ void hj_insert(unsigned long long* __restrict__ table, unsigned long long mask, unsigned long long key, int row) {
    const unsigned long long entry = (key & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)row;
    unsigned long long slot = hj_mix(key) & mask;
    while (atomicCAS(&table[slot], HJ_EMPTY, entry) != HJ_EMPTY) slot = (slot + 1) & mask;
}

// block-level candidate buffer: hits are rare (about one per hundred probes), a same-address atomic per hit would
// serialise the whole kernel
struct HjOut {
    uint2* buf;              // shared, HJ_CBUF entries
    unsigned* n;             // shared counter
    uint2* cand;
    unsigned long long cap;
    unsigned long long* cursor;
};
__device__ __forceinline__ void hj_emit(const HjOut& o, int a, int b) {
    const unsigned slot = atomicAdd(o.n, 1u);
    if (slot < HJ_CBUF) {
        o.buf[slot] = make_uint2((uint32_t)a, (uint32_t)b);
    } else {
        const unsigned long long pos = atomicAdd(o.cursor, 1ull);
        if (pos < o.cap) o.cand[pos] = make_uint2((uint32_t)a, (uint32_t)b);
    }
}
__device__ __forceinline__ void hj_flush(const HjOut& o, unsigned long long* base_s) {   // all threads of the block
    __syncthreads();
    const unsigned m = min(*o.n, (unsigned)HJ_CBUF);
    if (threadIdx.x == 0 && m) *base_s = atomicAdd(o.cursor, (unsigned long long)m);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < m; i += blockDim.x)
        if (*base_s + i < o.cap) o.cand[*base_s + i] = o.buf[i];
}

// H[r] and insertion into T1; eight lanes per row
__global__ void __launch_bounds__(256) k_hj_hash_rows(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                                                      unsigned long long* __restrict__ H, unsigned long long* __restrict__ t1,
                                                      unsigned long long mask1) {
    const int sub = threadIdx.x & 7;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    unsigned long long h = 0;
    if (r < n) {
        const int64_t b = __ldg(&indptr[r]), e = __ldg(&indptr[r + 1]);
        for (int64_t k = b + sub; k < e; k += 8) h += hj_g((uint32_t)__ldg(&indices[k]));
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (sub == 0 && r < n) {
        H[r] = h;
        hj_insert(t1, mask1, h, (int)r);
    }
}

// one-deletion keys of every row -> T2 (swap join): full 64-bit keys (a key cannot be recomputed from its row alone)
// with the rows in a parallel array, written after the slot is claimed and read by a later kernel; eight lanes per row.
// A key equal to HJ_EMPTY is stored as HJ_EMPTY - 1 (a collision like any other: matches are verified exactly).
__device__ __forceinline__ unsigned long long hj_key2(unsigned long long key) { return key == HJ_EMPTY ? HJ_EMPTY - 1 : key; }
__global__ void __launch_bounds__(256) k_hj_build_deletions(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                                                            const unsigned long long* __restrict__ H, unsigned long long* __restrict__ t2,
                                                            int32_t* __restrict__ t2_rows, unsigned long long mask2) {
    const int sub = threadIdx.x & 7;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    if (r >= n) return;
    const int64_t b = __ldg(&indptr[r]), e = __ldg(&indptr[r + 1]);
    const unsigned long long h = __ldg(&H[r]);
    for (int64_t k = b + sub; k < e; k += 8) {
        const unsigned long long key = hj_key2(h - hj_g((uint32_t)__ldg(&indices[k])));
        unsigned long long slot = hj_mix(key) & mask2;
        while (atomicCAS(&t2[slot], HJ_EMPTY, key) != HJ_EMPTY) slot = (slot + 1) & mask2;
        t2_rows[slot] = (int32_t)r;
    }
}

// probes of the one-deletion keys: MODE 1 against T1 (distance 1: |B| = |A| - 1), MODE 2 against T2 (distance-2 swaps:
// |B| = |A|, B > A so that a pair is reported once).  Row q of this rank is row q * world + rank; eight lanes per row.
template <int MODE>
__global__ void __launch_bounds__(256) k_hj_probe_deletions(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                                                            const unsigned long long* __restrict__ H, const unsigned long long* __restrict__ table,
                                                            const int32_t* __restrict__ t2_rows, unsigned long long mask, int rank, int world,
                                                            const unsigned char* __restrict__ is_query, uint2* __restrict__ cand,
                                                            unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    __shared__ uint2 cbuf[HJ_CBUF];
    __shared__ unsigned cbuf_n;
    __shared__ unsigned long long cbuf_base;
    if (threadIdx.x == 0) cbuf_n = 0;
    __syncthreads();
    const HjOut out{cbuf, &cbuf_n, cand, cand_cap, &counters->n_cand};
    const int sub = threadIdx.x & 7;
    const int64_t r = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3) * world + rank;
    if (r < n) {
        const int64_t b = __ldg(&indptr[r]), e = __ldg(&indptr[r + 1]);
        const int64_t want_len = MODE == 1 ? e - b - 1 : e - b;
        const unsigned long long h = __ldg(&H[r]);
        const bool a_is_q = is_query ? is_query[r] != 0 : true;
        for (int64_t k = b + sub; k < e; k += 8) {
            unsigned long long key = h - hj_g((uint32_t)__ldg(&indices[k]));
            if (MODE == 2) key = hj_key2(key);
            const unsigned long long tag = key & 0xffffffff00000000ull;
            unsigned long long slot = hj_mix(key) & mask;
            while (true) {
                const unsigned long long ent = __ldg(&table[slot]);
                if (ent == HJ_EMPTY) break;
                if (MODE == 1) {
                    if ((ent & 0xffffffff00000000ull) == tag) {   // T1: tag + row; the full key is the row's hash
                        const int rb = (int)(uint32_t)ent;
                        if (__ldg(&H[rb]) == key && __ldg(&indptr[rb + 1]) - __ldg(&indptr[rb]) == want_len && (a_is_q || is_query[rb]))
                            hj_emit(out, (int)r, rb);
                    }
                } else if (ent == key) {                          // T2: full keys, rows beside them
                    const int rb = __ldg(&t2_rows[slot]);
                    // rows with equal hashes (identical rows, or a collision) are the business of k_hj_probe_equal
                    if (rb > (int)r && __ldg(&H[rb]) != h && __ldg(&indptr[rb + 1]) - __ldg(&indptr[rb]) == want_len &&
                        (a_is_q || is_query[rb]))
                        hj_emit(out, (int)r, rb);
                }
                slot = (slot + 1) & mask;
            }
        }
    }
    hj_flush(out, &cbuf_base);
}

// distance-2 supersets: H(A) - g(x) - g(y) against T1 for every pair x < y of A's columns; a warp per row
__global__ void __launch_bounds__(256) k_hj_probe_pairs(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                                                        const unsigned long long* __restrict__ H, const unsigned long long* __restrict__ t1,
                                                        unsigned long long mask1, int rank, int world,
                                                        const unsigned char* __restrict__ is_query, uint2* __restrict__ cand,
                                                        unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    __shared__ uint2 cbuf[HJ_CBUF];
    __shared__ unsigned cbuf_n;
    __shared__ unsigned long long cbuf_base;
    if (threadIdx.x == 0) cbuf_n = 0;
    __syncthreads();
    const HjOut out{cbuf, &cbuf_n, cand, cand_cap, &counters->n_cand};
    const int lane = threadIdx.x & 31;
    const int64_t r = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * world + rank;
    if (r < n) {
        const int64_t b = __ldg(&indptr[r]), e = __ldg(&indptr[r + 1]);
        const int64_t want_len = e - b - 2;
        const unsigned long long h = __ldg(&H[r]);
        const bool a_is_q = is_query ? is_query[r] != 0 : true;
        for (int64_t i = b; i + 1 < e; ++i) {
            const unsigned long long hx = h - hj_g((uint32_t)__ldg(&indices[i]));
            for (int64_t j = i + 1 + lane; j < e; j += 32) {
                const unsigned long long key = hx - hj_g((uint32_t)__ldg(&indices[j]));
                const unsigned long long tag = key & 0xffffffff00000000ull;
                unsigned long long slot = hj_mix(key) & mask1;
                while (true) {
                    const unsigned long long ent = __ldg(&t1[slot]);
                    if (ent == HJ_EMPTY) break;
                    if ((ent & 0xffffffff00000000ull) == tag) {
                        const int rb = (int)(uint32_t)ent;
                        if (__ldg(&H[rb]) == key && __ldg(&indptr[rb + 1]) - __ldg(&indptr[rb]) == want_len && (a_is_q || is_query[rb]))
                            hj_emit(out, (int)r, rb);
                    }
                    slot = (slot + 1) & mask1;
                }
            }
        }
    }
    hj_flush(out, &cbuf_base);
}

// rows with equal hashes (identical rows; or, with probability 2^-64 per pair, different ones - the verify step decides):
// every row looks its own hash up in T1 and reports the equal-length rows after it
__global__ void __launch_bounds__(256) k_hj_probe_equal(const int64_t* __restrict__ indptr, int64_t n, const unsigned long long* __restrict__ H,
                                                        const unsigned long long* __restrict__ t1, unsigned long long mask1, int rank,
                                                        int world, const unsigned char* __restrict__ is_query, uint2* __restrict__ cand,
                                                        unsigned long long cand_cap, DevCounters* __restrict__ counters) {
    __shared__ uint2 cbuf[HJ_CBUF];
    __shared__ unsigned cbuf_n;
    __shared__ unsigned long long cbuf_base;
    if (threadIdx.x == 0) cbuf_n = 0;
    __syncthreads();
    const HjOut out{cbuf, &cbuf_n, cand, cand_cap, &counters->n_cand};
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * world + rank;
    if (r < n) {
        const unsigned long long key = __ldg(&H[r]), tag = key & 0xffffffff00000000ull;
        const int64_t len = __ldg(&indptr[r + 1]) - __ldg(&indptr[r]);
        const bool a_is_q = is_query ? is_query[r] != 0 : true;
        unsigned long long slot = hj_mix(key) & mask1;
        while (true) {
            const unsigned long long ent = __ldg(&t1[slot]);
            if (ent == HJ_EMPTY) break;
            if ((ent & 0xffffffff00000000ull) == tag) {
                const int rb = (int)(uint32_t)ent;
                if (rb > (int)r && __ldg(&H[rb]) == key && __ldg(&indptr[rb + 1]) - __ldg(&indptr[rb]) == len && (a_is_q || is_query[rb]))
                    hj_emit(out, (int)r, rb);
            }
            slot = (slot + 1) & mask1;
        }
    }
    hj_flush(out, &cbuf_base);
}

__global__ void k_mark_rows(const int32_t* __restrict__ rows, int64_t n, unsigned char* __restrict__ flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flags[rows[i]] = 1;
}

}  // namespace bf
