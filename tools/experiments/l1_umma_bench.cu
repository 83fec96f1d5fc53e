// Round-2 experiment: level 1 of the pair phase on tcgen05 (UTCIMMA / UTCQMMA) with TMEM accumulators, instrumented.
//   part 1: TMEM -> register epilogue rates (pure tcgen05.ld, + s32 max3 tree, + pack::16b with s16x2 max3, ...)
//   part 2: layout probe of an F16 accumulator written by kind::f8f6f4 (E4M3 +-1 operands), plain and pack::16b loads
//   part 3: the pipelined level-1 kernel (TMA producer, one MMA issuer, 4-16 epilogue warps, four 128-column
//           accumulators) on synthetic folds, K = 32 / 64 / 128, checked against the CPU, with clock64 stamps around
//           commit -> wake -> ld -> release -> reissue
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o l1_umma_bench l1_umma_bench.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", cudaGetErrorString(e_), __LINE__, #x); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle, one K = 32 byte chunk: core matrix = 8 rows x 16 B (128 B contiguous); the two K halves of a
// row group are 128 B apart (LBO), row groups 256 B apart (SBO)  [validated in umma_i8_test.cu]
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
template <int KIND>   // 0: kind::i8 (S8 x S8 -> S32), 1: kind::f8f6f4 (E4M3 x E4M3 -> F16)
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if constexpr (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                     "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                     "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
template <int KIND>
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    // c_format [4,6): F16 = 0, S32 = 2; a/b format [7,10) / [10,13): S8 = 1, E4M3 = 0; K-major both; N >> 3 at [17,23); M >> 4 at [24,29)
    return (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define O8(r, b) "=r"(r[b + 0]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3]), "=r"(r[b + 4]), "=r"(r[b + 5]), "=r"(r[b + 6]), "=r"(r[b + 7])
#define L16 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}"
#define L32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define L64 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, " \
            "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}"
// issue only; the caller waits (tmem_wait_ld) before touching r[]
template <bool PACK>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    if constexpr (PACK) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " L16 ", [%16];" : O8(r, 0), O8(r, 8) : "r"(taddr) : "memory");
    else asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 " L16 ", [%16];" : O8(r, 0), O8(r, 8) : "r"(taddr) : "memory");
}
template <bool PACK>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    if constexpr (PACK) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 ", [%32];" : O8(r, 0), O8(r, 8), O8(r, 16), O8(r, 24) : "r"(taddr) : "memory");
    else asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : O8(r, 0), O8(r, 8), O8(r, 16), O8(r, 24) : "r"(taddr) : "memory");
}
template <bool PACK>
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
    if constexpr (PACK)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 " L64 ", [%64];"
                     : O8(r, 0), O8(r, 8), O8(r, 16), O8(r, 24), O8(r, 32), O8(r, 40), O8(r, 48), O8(r, 56) : "r"(taddr) : "memory");
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 " L64 ", [%64];"
                     : O8(r, 0), O8(r, 8), O8(r, 16), O8(r, 24), O8(r, 32), O8(r, 40), O8(r, 48), O8(r, 56) : "r"(taddr) : "memory");
}

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }
template <int N> __device__ __forceinline__ int tree_s32(const uint32_t* r) {   // N = 32 or 64
    int m = (int)r[0];
#pragma unroll
    for (int k = 1; k + 1 < N; k += 2) m = max3(m, (int)r[k], (int)r[k + 1]);
    return max(m, (int)r[N - 1]);
}
template <int N> __device__ __forceinline__ uint32_t tree_s16x2(const uint32_t* r) {
    uint32_t m = r[0];
#pragma unroll
    for (int k = 1; k + 1 < N; k += 2) m = __vimax3_s16x2(m, r[k], r[k + 1]);
    return __vmaxs2(m, r[N - 1]);
}

// ------------------------------------------------------------------------------------------ part 1: epilogue rates
template <int MODE>
__global__ void __launch_bounds__((MODE == 4 || MODE == 6) ? 256 : 512, 1) k_epi_rate(int iters, uint32_t mul, uint32_t* out, long long* cycles) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * 128) & 511);
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {          // pure load
            uint32_t r[64];
            tmem_ld64<false>(base, r);
            tmem_wait_ld();
            acc |= r[0] | r[63];
        } else if constexpr (MODE == 1) {   // s32 max3 tree (round-1 epilogue)
            uint32_t r[64];
            tmem_ld64<false>(base, r);
            tmem_wait_ld();
            acc = (uint32_t)max((int)acc, tree_s32<64>(r));
        } else if constexpr (MODE == 2) {   // pack::16b load (64 registers) + s16x2 max3 tree
            uint32_t r[64];
            tmem_ld64<true>(base, r);
            tmem_wait_ld();
            acc = __vmaxs2(acc, tree_s16x2<64>(r));
        } else if constexpr (MODE == 3) {   // s32 load, IMAD pack of two values (FMA pipe), s16x2 max3 tree over 32
            uint32_t r[64], p[32];
            tmem_ld64<false>(base, r);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) p[k] = r[2 * k + 1] * mul + r[2 * k];   // mul = 65536 (opaque: stays an IMAD)
            acc = __vmaxs2(acc, tree_s16x2<32>(p));
        } else if constexpr (MODE == 4) {   // two loads in flight, one wait, s32 tree
            uint32_t r[128];
            tmem_ld64<false>(base, r);
            tmem_ld64<false>(base + 64, r + 64);
            tmem_wait_ld();
            acc = (uint32_t)max((int)acc, max(tree_s32<64>(r), tree_s32<64>(r + 64)));
        } else if constexpr (MODE == 5) {   // pack::16b x32 (32 registers) + s16x2 tree
            uint32_t r[32];
            tmem_ld32<true>(base, r);
            tmem_wait_ld();
            acc = __vmaxs2(acc, tree_s16x2<32>(r));
        } else if constexpr (MODE == 6) {   // two packed loads in flight
            uint32_t r[128];
            tmem_ld64<true>(base, r);
            tmem_ld64<true>(base + 128, r + 64);
            tmem_wait_ld();
            acc = __vmaxs2(acc, __vmaxs2(tree_s16x2<64>(r), tree_s16x2<64>(r + 64)));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

template <int MODE>
void run_epi_rate(const char* what, int regs_per_iter, uint32_t* d_out, long long* d_cyc) {
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        if (warps == 16 && (MODE == 4 || MODE == 6)) continue;   // 128 live registers per thread
        for (int rep = 0; rep < 2; ++rep) k_epi_rate<MODE><<<148, warps * 32>>>(iters, 65536u, d_out, d_cyc);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("  [epi] %-58s warps %2d: %7.1f registers/clk/SM\n", what, warps, (double)iters * warps * 32 * regs_per_iter / cyc);
    }
}

// ------------------------------------------------------------------------------------------ part 2: F16 accumulator probe
__device__ __forceinline__ uint4 expand_bytes(uint32_t bits16, uint32_t plus, uint32_t minus) {   // bit -> byte (clear: plus, set: minus)
    uint32_t w[4];
    for (int q = 0; q < 4; ++q) {
        const uint32_t spread = (((bits16 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
        w[q] = (0x01010101u * plus) ^ (spread * (plus ^ minus));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) k_probe(const uint32_t* foldA, const uint32_t* foldB, uint32_t* D, uint32_t* Dp) {
    __shared__ __align__(128) uint4 sA[128 * 2];
    __shared__ __align__(128) uint4 sB[256 * 2];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t plus = KIND == 0 ? 0x01u : 0x38u, minus = KIND == 0 ? 0xffu : 0xb8u;
    for (int r = tid; r < 128; r += 128) {
        sA[((r >> 3) * 2 + 0) * 8 + (r & 7)] = expand_bytes(foldA[r] & 0xffff, plus, minus);
        sA[((r >> 3) * 2 + 1) * 8 + (r & 7)] = expand_bytes(foldA[r] >> 16, plus, minus);
    }
    for (int r = tid; r < 256; r += 128) {
        sB[((r >> 3) * 2 + 0) * 8 + (r & 7)] = expand_bytes(foldB[r] & 0xffff, plus, minus);
        sB[((r >> 3) * 2 + 1) * 8 + (r & 7)] = expand_bytes(foldB[r] >> 16, plus, minus);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        umma<KIND>(tmem, umma_desc(smem_u32(sA)), umma_desc(smem_u32(sB)), make_idesc<KIND>(256), 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 256; c += 16) {
        uint32_t r[16];
        tmem_ld16<false>(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tmem_wait_ld();
        for (int k = 0; k < 16; ++k) D[tid * 256 + c + k] = r[k];
    }
    for (int c = 0; c < 256; c += 16) {   // packed: register k of the load at column c
        uint32_t r[16];
        tmem_ld16<true>(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tmem_wait_ld();
        for (int k = 0; k < 16; ++k) Dp[tid * 256 + c + k] = r[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static uint16_t f16_bits_of_int(int v) {   // exact for |v| <= 2048
    if (v == 0) return 0;
    uint16_t s = v < 0 ? 0x8000 : 0;
    unsigned a = (unsigned)(v < 0 ? -v : v);
    int e = 31 - __builtin_clz(a);
    unsigned mant = (a << (10 - e)) & 0x3ff;
    return (uint16_t)(s | ((e + 15) << 10) | mant);
}

// ------------------------------------------------------------------------------------------ part 3: the pipeline
constexpr int TILE = 128, GROUP = 4, CHUNK = TILE * 32;   // CHUNK: one K = 32 chunk of one 128-row tile (4 KB)
constexpr int NSLOT = 4;                                  // 128-column TMEM accumulators

template <int KB> struct Cfg {
    static constexpr int NCH = KB / 32;
    static constexpr int STAGE_BYTES = NCH * (1 + GROUP) * CHUNK;
    static constexpr int STAGES = KB == 32 ? 6 : (KB == 64 ? 4 : 2);
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int SMEM = BAR_OFF + 512;
};

struct Dbg {
    unsigned long long n_out;
    long long total, mma_wait_full, mma_wait_empty, epi_wait_full, lat_commit_wake, lat_ld, lat_release_reissue, n_sub, epi_tree;
};

// KB: bytes (= bits of the fold) per row; KIND as above; EW: epilogue warps (4, 8, 16); SPLIT: all groups share every
// accumulator (columns split) vs each group of 4 warps takes whole accumulators in turn
template <int KB, int KIND, int EW, bool SPLIT>
__global__ void __launch_bounds__((2 + EW) * 32, 1)
k_l1_umma(const unsigned char* __restrict__ planes, size_t plane_stride, const int2* __restrict__ items, int n_items, int thr,
          uint2* __restrict__ out, unsigned long long out_cap, Dbg* __restrict__ dbg) {
    using C = Cfg<KB>;
    constexpr int G = EW / 4;
    constexpr int CPT = SPLIT ? TILE / G : TILE;          // accumulator columns per epilogue thread
    constexpr bool PACK = KIND == 1;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* smem_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* smem_empty = smem_full + C::STAGES;
    uint64_t* tmem_full = smem_empty + C::STAGES;
    uint64_t* tmem_empty = tmem_full + NSLOT;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_empty + NSLOT);
    volatile long long* t_commit = reinterpret_cast<volatile long long*>(tmem_base_s + 2);   // [NSLOT]
    volatile long long* t_release = t_commit + NSLOT;                                          // [NSLOT]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = clock64();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&smem_full[s], 1); mbar_init(&smem_empty[s], 1); }
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], SPLIT ? EW : 4); }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        // ------------------------------- TMA producer -------------------------------
        uint32_t it = 0;
        for (int w0 = blockIdx.x; w0 < n_items; w0 += 32 * gridDim.x) {
            const int w = w0 + lane * gridDim.x;
            int2 mine = make_int2(0, 0);
            if (w < n_items) mine = __ldg(&items[w]);
            for (int l = 0; l < 32; ++l) {
                if (w0 + l * (int)gridDim.x >= n_items) break;
                const int Il = __shfl_sync(0xffffffffu, mine.x, l), Jp = __shfl_sync(0xffffffffu, mine.y, l);
                if (lane == 0) {
                    const int J0 = Jp & 0x1fffffff, cnt = ((unsigned)Jp >> 29) + 1;
                    const uint32_t stage = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    mbar_wait(&smem_empty[stage], ph ^ 1u);
                    unsigned char* sa = smem + stage * C::STAGE_BYTES;
                    mbar_arrive_expect_tx(&smem_full[stage], (uint32_t)(C::NCH * (1 + cnt)) * CHUNK);
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c) {
                        const unsigned char* pl = planes + (size_t)c * plane_stride;
                        bulk_g2s(sa + c * CHUNK, pl + (size_t)Il * CHUNK, CHUNK, &smem_full[stage]);
                        bulk_g2s(sa + C::NCH * CHUNK + c * (GROUP * CHUNK), pl + (size_t)J0 * CHUNK, (uint32_t)cnt * CHUNK, &smem_full[stage]);
                    }
                    ++it;
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer -------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<KIND>(TILE);
            uint32_t it = 0, sub_it = 0;
            long long w_full = 0, w_empty = 0, lat_rr = 0;
            int2 nxt = make_int2(0, 0);
            if ((int)blockIdx.x < n_items) nxt = __ldg(&items[blockIdx.x]);
            for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
                const int2 item = nxt;
                if (w + (int)gridDim.x < n_items) nxt = __ldg(&items[w + gridDim.x]);
                const int cnt = ((unsigned)item.y >> 29) + 1;
                const uint32_t stage = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                long long t0 = clock64();
                mbar_wait(&smem_full[stage], ph);
                w_full += clock64() - t0;
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                for (int jt = 0; jt < cnt; ++jt, ++sub_it) {
                    const uint32_t slot = sub_it % NSLOT, tph = (sub_it / NSLOT) & 1u;
                    t0 = clock64();
                    mbar_wait(&tmem_empty[slot], tph ^ 1u);
                    const long long t1 = clock64();
                    w_empty += t1 - t0;
                    if (sub_it >= NSLOT) lat_rr += t1 - t_release[slot];
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c)
                        umma<KIND>(tmem_base + slot * TILE, umma_desc(sa + c * CHUNK),
                                   umma_desc(sa + C::NCH * CHUNK + c * (GROUP * CHUNK) + jt * CHUNK), idesc, c > 0 ? 1u : 0u);
                    umma_commit(&tmem_full[slot]);
                    t_commit[slot] = clock64();
                }
                umma_commit(&smem_empty[stage]);
            }
            if (blockIdx.x == 0) {
                dbg->mma_wait_full = w_full;
                dbg->mma_wait_empty = w_empty;
                dbg->lat_release_reissue = lat_rr;
                dbg->n_sub = sub_it;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3, g = (warp - 2) >> 2;
        // thr: KIND 0 = KB - 2 d on the S32 accumulators; KIND 1 = the f16 bit pattern of (KB - 2 d): positive f16 values
        // compare like 16-bit integers and negative ones have the sign bit set, so the test is a signed 16-bit compare
        uint32_t sub_it = 0;
        long long w_full = 0, l_wake = 0, l_ld = 0, l_tree = 0;
        int2 nxt = make_int2(0, 0);
        if ((int)blockIdx.x < n_items) nxt = __ldg(&items[blockIdx.x]);
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int2 item = nxt;
            if (w + (int)gridDim.x < n_items) nxt = __ldg(&items[w + gridDim.x]);
            const int I = item.x, J0 = item.y & 0x1fffffff, cnt = ((unsigned)item.y >> 29) + 1;
            const uint32_t gi = (uint32_t)I * TILE + q * 32 + lane;
            for (int jt = 0; jt < cnt; ++jt, ++sub_it) {
                if (!SPLIT && (int)(sub_it % G) != g) continue;
                const uint32_t slot = sub_it % NSLOT, tph = (sub_it / NSLOT) & 1u;
                const long long t0 = clock64();
                mbar_wait(&tmem_full[slot], tph);
                const long long t1 = clock64();
                w_full += t1 - t0;
                l_wake += t1 - t_commit[slot];
                tc_fence_after();
                const int col0 = SPLIT ? g * CPT : 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * TILE + col0;
                constexpr int NREG = PACK ? CPT / 2 : CPT;
                uint32_t r[NREG];
                if constexpr (NREG == 128) { tmem_ld64<PACK>(taddr, r); tmem_ld64<PACK>(taddr + (PACK ? 128 : 64), r + 64); }
                else if constexpr (NREG == 64) tmem_ld64<PACK>(taddr, r);
                else if constexpr (NREG == 32) tmem_ld32<PACK>(taddr, r);
                else tmem_ld16<PACK>(taddr, r);
                tmem_wait_ld();
                const long long t2 = clock64();
                l_ld += t2 - t1;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { t_release[slot] = clock64(); mbar_arrive(&tmem_empty[slot]); }
                bool any;
                if constexpr (!PACK) {
                    int m = (int)r[0];
#pragma unroll
                    for (int k = 1; k + 1 < NREG; k += 2) m = max3(m, (int)r[k], (int)r[k + 1]);
                    m = max(m, (int)r[NREG - 1]);
                    any = m >= thr;
                } else {
                    uint32_t m = r[0];
#pragma unroll
                    for (int k = 1; k + 1 < NREG; k += 2) m = __vimax3_s16x2(m, r[k], r[k + 1]);
                    m = __vmaxs2(m, r[NREG - 1]);
                    any = (int)(short)(m & 0xffffu) >= thr || (int)(short)(m >> 16) >= thr;
                }
                l_tree += clock64() - t2;
                if (any) {
#pragma unroll
                    for (int k = 0; k < NREG; ++k) {
                        if constexpr (!PACK) {
                            if ((int)r[k] >= thr) {
                                const uint32_t gj = (uint32_t)(J0 + jt) * TILE + col0 + k;
                                if (gi != gj) {
                                    const unsigned long long pos = atomicAdd(&dbg->n_out, 1ull);
                                    if (pos < out_cap) out[pos] = make_uint2(gi, gj);
                                }
                            }
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if ((int)(short)((r[k] >> (16 * h)) & 0xffffu) >= thr) {
                                    const uint32_t gj = (uint32_t)(J0 + jt) * TILE + col0 + 2 * k + h;
                                    if (gi != gj) {
                                        const unsigned long long pos = atomicAdd(&dbg->n_out, 1ull);
                                        if (pos < out_cap) out[pos] = make_uint2(gi, gj);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (blockIdx.x == 0 && warp == 2 && lane == 0) {
            dbg->epi_wait_full = w_full;
            dbg->lat_commit_wake = l_wake;
            dbg->lat_ld = l_ld;
            dbg->epi_tree = l_tree;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) dbg->total = clock64() - t_start;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}


// ------------------------------------------------------------------------------------------ part 4: the throughput design
// Two 256-column accumulators (one MMA covers two column tiles), E4M3 +-1 operands, F16 accumulators read with pack::16b,
// EW epilogue warps splitting the 256 columns (each warp: its lane quadrant x 256 / (EW/4) columns), s16x2 max3 trees in
// four interleaved chains, no stamps in the hot loop.
struct Stamps { long long v[16]; };
#define STAMP(i) do { if constexpr (INSTR) { const long long t_ = clock64(); st.v[i] += t_ - tprev; tprev = t_; } } while (0)
template <int KB, int EW, bool INSTR>
__global__ void __launch_bounds__((2 + EW) * 32, 1)
k_l1_umma2(const unsigned char* __restrict__ planes, size_t plane_stride, const int2* __restrict__ items, int n_items, int thr,
           uint2* __restrict__ out, unsigned long long out_cap, Dbg* __restrict__ dbg, Stamps* __restrict__ stamps) {
    using C = Cfg<KB>;
    constexpr int G = EW / 4, CPT = 256 / G, NREG = CPT / 2;
    Stamps st;
    for (int i = 0; i < 16; ++i) st.v[i] = 0;
    long long tprev = clock64();
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* smem_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* smem_empty = smem_full + C::STAGES;
    uint64_t* tmem_full = smem_empty + C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&smem_full[s], 1); mbar_init(&smem_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EW); }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        uint32_t it = 0;
        for (int w0 = blockIdx.x; w0 < n_items; w0 += 32 * gridDim.x) {
            const int w = w0 + lane * gridDim.x;
            int2 mine = make_int2(0, 0);
            if (w < n_items) mine = __ldg(&items[w]);
            for (int l = 0; l < 32; ++l) {
                if (w0 + l * (int)gridDim.x >= n_items) break;
                const int Il = __shfl_sync(0xffffffffu, mine.x, l), Jp = __shfl_sync(0xffffffffu, mine.y, l);
                if (lane == 0) {
                    const int J0 = Jp & 0x1fffffff, cnt = ((unsigned)Jp >> 29) + 1;
                    const uint32_t stage = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                    mbar_wait(&smem_empty[stage], ph ^ 1u);
                    unsigned char* sa = smem + stage * C::STAGE_BYTES;
                    mbar_arrive_expect_tx(&smem_full[stage], (uint32_t)(C::NCH * (1 + cnt)) * CHUNK);
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c) {
                        const unsigned char* pl = planes + (size_t)c * plane_stride;
                        bulk_g2s(sa + c * CHUNK, pl + (size_t)Il * CHUNK, CHUNK, &smem_full[stage]);
                        bulk_g2s(sa + C::NCH * CHUNK + c * (GROUP * CHUNK), pl + (size_t)J0 * CHUNK, (uint32_t)cnt * CHUNK, &smem_full[stage]);
                    }
                    ++it;
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, sub_it = 0;
            int2 nxt = make_int2(0, 0);
            if ((int)blockIdx.x < n_items) nxt = __ldg(&items[blockIdx.x]);
            for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
                const int2 item = nxt;
                if (w + (int)gridDim.x < n_items) nxt = __ldg(&items[w + gridDim.x]);
                const int cnt = ((unsigned)item.y >> 29) + 1;
                const uint32_t stage = it % C::STAGES, ph = (it / C::STAGES) & 1u;
                const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                uint64_t da[C::NCH], db[C::NCH];
#pragma unroll
                for (int c = 0; c < C::NCH; ++c) {
                    da[c] = umma_desc(sa + c * CHUNK);
                    db[c] = umma_desc(sa + C::NCH * CHUNK + c * (GROUP * CHUNK));
                }
                STAMP(0);
                mbar_wait(&smem_full[stage], ph);
                tc_fence_after();
                STAMP(1);
                for (int sub = 0; 2 * sub < cnt; ++sub, ++sub_it) {
                    const uint32_t slot = sub_it & 1u, tph = (sub_it >> 1) & 1u;
                    const int ntiles = min(2, cnt - 2 * sub);
                    const uint32_t idesc = make_idesc<1>(TILE) + ((uint32_t)((ntiles - 1) * (TILE >> 3)) << 17);
                    STAMP(2);
                    mbar_wait(&tmem_empty[slot], tph ^ 1u);
                    STAMP(3);
                    tc_fence_after();
                    STAMP(4);
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c)
                        umma<1>(tmem_base + slot * 256, da[c], db[c] + (uint64_t)((sub * 2 * CHUNK) >> 4), idesc, c > 0 ? 1u : 0u);
                    STAMP(5);
                    umma_commit(&tmem_full[slot]);
                    STAMP(6);
                }
                umma_commit(&smem_empty[stage]);
                STAMP(7);
            }
            if (INSTR && blockIdx.x == 0) { st.v[15] = sub_it; stamps[0] = st; }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, g = (warp - 2) >> 2;
        const int col0 = g * CPT;
        uint32_t sub_it = 0;
        int2 nxt = make_int2(0, 0);
        if ((int)blockIdx.x < n_items) nxt = __ldg(&items[blockIdx.x]);
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int2 item = nxt;
            if (w + (int)gridDim.x < n_items) nxt = __ldg(&items[w + gridDim.x]);
            const int I = item.x, J0 = item.y & 0x1fffffff, cnt = ((unsigned)item.y >> 29) + 1;
            for (int sub = 0; 2 * sub < cnt; ++sub, ++sub_it) {
                const uint32_t slot = sub_it & 1u, tph = (sub_it >> 1) & 1u;
                const bool active = col0 < TILE * min(2, cnt - 2 * sub);
                STAMP(0);
                mbar_wait(&tmem_full[slot], tph);
                STAMP(1);
                tc_fence_after();
                STAMP(2);
                uint32_t r[NREG];
                if (active) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 256 + col0;
                    if constexpr (NREG == 64) tmem_ld64<true>(taddr, r); else tmem_ld32<true>(taddr, r);
                    STAMP(3);
                    tmem_wait_ld();
                    STAMP(4);
                }
                tc_fence_before();
                STAMP(5);
                if (lane == 0) mbar_arrive(&tmem_empty[slot]);
                STAMP(6);
                if (active) {
                    uint32_t m[4] = {r[0], r[1], r[2], r[3]};
#pragma unroll
                    for (int k = 4; k + 1 < NREG; k += 2) m[(k >> 1) & 3] = __vimax3_s16x2(m[(k >> 1) & 3], r[k], r[k + 1]);
                    const uint32_t mm = __vmaxs2(__vimax3_s16x2(m[0], m[1], m[2]), m[3]);
                    if ((int)(short)(mm & 0xffffu) >= thr || (int)(short)(mm >> 16) >= thr) {
                        const uint32_t gi = (uint32_t)I * TILE + q * 32 + lane;
#pragma unroll
                        for (int k = 0; k < NREG; ++k)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if ((int)(short)((r[k] >> (16 * h)) & 0xffffu) >= thr) {
                                    const uint32_t gj = (uint32_t)(J0 + 2 * sub) * TILE + col0 + 2 * k + h;
                                    if (gi != gj) {
                                        const unsigned long long pos = atomicAdd(&dbg->n_out, 1ull);
                                        if (pos < out_cap) out[pos] = make_uint2(gi, gj);
                                    }
                                }
                    }
                }
                STAMP(7);
            }
        }
        if (INSTR && blockIdx.x == 0 && lane == 0) { st.v[15] = sub_it; stamps[warp] = st; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------ host
struct Problem {
    int n_tiles, n_items;
    std::vector<uint64_t> sk[2];            // 128-bit sketch per row: sk[0], sk[1]
    std::vector<int2> items;
    int2* d_items = nullptr;
};

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

// planes for K = KB bits: plane c = bits 32 c .. 32 c + 31 of the fold, every row expanded to 32 bytes in the K = 32 layout
static std::vector<unsigned char> make_planes(const Problem& P, int KB, int kind, size_t* plane_stride) {
    const size_t rows = (size_t)P.n_tiles * TILE;
    const size_t stride = ((size_t)P.n_tiles + GROUP) * CHUNK;
    *plane_stride = stride;
    std::vector<unsigned char> buf(stride * (KB / 32), 0);
    const unsigned char plus = kind == 0 ? 0x01 : 0x38, minus = kind == 0 ? 0xff : 0xb8;
    for (size_t r = 0; r < rows; ++r) {
        // fold of the 128-bit sketch to KB bits
        uint64_t lo = P.sk[0][r], hi = P.sk[1][r];
        uint32_t words[4];
        if (KB == 128) { words[0] = (uint32_t)lo; words[1] = (uint32_t)(lo >> 32); words[2] = (uint32_t)hi; words[3] = (uint32_t)(hi >> 32); }
        else if (KB == 64) { const uint64_t f = lo ^ hi; words[0] = (uint32_t)f; words[1] = (uint32_t)(f >> 32); }
        else { const uint64_t f = lo ^ hi; words[0] = (uint32_t)f ^ (uint32_t)(f >> 32); }
        const size_t tile = r / TILE, row = r % TILE;
        for (int c = 0; c < KB / 32; ++c) {
            unsigned char* t = buf.data() + stride * c + tile * CHUNK;
            for (int k = 0; k < 32; ++k)
                t[((row >> 3) * 2 + (k >> 4)) * 128 + (row & 7) * 16 + (k & 15)] = ((words[c] >> k) & 1u) ? minus : plus;
        }
    }
    return buf;
}

static std::vector<uint2> cpu_reference(const Problem& P, int KB, int d) {
    std::vector<uint2> out;
    const size_t rows = (size_t)P.n_tiles * TILE;
    std::vector<uint64_t> f0(rows), f1(rows, 0);
    for (size_t r = 0; r < rows; ++r) {
        const uint64_t lo = P.sk[0][r], hi = P.sk[1][r];
        if (KB == 128) { f0[r] = lo; f1[r] = hi; }
        else if (KB == 64) f0[r] = lo ^ hi;
        else { const uint64_t f = lo ^ hi; f0[r] = (uint32_t)f ^ (uint32_t)(f >> 32); }
    }
    for (const int2& it : P.items) {
        const int I = it.x, J0 = it.y & 0x1fffffff, cnt = ((unsigned)it.y >> 29) + 1;
        for (int a = 0; a < TILE; ++a) {
            const size_t gi = (size_t)I * TILE + a;
            const uint64_t a0 = f0[gi], a1 = f1[gi];
            for (size_t gj = (size_t)J0 * TILE; gj < (size_t)(J0 + cnt) * TILE; ++gj) {
                const int p = __builtin_popcountll(a0 ^ f0[gj]) + __builtin_popcountll(a1 ^ f1[gj]);
                if (p <= d && gi != gj) out.push_back(make_uint2((uint32_t)gi, (uint32_t)gj));
            }
        }
    }
    return out;
}

static bool same_set(std::vector<uint2> a, std::vector<uint2> b) {
    auto lt = [](const uint2& x, const uint2& y) { return x.x != y.x ? x.x < y.x : x.y < y.y; };
    std::sort(a.begin(), a.end(), lt);
    std::sort(b.begin(), b.end(), lt);
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (a[i].x != b[i].x || a[i].y != b[i].y) return false;
    return true;
}

template <int KB, int KIND, int EW, bool SPLIT>
void run_pipeline(const Problem& P, const std::vector<uint2>& want, int d, uint2* d_out, unsigned long long out_cap, Dbg* d_dbg) {
    using C = Cfg<KB>;
    size_t stride;
    std::vector<unsigned char> planes = make_planes(P, KB, KIND, &stride);
    unsigned char* d_planes;
    CK(cudaMalloc(&d_planes, planes.size()));
    CK(cudaMemcpy(d_planes, planes.data(), planes.size(), cudaMemcpyHostToDevice));
    auto kern = k_l1_umma<KB, KIND, EW, SPLIT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    const int arg = KIND == 1 ? (int)f16_bits_of_int(KB - 2 * d) : KB - 2 * d;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e9f;
    Dbg h;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(d_dbg, 0, sizeof(Dbg)));
        CK(cudaEventRecord(e0));
        kern<<<148, (2 + EW) * 32, C::SMEM>>>(d_planes, stride, P.d_items, P.n_items, arg, d_out, out_cap, d_dbg);
        CK(cudaEventRecord(e1));
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { printf("  [pipe] K=%d kind=%d EW=%d split=%d: CUDA error %s\n", KB, KIND, EW, (int)SPLIT, cudaGetErrorString(err)); exit(1); }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CK(cudaMemcpy(&h, d_dbg, sizeof h, cudaMemcpyDeviceToHost));
    std::vector<uint2> got(std::min<unsigned long long>(h.n_out, out_cap));
    CK(cudaMemcpy(got.data(), d_out, got.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
    const bool ok = h.n_out <= out_cap && same_set(got, want);
    long long tp = 0;
    for (const int2& it : P.items) tp += ((unsigned)it.y >> 29) + 1;
    const double pairs = (double)tp * TILE * TILE;
    const double ns = (double)std::max<long long>(h.n_sub, 1);
    printf("  [pipe] K=%3d %s EW=%2d %s: %.3f ms  %6.1f pairs/clk/SM  survivors %llu %s | CTA0: %lld clk total, per tile pair %.0f; "
           "MMA waits: smem_full %.0f tmem_empty %.0f; release->reissue %.0f; epi(w2): wait_full %.0f commit->wake %.0f ld %.0f tree %.0f (per own sub-item)\n",
           KB, KIND == 0 ? "i8/S32 " : "f8/F16p", EW, SPLIT ? "split" : "whole", best, pairs / (best * 1e-3) / (148 * 1.965e9),
           (unsigned long long)h.n_out, ok ? "OK" : "MISMATCH", h.total, h.total / ns, h.mma_wait_full / ns, h.mma_wait_empty / ns,
           h.lat_release_reissue / ns, h.epi_wait_full / (ns / (SPLIT ? 1 : EW / 4)), h.lat_commit_wake / (ns / (SPLIT ? 1 : EW / 4)),
           h.lat_ld / (ns / (SPLIT ? 1 : EW / 4)), h.epi_tree / (ns / (SPLIT ? 1 : EW / 4)));
    CK(cudaFree(d_planes));
}

template <int KB, int EW, bool INSTR>
void run_pipeline2(const Problem& P, const std::vector<uint2>& want, int d, uint2* d_out, unsigned long long out_cap, Dbg* d_dbg) {
    using C = Cfg<KB>;
    size_t stride;
    std::vector<unsigned char> planes = make_planes(P, KB, 1, &stride);
    unsigned char* d_planes;
    CK(cudaMalloc(&d_planes, planes.size()));
    CK(cudaMemcpy(d_planes, planes.data(), planes.size(), cudaMemcpyHostToDevice));
    auto kern = k_l1_umma2<KB, EW, INSTR>;
    Stamps* d_st;
    CK(cudaMalloc(&d_st, 32 * sizeof(Stamps)));
    CK(cudaMemset(d_st, 0, 32 * sizeof(Stamps)));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    const int arg = (int)f16_bits_of_int(KB - 2 * d);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e9f;
    Dbg h;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemset(d_dbg, 0, sizeof(Dbg)));
        CK(cudaEventRecord(e0));
        kern<<<148, (2 + EW) * 32, C::SMEM>>>(d_planes, stride, P.d_items, P.n_items, arg, d_out, out_cap, d_dbg, d_st);
        CK(cudaEventRecord(e1));
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { printf("  [pipe2] K=%d EW=%d: CUDA error %s\n", KB, EW, cudaGetErrorString(err)); exit(1); }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CK(cudaMemcpy(&h, d_dbg, sizeof h, cudaMemcpyDeviceToHost));
    std::vector<uint2> got(std::min<unsigned long long>(h.n_out, out_cap));
    CK(cudaMemcpy(got.data(), d_out, got.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
    const bool ok = h.n_out <= out_cap && same_set(got, want);
    long long tp = 0;
    for (const int2& it : P.items) tp += ((unsigned)it.y >> 29) + 1;
    const double pairs = (double)tp * TILE * TILE;
    printf("  [pipe2] K=%3d f8/F16p 2x256 EW=%2d%s: %.3f ms  %6.1f pairs/clk/SM  survivors %llu %s\n", KB, EW, INSTR ? " instr" : "", best,
           pairs / (best * 1e-3) / (148 * 1.965e9), (unsigned long long)h.n_out, ok ? "OK" : "MISMATCH");
    if (INSTR) {
        Stamps hs[32];
        CK(cudaMemcpy(hs, d_st, sizeof hs, cudaMemcpyDeviceToHost));
        const double n = (double)std::max<long long>(hs[0].v[15], 1);
        printf("    CTA 0, cycles per sub-item (%.0f sub-items). MMA thread: item prologue %.0f | wait smem_full %.0f | sub prologue %.0f | wait tmem_empty %.0f | "
               "fence %.0f | mma issue %.0f | commit %.0f | stage commit %.0f\n", n, hs[0].v[0] / n, hs[0].v[1] / n, hs[0].v[2] / n, hs[0].v[3] / n, hs[0].v[4] / n,
               hs[0].v[5] / n, hs[0].v[6] / n, hs[0].v[7] / n);
        for (int w : {2, 3, 9, 17}) {
            if (w >= 2 + EW) continue;
            printf("    epilogue warp %2d: loop head %.0f | wait tmem_full %.0f | fence_after %.0f | ld issue %.0f | wait::ld %.0f | fence_before %.0f | arrive %.0f | tree+test %.0f\n",
                   w, hs[w].v[0] / n, hs[w].v[1] / n, hs[w].v[2] / n, hs[w].v[3] / n, hs[w].v[4] / n, hs[w].v[5] / n, hs[w].v[6] / n, hs[w].v[7] / n);
        }
    }
    CK(cudaFree(d_st));
    CK(cudaFree(d_planes));
}

int main(int argc, char** argv) {
    const int part = argc > 1 ? atoi(argv[1]) : 0;   // 0 = all
    uint32_t* d_out32;
    long long* d_cyc;
    CK(cudaMalloc(&d_out32, 4));
    CK(cudaMalloc(&d_cyc, 148 * 8));

    if (part == 0 || part == 1) {
        printf("part 1: TMEM -> registers epilogue rates (registers of 32 bits per clock per SM; a packed register holds two values)\n");
        run_epi_rate<0>("x64 load only", 64, d_out32, d_cyc);
        run_epi_rate<1>("x64 load + s32 max3 tree", 64, d_out32, d_cyc);
        run_epi_rate<2>("x64.pack::16b load + s16x2 max3 tree", 64, d_out32, d_cyc);
        run_epi_rate<3>("x64 load + IMAD pack + s16x2 max3 tree (64 s32 values)", 64, d_out32, d_cyc);
        run_epi_rate<4>("2 x x64 loads in flight + s32 max3 tree", 128, d_out32, d_cyc);
        run_epi_rate<5>("x32.pack::16b load + s16x2 max3 tree", 32, d_out32, d_cyc);
        run_epi_rate<6>("2 x x64.pack::16b loads in flight + s16x2 max3 tree", 128, d_out32, d_cyc);
    }

    if (part == 0 || part == 2) {
        printf("part 2: accumulator layout probe (M=128 N=256 K=32)\n");
        uint32_t hA[128], hB[256];
        for (auto& x : hA) x = (uint32_t)rnd();
        for (auto& x : hB) x = (uint32_t)rnd();
        hB[5] = hA[3];
        hB[6] = hA[3] ^ 0x10;
        uint32_t *dA, *dB, *dD, *dDp;
        CK(cudaMalloc(&dA, sizeof hA)); CK(cudaMalloc(&dB, sizeof hB)); CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMalloc(&dDp, 128 * 256 * 4));
        CK(cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice));
        static uint32_t hD[128 * 256], hDp[128 * 256];
        for (int kind = 0; kind < 2; ++kind) {
            CK(cudaMemset(dD, 0x7f, 128 * 256 * 4));
            if (kind == 0) k_probe<0><<<1, 128>>>(dA, dB, dD, dDp); else k_probe<1><<<1, 128>>>(dA, dB, dD, dDp);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  probe kind %d: CUDA error %s\n", kind, cudaGetErrorString(e)); return 1; }
            CK(cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hDp, dDp, sizeof hDp, cudaMemcpyDeviceToHost));
            int bad_s32 = 0, bad_h1 = 0, bad_h2 = 0, bad_p1 = 0, bad_p2 = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 256; ++j) {
                    const int want = 32 - 2 * __builtin_popcount(hA[i] ^ hB[j]);
                    const uint32_t cell = hD[i * 256 + j];
                    if ((int)cell != want) ++bad_s32;                                     // S32 per column
                    if ((cell & 0xffff) != f16_bits_of_int(want)) ++bad_h1;               // H1: f16 in the low half of column j
                    if (j < 128) {                                                        // H2: column j holds D[i][2j], D[i][2j+1]
                        const int w0 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[2 * j]), w1 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[2 * j + 1]);
                        if (cell != (uint32_t)(f16_bits_of_int(w0) | ((uint32_t)f16_bits_of_int(w1) << 16))) ++bad_h2;
                    }
                    // packed loads: P1: register k of the x16 load at column c holds columns (c + 2k, c + 2k + 1) -> covers 32 columns;
                    //               P2: register k holds the low halves of columns (c + k) and (c + k + 16)
                    const int c = j & ~15, k = j & 15;
                    const uint32_t pr = hDp[i * 256 + j];
                    if (c + 2 * k + 1 < 256) {
                        const int w0 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[c + 2 * k]), w1 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[c + 2 * k + 1]);
                        if (pr != (uint32_t)(f16_bits_of_int(w0) | ((uint32_t)f16_bits_of_int(w1) << 16))) ++bad_p1;
                    }
                    if (c + k + 16 < 256) {
                        const int w0 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[c + k]), w1 = 32 - 2 * __builtin_popcount(hA[i] ^ hB[c + k + 16]);
                        if (pr != (uint32_t)(f16_bits_of_int(w0) | ((uint32_t)f16_bits_of_int(w1) << 16))) ++bad_p2;
                    }
                }
            printf("  kind %s: mismatches  S32-per-column %d | f16-low-half-per-column (H1) %d | two-f16-per-column (H2) %d | packed load: adjacent columns (P1) %d, "
                   "k and k+16 (P2) %d\n", kind == 0 ? "i8 " : "f8 ", bad_s32, bad_h1, bad_h2, bad_p1, bad_p2);
            printf("    row 3: cells 4..7 = %08x %08x %08x %08x ; packed regs (col 0 load) 0..3 = %08x %08x %08x %08x ; expected f16(32)=%04x f16(30)=%04x\n",
                   hD[3 * 256 + 4], hD[3 * 256 + 5], hD[3 * 256 + 6], hD[3 * 256 + 7], hDp[3 * 256 + 0], hDp[3 * 256 + 1], hDp[3 * 256 + 2], hDp[3 * 256 + 3],
                   f16_bits_of_int(32), f16_bits_of_int(30));
        }
    }

    if (part == 0 || part == 3 || part == 4) {
        printf("part 3/4: pipelined level 1 on synthetic folds\n");
        Problem P;
        P.n_tiles = 2048;
        const size_t rows = (size_t)P.n_tiles * TILE;
        P.sk[0].resize(rows);
        P.sk[1].resize(rows);
        for (size_t r = 0; r < rows; ++r) { P.sk[0][r] = rnd(); P.sk[1][r] = rnd(); }
        const int per_tile = argc > 2 ? atoi(argv[2]) : 12;   // items per row tile (4 column tiles each, some partly filled)
        for (int I = 0; I < P.n_tiles; ++I) {
            int J = (int)(((long long)I * 7 + 3) % (P.n_tiles - 4 * per_tile - 4));
            for (int k = 0; k < per_tile; ++k) {
                const int cnt = (k % 5 == 4) ? 1 + (int)(rnd() % 3) : 4;
                P.items.push_back(make_int2(I, J | ((cnt - 1) << 29)));
                // plant near-duplicates: a few rows of tile J become copies of rows of tile I with 0..2 flipped bits
                for (int t = 0; t < 3; ++t) {
                    const size_t src = (size_t)I * TILE + rnd() % TILE, dst = (size_t)(J + (int)(rnd() % cnt)) * TILE + rnd() % TILE;
                    if (dst / TILE == (size_t)I) continue;
                    uint64_t lo = P.sk[0][src], hi = P.sk[1][src];
                    const int flips = (int)(rnd() % 3);
                    for (int f = 0; f < flips; ++f) { const int b = (int)(rnd() % 128); if (b < 64) lo ^= 1ull << b; else hi ^= 1ull << (b - 64); }
                    P.sk[0][dst] = lo;
                    P.sk[1][dst] = hi;
                }
                J += cnt;
            }
        }
        // interleave the items so that neighbouring CTAs do not all start on the same row tile
        P.n_items = (int)P.items.size();
        CK(cudaMalloc(&P.d_items, P.items.size() * sizeof(int2)));
        CK(cudaMemcpy(P.d_items, P.items.data(), P.items.size() * sizeof(int2), cudaMemcpyHostToDevice));
        const unsigned long long out_cap = 1ull << 22;
        uint2* d_out;
        Dbg* d_dbg;
        CK(cudaMalloc(&d_out, out_cap * sizeof(uint2)));
        CK(cudaMalloc(&d_dbg, sizeof(Dbg)));
        const int d = 1;
        for (int KB : {32, 64, 128}) {
            const std::vector<uint2> want = cpu_reference(P, KB, d);
            printf("  K = %d: %zu survivors expected (d = %d), %d items\n", KB, want.size(), d, P.n_items);
            if (part == 4) {
                if (KB == 32) { run_pipeline2<32, 16, false>(P, want, d, d_out, out_cap, d_dbg); run_pipeline2<32, 16, true>(P, want, d, d_out, out_cap, d_dbg); }
                else if (KB == 64) { run_pipeline2<64, 8, true>(P, want, d, d_out, out_cap, d_dbg); run_pipeline2<64, 16, true>(P, want, d, d_out, out_cap, d_dbg); }
                continue;
            }
            if (KB == 32) {
                run_pipeline<32, 0, 4, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 0, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 0, 8, false>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 0, 16, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 1, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 1, 8, false>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<32, 1, 16, true>(P, want, d, d_out, out_cap, d_dbg);
            } else if (KB == 64) {
                run_pipeline<64, 0, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 0, 8, false>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 0, 16, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 1, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 1, 8, false>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 1, 16, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<64, 1, 16, false>(P, want, d, d_out, out_cap, d_dbg);
            } else {
                run_pipeline<128, 0, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<128, 0, 16, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<128, 1, 8, true>(P, want, d, d_out, out_cap, d_dbg);
                run_pipeline<128, 1, 16, true>(P, want, d, d_out, out_cap, d_dbg);
            }
        }
    }
    return 0;
}
