// Round-2 experiment: where do the cycles of a tcgen05 accumulator round trip go?
//   A: issue n MMAs (M=128, N, K=32 each) + commit, then wait on the mbarrier: latency as a function of n and N
//   B: the same while 8 other warps hammer tcgen05.ld on other TMEM columns
//   C: mbarrier hand-off latency between two warps of one CTA: arrive at a stamped time, waiter already parked in
//      try_wait (all lanes), try_wait (one lane), or spinning on test_wait
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_latency umma_latency.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define O8(r, b) "=r"(r[b + 0]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3]), "=r"(r[b + 4]), "=r"(r[b + 5]), "=r"(r[b + 6]), "=r"(r[b + 7])
#define L32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

// mode 0: A (quiet), 1: B (8 warps of tcgen05.ld traffic)
__global__ void __launch_bounds__(320, 1) k_mma_latency(int n_mma, int N, int iters, int traffic, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];   // A 4 KB, B 8 KB (contents irrelevant for timing)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 12288 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t da = umma_desc(smem_u32(smem)), db = umma_desc(smem_u32(smem + 4096));
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            long long t_issue = 0;
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const long long a = clock64();
                for (int m = 0; m < n_mma; ++m) umma_i8(tmem + (uint32_t)((m * N) & 255), da, db, idesc, 0u);
                umma_commit(&bar);
                t_issue += clock64() - a;
                mbar_wait(&bar, it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const long long t1 = clock64();
            out[0] = (t1 - t0) / iters;
            out[1] = t_issue / iters;
            stop = 1;
        }
    } else if (warp >= 2 && traffic) {
        // tcgen05.ld traffic on columns 256.. (the MMAs write columns 0..255)
        uint32_t acc = 0;
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 1) * 128;
        while (!stop) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : O8(r, 0), O8(r, 8), O8(r, 16), O8(r, 24) : "r"(base) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc |= r[0] ^ r[31];
        }
        if (acc == 0x12345u) sink[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// C: warp 1 arrives (lane 0) at a stamped time, warp 0 waits.  mode 0: try_wait all lanes, 1: try_wait lane 0 + syncwarp, 2: test_wait spin lane 0
__global__ void __launch_bounds__(64, 1) k_handoff(int mode, int iters, int delay, long long* out) {
    __shared__ __align__(8) uint64_t bar_go, bar_back;
    __shared__ volatile long long t_arrive;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar_go, 1); mbar_init(&bar_back, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long sum = 0;
    for (int it = 0; it < iters; ++it) {
        if (warp == 0) {
            if (mode == 0) mbar_wait(&bar_go, it & 1);
            else if (mode == 1) { if (lane == 0) mbar_wait(&bar_go, it & 1); __syncwarp(); }
            else { if (lane == 0) mbar_spin(&bar_go, it & 1); __syncwarp(); }
            const long long t = clock64();
            sum += t - t_arrive;
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_back);
        } else {
            if (lane == 0) {
                const long long t0 = clock64();
                while (clock64() - t0 < delay) { }     // let the waiter park first
                t_arrive = clock64();
                mbar_arrive(&bar_go);
                mbar_wait(&bar_back, it & 1);
            }
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) out[0] = sum / iters;
}

int main() {
    long long* d_out;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMalloc(&d_sink, 4));
    long long h[2];
    printf("A/B: cycles from the first tcgen05.mma to the mbarrier flip seen by the issuing thread (kind::i8, M=128, K=32 per MMA)\n");
    for (int traffic = 0; traffic < 2; ++traffic)
        for (int N : {128, 256})
            for (int n_mma : {1, 2, 4, 8, 16}) {
                k_mma_latency<<<1, 320, 12288>>>(n_mma, N, 2000, traffic, d_out, d_sink);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
                printf("  %s N=%3d  %2d MMAs + commit: %5lld cycles per round trip (issue part %4lld)\n", traffic ? "with ld traffic" : "quiet          ", N, n_mma, h[0], h[1]);
            }
    printf("C: mbarrier hand-off latency, arrive -> waiter past the wait\n");
    for (int delay : {0, 2000})
        for (int mode = 0; mode < 3; ++mode) {
            k_handoff<<<1, 64>>>(mode, 2000, delay, d_out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
            printf("  waiter %s, parked %4d cycles before the arrive: %lld cycles\n",
                   mode == 0 ? "try_wait, all lanes     " : mode == 1 ? "try_wait, lane 0        " : "test_wait spin, lane 0  ", delay, h[0]);
        }
    return 0;
}
