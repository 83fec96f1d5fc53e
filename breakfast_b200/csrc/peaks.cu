// peaks.cu — register-resident integer-pipe microbenchmarks (roofline denominators for K3).
// MEASURED_PEAKS.json only records HBM and bf16 tensor peaks; the pair kernel is bound by the
// POPC / LOP3 / IADD pipes, so their lane-op rates are measured on the box with these kernels.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/breakfast_b200.h"

namespace {

constexpr int kIters = 2048;
constexpr int kChains = 16;

// mode 0: POPC only, 1: LOP3 (xor) only, 2: IADD3 only, 3: xor+popc+add (the pair kernel's inner op)
template <int MODE>
__global__ void __launch_bounds__(256) k_peak(uint32_t seed, uint32_t* out) {
    uint32_t x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = seed * (threadIdx.x + 1) + c * 0x9E3779B9u + blockIdx.x;
    uint32_t y = seed ^ threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (MODE == 0) {
                asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            } else if (MODE == 1) {
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            } else if (MODE == 2) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            } else {
                uint32_t t;
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(t) : "r"(x[c]), "r"(y + c));
                asm volatile("popc.b32 %0, %0;" : "+r"(t));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(t));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 0x12345678u) out[0] = s;  // keep the chains alive
}

// mode 4: mma.sync.m16n8k32 s8 (IMMA.16832.S8), 8 independent accumulators per warp; counted in int8 MACs
__global__ void __launch_bounds__(256) k_peak_imma(uint32_t seed, uint32_t* out) {
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0;
    const uint32_t a0 = seed * (threadIdx.x + 1) * 0x01010101u, a1 = a0 ^ 0xff00ff00u, a2 = a0 + 0x01000100u, a3 = ~a0;
    const uint32_t b0 = blockIdx.x * 0x01010101u + 1, b1 = ~b0;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + i), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 0x12345678) out[0] = (uint32_t)s;
}

int run_peak_imma(double* gmacs) {
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BF_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return BF_ERR_OOM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_peak_imma<<<blocks, 256>>>(rep + 1, d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || best > 1e29f) return BF_ERR_CUDA;
    const double macs = (double)blocks * 8.0 * kIters * 8.0 * (16 * 8 * 32);  // warps x iters x 8 MMAs x MACs per MMA
    *gmacs = macs / (best * 1e-3) / 1e9;
    return BF_OK;
}

// mode 5: tcgen05.mma kind::i8 (SASS UTCIMMA), M = 128, N = 256, K = 32 per instruction, operands in shared memory
// (K-major, no swizzle), accumulators in TMEM; one CTA per SM, one thread issues kUmmaIters MMAs alternating between
// two accumulators, then commits and waits.  The Blackwell-native int8 tensor rate (counted in MACs), which the
// level-1 kernel is NOT bound by: its accumulators live in registers (mma.sync), see DESIGN.md section 3.
constexpr int kUmmaIters = 4096;
__global__ void __launch_bounds__(128, 1) k_peak_umma_i8(uint32_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];   // A 4 KB, B 8 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 12288 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01ff01ffu;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
        auto desc = [](uint32_t a) { return (uint64_t)((a & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46); };
        const uint64_t da = desc(sa), db = desc(sa + 4096);
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll 1
        for (int it = 0; it < kUmmaIters; ++it)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(
                             tmem + (uint32_t)((it & 1) * 256)), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar_a) : "memory");
    }
    if (warp == 0) {   // the whole warp: tcgen05.ld is .sync.aligned
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tmem) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r0 == 0x12345678u) out[0] = r0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int run_peak_umma_i8(double* gmacs) {
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BF_ERR_CUDA;
    if (prop.major < 10) return BF_ERR_NO_DEVICE;
    const int blocks = prop.multiProcessorCount;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return BF_ERR_OOM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_peak_umma_i8<<<blocks, 128, 12288>>>(d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || best > 1e29f) return BF_ERR_CUDA;
    *gmacs = (double)blocks * kUmmaIters * (128.0 * 256.0 * 32.0) / (best * 1e-3) / 1e9;
    return BF_OK;
}

template <int MODE>
int run_peak(double* gops) {
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BF_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return BF_ERR_OOM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_peak<MODE><<<blocks, 256>>>(rep + 1, d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || best > 1e29f) return BF_ERR_CUDA;
    const double ops = (double)blocks * 256.0 * kIters * kChains;  // counted per lane, per primary op
    *gops = ops / (best * 1e-3) / 1e9;
    return BF_OK;
}

}  // namespace

extern "C" int bf_measure_peak(int32_t device, const char* name, double* gops_out) {
    if (!name || !gops_out) return BF_ERR_INVALID;
    *gops_out = 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return BF_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return BF_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return BF_ERR_CUDA;
    if (!strcmp(name, "popc32")) return run_peak<0>(gops_out);
    if (!strcmp(name, "lop3")) return run_peak<1>(gops_out);
    if (!strcmp(name, "iadd3")) return run_peak<2>(gops_out);
    if (!strcmp(name, "xor_popc_add")) return run_peak<3>(gops_out);
    if (!strcmp(name, "imma_s8")) return run_peak_imma(gops_out);
    if (!strcmp(name, "umma_i8")) return run_peak_umma_i8(gops_out);
    return BF_ERR_INVALID;
}
