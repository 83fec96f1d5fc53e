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
    return BF_ERR_INVALID;
}
