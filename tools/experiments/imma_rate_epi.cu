// throughput of mma.sync.aligned.m16n8k32.s32.s8.s8.s32 (SASS IMMA.16832.S8) on sm_100a, register resident
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(int iters, int* out) {
    int c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0xff00ff00u, a2 = a0 + 0x01000100u, a3 = ~a0;
    uint32_t b0 = blockIdx.x * 0x01010101u + 1, b1 = ~b0;
    int acc_m = -64;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                         : "=r"(c[i][0]), "=r"(c[i][1]), "=r"(c[i][2]), "=r"(c[i][3])
                         : "r"(a0 + i), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + it), "r"(b1), "r"(0));
        int m = -64;
#pragma unroll
        for (int i = 0; i < 8; ++i) { m = max(m, max(c[i][0], c[i][1])); m = max(m, max(c[i][2], c[i][3])); }
        acc_m = max(acc_m, m);
    }
    int s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s + acc_m == 0x12345678) out[0] = s;
}
int main() {
    int* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<blocks, 256>>>(iters, d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error\n"); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = (double)blocks * 8 * iters * 8;   // warps * iters * 8 per iter
        double macs = mmas * 16 * 8 * 32;
        printf("rep %d: %.3f ms  %.1f IMMA/clk/SM  %.0f MAC/clk/SM  (%.2f Pop/s)  = %.1f pairs(K=32)/clk/SM\n", rep, ms,
               mmas / (ms * 1e-3) / 1.965e9 / 148, macs / (ms * 1e-3) / 1.965e9 / 148, 2 * macs / (ms * 1e-3) / 1e15,
               mmas * 128 / (ms * 1e-3) / 1.965e9 / 148);
    }
}
