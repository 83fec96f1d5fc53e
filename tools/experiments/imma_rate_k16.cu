// throughput of mma.sync.aligned.m16n8k16.s32.s8.s8.s32 (SASS IMMA.16816.S8) on sm_100a, register resident, with
// and without the level-1 epilogue (fresh accumulators + 3-input max tree) - would a 16-bit fold double level 1?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <bool EPI>
__global__ void __launch_bounds__(256) k(int iters, int* out) {
    uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0xff00ff00u;
    uint32_t b0 = blockIdx.x * 0x01010101u + 1;
    int acc_m = -1000;
    int c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
        if (EPI) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                             : "=r"(c[i][0]), "=r"(c[i][1]), "=r"(c[i][2]), "=r"(c[i][3])
                             : "r"(a0 + it), "r"(a1), "r"(b0 + i), "r"(0));
            int m0 = max(max(c[0][0], c[0][1]), max(c[0][2], c[0][3]));
            int m1 = max(max(c[1][0], c[1][1]), max(c[1][2], c[1][3]));
#pragma unroll
            for (int i = 2; i < 8; i += 2) {
                m0 = max(m0, max(c[i][0], c[i][1])); m1 = max(m1, max(c[i][2], c[i][3]));
                m0 = max(m0, max(c[i + 1][0], c[i + 1][1])); m1 = max(m1, max(c[i + 1][2], c[i + 1][3]));
            }
            acc_m = max(acc_m, max(m0, m1));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(b0 + i));
        }
    }
    int s = acc_m;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 0x12345678) out[0] = s;
}
template <bool EPI>
void run(const char* name) {
    int* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<EPI><<<blocks, 256>>>(iters, d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error\n"); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = (double)blocks * 8 * iters * 8;
        printf("%s rep %d: %.3f ms  %.3f IMMA.16816/clk/SM = %.1f pairs(K=16)/clk/SM\n", name, rep, ms,
               mmas / (ms * 1e-3) / 1.965e9 / 148, mmas * 128 / (ms * 1e-3) / 1.965e9 / 148);
    }
    cudaFree(d);
}
int main() { run<false>("pure"); run<true>("with max tree"); }
