// Standalone check of the tcgen05 kind::i8 building block planned for the level-1 filter:
// D[128][256] (int32, TMEM) = A[128][32] (int8, +-1) x B[256][32]^T (int8, +-1), K-major operands in the
// no-swizzle canonical layout (core matrix = 8 rows x 16 bytes; LBO = 128 B between the two K halves,
// SBO = 256 B between 8-row groups).  Compares with a CPU result.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_test umma_i8_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // start address >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), layout none [61,64)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) k_test(const uint32_t* foldA, const uint32_t* foldB, int32_t* D) {
    __shared__ __align__(128) uint4 sA[128 * 2];   // 4 KB
    __shared__ __align__(128) uint4 sB[256 * 2];   // 8 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // expand folds to +-1 bytes: element (row r, k) at ((r/8)*2 + k/16)*128 + (r%8)*16 + k%16
    auto expand = [](uint32_t bits16) {
        uint4 v;
        uint32_t w[4];
        for (int q = 0; q < 4; ++q) {
            uint32_t nib = (bits16 >> (4 * q)) & 0xF;
            uint32_t spread = (nib * 0x00204081u) & 0x01010101u;
            w[q] = 0x01010101u ^ (spread * 0xFEu);
        }
        v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
        return v;
    };
    for (int r = tid; r < 128; r += 128) {
        uint32_t f = foldA[r];
        sA[((r >> 3) * 2 + 0) * 8 + (r & 7)] = expand(f & 0xFFFF);
        sA[((r >> 3) * 2 + 1) * 8 + (r & 7)] = expand(f >> 16);
    }
    for (int r = tid; r < 256; r += 128) {
        uint32_t f = foldB[r];
        sB[((r >> 3) * 2 + 0) * 8 + (r & 7)] = expand(f & 0xFFFF);
        sB[((r >> 3) * 2 + 1) * 8 + (r & 7)] = expand(f >> 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");   // generic-proxy smem writes -> visible to the async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
        // idesc: c=S32 (2<<4), a=int8 signed (1<<7), b=int8 signed (1<<10), K-major both, N=256 (32<<17), M=128 (8<<24)
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
            "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // wait for the MMA
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(&bar)));
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes x 256 columns
    for (int c = 0; c < 256; c += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int k = 0; k < 16; ++k) D[tid * 256 + c + k] = (int32_t)r[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
    uint32_t hA[128], hB[256];
    srand(7);
    for (auto& x : hA) x = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    for (auto& x : hB) x = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    hB[5] = hA[3];            // identical folds -> 32
    hB[6] = hA[3] ^ 0x10;     // one bit apart -> 30
    uint32_t *dA, *dB;
    int32_t* dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * 256 * 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0x7f, 128 * 256 * 4);
    k_test<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static int32_t hD[128 * 256];
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 256; ++j) {
            int want = 32 - 2 * __builtin_popcount(hA[i] ^ hB[j]);
            if (hD[i * 256 + j] != want) { if (bad < 10) printf("mismatch D[%d][%d] = %d want %d\n", i, j, hD[i * 256 + j], want); ++bad; }
        }
    printf("D[3][5]=%d D[3][6]=%d D[0][0]=%d ; mismatches: %d of %d\n", hD[3 * 256 + 5], hD[3 * 256 + 6], hD[0], bad, 128 * 256);
    return bad ? 2 : 0;
}
