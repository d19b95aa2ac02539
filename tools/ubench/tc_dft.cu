// Micro-benchmark: how fast is the tcgen05 tensor unit on the GEMM shapes a DFT pass would give it?
//
// A radix-16 pass of the fp32 FFT as a GEMM is a real 32 x 32 matrix (the complex DFT-16 as [re -im; im re]) applied to
// 32 x (columns) data: 128 flop per sample and pass, three passes for N = 4096, and fp32-grade accuracy needs the 3 x TF32 split
// (hi*hi + hi*lo + lo*hi): 1 152 tensor-flop per sample, 566 TFLOP/s at the 60 % roofline target of config 2.
// Two ways to map it onto tcgen05.mma (kind::tf32, K = 8 per instruction, operands in shared memory, accumulator in TMEM):
//   "data as A":  M = 128 data columns, N = 32, K = 32   (every flop useful; A = 4 KB of shared memory per instruction)
//   "DFT as A":   M = 64 (the smallest M; 32 rows used),  N = 256 data columns, K = 32   (half of the flops wasted)
// This program issues those instruction shapes back to back from one thread per SM (operands resident in shared memory, no data
// movement at all, results not read) and reports the rate - an UPPER bound for a DFT pass built on them.  bf16 M = 128, N = 256
// is run as a calibration point against the measured cuBLAS peak (MEASURED_PEAKS.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_dft tc_dft.cu && ./tc_dft
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    // K-major, no swizzle: 8-row x 16-byte core matrices; lbo = byte step between core matrices along K, sbo = along M / N
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

template <int TF32>
__global__ void __launch_bounds__(128, 1) mma_rate(int M, int N, int iters, long long *cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    // operands: [rows / 8][8 K-chunks][128-byte core matrix]; any finite bit pattern will do
    for (int i = threadIdx.x; i < (256 + 256) * 32; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = TF32 ? 0x3f800000u : 0x3f803f80u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 256 * 128;            // A: up to 256 rows x 32 K x 4 B = 32 KB, B likewise
        const uint32_t idesc = (1u << 4) | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {                                   // K = 32 (tf32) / 64 (bf16) in four instructions
                const uint64_t da = make_desc(a0 + k * 256, 128, 1024), db = make_desc(b0 + k * 256, 128, 1024);
                const uint32_t acc = (it | k) ? 1u : 0u;
                if (TF32)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        cyc[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int TF32> void run(const char *what, int M, int N, int useful_rows)
{
    long long *cyc;
    cudaMalloc(&cyc, 148 * 8);
    auto k = mma_rate<TF32>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 128, 65536>>>(M, N, 200, cyc);
    cudaEventRecord(e0);
    k<<<148, 128, 65536>>>(M, N, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const int K = TF32 ? 8 : 16;
    const double flop = 2.0 * M * N * K * 4.0 * iters * 148;
    printf("%-34s M=%3d N=%3d K=%2d x4 : %7.1f TFLOP/s issued, %7.1f useful (%d of %d rows), %6.1f cycles per instruction  (%s)\n", what, M, N, K,
           flop / (ms * 1e-3) / 1e12, flop / (ms * 1e-3) / 1e12 * useful_rows / M, useful_rows, M, (double)h[0] / (4.0 * iters), cudaGetErrorString(err));
    cudaFree(cyc);
}
int main()
{
    run<0>("bf16 calibration", 128, 256, 128);
    run<1>("tf32 large tile", 128, 256, 128);
    run<1>("tf32 data as A (DFT-16 pass)", 128, 32, 128);
    run<1>("tf32 data as A, two passes' worth", 128, 64, 128);
    run<1>("tf32 DFT matrix as A", 64, 256, 32);
    run<1>("tf32 DFT matrix as A", 64, 128, 32);
    return 0;
}
