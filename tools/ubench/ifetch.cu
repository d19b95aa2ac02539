// Micro-benchmark: does a loop body that does not fit the instruction caches cap the issue rate?
// Straight-line bodies of K scalar FFMAs (8 independent accumulators, issue-bound at 1 instruction / cycle / SM sub-partition)
// in a loop of TOTAL / K trips; 1, 2 and 3 warps per sub-partition.  "lockstep": the warps of a sub-partition run the body
// together (they share every fetched line); "skewed": warp w of a sub-partition starts w/3 of a body later, the way the frame
// streams of the render kernels sit at different places of their loop body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch ifetch.cu && ./ifetch
#include <cstdio>
#include <cuda_runtime.h>
#define TOTAL (3 << 20)

template <int K>
__global__ void __launch_bounds__(512, 1) body(float *out, long long *cyc, float a, float b, int skew)
{
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    if (skew) { const long long c0 = clock64(), d = (long long)(threadIdx.x >> 7) * skew; while (clock64() - c0 < d) { } }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < TOTAL / K; it++) {
#pragma unroll
        for (int i = 0; i < K; i++) x[i & 7] = fmaf(x[i & 7], a, b);
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int K> void run(int threads, bool skewed)
{
    float *out; long long *cyc, h[16];
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    const int skew = skewed ? K / 3 * 2 : 0;
    body<K><<<148, threads>>>(out, cyc, 0.999f, 0.25f, skew);
    body<K><<<148, threads>>>(out, cyc, 0.999f, 0.25f, skew);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < threads / 32; w++) mx = h[w] > mx ? h[w] : mx;
    const double inst = (double)(TOTAL / K) * K * (threads / 128);       // warp instructions per sub-partition
    printf("body %6d instr (%4d KB)  warps/SMSP %d  %-8s  IPC/SMSP %.3f  (%s)\n", K, K * 16 / 1024, threads / 128, skewed ? "skewed" : "lockstep", inst / mx, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int th : { 128, 256, 384 })
        for (int sk = 0; sk < (th > 128 ? 2 : 1); sk++) {
            run<1536>(th, sk); run<2048>(th, sk); run<2304>(th, sk); run<2560>(th, sk); run<2816>(th, sk); run<3072>(th, sk); run<3584>(th, sk); run<4096>(th, sk); run<6144>(th, sk);
        }
    return 0;
}
