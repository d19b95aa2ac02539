// pipes.cu — issue / pipe throughput microbenchmark for the instructions the render kernel is built
// from (sm_100a): scalar vs packed fp32 (FADD2/FMUL2/FFMA2), conversions, MUFU, shared-memory
// loads/stores/atomics, and mixes.  Prints warp-instructions per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define ITERS 4096
#define UNROLL 8

enum Op { FADD1, FFMA1, FMUL1, FADD2, FFMA2, FMUL2, FFMA2_SW, MIX_F2_ALU, MIX_F2_LDS, MIX_F1_LDS, LDS64, LDS128, STS64, I2F16, F2I, MUFU, ATOMS_SPREAD, ATOMS_SAME, PRMT, MIX_F2_F1, FFMA1_2SRC, OPS };
const char *names[] = {"FADD", "FFMA(3 src)", "FMUL", "FADD2", "FFMA2", "FMUL2", "FFMA2 swizzled(cmul)", "FADD2+IADD3 1:1", "FADD2+LDS.64 4:1", "FADD+LDS.64 4:1", "LDS.64", "LDS.128", "STS.64", "I2F.S16", "F2I", "MUFU.LG2", "ATOMS spread", "ATOMS same-addr-per-warp", "PRMT", "FADD2+FADD 1:1", "FFMA(2 distinct src)"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float *out, long long *cyc, int seed)
{
    __shared__ __align__(16) float sm[8192];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-3f;
    __syncthreads();
    float a[UNROLL], b[UNROLL];
    u64 A[UNROLL];
    int ia[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; j++) {
        a[j] = 1.0f + (float)(tid + j) * 1e-6f; b[j] = 0.5f + (float)j * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(A[j]) : "f"(a[j]), "f"(b[j]));
        ia[j] = tid * 17 + j + seed;
    }
    const float c = 0.999f + (float)seed * 1e-9f, d = 1e-7f;
    u64 C, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c), "f"(c));
    asm("mov.b64 %0, {%1, %2};" : "=l"(D) : "f"(d), "f"(d));
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (tid & 1023) * 16;
    unsigned aaddr_spread = (unsigned)__cvta_generic_to_shared(sm) + ((tid * 37 + seed) & 2047) * 4;
    unsigned aaddr_same = (unsigned)__cvta_generic_to_shared(sm) + ((tid >> 5) & 31) * 4;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            if (OP == FADD1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(d));
            if (OP == FFMA1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(c), "f"(b[j]));
            if (OP == FFMA1_2SRC) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[j]) : "f"(c));
            if (OP == FMUL1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(c));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(D));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[j]) : "l"(C), "l"(D));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(C));
            if (OP == FFMA2_SW) {   // complex multiply by (c, d): 2 packed ops with broadcast / swapped operands
                asm volatile("{\n\t.reg .f32 lo, hi, nd; .reg .b64 sw, t, w1, w2;\n\t"
                             "mov.b64 {lo, hi}, %0;\n\t mov.b64 sw, {hi, lo};\n\t neg.f32 nd, %2;\n\t"
                             "mov.b64 w1, {%1, %1};\n\t mov.b64 w2, {nd, %2};\n\t"
                             "mul.rn.f32x2 t, %0, w1;\n\t fma.rn.f32x2 %0, sw, w2, t;\n\t}" : "+l"(A[j]) : "f"(c), "f"(d));
            }
            if (OP == MIX_F2_ALU) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(D));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(seed));
            }
            if (OP == MIX_F2_F1) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(D));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(d));
            }
            if (OP == MIX_F2_LDS) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(D));
                if ((j & 3) == 0) { u64 t; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(saddr + 8 * (j >> 2))); A[(j + 5) & 7] ^= (t & 1); }
            }
            if (OP == MIX_F1_LDS) {
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(d));
                if ((j & 3) == 0) { u64 t; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(saddr + 8 * (j >> 2))); A[(j + 5) & 7] ^= (t & 1); }
            }
            if (OP == LDS64) { u64 t; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(saddr + 8 * (j & 1))); A[j] ^= t; }
            if (OP == LDS128) { u64 t, u; asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t), "=l"(u) : "r"(saddr)); A[j] ^= t + u; }
            if (OP == STS64) asm volatile("st.shared.b64 [%0], %1;" ::"r"(saddr + 8 * (j & 1)), "l"(A[j]) : "memory");
            if (OP == I2F16) asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(a[j]) : "h"((short)(ia[j] + it)));
            if (OP == F2I) asm volatile("cvt.rzi.sat.u32.f32 %0, %1;" : "=r"(ia[j]) : "f"(a[j]));
            if (OP == MUFU) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (OP == ATOMS_SPREAD) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(aaddr_spread + 128 * j) : "memory");
            if (OP == ATOMS_SAME) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(aaddr_same + 256 * j) : "memory");
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7610;" : "+r"(ia[j]) : "r"(seed));
        }
    }
    long long t1 = clock64();
    float s = 0; u64 S = 0; int is = 0;
#pragma unroll
    for (int j = 0; j < UNROLL; j++) { s += a[j] + b[j]; S += A[j]; is += ia[j]; }
    out[blockIdx.x * blockDim.x + tid] = s + (float)S + (float)is + sm[tid];
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(int threads, float *out, long long *cyc)
{
    int nsm = 148;
    cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 0);
    bench<OP><<<nsm, threads>>>(out, cyc, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nsm, threads>>>(out, cyc, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += (double)h[i]; avg /= nsm;
    int per_iter = UNROLL;
    if (OP == MIX_F2_ALU || OP == MIX_F2_F1) per_iter = 2 * UNROLL;
    if (OP == MIX_F2_LDS || OP == MIX_F1_LDS) per_iter = UNROLL + UNROLL / 4;
    if (OP == FFMA2_SW) per_iter = 2 * UNROLL;
    double winst = (double)ITERS * per_iter * (threads / 32);
    printf("%-28s threads/SM %4d  cycles %9.0f  warp-inst/cycle/SM %6.3f  (%.3f ms, err=%d)\n", names[OP], threads, avg, winst / avg, ms, (int)cudaGetLastError());
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {256, 512, 1024}) {
        run<FADD1>(threads, out, cyc); run<FFMA1>(threads, out, cyc); run<FFMA1_2SRC>(threads, out, cyc); run<FMUL1>(threads, out, cyc);
        run<FADD2>(threads, out, cyc); run<FFMA2>(threads, out, cyc); run<FMUL2>(threads, out, cyc); run<FFMA2_SW>(threads, out, cyc);
        run<MIX_F2_ALU>(threads, out, cyc); run<MIX_F2_F1>(threads, out, cyc); run<MIX_F2_LDS>(threads, out, cyc); run<MIX_F1_LDS>(threads, out, cyc);
        run<LDS64>(threads, out, cyc); run<LDS128>(threads, out, cyc); run<STS64>(threads, out, cyc);
        run<I2F16>(threads, out, cyc); run<F2I>(threads, out, cyc); run<MUFU>(threads, out, cyc);
        run<ATOMS_SPREAD>(threads, out, cyc); run<ATOMS_SAME>(threads, out, cyc); run<PRMT>(threads, out, cyc);
        printf("\n");
    }
    return 0;
}
