// Micro-benchmark: issue rate of the FP32 forms the fused kernels use (per SM sub-partition, cycles per warp instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_forms fma_forms.cu && ./fma_forms
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
#define NACC 8
struct cf { unsigned long long u; };
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *out, long long *cyc, float a, float b, float c0)
{
    float x[NACC];
    unsigned long long y[NACC];
    for (int i = 0; i < NACC; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = pk(x[i], x[i] + 1.0f); }
    const unsigned long long pa = pk(a, a), pb = pk(b, b);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                if (MODE == 0) x[i] = fmaf(x[i], a, b);                       // FFMA R, R, R, R
                if (MODE == 1) x[i] = fmaf(x[i], 1.0001f, b);                 // FFMA R, R, imm, R
                if (MODE == 2) x[i] = fmaf(x[i], 1.0001f, 0.5f);              // two immediates -> one must be a register / const
                if (MODE == 3) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(pa), "l"(pb));      // FFMA2 R, R, R, R
                if (MODE == 4) x[i] = __saturatef(fmaf(x[i], a, b));          // FFMA.SAT
                if (MODE == 5) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(pa));                     // FADD2
                if (MODE == 6) x[i] = x[i] + a;                                // FADD R, R, R
                if (MODE == 7) x[i] = x[i] * a;                                // FMUL R, R, R
                if (MODE == 8) x[i] = fmaf(x[i], x[(i + 1) % NACC], b);        // three distinct registers, no reuse
                if (MODE == 10) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(y[(i + 1) % NACC]));                               // FADD2 pair + pair
                if (MODE == 11) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(pa), "l"(y[(i + 3) % NACC]));                // FFMA2 pair * scalar + pair
                if (MODE == 12) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(y[(i + 1) % NACC]));                               // FMUL2 pair * pair
                if (MODE == 13) { float2 f; asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(y[(i + 1) % NACC]));           // FADD2 pair + swapped, half-negated pair
                                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(pk(f.y, -f.x))); }
                if (MODE == 14) { x[i] = fmaf(x[i], a, b); asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(y[(i + 1) % NACC])); } // FFMA + FADD2 alternating (two instructions)
                if (MODE == 15) { x[i] = __uint_as_float(__float_as_uint(x[i]) ^ (unsigned)it); asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(y[(i + 1) % NACC])); } // LOP3 + FADD2
                if (MODE == 9) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(y[(i + 1) % NACC]), "l"(y[(i + 3) % NACC]));   // FFMA2, three distinct pairs
            }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < NACC; i++) { s += x[i]; s += __uint_as_float((unsigned)y[i]) + __uint_as_float((unsigned)(y[i] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + c0;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int threads)
{
    float *out; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, threads>>>(out, cyc, 0.999f, 0.25f, 0.f);
    k<MODE><<<148, threads>>>(out, cyc, 0.999f, 0.25f, 0.f);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double inst_per_smsp = (double)ITER * 4 * NACC * (threads / 32) / 4.0;
    printf("%-44s warps/SMSP %d  cycles/warp-instr/SMSP %.3f  (%s)\n", name, threads / 128, h[0] / inst_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int th : { 256, 512 }) {
        run<0>("FFMA R,R,R,R (a, b in registers)", th);
        run<1>("FFMA R,R,imm,R", th);
        run<2>("FFMA R,R,imm,imm-as-written", th);
        run<3>("FFMA2 R,R,R,R", th);
        run<4>("FFMA.SAT R,R,R,R", th);
        run<5>("FADD2 R,R,R", th);
        run<6>("FADD R,R,R", th);
        run<7>("FMUL R,R,R", th);
        run<8>("FFMA three distinct registers", th);
        run<9>("FFMA2 three distinct pairs", th);
        run<10>("FADD2 pair + pair", th);
        run<11>("FFMA2 pair * scalar + pair", th);
        run<12>("FMUL2 pair * pair", th);
        run<13>("FADD2 pair + swapped half-negated pair", th);
        run<14>("FFMA + FADD2 alternating (per 2 instr)", th);
        run<15>("LOP3 + FADD2 alternating (per 2 instr)", th);
    }
    return 0;
}
