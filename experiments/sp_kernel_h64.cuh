// sp_kernel_h64.cuh — the N = 4096 headline kernel with HALF a 64-point transform per thread.
//
// render_r64_kernel (sp_kernel_r64.cuh) keeps 64 complex points in a thread: 232 registers, so only 8 FFT warps fit an SM
// (two per scheduler), and ncu shows what that costs: inside the butterfly blocks the two warps saturate the FMA pipe, but
// over the whole frame the pipe is busy 59 % of the time - whenever one of the two is in a load / exchange / epilogue phase
// the other cannot fill it alone (profiles/r02_ncu_summary_r64.txt).  This kernel keeps the 64 x 64 decomposition, the one
// shared-memory exchange, the joint histogram, the staging halves and the store warpgroup of render_r64_kernel, and splits
// every 64-point transform over TWO threads by one decimation-in-frequency step:
//     even outputs  = DFT32( x[a] + x[a + 32] ),     odd outputs = DFT32( (x[a] - x[a + 32]) * W64^a ),   a = 0..31
// Thread (t, h) forms its 32 inputs itself from all 64 points (the raw samples / the exchange row are read by both threads
// of a pair; h is warp-uniform, so the two variants do not diverge) and runs one DFT-32: 32 complex points, 104 registers,
// 16 FFT warps per SM (four per scheduler) with no extra exchange and no shuffles.  The price is the second read of every
// raw sample and exchange row (+27 % shared-memory wavefronts) and the duplicated decode (+7 % FMA-pipe work).
// Plain spectrogram layout with tensor-TMA row stores only; every option stays with render_r64_kernel.
// Replaces the hot loops of reference lib/worker.js:68-137 (+ lib/samples.js:313-400, lib/fft_nayuki.js:54-96).
#pragma once
#include "sp_kernel_r64.cuh"
#include <type_traits>

#ifndef SP_H64_FFT_REGS
#define SP_H64_FFT_REGS 104
#define SP_H64_STORE_REGS 40
#endif

namespace sp {

template <int FMT> struct H64Cfg {
    static constexpr int N = 4096, T = 64, STREAMS = 4, SPT = 128, FFT_THREADS = SPT * STREAMS, STORE_THREADS = 128, THREADS = FFT_THREADS + STORE_THREADS;
    static constexpr int STEPS = 4, F = STREAMS * STEPS;                         // 16 frames per tile
    static constexpr int FFT_REGS = SP_H64_FFT_REGS, STORE_REGS = SP_H64_STORE_REGS;   // <= 640 * 96, the CTA's allocation at launch
    static_assert(FFT_THREADS * FFT_REGS + STORE_THREADS * STORE_REGS <= THREADS * 96, "register budget");
    static constexpr int SWB = sample_width(FMT == FMT_RUNTIME ? CF64 : FMT);
    static constexpr bool OK = (FMT != FMT_RUNTIME) && SWB <= 8;                 // raw frame fits the exchange buffer
    static constexpr int XP = 66, X_BYTES = 64 * XP * 8;                         // as R64Cfg
    static constexpr int TW_PITCH = 12;                                          // float2 per column t: W^{2t j} j = 1..7, W^{16 t i} i = 1..3, W^t, pad
    static constexpr int ST_PITCH = 1026;
    static constexpr int TILE_BYTES = (STORE_THREADS / 32) * 4096;
    static constexpr size_t SMEM_BYTES = (size_t)STREAMS * X_BYTES + (size_t)F * ST_PITCH * 4 + (size_t)JH_SIZE * 4 + TILE_BYTES
                                       + (size_t)T * TW_PITCH * 8 + 1024 /* LUT */ + (size_t)F * 4 * 8 + 128 + 1024 /* LUT alignment */;
};

template <int I, int E, class Fn> __device__ __forceinline__ void static_for(Fn fn)
{
    if constexpr (I < E) {
        fn(std::integral_constant<int, I>{});
        static_for<I + 1, E>(fn);
    }
}

// bin k1 index of byte j of staging word m (m = 2 m' + h holds outputs 4 m' .. 4 m' + 3 of the pair's thread h: k1 = 2 (4 m' + j) + h)
__device__ __forceinline__ int h64_k1(int m, int j) { return 8 * (m >> 1) + 2 * j + (m & 1); }

// tw12: [64][12] float2 (see H64Cfg::TW_PITCH)
template <int FMT>
__global__ void __launch_bounds__(640, 1) render_h64_kernel(const Params p, const float2 *__restrict__ tw12, const __grid_constant__ CUtensorMap tmap)
{
    using B = H64Cfg<FMT>;
    constexpr int N = B::N, T = B::T, F = B::F;
    constexpr bool FLOAT_IN = FMT == CF32 || FMT == CF64 || FMT == FMT_RUNTIME;   // |X|^2 may be +inf / NaN
    extern __shared__ __align__(128) unsigned char smem_h64[];
    const unsigned lut_base = (smem_u32(smem_h64) + 1023u) & ~1023u;
    unsigned char *s_x = smem_h64 + (lut_base - smem_u32(smem_h64)) + 1024;           // [4][X_BYTES] exchange / raw frame
    unsigned *s_lut = reinterpret_cast<unsigned *>(s_x - 1024);                       // [256] RGBA indexed by the staged byte
    unsigned char *s_tiles = s_x + B::STREAMS * B::X_BYTES;                          // [4 store warps][4 boxes][32 rows][32 B] RGBA
    unsigned *s_stage = reinterpret_cast<unsigned *>(s_tiles + B::TILE_BYTES);       // [16][1026] colour bytes (4 bins per word)
    unsigned *s_jh = s_stage + F * B::ST_PITCH;                                       // [JH_SIZE] joint histogram
    float2 *s_tw = reinterpret_cast<float2 *>(s_jh + JH_SIZE);                        // [64][12]
    uint2 *s_mm = reinterpret_cast<uint2 *>(s_tw + T * B::TW_PITCH);                  // [16][4] per-warp min/max bit patterns of |X|^2
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(s_mm + F * 4);                    // [4]
    int *s_off = reinterpret_cast<int *>(s_mbar + B::STREAMS);                        // [4][2] misalignment of the staged frame
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_off + B::STREAMS * 2);          // [2] staging half h holds 8 finished frames
    uint64_t *s_empty = s_full + 2;                                                   // [2] staging half h has been stored

    const int tid = threadIdx.x;
    const int s = (tid >> 7) & 3;           // stream
    const int u = tid & 127;
    const int h = u >> 6;                   // which half of the pair's outputs (warp-uniform)
    const int t = u & 63;                   // pass A: input column; pass B: exchange row k0
    float2 *X = reinterpret_cast<float2 *>(s_x + (size_t)s * B::X_BYTES);
    unsigned char *raw = reinterpret_cast<unsigned char *>(X);
    uint64_t *mbar = s_mbar + s;

    for (int i = tid; i < JH_SIZE; i += B::THREADS) s_jh[i] = 0;
    const int cmax = p.cmap_len - 1;
    const JhConst jc = jh_const(p);
    for (int i = tid; i < 256; i += B::THREADS) s_lut[i] = i <= cmax ? p.lut[jc.rev ? cmax - i : i] : 0u;
    for (int i = tid; i < T * B::TW_PITCH; i += B::THREADS) s_tw[i] = tw12[i];
    const unsigned jh_base = smem_u32(s_jh) - (JH_MAGIC_BITS << 2);      // address of joint bin j = S.bits * 4 + jh_base (mod 2^32)

    // u == 0 of a stream: start the bulk copy of chunk-relative frame xr into the stream's buffer
    auto stage = [&](long long xr, unsigned par) {
        if (xr >= p.chunk_frames) xr = p.chunk_frames - 1;              // frames past the end of a partial tile redo the last one
        const long long xgl = p.frame_first + p.chunk_first + xr;
        const long long p0 = (long long)__dadd_rn(0.5, __dmul_rn(p.stride, (double)xgl)) - p.sample_base;   // lib/worker.js:72
        const unsigned long long off = (unsigned long long)p0 * B::SWB, a0 = off & ~15ull;
        s_off[s * 2 + par] = (int)(off - a0);
        tma_load_1d(raw, p.buf + a0, (unsigned)(((off - a0) + (unsigned long long)N * B::SWB + 15) & ~15ull), mbar);
    };

    if (u == 0 && tid < B::FFT_THREADS) mbar_init(mbar, 1);
    if (tid == 0) {
        for (int hh = 0; hh < 2; hh++) { mbar_init(s_full + hh, B::FFT_THREADS / 32); mbar_init(s_empty + hh, B::STORE_THREADS / 32); }
    }
    if (u == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    long long tile = blockIdx.x;
    unsigned kk = 0;                        // tiles done by this CTA (phase of the full / empty barriers)

    if (tid >= B::FFT_THREADS) {
        // ================= store warps: staged colour bytes -> LUT -> RGBA tiles -> tensor-TMA row stores (lib/worker.js:115-121) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B::STORE_REGS));
        const int ht = tid - B::FFT_THREADS;
        const int lane = ht & 31, wq = ht >> 5;
        unsigned char *tl = s_tiles + wq * 4096;
        const int r = 31 - lane;
        const unsigned toff = (unsigned)(r * 32 + (((r >> 2) & 1) << 4));       // 32-byte TMA swizzle: see render_r64_kernel
        for (; tile < p.ntiles; tile += gridDim.x, kk++) {
            const long long xr0 = tile * F;
#pragma unroll 1
            for (int hh = 0; hh < 2; hh++) {
                mbar_wait(s_full + hh, kk & 1);
                const size_t x0 = (size_t)(p.chunk_first + xr0) + 8 * hh;
                const bool live = xr0 + 8 * hh < p.chunk_frames;       // partial last tile: chunk_frames is a multiple of 8
#pragma unroll 1
                for (int i = 0; i < (live ? 8 : 0); i++) {
                    const int id = ht + B::STORE_THREADS * i, k0 = id & 63, m = id >> 6;
                    const unsigned *src = s_stage + (8 * hh) * B::ST_PITCH + m * 64 + k0;
                    unsigned w[8];
#pragma unroll
                    for (int f = 0; f < 8; f++) w[f] = src[f * B::ST_PITCH];
                    if (lane == 0) bulk_wait_read0();              // the previous iteration's boxes have left the tile
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint4 a, b;
                        a.x = lut_at(lut_base, w[0], j); a.y = lut_at(lut_base, w[1], j); a.z = lut_at(lut_base, w[2], j); a.w = lut_at(lut_base, w[3], j);
                        b.x = lut_at(lut_base, w[4], j); b.y = lut_at(lut_base, w[5], j); b.z = lut_at(lut_base, w[6], j); b.w = lut_at(lut_base, w[7], j);
                        *reinterpret_cast<uint4 *>(tl + j * 1024 + toff) = a;
                        *reinterpret_cast<uint4 *>(tl + j * 1024 + (toff ^ 16u)) = b;
                        if (k0 + 64 * h64_k1(m, j) == N / 2)       // bin n/2 -> image row 0 (clipped out of its box)
                            st_global_256(reinterpret_cast<uint32_t *>(p.image) + x0, a, b);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tma_store_2d(&tmap, smem_u32(tl + j * 1024), (int)x0, (N / 2 - ((k0 & 32) + 64 * h64_k1(m, j)) - 31) & (N - 1));
                        bulk_commit();
                    }
                }
                if (ht < 8 && live) {
                    // per-frame min / max of the half's frames, folded across the four warps of their stream, as dB
                    const int fl = 8 * hh + ht;
                    const long long xl = p.chunk_first + xr0 + fl;
                    const uint2 m0 = s_mm[fl * 4], m1 = s_mm[fl * 4 + 1], m2 = s_mm[fl * 4 + 2], m3 = s_mm[fl * 4 + 3];
                    const unsigned umn = min(min(m0.x, m1.x), min(m2.x, m3.x)), umx = max(max(m0.y, m1.y), max(m2.y, m3.y));
                    p.fmin[xl] = fminf(0.0f, fmaf(fast_log2(__uint_as_float(umn)), p.c1, p.c0));       // lib/worker.js:82,102
                    p.fmax[xl] = fmaxf(-200.0f, fmaf(fast_log2(__uint_as_float(umx)), p.c1, p.c0));    // lib/worker.js:83,103
                }
                __syncwarp();                                   // this warp is done reading the half (and s_mm)
                if (lane == 0) mbar_arrive(s_empty + hh);
            }
        }
        bulk_wait_read0();                                      // the tiles must outlive the last tensor stores
    } else {
    // ================= FFT warps: four frame streams of 128 threads =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B::FFT_REGS));
    auto stream_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(s + 1) : "memory"); };
    unsigned fpar = 0;                      // parity of this stream's frame counter (mbarrier phase, s_off slot)
    if (u == 0 && tile < p.ntiles) stage(tile * F + s, 0);
    const float4 *wrow = p.window_t + t;                                // [16][64] float4: quad q holds w[64 (4q + c) + t]
    float2 *Zw = X + h * B::XP + t;                                     // pass A writes Z[2 k' + h][t]
    const float4 *Zr = reinterpret_cast<const float4 *>(X + t * B::XP); // pass B reads row k0 = t

    while (tile < p.ntiles) {
        const long long xr0 = tile * F;
        const long long next_tile = tile + gridDim.x;
#pragma unroll 1
        for (int step = 0; step < B::STEPS; step++) {
            const int fl = step * B::STREAMS + s;                       // frame of the tile handled by this stream now
            const bool valid = xr0 + fl < p.chunk_frames;               // false: past the end of a partial last tile (outputs suppressed)
            const unsigned jbase = valid ? jh_base : jh_base + (smem_u32(s_stage + 8 * B::ST_PITCH) - smem_u32(s_jh));   // see render_r64_kernel
            int half = step >> 1;                                       // staging half of this frame
            asm volatile("" : "+r"(half));
            cf v[32];
            // ---------------- load + decode + window (lib/worker.js:70-75) + the decimation step of pass A ----------------
            mbar_wait(mbar, fpar);
            {
                const unsigned char *rp = raw + s_off[s * 2 + fpar];
                auto front = [&](auto H) {
                    static_for<0, 8>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        const float4 wa = __ldg(wrow + 64 * g), wb = __ldg(wrow + 64 * (g + 8));
                        const float was[4] = { wa.x, wa.y, wa.z, wa.w }, wbs[4] = { wb.x, wb.y, wb.z, wb.w };
                        static_for<0, 4>([&](auto C) {
                            constexpr int c = decltype(C)::value, a = 4 * g + c;
                            const cf da = decode_raw_cf<FMT>(rp, T * a + t, p.format), db = decode_raw_cf<FMT>(rp, T * (a + 32) + t, p.format);
                            if constexpr (a == 0 && decltype(H)::value == 0) {
                                // raw sample at p0 + n/2 (lib/worker.js:131-133); the power-of-two scale is exact
                                if (t == 0 && valid) p.fmid[p.chunk_first + xr0 + fl] = make_float2(cre(db) * raw_scale<FMT>(), cim(db) * raw_scale<FMT>());
                            }
                            const cf xa = cscale(da, was[c]);
                            if constexpr (decltype(H)::value == 0) v[a] = cfma2(db, cpk(wbs[c], wbs[c]), xa);
                            else v[a] = mul_w64<a>(cfma2(db, cpk(-wbs[c], -wbs[c]), xa));
                        });
                    });
                };
                if (h == 0) front(std::integral_constant<int, 0>{}); else front(std::integral_constant<int, 1>{});
            }
            fpar ^= 1;
            stream_bar();                                               // every thread of the stream has consumed the raw frame
            dft<32>(v);                                                 // v[k'] = output 2 k' + h of the column's DFT-64
            {
                // twiddles W^{t (2 k' + h)}, k' = j + 8 i: (W^{2 t j} * [W^t]) * W^{16 t i}; each product is followed by its exchange store
                const float4 *twp = reinterpret_cast<const float4 *>(s_tw + t * B::TW_PITCH);
                float2 w[8];
                const float4 a = twp[0], b = twp[1], c = twp[2], d = twp[3], e = twp[4], f = twp[5];
                w[0] = make_float2(1.0f, 0.0f);
                w[1] = make_float2(a.x, a.y); w[2] = make_float2(a.z, a.w); w[3] = make_float2(b.x, b.y); w[4] = make_float2(b.z, b.w);
                w[5] = make_float2(c.x, c.y); w[6] = make_float2(c.z, c.w); w[7] = make_float2(d.x, d.y);
                float2 hi[4];
                hi[1] = make_float2(d.z, d.w); hi[2] = make_float2(e.x, e.y); hi[3] = make_float2(e.z, e.w);
                if (h) {
                    const float2 wt = make_float2(f.x, f.y);
                    w[0] = wt;
#pragma unroll
                    for (int j = 1; j < 8; j++) w[j] = cun(cmul(cpk(w[j]), wt));
                    v[0] = cmul(v[0], wt);
                }
#pragma unroll
                for (int j = 1; j < 8; j++) v[j] = cmul(v[j], w[j]);
#pragma unroll
                for (int j = 0; j < 8; j++) cst(Zw + 2 * j * B::XP, v[j]);
#pragma unroll
                for (int i = 1; i < 4; i++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) v[8 * i + j] = cmul(v[8 * i + j], cun(cmul(cpk(hi[i]), w[j])));
#pragma unroll
                    for (int j = 0; j < 8; j++) cst(Zw + 2 * (8 * i + j) * B::XP, v[8 * i + j]);
                }
            }
            stream_bar();
            // ---------------- pass B: row k0 = t; the decimation step, then DFT-32: v[k'] is bin t + 64 (2 k' + h) ----------------
            {
                auto front = [&](auto H) {
                    static_for<0, 16>([&](auto M) {
                        constexpr int m = decltype(M)::value;
                        const float4 q0 = Zr[m], q1 = Zr[m + 16];
                        const cf za0 = cpk(q0.x, q0.y), za1 = cpk(q0.z, q0.w), zb0 = cpk(q1.x, q1.y), zb1 = cpk(q1.z, q1.w);
                        if constexpr (decltype(H)::value == 0) { v[2 * m] = cadd(za0, zb0); v[2 * m + 1] = cadd(za1, zb1); }
                        else { v[2 * m] = mul_w64<2 * m>(csub(za0, zb0)); v[2 * m + 1] = mul_w64<2 * m + 1>(csub(za1, zb1)); }
                    });
                };
                if (h == 0) front(std::integral_constant<int, 0>{}); else front(std::integral_constant<int, 1>{});
            }
            stream_bar();                                               // the exchange buffer is free: prefetch the stream's next frame
            if (u == 0) {
                if (step < B::STEPS - 1) stage(xr0 + fl + B::STREAMS, fpar);
                else if (next_tile < p.ntiles) stage(next_tile * F + s, fpar);
            }
            dft<32>(v);
            // first frame of this stream in staging half step/2: the store warps must be done with the half (previous tile)
            if ((step & 1) == 0) mbar_wait(s_empty + half, (kk + 1) & 1);

            // ---------------- per-bin epilogue (lib/worker.js:85-122) ----------------
            float amin = __int_as_float(0x7f800000), amax = 0.0f, prev = 0.0f;
            unsigned umin_i = 0x7f800000u, umax_i = 0u;
            unsigned *stg = s_stage + fl * B::ST_PITCH + h * 64 + t;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                unsigned yb[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 vi = cun(v[4 * m + j]);
                    const float abs2 = fmaf(vi.x, vi.x, vi.y * vi.y);
                    if constexpr (FLOAT_IN) {
                        umin_i = min(umin_i, __float_as_uint(abs2));
                        umax_i = max(umax_i, __float_as_uint(abs2));
                    } else if (j & 1) {
                        amin = fmin3(amin, prev, abs2);
                        amax = fmax3(amax, prev, abs2);
                    } else prev = abs2;
                    float Y;
                    const float S = jh_eval(fast_log2(abs2), jc, Y);    // 2^23 + joint index, 2^23 + (cmax - colour index)
                    red_shared_inc_addr(jbase + (__float_as_uint(S) << 2));
                    yb[j] = __float_as_uint(Y);
                }
                // outputs 4m .. 4m+3 of this thread: bins t + 64 (2 (4m + j) + h), four colour bytes in staging word 2m + h
                stg[m * 128] = __byte_perm(__byte_perm(yb[0], yb[1], 0x0040), __byte_perm(yb[2], yb[3], 0x0040), 0x5410);
            }
            unsigned umn, umx;
            if constexpr (FLOAT_IN) {
                umn = __reduce_min_sync(0xffffffffu, umin_i);
                umx = __reduce_max_sync(0xffffffffu, umax_i);
            } else {
                umn = __reduce_min_sync(0xffffffffu, __float_as_uint(amin));
                umx = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
            }
            if (umn < 0x00800000u || umx >= 0x7f800000u) {
                // rare (warp-uniform): |X|^2 == 0 (flushed: d0 = -inf), +inf or NaN in the frame: see render_r64_kernel
                unsigned nzero = 0, nbad = 0, nnan = 0;
                float mn = __int_as_float(0x7f800000), mx = 0.0f;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float2 vi = cun(v[i]);
                    const float abs2 = fmaf(vi.x, vi.x, vi.y * vi.y);
                    nzero += abs2 < 1.17549435e-38f ? 1u : 0u;
                    nbad += !(abs2 <= 3.402823466e38f) ? 1u : 0u;
                    nnan += abs2 != abs2 ? 1u : 0u;
                    mn = fminf(mn, abs2 < 1.17549435e-38f ? 0.0f : abs2);
                    mx = fmaxf(mx, abs2);
                }
                nzero = __reduce_add_sync(0xffffffffu, nzero);
                nbad = __reduce_add_sync(0xffffffffu, nbad);
                nnan = __reduce_add_sync(0xffffffffu, nnan);
                umn = __reduce_min_sync(0xffffffffu, __float_as_uint(mn));
                umx = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
                if ((t & 31) == 0 && valid) {
                    if (nzero) atomicAdd(&s_jh[JH_ZERO], nzero);
                    if (nbad) atomicAdd(&s_jh[JH_BAD], nbad);
                    if (nnan) {
                        float Yn;
                        const float Sn = jh_eval(__int_as_float(0x7fffffff), jc, Yn);
                        atomicSub(&s_jh[__float_as_uint(Sn) - JH_MAGIC_BITS], nnan);
                        atomicAdd(&s_jh[JH_NAN], nnan);
                    }
                }
            }
            if ((t & 31) == 0) s_mm[fl * 4 + (u >> 5)] = make_uint2(umn, umx);
            if (step & 1) {                 // this warp has staged its last frame of half step/2 (and its s_mm entries)
                __syncwarp();
                if ((t & 31) == 0) mbar_arrive(s_full + half);
            }
        } // steps
        tile = next_tile;
        kk++;
    } // tiles
    } // FFT warps

    __syncthreads();
    for (int i = tid; i < JH_SIZE; i += B::THREADS)
        if (s_jh[i]) atomicAdd(&p.j_hist[i], (unsigned long long)s_jh[i]);
}

} // namespace sp
