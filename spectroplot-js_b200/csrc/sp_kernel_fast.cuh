// sp_kernel_fast.cuh — the N = 4096 fast path of the fused render kernel (the headline size, and the
// second stage of the four-step path for N = 8192..65536).
//
// One persistent CTA per SM with TWO independent frame slots of 256 threads (16 warps/SM):
//   * a slot transforms one frame at a time (16 points per thread, 16 x 16 x 16 passes, packed
//     FADD2/FMUL2/FFMA2 arithmetic) and synchronises only with itself through a named barrier, so
//     the slots drift out of phase and one slot's shared-memory latency is covered by the other's
//     FMA work;
//   * raw sample bytes of the NEXT frame are staged into shared memory by one TMA bulk copy
//     (cp.async.bulk + mbarrier) issued by one thread as soon as the current frame is in registers;
//     that thread also does the double-precision frame-position arithmetic (lib/worker.js:72) and
//     leaves the 16-byte misalignment of the frame in shared memory for the other 255;
//   * per pass only SIX twiddles are loaded (w^1, w^2, w^3, w^4, w^8, w^12; three LDS.128) and the
//     other nine are formed as products in registers — the kernel is bound by the shared-memory /
//     LSU data pipe, not by the FMA pipe (ncu: profiles/r01_ncu_summary_p1_packed.txt);
//   * a padded exchange buffer per slot is reused in place by both exchanges (element (k0, a, b) at
//     k0*272 + 17*a + b: every access pattern is bank-conflict free); where shared memory allows
//     (<= 4 bytes per sample) two such buffers alternate per frame, so a frame costs two slot
//     barriers, not three;
//   * colour indices of 8 consecutive frames are packed in registers and each image row segment
//     (8 frames x RGBA = one 32-byte sector) is written with ONE 256-bit store (STG.E.ENL2.256) —
//     the transposed store of lib/worker.js:117;
//   * per-frame min / max of |X|^2 are folded with REDUX on the (order-preserving) bit patterns and
//     converted to dB once per frame, not per thread;
//   * tiles (8 frames) are handed out dynamically through a global counter (no tail imbalance).
// The kernel only ever sees FULL tiles of frames that lie inside the buffer (the engine routes the
// remainder, and every other option — waterfall, split-real, dB tap, odd widths, long colormaps —
// through render_kernel), so the hot loop carries no per-frame predicates.
// Replaces the hot loops of reference lib/worker.js:68-137 (+ lib/samples.js:313-400,
// lib/fft_nayuki.js:54-96).
#pragma once
#include "sp_kernels.cuh"

namespace sp {

template <int FMT, bool SUB> struct FastCfg {
    static constexpr int N = 4096, T = 256, SLOTS = 2, THREADS = T * SLOTS, F = 8;
    // sub-frame mode reads complex fp32 sub-sequences written by the pre-pass, whatever the capture format is
    static constexpr int SWB = SUB ? 8 : sample_width(FMT == FMT_RUNTIME ? CF64 : FMT);
    static constexpr bool STAGE = SUB || ((FMT != FMT_RUNTIME) && SWB <= 8);     // TMA-staged input
    static constexpr int RAW_BYTES = STAGE ? N * SWB + 32 : 0;
    static constexpr int PA = 272, P1 = 17;                                      // exchange pitches (float2)
    static constexpr int X_FLOAT2 = 16 * PA;
    static constexpr int WIN_PITCH = 20;                                         // floats per thread row (16 used): conflict-free LDS.128
    static constexpr int TW_PITCH = 6;                                           // float2 per row: w^1 w^2 w^3 w^4 w^8 w^12
    // two exchange buffers (alternating per frame) remove the write-after-read barrier between consecutive
    // frames; they fit next to a staged frame of up to 4 bytes per sample
    static constexpr bool DBX = STAGE && RAW_BYTES <= 4096 * 4 + 32;
    static constexpr size_t SLOT_BYTES = (size_t)X_FLOAT2 * 8 * (DBX ? 2 : 1) + RAW_BYTES + 2 * F * 8 * 8 + 2 * F * 16 + 64;
    static constexpr size_t SHARED_BYTES = (size_t)T * TW_PITCH * 8 + 16 * TW_PITCH * 8 + (SUB ? 0 : (size_t)T * WIN_PITCH * 4)
                                         + (size_t)CB_RAW * 4 + 256 * 4 + 256 * 4;
    static constexpr size_t SMEM_BYTES = SHARED_BYTES + SLOTS * SLOT_BYTES + 128;
};

__device__ __forceinline__ void slot_barrier(int slot)
{
    asm volatile("bar.sync %0, 256;" ::"r"(slot + 1) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one thread: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of dst vs the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_global_256(void *p, uint4 a, uint4 b)
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void red_shared_inc(unsigned *p)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

__device__ __forceinline__ void red_shared_inc_addr(unsigned addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// v[k] *= w^k, k = 1..15, from the six table entries tw[0..5] = w^1 w^2 w^3 w^4 w^8 w^12
__device__ __forceinline__ void twiddle16_from6(cf (&v)[16], const float4 t01, const float4 t23, const float4 t45)
{
    const float2 w1 = make_float2(t01.x, t01.y), w2 = make_float2(t01.z, t01.w), w3 = make_float2(t23.x, t23.y);
    const float2 w4 = make_float2(t23.z, t23.w), w8 = make_float2(t45.x, t45.y), w12 = make_float2(t45.z, t45.w);
    v[1] = cmul(v[1], w1);  v[2] = cmul(v[2], w2);  v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);  v[8] = cmul(v[8], w8);  v[12] = cmul(v[12], w12);
    v[5] = cmul(v[5], cun(cmul(cpk(w4), w1)));   v[6] = cmul(v[6], cun(cmul(cpk(w4), w2)));   v[7] = cmul(v[7], cun(cmul(cpk(w4), w3)));
    v[9] = cmul(v[9], cun(cmul(cpk(w8), w1)));   v[10] = cmul(v[10], cun(cmul(cpk(w8), w2))); v[11] = cmul(v[11], cun(cmul(cpk(w8), w3)));
    v[13] = cmul(v[13], cun(cmul(cpk(w12), w1))); v[14] = cmul(v[14], cun(cmul(cpk(w12), w2))); v[15] = cmul(v[15], cun(cmul(cpk(w12), w3)));
}

// tw6A: [256][6] float2 = W_4096^{t*k}; tw6B: [16][6] float2 = W_256^{b*k}; k = 1, 2, 3, 4, 8, 12
template <int FMT, bool SUB>
__global__ void __launch_bounds__(512, 1) render_fast_kernel(const Params p, const float2 *__restrict__ tw6A,
                                                             const float2 *__restrict__ tw6B, unsigned *__restrict__ tile_counter)
{
    using B = FastCfg<FMT, SUB>;
    constexpr int N = B::N, T = B::T, F = B::F;
    extern __shared__ __align__(128) unsigned char smem_fast[];
    // ---- shared by the slots
    float2 *s_twA = reinterpret_cast<float2 *>(smem_fast);                       // [256][6]
    float2 *s_twB = s_twA + T * B::TW_PITCH;                                     // [16][6]
    float *s_win = reinterpret_cast<float *>(s_twB + 16 * B::TW_PITCH);          // [256][20] (row t: window[256 a + t], a = 0..15)
    unsigned *s_cb = reinterpret_cast<unsigned *>(s_win + (SUB ? 0 : T * B::WIN_PITCH));   // [CB_RAW] raw dB histogram
    unsigned *s_cnt = s_cb + CB_RAW;                                             // [256] colour histogram
    unsigned *s_lut = s_cnt + 256;                                               // [256] RGBA
    unsigned char *slot_base = reinterpret_cast<unsigned char *>(s_lut + 256);
    slot_base += (128 - (reinterpret_cast<uintptr_t>(slot_base) & 127)) & 127;

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid % T;
    unsigned char *my = slot_base + (size_t)slot * B::SLOT_BYTES;
    float2 *X0 = reinterpret_cast<float2 *>(my);                                 // [1 or 2][16][272] in-place exchange
    unsigned char *raw = my + (size_t)B::X_FLOAT2 * 8 * (B::DBX ? 2 : 1);        // [RAW_BYTES] TMA destination
    uint2 *s_mm = reinterpret_cast<uint2 *>(raw + B::RAW_BYTES);                 // [2][F][8] per-warp min/max bit patterns of |X|^2
    ulonglong2 *s_pos = reinterpret_cast<ulonglong2 *>(s_mm + 2 * F * 8);        // [2][F] {source address, bytes | misalignment << 32}
    uint64_t *mbar = reinterpret_cast<uint64_t *>(s_pos + 2 * F);
    int *s_off = reinterpret_cast<int *>(mbar + 1);                              // [2] byte misalignment of the staged frame (by frame parity)
    unsigned *s_next = reinterpret_cast<unsigned *>(s_off + 2);                  // [2] tile ring

    for (int i = tid; i < CB_RAW; i += B::THREADS) s_cb[i] = 0;
    for (int i = tid; i < 256; i += B::THREADS) { s_cnt[i] = 0; s_lut[i] = i < p.cmap_len ? p.lut[i] : 0u; }
    for (int i = tid; i < T * B::TW_PITCH; i += B::THREADS) s_twA[i] = tw6A[i];
    for (int i = tid; i < 16 * B::TW_PITCH; i += B::THREADS) s_twB[i] = tw6B[i];
    if constexpr (!SUB)
        for (int i = tid; i < N; i += B::THREADS) s_win[(i & 255) * B::WIN_PITCH + (i >> 8)] = p.window[i];
    const int nfull = p.n_full, sub_r = p.sub_r;
    const unsigned cmax_u = (unsigned)(p.cmap_len - 1);
    const float gc5 = p.gc + 0.5f;
    const float l2c_k = -10.0f * p.c1, l2c_k0 = fmaf(-10.0f, p.c0, 2.5f);       // raw cB index = l2c_k * log2|X|^2 + l2c_k0
    const float l2c_g = p.gn * p.c1, l2c_g0 = fmaf(p.gn, p.c0, gc5);             // colour index  = l2c_g * log2|X|^2 + l2c_g0
    const cf l2c_kg = cpk(l2c_k, l2c_g), l2c_kg0 = cpk(l2c_k0, l2c_g0);
    const int k0p = t >> 4, lo4 = t & 15;                                        // pass B: (k0, b1); pass C: (k0, k1)
    const int binbase = k0p + 16 * lo4;                                          // bins binbase + 256*k2

    // frame position of chunk-relative frame xr (lib/worker.js:72: separate multiply and add, like JavaScript)
    auto frame_p0 = [&](long long xr) -> long long {
        const long long xgl = p.frame_first + p.chunk_first + xr;
        return (long long)__dadd_rn(0.5, __dmul_rn(p.stride, (double)xgl)) - p.sample_base;
    };
    // tile -> first chunk-relative frame
    auto tile_xr0 = [&](long long tile) -> long long { return (SUB ? tile / sub_r : tile) * F; };
    // lanes 0..7 of a slot's first warp, once per tile: where the 8 frames of tile `tl` start (always inside the buffer)
    auto compute_positions = [&](int rg, long long tl) {
        const long long xr = tile_xr0(tl) + t;
        ulonglong2 e;
        if constexpr (SUB) {
            e.x = reinterpret_cast<unsigned long long>(p.sub_in + ((size_t)xr * sub_r + (size_t)(tl % sub_r)) * N);
            e.y = (unsigned long long)(N * 8);
        } else {
            const long long p0 = frame_p0(xr);
            const unsigned long long off = (unsigned long long)p0 * B::SWB, a0 = off & ~15ull;
            e.x = reinterpret_cast<unsigned long long>(p.buf + a0);
            e.y = (((off - a0) + (unsigned long long)N * B::SWB + 15) & ~15ull) | ((off - a0) << 32);
        }
        s_pos[rg * F + t] = e;
    };
    // t == 0 of a slot: start the bulk copy of frame fi of the tile in ring slot rg and publish its misalignment
    auto stage = [&](int rg, int fi, int par) {
        const ulonglong2 e = s_pos[rg * F + fi];
        s_off[par] = (int)(e.y >> 32);
        tma_load_1d(raw, reinterpret_cast<const void *>(e.x), (unsigned)e.y, mbar);
    };

    if (t == 0) {
        mbar_init(mbar, 1);
        s_next[0] = atomicAdd(tile_counter, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long tile = s_next[0];
    if constexpr (B::STAGE) {
        if (t < 32 && tile < p.ntiles) {
            if (t < F) compute_positions(0, tile);
            __syncwarp();
            if (t == 0) stage(0, 0, 0);
        }
    }

    // per-frame min/max of the previous tile of this slot, folded across its 8 warps and converted to dB
    auto publish_minmax = [&](long long pxr0, int pring) {
        if (t < F) {
            const long long xl = p.chunk_first + pxr0 + t;
            unsigned umn = 0x7f800000u, umx = 0u;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint2 m = s_mm[(pring * F + t) * 8 + w];
                umn = min(umn, m.x); umx = max(umx, m.y);
            }
            const float mn = fminf(0.0f, fmaf(fast_log2(__uint_as_float(umn)), p.c1, p.c0));       // lib/worker.js:82,102
            const float mx = fmaxf(-200.0f, fmaf(fast_log2(__uint_as_float(umx)), p.c1, p.c0));    // lib/worker.js:83,103
            if constexpr (SUB) {
                atomicMin(reinterpret_cast<unsigned *>(p.fmin) + xl, f2ord(mn));
                atomicMax(reinterpret_cast<unsigned *>(p.fmax) + xl, f2ord(mx));
            } else { p.fmin[xl] = mn; p.fmax[xl] = mx; }
        }
    };
    long long prev_xr0 = -1;
    int ring = 0;
    unsigned fpar = 0;                     // parity of the frame counter of this slot (mbarrier phase, s_off slot)

    while (tile < p.ntiles) {
        unsigned fetched = 0;
        const int k0sub = SUB ? (int)(tile % sub_r) : 0;
        const long long xr0 = tile_xr0(tile);
        unsigned glo[16], ghi[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { glo[i] = 0; ghi[i] = 0; }

#pragma unroll 1
        for (int f = 0; f < F; f++) {
            cf v[16];
            // ---------------- load + decode + window (lib/worker.js:70-75) ----------------
            if constexpr (SUB) {
                mbar_wait(mbar, fpar);
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = cld(reinterpret_cast<const float2 *>(raw) + T * a + t);
            } else {
                if constexpr (B::STAGE) {
                    mbar_wait(mbar, fpar);
                    const unsigned char *rp = raw + s_off[fpar];
#pragma unroll
                    for (int a = 0; a < 16; a++) v[a] = cpk(decode_raw<FMT>(rp, T * a + t, p.format));
                } else {
                    const long long p0 = frame_p0(xr0 + f);
#pragma unroll
                    for (int a = 0; a < 16; a++) v[a] = cpk(decode_raw<FMT>(p.buf, p0 + T * a + t, p.format));
                }
                // raw sample at p0 + n/2 (lib/worker.js:131-133); the power-of-two scale is exact
                if (t == 0) p.fmid[p.chunk_first + xr0 + f] = make_float2(cre(v[8]) * raw_scale<FMT>(), cim(v[8]) * raw_scale<FMT>());
                const float4 *wrow = reinterpret_cast<const float4 *>(s_win + t * B::WIN_PITCH);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 w = wrow[q];
                    v[4 * q] = cscale(v[4 * q], w.x);         v[4 * q + 1] = cscale(v[4 * q + 1], w.y);
                    v[4 * q + 2] = cscale(v[4 * q + 2], w.z); v[4 * q + 3] = cscale(v[4 * q + 3], w.w);
                }
            }
            fpar ^= 1;

            // ---------------- pass A ----------------
            dft<16>(v);
            {
                const float4 *tw = reinterpret_cast<const float4 *>(s_twA + t * B::TW_PITCH);      // W_4096^{t*k}
                twiddle16_from6(v, tw[0], tw[1], tw[2]);
            }
            float2 *X = X0 + (B::DBX ? (int)fpar * B::X_FLOAT2 : 0);
            // bookkeeping of the slot's first warp, behind the first barrier of the frame (raw is consumed by then):
            // ask for the tile after this one early, park the answer one frame later, lay out its frame positions
            // another frame later, and stage the next frame of this slot while this one is transformed
            auto bookkeeping = [&]() {
                if (t < 32) {
                    if (t == 0) {
                        if (f == 0) fetched = atomicAdd(tile_counter, 1u);
                        if (f == 1) s_next[ring ^ 1] = fetched;
                    }
                    if constexpr (B::STAGE) {
                        if (f == 2) {
                            const long long nt = (long long)__shfl_sync(0xffffffffu, fetched, 0);
                            if (t < F && nt < p.ntiles) compute_positions(ring ^ 1, nt);
                        }
                        if (t == 0) {
                            if (f < F - 1) stage(ring, f + 1, fpar);
                            else if ((long long)fetched < p.ntiles) stage(ring ^ 1, 0, fpar);
                        }
                    }
                }
                if (f == 0 && prev_xr0 >= 0) publish_minmax(prev_xr0, ring ^ 1);
            };
            if constexpr (!B::DBX) {
                slot_barrier(slot);      // X is free (previous frame's pass-C reads) and raw is consumed
                bookkeeping();
            }
            {   // element (k, a1 = t/16, b1 = t%16)
                float2 *dst = X + B::P1 * k0p + lo4;
#pragma unroll
                for (int k = 0; k < 16; k++) cst(dst + k * B::PA, v[k]);
            }
            slot_barrier(slot);
            if constexpr (B::DBX) bookkeeping();
            // ---------------- pass B: thread (k0, b1), in place ----------------
            {
                float2 *col = X + k0p * B::PA + lo4;
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = cld(col + B::P1 * a);
                dft<16>(v);
                const float4 *tw = reinterpret_cast<const float4 *>(s_twB + lo4 * B::TW_PITCH);    // W_256^{b1*k}
                twiddle16_from6(v, tw[0], tw[1], tw[2]);
#pragma unroll
                for (int k = 0; k < 16; k++) cst(col + B::P1 * k, v[k]);
            }
            slot_barrier(slot);
            // ---------------- pass C: thread (k0, k1) ----------------
            {
                const float2 *row = X + k0p * B::PA + B::P1 * lo4;
#pragma unroll
                for (int b = 0; b < 16; b++) v[b] = cld(row + b);
                dft<16>(v);                  // v[k2] is bin binbase + 256*k2
            }

            // ---------------- per-bin epilogue (lib/worker.js:85-122) ----------------
            float amin = __int_as_float(0x7f800000), amax = 0.0f, prev = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float2 vi = cun(v[i]);
                const float abs2 = fmaf(vi.x, vi.x, vi.y * vi.y);
                if (i & 1) {                                 // 3-input min / max: one FMNMX3 per two bins (NaN never wins)
                    amin = fmin3(amin, prev, abs2);
                    amax = fmax3(amax, prev, abs2);
                } else prev = abs2;
                const float l2 = fast_log2(abs2);
                const float2 kg = cun(cfma2(cpk(l2, l2), l2c_kg, l2c_kg0));              // one FFMA2: (raw cB index, colour index)
                unsigned cr = min(__float2uint_rz(kg.x), (unsigned)CB_RAW_CAP);          // :105-106
                if constexpr (SUB || FMT == CF32 || FMT == CF64 || FMT == FMT_RUNTIME)
                    if (!(fabsf(kg.x) <= 3.0e9f)) cr = 1;                                 // NaN / -inf -> bin 0
                const unsigned g = min(__float2uint_rz(kg.y), cmax_u);                   // :111-112
                red_shared_inc(&s_cb[cr]);
                red_shared_inc(&s_cnt[g]);                                               // :113
                // byte f of (ghi:glo) = colour index of frame f of this tile
                glo[i] = __funnelshift_r(glo[i], ghi[i], 8);
                ghi[i] = __funnelshift_r(ghi[i], g, 8);
            }
            // |X|^2 >= 0 (or NaN, which fmin3 / fmax3 drop): the bit patterns order like the values
            const unsigned umn = __reduce_min_sync(0xffffffffu, __float_as_uint(amin));
            const unsigned umx = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
            if ((t & 31) == 0) s_mm[(ring * F + f) * 8 + (t >> 5)] = make_uint2(umn, umx);
        } // frames

        // ---------------- row stores: 8 frames x RGBA = one 32-byte sector per bin ----------------
        {
            const size_t x0 = (size_t)(p.chunk_first + xr0);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int kk = binbase + 256 * i;
                const int bin = SUB ? k0sub + sub_r * kk : kk;
                const int y = (nfull / 2 - bin) & (nfull - 1);                             // lib/worker.js:90
                uint32_t *row = reinterpret_cast<uint32_t *>(p.image) + (size_t)p.nframes * (size_t)y + x0;   // :117
                uint4 a, b;
                a.x = s_lut[glo[i] & 255]; a.y = s_lut[(glo[i] >> 8) & 255]; a.z = s_lut[(glo[i] >> 16) & 255]; a.w = s_lut[glo[i] >> 24];
                b.x = s_lut[ghi[i] & 255]; b.y = s_lut[(ghi[i] >> 8) & 255]; b.z = s_lut[(ghi[i] >> 16) & 255]; b.w = s_lut[ghi[i] >> 24];
                st_global_256(row, a, b);
            }
        }
        prev_xr0 = xr0;
        tile = s_next[ring ^ 1];          // written by t == 0 during frame 1, several slot barriers ago
        ring ^= 1;
    } // tiles

    if (prev_xr0 >= 0) {                  // min/max of the slot's last tile
        slot_barrier(slot);
        publish_minmax(prev_xr0, ring ^ 1);
    }
    __syncthreads();
    for (int i = tid; i < CB_RAW; i += B::THREADS) {
        const unsigned c = s_cb[i];
        if (c) {
            const int b = cb_bin_of_raw(i);
            if (b >= 0) atomicAdd(&p.cb_hist[b], (unsigned long long)c);
        }
    }
    for (int i = tid; i < p.cmap_len; i += B::THREADS)
        if (s_cnt[i]) atomicAdd(&p.c_hist[i], (unsigned long long)s_cnt[i]);
}

} // namespace sp
