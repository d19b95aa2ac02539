// sp_kernel_big.cuh — the latency-hiding variant of the fused render kernel for N = 4096
// (the headline size, and the second stage of the four-step path for N = 8192..65536).
//
// One persistent CTA per SM with THREE independent frame slots of 256 threads (24 warps/SM):
//   * each slot transforms one frame at a time (16 points per thread, 16 x 16 x 16 passes) and
//     synchronises only with itself through a named barrier, so the slots drift out of phase and
//     one slot's shared-memory / barrier latency is covered by the other two slots' FMA work;
//   * raw sample bytes of the NEXT frame are staged into shared memory by one TMA bulk copy
//     (cp.async.bulk + mbarrier) issued as soon as the current frame's samples are in registers,
//     so no thread ever waits on a global load in steady state;
//   * pass-A / pass-B twiddles and the fp32 window live in shared memory once per SM;
//   * ONE padded exchange buffer per slot is reused in place by all three exchanges
//     (element (k0, a, b) at k0*272 + 17*a + b: every access pattern is bank-conflict free);
//   * colour indices of 4 consecutive frames are packed in registers and each image row is written
//     as an aligned 16-byte segment (STG.128) — the transposed store of lib/worker.js:117;
//   * tiles (4 frames) are handed out dynamically through a global counter (no tail imbalance).
// Replaces the hot loops of reference lib/worker.js:68-137 (+ lib/samples.js:313-400,
// lib/fft_nayuki.js:54-96) for the spectrogram layout with cmap_len <= 256 and width % 4 == 0;
// everything else (waterfall, split-real, dB tap, odd widths) runs through render_kernel.
#pragma once
#include "sp_kernels.cuh"

namespace sp {

template <int FMT, int SLOTS_, int F_> struct BigCfg {
    static constexpr int N = 4096, T = 256, SLOTS = SLOTS_, THREADS = T * SLOTS, F = F_;   // F = 4 or 8 frames per tile
    static constexpr int SWB = sample_width(FMT == FMT_RUNTIME ? CF64 : FMT);
    static constexpr bool STAGE = (FMT != FMT_RUNTIME) && SWB <= 4;      // TMA-staged input
    static constexpr int RAW_BYTES = STAGE ? N * SWB + 32 : 0;
    static constexpr int PA = 272, P1 = 17;                              // exchange pitches (float2)
    static constexpr int X_FLOAT2 = 16 * PA;
};

__host__ __device__ inline size_t big_smem_bytes(int raw_bytes, int cmap_len, int slots, int f)
{
    return (size_t)15 * 256 * 8 + 4096 * 4 + (size_t)CB_RAW * 4 + (size_t)cmap_len * 8 + 15 * 16 * 8   // shared tables
         + (size_t)slots * ((size_t)16 * 272 * 8 + (size_t)raw_bytes + 2 * (size_t)f * 8 * 8 + 64);     // per slot
}

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
// one thread: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of dst vs the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int FMT, int SLOTS, int FT>
__global__ void __launch_bounds__(256 * SLOTS, 1) render_big_kernel(const Params p, unsigned *__restrict__ tile_counter)
{
    using B = BigCfg<FMT, SLOTS, FT>;
    constexpr int N = B::N, T = B::T, F = B::F;
    extern __shared__ __align__(128) unsigned char smem_big[];
    // ---- shared by the three slots
    float2 *s_twA = reinterpret_cast<float2 *>(smem_big);                        // [15][256]
    float *s_win = reinterpret_cast<float *>(s_twA + 15 * 256);                  // [4096]
    unsigned *s_cb = reinterpret_cast<unsigned *>(s_win + N);                    // [CB_RAW]
    uint2 *s_col = reinterpret_cast<uint2 *>(s_cb + CB_RAW);                     // [cmap_len] {RGBA, count}
    float2 *s_twB = reinterpret_cast<float2 *>(s_col + p.cmap_len);              // [15][16]
    unsigned char *slot_base = reinterpret_cast<unsigned char *>(s_twB + 15 * 16);
    slot_base += (16 - (reinterpret_cast<uintptr_t>(slot_base) & 15)) & 15;
    constexpr size_t SLOT_BYTES = (size_t)B::X_FLOAT2 * 8 + B::RAW_BYTES + 2 * F * 8 * 8 + 48;

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid % T;
    unsigned char *my = slot_base + (size_t)slot * SLOT_BYTES;
    float2 *X = reinterpret_cast<float2 *>(my);                                  // [16][272] in-place exchange
    unsigned char *raw = my + (size_t)B::X_FLOAT2 * 8;                           // [RAW_BYTES] TMA destination
    float2 *s_mm = reinterpret_cast<float2 *>(raw + B::RAW_BYTES);               // [2][F][8] per-warp min/max
    uint64_t *mbar = reinterpret_cast<uint64_t *>(s_mm + 2 * F * 8);
    long long *s_tile = reinterpret_cast<long long *>(mbar + 1);                 // [2] tile ring

    for (int i = tid; i < CB_RAW; i += B::THREADS) s_cb[i] = 0;
    for (int i = tid; i < p.cmap_len; i += B::THREADS) s_col[i] = make_uint2(p.lut[i], 0u);
    for (int i = tid; i < 15 * 256; i += B::THREADS) s_twA[i] = p.twA[i];
    for (int i = tid; i < 15 * 16; i += B::THREADS) s_twB[i] = p.twB[i];
    const bool sub = p.sub_r > 1;
    for (int i = tid; i < N; i += B::THREADS) s_win[i] = sub ? 1.0f : p.window[i];
    if (t == 0) {
        mbar_init(mbar, 1);
        s_tile[0] = (long long)atomicAdd(tile_counter, 1u);
        s_tile[1] = (long long)atomicAdd(tile_counter, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int nfull = p.n_full, sub_r = p.sub_r;
    const unsigned cmax_u = (unsigned)(p.cmap_len - 1);
    const float gc5 = p.gc + 0.5f;
    const float l2c_k = -10.0f * p.c1, l2c_k0 = fmaf(-10.0f, p.c0, 2.5f);       // raw cB index = l2c_k * log2|X|^2 + l2c_k0
    const float l2c_g = p.gn * p.c1, l2c_g0 = fmaf(p.gn, p.c0, gc5);             // colour index  = l2c_g * log2|X|^2 + l2c_g0
    const int k0p = t >> 4, lo4 = t & 15;                                        // pass B: (k0, b1); pass C: (k0, k1)
    const int binbase = k0p + 16 * lo4;                                          // bins binbase + 256*k2
    __syncthreads();

    // frame position of chunk-relative frame xr (clamped to the last frame of the chunk)
    auto frame_p0 = [&](long long xr) -> long long {
        const long long xgl = p.frame_first + p.chunk_first + xr;
        return (long long)__dadd_rn(0.5, __dmul_rn(p.stride, (double)xgl)) - p.sample_base;   // lib/worker.js:72
    };
    auto frame_inside = [&](long long p0) -> bool {
        return p0 >= 0 && (unsigned long long)(p0 + N) * (unsigned)sample_width(FMT == FMT_RUNTIME ? p.format : FMT) <= p.valid_bytes;
    };
    // tile -> first chunk-relative frame and sub-sequence
    auto tile_xr0 = [&](long long tile) -> long long { return (sub ? tile / sub_r : tile) * F; };

    unsigned parity = 0;
    long long tile = s_tile[0];
    int ring = 0;
    // prologue: stage the first frame of the first tile
    if constexpr (B::STAGE) {
        if (t == 0 && tile < p.ntiles && !sub) {
            long long xr = tile_xr0(tile);
            if (xr > p.chunk_frames - 1) xr = p.chunk_frames - 1;
            const long long p0 = frame_p0(xr);
            if (frame_inside(p0)) {
                const unsigned long long off = (unsigned long long)p0 * B::SWB, a0 = off & ~15ull;
                tma_load_1d(raw, p.buf + a0, (unsigned)(((off - a0) + (unsigned long long)N * B::SWB + 15) & ~15ull), mbar);
            }
        }
    }

    // per-frame min/max of the previous tile of this slot, folded across its 8 warps (runs one
    // tile late, behind a barrier that exists anyway, so no extra synchronisation per tile)
    auto publish_minmax = [&](long long pxr0, int pring) {
        if (t < F) {
            const long long xr = pxr0 + t;
            if (xr < p.chunk_frames) {
                const long long xl = p.chunk_first + xr;
                float mn = 0.0f, mx = -200.0f;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const float2 m = s_mm[(pring * F + t) * 8 + w];
                    mn = fminf(mn, m.x); mx = fmaxf(mx, m.y);
                }
                if (sub) {
                    atomicMin(reinterpret_cast<unsigned *>(p.fmin) + xl, f2ord(mn));
                    atomicMax(reinterpret_cast<unsigned *>(p.fmax) + xl, f2ord(mx));
                } else { p.fmin[xl] = mn; p.fmax[xl] = mx; }
            }
        }
    };
    long long prev_xr0 = -1;

    while (tile < p.ntiles) {
        const long long next_tile = s_tile[ring ^ 1];
        // thread 0 asks for the tile after next right away; the answer is parked in the ring during the last frame
        unsigned fetched = 0;
        if (t == 0) fetched = atomicAdd(tile_counter, 1u);
        const int k0sub = sub ? (int)(tile % sub_r) : 0;
        const long long xr0 = tile_xr0(tile);
        unsigned acc[16][F / 4];
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int h = 0; h < F / 4; h++) acc[i][h] = 0;

#pragma unroll 1
        for (int f = 0; f < F; f++) {
            const bool active = xr0 + f < p.chunk_frames;
            const long long xr = active ? xr0 + f : p.chunk_frames - 1;
            const long long xl = p.chunk_first + xr;
            float2 v[16];
            // ---------------- load + decode + window (lib/worker.js:70-75) ----------------
            if (sub) {
                const float2 *src = p.sub_in + ((size_t)xr * sub_r + k0sub) * N;
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = __ldg(src + T * a + t);
            } else {
                const long long p0 = frame_p0(xr);
                const bool inside = frame_inside(p0);
                if (inside) {
                    if constexpr (B::STAGE) {
                        mbar_wait(mbar, parity);
                        parity ^= 1;
                        const unsigned char *rp = raw + (((unsigned long long)p0 * B::SWB) & 15ull);
#pragma unroll
                        for (int a = 0; a < 16; a++) v[a] = decode_raw<FMT>(rp, T * a + t, p.format);
                    } else {
#pragma unroll
                        for (int a = 0; a < 16; a++) v[a] = decode_raw<FMT>(p.buf, p0 + T * a + t, p.format);
                    }
                } else {
                    const float inv = 1.0f / raw_scale<FMT>();
#pragma unroll
                    for (int a = 0; a < 16; a++) {
                        v[a] = decode_checked(p.buf, p0 + T * a + t, p.format, p.valid_bytes);
                        v[a].x *= inv; v[a].y *= inv;
                    }
                }
                if (t == 0 && active) p.fmid[xl] = make_float2(v[8].x * raw_scale<FMT>(), v[8].y * raw_scale<FMT>());
#pragma unroll
                for (int a = 0; a < 16; a++) { const float w = s_win[T * a + t]; v[a].x *= w; v[a].y *= w; }
            }

            // ---------------- pass A ----------------
            dft<16>(v);
#pragma unroll
            for (int k = 1; k < 16; k++) v[k] = cmul(v[k], s_twA[(k - 1) * T + t]);      // W_4096^{t*k}
            if (t == 0 && f == F - 1) s_tile[ring] = (long long)fetched;                 // ring slot of `tile` is free
            slot_barrier(slot);          // X is free (previous frame's pass-C reads) and raw is consumed
            if constexpr (B::STAGE) {
                if (t == 0 && !sub) {    // stage the next frame of this slot while this one is transformed
                    const bool last = (f == F - 1);
                    const long long ntile = last ? next_tile : tile;
                    if (ntile < p.ntiles) {
                        long long nxr = last ? tile_xr0(ntile) : xr0 + f + 1;
                        if (nxr > p.chunk_frames - 1) nxr = p.chunk_frames - 1;
                        const long long np0 = frame_p0(nxr);
                        if (frame_inside(np0)) {
                            const unsigned long long off = (unsigned long long)np0 * B::SWB, a0 = off & ~15ull;
                            tma_load_1d(raw, p.buf + a0, (unsigned)(((off - a0) + (unsigned long long)N * B::SWB + 15) & ~15ull), mbar);
                        }
                    }
                }
            }
            if (f == 0 && prev_xr0 >= 0) publish_minmax(prev_xr0, ring ^ 1);
            {   // element (k, a1 = t/16, b1 = t%16)
                float2 *dst = X + B::P1 * k0p + lo4;
#pragma unroll
                for (int k = 0; k < 16; k++) dst[k * B::PA] = v[k];
            }
            slot_barrier(slot);
            // ---------------- pass B: thread (k0, b1), in place ----------------
            {
                float2 *col = X + k0p * B::PA + lo4;
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = col[B::P1 * a];
                dft<16>(v);
#pragma unroll
                for (int k = 1; k < 16; k++) v[k] = cmul(v[k], s_twB[(k - 1) * 16 + lo4]);   // W_256^{b1*k}
#pragma unroll
                for (int k = 0; k < 16; k++) col[B::P1 * k] = v[k];
            }
            slot_barrier(slot);
            // ---------------- pass C: thread (k0, k1) ----------------
            {
                const float2 *row = X + k0p * B::PA + B::P1 * lo4;
#pragma unroll
                for (int b = 0; b < 16; b++) v[b] = row[b];
                dft<16>(v);                  // v[k2] is bin binbase + 256*k2
            }

            // ---------------- per-bin epilogue (lib/worker.js:85-122) ----------------
            float amin = __int_as_float(0x7f800000), amax = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float abs2 = fmaf(v[i].x, v[i].x, v[i].y * v[i].y);
                if (i & 1) {                                 // 3-input min / max: one FMNMX3 per two bins
                    const float prev = v[i - 1].x;           // (|X|^2 of bin i-1 parked in its dead re slot)
                    amin = fmin3(amin, prev, abs2);
                    amax = fmax3(amax, prev, abs2);
                } else v[i].x = abs2;
                const float l2 = fast_log2(abs2);
                const float kf = fmaf(l2, l2c_k, l2c_k0);
                unsigned cr = min(__float2uint_rz(kf), (unsigned)CB_RAW_CAP);            // :105-106
                if constexpr (FMT == CF32 || FMT == CF64 || FMT == FMT_RUNTIME)
                    if (!(fabsf(kf) <= 3.0e9f)) cr = 1;                                   // NaN / -inf -> bin 0
                const unsigned g = min(__float2uint_rz(fmaf(l2, l2c_g, l2c_g0)), cmax_u);  // :111-112
                if (active) {
                    atomicAdd(&s_cb[cr], 1u);
                    atomicAdd(&s_col[g].y, 1u);                                           // :113
                }
                // byte f of acc[i] (little endian over the F/4 words) = colour index of frame f
                if constexpr (F == 8) {
                    acc[i][0] = __funnelshift_r(acc[i][0], acc[i][1], 8);
                    acc[i][1] = __funnelshift_r(acc[i][1], g, 8);
                } else {
                    acc[i][0] = __funnelshift_r(acc[i][0], g, 8);
                }
            }
            float mn = fminf(0.0f, fmaf(fast_log2(amin), p.c1, p.c0));                       // :82,102
            float mx = fmaxf(-200.0f, fmaf(fast_log2(amax), p.c1, p.c0));                    // :83,103
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            }
            if ((t & 31) == 0) s_mm[(ring * F + f) * 8 + (t >> 5)] = make_float2(mn, mx);
        } // frames

        // ---------------- row stores: F frames x 4 bytes = aligned 16 / 32-byte segment per bin ----------------
        if (xr0 < p.chunk_frames) {
            const size_t x0 = (size_t)(p.chunk_first + xr0);
            const bool second = (F == 8) && (xr0 + 8 <= p.chunk_frames);                   // else width % 4 == 0: 4 frames left
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int kk = binbase + 256 * i;
                const int bin = sub ? k0sub + sub_r * kk : kk;
                const int y = (nfull / 2 - bin) & (nfull - 1);                             // lib/worker.js:90
                uint4 *row = reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(p.image) + (size_t)p.nframes * (size_t)y + x0);   // :117
                uint4 a;
                a.x = s_col[acc[i][0] & 255].x; a.y = s_col[(acc[i][0] >> 8) & 255].x;
                a.z = s_col[(acc[i][0] >> 16) & 255].x; a.w = s_col[acc[i][0] >> 24].x;
                row[0] = a;
                if constexpr (F == 8) {
                    if (second) {
                        a.x = s_col[acc[i][1] & 255].x; a.y = s_col[(acc[i][1] >> 8) & 255].x;
                        a.z = s_col[(acc[i][1] >> 16) & 255].x; a.w = s_col[acc[i][1] >> 24].x;
                        row[1] = a;
                    }
                }
            }
        }
        prev_xr0 = xr0;
        tile = next_tile;
        ring ^= 1;
    } // tiles

    if (prev_xr0 >= 0) {          // min/max of the slot's last tile
        slot_barrier(slot);
        publish_minmax(prev_xr0, ring ^ 1);
    }
    __syncthreads();
    flush_hist(s_cb, s_col, p, tid, B::THREADS, false);
}

} // namespace sp
