// sp_kernels.cuh — the fused render kernel: decode + window + FFT + dB + colormap + RGBA
// + histograms + per-frame min/max in ONE pass (replaces the hot loops of reference
// lib/worker.js:68-137 together with lib/samples.js:313-400 and lib/fft_nayuki.js:54-119).
//
// Work decomposition (N = FFT size):
//   * a "slot" of T = N/16 threads owns one frame at a time; each thread keeps 16 complex
//     points in registers.  The FFT is a decimation-in-frequency 16 x 16 x R transform
//     (1, 2 or 3 register passes) with shared-memory exchanges in padded, bank-conflict-free
//     layouts; twiddles are expanded from 4 table loads per pass.
//   * a slot renders 8 CONSECUTIVE frames and keeps their colour indices packed in registers
//     (16 bins x 8 bytes), so each image row is written as one aligned 32-byte segment
//     (2 x STG.128) — the transposed store of lib/worker.js:117 without a shared-memory tile.
//   * a CTA of 256 threads is persistent: it loops over tiles of 8*SLOTS frames and keeps the
//     two histograms in shared memory, flushing them to 64-bit global counters once.
#pragma once
#include "sp_device.cuh"

namespace sp {

struct Params {
    // input
    const uint8_t *buf;            // shard bytes; sample `sample_base` is at buf[0]
    unsigned long long valid_bytes;
    long long sample_base;
    int format;                    // runtime format (used by FMT_RUNTIME kernels and slow paths)
    int n_full;                    // FFT size of the message (n)
    double stride;                 // global fractional hop (lib/worker.js:50)
    long long frame_first;         // global index of local frame 0
    long long nframes;             // frames rendered by this launch (image width in pixels)
    const float *window;           // n_full fp32 coefficients (rounded once from double)
    const float2 *tw;              // twiddle table exp(-2 pi j i / n_kernel), i < n_kernel
    // dB / colour mapping
    float c0, c1;                  // d0 = c1 * log2(|X|^2) + c0
    float gn, gc, cmaxf;           // gray = trunc(0.5 + clamp(gn * d0 + gc, 0, cmax))
    int cmap_len;
    const uint32_t *lut;           // cmap_len packed RGBA words
    // output
    uint8_t *image;                // may be null
    int waterfall, channel_mode;
    float *fmin, *fmax;            // per local frame: min / max of d0 (reference init 0 / -200)
    float2 *fmid;                  // per local frame: raw mid sample (amp gauge)
    unsigned long long *cb_hist, *c_hist;
    float *db_out;                 // optional [nframes][n_full] d0 tap
    // sub-frame mode (n_full > kernel N): the kernel transforms n_full/sub_r point
    // sub-sequences produced by the radix-sub_r pre-pass; sub-frame (x, k0) yields the
    // bins k0 + sub_r*k'.  sub_r == 1 is the ordinary mode.
    int sub_r;
    const float2 *sub_in;          // [chunk frames][sub_r][N] complex, frame-major
    long long chunk_first;         // first local frame of this launch (multiple of 8)
    long long chunk_frames;        // frames rendered by this launch (== nframes when not chunked)
    long long ntiles;
};

template <int LOG2N> struct Cfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int P = N >= 16 ? 16 : N;          // points per thread
    static constexpr int T = N / P;                     // threads per frame
    static constexpr int PASSES = LOG2N <= 4 ? 1 : (LOG2N <= 8 ? 2 : 3);
    static constexpr int RL = PASSES == 1 ? P : (PASSES == 2 ? N / 16 : N / 256);   // last radix
    static constexpr int NB = P / RL;                   // butterflies per thread in the last pass
    static constexpr int THREADS = 256;
    static constexpr int SLOTS = THREADS / T;
    static constexpr int FRAMES_PER_SLOT = 8;
    static constexpr int TILE = SLOTS * FRAMES_PER_SLOT;
    // exchange layouts in float2 units (see DESIGN.md "shared-memory layouts")
    static constexpr int N1 = N / 16;
    static constexpr int PITCH_A = PASSES == 3 ? N1 + (RL % 16) : (T + 1);
    static constexpr int SIZE_A = PASSES >= 2 ? 16 * PITCH_A + (PASSES == 2 ? 1 : 0) : 0;
    static constexpr int P1 = RL + 1;
    static constexpr int PITCH_B = 16 * P1 + (RL % 16);
    static constexpr int SIZE_B = PASSES == 3 ? 16 * PITCH_B : 0;
    static constexpr int WPS = T > 32 ? T / 32 : 1;     // warps per slot
    static constexpr int SMEM_X = SLOTS * (SIZE_A + SIZE_B) > N * SLOTS ? SLOTS * (SIZE_A + SIZE_B) : N * SLOTS;
};

constexpr int CB_BINS = 1000;     // lib/worker.js:41

__host__ __device__ inline size_t main_smem_bytes(int smem_x_float2, int cmap_len)
{
    return (size_t)smem_x_float2 * 8 + (size_t)(CB_BINS + cmap_len) * 4 + (size_t)cmap_len * 4 + 64 * 8;
}

// first output bin of butterfly j of thread t in the last pass (bin = kbase + kstep * k)
template <class C> __device__ __forceinline__ int kbase_of(int t, int j)
{
    if constexpr (C::PASSES == 1) return 0;
    else if constexpr (C::PASSES == 2) return t + C::T * j;
    else { const int q = t + C::T * j; return (q / 16) + 16 * (q % 16); }
}

// order-preserving float <-> uint mapping for atomic min / max on floats
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <int LOG2N, int FMT>
__global__ void __launch_bounds__(256, 2) render_kernel(const Params p)
{
    using C = Cfg<LOG2N>;
    constexpr int N = C::N, P = C::P, T = C::T;
    constexpr int kstep = C::PASSES == 1 ? 1 : (C::PASSES == 2 ? 16 : 256);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *xbuf = reinterpret_cast<float2 *>(smem_raw);
    unsigned *s_cb = reinterpret_cast<unsigned *>(xbuf + C::SMEM_X);
    unsigned *s_c = s_cb + CB_BINS;
    unsigned *s_lut = s_c + p.cmap_len;
    float2 *s_mm = reinterpret_cast<float2 *>(s_lut + p.cmap_len);   // [64] per-warp min/max partials

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid % T;
    float2 *bufA = xbuf + slot * (C::SIZE_A + C::SIZE_B);
    float2 *bufB = bufA + C::SIZE_A;
    (void)bufA; (void)bufB;

    for (int i = tid; i < CB_BINS + p.cmap_len; i += C::THREADS) s_cb[i] = 0;
    for (int i = tid; i < p.cmap_len; i += C::THREADS) s_lut[i] = p.lut[i];

    const bool sub = p.sub_r > 1;
    const int nfull = p.n_full;
    const int sub_r = p.sub_r;

    // thread-invariant window coefficients (element T*a + t of the frame); ones in sub-frame mode
    float win[P];
#pragma unroll
    for (int a = 0; a < P; a++) win[a] = sub ? 1.0f : p.window[T * a + t];
    int kbase[C::NB];
#pragma unroll
    for (int j = 0; j < C::NB; j++) kbase[j] = kbase_of<C>(t, j);

    const bool wide_rows = !p.waterfall && p.image && p.cmap_len <= 256 && (p.nframes % 4 == 0)
                           && ((reinterpret_cast<uintptr_t>(p.image) & 15) == 0);
    const bool direct = p.image && !wide_rows;     // per-pixel 4-byte stores
    unsigned long long px_since_flush = 0;
    __syncthreads();

    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        // tile -> (first chunk-relative frame of this slot, sub-sequence index)
        long long xg; int k0sub = 0;
        if (sub) { k0sub = (int)(tile % sub_r); xg = tile / sub_r; } else xg = tile;
        const long long xr0 = xg * C::TILE + (long long)slot * C::FRAMES_PER_SLOT;   // chunk relative
        unsigned glo[P], ghi[P];
#pragma unroll
        for (int i = 0; i < P; i++) { glo[i] = 0; ghi[i] = 0; }

#pragma unroll 1
        for (int f = 0; f < C::FRAMES_PER_SLOT; f++) {
            const bool active = xr0 + f < p.chunk_frames;
            const long long xr = active ? xr0 + f : p.chunk_frames - 1;   // inactive slots redo the last frame
            const long long xl = p.chunk_first + xr;                      // local frame == image column
            float2 v[P];
            // ---------------- load + decode + window (lib/worker.js:70-75) ----------------
            if (sub) {
                const float2 *src = p.sub_in + ((size_t)xr * sub_r + k0sub) * N;
#pragma unroll
                for (int a = 0; a < P; a++) v[a] = src[T * a + t];
            } else {
                const long long xgl = p.frame_first + xl;
                // ~~(0.5 + stride * x): separate multiply and add like JavaScript (no FMA)
                const long long p0 = (long long)__dadd_rn(0.5, __dmul_rn(p.stride, (double)xgl)) - p.sample_base;
                const bool inside = p0 >= 0 &&
                    (unsigned long long)(p0 + N) * (unsigned)sample_width(FMT == FMT_RUNTIME ? p.format : FMT) <= p.valid_bytes;
                if (inside) {
#pragma unroll
                    for (int a = 0; a < P; a++) v[a] = decode_fast<FMT>(p.buf, p0 + T * a + t, p.format);
                } else {
#pragma unroll
                    for (int a = 0; a < P; a++) v[a] = decode_checked(p.buf, p0 + T * a + t, p.format, p.valid_bytes);
                }
                if (t == 0 && active) p.fmid[xl] = v[P / 2];   // raw sample at p0 + n/2 (lib/worker.js:131-133)
#pragma unroll
                for (int a = 0; a < P; a++) { v[a].x *= win[a]; v[a].y *= win[a]; }
            }

            // ---------------- FFT pass A: radix P over the slowest input digit ----------------
            dft<P>(v);
            if constexpr (C::PASSES >= 2) {
                {
                    const float2 w1 = p.tw[t], w2 = p.tw[2 * t], w4 = p.tw[4 * t], w8 = p.tw[8 * t];
                    twiddle16(v, w1, w2, w4, w8);             // W_N^{t*k0}
                }
#pragma unroll
                for (int k = 0; k < 16; k++) bufA[k * C::PITCH_A + t] = v[k];
                __syncthreads();
            }
            if constexpr (C::PASSES == 3) {
                // ------------ pass B: thread <-> (k0, b1), radix 16 over a1 ------------
                constexpr int R2 = C::RL;
                const int k0 = t / R2, b1 = t % R2;
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = bufA[k0 * C::PITCH_A + R2 * a + b1];
                dft<16>(v);
                {
                    const int e = 16 * b1;                    // W_{N/16}^{b1*k1} = W_N^{16*b1*k1}
                    const float2 w1 = p.tw[e], w2 = p.tw[2 * e], w4 = p.tw[4 * e], w8 = p.tw[8 * e];
                    twiddle16(v, w1, w2, w4, w8);
                }
#pragma unroll
                for (int k = 0; k < 16; k++) bufB[k0 * C::PITCH_B + k * C::P1 + b1] = v[k];
                __syncthreads();
            }
            // ---------------- last pass: NB radix-RL butterflies; v[j*RL + k] is bin kbase[j] + kstep*k ------
            if constexpr (C::PASSES == 2) {
#pragma unroll
                for (int j = 0; j < C::NB; j++) {
                    const int k0 = t + T * j;
                    float2 u[C::RL];
#pragma unroll
                    for (int b = 0; b < C::RL; b++) u[b] = bufA[k0 * C::PITCH_A + b];
                    dft<C::RL>(u);
#pragma unroll
                    for (int b = 0; b < C::RL; b++) v[j * C::RL + b] = u[b];
                }
            } else if constexpr (C::PASSES == 3) {
#pragma unroll
                for (int j = 0; j < C::NB; j++) {
                    const int q = t + T * j, k1 = q % 16, k0 = q / 16;
                    float2 u[C::RL];
#pragma unroll
                    for (int b = 0; b < C::RL; b++) u[b] = bufB[k0 * C::PITCH_B + k1 * C::P1 + b];
                    dft<C::RL>(u);
#pragma unroll
                    for (int b = 0; b < C::RL; b++) v[j * C::RL + b] = u[b];
                }
            }

            // ---------------- optional split-real post-process (lib/fft_nayuki.js:103-119) ----------------
            if (p.channel_mode) {
                // (the engine never combines channel mode with sub-frame mode, see sp_engine.cu)
                __syncthreads();                               // everyone is done reading bufA / bufB
                float2 *X = xbuf + slot * N;
#pragma unroll
                for (int j = 0; j < C::NB; j++)
#pragma unroll
                    for (int k = 0; k < C::RL; k++) X[kbase[j] + kstep * k] = v[j * C::RL + k];
                __syncthreads();
#pragma unroll
                for (int j = 0; j < C::NB; j++)
#pragma unroll
                    for (int k = 0; k < C::RL; k++) {
                        const int bin = kbase[j] + kstep * k;
                        const float2 a = v[j * C::RL + k];
                        const float2 b = X[(N - bin) & (N - 1)];
                        float2 r;
                        if (bin == 0) r = make_float2(a.x, 0.f);
                        else if (bin == N / 2) r = make_float2(0.f, 0.f);
                        else if (bin < N / 2) r = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                        else r = make_float2(0.5f * (b.y + a.y), 0.5f * (-b.x + a.x));
                        v[j * C::RL + k] = r;
                    }
                __syncthreads();                               // X aliases the exchange buffers
            }

            // ---------------- per-bin epilogue (lib/worker.js:85-122) ----------------
            float mn = 0.0f, mx = -200.0f;                     // lib/worker.js:82-83
#pragma unroll
            for (int j = 0; j < C::NB; j++)
#pragma unroll
                for (int k = 0; k < C::RL; k++) {
                    const int i = j * C::RL + k;
                    const float abs2 = fmaf(v[i].x, v[i].x, v[i].y * v[i].y);
                    const float d0 = fmaf(__log2f(abs2), p.c1, p.c0);        // dBfs - gain
                    mn = fminf(mn, d0);
                    mx = fmaxf(mx, d0);
                    int cb = js_trunc(fmaf(d0, -10.0f, 0.5f));                // lib/worker.js:105
                    cb = min(cb, CB_BINS - 1);
                    const float gf = fminf(fmaxf(fmaf(d0, p.gn, p.gc), 0.0f), p.cmaxf);
                    const int g = __float2int_rz(gf + 0.5f);                  // lib/worker.js:112
                    if (active) {
                        if (cb >= 0) atomicAdd(&s_cb[cb], 1u);               // lib/worker.js:106
                        atomicAdd(&s_c[g], 1u);                              // lib/worker.js:113
                    }
                    // byte f of (ghi:glo) = colour index of frame f of this slot
                    glo[i] = __funnelshift_r(glo[i], ghi[i], 8);
                    ghi[i] = __funnelshift_r(ghi[i], (unsigned)g, 8);
                    if (direct | (p.db_out != nullptr)) {
                        const int bin = sub ? k0sub + sub_r * (kbase[j] + kstep * k) : kbase[j] + kstep * k;
                        if (active) {
                            if (p.db_out) p.db_out[(size_t)xl * nfull + bin] = d0;
                            if (direct) {
                                const int y = (nfull / 2 - bin) & (nfull - 1);                 // lib/worker.js:90
                                const size_t px = p.waterfall
                                    ? (size_t)nfull * (size_t)(p.nframes - 1 - xl) + (size_t)(nfull - 1 - y)   // :116
                                    : (size_t)xl + (size_t)p.nframes * (size_t)y;                              // :117
                                reinterpret_cast<uint32_t *>(p.image)[px] = s_lut[g];
                            }
                        }
                    }
                }

            // ---------------- per-frame min / max (lib/worker.js:102-103) ----------------
#pragma unroll
            for (int off = (T < 32 ? T : 32) / 2; off > 0; off >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            }
            if constexpr (T <= 32) {
                if (t == 0 && active) {
                    if (sub) {
                        atomicMin(reinterpret_cast<unsigned *>(p.fmin) + xl, f2ord(mn));
                        atomicMax(reinterpret_cast<unsigned *>(p.fmax) + xl, f2ord(mx));
                    } else { p.fmin[xl] = mn; p.fmax[xl] = mx; }
                }
            } else {
                if ((t & 31) == 0) s_mm[(slot * C::FRAMES_PER_SLOT + f) * C::WPS + t / 32] = make_float2(mn, mx);
            }
        } // frames of the slot

        // ---------------- row stores: 8 frames x 4 bytes = one 32-byte segment per bin ----------------
        if (wide_rows && xr0 < p.chunk_frames) {
            const bool full8 = xr0 + 8 <= p.chunk_frames;
            const size_t x0 = (size_t)(p.chunk_first + xr0);
#pragma unroll
            for (int j = 0; j < C::NB; j++)
#pragma unroll
                for (int k = 0; k < C::RL; k++) {
                    const int i = j * C::RL + k;
                    const int bin = sub ? k0sub + sub_r * (kbase[j] + kstep * k) : kbase[j] + kstep * k;
                    const int y = (nfull / 2 - bin) & (nfull - 1);
                    uint32_t *row = reinterpret_cast<uint32_t *>(p.image) + (size_t)p.nframes * (size_t)y + x0;
                    uint4 a, b;
                    a.x = s_lut[glo[i] & 255]; a.y = s_lut[(glo[i] >> 8) & 255];
                    a.z = s_lut[(glo[i] >> 16) & 255]; a.w = s_lut[glo[i] >> 24];
                    reinterpret_cast<uint4 *>(row)[0] = a;
                    if (full8) {                               // else nframes % 4 == 0: exactly 4 frames left
                        b.x = s_lut[ghi[i] & 255]; b.y = s_lut[(ghi[i] >> 8) & 255];
                        b.z = s_lut[(ghi[i] >> 16) & 255]; b.w = s_lut[ghi[i] >> 24];
                        reinterpret_cast<uint4 *>(row)[1] = b;
                    }
                }
        }

        // ---------------- per-frame min/max across the warps of a slot ----------------
        if constexpr (T > 32) {
            __syncthreads();
            if (tid < C::SLOTS * C::FRAMES_PER_SLOT) {
                const int s = tid / C::FRAMES_PER_SLOT, f = tid % C::FRAMES_PER_SLOT;
                const long long xr = xg * C::TILE + (long long)s * C::FRAMES_PER_SLOT + f;
                if (xr < p.chunk_frames) {
                    const long long xl = p.chunk_first + xr;
                    float mn = 0.0f, mx = -200.0f;
                    for (int w = 0; w < C::WPS; w++) {
                        const float2 m = s_mm[(s * C::FRAMES_PER_SLOT + f) * C::WPS + w];
                        mn = fminf(mn, m.x); mx = fmaxf(mx, m.y);
                    }
                    if (sub) {
                        atomicMin(reinterpret_cast<unsigned *>(p.fmin) + xl, f2ord(mn));
                        atomicMax(reinterpret_cast<unsigned *>(p.fmax) + xl, f2ord(mx));
                    } else { p.fmin[xl] = mn; p.fmax[xl] = mx; }
                }
            }
            // s_mm is next written after the __syncthreads of the next tile's first frame
        }

        // keep the 32-bit shared counters from wrapping on very long captures
        px_since_flush += (unsigned long long)C::TILE * N;
        if (px_since_flush >= (1ull << 31)) {
            __syncthreads();
            for (int i = tid; i < CB_BINS + p.cmap_len; i += C::THREADS) {
                const unsigned c = s_cb[i];
                if (c) {
                    atomicAdd(i < CB_BINS ? &p.cb_hist[i] : &p.c_hist[i - CB_BINS], (unsigned long long)c);
                    s_cb[i] = 0;
                }
            }
            px_since_flush = 0;
            __syncthreads();
        }
    } // tiles

    __syncthreads();
    for (int i = tid; i < CB_BINS + p.cmap_len; i += C::THREADS) {
        const unsigned c = s_cb[i];
        if (c) atomicAdd(i < CB_BINS ? &p.cb_hist[i] : &p.c_hist[i - CB_BINS], (unsigned long long)c);
    }
}

// ---------------------------------------------------------------------------------------------
// Radix-R pre-pass for n_full = R * 4096 (R = 2, 4, 8, 16): decode + window + DFT_R over the
// slowest input digit + twiddle W_n^{b*k0}; writes R sub-sequences of 4096 points per frame,
// which render_kernel<12> then transforms in sub-frame mode (four-step FFT, scratch kept
// L2-resident by chunking the frames).
// ---------------------------------------------------------------------------------------------
template <int R, int FMT>
__global__ void __launch_bounds__(256) prepass_kernel(const Params p, float2 *__restrict__ out, const float2 *__restrict__ tw_full)
{
    constexpr int NS = 4096;
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long xr = gid / NS;                 // chunk-relative frame
    const int b = (int)(gid % NS);
    if (xr >= p.chunk_frames) return;
    const long long xl = p.chunk_first + xr;
    const long long xgl = p.frame_first + xl;
    const long long p0 = (long long)__dadd_rn(0.5, __dmul_rn(p.stride, (double)xgl)) - p.sample_base;
    const bool inside = p0 >= 0 &&
        (unsigned long long)(p0 + (long long)R * NS) * (unsigned)sample_width(FMT == FMT_RUNTIME ? p.format : FMT) <= p.valid_bytes;
    float2 v[R];
    if (inside) {
#pragma unroll
        for (int a = 0; a < R; a++) v[a] = decode_fast<FMT>(p.buf, p0 + NS * a + b, p.format);
    } else {
#pragma unroll
        for (int a = 0; a < R; a++) v[a] = decode_checked(p.buf, p0 + NS * a + b, p.format, p.valid_bytes);
    }
    if (b == 0) p.fmid[xl] = v[R / 2];             // sample p0 + n/2
#pragma unroll
    for (int a = 0; a < R; a++) { const float w = p.window[NS * a + b]; v[a].x *= w; v[a].y *= w; }
    dft<R>(v);
#pragma unroll
    for (int k = 1; k < R; k++) v[k] = cmul(v[k], tw_full[b * k]);
    float2 *dst = out + (size_t)xr * R * NS + b;
#pragma unroll
    for (int k = 0; k < R; k++) dst[(size_t)k * NS] = v[k];
}

} // namespace sp
