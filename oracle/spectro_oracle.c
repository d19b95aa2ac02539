/*
 * spectro_oracle.c — float64 CPU restatement of the spectroplot-js render worker.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product path (spectroplot-js_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the CPU arm.
 *
 * PARITY PIN: the reference (triq-org/spectroplot-js v1.2.1) ships no tests, golden
 * vectors or fixtures for this path, and no JavaScript engine exists in this image.
 * The pin is therefore the reference's own SOURCE executed by oracle/jsmini.py (an
 * ES-subset interpreter written for this purpose): tools/make_ref_golden.py runs the
 * unmodified lib/worker.js (+ samples.js, fft_nayuki.js, windows.js, the cmap modules)
 * on 36 messages and commits the replies under tests/golden/ref_js/;
 * tests/test_reference_js.py requires this file to reproduce every reply EXACTLY
 * (image bytes, both histograms, gauges, dBfs_min / dBfs_max).  Residual caveat: jsmini
 * takes Math.log10 / cos / sin from the platform libm where V8 uses its fdlibm port
 * (<= 1 ulp, visible only at exact quantisation ties).  Further cross-checks:
 * (i) the derived known-answers of SURVEY.md Appendix B, (ii) an independent numpy
 * restatement (oracle/np_restatement.py, different FFT), (iii) mathematical identities
 * (naive DFT, full-scale tone == 0 dB).  This file restates the reference line by line
 * (citations below are reference file:line).
 *
 * All arithmetic is IEEE double like JavaScript numbers.  Build with
 * -ffp-contract=off so no FMA contraction changes the operation order.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#define SPO_CB_HIST 1000

enum { F_CU4, F_CS4, F_CU8, F_CS8, F_CU12, F_CS12, F_CU16, F_CS16,
       F_CU32, F_CS32, F_CU64, F_CS64, F_CF32, F_CF64, F_COUNT };

typedef struct spo_request {
    const uint8_t *buffer;
    uint64_t byte_length;
    int32_t format, n;
    int64_t width;
    double block_norm, gain, range;
    const double *windowc;
    const uint8_t *cmap_rgb;
    int32_t cmap_len, channel_mode, waterfall, pad_;
} spo_request;

typedef struct spo_reply {
    uint8_t *image, *gauge_mins, *gauge_maxs, *gauge_amps;
    uint64_t *cB_hist, *c_hist;
    double dBfs_min, dBfs_max;
    /* taps (optional, may be NULL) */
    double *db;        /* [width][n] dBfs - gain in FFT bin order */
    uint16_t *gray;    /* [width][n] colour index in FFT bin order */
    int32_t *cbk;      /* [width][n] cB bin (-1 == dropped)        */
} spo_reply;

/* ------------------------------------------------------------------ JS number helpers */

/* `~~v`: ToInt32(v) — truncate, wrap modulo 2^32; NaN and +-Infinity give 0. */
static int32_t js_toint32(double v)
{
    if (!isfinite(v)) return 0;
    double t = trunc(v);
    double m = fmod(t, 4294967296.0);
    if (m < 0) m += 4294967296.0;
    if (m >= 2147483648.0) m -= 4294967296.0;
    return (int32_t)m;
}

/* store into a Uint8ClampedArray: clamp to [0,255], round half to even, NaN -> 0 */
static uint8_t js_u8clamped(double v)
{
    if (!(v > 0)) return 0;           /* NaN, -inf, <= 0 */
    if (v >= 255) return 255;
    return (uint8_t)nearbyint(v);     /* default rounding mode: ties to even */
}

/* ------------------------------------------------------------------ formats: lib/samples.js:22-155 */

static const char *const k_names[F_COUNT] = { "CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16",
    "CS16", "CU32", "CS32", "CU64", "CS64", "CF32", "CF64" };
static const int k_width[F_COUNT] = { 1, 1, 2, 2, 3, 3, 4, 4, 8, 8, 16, 16, 8, 16 };
static const int k_elem[F_COUNT]  = { 1, 1, 1, 1, 1, 1, 2, 2, 4, 4, 4, 4, 4, 8 };

int spo_format_from_name(const char *name)
{
    if (!name) return F_CU8;
    for (int i = 0; i < F_COUNT; i++)
        if (!strcasecmp(name, k_names[i])) return i;
    if (!strcasecmp(name, "DATA") || !strcasecmp(name, "COMPLEX16U")) return F_CU8;   /* :48 */
    if (!strcasecmp(name, "COMPLEX16S")) return F_CS8;                                  /* :55 */
    if (!strcasecmp(name, "CFILE") || !strcasecmp(name, "COMPLEX")) return F_CF32;      /* :126 */
    return F_CU8;                                                                       /* :149-155 */
}
int spo_sample_width(int f) { return (f < 0 || f >= F_COUNT) ? -1 : k_width[f]; }
int spo_element_size(int f) { return (f < 0 || f >= F_COUNT) ? -1 : k_elem[f]; }
const char *spo_format_name(int f) { return (f < 0 || f >= F_COUNT) ? "?" : k_names[f]; }

/* A typed-array view over the buffer: reading past `length` yields undefined -> NaN. */
typedef struct { const uint8_t *p; uint64_t nbytes; int fmt; } view_t;

static inline int in_range(const view_t *v, int64_t elem_index, int elem_size)
{
    if (elem_index < 0) return 0;
    uint64_t len = v->nbytes / (uint64_t)elem_size;     /* typed array length */
    return (uint64_t)elem_index < len;
}
static inline uint32_t rd_u32(const uint8_t *p) { uint32_t x; memcpy(&x, p, 4); return x; }

/* sampleI (c = 0) / sampleQ (c = 1) at sample index pos. */
static double sample_iq(const view_t *v, int64_t pos, int c)
{
    const uint8_t *p = v->p;
    switch (v->fmt) {
    case F_CU4: {                                   /* lib/samples.js:313-322 */
        int b0 = in_range(v, pos, 1) ? p[pos] : 0;     /* undefined & 0xf0 == 0 */
        int s = c ? (b0 & 0x0f) : ((b0 & 0xf0) >> 4);
        return (s - 7.5) * (1.0 / 7.5);
    }
    case F_CS4: {                                   /* lib/samples.js:325-334 */
        int b0 = in_range(v, pos, 1) ? p[pos] : 0;
        int s = c ? (b0 & 0x0f) : ((b0 & 0xf0) >> 4);
        if (s & 8) s -= 16;                         /* (x << 28) >> 28 sign extension */
        return s * (1.0 / 8.0);
    }
    case F_CU8: {                                   /* lib/samples.js:48-54,393-400 */
        if (!in_range(v, 2 * pos + c, 1)) return NAN;
        return ((double)p[2 * pos + c] - 127.5) * (1.0 / 127.5);
    }
    case F_CS8: {                                   /* lib/samples.js:55-61 */
        if (!in_range(v, 2 * pos + c, 1)) return NAN;
        return ((double)(int8_t)p[2 * pos + c] - 0) * (1.0 / 128.0);
    }
    case F_CU12: case F_CS12: {                     /* lib/samples.js:337-362 */
        /* I needs bytes 0,1; Q needs bytes 1,2 of the 3-byte group; a missing byte is
         * `undefined`, which the bit operators turn into 0 (ToInt32(undefined) = 0). */
        int64_t b = 3 * pos;
        int b0 = in_range(v, b, 1) ? p[b] : 0;
        int b1 = in_range(v, b + 1, 1) ? p[b + 1] : 0;
        int b2 = in_range(v, b + 2, 1) ? p[b + 2] : 0;
        int s = c ? ((b2 << 4) | ((b1 & 0xf0) >> 4)) : (((b1 & 0x0f) << 8) | b0);
        if (v->fmt == F_CU12) return (s - 2047.5) * (1.0 / 2047.5);
        if (s & 0x800) s -= 4096;
        return s * (1.0 / 2048.0);
    }
    case F_CU16: {                                  /* lib/samples.js:62-68 */
        if (!in_range(v, 2 * pos + c, 2)) return NAN;
        uint16_t x; memcpy(&x, p + 2 * (2 * pos + c), 2);
        return ((double)x - 32767.5) * (1.0 / 32768.0);
    }
    case F_CS16: {                                  /* lib/samples.js:69-75 */
        if (!in_range(v, 2 * pos + c, 2)) return NAN;
        int16_t x; memcpy(&x, p + 2 * (2 * pos + c), 2);
        return ((double)x - 0) * (1.0 / 32768.0);
    }
    case F_CU32: {                                  /* lib/samples.js:94-100 */
        if (!in_range(v, 2 * pos + c, 4)) return NAN;
        return ((double)rd_u32(p + 4 * (2 * pos + c)) - 2147483647.5) * (1.0 / 2147483648.0);
    }
    case F_CS32: {                                  /* lib/samples.js:101-107 */
        if (!in_range(v, 2 * pos + c, 4)) return NAN;
        return ((double)(int32_t)rd_u32(p + 4 * (2 * pos + c)) - 0) * (1.0 / 2147483648.0);
    }
    case F_CU64: case F_CS64: {                     /* lib/samples.js:365-390 (Uint32Array view) */
        int64_t w = 4 * pos + 2 * c;
        double b0 = in_range(v, w, 4) ? (double)rd_u32(p + 4 * w) : NAN;           /* low word  */
        double b1;
        if (!in_range(v, w + 1, 4)) b1 = (v->fmt == F_CS64) ? 0.0 : NAN;           /* undefined>>0 == 0 */
        else b1 = (v->fmt == F_CS64) ? (double)(int32_t)rd_u32(p + 4 * (w + 1))
                                     : (double)rd_u32(p + 4 * (w + 1));
        double s = b1 / 2147483648.0 + (b0 / 18446744073709551616.0);
        return (v->fmt == F_CU64) ? (s - 1.0) : s;
    }
    case F_CF32: {                                  /* lib/samples.js:126-132 */
        if (!in_range(v, 2 * pos + c, 4)) return NAN;
        float x; memcpy(&x, p + 4 * (2 * pos + c), 4);
        return ((double)x - 0) * 1.0;
    }
    case F_CF64: {                                  /* lib/samples.js:133-139 */
        if (!in_range(v, 2 * pos + c, 8)) return NAN;
        double x; memcpy(&x, p + 8 * (2 * pos + c), 8);
        return (x - 0) * 1.0;
    }
    }
    return NAN;
}

/* decode tap: iq[2*count] doubles */
int spo_decode(int fmt, const uint8_t *buf, uint64_t nbytes, int64_t first, int64_t count, double *iq)
{
    if (fmt < 0 || fmt >= F_COUNT) return -3;
    view_t v = { buf, nbytes, fmt };
    for (int64_t i = 0; i < count; i++) {
        iq[2 * i] = sample_iq(&v, first + i, 0);
        iq[2 * i + 1] = sample_iq(&v, first + i, 1);
    }
    return 0;
}

/* ------------------------------------------------------------------ windows: lib/windows.js:14-88 */

enum { W_RECT, W_BARTLETT, W_HAMMING, W_HANN, W_BLACKMAN, W_BLACKMAN_HARRIS, W_COUNT };

/* fills w[n]; returns the weight (sum accumulated in index order) */
double spo_window(int kind, int n, double *w)
{
    double weight = 0.0;
    for (int i = 0; i < n; ++i) {
        double x;
        switch (kind) {
        default:
        case W_RECT: x = 1.0; break;                                                        /* :15-22 */
        case W_BARTLETT: x = 1.0 - fabs((i - 0.5 * (n - 1)) / (0.5 * (n - 1))); break;        /* :25-32 */
        case W_HAMMING: x = 0.54 - 0.46 * cos(2.0 * M_PI * i / (n - 1)); break;               /* :35-45 */
        case W_HANN: x = 0.5 * (1.0 - cos(2.0 * M_PI * i / (n - 1))); break;                  /* :48-55 */
        case W_BLACKMAN:                                                                     /* :58-71 */
            x = 0.42 - (0.5 * cos((2.0 * M_PI * i) / (n - 1))) + (0.08 * cos((4.0 * M_PI * i) / (n - 1)));
            break;
        case W_BLACKMAN_HARRIS:                                                              /* :74-88 */
            x = 0.35875 - (0.48829 * cos((2.0 * M_PI * i) / (n - 1)))
                + (0.14128 * cos((4.0 * M_PI * i) / (n - 1)))
                - (0.01168 * cos((6.0 * M_PI * i) / (n - 1)));
            break;
        }
        w[i] = x;
        weight += x;
    }
    return weight;
}

/* ------------------------------------------------------------------ FFT: lib/fft_nayuki.js:29-119 */

typedef struct { int n, levels; double *ct, *st; uint32_t *rev; } fft_t;

static int fft_init(fft_t *f, int n)
{
    memset(f, 0, sizeof *f);
    int levels = -1;
    for (int i = 0; i < 31; i++) if ((1 << i) == n) levels = i;
    if (levels < 0) return -2;                       /* 'Length is not a power of 2' :38-39 */
    f->n = n; f->levels = levels;
    int h = n / 2 > 0 ? n / 2 : 1;
    f->ct = malloc(sizeof(double) * h);
    f->st = malloc(sizeof(double) * h);
    f->rev = malloc(sizeof(uint32_t) * n);
    for (int i = 0; i < n / 2; i++) {                /* :42-47 */
        f->ct[i] = cos(2 * M_PI * i / n);
        f->st[i] = sin(2 * M_PI * i / n);
    }
    for (int i = 0; i < n; i++) {                    /* reverseBits, :88-95 (tabulated here) */
        uint32_t y = 0, x = (uint32_t)i;
        for (int b = 0; b < levels; b++) { y = (y << 1) | (x & 1); x >>= 1; }
        f->rev[i] = y;
    }
    return 0;
}
static void fft_free(fft_t *f) { free(f->ct); free(f->st); free(f->rev); }

/* forward, unscaled, in place: bit-reversal then radix-2 decimation in time (:54-86) */
static void fft_transform(const fft_t *f, double *re, double *im)
{
    const int n = f->n;
    for (int i = 0; i < n; i++) {
        int j = (int)f->rev[i];
        if (j > i) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int size = 2; size <= n; size *= 2) {
        int half = size / 2, step = n / size;
        for (int base = 0; base < n; base += size) {
            for (int j = base, k = 0; j < base + half; j++, k += step) {
                int l = j + half;
                double c = f->ct[k], s = f->st[k];
                double tr = re[l] * c + im[l] * s;       /* :77 */
                double ti = -re[l] * s + im[l] * c;      /* :78 */
                re[l] = re[j] - tr;
                im[l] = im[j] - ti;
                re[j] += tr;
                im[j] += ti;
            }
        }
    }
}

/* two real channels out of one complex DFT (:103-119).  Note real[n/2] receives the
 * already zeroed imag[0] (:106-107). */
static void fft_splitreal(const fft_t *f, double *re, double *im)
{
    const int n = f->n;
    im[0] = 0;
    re[n / 2] = im[0];
    im[n / 2] = 0;
    for (int i = 1; i < n / 2; i++) {
        double lr = 0.5 * (re[i] + re[n - i]);
        double li = 0.5 * (im[i] - im[n - i]);
        double rr = 0.5 * (im[i] + im[n - i]);
        double ri = 0.5 * (-re[i] + re[n - i]);
        re[i] = lr; im[i] = li; re[n - i] = rr; im[n - i] = ri;
    }
}

/* test taps */
int spo_fft(int n, double *re, double *im)
{
    fft_t f; int rc = fft_init(&f, n); if (rc) return rc;
    fft_transform(&f, re, im); fft_free(&f); return 0;
}
int spo_splitreal(int n, double *re, double *im)
{
    fft_t f; int rc = fft_init(&f, n); if (rc) return rc;
    fft_splitreal(&f, re, im); fft_free(&f); return 0;
}

/* ------------------------------------------------------------------ renderFft: lib/worker.js:23-156 */

static int render_impl(const spo_request *rq, spo_reply *rp)
{
    if (!rq || !rp || !rq->buffer || !rq->windowc || !rq->cmap_rgb) return -1;
    if (rq->format < 0 || rq->format >= F_COUNT) return -3;
    if (rq->cmap_len < 1) return -7;
    if (rq->byte_length % (uint64_t)k_elem[rq->format]) return -6;   /* new TypedArray(buffer) throws */

    view_t view = { rq->buffer, rq->byte_length, rq->format };
    const double sampleCount = (double)rq->byte_length / k_width[rq->format];   /* samples.js:167 */

    const double block_norm_db = 10 * log10(rq->block_norm);                     /* :32 */
    const double gain = rq->gain, dB_range = rq->range;
    double dBfs_min = 0.0, dBfs_max = -200.0;                                    /* :35-36 */
    const int color_max = rq->cmap_len - 1;                                      /* :38 */
    const double color_norm = rq->cmap_len / -dB_range;                          /* :39 */

    const int n = rq->n;
    const int64_t width = rq->width;
    fft_t fft; int rc = fft_init(&fft, n); if (rc) return rc;
    const double stride = (sampleCount - n) / (double)(width - 1);               /* :50 */

    if (rp->cB_hist) memset(rp->cB_hist, 0, sizeof(uint64_t) * SPO_CB_HIST);
    if (rp->c_hist) memset(rp->c_hist, 0, sizeof(uint64_t) * rq->cmap_len);
    uint64_t cb_local[SPO_CB_HIST]; memset(cb_local, 0, sizeof cb_local);
    uint64_t *c_local = calloc((size_t)rq->cmap_len, sizeof(uint64_t));

    double *re = malloc(sizeof(double) * n), *im = malloc(sizeof(double) * n);

    for (int64_t x = 0; x < width; x++) {                                        /* :68 */
        const int64_t p0 = js_toint32(0.5 + stride * (double)x);                 /* :72 */
        for (int k = 0; k < n; k++) {
            int64_t pos = p0 + k;
            re[k] = rq->windowc[k] * sample_iq(&view, pos, 0);                   /* :73 */
            im[k] = rq->windowc[k] * sample_iq(&view, pos, 1);                   /* :74 */
        }
        fft_transform(&fft, re, im);                                             /* :77 */
        if (rq->channel_mode) fft_splitreal(&fft, re, im);                       /* :78-80 */

        double min_i = 0.0, max_i = -200.0;                                      /* :82-83 */
        for (int i = 0; i < n; i++) {                                            /* :85 */
            const int y = (i <= n / 2) ? n / 2 - i : n / 2 + n - i;              /* :90 */
            const double abs2 = re[i] * re[i] + im[i] * im[i];                   /* :92 */
            const double dBfs = 5 * log10(abs2) + block_norm_db + gain;          /* :93 */
            const double d0 = dBfs - gain;
            if (d0 < min_i) min_i = d0;                                          /* :102 */
            if (d0 > max_i) max_i = d0;                                          /* :103 */

            const int32_t cBabs = js_toint32(0.5 + d0 * -10);                    /* :105 */
            const int32_t kbin = cBabs >= SPO_CB_HIST ? SPO_CB_HIST - 1 : cBabs; /* :106 */
            if (kbin >= 0) cb_local[kbin] += 1;    /* negative index: a non-index property, dropped */

            const double grayU = color_max - dBfs * color_norm;                  /* :111 */
            const int32_t gray = js_toint32(0.5 + (grayU < 0 ? 0 : grayU > color_max ? color_max : grayU)); /* :112 */
            c_local[gray] += 1;                                                  /* :113 */
            if (rp->image) {
                const uint8_t *color = rq->cmap_rgb + 3 * gray;                  /* :114 */
                const int64_t j = rq->waterfall
                    ? (int64_t)n * (width - 1 - x) * 4 + (int64_t)(n - 1 - y) * 4  /* :116 */
                    : x * 4 + width * (int64_t)y * 4;                            /* :117 */
                rp->image[j + 0] = color[0];
                rp->image[j + 1] = color[1];
                rp->image[j + 2] = color[2];
                rp->image[j + 3] = 255;
            }
            if (rp->db) rp->db[x * n + i] = d0;
            if (rp->gray) rp->gray[x * n + i] = (uint16_t)gray;
            if (rp->cbk) rp->cbk[x * n + i] = kbin < 0 ? -1 : kbin;
        }
        if (min_i < dBfs_min) dBfs_min = min_i;                                  /* :124 */
        if (max_i > dBfs_max) dBfs_max = max_i;                                  /* :125 */

        if (rp->gauge_mins) rp->gauge_mins[x] = js_u8clamped(0.5 + (dB_range + min_i) * 256 / dB_range); /* :128 */
        if (rp->gauge_maxs) rp->gauge_maxs[x] = js_u8clamped(0.5 + (dB_range + max_i) * 256 / dB_range); /* :129 */

        const int64_t mid = p0 + n / 2;                                          /* :131 */
        const double mr = sample_iq(&view, mid, 0), mi = sample_iq(&view, mid, 1);
        const double a2 = mr * mr + mi * mi;
        const double dBfs_amp = 5 * log10(a2) + gain;                            /* :135 */
        if (rp->gauge_amps) rp->gauge_amps[x] = js_u8clamped(0.5 + (dB_range + dBfs_amp) * 256 / dB_range); /* :136 */
    }
    if (rp->cB_hist) memcpy(rp->cB_hist, cb_local, sizeof cb_local);
    if (rp->c_hist) memcpy(rp->c_hist, c_local, sizeof(uint64_t) * rq->cmap_len);
    rp->dBfs_min = dBfs_min;
    rp->dBfs_max = dBfs_max;
    free(re); free(im); free(c_local); fft_free(&fft);
    return 0;
}

int spo_render(const spo_request *rq, spo_reply *rp)
{
    if (!rq) return -1;
    /* width == 1: stride = (sampleCount - n) / 0 is +-Infinity or NaN, stride * 0 is NaN and ~~NaN == 0
     * (lib/worker.js:50,72): one frame at sample 0, which render_impl's own arithmetic reproduces */
    if (rq->width < 1) return -5;
    return render_impl(rq, rp);
}

/* ------------------------------------------------------------------ fan-out: lib/spectroplot.js:1206-1238 */
/*
 * The caller side of the protocol: slice the capture into `workers` disjoint chunks
 * (SampleView.slice, lib/samples.js:253-258), render each as its own message of
 * width ~~(width/workers) on its own thread, merge histograms / min / max and blit
 * tiles at `offset`.  This is the reference's CPU parallelism and therefore the CPU
 * baseline of the bench.  Trailing columns / samples stay unrendered like upstream.
 */
typedef struct { spo_request rq; spo_reply rp; int rc; } job_t;
static void *job_main(void *p) { job_t *j = p; j->rc = spo_render(&j->rq, &j->rp); return NULL; }

int spo_render_fanout(const spo_request *rq, spo_reply *rp, int workers)
{
    if (!rq || !rp || workers < 1) return -1;
    const int sw = k_width[rq->format];
    const int64_t endSample = (int64_t)(rq->byte_length / (uint64_t)sw);         /* :1207 */
    const int64_t sliceWidth = rq->width / workers;                              /* :1208 */
    const int64_t sliceLength = (int64_t)sw * (endSample / workers);             /* samples.js:256 */
    const int n = rq->n;
    job_t *jobs = calloc((size_t)workers, sizeof(job_t));
    pthread_t *th = calloc((size_t)workers, sizeof(pthread_t));
    for (int i = 0; i < workers; i++) {
        job_t *j = &jobs[i];
        j->rq = *rq;
        j->rq.buffer = rq->buffer + sliceLength * i;
        j->rq.byte_length = (uint64_t)sliceLength;
        j->rq.width = sliceWidth;
        j->rp.image = rp->image ? malloc((size_t)4 * sliceWidth * n) : NULL;
        j->rp.gauge_mins = rp->gauge_mins ? rp->gauge_mins + i * sliceWidth : NULL;
        j->rp.gauge_maxs = rp->gauge_maxs ? rp->gauge_maxs + i * sliceWidth : NULL;
        j->rp.gauge_amps = rp->gauge_amps ? rp->gauge_amps + i * sliceWidth : NULL;
        j->rp.cB_hist = calloc(SPO_CB_HIST, sizeof(uint64_t));
        j->rp.c_hist = calloc((size_t)rq->cmap_len, sizeof(uint64_t));
        pthread_create(&th[i], NULL, job_main, j);
    }
    int rc = 0;
    double mn = 0.0, mx = -200.0;                                                /* :1122-1123 */
    if (rp->cB_hist) memset(rp->cB_hist, 0, sizeof(uint64_t) * SPO_CB_HIST);
    if (rp->c_hist) memset(rp->c_hist, 0, sizeof(uint64_t) * rq->cmap_len);
    for (int i = 0; i < workers; i++) {
        pthread_join(th[i], NULL);
        job_t *j = &jobs[i];
        if (j->rc && !rc) rc = j->rc;
        if (!j->rc) {
            if (j->rp.dBfs_min < mn) mn = j->rp.dBfs_min;                        /* :1230 */
            if (j->rp.dBfs_max > mx) mx = j->rp.dBfs_max;                        /* :1231 */
            if (rp->cB_hist) for (int b = 0; b < SPO_CB_HIST; b++) rp->cB_hist[b] += j->rp.cB_hist[b];
            if (rp->c_hist) for (int b = 0; b < rq->cmap_len; b++) rp->c_hist[b] += j->rp.c_hist[b];
            if (rp->image) {
                /* putImageData(tile, offset, 0) / (0, width - sliceWidth - offset): :1244 */
                const int64_t off = (int64_t)i * sliceWidth;
                if (!rq->waterfall) {
                    for (int y = 0; y < n; y++)
                        memcpy(rp->image + 4 * (off + rq->width * (int64_t)y),
                               j->rp.image + 4 * sliceWidth * (int64_t)y, (size_t)4 * sliceWidth);
                } else {
                    const int64_t row0 = rq->width - sliceWidth - off;
                    memcpy(rp->image + 4 * (int64_t)n * row0, j->rp.image, (size_t)4 * sliceWidth * n);
                }
            }
        }
        free(j->rp.image); free(j->rp.cB_hist); free(j->rp.c_hist);
    }
    rp->dBfs_min = mn; rp->dBfs_max = mx;
    free(jobs); free(th);
    return rc;
}

/* ------------------------------------------------------------------ computed colormaps */

/* lib/soxcmap.js:12-46 */
void spo_cmap_sox(int stops, uint8_t *rgb)
{
    for (int i = 0; i < stops; ++i) {
        double x = i / (stops - 1.0), c0, c1, c2;
        if (x < .13) c0 = 0; else if (x < .73) c0 = 1 * sin((x - .13) / .60 * M_PI / 2); else c0 = 1;
        if (x < .60) c1 = 0; else if (x < .91) c1 = 1 * sin((x - .60) / .31 * M_PI / 2); else c1 = 1;
        if (x < .60) c2 = .5 * sin((x - .00) / .60 * M_PI); else if (x < .78) c2 = 0; else c2 = (x - .78) / .22;
        /* Math.round: floor(v + 0.5) for these non-negative values */
        rgb[3 * i + 0] = (uint8_t)floor(255 * c0 + 0.5);
        rgb[3 * i + 1] = (uint8_t)floor(255 * c1 + 0.5);
        rgb[3 * i + 2] = (uint8_t)floor(255 * c2 + 0.5);
    }
}

/* lib/naivecmap.js:13-79; kind 0 naive, 1 grayscale, 2 roentgen, 3 phosphor */
void spo_cmap_naive(int kind, int stops, uint8_t *rgb)
{
    for (int i = 0; i < stops; ++i) {
        double r = 0, g = 0, b = 0;
        if (kind == 0) {
            if (i < stops / 4.0) { b = i * 128 / (stops / 4.0); g = 0; r = 0; }
            else if (i < stops / 2.0) { b = 256 - i / 2.0; g = 0; r = i - stops / 4.0; }
            else if (i < stops * 3 / 4.0) { b = 0; g = i - stops / 2.0; r = 255; }
            else { b = i - stops * 3 / 4.0; g = 255; r = 255; }
        } else if (kind == 1) {
            r = g = b = i * 255 / (double)stops;
        } else if (kind == 2) {
            r = g = b = 255 - (i * 255 / (double)stops);
        } else {
            if (i < stops / 2.0) { r = 0; g = i * 191 / (stops / 2.0); b = 0; }
            else {
                r = (i - stops / 2.0) * 255 / (stops / 2.0);
                g = 191 + (i - stops / 2.0) * 64 / (stops / 2.0);
                b = (i - stops / 2.0) * 255 / (stops / 2.0);
            }
        }
        /* the table holds ~~v (plain numbers, e.g. 256 for naive blue at i = stops/4 ... );
         * the Uint8ClampedArray store in the worker clamps to 255 */
        int32_t ri = js_toint32(r), gi = js_toint32(g), bi = js_toint32(b);
        rgb[3 * i + 0] = (uint8_t)(ri < 0 ? 0 : ri > 255 ? 255 : ri);
        rgb[3 * i + 1] = (uint8_t)(gi < 0 ? 0 : gi > 255 ? 255 : gi);
        rgb[3 * i + 2] = (uint8_t)(bi < 0 ? 0 : bi > 255 ? 255 : bi);
    }
}

/* ------------------------------------------------------------------ synthetic capture generator */
/*
 * Integer-only and counter-based: a pure function of
 * (seed, sample index i, total capture length S).  The CUDA generator in
 * spectroplot-js_b200/csrc implements the same arithmetic; tests compare them bit for bit.
 *   tone A : +fs/8, -6 dBFS           (phase = 512*i mod 4096 LUT steps)
 *   tone B : linear chirp -fs/4 -> +fs/4 over the capture, -20 dBFS (64-bit wrapping phase)
 *   noise  : sum of four hashed 16-bit uniforms per component, about -50 dBFS rms
 * The result is a signed 16-bit I/Q pair which is then re-quantised to the target format.
 */
void spo_synth_lut(int16_t *lut)
{
    for (int j = 0; j < 4096; j++) lut[j] = (int16_t)lround(32767.0 * sin(2.0 * M_PI * j / 4096.0));
}

static inline uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline int32_t sum4x16(uint64_t r)
{
    return (int32_t)((r & 0xFFFF) + ((r >> 16) & 0xFFFF) + ((r >> 32) & 0xFFFF) + (r >> 48)) - 131070;
}
static inline int32_t clamp16(int32_t v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }

static void synth_sample16(const int16_t *lut, uint64_t seed, uint64_t i, uint64_t S, int32_t *I, int32_t *Q)
{
    uint32_t ph1 = (uint32_t)((i * 512u) & 4095u);
    int32_t c1 = lut[(ph1 + 1024u) & 4095u] >> 1, s1 = lut[ph1] >> 1;
    uint64_t f0 = (uint64_t)0 - ((uint64_t)1 << 62);
    uint64_t delta = S ? (((uint64_t)1 << 63) / S) : 0;
    uint64_t a = i, b = i - 1;                /* tri = i(i-1)/2 mod 2^64 */
    if (a & 1) b >>= 1; else a >>= 1;
    uint64_t tri = a * b;
    uint64_t ph2 = f0 * i + delta * tri;
    uint32_t idx2 = (uint32_t)(ph2 >> 52);
    int32_t c2 = (lut[(idx2 + 1024u) & 4095u] * 3277) >> 15, s2 = (lut[idx2] * 3277) >> 15;
    int32_t nI = (sum4x16(splitmix64(seed ^ (2 * i))) * 11) >> 12;
    int32_t nQ = (sum4x16(splitmix64(seed ^ (2 * i + 1))) * 11) >> 12;
    *I = clamp16(c1 + c2 + nI);
    *Q = clamp16(s1 + s2 + nQ);
}

static void synth_pack(int fmt, uint8_t *dst, uint64_t k, int32_t I, int32_t Q)
{
    switch (fmt) {
    case F_CU4: dst[k] = (uint8_t)((((I >> 12) + 8) << 4) | ((Q >> 12) + 8)); break;
    case F_CS4: dst[k] = (uint8_t)((((I >> 12) & 15) << 4) | ((Q >> 12) & 15)); break;
    case F_CU8: dst[2 * k] = (uint8_t)((I >> 8) + 128); dst[2 * k + 1] = (uint8_t)((Q >> 8) + 128); break;
    case F_CS8: dst[2 * k] = (uint8_t)(I >> 8); dst[2 * k + 1] = (uint8_t)(Q >> 8); break;
    case F_CU12: case F_CS12: {
        uint32_t a = (uint32_t)((I >> 4) + (fmt == F_CU12 ? 2048 : 0)) & 0xFFF;
        uint32_t b = (uint32_t)((Q >> 4) + (fmt == F_CU12 ? 2048 : 0)) & 0xFFF;
        dst[3 * k] = (uint8_t)(a & 0xFF);
        dst[3 * k + 1] = (uint8_t)((a >> 8) | ((b & 0xF) << 4));
        dst[3 * k + 2] = (uint8_t)(b >> 4);
        break;
    }
    case F_CU16: { uint16_t v[2] = { (uint16_t)(I + 32768), (uint16_t)(Q + 32768) }; memcpy(dst + 4 * k, v, 4); break; }
    case F_CS16: { int16_t v[2] = { (int16_t)I, (int16_t)Q }; memcpy(dst + 4 * k, v, 4); break; }
    case F_CU32: { uint32_t v[2] = { ((uint32_t)I << 16) + 0x80000000u, ((uint32_t)Q << 16) + 0x80000000u }; memcpy(dst + 8 * k, v, 8); break; }
    case F_CS32: { uint32_t v[2] = { (uint32_t)I << 16, (uint32_t)Q << 16 }; memcpy(dst + 8 * k, v, 8); break; }
    case F_CU64: case F_CS64: {
        uint64_t off = fmt == F_CU64 ? 0x8000000000000000ull : 0;
        uint64_t v[2] = { ((uint64_t)(int64_t)I << 48) + off, ((uint64_t)(int64_t)Q << 48) + off };
        memcpy(dst + 16 * k, v, 16); break;
    }
    case F_CF32: { float v[2] = { (float)I / 32768.0f, (float)Q / 32768.0f }; memcpy(dst + 8 * k, v, 8); break; }
    case F_CF64: { double v[2] = { (double)I / 32768.0, (double)Q / 32768.0 }; memcpy(dst + 16 * k, v, 16); break; }
    }
}

int spo_synth_fill(uint8_t *dst, int fmt, uint64_t first, uint64_t count, uint64_t total, uint64_t seed)
{
    if (fmt < 0 || fmt >= F_COUNT) return -3;
    int16_t lut[4096]; spo_synth_lut(lut);
    for (uint64_t k = 0; k < count; k++) {
        int32_t I, Q; synth_sample16(lut, seed, first + k, total, &I, &Q);
        synth_pack(fmt, dst, k, I, Q);
    }
    return 0;
}
