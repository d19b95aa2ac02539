/*
 * spectro_b200.h — C ABI of the B200-native spectrogram render engine.
 *
 * This is the drop-in boundary for ONE path of triq-org/spectroplot-js: the
 * per-message render done by its Web Worker (reference lib/worker.js:23-156,
 * `renderFft(ctx)`), i.e.  raw I/Q bytes -> decode -> window -> FFT -> dB ->
 * colormap -> RGBA image + dB/colour histograms + per-frame gauges.
 *
 * The reference has no native interface; a maintainer binds these entry
 * points from a Node.js N-API addon underneath a
 * worker-protocol shim, so `new Spectroplot(options)`, `setOption`,
 * `setOptions`, `setData` and the lib/worker.js message protocol stay as
 * they are.
 *
 * Plain C: pointers and sizes only.  No torch / STL types cross this line.
 * All entry points return 0 on success or a negative SP_E_* code; the text
 * of the last failure is available from sp_last_error().  Nothing aborts.
 *
 * Threading: one in-flight call per engine (the reference relies on
 * "sequential worker execution", lib/spectroplot.js:89); engines are
 * independent of each other.
 *
 * Ownership: the caller owns every buffer named in a request / reply.  The
 * engine never frees them and never retains them past the return of the call.
 */
#ifndef SPECTRO_B200_H
#define SPECTRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_ABI_VERSION 1

/* ---- sample formats: reference lib/samples.js:30-155 (format table) ---- */
enum sp_format {
    SP_CU4 = 0,  /* lib/samples.js:30   1 byte / sample, I = high nibble           */
    SP_CS4 = 1,  /* lib/samples.js:39   1 byte / sample, two's complement nibbles  */
    SP_CU8 = 2,  /* lib/samples.js:48   aliases DATA, COMPLEX16U, and any unknown  */
    SP_CS8 = 3,  /* lib/samples.js:55   alias COMPLEX16S                           */
    SP_CU12 = 4, /* lib/samples.js:76   3 bytes / sample, packed (iiqIQQ)          */
    SP_CS12 = 5, /* lib/samples.js:85                                              */
    SP_CU16 = 6, /* lib/samples.js:62                                              */
    SP_CS16 = 7, /* lib/samples.js:69                                              */
    SP_CU32 = 8, /* lib/samples.js:94                                              */
    SP_CS32 = 9, /* lib/samples.js:101                                             */
    SP_CU64 = 10,/* lib/samples.js:108  read as two uint32 words                   */
    SP_CS64 = 11,/* lib/samples.js:117                                             */
    SP_CF32 = 12,/* lib/samples.js:126  aliases CFILE, COMPLEX                     */
    SP_CF64 = 13,/* lib/samples.js:133                                             */
    SP_FORMAT_COUNT = 14
};

/* ---- error codes ---- */
enum sp_error {
    SP_OK = 0,
    SP_E_INVAL = -1,        /* null pointer / malformed request                        */
    SP_E_BAD_N = -2,        /* n is not a power of two (reference throws a string,
                               lib/fft_nayuki.js:38-39) or outside [SP_MIN_N, SP_MAX_N] */
    SP_E_BAD_FORMAT = -3,   /* format enum out of range                                */
    SP_E_TOO_SHORT = -4,    /* a SHARD of a message with sampleCount < n (a whole such message is
                               rendered like the reference renders it: undefined -> NaN frames) */
    SP_E_BAD_WIDTH = -5,    /* width < 1 (width == 1 renders one frame at sample 0 like
                               lib/worker.js:50,72: stride = x/0, ~~(0.5 + NaN) == 0)   */
    SP_E_RAGGED = -6,       /* byte_length not a multiple of the typed-array element
                               size (the reference's `new Int16Array(buffer)` throws)  */
    SP_E_BAD_CMAP = -7,     /* cmap_len < 2 or > SP_MAX_CMAP                           */
    SP_E_CUDA = -8,         /* CUDA runtime failure; see sp_last_error                 */
    SP_E_NO_DEVICE = -9,    /* no usable sm_100 device: there is NO CPU fallback       */
    SP_E_RANGE = -10,       /* shard fields inconsistent with the global message       */
    SP_E_ALIGN = -11,       /* device-resident buffer not 16-byte aligned              */
    SP_E_NCCL = -12         /* sp_render_shards: NCCL could not be loaded / initialised
                               (libnccl.so.2 is opened with dlopen when a multi-device engine
                               is created), or the merge collective failed               */
};

#define SP_MIN_N 2          /* reference accepts any power of two (lib/fft_nayuki.js:38-39); kernels cover
                               2..262144 (n = 1 draws nothing upstream: its row index n/2 - i is fractional) */
#define SP_MAX_N 262144
#define SP_CB_HIST_SIZE 1000 /* lib/worker.js:41: centi-Bel bins, 0.0 .. -100.0 dB        */
#define SP_MAX_CMAP 4096     /* custom RGB[] maps of any length (lib/spectroplot.js:245)  */

/* request flags */
#define SP_F_BUFFER_ON_DEVICE 1u /* request.buffer is a device pointer (16-byte aligned, readable
                                    up to the next multiple of 16 bytes)                      */
#define SP_F_REPLY_ON_DEVICE  2u /* every non-null pointer in the reply is a device pointer   */
#define SP_F_NO_IMAGE         4u /* skip the RGBA image (histogram / autorange pre-pass only) */

/*
 * One render request == one worker message
 * (reference lib/spectroplot.js:1213-1226 builds it, lib/worker.js:23-62 reads it).
 *
 * The shard_* fields extend the message so that a long capture can be split by
 * contiguous FRAME range across GPUs and still reproduce the unsharded message
 * bit for bit (frame positions always come from the GLOBAL stride).  Leave them
 * zero for an ordinary, whole message.
 */
typedef struct sp_request {
    const void *buffer;       /* raw interleaved I/Q bytes (ctx.buffer)                         */
    uint64_t byte_length;     /* bytes available at `buffer`                                    */
    int32_t format;           /* enum sp_format (ctx.format after SampleView's alias table)     */
    int32_t n;                /* FFT size (ctx.n), power of two                                 */
    int64_t width;            /* number of frames == pixel columns (ctx.width)                  */
    double block_norm;        /* 1 / window weight (ctx.block_norm, lib/spectroplot.js:1116)    */
    double gain;              /* dB (ctx.gain)                                                  */
    double range;             /* dB (ctx.range)                                                 */
    const double *windowc;    /* n window coefficients (ctx.windowc)                            */
    const uint8_t *cmap_rgb;  /* cmap_len x 3 bytes, endpoint overwrite already applied by the
                                 caller as in lib/spectroplot.js:1129-1130 (ctx.cmap)           */
    int32_t cmap_len;
    int32_t channel_mode;     /* ctx.channelMode: run the split-real post-process               */
    int32_t waterfall;        /* ctx.waterfall: frame-contiguous layout (lib/worker.js:116)     */
    uint32_t flags;           /* SP_F_*                                                         */

    /* ---- frame-range shard of a larger message (all zero => not a shard) ---- */
    uint64_t total_byte_length; /* byte length of the WHOLE capture the global stride refers to */
    int64_t total_width;        /* width of the WHOLE message                                    */
    int64_t frame_first;        /* first global frame index rendered by this call               */
    uint64_t buffer_first_sample; /* global sample index of buffer[0]                           */
} sp_request;

/*
 * Reply == the worker's postMessage payload (reference lib/worker.js:140-155).
 * Caller allocates; null pointers mean "not wanted".
 */
typedef struct sp_reply {
    uint8_t *image;       /* 4*width*n bytes RGBA.  spectrogram: row-major [n][width];
                             waterfall: [width][n] (lib/worker.js:115-121)                   */
    uint8_t *gauge_mins;  /* width bytes (lib/worker.js:128)                                 */
    uint8_t *gauge_maxs;  /* width bytes (lib/worker.js:129)                                 */
    uint8_t *gauge_amps;  /* width bytes (lib/worker.js:131-136)                             */
    uint64_t *cB_hist;    /* SP_CB_HIST_SIZE counters (lib/worker.js:42,106)                 */
    uint64_t *c_hist;     /* cmap_len counters (lib/worker.js:43,113)                        */
    double dBfs_min;      /* min of (dBfs - gain), initial 0.0   (lib/worker.js:35,124)      */
    double dBfs_max;      /* max of (dBfs - gain), initial -200  (lib/worker.js:36,125)      */
    float device_ms;      /* out: device time of the kernels of this call (CUDA events)      */
    int32_t kernel_launches; /* out: number of engine kernels launched by this call          */
    double *minmax_dev;   /* optional, SP_F_REPLY_ON_DEVICE only: device double[2] that receives
                             {dBfs_min, dBfs_max} without a host round trip (multi-GPU merge) */
} sp_reply;

typedef struct sp_engine sp_engine;

/* Library / ABI identification.  sp_build_id(): hash of the kernel sources the library was built from
 * (measurement records name the build they belong to). */
int sp_abi_version(void);
const char *sp_build_id(void);

/* Format helpers: lib/samples.js:22,30-155.  Name matching is case-insensitive and
 * follows the reference's alias table; an unknown name maps to SP_CU8 like the
 * reference's final `else` (lib/samples.js:149-155). */
int sp_format_from_name(const char *name);
const char *sp_format_name(int format);
int sp_sample_width(int format);   /* bytes per complex sample, <0 on bad format */
int sp_element_size(int format);   /* typed-array element size in bytes           */

/* Create an engine on the listed CUDA devices (ndev >= 1; device_ids may be NULL
 * to mean devices 0 .. ndev-1).  With ndev > 1, sp_render() splits a whole
 * host-buffer message by contiguous frame range (cuts on multiples of 8 frames,
 * global frame positions, an n-sample halo per range) across the devices, one
 * host thread per device, each device writing its column band / row block
 * straight into the caller's image; the histograms and min / max are merged on
 * the host like lib/spectroplot.js:1229-1238 (ndev x ~10 KB).  The result is
 * the single-device result (bit-identical when the width is a multiple of 8).
 * Device-resident shards are rendered and merged over NVLink by
 * sp_render_shards() (one NCCL communicator per device, ncclCommInitAll at
 * creation; libnccl is opened with dlopen, so a single-GPU host needs none).
 * The taps and memory helpers address the device chosen with sp_select_device().
 * Fails with SP_E_NO_DEVICE when no sm_100 GPU is present. */
int sp_create(sp_engine **out, const int *device_ids, int ndev);
void sp_destroy(sp_engine *e);
const char *sp_last_error(sp_engine *e); /* e may be NULL: last error of sp_create */

/* Use an existing CUDA stream (a cudaStream_t passed as void*) for device 0 of the
 * engine instead of the engine's own stream; NULL restores the engine's stream. */
int sp_set_stream(sp_engine *e, void *cuda_stream);

/* The path itself: replaces renderFft(ctx), reference lib/worker.js:23-156. */
int sp_render(sp_engine *e, const sp_request *rq, sp_reply *rp);

/* Same, but returns after enqueueing when every buffer is device resident
 * (SP_F_BUFFER_ON_DEVICE | SP_F_REPLY_ON_DEVICE); dBfs_min/max are then read
 * with sp_render_finish().  Used to time the kernels without a host sync. */
int sp_render_enqueue(sp_engine *e, const sp_request *rq, sp_reply *rp);
int sp_render_finish(sp_engine *e, sp_reply *rp);

/* Asynchronous form of sp_render() for host buffers on a single-device engine: the message is enqueued (chunked copies in,
 * kernels, copies out on the engine's streams) and the call returns a ticket; sp_render_wait() blocks until that message's
 * reply is complete and fills dBfs_min / dBfs_max.  Up to TWO messages may be in flight: the copy-out tail of one overlaps
 * the copy-in head of the next (a third call first collects the oldest).  Replaces the reference's posting of messages to
 * several workers and collecting the replies as they arrive (lib/spectroplot.js:1206-1238, worker.onmessage).  `rp` and
 * every buffer `rq` / `rp` point to must stay alive and untouched until the ticket has been waited for; page-locked host
 * memory (sp_host_alloc_pinned) is needed for the copies to overlap.  Short messages and device-resident ones are simply
 * rendered before the call returns.  Results are bit-identical to sp_render(). */
int sp_render_async(sp_engine *e, const sp_request *rq, sp_reply *rp, int *ticket);
int sp_render_wait(sp_engine *e, int ticket);

/* Device-resident shards of ONE message on a multi-device engine: rq[g] / rp[g] (g = 0 .. ndev-1) describe the
 * frame-range shard that lives on device g - bytes, image band, gauges, both histograms and minmax_dev are device
 * pointers on THAT device (SP_F_BUFFER_ON_DEVICE | SP_F_REPLY_ON_DEVICE, shard fields set; allocate with
 * sp_select_device + sp_device_alloc).  Every device renders its shard on its own stream, then one grouped NCCL
 * all-reduce over NVLink merges cB_hist / c_hist (sum, u64) and {dBfs_min, dBfs_max} (min / max, f64) in place:
 * the caller-side merge of lib/spectroplot.js:1229-1238 without a host copy.  On return every reply carries the
 * statistics of the WHOLE message; the image bands and gauges stay per device.  SP_E_NCCL when NCCL is missing. */
int sp_render_shards(sp_engine *e, const sp_request *rq, sp_reply *rp);

/* Which device of a multi-device engine the taps, the memory helpers and sp_synth_fill address (default 0). */
int sp_select_device(sp_engine *e, int index);

/* Several zoom levels of ONE capture in one pass over the capture bytes: the buffer crosses PCIe
 * once (or is bound, if device-resident) and level i is rendered as the message `rq` with
 * width = widths[i], i.e. with its OWN stride (sampleCount - n)/(widths[i] - 1) exactly as the
 * reference does when the user zooms (one processData() per zoom step: lib/spectroplot.js:513-527,
 * 1103-1104; SURVEY.md A.6 — the frames of zoom 1 are not a subset of zoom z's frames).
 * replies[i] is the reply of level i; rq->width is ignored.  Stops at the first failing level. */
int sp_render_zooms(sp_engine *e, const sp_request *rq, int nlevels, const int64_t *widths, sp_reply *replies);

/* Test tap: decode `count` samples starting at sample `first` to interleaved
 * fp32 I/Q (iq[2*count], host memory) with the SAME device function the fused
 * kernel uses.  Replaces SampleView.sampleI/Q, lib/samples.js:313-400.
 * Out-of-range samples decode to NaN like the reference's `undefined`. */
int sp_decode(sp_engine *e, int format, const void *bytes, uint64_t nbytes,
              uint64_t first, uint64_t count, float *iq);

/* Test tap: per-bin dB values (dBfs - gain, fp32) of a whole message, row-major
 * [width][n] in FFT bin order, host memory.  Same kernels as sp_render. */
int sp_render_db(sp_engine *e, const sp_request *rq, float *db);

/* Device memory helpers for callers that keep captures resident in HBM
 * (the bench, the multi-GPU host layer).  Device 0 of the engine unless noted. */
int sp_device_alloc(sp_engine *e, uint64_t nbytes, void **dptr);
int sp_device_free(sp_engine *e, void *dptr);
int sp_memcpy_h2d(sp_engine *e, void *dst_dev, const void *src_host, uint64_t nbytes);
int sp_memcpy_d2h(sp_engine *e, void *dst_host, const void *src_dev, uint64_t nbytes);
int sp_host_alloc_pinned(uint64_t nbytes, void **hptr);
int sp_host_free_pinned(void *hptr);
int sp_device_sync(sp_engine *e);

/* Deterministic integer-only synthetic capture (two tones + hashed noise), a pure
 * function of (seed, sample index, total_samples); bit-identical to the oracle's
 * generator.  Writes `count` samples of `format` starting at global sample
 * `first` into device memory `dst_dev`. */
int sp_synth_fill(sp_engine *e, void *dst_dev, int format, uint64_t first, uint64_t count,
                  uint64_t total_samples, uint64_t seed);
void sp_synth_lut(int16_t *lut4096);

/* Per-launch timing of the dominant (render) kernel: keep `slots` CUDA event pairs and
 * record one around every render-kernel launch (ring).  sp_profile_read() synchronises and
 * returns up to `max` most recent durations in milliseconds (oldest first) and clears the ring. */
int sp_profile_enable(sp_engine *e, int slots);
/* Bracket only every `every`-th render-kernel launch (an event pair between dependent kernels costs the stream a few
 * microseconds: the ring then samples the launches instead of slowing every step).  Reset to 1 by sp_profile_enable. */
int sp_profile_sample(sp_engine *e, int every);
int sp_profile_read(sp_engine *e, float *ms, int max);

/* Introspection used by the bench / tests. */
int sp_device_count(sp_engine *e);
int sp_sm_count(sp_engine *e);
const char *sp_kernel_plan(sp_engine *e, int format, int n, int channel_mode); /* human readable */

#ifdef __cplusplus
}
#endif
#endif /* SPECTRO_B200_H */
