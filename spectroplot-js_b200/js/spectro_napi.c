/*
 * spectro_napi.c — thin N-API addon over the C ABI (include/spectro_b200.h).
 *
 * There is no Node.js toolchain in the build image (no node, no node_api.h), so the test suite compiles this file
 * against tests/napi_stub/node_api.h (the Node-API signatures it uses, -Wall -Wextra -Werror) and runs it on a small fake
 * Node-API runtime (tests/napi_stub/fake_napi.c, tests/test_napi_addon.py): create / render / destroy are called as
 * js/gpu_worker.js calls them and the outputs are checked against the reference-worker fixtures on the GPU.
 * It is the binding a maintainer of spectroplot-js adds so that lib/worker.js's renderFft(ctx)
 * (reference lib/worker.js:23-156) runs on the GPU; gpu_worker.js puts the worker message
 * protocol on top of it.  Build where Node is available:
 *     cc -shared -fPIC -I$(node -p "require('node-addon-api').include_dir || process.execPath+'/../../include/node'") \
 *        -I../../include spectro_napi.c -L../lib -lspectro_b200 -o spectro_napi.node
 *
 * Exports:  create(device | [devices]) -> external;  destroy(engine);
 *           render(engine, ctx) -> { cB_hist, c_hist, dBfs_min, dBfs_max, gauge_mins, gauge_maxs, gauge_amps, image }
 * where ctx = { buffer:ArrayBuffer, format:int, n, width, block_norm, gain, range, windowc:Float64Array,
 *               cmap:Uint8Array(len*3), channelMode:bool, waterfall:bool }.
 * It only unwraps ArrayBuffer / TypedArray pointers, calls sp_render, and wraps the outputs.
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>
#include "spectro_b200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call " failed"); return NULL; } } while (0)

static double get_num(napi_env env, napi_value obj, const char *key)
{
    napi_value v; double d = 0;
    napi_get_named_property(env, obj, key, &v);
    napi_get_value_double(env, v, &d);
    return d;
}
static int get_bool(napi_env env, napi_value obj, const char *key)
{
    napi_value v; bool b = false;
    napi_get_named_property(env, obj, key, &v);
    napi_coerce_to_bool(env, v, &v);
    napi_get_value_bool(env, v, &b);
    return b ? 1 : 0;
}

static napi_value Create(napi_env env, napi_callback_info info)
{
    /* create(device) or create([device, device, ...]): several devices make one engine whose render() shards a
     * message by frame range across them (sp_create with ndev > 1) */
    size_t argc = 1; napi_value argv[1]; int32_t devs[64] = { 0 }; uint32_t ndev = 1;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    bool is_array = false;
    if (argc > 0) napi_is_array(env, argv[0], &is_array);
    if (is_array) {
        NAPI_OK(napi_get_array_length(env, argv[0], &ndev));
        if (ndev < 1 || ndev > 64) { napi_throw_error(env, NULL, "create: 1 .. 64 devices"); return NULL; }
        for (uint32_t i = 0; i < ndev; i++) {
            napi_value v;
            NAPI_OK(napi_get_element(env, argv[0], i, &v));
            NAPI_OK(napi_get_value_int32(env, v, &devs[i]));
        }
    } else if (argc > 0) napi_get_value_int32(env, argv[0], &devs[0]);
    sp_engine *e = NULL;
    int rc = sp_create(&e, devs, (int)ndev);
    if (rc) { napi_throw_error(env, NULL, sp_last_error(NULL)); return NULL; }   /* no CPU fallback */
    napi_value ext;
    NAPI_OK(napi_create_external(env, e, NULL, NULL, &ext));
    return ext;
}

static napi_value Destroy(napi_env env, napi_callback_info info)
{
    size_t argc = 1; napi_value argv[1]; void *e = NULL;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    napi_get_value_external(env, argv[0], &e);
    sp_destroy((sp_engine *)e);
    return NULL;
}

static napi_value Render(napi_env env, napi_callback_info info)
{
    size_t argc = 2; napi_value argv[2]; void *eh = NULL;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    napi_get_value_external(env, argv[0], &eh);
    sp_engine *e = (sp_engine *)eh;
    napi_value ctx = argv[1], v;

    sp_request rq; memset(&rq, 0, sizeof rq);
    void *p; size_t len;
    napi_get_named_property(env, ctx, "buffer", &v);
    NAPI_OK(napi_get_arraybuffer_info(env, v, &p, &len));
    rq.buffer = p; rq.byte_length = len;
    rq.format = (int32_t)get_num(env, ctx, "format");
    rq.n = (int32_t)get_num(env, ctx, "n");
    rq.width = (int64_t)get_num(env, ctx, "width");
    rq.block_norm = get_num(env, ctx, "block_norm");
    rq.gain = get_num(env, ctx, "gain");
    rq.range = get_num(env, ctx, "range");
    napi_typedarray_type tt; napi_value ab; size_t off;
    napi_get_named_property(env, ctx, "windowc", &v);
    NAPI_OK(napi_get_typedarray_info(env, v, &tt, &len, &p, &ab, &off));       /* Float64Array(n) */
    if (tt != napi_float64_array) { napi_throw_type_error(env, NULL, "windowc must be a Float64Array"); return NULL; }
    if (rq.n < SP_MIN_N || rq.n > SP_MAX_N || (rq.n & (rq.n - 1))) { napi_throw_range_error(env, NULL, "Length is not a power of 2"); return NULL; }
    if (len < (size_t)rq.n) { napi_throw_range_error(env, NULL, "windowc is shorter than n"); return NULL; }   /* sp_render reads n doubles */
    rq.windowc = (const double *)p;
    napi_get_named_property(env, ctx, "cmap", &v);
    NAPI_OK(napi_get_typedarray_info(env, v, &tt, &len, &p, &ab, &off));       /* Uint8Array / Uint8ClampedArray(len*3) */
    if (tt != napi_uint8_array && tt != napi_uint8_clamped_array) { napi_throw_type_error(env, NULL, "cmap must be a Uint8Array or Uint8ClampedArray"); return NULL; }
    if (len % 3 != 0 || len < 6 || len > 3 * (size_t)SP_MAX_CMAP) { napi_throw_range_error(env, NULL, "cmap must hold 2 .. 4096 RGB triples"); return NULL; }
    rq.cmap_rgb = (const uint8_t *)p; rq.cmap_len = (int32_t)(len / 3);
    /* the outputs below are sized from width and n: range-check them before allocating (sp_render validates the rest) */
    if (rq.width < 1 || (double)rq.width * (double)rq.n > 17179869184.0) { napi_throw_range_error(env, NULL, "width out of range"); return NULL; }
    rq.channel_mode = get_bool(env, ctx, "channelMode");
    rq.waterfall = get_bool(env, ctx, "waterfall");

    /* outputs: fresh ArrayBuffers the worker shim transfers back (lib/worker.js:150-155) */
    sp_reply rp; memset(&rp, 0, sizeof rp);
    const size_t W = (size_t)rq.width, N = (size_t)rq.n;
    napi_value ab_img, ab_min, ab_max, ab_amp, ab_cb, ab_c;
    NAPI_OK(napi_create_arraybuffer(env, 4 * W * N, (void **)&rp.image, &ab_img));
    NAPI_OK(napi_create_arraybuffer(env, W, (void **)&rp.gauge_mins, &ab_min));
    NAPI_OK(napi_create_arraybuffer(env, W, (void **)&rp.gauge_maxs, &ab_max));
    NAPI_OK(napi_create_arraybuffer(env, W, (void **)&rp.gauge_amps, &ab_amp));
    NAPI_OK(napi_create_arraybuffer(env, 8 * SP_CB_HIST_SIZE, (void **)&rp.cB_hist, &ab_cb));
    NAPI_OK(napi_create_arraybuffer(env, 8 * (size_t)rq.cmap_len, (void **)&rp.c_hist, &ab_c));

    int rc = sp_render(e, &rq, &rp);
    if (rc) { napi_throw_error(env, NULL, sp_last_error(e)); return NULL; }      /* shim rejects the promise */

    napi_value out, t;
    NAPI_OK(napi_create_object(env, &out));
    napi_create_typedarray(env, napi_uint8_clamped_array, 4 * W * N, ab_img, 0, &t); napi_set_named_property(env, out, "image", t);
    napi_create_typedarray(env, napi_uint8_clamped_array, W, ab_min, 0, &t); napi_set_named_property(env, out, "gauge_mins", t);
    napi_create_typedarray(env, napi_uint8_clamped_array, W, ab_max, 0, &t); napi_set_named_property(env, out, "gauge_maxs", t);
    napi_create_typedarray(env, napi_uint8_clamped_array, W, ab_amp, 0, &t); napi_set_named_property(env, out, "gauge_amps", t);
    napi_create_typedarray(env, napi_biguint64_array, SP_CB_HIST_SIZE, ab_cb, 0, &t); napi_set_named_property(env, out, "cB_hist", t);
    napi_create_typedarray(env, napi_biguint64_array, (size_t)rq.cmap_len, ab_c, 0, &t); napi_set_named_property(env, out, "c_hist", t);
    napi_create_double(env, rp.dBfs_min, &t); napi_set_named_property(env, out, "dBfs_min", t);
    napi_create_double(env, rp.dBfs_max, &t); napi_set_named_property(env, out, "dBfs_max", t);
    napi_create_double(env, rp.device_ms, &t); napi_set_named_property(env, out, "device_ms", t);
    return out;
}

static napi_value Init(napi_env env, napi_value exports)
{
    napi_value fn;
    napi_create_function(env, "create", NAPI_AUTO_LENGTH, Create, NULL, &fn); napi_set_named_property(env, exports, "create", fn);
    napi_create_function(env, "destroy", NAPI_AUTO_LENGTH, Destroy, NULL, &fn); napi_set_named_property(env, exports, "destroy", fn);
    napi_create_function(env, "render", NAPI_AUTO_LENGTH, Render, NULL, &fn); napi_set_named_property(env, exports, "render", fn);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
