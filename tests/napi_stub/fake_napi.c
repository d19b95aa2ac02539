/*
 * fake_napi.c — a minimal Node-API runtime for testing js/spectro_napi.c without Node.js (TEST INFRASTRUCTURE).
 * Values are plain C structs (number, boolean, undefined, object with named properties, ArrayBuffer, TypedArray,
 * external, function); the fk_* functions let a ctypes harness build arguments, call exported functions and read
 * results.  Type checks mirror Node's: asking an ArrayBuffer for typed-array info (or the reverse) fails with the
 * documented status, a pending exception makes the call return NULL.  Nothing is freed (short-lived test process).
 */
#include "node_api.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { K_UNDEF, K_NUM, K_BOOL, K_OBJ, K_AB, K_TA, K_EXT, K_FN, K_ARR };
struct prop { char *key; napi_value val; struct prop *next; };
struct napi_value__ {
    int kind; double num; int b;
    struct prop *props;                                   /* K_OBJ */
    void *data; size_t len;                               /* K_AB: bytes; K_TA: element pointer / element count; K_EXT: pointer */
    napi_typedarray_type tt; napi_value ab; size_t off;   /* K_TA */
    napi_callback cb;                                     /* K_FN */
    napi_value *items; uint32_t nitems;                   /* K_ARR */
};
struct napi_env__ { char err[512]; int pending; struct napi_value__ undef; };
struct napi_callback_info__ { size_t argc; napi_value *argv; };

static napi_value nv(int kind) { napi_value v = (napi_value)calloc(1, sizeof *v); v->kind = kind; return v; }
static const size_t ELEM[] = { 1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8 };

napi_status napi_get_named_property(napi_env env, napi_value o, const char *k, napi_value *r)
{
    if (!o || o->kind != K_OBJ) return napi_object_expected;
    for (struct prop *p = o->props; p; p = p->next) if (!strcmp(p->key, k)) { *r = p->val; return napi_ok; }
    *r = &env->undef;
    return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value o, const char *k, napi_value v)
{
    (void)env;
    if (!o || o->kind != K_OBJ) return napi_object_expected;
    for (struct prop *p = o->props; p; p = p->next) if (!strcmp(p->key, k)) { p->val = v; return napi_ok; }
    struct prop *p = (struct prop *)calloc(1, sizeof *p);
    p->key = strdup(k); p->val = v; p->next = o->props; o->props = p;
    return napi_ok;
}
napi_status napi_get_value_double(napi_env env, napi_value v, double *r) { (void)env; if (!v || v->kind != K_NUM) return napi_number_expected; *r = v->num; return napi_ok; }
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t *r) { (void)env; if (!v || v->kind != K_NUM) return napi_number_expected; *r = (int32_t)v->num; return napi_ok; }
napi_status napi_get_value_bool(napi_env env, napi_value v, bool *r) { (void)env; if (!v || v->kind != K_BOOL) return napi_boolean_expected; *r = v->b != 0; return napi_ok; }
napi_status napi_coerce_to_bool(napi_env env, napi_value v, napi_value *r)
{
    (void)env;
    napi_value o = nv(K_BOOL);
    o->b = v && (v->kind == K_BOOL ? v->b : v->kind == K_NUM ? (v->num != 0 && v->num == v->num) : v->kind != K_UNDEF);
    *r = o;
    return napi_ok;
}
napi_status napi_get_cb_info(napi_env env, napi_callback_info ci, size_t *argc, napi_value *argv, napi_value *this_arg, void **data)
{
    size_t want = *argc;
    for (size_t i = 0; i < want; i++) argv[i] = i < ci->argc ? ci->argv[i] : &env->undef;
    *argc = ci->argc;
    if (this_arg) *this_arg = &env->undef;
    if (data) *data = NULL;
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char *code, const char *msg)
{
    (void)code;
    snprintf(env->err, sizeof env->err, "%s", msg ? msg : "");
    env->pending = 1;
    return napi_ok;
}
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg) { return napi_throw_error(env, code, msg); }
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg) { return napi_throw_error(env, code, msg); }
napi_status napi_create_external(napi_env env, void *data, napi_finalize f, void *hint, napi_value *r) { (void)env; (void)f; (void)hint; napi_value v = nv(K_EXT); v->data = data; *r = v; return napi_ok; }
napi_status napi_get_value_external(napi_env env, napi_value v, void **r) { (void)env; if (!v || v->kind != K_EXT) return napi_invalid_arg; *r = v->data; return napi_ok; }
napi_status napi_get_arraybuffer_info(napi_env env, napi_value v, void **data, size_t *len)
{
    (void)env;
    if (!v || v->kind != K_AB) return napi_invalid_arg;
    if (data) *data = v->data;
    if (len) *len = v->len;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type *type, size_t *length, void **data, napi_value *ab, size_t *off)
{
    (void)env;
    if (!v || v->kind != K_TA) return napi_invalid_arg;
    if (type) *type = v->tt;
    if (length) *length = v->len;
    if (data) *data = v->data;
    if (ab) *ab = v->ab;
    if (off) *off = v->off;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t n, void **data, napi_value *r)
{
    (void)env;
    napi_value v = nv(K_AB);
    v->data = calloc(n ? n : 1, 1); v->len = n;
    if (data) *data = v->data;
    *r = v;
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t off, napi_value *r)
{
    (void)env;
    if (!ab || ab->kind != K_AB) return napi_invalid_arg;
    if (off % ELEM[type] || off + length * ELEM[type] > ab->len) return napi_invalid_arg;      /* RangeError in Node */
    napi_value v = nv(K_TA);
    v->tt = type; v->len = length; v->ab = ab; v->off = off; v->data = (char *)ab->data + off;
    *r = v;
    return napi_ok;
}
napi_status napi_is_array(napi_env env, napi_value v, bool *r) { (void)env; *r = v && v->kind == K_ARR; return napi_ok; }
napi_status napi_get_array_length(napi_env env, napi_value v, uint32_t *r) { (void)env; if (!v || v->kind != K_ARR) return napi_array_expected; *r = v->nitems; return napi_ok; }
napi_status napi_get_element(napi_env env, napi_value v, uint32_t i, napi_value *r) { if (!v || v->kind != K_ARR) return napi_array_expected; *r = i < v->nitems ? v->items[i] : &env->undef; return napi_ok; }
napi_status napi_create_object(napi_env env, napi_value *r) { (void)env; *r = nv(K_OBJ); return napi_ok; }
napi_status napi_create_double(napi_env env, double d, napi_value *r) { (void)env; napi_value v = nv(K_NUM); v->num = d; *r = v; return napi_ok; }
napi_status napi_create_function(napi_env env, const char *name, size_t len, napi_callback cb, void *data, napi_value *r)
{
    (void)env; (void)name; (void)len; (void)data;
    napi_value v = nv(K_FN); v->cb = cb; *r = v;
    return napi_ok;
}

/* ---- harness side (ctypes) ---- */
napi_value fake_napi_module_init(napi_env env, napi_value exports);

napi_env fk_env_new(void) { napi_env e = (napi_env)calloc(1, sizeof *e); e->undef.kind = K_UNDEF; return e; }
napi_value fk_load(napi_env env) { napi_value ex = nv(K_OBJ); return fake_napi_module_init(env, ex); }
napi_value fk_number(double d) { napi_value v = nv(K_NUM); v->num = d; return v; }
napi_value fk_bool(int b) { napi_value v = nv(K_BOOL); v->b = b; return v; }
napi_value fk_object(void) { return nv(K_OBJ); }
napi_value fk_array(uint32_t n, napi_value *items) { napi_value v = nv(K_ARR); v->items = (napi_value *)malloc(sizeof(napi_value) * (n ? n : 1)); memcpy(v->items, items, sizeof(napi_value) * n); v->nitems = n; return v; }
napi_value fk_arraybuffer(const void *bytes, size_t n) { napi_value v = nv(K_AB); v->data = malloc(n ? n : 1); memcpy(v->data, bytes, n); v->len = n; return v; }
napi_value fk_typedarray(int type, napi_value ab, size_t off, size_t length) { napi_value r = NULL; napi_create_typedarray(NULL, (napi_typedarray_type)type, length, ab, off, &r); return r; }
void fk_set(napi_value o, const char *k, napi_value v) { napi_set_named_property(NULL, o, k, v); }
napi_value fk_get(napi_env env, napi_value o, const char *k) { napi_value r = NULL; napi_get_named_property(env, o, k, &r); return r; }
int fk_kind(napi_value v) { return v ? v->kind : -1; }
double fk_num(napi_value v) { return v->num; }
void *fk_data(napi_value v) { return v->data; }
size_t fk_len(napi_value v) { return v->len; }
int fk_ta_type(napi_value v) { return (int)v->tt; }
napi_value fk_call(napi_env env, napi_value fn, size_t argc, napi_value *argv)
{
    struct napi_callback_info__ ci = { argc, argv };
    env->pending = 0; env->err[0] = 0;
    napi_value r = fn->cb(env, &ci);
    return env->pending ? NULL : (r ? r : &env->undef);
}
const char *fk_error(napi_env env) { return env->pending ? env->err : NULL; }
