/* node_api_min.h — hand-declared subset of Node's <node_api.h> (N-API v6), ONLY used to
 * syntax-check ycnr_napi.cc in images that have no Node headers (g++ -fsyntax-only
 * -DYCNR_NAPI_MIN).  A real build includes <node_api.h> from node-gyp instead. */
#ifndef YCNR_NODE_API_MIN_H
#define YCNR_NODE_API_MIN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* data, void* hint);
typedef struct {
  const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
  napi_value value; int attributes; void* data;
} napi_property_descriptor;
typedef struct napi_module {
  int nm_version; unsigned int nm_flags; const char* nm_filename;
  napi_value (*nm_register_func)(napi_env, napi_value); const char* nm_modname; void* nm_priv; void* reserved[4];
} napi_module;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t*);
napi_status napi_get_value_double(napi_env, napi_value, double*);
napi_status napi_get_value_bool(napi_env, napi_value, bool*);
napi_status napi_get_value_int64(napi_env, napi_value, int64_t*);
napi_status napi_create_array(napi_env, napi_value*);
napi_status napi_set_element(napi_env, napi_value object, uint32_t index, napi_value value);
napi_status napi_get_element(napi_env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_array_length(napi_env, napi_value value, uint32_t* result);
napi_status napi_get_named_property(napi_env, napi_value object, const char* name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* name, bool* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* name, napi_value value);
napi_status napi_create_object(napi_env, napi_value*);
napi_status napi_create_int32(napi_env, int32_t, napi_value*);
napi_status napi_create_int64(napi_env, int64_t, napi_value*);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_get_undefined(napi_env, napi_value*);
napi_status napi_create_external(napi_env, void* data, napi_finalize, void* hint, napi_value*);
napi_status napi_get_value_external(napi_env, napi_value, void** result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_define_properties(napi_env, napi_value object, size_t count, const napi_property_descriptor*);
void napi_module_register(napi_module*);
#define NAPI_MODULE_X(modname, regfunc, priv, flags) \
  static napi_module _module = {1, flags, __FILE__, regfunc, #modname, priv, {0}}; \
  static void _register_##modname(void) __attribute__((constructor)); \
  static void _register_##modname(void) { napi_module_register(&_module); }
#define NAPI_MODULE(modname, regfunc) NAPI_MODULE_X(modname, regfunc, NULL, 0)
#ifdef __cplusplus
}
#endif
#endif
