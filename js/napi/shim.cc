// js/napi/shim.cc — N-API binding of include/rfx.h for a box that has Node + node_api.h.
//
// NOT COMPILED IN THIS IMAGE: there is no `node`, no Node headers and no JS engine here, so this
// file is the reference-side binding a maintainer would add, kept deliberately thin: every exported function
// forwards 1:1 to one `rfx_*` entry point of librfx.so and throws a JS Error carrying rfx_last_error() on a
// non-zero status.  Build (on such a box):
//   c++ -shared -fPIC -I$(node -p "require('node:path').dirname(process.execPath)")/../include/node \
//       -I../../include shim.cc -L../../realism_effects_b200/csrc -lrfx -o rfx_napi.node
#include <node_api.h>

#include <cstring>
#include <string>

#include "rfx.h"

namespace {

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, nullptr, "N-API call failed: " #call); return nullptr; } } while (0)

napi_value throw_status(napi_env env, rfx_ctx* ctx, rfx_status st, const char* what) {
  std::string msg = std::string(what) + ": rfx status " + std::to_string((int)st) + " — " + (ctx ? rfx_last_error(ctx) : "no context");
  napi_throw_error(env, nullptr, msg.c_str());
  return nullptr;
}
template <class T>
T* unwrap(napi_env env, napi_value v) { void* p = nullptr; napi_get_value_external(env, v, &p); return (T*)p; }
napi_value external(napi_env env, void* p) { napi_value v; napi_create_external(env, p, nullptr, nullptr, &v); return v; }

// Float32Array(16) -> float[16]
bool read_f32(napi_env env, napi_value v, float* dst, size_t n) {
  napi_typedarray_type t; size_t len; void* data; napi_value ab; size_t off;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok || t != napi_float32_array || len < n) return false;
  std::memcpy(dst, data, n * sizeof(float));
  return true;
}
bool read_camera(napi_env env, napi_value o, rfx_camera* c) {  // {projection, projectionInverse, matrixWorld, matrixWorldInverse, near, far}
  napi_value v; double d;
  if (napi_get_named_property(env, o, "projection", &v) != napi_ok || !read_f32(env, v, c->projection, 16)) return false;
  if (napi_get_named_property(env, o, "projectionInverse", &v) != napi_ok || !read_f32(env, v, c->projection_inverse, 16)) return false;
  if (napi_get_named_property(env, o, "matrixWorld", &v) != napi_ok || !read_f32(env, v, c->camera_matrix_world, 16)) return false;
  if (napi_get_named_property(env, o, "matrixWorldInverse", &v) != napi_ok || !read_f32(env, v, c->view_matrix, 16)) return false;
  napi_get_named_property(env, o, "near", &v); napi_get_value_double(env, v, &d); c->near_plane = (float)d;
  napi_get_named_property(env, o, "far", &v); napi_get_value_double(env, v, &d); c->far_plane = (float)d;
  c->perspective = 1;
  return true;
}

// ctxCreate(device) -> external
napi_value CtxCreate(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  int32_t dev = 0; napi_get_value_int32(env, argv[0], &dev);
  rfx_ctx* ctx = nullptr;
  rfx_status st = rfx_ctx_create(dev, &ctx);
  if (st != RFX_OK) return throw_status(env, nullptr, st, "rfx_ctx_create");
  return external(env, ctx);
}
napi_value CtxDestroy(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ctx_destroy(unwrap<rfx_ctx>(env, argv[0]));
  return nullptr;
}
// blueNoiseSet(ctx, Uint8Array rgba, size)
napi_value BlueNoiseSet(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ctx* ctx = unwrap<rfx_ctx>(env, argv[0]);
  napi_typedarray_type t; size_t len; void* data; napi_value ab; size_t off;
  NAPI_OK(napi_get_typedarray_info(env, argv[1], &t, &len, &data, &ab, &off));
  uint32_t size = 0; napi_get_value_uint32(env, argv[2], &size);
  rfx_status st = rfx_blue_noise_set(ctx, (const uint8_t*)data, size, size);
  return st == RFX_OK ? nullptr : throw_status(env, ctx, st, "rfx_blue_noise_set");
}
// chainCreate(ctx, optionsObject) -> external ; options keys = SSGIOptions.js names
napi_value ChainCreate(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ctx* ctx = unwrap<rfx_ctx>(env, argv[0]);
  rfx_ssgi_chain_options o{};
  auto num = [&](const char* k, double dflt) { napi_value v; double d = dflt; bool has = false; napi_has_named_property(env, argv[1], k, &has);
                                               if (has) { napi_get_named_property(env, argv[1], k, &v); napi_get_value_double(env, v, &d); } return d; };
  o.width = (uint32_t)num("width", 0); o.height = (uint32_t)num("height", 0);
  o.denoise_iterations = (int32_t)num("denoiseIterations", 1); o.steps = (int32_t)num("steps", 20); o.refine_steps = (int32_t)num("refineSteps", 5);
  o.distance = (float)num("distance", 10); o.thickness = (float)num("thickness", 10); o.env_blur = (float)num("envBlur", 0.5);
  o.radius = (float)num("radius", 3); o.phi = (float)num("phi", 0.5); o.luma_phi = (float)num("lumaPhi", 5); o.depth_phi = (float)num("depthPhi", 2);
  o.normal_phi = (float)num("normalPhi", 50); o.roughness_phi = (float)num("roughnessPhi", 50); o.specular_phi = (float)num("specularPhi", 50);
  o.ssgi_flags = (uint32_t)num("flags", RFX_SSGI_USE_DIRECT_LIGHT); o.mode = (int32_t)num("mode", RFX_MODE_SSGI);
  o.blue_noise_start = (int32_t)num("blueNoiseStart", 1234567);
  rfx_ssgi_chain* ch = nullptr;
  rfx_status st = rfx_ssgi_chain_create(ctx, &o, &ch);
  return st == RFX_OK ? external(env, ch) : throw_status(env, ctx, st, "rfx_ssgi_chain_create");
}
// chainRenderHost(ctx, chain, camera, depth:Float32Array, gbuffer:Float32Array, velocity:Float32Array, direct:Uint16Array|null,
//                 cameraPos:Float32Array(3), cameraMoved:boolean, out:Float32Array[, sync:boolean = true])
napi_value ChainRenderHost(napi_env env, napi_callback_info info) {
  size_t argc = 11; napi_value argv[11]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ctx* ctx = unwrap<rfx_ctx>(env, argv[0]);
  rfx_ssgi_chain* ch = unwrap<rfx_ssgi_chain>(env, argv[1]);
  rfx_ssgi_host_frame f{};
  if (!read_camera(env, argv[2], &f.cam)) { napi_throw_type_error(env, nullptr, "camera: expected Float32Array(16) matrices"); return nullptr; }
  auto ptr = [&](napi_value v) -> void* { napi_valuetype vt; napi_typeof(env, v, &vt); if (vt == napi_null || vt == napi_undefined) return nullptr;
                                          napi_typedarray_type t; size_t len; void* data; napi_value ab; size_t off;
                                          return napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) == napi_ok ? data : nullptr; };
  f.depth = (const float*)ptr(argv[3]); f.gbuffer = (const float*)ptr(argv[4]); f.velocity = (const float*)ptr(argv[5]);
  f.direct_light = (const uint16_t*)ptr(argv[6]);
  read_f32(env, argv[7], f.camera_pos, 3);
  bool moved = true; napi_get_value_bool(env, argv[8], &moved); f.camera_moved = moved;
  f.out_composed = (float*)ptr(argv[9]);
  // `sync` selects the synchronous call; otherwise the frame is only enqueued (pipelined H2D / kernels / D2H) and the caller
  // pairs it with chainWaitHost(chain, 1) - the typed arrays of a frame must stay alive and untouched until it completed.
  bool sync = true;
  if (argc > 10) napi_get_value_bool(env, argv[10], &sync);
  rfx_status st = sync ? rfx_ssgi_chain_render_host(ch, &f) : rfx_ssgi_chain_submit_host(ch, &f);
  return st == RFX_OK ? nullptr : throw_status(env, ctx, st, sync ? "rfx_ssgi_chain_render_host" : "rfx_ssgi_chain_submit_host");
}
// chainWaitHost(ctx, chain, maxInFlight)
napi_value ChainWaitHost(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ctx* ctx = unwrap<rfx_ctx>(env, argv[0]);
  int32_t n = 0; napi_get_value_int32(env, argv[2], &n);
  rfx_status st = rfx_ssgi_chain_wait_host(unwrap<rfx_ssgi_chain>(env, argv[1]), n);
  return st == RFX_OK ? nullptr : throw_status(env, ctx, st, "rfx_ssgi_chain_wait_host");
}
napi_value ChainReset(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ssgi_chain_reset(unwrap<rfx_ssgi_chain>(env, argv[0]));
  return nullptr;
}
napi_value ChainDestroy(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rfx_ssgi_chain_destroy(unwrap<rfx_ssgi_chain>(env, argv[0]));
  return nullptr;
}

napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"ctxCreate", nullptr, CtxCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"ctxDestroy", nullptr, CtxDestroy, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"blueNoiseSet", nullptr, BlueNoiseSet, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"chainCreate", nullptr, ChainCreate, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"chainRenderHost", nullptr, ChainRenderHost, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"chainWaitHost", nullptr, ChainWaitHost, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"chainReset", nullptr, ChainReset, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"chainDestroy", nullptr, ChainDestroy, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
