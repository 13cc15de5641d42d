// rfx_api.cu — implementation of the C ABI declared in include/rfx.h.
//
// Host-side responsibilities: argument validation, plane bookkeeping, the small per-context
// tables the kernels index by 8-bit blue-noise values, the env-map mip chain, and the native
// SSGI chain (the mirror of SSGIEffect.update / Denoiser.render frame logic).  No torch types, no
// exceptions across the boundary, no CPU fallback: every compute entry point launches a kernel.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rfx_kernels.h"

using namespace rfx;

struct rfx_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // blue noise
  uchar4* blue = nullptr;
  int blue_size = 0;
  // tables
  float2* rot_table = nullptr;  // [256]
  float* step_table = nullptr;  // [steps-1][256]
  int step_table_steps = 0;
  // env
  bool env_set = false;
  double env_total = 0.0;  // totalSumValue of the device-built tables (rfx_env_build)
  EnvD env{};
  std::vector<void*> env_allocs;
  // fast-math kernel variants (SFU lg2/ex2 instead of libm polynomials); 0 selects the exact-libm variants
  int fast_math = 1;
  // scratch: decoded G-buffer (float4 normal.xyz + roughness) for the fast Poisson kernel
  void* nrd = nullptr;
  int nrd_w = 0, nrd_h = 0;
  size_t nrd_pitch = 0;
  bool nrd_reuse = false;  // set by the native chain: the scratch already holds this frame's decode
  // scratch: view-space z plane for the SSGI march
  void* viewz = nullptr;
  int viewz_w = 0, viewz_h = 0;
  size_t viewz_pitch = 0;
  bool viewz_reuse = false;  // set by the native chain for the 2nd.. row block of a frame
  int k1_phase = 0;          // set by the native chain: 0 fused K1, 1 ray march only, 2 shading from the march records
  void* k1rec = nullptr;     // march records, 2 x float4 per pixel
  size_t k1rec_pitch = 0;
  int k1rec_w = 0, k1rec_h = 0;
  const RowSegs* segs_override = nullptr;  // set by the native chain: all owned row blocks in ONE launch
  int k3_tma = 1;     // RFX_K3_TMA=0 disables the TMA-staged tap tiles of the Poisson passes >= 1 (same bytes out; measured 0.39 vs 0.42 ms per pass at 4K)
  int k1_batch = 4;   // RFX_K1_BATCH: march steps fetched together (1, 2, 4); measured at 4K: 1.195 / 1.119 / 1.079 ms
  int compose_mode = 0;  // RFX_COMPOSE_MODE: arithmetic of the fused K4: 0 IEEE (default: 4.4e-4 of the 4K pixels outside 1e-3), 1 SFU, 2 SFU + Newton (both 6.9e-4, 0.06 ms faster)
  int debug_no_a_carry = 0;  // RFX_DEBUG_NO_A_CARRY=1: groups leave discarded texels of the A target as they are (the pre-fix behaviour; shows what tests/test_gpu_chain.py's in-process group test catches)
  int k1_vz_tiled = 0;  // RFX_K1_VZ_TILED=1: 8x4-tiled viewZ scratch for the fast K1 (experiment, profiles/r02_optimisation_log.txt)
  int legacy_k1 = 0;  // RFX_LEGACY_K1=1 in the environment: the round-1 fast K1 kernel (A/B timing)
  const PeerPV* peer_accumulated = nullptr;  // set by the native chain in a row-sharded group: K1's `accumulated` rows live on their owners
};

static rfx_status fail(rfx_ctx* c, rfx_status st, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return st;
}
#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(ctx, RFX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static const float kPi = 3.1415926535897932384626433832795f;

extern "C" {

int rfx_version(void) { return RFX_VERSION; }
uint32_t rfx_format_bytes(int32_t f) { return f == RFX_FMT_R32F ? 4u : f == RFX_FMT_RGBA32F ? 16u : f == RFX_FMT_RGBA16F ? 8u : f == RFX_FMT_RGBA8 ? 4u : 0u; }

rfx_status rfx_ctx_create(int device, rfx_ctx** out) {
  if (!out) return RFX_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return RFX_ERR_CUDA;
  rfx_ctx* ctx = new rfx_ctx();
  ctx->device = device;
  if (const char* e = getenv("RFX_LEGACY_K1")) ctx->legacy_k1 = atoi(e);
  if (const char* e = getenv("RFX_K3_TMA")) ctx->k3_tma = atoi(e);
  if (const char* e = getenv("RFX_K1_BATCH")) ctx->k1_batch = atoi(e);
  if (const char* e = getenv("RFX_K1_VZ_TILED")) ctx->k1_vz_tiled = atoi(e);
  if (const char* e = getenv("RFX_DEBUG_NO_A_CARRY")) ctx->debug_no_a_carry = atoi(e);
  if (const char* e = getenv("RFX_COMPOSE_MODE")) ctx->compose_mode = atoi(e);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return RFX_ERR_CUDA;
  }
  // (sin, cos) of the fp32 angle (k/255)*2*pi, correctly rounded: the 256 values blueNoise().r/.g can take
  float2 rot[256];
  for (int k = 0; k < 256; k++) {
    volatile float r = (float)k / 255.0f;
    volatile float r2 = r * 2.0f;
    volatile float ang = r2 * kPi;
    rot[k].x = (float)std::sin((double)ang);
    rot[k].y = (float)std::cos((double)ang);
  }
  if (cudaMalloc(&ctx->rot_table, sizeof rot) != cudaSuccess || cudaMemcpy(ctx->rot_table, rot, sizeof rot, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RFX_ERR_CUDA;
  }
  *out = ctx;
  return RFX_OK;
}

static void free_env(rfx_ctx* ctx) {
  for (void* p : ctx->env_allocs) cudaFree(p);
  ctx->env_allocs.clear();
  ctx->env_set = false;
  ctx->env = EnvD{};
}

void rfx_ctx_destroy(rfx_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_env(ctx);
  cudaFree(ctx->blue);
  cudaFree(ctx->rot_table);
  cudaFree(ctx->step_table);
  cudaFree(ctx->nrd);
  cudaFree(ctx->viewz);
  cudaFree(ctx->k1rec);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
rfx_status rfx_ctx_set_fast_math(rfx_ctx* ctx, int32_t enable) {
  if (!ctx) return RFX_ERR_INVALID_ARG;
  ctx->fast_math = enable ? 1 : 0;
  return RFX_OK;
}
const char* rfx_last_error(const rfx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* rfx_ctx_stream(rfx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
rfx_status rfx_ctx_sync(rfx_ctx* ctx) {
  if (!ctx) return RFX_ERR_INVALID_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  return RFX_OK;
}
uint64_t rfx_launch_count(const rfx_ctx* ctx) { return ctx ? ctx->launches : 0; }

rfx_status rfx_blue_noise_set(rfx_ctx* ctx, const uint8_t* rgba8, uint32_t w, uint32_t h) {
  if (!ctx || !rgba8 || w != h || w == 0) return fail(ctx, RFX_ERR_INVALID_ARG, "blue noise must be a square RGBA8 image");
  CU(cudaSetDevice(ctx->device));
  cudaFree(ctx->blue);
  ctx->blue = nullptr;
  CU(cudaMalloc(&ctx->blue, (size_t)w * h * 4));
  CU(cudaMemcpy(ctx->blue, rgba8, (size_t)w * h * 4, cudaMemcpyHostToDevice));
  ctx->blue_size = (int)w;
  return RFX_OK;
}

// ---- planes -----------------------------------------------------------------------------
rfx_status rfx_plane_alloc(rfx_ctx* ctx, int32_t format, uint32_t w, uint32_t h, rfx_plane* out) {
  if (!ctx || !out || w == 0 || h == 0 || rfx_format_bytes(format) == 0) return fail(ctx, RFX_ERR_INVALID_ARG, "plane_alloc: bad arguments");
  CU(cudaSetDevice(ctx->device));
  size_t pitch = ((size_t)w * rfx_format_bytes(format) + 255) & ~(size_t)255;
  void* p = nullptr;
  CU(cudaMalloc(&p, pitch * h));
  CU(cudaMemsetAsync(p, 0, pitch * h, ctx->stream));  // targets are zero after allocation
  CU(cudaStreamSynchronize(ctx->stream));
  out->ptr = p; out->width = w; out->height = h; out->pitch = pitch; out->format = format; out->_reserved = 0;
  return RFX_OK;
}
rfx_status rfx_plane_free(rfx_ctx* ctx, rfx_plane* p) {
  if (!ctx || !p) return RFX_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaFree(p->ptr));
  p->ptr = nullptr;
  return RFX_OK;
}
static cudaStream_t pick(rfx_ctx* ctx, void* s) { return s ? (cudaStream_t)s : ctx->stream; }
rfx_status rfx_plane_clear(rfx_ctx* ctx, void* stream, const rfx_plane* p) {
  if (!ctx || !p || !p->ptr) return RFX_ERR_INVALID_ARG;
  CU(cudaMemsetAsync(p->ptr, 0, p->pitch * p->height, pick(ctx, stream)));
  return RFX_OK;
}
rfx_status rfx_plane_upload(rfx_ctx* ctx, void* stream, const rfx_plane* dst, const void* host, uint64_t host_pitch) {
  if (!ctx || !dst || !dst->ptr || !host) return RFX_ERR_INVALID_ARG;
  size_t row = (size_t)dst->width * rfx_format_bytes(dst->format);
  if (host_pitch == 0) host_pitch = row;
  if (host_pitch == row && dst->pitch == row) CU(cudaMemcpyAsync(dst->ptr, host, row * dst->height, cudaMemcpyHostToDevice, pick(ctx, stream)));  // one contiguous DMA
  else CU(cudaMemcpy2DAsync(dst->ptr, dst->pitch, host, host_pitch, row, dst->height, cudaMemcpyHostToDevice, pick(ctx, stream)));
  return RFX_OK;
}
rfx_status rfx_plane_download(rfx_ctx* ctx, void* stream, const rfx_plane* src, void* host, uint64_t host_pitch) {
  if (!ctx || !src || !src->ptr || !host) return RFX_ERR_INVALID_ARG;
  size_t row = (size_t)src->width * rfx_format_bytes(src->format);
  if (host_pitch == 0) host_pitch = row;
  if (host_pitch == row && src->pitch == row) CU(cudaMemcpyAsync(host, src->ptr, row * src->height, cudaMemcpyDeviceToHost, pick(ctx, stream)));
  else CU(cudaMemcpy2DAsync(host, host_pitch, src->ptr, src->pitch, row, src->height, cudaMemcpyDeviceToHost, pick(ctx, stream)));
  return RFX_OK;
}
rfx_status rfx_plane_download_rows(rfx_ctx* ctx, void* stream, const rfx_plane* src, void* host, uint32_t row0, uint32_t row1) {
  if (!ctx || !src || !src->ptr || !host || row0 >= row1 || row1 > src->height) return RFX_ERR_INVALID_ARG;
  const size_t row = (size_t)src->width * rfx_format_bytes(src->format);
  const unsigned char* p = (const unsigned char*)src->ptr + (size_t)row0 * src->pitch;
  if (src->pitch == row) CU(cudaMemcpyAsync(host, p, row * (row1 - row0), cudaMemcpyDeviceToHost, pick(ctx, stream)));
  else CU(cudaMemcpy2DAsync(host, row, p, src->pitch, row, row1 - row0, cudaMemcpyDeviceToHost, pick(ctx, stream)));
  return RFX_OK;
}
rfx_status rfx_host_alloc(rfx_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return RFX_ERR_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return RFX_OK;
}
rfx_status rfx_host_free(rfx_ctx* ctx, void* p) {
  if (!ctx) return RFX_ERR_INVALID_ARG;
  CU(cudaFreeHost(p));
  return RFX_OK;
}

// ---- env map ----------------------------------------------------------------------------
rfx_status rfx_env_clear(rfx_ctx* ctx) {
  if (!ctx) return RFX_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_env(ctx);
  return RFX_OK;
}
rfx_status rfx_env_set(rfx_ctx* ctx, const rfx_env_desc* e) {
  if (!ctx || !e || !e->map_rgba16f || e->width == 0 || e->height == 0) return fail(ctx, RFX_ERR_INVALID_ARG, "env_set: bad arguments");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  free_env(ctx);
  int w = (int)e->width, h = (int)e->height, l = 0;
  EnvD d{};
  for (;;) {
    if (l >= 16) return fail(ctx, RFX_ERR_UNSUPPORTED, "env map too large");
    size_t pitch = ((size_t)w * 8 + 255) & ~(size_t)255;
    void* p = nullptr;
    CU(cudaMalloc(&p, pitch * h));
    ctx->env_allocs.push_back(p);
    d.mip[l] = PV{(const unsigned char*)p, w, h, (long long)pitch};
    if (l == 0) {
      CU(cudaMemcpy2DAsync(p, pitch, e->map_rgba16f, (size_t)w * 8, (size_t)w * 8, h, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      cudaError_t ce = launch_env_downsample(d.mip[l - 1], OutV{(unsigned char*)p, (long long)pitch}, w, h, ctx->stream);
      if (ce != cudaSuccess) return fail(ctx, RFX_ERR_CUDA, "env downsample: %s", cudaGetErrorString(ce));
      ctx->launches++;
    }
    l++;
    if (w == 1 && h == 1) break;
    w = w > 1 ? w >> 1 : 1;
    h = h > 1 ? h >> 1 : 1;
  }
  d.levels = l;
  d.size_x = (float)e->width;
  d.size_y = (float)e->height;
  d.total_sum_whole = e->total_sum_whole;
  d.total_sum_decimal = e->total_sum_decimal;
  ctx->env_total = (double)e->total_sum_whole + (double)e->total_sum_decimal;
  if (e->marginal && e->conditional) {
    float *m = nullptr, *c = nullptr;
    CU(cudaMalloc(&m, (size_t)e->height * 4));
    ctx->env_allocs.push_back(m);
    CU(cudaMalloc(&c, (size_t)e->width * e->height * 4));
    ctx->env_allocs.push_back(c);
    CU(cudaMemcpyAsync(m, e->marginal, (size_t)e->height * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(c, e->conditional, (size_t)e->width * e->height * 4, cudaMemcpyHostToDevice, ctx->stream));
    d.marginal = PV{(const unsigned char*)m, (int)e->height, 1, (long long)e->height * 4};  // image {width: height, height: 1}
    d.conditional = PV{(const unsigned char*)c, (int)e->width, (int)e->height, (long long)e->width * 4};
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->env = d;
  ctx->env_set = true;
  return RFX_OK;
}

// Env map + importance-sampling tables built ON THE DEVICE (replaces the reference's Web Worker,
// src/ssgi/utils/EquirectHdrInfoUniform.js:323-358 -> gatherData :149-245): uploads mip 0, builds the mip chain, the per-row
// cumulative distributions and the two inverse-CDF tables with the JS loops' summation order (bit-identical tables).
rfx_status rfx_env_build(rfx_ctx* ctx, const void* map_rgba16f, uint32_t width, uint32_t height, int32_t flip_y) {
  if (!ctx || !map_rgba16f || width == 0 || height == 0) return fail(ctx, RFX_ERR_INVALID_ARG, "env_build: bad arguments");
  rfx_env_desc e{};
  e.map_rgba16f = map_rgba16f; e.width = width; e.height = height;
  rfx_status st = rfx_env_set(ctx, &e);  // mip chain; no tables yet
  if (st != RFX_OK) return st;
  const size_t n = (size_t)width * height;
  float *cdf_c = nullptr, *cdf_m = nullptr, *marg = nullptr, *cond = nullptr;
  double *row_sum = nullptr, *total = nullptr;
  CU(cudaMalloc(&cdf_c, n * 4)); CU(cudaMalloc(&cdf_m, (size_t)height * 4)); CU(cudaMalloc(&row_sum, (size_t)height * 8)); CU(cudaMalloc(&total, 8));
  CU(cudaMalloc(&marg, (size_t)height * 4)); ctx->env_allocs.push_back(marg);
  CU(cudaMalloc(&cond, n * 4)); ctx->env_allocs.push_back(cond);
  cudaError_t ce = launch_env_cdf(ctx->env.mip[0], flip_y ? 1 : 0, cdf_c, cdf_m, row_sum, total, marg, cond, ctx->stream);
  ctx->launches += 3;
  double t = 0.0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&t, total, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(cdf_c); cudaFree(cdf_m); cudaFree(row_sum); cudaFree(total);
  if (ce != cudaSuccess) return fail(ctx, RFX_ERR_CUDA, "env_build: %s", cudaGetErrorString(ce));
  ctx->env.marginal = PV{(const unsigned char*)marg, (int)height, 1, (long long)height * 4};
  ctx->env.conditional = PV{(const unsigned char*)cond, (int)width, (int)height, (long long)width * 4};
  const double whole = (double)(int32_t)t;  // ~~totalSumValue (EquirectHdrInfoUniform.js:346-349)
  ctx->env.total_sum_whole = (float)whole;
  ctx->env.total_sum_decimal = (float)(t - whole);
  ctx->env_total = t;
  return RFX_OK;
}
// the tables of the current environment (host copies): marginal[height], conditional[width*height], totalSum; any pointer may be NULL
rfx_status rfx_env_tables_download(rfx_ctx* ctx, float* marginal, float* conditional, double* total_sum) {
  if (!ctx) return RFX_ERR_INVALID_ARG;
  if (!ctx->env_set || !ctx->env.marginal.p) return fail(ctx, RFX_ERR_NOT_READY, "env_tables_download: no importance-sampling tables");
  CU(cudaStreamSynchronize(ctx->stream));
  if (marginal) CU(cudaMemcpy(marginal, ctx->env.marginal.p, (size_t)ctx->env.marginal.w * 4, cudaMemcpyDeviceToHost));
  if (conditional) CU(cudaMemcpy(conditional, ctx->env.conditional.p, (size_t)ctx->env.conditional.w * ctx->env.conditional.h * 4, cudaMemcpyDeviceToHost));
  if (total_sum) *total_sum = ctx->env_total;
  return RFX_OK;
}

}  // extern "C"

// ---- helpers ------------------------------------------------------------------------------
static bool plane_ok(const rfx_plane* p, int fmt) {  // kernels address planes with 32-bit byte offsets
  return p && p->ptr && p->format == fmt && p->pitch * (uint64_t)p->height < (1ull << 32) && (p->pitch % rfx_format_bytes(fmt)) == 0;
}
static bool pv(const rfx_plane* p, int fmt, PV& out) {
  if (!plane_ok(p, fmt)) return false;
  out = PV{(const unsigned char*)p->ptr, (int)p->width, (int)p->height, (long long)p->pitch};
  return true;
}
static bool ov(const rfx_plane* p, int fmt, OutV& out) {
  if (!plane_ok(p, fmt)) return false;
  out = OutV{(unsigned char*)p->ptr, (long long)p->pitch};
  return true;
}
static void cam_to_dev(const rfx_camera& c, CamD& d) {
  memcpy(d.projection.m, c.projection, 64);
  memcpy(d.projection_inverse.m, c.projection_inverse, 64);
  memcpy(d.camera_matrix_world.m, c.camera_matrix_world, 64);
  memcpy(d.view_matrix.m, c.view_matrix, 64);
  d.near_plane = c.near_plane;
  d.far_plane = c.far_plane;
  d.perspective = c.perspective;
}
// pcg4d  (reference src/utils/shader/blue_noise.glsl:17-28), integer, bit-exact
static void pcg4d(uint32_t v[4]) {
  for (int i = 0; i < 4; i++) v[i] = v[i] * 1664525u + 1013904223u;
  v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
  for (int i = 0; i < 4; i++) v[i] ^= v[i] >> 16;
  v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}
static rfx_status blue_for(rfx_ctx* ctx, int index, BlueD& b) {
  if (!ctx->blue) return fail(ctx, RFX_ERR_NOT_READY, "blue noise texture not set (rfx_blue_noise_set)");
  uint32_t ui = (uint32_t)index;
  uint32_t s1[4] = {ui, ui * 15843u, ui * 31u + 4566u, ui * 2345u + 58585u};  // rng_initialize :13
  pcg4d(s1);
  b.tex = ctx->blue;
  b.size = ctx->blue_size;
  b.shift.sx = (int)((s1[0] % 0x0fffffffu) % (uint32_t)ctx->blue_size);  // shift2 :31-34
  b.shift.sy = (int)((s1[1] % 0x0fffffffu) % (uint32_t)ctx->blue_size);
  b.index = index;
  b.mask = (ctx->blue_size & (ctx->blue_size - 1)) == 0 ? ctx->blue_size - 1 : 0;
  return RFX_OK;
}
static void rows(uint32_t row0, uint32_t row1, uint32_t H, int& r0, int& r1) {
  if (row0 == 0 && row1 == 0) { r0 = 0; r1 = (int)H; }
  else { r0 = (int)row0; r1 = (int)(row1 > H ? H : row1); }
}
// the row segments of a launch: [r0,r1) of the call, or the multi-block table the native chain installed for this launch
static void set_segs(rfx_ctx* ctx, int r0, int r1, RowSegs& s) {
  if (ctx->segs_override) s = *ctx->segs_override;
  else s = make_segs(&r0, &r1, 1);
}
// mat4 * mat4 with the oracle's lowering: s = ((a0*b0 + a1*b1) + a2*b2) + a3*b3, no contraction
static void matmul(const float* A, const float* B, float* R) {
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      volatile float s = 0.0f;
      for (int k = 0; k < 4; k++) {
        volatile float p = A[k * 4 + r] * B[c * 4 + k];
        s = s + p;
      }
      R[c * 4 + r] = s;
    }
}
static rfx_status ensure_step_table(rfx_ctx* ctx, int steps) {
  if (ctx->step_table && ctx->step_table_steps == steps) return RFX_OK;
  if (steps < 1 || steps > 4096) return fail(ctx, RFX_ERR_INVALID_ARG, "steps out of range");
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->step_table);
  ctx->step_table = nullptr;
  int rows_n = steps + 4;  // rows 0..steps-2 hold cs(1..steps-1, b); spare zero rows: the fast march reads up to 3 steps ahead
  std::vector<float> t((size_t)rows_n * 256, 0.0f);
  for (int i = 1; i < steps; i++)
    for (int k = 0; k < 256; k++) {  // ssgi.frag:453   cs = 1. - exp(-0.25 * pow(float(i) + random.b - 0.5, 2.))
      volatile float b = (float)k / 255.0f;
      volatile float u = (float)i + b;
      volatile float v = u - 0.5f;
      volatile float p = (float)std::pow((double)v, 2.0);
      volatile float q = -0.25f * p;
      volatile float e = (float)std::exp((double)q);
      t[(size_t)(i - 1) * 256 + k] = 1.0f - e;
    }
  CU(cudaMalloc(&ctx->step_table, t.size() * 4));
  CU(cudaMemcpy(ctx->step_table, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
  ctx->step_table_steps = steps;
  return RFX_OK;
}
#define LAUNCHED(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, RFX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));      \
    ctx->launches++;                                                                                     \
  } while (0)

extern "C" {

rfx_status rfx_ssgi_trace_launch(rfx_ctx* ctx, void* stream, const rfx_ssgi_params* p, const rfx_plane* depth, const rfx_plane* gbuffer,
                                 const rfx_plane* velocity, const rfx_plane* direct_light, const rfx_plane* accumulated, const rfx_plane* out,
                                 uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "ssgi_trace: null argument");
  SsgiArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !pv(gbuffer, RFX_FMT_RGBA32F, a.gb) || !ov(out, RFX_FMT_RGBA32F, a.out))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "ssgi_trace: depth must be R32F, gbuffer/out RGBA32F");
  if (velocity && !pv(velocity, RFX_FMT_RGBA32F, a.velocity)) return fail(ctx, RFX_ERR_BAD_FORMAT, "ssgi_trace: velocity must be RGBA32F");
  if (direct_light && !pv(direct_light, RFX_FMT_RGBA16F, a.direct)) return fail(ctx, RFX_ERR_BAD_FORMAT, "ssgi_trace: direct light must be RGBA16F");
  if (accumulated && !pv(accumulated, RFX_FMT_RGBA32F, a.accumulated)) return fail(ctx, RFX_ERR_BAD_FORMAT, "ssgi_trace: accumulated must be RGBA32F");
  a.W = (int)out->width; a.H = (int)out->height;
  const int TW = a.depth.w, TH = a.depth.h;  // the input planes share one size; the render target may be smaller (resolutionScale < 1, SSGIPass.js:52-57)
  if (a.gb.w != TW || a.gb.h != TH || (a.velocity.p && (a.velocity.w != TW || a.velocity.h != TH)) ||
      (a.direct.p && (a.direct.w != TW || a.direct.h != TH)) || (a.accumulated.p && (a.accumulated.w != TW || a.accumulated.h != TH)))
    return fail(ctx, RFX_ERR_SIZE_MISMATCH, "ssgi_trace: depth / gbuffer / velocity / direct light / accumulated must have one size");
  if (a.W > TW || a.H > TH) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "ssgi_trace: the output may be smaller than the input planes (resolutionScale <= 1), not larger");
  a.scaled = a.W != TW || a.H != TH;
  if (p->steps < 1 || p->refine_steps < 0 || (p->mode != RFX_MODE_SSGI && p->mode != RFX_MODE_SSR)) return fail(ctx, RFX_ERR_INVALID_ARG, "ssgi_trace: bad steps/mode");
  rows(row0, row1, out->height, a.row0, a.row1);
  set_segs(ctx, a.row0, a.row1, a.segs);
  cam_to_dev(p->cam, a.cam);
  a.ray_distance = p->ray_distance; a.thickness = p->thickness; a.env_blur = p->env_blur; a.max_env_mip = p->max_env_map_mip_level;
  a.near_minus_far = p->cam.near_plane - p->cam.far_plane;  // SSGIPass.js:85-87
  a.far_minus_near = p->cam.far_plane - p->cam.near_plane;
  a.near_mul_far = p->cam.near_plane * p->cam.far_plane;
  a.steps = p->steps; a.refine_steps = p->refine_steps; a.mode = p->mode; a.flags = p->flags;
  if (p->flags & (RFX_SSGI_USE_ENVMAP | RFX_SSGI_IMPORTANCE_SAMPLING)) {
    if (!ctx->env_set) return fail(ctx, RFX_ERR_NOT_READY, "ssgi_trace: env map requested but rfx_env_set was not called");
    if ((p->flags & RFX_SSGI_IMPORTANCE_SAMPLING) && !ctx->env.marginal.p) return fail(ctx, RFX_ERR_NOT_READY, "ssgi_trace: importance sampling needs CDF tables");
    a.env = ctx->env;
  }
  rfx_status st = blue_for(ctx, p->blue_noise_index, a.blue);
  if (st != RFX_OK) return st;
  if (p->blue_noise_index == 0) return fail(ctx, RFX_ERR_UNSUPPORTED, "ssgi_trace: blue_noise_index 0 (tiled lookup) is not used by this pass");
  st = ensure_step_table(ctx, p->steps);
  if (st != RFX_OK) return st;
  a.rot_table = ctx->rot_table;
  a.step_table = ctx->step_table;
  a.fast = ctx->fast_math;
  {  // perspective sparsity pattern: [P00 0 P20 0; 0 P11 P21 0; 0 0 P22 P32; 0 0 -1 0] (column-major m[col*4+row])
    const float* M = p->cam.projection;
    a.proj_sparse = M[1] == 0.0f && M[2] == 0.0f && M[3] == 0.0f && M[4] == 0.0f && M[6] == 0.0f && M[7] == 0.0f && M[11] == -1.0f &&
                    M[12] == 0.0f && M[13] == 0.0f && M[15] == 0.0f;
  }
  // view-space z plane (scratch): the march taps read it instead of converting depth -> viewZ per tap
  if (!ctx->viewz || ctx->viewz_w != TW || ctx->viewz_h != TH) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->viewz);
    ctx->viewz = nullptr;
    ctx->viewz_pitch = ((size_t)TW * 4 + 255) & ~(size_t)255;
    CU(cudaMalloc(&ctx->viewz, std::max(ctx->viewz_pitch * TH, (size_t)((TW + 7) / 8) * ((TH + 3) / 4) * 128)));  // row-major or 8x4 tiles
    ctx->viewz_w = TW; ctx->viewz_h = TH;
  }
  {  // fast fused kernel: projection rows in texel units (0.5 W P00, 0.5 W P20, 0.5 H P11, 0.5 H P21), word pitch of the viewZ plane
    const float* M = p->cam.projection;
    const float hw = 0.5f * (float)a.W, hh = 0.5f * (float)a.H;
    a.ps_x0 = hw * M[0]; a.ps_x2 = hw * M[8]; a.ps_y1 = hh * M[5]; a.ps_y2 = hh * M[9]; a.ps_hw = hw; a.ps_hh = hh;
    a.vz_pitchw = (int)(ctx->viewz_pitch / 4);
    a.legacy_fast = ctx->legacy_k1 || a.scaled;  // the fused fast kernel addresses texels by pixel index: a scaled target takes the general kernel
    a.march_batch = ctx->k1_batch;
    a.vz_tiles_x = (TW + 7) / 8;
    a.vz_tiled = ctx->k1_vz_tiled && a.fast && !a.legacy_fast && ctx->k1_phase == 0;  // only the fused fast kernel reads the tiled layout
    if (ctx->peer_accumulated) a.acc_peer = *ctx->peer_accumulated;
  }
  a.phase = a.scaled ? 0 : ctx->k1_phase;
  if (a.phase != 0 && !a.fast) {  // the split phases exist for the fast variant: otherwise phase 1 is empty and phase 2 is the fused kernel
    if (a.phase == 1) return RFX_OK;
    a.phase = 0;
  }
  if (a.phase != 0) {
    if (!ctx->k1rec || ctx->k1rec_w != a.W || ctx->k1rec_h != a.H) {
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->k1rec);
      ctx->k1rec = nullptr;
      ctx->k1rec_pitch = (size_t)a.W * 32;
      CU(cudaMalloc(&ctx->k1rec, ctx->k1rec_pitch * a.H));
      ctx->k1rec_w = a.W; ctx->k1rec_h = a.H;
    }
    a.rec = (unsigned char*)ctx->k1rec;
    a.rec_pitch = (long long)ctx->k1rec_pitch;
  }
  if (!ctx->viewz_reuse && a.phase != 2) LAUNCHED(launch_viewz(a, OutV{(unsigned char*)ctx->viewz, (long long)ctx->viewz_pitch}, pick(ctx, stream)));
  a.viewz = PV{(const unsigned char*)ctx->viewz, TW, TH, (long long)ctx->viewz_pitch};
  LAUNCHED(launch_ssgi(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_temporal_reproject_launch(rfx_ctx* ctx, void* stream, const rfx_temporal_params* p, const rfx_plane* input, const rfx_plane* velocity,
                                         const rfx_plane* history0, const rfx_plane* history1, const rfx_plane* out0, const rfx_plane* out1,
                                         uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !input || !out0) return fail(ctx, RFX_ERR_INVALID_ARG, "temporal: null argument");
  TemporalArgs a{};
  if (p->texture_count != 1 && p->texture_count != 2) return fail(ctx, RFX_ERR_INVALID_ARG, "temporal: texture_count must be 1 or 2");
  a.input_half = input->format == RFX_FMT_RGBA16F;
  if (!pv(input, a.input_half ? RFX_FMT_RGBA16F : RFX_FMT_RGBA32F, a.input)) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: input must be RGBA32F or RGBA16F");
  if (p->input_type != RFX_INPUT_DIFFUSE && a.input_half) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: packed inputs must be RGBA32F");
  if (!pv(velocity, RFX_FMT_RGBA32F, a.velocity)) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: velocity must be RGBA32F");
  a.hist_f32 = history0 && history0->format == RFX_FMT_RGBA32F;  // the FloatType FramebufferTexture history of denoiseMode "full_temporal" / "temporal"
  const int hfmt = a.hist_f32 ? RFX_FMT_RGBA32F : RFX_FMT_RGBA16F;
  if (!pv(history0, hfmt, a.hist0)) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: history must be RGBA16F or RGBA32F");
  a.out_half = out0->format == RFX_FMT_RGBA16F;
  if (!ov(out0, a.out_half ? RFX_FMT_RGBA16F : RFX_FMT_RGBA32F, a.out0)) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: out must be RGBA32F or RGBA16F");
  if (p->texture_count == 2) {
    if (!pv(history1, hfmt, a.hist1) || !ov(out1, out0->format, a.out1)) return fail(ctx, RFX_ERR_BAD_FORMAT, "temporal: second plane missing / wrong format");
  }
  a.W = (int)out0->width; a.H = (int)out0->height;
  if (a.velocity.w != a.W || a.velocity.h != a.H || a.hist0.w != a.W || a.hist0.h != a.H || a.input.w > a.W || a.input.h > a.H)
    return fail(ctx, RFX_ERR_SIZE_MISMATCH, "temporal: plane sizes differ (only the input may be smaller: resolutionScale < 1)");
  a.in_scaled = a.input.w != a.W || a.input.h != a.H;
  if (a.in_scaled && a.input_half) return fail(ctx, RFX_ERR_UNSUPPORTED, "temporal: a scaled input is the RGBA32F SSGI target");
  rows(row0, row1, out0->height, a.row0, a.row1);
  set_segs(ctx, a.row0, a.row1, a.segs);
  cam_to_dev(p->cam, a.cam);
  memcpy(a.prev_view.m, p->prev_view_matrix, 64);
  memcpy(a.prev_world.m, p->prev_camera_matrix_world, 64);
  memcpy(a.prev_proj.m, p->prev_projection, 64);
  memcpy(a.prev_proj_inv.m, p->prev_projection_inverse, 64);
  matmul(p->prev_projection, p->prev_view_matrix, a.prev_proj_view.m);
  memcpy(a.camera_pos, p->camera_pos, 12);
  a.max_blend = p->max_blend; a.clamp_intensity = p->neighborhood_clamp_intensity; a.keep_data = p->keep_data; a.confidence_power = p->confidence_power;
  a.inv_w = (float)(1.0 / (double)a.W);  // TemporalReprojectPass.js:135: JS doubles, uploaded as float32
  a.inv_h = (float)(1.0 / (double)a.H);
  a.full_accumulate = p->full_accumulate; a.texture_count = p->texture_count; a.input_type = p->input_type; a.log_transform = p->log_transform;
  a.rs0 = p->reproject_specular[0]; a.rs1 = p->reproject_specular[1]; a.history_linear = p->history_linear;
  a.fast = ctx->fast_math;
  LAUNCHED(launch_temporal(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_poisson_denoise_launch(rfx_ctx* ctx, void* stream, const rfx_poisson_params* p, const rfx_plane* depth, const rfx_plane* gb,
                                      const rfx_plane* in0, const rfx_plane* in1, const rfx_plane* out0, const rfx_plane* out1, uint32_t row0,
                                      uint32_t row1) {
  if (!ctx || !p || !in0 || !out0) return fail(ctx, RFX_ERR_INVALID_ARG, "poisson: null argument");
  if (p->texture_count != 1 && p->texture_count != 2) return fail(ctx, RFX_ERR_INVALID_ARG, "poisson: texture_count must be 1 or 2");
  PoissonArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !pv(gb, RFX_FMT_RGBA32F, a.gb)) return fail(ctx, RFX_ERR_BAD_FORMAT, "poisson: depth R32F + gbuffer/normal RGBA32F required");
  a.in_half = in0->format == RFX_FMT_RGBA16F;
  if (!pv(in0, a.in_half ? RFX_FMT_RGBA16F : RFX_FMT_RGBA32F, a.in0)) return fail(ctx, RFX_ERR_BAD_FORMAT, "poisson: in0 must be RGBA32F or RGBA16F");
  if (!ov(out0, RFX_FMT_RGBA16F, a.out0)) return fail(ctx, RFX_ERR_BAD_FORMAT, "poisson: out must be RGBA16F");
  if (p->texture_count == 2) {
    if (!pv(in1, in0->format, a.in1) || !ov(out1, RFX_FMT_RGBA16F, a.out1)) return fail(ctx, RFX_ERR_BAD_FORMAT, "poisson: second plane missing / wrong format");
  } else {
    a.in1 = a.in0;
  }
  if (p->input_linear && !a.in_half) return fail(ctx, RFX_ERR_UNSUPPORTED, "poisson: LINEAR inputs must be RGBA16F");
  a.W = (int)out0->width; a.H = (int)out0->height;
  if (a.depth.w != a.W || a.depth.h != a.H || a.gb.w != a.W || a.gb.h != a.H || a.in0.w != a.W || a.in0.h != a.H)
    return fail(ctx, RFX_ERR_SIZE_MISMATCH, "poisson: plane sizes differ");
  if (out0->ptr == in0->ptr || (out1 && in1 && out1->ptr == in1->ptr)) return fail(ctx, RFX_ERR_INVALID_ARG, "poisson: in-place filtering is not allowed");
  rows(row0, row1, out0->height, a.row0, a.row1);
  set_segs(ctx, a.row0, a.row1, a.segs);
  a.radius = p->radius; a.phi = p->phi; a.luma_phi = p->luma_phi; a.depth_phi = p->depth_phi; a.normal_phi = p->normal_phi;
  a.roughness_phi = p->roughness_phi; a.specular_phi = p->specular_phi;
  a.texture_count = p->texture_count; a.spec0 = p->is_texture_specular[0]; a.spec1 = p->is_texture_specular[1];
  a.gbuffer_texture = p->gbuffer_texture; a.input_linear = p->input_linear;
  rfx_status st = blue_for(ctx, p->blue_noise_index, a.blue);
  if (st != RFX_OK) return st;
  if (p->blue_noise_index == 0) return fail(ctx, RFX_ERR_UNSUPPORTED, "poisson: blue_noise_index 0 is not used by this pass");
  a.rot_table = ctx->rot_table;
  const bool fast = ctx->fast_math && ((p->input_linear && a.in_half) || (!p->input_linear && !a.in_half));
  if (!fast) {
    LAUNCHED(launch_poisson(a, pick(ctx, stream)));
    return RFX_OK;
  }
  // fast variant: decode the G-buffer once into the context scratch (the native chain reuses it across the passes of a frame)
  if (!ctx->nrd || ctx->nrd_w != a.W || ctx->nrd_h != a.H) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->nrd);
    ctx->nrd = nullptr;
    ctx->nrd_pitch = ((size_t)a.W * 16 + 255) & ~(size_t)255;
    CU(cudaMalloc(&ctx->nrd, ctx->nrd_pitch * a.H));
    ctx->nrd_w = a.W; ctx->nrd_h = a.H;
    ctx->nrd_reuse = false;
  }
  if (!ctx->nrd_reuse)  // (the chain's first pass of a frame has the widest rows of all its passes, so its decode serves the later ones)
    LAUNCHED(launch_gbuffer_decode(a.gb, OutV{(unsigned char*)ctx->nrd, (long long)ctx->nrd_pitch}, a.W, a.H, p->gbuffer_texture ? 1 : 0, a.segs,
                                   (int)std::ceil(p->radius * std::max(1.0f, (float)a.H / (float)a.W)) + 1, pick(ctx, stream)));
  a.nrd = PV{(const unsigned char*)ctx->nrd, a.W, a.H, (long long)ctx->nrd_pitch};
  {
    const float SQ = 1.41421356237f;
    const float px[8] = {-1.0f, 0.0f, 1.0f, 0.0f, -0.25f * SQ, 0.25f * SQ, 0.25f * SQ, -0.25f * SQ};
    const float py[8] = {0.0f, -1.0f, 0.0f, 1.0f, -0.25f * SQ, -0.25f * SQ, 0.25f * SQ, 0.25f * SQ};
    for (int i = 0; i < 8; i++) { a.tap_ox[i] = px[i] / (float)a.W; a.tap_oy[i] = py[i] / (float)a.H; }  // offset / resolution
  }
  LAUNCHED(launch_poisson_fast(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_gi_compose_launch(rfx_ctx* ctx, void* stream, const rfx_compose_params* p, const rfx_plane* depth, const rfx_plane* gb,
                                 const rfx_plane* dgi, const rfx_plane* sgi, const rfx_plane* scene, const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "gi_compose: null argument");
  ComposeArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !pv(gb, RFX_FMT_RGBA32F, a.gb) || !ov(out, RFX_FMT_RGBA32F, a.out))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "gi_compose: depth R32F, gbuffer RGBA32F, out RGBA32F required");
  if (p->input_type != RFX_INPUT_DIFFUSE_SPECULAR && p->input_type != RFX_INPUT_DIFFUSE && p->input_type != RFX_INPUT_SPECULAR)
    return fail(ctx, RFX_ERR_INVALID_ARG, "gi_compose: bad input_type");
  // DenoiserComposePass.js:23-33: diffuseSpecular binds both GI textures, diffuse only the first, specular only the second
  const bool need_d = p->input_type != RFX_INPUT_SPECULAR, need_s = p->input_type != RFX_INPUT_DIFFUSE;
  const rfx_plane* first = need_d ? dgi : sgi;
  a.gi_f32 = first && first->format == RFX_FMT_RGBA32F;  // denoiseMode "full_temporal": the temporal pass's FloatType NEAREST targets
  const int gfmt = a.gi_f32 ? RFX_FMT_RGBA32F : RFX_FMT_RGBA16F;
  if (need_d && !pv(dgi, gfmt, a.diffuse)) return fail(ctx, RFX_ERR_BAD_FORMAT, "gi_compose: diffuse GI must be RGBA16F (or both RGBA32F)");
  if (need_s && !pv(sgi, gfmt, a.specular)) return fail(ctx, RFX_ERR_BAD_FORMAT, "gi_compose: specular GI must be RGBA16F (or both RGBA32F)");
  if (p->input_type == RFX_INPUT_SPECULAR && scene && !pv(scene, RFX_FMT_RGBA16F, a.scene)) return fail(ctx, RFX_ERR_BAD_FORMAT, "gi_compose: scene must be RGBA16F");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.depth.w != a.W || a.depth.h != a.H || a.gb.w != a.W || a.gb.h != a.H || (a.diffuse.p && (a.diffuse.w != a.W || a.diffuse.h != a.H)) ||
      (a.specular.p && (a.specular.w != a.W || a.specular.h != a.H)) || (a.scene.p && (a.scene.w != a.W || a.scene.h != a.H)))
    return fail(ctx, RFX_ERR_SIZE_MISMATCH, "gi_compose: plane sizes differ");
  rows(row0, row1, out->height, a.row0, a.row1);
  set_segs(ctx, a.row0, a.row1, a.segs);
  cam_to_dev(p->cam, a.cam);
  a.input_type = p->input_type;
  a.fast = ctx->fast_math;
  LAUNCHED(launch_gi_compose(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_ssgi_compose_launch(rfx_ctx* ctx, void* stream, const rfx_ssgi_compose_params* p, const rfx_plane* depth, const rfx_plane* gi, const rfx_plane* scene,
                                   const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "ssgi_compose: null argument");
  SsgiComposeArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !pv(gi, RFX_FMT_RGBA32F, a.gi) || !pv(scene, RFX_FMT_RGBA16F, a.scene) || !ov(out, RFX_FMT_RGBA16F, a.out))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "ssgi_compose: depth R32F, gi RGBA32F, scene/out RGBA16F required");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.depth.w != a.W || a.depth.h != a.H || a.gi.w != a.W || a.scene.w != a.W) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "ssgi_compose: plane sizes differ");
  rows(row0, row1, out->height, a.row0, a.row1);
  if (p) {
    a.use_fog = p->use_fog; a.fog_exp2 = p->fog_exp2; a.perspective = p->perspective; a.is_debug = p->is_debug;
    memcpy(a.fog_color, p->fog_color, 12);
    a.fog_near = p->fog_near; a.fog_far = p->fog_far; a.fog_density = p->fog_density; a.camera_near = p->camera_near; a.camera_far = p->camera_far;
  }
  LAUNCHED(launch_ssgi_compose(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_hbao_launch(rfx_ctx* ctx, void* stream, const rfx_hbao_params* p, const rfx_plane* depth, const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "hbao: null argument");
  HbaoArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !ov(out, RFX_FMT_RGBA16F, a.out)) return fail(ctx, RFX_ERR_BAD_FORMAT, "hbao: depth R32F, out RGBA16F required");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.depth.w != a.W || a.depth.h != a.H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "hbao: plane sizes differ");
  if (p->spp < 0) return fail(ctx, RFX_ERR_INVALID_ARG, "hbao: spp < 0");
  rows(row0, row1, out->height, a.row0, a.row1);
  memcpy(a.projection_view.m, p->projection_view, 64);
  memcpy(a.projection_inverse.m, p->projection_inverse, 64);
  memcpy(a.camera_matrix_world.m, p->camera_matrix_world, 64);
  a.ao_distance = p->ao_distance; a.distance_power = p->distance_power; a.bias = p->bias; a.thickness = p->thickness; a.spp = p->spp;
  rfx_status st = blue_for(ctx, p->blue_noise_index, a.blue);
  if (st != RFX_OK) return st;
  if (p->blue_noise_index == 0) return fail(ctx, RFX_ERR_UNSUPPORTED, "hbao: blue_noise_index 0 is not used by this pass");
  a.rot_table = ctx->rot_table;
  LAUNCHED(launch_hbao(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_ao_compose_launch(rfx_ctx* ctx, void* stream, const rfx_ao_compose_params* p, const rfx_plane* depth, const rfx_plane* ao,
                                 const rfx_plane* input, const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "ao_compose: null argument");
  AoComposeArgs a{};
  if (!pv(depth, RFX_FMT_R32F, a.depth) || !pv(ao, RFX_FMT_RGBA16F, a.ao) || !pv(input, RFX_FMT_RGBA16F, a.input) || !ov(out, RFX_FMT_RGBA16F, a.out))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "ao_compose: depth R32F, ao/input/out RGBA16F required");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.depth.w != a.W || a.depth.h != a.H || a.ao.w != a.W || a.input.w != a.W) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "ao_compose: plane sizes differ");
  rows(row0, row1, out->height, a.row0, a.row1);
  a.power = p->power;
  memcpy(a.color, p->color, 12);
  LAUNCHED(launch_ao_compose(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_motion_blur_launch(rfx_ctx* ctx, void* stream, const rfx_motion_blur_params* p, const rfx_plane* velocity, const rfx_plane* input,
                                  const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "motion_blur: null argument");
  MotionBlurArgs a{};
  if (!pv(velocity, RFX_FMT_RGBA32F, a.velocity) || !pv(input, RFX_FMT_RGBA16F, a.input) || !ov(out, RFX_FMT_RGBA16F, a.out))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "motion_blur: velocity RGBA32F, input/out RGBA16F required");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.velocity.w != a.W || a.velocity.h != a.H || a.input.w != a.W || a.input.h != a.H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "motion_blur: plane sizes differ");
  if (input->ptr == out->ptr) return fail(ctx, RFX_ERR_INVALID_ARG, "motion_blur: in-place is not allowed");
  if (p->samples < 1) return fail(ctx, RFX_ERR_INVALID_ARG, "motion_blur: samples < 1");
  rows(row0, row1, out->height, a.row0, a.row1);
  a.intensity = p->intensity; a.jitter = p->jitter; a.delta_time = p->delta_time; a.res_x = p->resolution[0]; a.res_y = p->resolution[1];
  a.samples = p->samples;
  rfx_status st = blue_for(ctx, p->frame, a.blue);
  if (st != RFX_OK) return st;
  LAUNCHED(launch_motion_blur(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_traa_compose_launch(rfx_ctx* ctx, void* stream, const rfx_plane* acc, const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "traa_compose: null argument");
  TraaComposeArgs a{};
  if (!pv(acc, RFX_FMT_RGBA16F, a.acc) || !ov(out, RFX_FMT_RGBA16F, a.out)) return fail(ctx, RFX_ERR_BAD_FORMAT, "traa_compose: RGBA16F planes required");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.acc.w != a.W || a.acc.h != a.H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "traa_compose: plane sizes differ");
  rows(row0, row1, out->height, a.row0, a.row1);
  LAUNCHED(launch_traa_compose(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_effects_launch(rfx_ctx* ctx, void* stream, const rfx_effects_params* p, const rfx_plane* input, const rfx_plane* depth, const rfx_plane* velocity,
                              const rfx_plane* out, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !input || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "effects: null argument");
  if (p->n_effects < 1 || p->n_effects > 4) return fail(ctx, RFX_ERR_INVALID_ARG, "effects: n_effects must be 1..4");
  EffectsArgs a{};
  if (!pv(input, RFX_FMT_RGBA16F, a.input) || !ov(out, RFX_FMT_RGBA16F, a.out)) return fail(ctx, RFX_ERR_BAD_FORMAT, "effects: input / out must be RGBA16F");
  if (input->ptr == out->ptr) return fail(ctx, RFX_ERR_INVALID_ARG, "effects: out may not alias input (neighbour taps)");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.input.w != a.W || a.input.h != a.H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "effects: plane sizes differ");
  for (int i = 0; i < p->n_effects; i++) {
    const int e = p->effects[i];
    a.effects[i] = e;
    if (e < RFX_FX_SHARPNESS || e > RFX_FX_SPARKLE) return fail(ctx, RFX_ERR_INVALID_ARG, "effects: unknown effect id %d", e);
    if (e == RFX_FX_GRADUAL_BACKGROUND && (!pv(depth, RFX_FMT_R32F, a.depth) || a.depth.w != a.W || a.depth.h != a.H))
      return fail(ctx, RFX_ERR_BAD_FORMAT, "effects: GradualBackground needs an R32F depth plane of the output size");
    if (e == RFX_FX_SPARKLE && (!pv(velocity, RFX_FMT_RGBA32F, a.velocity) || a.velocity.w != a.W || a.velocity.h != a.H))
      return fail(ctx, RFX_ERR_BAD_FORMAT, "effects: Sparkle needs an RGBA32F velocity plane of the output size");
  }
  a.n_effects = p->n_effects;
  cam_to_dev(p->cam, a.cam);
  a.texel_x = (float)(1.0 / a.W); a.texel_y = (float)(1.0 / a.H);
  a.sharpness = p->sharpness; a.alphax = p->alphax; a.alphay = p->alphay; a.aberration = p->aberration;
  memcpy(a.bg, p->background_color, 12);
  a.max_distance = p->max_distance; a.spread = p->spread; a.intensity = p->intensity; a.sparkle_perspective = p->sparkle_perspective;
  rows(row0, row1, out->height, a.row0, a.row1);
  LAUNCHED(launch_effects(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_taa_launch(rfx_ctx* ctx, void* stream, const rfx_taa_params* p, const rfx_plane* input, const rfx_plane* history, const rfx_plane* out,
                          uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !input || !out) return fail(ctx, RFX_ERR_INVALID_ARG, "taa: null argument");
  TaaArgs a{};
  if (!pv(input, RFX_FMT_RGBA16F, a.input) || !ov(out, RFX_FMT_RGBA8, a.out)) return fail(ctx, RFX_ERR_BAD_FORMAT, "taa: input RGBA16F, out RGBA8");
  a.W = (int)out->width; a.H = (int)out->height;
  if (a.input.w != a.W || a.input.h != a.H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "taa: plane sizes differ");
  if (p->camera_not_moved_frames != 0.0f && (!pv(history, RFX_FMT_RGBA8, a.history) || a.history.w != a.W || a.history.h != a.H))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "taa: an RGBA8 history plane of the output size is needed once the camera stands still");
  a.camera_not_moved_frames = p->camera_not_moved_frames; a.srgb_output = p->srgb_output;
  rows(row0, row1, out->height, a.row0, a.row1);
  LAUNCHED(launch_taa(a, pick(ctx, stream)));
  return RFX_OK;
}

rfx_status rfx_gbuffer_ingest_launch(rfx_ctx* ctx, void* stream, const rfx_ingest_params* p, const rfx_plane* albedo, const rfx_plane* normal,
                                     const rfx_plane* material, const rfx_plane* emissive, const rfx_plane* motion, const rfx_plane* depth,
                                     const rfx_plane* out_gbuffer, const rfx_plane* out_velocity, uint32_t row0, uint32_t row1) {
  if (!ctx || !p || !albedo || !normal || !material || !depth || (!out_gbuffer && !out_velocity)) return fail(ctx, RFX_ERR_INVALID_ARG, "gbuffer_ingest: null argument");
  IngestArgs a{};
  auto one_of = [](const rfx_plane* pl, int f0, int f1, PV& out, int& is_second) {
    if (pl && pl->format == f1) { is_second = 1; return pv(pl, f1, out); }
    is_second = 0;
    return pv(pl, f0, out);
  };
  if (!one_of(albedo, RFX_FMT_RGBA8, RFX_FMT_RGBA16F, a.albedo, a.albedo_half) || !one_of(material, RFX_FMT_RGBA8, RFX_FMT_RGBA16F, a.material, a.material_half) ||
      !one_of(normal, RFX_FMT_RGBA16F, RFX_FMT_RGBA32F, a.normal, a.normal_f32) || !pv(depth, RFX_FMT_R32F, a.depth) ||
      (emissive && !pv(emissive, RFX_FMT_RGBA16F, a.emissive)) || (motion && !one_of(motion, RFX_FMT_RGBA16F, RFX_FMT_RGBA32F, a.motion, a.motion_f32)) ||
      (out_gbuffer && !ov(out_gbuffer, RFX_FMT_RGBA32F, a.out_gb)) || (out_velocity && !ov(out_velocity, RFX_FMT_RGBA32F, a.out_vel)))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "gbuffer_ingest: albedo / material RGBA8|RGBA16F, normal / motion RGBA16F|RGBA32F, emissive RGBA16F, depth R32F, outputs RGBA32F");
  a.W = (int)depth->width; a.H = (int)depth->height;
  for (const rfx_plane* pl : {albedo, normal, material, emissive, motion, out_gbuffer, out_velocity})
    if (pl && ((int)pl->width != a.W || (int)pl->height != a.H)) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "gbuffer_ingest: plane sizes differ");
  a.normalize_normals = p->normalize_normals;
  a.motion_sx = p->motion_scale[0]; a.motion_sy = p->motion_scale[1];
  rows(row0, row1, depth->height, a.row0, a.row1);
  LAUNCHED(launch_gbuffer_ingest(a, pick(ctx, stream)));
  return RFX_OK;
}

}  // extern "C"

// ==========================================================================================
// native SSGI chain
// ==========================================================================================
struct IPlane {  // chain-internal plane (formats of k_chain.cu): raw pitched allocation
  void* p = nullptr;
  size_t pitch = 0;
};
struct rfx_ssgi_chain {
  rfx_ctx* ctx;
  rfx_ssgi_chain_options opt;
  rfx_plane ssgi_out{}, tr[2]{}, dnA[2]{}, dnB[2]{}, composed{};
  rfx_plane fb{};  // denoise_mode != full: the FramebufferTexture copy of the temporal target's attachment 0 (TemporalReprojectPass.js:134-152,197-200)
  // fast chain (fast_math on at creation, mode SSGI): interleaved internal planes; history planes are double-buffered by frame
  // parity so that, in a row-sharded group, no rank overwrites rows a peer may still be reading (see rfx_group_*)
  bool fastpath = false;
  IPlane nrdz, tr32, dnA16[2], dnB16[2];  // dnA16[1] is used by row-sharded groups only (the A target double-buffered like B, see chain_render_fast)
  rfx_plane composed2[2]{};  // fast chain: composed of frame parity 0 / 1 (composed2[cur] is output 0)
  struct rfx_group* group = nullptr;  // row-sharded group this chain is attached to (rfx_group_attach_chain)
  PeerPV peer_composed[2]{}, peer_dn[2]{}, peer_dnA[2]{};
  uint64_t frame_idx = 0;    // frames completed (advances with the frame's last launch)
  bool views_valid = false;  // tr[]/dnB[] hold the split views of the current frame's interleaved planes
  // host-buffer entry points: two staging sets so frame i+1 uploads while frame i renders; H2D, kernels and D2H each get
  // their own stream and are ordered by events only (see rfx_ssgi_chain_submit_host)
  rfx_plane in_depth[2]{}, in_gb[2]{}, in_vel[2]{}, in_direct[2]{};
  bool have_staging = false;
  cudaStream_t s_up = nullptr, s_dn = nullptr;
  cudaEvent_t ev_up[2]{}, ev_rendered[2]{}, ev_dn[2]{};
  uint64_t host_submitted = 0;
  int dn_buf[2] = {-1, -1};  // fast chain: which composed2[] buffer the D2H of staging set 0 / 1 reads
  // cross-frame state (TemporalReprojectPass.js:203-213)
  bool have_prev = false;
  float prev_view[16], prev_world[16], prev_proj[16], prev_proj_inv[16], prev_pos[3];
  float keep_data = 0.0f;  // SSGIEffect's constructor resets the denoiser (makeOptionsReactive -> reset())
  // blue-noise counters: one closure per material (BlueNoiseUtils.js:17-33)
  int32_t bn_trace = 0, bn_poisson = 0;
  // optional per-pass event timing
  bool profiling = false;
  struct Span { cudaEvent_t a, b; int slot; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
};

static cudaEvent_t chain_event(rfx_ssgi_chain* ch) {
  if (!ch->event_pool.empty()) { cudaEvent_t e = ch->event_pool.back(); ch->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct SpanGuard {  // records an event pair around one launch when profiling is on
  rfx_ssgi_chain* ch; cudaStream_t s; cudaEvent_t a = nullptr; int slot;
  SpanGuard(rfx_ssgi_chain* c, cudaStream_t st, int sl) : ch(c), s(st), slot(sl) { if (ch->profiling) { a = chain_event(ch); cudaEventRecord(a, s); } }
  ~SpanGuard() { if (a) { cudaEvent_t b = chain_event(ch); cudaEventRecord(b, s); ch->spans.push_back({a, b, slot}); } }
};

static int32_t next_blue(int32_t start, int32_t& counter) {  // BlueNoiseUtils.js:25-28
  const int64_t highest = 0x7fffffff;
  counter = (int32_t)(((int64_t)start + (int64_t)counter + 1) % highest);
  return counter;
}

extern "C" {

rfx_status rfx_ssgi_chain_create(rfx_ctx* ctx, const rfx_ssgi_chain_options* opt, rfx_ssgi_chain** out) {
  if (!ctx || !opt || !out || opt->width == 0 || opt->height == 0 || opt->denoise_iterations < 0) return fail(ctx, RFX_ERR_INVALID_ARG, "chain_create: bad arguments");
  rfx_ssgi_chain* ch = new rfx_ssgi_chain();
  ch->ctx = ctx;
  ch->opt = *opt;
  rfx_status st = RFX_OK;
  auto alloc = [&](int fmt, rfx_plane* p) { if (st == RFX_OK) st = rfx_plane_alloc(ctx, fmt, opt->width, opt->height, p); };
  auto ialloc = [&](size_t texel, IPlane* p) {
    if (st != RFX_OK) return;
    p->pitch = ((size_t)opt->width * texel + 255) & ~(size_t)255;
    if (p->pitch * opt->height >= (1ull << 32)) { st = fail(ctx, RFX_ERR_UNSUPPORTED, "chain_create: internal plane exceeds 4 GiB"); return; }
    if (cudaMalloc(&p->p, p->pitch * opt->height) != cudaSuccess || cudaMemset(p->p, 0, p->pitch * opt->height) != cudaSuccess) st = fail(ctx, RFX_ERR_CUDA, "chain_create: cudaMalloc failed");
  };
  CU(cudaSetDevice(ctx->device));
  if (opt->denoise_mode < RFX_DENOISE_FULL || opt->denoise_mode > RFX_DENOISE_TEMPORAL) {
    delete ch;
    return fail(ctx, RFX_ERR_UNSUPPORTED, "chain_create: denoise_mode must be full / full_temporal / temporal (\"denoised\" hands an array of textures to a sampler in the reference and cannot run there either)");
  }
  const bool scaled = opt->resolution_scale != 0.0f && opt->resolution_scale != 1.0f;
  if (scaled && !(opt->resolution_scale > 0.0f && opt->resolution_scale < 1.0f)) { delete ch; return fail(ctx, RFX_ERR_INVALID_ARG, "chain_create: resolution_scale must be in (0, 1]"); }
  const uint32_t sw = scaled ? (uint32_t)((double)opt->width * (double)opt->resolution_scale) : opt->width;   // renderTarget.setSize(width * scale, height * scale)
  const uint32_t sh = scaled ? (uint32_t)((double)opt->height * (double)opt->resolution_scale) : opt->height;
  if (sw == 0 || sh == 0) { delete ch; return fail(ctx, RFX_ERR_INVALID_ARG, "chain_create: resolution_scale leaves an empty SSGI target"); }
  // latched: the history formats differ between the paths; the fused fast chain needs the full-size SSGI target and the Poisson history
  ch->fastpath = ctx->fast_math && opt->mode == RFX_MODE_SSGI && opt->denoise_mode == RFX_DENOISE_FULL && !scaled;
  if (st == RFX_OK) st = rfx_plane_alloc(ctx, RFX_FMT_RGBA32F, sw, sh, &ch->ssgi_out);
  if (opt->denoise_mode != RFX_DENOISE_FULL) alloc(RFX_FMT_RGBA32F, &ch->fb);
  if (ch->fastpath) {
    ialloc(16, &ch->nrdz); ialloc(32, &ch->tr32); ialloc(16, &ch->dnA16[0]); ialloc(16, &ch->dnA16[1]); ialloc(16, &ch->dnB16[0]); ialloc(16, &ch->dnB16[1]);
    alloc(RFX_FMT_RGBA32F, &ch->composed2[0]); alloc(RFX_FMT_RGBA32F, &ch->composed2[1]);
  } else {
    for (int i = 0; i < 2; i++) { alloc(RFX_FMT_RGBA32F, &ch->tr[i]); alloc(RFX_FMT_RGBA16F, &ch->dnA[i]); alloc(RFX_FMT_RGBA16F, &ch->dnB[i]); }
    alloc(RFX_FMT_RGBA32F, &ch->composed);
  }
  if (st != RFX_OK) { rfx_ssgi_chain_destroy(ch); return st; }
  *out = ch;
  return RFX_OK;
}

void rfx_ssgi_chain_destroy(rfx_ssgi_chain* ch) {
  if (!ch) return;
  rfx_ctx* ctx = ch->ctx;
  cudaStreamSynchronize(ctx->stream);
  if (ch->s_up) cudaStreamSynchronize(ch->s_up);
  if (ch->s_dn) cudaStreamSynchronize(ch->s_dn);
  rfx_plane* all[] = {&ch->fb, &ch->ssgi_out, &ch->tr[0], &ch->tr[1], &ch->dnA[0], &ch->dnA[1], &ch->dnB[0], &ch->dnB[1], &ch->composed,
                      &ch->in_depth[0], &ch->in_gb[0], &ch->in_vel[0], &ch->in_direct[0], &ch->in_depth[1], &ch->in_gb[1], &ch->in_vel[1], &ch->in_direct[1]};
  for (rfx_plane* p : all) if (p->ptr) rfx_plane_free(ctx, p);
  for (rfx_plane* p : {&ch->composed2[0], &ch->composed2[1]}) if (p->ptr) rfx_plane_free(ctx, p);
  for (IPlane* p : {&ch->nrdz, &ch->tr32, &ch->dnA16[0], &ch->dnA16[1], &ch->dnB16[0], &ch->dnB16[1]}) if (p->p) cudaFree(p->p);
  for (int i = 0; i < 2; i++) for (cudaEvent_t e : {ch->ev_up[i], ch->ev_rendered[i], ch->ev_dn[i]}) if (e) cudaEventDestroy(e);
  if (ch->s_up) cudaStreamDestroy(ch->s_up);
  if (ch->s_dn) cudaStreamDestroy(ch->s_dn);
  for (auto& sp : ch->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (cudaEvent_t e : ch->event_pool) cudaEventDestroy(e);
  delete ch;
}

rfx_status rfx_ssgi_chain_set_profiling(rfx_ssgi_chain* ch, int32_t enable) {
  if (!ch) return RFX_ERR_INVALID_ARG;
  ch->profiling = enable != 0;
  return RFX_OK;
}
rfx_status rfx_ssgi_chain_get_profile(rfx_ssgi_chain* ch, double* ms, uint64_t* launches) {
  if (!ch || !ms || !launches) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = ch->ctx;
  CU(cudaDeviceSynchronize());
  for (auto& sp : ch->spans) {
    float t = 0.0f;
    CU(cudaEventElapsedTime(&t, sp.a, sp.b));
    ms[sp.slot] += (double)t;
    launches[sp.slot] += 1;
    ch->event_pool.push_back(sp.a);
    ch->event_pool.push_back(sp.b);
  }
  ch->spans.clear();
  return RFX_OK;
}

rfx_status rfx_ssgi_chain_set_options(rfx_ssgi_chain* ch, const rfx_ssgi_chain_options* opt) {
  if (!ch || !opt) return RFX_ERR_INVALID_ARG;
  if (opt->width != ch->opt.width || opt->height != ch->opt.height) return fail(ch->ctx, RFX_ERR_SIZE_MISMATCH, "chain_set_options: size change needs a new chain");
  if (opt->denoise_iterations < 0 || opt->steps < 1 || opt->refine_steps < 0) return fail(ch->ctx, RFX_ERR_INVALID_ARG, "chain_set_options: bad option value");
  if (opt->denoise_mode != ch->opt.denoise_mode || opt->mode != ch->opt.mode) return fail(ch->ctx, RFX_ERR_UNSUPPORTED, "chain_set_options: mode / denoise_mode are constructor options (Denoiser.js:17-64): create a new chain");
  if (opt->resolution_scale != ch->opt.resolution_scale) return fail(ch->ctx, RFX_ERR_SIZE_MISMATCH, "chain_set_options: resolution_scale resizes the SSGI target (SSGIEffect.js:193-196 calls setSize): create a new chain");
  const int32_t start = ch->opt.blue_noise_start;
  ch->opt = *opt;
  ch->opt.blue_noise_start = start;  // the blue-noise closures keep their start index for the life of the material
  ch->keep_data = 0.0f;              // every reference setter ends with this.reset()
  return RFX_OK;
}

rfx_status rfx_ssgi_chain_reset(rfx_ssgi_chain* ch) {
  if (!ch) return RFX_ERR_INVALID_ARG;
  ch->keep_data = 0.0f;  // TemporalReprojectPass.reset()  :158-160
  return RFX_OK;
}

rfx_status rfx_ssgi_chain_output(rfx_ssgi_chain* ch, int32_t which, rfx_plane* out) {
  if (!ch || !out) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = ch->ctx;
  if (which < 0 || which > 5) return fail(ctx, RFX_ERR_INVALID_ARG, "chain_output: which must be 0..5");
  if (ch->fastpath) {
    const int last = (int)((ch->frame_idx + 1) & 1);  // parity of the most recently completed frame
    if (which == 0) { *out = ch->composed2[last]; return RFX_OK; }
    if (which == 1) { *out = ch->ssgi_out; return RFX_OK; }
    // views of the interleaved planes in the reference's formats, refreshed on the context stream
    if (!ch->tr[0].ptr) {
      rfx_status st = RFX_OK;
      for (int i = 0; i < 2 && st == RFX_OK; i++) {
        st = rfx_plane_alloc(ctx, RFX_FMT_RGBA32F, ch->opt.width, ch->opt.height, &ch->tr[i]);
        if (st == RFX_OK) st = rfx_plane_alloc(ctx, RFX_FMT_RGBA16F, ch->opt.width, ch->opt.height, &ch->dnB[i]);
      }
      if (st != RFX_OK) return st;
    }
    if (!ch->views_valid) {
      const int W = (int)ch->opt.width, H = (int)ch->opt.height;
      LAUNCHED(launch_split_tr(PV{(const unsigned char*)ch->tr32.p, W, H, (long long)ch->tr32.pitch}, OutV{(unsigned char*)ch->tr[0].ptr, (long long)ch->tr[0].pitch},
                               OutV{(unsigned char*)ch->tr[1].ptr, (long long)ch->tr[1].pitch}, W, H, ctx->stream));
      LAUNCHED(launch_split_dn(PV{(const unsigned char*)ch->dnB16[last].p, W, H, (long long)ch->dnB16[last].pitch}, OutV{(unsigned char*)ch->dnB[0].ptr, (long long)ch->dnB[0].pitch},
                               OutV{(unsigned char*)ch->dnB[1].ptr, (long long)ch->dnB[1].pitch}, W, H, ctx->stream));
      ch->views_valid = true;
    }
    *out = which == 2 ? ch->tr[0] : which == 3 ? ch->tr[1] : which == 4 ? ch->dnB[0] : ch->dnB[1];
    return RFX_OK;
  }
  switch (which) {
    case 0: *out = ch->opt.denoise_mode == RFX_DENOISE_TEMPORAL ? ch->tr[0] : ch->composed; break;  // denoiser.texture (Denoiser.js:67-78)
    case 1: *out = ch->ssgi_out; break;
    case 2: *out = ch->tr[0]; break;
    case 3: *out = ch->tr[1]; break;
    case 4: *out = ch->dnB[0]; break;
    default: *out = ch->dnB[1]; break;
  }
  return RFX_OK;
}

// installs the row-segment table of launch k (all owned blocks) for the duration of one launch call
struct SegScope {
  rfx_ctx* ctx;
  RowSegs segs;
  SegScope(rfx_ctx* c, const uint32_t* ranges, uint32_t n_blocks, uint32_t n_launches, uint32_t k) : ctx(c) {
    if (!ranges || n_blocks <= 1) return;
    int r0[RFX_MAX_SEGS], r1[RFX_MAX_SEGS];
    for (uint32_t b = 0; b < n_blocks; b++) { r0[b] = (int)ranges[(b * n_launches + k) * 2]; r1[b] = (int)ranges[(b * n_launches + k) * 2 + 1]; }
    segs = make_segs(r0, r1, (int)n_blocks);
    ctx->segs_override = &segs;
  }
  ~SegScope() { ctx->segs_override = nullptr; }
};

// ------------------------------------------------------------------------------------------
// fast chain (k_chain.cu): same frame logic as chain_render_impl below, interleaved internal planes, compose fused into the
// last Poisson pass.  Launch indices k (for `ranges` and [k_begin, k_end)) are those of the reference chain — K1, K2, K3 pass
// 0..2n-1, K4 — the K4 range selects the rows the last pass composes.
// ------------------------------------------------------------------------------------------
static PV ipv(const IPlane& p, int w, int h) { return PV{(const unsigned char*)p.p, w, h, (long long)p.pitch}; }
static OutV iov(const IPlane& p) { return OutV{(unsigned char*)p.p, (long long)p.pitch}; }
static PV rpv(const rfx_plane& p) { return PV{(const unsigned char*)p.ptr, (int)p.width, (int)p.height, (long long)p.pitch}; }
static RowSegs segs_for(const uint32_t* ranges, uint32_t n_blocks, uint32_t n_launches, uint32_t k, int H) {
  int r0[RFX_MAX_SEGS], r1[RFX_MAX_SEGS];
  if (!ranges) { r0[0] = 0; r1[0] = H; return make_segs(r0, r1, 1); }
  for (uint32_t b = 0; b < n_blocks; b++) { r0[b] = (int)ranges[(b * n_launches + k) * 2]; r1[b] = (int)ranges[(b * n_launches + k) * 2 + 1]; }
  return make_segs(r0, r1, (int)n_blocks);
}
static void peer_single(PeerPV& pp, PV local) {
  pp = PeerPV{};
  pp.local = local;
  pp.n = 1;
  pp.own0 = 0; pp.own1 = local.h;
}

// 2-D TMA descriptor over a plane of 16-byte texels, addressed as rows of 4-byte elements (box dims are limited to 256 elements)
static bool encode_texel_map(CUtensorMap* map, const void* base, int W, int H, size_t pitch, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = (EncodeFn)p;
  }
  if (box_w * 4 > 256 || box_h > 256) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)W * 4, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_w * 4, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static rfx_status chain_render_fast(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f, const uint32_t* ranges, uint32_t n_blocks,
                                    uint32_t k_begin, uint32_t k_end, int k1_phase) {
  rfx_ctx* ctx = ch->ctx;
  const rfx_ssgi_chain_options& o = ch->opt;
  const int W = (int)o.width, H = (int)o.height;
  const uint32_t n_launches = 3u + 2u * (uint32_t)o.denoise_iterations;
  if (!ranges) n_blocks = 1;
  auto on = [&](uint32_t k) { return k >= k_begin && k < k_end; };
  const cudaStream_t cs = stream ? (cudaStream_t)stream : ctx->stream;
  const int cur = (int)(ch->frame_idx & 1), prev = cur ^ 1;
  rfx_status st = RFX_OK;
  PV depth, gb, vel;
  if (!pv(f->depth, RFX_FMT_R32F, depth) || !pv(f->gbuffer, RFX_FMT_RGBA32F, gb) || !pv(f->velocity, RFX_FMT_RGBA32F, vel))
    return fail(ctx, RFX_ERR_BAD_FORMAT, "chain: depth must be R32F, gbuffer / velocity RGBA32F");
  if (depth.w != W || depth.h != H || gb.w != W || gb.h != H || vel.w != W || vel.h != H) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "chain: input planes must match the chain size");
  CamD cam;
  cam_to_dev(f->cam, cam);
  uint32_t k = 0;
  ch->views_valid = false;
  // ---- K1
  if (on(k)) {
    rfx_ssgi_params sp{};
    sp.cam = f->cam;
    sp.ray_distance = o.distance; sp.thickness = o.thickness; sp.env_blur = o.env_blur;
    sp.max_env_map_mip_level = ctx->env_set ? (float)((int)std::floor(std::log2((double)std::max(ctx->env.size_x, ctx->env.size_y))) + 1) : 0.0f;
    sp.steps = o.steps; sp.refine_steps = o.refine_steps; sp.mode = o.mode; sp.flags = o.ssgi_flags;
    sp.blue_noise_index = k1_phase == 2 ? ch->bn_trace : next_blue(o.blue_noise_start, ch->bn_trace);
    SegScope seg_scope(ctx, ranges, n_blocks, n_launches, k);
    ctx->k1_phase = k1_phase;
    ctx->peer_accumulated = ch->group ? &ch->peer_composed[prev] : nullptr;
    {
      SpanGuard g(ch, cs, 0);
      st = rfx_ssgi_trace_launch(ctx, stream, &sp, f->depth, f->gbuffer, nullptr, f->direct_light, &ch->composed2[prev], &ch->ssgi_out, ranges ? ranges[k * 2] : 0u,
                                 ranges ? ranges[k * 2 + 1] : 0u);
    }
    ctx->k1_phase = 0;
    ctx->peer_accumulated = nullptr;
    if (st != RFX_OK) return st;
  }
  k++;
  // ---- K2
  if (on(k)) {
    CTemporalArgs a{};
    a.input = rpv(ch->ssgi_out); a.velocity = vel;
    if (ch->group) a.hist = ch->peer_dn[prev]; else peer_single(a.hist, ipv(ch->dnB16[prev], W, H));
    a.out = iov(ch->tr32);
    a.W = W; a.H = H;
    a.segs = segs_for(ranges, n_blocks, n_launches, k, H);
    a.cam = cam;
    if (!ch->have_prev) {
      memcpy(ch->prev_view, f->cam.view_matrix, 64); memcpy(ch->prev_world, f->cam.camera_matrix_world, 64);
      memcpy(ch->prev_proj, f->cam.projection, 64); memcpy(ch->prev_proj_inv, f->cam.projection_inverse, 64);
      memcpy(ch->prev_pos, f->camera_pos, 12);
      ch->have_prev = true;
    }
    memcpy(a.prev_world.m, ch->prev_world, 64); memcpy(a.prev_proj_inv.m, ch->prev_proj_inv, 64);
    matmul(ch->prev_proj, ch->prev_view, a.prev_proj_view.m);
    memcpy(a.camera_pos, f->camera_pos, 12);
    a.max_blend = 1.0f; a.clamp_intensity = 0.5f; a.keep_data = ch->keep_data; a.confidence_power = 0.75f;  // Denoiser.js:26-43, TemporalReprojectPass.js:17-32
    a.inv_w = (float)(1.0 / (double)W); a.inv_h = (float)(1.0 / (double)H);
    a.full_accumulate = f->camera_moved ? 0 : 1;
    {
      SpanGuard g(ch, cs, 1);
      LAUNCHED(launch_ctemporal(a, cs));
    }
    ch->keep_data = 1.0f;
    memcpy(ch->prev_world, f->cam.camera_matrix_world, 64); memcpy(ch->prev_view, f->cam.view_matrix, 64);
    memcpy(ch->prev_proj, f->cam.projection, 64); memcpy(ch->prev_proj_inv, f->cam.projection_inverse, 64);
    memcpy(ch->prev_pos, f->camera_pos, 12);
  }
  k++;
  // ---- K3 (+ fused K4)
  const int n_pass = 2 * o.denoise_iterations;
  const int halo = (int)std::ceil(o.radius * std::max(1.0f, (float)H / (float)W)) + 1;  // rows a Poisson tap can reach (the offset is rotated AFTER the division by the resolution)
  bool decoded = false;
  auto decode = [&](const RowSegs& segs) -> rfx_status {
    if (decoded) return RFX_OK;
    CDecodeArgs d{gb, depth, iov(ch->nrdz), W, H};
    LAUNCHED(launch_cdecode(d, segs, halo, cs));
    decoded = true;
    return RFX_OK;
  };
  const uint32_t k_compose = 2u + (uint32_t)n_pass;
  for (int i = 0; i < n_pass; i++, k++) {
    if (!on(k)) continue;
    const bool horizontal = (i % 2) == 0, last = i == n_pass - 1;
    CPoissonArgs a{};
    a.segs = segs_for(ranges, n_blocks, n_launches, k, H);
    if ((st = decode(a.segs)) != RFX_OK) return st;  // the first pass of a frame has the widest rows of all its passes
    a.nrdz = ipv(ch->nrdz, W, H);
    a.first = i == 0;
    // Target A (even passes).  A `discard`ed pixel keeps its texel = last frame's LAST even pass there (A2), and the LINEAR taps of the next
    // pass read such texels at silhouettes.  One GPU: A is single-buffered and a discard is simply no write.  In a row-sharded group a
    // rank's A rows outside its band hold the result of whichever even pass last covered them (the ranges shrink pass by pass), not the
    // last one's, so A is double-buffered by frame parity like B and the discarded texel is carried from the rank that OWNS the row.
    const int acur = (ch->group && !ctx->debug_no_a_carry) ? cur : 0;  // RFX_DEBUG_NO_A_CARRY: the pre-fix behaviour (single-buffered A, discard = no write)
    a.in = i == 0 ? ipv(ch->tr32, W, H) : ipv(horizontal ? ch->dnB16[cur] : ch->dnA16[acur], W, H);
    a.out = iov(horizontal ? ch->dnA16[acur] : ch->dnB16[cur]);
    if (!horizontal) { if (ch->group) a.carry = ch->peer_dn[prev]; else peer_single(a.carry, ipv(ch->dnB16[prev], W, H)); }
    else if (ch->group && !ctx->debug_no_a_carry) a.carry = ch->peer_dnA[prev];
    a.W = W; a.H = H;
    a.radius = o.radius; a.phi = o.phi; a.luma_phi = o.luma_phi; a.depth_phi = o.depth_phi; a.normal_phi = o.normal_phi;
    a.roughness_phi = o.roughness_phi; a.specular_phi = o.specular_phi;
    if ((st = blue_for(ctx, next_blue(o.blue_noise_start, ch->bn_poisson), a.blue)) != RFX_OK) return st;
    a.rot_table = ctx->rot_table;
    a.reach_x = (int)std::ceil(o.radius * std::max(1.0f, (float)W / (float)H)) + 2;
    a.reach_y = (int)std::ceil(o.radius * std::max(1.0f, (float)H / (float)W)) + 2;
    {
      const float SQ = 1.41421356237f;
      const float px[8] = {-1.0f, 0.0f, 1.0f, 0.0f, -0.25f * SQ, 0.25f * SQ, 0.25f * SQ, -0.25f * SQ};
      const float py[8] = {0.0f, -1.0f, 0.0f, 1.0f, -0.25f * SQ, -0.25f * SQ, 0.25f * SQ, 0.25f * SQ};
      for (int t = 0; t < 8; t++) { a.tap_ox[t] = px[t] / (float)W; a.tap_oy[t] = py[t] / (float)H; }
    }
    if (last && on(k_compose)) {
      a.compose = 1;
      a.compose_mode = ctx->compose_mode;
      a.csegs = segs_for(ranges, n_blocks, n_launches, k_compose, H);
      a.gb = gb;
      a.composed = OutV{(unsigned char*)ch->composed2[cur].ptr, (long long)ch->composed2[cur].pitch};
      if (ch->group) a.composed_carry = ch->peer_composed[prev]; else peer_single(a.composed_carry, rpv(ch->composed2[prev]));
      a.cam = cam;
    }
    SpanGuard g(ch, cs, i == 0 ? 2 : 3);
    if (ctx->k3_tma && i > 0) {  // experiment: TMA-staged tap tiles for the LINEAR passes
      CPoissonTmaArgs t{};
      t.a = a;
      t.box_w = (kTileW + 2 * a.reach_x) | 1;  // odd row pitch in texels: consecutive tile rows start 4 banks apart
      t.box_h = kTileH + 2 * a.reach_y;
      if (encode_texel_map(&t.map_in, a.in.p, W, H, (size_t)a.in.pitch, t.box_w, t.box_h) && encode_texel_map(&t.map_nrdz, a.nrdz.p, W, H, (size_t)a.nrdz.pitch, t.box_w, t.box_h)) {
        LAUNCHED(launch_cpoisson_tma(t, cs));
        continue;
      }
    }
    LAUNCHED(launch_cpoisson(a, cs));
  }
  // ---- K4 stand-alone (no Poisson pass to ride on, or the caller split the frame between the last pass and K4)
  k = k_compose;
  if (on(k) && (n_pass == 0 || !on(k - 1))) {
    CComposeArgs a{};
    a.segs = segs_for(ranges, n_blocks, n_launches, k, H);
    if ((st = decode(a.segs)) != RFX_OK) return st;
    a.nrdz = ipv(ch->nrdz, W, H); a.gb = gb; a.dn = ipv(ch->dnB16[cur], W, H);
    a.composed = OutV{(unsigned char*)ch->composed2[cur].ptr, (long long)ch->composed2[cur].pitch};
    if (ch->group) a.composed_carry = ch->peer_composed[prev]; else peer_single(a.composed_carry, rpv(ch->composed2[prev]));
    a.W = W; a.H = H; a.cam = cam;
    a.compose_mode = ctx->compose_mode;
    SpanGuard g(ch, cs, 4);
    LAUNCHED(launch_ccompose(a, cs));
  }
  if (on(k_compose)) ch->frame_idx++;  // the frame is complete: its planes become `prev`
  return RFX_OK;
}

// One frame of the chain.  `ranges` == nullptr: whole planes, one block.  Otherwise ranges[(blk*n_launches + k)*2 + {0,1}] =
// output rows [a,b) of launch k (chain order: K1, K2, K3 pass 0..2*iterations-1, K4) for row block `blk` of this rank
// (row-block sharding with locally recomputed halos, realism_effects_b200/parallel.py).  Only launches k in
// [k_begin, k_end) are issued, so a frame can be split into phases (K1 | the rest) between which the caller waits for a
// different all-gather; per-frame state advances with the launch that consumes it.
static rfx_status chain_render_impl(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f, const uint32_t* ranges, uint32_t n_blocks,
                                    uint32_t k_begin, uint32_t k_end, int k1_phase = 0) {
  if (ch->fastpath) return chain_render_fast(ch, stream, f, ranges, n_blocks, k_begin, k_end, k1_phase);
  rfx_ctx* ctx = ch->ctx;
  const rfx_ssgi_chain_options& o = ch->opt;
  rfx_status st = RFX_OK;
  const bool dm_full = o.denoise_mode == RFX_DENOISE_FULL;
  if (!dm_full && ranges) return fail(ctx, RFX_ERR_UNSUPPORTED, "chain: row-range rendering is implemented for denoise_mode full only");
  if (ranges && ch->ssgi_out.height != o.height) return fail(ctx, RFX_ERR_UNSUPPORTED, "chain: row-range rendering is implemented for resolution_scale 1 only");
  const uint32_t n_launches = 3u + 2u * (uint32_t)o.denoise_iterations;  // K1, K2, K3 passes, K4 (both modes: DenoiserComposePass runs for inputType specular too)
  if (!ranges) n_blocks = 1;
  // what SSGIPass samples as accumulatedTexture = denoiser.texture (Denoiser.js:67-78): the compose target, or the temporal pass's first texture
  rfx_plane* accumulated = o.denoise_mode == RFX_DENOISE_TEMPORAL ? &ch->tr[0] : &ch->composed;
  auto R0 = [&](uint32_t blk, uint32_t k) -> uint32_t { return ranges ? ranges[(blk * n_launches + k) * 2] : 0u; };
  auto R1 = [&](uint32_t blk, uint32_t k) -> uint32_t { return ranges ? ranges[(blk * n_launches + k) * 2 + 1] : 0u; };
  auto on = [&](uint32_t k) { return k >= k_begin && k < k_end; };
  const cudaStream_t cs = stream ? (cudaStream_t)stream : ctx->stream;
  uint32_t k = 0;
  // ---- K1  SSGIPass.render (src/ssgi/pass/SSGIPass.js:68-95)
  if (on(k)) {
    rfx_ssgi_params sp{};
    sp.cam = f->cam;
    sp.ray_distance = o.distance; sp.thickness = o.thickness; sp.env_blur = o.env_blur;
    sp.max_env_map_mip_level = ctx->env_set ? (float)((int)std::floor(std::log2((double)std::max(ctx->env.size_x, ctx->env.size_y))) + 1) : 0.0f;  // Utils.js:30-34
    sp.steps = o.steps; sp.refine_steps = o.refine_steps; sp.mode = o.mode; sp.flags = o.ssgi_flags;
    sp.blue_noise_index = k1_phase == 2 ? ch->bn_trace : next_blue(o.blue_noise_start, ch->bn_trace);  // the shading phase reuses the march phase's index
    // velocityTexture is a null sampler in the shipped wiring
    for (uint32_t blk = 0; blk < 1; blk++) {  // ONE launch covers every owned row block (SegScope installs the segment table)
      SegScope seg_scope(ctx, ranges, n_blocks, n_launches, k);
      ctx->viewz_reuse = blk > 0;  // the view-z plane depends on the depth plane only: one prepass per frame
      ctx->k1_phase = k1_phase;
      SpanGuard g(ch, cs, 0);
      st = rfx_ssgi_trace_launch(ctx, stream, &sp, f->depth, f->gbuffer, nullptr, f->direct_light, accumulated, &ch->ssgi_out, R0(blk, k), R1(blk, k));
    }
    ctx->viewz_reuse = false;
    ctx->k1_phase = 0;
    if (st != RFX_OK) return st;
  }
  k++;
  // ---- K2  TemporalReprojectPass.render (TemporalReprojectPass.js:162-214), options from Denoiser.js:26-43 + SSGIEffect.js:74-77
  const int tc = o.mode == RFX_MODE_SSGI ? 2 : 1;
  if (on(k)) {
    rfx_temporal_params tp{};
    tp.cam = f->cam;
    if (!ch->have_prev) {  // the constructor clones the current camera matrices (TemporalReprojectPass.js:94-97)
      memcpy(ch->prev_view, f->cam.view_matrix, 64); memcpy(ch->prev_world, f->cam.camera_matrix_world, 64);
      memcpy(ch->prev_proj, f->cam.projection, 64); memcpy(ch->prev_proj_inv, f->cam.projection_inverse, 64);
      memcpy(ch->prev_pos, f->camera_pos, 12);
      ch->have_prev = true;
    }
    memcpy(tp.prev_view_matrix, ch->prev_view, 64); memcpy(tp.prev_camera_matrix_world, ch->prev_world, 64);
    memcpy(tp.prev_projection, ch->prev_proj, 64); memcpy(tp.prev_projection_inverse, ch->prev_proj_inv, 64);
    memcpy(tp.camera_pos, f->camera_pos, 12); memcpy(tp.prev_camera_pos, ch->prev_pos, 12);
    tp.max_blend = 1.0f; tp.neighborhood_clamp_intensity = 0.5f; tp.keep_data = ch->keep_data; tp.confidence_power = 0.75f;
    tp.full_accumulate = f->camera_moved ? 0 : 1;  // options.fullAccumulate && !didCameraMove
    tp.log_transform = 1; tp.history_linear = 1;
    if (o.mode == RFX_MODE_SSGI) { tp.texture_count = 2; tp.input_type = RFX_INPUT_DIFFUSE_SPECULAR; tp.reproject_specular[0] = 0; tp.reproject_specular[1] = 1; }
    else { tp.texture_count = 1; tp.input_type = RFX_INPUT_SPECULAR; tp.reproject_specular[0] = 1; tp.reproject_specular[1] = 1; }
    for (uint32_t blk = 0; blk < 1; blk++) {  // ONE launch covers every owned row block (SegScope installs the segment table)
      SegScope seg_scope(ctx, ranges, n_blocks, n_launches, k);
      SpanGuard g(ch, cs, 1);
      // without a denoise pass overrideAccumulatedTextures stays empty: BOTH accumulated textures are the one FramebufferTexture
      rfx_plane* h0 = dm_full ? &ch->dnB[0] : &ch->fb;
      rfx_plane* h1 = dm_full ? &ch->dnB[1] : &ch->fb;
      st = rfx_temporal_reproject_launch(ctx, stream, &tp, &ch->ssgi_out, f->velocity, h0, tc == 2 ? h1 : nullptr, &ch->tr[0],
                                         tc == 2 ? &ch->tr[1] : nullptr, R0(blk, k), R1(blk, k));
    }
    if (st != RFX_OK) return st;
    if (!dm_full)  // renderer.copyFramebufferToTexture(tmpVec2, this.framebufferTexture) after the draw (:197-200)
      CU(cudaMemcpy2DAsync(ch->fb.ptr, ch->fb.pitch, ch->tr[0].ptr, ch->tr[0].pitch, (size_t)ch->tr[0].width * 16, ch->tr[0].height, cudaMemcpyDeviceToDevice, cs));
    ch->keep_data = 1.0f;  // :195
    memcpy(ch->prev_world, f->cam.camera_matrix_world, 64); memcpy(ch->prev_view, f->cam.view_matrix, 64);
    memcpy(ch->prev_proj, f->cam.projection, 64); memcpy(ch->prev_proj_inv, f->cam.projection_inverse, 64);
    memcpy(ch->prev_pos, f->camera_pos, 12);
  }
  k++;
  // ---- K3  PoissonDenoisePass.render (PoissonDenoisePass.js:135-149)
  rfx_poisson_params pp{};
  pp.radius = o.radius; pp.phi = o.phi; pp.luma_phi = o.luma_phi; pp.depth_phi = o.depth_phi; pp.normal_phi = o.normal_phi;
  pp.roughness_phi = o.roughness_phi; pp.specular_phi = o.specular_phi;
  pp.texture_count = tc; pp.gbuffer_texture = 1;
  if (o.mode == RFX_MODE_SSGI) { pp.is_texture_specular[0] = 0; pp.is_texture_specular[1] = 1; } else { pp.is_texture_specular[0] = 1; pp.is_texture_specular[1] = 1; }
  bool decoded = false;
  for (int i = 0; i < 2 * o.denoise_iterations; i++, k++) {
    if (!on(k) || !dm_full) continue;  // "full_temporal" / "temporal": no denoise pass (Denoiser.js:47-52)
    const bool horizontal = (i % 2) == 0;
    rfx_plane* inp = i == 0 ? ch->tr : (horizontal ? ch->dnB : ch->dnA);
    rfx_plane* outp = horizontal ? ch->dnA : ch->dnB;
    pp.input_linear = i == 0 ? 0 : 1;
    pp.blue_noise_index = next_blue(o.blue_noise_start, ch->bn_poisson);
    for (uint32_t blk = 0; blk < 1; blk++) {  // ONE launch covers every owned row block (SegScope installs the segment table)
      SegScope seg_scope(ctx, ranges, n_blocks, n_launches, k);
      ctx->nrd_reuse = decoded;  // the G-buffer does not change within a frame: decode it once, reuse it afterwards
      SpanGuard g(ch, cs, i == 0 ? 2 : 3);
      st = rfx_poisson_denoise_launch(ctx, stream, &pp, f->depth, f->gbuffer, &inp[0], tc == 2 ? &inp[1] : nullptr, &outp[0], tc == 2 ? &outp[1] : nullptr,
                                      R0(blk, k), R1(blk, k));
      decoded = true;
    }
    ctx->nrd_reuse = false;
    if (st != RFX_OK) return st;
  }
  // ---- K4  DenoiserComposePass.render ("full" and "full_temporal": Denoiser.js:55-64)
  rfx_plane* gi = dm_full ? ch->dnB : ch->tr;  // composerInputTextures = denoisePass?.texture ?? the temporal textures
  if (on(k) && o.denoise_mode != RFX_DENOISE_TEMPORAL) {
    rfx_compose_params cp{};
    cp.cam = f->cam;
    cp.input_type = o.mode == RFX_MODE_SSGI ? RFX_INPUT_DIFFUSE_SPECULAR : RFX_INPUT_SPECULAR;  // SSGIEffect.js:70-77
    for (uint32_t blk = 0; blk < 1; blk++) {  // ONE launch covers every owned row block (SegScope installs the segment table)
      SegScope seg_scope(ctx, ranges, n_blocks, n_launches, k);
      SpanGuard g(ch, cs, 4);
      if (o.mode == RFX_MODE_SSGI) st = rfx_gi_compose_launch(ctx, stream, &cp, f->depth, f->gbuffer, &gi[0], &gi[1], nullptr, &ch->composed, R0(blk, k), R1(blk, k));
      else st = rfx_gi_compose_launch(ctx, stream, &cp, f->depth, f->gbuffer, nullptr, &gi[0], f->direct_light, &ch->composed, R0(blk, k), R1(blk, k));  // scene = the composer input buffer (Denoiser.js:100-102)
    }
    if (st != RFX_OK) return st;
  }
  return RFX_OK;
}

rfx_status rfx_ssgi_chain_render(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f) {
  if (!ch || !f) return RFX_ERR_INVALID_ARG;
  return chain_render_impl(ch, stream, f, nullptr, 1, 0, 0xffffffffu);
}
static rfx_status check_ranges(rfx_ssgi_chain* ch, const uint32_t* ranges, uint32_t n_launches, uint32_t n_blocks) {
  const uint32_t expect = 3u + 2u * (uint32_t)ch->opt.denoise_iterations;
  if (n_launches != expect || n_blocks == 0 || n_blocks > RFX_MAX_SEGS) return fail(ch->ctx, RFX_ERR_INVALID_ARG, "chain ranges: expected %u launches per block, got %u (blocks %u)", expect, n_launches, n_blocks);
  for (uint32_t i = 0; i < n_launches * n_blocks; i++)
    if (ranges[2 * i] >= ranges[2 * i + 1] || ranges[2 * i + 1] > ch->opt.height) return fail(ch->ctx, RFX_ERR_INVALID_ARG, "chain ranges: bad range %u", i);
  return RFX_OK;
}
rfx_status rfx_ssgi_chain_render_ranges(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f, const uint32_t* ranges, uint32_t n_launches) {
  if (!ch || !f || !ranges) return RFX_ERR_INVALID_ARG;
  rfx_status st = check_ranges(ch, ranges, n_launches, 1);
  return st != RFX_OK ? st : chain_render_impl(ch, stream, f, ranges, 1, 0, 0xffffffffu);
}
rfx_status rfx_ssgi_chain_render_blocks(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f, const uint32_t* ranges, uint32_t n_launches,
                                        uint32_t n_blocks, uint32_t k_begin, uint32_t k_end) {
  if (!ch || !f || !ranges) return RFX_ERR_INVALID_ARG;
  rfx_status st = check_ranges(ch, ranges, n_launches, n_blocks);
  return st != RFX_OK ? st : chain_render_impl(ch, stream, f, ranges, n_blocks, k_begin, k_end);
}

// A frame in three parts, so a row-sharded caller can put its waits for the exchanged planes exactly where the data is needed:
// part 0 = K1 ray march (reads depth / G-buffer only), part 1 = K1 shading (samples last frame's `composed`), part 2 = K2..K4
// (K2 samples last frame's dnB history).  Parts 0+1 together write the same bytes as the fused K1.  ranges may be NULL.
rfx_status rfx_ssgi_chain_render_part(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f, const uint32_t* ranges, uint32_t n_launches,
                                      uint32_t n_blocks, uint32_t part) {
  if (!ch || !f || part > 2) return RFX_ERR_INVALID_ARG;
  if (ranges) {
    rfx_status st = check_ranges(ch, ranges, n_launches, n_blocks);
    if (st != RFX_OK) return st;
  }
  if (part == 2) return chain_render_impl(ch, stream, f, ranges, ranges ? n_blocks : 1, 1, 0xffffffffu);
  if (ch->fastpath) {  // the fast K1 is one fused kernel (its diffuse rays are compacted inside the block): part 0 is empty, part 1 is K1
    if (part == 0) return RFX_OK;
    return chain_render_impl(ch, stream, f, ranges, ranges ? n_blocks : 1, 0, 1, 0);
  }
  return chain_render_impl(ch, stream, f, ranges, ranges ? n_blocks : 1, 0, 1, part == 0 ? 1 : 2);
}

// Host-buffer path.  submit enqueues one frame and returns: the four input planes go H2D on a copy stream into staging set
// (frame & 1), the chain runs on the context stream once that upload's event fires, and `composed` goes D2H on a third stream once
// the frame's last kernel is done.  Frame i+1 therefore uploads (and frame i-1 downloads) while frame i renders; PCIe is full
// duplex, so steady-state time per frame is max(H2D, kernels, D2H) instead of their sum.  Hazards, all resolved on the device:
//   staging set reuse   - the upload of frame i+2 waits for frame i's kernels (ev_rendered);
//   `composed` reuse    - K4 of frame i+1 (the only writer) waits for frame i's D2H (ev_dn); K1 of frame i+1 only reads it.
rfx_status rfx_ssgi_chain_submit_host(rfx_ssgi_chain* ch, const rfx_ssgi_host_frame* hf) {
  if (!ch || !hf || !hf->depth || !hf->gbuffer || !hf->velocity || !hf->out_composed) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = ch->ctx;
  rfx_status st = RFX_OK;
  if (!ch->have_staging) {
    auto alloc = [&](int fmt, rfx_plane* p) { if (st == RFX_OK) st = rfx_plane_alloc(ctx, fmt, ch->opt.width, ch->opt.height, p); };
    for (int i = 0; i < 2; i++) {
      alloc(RFX_FMT_R32F, &ch->in_depth[i]); alloc(RFX_FMT_RGBA32F, &ch->in_gb[i]); alloc(RFX_FMT_RGBA32F, &ch->in_vel[i]); alloc(RFX_FMT_RGBA16F, &ch->in_direct[i]);
    }
    if (st != RFX_OK) return st;
    CU(cudaStreamCreateWithFlags(&ch->s_up, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ch->s_dn, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CU(cudaEventCreateWithFlags(&ch->ev_up[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ch->ev_rendered[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ch->ev_dn[i], cudaEventDisableTiming));
    }
    ch->have_staging = true;
  }
  const int set = (int)(ch->host_submitted & 1);
  if (ch->host_submitted >= 2) CU(cudaStreamWaitEvent(ch->s_up, ch->ev_rendered[set], 0));
  if ((st = rfx_plane_upload(ctx, ch->s_up, &ch->in_depth[set], hf->depth, 0)) != RFX_OK) return st;
  if ((st = rfx_plane_upload(ctx, ch->s_up, &ch->in_gb[set], hf->gbuffer, 0)) != RFX_OK) return st;
  if (hf->direct_light && (st = rfx_plane_upload(ctx, ch->s_up, &ch->in_direct[set], hf->direct_light, 0)) != RFX_OK) return st;
  if ((st = rfx_plane_upload(ctx, ch->s_up, &ch->in_vel[set], hf->velocity, 0)) != RFX_OK) return st;
  CU(cudaEventRecord(ch->ev_up[set], ch->s_up));
  CU(cudaStreamWaitEvent(ctx->stream, ch->ev_up[set], 0));
  rfx_ssgi_frame f{};
  f.cam = hf->cam;
  f.depth = &ch->in_depth[set]; f.gbuffer = &ch->in_gb[set]; f.velocity = &ch->in_vel[set]; f.direct_light = hf->direct_light ? &ch->in_direct[set] : nullptr;
  memcpy(f.camera_pos, hf->camera_pos, 12);
  f.camera_moved = hf->camera_moved;
  const rfx_plane* result = &ch->composed;
  if (ch->fastpath) {
    // `composed` is double-buffered by frame parity: this frame writes composed2[cur], whose previous reader is the D2H of two
    // frames ago (long finished in steady state); frame i-1's D2H keeps running under this frame's kernels
    const int cur = (int)(ch->frame_idx & 1);
    for (int q = 0; q < 2; q++)
      if (ch->dn_buf[q] == cur) CU(cudaStreamWaitEvent(ctx->stream, ch->ev_dn[q], 0));
    if ((st = chain_render_impl(ch, nullptr, &f, nullptr, 1, 0, 0xffffffffu)) != RFX_OK) return st;
    result = &ch->composed2[cur];
    ch->dn_buf[set] = cur;
  } else {
    const uint32_t n_launches = 3u + 2u * (uint32_t)ch->opt.denoise_iterations;
    const uint32_t split = n_launches - 1;  // K4 is the only launch that writes `composed`
    if (split && (st = chain_render_impl(ch, nullptr, &f, nullptr, 1, 0, split)) != RFX_OK) return st;
    if (ch->host_submitted >= 1) CU(cudaStreamWaitEvent(ctx->stream, ch->ev_dn[set ^ 1], 0));
    if ((st = chain_render_impl(ch, nullptr, &f, nullptr, 1, split, 0xffffffffu)) != RFX_OK) return st;
  }
  CU(cudaEventRecord(ch->ev_rendered[set], ctx->stream));
  CU(cudaStreamWaitEvent(ch->s_dn, ch->ev_rendered[set], 0));
  if ((st = rfx_plane_download(ctx, ch->s_dn, result, hf->out_composed, 0)) != RFX_OK) return st;
  CU(cudaEventRecord(ch->ev_dn[set], ch->s_dn));
  ch->host_submitted++;
  return RFX_OK;
}

// Blocks until at most `max_in_flight` (0 or 1) of the submitted frames are still incomplete.  A complete frame's out_composed is
// filled and its input buffers may be overwritten.  Frames complete in submission order.
rfx_status rfx_ssgi_chain_wait_host(rfx_ssgi_chain* ch, int32_t max_in_flight) {
  if (!ch || max_in_flight < 0) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = ch->ctx;
  if (ch->host_submitted == 0 || !ch->have_staging) return RFX_OK;
  if (max_in_flight == 0) { CU(cudaEventSynchronize(ch->ev_dn[(ch->host_submitted - 1) & 1])); }
  else if (max_in_flight == 1 && ch->host_submitted >= 2) { CU(cudaEventSynchronize(ch->ev_dn[(ch->host_submitted - 2) & 1])); }
  return RFX_OK;
}

// synchronous form: one frame in, its result out
rfx_status rfx_ssgi_chain_render_host(rfx_ssgi_chain* ch, const rfx_ssgi_host_frame* hf) {
  rfx_status st = rfx_ssgi_chain_submit_host(ch, hf);
  return st != RFX_OK ? st : rfx_ssgi_chain_wait_host(ch, 0);
}

}  // extern "C"

#include "rfx_group.inl"
