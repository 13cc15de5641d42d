// k_ingest.cu — G-buffer ingest (include/rfx.h rfx_gbuffer_ingest_launch): a conventional deferred
// renderer's SoA planes -> the reference's packed `gBuffer` (packGBuffer, src/gbuffer/shader/gbuffer_packing.glsl:166-178) and
// `velocity` (src/temporal-reproject/material/VelocityDepthNormalMaterial.js:76-83,186-188) planes, so a non-three.js host does
// not need the reference's rasteriser materials.
//
// A pure stream: per pixel 4-8 (albedo) + 8-16 (normal) + 4-8 (material) + 8 (emissive) + 8-16 (motion) + 4 (depth) bytes in,
// 16 + 16 bytes out (68 B with the 8-bit / fp16 input formats) and ~120 instructions — HBM-bound.  One thread per pixel, a warp
// covers 32 consecutive pixels of a row, so every load / store instruction of a warp is one contiguous 128 B .. 512 B run.
// All arithmetic is the oracle's (IEEE, -fmad=false): the packed planes are BIT-identical to the CPU parity restatement of the same packers.
#include "rfx_device.cuh"
#include "rfx_kernels.h"

namespace rfx {

namespace {

constexpr float kOneSafe = 0.999999f;

RFX_D unsigned to_byte(float v) {  // uvec4(v * 255.0) of min(v + NON_ZERO_OFFSET, ONE_SAFE): truncation, negative / NaN -> 0
  const float t = fminf(v + RFX_NON_ZERO_OFFSET, kOneSafe) * 255.0f;
  return t > 0.0f ? (unsigned)t : 0u;
}
// vec4ToFloat :143-149
RFX_D float vec4ToFloat(v4 v) { return __uint_as_float((to_byte(v.w) << 24) | (to_byte(v.z) << 16) | (to_byte(v.y) << 8) | to_byte(v.x)); }
// color2float :17-22
RFX_D float color2float(float r, float g, float b) {
  const float cp = 256.0f, cp1 = 257.0f;
  r = fminf(r + RFX_NON_ZERO_OFFSET, kOneSafe); g = fminf(g + RFX_NON_ZERO_OFFSET, kOneSafe); b = fminf(b + RFX_NON_ZERO_OFFSET, kOneSafe);
  return floorf(r * cp + 0.5f) + floorf(b * cp + 0.5f) * cp1 + floorf(g * cp + 0.5f) * cp1 * cp1;
}
// encodeOctWrap + packNormal :36-50,61
RFX_D float packNormal(v3 n) {
  n = n / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
  float ox = n.x, oy = n.y;
  if (!(n.z > 0.0f)) {  // OctWrap
    const float wx = 1.0f - fabsf(n.y), wy = 1.0f - fabsf(n.x);
    ox = n.x < 0.0f ? -wx : wx;
    oy = n.y < 0.0f ? -wy : wy;
  }
  ox = ox * 0.5f + 0.5f; oy = oy * 0.5f + 0.5f;
  return __uint_as_float(packHalf2x16(ox, oy));
}
// encodeRGBE8 :127-134 (maxComponent > 0).  log2 is the correctly rounded one of the exact variants (double, rounded once);
// exp2 of the integral exponent is exact through ldexp.
RFX_D v4 encodeRGBE8(v3 rgb, float mx) {
  const float fExp = ceilf((float)log2((double)mx));
  const float s = ldexpf(1.0f, (int)fExp);
  return mk4(__fdiv_rn(rgb.x, s), __fdiv_rn(rgb.y, s), __fdiv_rn(rgb.z, s), __fdiv_rn(fExp + 128.0f, 255.0f));
}

template <bool H>
RFX_D v4 ld_color(const PV& t, int x, int y) {  // RGBA16F or RGBA8 (unorm: c / 255, as GL decodes it)
  if (H) return ld_h4(t, x, y);
  const uchar4 c = __ldg((const uchar4*)(t.p + pv_off(t, x, y, 4)));
  return mk4(__fdiv_rn((float)c.x, 255.0f), __fdiv_rn((float)c.y, 255.0f), __fdiv_rn((float)c.z, 255.0f), __fdiv_rn((float)c.w, 255.0f));
}
template <bool F32>
RFX_D v4 ld_vec(const PV& t, int x, int y) { return F32 ? f4v(ld_f4(t, x, y)) : ld_h4(t, x, y); }

template <bool ALB_H, bool MAT_H, bool NRM_F32, bool MOT_F32>
__global__ void __launch_bounds__(256) gbuffer_ingest_kernel(const __grid_constant__ IngestArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = a.row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.W || y >= a.row1) return;
  const float d = ld_r32f(a.depth, x, y);
  if (d == 1.0f) {  // the cleared targets' texel
    if (a.out_gb.p) st_f4(a.out_gb.p, a.out_gb.pitch, x, y, make_float4(0.0f, 0.0f, 0.0f, 1.0f));
    if (a.out_vel.p) st_f4(a.out_vel.p, a.out_vel.pitch, x, y, make_float4(0.0f, 0.0f, 0.0f, 1.0f));
    return;
  }
  v3 n = xyz(ld_vec<NRM_F32>(a.normal, x, y));
  if (a.normalize_normals) n = normalize(n);
  const float pn = packNormal(n);
  if (a.out_gb.p) {
    const v4 diffuse = ld_color<ALB_H>(a.albedo, x, y), mt = ld_color<MAT_H>(a.material, x, y);
    float ga = 0.0f;
    if (a.emissive.p) {
      const v3 em = xyz(ld_h4(a.emissive, x, y));
      const float mx = fmaxf(fmaxf(em.x, em.y), em.z);
      if (mx > 0.0f) ga = vec4ToFloat(encodeRGBE8(em, mx));
    }
    st_f4(a.out_gb.p, a.out_gb.pitch, x, y, make_float4(vec4ToFloat(diffuse), pn, color2float(mt.x, mt.y, 0.0f), ga));
  }
  if (a.out_vel.p) {
    float mx = 0.0f, my = 0.0f;
    if (a.motion.p) { const v4 m = ld_vec<MOT_F32>(a.motion, x, y); mx = m.x * a.motion_sx; my = m.y * a.motion_sy; }
    st_f4(a.out_vel.p, a.out_vel.pitch, x, y, make_float4(mx, my, pn, d));
  }
}

template <bool A, bool B, bool C>
cudaError_t launch3(const IngestArgs& a, dim3 grid, cudaStream_t s) {
  if (a.motion_f32) gbuffer_ingest_kernel<A, B, C, true><<<grid, 256, 0, s>>>(a);
  else gbuffer_ingest_kernel<A, B, C, false><<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}
template <bool A, bool B>
cudaError_t launch2(const IngestArgs& a, dim3 grid, cudaStream_t s) { return a.normal_f32 ? launch3<A, B, true>(a, grid, s) : launch3<A, B, false>(a, grid, s); }
template <bool A>
cudaError_t launch1(const IngestArgs& a, dim3 grid, cudaStream_t s) { return a.material_half ? launch2<A, true>(a, grid, s) : launch2<A, false>(a, grid, s); }

}  // namespace

cudaError_t launch_gbuffer_ingest(const IngestArgs& a, cudaStream_t s) {
  dim3 grid((a.W + 31) / 32, (a.row1 - a.row0 + 7) / 8);
  return a.albedo_half ? launch1<true>(a, grid, s) : launch1<false>(a, grid, s);
}

}  // namespace rfx
