// oracle/glsl.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// A tiny scalar-fp32 emulation of the GLSL ES 3.00 vocabulary the reference shaders use
// (vec2/3/4, mat4, built-ins, packHalf2x16, GL sampler state), so that oracle/rfx_oracle.cpp
// can restate the reference fragment shaders almost line by line.
//
// Decisions where GL leaves behaviour implementation-defined (SURVEY.md §7 "Hard parts"):
//  * packHalf2x16 rounds to nearest-even; half denormals are kept.
//  * bilinear weights are full fp32 (software-GL behaviour, not 8-bit fixed point).
//  * texelFetch outside the image clamps to the edge texel.
//  * derivatives are "fine" 2x2-quad differences, evaluated for all four quad pixels even
//    when a neighbour executed `discard`.
//  * a null sampler returns (0,0,0,1).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace gl {

// min / max: NaN-ignoring IEEE fmin/fmax for floats (what GPU min/max instructions and CUDA's fminf/fmaxf do)
inline float gmin(float a, float b) { return std::fmin(a, b); }
inline float gmax(float a, float b) { return std::fmax(a, b); }
inline int gmin(int a, int b) { return a < b ? a : b; }
inline int gmax(int a, int b) { return a > b ? a : b; }

// ---------------------------------------------------------------- vectors
struct vec2 {
  float x = 0, y = 0;
  vec2() {}
  explicit vec2(float s) : x(s), y(s) {}
  vec2(float x_, float y_) : x(x_), y(y_) {}
};
struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  vec3(vec2 a, float z_) : x(a.x), y(a.y), z(z_) {}
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
  float x = 0, y = 0, z = 0, w = 0;
  vec4() {}
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
  vec4(vec3 a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) {}
  vec3 xyz() const { return vec3(x, y, z); }
  vec2 xy() const { return vec2(x, y); }
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
};

#define GL_BINOP(V, op)                                                     \
  inline V operator op(V a, V b);                                           \
  inline V operator op(V a, float s);                                       \
  inline V operator op(float s, V a);
#define GL_DEF2(op)                                                                        \
  inline vec2 operator op(vec2 a, vec2 b) { return vec2(a.x op b.x, a.y op b.y); }         \
  inline vec2 operator op(vec2 a, float s) { return vec2(a.x op s, a.y op s); }            \
  inline vec2 operator op(float s, vec2 a) { return vec2(s op a.x, s op a.y); }
#define GL_DEF3(op)                                                                             \
  inline vec3 operator op(vec3 a, vec3 b) { return vec3(a.x op b.x, a.y op b.y, a.z op b.z); }  \
  inline vec3 operator op(vec3 a, float s) { return vec3(a.x op s, a.y op s, a.z op s); }       \
  inline vec3 operator op(float s, vec3 a) { return vec3(s op a.x, s op a.y, s op a.z); }
#define GL_DEF4(op)                                                                                           \
  inline vec4 operator op(vec4 a, vec4 b) { return vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }    \
  inline vec4 operator op(vec4 a, float s) { return vec4(a.x op s, a.y op s, a.z op s, a.w op s); }           \
  inline vec4 operator op(float s, vec4 a) { return vec4(s op a.x, s op a.y, s op a.z, s op a.w); }
GL_DEF2(+) GL_DEF2(-) GL_DEF2(*)
GL_DEF3(+) GL_DEF3(-) GL_DEF3(*)
GL_DEF4(+) GL_DEF4(-) GL_DEF4(*)
// division: vector / vector and scalar / vector are component-wise IEEE divisions; vector / SCALAR is lowered the
// way GLSL compilers lower it — one IEEE reciprocal, then a multiply per component.
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec4 operator/(vec4 a, vec4 b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline vec2 operator/(float s, vec2 a) { return vec2(s / a.x, s / a.y); }
inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec2 operator/(vec2 a, float s) { float r = 1.0f / s; return vec2(a.x * r, a.y * r); }
inline vec3 operator/(vec3 a, float s) { float r = 1.0f / s; return vec3(a.x * r, a.y * r, a.z * r); }
inline vec4 operator/(vec4 a, float s) { float r = 1.0f / s; return vec4(a.x * r, a.y * r, a.z * r, a.w * r); }
inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline vec4& operator+=(vec4& a, vec4 b) { a = a + b; return a; }
inline vec4& operator*=(vec4& a, float s) { a = a * s; return a; }

// ---------------------------------------------------------------- built-ins
// fp32 lowering rules (mirrored 1:1 by realism_effects_b200/csrc/rfx_device.cuh so that the
// CUDA kernels, compiled with -fmad=false, produce the same roundings):
//   dot      = fma chain, last component outermost
//   mix      = fma(y, a, x*(1-a))
//   M*v, v*M = fma chain per component, .w term innermost
//   cross    = fma(a.y, b.z, -(b.y*a.z)) ...
//   normalize(a) = a * (1/sqrt(dot(a,a)))
//   everything else: one IEEE operation per GLSL operator, no contraction (-ffp-contract=off)
// Transcendentals are correctly rounded (evaluated in double, rounded once to fp32).
inline float fma_(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float sincr(float x) { return (float)std::sin((double)x); }
inline float coscr(float x) { return (float)std::cos((double)x); }
inline float expcr(float x) { return (float)std::exp((double)x); }
inline float logcr(float x) { return (float)std::log((double)x); }
inline float log2cr(float x) { return (float)std::log2((double)x); }
inline float exp2cr(float x) { return (float)std::exp2((double)x); }
inline float powcr(float x, float y) { return (float)std::pow((double)x, (double)y); }
inline float atan2cr(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float acoscr(float x) { return (float)std::acos((double)x); }

inline float dot(vec2 a, vec2 b) { return fma_(a.y, b.y, a.x * b.x); }
inline float dot(vec3 a, vec3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
inline float dot(vec4 a, vec4 b) { return fma_(a.w, b.w, fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x))); }
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 cross(vec3 a, vec3 b) {
  return vec3(fma_(a.y, b.z, -(b.y * a.z)), fma_(a.z, b.x, -(b.z * a.x)), fma_(a.x, b.y, -(b.x * a.y)));
}
inline vec3 reflect(vec3 I, vec3 N) {
  float k = 2.0f * dot(N, I);
  return vec3(fma_(-k, N.x, I.x), fma_(-k, N.y, I.y), fma_(-k, N.z, I.z));
}
inline float clampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return vec3(clampf(v.x, lo.x, hi.x), clampf(v.y, lo.y, hi.y), clampf(v.z, lo.z, hi.z)); }
inline float mixf(float x, float y, float a) { return fma_(y, a, x * (1.0f - a)); }
inline vec3 mix(vec3 x, vec3 y, float a) { return vec3(mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)); }
inline vec2 mix(vec2 x, vec2 y, float a) { return vec2(mixf(x.x, y.x, a), mixf(x.y, y.y, a)); }
inline float stepf(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstepf(float e0, float e1, float x) {
  float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
inline float modf_gl(float x, float y) { return x - y * std::floor(x / y); }
inline vec3 vmin(vec3 a, vec3 b) { return vec3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
inline vec3 vmax(vec3 a, vec3 b) { return vec3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
inline vec4 vmax(vec4 a, vec4 b) { return vec4(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z), gmax(a.w, b.w)); }
inline vec3 vabs(vec3 a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline vec3 vlog(vec3 a) { return vec3(logcr(a.x), logcr(a.y), logcr(a.z)); }
inline vec3 vexp(vec3 a) { return vec3(expcr(a.x), expcr(a.y), expcr(a.z)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

// ---------------------------------------------------------------- mat4 (column-major)
struct mat4 {
  float m[16];
  float at(int col, int row) const { return m[col * 4 + row]; }  // GLSL M[col][row]
};
inline vec4 operator*(const mat4& M, vec4 v) {  // M * v
  vec4 r;
  for (int row = 0; row < 4; row++)
    r[row] = fma_(M.at(0, row), v.x, fma_(M.at(1, row), v.y, fma_(M.at(2, row), v.z, M.at(3, row) * v.w)));
  return r;
}
inline vec4 operator*(vec4 v, const mat4& M) {  // v * M  == transpose(M) * v
  vec4 r;
  for (int col = 0; col < 4; col++)
    r[col] = fma_(v.x, M.at(col, 0), fma_(v.y, M.at(col, 1), fma_(v.z, M.at(col, 2), v.w * M.at(col, 3))));
  return r;
}
inline mat4 operator*(const mat4& A, const mat4& B) {
  mat4 R;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      float s = 0;
      for (int k = 0; k < 4; k++) s += A.at(k, r) * B.at(c, k);
      R.m[c * 4 + r] = s;
    }
  return R;
}
inline mat4 load_mat4(const float* p) { mat4 M; std::memcpy(M.m, p, 64); return M; }

// ---------------------------------------------------------------- bit casts / half
inline uint32_t floatBitsToUint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// fp32 -> fp16, round-to-nearest-even, denormals kept, overflow -> inf
inline uint16_t float_to_half(float f) {
  uint32_t x = floatBitsToUint(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) return (uint16_t)(sign | (ax > 0x7f800000u ? 0x7e00u : 0x7c00u));  // nan / inf
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // >= 65520 rounds to inf
  if (ax < 0x33000001u) return (uint16_t)sign;                // < 2^-25 (or == 2^-25: ties to even 0)
  int32_t e = (int32_t)(ax >> 23) - 127;
  uint32_t mant = (ax & 0x7fffffu) | 0x800000u;
  if (e < -14) {  // half denormal
    int shift = -14 - e + 13;  // bits to drop from the 24-bit significand
    uint32_t q = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | q);
  }
  uint32_t q = ((uint32_t)(e + 15) << 10) | ((mant >> 13) & 0x3ffu);
  uint32_t rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;  // may carry into exponent (correct)
  return (uint16_t)(sign | q);
}
inline float half_to_float(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu;
  uint32_t m = h & 0x3ffu;
  if (e == 0) {
    if (m == 0) return uintBitsToFloat(sign);
    float v = std::ldexp((float)m, -24);
    return sign ? -v : v;
  }
  if (e == 31) return uintBitsToFloat(sign | 0x7f800000u | (m << 13));
  return uintBitsToFloat(sign | ((e + 112) << 23) | (m << 13));
}
inline uint32_t packHalf2x16(vec2 v) { return (uint32_t)float_to_half(v.x) | ((uint32_t)float_to_half(v.y) << 16); }
inline vec2 unpackHalf2x16(uint32_t u) { return vec2(half_to_float((uint16_t)(u & 0xffffu)), half_to_float((uint16_t)(u >> 16))); }

// ---------------------------------------------------------------- samplers
enum Fmt { F_R32F = 0, F_RGBA32F = 1, F_RGBA16F = 2, F_RGBA8 = 3 };
struct Tex {
  const void* data = nullptr;  // tightly packed, row 0 = v 0
  int w = 0, h = 0;
  int fmt = F_RGBA32F;
  bool linear = false;
  bool repeat = false;
  bool null() const { return data == nullptr; }
};
inline int wrapi(int i, int n, bool repeat) {
  if (repeat) { i %= n; if (i < 0) i += n; return i; }
  return gmin(gmax(i, 0), n - 1);
}
inline vec4 texel(const Tex& t, int ix, int iy) {
  ix = wrapi(ix, t.w, t.repeat);
  iy = wrapi(iy, t.h, t.repeat);
  size_t i = (size_t)iy * t.w + ix;
  switch (t.fmt) {
    case F_R32F: { float v = ((const float*)t.data)[i]; return vec4(v, 0, 0, 1); }
    case F_RGBA32F: { const float* p = (const float*)t.data + 4 * i; return vec4(p[0], p[1], p[2], p[3]); }
    case F_RGBA16F: { const uint16_t* p = (const uint16_t*)t.data + 4 * i;
      return vec4(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3])); }
    default: { const uint8_t* p = (const uint8_t*)t.data + 4 * i;
      return vec4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f); }
  }
}
inline vec4 texelFetch(const Tex& t, int ix, int iy) {
  if (t.null()) return vec4(0, 0, 0, 1);
  return texel(t, ix, iy);
}
// textureLod(tex, uv, 0.) honouring the texture's filter / wrap state
inline vec4 textureLod0(const Tex& t, vec2 uv) {
  if (t.null()) return vec4(0, 0, 0, 1);
  if (!t.linear) {
    int ix = (int)std::floor(uv.x * (float)t.w);
    int iy = (int)std::floor(uv.y * (float)t.h);
    return texel(t, ix, iy);
  }
  float fx = uv.x * (float)t.w - 0.5f, fy = uv.y * (float)t.h - 0.5f;
  float x0 = std::floor(fx), y0 = std::floor(fy);
  float ax = fx - x0, ay = fy - y0;
  int ix = (int)x0, iy = (int)y0;
  vec4 t00 = texel(t, ix, iy), t10 = texel(t, ix + 1, iy), t01 = texel(t, ix, iy + 1), t11 = texel(t, ix + 1, iy + 1);
  float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
  vec4 r;
  for (int c = 0; c < 4; c++) r[c] = fma_(t11[c], w11, fma_(t01[c], w01, fma_(t10[c], w10, t00[c] * w00)));
  return r;
}

// mip chain (box filter) of an RGBA16F texture: levels stored as separate Tex
struct MipTex {
  static const int MAXL = 16;
  Tex level[MAXL];
  int levels = 0;
};
inline vec4 textureLodTrilinear(const MipTex& m, vec2 uv, float lod) {
  if (m.levels == 0) return vec4(0, 0, 0, 1);
  float l = clampf(lod, 0.0f, (float)(m.levels - 1));
  int l0 = (int)std::floor(l);
  int l1 = gmin(l0 + 1, m.levels - 1);
  float f = l - (float)l0;
  vec4 a = textureLod0(m.level[l0], uv);
  if (f == 0.0f || l1 == l0) return a;
  vec4 b = textureLod0(m.level[l1], uv);
  vec4 r;
  for (int c = 0; c < 4; c++) r[c] = mixf(a[c], b[c], f);
  return r;
}

}  // namespace gl
