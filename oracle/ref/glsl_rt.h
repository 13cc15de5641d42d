// oracle/ref/glsl_rt.h — TEST INFRASTRUCTURE ONLY.
//
// A GLSL ES 3.00 *language runtime* for C++20: enough of the GLSL type system (vecN / ivecN / uvecN with full
// swizzles, mat2/3/4, the built-in function library, sampler2D, quad derivatives, `discard`) that the reference's
// fragment shaders — the text under /root/reference/src/**/shader/*.frag|*.glsl, assembled the way the reference's
// JS assembles it (oracle/ref/assemble.py) and mechanically re-spelled where GLSL and C++ differ in syntax
// (oracle/ref/transpile.py) — compile with g++ and run on the CPU.  Nothing in this file restates any shader: it
// plays the role of the GLSL compiler + GL driver.  The compiled shaders (oracle/_ref/*.so) are "the reference run
// here": they pin oracle/rfx_oracle.cpp (tests/test_reference_glsl.py) and generate tests/golden/ref_glsl/*.npz.
//
// Where GLSL ES leaves behaviour to the implementation, this runtime takes the same decisions as oracle/glsl.h
// (it reuses its scalar helpers for exactly that reason): RNE packHalf2x16, fp32 bilinear weights, correctly
// rounded transcendentals, dot / mix / M*v as fma chains, vector / scalar as reciprocal-multiply, NaN-ignoring
// min / max, "fine" 2x2-quad derivatives with all four pixels of a quad executing every derivative site (the
// behaviour of the SIMD software rasterisers BASELINE.json names — swiftshader / llvmpipe — where control flow is
// a lane mask and helper lanes keep computing), null sampler = (0,0,0,1).
#pragma once
#include "../glsl.h"

#include <array>
#include <cstdint>
#include <cstring>
#include <source_location>
#include <type_traits>
#include <vector>

namespace rt {

using uint = unsigned int;

template <class T, int N>
struct vec;

// ------------------------------------------------------------------------------------------------ swizzle proxy
// Lives inside the anonymous union of vec<T,N>, aliasing its storage.
template <class T, int N, int... I>
struct swz {
  T d[N];
  static constexpr int M = sizeof...(I);
  using V = vec<T, M>;
  operator V() const {
    V r;
    const int idx[M] = {I...};
    for (int k = 0; k < M; k++) r.d[k] = d[idx[k]];
    return r;
  }
  swz& operator=(const V& v) {
    const V t = v;  // the source may alias this storage (v.xy = v.yx)
    const int idx[M] = {I...};
    for (int k = 0; k < M; k++) d[idx[k]] = t.d[k];
    return *this;
  }
  swz& operator=(const swz& o) { return *this = V(o); }
  template <int N2, int... J>
  swz& operator=(const swz<T, N2, J...>& o) { return *this = vec<T, sizeof...(J)>(o); }
  swz& operator+=(const V& v) { return *this = V(*this) + v; }
  swz& operator-=(const V& v) { return *this = V(*this) - v; }
  swz& operator*=(const V& v) { return *this = V(*this) * v; }
  swz& operator/=(const V& v) { return *this = V(*this) / v; }
  swz& operator+=(T s) { return *this = V(*this) + s; }
  swz& operator-=(T s) { return *this = V(*this) - s; }
  swz& operator*=(T s) { return *this = V(*this) * s; }
  swz& operator/=(T s) { return *this = V(*this) / s; }
  T operator[](int i) const { const int idx[M] = {I...}; return d[idx[i]]; }
};

// ------------------------------------------------------------------------------------------------ vectors
#define RT_VEC_COMMON(N)                                                              \
  T& operator[](int i) { return d[i]; }                                               \
  const T& operator[](int i) const { return d[i]; }                                   \
  vec(const vec& o) { for (int k = 0; k < N; k++) d[k] = o.d[k]; }                    \
  vec& operator=(const vec& o) { for (int k = 0; k < N; k++) d[k] = o.d[k]; return *this; } \
  template <class U>                                                                  \
  explicit vec(const vec<U, N>& o) { for (int k = 0; k < N; k++) d[k] = (T)o.d[k]; }  \
  template <class U, int N2, int... J>                                                \
  explicit vec(const swz<U, N2, J...>& o) { const vec<U, sizeof...(J)> t = o; static_assert(sizeof...(J) >= N); for (int k = 0; k < N; k++) d[k] = (T)t.d[k]; }

template <class T>
struct vec<T, 2> {
  union {
    T d[2];
    struct { T x, y; };
    struct { T r, g; };
    struct { T s, t; };
#include "swz2.inc"
  };
  vec() : d{T(0), T(0)} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic_v<A>>>
  explicit vec(A a) : d{(T)a, (T)a} {}
  vec(T a, T b) : d{a, b} {}
  explicit vec(const vec<T, 3>& o);
  explicit vec(const vec<T, 4>& o);
  RT_VEC_COMMON(2)
};
template <class T>
struct vec<T, 3> {
  union {
    T d[3];
    struct { T x, y, z; };
    struct { T r, g, b; };
    struct { T s, t, p; };
#include "swz3.inc"
  };
  vec() : d{T(0), T(0), T(0)} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic_v<A>>>
  explicit vec(A a) : d{(T)a, (T)a, (T)a} {}
  vec(T a, T b, T c) : d{a, b, c} {}
  vec(const vec<T, 2>& a, T c) : d{a.d[0], a.d[1], c} {}
  vec(T a, const vec<T, 2>& b) : d{a, b.d[0], b.d[1]} {}
  explicit vec(const vec<T, 4>& o);
  RT_VEC_COMMON(3)
};
template <class T>
struct vec<T, 4> {
  union {
    T d[4];
    struct { T x, y, z, w; };
    struct { T r, g, b, a; };
    struct { T s, t, p, q; };
#include "swz4.inc"
  };
  vec() : d{T(0), T(0), T(0), T(0)} {}
  template <class A, class = std::enable_if_t<std::is_arithmetic_v<A>>>
  explicit vec(A a_) : d{(T)a_, (T)a_, (T)a_, (T)a_} {}
  vec(T a_, T b_, T c, T e) : d{a_, b_, c, e} {}
  vec(const vec<T, 3>& v, T e) : d{v.d[0], v.d[1], v.d[2], e} {}
  vec(T a_, const vec<T, 3>& v) : d{a_, v.d[0], v.d[1], v.d[2]} {}
  vec(const vec<T, 2>& u, const vec<T, 2>& v) : d{u.d[0], u.d[1], v.d[0], v.d[1]} {}
  vec(const vec<T, 2>& u, T c, T e) : d{u.d[0], u.d[1], c, e} {}
  vec(T a_, const vec<T, 2>& u, T e) : d{a_, u.d[0], u.d[1], e} {}
  vec(T a_, T b_, const vec<T, 2>& u) : d{a_, b_, u.d[0], u.d[1]} {}
  RT_VEC_COMMON(4)
};
template <class T> vec<T, 2>::vec(const vec<T, 3>& o) : d{o.d[0], o.d[1]} {}
template <class T> vec<T, 2>::vec(const vec<T, 4>& o) : d{o.d[0], o.d[1]} {}
template <class T> vec<T, 3>::vec(const vec<T, 4>& o) : d{o.d[0], o.d[1], o.d[2]} {}

using vec2 = vec<float, 2>; using vec3 = vec<float, 3>; using vec4 = vec<float, 4>;
using ivec2 = vec<int, 2>; using ivec3 = vec<int, 3>; using ivec4 = vec<int, 4>;
using uvec2 = vec<uint, 2>; using uvec3 = vec<uint, 3>; using uvec4 = vec<uint, 4>;
using bvec2 = vec<bool, 2>; using bvec3 = vec<bool, 3>; using bvec4 = vec<bool, 4>;

// Non-template overloads per concrete type, so that swizzle proxies convert implicitly in every operand position.
#define RT_ARITH(T, N, op)                                                                                                         \
  inline vec<T, N> operator op(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] op b.d[k]; return r; } \
  inline vec<T, N> operator op(const vec<T, N>& a, T s) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] op s; return r; }     \
  inline vec<T, N> operator op(T s, const vec<T, N>& a) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = s op a.d[k]; return r; }     \
  inline vec<T, N>& operator op##=(vec<T, N>& a, const vec<T, N>& b) { a = a op b; return a; }                                            \
  inline vec<T, N>& operator op##=(vec<T, N>& a, T s) { a = a op s; return a; }
#define RT_INTOPS(T, N) RT_ARITH(T, N, +) RT_ARITH(T, N, -) RT_ARITH(T, N, *) RT_ARITH(T, N, /) RT_ARITH(T, N, %) RT_ARITH(T, N, ^) RT_ARITH(T, N, &) RT_ARITH(T, N, |) RT_ARITH(T, N, >>) RT_ARITH(T, N, <<)
RT_INTOPS(int, 2) RT_INTOPS(int, 3) RT_INTOPS(int, 4) RT_INTOPS(uint, 2) RT_INTOPS(uint, 3) RT_INTOPS(uint, 4)
// float vectors: + - * componentwise; division: v / v and s / v are IEEE per component, v / s is one IEEE reciprocal and a
// multiply per component (how GLSL compilers lower it; oracle/glsl.h)
#define RT_FLOATOPS(N)                                                                                                             \
  RT_ARITH(float, N, +) RT_ARITH(float, N, -) RT_ARITH(float, N, *)                                                                \
  inline vec<float, N> operator/(const vec<float, N>& a, const vec<float, N>& b) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] / b.d[k]; return r; } \
  inline vec<float, N> operator/(float s, const vec<float, N>& a) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = s / a.d[k]; return r; } \
  inline vec<float, N> operator/(const vec<float, N>& a, float s) { const float q = 1.0f / s; vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] * q; return r; } \
  inline vec<float, N>& operator/=(vec<float, N>& a, const vec<float, N>& b) { a = a / b; return a; }                              \
  inline vec<float, N>& operator/=(vec<float, N>& a, float s) { a = a / s; return a; }                                             \
  inline vec<float, N> operator-(const vec<float, N>& a) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = -a.d[k]; return r; } \
  inline bool operator==(const vec<float, N>& a, const vec<float, N>& b) { for (int k = 0; k < N; k++) if (!(a.d[k] == b.d[k])) return false; return true; } \
  inline bool operator!=(const vec<float, N>& a, const vec<float, N>& b) { return !(a == b); }
RT_FLOATOPS(2) RT_FLOATOPS(3) RT_FLOATOPS(4)
#define RT_INTNEG(T, N) inline vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = (T)(0 - a.d[k]); return r; }
RT_INTNEG(int, 2) RT_INTNEG(int, 3) RT_INTNEG(int, 4)

// ------------------------------------------------------------------------------------------------ scalar built-ins
inline float radians(float x) { return x * 0.017453292519943295f; }
inline float sin(float x) { return gl::sincr(x); }
inline float cos(float x) { return gl::coscr(x); }
inline float tan(float x) { return (float)std::tan((double)x); }
inline float asin(float x) { return (float)std::asin((double)x); }
inline float acos(float x) { return gl::acoscr(x); }
inline float atan(float y, float x) { return gl::atan2cr(y, x); }
inline float atan(float x) { return (float)std::atan((double)x); }
inline float pow(float x, float y) { return gl::powcr(x, y); }
inline float exp(float x) { return gl::expcr(x); }
inline float log(float x) { return gl::logcr(x); }
inline float exp2(float x) { return gl::exp2cr(x); }
inline float log2(float x) { return gl::log2cr(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float round(float x) { return std::round(x); }
inline float trunc(float x) { return std::trunc(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float mod(float x, float y) { return gl::modf_gl(x, y); }
inline float min(float a, float b) { return gl::gmin(a, b); }
inline float max(float a, float b) { return gl::gmax(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return gl::clampf(x, lo, hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return gl::mixf(x, y, a); }
inline float step(float edge, float x) { return gl::stepf(edge, x); }
inline float smoothstep(float e0, float e1, float x) { return gl::smoothstepf(e0, e1, x); }
inline bool isnan(float x) { return std::isnan(x); }
inline bool isinf(float x) { return std::isinf(x); }
inline uint floatBitsToUint(float f) { return gl::floatBitsToUint(f); }
inline int floatBitsToInt(float f) { return (int)gl::floatBitsToUint(f); }
inline float uintBitsToFloat(uint u) { return gl::uintBitsToFloat(u); }
inline float intBitsToFloat(int u) { return gl::uintBitsToFloat((uint)u); }
inline uint packHalf2x16(const vec2& v) { return gl::packHalf2x16(gl::vec2(v.x, v.y)); }
inline vec2 unpackHalf2x16(uint u) { const gl::vec2 r = gl::unpackHalf2x16(u); return vec2(r.x, r.y); }

// ------------------------------------------------------------------------------------------------ vector built-ins
#define RT_MAP1(N, f) inline vec<float, N> f(const vec<float, N>& a) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = f(a.d[k]); return r; }
#define RT_MAP2(N, f)                                                                                                                               \
  inline vec<float, N> f(const vec<float, N>& a, const vec<float, N>& b) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = f(a.d[k], b.d[k]); return r; } \
  inline vec<float, N> f(const vec<float, N>& a, float b) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = f(a.d[k], b); return r; }
#define RT_VFUNCS(N)                                                                                                                                \
  RT_MAP1(N, sin) RT_MAP1(N, cos) RT_MAP1(N, tan) RT_MAP1(N, exp) RT_MAP1(N, log) RT_MAP1(N, exp2) RT_MAP1(N, log2) RT_MAP1(N, sqrt)                  \
  RT_MAP1(N, inversesqrt) RT_MAP1(N, abs) RT_MAP1(N, sign) RT_MAP1(N, floor) RT_MAP1(N, ceil) RT_MAP1(N, round) RT_MAP1(N, fract) RT_MAP1(N, radians) \
  RT_MAP2(N, min) RT_MAP2(N, max) RT_MAP2(N, mod) RT_MAP2(N, pow)                                                                                    \
  inline vec<float, N> step(const vec<float, N>& e, const vec<float, N>& x) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = step(e.d[k], x.d[k]); return r; } \
  inline vec<float, N> step(float e, const vec<float, N>& x) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = step(e, x.d[k]); return r; }     \
  inline vec<float, N> clamp(const vec<float, N>& x, const vec<float, N>& lo, const vec<float, N>& hi) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = clamp(x.d[k], lo.d[k], hi.d[k]); return r; } \
  inline vec<float, N> clamp(const vec<float, N>& x, float lo, float hi) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = clamp(x.d[k], lo, hi); return r; } \
  inline vec<float, N> mix(const vec<float, N>& x, const vec<float, N>& y, float a) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = mix(x.d[k], y.d[k], a); return r; } \
  inline vec<float, N> mix(const vec<float, N>& x, const vec<float, N>& y, const vec<float, N>& a) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = mix(x.d[k], y.d[k], a.d[k]); return r; } \
  inline vec<float, N> smoothstep(float e0, float e1, const vec<float, N>& x) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = smoothstep(e0, e1, x.d[k]); return r; } \
  inline vec<float, N> smoothstep(const vec<float, N>& e0, const vec<float, N>& e1, const vec<float, N>& x) { vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = smoothstep(e0.d[k], e1.d[k], x.d[k]); return r; } \
  inline float length(const vec<float, N>& a) { return std::sqrt(dot(a, a)); }                                                                       \
  inline float distance(const vec<float, N>& a, const vec<float, N>& b) { return length(a - b); }                                                   \
  inline vec<float, N> normalize(const vec<float, N>& a) { return a * (1.0f / std::sqrt(dot(a, a))); }                                             \
  inline vec<float, N> reflect(const vec<float, N>& I, const vec<float, N>& n) { const float k2 = 2.0f * dot(n, I); vec<float, N> r; for (int k = 0; k < N; k++) r.d[k] = gl::fma_(-k2, n.d[k], I.d[k]); return r; } \
  inline vec<bool, N> lessThan(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] < b.d[k]; return r; } \
  inline vec<bool, N> greaterThan(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] > b.d[k]; return r; } \
  inline vec<bool, N> lessThanEqual(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] <= b.d[k]; return r; } \
  inline vec<bool, N> greaterThanEqual(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] >= b.d[k]; return r; } \
  inline vec<bool, N> equal(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] == b.d[k]; return r; } \
  inline vec<bool, N> notEqual(const vec<float, N>& a, const vec<float, N>& b) { vec<bool, N> r; for (int k = 0; k < N; k++) r.d[k] = a.d[k] != b.d[k]; return r; } \
  inline bool any(const vec<bool, N>& a) { for (int k = 0; k < N; k++) if (a.d[k]) return true; return false; }                                     \
  inline bool all(const vec<bool, N>& a) { for (int k = 0; k < N; k++) if (!a.d[k]) return false; return true; }
// dot: fma chain, last component outermost (oracle/glsl.h)
inline float dot(const vec2& a, const vec2& b) { return gl::fma_(a.y, b.y, a.x * b.x); }
inline float dot(const vec3& a, const vec3& b) { return gl::fma_(a.z, b.z, gl::fma_(a.y, b.y, a.x * b.x)); }
inline float dot(const vec4& a, const vec4& b) { return gl::fma_(a.w, b.w, gl::fma_(a.z, b.z, gl::fma_(a.y, b.y, a.x * b.x))); }
RT_VFUNCS(2) RT_VFUNCS(3) RT_VFUNCS(4)
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(gl::fma_(a.y, b.z, -(b.y * a.z)), gl::fma_(a.z, b.x, -(b.z * a.x)), gl::fma_(a.x, b.y, -(b.x * a.y)));
}
#define RT_IMINMAX(T, N)                                                                                                                 \
  inline vec<T, N> min(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = min(a.d[k], b.d[k]); return r; } \
  inline vec<T, N> max(const vec<T, N>& a, const vec<T, N>& b) { vec<T, N> r; for (int k = 0; k < N; k++) r.d[k] = max(a.d[k], b.d[k]); return r; } \
  inline vec<T, N> clamp(const vec<T, N>& x, const vec<T, N>& lo, const vec<T, N>& hi) { return min(max(x, lo), hi); }
RT_IMINMAX(int, 2) RT_IMINMAX(int, 3) RT_IMINMAX(int, 4) RT_IMINMAX(uint, 2) RT_IMINMAX(uint, 3) RT_IMINMAX(uint, 4)

// GLSL `inout` / `out` parameter: copy in at the call, copy out when the call's full expression ends.  Binds to a plain lvalue
// or to a swizzle.
template <class V>
struct io {
  V v;
  void* dst;
  void (*back)(void*, const V&);
  io(V& r) : v(r), dst(&r), back([](void* d, const V& x) { *(V*)d = x; }) {}
  template <class T, int N, int... I>
  io(swz<T, N, I...>& s) : v(s), dst(&s), back([](void* d, const V& x) { *(swz<T, N, I...>*)d = x; }) {}
  io(const io&) = delete;
  ~io() { back(dst, v); }
};

template <class T, class... A>
inline std::array<T, sizeof...(A)> make_arr(A... a) { return std::array<T, sizeof...(A)>{T(a)...}; }

// ------------------------------------------------------------------------------------------------ matrices (column-major)
template <int N>
struct mat {
  vec<float, N> c[N];
  mat() {}
  explicit mat(float diag) { for (int k = 0; k < N; k++) c[k].d[k] = diag; }
  vec<float, N>& operator[](int i) { return c[i]; }
  const vec<float, N>& operator[](int i) const { return c[i]; }
};
struct mat2 : mat<2> {
  mat2() {}
  explicit mat2(float s) : mat<2>(s) {}
  mat2(const mat<2>& m) : mat<2>(m) {}
  mat2(float a, float b, float cc, float e) { c[0] = vec2(a, b); c[1] = vec2(cc, e); }
  mat2(const vec2& a, const vec2& b) { c[0] = a; c[1] = b; }
};
struct mat4;
struct mat3 : mat<3> {
  mat3() {}
  explicit mat3(float s) : mat<3>(s) {}
  mat3(const mat<3>& m) : mat<3>(m) {}
  mat3(const vec3& a, const vec3& b, const vec3& cc) { c[0] = a; c[1] = b; c[2] = cc; }
  mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
  explicit mat3(const mat4& m);
};
struct mat4 : mat<4> {
  mat4() {}
  explicit mat4(float s) : mat<4>(s) {}
  mat4(const mat<4>& m) : mat<4>(m) {}
  mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& e) { c[0] = a; c[1] = b; c[2] = cc; c[3] = e; }
};
inline mat3::mat3(const mat4& m) { for (int k = 0; k < 3; k++) c[k] = vec3(m.c[k].x, m.c[k].y, m.c[k].z); }
// M * v and v * M: fma chain per component, last term innermost (oracle/glsl.h); mat2 products are plain mul/add
inline vec4 operator*(const mat4& M, const vec4& v) {
  vec4 r;
  for (int row = 0; row < 4; row++) r.d[row] = gl::fma_(M.c[0].d[row], v.x, gl::fma_(M.c[1].d[row], v.y, gl::fma_(M.c[2].d[row], v.z, M.c[3].d[row] * v.w)));
  return r;
}
inline vec4 operator*(const vec4& v, const mat4& M) {
  vec4 r;
  for (int col = 0; col < 4; col++) r.d[col] = gl::fma_(v.x, M.c[col].x, gl::fma_(v.y, M.c[col].y, gl::fma_(v.z, M.c[col].z, v.w * M.c[col].w)));
  return r;
}
inline vec3 operator*(const mat3& M, const vec3& v) {
  vec3 r;
  for (int row = 0; row < 3; row++) r.d[row] = gl::fma_(M.c[0].d[row], v.x, gl::fma_(M.c[1].d[row], v.y, M.c[2].d[row] * v.z));
  return r;
}
inline vec3 operator*(const vec3& v, const mat3& M) {
  vec3 r;
  for (int col = 0; col < 3; col++) r.d[col] = gl::fma_(v.x, M.c[col].x, gl::fma_(v.y, M.c[col].y, v.z * M.c[col].z));
  return r;
}
inline vec2 operator*(const mat2& M, const vec2& v) { return vec2(M.c[0].x * v.x + M.c[1].x * v.y, M.c[0].y * v.x + M.c[1].y * v.y); }
inline vec2 operator*(const vec2& v, const mat2& M) { return vec2(v.x * M.c[0].x + v.y * M.c[0].y, v.x * M.c[1].x + v.y * M.c[1].y); }
template <int N>
inline mat<N> matmul(const mat<N>& A, const mat<N>& B) {
  mat<N> R;
  for (int cc = 0; cc < N; cc++)
    for (int r = 0; r < N; r++) {
      float s = 0;
      for (int k = 0; k < N; k++) s += A.c[k].d[r] * B.c[cc].d[k];
      R.c[cc].d[r] = s;
    }
  return R;
}
inline mat4 operator*(const mat4& A, const mat4& B) { return mat4(matmul<4>(A, B)); }
inline mat3 operator*(const mat3& A, const mat3& B) { return mat3(matmul<3>(A, B)); }
inline mat2 operator*(const mat2& A, const mat2& B) { return mat2(matmul<2>(A, B)); }
#define RT_MATSCALE(M, N)                                                                                              \
  inline M operator*(float s, const M& A) { M R; for (int k = 0; k < N; k++) R.c[k] = s * A.c[k]; return R; }            \
  inline M operator*(const M& A, float s) { M R; for (int k = 0; k < N; k++) R.c[k] = A.c[k] * s; return R; }            \
  inline M transpose(const M& A) { M R; for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) R.c[i].d[j] = A.c[j].d[i]; return R; }
RT_MATSCALE(mat2, 2) RT_MATSCALE(mat3, 3) RT_MATSCALE(mat4, 4)

// ------------------------------------------------------------------------------------------------ quad derivatives
// Every pixel of a 2x2 quad runs the whole shader; the k-th derivative call of a run records its argument.  The run
// is repeated with the previous run's records of all four pixels supplying the neighbours' values, until the records
// stop changing (two or three runs).  dFdx = v(x|1) - v(x&~1) within the row, dFdy likewise ("fine" derivatives).
struct DerivRec { uint32_t site, occ; float v[4]; };
struct QuadCtx {
  std::vector<DerivRec> rec[2][4];  // [generation][lane]
  int gen = 0, lane = 0;
  bool have_prev = false;
};
inline thread_local QuadCtx* g_quad = nullptr;
using srcloc = std::source_location;
inline uint32_t site_of(const srcloc& l) { return (uint32_t)l.line() * 4096u + (uint32_t)l.column(); }

// A derivative site is identified by its position in the (generated) source and by how many times this invocation has
// reached it, so that the four pixels of a quad are matched site by site even if one of them skipped another site.
template <int N>
inline void deriv_pair(const vec<float, N>& v, vec<float, N>& dx, vec<float, N>& dy, uint32_t site) {
  QuadCtx& q = *g_quad;
  auto& mine = q.rec[q.gen][q.lane];
  DerivRec r{};
  r.site = site;
  for (const DerivRec& e : mine) r.occ += e.site == site;
  for (int i = 0; i < N; i++) r.v[i] = v.d[i];
  mine.push_back(r);
  dx = vec<float, N>(); dy = vec<float, N>();
  if (!q.have_prev) return;
  auto val = [&](int lane, int i) -> float {
    for (const DerivRec& e : q.rec[q.gen ^ 1][lane])
      if (e.site == r.site && e.occ == r.occ) return e.v[i];
    return v.d[i];  // the neighbour never reached this site: no difference in that direction
  };
  const int l = q.lane;
  for (int i = 0; i < N; i++) {
    dx.d[i] = val(l | 1, i) - val(l & ~1, i);
    dy.d[i] = val(l | 2, i) - val(l & ~2, i);
  }
}
inline float dFdx(float v, srcloc L = srcloc::current()) { vec<float, 2> a(v, 0.0f), dx, dy; deriv_pair<2>(a, dx, dy, site_of(L)); return dx.x; }
inline float dFdy(float v, srcloc L = srcloc::current()) { vec<float, 2> a(v, 0.0f), dx, dy; deriv_pair<2>(a, dx, dy, site_of(L)); return dy.x; }
inline float fwidth(float v, srcloc L = srcloc::current()) { vec<float, 2> a(v, 0.0f), dx, dy; deriv_pair<2>(a, dx, dy, site_of(L)); return std::fabs(dx.x) + std::fabs(dy.x); }
#define RT_DERIV(N)                                                                                                                                          \
  inline vec<float, N> dFdx(const vec<float, N>& v, srcloc L = srcloc::current()) { vec<float, N> dx, dy; deriv_pair<N>(v, dx, dy, site_of(L)); return dx; }   \
  inline vec<float, N> dFdy(const vec<float, N>& v, srcloc L = srcloc::current()) { vec<float, N> dx, dy; deriv_pair<N>(v, dx, dy, site_of(L)); return dy; }   \
  inline vec<float, N> fwidth(const vec<float, N>& v, srcloc L = srcloc::current()) { vec<float, N> dx, dy; deriv_pair<N>(v, dx, dy, site_of(L)); return abs(dx) + abs(dy); }
RT_DERIV(2) RT_DERIV(3) RT_DERIV(4)

// ------------------------------------------------------------------------------------------------ samplers
struct TexDesc {
  gl::Tex base;
  gl::MipTex mips;  // levels > 0: mip-mapped (trilinear)
};
struct sampler2D {
  const TexDesc* t = nullptr;
};
inline vec4 to_rt(const gl::vec4& v) { return vec4(v.x, v.y, v.z, v.w); }
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float lod) {
  if (!s.t || s.t->base.null()) return vec4(0.0f, 0.0f, 0.0f, 1.0f);
  if (s.t->mips.levels > 0) return to_rt(gl::textureLodTrilinear(s.t->mips, gl::vec2(uv.x, uv.y), lod));
  return to_rt(gl::textureLod0(s.t->base, gl::vec2(uv.x, uv.y)));
}
// implicit LOD: lambda = log2(max(|dFdx(uv * size)|, |dFdy(uv * size)|)) for mip-mapped textures, level 0 otherwise.  The
// derivative site is recorded for every texture() call so that the call sequence is the same for all four quad pixels.
inline vec4 texture(const sampler2D& s, const vec2& uv, srcloc L = srcloc::current()) {
  vec2 dx, dy;
  deriv_pair<2>(uv, dx, dy, site_of(L));
  if (!s.t || s.t->base.null()) return vec4(0.0f, 0.0f, 0.0f, 1.0f);
  if (s.t->mips.levels == 0) return to_rt(gl::textureLod0(s.t->base, gl::vec2(uv.x, uv.y)));
  const vec2 sz((float)s.t->base.w, (float)s.t->base.h);
  const float rho = gl::gmax(length(dx * sz), length(dy * sz));
  const float lambda = rho > 0.0f ? gl::log2cr(rho) : -1000.0f;
  return to_rt(gl::textureLodTrilinear(s.t->mips, gl::vec2(uv.x, uv.y), lambda));
}
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int) {
  if (!s.t) return vec4(0.0f, 0.0f, 0.0f, 1.0f);
  return to_rt(gl::texelFetch(s.t->base, p.x, p.y));
}
inline ivec2 textureSize(const sampler2D& s, int) { return s.t ? ivec2(s.t->base.w, s.t->base.h) : ivec2(0, 0); }

// ------------------------------------------------------------------------------------------------ invocation state
// Base class of every transpiled shader: the per-fragment built-ins and the helper-lane bookkeeping.
struct FragBase {
  vec4 gl_FragCoord;
  bool rt_discarded = false;  // `discard` executed: no write, the lane lives on as a helper for derivatives
  bool rt_returned = false;   // main() executed `return`: outputs frozen, the lane lives on as a helper
};

inline void store_texel(void* base, int fmt, size_t idx, const vec4& v) {
  switch (fmt) {
    case gl::F_R32F: ((float*)base)[idx] = v.x; break;
    case gl::F_RGBA32F: std::memcpy((float*)base + 4 * idx, v.d, 16); break;
    case gl::F_RGBA16F: { uint16_t* p = (uint16_t*)base + 4 * idx; for (int c = 0; c < 4; c++) p[c] = gl::float_to_half(v.d[c]); break; }
    default: { uint8_t* p = (uint8_t*)base + 4 * idx; for (int c = 0; c < 4; c++) p[c] = (uint8_t)std::lround(gl::clampf(v.d[c], 0.0f, 1.0f) * 255.0f); break; }
  }
}

// Runs shader S (a FragBase with `vUv`, rt_main(), rt_outputs()) over a W x H target.  Targets keep their previous
// contents where the fragment was discarded.
template <class S>
void run_fullscreen(const S& proto, int W, int H, int n_out, void* const* out_data, const int* out_fmt) {
  const int QW = (W + 1) / 2, QH = (H + 1) / 2;
#pragma omp parallel for schedule(dynamic, 4)
  for (int qy = 0; qy < QH; qy++) {
    QuadCtx q;
    g_quad = &q;
    for (int qx = 0; qx < QW; qx++) {
      vec4 outv[4][8];
      bool dead[4] = {false, false, false, false};
      q.have_prev = false;
      for (int pass = 0; pass < 6; pass++) {
        for (int lane = 0; lane < 4; lane++) {
          const int px = 2 * qx + (lane & 1), py = 2 * qy + (lane >> 1);
          S s = proto;
          s.vUv = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
          s.gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
          q.lane = lane;
          q.rec[q.gen][lane].clear();
          s.rt_main();
          s.rt_outputs(outv[lane]);
          dead[lane] = s.rt_discarded;
        }
        bool same = q.have_prev;
        if (same)
          for (int lane = 0; lane < 4 && same; lane++) {
            const auto &a = q.rec[q.gen][lane], &b = q.rec[q.gen ^ 1][lane];
            same = a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(DerivRec)) == 0);
          }
        bool none = true;
        for (int lane = 0; lane < 4; lane++) none = none && q.rec[q.gen][lane].empty();
        q.gen ^= 1;
        q.have_prev = true;
        if (same || none) break;
      }
      for (int lane = 0; lane < 4; lane++) {
        const int px = 2 * qx + (lane & 1), py = 2 * qy + (lane >> 1);
        if (px >= W || py >= H || dead[lane]) continue;
        for (int o = 0; o < n_out; o++) store_texel(out_data[o], out_fmt[o], (size_t)py * W + px, outv[lane][o]);
      }
    }
    g_quad = nullptr;
  }
}

}  // namespace rt
