// oracle/rfx_oracle.cpp — TEST INFRASTRUCTURE ONLY.  PARITY PINNED BY THE REFERENCE'S OWN SHADERS (see below).
//
// CPU restatement (scalar fp32, C++17 + OpenMP) of the reference's per-pixel hot path, one
// function per reference fragment shader, each citing the reference file:line it follows
// (paths are relative to the reference checkout).  It exists so that tests/ can check the
// CUDA kernels, so that __graft_entry__.smoke() can check one small frame, and so that
// bench.py can time a CPU baseline.  The product (realism_effects_b200/) never links,
// imports or executes anything in this directory.
//
// HOW IT IS PINNED.  The reference ships no tests, golden images or known-answer vectors
//  and there is no GL / JS engine here, but its algorithm IS its GLSL
// text: oracle/ref/ compiles that text (read from the reference checkout, assembled the
// way the reference's JS assembles it) for the CPU on a GLSL language runtime
// (oracle/ref/glsl_rt.h) into oracle/_ref/*.so — the reference run here.  This
// restatement equals those shaders BIT FOR BIT on every plane of every pass: 17 chain
// configurations + all effect passes (tools/pin_oracle.py), 460 random option sets
// (tools/fuzz_pin.py), the chain at 1920x1080 x 3 frames, 3840x2160 x 2 and 7680x4320 x 2
// frames (profiles/r02_pin_oracle_*.json, r02_fuzz_pin_seed*.json), checked live by
// tests/test_reference_glsl.py and, where the checkout is absent, against the committed
// outputs of the reference shaders (tests/golden/, tests/test_oracle_chain_cpu.py).
// What that does NOT pin: the reference's JavaScript host logic (uniform wiring, frame
// sequencing, env CDF tables), which tests/chain_harness.py, tests/refglsl.py and
// realism_effects_b200/synth.py restate from the JS, and the behaviour GLSL leaves to the
// GL implementation, which oracle/glsl.h fixes once for both sides.
//
// Implementation-defined GL behaviour is fixed as documented in oracle/glsl.h.
#include "glsl.h"
#include "../include/rfx.h"

#include <vector>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace gl;

namespace {

const float PI_F = 3.1415926535897932384626433832795f;

// ======================================================================================
// gbuffer_packing.glsl  (src/gbuffer/shader/gbuffer_packing.glsl)
// ======================================================================================
const float ONE_SAFE = 0.999999f;
const float NON_ZERO_OFFSET = 0.0001f;
const float c_precision = 256.0f;
const float c_precisionp1 = c_precision + 1.0f;

struct Material {
  vec4 diffuse;
  vec3 normal;
  float roughness = 0, metalness = 0;
  vec3 emissive;
};

// :24-34
vec3 float2color(float value) {
  vec3 color;
  color.x = modf_gl(value, c_precisionp1) / c_precision;
  color.z = modf_gl(std::floor(value / c_precisionp1), c_precisionp1) / c_precision;
  color.y = std::floor(value / (c_precisionp1 * c_precisionp1)) / c_precision;
  color = color - NON_ZERO_OFFSET;
  color = vmax(color, vec3(0.0f));
  return color;
}
// :17-22
float color2float(vec3 color) {
  color = vmin(color + NON_ZERO_OFFSET, vec3(ONE_SAFE));
  return std::floor(color.x * c_precision + 0.5f) + std::floor(color.z * c_precision + 0.5f) * c_precisionp1 +
         std::floor(color.y * c_precision + 0.5f) * c_precisionp1 * c_precisionp1;
}
// :36-50
vec2 OctWrap(vec2 v) {
  vec2 w = vec2(1.0f - std::fabs(v.y), 1.0f - std::fabs(v.x));
  if (v.x < 0.0f) w.x = -w.x;
  if (v.y < 0.0f) w.y = -w.y;
  return w;
}
vec2 encodeOctWrap(vec3 n) {
  n = n / (std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z));
  vec2 xy = n.z > 0.0f ? vec2(n.x, n.y) : OctWrap(vec2(n.x, n.y));
  xy = xy * 0.5f + 0.5f;
  return xy;
}
// :52-59
vec3 decodeOctWrap(vec2 f) {
  f = f * 2.0f - 1.0f;
  vec3 n = vec3(f.x, f.y, 1.0f - std::fabs(f.x) - std::fabs(f.y));
  float t = gmax(-n.z, 0.0f);
  n.x += n.x >= 0.0f ? -t : t;
  n.y += n.y >= 0.0f ? -t : t;
  return normalize(n);
}
float packNormal(vec3 normal) { return uintBitsToFloat(packHalf2x16(encodeOctWrap(normal))); }          // :61
vec3 unpackNormal(float packedNormal) { return decodeOctWrap(unpackHalf2x16(floatBitsToUint(packedNormal))); }  // :63

// :65-83
vec4 packTwoVec4(vec4 v1, vec4 v2) {
  v1 = v1 + NON_ZERO_OFFSET;
  v2 = v2 + NON_ZERO_OFFSET;
  uint32_t v1r = packHalf2x16(vec2(v1.x, v1.y));
  uint32_t v1g = packHalf2x16(vec2(v1.z, v1.w));
  uint32_t v2r = packHalf2x16(vec2(v2.x, v2.y));
  uint32_t v2g = packHalf2x16(vec2(v2.z, v2.w));
  return vec4(uintBitsToFloat(v1r), uintBitsToFloat(v1g), uintBitsToFloat(v2r), uintBitsToFloat(v2g));
}
// :85-98
void unpackTwoVec4(vec4 encoded, vec4& v1, vec4& v2) {
  vec2 a = unpackHalf2x16(floatBitsToUint(encoded.x));
  vec2 b = unpackHalf2x16(floatBitsToUint(encoded.y));
  vec2 c = unpackHalf2x16(floatBitsToUint(encoded.z));
  vec2 d = unpackHalf2x16(floatBitsToUint(encoded.w));
  v1 = vec4(a.x, a.y, b.x, b.y) - NON_ZERO_OFFSET;
  v2 = vec4(c.x, c.y, d.x, d.y) - NON_ZERO_OFFSET;
}
// :127-141
vec4 encodeRGBE8(vec3 rgb) {
  vec4 e;
  float maxComponent = gmax(gmax(rgb.x, rgb.y), rgb.z);
  float fExp = std::ceil(log2cr(maxComponent));
  float s = exp2cr(fExp);
  e.x = rgb.x / s; e.y = rgb.y / s; e.z = rgb.z / s;
  e.w = (fExp + 128.0f) / 255.0f;
  return e;
}
vec3 decodeRGBE8(vec4 rgbe) {
  float fExp = rgbe.w * 255.0f - 128.0f;
  return rgbe.xyz() * exp2cr(fExp);
}
// :143-164
float vec4ToFloat(vec4 v) {
  for (int i = 0; i < 4; i++) v[i] = gmin(v[i] + NON_ZERO_OFFSET, ONE_SAFE);
  auto q = [](float t) -> uint32_t { return t > 0.0f ? (uint32_t)t : 0u; };  // uvec4(): truncation; negative / NaN -> 0 (defined here)
  uint32_t r = q(v.x * 255.0f), g = q(v.y * 255.0f), b = q(v.z * 255.0f), a = q(v.w * 255.0f);
  return uintBitsToFloat((a << 24) | (b << 16) | (g << 8) | r);
}
vec4 floatToVec4(float f) {
  uint32_t value = floatBitsToUint(f);
  vec4 v;
  v.x = (float)(value & 0xFFu) / 255.0f;
  v.y = (float)((value >> 8) & 0xFFu) / 255.0f;
  v.z = (float)((value >> 16) & 0xFFu) / 255.0f;
  v.w = (float)((value >> 24) & 0xFFu) / 255.0f;
  v = v - NON_ZERO_OFFSET;
  v = vmax(v, vec4(0.0f));
  return v;
}
// :166-178
vec4 packGBuffer(vec4 diffuse, vec3 normal, float roughness, float metalness, vec3 emissive) {
  vec4 g;
  g.x = vec4ToFloat(diffuse);
  g.y = packNormal(normal);
  g.z = color2float(vec3(roughness, metalness, 0.0f));
  g.w = vec4ToFloat(encodeRGBE8(emissive));
  return g;
}
// :181-196
Material getMaterial(const Tex& gBufferTexture, vec2 uv) {
  vec4 g = textureLod0(gBufferTexture, uv);
  Material m;
  m.diffuse = floatToVec4(g.x);
  m.normal = unpackNormal(g.y);
  vec3 rm = float2color(g.z);
  m.roughness = rm.x;
  m.metalness = rm.y;
  m.emissive = decodeRGBE8(floatToVec4(g.w));
  return m;
}

// ======================================================================================
// blue_noise.glsl  (src/utils/shader/blue_noise.glsl:9-48)
// ======================================================================================
struct BlueNoise {
  Tex tex;  // RGBA8, nearest, repeat
  void pcg4d(uint32_t v[4]) const {  // :17-28
    for (int i = 0; i < 4; i++) v[i] = v[i] * 1664525u + 1013904223u;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
    for (int i = 0; i < 4; i++) v[i] ^= v[i] >> 16;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
  }
  // blueNoise(uv, index) :37-45.   p = vUv * resolution
  vec4 sample(vec2 vUv, vec2 resolution, int index) const {
    if (index == 0) {  // :38-39  textureLod(blueNoiseTexture, uv * resolution / blueNoiseSize, 0.)
      vec2 blueNoiseSize((float)tex.w, (float)tex.h);
      vec2 c = vUv * resolution / blueNoiseSize;
      return textureLod0(tex, c);
    }
    vec2 p = vUv * resolution;
    int px = (int)p.x, py = (int)p.y;  // ivec2(p) truncates
    uint32_t ui = (uint32_t)index;     // int arithmetic wraps; uvec4(int) reinterprets
    uint32_t s1[4] = {ui, ui * 15843u, ui * 31u + 4566u, ui * 2345u + 58585u};  // :13
    pcg4d(s1);                                                                  // :32
    int sx = (px + (int)(s1[0] % 0x0fffffffu)) % tex.w;                          // :33
    int sy = (py + (int)(s1[1] % 0x0fffffffu)) % tex.h;
    return texelFetch(tex, sx, sy);
  }
};

// three.js <packing> chunk (three r156): perspectiveDepthToViewZ / orthographicDepthToViewZ
float perspectiveDepthToViewZ(float depth, float n, float f) { return (n * f) / ((f - n) * depth - f); }
float orthographicDepthToViewZ(float depth, float n, float f) { return depth * (n - f) - n; }

float luminanceSsgi(vec3 a) { return dot(vec3(0.2125f, 0.7154f, 0.0721f), a); }

struct Camera {
  mat4 projectionMatrix, projectionMatrixInverse, cameraMatrixWorld, viewMatrix;
  float cameraNear, cameraFar;
  bool perspective;
  explicit Camera(const rfx_camera& c)
      : projectionMatrix(load_mat4(c.projection)), projectionMatrixInverse(load_mat4(c.projection_inverse)),
        cameraMatrixWorld(load_mat4(c.camera_matrix_world)), viewMatrix(load_mat4(c.view_matrix)),
        cameraNear(c.near_plane), cameraFar(c.far_plane), perspective(c.perspective != 0) {}
};

inline vec2 pixelUv(int x, int y, int W, int H) { return vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H); }

// "fine" quad derivative width: |f(x^1,y)-f(x,y)| + |f(x,y^1)-f(x,y)|
template <class F>
float fwidthf(F f, int x, int y) {
  float c = f(x, y);
  return std::fabs(f(x ^ 1, y) - c) + std::fabs(f(x, y ^ 1) - c);
}
template <class F>
vec3 fwidth3(F f, int x, int y) {
  vec3 c = f(x, y);
  return vabs(f(x ^ 1, y) - c) + vabs(f(x, y ^ 1) - c);
}

// ======================================================================================
// K1  ssgi.frag + ssgi_utils.frag
// ======================================================================================
struct EnvInfo {  // struct EquirectHdrInfo  ssgi.frag:27-34
  MipTex map;     // linear-mipmap-linear, clamp  (SSGIEffect.js:324-329)
  Tex marginalWeights, conditionalWeights;  // R32F nearest (EquirectHdrInfoUniform.js:287-303)
  vec2 size;
  float totalSumWhole = 1, totalSumDecimal = 0;
};

struct SsgiShader {
  // uniforms
  Camera cam;
  Tex accumulatedTexture, depthTexture, velocityTexture, directLightTexture, gBufferTexture;
  BlueNoise bn;
  EnvInfo envMapInfo;
  float maxEnvMapMipLevel, rayDistance, thickness, envBlur;
  vec2 resolution;
  float nearMinusFar, nearMulFar, farMinusNear;
  int steps, refineSteps, mode, blueNoiseIndex;
  bool importanceSampling, missedRays, useDirectLight, USE_ENVMAP;
  // per-invocation globals
  vec2 vUv;
  Material mat;

  static constexpr float EPSILON = 0.00001f;
  static constexpr float ONE_MINUS_EPSILON = 1.0f - 0.00001f;

  SsgiShader(const rfx_ssgi_params& p) : cam(p.cam) {
    maxEnvMapMipLevel = p.max_env_map_mip_level; rayDistance = p.ray_distance; thickness = p.thickness; envBlur = p.env_blur;
    steps = p.steps; refineSteps = p.refine_steps; mode = p.mode; blueNoiseIndex = p.blue_noise_index;
    importanceSampling = p.flags & RFX_SSGI_IMPORTANCE_SAMPLING; missedRays = p.flags & RFX_SSGI_MISSED_RAYS;
    useDirectLight = p.flags & RFX_SSGI_USE_DIRECT_LIGHT; USE_ENVMAP = p.flags & RFX_SSGI_USE_ENVMAP;
    // SSGIPass.js:85-87
    nearMinusFar = cam.cameraNear - cam.cameraFar; farMinusNear = cam.cameraFar - cam.cameraNear; nearMulFar = cam.cameraNear * cam.cameraFar;
  }

  // ssgi_utils.frag:7-13
  float getViewZ(float depth) const {
    if (cam.perspective) return nearMulFar / (farMinusNear * depth - cam.cameraFar);
    return depth * nearMinusFar - cam.cameraNear;
  }
  // :17-24
  vec3 getViewPosition(float viewZ) const {
    float clipW = cam.projectionMatrix.at(2, 3) * viewZ + cam.projectionMatrix.at(3, 3);
    vec4 clipPosition = vec4((vec3(vUv, viewZ) - 0.5f) * 2.0f, 1.0f);
    clipPosition = clipPosition * clipW;
    vec3 p = (cam.projectionMatrixInverse * clipPosition).xyz();
    p.z = viewZ;
    return p;
  }
  // :26-33
  vec2 viewSpaceToScreenSpace(vec3 position) const {
    vec4 pc = cam.projectionMatrix * vec4(position, 1.0f);
    vec2 xy = vec2(pc.x, pc.y) / pc.w;
    xy = xy * 0.5f + 0.5f;
    return xy;
  }
  // :64-74
  static vec2 equirectDirectionToUv(vec3 direction) {
    vec2 uv = vec2(atan2cr(direction.z, direction.x), acoscr(direction.y));
    uv = uv / vec2(2.0f * PI_F, PI_F);
    uv.x += 0.5f;
    uv.y = 1.0f - uv.y;
    return uv;
  }
  // :77-86
  static vec3 equirectUvToDirection(vec2 uv) {
    uv.x -= 0.5f;
    uv.y = 1.0f - uv.y;
    float theta = uv.x * 2.0f * PI_F;
    float phi = uv.y * PI_F;
    float sinPhi = sincr(phi);
    return vec3(sinPhi * coscr(theta), coscr(phi), sinPhi * sincr(theta));
  }
  // :108-151
  static vec3 F_Schlick(vec3 f0, float theta) { return f0 + (vec3(1.0f) - f0) * powcr(1.0f - theta, 5.0f); }
  static float F_Schlick(float f0, float f90, float theta) { return f0 + (f90 - f0) * powcr(1.0f - theta, 5.0f); }
  static float D_GTR(float roughness, float NoH, float k) {
    float a2 = powcr(roughness, 2.0f);
    return a2 / (PI_F * powcr((NoH * NoH) * (a2 * a2 - 1.0f) + 1.0f, k));
  }
  static float SmithG(float NDotV, float alphaG) {
    float a = alphaG * alphaG;
    float b = NDotV * NDotV;
    return (2.0f * NDotV) / (NDotV + std::sqrt(a + b - a * b));
  }
  static float GGXVNDFPdf(float NoH, float NoV, float roughness) {
    float D = D_GTR(roughness, NoH, 2.0f);
    float G1 = SmithG(NoV, roughness * roughness);
    return (D * G1) / gmax(0.00001f, 4.0f * NoV);
  }
  static float GeometryTerm(float NoL, float NoV, float roughness) {
    float a2 = roughness * roughness;
    return SmithG(NoV, a2) * SmithG(NoL, a2);
  }
  static vec3 evalDisneyDiffuse(float NoL, float NoV, float LoH, float roughness, float metalness) {
    float FD90 = 0.5f + 2.0f * roughness * powcr(LoH, 2.0f);
    float a = F_Schlick(1.0f, FD90, NoL);
    float b = F_Schlick(1.0f, FD90, NoV);
    return vec3((a * b / PI_F) * (1.0f - metalness));
  }
  static vec3 evalDisneySpecular(float roughness, float NoH, float NoV, float NoL) {
    float D = D_GTR(roughness, NoH, 2.0f);
    float G = GeometryTerm(NoL, NoV, powcr(0.5f + roughness * 0.5f, 2.0f));
    return vec3(D * G / (4.0f * NoL * NoV));
  }
  // :153-170
  static vec3 SampleGGXVNDF(vec3 V, float ax, float ay, float r1, float r2) {
    vec3 Vh = normalize(vec3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    vec3 T1 = lensq > 0.0f ? vec3(-Vh.y, Vh.x, 0.0f) * inversesqrt(lensq) : vec3(1.0f, 0.0f, 0.0f);
    vec3 T2 = cross(Vh, T1);
    float r = std::sqrt(r1);
    float phi = 2.0f * PI_F * r2;
    float t1 = r * coscr(phi);
    float t2 = r * sincr(phi);
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * std::sqrt(1.0f - t1 * t1) + s * t2;
    vec3 Nh = t1 * T1 + t2 * T2 + std::sqrt(gmax(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(vec3(ax * Nh.x, ay * Nh.y, gmax(0.0f, Nh.z)));
  }
  // :172-180
  static void Onb(vec3 N, vec3& T, vec3& B) {
    vec3 up = std::fabs(N.z) < 0.9999999f ? vec3(0, 0, 1) : vec3(1, 0, 0);
    T = normalize(cross(up, N));
    B = cross(N, T);
  }
  static vec3 ToLocal(vec3 X, vec3 Y, vec3 Z, vec3 V) { return vec3(dot(V, X), dot(V, Y), dot(V, Z)); }
  static vec3 ToWorld(vec3 X, vec3 Y, vec3 Z, vec3 V) { return V.x * X + V.y * Y + V.z * Z; }
  // :183-191
  static vec3 cosineSampleHemisphere(vec3 n, vec2 u) {
    float r = std::sqrt(u.x);
    float theta = 2.0f * PI_F * u.y;
    vec3 b = normalize(cross(n, vec3(0.0f, 1.0f, 1.0f)));
    vec3 t = cross(b, n);
    return normalize(r * sincr(theta) * b + std::sqrt(1.0f - u.x) * n + r * coscr(theta) * t);
  }
  static float misHeuristic(float a, float b) { float aa = a * a, bb = b * b; return aa / (aa + bb); }  // :227-231

  // uv the CDF lookup produces for a pixel (only depends on the blue noise) — needed at all
  // four quad pixels for the implicit-LOD fetch at ssgi_utils.frag:218
  vec2 envCdfUv(int x, int y) const {
    int W = (int)resolution.x, H = (int)resolution.y;
    vec4 random = bn.sample(pixelUv(x, y, W, H), resolution, blueNoiseIndex);
    float v = textureLod0(envMapInfo.marginalWeights, vec2(random.x, 0.0f)).x;      // :212
    float u = textureLod0(envMapInfo.conditionalWeights, vec2(random.y, v)).x;      // :213
    return vec2(u, v);
  }
  // ssgi_utils.frag:210-225
  float sampleEquirectProbability(int px, int py, vec3& direction) const {
    vec2 uv = envCdfUv(px, py);
    direction = equirectUvToDirection(uv);
    // :218  texture(info.map, uv) — implicit LOD from quad derivatives (A3)
    vec2 uvx = envCdfUv(px ^ 1, py), uvy = envCdfUv(px, py ^ 1);
    vec2 ddx = (uvx - uv) * envMapInfo.size, ddy = (uvy - uv) * envMapInfo.size;
    float rho = gmax(length(ddx), length(ddy));
    float lambda = rho > 0.0f ? log2cr(rho) : -1000.0f;
    vec3 color = textureLodTrilinear(envMapInfo.map, uv, lambda).xyz();
    float totalSum = envMapInfo.totalSumWhole + envMapInfo.totalSumDecimal;
    float lum = luminanceSsgi(color);
    float pdf = lum / totalSum;
    return envMapInfo.size.x * envMapInfo.size.y * pdf;
  }

  // ssgi.frag:93-100
  static void calculateAngles(vec3& h, vec3& l, vec3& v, vec3& n, float& NoL, float& NoH, float& LoH, float& VoH) {
    h = normalize(v + l);
    NoL = clampf(dot(n, l), EPSILON, ONE_MINUS_EPSILON);
    NoH = clampf(dot(n, h), EPSILON, ONE_MINUS_EPSILON);
    LoH = clampf(dot(l, h), EPSILON, ONE_MINUS_EPSILON);
    VoH = clampf(dot(v, h), EPSILON, ONE_MINUS_EPSILON);
  }

  // ssgi.frag:311-346
  vec3 getEnvColor(vec3 l, float roughness, bool isDiffuseSample, bool isEnvSample) const {
    if (!USE_ENVMAP) return vec3(0.0f);
    vec3 reflectedWS = normalize((vec4(l, 0.0f) * cam.viewMatrix).xyz());
    float mip = envBlur * maxEnvMapMipLevel;
    if (!isDiffuseSample && roughness < 0.15f) mip *= roughness / 0.15f;
    vec3 envMapSample = textureLodTrilinear(envMapInfo.map, equirectDirectionToUv(reflectedWS), mip).xyz();  // ssgi_utils.frag:90-92
    float maxEnvLum = isEnvSample ? 100.0f : 25.0f;
    if (maxEnvLum != 0.0f) {
      float envLum = luminanceSsgi(envMapSample);
      if (envLum > maxEnvLum) envMapSample *= maxEnvLum / envLum;
    }
    return envMapSample;
  }
  // :348-360
  static float getSaturation(vec3 c) {
    float maxComponent = gmax(gmax(c.x, c.y), c.z);
    float minComponent = gmin(gmin(c.x, c.y), c.z);
    float delta = maxComponent - minComponent;
    if (maxComponent == minComponent) return 0.0f;
    return delta / maxComponent;
  }

  // :477-503
  vec2 BinarySearch(vec3& dir, vec3& hitPos) const {
    dir *= 0.5f;
    hitPos -= dir;
    for (int i = 0; i < refineSteps; i++) {
      vec2 uv = viewSpaceToScreenSpace(hitPos);
      float unpackedDepth = textureLod0(depthTexture, uv).x;
      float z = getViewZ(unpackedDepth);
      float rayHitDepthDifference = z - hitPos.z;
      dir *= 0.5f;
      if (rayHitDepthDifference >= 0.0f) hitPos -= dir; else hitPos += dir;
    }
    return viewSpaceToScreenSpace(hitPos);
  }
  // :441-475
  vec2 RayMarch(vec3& dir, vec3& hitPos, vec4 random) const {
    dir *= rayDistance / (float)steps;
    vec2 uv;
    for (int i = 1; i < steps; i++) {
      float cs = 1.0f - expcr(-0.25f * powcr((float)i + random.z - 0.5f, 2.0f));
      hitPos += dir * cs;
      uv = viewSpaceToScreenSpace(hitPos);
      float unpackedDepth = textureLod0(depthTexture, uv).x;
      float z = getViewZ(unpackedDepth);
      float rayHitDepthDifference = z - hitPos.z;
      if (rayHitDepthDifference >= 0.0f && rayHitDepthDifference < thickness) {
        if (refineSteps == 0) return uv;
        return BinarySearch(dir, hitPos);
      }
    }
    hitPos = vec3(10.0e9f);
    return uv;
  }

  // :362-439
  vec3 doSample(vec3 viewPos, vec3 viewNormal, float metalness, float roughness, bool isDiffuseSample, bool isEnvSample,
                float NoV, float NoL, float NoH, float LoH, float VoH, vec4 random, vec3& l, vec3& hitPos, bool& isMissedRay,
                vec3& brdf, float& pdf) const {
    float cosTheta = gmax(0.0f, dot(viewNormal, l));
    if (isDiffuseSample) {
      brdf = evalDisneyDiffuse(NoL, NoV, LoH, roughness, metalness);
      pdf = NoL / PI_F;
    } else {
      brdf = evalDisneySpecular(roughness, NoH, NoV, NoL);
      pdf = GGXVNDFPdf(NoH, NoV, roughness);
    }
    brdf *= cosTheta;
    pdf = gmax(EPSILON, pdf);
    hitPos = viewPos;
    vec2 coords = RayMarch(l, hitPos, random);
    bool allowMissedRays = missedRays;
    isMissedRay = hitPos.x == 10.0e9f;
    vec3 envMapSample = vec3(0.0f);
    if (isMissedRay && !allowMissedRays) return getEnvColor(l, roughness, isDiffuseSample, isEnvSample);
    vec4 velocity = textureLod0(velocityTexture, coords);  // :400 (null sampler => zero velocity, D4)
    vec2 reprojectedUv = coords - vec2(velocity.x, velocity.y);
    vec3 SSGI;
    vec3 envColor = getEnvColor(l, roughness, isDiffuseSample, isEnvSample);
    if (reprojectedUv.x >= 0.0f && reprojectedUv.x <= 1.0f && reprojectedUv.y >= 0.0f && reprojectedUv.y <= 1.0f) {
      vec4 reprojectedGI = textureLod0(accumulatedTexture, reprojectedUv);
      float saturation = getSaturation(mat.diffuse.xyz());
      vec3 rgi = mix(reprojectedGI.xyz(), vec3(luminanceSsgi(reprojectedGI.xyz())), (1.0f - roughness) * saturation * 0.4f);
      SSGI = rgi;
      float border = 0.15f;
      float borderFactor = smoothstepf(0.0f, border, coords.x) * smoothstepf(1.0f, 1.0f - border, coords.x) *
                           smoothstepf(0.0f, border, coords.y) * smoothstepf(1.0f, 1.0f - border, coords.y);
      borderFactor = std::sqrt(borderFactor);
      SSGI = mix(envColor, SSGI, borderFactor);
    } else {
      return envColor;
    }
    if (allowMissedRays) {
      float ssgiLum = luminanceSsgi(SSGI);
      float envLum = luminanceSsgi(envMapSample);
      if (envLum > ssgiLum) SSGI = envMapSample;
    }
    return SSGI;
  }

  // main()  :105-309.  Returns the packed RGBA32F output texel.
  vec4 mainPx(int px, int py) {
    int W = (int)resolution.x, H = (int)resolution.y;
    vUv = pixelUv(px, py, W, H);
    float unpackedDepth = textureLod0(depthTexture, vUv).x;
    if (unpackedDepth == 1.0f) {  // :109-113
      vec4 directLight = textureLod0(directLightTexture, vUv);
      return packTwoVec4(directLight, directLight);
    }
    mat = getMaterial(gBufferTexture, vUv);
    float roughnessSq = clampf(mat.roughness * mat.roughness, 0.000001f, 1.0f);
    float viewZ = getViewZ(unpackedDepth);
    vec3 viewPos = getViewPosition(viewZ);
    vec3 viewDir = normalize(viewPos);
    vec3 worldNormal = mat.normal;
    vec3 viewNormal = normalize((vec4(worldNormal, 0.0f) * cam.cameraMatrixWorld).xyz());
    vec3 n = viewNormal;
    vec3 v = -viewDir;
    float NoV = gmax(EPSILON, dot(n, v));
    vec3 V = (vec4(v, 0.0f) * cam.viewMatrix).xyz();
    vec3 N = worldNormal;
    vec4 random;
    vec3 H_, l, h, F, T, B, envMisDir, gi;
    vec3 diffuseGI(0.0f), specularGI(0.0f), brdf, hitPos, specularHitPos;  // zero-initialised (A5)
    Onb(N, T, B);
    V = ToLocal(T, B, N, V);
    vec3 f0 = mix(vec3(0.04f), mat.diffuse.xyz(), mat.metalness);
    float NoL, NoH, LoH, VoH, diffW, specW, invW, pdf, diffuseSamples = 0.0f, specularSamples = 0.0f;
    bool isDiffuseSample, isMissedRay;

    random = bn.sample(vUv, resolution, blueNoiseIndex);
    H_ = SampleGGXVNDF(V, roughnessSq, roughnessSq, random.x, random.y);
    if (H_.z < 0.0f) H_ = -H_;
    l = normalize(reflect(-V, H_));
    l = ToWorld(T, B, N, l);
    l = (vec4(l, 0.0f) * cam.cameraMatrixWorld).xyz();
    l = normalize(l);
    calculateAngles(h, l, v, n, NoL, NoH, LoH, VoH);

    if (mode == RFX_MODE_SSGI) {
      F = F_Schlick(f0, VoH);
      diffW = (1.0f - mat.metalness) * luminanceSsgi(mat.diffuse.xyz());
      specW = luminanceSsgi(F);
      diffW = gmax(diffW, EPSILON);
      specW = gmax(specW, EPSILON);
      invW = 1.0f / (diffW + specW);
      diffW *= invW;
      isDiffuseSample = random.z < diffW;
    } else {
      isDiffuseSample = false;
    }

    float emsPdf = 1.0f, emsProbability = 0.0f;
    bool emsIsEnvSample = false;  // uninitialised in GLSL => false / 0 (A5)
    envMisDir = vec3(0.0f);

    if (importanceSampling) {  // :197-215
      emsPdf = sampleEquirectProbability(px, py, envMisDir);
      envMisDir = normalize((vec4(envMisDir, 0.0f) * cam.cameraMatrixWorld).xyz());
      emsProbability = dot(envMisDir, viewNormal);
      emsProbability *= mat.roughness;
      emsProbability = gmin(ONE_MINUS_EPSILON, emsProbability);
      emsIsEnvSample = random.w < emsProbability;
      if (emsIsEnvSample) {
        emsPdf /= 1.0f - emsProbability;
        l = envMisDir;
        calculateAngles(h, l, v, n, NoL, NoH, LoH, VoH);
      } else {
        emsPdf = 1.0f - emsProbability;
      }
    }

    vec3 diffuseRay = emsIsEnvSample ? envMisDir : cosineSampleHemisphere(viewNormal, vec2(random.x, random.y));
    vec3 specularRay = emsIsEnvSample ? envMisDir : l;

    if (mode == RFX_MODE_SSGI && isDiffuseSample) {  // :222-242
      l = diffuseRay;
      calculateAngles(h, l, v, n, NoL, NoH, LoH, VoH);
      gi = doSample(viewPos, viewNormal, mat.metalness, roughnessSq, isDiffuseSample, emsIsEnvSample, NoV, NoL, NoH, LoH, VoH, random, l,
                    hitPos, isMissedRay, brdf, pdf);
      gi *= brdf;
      if (emsIsEnvSample) gi *= misHeuristic(emsPdf, pdf); else gi /= pdf;
      gi /= emsPdf;
      diffuseSamples++;
      diffuseGI = mix(diffuseGI, gi, 1.0f / diffuseSamples);
    }

    l = specularRay;  // :246-265
    calculateAngles(h, l, v, n, NoL, NoH, LoH, VoH);
    gi = doSample(viewPos, viewNormal, mat.metalness, roughnessSq, isDiffuseSample, emsIsEnvSample, NoV, NoL, NoH, LoH, VoH, random, l,
                  hitPos, isMissedRay, brdf, pdf);
    gi *= brdf;
    if (emsIsEnvSample) gi *= misHeuristic(emsPdf, pdf); else gi /= pdf;
    gi /= emsPdf;
    specularHitPos = hitPos;
    specularSamples++;
    specularGI = mix(specularGI, gi, 1.0f / specularSamples);

    if (useDirectLight) {  // :267-272
      vec3 directLight = textureLod0(directLightTexture, vUv).xyz();
      diffuseGI += directLight;
      specularGI += directLight;
    }

    vec4 gDiffuse, gSpecular;
    if (mode == RFX_MODE_SSGI) {
      if (diffuseSamples == 0.0f) diffuseGI = vec3(-1.0f);
      gDiffuse = vec4(diffuseGI, mat.roughness);
    }
    float rayLength = 0.0f;
    vec3 cameraPosWS = vec3(cam.cameraMatrixWorld.at(3, 0), cam.cameraMatrixWorld.at(3, 1), cam.cameraMatrixWorld.at(3, 2));
    isMissedRay = hitPos.x > 10.0e8f;
    if (!isMissedRay) {
      vec4 hitPosWS = cam.cameraMatrixWorld * vec4(specularHitPos, 1.0f);
      rayLength = distance(cameraPosWS, hitPosWS.xyz());
    }
    uint32_t packedRoughnessRayLength = packHalf2x16(vec2(rayLength, mat.roughness));
    float a = uintBitsToFloat(packedRoughnessRayLength);
    if (mode == RFX_MODE_SSGI) {
      gSpecular = vec4(specularGI, rayLength);
      return packTwoVec4(gDiffuse, gSpecular);
    }
    return vec4(specularGI, a);
  }
};

// ======================================================================================
// K2  temporal_reproject.frag + reproject.frag
// ======================================================================================
struct TemporalShader {
  Camera cam;
  mat4 prevViewMatrix, prevCameraMatrixWorld, prevProjectionMatrix, prevProjectionMatrixInverse;
  vec3 cameraPos;
  Tex inputTexture, velocityTexture, accumulatedTexture[2];
  float maxBlend, neighborhoodClampIntensity, keepData, confidencePower;
  bool fullAccumulate, logTransform;
  int textureCount, inputType;
  bool reprojectSpecular[2];
  vec2 invTexSize;
  int W, H;

  // invocation globals (reproject.frag:3-7)
  vec2 vUv, velocity;
  vec3 worldNormal, worldPos, viewDir;
  float depth, curvature, viewAngle, rayLength, angleMix;
  float roughness, moveFactor;
  vec3 reprojectedUvDiffuse, reprojectedUvSpecular;

  TemporalShader(const rfx_temporal_params& p)
      : cam(p.cam), prevViewMatrix(load_mat4(p.prev_view_matrix)), prevCameraMatrixWorld(load_mat4(p.prev_camera_matrix_world)),
        prevProjectionMatrix(load_mat4(p.prev_projection)), prevProjectionMatrixInverse(load_mat4(p.prev_projection_inverse)),
        cameraPos(p.camera_pos[0], p.camera_pos[1], p.camera_pos[2]) {
    maxBlend = p.max_blend; neighborhoodClampIntensity = p.neighborhood_clamp_intensity; keepData = p.keep_data;
    confidencePower = p.confidence_power; fullAccumulate = p.full_accumulate; logTransform = p.log_transform;
    textureCount = p.texture_count; inputType = p.input_type;
    reprojectSpecular[0] = p.reproject_specular[0]; reprojectSpecular[1] = p.reproject_specular[1];
  }

  float getViewZ(float d) const {  // reproject.frag:13-19
    return cam.perspective ? perspectiveDepthToViewZ(d, cam.cameraNear, cam.cameraFar) : orthographicDepthToViewZ(d, cam.cameraNear, cam.cameraFar);
  }
  static vec3 screenSpaceToWorldSpace(vec2 uv, float depth, const mat4& curMatrixWorld, const mat4& projMatrixInverse) {  // :21-28
    vec4 ndc = vec4((uv.x - 0.5f) * 2.0f, (uv.y - 0.5f) * 2.0f, (depth - 0.5f) * 2.0f, 1.0f);
    vec4 clip = projMatrixInverse * ndc;
    vec4 view = curMatrixWorld * (clip / clip.w);
    return view.xyz();
  }
  void transformColor(vec3& c) const { if (logTransform) c = vlog(c + 1.0f); }      // :42
  void undoColorTransform(vec3& c) const { if (logTransform) c = vexp(c) - 1.0f; }  // :43

  // :97-105
  void getVelocityNormalDepth(vec2 uv, vec2& vel, vec3& normal, float& d) const {
    vec4 t = textureLod0(velocityTexture, uv);
    vel = vec2(t.x, t.y);
    normal = unpackNormal(t.z);
    d = t.w;
  }
  // :53-81
  void getNeighborhoodAABB(int clampRadius, vec3& minC, vec3& maxC, bool isSpecular) const {
    for (int x = -clampRadius; x <= clampRadius; x++) {
      for (int y = -clampRadius; y <= clampRadius; y++) {
        vec2 offset = vec2((float)x, (float)y) * invTexSize;
        vec2 neighborUv = vUv + offset;
        vec4 neighborTexel;
        if (inputType == RFX_INPUT_DIFFUSE_SPECULAR) {
          vec4 t1, t2;
          unpackTwoVec4(textureLod0(inputTexture, neighborUv), t1, t2);
          neighborTexel = isSpecular ? t2 : t1;
        } else {
          neighborTexel = textureLod0(inputTexture, neighborUv);
        }
        if (neighborTexel.x >= 0.0f) {
          minC = vmin(neighborTexel.xyz(), minC);
          maxC = vmax(neighborTexel.xyz(), maxC);
        }
      }
    }
  }
  // :83-95
  void clampNeighborhood(vec3& color, vec3 inputColor, int clampRadius, bool isSpecular) const {
    undoColorTransform(inputColor);
    vec3 minC = inputColor, maxC = inputColor;
    getNeighborhoodAABB(clampRadius, minC, maxC, isSpecular);
    transformColor(minC);
    transformColor(maxC);
    color = clamp(color, minC, maxC);
  }
  // :130-167
  float validateReprojectedUV(vec2 reprojectedUv) {
    if (reprojectedUv.x > 1.0f || reprojectedUv.x < 0.0f || reprojectedUv.y > 1.0f || reprojectedUv.y < 0.0f) return 0.0f;
    vec2 lastVelocity; vec3 lastWorldNormal; float lastDepth;
    getVelocityNormalDepth(reprojectedUv, lastVelocity, lastWorldNormal, lastDepth);
    vec3 lastWorldPos = screenSpaceToWorldSpace(reprojectedUv, lastDepth, prevCameraMatrixWorld, prevProjectionMatrixInverse);
    vec3 lastViewPos = (prevViewMatrix * vec4(lastWorldPos, 1.0f)).xyz();
    vec3 lastViewDir = normalize(lastViewPos);
    vec3 lastViewNormal = (vec4(lastWorldNormal, 0.0f) * prevViewMatrix).xyz();
    float lastViewAngle = dot(-lastViewDir, lastViewNormal);
    angleMix = std::fabs(lastViewAngle - viewAngle);
    float viewZ = std::fabs(getViewZ(depth));
    float distFactor = 1.0f + 1.0f / (viewZ + 1.0f);
    float disoccl = 0.0f;
    disoccl += length(worldPos - lastWorldPos) / 10.0f * distFactor;                            // :120-123
    disoccl += std::fabs(dot(worldPos - lastWorldPos, worldNormal)) / 20.0f * distFactor;       // :111-118
    disoccl += gmin(1.0f - dot(worldNormal, lastWorldNormal), 1.0f) / 1.0f * distFactor;    // :125-128
    float confidence = 1.0f - gmin(disoccl, 1.0f);
    confidence = gmax(confidence, 0.0f);
    confidence = powcr(confidence, confidencePower);
    return confidence;
  }
  // :169-193
  vec2 reprojectHitPoint(vec3 rayOrig, float rayLen) const {
    if (curvature > 0.05f || rayLen < 0.01f) return vec2(-1.0f);
    vec3 cameraRay = normalize(rayOrig - cameraPos);
    vec3 parallaxHitPoint = cameraPos + cameraRay * rayLen;
    vec4 rh = prevProjectionMatrix * prevViewMatrix * vec4(parallaxHitPoint, 1.0f);
    vec2 xy = vec2(rh.x, rh.y) / rh.w;
    xy = xy * 0.5f + 0.5f;
    return xy;
  }
  // :195-209
  vec3 getReprojectedUV(bool doReprojectSpecular) {
    vec2 reprojectedUv = doReprojectSpecular ? reprojectHitPoint(worldPos, rayLength) : vUv - velocity;
    float confidence = validateReprojectedUV(reprojectedUv);
    return vec3(reprojectedUv, confidence);
  }
  // :212-255
  vec4 BiCubicCatmullRom5Tap(const Tex& tex, vec2 P) const {
    vec2 UV = P / invTexSize;
    vec2 tc = vec2(std::floor(UV.x - 0.5f), std::floor(UV.y - 0.5f)) + 0.5f;
    vec2 f = UV - tc;
    vec2 f2 = f * f;
    vec2 f3 = f2 * f;
    vec2 w0 = f2 - 0.5f * (f3 + f);
    vec2 w1 = 1.5f * f3 - 2.5f * f2 + vec2(1.0f);
    vec2 w3 = 0.5f * (f3 - f2);
    vec2 w2 = vec2(1.0f) - w0 - w1 - w3;
    vec2 Weight[3] = {w0, w1 + w2, w3};
    vec2 Sample[3] = {tc - vec2(1.0f), tc + w2 / Weight[1], tc + vec2(2.0f)};
    for (int i = 0; i < 3; i++) Sample[i] = Sample[i] * invTexSize;
    float sw[5] = {Weight[1].x * Weight[0].y, Weight[0].x * Weight[1].y, Weight[1].x * Weight[1].y, Weight[2].x * Weight[1].y,
                   Weight[1].x * Weight[2].y};
    vec4 Ct = textureLod0(tex, vec2(Sample[1].x, Sample[0].y)) * sw[0];
    vec4 Cl = textureLod0(tex, vec2(Sample[0].x, Sample[1].y)) * sw[1];
    vec4 Cc = textureLod0(tex, vec2(Sample[1].x, Sample[1].y)) * sw[2];
    vec4 Cr = textureLod0(tex, vec2(Sample[2].x, Sample[1].y)) * sw[3];
    vec4 Cb = textureLod0(tex, vec2(Sample[1].x, Sample[2].y)) * sw[4];
    float WeightMultiplier = 1.0f / (sw[0] + sw[1] + sw[2] + sw[3] + sw[4]);
    return vmax((Ct + Cl + Cc + Cr + Cb) * WeightMultiplier, vec4(0.0f));
  }

  // temporal_reproject.frag:83-122
  void reproject(vec4& inp, vec4& acc, const Tex& accumulatedTex, bool wasSampled, bool doReprojectSpecular) {
    vec3 uvc = doReprojectSpecular ? reprojectedUvSpecular : reprojectedUvDiffuse;
    vec2 uv = vec2(uvc.x, uvc.y);
    acc = BiCubicCatmullRom5Tap(accumulatedTex, uv);
    vec3 accRgb = acc.xyz();
    transformColor(accRgb);
    if (!wasSampled) {
      inp = vec4(accRgb, inp.w);
      acc = vec4(accRgb, acc.w);
      return;
    }
    acc.w++;
    vec3 clampedColor = accRgb;
    int clampRadius = doReprojectSpecular && roughness < 0.25f ? 1 : 2;
    clampNeighborhood(clampedColor, inp.xyz(), clampRadius, doReprojectSpecular);
    float r = doReprojectSpecular ? roughness : 1.0f;
    float clampAggressiveness = gmin(1.0f, uvc.z * r);
    float clampIntensity = mixf(0.0f, gmin(1.0f, moveFactor * 50.0f + neighborhoodClampIntensity), clampAggressiveness);
    vec3 newColor = mix(accRgb, clampedColor, clampIntensity);
    float colorDiff = gmin(length(newColor - accRgb), 1.0f);
    acc.w *= 1.0f - colorDiff;
    acc = vec4(newColor, acc.w);
  }
  // :42-79
  vec4 accumulate(vec4 inp, vec4 acc, bool doReprojectSpecular) const {
    vec3 ruc = doReprojectSpecular ? reprojectedUvSpecular : reprojectedUvDiffuse;
    float confidence = ruc.z;
    confidence = powcr(confidence, confidencePower);
    float accumBlend = 1.0f - 1.0f / (acc.w + 1.0f);
    accumBlend = mixf(0.0f, accumBlend, confidence);
    float maxValue = (fullAccumulate ? 1.0f : maxBlend) * keepData;
    if (inputType != RFX_INPUT_DIFFUSE) {
      const float roughnessMaximum = 0.1f;
      if (doReprojectSpecular && roughness >= 0.0f && roughness < roughnessMaximum) {
        float maxRoughnessValue = mixf(0.0f, maxValue, roughness / roughnessMaximum);
        maxValue = mixf(maxValue, maxRoughnessValue, gmin(100.0f * moveFactor, 1.0f));
      }
    }
    float temporalReprojectMix = gmin(accumBlend, maxValue);
    float a = 1.0f / (1.0f - temporalReprojectMix) - 1.0f;
    a = gmin(65536.0f, a);
    vec3 rgb = mix(inp.xyz(), acc.xyz(), temporalReprojectMix);
    undoColorTransform(rgb);
    return vec4(rgb, a);
  }

  float depthAt(int x, int y) const { return textureLod0(velocityTexture, pixelUv(x, y, W, H)).w; }
  vec3 normalAt(int x, int y) const { return unpackNormal(textureLod0(velocityTexture, pixelUv(x, y, W, H)).z); }

  // main()  temporal_reproject.frag:178-208.  returns false on discard
  bool mainPx(int px, int py, vec4 out[2]) {
    vUv = pixelUv(px, py, W, H);
    roughness = 1.0f; moveFactor = 0.0f; rayLength = 0.0f;  // reproject.frag:5-7 (rayLength uninitialised global => 0)
    reprojectedUvDiffuse = vec3(-1.0f); reprojectedUvSpecular = vec3(-1.0f);
    getVelocityNormalDepth(vUv, velocity, worldNormal, depth);
    vec4 inputTexel[2], accumulatedTexel[2];
    bool sampledThisFrame[2] = {false, false};
    // getTexels :130-145
    if (inputType == RFX_INPUT_DIFFUSE_SPECULAR) {
      unpackTwoVec4(textureLod0(inputTexture, vUv), inputTexel[0], inputTexel[1]);
    } else {
      inputTexel[0] = textureLod0(inputTexture, vUv);
    }
    int nIn = inputType == RFX_INPUT_DIFFUSE_SPECULAR ? 2 : 1;
    for (int i = 0; i < nIn; i++) {  // preprocessInput :124-128
      sampledThisFrame[i] = inputTexel[i].x >= 0.0f;
      vec3 rgb = vmax(inputTexel[i].xyz(), vec3(0.0f));
      transformColor(rgb);
      inputTexel[i] = vec4(rgb, inputTexel[i].w);
    }
    if (inputType != RFX_INPUT_DIFFUSE) {  // :188-193
      if (depth == 1.0f && fwidthf([&](int x, int y) { return depthAt(x, y); }, px, py) == 0.0f) return false;
    }
    curvature = length(fwidth3([&](int x, int y) { return normalAt(x, y); }, px, py));  // :195, reproject.frag:265-269
    // computeGVariables :147-153
    worldPos = screenSpaceToWorldSpace(vUv, depth, cam.cameraMatrixWorld, cam.projectionMatrixInverse);
    vec3 viewPos = (cam.viewMatrix * vec4(worldPos, 1.0f)).xyz();
    viewDir = normalize(viewPos);
    vec3 viewNormal = (vec4(worldNormal, 0.0f) * cam.viewMatrix).xyz();
    viewAngle = dot(-viewDir, viewNormal);
    // getRoughnessRayLength :167-176
    if (inputType == RFX_INPUT_DIFFUSE_SPECULAR) {
      rayLength = inputTexel[1].w;
      roughness = clampf(inputTexel[0].w, 0.0f, 1.0f);
    } else if (inputType == RFX_INPUT_SPECULAR) {
      vec2 data = unpackHalf2x16(floatBitsToUint(inputTexel[0].w));
      rayLength = data.x;
      roughness = clampf(data.y, 0.0f, 1.0f);
    }
    // computeReprojectedUv :155-165
    reprojectedUvDiffuse = getReprojectedUV(false);
    if (inputType == RFX_INPUT_DIFFUSE_SPECULAR || inputType == RFX_INPUT_SPECULAR) {
      reprojectedUvSpecular = getReprojectedUV(true);
      if (reprojectedUvSpecular.x == -1.0f) reprojectedUvSpecular = reprojectedUvDiffuse;
    }
    moveFactor = gmin(dot(velocity, velocity) * 10000.0f, 1.0f);
    for (int i = 0; i < textureCount; i++) {
      reproject(inputTexel[i], accumulatedTexel[i], accumulatedTexture[i], sampledThisFrame[i], reprojectSpecular[i]);
      out[i] = accumulate(inputTexel[i], accumulatedTexel[i], reprojectSpecular[i]);
    }
    return true;
  }
};

// ======================================================================================
// K3  poisson_denoise.frag
// ======================================================================================
struct PoissonShader {
  Tex inputTexture, inputTexture2, depthTexture, normalTexture, gBufferTexture;
  BlueNoise bn;
  float radius, phi, lumaPhi, depthPhi, normalPhi, roughnessPhi, specularPhi;
  vec2 resolution;
  int textureCount, blueNoiseIndex, W, H;
  bool isTextureSpecular[2], GBUFFER_TEXTURE;

  struct InputTexel { vec3 rgb; float a, luminance, w, totalWeight; bool isSpecular; };

  static float luminanceP(vec3 a) { return powcr(dot(vec3(0.2125f, 0.7154f, 0.0721f), a), 0.125f); }  // :28

  float depthAt(int x, int y) const { return textureLod0(depthTexture, pixelUv(x, y, W, H)).x; }
  vec3 normalAt(int x, int y) const {  // getNormal(mat) :80-87 with mat = getMaterial(gBufferTexture, vUv)
    vec2 uv = pixelUv(x, y, W, H);
    if (GBUFFER_TEXTURE) return getMaterial(gBufferTexture, uv).normal;
    return unpackNormal(textureLod0(normalTexture, uv).z);
  }

  // returns false on discard
  bool mainPx(int px, int py, vec4 out[2]) const {
    vec2 vUv = pixelUv(px, py, W, H);
    float depth = textureLod0(depthTexture, vUv).x;
    if (depth == 1.0f && fwidthf([&](int x, int y) { return depthAt(x, y); }, px, py) == 0.0f) return false;  // :129-132
    InputTexel inputs[2];
    for (int i = 0; i < textureCount; i++) {  // :138-164
      vec4 t = isTextureSpecular[i] ? textureLod0(textureCount == 1 ? inputTexture : inputTexture2, vUv) : textureLod0(inputTexture, vUv);
      float age = 1.0f / powcr(t.w + 1.0f, 1.2f * phi);
      vec3 rgb = t.xyz() * 1.0003f;
      rgb = vlog(rgb + 1.0f);
      inputs[i] = InputTexel{rgb, t.w, luminanceP(rgb), age, 1.0f, isTextureSpecular[i]};
    }
    Material mat = getMaterial(gBufferTexture, vUv);  // :167 (null sampler without GBUFFER_TEXTURE, A8)
    vec3 normal = normalAt(px, py);
    float glossiness = gmax(0.0f, 4.0f * (1.0f - mat.roughness / 0.25f));
    float specularFactor = expcr(-glossiness * specularPhi);
    float flatness = 1.0f - gmin(length(fwidth3([&](int x, int y) { return normalAt(x, y); }, px, py)), 1.0f);
    flatness = powcr(flatness, 2.0f) * 0.75f + 0.25f;
    vec4 random = bn.sample(vUv, resolution, blueNoiseIndex);
    float r = radius;
    float angle = random.x * 2.0f * PI_F;
    float s = sincr(angle), c = coscr(angle);
    // mat2 rm = r * flatness * mat2(c, -s, s, c): columns (c,-s) and (s,c)
    float k = r * flatness;
    float m00 = k * c, m01 = k * -s, m10 = k * s, m11 = k * c;  // m[col][row]
    const float SQ = 1.41421356237f;
    const vec2 POISSON[8] = {vec2(-1.0f, 0.0f), vec2(0.0f, -1.0f), vec2(1.0f, 0.0f), vec2(0.0f, 1.0f), vec2(-0.25f * SQ, -0.25f * SQ),
                             vec2(0.25f * SQ, -0.25f * SQ), vec2(0.25f * SQ, 0.25f * SQ), vec2(-0.25f * SQ, 0.25f * SQ)};
    for (int i = 0; i < 8; i++) {
      vec2 o = POISSON[i] / resolution;
      vec2 neighborUv = vUv + vec2(m00 * o.x + m10 * o.y, m01 * o.x + m11 * o.y);
      // getBasicNeighborWeight :52-78
      float wBasic;
      {
        vec3 neighborNormal; float neighborDepth; float neighborRoughness = 0.0f;
        if (GBUFFER_TEXTURE) {
          Material nm = getMaterial(gBufferTexture, neighborUv);
          neighborNormal = nm.normal; neighborRoughness = nm.roughness;
          neighborDepth = textureLod0(depthTexture, neighborUv).x;
        } else {
          vec4 t = textureLod0(normalTexture, neighborUv);
          neighborNormal = unpackNormal(t.z);
          neighborDepth = t.w;
        }
        if (neighborDepth == 1.0f) {
          wBasic = 0.0f;
        } else {
          float normalDiff = 1.0f - gmax(dot(normal, neighborNormal), 0.0f);
          float depthDiff = 10000.0f * std::fabs(depth - neighborDepth);
          if (GBUFFER_TEXTURE) {
            float roughnessDiff = std::fabs(mat.roughness - neighborRoughness);
            wBasic = expcr(-normalDiff * normalPhi - depthDiff * depthPhi - roughnessDiff * roughnessPhi);
          } else {
            wBasic = expcr(-normalDiff * normalPhi - depthDiff * depthPhi);
          }
        }
      }
      for (int j = 0; j < textureCount; j++) {  // applyWeight :102-124
        InputTexel& inp = inputs[j];
        float w = wBasic;
        vec4 t;
        if (inp.isSpecular) {
          t = textureLod0(textureCount == 1 ? inputTexture : inputTexture2, neighborUv);
          w *= specularFactor;
        } else {
          t = textureLod0(inputTexture, neighborUv);
        }
        vec3 trgb = vlog(t.xyz() + 1.0f);
        float disocclW = powcr(w, 0.1f);
        float lumaDiff = std::fabs(inp.luminance - luminanceP(trgb));
        lumaDiff = gmin(lumaDiff, 0.5f);
        float lumaFactor = expcr(-lumaDiff * lumaPhi);
        w = mixf(w * lumaFactor, disocclW, inp.w) * inp.w;
        w *= stepf(0.0001f, w);
        inp.rgb += w * trgb;
        inp.totalWeight += w;
      }
    }
    for (int j = 0; j < textureCount; j++) {  // outputTexel :94-100
      vec3 rgb = inputs[j].rgb / inputs[j].totalWeight;
      rgb = vexp(rgb) - 1.0f;
      out[j] = vec4(rgb, inputs[j].a);
    }
    return true;
  }
};

// ======================================================================================
// K4  DenoiserComposePass.js:58-85 + denoiser_compose_functions.glsl
// ======================================================================================
struct ComposeShader {
  Camera cam;
  Tex depthTexture, gBufferTexture, diffuseGiTexture, specularGiTexture, sceneTexture;
  int inputType, W, H;
  static constexpr float EPSILON = 1e-6f;  // three <common>

  ComposeShader(const rfx_compose_params& p) : cam(p.cam), inputType(p.input_type) {}
  float depthAt(int x, int y) const { return textureLod0(depthTexture, pixelUv(x, y, W, H)).x; }

  bool mainPx(int px, int py, vec4& out) const {
    vec2 vUv = pixelUv(px, py, W, H);
    float depth = textureLod0(depthTexture, vUv).x;
    if (depth == 1.0f && fwidthf([&](int x, int y) { return depthAt(x, y); }, px, py) == 0.0f) return false;
    Material mat = getMaterial(gBufferTexture, vUv);
    vec3 viewNormal = (vec4(mat.normal, 0.0f) * cam.cameraMatrixWorld).xyz();
    float gz = cam.perspective ? perspectiveDepthToViewZ(depth, cam.cameraNear, cam.cameraFar) : orthographicDepthToViewZ(depth, cam.cameraNear, cam.cameraFar);
    float viewZ = -gz;
    // getViewPosition  denoiser_compose_functions.glsl:13-20
    float clipW = cam.projectionMatrix.at(2, 3) * viewZ + cam.projectionMatrix.at(3, 3);
    vec4 clipPosition = vec4((vec3(vUv, viewZ) - 0.5f) * 2.0f, 1.0f);
    clipPosition = clipPosition * clipW;
    vec3 viewPos = (cam.projectionMatrixInverse * clipPosition).xyz();
    viewPos.z = -viewZ;
    vec3 viewDir = normalize(viewPos);
    vec4 diffuseGi = textureLod0(diffuseGiTexture, vUv);
    vec4 specularGi = textureLod0(specularGiTexture, vUv);
    // constructGlobalIllumination :53-107
    vec3 cameraRay = viewDir;
    float roughness = mat.roughness * mat.roughness;
    vec3 normal = (vec4(viewNormal, 0.0f) * cam.viewMatrix).xyz();
    vec3 T, B;
    vec3 v = -cameraRay;
    vec3 V = (vec4(v, 0.0f) * cam.viewMatrix).xyz();
    vec3 N = normal;
    SsgiShader::Onb(N, T, B);
    V = SsgiShader::ToLocal(T, B, N, V);
    vec3 Hh = SsgiShader::SampleGGXVNDF(V, roughness, roughness, 0.25f, 0.25f);
    if (Hh.z < 0.0f) Hh = -Hh;
    vec3 l = normalize(reflect(-V, Hh));
    l = SsgiShader::ToWorld(T, B, N, l);
    l = (vec4(l, 1.0f) * cam.cameraMatrixWorld).xyz();
    l = normalize(l);
    if (dot(viewNormal, l) < 0.0f) l = -l;
    vec3 h = normalize(v + l);
    float VoH = gmax(EPSILON, dot(v, h));
    vec3 diffuse = mat.diffuse.xyz();
    vec3 f0 = mix(vec3(0.04f), diffuse, mat.metalness);
    vec3 F = f0 + (vec3(1.0f) - f0) * powcr(1.0f - VoH, 5.0f);
    // #if inputType != TYPE_SPECULAR ... #else textureLod(sceneTexture, vUv, 0.).rgb   (denoiser_compose_functions.glsl:97-101)
    vec3 diffuseComponent = inputType != RFX_INPUT_SPECULAR ? diffuse * (1.0f - mat.metalness) * (vec3(1.0f) - F) * diffuseGi.xyz() : textureLod0(sceneTexture, vUv).xyz();
    vec3 specularComponent = specularGi.xyz() * F;
    vec3 gi = diffuseComponent + specularComponent + mat.emissive;
    out = vec4(gi, 1.0f);
    return true;
  }
};

// ======================================================================================
// K6  hbao.frag + hbao_utils.glsl (line 1's bad include dropped)
// ======================================================================================
struct HbaoShader {
  mat4 projectionViewMatrix, projectionMatrixInverse, cameraMatrixWorld;
  Tex depthTexture;
  BlueNoise bn;
  vec2 resolution;
  float aoDistance, distancePower, bias, thickness;
  int spp, blueNoiseIndex, W, H;

  vec3 getWorldPos(float depth, vec2 coord) const {  // hbao_utils.glsl:19-29
    float z = depth * 2.0f - 1.0f;
    vec4 clipSpacePosition = vec4(coord.x * 2.0f - 1.0f, coord.y * 2.0f - 1.0f, z, 1.0f);
    vec4 viewSpacePosition = projectionMatrixInverse * clipSpacePosition;
    vec4 ws = cameraMatrixWorld * viewSpacePosition;
    return ws.xyz() / ws.w;
  }
  vec3 computeWorldNormal(vec2 vUv) const {  // :46-68
    vec2 size((float)depthTexture.w, (float)depthTexture.h);
    int ix = (int)(vUv.x * size.x), iy = (int)(vUv.y * size.y);
    auto D = [&](int dx, int dy) { return texelFetch(depthTexture, ix + dx, iy + dy).x; };
    float c0 = D(0, 0), l2 = D(-2, 0), l1 = D(-1, 0), r1 = D(1, 0), r2 = D(2, 0), b2 = D(0, -2), b1 = D(0, -1), t1 = D(0, 1), t2 = D(0, 2);
    float dl = std::fabs((2.0f * l1 - l2) - c0);
    float dr = std::fabs((2.0f * r1 - r2) - c0);
    float db = std::fabs((2.0f * b1 - b2) - c0);
    float dt = std::fabs((2.0f * t1 - t2) - c0);
    vec3 ce = getWorldPos(c0, vUv);
    vec3 dpdx = (dl < dr) ? ce - getWorldPos(l1, vUv - vec2(1.0f / size.x, 0.0f)) : -ce + getWorldPos(r1, vUv + vec2(1.0f / size.x, 0.0f));
    vec3 dpdy = (db < dt) ? ce - getWorldPos(b1, vUv - vec2(0.0f, 1.0f / size.y)) : -ce + getWorldPos(t1, vUv + vec2(0.0f, 1.0f / size.y));
    return normalize(cross(dpdx, dpdy));
  }
  bool mainPx(int px, int py, vec4& out) const {  // hbao.frag:64-96
    vec2 vUv = pixelUv(px, py, W, H);
    float depth = textureLod0(depthTexture, vUv).x;
    if (depth == 1.0f) return false;
    vec3 cameraPosition = (cameraMatrixWorld * vec4(0.0f, 0.0f, 0.0f, 1.0f)).xyz();
    vec3 worldPos = getWorldPos(depth, vUv);
    vec3 worldNormal = computeWorldNormal(vUv);
    float ao = 0.0f, totalWeight = 0.0f;
    for (int i = 0; i < spp; i++) {  // getOcclusion :21-62 (same blue-noise value every iteration, A9)
      vec4 blueNoise = bn.sample(vUv, resolution, blueNoiseIndex);
      vec3 sampleWorldDir = SsgiShader::cosineSampleHemisphere(worldNormal, vec2(blueNoise.x, blueNoise.y));
      vec3 sampleWorldPos = worldPos + aoDistance * powcr(blueNoise.z, distancePower + 1.0f) * sampleWorldDir;
      vec4 sampleUv = projectionViewMatrix * vec4(sampleWorldPos, 1.0f);
      vec2 suv = vec2(sampleUv.x, sampleUv.y) / sampleUv.w;
      suv = suv * 0.5f + 0.5f;
      float sampleDepth = textureLod0(depthTexture, suv).x;
      float deltaDepth = depth - sampleDepth;
      float d = distance(sampleWorldPos, cameraPosition);
      deltaDepth *= 0.001f * d * d;
      float th = thickness * 0.01f;
      float theta = dot(worldNormal, sampleWorldDir);
      totalWeight += theta;
      if (deltaDepth < th) {
        float horizon = sampleDepth + deltaDepth * bias * 1000.0f;
        float occlusion = gmax(0.0f, horizon - depth) * theta;
        float m = gmax(0.0f, 1.0f - deltaDepth / th);
        occlusion = 10.0f * occlusion * m / d;
        occlusion = std::sqrt(occlusion);
        ao += occlusion;
      }
    }
    if (totalWeight > 0.0f) ao /= totalWeight;
    ao = clampf(1.0f - ao, 0.0f, 1.0f);
    out = vec4(worldNormal, ao);
    return true;
  }
};

// ---- helpers to build Tex from raw pointers ------------------------------------------
Tex mk(const void* p, int w, int h, int fmt, bool linear = false, bool repeat = false) {
  Tex t; t.data = p; t.w = w; t.h = h; t.fmt = fmt; t.linear = linear; t.repeat = repeat; return t;
}
void store_rgba32f(float* dst, int W, int x, int y, vec4 v) { float* p = dst + 4 * ((size_t)y * W + x); p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
void store_rgba16f(uint16_t* dst, int W, int x, int y, vec4 v) {
  uint16_t* p = dst + 4 * ((size_t)y * W + x);
  p[0] = float_to_half(v.x); p[1] = float_to_half(v.y); p[2] = float_to_half(v.z); p[3] = float_to_half(v.w);
}

}  // namespace

// ======================================================================================
// extern "C" surface used by tests / smoke / bench cpu_baseline (ctypes)
// ======================================================================================
extern "C" {

struct orc_env {
  const uint16_t* mips[16];  // RGBA16F per level (level 0 = w x h), built by orc_env_build_mips
  int32_t mip_w[16], mip_h[16];
  int32_t levels;
  int32_t width, height;
  const float* marginal;     // `height` floats
  const float* conditional;  // width*height floats
  float total_sum_whole, total_sum_decimal;
};

// number of mip levels of a w x h texture
int orc_mip_levels(int w, int h) { int l = 1; while (w > 1 || h > 1) { w = gmax(1, w >> 1); h = gmax(1, h >> 1); l++; } return l; }

// Box-filter mip level: dst(w1 x h1) from src(w0 x h0), RGBA16F, average of the 2x2 (or 2x1 / 1x2) footprint in fp32,
// stored RNE.  (GL generateMipmap is implementation-defined; the box filter is the universal choice for power-of-two sizes.)
void orc_env_downsample(const uint16_t* src, int w0, int h0, uint16_t* dst) {
  int w1 = gmax(1, w0 >> 1), h1 = gmax(1, h0 >> 1);
  for (int y = 0; y < h1; y++)
    for (int x = 0; x < w1; x++) {
      int x0 = gmin(2 * x, w0 - 1), x1 = gmin(2 * x + 1, w0 - 1), y0 = gmin(2 * y, h0 - 1), y1 = gmin(2 * y + 1, h0 - 1);
      for (int c = 0; c < 4; c++) {
        float a = half_to_float(src[4 * ((size_t)y0 * w0 + x0) + c]), b = half_to_float(src[4 * ((size_t)y0 * w0 + x1) + c]);
        float d = half_to_float(src[4 * ((size_t)y1 * w0 + x0) + c]), e = half_to_float(src[4 * ((size_t)y1 * w0 + x1) + c]);
        dst[4 * ((size_t)y * w1 + x) + c] = float_to_half(((a + b) + (d + e)) * 0.25f);
      }
    }
}

static void fill_env(EnvInfo& e, const orc_env* env) {
  if (!env) return;
  e.map.levels = env->levels;
  for (int l = 0; l < env->levels; l++) e.map.level[l] = mk(env->mips[l], env->mip_w[l], env->mip_h[l], F_RGBA16F, true, false);
  e.marginalWeights = mk(env->marginal, env->height, 1, F_R32F);
  e.conditionalWeights = mk(env->conditional, env->width, env->height, F_R32F);
  e.size = vec2((float)env->width, (float)env->height);
  e.totalSumWhole = env->total_sum_whole;
  e.totalSumDecimal = env->total_sum_decimal;
}

// K1.  depth R32F, gbuffer RGBA32F, velocity RGBA32F|NULL, direct_light RGBA16F|NULL, accumulated RGBA32F|NULL, out RGBA32F
// W x H: the render target (= `resolution`); TW x TH: the size of the input planes — larger than W x H when resolutionScale < 1
// (SSGIPass.js:52-57: only the SSGI target is scaled, the G-buffer / depth / denoiser textures stay at full size)
void orc_ssgi_trace(const rfx_ssgi_params* p, int W, int H, const float* depth, const float* gbuffer, const float* velocity,
                    const uint16_t* direct_light, const float* accumulated, const orc_env* env, const uint8_t* blue_noise, int bn_w, int bn_h,
                    float* out, int TW, int TH) {
#pragma omp parallel
  {
    SsgiShader s(*p);
    s.depthTexture = mk(depth, TW, TH, F_R32F);
    s.gBufferTexture = mk(gbuffer, TW, TH, F_RGBA32F);
    s.velocityTexture = mk(velocity, TW, TH, F_RGBA32F);
    s.directLightTexture = mk(direct_light, TW, TH, F_RGBA16F, true);
    s.accumulatedTexture = mk(accumulated, TW, TH, F_RGBA32F);
    s.bn.tex = mk(blue_noise, bn_w, bn_h, F_RGBA8, false, true);
    s.resolution = vec2((float)W, (float)H);
    fill_env(s.envMapInfo, env);
#pragma omp for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) store_rgba32f(out, W, x, y, s.mainPx(x, y));
  }
}

// K2.  input RGBA32F (diffuseSpecular / specular) or RGBA16F (diffuse); history RGBA16F; out fp32 (out_half=0) or fp16 (out_half=1).
// Discarded pixels are left untouched in out0/out1.
// history_float: the history planes are RGBA32F (denoiseMode "full_temporal" / "temporal": the FramebufferTexture copy of the FloatType target)
void orc_temporal_reproject(const rfx_temporal_params* p, int W, int H, const void* input, int input_half, const float* velocity,
                            const void* history0, const void* history1, void* out0, void* out1, int out_half, int history_float, int IW, int IH) {
  // IW x IH: the size of inputTexture (the SSGI target: smaller than W x H when resolutionScale < 1; NEAREST, fetched by uv)
#pragma omp parallel
  {
    TemporalShader s(*p);
    s.W = W; s.H = H;
    s.inputTexture = mk(input, IW, IH, input_half ? F_RGBA16F : F_RGBA32F, input_half != 0);
    s.velocityTexture = mk(velocity, W, H, F_RGBA32F);
    s.accumulatedTexture[0] = mk(history0, W, H, history_float ? F_RGBA32F : F_RGBA16F, p->history_linear != 0);
    s.accumulatedTexture[1] = mk(history1, W, H, history_float ? F_RGBA32F : F_RGBA16F, p->history_linear != 0);
    s.invTexSize = vec2((float)(1.0 / W), (float)(1.0 / H));  // TemporalReprojectPass.js:135 (JS doubles -> float32 uniform)
    void* outs[2] = {out0, out1};
#pragma omp for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        vec4 o[2];
        if (!s.mainPx(x, y, o)) continue;
        for (int i = 0; i < p->texture_count; i++) {
          if (out_half) store_rgba16f((uint16_t*)outs[i], W, x, y, o[i]); else store_rgba32f((float*)outs[i], W, x, y, o[i]);
        }
      }
  }
}

// K3.  in0/in1: RGBA32F (in_half=0) or RGBA16F (in_half=1); out RGBA16F; discarded pixels untouched.
void orc_poisson_denoise(const rfx_poisson_params* p, int W, int H, const float* depth, const float* gbuffer_or_normal, const void* in0,
                         const void* in1, int in_half, const uint8_t* blue_noise, int bn_w, int bn_h, uint16_t* out0, uint16_t* out1) {
  PoissonShader s;
  s.W = W; s.H = H;
  s.depthTexture = mk(depth, W, H, F_R32F);
  s.GBUFFER_TEXTURE = p->gbuffer_texture != 0;
  if (s.GBUFFER_TEXTURE) s.gBufferTexture = mk(gbuffer_or_normal, W, H, F_RGBA32F); else s.normalTexture = mk(gbuffer_or_normal, W, H, F_RGBA32F);
  int fmt = in_half ? F_RGBA16F : F_RGBA32F;
  s.inputTexture = mk(in0, W, H, fmt, p->input_linear != 0);
  s.inputTexture2 = mk(in1, W, H, fmt, p->input_linear != 0);
  s.bn.tex = mk(blue_noise, bn_w, bn_h, F_RGBA8, false, true);
  s.radius = p->radius; s.phi = p->phi; s.lumaPhi = p->luma_phi; s.depthPhi = p->depth_phi; s.normalPhi = p->normal_phi;
  s.roughnessPhi = p->roughness_phi; s.specularPhi = p->specular_phi;
  s.resolution = vec2((float)W, (float)H);
  s.textureCount = p->texture_count; s.blueNoiseIndex = p->blue_noise_index;
  s.isTextureSpecular[0] = p->is_texture_specular[0]; s.isTextureSpecular[1] = p->is_texture_specular[1];
  uint16_t* outs[2] = {out0, out1};
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec4 o[2];
      if (!s.mainPx(x, y, o)) continue;
      for (int i = 0; i < p->texture_count; i++) store_rgba16f(outs[i], W, x, y, o[i]);
    }
}

// K4.  diffuse/specular RGBA16F (nearest fetch at pixel centre; filter irrelevant), out RGBA32F; discarded pixels untouched.
// diffuse_gi / specular_gi / scene may be NULL (null sampler => (0,0,0,1)): DenoiserComposePass.js:23-33 binds only the textures of its
// inputType; `scene` is the composer input buffer (Denoiser.js:100-102, LINEAR RGBA16F), read by TYPE_SPECULAR only
// gi_float: diffuse_gi / specular_gi are RGBA32F NEAREST (denoiseMode "full_temporal": the temporal pass's FloatType targets)
void orc_gi_compose(const rfx_compose_params* p, int W, int H, const float* depth, const float* gbuffer, const void* diffuse_gi,
                    const void* specular_gi, const uint16_t* scene, float* out, int gi_float) {
  ComposeShader s(*p);
  s.W = W; s.H = H;
  s.depthTexture = mk(depth, W, H, F_R32F);
  s.gBufferTexture = mk(gbuffer, W, H, F_RGBA32F);
  s.diffuseGiTexture = mk(diffuse_gi, W, H, gi_float ? F_RGBA32F : F_RGBA16F, !gi_float);
  s.specularGiTexture = mk(specular_gi, W, H, gi_float ? F_RGBA32F : F_RGBA16F, !gi_float);
  s.sceneTexture = mk(scene, W, H, F_RGBA16F, true);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec4 o;
      if (s.mainPx(x, y, o)) store_rgba32f(out, W, x, y, o);
    }
}

// K5.  ssgi_compose.frag:20-44 (+ three.js <fog_fragment> on vFogDepth = -getViewZ(depth) * 0.4, SSGIEffect.js:34-43).  gi RGBA32F, scene RGBA16F,
// out RGBA16F; p may be NULL (no fog, no debug)
void orc_ssgi_compose(const rfx_ssgi_compose_params* p, int W, int H, const float* depth, const float* gi, const uint16_t* scene, uint16_t* out) {
  Tex d = mk(depth, W, H, F_R32F), g = mk(gi, W, H, F_RGBA32F), sc = mk(scene, W, H, F_RGBA16F, true);
#pragma omp parallel for
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec2 uv = pixelUv(x, y, W, H);
      if (p && p->is_debug) { store_rgba16f(out, W, x, y, textureLod0(g, uv)); continue; }
      float depthv = textureLod0(d, uv).x;
      vec3 c;
      if (depthv == 1.0f) {
        c = textureLod0(sc, uv).xyz();
      } else {
        c = textureLod0(g, uv).xyz();
        if (p && p->use_fog) {
          float gz = p->perspective ? perspectiveDepthToViewZ(depthv, p->camera_near, p->camera_far) : orthographicDepthToViewZ(depthv, p->camera_near, p->camera_far);
          float viewZ = gz * 0.4f;
          float vFogDepth = -viewZ;
          float fogFactor = p->fog_exp2 ? 1.0f - expcr(-p->fog_density * p->fog_density * vFogDepth * vFogDepth) : smoothstepf(p->fog_near, p->fog_far, vFogDepth);
          c = mix(c, vec3(p->fog_color[0], p->fog_color[1], p->fog_color[2]), fogFactor);
        }
      }
      store_rgba16f(out, W, x, y, vec4(c, 1.0f));
    }
}

// K6.  out RGBA16F; discarded pixels untouched
void orc_hbao(const rfx_hbao_params* p, int W, int H, const float* depth, const uint8_t* blue_noise, int bn_w, int bn_h, uint16_t* out) {
  HbaoShader s;
  s.W = W; s.H = H;
  s.projectionViewMatrix = load_mat4(p->projection_view);
  s.projectionMatrixInverse = load_mat4(p->projection_inverse);
  s.cameraMatrixWorld = load_mat4(p->camera_matrix_world);
  s.depthTexture = mk(depth, W, H, F_R32F);
  s.bn.tex = mk(blue_noise, bn_w, bn_h, F_RGBA8, false, true);
  s.resolution = vec2((float)W, (float)H);
  s.aoDistance = p->ao_distance; s.distancePower = p->distance_power; s.bias = p->bias; s.thickness = p->thickness;
  s.spp = p->spp; s.blueNoiseIndex = p->blue_noise_index;
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec4 o;
      if (s.mainPx(x, y, o)) store_rgba16f(out, W, x, y, o);
    }
}

// K7.  ao_compose.frag:6-16.  ao RGBA16F linear (pixel-centre fetch), input/out RGBA16F
void orc_ao_compose(const rfx_ao_compose_params* p, int W, int H, const float* depth, const uint16_t* ao, const uint16_t* input, uint16_t* out) {
  Tex d = mk(depth, W, H, F_R32F), a = mk(ao, W, H, F_RGBA16F, true), in = mk(input, W, H, F_RGBA16F, true);
  vec3 color(p->color[0], p->color[1], p->color[2]);
#pragma omp parallel for
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec2 uv = pixelUv(x, y, W, H);
      float unpackedDepth = textureLod0(d, uv).x;
      float aov = unpackedDepth > 0.9999f ? 1.0f : textureLod0(a, uv).w;
      aov = powcr(aov, p->power);
      vec3 aoColor = mix(color, vec3(1.0f), aov);
      vec4 inputColor = textureLod0(in, uv);
      aoColor *= inputColor.xyz();
      store_rgba16f(out, W, x, y, vec4(aoColor, inputColor.w));
    }
}

// K8.  motion_blur.frag:11-44.  velocity RGBA32F nearest, input RGBA16F linear, out RGBA16F
void orc_motion_blur(const rfx_motion_blur_params* p, int W, int H, const float* velocity, const uint16_t* input, const uint8_t* blue_noise,
                     int bn_w, int bn_h, uint16_t* out) {
  Tex vel = mk(velocity, W, H, F_RGBA32F), in = mk(input, W, H, F_RGBA16F, true);
  BlueNoise bn; bn.tex = mk(blue_noise, bn_w, bn_h, F_RGBA8, false, true);
  vec2 resolution(p->resolution[0], p->resolution[1]);
  float samplesFloat = (float)p->samples;
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec2 vUv = pixelUv(x, y, W, H);
      vec4 inputColor = textureLod0(in, vUv);
      vec4 vt = textureLod0(vel, vUv);
      vec2 velocityv(vt.x, vt.y);
      bool didMove = dot(velocityv, velocityv) > 0.000000001f;
      if (!didMove) { store_rgba16f(out, W, x, y, inputColor); continue; }
      velocityv = velocityv * p->intensity;
      vec4 blueNoise = bn.sample(vUv, resolution, p->frame);
      vec2 jitterOffset = p->jitter * velocityv * vec2(blueNoise.x, blueNoise.y);
      float frameSpeed = (1.0f / 100.0f) / p->delta_time;
      vec2 startUv = vUv + (jitterOffset - velocityv * 0.5f) * frameSpeed;
      vec2 endUv = vUv + (jitterOffset + velocityv * 0.5f) * frameSpeed;
      startUv = vec2(gmax(0.0f, startUv.x), gmax(0.0f, startUv.y));
      endUv = vec2(gmin(1.0f, endUv.x), gmin(1.0f, endUv.y));
      vec3 motionBlurredColor = inputColor.xyz();
      for (float i = 0.0f; i <= samplesFloat; i++) {
        vec2 reprojectedUv = mix(startUv, endUv, i / samplesFloat);
        motionBlurredColor += textureLod0(in, reprojectedUv).xyz();
      }
      motionBlurredColor /= samplesFloat + 2.0f;
      store_rgba16f(out, W, x, y, vec4(motionBlurredColor, inputColor.w));
    }
}

// K9.  traa_compose.frag:3-6
void orc_traa_compose(int W, int H, const uint16_t* accumulated, uint16_t* out) {
  Tex a = mk(accumulated, W, H, F_RGBA16F, true);
#pragma omp parallel for
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) store_rgba16f(out, W, x, y, vec4(textureLod0(a, pixelUv(x, y, W, H)).xyz(), 1.0f));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Cosmetic effects, merged like postprocessing's EffectPass: one input buffer, colour chained.
// ---------------------------------------------------------------------------------------------------------------------
namespace fx {
const float PI_COMMON = 3.141592653589793f;  // three <common>
// three <common> rand()
inline float rand2(vec2 uv) {
  const float a = 12.9898f, b = 78.233f, c = 43758.5453f;
  float dt = dot(uv, vec2(a, b)), sn = modf_gl(dt, PI_COMMON);
  float v = sincr(sn) * c;
  return v - std::floor(v);
}
inline vec4 add4(vec4 a, vec4 b) { return a + b; }
// SharpnessEffect.js:8-29
inline vec4 sharpness(const Tex& inputTexture, vec4 inputColor, vec2 uv, vec2 texelSize, float sharp) {
  auto T = [&](vec2 o) { return textureLod0(inputTexture, uv + o * texelSize); };
  vec4 blurredPixel = textureLod0(inputTexture, uv - 1.0f * texelSize);
  blurredPixel += T(vec2(0.0f, -1.0f));
  blurredPixel += T(vec2(1.0f, -1.0f));
  blurredPixel += T(vec2(-1.0f, 0.0f));
  blurredPixel += inputColor;
  blurredPixel += T(vec2(1.0f, 0.0f));
  blurredPixel += T(vec2(-1.0f, 1.0f));
  blurredPixel += T(vec2(0.0f, 1.0f));
  blurredPixel += textureLod0(inputTexture, uv + 1.0f * texelSize);
  blurredPixel = blurredPixel / 9.0f;
  vec4 sharpDiff = inputColor - blurredPixel;
  vec4 sharpenedPixel = inputColor + sharpDiff * sharp;
  return vec4(gmax(sharpenedPixel.x, 0.0f), gmax(sharpenedPixel.y, 0.0f), gmax(sharpenedPixel.z, 0.0f), sharpenedPixel.w);
}
// LensDistortionEffect.js:14-45
inline vec4 lens_distortion(const Tex& inputTexture, vec2 vUv, vec2 resolution, float alphax, float alphay, float aberration) {
  float x = (2.0f * vUv.x - 1.0f) / 1.0f;
  float y = (2.0f * vUv.y - 1.0f) / 1.0f;
  float r = x * x + y * y;
  float x3 = x / (1.0f - alphax * r);
  float y3 = y / (1.0f - alphay * r);
  float x2 = x / (1.0f - alphax * (x3 * x3 + y3 * y3));
  float y2 = y / (1.0f - alphay * (x3 * x3 + y3 * y3));
  float i2 = (x2 + 1.0f) * 1.0f / 2.0f;
  float j2 = (y2 + 1.0f) * 1.0f / 2.0f;
  vec2 duv = vec2(i2, j2);
  vec2 rOffset = vec2(1.0f / resolution.x, 0.0f);
  vec2 gOffset = vec2(0.0f, 1.0f / resolution.y);
  vec2 bOffset = vec2(1.0f / resolution.x, 1.0f / resolution.y);
  vec4 rValue = textureLod0(inputTexture, duv - aberration * rOffset);
  vec4 gValue = textureLod0(inputTexture, duv - aberration * gOffset);
  vec4 bValue = textureLod0(inputTexture, duv - aberration * bOffset);
  return vec4(rValue.x, gValue.y, bValue.z, 1.0f);
}
// getViewPosition of GradualBackgroundEffect.js:22-29 / SparkleEffect.js:29-36
inline vec3 view_position(const Camera& cam, vec2 vUv, float viewZ) {
  float clipW = cam.projectionMatrix.at(2, 3) * viewZ + cam.projectionMatrix.at(3, 3);
  vec4 clipPosition = vec4((vec3(vUv, viewZ) - 0.5f) * 2.0f, 1.0f);
  clipPosition = clipPosition * clipW;
  vec3 p = (cam.projectionMatrixInverse * clipPosition).xyz();
  p.z = viewZ;
  return p;
}
// GradualBackgroundEffect.js:31-46
inline vec4 gradual_background(const Camera& cam, const Tex& depthTexture, vec4 inputColor, vec2 uv, vec3 backgroundColor, float maxDistance) {
  float depth = textureLod0(depthTexture, uv).x;
  float viewZ = cam.perspective ? perspectiveDepthToViewZ(depth, cam.cameraNear, cam.cameraFar) : orthographicDepthToViewZ(depth, cam.cameraNear, cam.cameraFar);
  vec3 viewPos = view_position(cam, uv, viewZ);
  vec3 worldPos = (cam.cameraMatrixWorld * vec4(viewPos, 1.0f)).xyz();
  float distToCenter = length(vec2(worldPos.x, worldPos.z)) + gmax(0.0f, -worldPos.y);
  float fade = clampf(powcr(distToCenter, 0.1f) * 15.0f - maxDistance, 0.0f, 1.0f);
  vec3 color = mix(inputColor.xyz(), backgroundColor, fade);
  return vec4(color, 1.0f);
}
// SparkleEffect.js:38-43
inline float nn(vec2 n) {
  vec2 b = vec2(std::floor(n.x), std::floor(n.y));
  vec2 fr = vec2(n.x - std::floor(n.x), n.y - std::floor(n.y));
  vec2 f = vec2(smoothstepf(0.0f, 1.0f, fr.x), smoothstepf(0.0f, 1.0f, fr.y));
  return mixf(mixf(rand2(b), rand2(b + vec2(1.0f, 0.0f)), f.x), mixf(rand2(b + vec2(0.0f, 1.0f)), rand2(b + vec2(1.0f, 1.0f)), f.x), f.y);
}
// SparkleEffect.js:45-99
inline vec4 sparkle(const Camera& cam, bool perspective_defined, const Tex& velocityTexture, vec4 inputColor, vec2 uv, float spread, float intensity) {
  vec4 velocityTexel = textureLod0(velocityTexture, uv);
  float depth = velocityTexel.w;
  if (depth == 0.0f || depth == 1.0f) return inputColor;
  vec3 normal = unpackNormal(velocityTexel.z);
  vec3 viewNormal = normalize((cam.viewMatrix * vec4(normal, 0.0f)).xyz());
  float viewZ = perspective_defined ? perspectiveDepthToViewZ(depth, cam.cameraNear, cam.cameraFar) : orthographicDepthToViewZ(depth, cam.cameraNear, cam.cameraFar);
  vec3 viewPos = view_position(cam, uv, viewZ);
  vec3 viewDir = normalize(viewPos);
  vec3 worldPos = (cam.cameraMatrixWorld * vec4(viewPos, 1.0f)).xyz();
  if (worldPos.y < 0.01f) return inputColor;
  vec3 cameraPos = (cam.cameraMatrixWorld * vec4(0.0f, 0.0f, 0.0f, 1.0f)).xyz();
  float dist = length(worldPos - cameraPos);
  float distFactor = expcr(-dist * 0.005f);
  float facing = gmax(dot(-viewDir, viewNormal), 0.0f);
  facing = powcr(facing, 4.0f);
  vec3 nw = normalize(worldPos);
  vec2 offset = vec2(nw.x, nw.z) * 1000.0f + vec2(normal.x, normal.z) * 500.0f;
  float noise = nn(offset);
  noise = powcr(noise, 500.0f * spread);
  float lum = dot(inputColor.xyz(), vec3(0.299f, 0.587f, 0.114f));
  lum = smoothstepf(0.15f, 1.0f, lum);
  float sparkleFactor = noise * lum * facing * distFactor * 5000.0f * intensity;
  vec3 c = inputColor.xyz();
  vec3 color = c + vec3(powcr(c.x, 4.0f), powcr(c.y, 4.0f), powcr(c.z, 4.0f)) * sparkleFactor;
  return vec4(color, 1.0f);
}
}  // namespace fx

extern "C" {
// input RGBA16F linear; depth R32F | NULL; velocity RGBA32F | NULL; out RGBA16F
void orc_effects(const rfx_effects_params* p, int W, int H, const uint16_t* input, const float* depth, const float* velocity, uint16_t* out) {
  Tex in = mk(input, W, H, F_RGBA16F, true), d = mk(depth, W, H, F_R32F), v = mk(velocity, W, H, F_RGBA32F);
  Camera cam(p->cam);
  vec2 resolution((float)W, (float)H), texelSize((float)(1.0 / W), (float)(1.0 / H));  // postprocessing: texelSize = 1 / size (JS doubles -> float32)
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec2 uv = pixelUv(x, y, W, H);
      vec4 c = textureLod0(in, uv);  // EffectPass: color0 = texture2D(inputBuffer, vUv)
      for (int e = 0; e < p->n_effects; e++) {
        switch (p->effects[e]) {
          case RFX_FX_SHARPNESS: c = fx::sharpness(in, c, uv, texelSize, p->sharpness); break;
          case RFX_FX_LENS_DISTORTION: c = fx::lens_distortion(in, uv, resolution, p->alphax, p->alphay, p->aberration); break;
          case RFX_FX_GRADUAL_BACKGROUND:
            c = fx::gradual_background(cam, d, c, uv, vec3(p->background_color[0], p->background_color[1], p->background_color[2]), p->max_distance); break;
          case RFX_FX_SPARKLE: c = fx::sparkle(cam, p->sparkle_perspective != 0, v, c, uv, p->spread, p->intensity); break;
          default: break;
        }
      }
      store_rgba16f(out, W, x, y, c);
    }
}
// TAAPass: input RGBA16F, history RGBA8, out RGBA8
void orc_taa(const rfx_taa_params* p, int W, int H, const uint16_t* input, const uint8_t* history, uint8_t* out) {
  Tex in = mk(input, W, H, F_RGBA16F, true), acc = mk(history, W, H, F_RGBA8);
#pragma omp parallel for
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      vec2 uv = pixelUv(x, y, W, H);
      vec4 color = textureLod0(in, uv);
      if (p->srgb_output) {  // three LinearTosRGB
        for (int i = 0; i < 3; i++) {
          float v = color[i];
          float hi = powcr(v, 0.41666f) * 1.055f - 0.055f, lo = v * 12.92f;
          color[i] = mixf(hi, lo, v <= 0.0031308f ? 1.0f : 0.0f);
        }
      }
      vec4 o = color;
      if (!(p->camera_not_moved_frames == 0.0f)) {
        vec4 a = textureLod0(acc, uv);
        float t = 1.0f / (p->camera_not_moved_frames + 1.0f);
        for (int i = 0; i < 4; i++) o[i] = mixf(a[i], color[i], t);
      }
      uint8_t* q = out + 4 * ((size_t)y * W + x);
      for (int i = 0; i < 4; i++) q[i] = (uint8_t)std::lround(clampf(o[i], 0.0f, 1.0f) * 255.0f);
    }
}
}  // extern "C"

extern "C" {
// G-buffer ingest (include/rfx.h rfx_gbuffer_ingest_launch): packGBuffer :166-178 + the velocity layout of
// VelocityDepthNormalMaterial.js:76-83,186-188 over SoA planes.  fmt_*: gl::Fmt of the plane (albedo / material RGBA8 | RGBA16F,
// normal / motion RGBA16F | RGBA32F).  emissive / motion may be NULL.  out_* may be NULL.
void orc_gbuffer_ingest(const rfx_ingest_params* p, int W, int H, const void* albedo, int fmt_albedo, const void* normal, int fmt_normal,
                        const void* material, int fmt_material, const uint16_t* emissive, const void* motion, int fmt_motion,
                        const float* depth, float* out_gbuffer, float* out_velocity) {
  Tex ta = mk(albedo, W, H, fmt_albedo), tn = mk(normal, W, H, fmt_normal), tm = mk(material, W, H, fmt_material);
  Tex te = mk(emissive, W, H, F_RGBA16F), tv = mk(motion, W, H, fmt_motion);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      float d = depth[(size_t)y * W + x];
      if (d == 1.0f) {
        if (out_gbuffer) store_rgba32f(out_gbuffer, W, x, y, vec4(0, 0, 0, 1));
        if (out_velocity) store_rgba32f(out_velocity, W, x, y, vec4(0, 0, 0, 1));
        continue;
      }
      vec4 nt = texel(tn, x, y);
      vec3 n = nt.xyz();
      if (p->normalize_normals) n = normalize(n);
      if (out_gbuffer) {
        vec4 diffuse = texel(ta, x, y), mt = texel(tm, x, y);
        vec3 em = emissive ? texel(te, x, y).xyz() : vec3(0.0f);
        vec4 g;
        g.x = vec4ToFloat(diffuse);
        g.y = packNormal(n);
        g.z = color2float(vec3(mt.x, mt.y, 0.0f));
        float mx = gmax(gmax(em.x, em.y), em.z);
        g.w = mx > 0.0f ? vec4ToFloat(encodeRGBE8(em)) : 0.0f;  // zero radiance: the shader's log2(0) path decodes to 0 everywhere
        store_rgba32f(out_gbuffer, W, x, y, g);
      }
      if (out_velocity) {
        vec4 mv = motion ? texel(tv, x, y) : vec4(0, 0, 0, 0);
        store_rgba32f(out_velocity, W, x, y, vec4(mv.x * p->motion_scale[0], mv.y * p->motion_scale[1], packNormal(n), d));
      }
    }
}

// ---- leaf functions exported for unit tests -------------------------------------------
uint16_t orc_float_to_half(float f) { return float_to_half(f); }
float orc_half_to_float(uint16_t h) { return half_to_float(h); }
void orc_pack_gbuffer(const float* diffuse4, const float* normal3, float roughness, float metalness, const float* emissive3, float* out4) {
  vec4 g = packGBuffer(vec4(diffuse4[0], diffuse4[1], diffuse4[2], diffuse4[3]), vec3(normal3[0], normal3[1], normal3[2]), roughness, metalness,
                       vec3(emissive3[0], emissive3[1], emissive3[2]));
  out4[0] = g.x; out4[1] = g.y; out4[2] = g.z; out4[3] = g.w;
}
// out: diffuse4, normal3, roughness, metalness, emissive3 = 12 floats
void orc_unpack_gbuffer(const float* g4, float* out12) {
  Tex t = mk(g4, 1, 1, F_RGBA32F);
  Material m = getMaterial(t, vec2(0.5f, 0.5f));
  float o[12] = {m.diffuse.x, m.diffuse.y, m.diffuse.z, m.diffuse.w, m.normal.x, m.normal.y, m.normal.z, m.roughness, m.metalness,
                 m.emissive.x, m.emissive.y, m.emissive.z};
  std::memcpy(out12, o, sizeof(o));
}
float orc_pack_normal(const float* n3) { return packNormal(vec3(n3[0], n3[1], n3[2])); }
void orc_pack_two_vec4(const float* a4, const float* b4, float* out4) {
  vec4 r = packTwoVec4(vec4(a4[0], a4[1], a4[2], a4[3]), vec4(b4[0], b4[1], b4[2], b4[3]));
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
void orc_unpack_two_vec4(const float* e4, float* a4, float* b4) {
  vec4 a, b;
  unpackTwoVec4(vec4(e4[0], e4[1], e4[2], e4[3]), a, b);
  a4[0] = a.x; a4[1] = a.y; a4[2] = a.z; a4[3] = a.w; b4[0] = b.x; b4[1] = b.y; b4[2] = b.z; b4[3] = b.w;
}
// blue-noise texel coordinate (index != 0) for pixel (x,y)
void orc_blue_noise_coord(int x, int y, int index, int size, int* sx, int* sy) {
  BlueNoise bn;
  uint32_t ui = (uint32_t)index;
  uint32_t s1[4] = {ui, ui * 15843u, ui * 31u + 4566u, ui * 2345u + 58585u};
  bn.pcg4d(s1);
  *sx = (x + (int)(s1[0] % 0x0fffffffu)) % size;
  *sy = (y + (int)(s1[1] % 0x0fffffffu)) % size;
}
int orc_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp master
    n = omp_get_num_threads();
  }
#endif
  return n;
}

}  // extern "C"
