// k_temporal.cu — K2 temporal reprojection / accumulation (sm_100a).
//
// Replaces the fullscreen draw of reference src/temporal-reproject/TemporalReprojectPass.js:192-193
// (shader src/temporal-reproject/shader/temporal_reproject.frag:178-208 + reproject.frag).
// Used with 2 planes (SSGI: diffuse + specular, packed fp16x8 input) and with 1 plane (TRAA).
#include "rfx_kernels.h"

namespace rfx {

struct TState {  // the shader's invocation globals (reproject.frag:3-7)
  v2 vUv, velocity;
  v3 worldNormal, worldPos;
  float depth, curvature, viewAngle, rayLength, roughness, moveFactor;
};

RFX_D v3 screenSpaceToWorldSpace(v2 uv, float depth, const M4& world, const M4& projInv) {  // reproject.frag:21-28
  v4 ndc = mk4((uv.x - 0.5f) * 2.0f, (uv.y - 0.5f) * 2.0f, (depth - 0.5f) * 2.0f, 1.0f);
  v4 clip = mul(projInv, ndc);
  v4 view = mul(world, mk4(clip.x / clip.w, clip.y / clip.w, clip.z / clip.w, clip.w / clip.w));
  return xyz(view);
}

template <bool LOG>
RFX_D v3 transformColor(v3 c) { return LOG ? vlog1p_(c) : c; }
template <bool LOG>
RFX_D v3 undoColorTransform(v3 c) { return LOG ? vexpm1_(c) : c; }

RFX_D float getViewZ(const TemporalArgs& a, float d) {
  return a.cam.perspective ? perspectiveDepthToViewZ(d, a.cam.near_plane, a.cam.far_plane) : orthographicDepthToViewZ(d, a.cam.near_plane, a.cam.far_plane);
}

// validateReprojectedUV  reproject.frag:130-167
RFX_D float validateReprojectedUV(const TemporalArgs& a, const TState& s, v2 ruv) {
  if (ruv.x > 1.0f || ruv.x < 0.0f || ruv.y > 1.0f || ruv.y < 0.0f) return 0.0f;
  const float4 t = tex_f4_nearest(a.velocity, ruv);
  const v3 lastWorldNormal = unpackNormal(t.z);
  const float lastDepth = t.w;
  const v3 lastWorldPos = screenSpaceToWorldSpace(ruv, lastDepth, a.prev_world, a.prev_proj_inv);
  // (lastViewAngle / angleMix are computed by the shader but never used)
  const float viewZ = fabsf(getViewZ(a, s.depth));
  const float distFactor = 1.0f + 1.0f / (viewZ + 1.0f);
  const v3 dpos = s.worldPos - lastWorldPos;
  float disoccl = 0.0f;
  disoccl += length(dpos) / 10.0f * distFactor;
  disoccl += fabsf(dot(dpos, s.worldNormal)) / 20.0f * distFactor;
  disoccl += fminf(1.0f - dot(s.worldNormal, lastWorldNormal), 1.0f) / 1.0f * distFactor;
  float confidence = 1.0f - fminf(disoccl, 1.0f);
  confidence = fmaxf(confidence, 0.0f);
  return powf(confidence, a.confidence_power);
}

// reprojectHitPoint  reproject.frag:169-193
RFX_D v2 reprojectHitPoint(const TemporalArgs& a, const TState& s) {
  if (s.curvature > 0.05f || s.rayLength < 0.01f) return mk2(-1.0f, -1.0f);
  const v3 cameraPos = mk3(a.camera_pos[0], a.camera_pos[1], a.camera_pos[2]);
  const v3 cameraRay = normalize(s.worldPos - cameraPos);
  const v3 hit = cameraPos + cameraRay * s.rayLength;
  const v4 rh = mul(a.prev_proj_view, mk4(hit, 1.0f));
  return mk2(rh.x / rh.w * 0.5f + 0.5f, rh.y / rh.w * 0.5f + 0.5f);
}

template <bool HLIN>
RFX_D v4 fetch_hist(const PV& t, v2 uv) { return HLIN ? tex_h4_linear(t, uv) : tex_h4_nearest(t, uv); }

// BiCubicCatmullRom5Tap  reproject.frag:212-255
template <bool HLIN>
RFX_D v4 catmull5(const TemporalArgs& a, const PV& tex, v2 P) {
  const v2 inv = mk2(a.inv_w, a.inv_h);
  const v2 UV = P / inv;
  const v2 tc = mk2(floorf(UV.x - 0.5f) + 0.5f, floorf(UV.y - 0.5f) + 0.5f);
  const v2 f = UV - tc;
  const v2 f2 = f * f;
  const v2 f3 = f2 * f;
  const v2 w0 = f2 - 0.5f * (f3 + f);
  const v2 w1 = 1.5f * f3 - 2.5f * f2 + mk2(1.0f, 1.0f);
  const v2 w3 = 0.5f * (f3 - f2);
  const v2 w2 = mk2(1.0f, 1.0f) - w0 - w1 - w3;
  const v2 W0 = w0, W1 = w1 + w2, W2 = w3;
  const v2 S0 = (tc - mk2(1.0f, 1.0f)) * inv, S1 = (tc + w2 / W1) * inv, S2 = (tc + mk2(2.0f, 2.0f)) * inv;
  const float sw0 = W1.x * W0.y, sw1 = W0.x * W1.y, sw2 = W1.x * W1.y, sw3 = W2.x * W1.y, sw4 = W1.x * W2.y;
  const v4 Ct = fetch_hist<HLIN>(tex, mk2(S1.x, S0.y)), Cl = fetch_hist<HLIN>(tex, mk2(S0.x, S1.y)), Cc = fetch_hist<HLIN>(tex, mk2(S1.x, S1.y)),
           Cr = fetch_hist<HLIN>(tex, mk2(S2.x, S1.y)), Cb = fetch_hist<HLIN>(tex, mk2(S1.x, S2.y));
  const float wm = 1.0f / (sw0 + sw1 + sw2 + sw3 + sw4);
  v4 r;
  r.x = fmaxf(((((Ct.x * sw0 + Cl.x * sw1) + Cc.x * sw2) + Cr.x * sw3) + Cb.x * sw4) * wm, 0.0f);
  r.y = fmaxf(((((Ct.y * sw0 + Cl.y * sw1) + Cc.y * sw2) + Cr.y * sw3) + Cb.y * sw4) * wm, 0.0f);
  r.z = fmaxf(((((Ct.z * sw0 + Cl.z * sw1) + Cc.z * sw2) + Cr.z * sw3) + Cb.z * sw4) * wm, 0.0f);
  r.w = fmaxf(((((Ct.w * sw0 + Cl.w * sw1) + Cc.w * sw2) + Cr.w * sw3) + Cb.w * sw4) * wm, 0.0f);
  return r;
}

template <int TC, int ITYPE, bool LOG, bool HLIN>
__global__ void __launch_bounds__(kThreads) temporal_kernel(const __grid_constant__ TemporalArgs a) {
  int x, y;
  block_pixel(x, y, a.row0 & ~1);
  const bool active = x < a.W && y < a.H && y >= a.row0 && y < a.row1;
  const int xc = min(x, a.W - 1), yc = min(y, a.H - 1);
  TState s;
  s.vUv = pixel_uv(x, y, a.W, a.H);
  s.roughness = 1.0f; s.moveFactor = 0.0f; s.rayLength = 0.0f;

  // getVelocityNormalDepth  reproject.frag:97-105
  const float4 vt = ld_f4(a.velocity, xc, yc);
  s.velocity = mk2(vt.x, vt.y);
  s.worldNormal = unpackNormal(vt.z);
  s.depth = vt.w;
  const float fwd = fwidth_f(s.depth);
  s.curvature = length(fwidth_3(s.worldNormal));  // getCurvature :265-269
  if (!active) return;

  // getTexels + preprocessInput  temporal_reproject.frag:124-145
  v4 inp[2];
  bool sampled[2] = {false, false};
  if (ITYPE == RFX_INPUT_DIFFUSE_SPECULAR) {
    unpackTwoVec4(ld_f4(a.input, x, y), inp[0], inp[1]);
  } else if (a.input_half) {
    inp[0] = tex_h4_linear(a.input, s.vUv);  // composer buffer: LINEAR, fetched at the pixel centre
  } else {
    inp[0] = f4v(ld_f4(a.input, x, y));
  }
  constexpr int NIN = ITYPE == RFX_INPUT_DIFFUSE_SPECULAR ? 2 : 1;
#pragma unroll
  for (int i = 0; i < NIN; i++) {
    sampled[i] = inp[i].x >= 0.0f;
    v3 c = transformColor<LOG>(vmax(xyz(inp[i]), mk3(0.0f)));
    inp[i] = mk4(c, inp[i].w);
  }
  if (ITYPE != RFX_INPUT_DIFFUSE) {
    if (s.depth == 1.0f && fwd == 0.0f) return;  // discard :188-193
  }
  // computeGVariables :147-153
  s.worldPos = screenSpaceToWorldSpace(s.vUv, s.depth, a.cam.camera_matrix_world, a.cam.projection_inverse);
  {
    const v3 viewPos = xyz(mul(a.cam.view_matrix, mk4(s.worldPos, 1.0f)));
    const v3 viewDir = normalize(viewPos);
    const v3 viewNormal = mul_dir_left(s.worldNormal, a.cam.view_matrix);
    s.viewAngle = dot(-viewDir, viewNormal);
  }
  // getRoughnessRayLength :167-176
  if (ITYPE == RFX_INPUT_DIFFUSE_SPECULAR) {
    s.rayLength = inp[1].w;
    s.roughness = clampf(inp[0].w, 0.0f, 1.0f);
  } else if (ITYPE == RFX_INPUT_SPECULAR) {
    v2 d = unpackHalf2x16(__float_as_uint(inp[0].w));
    s.rayLength = d.x;
    s.roughness = clampf(d.y, 0.0f, 1.0f);
  }
  // computeReprojectedUv :155-165
  v3 ruvD, ruvS;
  {
    v2 r = s.vUv - s.velocity;
    ruvD = mk3(r.x, r.y, validateReprojectedUV(a, s, r));
    ruvS = mk3(-1.0f);
    if (ITYPE != RFX_INPUT_DIFFUSE) {
      v2 h = reprojectHitPoint(a, s);
      ruvS = mk3(h.x, h.y, validateReprojectedUV(a, s, h));
      if (ruvS.x == -1.0f) ruvS = ruvD;
    }
  }
  s.moveFactor = fminf(dot(s.velocity, s.velocity) * 10000.0f, 1.0f);

  const int rs[2] = {a.rs0, a.rs1};
#pragma unroll
  for (int i = 0; i < TC; i++) {
    const bool spec = rs[i] != 0;
    const v3 uvc = spec ? ruvS : ruvD;
    const PV& hist = i == 0 ? a.hist0 : a.hist1;
    // reproject()  temporal_reproject.frag:83-122
    v4 acc = catmull5<HLIN>(a, hist, mk2(uvc.x, uvc.y));
    v3 accRgb = transformColor<LOG>(xyz(acc));
    float accA = acc.w;
    v3 inRgb = xyz(inp[i]);
    if (!sampled[i]) {
      inRgb = accRgb;
    } else {
      accA += 1.0f;
      const int clampRadius = (spec && s.roughness < 0.25f) ? 1 : 2;
      // clampNeighborhood  reproject.frag:53-95
      v3 inLin = undoColorTransform<LOG>(inRgb);
      v3 mn = inLin, mx = inLin;
      for (int dx = -clampRadius; dx <= clampRadius; dx++) {
        for (int dy = -clampRadius; dy <= clampRadius; dy++) {
          const v2 nuv = mk2(s.vUv.x + (float)dx * a.inv_w, s.vUv.y + (float)dy * a.inv_h);
          v4 nt;
          if (ITYPE == RFX_INPUT_DIFFUSE_SPECULAR) {
            v4 t1, t2;
            unpackTwoVec4(tex_f4_nearest(a.input, nuv), t1, t2);
            nt = spec ? t2 : t1;
          } else if (a.input_half) {
            nt = tex_h4_linear(a.input, nuv);
          } else {
            nt = f4v(tex_f4_nearest(a.input, nuv));
          }
          if (nt.x >= 0.0f) {
            mn = vmin(xyz(nt), mn);
            mx = vmax(xyz(nt), mx);
          }
        }
      }
      mn = transformColor<LOG>(mn);
      mx = transformColor<LOG>(mx);
      const v3 clamped = mk3(clampf(accRgb.x, mn.x, mx.x), clampf(accRgb.y, mn.y, mx.y), clampf(accRgb.z, mn.z, mx.z));
      const float r = spec ? s.roughness : 1.0f;
      const float clampAggressiveness = fminf(1.0f, uvc.z * r);
      const float clampIntensity = mixf(0.0f, fminf(1.0f, s.moveFactor * 50.0f + a.clamp_intensity), clampAggressiveness);
      const v3 newColor = mix(accRgb, clamped, clampIntensity);
      const float colorDiff = fminf(length(newColor - accRgb), 1.0f);
      accA *= 1.0f - colorDiff;
      accRgb = newColor;
    }
    // accumulate()  temporal_reproject.frag:42-79
    const float confidence = powf(uvc.z, a.confidence_power);
    float accumBlend = 1.0f - 1.0f / (accA + 1.0f);
    accumBlend = mixf(0.0f, accumBlend, confidence);
    float maxValue = (a.full_accumulate ? 1.0f : a.max_blend) * a.keep_data;
    if (ITYPE != RFX_INPUT_DIFFUSE) {
      if (spec && s.roughness >= 0.0f && s.roughness < 0.1f) {
        const float maxRoughnessValue = mixf(0.0f, maxValue, s.roughness / 0.1f);
        maxValue = mixf(maxValue, maxRoughnessValue, fminf(100.0f * s.moveFactor, 1.0f));
      }
    }
    const float tmix = fminf(accumBlend, maxValue);
    float oa = 1.0f / (1.0f - tmix) - 1.0f;
    oa = fminf(65536.0f, oa);
    const v3 orgb = undoColorTransform<LOG>(mix(inRgb, accRgb, tmix));
    const OutV& o = i == 0 ? a.out0 : a.out1;
    if (a.out_half) st_h4(o.p, o.pitch, x, y, mk4(orgb, oa));
    else st_f4(o.p, o.pitch, x, y, make_float4(orgb.x, orgb.y, orgb.z, oa));
  }
}

cudaError_t launch_temporal(const TemporalArgs& a, cudaStream_t s) {
  const int rb = a.row0 & ~1;
  dim3 grid((a.W + kTileW - 1) / kTileW, (a.row1 - rb + kTileH - 1) / kTileH);
#define RFX_LT(TC, IT) \
  do { if (a.log_transform) { if (a.history_linear) temporal_kernel<TC, IT, true, true><<<grid, kThreads, 0, s>>>(a); \
                              else temporal_kernel<TC, IT, true, false><<<grid, kThreads, 0, s>>>(a); } \
       else { if (a.history_linear) temporal_kernel<TC, IT, false, true><<<grid, kThreads, 0, s>>>(a); \
              else temporal_kernel<TC, IT, false, false><<<grid, kThreads, 0, s>>>(a); } } while (0)
  if (a.input_type == RFX_INPUT_DIFFUSE_SPECULAR && a.texture_count == 2) RFX_LT(2, RFX_INPUT_DIFFUSE_SPECULAR);
  else if (a.input_type == RFX_INPUT_DIFFUSE && a.texture_count == 1) RFX_LT(1, RFX_INPUT_DIFFUSE);
  else if (a.input_type == RFX_INPUT_SPECULAR && a.texture_count == 1) RFX_LT(1, RFX_INPUT_SPECULAR);
  else return cudaErrorNotSupported;
#undef RFX_LT
  return cudaGetLastError();
}

}  // namespace rfx
