// k_chain.cu — the fast SSGI chain's kernels K2 / K3 / K4 on chain-internal plane formats (sm_100a).
//
// Same passes as k_temporal.cu / k_denoise.cu (reference: temporal_reproject.frag + reproject.frag, poisson_denoise.frag:126-208,
// DenoiserComposePass.js:58-85 + denoiser_compose_functions.glsl:53-107) with the same tap geometry, texel choices and per-pixel
// decisions (those stay in the IEEE arithmetic the oracle uses), restructured around what bounds them on B200 — instruction issue
// and L1 wavefronts, not HBM (profiles/):
//   * the two denoise planes travel as ONE 16-byte texel {diffuse rgba, specular rgba} in fp16 (one LDG.128 per bilinear corner
//     instead of two LDG.64, one address computation instead of two), the temporal targets as one 32-byte fp32 texel;
//   * the tap texel of the decoded G-buffer carries depth and roughness too (nrdz, one LDG.128 per tap instead of LDG.128 + LDG.32);
//   * the GI compose runs inside the last Poisson pass (its only neighbourhood is the pixel itself in the fast variant), which
//     removes a launch and a round trip of depth + G-buffer + both denoise planes;
//   * value arithmetic downstream of the decisions uses the SFU (rcp / rsqrt / lg2 / ex2 .approx).
// Row-sharded frames (PEER = true) read rows of last frame's history owned by another rank in place, through peer-mapped
// pointers (NVLink loads), instead of receiving replicated planes.
#include "rfx_kernels.h"

namespace rfx {

#define C_LN2 0.69314718055994530942f
#define C_LOG2E 1.44269504088896340736f

// ------------------------------------------------------------------------------------------------------------------
// decode prepass: gBuffer + depth -> nrdz
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cdecode_kernel(const CDecodeArgs a, int row0, int row1) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.W || y >= row1) return;
  const float4 g = ld_f4(a.gb, x, y);
  const v3 n = unpackNormal(g.y);
  const float k = mod_gl(g.z, 257.0f);  // float2color(...).r * 256 : the 9-bit roughness code (gbuffer_packing.glsl:24-34)
  st_f4(a.nrdz.p, a.nrdz.pitch, x, y, nrdz_pack(n, fminf(fmaxf(k, 0.0f), 256.0f), ld_r32f(a.depth, x, y)));
}
cudaError_t launch_cdecode(const CDecodeArgs& a, const RowSegs& segs, int halo, cudaStream_t s) {
  for (int k = 0; k < segs.n; k++) {
    int r0 = max(0, segs.r0[k] - halo), r1 = min(a.H, segs.r1[k] + halo);
    if (k > 0) r0 = max(r0, min(a.H, segs.r1[k - 1] + halo));
    if (r0 >= r1) continue;
    dim3 grid((a.W + 31) / 32, (r1 - r0 + 7) / 8);
    cdecode_kernel<<<grid, 256, 0, s>>>(a, r0, r1);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// K2 (SSGI form: 2 planes, packed K1 input, logTransform, LINEAR history)
// ------------------------------------------------------------------------------------------------------------------
struct CTState {
  v2 vUv, velocity;
  v3 worldNormal, worldPos;
  float depth, curvature, rayLength, roughness, moveFactor;
};

// reproject.frag:21-28; EXACT = the IEEE perspective divide (the result feeds texel choices), else SFU reciprocal (distances only)
template <bool EXACT>
RFX_D v3 c_screen_to_world(v2 uv, float depth, const M4& world, const M4& projInv) {
  const v4 ndc = mk4((uv.x - 0.5f) * 2.0f, (uv.y - 0.5f) * 2.0f, (depth - 0.5f) * 2.0f, 1.0f);
  const v4 clip = mul(projInv, ndc);
  const float r = EXACT ? 1.0f / clip.w : fx_rcp(clip.w);
  return xyz(mul(world, mk4(clip.x * r, clip.y * r, clip.z * r, clip.w * r)));
}
RFX_D v3 c_log1p(v3 c) { return mk3(fx_lg2(c.x + 1.0f) * C_LN2, fx_lg2(c.y + 1.0f) * C_LN2, fx_lg2(c.z + 1.0f) * C_LN2); }
RFX_D v3 c_expm1(v3 c) { return mk3(fx_ex2(c.x * C_LOG2E) - 1.0f, fx_ex2(c.y * C_LOG2E) - 1.0f, fx_ex2(c.z * C_LOG2E) - 1.0f); }
RFX_D float c_pow(float x, float p) { return fx_ex2(p * fx_lg2(x)); }
RFX_D v3 c_unpack_normal(float packed) {  // unpackNormal with an SFU normalise
  v2 f = unpackHalf2x16(__float_as_uint(packed));
  f = f * 2.0f - 1.0f;
  v3 n = mk3(f.x, f.y, 1.0f - fabsf(f.x) - fabsf(f.y));
  const float t = fmaxf(-n.z, 0.0f);
  n.x += n.x >= 0.0f ? -t : t;
  n.y += n.y >= 0.0f ? -t : t;
  return fx_normalize(n);
}

// validateReprojectedUV  reproject.frag:130-167
RFX_D float c_validate(const CTemporalArgs& a, const CTState& s, v2 ruv, float distFactor) {
  if (ruv.x > 1.0f || ruv.x < 0.0f || ruv.y > 1.0f || ruv.y < 0.0f) return 0.0f;
  const float4 t = tex_f4_nearest(a.velocity, ruv);
  const v3 lastWorldNormal = c_unpack_normal(t.z);
  const v3 lastWorldPos = c_screen_to_world<false>(ruv, t.w, a.prev_world, a.prev_proj_inv);
  const v3 dpos = s.worldPos - lastWorldPos;
  float disoccl = fx_length(dpos) * 0.1f * distFactor;
  disoccl += fabsf(dot(dpos, s.worldNormal)) * 0.05f * distFactor;
  disoccl += fminf(1.0f - dot(s.worldNormal, lastWorldNormal), 1.0f) * distFactor;
  const float confidence = fmaxf(1.0f - fminf(disoccl, 1.0f), 0.0f);
  return c_pow(confidence, a.confidence_power);
}

// reprojectHitPoint  reproject.frag:169-193 — IEEE throughout: the result selects history / velocity texels
RFX_D v2 c_reproject_hit(const CTemporalArgs& a, const CTState& s) {
  if (s.curvature > 0.05f || s.rayLength < 0.01f) return mk2(-1.0f, -1.0f);
  const v3 cameraPos = mk3(a.camera_pos[0], a.camera_pos[1], a.camera_pos[2]);
  const v3 cameraRay = normalize(s.worldPos - cameraPos);
  const v3 hit = cameraPos + cameraRay * s.rayLength;
  const v4 rh = mul(a.prev_proj_view, mk4(hit, 1.0f));
  const v2 q = mk2(rh.x, rh.y) / rh.w;
  return mk2(q.x * 0.5f + 0.5f, q.y * 0.5f + 0.5f);
}

// one LINEAR fetch of plane `PL` (0 diffuse, 1 specular) of the interleaved dn history
template <int PL, bool PEER>
RFX_D v4 c_hist_bilinear(const PeerPV& h, v2 uv) {
  const Bilin b = bilin_setup(uv, h.local.w, h.local.h);
  const unsigned char* r0 = PEER ? peer_row_base(h, b.y0) : h.local.p;
  const unsigned char* r1 = PEER ? peer_row_base(h, b.y1) : h.local.p;
  const unsigned o0 = (unsigned)b.y0 * (unsigned)h.local.pitch + 8u * PL, o1 = (unsigned)b.y1 * (unsigned)h.local.pitch + 8u * PL;
  const uint2 t00 = __ldg((const uint2*)(r0 + (o0 + (unsigned)b.x0 * 16u))), t10 = __ldg((const uint2*)(r0 + (o0 + (unsigned)b.x1 * 16u)));
  const uint2 t01 = __ldg((const uint2*)(r1 + (o1 + (unsigned)b.x0 * 16u))), t11 = __ldg((const uint2*)(r1 + (o1 + (unsigned)b.x1 * 16u)));
  return mk4(bilin_blend(b, h_lo(t00.x), h_lo(t10.x), h_lo(t01.x), h_lo(t11.x)), bilin_blend(b, h_hi(t00.x), h_hi(t10.x), h_hi(t01.x), h_hi(t11.x)),
             bilin_blend(b, h_lo(t00.y), h_lo(t10.y), h_lo(t01.y), h_lo(t11.y)), bilin_blend(b, h_hi(t00.y), h_hi(t10.y), h_hi(t01.y), h_hi(t11.y)));
}

// BiCubicCatmullRom5Tap  reproject.frag:212-255 (texel choice: IEEE; weights: SFU reciprocals)
template <int PL, bool PEER>
RFX_D v4 c_catmull5(const CTemporalArgs& a, v2 P) {
  const v2 inv = mk2(a.inv_w, a.inv_h);
  const v2 UV = mk2(P.x / inv.x, P.y / inv.y);
  const v2 tc = mk2(floorf(UV.x - 0.5f) + 0.5f, floorf(UV.y - 0.5f) + 0.5f);
  const v2 f = UV - tc;
  const v2 f2 = f * f;
  const v2 f3 = f2 * f;
  const v2 w0 = f2 - 0.5f * (f3 + f);
  const v2 w1 = 1.5f * f3 - 2.5f * f2 + mk2(1.0f, 1.0f);
  const v2 w3 = 0.5f * (f3 - f2);
  const v2 w2 = mk2(1.0f, 1.0f) - w0 - w1 - w3;
  const v2 W0 = w0, W1 = w1 + w2, W2 = w3;
  const v2 S0 = (tc - mk2(1.0f, 1.0f)) * inv, S1 = mk2((tc.x + w2.x * fx_rcp(W1.x)) * inv.x, (tc.y + w2.y * fx_rcp(W1.y)) * inv.y), S2 = (tc + mk2(2.0f, 2.0f)) * inv;
  const float sw0 = W1.x * W0.y, sw1 = W0.x * W1.y, sw2 = W1.x * W1.y, sw3 = W2.x * W1.y, sw4 = W1.x * W2.y;
  const v4 Ct = c_hist_bilinear<PL, PEER>(a.hist, mk2(S1.x, S0.y)), Cl = c_hist_bilinear<PL, PEER>(a.hist, mk2(S0.x, S1.y)),
           Cc = c_hist_bilinear<PL, PEER>(a.hist, mk2(S1.x, S1.y)), Cr = c_hist_bilinear<PL, PEER>(a.hist, mk2(S2.x, S1.y)),
           Cb = c_hist_bilinear<PL, PEER>(a.hist, mk2(S1.x, S2.y));
  const float wm = fx_rcp(sw0 + sw1 + sw2 + sw3 + sw4);
  v4 r;
  r.x = fmaxf(fma_(Cb.x, sw4, fma_(Cr.x, sw3, fma_(Cc.x, sw2, fma_(Cl.x, sw1, Ct.x * sw0)))) * wm, 0.0f);
  r.y = fmaxf(fma_(Cb.y, sw4, fma_(Cr.y, sw3, fma_(Cc.y, sw2, fma_(Cl.y, sw1, Ct.y * sw0)))) * wm, 0.0f);
  r.z = fmaxf(fma_(Cb.z, sw4, fma_(Cr.z, sw3, fma_(Cc.z, sw2, fma_(Cl.z, sw1, Ct.z * sw0)))) * wm, 0.0f);
  r.w = fmaxf(fma_(Cb.w, sw4, fma_(Cr.w, sw3, fma_(Cc.w, sw2, fma_(Cl.w, sw1, Ct.w * sw0)))) * wm, 0.0f);
  return r;
}

template <int PL, bool PEER>
RFX_D float4 c_temporal_plane(const CTemporalArgs& a, const CTState& s, v3 uvc, v4 inp, bool sampled, v3 mn, v3 mx) {
  constexpr bool spec = PL == 1;
  // reproject()  temporal_reproject.frag:83-122
  const v4 acc = c_catmull5<PL, PEER>(a, mk2(uvc.x, uvc.y));
  v3 accRgb = c_log1p(xyz(acc));
  float accA = acc.w;
  v3 inRgb = xyz(inp);
  if (!sampled) {
    inRgb = accRgb;
  } else {
    accA += 1.0f;
    const v3 lo = c_log1p(mn), hi = c_log1p(mx);
    const v3 clamped = mk3(clampf(accRgb.x, lo.x, hi.x), clampf(accRgb.y, lo.y, hi.y), clampf(accRgb.z, lo.z, hi.z));
    const float r = spec ? s.roughness : 1.0f;
    const float clampAggressiveness = fminf(1.0f, uvc.z * r);
    const float clampIntensity = fminf(1.0f, fma_(s.moveFactor, 50.0f, a.clamp_intensity)) * clampAggressiveness;  // mix(0., x, t) = x*t
    const v3 newColor = mix(accRgb, clamped, clampIntensity);
    const float colorDiff = fminf(fx_length(newColor - accRgb), 1.0f);
    accA *= 1.0f - colorDiff;
    accRgb = newColor;
  }
  // accumulate()  temporal_reproject.frag:42-79
  const float confidence = c_pow(uvc.z, a.confidence_power);
  const float accumBlend = (1.0f - fx_rcp(accA + 1.0f)) * confidence;
  float maxValue = (a.full_accumulate ? 1.0f : a.max_blend) * a.keep_data;
  if (spec && s.roughness >= 0.0f && s.roughness < 0.1f) {
    const float maxRoughnessValue = maxValue * (s.roughness * 10.0f);
    maxValue = mixf(maxValue, maxRoughnessValue, fminf(100.0f * s.moveFactor, 1.0f));
  }
  const float tmix = fminf(accumBlend, maxValue);
  const float oa = fminf(65536.0f, fx_rcp(1.0f - tmix) - 1.0f);
  const v3 orgb = c_expm1(mix(inRgb, accRgb, tmix));
  return make_float4(orgb.x, orgb.y, orgb.z, oa);
}

template <bool PEER>
__global__ void __launch_bounds__(kThreads, 4) ctemporal_kernel(const __grid_constant__ CTemporalArgs a) {
  int x, y;
  const bool in_rows = seg_pixel(a.segs, x, y);
  const bool active = x < a.W && y < a.H && in_rows;
  const int xc = min(x, a.W - 1), yc = min(y, a.H - 1);
  CTState s;
  s.vUv = pixel_uv(x, y, a.W, a.H);
  // getVelocityNormalDepth  reproject.frag:97-105
  const float4 vt = ld_f4(a.velocity, xc, yc);
  s.velocity = mk2(vt.x, vt.y);
  s.worldNormal = c_unpack_normal(vt.z);
  s.depth = vt.w;
  const float fwd = fwidth_f(s.depth);
  s.curvature = fx_length(fwidth_3(s.worldNormal));
  if (!active) return;
  if (s.depth == 1.0f && fwd == 0.0f) return;  // discard  temporal_reproject.frag:188-193 (the target keeps its texel)

  v4 inp[2];
  unpackTwoVec4(ld_f4(a.input, x, y), inp[0], inp[1]);
  bool sampled[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    sampled[i] = inp[i].x >= 0.0f;
    const v3 c = c_log1p(vmax(xyz(inp[i]), mk3(0.0f)));
    inp[i] = mk4(c, inp[i].w);
  }
  s.worldPos = c_screen_to_world<true>(s.vUv, s.depth, a.cam.camera_matrix_world, a.cam.projection_inverse);
  s.rayLength = inp[1].w;
  s.roughness = clampf(inp[0].w, 0.0f, 1.0f);
  v3 ruvD, ruvS;
  {
    const float gz = a.cam.perspective ? (a.cam.near_plane * a.cam.far_plane) * fx_rcp((a.cam.far_plane - a.cam.near_plane) * s.depth - a.cam.far_plane)
                                       : orthographicDepthToViewZ(s.depth, a.cam.near_plane, a.cam.far_plane);
    const float distFactor = 1.0f + fx_rcp(fabsf(gz) + 1.0f);
    const v2 r = s.vUv - s.velocity;
    ruvD = mk3(r.x, r.y, c_validate(a, s, r, distFactor));
    const v2 h = c_reproject_hit(a, s);
    ruvS = mk3(h.x, h.y, c_validate(a, s, h, distFactor));
    if (ruvS.x == -1.0f) ruvS = ruvD;
  }
  s.moveFactor = fminf(dot(s.velocity, s.velocity) * 10000.0f, 1.0f);

  // neighbourhood AABBs: one merged 5x5 sweep on the packed fp16 pairs (see k_temporal.cu)
  v3 mn[2], mx[2];
  int radius[2] = {2, s.roughness < 0.25f ? 1 : 2};
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const v3 inLin = c_expm1(xyz(inp[i]));
    mn[i] = inLin; mx[i] = inLin;
  }
  if (sampled[0] || sampled[1]) {
    const __half2 pinf = __floats2half2_rn(INFINITY, INFINITY), ninf = __floats2half2_rn(-INFINITY, -INFINITY);
    __half2 mn_rg[2] = {pinf, pinf}, mn_ba[2] = {pinf, pinf}, mx_rg[2] = {ninf, ninf}, mx_ba[2] = {ninf, ninf};
    for (int dy = -2; dy <= 2; dy++) {
      const int ty = clampi(y + dy, a.H);
#pragma unroll
      for (int dx = -2; dx <= 2; dx++) {
        const int tx = clampi(x + dx, a.W);
        const float4 e = ld_f4(a.input, tx, ty);
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const unsigned urg = __float_as_uint(i ? e.z : e.x), uba = __float_as_uint(i ? e.w : e.y);
          const __half2 rg = *reinterpret_cast<const __half2*>(&urg), ba = *reinterpret_cast<const __half2*>(&uba);
          const bool inside = abs(dx) <= radius[i] && abs(dy) <= radius[i];
          if (inside && (__low2float(rg) - RFX_NON_ZERO_OFFSET) >= 0.0f) {
            mn_rg[i] = __hmin2(mn_rg[i], rg); mx_rg[i] = __hmax2(mx_rg[i], rg);
            mn_ba[i] = __hmin2(mn_ba[i], ba); mx_ba[i] = __hmax2(mx_ba[i], ba);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const v3 lo = mk3(__low2float(mn_rg[i]) - RFX_NON_ZERO_OFFSET, __high2float(mn_rg[i]) - RFX_NON_ZERO_OFFSET, __low2float(mn_ba[i]) - RFX_NON_ZERO_OFFSET);
      const v3 hi = mk3(__low2float(mx_rg[i]) - RFX_NON_ZERO_OFFSET, __high2float(mx_rg[i]) - RFX_NON_ZERO_OFFSET, __low2float(mx_ba[i]) - RFX_NON_ZERO_OFFSET);
      mn[i] = vmin(lo, mn[i]);
      mx[i] = vmax(hi, mx[i]);
    }
  }
  float4* o = (float4*)(a.out.p + ((unsigned)y * (unsigned)a.out.pitch + (unsigned)x * 32u));
  o[0] = c_temporal_plane<0, PEER>(a, s, ruvD, inp[0], sampled[0], mn[0], mx[0]);
  o[1] = c_temporal_plane<1, PEER>(a, s, ruvS, inp[1], sampled[1], mn[1], mx[1]);
}

cudaError_t launch_ctemporal(const CTemporalArgs& a, cudaStream_t s) {
  dim3 grid((a.W + kTileW - 1) / kTileW, a.segs.tiles);
  if (a.hist.n > 1) ctemporal_kernel<true><<<grid, kThreads, 0, s>>>(a);
  else ctemporal_kernel<false><<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// K4 as a device function: constructGlobalIllumination (denoiser_compose_functions.glsl:53-107)
// ------------------------------------------------------------------------------------------------------------------
// The arithmetic is the exact K4's (IEEE division / sqrt, k_denoise.cu: gi_compose_kernel) — measured at 4K (tools/parity_at_size.py): with SFU
// reciprocals here another 2.4e-4 of the pixels leave the 1e-3 band, because the two fp16 inputs already sit up to one fp16 ulp
// (9.8e-4) from the oracle's; the pixel-centre fetch is the centre texel and pow(x, 5) is multiplies, as in the round-1 fast K4.
RFX_D float4 c_compose(const CamD& cam, int x, int y, int W, int H, float4 g, float rough0, float depth, v3 dgi, v3 sgi) {
  const v2 vUv = pixel_uv(x, y, W, H);
  const v3 diffuse = xyz(floatToVec4(g.x));
  const v3 wn = unpackNormal(g.y);
  const float metalness = gb_metalness(g.z);
  const v3 emissive = decodeRGBE8(floatToVec4(g.w));
  const v3 viewNormal = mul_dir_left(wn, cam.camera_matrix_world);
  const float gz = cam.perspective ? perspectiveDepthToViewZ(depth, cam.near_plane, cam.far_plane) : orthographicDepthToViewZ(depth, cam.near_plane, cam.far_plane);
  const float viewZ = -gz;
  const float clipW = cam.projection.m[2 * 4 + 3] * viewZ + cam.projection.m[3 * 4 + 3];
  v4 clip = mk4((vUv.x - 0.5f) * 2.0f, (vUv.y - 0.5f) * 2.0f, (viewZ - 0.5f) * 2.0f, 1.0f);
  clip = mk4(clip.x * clipW, clip.y * clipW, clip.z * clipW, clip.w * clipW);
  v3 viewPos = xyz(mul(cam.projection_inverse, clip));
  viewPos.z = -viewZ;
  const v3 viewDir = normalize(viewPos);
  const float roughness = rough0 * rough0;
  const v3 N = mul_dir_left(viewNormal, cam.view_matrix);
  v3 T, B;
  const v3 v = -viewDir;
  v3 V = mul_dir_left(v, cam.view_matrix);
  Onb(N, T, B);
  V = ToLocal(T, B, N, V);
  v3 Hh = SampleGGXVNDF_cs(V, roughness, roughness, 0.25f, -4.37113883e-08f, 1.0f);  // r2 = 0.25: (cos, sin) of fp32(pi/2)
  if (Hh.z < 0.0f) Hh = -Hh;
  v3 l = normalize(reflect(-V, Hh));
  l = ToWorld(T, B, N, l);
  l = xyz(mul(mk4(l, 1.0f), cam.camera_matrix_world));
  l = normalize(l);
  if (dot(viewNormal, l) < 0.0f) l = -l;
  const v3 h = normalize(v + l);
  const float VoH = fmaxf(1e-6f, dot(v, h));
  const v3 f0 = mix(mk3(0.04f), diffuse, metalness);
  const float omv = 1.0f - VoH, omv2 = omv * omv;
  const v3 F = f0 + (mk3(1.0f) - f0) * (omv2 * omv2 * omv);
  const v3 gi = diffuse * (1.0f - metalness) * (mk3(1.0f) - F) * dgi + sgi * F + emissive;
  return make_float4(gi.x, gi.y, gi.z, 1.0f);
}

// ------------------------------------------------------------------------------------------------------------------
// K3  Poisson pass on the interleaved planes.  FIRST: `in` = tr (fp32, NEAREST); else `in` = dn (fp16, LINEAR).
// Arithmetic = poisson_fast_kernel (k_denoise.cu): colours in log2 units, merged exponents.
// ------------------------------------------------------------------------------------------------------------------
RFX_D bool in_segs(const RowSegs& s, int y) {
  bool r = false;
  for (int k = 0; k < s.n; k++) r = r || (y >= s.r0[k] && y < s.r1[k]);
  return r;
}
RFX_D v3 cp_log1p(v3 c) { return mk3(fx_lg2(c.x + 1.0f), fx_lg2(c.y + 1.0f), fx_lg2(c.z + 1.0f)); }
#define CP_LUM_C (-0.06609580f) /* 0.125 * log2(ln 2) */
RFX_D float cp_lum(v3 c2) { return fx_ex2(fma_(0.125f, fx_lg2(dot(mk3(0.2125f, 0.7154f, 0.0721f), c2)), CP_LUM_C)); }

// The two planes ride in the two halves of packed fp32x2 registers (f2: .lo = diffuse plane, .hi = specular plane).
struct CTexel2 { f2 r, g, b, a; };
// INTERIOR: every tap of this block stays inside the image, so no index is clamped and the four bilinear corners are
// p, p + 16, p + pitch, p + pitch + 16 (immediate offsets off two address computations).
template <bool FIRST, bool ALPHA, bool INTERIOR>
RFX_D CTexel2 cp_fetch(const CPoissonArgs& a, float fxn, float fyn, int nx, int ny) {
  CTexel2 t;
  t.a = mkf2(0.0f);
  if (FIRST) {  // NEAREST fp32 pair
    const float4* p = (const float4*)(a.in.p + ((unsigned)ny * (unsigned)a.in.pitch + (unsigned)nx * 32u));
    const float4 u0 = __ldg(p), u1 = __ldg(p + 1);
    t.r = mkf2(u0.x, u1.x); t.g = mkf2(u0.y, u1.y); t.b = mkf2(u0.z, u1.z);
    if (ALPHA) t.a = mkf2(u0.w, u1.w);
  } else {      // LINEAR fp16 pair: one bilinear setup, one LDG.128 per corner (fxn = uv.x * W, the product bilin_setup forms)
    const float fx = fxn - 0.5f, fy = fyn - 0.5f;
    float x0f, y0f;
    const int ix = floor_i(fx, x0f), iy = floor_i(fy, y0f);
    const float ax = fx - x0f, ay = fy - y0f;
    const float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
    uint4 t00, t10, t01, t11;
    if (INTERIOR) {
      const unsigned char* p0 = a.in.p + ((unsigned)iy * (unsigned)a.in.pitch + (unsigned)ix * 16u);
      const unsigned char* p1 = p0 + a.in.pitch;
      t00 = __ldg((const uint4*)p0); t10 = __ldg((const uint4*)(p0 + 16)); t01 = __ldg((const uint4*)p1); t11 = __ldg((const uint4*)(p1 + 16));
    } else {
      const int x0 = clamp_idx(ix, a.W - 1), x1 = clamp_idx(ix + 1, a.W - 1), y0 = clamp_idx(iy, a.H - 1), y1 = clamp_idx(iy + 1, a.H - 1);
      const unsigned r0 = (unsigned)y0 * (unsigned)a.in.pitch, r1 = (unsigned)y1 * (unsigned)a.in.pitch;
      t00 = __ldg((const uint4*)(a.in.p + (r0 + (unsigned)x0 * 16u))); t10 = __ldg((const uint4*)(a.in.p + (r0 + (unsigned)x1 * 16u)));
      t01 = __ldg((const uint4*)(a.in.p + (r1 + (unsigned)x0 * 16u))); t11 = __ldg((const uint4*)(a.in.p + (r1 + (unsigned)x1 * 16u)));
    }
    // weighted sum in the sampler's order: fma(t11, w11, fma(t01, w01, fma(t10, w10, t00 * w00)))
#define CP_BL(F, M0, M1) f2fma(mkf2(F(t11.M0), F(t11.M1)), mkf2(w11), f2fma(mkf2(F(t01.M0), F(t01.M1)), mkf2(w01), \
                         f2fma(mkf2(F(t10.M0), F(t10.M1)), mkf2(w10), f2mul(mkf2(F(t00.M0), F(t00.M1)), mkf2(w00)))))
    t.r = CP_BL(h_lo, x, z); t.g = CP_BL(h_hi, x, z); t.b = CP_BL(h_lo, y, w);
    if (ALPHA) t.a = CP_BL(h_hi, y, w);
#undef CP_BL
  }
  return t;
}
RFX_D f2 cp_lum2(f2 r, f2 g, f2 b) {  // luminance(): pow(dot(w, c), 0.125) on log2-unit colours, both planes
  const f2 d = f2fma(b, mkf2(0.0721f), f2fma(g, mkf2(0.7154f), f2mul(r, mkf2(0.2125f))));
  return f2ex2(f2fma(f2lg2(d), mkf2(0.125f), mkf2(CP_LUM_C)));
}

template <bool FIRST, bool COMPOSE, bool INTERIOR>
RFX_D void cpoisson_body(const CPoissonArgs& a, int x, int y, float4 nc, float fwn) {
  const v2 vUv = pixel_uv(x, y, a.W, a.H);
  const float depth = nc.w;
  const v3 normal = mk3(nc.x, nc.y, nc.z);
  const float roughness = nrdz_roughness(nc);
  const float resx = (float)a.W, resy = (float)a.H;
  f2 accr, accg, accb, tw = mkf2(1.0f), lumc, age, alpha;
  {
    const CTexel2 c = cp_fetch<FIRST, true, INTERIOR>(a, vUv.x * resx, vUv.y * resy, x, y);
    alpha = c.a;
    age = f2ex2(f2mul(f2lg2(f2add(alpha, mkf2(1.0f))), mkf2(-1.2f * a.phi)));
    accr = f2lg2(f2fma(c.r, mkf2(1.0003f), mkf2(1.0f))); accg = f2lg2(f2fma(c.g, mkf2(1.0003f), mkf2(1.0f))); accb = f2lg2(f2fma(c.b, mkf2(1.0003f), mkf2(1.0f)));
    lumc = cp_lum2(accr, accg, accb);
  }
  const float glossiness = fmaxf(0.0f, 4.0f * (1.0f - roughness * 4.0f));
  const float sarg2 = (-glossiness * a.specular_phi) * C_LOG2E;   // specular plane only: w *= exp(-glossiness * specularPhi)
  const f2 sargp = mkf2(0.0f, sarg2), swdp = mkf2(1.0f, fx_ex2(sarg2 * 0.1f));
  const f2 nlphi2 = mkf2(-a.luma_phi * C_LOG2E);
  float flatness = 1.0f - fminf(fwn, 1.0f);
  flatness = flatness * flatness * 0.75f + 0.25f;
  const uchar4 bn = __ldg(a.blue.tex + blue_index(a.blue, x, y));
  const float2 sc = __ldg(a.rot_table + bn.x);
  const float k = a.radius * flatness;
  const float m00 = k * sc.y, m01 = k * -sc.x, m10 = k * sc.x, m11 = k * sc.y;
  const f2 oma = f2sub(mkf2(1.0f), age);
  const f2 agewd = f2mul(age, swdp);  // age * (1, specular share of pow(w, 0.1))

#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float ox = a.tap_ox[i], oy = a.tap_oy[i];
    const v2 nuv = mk2(vUv.x + (m00 * ox + m10 * oy), vUv.y + (m01 * ox + m11 * oy));  // tap position: the oracle's arithmetic, op for op
    const float fxn = nuv.x * resx, fyn = nuv.y * resy;
    int nx = floor_i(fxn), ny = floor_i(fyn);
    if (!INTERIOR) { nx = clamp_idx(nx, a.W - 1); ny = clamp_idx(ny, a.H - 1); }
    const float4 nn = ld_f4(a.nrdz, nx, ny);
    if (nn.w == 1.0f) continue;  // background tap: wBasic = 0
    const float normalDiff = 1.0f - fmaxf(dot(normal, mk3(nn.x, nn.y, nn.z)), 0.0f);
    const float depthDiff = 10000.0f * fabsf(depth - nn.w);
    const float roughnessDiff = fabsf(roughness - nrdz_roughness(nn));
    const float A2 = (-normalDiff * a.normal_phi - depthDiff * a.depth_phi - roughnessDiff * a.roughness_phi) * C_LOG2E;
    const float wdA = fx_ex2(A2 * 0.1f);
    const CTexel2 c = cp_fetch<FIRST, false, INTERIOR>(a, fxn, fyn, nx, ny);
    const f2 lr = f2lg2(f2add(c.r, mkf2(1.0f))), lg = f2lg2(f2add(c.g, mkf2(1.0f))), lb = f2lg2(f2add(c.b, mkf2(1.0f)));
    const f2 dl = f2sub(lumc, cp_lum2(lr, lg, lb));
    const f2 lumaDiff = mkf2(fminf(fabsf(f2lo(dl)), 0.5f), fminf(fabsf(f2hi(dl)), 0.5f));
    const f2 wl = f2ex2(f2fma(lumaDiff, nlphi2, f2add(sargp, mkf2(A2))));         // w * lumaFactor
    const f2 wm = f2mul(f2fma(agewd, mkf2(wdA), f2mul(wl, oma)), age);            // mix(w * lumaFactor, pow(w, 0.1), age) * age
    const f2 w = mkf2(f2lo(wm) < 0.0001f ? 0.0f : f2lo(wm), f2hi(wm) < 0.0001f ? 0.0f : f2hi(wm));
    accr = f2fma(w, lr, accr); accg = f2fma(w, lg, accg); accb = f2fma(w, lb, accb);
    tw = f2add(tw, w);
  }
  const f2 inv = mkf2(fx_rcp(f2lo(tw)), fx_rcp(f2hi(tw)));
  const f2 orr = f2sub(f2ex2(f2mul(accr, inv)), mkf2(1.0f)), og = f2sub(f2ex2(f2mul(accg, inv)), mkf2(1.0f)), ob = f2sub(f2ex2(f2mul(accb, inv)), mkf2(1.0f));
  uint4 q;
  q.x = pack_h2(f2lo(orr), f2lo(og)); q.y = pack_h2(f2lo(ob), f2lo(alpha)); q.z = pack_h2(f2hi(orr), f2hi(og)); q.w = pack_h2(f2hi(ob), f2hi(alpha));
  *((uint4*)(a.out.p + ((unsigned)y * (unsigned)a.out.pitch + (unsigned)x * 16u))) = q;
  if (COMPOSE) {
    if (in_segs(a.csegs, y)) {  // K4 reads the fp16 texel just stored (DenoiserComposePass.js:66-67)
      const float4 g = ld_f4(a.gb, x, y);
      st_f4(a.composed.p, a.composed.pitch, x, y,
            c_compose(a.cam, x, y, a.W, a.H, g, roughness, depth, mk3(h_lo(q.x), h_hi(q.x), h_lo(q.y)), mk3(h_lo(q.z), h_hi(q.z), h_lo(q.w))));
    }
  }
}

template <bool FIRST, bool COMPOSE>
__global__ void __launch_bounds__(kThreads, 4) cpoisson_kernel(const __grid_constant__ CPoissonArgs a) {
  int x, y;
  const bool in_rows = seg_pixel(a.segs, x, y);
  const bool active = x < a.W && y < a.H && in_rows;
  const int xc = min(x, a.W - 1), yc = min(y, a.H - 1);
  const float4 nc = ld_f4(a.nrdz, xc, yc);
  const float depth = nc.w;
  const float fwd = fwidth_f(depth);
  const float fwn = fx_length(fwidth_3(mk3(nc.x, nc.y, nc.z)));
  if (!active) return;
  if (depth == 1.0f && fwd == 0.0f) {  // discard :129-132: the target keeps its texel (double-buffered targets: carry it forward)
    if (a.carry.local.p) {
      const unsigned off = (unsigned)y * (unsigned)a.carry.local.pitch + (unsigned)x * 16u;
      *((uint4*)(a.out.p + ((unsigned)y * (unsigned)a.out.pitch + (unsigned)x * 16u))) = *((const uint4*)(peer_row_base(a.carry, y) + off));
    }
    if (COMPOSE && a.composed_carry.local.p && in_segs(a.csegs, y)) {
      const unsigned off = (unsigned)y * (unsigned)a.composed_carry.local.pitch + (unsigned)x * 16u;
      st_f4(a.composed.p, a.composed.pitch, x, y, *((const float4*)(peer_row_base(a.composed_carry, y) + off)));
    }
    return;
  }
  // block-uniform: can any tap of this 16x16 tile leave the image?
  int lx, ly;
  lane_to_pixel(threadIdx.x & 31, lx, ly);
  const int bx0 = blockIdx.x * kTileW, by0 = y - ((int)((threadIdx.x >> 6) << 2) + ly);
  const bool interior = bx0 - a.reach_x >= 0 && bx0 + kTileW - 1 + a.reach_x <= a.W - 1 && by0 - a.reach_y >= 0 && by0 + kTileH - 1 + a.reach_y <= a.H - 1;
  if (interior) cpoisson_body<FIRST, COMPOSE, true>(a, x, y, nc, fwn);
  else cpoisson_body<FIRST, COMPOSE, false>(a, x, y, nc, fwn);
}

cudaError_t launch_cpoisson(const CPoissonArgs& a, cudaStream_t s) {
  dim3 grid((a.W + kTileW - 1) / kTileW, a.segs.tiles);
  if (a.first) { if (a.compose) cpoisson_kernel<true, true><<<grid, kThreads, 0, s>>>(a); else cpoisson_kernel<true, false><<<grid, kThreads, 0, s>>>(a); }
  else { if (a.compose) cpoisson_kernel<false, true><<<grid, kThreads, 0, s>>>(a); else cpoisson_kernel<false, false><<<grid, kThreads, 0, s>>>(a); }
  return cudaGetLastError();
}

// stand-alone K4 (denoiseIterations == 0: no Poisson pass to ride on)
__global__ void __launch_bounds__(kThreads) ccompose_kernel(const __grid_constant__ CComposeArgs a) {
  int x, y;
  const bool in_rows = seg_pixel(a.segs, x, y);
  const bool active = x < a.W && y < a.H && in_rows;
  const int xc = min(x, a.W - 1), yc = min(y, a.H - 1);
  const float4 nc = ld_f4(a.nrdz, xc, yc);
  const float fwd = fwidth_f(nc.w);
  if (!active) return;
  if (nc.w == 1.0f && fwd == 0.0f) {
    if (a.composed_carry.local.p) {
      const unsigned off = (unsigned)y * (unsigned)a.composed_carry.local.pitch + (unsigned)x * 16u;
      st_f4(a.composed.p, a.composed.pitch, x, y, *((const float4*)(peer_row_base(a.composed_carry, y) + off)));
    }
    return;
  }
  const uint4 q = __ldg((const uint4*)(a.dn.p + ((unsigned)y * (unsigned)a.dn.pitch + (unsigned)x * 16u)));
  const float4 g = ld_f4(a.gb, x, y);
  st_f4(a.composed.p, a.composed.pitch, x, y,
        c_compose(a.cam, x, y, a.W, a.H, g, nrdz_roughness(nc), nc.w, mk3(h_lo(q.x), h_hi(q.x), h_lo(q.y)), mk3(h_lo(q.z), h_hi(q.z), h_lo(q.w))));
}
cudaError_t launch_ccompose(const CComposeArgs& a, cudaStream_t s) {
  dim3 grid((a.W + kTileW - 1) / kTileW, a.segs.tiles);
  ccompose_kernel<<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// chain_output() views: interleaved -> the reference's plane formats
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tr_kernel(PV tr, OutV o0, OutV o1, int W, int H) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float4* p = (const float4*)(tr.p + ((unsigned)y * (unsigned)tr.pitch + (unsigned)x * 32u));
  st_f4(o0.p, o0.pitch, x, y, p[0]);
  st_f4(o1.p, o1.pitch, x, y, p[1]);
}
__global__ void __launch_bounds__(256) split_dn_kernel(PV dn, OutV o0, OutV o1, int W, int H) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const uint4 q = *((const uint4*)(dn.p + ((unsigned)y * (unsigned)dn.pitch + (unsigned)x * 16u)));
  *((uint2*)(o0.p + ((unsigned)y * (unsigned)o0.pitch + (unsigned)x * 8u))) = make_uint2(q.x, q.y);
  *((uint2*)(o1.p + ((unsigned)y * (unsigned)o1.pitch + (unsigned)x * 8u))) = make_uint2(q.z, q.w);
}
cudaError_t launch_split_tr(PV tr, OutV o0, OutV o1, int W, int H, cudaStream_t s) {
  split_tr_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), 256, 0, s>>>(tr, o0, o1, W, H);
  return cudaGetLastError();
}
cudaError_t launch_split_dn(PV dn, OutV o0, OutV o1, int W, int H, cudaStream_t s) {
  split_dn_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), 256, 0, s>>>(dn, o0, o1, W, H);
  return cudaGetLastError();
}

}  // namespace rfx
