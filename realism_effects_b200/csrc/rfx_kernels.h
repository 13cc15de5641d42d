// rfx_kernels.h — internal launch interface between the C-ABI layer (rfx_api.cu) and the
// sm_100a kernels (k_*.cu).  Everything here is plain structs of device pointers + uniforms.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encode entry point is resolved at run time)

#include "rfx_device.cuh"
#include "../../include/rfx.h"

namespace rfx {

struct OutV {  // writable plane view
  unsigned char* p;
  long long pitch;
};

// Output rows of one launch: up to RFX_MAX_SEGS disjoint row segments [r0,r1) (one per row block this rank owns), covered by a
// single grid: blockIdx.y walks the 16-row tiles of segment 0, then segment 1, ...  One launch per pass regardless of how
// many row blocks a rank owns (a launch per block costs ~10 % in launch gaps and partial last waves).
#define RFX_MAX_SEGS 16
struct RowSegs {
  int n, tiles;                  // segments, total 16-row tiles (= gridDim.y)
  int r0[RFX_MAX_SEGS], r1[RFX_MAX_SEGS];
  int tile0[RFX_MAX_SEGS + 1];   // first tile index of each segment
};

#define RFX_MAX_PEERS 8
struct PeerPV {
  PV local;                                    // this rank's allocation (full-frame sized); rows [own0, own1) are valid here
  const unsigned char* base[RFX_MAX_PEERS];    // every rank's allocation of the same plane (peer-mapped), same pitch
  int bound[RFX_MAX_PEERS + 1];                // rank k owns rows [bound[k], bound[k+1]) of the data being read
  int n;                                       // ranks; <= 1: single GPU (local only)
  int own0, own1;
};
RFX_D const unsigned char* peer_row_base(const PeerPV& p, int y) {
  if (p.n <= 1 || (y >= p.own0 && y < p.own1)) return p.local.p;
  int k = 0;
#pragma unroll
  for (int i = 1; i < RFX_MAX_PEERS; i++) k += (i < p.n && y >= p.bound[i]) ? 1 : 0;
  return p.base[k];
}

struct CamD {  // device copy of rfx_camera
  M4 projection, projection_inverse, camera_matrix_world, view_matrix;
  float near_plane, far_plane;
  int perspective;
};

struct BlueD {
  const uchar4* tex;  // size x size RGBA8, GL texel order
  int size;
  BlueShift shift;    // (pcg4d(seed(index)).xy % 0x0fffffff) % size, computed on the host
  int index;          // raw index (0 selects the tiled lookup of blue_noise.glsl:38-39)
  int mask;           // size - 1 when size is a power of two (the shipped 128 x 128 texture), else 0
};
RFX_D int blue_index(const BlueD& b, int x, int y) {
  if (b.mask) return ((y + b.shift.sy) & b.mask) * b.size + ((x + b.shift.sx) & b.mask);
  return ((y + b.shift.sy) % b.size) * b.size + ((x + b.shift.sx) % b.size);
}

struct EnvD {
  PV mip[16];  // RGBA16F levels
  int levels;
  PV marginal, conditional;  // R32F
  float size_x, size_y, total_sum_whole, total_sum_decimal;
};

// ---- K3 ----------------------------------------------------------------------------------
struct PoissonArgs {
  PV depth, gb, in0, in1;
  OutV out0, out1;
  int W, H, row0, row1;
  RowSegs segs;
  float radius, phi, luma_phi, depth_phi, normal_phi, roughness_phi, specular_phi;
  int texture_count, spec0, spec1, gbuffer_texture, input_linear, in_half;
  BlueD blue;
  const float2* rot_table;  // [256] (sin, cos) of (k/255)*2*pi, correctly rounded
  // fast variant only:
  PV nrd;                       // float4 (normal.xyz, roughness) from the decode prepass
  float tap_ox[8], tap_oy[8];   // POISSON[i] / resolution
};
cudaError_t launch_poisson(const PoissonArgs& a, cudaStream_t s);       // exact-libm variant (any configuration)
cudaError_t launch_poisson_fast(const PoissonArgs& a, cudaStream_t s);  // SFU variant (GBUFFER_TEXTURE configurations)
cudaError_t launch_gbuffer_decode(PV gb, OutV nrd, int W, int H, int gbuffer_texture, const RowSegs& segs, int halo, cudaStream_t s);

// ---- K4 / K5 -------------------------------------------------------------------------------
struct ComposeArgs {
  PV depth, gb, diffuse, specular, scene;  // diffuse / specular / scene: p == nullptr = not bound (null sampler)
  OutV out;
  int W, H, row0, row1;
  RowSegs segs;
  CamD cam;
  int input_type;
  int gi_f32;  // diffuse / specular are RGBA32F NEAREST (denoiseMode "full_temporal": the temporal pass's targets)
  int fast;
};
cudaError_t launch_gi_compose(const ComposeArgs& a, cudaStream_t s);

struct SsgiComposeArgs {
  PV depth, gi, scene;
  OutV out;
  int W, H, row0, row1;
  int use_fog, fog_exp2, perspective, is_debug;
  float fog_color[3], fog_near, fog_far, fog_density, camera_near, camera_far;
};
cudaError_t launch_ssgi_compose(const SsgiComposeArgs& a, cudaStream_t s);

// ---- K2 ----------------------------------------------------------------------------------
struct TemporalArgs {
  PV input, velocity, hist0, hist1;
  OutV out0, out1;
  int W, H, row0, row1;
  RowSegs segs;
  CamD cam;
  M4 prev_view, prev_world, prev_proj, prev_proj_inv;
  M4 prev_proj_view;  // prevProjectionMatrix * prevViewMatrix (reproject.frag:183), fma-lowered on the host
  float camera_pos[3];
  float max_blend, clamp_intensity, keep_data, confidence_power;
  float inv_w, inv_h;  // invTexSize
  int full_accumulate, texture_count, input_type, log_transform, rs0, rs1, history_linear;
  int input_half, out_half;
  int in_scaled;  // inputTexture (the SSGI target) is smaller than the output (resolutionScale < 1): NEAREST fetch by uv
  int hist_f32;  // history planes are RGBA32F (denoiseMode "full_temporal" / "temporal")
  int fast;  // SFU variants of log/exp/pow
};
cudaError_t launch_temporal(const TemporalArgs& a, cudaStream_t s);

// ---- K1 ----------------------------------------------------------------------------------
struct SsgiArgs {
  PV depth, gb, velocity, direct, accumulated;  // velocity/direct/accumulated may have p == nullptr
  OutV out;
  int W, H, row0, row1;
  RowSegs segs;
  CamD cam;
  float ray_distance, thickness, env_blur, max_env_mip;
  float near_minus_far, near_mul_far, far_minus_near;
  int steps, refine_steps, mode;
  unsigned flags;
  BlueD blue;
  EnvD env;
  const float2* rot_table;   // [256] (sin, cos)
  const float* step_table;   // [steps][256]  cs(i, b) = 1 - exp(-0.25 (i + b - 0.5)^2), row i-1 for step i
  PV viewz;                  // R32F: getViewZ(depth) per texel (launch_viewz prepass)
  int proj_sparse;           // projection matrix has the perspective sparsity pattern (exact-zero terms dropped)
  int fast;                  // SFU variants of the continuous transcendentals
  int phase;                 // 0 fused; 1 ray march only -> rec; 2 shading from rec (k_ssgi.cu "Split-phase K1")
  unsigned char* rec;        // 2 x float4 per pixel (diffuse ray, specular ray)
  long long rec_pitch;
  // fast fused kernel (ssgi_fast_kernel)
  float ps_x0, ps_x2, ps_y1, ps_y2, ps_hw, ps_hh;  // projection rows scaled to texel units: tx = (ps_x0*x + ps_x2*z) / -z + ps_hw
  int vz_pitchw;             // viewZ pitch in 4-byte words
  int vz_tiled;              // experiment (RFX_K1_VZ_TILED=1): the viewZ scratch is stored as 8x4-texel tiles (one 128-B line each) for the fast kernel's gathers
  int vz_tiles_x;            // tiles per tile row
  int legacy_fast;           // 1: use the round-1 fast kernel (tools/ A/B comparisons)
  int scaled;                // the render target (W x H) is smaller than the input planes (resolutionScale < 1): texels are fetched by uv
  int march_batch;           // march steps fetched together before they are tested: 1, 2 or 4
  PeerPV acc_peer;           // `accumulated` in a row-sharded group (n > 1): rows live on their owners
};
cudaError_t launch_ssgi(const SsgiArgs& a, cudaStream_t s);
cudaError_t launch_viewz(const SsgiArgs& a, OutV vz, cudaStream_t s);

// ---- K6 / K7 / K8 / K9 -----------------------------------------------------------------------
struct HbaoArgs {
  PV depth;
  OutV out;
  int W, H, row0, row1;
  M4 projection_view, projection_inverse, camera_matrix_world;
  float ao_distance, distance_power, bias, thickness;
  int spp;
  BlueD blue;
  const float2* rot_table;
};
cudaError_t launch_hbao(const HbaoArgs& a, cudaStream_t s);

struct AoComposeArgs {
  PV depth, ao, input;
  OutV out;
  int W, H, row0, row1;
  float power, color[3];
};
cudaError_t launch_ao_compose(const AoComposeArgs& a, cudaStream_t s);

struct MotionBlurArgs {
  PV velocity, input;
  OutV out;
  int W, H, row0, row1;
  float intensity, jitter, delta_time, res_x, res_y;
  int samples;
  BlueD blue;
};
cudaError_t launch_motion_blur(const MotionBlurArgs& a, cudaStream_t s);

struct TraaComposeArgs {
  PV acc;
  OutV out;
  int W, H, row0, row1;
};
cudaError_t launch_traa_compose(const TraaComposeArgs& a, cudaStream_t s);

// merged cosmetic effects + TAAPass (k_fx.cu)
struct EffectsArgs {
  PV input, depth, velocity;
  OutV out;
  CamD cam;
  int W, H, row0, row1;
  int n_effects, effects[4];
  float texel_x, texel_y;  // postprocessing's texelSize = 1 / size (a JS double rounded to fp32)
  float sharpness, alphax, alphay, aberration, bg[3], max_distance, spread, intensity;
  int sparkle_perspective;
};
cudaError_t launch_effects(const EffectsArgs& a, cudaStream_t s);
struct TaaArgs {
  PV input, history;
  OutV out;
  int W, H, row0, row1;
  float camera_not_moved_frames;
  int srgb_output;
};
cudaError_t launch_taa(const TaaArgs& a, cudaStream_t s);

// G-buffer ingest (k_ingest.cu)
struct IngestArgs {
  PV albedo, normal, material, emissive, motion, depth;  // emissive.p / motion.p may be null
  OutV out_gb, out_vel;                                  // either .p may be null
  int W, H, row0, row1;
  int albedo_half, material_half, normal_f32, motion_f32, normalize_normals;
  float motion_sx, motion_sy;
};
cudaError_t launch_gbuffer_ingest(const IngestArgs& a, cudaStream_t s);

// env mip chain: dst (w1 x h1) = box filter of src (w0 x h0), RGBA16F
cudaError_t launch_env_downsample(PV src, OutV dst, int w1, int h1, cudaStream_t s);
// importance-sampling tables of an equirect map on the device (gatherData, EquirectHdrInfoUniform.js:149-245); 3 launches
cudaError_t launch_env_cdf(PV map, int flip_y, float* cdf_c, float* cdf_m, double* row_sum, double* total, float* marginal, float* conditional, cudaStream_t s);


// ==========================================================================================
// fast chain (k_chain.cu): the SSGI-mode chain of rfx_ssgi_chain_* with fast_math on.  Same passes, same tap geometry and
// decisions as the per-pass kernels above, but chain-internal plane formats chosen for the tap loops:
//   nrdz  16 B  (n.xyz, depth) + roughness code in the low mantissa bits          (decode prepass, once per frame)
//   tr    32 B  {diffuse rgba fp32, specular rgba fp32}                            K2 -> K3 pass 0 (NEAREST)
//   dn    16 B  {diffuse rgba fp16, specular rgba fp16}                            K3 ping-pong (LINEAR), K2 history, K4 input
// and the last Poisson pass also does the GI compose for its pixel (K4 has no neighbourhood in the fast variant).
// Row-sharded multi-GPU frames read last frame's `composed` / `dn` rows owned by other ranks in place over NVLink (PeerPV).
// ==========================================================================================
struct CDecodeArgs { PV gb, depth; OutV nrdz; int W, H; };
cudaError_t launch_cdecode(const CDecodeArgs& a, const RowSegs& segs, int halo, cudaStream_t s);

struct CTemporalArgs {
  PV input, velocity;   // K1 output (packed), velocity plane
  PeerPV hist;          // dn of the previous frame
  OutV out;             // tr
  int W, H;
  RowSegs segs;
  CamD cam;
  M4 prev_world, prev_proj_inv, prev_proj_view;
  float camera_pos[3];
  float max_blend, clamp_intensity, keep_data, confidence_power;
  float inv_w, inv_h;
  int full_accumulate;
};
cudaError_t launch_ctemporal(const CTemporalArgs& a, cudaStream_t s);

struct CPoissonArgs {
  PV nrdz, in;          // in: tr (pass 0) or dn (passes >= 1)
  OutV out;             // dn
  int W, H;
  RowSegs segs;
  float radius, phi, luma_phi, depth_phi, normal_phi, roughness_phi, specular_phi;
  BlueD blue;
  const float2* rot_table;
  float tap_ox[8], tap_oy[8];
  int first;            // pass 0: `in` is tr, NEAREST
  int reach_x, reach_y; // pixels a tap (incl. its bilinear footprint) can lie from its pixel: blocks further than that from every border skip all clamping
  // `out` is double-buffered by frame parity when it is the history plane: a discarded pixel (the reference's "target keeps its
  // texel") then copies last frame's texel forward.  carry.local.p == nullptr: single-buffered target, no write.
  PeerPV carry;
  // fused GI compose (last pass): rows of `csegs` also write `composed` (discarded pixels carry last frame's texel forward)
  int compose;
  int compose_mode;     // arithmetic of the fused K4: 0 IEEE, 1 SFU, 2 SFU + one Newton step (k_chain.cu: c_compose_t)
  RowSegs csegs;
  PV gb;
  OutV composed;
  PeerPV composed_carry;
  CamD cam;
};
cudaError_t launch_cpoisson(const CPoissonArgs& a, cudaStream_t s);
struct CPoissonTmaArgs {  // passes >= 1 with TMA-staged tiles (experiment, RFX_K3_TMA=1)
  CPoissonArgs a;
  CUtensorMap map_in, map_nrdz;  // 2-D maps over the 16-byte texel planes as rows of 4-byte elements; box = box_w*4 x box_h
  int box_w, box_h;              // texels: 16 + 2 * reach_x, 16 + 2 * reach_y
};
cudaError_t launch_cpoisson_tma(const CPoissonTmaArgs& t, cudaStream_t s);

struct CComposeArgs {   // stand-alone K4 over dn (denoiseIterations == 0)
  int compose_mode;
  PV nrdz, gb, dn;
  OutV composed;
  PeerPV composed_carry;
  int W, H;
  RowSegs segs;
  CamD cam;
};
cudaError_t launch_ccompose(const CComposeArgs& a, cudaStream_t s);

// chain_output() views of the interleaved planes in the reference's formats
cudaError_t launch_split_tr(PV tr, OutV o0, OutV o1, int W, int H, cudaStream_t s);   // 32 B -> 2 x RGBA32F
cudaError_t launch_split_dn(PV dn, OutV o0, OutV o1, int W, int H, cudaStream_t s);   // 16 B -> 2 x RGBA16F

// common launch geometry: 256-thread blocks, 8 warps as 2 x 4 warp tiles of 8x4 pixels => 16x16 pixel tile
constexpr int kTileW = 16, kTileH = 16, kThreads = 256;
RFX_D void block_pixel(int& x, int& y, int row_base) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lx, ly;
  lane_to_pixel(lane, lx, ly);
  x = blockIdx.x * kTileW + ((warp & 1) << 3) + lx;
  y = row_base + blockIdx.y * kTileH + ((warp >> 1) << 2) + ly;
}
// pixel of this thread for a multi-segment launch; returns whether its row lies inside the segment (quads stay aligned to
// even rows because every segment's tiles start at r0 & ~1)
RFX_D bool seg_pixel(const RowSegs& s, int& x, int& y) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lx, ly;
  lane_to_pixel(lane, lx, ly);
  int k = 0;
  const int ty = blockIdx.y;
  while (k + 1 < s.n && ty >= s.tile0[k + 1]) k++;
  x = blockIdx.x * kTileW + ((warp & 1) << 3) + lx;
  y = (s.r0[k] & ~1) + (ty - s.tile0[k]) * kTileH + ((warp >> 1) << 2) + ly;
  return y >= s.r0[k] && y < s.r1[k];
}
// host: build the segment table
inline RowSegs make_segs(const int* r0, const int* r1, int n) {
  RowSegs s{};
  s.n = n;
  int t = 0;
  for (int i = 0; i < n; i++) {
    s.r0[i] = r0[i]; s.r1[i] = r1[i]; s.tile0[i] = t;
    t += (r1[i] - (r0[i] & ~1) + kTileH - 1) / kTileH;
  }
  s.tile0[n] = t;
  s.tiles = t;
  return s;
}

}  // namespace rfx
