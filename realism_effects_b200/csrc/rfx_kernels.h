// rfx_kernels.h — internal launch interface between the C-ABI layer (rfx_api.cu) and the
// sm_100a kernels (k_*.cu).  Everything here is plain structs of device pointers + uniforms.
#pragma once
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
};

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
  PV depth, gb, diffuse, specular;
  OutV out;
  int W, H, row0, row1;
  RowSegs segs;
  CamD cam;
  int input_type;
  int fast;
};
cudaError_t launch_gi_compose(const ComposeArgs& a, cudaStream_t s);

struct SsgiComposeArgs {
  PV depth, gi, scene;
  OutV out;
  int W, H, row0, row1;
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

// env mip chain: dst (w1 x h1) = box filter of src (w0 x h0), RGBA16F
cudaError_t launch_env_downsample(PV src, OutV dst, int w1, int h1, cudaStream_t s);

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
