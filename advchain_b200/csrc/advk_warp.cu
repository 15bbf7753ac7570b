// advk_warp.cu -- single-stage image / prediction warps (bi/tri-linear or nearest gather) with
// the sampling grid produced on the fly: either an affine grid (theta * base coordinates) or a
// dense interleaved field.  Forward = gather; backward = atomic scatter to the source gradient
// plus the spatial Jacobian (gradient w.r.t. the grid), reduced to d x (d+1) per sample for the
// affine case.
//
// Replaces F.affine_grid + F.grid_sample at adv_affine.py:297-313 and F.grid_sample at
// adv_morph.py:546-557 and their autograd backward (ATen grid_sampler_{2,3}d_backward,
// affine_grid_generator_backward).
#include "advk_common.cuh"

namespace advk {

constexpr int WARP_THREADS = 256;

template <int DIM>
struct Stencil {
  Axis ax, ay, az;
};

// Sampling coordinates of output voxel (z,y,x) of sample n.
template <int DIM, bool FIELD>
__device__ __forceinline__ void sample_coords(const Dims& g, int n, int z, int y, int x, i64 p,
                                              const float* __restrict__ theta,
                                              const void* __restrict__ field, float& cx, float& cy,
                                              float& cz, float& rx, float& ry, float& rz) {
  if (FIELD) {
    if (DIM == 2) {
      float2 f = reinterpret_cast<const float2*>(field)[(i64)n * g.S + p];
      rx = f.x; ry = f.y; rz = 0.f;
    } else {
      float4 f = reinterpret_cast<const float4*>(field)[(i64)n * g.S + p];
      rx = f.x; ry = f.y; rz = f.z;
    }
    cx = clampf(rx, -1.f, 1.f); cy = clampf(ry, -1.f, 1.f); cz = clampf(rz, -1.f, 1.f);
  } else {
    float bx = base_coord(x, g.W, 0.f), by = base_coord(y, g.H, 0.f);
    if (DIM == 2) {
      const float* t = theta + n * 6;
      // the K = 3 dot product of affine_grid's bmm, accumulated in order with one rounding per step
      cx = __fadd_rn(__fmaf_rn(t[1], by, __fmul_rn(t[0], bx)), t[2]);
      cy = __fadd_rn(__fmaf_rn(t[4], by, __fmul_rn(t[3], bx)), t[5]);
      cz = 0.f;
    } else {
      float bz = base_coord(z, g.D, 0.f);
      const float* t = theta + n * 12;
      cx = __fadd_rn(__fmaf_rn(t[2], bz, __fmaf_rn(t[1], by, __fmul_rn(t[0], bx))), t[3]);
      cy = __fadd_rn(__fmaf_rn(t[6], bz, __fmaf_rn(t[5], by, __fmul_rn(t[4], bx))), t[7]);
      cz = __fadd_rn(__fmaf_rn(t[10], bz, __fmaf_rn(t[9], by, __fmul_rn(t[8], bx))), t[11]);
    }
    rx = cx; ry = cy; rz = cz;
  }
}

template <int DIM, bool FIELD>
__global__ void __launch_bounds__(WARP_THREADS)
warp_fwd_kernel(Dims g, int C, const float* __restrict__ src, const float* __restrict__ theta,
                const void* __restrict__ field, int pad, int interp,
                const float* __restrict__ padv, float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  const int x = (int)(p % g.W);
  const int y = (int)((p / g.W) % g.H);
  const int z = (int)(p / ((i64)g.W * g.H));
  float cx, cy, cz, rx, ry, rz;
  sample_coords<DIM, FIELD>(g, n, z, y, x, p, theta, field, cx, cy, cz, rx, ry, rz);
  Axis ax = make_axis(cx, g.W, pad, interp);
  Axis ay = make_axis(cy, g.H, pad, interp);
  Axis az;
  if (DIM == 3) az = make_axis(cz, g.D, pad, interp);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; az.mult = 0.f; }
  const float pv = padv ? padv[n] : 0.f;
  const i64 HW = (i64)g.H * g.W;
  for (int c = 0; c < C; ++c) {
    const float* s = src + ((i64)n * C + c) * g.S;
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
      bool vz = dz ? az.v1 : az.v0;
      float wz = dz ? az.w1 : az.w0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        bool vy = dy ? ay.v1 : ay.v0;
        float wy = dy ? ay.w1 : ay.w0;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          bool vx = dx ? ax.v1 : ax.v0;
          float wx = dx ? ax.w1 : ax.w0;
          if (vx && vy && vz) {
            i64 q = (i64)(az.i0 + dz) * HW + (i64)(ay.i0 + dy) * g.W + (ax.i0 + dx);
            acc += (__ldg(s + q) - pv) * (wx * wy * wz);
          }
        }
      }
    }
    out[((i64)n * C + c) * g.S + p] = acc + pv;
  }
}

// Backward. NG = d*(d+1) entries of g_theta per sample for the affine case.
template <int DIM, bool FIELD>
__global__ void __launch_bounds__(WARP_THREADS)
warp_bwd_kernel(Dims g, int C, const float* __restrict__ g_out, const float* __restrict__ src,
                const float* __restrict__ theta, const void* __restrict__ field, int pad,
                int interp, const float* __restrict__ padv, float* __restrict__ g_src,
                float* __restrict__ g_theta, void* __restrict__ g_field) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[12 * 32];
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < g.S;
  float ggx = 0.f, ggy = 0.f, ggz = 0.f;
  float bx = 0.f, by = 0.f, bz = 0.f;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (live) {
    const int x = (int)(p % g.W);
    const int y = (int)((p / g.W) % g.H);
    const int z = (int)(p / ((i64)g.W * g.H));
    float cx, cy, cz;
    sample_coords<DIM, FIELD>(g, n, z, y, x, p, theta, field, cx, cy, cz, rx, ry, rz);
    if (!FIELD) {
      bx = base_coord(x, g.W, 0.f); by = base_coord(y, g.H, 0.f);
      bz = (DIM == 3) ? base_coord(z, g.D, 0.f) : 0.f;
    }
    Axis ax = make_axis(cx, g.W, pad, interp);
    Axis ay = make_axis(cy, g.H, pad, interp);
    Axis az;
    if (DIM == 3) az = make_axis(cz, g.D, pad, interp);
    else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; az.mult = 0.f; }
    const float pv = padv ? padv[n] : 0.f;
    const i64 HW = (i64)g.H * g.W;
    for (int c = 0; c < C; ++c) {
      const i64 cb = ((i64)n * C + c) * g.S;
      const float go = g_out[cb + p];
      const float* s = src + cb;
#pragma unroll
      for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
        bool vz = dz ? az.v1 : az.v0;
        float wz = dz ? az.w1 : az.w0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          bool vy = dy ? ay.v1 : ay.v0;
          float wy = dy ? ay.w1 : ay.w0;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            bool vx = dx ? ax.v1 : ax.v0;
            float wx = dx ? ax.w1 : ax.w0;
            if (vx && vy && vz) {
              i64 q = (i64)(az.i0 + dz) * HW + (i64)(ay.i0 + dy) * g.W + (ax.i0 + dx);
              if (g_src) atomicAdd(g_src + cb + q, go * (wx * wy * wz));
              float v = (__ldg(s + q) - pv) * go;
              ggx += (dx ? v : -v) * (wy * wz);
              ggy += (dy ? v : -v) * (wx * wz);
              if (DIM == 3) ggz += (dz ? v : -v) * (wx * wy);
            }
          }
        }
      }
    }
    ggx *= ax.mult; ggy *= ay.mult; ggz *= az.mult;
  }
  if (FIELD) {
    if (!live || !g_field) return;
    // the consumer clamps the stored field to [-1,1] (torch.clamp passes gradient on the
    // closed interval, quirk Q7)
    if (!(rx >= -1.f && rx <= 1.f)) ggx = 0.f;
    if (!(ry >= -1.f && ry <= 1.f)) ggy = 0.f;
    if (DIM == 2) {
      reinterpret_cast<float2*>(g_field)[(i64)n * g.S + p] = make_float2(ggx, ggy);
    } else {
      if (!(rz >= -1.f && rz <= 1.f)) ggz = 0.f;
      reinterpret_cast<float4*>(g_field)[(i64)n * g.S + p] = make_float4(ggx, ggy, ggz, 0.f);
    }
  } else {
    if (!g_theta) return;
    constexpr int NG = DIM * (DIM + 1);
    float v[NG];
    if (DIM == 2) {
      v[0] = ggx * bx; v[1] = ggx * by; v[2] = ggx;
      v[3] = ggy * bx; v[4] = ggy * by; v[5] = ggy;
    } else {
      v[0] = ggx * bx; v[1] = ggx * by; v[2] = ggx * bz; v[3] = ggx;
      v[4] = ggy * bx; v[5] = ggy * by; v[6] = ggy * bz; v[7] = ggy;
      v[8] = ggz * bx; v[9] = ggz * by; v[10] = ggz * bz; v[11] = ggz;
    }
    block_sum<NG>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NG; ++k) atomicAdd(g_theta + n * NG + k, v[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Bicubic (2-D only, like F.grid_sample): GridSampler.cuh grid_sampler_2d_kernel /
// grid_sampler_2d_backward_kernel, mode == Bicubic.  The grid coordinate is only un-normalised (no
// clipping); every one of the 4 x 4 taps runs its INTEGER coordinate through the padding rule
// (clip / reflect) and a bounds test, forward and backward alike; the grid gradient uses the
// derivative of the cubic-convolution coefficients (A = -0.75) and treats the tap positions as
// constants, exactly as ATen does.
constexpr float CUBIC_A = -0.75f;
__device__ __forceinline__ float cubic_conv1(float x) { return ((CUBIC_A + 2.f) * x - (CUBIC_A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic_conv2(float x) {
  return ((CUBIC_A * x - 5.f * CUBIC_A) * x + 8.f * CUBIC_A) * x - 4.f * CUBIC_A;
}
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  c[0] = cubic_conv2(t + 1.f);
  c[1] = cubic_conv1(t);
  const float u = 1.f - t;
  c[2] = cubic_conv1(u);
  c[3] = cubic_conv2(u + 1.f);
}
__device__ __forceinline__ void cubic_coeffs_grad(float t, float (&c)[4]) {
  float x = -1.f - t;
  c[0] = (-3.f * CUBIC_A * x - 10.f * CUBIC_A) * x - 8.f * CUBIC_A;
  x = -t;
  c[1] = (-3.f * (CUBIC_A + 2.f) * x - 2.f * (CUBIC_A + 3.f)) * x;
  x = 1.f - t;
  c[2] = (3.f * (CUBIC_A + 2.f) * x - 2.f * (CUBIC_A + 3.f)) * x;
  x = 2.f - t;
  c[3] = (3.f * CUBIC_A * x - 10.f * CUBIC_A) * x + 8.f * CUBIC_A;
}
// compute_coordinates on a tap index (align_corners=True) + bounds test; returns -1 when outside
__device__ __forceinline__ int cubic_tap(float c, int size, int pad) {
  if (pad == ADVK_PAD_REFLECTION) {
    float m;
    c = gs_reflect(c, 0, 2 * (size - 1), m);
  }
  if (pad != ADVK_PAD_ZEROS) c = fminf((float)(size - 1), fmaxf(c, 0.f));
  if (!(c <= 2147483646.f && c >= -2147483648.f)) c = -100.f;
  const int i = (int)c;
  return (i >= 0 && i < size) ? i : -1;
}

struct CubicTaps {
  int ix[4], iy[4];
  float cx[4], cy[4], gx[4], gy[4];
  float mx, my;
};
__device__ __forceinline__ CubicTaps make_cubic(float coordx, float coordy, const Dims& g, int pad) {
  CubicTaps t;
  t.mx = (float)(g.W - 1) / 2.f; t.my = (float)(g.H - 1) / 2.f;
  const float fx = ((coordx + 1.f) / 2.f) * (float)(g.W - 1), fy = ((coordy + 1.f) / 2.f) * (float)(g.H - 1);
  const float nx = floorf(fx), ny = floorf(fy);
  cubic_coeffs(fx - nx, t.cx); cubic_coeffs(fy - ny, t.cy);
  cubic_coeffs_grad(fx - nx, t.gx); cubic_coeffs_grad(fy - ny, t.gy);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.ix[i] = cubic_tap(nx - 1.f + (float)i, g.W, pad);
    t.iy[i] = cubic_tap(ny - 1.f + (float)i, g.H, pad);
  }
  return t;
}

template <bool FIELD>
__global__ void __launch_bounds__(WARP_THREADS)
warp_bicubic_fwd_kernel(Dims g, int C, const float* __restrict__ src, const float* __restrict__ theta,
                        const void* __restrict__ field, int pad, const float* __restrict__ padv,
                        float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  const int x = (int)(p % g.W), y = (int)(p / g.W);
  float cx, cy, cz, rx, ry, rz;
  sample_coords<2, FIELD>(g, n, 0, y, x, p, theta, field, cx, cy, cz, rx, ry, rz);
  const CubicTaps t = make_cubic(cx, cy, g, pad);
  const float pv = padv ? padv[n] : 0.f;
  for (int c = 0; c < C; ++c) {
    const float* s = src + ((i64)n * C + c) * g.S;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float row = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = (t.ix[i] >= 0 && t.iy[j] >= 0) ? __ldg(s + (i64)t.iy[j] * g.W + t.ix[i]) - pv : 0.f;
        row += v * t.cx[i];
      }
      acc += row * t.cy[j];
    }
    out[((i64)n * C + c) * g.S + p] = acc + pv;
  }
}

template <bool FIELD>
__global__ void __launch_bounds__(WARP_THREADS)
warp_bicubic_bwd_kernel(Dims g, int C, const float* __restrict__ g_out, const float* __restrict__ src,
                        const float* __restrict__ theta, const void* __restrict__ field, int pad,
                        const float* __restrict__ padv, float* __restrict__ g_src,
                        float* __restrict__ g_theta, void* __restrict__ g_field) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[6 * 32];
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < g.S;
  float ggx = 0.f, ggy = 0.f, bx = 0.f, by = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;
  if (live) {
    const int x = (int)(p % g.W), y = (int)(p / g.W);
    float cx, cy, cz;
    sample_coords<2, FIELD>(g, n, 0, y, x, p, theta, field, cx, cy, cz, rx, ry, rz);
    if (!FIELD) { bx = base_coord(x, g.W, 0.f); by = base_coord(y, g.H, 0.f); }
    const CubicTaps t = make_cubic(cx, cy, g, pad);
    const float pv = padv ? padv[n] : 0.f;
    for (int c = 0; c < C; ++c) {
      const i64 cb = ((i64)n * C + c) * g.S;
      const float go = g_out[cb + p];
      const float* s = src + cb;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (t.ix[i] >= 0 && t.iy[j] >= 0) {
            const i64 q = (i64)t.iy[j] * g.W + t.ix[i];
            if (g_src) atomicAdd(g_src + cb + q, go * t.cx[i] * t.cy[j]);
            const float v = __ldg(s + q) - pv;
            ggx -= v * t.gx[i] * t.cy[j] * go;
            ggy -= v * t.gy[j] * t.cx[i] * go;
          }
        }
      }
    }
    ggx *= t.mx; ggy *= t.my;
  }
  if (FIELD) {
    if (!live || !g_field) return;
    if (!(rx >= -1.f && rx <= 1.f)) ggx = 0.f;
    if (!(ry >= -1.f && ry <= 1.f)) ggy = 0.f;
    reinterpret_cast<float2*>(g_field)[(i64)n * g.S + p] = make_float2(ggx, ggy);
  } else {
    if (!g_theta) return;
    float v[6] = {ggx * bx, ggx * by, ggx, ggy * bx, ggy * by, ggy};
    block_sum<6>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) atomicAdd(g_theta + n * 6 + k, v[k]);
    }
  }
}

template <bool FIELD>
static int launch_fwd(const advk_geom* gg, int C, const float* src, const float* theta,
                      const void* field, int pad, int interp, const float* padv, float* out,
                      void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(C >= 1 && src && out && (FIELD ? field != nullptr : theta != nullptr), "null pointer");
  ADVK_REQUIRE(pad >= 0 && pad <= 2 && interp >= 0 && interp <= 2, "bad pad/interp mode");
  ADVK_REQUIRE(interp != ADVK_INTERP_BICUBIC || gg->d == 2, "bicubic interpolation is 2-D only (like F.grid_sample)");
  dim3 grid(blocks_for(g.S, WARP_THREADS), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (interp == ADVK_INTERP_BICUBIC)
    ADVK_LAUNCH(K_warp_fwd, st, launch_pdl((warp_bicubic_fwd_kernel<FIELD>), grid, WARP_THREADS, 0, st, g, C, src, theta, field, pad, padv, out));
  else if (gg->d == 2)
    ADVK_LAUNCH(K_warp_fwd, st, launch_pdl((warp_fwd_kernel<2, FIELD>), grid, WARP_THREADS, 0, st, g, C, src, theta, field, pad, interp, padv, out));
  else
    ADVK_LAUNCH(K_warp_fwd, st, launch_pdl((warp_fwd_kernel<3, FIELD>), grid, WARP_THREADS, 0, st, g, C, src, theta, field, pad, interp, padv, out));
  return check_launch("warp_fwd");
}

template <bool FIELD>
static int launch_bwd(const advk_geom* gg, int C, const float* g_out, const float* src,
                      const float* theta, const void* field, int pad, int interp,
                      const float* padv, float* g_src, float* g_theta, void* g_field,
                      void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(C >= 1 && src && g_out && (FIELD ? field != nullptr : theta != nullptr), "null pointer");
  ADVK_REQUIRE(pad >= 0 && pad <= 2 && interp >= 0 && interp <= 2, "bad pad/interp mode");
  ADVK_REQUIRE(interp != ADVK_INTERP_BICUBIC || gg->d == 2, "bicubic interpolation is 2-D only (like F.grid_sample)");
  dim3 grid(blocks_for(g.S, WARP_THREADS), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (interp == ADVK_INTERP_BICUBIC)
    ADVK_LAUNCH(K_warp_bwd, st, launch_pdl((warp_bicubic_bwd_kernel<FIELD>), grid, WARP_THREADS, 0, st, g, C, g_out, src, theta, field, pad, padv, g_src, g_theta, g_field));
  else if (gg->d == 2)
    ADVK_LAUNCH(K_warp_bwd, st, launch_pdl((warp_bwd_kernel<2, FIELD>), grid, WARP_THREADS, 0, st, g, C, g_out, src, theta, field, pad, interp, padv, g_src, g_theta, g_field));
  else
    ADVK_LAUNCH(K_warp_bwd, st, launch_pdl((warp_bwd_kernel<3, FIELD>), grid, WARP_THREADS, 0, st, g, C, g_out, src, theta, field, pad, interp, padv, g_src, g_theta, g_field));
  return check_launch("warp_bwd");
}

}  // namespace advk

using namespace advk;

extern "C" int advk_warp_affine_fwd(const advk_geom* g, int C, const float* src, const float* theta,
                                    int pad_mode, int interp, const float* pad_values, float* out,
                                    void* stream) {
  return launch_fwd<false>(g, C, src, theta, nullptr, pad_mode, interp, pad_values, out, stream);
}
extern "C" int advk_warp_affine_bwd(const advk_geom* g, int C, const float* g_out, const float* src,
                                    const float* theta, int pad_mode, int interp,
                                    const float* pad_values, float* g_src, float* g_theta,
                                    void* stream) {
  return launch_bwd<false>(g, C, g_out, src, theta, nullptr, pad_mode, interp, pad_values, g_src, g_theta, nullptr, stream);
}
extern "C" int advk_warp_field_fwd(const advk_geom* g, int C, const float* src, const void* field,
                                   int pad_mode, int interp, const float* pad_values, float* out,
                                   void* stream) {
  return launch_fwd<true>(g, C, src, nullptr, field, pad_mode, interp, pad_values, out, stream);
}
extern "C" int advk_warp_field_bwd(const advk_geom* g, int C, const float* g_out, const float* src,
                                   const void* field, int pad_mode, int interp,
                                   const float* pad_values, float* g_src, void* g_field,
                                   void* stream) {
  return launch_bwd<true>(g, C, g_out, src, nullptr, field, pad_mode, interp, pad_values, g_src, nullptr, g_field, stream);
}
