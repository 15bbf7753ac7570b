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
      cx = t[0] * bx + t[1] * by + t[2];
      cy = t[3] * bx + t[4] * by + t[5];
      cz = 0.f;
    } else {
      float bz = base_coord(z, g.D, 0.f);
      const float* t = theta + n * 12;
      cx = t[0] * bx + t[1] * by + t[2] * bz + t[3];
      cy = t[4] * bx + t[5] * by + t[6] * bz + t[7];
      cz = t[8] * bx + t[9] * by + t[10] * bz + t[11];
    }
    rx = cx; ry = cy; rz = cz;
  }
}

template <int DIM, bool FIELD>
__global__ void __launch_bounds__(WARP_THREADS)
warp_fwd_kernel(Dims g, int C, const float* __restrict__ src, const float* __restrict__ theta,
                const void* __restrict__ field, int pad, int interp,
                const float* __restrict__ padv, float* __restrict__ out) {
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

template <bool FIELD>
static int launch_fwd(const advk_geom* gg, int C, const float* src, const float* theta,
                      const void* field, int pad, int interp, const float* padv, float* out,
                      void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(C >= 1 && src && out && (FIELD ? field != nullptr : theta != nullptr), "null pointer");
  ADVK_REQUIRE(pad >= 0 && pad <= 2 && interp >= 0 && interp <= 1, "bad pad/interp mode");
  dim3 grid(blocks_for(g.S, WARP_THREADS), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (gg->d == 2)
    ADVK_LAUNCH(K_warp_fwd, st, warp_fwd_kernel<2, FIELD><<<grid, WARP_THREADS, 0, st>>>(g, C, src, theta, field, pad, interp, padv, out));
  else
    ADVK_LAUNCH(K_warp_fwd, st, warp_fwd_kernel<3, FIELD><<<grid, WARP_THREADS, 0, st>>>(g, C, src, theta, field, pad, interp, padv, out));
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
  ADVK_REQUIRE(pad >= 0 && pad <= 2 && interp >= 0 && interp <= 1, "bad pad/interp mode");
  dim3 grid(blocks_for(g.S, WARP_THREADS), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (gg->d == 2)
    ADVK_LAUNCH(K_warp_bwd, st, warp_bwd_kernel<2, FIELD><<<grid, WARP_THREADS, 0, st>>>(g, C, g_out, src, theta, field, pad, interp, padv, g_src, g_theta, g_field));
  else
    ADVK_LAUNCH(K_warp_bwd, st, warp_bwd_kernel<3, FIELD><<<grid, WARP_THREADS, 0, st>>>(g, C, g_out, src, theta, field, pad, interp, padv, g_src, g_theta, g_field));
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
