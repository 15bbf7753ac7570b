// advk_morph.cu -- AdvMorph deformation-field build (DemonsCompose, adv_morph.py:454-491) and its
// hand-derived backward (the reference relies on autograd; derivation in SURVEY.md section 9).
//
// Forward, per sign s (scale = s*epsilon):
//   u_lr  = G (*) (scale * v)                         low-res depthwise Gaussian, zero pad
//   phi_0 = base + upsample(u_lr) / 2^n               linear, align_corners=False
//   phi_k = phi_{k-1} o phi_{k-1}, k = 1..n            grid_sample(border, align_corners=True)
//   r     = compose_with_base(phi_n - phi_0 + base) - base        (quirk Q1: minus phi_0)
//   field = G (*) r + base                             full-res separable Gaussian, zero pad
// (the final clamp to [-1,1] is applied by the consumers on load).
//
// Fields are interleaved float2 / float4 per voxel so that every stencil corner is ONE vector
// load and every backward scatter ONE vector reduction (REDG.E.ADD.F32x2/x4).
#include <stdlib.h>
#include "advk_common.cuh"
#include "advk_adjoint.cuh"

namespace advk {

constexpr int KT = 9;   // Gaussian taps (sigma = 1 -> 2*int(4*sigma+0.5)+1, adv_morph.py:394-400)
constexpr int KR = 4;

struct MorphCfg {
  int Dl, Hl, Wl;
  float w[KT];
  float sD, sH, sW;   // ATen upsample scales in/out
};

static bool make_cfg(const advk_morph_cfg* c, const Dims& g, int d, MorphCfg& o) {
  if (!c || c->ktaps != KT) return false;
  o.Dl = c->lr[0]; o.Hl = c->lr[1]; o.Wl = c->lr[2];
  if (o.Dl < 1 || o.Hl < 1 || o.Wl < 1 || (d == 2 && o.Dl != 1)) return false;
  for (int i = 0; i < KT; ++i) o.w[i] = c->gauss[i];
  o.sD = (float)o.Dl / (float)g.D; o.sH = (float)o.Hl / (float)g.H; o.sW = (float)o.Wl / (float)g.W;
  return true;
}

template <int DIM> struct V;
template <> struct V<2> {
  typedef float2 T;
  static __device__ __forceinline__ T make(float x, float y, float) { return make_float2(x, y); }
  static __device__ __forceinline__ float z(const T&) { return 0.f; }
};
template <> struct V<3> {
  typedef float4 T;
  static __device__ __forceinline__ T make(float x, float y, float z) { return make_float4(x, y, z, 0.f); }
  static __device__ __forceinline__ float z(const T& v) { return v.z; }
};

// Hides how a per-sample base pointer was computed from the optimiser, so that base + 32-bit voxel offset
// stays ONE IMAD.WIDE per address instead of a re-derived 64-bit (n*S + offset)*sizeof chain.
template <typename P>
__device__ __forceinline__ P* opaque_ptr(P* p) {
  unsigned long long v = (unsigned long long)p;
  asm("" : "+l"(v));
  P* q = (P*)v;
  __builtin_assume(__isGlobal(q));       // keep LDG / REDG / STG (the barrier hides the address space too)
  return q;
}

// ---------------------------------------------------------------------------------------
// Low-res depthwise Gaussian (direct 9^d taps; the lattice is tiny: 16x16 / 8x8x8).
// in/out planar [NC][Dl][Hl][Wl]; out = G (*) (scale*in).  Self-adjoint -> also used backward.
template <int DIM>
__global__ void lowres_smooth_kernel(MorphCfg c, int NC, const float* __restrict__ in, float scale,
                                     float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)NC * lr) return;
  int x = (int)(idx % c.Wl);
  int y = (int)((idx / c.Wl) % c.Hl);
  int z = (int)((idx / ((i64)c.Wl * c.Hl)) % c.Dl);
  i64 nc = idx / lr;
  const float* src = in + nc * lr;
  float acc = 0.f;
  for (int kz = 0; kz < (DIM == 3 ? KT : 1); ++kz) {
    int zz = (DIM == 3) ? z + kz - KR : 0;
    if (zz < 0 || zz >= c.Dl) continue;
    float wz = (DIM == 3) ? c.w[kz] : 1.f;
    for (int ky = 0; ky < KT; ++ky) {
      int yy = y + ky - KR;
      if (yy < 0 || yy >= c.Hl) continue;
      float wzy = wz * c.w[ky];
      for (int kx = 0; kx < KT; ++kx) {
        int xx = x + kx - KR;
        if (xx < 0 || xx >= c.Wl) continue;
        acc += (wzy * c.w[kx]) * (scale * src[((i64)zz * c.Hl + yy) * c.Wl + xx]);
      }
    }
  }
  out[idx] = acc;
}

// Same smoothing, separable, one CTA per (sample, channel) with the whole lattice in shared memory
// (the default lattices are 16x16 / 8x8x8; the direct kernel above stays for lattices > 4096).
constexpr int LRS_MAX = 4096;
template <int DIM>
__global__ void __launch_bounds__(256)
lowres_smooth_smem_kernel(MorphCfg c, const float* __restrict__ in, float scale, float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float a[LRS_MAX];
  __shared__ float b[LRS_MAX];
  const int lr = c.Dl * c.Hl * c.Wl;
  const float* src = in + (i64)blockIdx.x * lr;
  for (int i = threadIdx.x; i < lr; i += blockDim.x) a[i] = scale * src[i];
  __syncthreads();
  for (int i = threadIdx.x; i < lr; i += blockDim.x) {          // along W
    int x = i % c.Wl, row = i - x;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) { int xx = x + k - KR; if (xx >= 0 && xx < c.Wl) acc += c.w[k] * a[row + xx]; }
    b[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < lr; i += blockDim.x) {          // along H
    int y = (i / c.Wl) % c.Hl;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) { int yy = y + k - KR; if (yy >= 0 && yy < c.Hl) acc += c.w[k] * b[i + (yy - y) * c.Wl]; }
    a[i] = acc;
  }
  __syncthreads();
  float* dst = out + (i64)blockIdx.x * lr;
  for (int i = threadIdx.x; i < lr; i += blockDim.x) {          // along D
    float acc;
    if (DIM == 3) {
      int z = i / (c.Wl * c.Hl);
      acc = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) { int zz = z + k - KR; if (zz >= 0 && zz < c.Dl) acc += c.w[k] * a[i + (zz - z) * c.Wl * c.Hl]; }
    } else acc = a[i];
    dst[i] = acc;
  }
}

template <int DIM>
static void launch_lowres_smooth(const MorphCfg& c, int NC, const float* in, float scale, float* out, cudaStream_t st) {
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  if (lr <= LRS_MAX)
    ADVK_LAUNCH(K_lowres_smooth, st, launch_pdl((lowres_smooth_smem_kernel<DIM>), NC, 256, 0, st, c, in, scale, out));
  else
    ADVK_LAUNCH(K_lowres_smooth, st, launch_pdl((lowres_smooth_kernel<DIM>), blocks_for(NC * lr, 128), 128, 0, st, c, NC, in, scale, out));
}

// upsampled velocity at one voxel; u_lr planar [N][DIM][lr]
template <int DIM>
__device__ __forceinline__ void upsample_u(const MorphCfg& c, const Dims& g, const float* __restrict__ u_lr,
                                           int n, int z, int y, int x, float (&u)[3]) {
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  UpAxis ux = up_axis(x, c.Wl, c.sW), uy = up_axis(y, c.Hl, c.sH);
  UpAxis uz;
  if (DIM == 3) uz = up_axis(z, c.Dl, c.sD);
  else { uz.i0 = uz.i1 = 0; uz.l0 = 1.f; uz.l1 = 0.f; }
#pragma unroll
  for (int ch = 0; ch < DIM; ++ch) {
    const float* s = u_lr + ((i64)n * DIM + ch) * lr;
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
      int zi = dz ? uz.i1 : uz.i0;
      float lz = dz ? uz.l1 : uz.l0;
      const float* r0 = s + ((i64)zi * c.Hl + uy.i0) * c.Wl;
      const float* r1 = s + ((i64)zi * c.Hl + uy.i1) * c.Wl;
      float a = uy.l0 * (ux.l0 * __ldg(r0 + ux.i0) + ux.l1 * __ldg(r0 + ux.i1)) +
                uy.l1 * (ux.l0 * __ldg(r1 + ux.i0) + ux.l1 * __ldg(r1 + ux.i1));
      acc += lz * a;
    }
    u[ch] = acc;
  }
  if (DIM == 2) u[2] = 0.f;
}

template <int DIM>
__global__ void __launch_bounds__(256)
init_phi0_kernel(MorphCfg c, Dims g, const float* __restrict__ u_lr, float inv2n,
                 typename V<DIM>::T* __restrict__ phi0) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  int x, y, z;
  voxel_xyz(g, (unsigned)p, x, y, z);
  float u[3];
  upsample_u<DIM>(c, g, u_lr, n, z, y, x, u);
  phi0[(i64)n * g.S + p] = V<DIM>::make(base_coord_s(x, g.W, g.stW) + u[0] * inv2n,
                                         base_coord_s(y, g.H, g.stH) + u[1] * inv2n,
                                         (DIM == 3 ? base_coord_s(z, g.D, g.stD) + u[2] * inv2n : 0.f));
}

template <int DIM>
__global__ void __launch_bounds__(256)
unorm2_kernel(MorphCfg c, Dims g, const float* __restrict__ u_lr, float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[32];
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  float v[1] = {0.f};
  if (p < g.S) {
    int x, y, z;
    voxel_xyz(g, (unsigned)p, x, y, z);
    float u[3];
    upsample_u<DIM>(c, g, u_lr, n, z, y, x, u);
    v[0] = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(out, v[0]);
}

// Row-wise version of the two kernels above: a CTA owns 8 whole rows of one z-plane.  The z and
// y interpolation weights are shared by a whole row, so the CTA first reduces the lattice to one line
// of Wl values per (channel, row) in shared memory and every voxel only interpolates along x
// (2 shared reads per channel instead of 2^d global ones; ~35 instead of ~150 instructions per voxel).
// phi0 == nullptr: only the sum of |u|^2 is produced (unorm2).
constexpr int LRW_MAX = 64;
template <int DIM>
__global__ void __launch_bounds__(256)
init_phi0_rows_kernel(MorphCfg c, Dims g, const float* __restrict__ u_lr, float inv2n,
                      typename V<DIM>::T* __restrict__ phi0, float* __restrict__ norm2) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float rowbuf[DIM][8][LRW_MAX];
  __shared__ float red[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int z = (DIM == 3) ? blockIdx.z % g.D : 0;
  const int n = (DIM == 3) ? blockIdx.z / g.D : blockIdx.z;
  const int y = blockIdx.y * 8 + ty;
  const int lr = c.Dl * c.Hl * c.Wl;
  UpAxis uz;
  if (DIM == 3) uz = up_axis(z, c.Dl, c.sD);
  else { uz.i0 = uz.i1 = 0; uz.l0 = 1.f; uz.l1 = 0.f; }
  for (int i = threadIdx.x; i < DIM * 8 * c.Wl; i += 256) {
    const int xl = i % c.Wl, r = (i / c.Wl) & 7, ch = i / (8 * c.Wl);
    const int yy = min(blockIdx.y * 8 + r, g.H - 1);
    const UpAxis uy = up_axis(yy, c.Hl, c.sH);
    const float* s = u_lr + ((i64)n * DIM + ch) * lr + xl;
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
      const int zi = dz ? uz.i1 : uz.i0;
      const float lz = dz ? uz.l1 : uz.l0;
      acc += lz * (uy.l0 * __ldg(s + (zi * c.Hl + uy.i0) * c.Wl) + uy.l1 * __ldg(s + (zi * c.Hl + uy.i1) * c.Wl));
    }
    rowbuf[ch][r][xl] = acc;
  }
  __syncthreads();
  float v[1] = {0.f};
  if (y < g.H) {
    // the CTA owns whole rows: the lattice lines above are shared by every x of the row
    const float by = base_coord_s(y, g.H, g.stH), bz = (DIM == 3) ? base_coord_s(z, g.D, g.stD) : 0.f;
    const i64 rowp = (i64)n * g.S + ((i64)z * g.H + y) * g.W;
    for (int x = tx; x < g.W; x += 32) {
      const UpAxis ux = up_axis(x, c.Wl, c.sW);
      float u[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < DIM; ++ch) u[ch] = ux.l0 * rowbuf[ch][ty][ux.i0] + ux.l1 * rowbuf[ch][ty][ux.i1];
      v[0] += u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
      if (phi0)
        phi0[rowp + x] = V<DIM>::make(base_coord_s(x, g.W, g.stW) + u[0] * inv2n, by + u[1] * inv2n,
                                      (DIM == 3 ? bz + u[2] * inv2n : 0.f));
    }
  }
  if (norm2) {
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(norm2, v[0]);
  }
}

template <int DIM>
static void launch_init_phi0(const MorphCfg& c, const Dims& g, const float* u_lr, float inv2n, void* phi0,
                             float* norm2, cudaStream_t st) {
  typedef typename V<DIM>::T T;
  const i64 gz = (DIM == 3) ? (i64)g.N * g.D : g.N;
  if (c.Wl <= LRW_MAX && gz <= 65535) {
    dim3 grid(1, (g.H + 7) / 8, (unsigned)gz);
    ADVK_LAUNCH(phi0 ? K_init_phi0 : K_unorm2, st, launch_pdl((init_phi0_rows_kernel<DIM>), grid, 256, 0, st, c, g, u_lr, inv2n, (T*)phi0, norm2));
  } else {
    dim3 grid(blocks_for(g.S, 256), g.N);
    if (phi0) ADVK_LAUNCH(K_init_phi0, st, launch_pdl((init_phi0_kernel<DIM>), grid, 256, 0, st, c, g, u_lr, inv2n, (T*)phi0));
    else ADVK_LAUNCH(K_unorm2, st, launch_pdl((unorm2_kernel<DIM>), grid, 256, 0, st, c, g, u_lr, norm2));
  }
}

// grid_sample(base, c, border) along one axis: linear interpolation of the linspace.
__device__ __forceinline__ float compose_axis(float c, int size, float step, float& mult) {
  float i = gs_index(c, size, ADVK_PAD_BORDER, mult);
  float f = floorf(i);
  int i0 = (int)f;
  float v = base_coord_s(i0, size, step) * ((f + 1.f) - i);
  if (i0 + 1 < size) v += base_coord_s(i0 + 1, size, step) * (i - f);
  return v;
}

// ---------------------------------------------------------------------------------------
// One squaring step: out(p) = sample(in, at in(p)), border padding, align_corners=True.
template <int DIM>
__global__ void __launch_bounds__(256, 8)
ss_step_kernel(Dims g, const typename V<DIM>::T* __restrict__ in, typename V<DIM>::T* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename V<DIM>::T T;
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;           // S < 2^31 (host-checked)
  if (p >= g.S) return;
  const T* src = in + (i64)n * g.S;
  const T f = __ldg(src + p);
  Axis ax = make_axis_border(f.x, g.W);
  Axis ay = make_axis_border(f.y, g.H);
  Axis az;
  if (DIM == 3) az = make_axis_border(V<DIM>::z(f), g.D);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; }
  const int HW = g.H * g.W;
  const T* c000 = src + (az.i0 * HW + ay.i0 * g.W + ax.i0);     // only dereferenced at in-bounds corners
  float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
  for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
    const bool vz = dz ? az.v1 : az.v0;
    const float wz = dz ? az.w1 : az.w0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const bool vy = dy ? ay.v1 : ay.v0;
      const float wy = dy ? ay.w1 : ay.w0;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const bool vx = dx ? ax.v1 : ax.v0;
        const float wx = dx ? ax.w1 : ax.w0;
        if (vx && vy && vz) {
          const T s = __ldg(c000 + (dz * HW + dy * g.W + dx));
          const float w = wx * wy * wz;
          ox += s.x * w; oy += s.y * w;
          if (DIM == 3) oz += V<DIM>::z(s) * w;
        }
      }
    }
  }
  out[(i64)n * g.S + p] = V<DIM>::make(ox, oy, oz);
}

// Same step with no validity predicates (see ss_step_bwd_lean_kernel: under border padding a corner
// outside the volume has weight exactly 0, so it is redirected to corner 0 of its axis and adds s * 0).
//
// EMIT_R (the LAST step of a 3-D build): also writes r = compose_with_base(phi_n - phi_0 + base) - base, the
// input of the full-resolution Gaussian (quirk Q1; smooth_in<DIM, 0> is the same arithmetic), so that the
// smoothing kernel stages ONE field by TMA and evaluates no input map at its halo voxels.
template <int DIM, bool EMIT_R>
__global__ void __launch_bounds__(256, 8)
ss_step_lean_kernel(Dims g, const typename V<DIM>::T* __restrict__ in, typename V<DIM>::T* __restrict__ out,
                    const typename V<DIM>::T* __restrict__ phi0, typename V<DIM>::T* __restrict__ rout) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename V<DIM>::T T;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;           // S < 2^31 (host-checked)
  if (p >= g.S) return;
  const i64 nb = (i64)blockIdx.y * g.S;
  const T* src = opaque_ptr(in + nb);
  const T f = __ldg(src + p);
  Axis ax = make_axis_border(f.x, g.W);
  Axis ay = make_axis_border(f.y, g.H);
  Axis az;
  if (DIM == 3) az = make_axis_border(V<DIM>::z(f), g.D);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; }
  const int HW = g.H * g.W;
  const int dxo = ax.v1 ? 1 : 0, dyo = ay.v1 ? g.W : 0, dzo = (DIM == 3 && az.v1) ? HW : 0;
  const T* c000 = src + (az.i0 * HW + ay.i0 * g.W + ax.i0);
  float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
  for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const T s = __ldg(c000 + (dz * dzo + dy * dyo + dx * dxo));
        const float w = (dx ? ax.w1 : ax.w0) * (dy ? ay.w1 : ay.w0) * (dz ? az.w1 : az.w0);
        ox += s.x * w; oy += s.y * w;
        if (DIM == 3) oz += V<DIM>::z(s) * w;
      }
    }
  }
  (out + nb)[p] = V<DIM>::make(ox, oy, oz);
  if (EMIT_R) {
    const T q = __ldg(opaque_ptr(phi0 + nb) + p);
    int x, y, z;
    voxel_xyz(g, (unsigned)p, x, y, z);
    float m;
    const float bx = base_coord_s(x, g.W, g.stW), by = base_coord_s(y, g.H, g.stH);
    const float rx = compose_axis((ox - q.x) + bx, g.W, g.stW, m) - bx;
    const float ry = compose_axis((oy - q.y) + by, g.H, g.stH, m) - by;
    float rz = 0.f;
    if (DIM == 3) {
      const float bz = base_coord_s(z, g.D, g.stD);
      rz = compose_axis((oz - V<DIM>::z(q)) + bz, g.D, g.stD, m) - bz;
    }
    (opaque_ptr(rout + nb))[p] = V<DIM>::make(rx, ry, rz);
  }
}

// A CTA of the tile adjoint owns a 32 (x) x 8 (y) tile of one z plane, one warp per row: block index ->
// (tile x, tile y, z) by multiply-high division.
struct TileMap {
  FastDiv ftx, fty;     // tiles along x, tiles along y
  int tx, ty;
};
static TileMap make_tilemap(const Dims& g) {
  TileMap tm;
  tm.tx = (g.W + 31) / 32; tm.ty = (g.H + 7) / 8;
  tm.ftx = make_fastdiv((unsigned)tm.tx); tm.fty = make_fastdiv((unsigned)tm.ty);
  return tm;
}

// Backward of one squaring step phi_k = phi_{k-1} o phi_{k-1}:
//     dL/dphi_{k-1}(y) = sum_x g_k(x) * w(phi_{k-1}(x), y)            scatter adjoint of the gather
//                      + mult * < d(sample)/d(coord) at y , g_k(y) >   spatial-Jacobian term
// accumulated with vector REDs into ONE buffer `out` that must be zero on entry.
//
// ss_step_bwd_plain_kernel is the straightforward form (one RED per corner, validity predicates, the
// caller zeroes `out` with a memset node): the A/B predecessor of the lean kernel below and the
// independent implementation the parity tests compare it with (advk_morph_tune bit 0).
// What was measured and dropped on B200 (shared-memory tiles, warp-box hand-offs along y and z, planar
// field levels, stores before the REDs) is recorded in DESIGN.md section 3.2 and profiles/r01l_ssb_taken_apart.md.
template <int DIM>
__global__ void __launch_bounds__(256)
ss_step_bwd_plain_kernel(Dims g, const typename V<DIM>::T* __restrict__ phi_prev, const typename V<DIM>::T* __restrict__ up,
                         typename V<DIM>::T* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename V<DIM>::T T;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;           // S < 2^31 (host-checked)
  if (p >= g.S) return;
  const i64 nb = (i64)blockIdx.y * g.S;
  const T* src = phi_prev + nb;
  T* dst = out + nb;
  const T f = __ldg(src + p), go = up[nb + p];
  Axis ax = make_axis_border(f.x, g.W);
  Axis ay = make_axis_border(f.y, g.H);
  Axis az;
  if (DIM == 3) az = make_axis_border(V<DIM>::z(f), g.D);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; az.mult = 0.f; }
  const int HW = g.H * g.W;
  const float gx = go.x, gy = go.y, gz = V<DIM>::z(go);
  float jx = 0.f, jy = 0.f, jz = 0.f;
#pragma unroll
  for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        if (!((dx ? ax.v1 : ax.v0) && (dy ? ay.v1 : ay.v0) && (dz ? az.v1 : az.v0))) continue;
        const float wx = dx ? ax.w1 : ax.w0, wy = dy ? ay.w1 : ay.w0, wz = dz ? az.w1 : az.w0;
        const int a = (az.i0 + dz) * HW + (ay.i0 + dy) * g.W + ax.i0 + dx;
        const T s = __ldg(src + a);
        const float dot = s.x * gx + s.y * gy + V<DIM>::z(s) * gz;
        jx += (dx ? dot : -dot) * (wy * wz);
        jy += (dy ? dot : -dot) * (wx * wz);
        if (DIM == 3) jz += (dz ? dot : -dot) * (wx * wy);
        const float w = wx * wy * wz;
        atomicAdd(dst + a, V<DIM>::make(gx * w, gy * w, gz * w));
      }
    }
  }
  atomicAdd(dst + p, V<DIM>::make(jx * ax.mult, jy * ay.mult, jz * az.mult));
}

// Lean adjoint.  Taking the plain kernel apart on B200 (profiles/r01l_ssb_taken_apart.md:
// 50 us per launch at 128^3; 35 us with the REDs turned into register adds, 42 us with the gathers turned
// into register moves, 29 us with both gone -- and 48 us with the REDs turned into plain stores) shows that
// the skeleton, i.e. instruction issue, is the larger part of a launch, so this variant spends fewer
// instructions on the same arithmetic:
//   * per-sample base pointers are made opaque to the optimiser, so that a corner address is one
//     IMAD.WIDE of a 32-bit voxel offset instead of a re-derived 64-bit (n*S + offset)*16 chain;
//   * the x hand-off ships the upstream value once and ONE weight per corner row (the receiver forms
//     g*w0 + g_prev*w1_prev with FMAs) instead of a 3-vector per row, on one address test per voxel (a
//     neighbour whose corner (0,0,0) is one voxel further sees the same rows with the same validity);
//   * the Jacobian term is built from the 2^d dot products <phi(corner), g> by separable differences
//     (interpolate along x, difference along y, ...) instead of three weighted sums over all corners.
// Same result as the plain kernel up to fp32 summation order.
template <int DIM, bool ZS>
__global__ void __launch_bounds__(256)
ss_step_bwd_lean_kernel(Dims g, const typename V<DIM>::T* __restrict__ phi_prev, typename V<DIM>::T* up,
                        typename V<DIM>::T* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename V<DIM>::T T;
  constexpr int NZ = DIM == 3 ? 2 : 1;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;           // S < 2^31 (host-checked)
  const bool live = p < g.S;
  const i64 nb = (i64)blockIdx.y * g.S;
  const T* src = opaque_ptr(phi_prev + nb);
  T* dst = opaque_ptr(out + nb);
  T* upn = opaque_ptr(up + nb);
  T f = V<DIM>::make(0.f, 0.f, 0.f), go = V<DIM>::make(0.f, 0.f, 0.f);
  if (live) {
    f = __ldg(src + p);
    go = upn[p];
  }
  Axis ax = make_axis_border(f.x, g.W);
  Axis ay = make_axis_border(f.y, g.H);
  Axis az;
  if (DIM == 3) az = make_axis_border(V<DIM>::z(f), g.D);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; az.mult = 0.f; }
  const int HW = g.H * g.W;
  const float gx = go.x, gy = go.y, gz = V<DIM>::z(go);
  // Border padding: corner 0 of an axis is always inside, and corner 1 is outside exactly when the clipped
  // coordinate sits ON the last voxel -- where its weight (x - floor x) is exactly 0 and d(index)/d(coord)
  // (`mult`) is 0 as well.  So an outside corner is simply redirected to corner 0 of that axis: whatever
  // is gathered there meets a zero weight or a zero `mult`, whatever is scattered there is a zero.  No
  // validity predicate is left on any gather or RED (a dead lane of the last CTA works on zeros at the
  // in-bounds voxel its zero coordinate points to, and skips the REDs).
  const int dxo = ax.v1 ? 1 : 0, dyo = ay.v1 ? g.W : 0, dzo = (DIM == 3 && az.v1) ? HW : 0;
  const int a000 = az.i0 * HW + ay.i0 * g.W + ax.i0;
  const T* s000 = src + a000;
  // gathers -> d = <phi(corner), g>
  float d[NZ][2][2];
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const T* r = s000 + (dz * dzo + dy * dyo);
      const T s0 = __ldg(r), s1 = __ldg(r + dxo);
      d[dz][dy][0] = s0.x * gx + s0.y * gy + V<DIM>::z(s0) * gz;
      d[dz][dy][1] = s1.x * gx + s1.y * gy + V<DIM>::z(s1) * gz;
    }
  }
  // Jacobian term: d(sample)/d(coord) = differences of the interpolated dot products
  float jx, jy, jz = 0.f;
  {
    float jxz[NZ], eyd[NZ], ey[NZ];
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz) {
      const float e0 = ax.w0 * d[dz][0][0] + ax.w1 * d[dz][0][1], e1 = ax.w0 * d[dz][1][0] + ax.w1 * d[dz][1][1];
      jxz[dz] = ay.w0 * (d[dz][0][1] - d[dz][0][0]) + ay.w1 * (d[dz][1][1] - d[dz][1][0]);
      eyd[dz] = e1 - e0;
      ey[dz] = ay.w0 * e0 + ay.w1 * e1;
    }
    if (DIM == 3) {
      jx = az.w0 * jxz[0] + az.w1 * jxz[NZ - 1];
      jy = az.w0 * eyd[0] + az.w1 * eyd[NZ - 1];
      jz = ey[NZ - 1] - ey[0];
    } else {
      jx = jxz[0]; jy = eyd[0];
    }
  }
  // x hand-off (every shuffle is executed by all lanes): lane i's corner-1 column is lane i+1's corner-0
  // column when that lane's corner (0,0,0) is one voxel further -- then both see the same rows.
  const int next = __shfl_down_sync(FULL, live ? a000 : -1, 1);
  const bool hand = live && ax.v1 && lane < 31 && next == a000 + 1;
  const float px = __shfl_up_sync(FULL, gx, 1), py = __shfl_up_sync(FULL, gy, 1);
  const float pz = DIM == 3 ? __shfl_up_sync(FULL, gz, 1) : 0.f;
  float w0[NZ][2], w1[NZ][2], ws[NZ][2];
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float wr = (dy ? ay.w1 : ay.w0) * (dz ? az.w1 : az.w0);
      w0[dz][dy] = ax.w0 * wr; w1[dz][dy] = ax.w1 * wr;
      const float got = __shfl_up_sync(FULL, hand ? w1[dz][dy] : 0.f, 1);
      ws[dz][dy] = lane ? got : 0.f;
    }
  }
  if (live) {
    T* d000 = dst + a000;
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float a = w0[dz][dy], b = ws[dz][dy];
        atomicAdd(d000 + (dz * dzo + dy * dyo), V<DIM>::make(gx * a + px * b, gy * a + py * b, gz * a + pz * b));
      }
    }
    if (ax.v1 && !hand) {                                           // row ends and non-smooth spots only
#pragma unroll
      for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const float c = w1[dz][dy];
          atomicAdd(d000 + (dz * dzo + dy * dyo + 1), V<DIM>::make(gx * c, gy * c, gz * c));
        }
      }
    }
    atomicAdd(dst + p, V<DIM>::make(jx * ax.mult, jy * ay.mult, jz * az.mult));
    if (ZS) upn[p] = V<DIM>::make(0.f, 0.f, 0.f);               // AFTER the REDs: a store to a line whose load is in flight stalls the LSU
  }
}

// Tile adjoint (the default where the tiles are full; advk_morph_tune bits 3 / 5 force it on / off): the lean
// kernel on a 32 (x) x 8 (y) tile of one z plane, one warp per row,
// with a second hand-off ALONG Y through shared memory.  The RED payload is what keeps the L1 -> crossbar port
// busy in the lean kernel (6.1 M sectors = 194 MB per launch at 128^3, 70 % of the port's cycles): the corner
// rows (y0+1, z0 | z0+1) of the voxel (x, y) are the corner rows (y0, ...) of the voxel (x, y+1) whenever that
// voxel's corner (0,0,0) lies one row further (always, for a smooth field), so the upper rows travel to the
// thread one row up as ONE float4 each (value xyz + the target voxel offset in w) and are added to its lower-row
// REDs: 2 corner REDs + the Jacobian RED per voxel instead of 4 + 1 (the last row of a tile keeps its own upper
// rows: 3.25 on average).  A receiver whose address does not match -- or that lies outside the volume -- issues
// the sender's REDs at the sender's addresses, so the result does not depend on the field being smooth.
// Measured at 128^3 (gpurun_out/r02w): 715 against 733 us per 16 launches, 441.3 against 438.3 it/s -- the RED
// sectors drop by a third, the launch by 2.4 %: with the REDs thinned the kernel sits on its two dependent
// round trips (phi(p), g(p) -> corners) at 50 % occupancy.
// what a thread of the tile adjoint does with its voxel once phi(p) = f and g(p) = go are in registers
// (hand_s: this iteration's [NZ][8][32] hand-off buffer; contains the CTA's one barrier)
template <int DIM, bool ZS>
__device__ __forceinline__ void ssb_tile_voxel(const Dims& g, const typename V<DIM>::T* src, typename V<DIM>::T* dst,
                                               typename V<DIM>::T* upn, const typename V<DIM>::T f,
                                               const typename V<DIM>::T go, const bool live, const int p,
                                               const int lane, const int row, float4 (*hand_s)[8][32]) {
  typedef typename V<DIM>::T T;
  constexpr int NZ = DIM == 3 ? 2 : 1;
  const unsigned FULL = 0xffffffffu;
  const int HW = g.H * g.W;
  Axis ax = make_axis_border(f.x, g.W);
  Axis ay = make_axis_border(f.y, g.H);
  Axis az;
  if (DIM == 3) az = make_axis_border(V<DIM>::z(f), g.D);
  else { az.i0 = 0; az.w0 = 1.f; az.w1 = 0.f; az.v0 = true; az.v1 = false; az.mult = 0.f; }
  const float gx = go.x, gy = go.y, gz = V<DIM>::z(go);
  // outside corners are redirected to corner 0 of their axis (weight / mult exactly 0 there): see the lean kernel
  const int dxo = ax.v1 ? 1 : 0, dyo = ay.v1 ? g.W : 0, dzo = (DIM == 3 && az.v1) ? HW : 0;
  const int a000 = az.i0 * HW + ay.i0 * g.W + ax.i0;
  // x hand-off by weight (shuffles, executed by all lanes)
  const int next = __shfl_down_sync(FULL, live ? a000 : -1, 1);
  const bool hand = live && ax.v1 && lane < 31 && next == a000 + 1;
  const float px = __shfl_up_sync(FULL, gx, 1), py = __shfl_up_sync(FULL, gy, 1);
  const float pz = DIM == 3 ? __shfl_up_sync(FULL, gz, 1) : 0.f;
  float w1[NZ][2];
  float vx[NZ][2], vy[NZ][2], vz[NZ][2];
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float wr = (dy ? ay.w1 : ay.w0) * (dz ? az.w1 : az.w0);
      const float a = ax.w0 * wr;
      w1[dz][dy] = ax.w1 * wr;
      const float got = __shfl_up_sync(FULL, hand ? w1[dz][dy] : 0.f, 1);
      const float bw = lane ? got : 0.f;
      vx[dz][dy] = gx * a + px * bw; vy[dz][dy] = gy * a + py * bw; vz[dz][dy] = gz * a + pz * bw;
    }
  }
  // y hand-off: the upper rows go to the thread one row up (the last row of the tile keeps them)
  const bool send = live && ay.v1 && row < 7;
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz)
    hand_s[dz][row][lane] = make_float4(vx[dz][1], vy[dz][1], vz[dz][1],
                                        __int_as_float(send ? a000 + dz * dzo + dyo : -1));
  __syncthreads();
  T* d000 = dst + a000;
  if (row) {
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz) {
      const float4 h = hand_s[dz][row - 1][lane];
      const int ha = __float_as_int(h.w);
      if (ha >= 0) {
        if (live && ha == a000 + dz * dzo) { vx[dz][0] += h.x; vy[dz][0] += h.y; vz[dz][0] += h.z; }
        else atomicAdd(dst + ha, V<DIM>::make(h.x, h.y, h.z));     // non-smooth spot / receiver outside the volume
      }
    }
  }
  if (live) {
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz) {
      atomicAdd(d000 + dz * dzo, V<DIM>::make(vx[dz][0], vy[dz][0], vz[dz][0]));
      if (!send) atomicAdd(d000 + (dz * dzo + dyo), V<DIM>::make(vx[dz][1], vy[dz][1], vz[dz][1]));
    }
    if (ax.v1 && !hand) {                                           // row ends and non-smooth spots only
#pragma unroll
      for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const float c = w1[dz][dy];
          atomicAdd(d000 + (dz * dzo + dy * dyo + 1), V<DIM>::make(gx * c, gy * c, gz * c));
        }
      }
    }
  }
  // gathers -> d = <phi(corner), g>, Jacobian term by separable differences (as in the lean kernel)
  const T* s000 = src + a000;
  float d[NZ][2][2];
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const T* r = s000 + (dz * dzo + dy * dyo);
      const T s0 = __ldg(r), s1 = __ldg(r + dxo);
      d[dz][dy][0] = s0.x * gx + s0.y * gy + V<DIM>::z(s0) * gz;
      d[dz][dy][1] = s1.x * gx + s1.y * gy + V<DIM>::z(s1) * gz;
    }
  }
  float jx, jy, jz = 0.f;
  {
    float jxz[NZ], eyd[NZ], ey[NZ];
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz) {
      const float e0 = ax.w0 * d[dz][0][0] + ax.w1 * d[dz][0][1], e1 = ax.w0 * d[dz][1][0] + ax.w1 * d[dz][1][1];
      jxz[dz] = ay.w0 * (d[dz][0][1] - d[dz][0][0]) + ay.w1 * (d[dz][1][1] - d[dz][1][0]);
      eyd[dz] = e1 - e0;
      ey[dz] = ay.w0 * e0 + ay.w1 * e1;
    }
    if (DIM == 3) {
      jx = az.w0 * jxz[0] + az.w1 * jxz[NZ - 1];
      jy = az.w0 * eyd[0] + az.w1 * eyd[NZ - 1];
      jz = ey[NZ - 1] - ey[0];
    } else {
      jx = jxz[0]; jy = eyd[0];
    }
  }
  if (live) {
    atomicAdd(dst + p, V<DIM>::make(jx * ax.mult, jy * ay.mult, jz * az.mult));
    if (ZS) upn[p] = V<DIM>::make(0.f, 0.f, 0.f);               // AFTER the REDs (see the lean kernel)
  }
}

template <int DIM, bool ZS>
__global__ void __launch_bounds__(256)
ss_step_bwd_tile_kernel(Dims g, TileMap tm, const typename V<DIM>::T* __restrict__ phi_prev, typename V<DIM>::T* up,
                        typename V<DIM>::T* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename V<DIM>::T T;
  constexpr int NZ = DIM == 3 ? 2 : 1;
  __shared__ float4 hand_s[NZ][8][32];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const unsigned b = blockIdx.x;
  const unsigned q = fast_div(b, tm.ftx);                         // b = (z * ty + tyi) * tx + txi
  const int txi = (int)(b - q * (unsigned)tm.tx);
  const unsigned z = fast_div(q, tm.fty);
  const int tyi = (int)(q - z * (unsigned)tm.ty);
  const int x = txi * 32 + lane, y = tyi * 8 + row;
  const bool live = x < g.W && y < g.H;
  const int p = live ? ((int)z * g.H + y) * g.W + x : 0;
  const i64 nb = (i64)blockIdx.y * g.S;
  const T* src = opaque_ptr(phi_prev + nb);
  T* dst = opaque_ptr(out + nb);
  T* upn = opaque_ptr(up + nb);
  T f = V<DIM>::make(0.f, 0.f, 0.f), go = V<DIM>::make(0.f, 0.f, 0.f);
  if (live) {
    f = __ldg(src + p);
    go = upn[p];
  }
  ssb_tile_voxel<DIM, ZS>(g, src, dst, upn, f, go, live, p, lane, row, hand_s);
}

// advk_morph_tune / ADVK_SSB_MODE: bit 0 = the plain predecessors of the lean squaring-step kernels
// (predicated forward step, one-RED-per-corner adjoint with memset nodes); bit 1 = the two-launch predecessor
// (smooth3d_xy + smooth3d_z) of the TMA-staged 3-D smoothing kernel (advk_smooth_tma.cuh); bit 2 = the lean
// adjoint zeroes its consumed buffer itself (predecessor of the side-stream memsets, field_bwd); bit 3 = the tile
// adjoint with the y hand-off through shared memory (ss_step_bwd_tile_kernel) whatever the geometry (default: where
// the 32 x 8 tiles are at least 97 % full); bit 5 = the lean linear adjoint whatever the geometry.  Default 0.
// Measured and removed (gpurun_out/r02w, r02y): the forward step on 32 x 8 tiles (1.16
// instead of 1.5 L1 fills per voxel: 2 % slower), and the x1 corners of either step taken from the neighbour lane
// by shuffle (4 gathers instead of 8: forward +1.1 %, adjoint +1.8 % of the whole iteration, and not bit-identical
// because the weight products associate differently); persistent forms that load phi(p), g(p) of their NEXT tile
// before they work on the current one (the adjoint at 64 registers / 4 CTAs per SM: 745 against 713 us per 16
// launches; the forward step at 40 registers / 6 CTAs: no change -- gpurun_out/r02D).
static int g_ssb_mode = -1;
static int ssb_mode() {
  if (g_ssb_mode < 0) {
    const char* e = getenv("ADVK_SSB_MODE");
    g_ssb_mode = e ? (atoi(e) & 63) : 0;
  }
  return g_ssb_mode;
}

// A side stream per device for work that can run BESIDE the calling stream (the zeroing of scatter targets):
// forked and joined with events only, so it is legal under stream capture and never synchronises the host.
struct SideStream {
  cudaStream_t s;
  cudaEvent_t ev[16];
  int next;
  cudaEvent_t event() { cudaEvent_t e = ev[next]; next = (next + 1) & 15; return e; }
  // memset `p` on the side stream after everything enqueued on `st` so far; returns the event to wait for
  cudaEvent_t zero_beside(cudaStream_t st, void* p, size_t bytes) {
    cudaEvent_t a = event(), b = event();
    cudaEventRecord(a, st);
    cudaStreamWaitEvent(s, a, 0);
    cudaMemsetAsync(p, 0, bytes, s);
    cudaEventRecord(b, s);
    return b;
  }
};
static SideStream* side_stream(cudaStream_t st) {
  static SideStream* tab[64] = {nullptr};
  static bool failed[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || failed[dev]) return nullptr;
  if (!tab[dev]) {
    // creating a stream is not a capturable operation: a first use inside a capture keeps the in-kernel
    // zeroing (the solver's graph loop runs an eager warm-up iteration before it captures)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      (void)cudaGetLastError();
      return nullptr;
    }
    SideStream* n = new SideStream();
    n->next = 0;
    bool ok = cudaStreamCreateWithFlags(&n->s, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 16; ++i) ok = cudaEventCreateWithFlags(&n->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); failed[dev] = true; delete n; return nullptr; }
    tab[dev] = n;
  }
  return tab[dev];
}

template <int DIM>
static void launch_ss_step_bwd(const Dims& g, const typename V<DIM>::T* phi_prev, typename V<DIM>::T* up,
                               typename V<DIM>::T* out, bool may_zero_up, cudaStream_t st) {
  dim3 grid(blocks_for(g.S, 256), g.N);
  // Tile adjoint by default where the tiles are (almost) full -- it buys 2.4 % per launch at 128^3, less than a
  // partly filled tile column costs; bit 3 forces it, bit 5 forces the lean linear kernel.
  const TileMap tm = make_tilemap(g);
  const bool tile_fits = (i64)g.W * g.H * 100 >= (i64)97 * (tm.tx * 32) * (tm.ty * 8);
  if ((ssb_mode() & 1) == 0 && ((ssb_mode() & 8) || (tile_fits && !(ssb_mode() & 32)))) {
    dim3 tg((unsigned)((i64)tm.tx * tm.ty * g.D), g.N);
    if (may_zero_up) ADVK_LAUNCH(K_ss_step_bwd, st, (launch_pdl((ss_step_bwd_tile_kernel<DIM, true>), tg, 256, 0, st, g, tm, phi_prev, up, out)));
    else ADVK_LAUNCH(K_ss_step_bwd, st, (launch_pdl((ss_step_bwd_tile_kernel<DIM, false>), tg, 256, 0, st, g, tm, phi_prev, up, out)));
    return;
  }
  if (ssb_mode() & 1) ADVK_LAUNCH(K_ss_step_bwd, st, (launch_pdl((ss_step_bwd_plain_kernel<DIM>), grid, 256, 0, st, g, phi_prev, up, out)));
  else if (may_zero_up) ADVK_LAUNCH(K_ss_step_bwd, st, (launch_pdl((ss_step_bwd_lean_kernel<DIM, true>), grid, 256, 0, st, g, phi_prev, up, out)));
  else ADVK_LAUNCH(K_ss_step_bwd, st, (launch_pdl((ss_step_bwd_lean_kernel<DIM, false>), grid, 256, 0, st, g, phi_prev, up, out)));
}

// ---------------------------------------------------------------------------------------
// Full-res separable Gaussian with fused pre/post maps.
//   MODE 0 (forward):  in  = compose_with_base(phi_n - phi_0 + base) - base
//                      out = smooth + base                            -> field_out
//   MODE 1 (backward): in  = g_field * [ -1 <= field <= 1 ]
//                      out = smooth * [ 0 < pix(phi_n - phi_0 + base) < size-1 ]   -> g_off
// A, B are the two fields the input map reads at halo positions (phi_n, phi_0 | g_field, field);
// C, D are read at the centre by the backward output map (phi_n, phi_0).
template <int DIM>
struct SmoothArgs {
  const typename V<DIM>::T* A;
  const typename V<DIM>::T* B;
  const typename V<DIM>::T* C;
  const typename V<DIM>::T* D;
  typename V<DIM>::T* out;
  float w[KT];
};

template <int DIM, int MODE>
__device__ __forceinline__ void smooth_in(const SmoothArgs<DIM>& a, const Dims& g, i64 idx, int z,
                                          int y, int x, float (&r)[3]) {
  typedef typename V<DIM>::T T;
  T p = a.A[idx], q = a.B[idx];
  if (MODE == 0) {
    float m;
    float bx = base_coord_s(x, g.W, g.stW), by = base_coord_s(y, g.H, g.stH);
    r[0] = compose_axis((p.x - q.x) + bx, g.W, g.stW, m) - bx;
    r[1] = compose_axis((p.y - q.y) + by, g.H, g.stH, m) - by;
    if (DIM == 3) {
      float bz = base_coord_s(z, g.D, g.stD);
      r[2] = compose_axis((V<DIM>::z(p) - V<DIM>::z(q)) + bz, g.D, g.stD, m) - bz;
    } else r[2] = 0.f;
  } else {
    r[0] = (q.x >= -1.f && q.x <= 1.f) ? p.x : 0.f;
    r[1] = (q.y >= -1.f && q.y <= 1.f) ? p.y : 0.f;
    r[2] = (DIM == 3 && V<DIM>::z(q) >= -1.f && V<DIM>::z(q) <= 1.f) ? V<DIM>::z(p) : 0.f;
  }
}

template <int DIM, int MODE>
__device__ __forceinline__ void smooth_out(const SmoothArgs<DIM>& a, const Dims& g, i64 idx, int z,
                                           int y, int x, const float (&s)[3]) {
  typedef typename V<DIM>::T T;
  if (MODE == 0) {
    a.out[idx] = V<DIM>::make(s[0] + base_coord_s(x, g.W, g.stW), s[1] + base_coord_s(y, g.H, g.stH),
                              DIM == 3 ? s[2] + base_coord_s(z, g.D, g.stD) : 0.f);
  } else {
    T p = a.C[idx], q = a.D[idx];
    float mx, my, mz = 0.f;
    gs_index((p.x - q.x) + base_coord_s(x, g.W, g.stW), g.W, ADVK_PAD_BORDER, mx);
    gs_index((p.y - q.y) + base_coord_s(y, g.H, g.stH), g.H, ADVK_PAD_BORDER, my);
    if (DIM == 3) gs_index((V<DIM>::z(p) - V<DIM>::z(q)) + base_coord_s(z, g.D, g.stD), g.D, ADVK_PAD_BORDER, mz);
    a.out[idx] = V<DIM>::make(mx != 0.f ? s[0] : 0.f, my != 0.f ? s[1] : 0.f, mz != 0.f ? s[2] : 0.f);
  }
}

constexpr int S2_TX = 32, S2_TY = 16, S2_THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(S2_THREADS)
smooth2d_kernel(Dims g, SmoothArgs<2> a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float s_in[2][S2_TY + 2 * KR][S2_TX + 2 * KR];
  __shared__ float s_tmp[2][S2_TY + 2 * KR][S2_TX];
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * S2_TX, y0 = blockIdx.y * S2_TY;
  constexpr int IW = S2_TX + 2 * KR, IH = S2_TY + 2 * KR;
  for (int i = threadIdx.x; i < IW * IH; i += S2_THREADS) {
    int ly = i / IW, lx = i % IW;
    int gy = y0 + ly - KR, gx = x0 + lx - KR;
    float r[3] = {0.f, 0.f, 0.f};
    if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W)
      smooth_in<2, MODE>(a, g, (i64)n * g.S + (i64)gy * g.W + gx, 0, gy, gx, r);
    s_in[0][ly][lx] = r[0]; s_in[1][ly][lx] = r[1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S2_TX * IH; i += S2_THREADS) {
    int ly = i / S2_TX, lx = i % S2_TX;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) { a0 += a.w[k] * s_in[0][ly][lx + k]; a1 += a.w[k] * s_in[1][ly][lx + k]; }
    s_tmp[0][ly][lx] = a0; s_tmp[1][ly][lx] = a1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S2_TX * S2_TY; i += S2_THREADS) {
    int ly = i / S2_TX, lx = i % S2_TX;
    int gy = y0 + ly, gx = x0 + lx;
    if (gy >= g.H || gx >= g.W) continue;
    float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KT; ++k) { s[0] += a.w[k] * s_tmp[0][ly + k][lx]; s[1] += a.w[k] * s_tmp[1][ly + k][lx]; }
    smooth_out<2, MODE>(a, g, (i64)n * g.S + (i64)gy * g.W + gx, 0, gy, gx, s);
  }
}

// 3-D: two launches.
//   smooth3d_xy: one CTA per 64 x 16 tile of ONE z-plane: input map (compose-with-base | clamp mask)
//     on tile + 4-voxel halo into shared memory, x pass and y pass with register windows (every
//     thread produces 4 consecutive outputs from 12 inputs: 3 LDS.128 per 36 FMA), result t -> `tmp`.
//   smooth3d_z: one thread per (x,y) column and 16-plane chunk, 9-plane register ring (fully unrolled),
//     output map (re-add base | border mask) fused.
// The previous single kernel (xy tile marching in z) staged every plane 1.5 x 1.9 times through the
// 90-instruction input map and issued 960 warp instructions per 32 voxels (profiles/r01p); this
// split needs ~300 and exposes D times more CTAs.
constexpr int SX_TX = 64, SX_TY = 16, SX_THREADS = 256;
constexpr int SX_IW = SX_TX + 2 * KR, SX_IH = SX_TY + 2 * KR;
constexpr int SZ_CHUNK = 16;

template <int MODE>
__global__ void __launch_bounds__(SX_THREADS)
smooth3d_xy_kernel(Dims g, SmoothArgs<3> a, float4* __restrict__ tmp) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ __align__(16) float s_in[3][SX_IH][SX_IW];
  __shared__ __align__(16) float s_tmp[3][SX_IH][SX_TX];
  const int z = blockIdx.z % g.D, n = blockIdx.z / g.D;
  const int x0 = blockIdx.x * SX_TX, y0 = blockIdx.y * SX_TY;
  const i64 pb = (i64)n * g.S + (i64)z * g.H * g.W;
  for (int i = threadIdx.x; i < SX_IW * SX_IH; i += SX_THREADS) {
    const int ly = i / SX_IW, lx = i - ly * SX_IW;
    const int gy = y0 + ly - KR, gx = x0 + lx - KR;
    float r[3] = {0.f, 0.f, 0.f};
    if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) smooth_in<3, MODE>(a, g, pb + (i64)gy * g.W + gx, z, gy, gx, r);
    s_in[0][ly][lx] = r[0]; s_in[1][ly][lx] = r[1]; s_in[2][ly][lx] = r[2];
  }
  __syncthreads();
  float w[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) w[k] = a.w[k];
  // x pass: task = (component, row, run of 4 outputs)
  constexpr int RUNS = SX_TX / 4;
  for (int task = threadIdx.x; task < 3 * SX_IH * RUNS; task += SX_THREADS) {
    const int c = task / (SX_IH * RUNS);
    const int rem = task - c * (SX_IH * RUNS);
    const int ly = rem / RUNS, xs = (rem - ly * RUNS) * 4;
    const float4 i0 = *reinterpret_cast<const float4*>(&s_in[c][ly][xs]);
    const float4 i1 = *reinterpret_cast<const float4*>(&s_in[c][ly][xs + 4]);
    const float4 i2 = *reinterpret_cast<const float4*>(&s_in[c][ly][xs + 8]);
    const float in[12] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w};
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KT; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += w[k] * in[j + k];
    }
    *reinterpret_cast<float4*>(&s_tmp[c][ly][xs]) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // y pass: thread = (x, group of 4 rows), all three components
  const int tx = threadIdx.x % SX_TX, yg = threadIdx.x / SX_TX;
  float o[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float in[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) in[i] = s_tmp[c][yg * 4 + i][tx];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k) acc += w[k] * in[j + k];
      o[c][j] = acc;
    }
  }
  const int gx = x0 + tx;
  if (gx < g.W) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + yg * 4 + j;
      if (gy < g.H) tmp[pb + (i64)gy * g.W + gx] = make_float4(o[0][j], o[1][j], o[2][j], 0.f);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
smooth3d_z_kernel(Dims g, SmoothArgs<3> a, const float4* __restrict__ tmp, int nzc) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z / nzc, zb = (blockIdx.z % nzc) * SZ_CHUNK;
  if (x >= g.W || y >= g.H) return;
  const i64 HW = (i64)g.H * g.W;
  const i64 col = (i64)n * g.S + (i64)y * g.W + x;
  float w[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) w[k] = a.w[k];
  float rx[KT], ry[KT], rz[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) rx[k] = ry[k] = rz[k] = 0.f;
#pragma unroll
  for (int i = 0; i < SZ_CHUNK + 2 * KR; ++i) {
    const int pz = zb - KR + i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pz >= 0 && pz < g.D) v = __ldg(tmp + col + (i64)pz * HW);
#pragma unroll
    for (int k = 0; k < KT - 1; ++k) { rx[k] = rx[k + 1]; ry[k] = ry[k + 1]; rz[k] = rz[k + 1]; }
    rx[KT - 1] = v.x; ry[KT - 1] = v.y; rz[KT - 1] = v.z;
    const int zo = pz - KR;
    if (i >= 2 * KR && zo < g.D) {
      float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < KT; ++k) { s[0] += w[k] * rx[k]; s[1] += w[k] * ry[k]; s[2] += w[k] * rz[k]; }
      smooth_out<3, MODE>(a, g, col + (i64)zo * HW, zo, y, x, s);
    }
  }
}

template <int DIM, int MODE>
static void launch_smooth(const Dims& g, const MorphCfg& c, const void* A, const void* B, const void* C,
                          const void* D, void* out, void* tmp, cudaStream_t st) {
  typedef typename V<DIM>::T T;
  SmoothArgs<DIM> a;
  a.A = (const T*)A; a.B = (const T*)B; a.C = (const T*)C; a.D = (const T*)D; a.out = (T*)out;
  for (int i = 0; i < KT; ++i) a.w[i] = c.w[i];
  if (DIM == 2) {
    dim3 grid((g.W + S2_TX - 1) / S2_TX, (g.H + S2_TY - 1) / S2_TY, g.N);
    ADVK_LAUNCH(MODE ? K_smooth_bwd : K_smooth_fwd, st, launch_pdl((smooth2d_kernel<MODE>), grid, S2_THREADS, 0, st, g, *reinterpret_cast<SmoothArgs<2>*>(&a)));
  } else {
    SmoothArgs<3>& a3 = *reinterpret_cast<SmoothArgs<3>*>(&a);
    dim3 gxy((g.W + SX_TX - 1) / SX_TX, (g.H + SX_TY - 1) / SX_TY, (unsigned)(g.N * g.D));
    ADVK_LAUNCH(MODE ? K_smooth_bwd : K_smooth_fwd, st, launch_pdl((smooth3d_xy_kernel<MODE>), gxy, SX_THREADS, 0, st, g, a3, (float4*)tmp));
    const int nzc = (g.D + SZ_CHUNK - 1) / SZ_CHUNK;
    dim3 gz((g.W + 31) / 32, (g.H + 7) / 8, (unsigned)(g.N * nzc));
    ADVK_LAUNCH(MODE ? K_smooth_bwd : K_smooth_fwd, st, launch_pdl((smooth3d_z_kernel<MODE>), gz, 256, 0, st, g, a3, (const float4*)tmp, nzc));
  }
}

}  // namespace advk
#include "advk_smooth_tma.cuh"
namespace advk {

// ---------------------------------------------------------------------------------------
// Adjoint of the align_corners=False linear upsample along ONE axis, gather form (deterministic).
// in  viewed as [outer][n_in ][inner] elements of T, out as [outer][n_out][inner];
// out[o][j][i] = sum_p w(p -> j) * val(o,p,i),  val = (a - b) * vs  (b nullable).
template <typename T> __device__ __forceinline__ T t_zero();
template <> __device__ __forceinline__ float t_zero<float>() { return 0.f; }
template <> __device__ __forceinline__ float2 t_zero<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ float4 t_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void t_fma(float& acc, float w, float a, float b, bool hb) { acc += w * (hb ? a - b : a); }
__device__ __forceinline__ void t_fma(float2& acc, float w, float2 a, float2 b, bool hb) {
  acc.x += w * (hb ? a.x - b.x : a.x); acc.y += w * (hb ? a.y - b.y : a.y);
}
__device__ __forceinline__ void t_fma(float4& acc, float w, float4 a, float4 b, bool hb) {
  acc.x += w * (hb ? a.x - b.x : a.x); acc.y += w * (hb ? a.y - b.y : a.y); acc.z += w * (hb ? a.z - b.z : a.z);
}
__device__ __forceinline__ float t_add(float a, float b) { return a + b; }
__device__ __forceinline__ float2 t_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float4 t_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, 0.f); }
__device__ __forceinline__ float t_scale(float a, float s) { return a * s; }
__device__ __forceinline__ float2 t_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float4 t_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, 0.f); }

template <typename T>
__global__ void __launch_bounds__(256)
adjoint_axis_kernel(const T* __restrict__ a, const T* __restrict__ a2, const T* __restrict__ b, float vs,
                    T* __restrict__ out, i64 outer, int n_in, int n_out, i64 inner, float scale) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= outer * n_out * inner) return;
  i64 i = idx % inner;
  int j = (int)((idx / inner) % n_out);
  i64 o = idx / (inner * n_out);
  // voxels p whose source coordinate scale*(p+0.5)-0.5 lies in (j-1, j+1), widened by one
  int plo = (int)floorf(((float)j - 0.5f) / scale - 0.5f) - 1;
  int phi = (int)ceilf(((float)j + 1.5f) / scale - 0.5f) + 1;
  if (plo < 0) plo = 0;
  if (phi > n_in - 1) phi = n_in - 1;
  if (j == 0) plo = 0;                 // src is clamped at 0 from below
  T acc = t_zero<T>();
  const bool hb = (b != nullptr);
  for (int p = plo; p <= phi; ++p) {
    UpAxis u = up_axis(p, n_out, scale);
    float w = (u.i0 == j ? u.l0 : 0.f) + (u.i1 == j ? u.l1 : 0.f);
    if (w != 0.f) {
      i64 q = (o * n_in + p) * inner + i;
      T av = a[q];
      if (a2) av = t_add(av, a2[q]);
      t_fma(acc, w, av, hb ? b[q] : av, hb);
    }
  }
  out[idx] = t_scale(acc, vs);
}

// Same adjoint for the FIRST (largest) axis, where `inner` is a whole plane / row: a CTA owns 32
// consecutive inner positions and all of the axis; its 8 warps split the axis, every thread streams its
// slice once (no window overlap: each input is read exactly once), keeps the running pair of output
// nodes in registers and flushes to a shared accumulator when the pair changes.
constexpr int ADJ_NOUT_MAX = 64;
__device__ __forceinline__ void sm_add(float4* d, float4 v) {
  atomicAdd(&d->x, v.x); atomicAdd(&d->y, v.y); atomicAdd(&d->z, v.z);
}
__device__ __forceinline__ void sm_add(float2* d, float2 v) { atomicAdd(&d->x, v.x); atomicAdd(&d->y, v.y); }
__device__ __forceinline__ void sm_add(float* d, float v) { atomicAdd(d, v); }

template <typename T>
__global__ void __launch_bounds__(256)
adjoint_axis_big_kernel(const T* __restrict__ a, const T* __restrict__ b, float vs, T* __restrict__ out,
                        int n_in, int n_out, i64 inner, float scale, i64 sl_in, i64 sp_in, i64 so_in,
                        i64 sl_out, i64 sj_out, i64 so_out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  // lanes run over `inner` positions (stride sl_*), the axis has stride sp_in / sj_out, blockIdx.y strides
  // so_*: the same kernel serves [outer][axis][inner] (lanes = inner) and [outer][axis] (lanes = outer)
  extern __shared__ float4 adj_smem4[];
  T* acc = reinterpret_cast<T*>(adj_smem4);                 // [n_out][32]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const i64 o = blockIdx.y;
  const i64 i = (i64)blockIdx.x * 32 + lane;
  for (int k = threadIdx.x; k < n_out * 32; k += 256) acc[k] = t_zero<T>();
  __syncthreads();
  if (i < inner) {
    const int chunk = (n_in + 7) / 8;
    const int p0 = wrp * chunk, p1 = min(n_in, p0 + chunk);
    const bool hb = (b != nullptr);
    int c0 = -1, c1 = -1;
    T s0 = t_zero<T>(), s1 = t_zero<T>();
    for (int p = p0; p < p1; ++p) {
      const UpAxis u = up_axis(p, n_out, scale);
      if (u.i0 != c0 || u.i1 != c1) {
        if (c0 >= 0) { sm_add(acc + c0 * 32 + lane, s0); sm_add(acc + c1 * 32 + lane, s1); }
        c0 = u.i0; c1 = u.i1; s0 = t_zero<T>(); s1 = t_zero<T>();
      }
      const i64 q = o * so_in + (i64)p * sp_in + i * sl_in;
      const T av = a[q];
      const T bv = hb ? b[q] : av;
      t_fma(s0, u.l0, av, bv, hb);
      t_fma(s1, u.l1, av, bv, hb);
    }
    if (c0 >= 0) { sm_add(acc + c0 * 32 + lane, s0); sm_add(acc + c1 * 32 + lane, s1); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_out * 32; k += 256) {
    const int j = k >> 5, l = k & 31;
    const i64 ii = (i64)blockIdx.x * 32 + l;
    if (ii < inner) out[o * so_out + (i64)j * sj_out + ii * sl_out] = t_scale(acc[k], vs);
  }
}

template <typename T>
static void launch_adjoint_axis(const T* a, const T* a2, const T* b, float vs, T* out, i64 outer, int n_in,
                                int n_out, i64 inner, float scale, cudaStream_t st) {
  i64 tot = outer * n_out * inner;
  if (!a2 && inner >= 32 && n_out <= ADJ_NOUT_MAX && outer <= 65535) {
    dim3 grid((unsigned)((inner + 31) / 32), (unsigned)outer);
    ADVK_LAUNCH(K_adjoint_axis, st, launch_pdl((adjoint_axis_big_kernel<T>), grid, 256, sizeof(T) * 32 * n_out, st, 
        a, b, vs, out, n_in, n_out, inner, scale, 1, inner, (i64)n_in * inner, 1, inner, (i64)n_out * inner));
    return;
  }
  if (!a2 && inner == 1 && outer >= 32 && n_out <= ADJ_NOUT_MAX) {      // last axis: lanes over the rows
    dim3 grid((unsigned)((outer + 31) / 32), 1);
    ADVK_LAUNCH(K_adjoint_axis, st, launch_pdl((adjoint_axis_big_kernel<T>), grid, 256, sizeof(T) * 32 * n_out, st, 
        a, b, vs, out, n_in, n_out, outer, scale, n_in, 1, 0, n_out, 1, 0));
    return;
  }
  ADVK_LAUNCH(K_adjoint_axis, st, launch_pdl((adjoint_axis_kernel<T>), blocks_for(tot, 256), 256, 0, st, a, a2, b, vs, out, outer, n_in, n_out, inner, scale));
}

template <int DIM>
__global__ void aos_to_planar_kernel(const typename V<DIM>::T* __restrict__ in, float* __restrict__ out,
                                     int N, i64 lr) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)N * lr) return;
  i64 n = idx / lr, q = idx % lr;
  typename V<DIM>::T v = in[idx];
  out[(n * DIM + 0) * lr + q] = v.x;
  out[(n * DIM + 1) * lr + q] = v.y;
  if (DIM == 3) out[(n * DIM + 2) * lr + q] = V<DIM>::z(v);
}

// ---------------------------------------------------------------------------------------
template <int DIM>
static int field_fwd(const Dims& g, const MorphCfg& c, const float* v, float scale, int nb, float* u_lr,
                     void* levels, void* field_out, float* norm2_out, cudaStream_t st) {
  typedef typename V<DIM>::T T;
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  int NC = g.N * DIM;
  launch_lowres_smooth<DIM>(c, NC, v, scale, u_lr, st);
  T* L = (T*)levels;
  i64 F = (i64)g.N * g.S;
  dim3 grid(blocks_for(g.S, 256), g.N);
  float inv2n = 1.0f / (float)(1u << nb);
  if (norm2_out) cudaMemsetAsync(norm2_out, 0, sizeof(float), st);
  launch_init_phi0<DIM>(c, g, u_lr, inv2n, L, norm2_out, st);     // sum |u|^2 as a by-product when asked for
  const bool lean = (ssb_mode() & 1) == 0;
  // 3-D: the last step also writes the smoothing input r into the scratch level, for the TMA-staged Gaussian
  const bool fused = DIM == 3 && lean && (ssb_mode() & 2) == 0 && ft_encoder() != nullptr;
  for (int k = 1; k <= nb; ++k) {
    if (lean && fused && k == nb)
      ADVK_LAUNCH(K_ss_step, st, (launch_pdl((ss_step_lean_kernel<DIM, true>), grid, 256, 0, st, g, L + (k - 1) * F, L + k * F, L, L + (nb + 1) * F)));
    else if (lean) ADVK_LAUNCH(K_ss_step, st, (launch_pdl((ss_step_lean_kernel<DIM, false>), grid, 256, 0, st, g, L + (k - 1) * F, L + k * F, nullptr, nullptr)));
    else ADVK_LAUNCH(K_ss_step, st, launch_pdl((ss_step_kernel<DIM>), grid, 256, 0, st, g, L + (k - 1) * F, L + k * F));
  }
  if (!(fused && launch_smooth_tma<0>(g, c, L + (nb + 1) * F, nullptr, nullptr, nullptr, field_out, st)))
    launch_smooth<DIM, 0>(g, c, L + nb * F, L, nullptr, nullptr, field_out, L + (nb + 1) * F, st);
  return check_launch("morph_field_fwd");
}

template <int DIM>
static size_t lr_scratch_floats(const Dims& g, const MorphCfg& c) {
  size_t e = sizeof(typename V<DIM>::T) / sizeof(float);
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  i64 p1 = (DIM == 3) ? (i64)g.N * c.Dl * g.H * g.W : 0;
  i64 p2 = (i64)g.N * c.Dl * c.Hl * g.W;
  i64 p3 = (i64)g.N * lr;
  return (size_t)(e * (p1 + p2 + p3) + (i64)g.N * DIM * lr);
}

template <int DIM>
static int field_bwd(const Dims& g, const MorphCfg& c, float scale, int nb, const void* levels,
                     const void* field_out, const void* g_field, void* scratch, float* lr_scratch,
                     float* g_v, cudaStream_t st) {
  typedef typename V<DIM>::T T;
  const T* L = (const T*)levels;
  i64 F = (i64)g.N * g.S;
  T* g_off = (T*)scratch;
  T* buf[2] = {g_off + F, g_off + 2 * F};
  // (9)+(8)+(7): clamp mask, Gaussian (self-adjoint), compose-with-base border mask -> g_off = dL/d(off)
  const bool fused = DIM == 3 && (ssb_mode() & 2) == 0 && ft_encoder() != nullptr;
  if (!(fused && launch_smooth_tma<1>(g, c, g_field, field_out, L + nb * F, L, g_off, st)))
    launch_smooth<DIM, 1>(g, c, g_field, field_out, L + nb * F, L, g_off, g_off + 3 * F, st);
  // dL/dphi_n = g_off ; walk the squaring steps back through scatter targets that must be zero on entry
  // (g_off is kept for the Q1 subtraction below)
  const bool self_zero = (ssb_mode() & 1) == 0;       // the lean adjoint can zero the buffer it consumed
  SideStream* ss = (self_zero && (ssb_mode() & 4) == 0 && nb >= 3) ? side_stream(st) : nullptr;
  T* cur = g_off;
  if (ss) {
    // THREE targets in rotation, zeroed by memsets on a side stream: launch j reads B[(j-1)%3] and scatters into
    // B[j%3]; once it has run, its input is dead and is zeroed BESIDE launch j+1 for launch j+2.  The adjoint
    // kernels sit at ~30 % of the DRAM bandwidth (they are bound by L1<->L2 requests), so the 32 MB memset hides
    // behind them; zeroing inside the kernel costs 4.7 us of every launch (ss_step_bwd_lean<.., true> vs
    // <.., false>: 44.3 vs 39.6 us at 128^3).  Forks and joins are events: capturable, nothing is synchronised.
    T* B[3] = {g_off + F, g_off + 2 * F, g_off + 3 * F};
    cudaEvent_t pending[3] = {nullptr, nullptr, nullptr};
    cudaMemsetAsync(B[0], 0, sizeof(T) * F * 2, st);
    pending[2] = ss->zero_beside(st, B[2], sizeof(T) * F);           // (also orders it after the smoothing's scratch use)
    for (int j = 0; j < nb; ++j) {
      const int k = nb - j;
      T* nxt = B[j % 3];
      if (pending[j % 3]) { cudaStreamWaitEvent(st, pending[j % 3], 0); pending[j % 3] = nullptr; }
      launch_ss_step_bwd<DIM>(g, L + (k - 1) * F, cur, nxt, false, st);
      if (j >= 1 && j + 2 < nb) pending[(j - 1) % 3] = ss->zero_beside(st, B[(j - 1) % 3], sizeof(T) * F);
      cur = nxt;
    }
    for (int i = 0; i < 3; ++i)
      if (pending[i]) cudaStreamWaitEvent(st, pending[i], 0);       // every side-stream branch joins before we return
  } else {
    // two targets: zeroed by the lean kernel itself (the buffer it consumed) or, plain kernels, by memset nodes
    if (self_zero) cudaMemsetAsync(buf[0], 0, sizeof(T) * F * (nb > 1 ? 2 : 1), st);
    for (int k = nb; k >= 1; --k) {
      T* nxt = buf[(nb - k) & 1];
      if (!self_zero) cudaMemsetAsync(nxt, 0, sizeof(T) * F, st);
      launch_ss_step_bwd<DIM>(g, L + (k - 1) * F, cur, nxt, cur != g_off, st);
      cur = nxt;
    }
  }
  const T* curS = cur;
  const T* curJ = nullptr;
  // dL/dphi_0 = S + J - g_off (quirk Q1);  dL/du = that / 2^n;  then the upsample adjoint per axis
  float inv2n = 1.0f / (float)(1u << nb);
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  T* s1 = (T*)lr_scratch;
  const T* a = curS;
  const T* a2 = curJ;
  const T* b = g_off;
  float vs = inv2n;
  if (DIM == 3) {
    launch_adjoint_axis<T>(a, a2, b, vs, s1, g.N, g.D, c.Dl, (i64)g.H * g.W, c.sD, st);
    a = s1; a2 = nullptr; b = nullptr; vs = 1.f;
    s1 += (i64)g.N * c.Dl * g.H * g.W;
  }
  T* s2 = s1 + (i64)g.N * c.Dl * c.Hl * g.W;
  // the last two axes in one launch (advk_adjoint.cuh); per-axis kernels when a row does not fit shared memory
  if (a2 || !launch_adjoint_hw<T>(K_adjoint_axis, a, b, vs, s2, (i64)g.N * c.Dl, g.H, g.W, c.Hl, c.Wl, c.sH, c.sW, st)) {
    launch_adjoint_axis<T>(a, a2, b, vs, s1, (i64)g.N * c.Dl, g.H, c.Hl, g.W, c.sH, st);
    launch_adjoint_axis<T>(s1, nullptr, nullptr, 1.f, s2, (i64)g.N * c.Dl * c.Hl, g.W, c.Wl, 1, c.sW, st);
  }
  float* planar = (float*)(s2 + (i64)g.N * lr);
  ADVK_LAUNCH(K_aos_to_planar, st, launch_pdl((aos_to_planar_kernel<DIM>), blocks_for(g.N * lr, 128), 128, 0, st, s2, planar, g.N, lr));
  int NC = g.N * DIM;
  launch_lowres_smooth<DIM>(c, NC, planar, scale, g_v, st);
  return check_launch("morph_field_bwd");
}

}  // namespace advk

using namespace advk;

extern "C" int advk_morph_tune(int ssb_mode_mask) {
  int prev = ssb_mode();
  if (ssb_mode_mask >= 0) g_ssb_mode = ssb_mode_mask & 63;
  return prev;
}

extern "C" int advk_morph_unorm2(const advk_geom* gg, const advk_morph_cfg* cfg, const float* v,
                                 float scale, float* u_lr, float* out_norm2, void* stream) {
  Dims g; MorphCfg c;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(make_cfg(cfg, g, gg->d, c), "bad morph config (ktaps must be 9)");
  ADVK_REQUIRE(v && u_lr && out_norm2, "null pointer");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  cudaStream_t st = (cudaStream_t)stream;
  i64 lr = (i64)c.Dl * c.Hl * c.Wl;
  int NC = g.N * gg->d;
  cudaMemsetAsync(out_norm2, 0, sizeof(float), st);
  dim3 grid(blocks_for(g.S, 256), g.N);
  if (gg->d == 2) {
    launch_lowres_smooth<2>(c, NC, v, scale, u_lr, st);
    launch_init_phi0<2>(c, g, u_lr, 0.f, nullptr, out_norm2, st);
  } else {
    launch_lowres_smooth<3>(c, NC, v, scale, u_lr, st);
    launch_init_phi0<3>(c, g, u_lr, 0.f, nullptr, out_norm2, st);
  }
  return check_launch("morph_unorm2");
}

extern "C" int advk_morph_field_fwd(const advk_geom* gg, const advk_morph_cfg* cfg, const float* v,
                                    float scale, int nb_steps, float* u_lr, void* levels,
                                    void* field_out, float* norm2_out, void* stream) {
  Dims g; MorphCfg c;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(make_cfg(cfg, g, gg->d, c), "bad morph config (ktaps must be 9)");
  ADVK_REQUIRE(v && u_lr && levels && field_out, "null pointer");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  ADVK_REQUIRE(nb_steps >= 1 && nb_steps <= 30, "nb_steps out of range");
  ADVK_REQUIRE((i64)g.N * g.D <= 65535, "batch x depth too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  return gg->d == 2 ? field_fwd<2>(g, c, v, scale, nb_steps, u_lr, levels, field_out, norm2_out, st)
                    : field_fwd<3>(g, c, v, scale, nb_steps, u_lr, levels, field_out, norm2_out, st);
}

extern "C" size_t advk_morph_lr_scratch_floats(const advk_geom* gg, const advk_morph_cfg* cfg) {
  Dims g; MorphCfg c;
  if (!make_dims(gg, g) || !make_cfg(cfg, g, gg->d, c)) return 0;
  return gg->d == 2 ? lr_scratch_floats<2>(g, c) : lr_scratch_floats<3>(g, c);
}

extern "C" int advk_morph_field_bwd(const advk_geom* gg, const advk_morph_cfg* cfg, float scale,
                                    int nb_steps, const void* levels, const void* field_out,
                                    const void* g_field, void* scratch, float* lr_scratch,
                                    float* g_v, void* stream) {
  Dims g; MorphCfg c;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(make_cfg(cfg, g, gg->d, c), "bad morph config (ktaps must be 9)");
  ADVK_REQUIRE(levels && field_out && g_field && scratch && lr_scratch && g_v, "null pointer");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  ADVK_REQUIRE(nb_steps >= 1 && nb_steps <= 30, "nb_steps out of range");
  ADVK_REQUIRE((i64)g.N * g.D <= 65535, "batch x depth too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  return gg->d == 2 ? field_bwd<2>(g, c, scale, nb_steps, levels, field_out, g_field, scratch, lr_scratch, g_v, st)
                    : field_bwd<3>(g, c, scale, nb_steps, levels, field_out, g_field, scratch, lr_scratch, g_v, st);
}
