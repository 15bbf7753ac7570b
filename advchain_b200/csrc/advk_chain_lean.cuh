// advk_chain_lean.cuh -- the chain stages specialised at compile time for the configuration every
// BASELINE workload runs: zeros padding, bi/tri-linear interpolation (adv_morph.py:556-557, adv_affine.py:
// 312-313 with the constructors' defaults), no pad values.  Included by advk_chain.cu.
//
// One launch per stage (a resampling needs the COMPLETE output of the stage before it, so stages are
// kernel boundaries either way); what is specialised away against the generic executor in advk_chain.cu:
//   * no Stage interpreter, no run-time pad / interp branches: FIELD vs AFFINE, mask mode, channel layout
//     are template parameters;
//   * no validity predicates: a corner outside the volume gets weight 0 and its index is clamped into the
//     volume (what the lean squaring-step kernels do for border padding, advk_morph.cu), so every gather is
//     an unconditional load from base + 32-bit offset with ONE opaque per-sample base pointer;
//   * the spatial Jacobian is built by separable differences (interpolate along x, difference along y, ...)
//     instead of three weighted sums over all corners;
//   * the adjoint's scatter hands the x1-corner contributions to lane+1 on ONE address test per voxel; the
//     channel-packed variant ships the upstream float4 once plus one weight per corner row;
//   * every kernel gets the register budget of its own stage (the fused cooperative kernel ran every phase
//     at the budget of its hungriest one and spilled).
// Arithmetic is the generic executor's: same weights ((wx*wy)*wz, ATen's order), same corner order in the
// gather sums, so both paths agree to fp32 summation order (tests/test_gpu_kernels.py).
#pragma once

namespace advk {

template <typename P>
__device__ __forceinline__ P* lean_opaque(P* p) {
  unsigned long long v = (unsigned long long)p;
  asm("" : "+l"(v));
  P* q = (P*)v;
  __builtin_assume(__isGlobal(q));
  return q;
}

// One axis: clamped corner indices, weights zeroed where the corner is outside (zeros padding), validity
// as 0/1 factors for the Jacobian.  Same pixel coordinate / weights as make_axis(ZEROS, LINEAR).
struct LAxis { int o0, o1; float w0, w1, k0, k1; };

__device__ __forceinline__ LAxis lean_axis(float coord, int size) {
  LAxis a;
  const float mx = (float)(size - 1);
  const float ix = ((coord + 1.f) / 2.f) * mx;
  const float f = floorf(ix);
  const bool v0 = f >= 0.f && f <= mx;              // false for NaN: ATen's "far outside"
  const bool v1 = f >= -1.f && f <= mx - 1.f;
  a.k0 = v0 ? 1.f : 0.f; a.k1 = v1 ? 1.f : 0.f;
  a.w0 = v0 ? (f + 1.f) - ix : 0.f;
  a.w1 = v1 ? ix - f : 0.f;
  a.o0 = (int)fminf(fmaxf(f, 0.f), mx);
  a.o1 = (int)fminf(fmaxf(f + 1.f, 0.f), mx);
  return a;
}

template <int DIM> struct LGeo {
  LAxis x, y, z;
  int a000, dxo, dyo, dzo;       // corner (0,0,0) offset and the (0 or one-row/one-plane) steps to corner 1
};

template <int DIM>
__device__ __forceinline__ void lean_geo(float cx, float cy, float cz, const Dims& g, LGeo<DIM>& G) {
  G.x = lean_axis(cx, g.W);
  G.y = lean_axis(cy, g.H);
  if (DIM == 3) G.z = lean_axis(cz, g.D);
  else { G.z.o0 = G.z.o1 = 0; G.z.w0 = 1.f; G.z.w1 = 0.f; G.z.k0 = 1.f; G.z.k1 = 0.f; }
  const int HW = g.H * g.W;
  G.a000 = G.z.o0 * HW + G.y.o0 * g.W + G.x.o0;
  G.dxo = G.x.o1 - G.x.o0;
  G.dyo = (G.y.o1 - G.y.o0) * g.W;
  G.dzo = DIM == 3 ? (G.z.o1 - G.z.o0) * HW : 0;
}

#define LEAN_W(G, dx, dy, dz) ((((dx) ? (G).x.w1 : (G).x.w0) * ((dy) ? (G).y.w1 : (G).y.w0)) * ((dz) ? (G).z.w1 : (G).z.w0))
#define LEAN_OFF(G, dx, dy, dz) ((dz) * (G).dzo + (dy) * (G).dyo + (dx) * (G).dxo)

// sampling coordinates of voxel p of one sample: c* clamped (what is sampled), r* raw field values (for the
// clamp-range gradient mask), b* base coordinates (affine: for the theta gradient)
template <int DIM, bool FIELD>
__device__ __forceinline__ void lean_coords(const Dims& g, const typename FieldT<DIM>::type* __restrict__ phi_n,
                                            const float* __restrict__ th, int p, bool ok, float& cx, float& cy,
                                            float& cz, float& rx, float& ry, float& rz, float& bx, float& by,
                                            float& bz) {
  if (FIELD) {
    rx = ry = rz = 0.f;
    if (ok) {
      if (DIM == 2) {
        const float2 f = __ldg(reinterpret_cast<const float2*>(phi_n) + p);
        rx = f.x; ry = f.y;
      } else {
        const float4 f = __ldg(reinterpret_cast<const float4*>(phi_n) + p);
        rx = f.x; ry = f.y; rz = f.z;
      }
    }
    cx = clampf(rx, -1.f, 1.f); cy = clampf(ry, -1.f, 1.f); cz = clampf(rz, -1.f, 1.f);
    bx = by = bz = 0.f;
  } else {
    int x, y, z;
    voxel_xyz(g, (unsigned)p, x, y, z);
    bx = base_coord_s(x, g.W, g.stW, 0.f); by = base_coord_s(y, g.H, g.stH, 0.f);
    if (DIM == 2) {
      cx = __fadd_rn(__fmaf_rn(__ldg(th + 1), by, __fmul_rn(__ldg(th + 0), bx)), __ldg(th + 2));
      cy = __fadd_rn(__fmaf_rn(__ldg(th + 4), by, __fmul_rn(__ldg(th + 3), bx)), __ldg(th + 5));
      cz = 0.f; bz = 0.f;
    } else {
      bz = base_coord_s(z, g.D, g.stD, 0.f);
      cx = __fadd_rn(__fmaf_rn(__ldg(th + 2), bz, __fmaf_rn(__ldg(th + 1), by, __fmul_rn(__ldg(th + 0), bx))), __ldg(th + 3));
      cy = __fadd_rn(__fmaf_rn(__ldg(th + 6), bz, __fmaf_rn(__ldg(th + 5), by, __fmul_rn(__ldg(th + 4), bx))), __ldg(th + 7));
      cz = __fadd_rn(__fmaf_rn(__ldg(th + 10), bz, __fmaf_rn(__ldg(th + 9), by, __fmul_rn(__ldg(th + 8), bx))), __ldg(th + 11));
    }
    rx = cx; ry = cy; rz = cz;
  }
}

// ------------------------------------------------------------------------------------------- forward

struct LeanFwd {
  Dims g; int C;
  const float* src; float* dst;           // planar N x C x S, or packed float4 [N][C/4][S]
  const void* phi; const float* theta;
  const float* msrc; float* mdst; int binarize;
  int clamp; float lo, hi;
};

// mask of the stage: MASKM 0 none, 1 generated (source mask = ones: sum of the in-bounds weights), 2 gathered
template <int DIM, int MASKM>
__device__ __forceinline__ void lean_mask_out(const LeanFwd& a, const LGeo<DIM>& G, i64 nS, int p) {
  if (MASKM == 0) return;
  float m = 0.f;
  const float* ms = MASKM == 2 ? lean_opaque(a.msrc + nS) + G.a000 : nullptr;
#pragma unroll
  for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        m += (MASKM == 2 ? __ldg(ms + LEAN_OFF(G, dx, dy, dz)) : 1.f) * LEAN_W(G, dx, dy, dz);
  if (a.binarize) m = (m != 0.f) ? 1.f : 0.f;
  a.mdst[nS + p] = m;
}

// VM_SRC / VM_DST (C == 1 image chains): the source / destination carries the valid-region mask interleaved
// with the value, float2 {value, mask} per voxel, so that the next resampling needs ONE 8-byte gather per
// corner for both (an affine map turns every warp-wide gather into ~10 cache-line wavefronts: the count of
// gather instructions is what such a stage pays for).
template <int DIM, bool FIELD, int MASKM, bool VM_SRC, bool VM_DST>
__global__ void __launch_bounds__(256)
lean_warp_fwd_kernel(const __grid_constant__ LeanFwd a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename FieldT<DIM>::type T;
  static_assert(!VM_SRC || MASKM == 2, "an interleaved source carries the mask to gather");
  static_assert(!VM_DST || MASKM != 0, "an interleaved destination needs a mask");
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= a.g.S) return;
  const int n = blockIdx.y;
  const i64 nS = (i64)n * a.g.S;
  float cx, cy, cz, rx, ry, rz, bx, by, bz;
  lean_coords<DIM, FIELD>(a.g, FIELD ? lean_opaque(reinterpret_cast<const T*>(a.phi) + nS) : nullptr,
                          FIELD ? nullptr : a.theta + n * (DIM * (DIM + 1)), p, true, cx, cy, cz, rx, ry, rz, bx, by, bz);
  LGeo<DIM> G;
  lean_geo<DIM>(cx, cy, cz, a.g, G);
  if (VM_SRC || VM_DST) {                       // C == 1 (host-checked)
    float acc = 0.f, m = 0.f;
    if (VM_SRC) {
      const float2* s = lean_opaque(reinterpret_cast<const float2*>(a.src) + nS) + G.a000;
#pragma unroll
      for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float2 v = __ldg(s + LEAN_OFF(G, dx, dy, dz));
            const float w = LEAN_W(G, dx, dy, dz);
            acc += v.x * w; m += v.y * w;
          }
    } else {
      const float* s = lean_opaque(a.src + nS) + G.a000;
      const float* ms = MASKM == 2 ? lean_opaque(a.msrc + nS) + G.a000 : nullptr;
#pragma unroll
      for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float w = LEAN_W(G, dx, dy, dz);
            acc += __ldg(s + LEAN_OFF(G, dx, dy, dz)) * w;
            m += (MASKM == 2 ? __ldg(ms + LEAN_OFF(G, dx, dy, dz)) : 1.f) * w;
          }
    }
    if (a.clamp) acc = clampf(acc, a.lo, a.hi);
    if (a.binarize) m = (m != 0.f) ? 1.f : 0.f;
    if (VM_DST) {
      (lean_opaque(reinterpret_cast<float2*>(a.dst) + nS))[p] = make_float2(acc, m);
    } else {
      a.dst[nS + p] = acc;
      a.mdst[nS + p] = m;
    }
    return;
  }
  const float* s = lean_opaque(a.src + nS * a.C) + G.a000;
  float* d = lean_opaque(a.dst + nS * a.C) + p;
  for (int c = 0; c < a.C; ++c, s += a.g.S, d += a.g.S) {
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) acc += __ldg(s + LEAN_OFF(G, dx, dy, dz)) * LEAN_W(G, dx, dy, dz);
    if (a.clamp) acc = clampf(acc, a.lo, a.hi);
    *d = acc;
  }
  lean_mask_out<DIM, MASKM>(a, G, nS, p);
}

// channel-packed (C % 4 == 0, the K-class prediction path): the source is always float4-per-voxel-per-group
// (the chain input is packed once by lean_pack_kernel: 8 LDG.128 per voxel instead of 32 LDG.32 whose
// rotated warp-wide footprint costs ~10 wavefronts each); DST_PK: packed (intermediate) or planar (chain output)
template <int DIM, bool FIELD, int MASKM, bool DST_PK>
__global__ void __launch_bounds__(256)
lean_warp_fwd_pk_kernel(const __grid_constant__ LeanFwd a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename FieldT<DIM>::type T;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= a.g.S) return;
  const int n = blockIdx.y;
  const i64 nS = (i64)n * a.g.S;
  const int S = (int)a.g.S;
  float cx, cy, cz, rx, ry, rz, bx, by, bz;
  lean_coords<DIM, FIELD>(a.g, FIELD ? lean_opaque(reinterpret_cast<const T*>(a.phi) + nS) : nullptr,
                          FIELD ? nullptr : a.theta + n * (DIM * (DIM + 1)), p, true, cx, cy, cz, rx, ry, rz, bx, by, bz);
  LGeo<DIM> G;
  lean_geo<DIM>(cx, cy, cz, a.g, G);
  const int CG = a.C >> 2;
  for (int cg = 0; cg < CG; ++cg) {
    const i64 gb = ((i64)n * CG + cg) * a.g.S;                 // group base in voxels (x4 floats either way)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* s = lean_opaque(reinterpret_cast<const float4*>(a.src) + gb) + G.a000;
#pragma unroll
    for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const float4 v = __ldg(s + LEAN_OFF(G, dx, dy, dz));
          const float w = LEAN_W(G, dx, dy, dz);
          acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w; acc.w += v.w * w;
        }
    if (DST_PK) {
      (lean_opaque(reinterpret_cast<float4*>(a.dst) + gb))[p] = acc;
    } else {
      float* d = lean_opaque(a.dst + 4 * gb) + p;
      d[0] = acc.x; d[S] = acc.y; d[2 * S] = acc.z; d[3 * S] = acc.w;
    }
  }
  lean_mask_out<DIM, MASKM>(a, G, nS, p);
}

struct LeanInt {
  Dims g; int C;
  int order; float ns; int use_ig; float ig; BiasCfg b;
  const float* src; const float* delta; const float* low; float* dst;
  int clamp; float lo, hi;
  const float* g_dst; float* g_src; float* g_delta; float* g_up;
};

template <int DIM, bool BWD>
__global__ void __launch_bounds__(256)
lean_intensity_kernel(const __grid_constant__ LeanInt a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= a.g.S) return;
  const int n = blockIdx.y;
  float bv = 1.f, braw = 1.f;
  bool pass = true;
  if (a.order != 0) {
    int x = 0, y = 0, z = 0;
    if (a.b.upsample) voxel_xyz(a.g, (unsigned)p, x, y, z);
    bv = bias_value(a.b, bias_up<DIM>(a.b, a.low + (i64)n * a.b.lD * a.b.lH * a.b.lW, z, y, x, a.g, p), braw, pass);
  }
  i64 q = (i64)n * a.C * a.g.S + p;
  float gb = 0.f;
  for (int c = 0; c < a.C; ++c, q += a.g.S) {
    const float x0 = a.src[q];
    const float dl = (a.order != 1) ? a.delta[q] : 0.f;
    if (!BWD) {
      float val = intensity_point(a.order, x0, dl, a.ns, bv, a.use_ig, a.ig);
      if (a.clamp) val = clampf(val, a.lo, a.hi);
      a.dst[q] = val;
    } else {
      float go = a.g_dst[q];
      if (a.clamp) {
        const float val = intensity_point(a.order, x0, dl, a.ns, bv, a.use_ig, a.ig);
        if (!(val >= a.lo && val <= a.hi)) go = 0.f;
      }
      float gd;
      const float gi = intensity_point_bwd(a.order, go, x0, dl, a.ns, bv, a.use_ig, a.ig, gd, gb);
      if (a.g_delta) a.g_delta[q] = gd;
      if (a.g_src) a.g_src[q] = gi;
    }
  }
  if (BWD && a.g_up) a.g_up[(i64)n * a.g.S + p] = pass ? gb * (a.b.use_log ? braw : 1.f) : 0.f;
}

// Row-wise version for the bias stage with an upsampled field (orders 1-3, every BASELINE workload): a CTA owns
// 8 whole rows of one z-plane, one warp per row.  The z and y interpolation weights of the bias lattice are
// shared by a whole row, so the CTA first reduces the low-res field to one line of lW values per row in shared
// memory; a voxel then interpolates along x only (2 shared reads instead of 2^d global gathers + 3 index
// computations: the generic kernel above issues 273 warp instructions per 32 voxels and is issue-bound at
// 80 %, gpurun_out/r02i).
constexpr int LIR_MAX = 128;
template <int DIM, bool BWD>
__global__ void __launch_bounds__(256)
lean_intensity_rows_kernel(const __grid_constant__ LeanInt a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float rowbuf[8][LIR_MAX];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int z = (DIM == 3) ? blockIdx.z % a.g.D : 0;
  const int n = (DIM == 3) ? blockIdx.z / a.g.D : blockIdx.z;
  const int y = blockIdx.y * 8 + ty;
  const BiasCfg& b = a.b;
  {
    UpAxis uz;
    if (DIM == 3) uz = up_axis(z, b.lD, b.sD);
    else { uz.i0 = uz.i1 = 0; uz.l0 = 1.f; uz.l1 = 0.f; }
    const float* low_n = a.low + (i64)n * b.lD * b.lH * b.lW;
    for (int i = threadIdx.x; i < 8 * b.lW; i += 256) {
      const int r = i / b.lW, xl = i - r * b.lW;
      const int yy = min((int)blockIdx.y * 8 + r, a.g.H - 1);
      const UpAxis uy = up_axis(yy, b.lH, b.sH);
      float acc = 0.f;
#pragma unroll
      for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
        const int zi = dz ? uz.i1 : uz.i0;
        const float lz = dz ? uz.l1 : uz.l0;
        acc += lz * (uy.l0 * __ldg(low_n + (zi * b.lH + uy.i0) * b.lW + xl) + uy.l1 * __ldg(low_n + (zi * b.lH + uy.i1) * b.lW + xl));
      }
      rowbuf[r][xl] = acc;
    }
  }
  __syncthreads();
  if (y >= a.g.H) return;
  const int S = (int)a.g.S;
  const int rowp = (z * a.g.H + y) * a.g.W;
  const i64 nCS = (i64)n * a.C * a.g.S;
  for (int x = tx; x < a.g.W; x += 32) {
    const UpAxis ux = up_axis(x, b.lW, b.sW);
    float braw;
    bool pass;
    const float bv = bias_value(b, ux.l0 * rowbuf[ty][ux.i0] + ux.l1 * rowbuf[ty][ux.i1], braw, pass);
    const int p = rowp + x;
    i64 q = nCS + p;
    float gb = 0.f;
    for (int c = 0; c < a.C; ++c, q += S) {
      const float x0 = a.src[q];
      const float dl = (a.order != 1) ? a.delta[q] : 0.f;
      if (!BWD) {
        float val = intensity_point(a.order, x0, dl, a.ns, bv, a.use_ig, a.ig);
        if (a.clamp) val = clampf(val, a.lo, a.hi);
        a.dst[q] = val;
      } else {
        float go = a.g_dst[q];
        if (a.clamp) {
          const float val = intensity_point(a.order, x0, dl, a.ns, bv, a.use_ig, a.ig);
          if (!(val >= a.lo && val <= a.hi)) go = 0.f;
        }
        float gd;
        const float gi = intensity_point_bwd(a.order, go, x0, dl, a.ns, bv, a.use_ig, a.ig, gd, gb);
        if (a.g_delta) a.g_delta[q] = gd;
        if (a.g_src) a.g_src[q] = gi;
      }
    }
    if (BWD && a.g_up) a.g_up[(i64)n * a.g.S + p] = pass ? gb * (b.use_log ? braw : 1.f) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------- adjoint

struct LeanBwd {
  Dims g; int C; int tps; FastDiv ftps; unsigned n_tiles;
  const float* src; const float* g_dst; float* g_src;
  const void* phi; const float* theta; void* g_phi; float* g_theta;
  int clamp; float lo, hi;
};

// separable Jacobian of the interpolated value from the 2^d corner values v (already multiplied by the
// upstream gradient, or dot products with it): j = d(sum_k v_k w_k)/d(pixel coordinate), corners outside
// the volume contribute 0 (k* = validity, w* = weights zeroed outside)
template <int DIM>
__device__ __forceinline__ void lean_jacobian(const LGeo<DIM>& G, const float (&v)[2][2][2], float& jx, float& jy, float& jz) {
  constexpr int NZ = DIM == 3 ? 2 : 1;
  float jxz[2], jyz[2], e[2];
#pragma unroll
  for (int dz = 0; dz < NZ; ++dz) {
    const float m0 = G.x.w0 * v[dz][0][0] + G.x.w1 * v[dz][0][1], m1 = G.x.w0 * v[dz][1][0] + G.x.w1 * v[dz][1][1];
    const float d0 = G.x.k1 * v[dz][0][1] - G.x.k0 * v[dz][0][0], d1 = G.x.k1 * v[dz][1][1] - G.x.k0 * v[dz][1][0];
    jxz[dz] = G.y.w0 * d0 + G.y.w1 * d1;
    jyz[dz] = G.y.k1 * m1 - G.y.k0 * m0;
    e[dz] = G.y.w0 * m0 + G.y.w1 * m1;
  }
  if (DIM == 3) {
    jx = G.z.w0 * jxz[0] + G.z.w1 * jxz[1];
    jy = G.z.w0 * jyz[0] + G.z.w1 * jyz[1];
    jz = G.z.k1 * e[1] - G.z.k0 * e[0];
  } else {
    jx = jxz[0]; jy = jyz[0]; jz = 0.f;
  }
}

template <int DIM>
__device__ __forceinline__ void lean_theta_acc(float (&acc)[DIM * (DIM + 1)], float ggx, float ggy, float ggz, float bx,
                                               float by, float bz) {
  if (DIM == 2) {
    acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx;
    acc[3] += ggy * bx; acc[4] += ggy * by; acc[5] += ggy;
  } else {
    acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx * bz; acc[3] += ggx;
    acc[4] += ggy * bx; acc[5] += ggy * by; acc[6] += ggy * bz; acc[7] += ggy;
    acc[8] += ggz * bx; acc[9] += ggz * by; acc[10] += ggz * bz; acc[11] += ggz;
  }
}

template <int DIM>
__device__ __forceinline__ void lean_theta_flush(float (&acc)[DIM * (DIM + 1)], float* red, float* g_theta, int n) {
  constexpr int NG = DIM * (DIM + 1);
  block_sum<NG>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NG; ++i) atomicAdd(g_theta + n * NG + i, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
}

template <int DIM>
__device__ __forceinline__ void lean_store_gphi(void* g_phi, i64 idx, float ggx, float ggy, float ggz, float rx, float ry,
                                                float rz) {
  // consumers clamp the stored field to [-1,1]; torch.clamp passes gradient on the closed interval
  if (!(rx >= -1.f && rx <= 1.f)) ggx = 0.f;
  if (!(ry >= -1.f && ry <= 1.f)) ggy = 0.f;
  if (DIM == 2) {
    reinterpret_cast<float2*>(g_phi)[idx] = make_float2(ggx, ggy);
  } else {
    if (!(rz >= -1.f && rz <= 1.f)) ggz = 0.f;
    reinterpret_cast<float4*>(g_phi)[idx] = make_float4(ggx, ggy, ggz, 0.f);
  }
}

// hand-off key of a voxel: corner (0,0,0) offset (S < 2^29, host-checked) + whether the y / z steps are
// real; lane+1 takes this lane's x1 column exactly when its key is this key + 4 (same rows, one voxel on)
template <int DIM>
__device__ __forceinline__ int lean_key(const LGeo<DIM>& G, bool ok) {
  return ok ? ((G.a000 << 2) | (G.dyo ? 1 : 0) | (G.dzo ? 2 : 0)) : -1;
}

template <int DIM, bool FIELD, bool VM_SRC, int MINB, bool THACC = false>
__global__ void __launch_bounds__(256, MINB)
lean_warp_bwd_kernel(const __grid_constant__ LeanBwd a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename FieldT<DIM>::type T;
  constexpr int NG = DIM * (DIM + 1);
  constexpr int NZ = DIM == 3 ? 2 : 1;
  const unsigned FULL = 0xffffffffu;
  __shared__ float red[12 * 32];
  const int lane = threadIdx.x & 31;
  const bool want_theta = THACC && !FIELD && a.g_theta != nullptr;
  const bool scatter = a.g_src != nullptr;
  const int S = (int)a.g.S;
  float acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
  int cur_n = -1;
  for (unsigned t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int n = (int)fast_div(t, a.ftps);
    const int p = (int)(t - (unsigned)n * (unsigned)a.tps) * 256 + threadIdx.x;
    const bool ok = p < S;
    if (want_theta && n != cur_n) {              // block-uniform
      if (cur_n >= 0) lean_theta_flush<DIM>(acc, red, a.g_theta, cur_n);
      cur_n = n;
    }
    const i64 nS = (i64)n * a.g.S;
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    lean_coords<DIM, FIELD>(a.g, FIELD ? lean_opaque(reinterpret_cast<const T*>(a.phi) + nS) : nullptr,
                            FIELD ? nullptr : a.theta + n * NG, ok ? p : 0, ok, cx, cy, cz, rx, ry, rz, bx, by, bz);
    LGeo<DIM> G;
    lean_geo<DIM>(cx, cy, cz, a.g, G);
    bool hand = false;
    if (scatter) {                               // (every shuffle is executed by all lanes)
      const int key = lean_key<DIM>(G, ok);
      const int nkey = __shfl_down_sync(FULL, key, 1);
      hand = ok && G.dxo && lane < 31 && nkey == key + 4;
    }
    // VM_SRC (C == 1): the stashed source is float2 {value, mask} per voxel
    const float* s = lean_opaque(a.src + (VM_SRC ? 2 * nS : nS * a.C)) + (VM_SRC ? 2 : 1) * G.a000;
    const float* gd = lean_opaque(a.g_dst + nS * a.C);
    // (no ternary with nullptr here: it would turn the pointer generic and every RED into the three-way
    // generic-address atomic sequence; a NULL g_src is never dereferenced)
    float* gs = lean_opaque(a.g_src + nS * a.C) + G.a000;
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    for (int c = 0; c < a.C; ++c, s += S, gd += S) {
      float go = ok ? gd[p] : 0.f;
      float v[2][2][2];
#pragma unroll
      for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) v[dz][dy][dx] = __ldg(s + (VM_SRC ? 2 : 1) * LEAN_OFF(G, dx, dy, dz));
      if (a.clamp) {                             // same summation order as the forward kernel
        float val = 0.f;
#pragma unroll
        for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) val += v[dz][dy][dx] * LEAN_W(G, dx, dy, dz);
        if (!(val >= a.lo && val <= a.hi)) go = 0.f;
      }
      float jx, jy, jz;
      lean_jacobian<DIM>(G, v, jx, jy, jz);
      ggx += jx * go; ggy += jy * go; ggz += jz * go;
      if (scatter) {
#pragma unroll
        for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const float wr = (dy ? G.y.w1 : G.y.w0) * (dz ? G.z.w1 : G.z.w0);
            float c0 = go * (G.x.w0 * wr);
            const float c1 = go * (G.x.w1 * wr);
            const float rcv = __shfl_up_sync(FULL, hand ? c1 : 0.f, 1);
            if (lane) c0 += rcv;
            float* r = gs + LEAN_OFF(G, 0, dy, dz);
            if (c0 != 0.f) atomicAdd(r, c0);
            if (!hand && c1 != 0.f) atomicAdd(r + G.dxo, c1);
          }
        gs += S;
      }
    }
    if (!ok) continue;
    ggx *= (float)(a.g.W - 1) / 2.f; ggy *= (float)(a.g.H - 1) / 2.f; ggz *= (float)(a.g.D - 1) / 2.f;
    if (FIELD) {
      if (a.g_phi) lean_store_gphi<DIM>(a.g_phi, nS + p, ggx, ggy, ggz, rx, ry, rz);
    } else if (THACC) {
      if (want_theta) lean_theta_acc<DIM>(acc, ggx, ggy, ggz, bx, by, bz);
    } else if (a.g_phi) {                        // affine stage: dL/d(coordinate) for lean_theta_reduce_kernel
      if (DIM == 2) reinterpret_cast<float2*>(a.g_phi)[nS + p] = make_float2(ggx, ggy);
      else reinterpret_cast<float4*>(a.g_phi)[nS + p] = make_float4(ggx, ggy, ggz, 0.f);
    }
  }
  if (want_theta && cur_n >= 0) lean_theta_flush<DIM>(acc, red, a.g_theta, cur_n);
}

// channel-packed adjoint: the stashed source and g_src (scatter target) are always packed;
// GD_PK: layout of the upstream gradient (planar = the user's g_out at the last stage)
template <int DIM, bool FIELD, bool GD_PK, int MINB, bool THACC = false>
__global__ void __launch_bounds__(256, MINB)
lean_warp_bwd_pk_kernel(const __grid_constant__ LeanBwd a) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  typedef typename FieldT<DIM>::type T;
  constexpr int NG = DIM * (DIM + 1);
  constexpr int NZ = DIM == 3 ? 2 : 1;
  const unsigned FULL = 0xffffffffu;
  __shared__ float red[12 * 32];
  const int lane = threadIdx.x & 31;
  const bool want_theta = THACC && !FIELD && a.g_theta != nullptr;
  const bool scatter = a.g_src != nullptr;
  const int S = (int)a.g.S;
  const int CG = a.C >> 2;
  float acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
  int cur_n = -1;
  for (unsigned t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int n = (int)fast_div(t, a.ftps);
    const int p = (int)(t - (unsigned)n * (unsigned)a.tps) * 256 + threadIdx.x;
    const bool ok = p < S;
    if (want_theta && n != cur_n) {
      if (cur_n >= 0) lean_theta_flush<DIM>(acc, red, a.g_theta, cur_n);
      cur_n = n;
    }
    const i64 nS = (i64)n * a.g.S;
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    lean_coords<DIM, FIELD>(a.g, FIELD ? lean_opaque(reinterpret_cast<const T*>(a.phi) + nS) : nullptr,
                            FIELD ? nullptr : a.theta + n * NG, ok ? p : 0, ok, cx, cy, cz, rx, ry, rz, bx, by, bz);
    LGeo<DIM> G;
    lean_geo<DIM>(cx, cy, cz, a.g, G);
    bool hand = false;
    float w0[NZ][2], w1[NZ][2], ws[NZ][2];
#pragma unroll
    for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float wr = (dy ? G.y.w1 : G.y.w0) * (dz ? G.z.w1 : G.z.w0);
        w0[dz][dy] = G.x.w0 * wr; w1[dz][dy] = G.x.w1 * wr; ws[dz][dy] = 0.f;
      }
    if (scatter) {
      const int key = lean_key<DIM>(G, ok);
      const int nkey = __shfl_down_sync(FULL, key, 1);
      hand = ok && G.dxo && lane < 31 && nkey == key + 4;
#pragma unroll
      for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const float got = __shfl_up_sync(FULL, hand ? w1[dz][dy] : 0.f, 1);
          ws[dz][dy] = lane ? got : 0.f;
        }
    }
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
    for (int cg = 0; cg < CG; ++cg) {
      const i64 gb = ((i64)n * CG + cg) * a.g.S;
      float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        if (GD_PK) go = (lean_opaque(reinterpret_cast<const float4*>(a.g_dst) + gb))[p];
        else {
          const float* gd = lean_opaque(a.g_dst + 4 * gb) + p;
          go = make_float4(gd[0], gd[S], gd[2 * S], gd[3 * S]);
        }
      }
      float v[2][2][2];
      {
        const float4* s = lean_opaque(reinterpret_cast<const float4*>(a.src) + gb) + G.a000;
#pragma unroll
        for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const float4 q = __ldg(s + LEAN_OFF(G, dx, dy, dz));
              v[dz][dy][dx] = q.x * go.x + q.y * go.y + q.z * go.z + q.w * go.w;
            }
      }
      float jx, jy, jz;
      lean_jacobian<DIM>(G, v, jx, jy, jz);
      ggx += jx; ggy += jy; ggz += jz;
      if (scatter) {
        // the sender ships its upstream value once and one weight per corner row; the receiver forms
        // g*w0 + g_prev*w1_prev with FMAs
        const float px = __shfl_up_sync(FULL, go.x, 1), py = __shfl_up_sync(FULL, go.y, 1);
        const float pz = __shfl_up_sync(FULL, go.z, 1), pw = __shfl_up_sync(FULL, go.w, 1);
        if (ok) {
          float4* gs = lean_opaque(reinterpret_cast<float4*>(a.g_src) + gb) + G.a000;
#pragma unroll
          for (int dz = 0; dz < NZ; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const float u = w0[dz][dy], b = ws[dz][dy];
              float4* r = gs + LEAN_OFF(G, 0, dy, dz);
              if (u != 0.f || b != 0.f)
                atomicAdd(r, make_float4(go.x * u + px * b, go.y * u + py * b, go.z * u + pz * b, go.w * u + pw * b));
              const float c = w1[dz][dy];
              if (!hand && c != 0.f) atomicAdd(r + G.dxo, make_float4(go.x * c, go.y * c, go.z * c, go.w * c));
            }
        }
      }
    }
    if (!ok) continue;
    ggx *= (float)(a.g.W - 1) / 2.f; ggy *= (float)(a.g.H - 1) / 2.f; ggz *= (float)(a.g.D - 1) / 2.f;
    if (FIELD) {
      if (a.g_phi) lean_store_gphi<DIM>(a.g_phi, nS + p, ggx, ggy, ggz, rx, ry, rz);
    } else if (THACC) {
      if (want_theta) lean_theta_acc<DIM>(acc, ggx, ggy, ggz, bx, by, bz);
    } else if (a.g_phi) {                        // affine stage: dL/d(coordinate) for lean_theta_reduce_kernel
      if (DIM == 2) reinterpret_cast<float2*>(a.g_phi)[nS + p] = make_float2(ggx, ggy);
      else reinterpret_cast<float4*>(a.g_phi)[nS + p] = make_float4(ggx, ggy, ggz, 0.f);
    }
  }
  if (want_theta && cur_n >= 0) lean_theta_flush<DIM>(acc, red, a.g_theta, cur_n);
}

// theta gradient of an affine stage from the per-voxel coordinate gradient the adjoint kernel left in `gc`
// (float2 / float4 per voxel): g_theta[n] += sum_p gc(p) (x) [base(p), 1]  (adv_affine.py:316-324 through
// F.affine_grid's adjoint).  The adjoint kernels used to carry these d(d+1) accumulators themselves: 100
// registers, two resident CTAs per SM and a block reduction behind every CTA made the affine adjoints twice as
// slow as the field adjoints of the same stencil (125.7 vs 65.1 us packed, 91.3 vs 39.6 us C = 1 at 128^3,
// gpurun_out/r02q); a streaming pass over 16 bytes per voxel costs less than the difference.
template <int DIM>
__global__ void __launch_bounds__(256)
lean_theta_reduce_kernel(Dims g, const void* __restrict__ gc, float* __restrict__ g_theta) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  constexpr int NG = DIM * (DIM + 1);
  __shared__ float red[12 * 32];
  const int n = blockIdx.y;
  const int S = (int)g.S;
  const i64 nS = (i64)n * g.S;
  float acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
  // 4 CTAs per SM; every thread keeps the loads of FOUR grid-strided voxels in flight before it touches any of
  // them (one load per trip ran at 1.7 TB/s: 18.9 us for 32 MB at 128^3, gpurun_out/r02u; more, smaller CTAs
  // instead -- 16 per SM, 3.5 trips each, a block reduction and d(d+1) atomics behind every one -- ran at 0.8
  // TB/s: 40.4 us, profiles/r02A_launch_list.md).  A voxel past the end loads nothing and adds zeros.
  const int stride = gridDim.x * 256;
  for (int p0 = blockIdx.x * 256 + threadIdx.x; p0 < S; p0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const i64 p = (i64)p0 + (i64)k * stride;               // (S can sit just below 2^31)
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < (i64)S) {
        if (DIM == 2) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(gc) + nS + p);
          v[k].x = t.x; v[k].y = t.y;
        } else {
          v[k] = __ldg(reinterpret_cast<const float4*>(gc) + nS + p);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const i64 pl = (i64)p0 + (i64)k * stride;
      const int p = pl < (i64)S ? (int)pl : S - 1;
      int x, y, z;
      voxel_xyz(g, (unsigned)p, x, y, z);
      const float bx = base_coord_s(x, g.W, g.stW, 0.f), by = base_coord_s(y, g.H, g.stH, 0.f);
      const float bz = DIM == 3 ? base_coord_s(z, g.D, g.stD, 0.f) : 0.f;
      lean_theta_acc<DIM>(acc, v[k].x, v[k].y, DIM == 3 ? v[k].z : 0.f, bx, by, bz);
    }
  }
  lean_theta_flush<DIM>(acc, red, g_theta, n);
}

// planar [N*C][S] -> packed [N*CG][S] float4 (the chain input of the prediction path, once per pass)
__global__ void __launch_bounds__(256)
lean_pack_kernel(const float* __restrict__ src, float4* __restrict__ dst, int S) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= S) return;
  const i64 gb = (i64)blockIdx.y * S;
  const float* s = src + 4 * gb + p;
  dst[gb + p] = make_float4(s[0], s[S], s[2 * S], s[3 * S]);
}

// packed [N*CG][S] float4 -> planar [N*C][S]
__global__ void __launch_bounds__(256)
lean_unpack_kernel(const float4* __restrict__ src, float* __restrict__ dst, int S) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= S) return;
  const i64 gb = (i64)blockIdx.y * S;
  const float4 v = src[gb + p];
  float* d = dst + 4 * gb + p;
  d[0] = v.x; d[S] = v.y; d[2 * S] = v.z; d[3 * S] = v.w;
}

}  // namespace advk
