// advk_chain.cu -- the fused chain-apply kernels: ONE persistent cooperative launch applies a
// whole chain of transforms (forward), ONE applies its adjoint (backward).
//
// A chain is a short program of stages over an N x C x S tensor:
//     INTENSITY   x -> x + eps*delta, x * bias(cp)        adv_noise.py:79-90, adv_bias.py:152-188
//     WARP_FIELD  x -> grid_sample(x, phi)                adv_morph.py:524-558
//     WARP_AFFINE x -> grid_sample(x, affine_grid(theta)) adv_affine.py:289-314
// in any order (the reference's solver.forward / predict_backward loops, adv_compose_solver.py:
// 148-176, 210-219, issue one ATen call sequence per transform).  Two resampling stages in a row
// are a gather of a gather with unbounded displacement, so each stage needs the COMPLETE output
// of the previous one; the stages therefore run as phases of one grid separated by grid-wide
// barriers, with the (N x C x S) intermediates living in L2 (8 MB per channel at 128^3 against
// 126 MB of L2).  The same executor runs
//   * the image chain            noise -> bias -> morph -> affine (+ the if_norm_image clamp),
//     with a one-channel "valid region" mask riding along the geometric stages, and
//   * the prediction warp-back   affine^-1 -> morph(-v) on K channels + the mask, binarised.
// Backward: zero the scatter targets, then walk the stages in reverse; a warp stage scatters the
// upstream gradient to its source (fp32 RED), computes the spatial Jacobian from the stashed
// source, and reduces it to d x (d+1) per sample (affine) or writes it per voxel (field).
//
// Work distribution: the flattened (sample, voxel) space is cut into tiles of 256 voxels that
// never straddle a sample; block b owns a contiguous range of tiles (good L1 locality for the
// gathers, and per-sample reductions are flushed only when the sample changes).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "advk_intensity.cuh"

namespace cg = cooperative_groups;

namespace advk {

constexpr int CT = 256;
constexpr int MAX_STAGES = ADVK_CHAIN_MAX_STAGES;

struct Stage {
  int kind;
  // intensity
  int order; float ns; int use_ig; float ig; BiasCfg b;
  const float* delta; const float* low;
  // warp
  const void* phi; const float* theta; int pad, interp; const float* pv;
  // data path
  const float* src; float* dst;
  const float* msrc; float* mdst; int binarize;
  int src_pk, dst_pk;        // channel-packed (float4 per voxel per group of 4 channels) source / destination
  int src_vm, dst_vm;        // lean path, C == 1: float2 {value, mask} per voxel (see advk_chain_lean.cuh)
  // adjoint
  const float* g_dst; float* g_src; int zero_g_src;
  float* g_src_user;         // packed mode, first stage: planar user buffer the packed g_src is unpacked into
  float* g_delta; float* g_up; void* g_phi; float* g_theta;
};

struct Program {
  Dims g; int C; int n; int first_bwd;
  int do_clamp; float lo, hi; int want_mask;
  int pack;         // 0 planar intermediates; 4: intermediates and scatter targets channel-packed
  int tps;          // tiles per sample
  FastDiv ftps;
  int interleave;   // 0: block owns a contiguous tile range; 1: tiles dealt round-robin
  i64 n_tiles;
  int lean;         // runs the specialised per-stage kernels (advk_chain_lean.cuh)
  const float* user_src; float* src_packed;   // lean + pack: the chain input and its packed copy (stash)
  float* g_coord;   // adjoint, lean: per-voxel coordinate gradient of an affine stage (tail of `scratch`)
  Stage st[MAX_STAGES];
};

}  // namespace advk
#include "advk_chain_lean.cuh"
namespace advk {

template <int DIM> struct Stencil { Axis x, y, z; };

template <int DIM>
__device__ __forceinline__ Stencil<DIM> make_stencil(float cx, float cy, float cz, const Dims& g, int pad, int interp) {
  Stencil<DIM> s;
  s.x = make_axis(cx, g.W, pad, interp);
  s.y = make_axis(cy, g.H, pad, interp);
  if (DIM == 3) s.z = make_axis(cz, g.D, pad, interp);
  else { s.z.i0 = 0; s.z.w0 = 1.f; s.z.w1 = 0.f; s.z.v0 = true; s.z.v1 = false; s.z.mult = 0.f; }
  return s;
}

// The 2^d corners of one sampling stencil, resolved ONCE per voxel and reused by every channel:
// 32-bit voxel offsets inside the sample (S < 2^31, host-checked), interpolation weights in ATen's
// multiplication order (wx*wy)*wz, and a validity bit per corner.  Corner k: dx = k&1, dy = (k>>1)&1,
// dz = k>>2 -- the order the gathers are summed in (z outer, x inner, like grid_sample).
template <int DIM> struct Corners {
  static constexpr int NC = (DIM == 3) ? 8 : 4;
  int off[NC];
  float w[NC];
  unsigned mask;
};

template <int DIM>
__device__ __forceinline__ void make_corners(const Stencil<DIM>& s, const Dims& g, bool ok, Corners<DIM>& t) {
  const int HW = g.H * g.W;
  const int base = s.z.i0 * HW + s.y.i0 * g.W + s.x.i0;
  t.mask = 0u;
#pragma unroll
  for (int k = 0; k < Corners<DIM>::NC; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const bool v = ok && (dx ? s.x.v1 : s.x.v0) && (dy ? s.y.v1 : s.y.v0) && (dz ? s.z.v1 : s.z.v0);
    t.off[k] = base + dx + dy * g.W + dz * HW;
    t.w[k] = ((dx ? s.x.w1 : s.x.w0) * (dy ? s.y.w1 : s.y.w0)) * (dz ? s.z.w1 : s.z.w0);
    t.mask |= (v ? 1u : 0u) << k;
  }
}

struct Vox { int n, x, y, z; int p; bool ok; };

__device__ __forceinline__ Vox tile_voxel(const Program& P, unsigned t) {
  Vox v;
  const unsigned n = fast_div(t, P.ftps);
  v.n = (int)n;
  v.p = (int)((t - n * (unsigned)P.tps) * CT + threadIdx.x);
  v.ok = v.p < P.g.S;
  voxel_xyz(P.g, (unsigned)v.p, v.x, v.y, v.z);
  return v;
}

// tile loop: for (t = t0; t < t1; t += dt)
__device__ __forceinline__ void tile_range(const Program& P, unsigned& t0, unsigned& t1, unsigned& dt) {
  const unsigned nt = (unsigned)P.n_tiles;
  if (P.interleave) {
    t0 = blockIdx.x; t1 = nt; dt = gridDim.x;
  } else {
    const unsigned per = (nt + gridDim.x - 1) / gridDim.x;
    t0 = blockIdx.x * per;
    t1 = t0 + per < nt ? t0 + per : nt;
    dt = 1;
  }
}

// sampling coordinates of output voxel v; r* = raw field values (for the clamp-range mask),
// b* = base coordinates (affine only)
template <int DIM, bool FIELD>
__device__ __forceinline__ void stage_coords(const Program& P, const Stage& s, const Vox& v, float& cx, float& cy,
                                             float& cz, float& rx, float& ry, float& rz, float& bx, float& by,
                                             float& bz) {
  const Dims& g = P.g;
  if (FIELD) {
    rx = ry = rz = 0.f;
    if (v.ok) {
      if (DIM == 2) {
        float2 f = __ldg(reinterpret_cast<const float2*>(s.phi) + ((i64)v.n * g.S + v.p));
        rx = f.x; ry = f.y;
      } else {
        float4 f = __ldg(reinterpret_cast<const float4*>(s.phi) + ((i64)v.n * g.S + v.p));
        rx = f.x; ry = f.y; rz = f.z;
      }
    }
    cx = clampf(rx, -1.f, 1.f); cy = clampf(ry, -1.f, 1.f); cz = clampf(rz, -1.f, 1.f);
    bx = by = bz = 0.f;
  } else {
    bx = base_coord_s(v.x, g.W, g.stW, 0.f); by = base_coord_s(v.y, g.H, g.stH, 0.f);
    if (DIM == 2) {
      const float* t = s.theta + v.n * 6;
      // the K = 3 dot product of affine_grid's bmm, accumulated in order with one rounding per step
      cx = __fadd_rn(__fmaf_rn(t[1], by, __fmul_rn(t[0], bx)), t[2]);
      cy = __fadd_rn(__fmaf_rn(t[4], by, __fmul_rn(t[3], bx)), t[5]);
      cz = 0.f; bz = 0.f;
    } else {
      bz = base_coord_s(v.z, g.D, g.stD, 0.f);
      const float* t = s.theta + v.n * 12;
      cx = __fadd_rn(__fmaf_rn(t[2], bz, __fmaf_rn(t[1], by, __fmul_rn(t[0], bx))), t[3]);
      cy = __fadd_rn(__fmaf_rn(t[6], bz, __fmaf_rn(t[5], by, __fmul_rn(t[4], bx))), t[7]);
      cz = __fadd_rn(__fmaf_rn(t[10], bz, __fmaf_rn(t[9], by, __fmul_rn(t[8], bx))), t[11]);
    }
    rx = cx; ry = cy; rz = cz;
  }
}

// ------------------------------------------------------------------------------------ forward

template <int DIM>
__device__ void stage_intensity_fwd(const Program& P, const Stage& s, bool last) {
  const Dims& g = P.g;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (!v.ok) continue;
    float bv = 1.f;
    if (s.order != 0) {
      float braw; bool pass;
      bv = bias_value(s.b, bias_up<DIM>(s.b, s.low + (i64)v.n * s.b.lD * s.b.lH * s.b.lW, v.z, v.y, v.x, g, v.p), braw, pass);
    }
    i64 q = (i64)v.n * P.C * g.S + v.p;
    for (int c = 0; c < P.C; ++c, q += g.S) {
      float val = intensity_point(s.order, s.src[q], s.order != 1 ? s.delta[q] : 0.f, s.ns, bv, s.use_ig, s.ig);
      if (clamp) val = clampf(val, P.lo, P.hi);
      s.dst[q] = val;
    }
  }
}

template <int DIM>
__device__ __forceinline__ float gather_corners(const Corners<DIM>& ct, const float* __restrict__ src, float pv) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < Corners<DIM>::NC; ++k)
    if (ct.mask & (1u << k)) acc += (__ldg(src + ct.off[k]) - pv) * ct.w[k];
  return acc + pv;
}

template <int DIM, bool FIELD, bool ZL>
__device__ void stage_warp_fwd(const Program& P, const Stage& s, bool last) {
  const Dims& g = P.g;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (!v.ok) continue;
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    stage_coords<DIM, FIELD>(P, s, v, cx, cy, cz, rx, ry, rz, bx, by, bz);
    Stencil<DIM> st = make_stencil<DIM>(cx, cy, cz, g, ZL ? (int)ADVK_PAD_ZEROS : s.pad, ZL ? (int)ADVK_INTERP_LINEAR : s.interp);
    Corners<DIM> ct;
    make_corners<DIM>(st, g, true, ct);
    const float pv = s.pv ? s.pv[v.n] : 0.f;
    const float* src = s.src + (i64)v.n * P.C * g.S;
    float* dst = s.dst + (i64)v.n * P.C * g.S + v.p;
    for (int c = 0; c < P.C; ++c, src += g.S, dst += g.S) {
      float val = gather_corners<DIM>(ct, src, pv);
      if (clamp) val = clampf(val, P.lo, P.hi);
      *dst = val;
    }
    if (s.mdst) {
      float m;
      if (s.msrc) m = gather_corners<DIM>(ct, s.msrc + (i64)v.n * g.S, pv);
      else {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < Corners<DIM>::NC; ++k)
          if (ct.mask & (1u << k)) acc += (1.f - pv) * ct.w[k];
        m = acc + pv;
      }
      if (s.binarize) m = (m != 0.f) ? 1.f : 0.f;
      s.mdst[(i64)v.n * g.S + v.p] = m;
    }
  }
}

// ---- channel-packed variants (C % 4 == 0, warp stages only: the K-class prediction path) -----------
// Intermediates are stored as one float4 per voxel per group of 4 channels, so a corner is ONE 16-byte
// gather / ONE vector RED for 4 channels instead of 4 scalar ones; only the user-facing tensors (chain
// input, output, their gradients) stay planar N x C x S.
__device__ __forceinline__ float4 f4_fma(float4 a, float w, float4 acc) {
  return make_float4(acc.x + a.x * w, acc.y + a.y * w, acc.z + a.z * w, acc.w + a.w * w);
}

template <int DIM>
__device__ __forceinline__ float4 load_corner4(const Stage& s, bool packed, const float* __restrict__ base, i64 S,
                                               int off, float pv) {
  float4 v;
  if (packed) v = __ldg(reinterpret_cast<const float4*>(base) + off);
  else v = make_float4(__ldg(base + off), __ldg(base + S + off), __ldg(base + 2 * S + off), __ldg(base + 3 * S + off));
  return make_float4(v.x - pv, v.y - pv, v.z - pv, v.w - pv);
}

template <int DIM, bool FIELD, bool ZL>
__device__ void stage_warp_fwd_pk(const Program& P, const Stage& s, bool last) {
  const Dims& g = P.g;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  const int CG = P.C >> 2;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (!v.ok) continue;
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    stage_coords<DIM, FIELD>(P, s, v, cx, cy, cz, rx, ry, rz, bx, by, bz);
    Stencil<DIM> st = make_stencil<DIM>(cx, cy, cz, g, ZL ? (int)ADVK_PAD_ZEROS : s.pad, ZL ? (int)ADVK_INTERP_LINEAR : s.interp);
    Corners<DIM> ct;
    make_corners<DIM>(st, g, true, ct);
    const float pv = s.pv ? s.pv[v.n] : 0.f;
    for (int cg = 0; cg < CG; ++cg) {
      // group base: packed = float4 index (n*CG+cg)*S, planar = channel (n*C + 4cg)
      const float* base = s.src_pk ? s.src + 4 * ((i64)v.n * CG + cg) * g.S : s.src + ((i64)v.n * P.C + 4 * cg) * g.S;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < Corners<DIM>::NC; ++k)
        if (ct.mask & (1u << k)) acc = f4_fma(load_corner4<DIM>(s, s.src_pk, base, g.S, ct.off[k], pv), ct.w[k], acc);
      float4 val = make_float4(acc.x + pv, acc.y + pv, acc.z + pv, acc.w + pv);
      if (clamp) val = make_float4(clampf(val.x, P.lo, P.hi), clampf(val.y, P.lo, P.hi), clampf(val.z, P.lo, P.hi), clampf(val.w, P.lo, P.hi));
      if (s.dst_pk) {
        reinterpret_cast<float4*>(s.dst)[((i64)v.n * CG + cg) * g.S + v.p] = val;
      } else {
        float* d = s.dst + ((i64)v.n * P.C + 4 * cg) * g.S + v.p;
        d[0] = val.x; d[g.S] = val.y; d[2 * g.S] = val.z; d[3 * g.S] = val.w;
      }
    }
    if (s.mdst) {
      float m;
      if (s.msrc) m = gather_corners<DIM>(ct, s.msrc + (i64)v.n * g.S, pv);
      else {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < Corners<DIM>::NC; ++k)
          if (ct.mask & (1u << k)) acc += (1.f - pv) * ct.w[k];
        m = acc + pv;
      }
      if (s.binarize) m = (m != 0.f) ? 1.f : 0.f;
      s.mdst[(i64)v.n * g.S + v.p] = m;
    }
  }
}

template <int DIM, bool PK>
__device__ __forceinline__ void run_stage_fwd(const Program& P, int k) {
  const Stage& s = P.st[k];
  const bool last = (k == P.n - 1);
  if (PK) {
    const bool zl = s.pad == ADVK_PAD_ZEROS && s.interp == ADVK_INTERP_LINEAR;    // the default configuration
    if (s.kind == ADVK_STAGE_WARP_FIELD) { if (zl) stage_warp_fwd_pk<DIM, true, true>(P, s, last); else stage_warp_fwd_pk<DIM, true, false>(P, s, last); }
    else { if (zl) stage_warp_fwd_pk<DIM, false, true>(P, s, last); else stage_warp_fwd_pk<DIM, false, false>(P, s, last); }
    return;
  }
  if (s.kind == ADVK_STAGE_INTENSITY) stage_intensity_fwd<DIM>(P, s, last);
  else {
    const bool zl = s.pad == ADVK_PAD_ZEROS && s.interp == ADVK_INTERP_LINEAR;
    if (s.kind == ADVK_STAGE_WARP_FIELD) { if (zl) stage_warp_fwd<DIM, true, true>(P, s, last); else stage_warp_fwd<DIM, true, false>(P, s, last); }
    else { if (zl) stage_warp_fwd<DIM, false, true>(P, s, last); else stage_warp_fwd<DIM, false, false>(P, s, last); }
  }
}

template <int DIM, int MINB, bool PK>
__global__ void __launch_bounds__(CT, MINB)
chain_fwd_kernel(const __grid_constant__ Program P) {
  cg::grid_group grid = cg::this_grid();
  for (int k = 0; k < P.n; ++k) {
    if (k) grid.sync();
    run_stage_fwd<DIM, PK>(P, k);
  }
}

// one stage per launch (used when a cooperative launch is not possible / for A-B timing)
template <int DIM, int MINB, bool PK>
__global__ void __launch_bounds__(CT, MINB)
chain_fwd_stage_kernel(const __grid_constant__ Program P, int k) {
  run_stage_fwd<DIM, PK>(P, k);
}

// ------------------------------------------------------------------------------------ backward

__device__ void zero_buffers(const Program& P) {
  const i64 n4 = ((i64)P.g.N * P.C * P.g.S) / 4;     // host guarantees N*C*S % 4 == 0 or handles tail
  const i64 tail0 = n4 * 4, tot = (i64)P.g.N * P.C * P.g.S;
  for (int k = P.first_bwd; k < P.n; ++k) {
    const Stage& s = P.st[k];
    if (!s.zero_g_src || !s.g_src) continue;
    float4* d4 = reinterpret_cast<float4*>(s.g_src);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (i64)gridDim.x * blockDim.x)
      d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (i64 i = tail0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (i64)gridDim.x * blockDim.x)
      s.g_src[i] = 0.f;
  }
}

template <int DIM>
__device__ void stage_intensity_bwd(const Program& P, const Stage& s, bool last) {
  const Dims& g = P.g;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (!v.ok) continue;
    float bv = 1.f, braw = 1.f;
    bool pass = true;
    if (s.order != 0)
      bv = bias_value(s.b, bias_up<DIM>(s.b, s.low + (i64)v.n * s.b.lD * s.b.lH * s.b.lW, v.z, v.y, v.x, g, v.p), braw, pass);
    float gb = 0.f;
    i64 q = (i64)v.n * P.C * g.S + v.p;
    for (int c = 0; c < P.C; ++c, q += g.S) {
      float go = s.g_dst[q];
      const float x0 = s.src[q];
      const float dl = (s.order != 1) ? s.delta[q] : 0.f;
      if (clamp) {
        float val = intensity_point(s.order, x0, dl, s.ns, bv, s.use_ig, s.ig);
        if (!(val >= P.lo && val <= P.hi)) go = 0.f;
      }
      float gd;
      float gi = intensity_point_bwd(s.order, go, x0, dl, s.ns, bv, s.use_ig, s.ig, gd, gb);
      if (s.g_delta) s.g_delta[q] = gd;
      if (s.g_src) s.g_src[q] = gi;
    }
    if (s.g_up) s.g_up[(i64)v.n * g.S + v.p] = pass ? gb * (s.b.use_log ? braw : 1.f) : 0.f;
  }
}

// Adjoint of a warp stage.  Per voxel the corner table is built once; per channel the kernel loads the
// 2^d source values (needed for the coordinate gradient and the clamp mask) and scatters go*w.  The
// scatter is lane-combined like the squaring-step backward: along a corner row the x1 corner of lane i
// is the x0 corner of lane i+1 for (near-)unit x spacing, so one shuffle per (row, channel) hands that
// contribution over and the RED count halves; the hand-off pattern is decided once per voxel and shared
// by all channels.  The coordinate gradient sum_c (src_c - pv) * go_c is accumulated per corner and
// turned into d/dx, d/dy, d/dz after the channel loop.
template <int DIM, bool FIELD, bool ZL>
__device__ void stage_warp_bwd(const Program& P, const Stage& s, bool last, float* red) {
  constexpr int NG = DIM * (DIM + 1);
  constexpr int NC = Corners<DIM>::NC;
  constexpr int NR = NC / 2;
  const unsigned FULL = 0xffffffffu;
  const Dims& g = P.g;
  const int lane = threadIdx.x & 31;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  const bool want_theta = !FIELD && s.g_theta != nullptr;
  float acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
  int cur_n = -1;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (want_theta && v.n != cur_n) {            // block-uniform
      if (cur_n >= 0) {
        block_sum<NG>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
          for (int i = 0; i < NG; ++i) atomicAdd(s.g_theta + cur_n * NG + i, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < NG; ++i) acc[i] = 0.f;
      }
      cur_n = v.n;
    }
    // no early exit for !v.ok: the whole warp takes part in the shuffles below
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    stage_coords<DIM, FIELD>(P, s, v, cx, cy, cz, rx, ry, rz, bx, by, bz);
    Stencil<DIM> st = make_stencil<DIM>(cx, cy, cz, g, ZL ? (int)ADVK_PAD_ZEROS : s.pad, ZL ? (int)ADVK_INTERP_LINEAR : s.interp);
    Corners<DIM> ct;
    make_corners<DIM>(st, g, v.ok, ct);
    const float pv = s.pv ? s.pv[v.n] : 0.f;
    const bool scatter = s.g_src != nullptr;
    unsigned hand = 0u;                            // bit r: row r's x1 contribution goes to lane+1
    if (scatter) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int k0 = 2 * r, k1 = k0 + 1;
        const int a0_next = __shfl_down_sync(FULL, (ct.mask & (1u << k0)) ? ct.off[k0] : -1, 1);
        if ((ct.mask & (1u << k1)) && lane < 31 && a0_next == ct.off[k1]) hand |= 1u << r;
      }
    }
    float tk[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) tk[k] = 0.f;
    const float* src = s.src + (i64)v.n * P.C * g.S;
    const float* gd = s.g_dst + (i64)v.n * P.C * g.S + v.p;
    float* gs = scatter ? s.g_src + (i64)v.n * P.C * g.S : nullptr;
    for (int c = 0; c < P.C; ++c, src += g.S, gd += g.S) {
      float go = v.ok ? *gd : 0.f;
      float val[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) val[k] = (ct.mask & (1u << k)) ? __ldg(src + ct.off[k]) - pv : 0.f;
      if (clamp) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += val[k] * ct.w[k];
        a += pv;
        if (!(a >= P.lo && a <= P.hi)) go = 0.f;
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) tk[k] += val[k] * go;
      if (scatter) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int k0 = 2 * r, k1 = k0 + 1;
          const bool h = (hand >> r) & 1u;
          float c0 = go * ct.w[k0];
          const float c1 = go * ct.w[k1];
          const float rcv = __shfl_up_sync(FULL, h ? c1 : 0.f, 1);
          if (lane > 0) c0 += rcv;
          if (ct.mask & (1u << k0)) atomicAdd(gs + ct.off[k0], c0);
          if ((ct.mask & (1u << k1)) && !h) atomicAdd(gs + ct.off[k1], c1);
        }
        gs += g.S;
      }
    }
    if (!v.ok) continue;
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
      const float wx = dx ? st.x.w1 : st.x.w0, wy = dy ? st.y.w1 : st.y.w0, wz = dz ? st.z.w1 : st.z.w0;
      ggx += (dx ? tk[k] : -tk[k]) * (wy * wz);
      ggy += (dy ? tk[k] : -tk[k]) * (wx * wz);
      if (DIM == 3) ggz += (dz ? tk[k] : -tk[k]) * (wx * wy);
    }
    ggx *= st.x.mult; ggy *= st.y.mult; ggz *= st.z.mult;
    if (FIELD) {
      if (s.g_phi) {
        // consumers clamp the stored field to [-1,1]; torch.clamp passes gradient on the closed interval
        if (!(rx >= -1.f && rx <= 1.f)) ggx = 0.f;
        if (!(ry >= -1.f && ry <= 1.f)) ggy = 0.f;
        if (DIM == 2) {
          reinterpret_cast<float2*>(s.g_phi)[(i64)v.n * g.S + v.p] = make_float2(ggx, ggy);
        } else {
          if (!(rz >= -1.f && rz <= 1.f)) ggz = 0.f;
          reinterpret_cast<float4*>(s.g_phi)[(i64)v.n * g.S + v.p] = make_float4(ggx, ggy, ggz, 0.f);
        }
      }
    } else if (want_theta) {
      if (DIM == 2) {
        acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx;
        acc[3] += ggy * bx; acc[4] += ggy * by; acc[5] += ggy;
      } else {
        acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx * bz; acc[3] += ggx;
        acc[4] += ggy * bx; acc[5] += ggy * by; acc[6] += ggy * bz; acc[7] += ggy;
        acc[8] += ggz * bx; acc[9] += ggz * by; acc[10] += ggz * bz; acc[11] += ggz;
      }
    }
  }
  if (want_theta && cur_n >= 0) {
    block_sum<NG>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NG; ++i) atomicAdd(s.g_theta + cur_n * NG + i, acc[i]);
    }
  }
}

template <int DIM, bool FIELD, bool ZL>
__device__ void stage_warp_bwd_pk(const Program& P, const Stage& s, bool last, float* red) {
  constexpr int NG = DIM * (DIM + 1);
  constexpr int NC = Corners<DIM>::NC;
  constexpr int NR = NC / 2;
  const unsigned FULL = 0xffffffffu;
  const Dims& g = P.g;
  const int lane = threadIdx.x & 31;
  const int CG = P.C >> 2;
  unsigned t0, t1, dt;
  tile_range(P, t0, t1, dt);
  const bool clamp = last && P.do_clamp;
  const bool want_theta = !FIELD && s.g_theta != nullptr;
  const bool gd_pk = !last;                       // upstream of the last stage is the user's planar g_out
  float acc[NG];
#pragma unroll
  for (int i = 0; i < NG; ++i) acc[i] = 0.f;
  int cur_n = -1;
  for (unsigned t = t0; t < t1; t += dt) {
    Vox v = tile_voxel(P, t);
    if (want_theta && v.n != cur_n) {            // block-uniform
      if (cur_n >= 0) {
        block_sum<NG>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
          for (int i = 0; i < NG; ++i) atomicAdd(s.g_theta + cur_n * NG + i, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < NG; ++i) acc[i] = 0.f;
      }
      cur_n = v.n;
    }
    float cx, cy, cz, rx, ry, rz, bx, by, bz;
    stage_coords<DIM, FIELD>(P, s, v, cx, cy, cz, rx, ry, rz, bx, by, bz);
    Stencil<DIM> st = make_stencil<DIM>(cx, cy, cz, g, ZL ? (int)ADVK_PAD_ZEROS : s.pad, ZL ? (int)ADVK_INTERP_LINEAR : s.interp);
    Corners<DIM> ct;
    make_corners<DIM>(st, g, v.ok, ct);
    const float pv = s.pv ? s.pv[v.n] : 0.f;
    const bool scatter = s.g_src != nullptr;
    unsigned hand = 0u;
    if (scatter) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int k0 = 2 * r, k1 = k0 + 1;
        const int a0_next = __shfl_down_sync(FULL, (ct.mask & (1u << k0)) ? ct.off[k0] : -1, 1);
        if ((ct.mask & (1u << k1)) && lane < 31 && a0_next == ct.off[k1]) hand |= 1u << r;
      }
    }
    float tk[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) tk[k] = 0.f;
    for (int cg = 0; cg < CG; ++cg) {
      const float* base = s.src_pk ? s.src + 4 * ((i64)v.n * CG + cg) * g.S : s.src + ((i64)v.n * P.C + 4 * cg) * g.S;
      float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v.ok) {
        if (gd_pk) go = reinterpret_cast<const float4*>(s.g_dst)[((i64)v.n * CG + cg) * g.S + v.p];
        else {
          const float* gd = s.g_dst + ((i64)v.n * P.C + 4 * cg) * g.S + v.p;
          go = make_float4(gd[0], gd[g.S], gd[2 * g.S], gd[3 * g.S]);
        }
      }
      if (clamp) {                                 // last stage of an image chain with C % 4 == 0 (rare)
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NC; ++k)
          if (ct.mask & (1u << k)) a = f4_fma(load_corner4<DIM>(s, s.src_pk, base, g.S, ct.off[k], pv), ct.w[k], a);
        if (!(a.x + pv >= P.lo && a.x + pv <= P.hi)) go.x = 0.f;
        if (!(a.y + pv >= P.lo && a.y + pv <= P.hi)) go.y = 0.f;
        if (!(a.z + pv >= P.lo && a.z + pv <= P.hi)) go.z = 0.f;
        if (!(a.w + pv >= P.lo && a.w + pv <= P.hi)) go.w = 0.f;
      }
      float4* gs = scatter ? reinterpret_cast<float4*>(s.g_src) + ((i64)v.n * CG + cg) * g.S : nullptr;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int k0 = 2 * r, k1 = k0 + 1;
        const bool m0 = ct.mask & (1u << k0), m1 = ct.mask & (1u << k1);
        if (m0) {
          const float4 a = load_corner4<DIM>(s, s.src_pk, base, g.S, ct.off[k0], pv);
          tk[k0] += a.x * go.x + a.y * go.y + a.z * go.z + a.w * go.w;
        }
        if (m1) {
          const float4 a = load_corner4<DIM>(s, s.src_pk, base, g.S, ct.off[k1], pv);
          tk[k1] += a.x * go.x + a.y * go.y + a.z * go.z + a.w * go.w;
        }
        if (scatter) {
          const bool h = (hand >> r) & 1u;
          const float w0 = ct.w[k0], w1 = ct.w[k1];
          float4 c0 = make_float4(go.x * w0, go.y * w0, go.z * w0, go.w * w0);
          const float4 c1 = make_float4(go.x * w1, go.y * w1, go.z * w1, go.w * w1);
          const float r0 = __shfl_up_sync(FULL, h ? c1.x : 0.f, 1);
          const float r1 = __shfl_up_sync(FULL, h ? c1.y : 0.f, 1);
          const float r2 = __shfl_up_sync(FULL, h ? c1.z : 0.f, 1);
          const float r3 = __shfl_up_sync(FULL, h ? c1.w : 0.f, 1);
          if (lane > 0) { c0.x += r0; c0.y += r1; c0.z += r2; c0.w += r3; }
          if (m0) atomicAdd(gs + ct.off[k0], c0);
          if (m1 && !h) atomicAdd(gs + ct.off[k1], c1);
        }
      }
    }
    if (!v.ok) continue;
    float ggx = 0.f, ggy = 0.f, ggz = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
      const float wx = dx ? st.x.w1 : st.x.w0, wy = dy ? st.y.w1 : st.y.w0, wz = dz ? st.z.w1 : st.z.w0;
      ggx += (dx ? tk[k] : -tk[k]) * (wy * wz);
      ggy += (dy ? tk[k] : -tk[k]) * (wx * wz);
      if (DIM == 3) ggz += (dz ? tk[k] : -tk[k]) * (wx * wy);
    }
    ggx *= st.x.mult; ggy *= st.y.mult; ggz *= st.z.mult;
    if (FIELD) {
      if (s.g_phi) {
        if (!(rx >= -1.f && rx <= 1.f)) ggx = 0.f;
        if (!(ry >= -1.f && ry <= 1.f)) ggy = 0.f;
        if (DIM == 2) {
          reinterpret_cast<float2*>(s.g_phi)[(i64)v.n * g.S + v.p] = make_float2(ggx, ggy);
        } else {
          if (!(rz >= -1.f && rz <= 1.f)) ggz = 0.f;
          reinterpret_cast<float4*>(s.g_phi)[(i64)v.n * g.S + v.p] = make_float4(ggx, ggy, ggz, 0.f);
        }
      }
    } else if (want_theta) {
      if (DIM == 2) {
        acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx;
        acc[3] += ggy * bx; acc[4] += ggy * by; acc[5] += ggy;
      } else {
        acc[0] += ggx * bx; acc[1] += ggx * by; acc[2] += ggx * bz; acc[3] += ggx;
        acc[4] += ggy * bx; acc[5] += ggy * by; acc[6] += ggy * bz; acc[7] += ggy;
        acc[8] += ggz * bx; acc[9] += ggz * by; acc[10] += ggz * bz; acc[11] += ggz;
      }
    }
  }
  if (want_theta && cur_n >= 0) {
    block_sum<NG>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NG; ++i) atomicAdd(s.g_theta + cur_n * NG + i, acc[i]);
    }
  }
}

// packed g_src of the first stage -> the user's planar N x C x S gradient
__device__ void unpack_g_src(const Program& P) {
  const Stage& s = P.st[P.first_bwd];
  if (!P.pack || !s.g_src_user || !s.g_src) return;
  const Dims& g = P.g;
  const int CG = P.C >> 2;
  const i64 tot = (i64)g.N * CG * g.S;
  const float4* src = reinterpret_cast<const float4*>(s.g_src);
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (i64)gridDim.x * blockDim.x) {
    const i64 ng = i / g.S, p = i - ng * g.S;       // one 64-bit divide per element of a streaming pass
    const float4 v = src[i];
    float* d = s.g_src_user + (ng * 4) * g.S + p;    // channel (n*C + 4cg) = 4*(n*CG+cg)
    d[0] = v.x; d[g.S] = v.y; d[2 * g.S] = v.z; d[3 * g.S] = v.w;
  }
}

template <int DIM, bool PK>
__device__ __forceinline__ void run_stage_bwd(const Program& P, int k, float* red) {
  const Stage& s = P.st[k];
  const bool last = (k == P.n - 1);
  if (PK) {
    const bool zl = s.pad == ADVK_PAD_ZEROS && s.interp == ADVK_INTERP_LINEAR;
    if (s.kind == ADVK_STAGE_WARP_FIELD) { if (zl) stage_warp_bwd_pk<DIM, true, true>(P, s, last, red); else stage_warp_bwd_pk<DIM, true, false>(P, s, last, red); }
    else { if (zl) stage_warp_bwd_pk<DIM, false, true>(P, s, last, red); else stage_warp_bwd_pk<DIM, false, false>(P, s, last, red); }
    return;
  }
  if (s.kind == ADVK_STAGE_INTENSITY) stage_intensity_bwd<DIM>(P, s, last);
  else {
    const bool zl = s.pad == ADVK_PAD_ZEROS && s.interp == ADVK_INTERP_LINEAR;
    if (s.kind == ADVK_STAGE_WARP_FIELD) { if (zl) stage_warp_bwd<DIM, true, true>(P, s, last, red); else stage_warp_bwd<DIM, true, false>(P, s, last, red); }
    else { if (zl) stage_warp_bwd<DIM, false, true>(P, s, last, red); else stage_warp_bwd<DIM, false, false>(P, s, last, red); }
  }
}

template <int DIM, int MINB, bool PK>
__global__ void __launch_bounds__(CT, MINB)
chain_bwd_kernel(const __grid_constant__ Program P) {
  __shared__ float red[12 * 32];
  cg::grid_group grid = cg::this_grid();
  zero_buffers(P);
  for (int k = P.n - 1; k >= P.first_bwd; --k) {
    grid.sync();
    run_stage_bwd<DIM, PK>(P, k, red);
  }
  if (PK && P.st[P.first_bwd].g_src_user) {
    grid.sync();
    unpack_g_src(P);
  }
}

template <int DIM, int MINB, bool PK>
__global__ void __launch_bounds__(CT, MINB)
chain_bwd_stage_kernel(const __grid_constant__ Program P, int k) {
  __shared__ float red[12 * 32];
  if (k == -1) zero_buffers(P);
  else if (k == -2) unpack_g_src(P);
  else run_stage_bwd<DIM, PK>(P, k, red);
}

// ------------------------------------------------------------------------------------ host

static int g_coop = -1;   // -1: decide on first use
static int g_pack = -1;

static bool pack_enabled() {
  if (g_pack < 0) {
    const char* e = getenv("ADVK_CHAIN_PACK");
    g_pack = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pack == 1;
}

static bool coop_enabled() {
  int& v = g_coop;
  if (v < 0) {
    const char* e = getenv("ADVK_CHAIN_COOP");
    int dev = 0, sup = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sup, cudaDevAttrCooperativeLaunch, dev);
    v = (sup && !(e && e[0] == '0')) ? 1 : 0;
  }
  return v == 1;
}

template <typename K>
static int coop_grid(K kernel) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CT, 0);
  if (occ < 1) occ = 1;
  return sms * occ;
}

static int g_lean = -1;
static bool lean_enabled() {
  if (g_lean < 0) {
    const char* e = getenv("ADVK_CHAIN_LEAN");
    g_lean = (e && e[0] == '0') ? 0 : 1;
  }
  return g_lean == 1;
}

// zeros padding + linear interpolation + no pad values on every warp stage (the defaults of the reference's
// constructors and of every BASELINE workload); anything else runs the generic executor above
static bool lean_eligible(const Program& P) {
  if (!lean_enabled() || P.g.S >= (1LL << 29) || P.g.N > 65535) return false;
  for (int k = 0; k < P.n; ++k) {
    const Stage& s = P.st[k];
    if (s.kind == ADVK_STAGE_INTENSITY) continue;
    if (s.pad != ADVK_PAD_ZEROS || s.interp != ADVK_INTERP_LINEAR || s.pv) return false;
  }
  if (P.pack && P.do_clamp) return false;
  return true;
}

// intermediates are laid out in slots rounded up to 4 floats so that every slot is 16-byte aligned
static inline i64 slot_floats(i64 n) { return (n + 3) & ~(i64)3; }

// Stash layout (floats): for k = 0 .. n-2: dst of stage k (N*C*S), followed -- if want_mask and stage k is a
// warp stage other than the last warp stage (whose mask output is `mask_out`) -- by its N*S mask; then, for the
// channel-packed prediction path, one N*C*S slot for the packed copy of the chain input.
// `bwd`: the adjoint rebuilds the forward's data-path pointers (same stash layout, masks included) but has
// no chain / mask outputs.
static bool build_program(const advk_chain_desc* d, Program& P, const float* src, const float* mask_src,
                          float* stash, float* out, float* mask_out, bool bwd = false) {
  if (!d || !make_dims(&d->g, P.g) || d->C < 1 || d->n_stages < 1 || d->n_stages > MAX_STAGES) return false;
  P.C = d->C; P.n = d->n_stages; P.first_bwd = 0;
  P.do_clamp = d->do_clamp; P.lo = d->clamp_lo; P.hi = d->clamp_hi; P.want_mask = d->want_mask;
  if (P.g.S >= 0x7fffffffLL) return false;               // 32-bit voxel offsets inside a sample
  P.tps = (int)((P.g.S + CT - 1) / CT);
  P.ftps = make_fastdiv((unsigned)P.tps);
  P.n_tiles = (i64)P.tps * P.g.N;
  if (P.n_tiles >= 0x7fffffffLL) return false;
  const i64 ncs = slot_floats((i64)P.g.N * P.C * P.g.S), ns = slot_floats((i64)P.g.N * P.g.S);
  int last_warp = -1;
  bool all_warps = true;
  for (int k = 0; k < P.n; ++k) {
    if (d->stages[k].kind != ADVK_STAGE_INTENSITY) last_warp = k;
    else all_warps = false;
  }
  P.pack = (pack_enabled() && all_warps && P.n >= 2 && (P.C % 4) == 0) ? 4 : 0;
  float* cursor = stash;
  const float* cur = src;
  const float* mcur = mask_src;
  for (int k = 0; k < P.n; ++k) {
    const advk_chain_stage& a = d->stages[k];
    Stage& s = P.st[k];
    memset(&s, 0, sizeof(s));
    s.kind = a.kind;
    if (a.kind == ADVK_STAGE_INTENSITY) {
      if (a.intensity_order < 0 || a.intensity_order > 3) return false;
      s.order = a.intensity_order; s.ns = a.noise_scale; s.use_ig = a.use_ignore; s.ig = a.ignore_value;
      if (s.order != 1 && !a.delta) return false;
      if (s.order != 0) {
        if (!make_bias(a.bias, d->g.d, s.b) || !a.low) return false;
        if (!s.b.upsample && !(s.b.lD == P.g.D && s.b.lH == P.g.H && s.b.lW == P.g.W)) return false;
      }
      s.delta = a.delta; s.low = a.low;
    } else if (a.kind == ADVK_STAGE_WARP_FIELD) {
      if (!a.field) return false;
      s.phi = a.field;
    } else if (a.kind == ADVK_STAGE_WARP_AFFINE) {
      if (!a.theta) return false;
      s.theta = a.theta;
    } else return false;
    if (a.kind != ADVK_STAGE_INTENSITY) {
      if (a.pad_mode < 0 || a.pad_mode > 2 || a.interp < 0 || a.interp > 1) return false;
      s.pad = a.pad_mode; s.interp = a.interp; s.pv = a.pad_values;
    }
    s.src = cur;
    s.src_pk = (P.pack && k > 0) ? 1 : 0;
    s.dst_pk = (P.pack && k < P.n - 1) ? 1 : 0;
    if (k == P.n - 1) s.dst = out;
    else {
      if (!stash) return false;
      s.dst = cursor;
      cursor += ncs;
    }
    cur = s.dst;
    if (P.want_mask && a.kind != ADVK_STAGE_INTENSITY) {
      s.msrc = mcur;
      if (k == last_warp) { s.mdst = mask_out; s.binarize = d->binarize_mask; if (!mask_out && !bwd) return false; }
      else { s.mdst = cursor; cursor += ns; }        // right behind the stage's data slot (see dst_vm)
      mcur = s.mdst;
    }
  }
  P.user_src = src;
  P.g_coord = nullptr;
  P.src_packed = (P.pack && stash) ? cursor : nullptr;
  P.lean = lean_eligible(P) ? 1 : 0;
  if (P.lean && P.C == 1 && !P.pack) {
    // value + mask interleaved between two consecutive warp stages (the mask slot follows the data slot)
    for (int k = 0; k + 1 < P.n; ++k) {
      Stage& s = P.st[k];
      Stage& t = P.st[k + 1];
      if (s.kind != ADVK_STAGE_INTENSITY && t.kind != ADVK_STAGE_INTENSITY && s.mdst && s.mdst == s.dst + ncs &&
          t.msrc == s.mdst) {
        s.dst_vm = 1; t.src_vm = 1;
      }
    }
  }
  return true;
}

// Generic executor: compiled for 4 (forward) / 3 (adjoint: corner table + per-corner partial sums + theta
// accumulators) resident CTAs per SM; tiles dealt round-robin (measured 5-15 % faster than contiguous ranges).
constexpr int GEN_MINB_FWD = 4, GEN_MINB_BWD = 3;

template <int DIM, int MINB, bool PK>
static int launch_fwd_p(Program& P, cudaStream_t st) {
  if (coop_enabled() && P.n > 1) {
    static int gmax = coop_grid(chain_fwd_kernel<DIM, MINB, PK>);
    int grid = (int)(P.n_tiles < gmax ? P.n_tiles : gmax);
    void* args[] = {(void*)&P};
    ADVK_LAUNCH(K_chain_fwd, st,
                cudaLaunchCooperativeKernel((void*)chain_fwd_kernel<DIM, MINB, PK>, dim3(grid), dim3(CT), args, 0, st));
  } else {
    int grid = (int)(P.n_tiles < 148 * 16 ? P.n_tiles : 148 * 16);
    for (int k = 0; k < P.n; ++k)
      ADVK_LAUNCH(K_chain_fwd_stage, st, (chain_fwd_stage_kernel<DIM, MINB, PK><<<grid, CT, 0, st>>>(P, k)));
  }
  return check_launch("chain_apply_fwd");
}
template <int DIM, int MINB>
static int launch_fwd_t(Program& P, cudaStream_t st) {
  return P.pack ? launch_fwd_p<DIM, MINB, true>(P, st) : launch_fwd_p<DIM, MINB, false>(P, st);
}

template <int DIM, int MINB, bool PK>
static int launch_bwd_p(Program& P, cudaStream_t st) {
  if (coop_enabled()) {
    static int gmax = coop_grid(chain_bwd_kernel<DIM, MINB, PK>);
    int grid = (int)(P.n_tiles < gmax ? P.n_tiles : gmax);
    void* args[] = {(void*)&P};
    ADVK_LAUNCH(K_chain_bwd, st,
                cudaLaunchCooperativeKernel((void*)chain_bwd_kernel<DIM, MINB, PK>, dim3(grid), dim3(CT), args, 0, st));
  } else {
    int grid = (int)(P.n_tiles < 148 * 16 ? P.n_tiles : 148 * 16);
    ADVK_LAUNCH(K_chain_bwd_stage, st, (chain_bwd_stage_kernel<DIM, MINB, PK><<<grid, CT, 0, st>>>(P, -1)));
    for (int k = P.n - 1; k >= P.first_bwd; --k)
      ADVK_LAUNCH(K_chain_bwd_stage, st, (chain_bwd_stage_kernel<DIM, MINB, PK><<<grid, CT, 0, st>>>(P, k)));
    if (PK && P.st[P.first_bwd].g_src_user)
      ADVK_LAUNCH(K_chain_bwd_stage, st, (chain_bwd_stage_kernel<DIM, MINB, PK><<<grid, CT, 0, st>>>(P, -2)));
  }
  return check_launch("chain_apply_bwd");
}
template <int DIM, int MINB>
static int launch_bwd_t(Program& P, cudaStream_t st) {
  return P.pack ? launch_bwd_p<DIM, MINB, true>(P, st) : launch_bwd_p<DIM, MINB, false>(P, st);
}


// ---- lean per-stage launches (advk_chain_lean.cuh) ------------------------------------------------
static void lean_fill_int(const Program& P, const Stage& s, bool last, LeanInt& a) {
  a.g = P.g; a.C = P.C; a.order = s.order; a.ns = s.ns; a.use_ig = s.use_ig; a.ig = s.ig; a.b = s.b;
  a.src = s.src; a.delta = s.delta; a.low = s.low; a.dst = s.dst;
  a.clamp = (last && P.do_clamp) ? 1 : 0; a.lo = P.lo; a.hi = P.hi;
  a.g_dst = s.g_dst; a.g_src = s.g_src; a.g_delta = s.g_delta; a.g_up = s.g_up;
}

// bias stage with an upsampled field -> the row-wise intensity kernel (ADVK_LEAN_INT_ROWS=0: the per-voxel one)
static bool lean_int_rows(const Program& P, const Stage& s) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ADVK_LEAN_INT_ROWS"); on = (e && e[0] == '0') ? 0 : 1; }
  const i64 gz = (i64)P.g.N * P.g.D;
  return on && s.order != 0 && s.b.upsample && s.b.lW <= LIR_MAX && gz <= 65535 && P.g.S < 0x7fffffffLL;
}
template <int DIM>
static dim3 lean_int_grid(const Program& P) {
  return dim3(1, (unsigned)((P.g.H + 7) / 8), (unsigned)(DIM == 3 ? P.g.N * P.g.D : P.g.N));
}

template <int DIM, bool FIELD>
static void lean_launch_warp_fwd(const Program& P, const Stage& s, int k, cudaStream_t st) {
  const bool last = (k == P.n - 1);
  LeanFwd a;
  a.g = P.g; a.C = P.C; a.src = s.src; a.dst = s.dst; a.phi = s.phi; a.theta = s.theta;
  a.msrc = s.msrc; a.mdst = s.mdst; a.binarize = s.binarize;
  a.clamp = (last && P.do_clamp) ? 1 : 0; a.lo = P.lo; a.hi = P.hi;
  const dim3 grid(blocks_for(P.g.S, 256), P.g.N);
  const int mm = s.mdst ? (s.msrc ? 2 : 1) : 0;
  if (!P.pack) {
#define ADVK_LF(MM, VS, VD) ADVK_LAUNCH(K_chain_img_fwd, st, (launch_pdl((lean_warp_fwd_kernel<DIM, FIELD, MM, VS, VD>), grid, 256, 0, st, a)))
    if (s.src_vm && s.dst_vm) ADVK_LF(2, true, true);
    else if (s.src_vm) ADVK_LF(2, true, false);
    else if (s.dst_vm && mm == 2) ADVK_LF(2, false, true);
    else if (s.dst_vm) ADVK_LF(1, false, true);
    else if (mm == 0) ADVK_LF(0, false, false);
    else if (mm == 1) ADVK_LF(1, false, false);
    else ADVK_LF(2, false, false);
#undef ADVK_LF
  } else {
    if (k == 0) a.src = P.src_packed;            // the packed copy of the chain input
#define ADVK_LF(MM, DP) ADVK_LAUNCH(K_chain_pk_fwd, st, (launch_pdl((lean_warp_fwd_pk_kernel<DIM, FIELD, MM, DP>), grid, 256, 0, st, a)))
#define ADVK_LF2(DP) do { if (mm == 0) ADVK_LF(0, DP); else if (mm == 1) ADVK_LF(1, DP); else ADVK_LF(2, DP); } while (0)
    if (s.dst_pk) ADVK_LF2(true);
    else ADVK_LF2(false);
#undef ADVK_LF2
#undef ADVK_LF
  }
}

template <int DIM>
static int lean_fwd(Program& P, cudaStream_t st) {
  const dim3 grid(blocks_for(P.g.S, 256), P.g.N);
  if (P.pack) {
    if (!P.src_packed) { set_error("chain_apply_fwd: stash is NULL"); return ADVK_ERR_ARG; }
    const dim3 gp(blocks_for(P.g.S, 256), (unsigned)(P.g.N * (P.C >> 2)));
    ADVK_LAUNCH(K_chain_pk_fwd, st, (launch_pdl((lean_pack_kernel), gp, 256, 0, st, P.user_src, reinterpret_cast<float4*>(P.src_packed), (int)P.g.S)));
  }
  for (int k = 0; k < P.n; ++k) {
    const Stage& s = P.st[k];
    if (s.kind == ADVK_STAGE_INTENSITY) {
      LeanInt a;
      lean_fill_int(P, s, k == P.n - 1, a);
      if (lean_int_rows(P, s)) ADVK_LAUNCH(K_chain_img_fwd, st, (launch_pdl((lean_intensity_rows_kernel<DIM, false>), lean_int_grid<DIM>(P), 256, 0, st, a)));
      else ADVK_LAUNCH(K_chain_img_fwd, st, (launch_pdl((lean_intensity_kernel<DIM, false>), grid, 256, 0, st, a)));
    } else if (s.kind == ADVK_STAGE_WARP_FIELD) lean_launch_warp_fwd<DIM, true>(P, s, k, st);
    else lean_launch_warp_fwd<DIM, false>(P, s, k, st);
  }
  return check_launch("chain_apply_fwd (lean)");
}

static int lean_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
  }
  return sms;
}

// Resident CTAs per SM the lean adjoint kernels are compiled for (their register budget).  Measured at 128^3
// (gpurun_out/r02g, ncu, us per launch; 0 = no target, the compiler's own choice of 77-102 registers):
//                              0      2      3      4
//   field  warp, C = 1        52.4   60.2   44.3   38.3     -> 4 (64 registers, 8-24 bytes of spill)
//   affine warp, C = 1        89.9    -     84.2   99.8     -> 3
//   field  warp, packed K=4   63.1   75.4   64.0   66.1     -> 0
//   affine warp, packed K=4  121.0    -    125.2  136.2     -> 0
// ADVK_LEAN_MINB overrides all four (experiments).
static int g_lean_minb = -2;
static int lean_minb(bool field, bool pack) {
  if (g_lean_minb == -2) {
    const char* e = getenv("ADVK_LEAN_MINB");
    g_lean_minb = e ? atoi(e) : -1;
  }
  if (g_lean_minb >= 0) return g_lean_minb;
  return pack ? 0 : (field ? 4 : 3);
}

// The theta gradient of an affine stage.  Split: the adjoint kernel writes dL/d(coordinate) per voxel into the tail
// of `scratch` and lean_theta_reduce_kernel sums it; in-kernel: the adjoint kernel carries the d(d+1) accumulators
// itself.  Measured at 128^3 (live CUDA events per iteration, gpurun_out/r02r): channel-packed K = 4 chain adjoint
// 214.9 -> 177.3 us with the split, C = 1 image chain adjoint 128.9 -> 135.4 us (its affine stage is bound by the
// sectors a rotated line of scalar REDs touches, 6.5 M against 1.9 M for the field stage, not by registers).
// ADVK_LEAN_THETA: 0 = never split, 1 = packed chains only (default), 2 = always.
// (Measured and removed, gpurun_out/r02F / r02G: a SHEARED traversal of the affine adjoints -- the volume walked along
// the direction the affine map sends to the source x axis, so that a warp's gathers and REDs fall on an axis-aligned
// source line -- parity green, but 4-8 us slower per kernel, with exact or approximate divisions alike: the upstream
// reads and coordinate-gradient writes along slanted output rows cost more than the aligned source rows return.)
static int g_lean_theta = -1;
static bool lean_theta_split(bool pack) {
  if (g_lean_theta < 0) {
    const char* e = getenv("ADVK_LEAN_THETA");
    g_lean_theta = e ? atoi(e) : 1;
  }
  return g_lean_theta >= 2 || (g_lean_theta == 1 && pack);
}
static int g_lean_minb_aff = -2;
static int lean_minb_affine(bool pack) {      // register budget of the split affine adjoints (ADVK_LEAN_MINB_AFFINE)
  if (g_lean_minb_aff == -2) {
    const char* e = getenv("ADVK_LEAN_MINB_AFFINE");
    g_lean_minb_aff = e ? atoi(e) : -1;
  }
  if (g_lean_minb_aff >= 0) return g_lean_minb_aff;
  return pack ? 3 : 4;        // 80 / 62 registers, no spill (unconstrained: 124 / 73)
}

template <int DIM, bool FIELD>
static void lean_launch_warp_bwd(const Program& P, const Stage& s, int k, cudaStream_t st) {
  const bool last = (k == P.n - 1);
  LeanBwd a;
  a.g = P.g; a.C = P.C; a.tps = P.tps; a.ftps = P.ftps; a.n_tiles = (unsigned)P.n_tiles;
  a.src = s.src; a.g_dst = s.g_dst; a.g_src = s.g_src; a.phi = s.phi; a.theta = s.theta;
  a.g_phi = s.g_phi; a.g_theta = s.g_theta;
  a.clamp = (last && P.do_clamp) ? 1 : 0; a.lo = P.lo; a.hi = P.hi;
  unsigned grid = (unsigned)P.n_tiles;
  const bool split = !FIELD && s.g_theta && P.g_coord && lean_theta_split(P.pack != 0);
  if (split) { a.g_phi = P.g_coord; a.g_theta = nullptr; }
  // in-kernel theta reduction (predecessor): a block walks several tiles so that the per-sample atomics stay few
  else if (!FIELD && s.g_theta) { const unsigned cap = (unsigned)lean_sms() * 16u; if (grid > cap) grid = cap; }
  const int minb = split ? lean_minb_affine(P.pack != 0) : lean_minb(FIELD, P.pack != 0);
  constexpr bool TA = !FIELD;                 // only affine stages have an in-kernel-accumulating variant
#define ADVK_LB1(KID, KERN, B1, ACC)                                                                     \
  do {                                                                                                   \
    if (minb >= 4) ADVK_LAUNCH(KID, st, (launch_pdl((KERN<DIM, FIELD, B1, 4, ACC>), grid, 256, 0, st, a)));        \
    else if (minb == 3) ADVK_LAUNCH(KID, st, (launch_pdl((KERN<DIM, FIELD, B1, 3, ACC>), grid, 256, 0, st, a)));   \
    else if (minb == 2) ADVK_LAUNCH(KID, st, (launch_pdl((KERN<DIM, FIELD, B1, 2, ACC>), grid, 256, 0, st, a)));   \
    else ADVK_LAUNCH(KID, st, (launch_pdl((KERN<DIM, FIELD, B1, 0, ACC>), grid, 256, 0, st, a)));                  \
  } while (0)
#define ADVK_LB(KID, KERN, B1)                                                                           \
  do {                                                                                                   \
    if (TA && !split && s.g_theta) ADVK_LB1(KID, KERN, B1, TA);                                          \
    else ADVK_LB1(KID, KERN, B1, false);                                                                 \
  } while (0)
  if (!P.pack) {
    if (s.src_vm) ADVK_LB(K_chain_img_bwd, lean_warp_bwd_kernel, true);
    else ADVK_LB(K_chain_img_bwd, lean_warp_bwd_kernel, false);
  } else {
    if (k == 0) a.src = P.src_packed;
    if (!last) ADVK_LB(K_chain_pk_bwd, lean_warp_bwd_pk_kernel, true);
    else ADVK_LB(K_chain_pk_bwd, lean_warp_bwd_pk_kernel, false);
  }
#undef ADVK_LB
#undef ADVK_LB1
  if (split) {
    // ~4 CTAs per SM over the whole batch, 4 loads in flight per thread (lean_theta_reduce_kernel); every CTA ends
    // in d(d+1) atomics on its sample's gradient
    unsigned gx = (unsigned)((lean_sms() * 4 + P.g.N - 1) / P.g.N);
    if (gx > (unsigned)P.tps) gx = (unsigned)P.tps;
    if (gx < 1) gx = 1;
    const KernelId kid = P.pack ? K_chain_pk_bwd : K_chain_img_bwd;
    ADVK_LAUNCH(kid, st, (launch_pdl((lean_theta_reduce_kernel<DIM>), dim3(gx, (unsigned)P.g.N), 256, 0, st, P.g, P.g_coord, s.g_theta)));
  }
}

template <int DIM>
static int lean_bwd(Program& P, cudaStream_t st) {
  if (P.pack && !P.src_packed) { set_error("chain_apply_bwd: stash is NULL"); return ADVK_ERR_ARG; }
  const i64 tot = (i64)P.g.N * P.C * P.g.S;
  for (int k = P.first_bwd; k < P.n; ++k) {
    const Stage& s = P.st[k];
    if (s.zero_g_src && s.g_src) cudaMemsetAsync(s.g_src, 0, sizeof(float) * tot, st);
  }
  const dim3 grid(blocks_for(P.g.S, 256), P.g.N);
  for (int k = P.n - 1; k >= P.first_bwd; --k) {
    const Stage& s = P.st[k];
    const bool last = (k == P.n - 1);
    if (s.kind == ADVK_STAGE_INTENSITY) {
      LeanInt a;
      lean_fill_int(P, s, last, a);
      if (lean_int_rows(P, s)) ADVK_LAUNCH(K_chain_img_bwd, st, (launch_pdl((lean_intensity_rows_kernel<DIM, true>), lean_int_grid<DIM>(P), 256, 0, st, a)));
      else ADVK_LAUNCH(K_chain_img_bwd, st, (launch_pdl((lean_intensity_kernel<DIM, true>), grid, 256, 0, st, a)));
    } else if (s.kind == ADVK_STAGE_WARP_FIELD) lean_launch_warp_bwd<DIM, true>(P, s, k, st);
    else lean_launch_warp_bwd<DIM, false>(P, s, k, st);
  }
  const Stage& f = P.st[P.first_bwd];
  if (P.pack && f.g_src_user && f.g_src) {
    const dim3 gu(blocks_for(P.g.S, 256), (unsigned)(P.g.N * (P.C >> 2)));
    ADVK_LAUNCH(K_chain_pk_bwd, st, (launch_pdl((lean_unpack_kernel), gu, 256, 0, st, reinterpret_cast<const float4*>(f.g_src), f.g_src_user, (int)P.g.S)));
  }
  return check_launch("chain_apply_bwd (lean)");
}

template <int DIM>
static int launch_fwd(Program& P, cudaStream_t st) {
  if (P.lean) return lean_fwd<DIM>(P, st);
  P.interleave = 1;
  return launch_fwd_t<DIM, GEN_MINB_FWD>(P, st);
}

template <int DIM>
static int launch_bwd(Program& P, cudaStream_t st) {
  if (P.lean) return lean_bwd<DIM>(P, st);
  P.interleave = 1;
  return launch_bwd_t<DIM, GEN_MINB_BWD>(P, st);
}

}  // namespace advk

using namespace advk;

extern "C" int advk_chain_set_lean(int enable) {
  int prev = lean_enabled() ? 1 : 0;
  g_lean = enable ? 1 : 0;
  return prev;
}

extern "C" int advk_chain_set_packed(int enable) {
  int prev = pack_enabled() ? 1 : 0;
  g_pack = enable ? 1 : 0;
  return prev;
}

extern "C" int advk_chain_set_cooperative(int enable) {
  int prev = coop_enabled() ? 1 : 0;
  if (enable) {
    g_coop = -1;
    (void)coop_enabled();      // re-probe the device attribute / environment
  } else {
    g_coop = 0;
  }
  return prev;
}

extern "C" int advk_chain_workspace_floats(const advk_chain_desc* d, size_t* stash_floats,
                                           size_t* scratch_floats) {
  Dims g;
  ADVK_REQUIRE(d && make_dims(&d->g, g) && d->C >= 1 && d->n_stages >= 1 && d->n_stages <= MAX_STAGES,
               "bad chain descriptor");
  const i64 ncs = slot_floats((i64)g.N * d->C * g.S), ns = slot_floats((i64)g.N * g.S);
  int warps = 0;
  for (int k = 0; k < d->n_stages; ++k)
    if (d->stages[k].kind != ADVK_STAGE_INTENSITY) ++warps;
  // + one slot for the packed copy of the chain input (channel-packed prediction path, lean kernels)
  const bool may_pack = warps == d->n_stages && d->n_stages >= 2 && (d->C % 4) == 0;
  if (stash_floats)
    *stash_floats = (size_t)((i64)(d->n_stages - 1) * ncs + ((d->want_mask && warps > 1) ? (i64)(warps - 1) * ns : 0) +
                             (may_pack ? ncs : 0));
  // one slot per stage boundary, plus one for the packed gradient of the chain input (packed mode)
  // ... and, behind them, 4 floats per voxel for the coordinate gradient of an affine stage (lean kernels)
  bool affine = false;
  for (int k = 0; k < d->n_stages; ++k)
    if (d->stages[k].kind == ADVK_STAGE_WARP_AFFINE) affine = true;
  if (scratch_floats) *scratch_floats = (size_t)((i64)d->n_stages * ncs + (affine ? 4 * ns : 0));
  return ADVK_OK;
}

extern "C" int advk_chain_apply_fwd(const advk_chain_desc* d, const float* src, const float* mask_src,
                                    float* stash, float* out, float* mask_out, void* stream) {
  Program P;
  ADVK_REQUIRE(src && out, "null pointer");
  ADVK_REQUIRE(build_program(d, P, src, mask_src, stash, out, mask_out), "bad chain descriptor");
  cudaStream_t st = (cudaStream_t)stream;
  return d->g.d == 2 ? launch_fwd<2>(P, st) : launch_fwd<3>(P, st);
}

extern "C" int advk_chain_apply_bwd(const advk_chain_desc* d, const float* g_out, const float* src,
                                    const float* stash, float* scratch, float* g_src, void* stream) {
  Program P;
  ADVK_REQUIRE(src && g_out, "null pointer");
  // data-path pointers are rebuilt exactly as in the forward (same stash layout; the last stage's dst and
  // the mask outputs are unused)
  ADVK_REQUIRE(build_program(d, P, src, nullptr, const_cast<float*>(stash), nullptr, nullptr, true),
               "bad chain descriptor");
  P.want_mask = 0;
  const i64 slot = slot_floats((i64)P.g.N * P.C * P.g.S);
  // gradient plumbing: g_dst of stage k = g_src of stage k+1 (scratch slot k), last = g_out
  int first = P.n;   // earliest stage that produces a requested gradient
  for (int k = 0; k < P.n; ++k) {
    const advk_chain_stage& a = d->stages[k];
    Stage& s = P.st[k];
    s.g_delta = a.g_delta; s.g_up = a.g_up; s.g_phi = a.g_field; s.g_theta = a.g_theta;
    bool wants = a.g_delta || a.g_up || a.g_field || a.g_theta;
    if (wants && k < first) first = k;
  }
  if (g_src) first = 0;
  if (first == P.n) return ADVK_OK;       // nothing requested
  P.first_bwd = first;
  if (first < P.n - 1 || (P.pack && g_src)) ADVK_REQUIRE(scratch != nullptr, "scratch is NULL");
  P.g_coord = scratch ? scratch + (i64)P.n * slot : nullptr;   // behind the n stage slots (advk_chain_workspace_floats)
  for (int k = P.n - 1; k >= first; --k) {
    Stage& s = P.st[k];
    s.g_dst = (k == P.n - 1) ? g_out : scratch + (i64)k * slot;
    if (k == 0) {                                 // g_src may be NULL: gradient w.r.t. the chain input not wanted
      if (P.pack && g_src) { s.g_src = scratch + (i64)(P.n - 1) * slot; s.g_src_user = g_src; }
      else s.g_src = g_src;
    }
    else if (k == first) s.g_src = nullptr;       // nobody consumes it
    else s.g_src = scratch + (i64)(k - 1) * slot;
    s.zero_g_src = (s.kind != ADVK_STAGE_INTENSITY && s.g_src != nullptr) ? 1 : 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return d->g.d == 2 ? launch_bwd<2>(P, st) : launch_bwd<3>(P, st);
}
