// advk_intensity.cuh -- device helpers shared by the single-stage intensity kernels
// (advk_intensity.cu) and the fused chain kernels (advk_chain.cu): bias-field evaluation
// (adv_bias.py:313-356) and the per-voxel noise/bias stage arithmetic with its adjoint
// (adv_noise.py:79-90, adv_bias.py:176-188).
#pragma once
#include "advk_common.cuh"

namespace advk {

struct BiasCfg {
  int nD, nH, nW;     // control points
  int lD, lH, lW;     // low-res field
  const float* AD; const float* AH; const float* AW;
  float sD, sH, sW;
  int upsample, use_log;
  float mag;
};

static inline bool make_bias(const advk_bias_cfg* b, int d, BiasCfg& o) {
  if (!b) return false;
  o.nD = b->n_cp[0]; o.nH = b->n_cp[1]; o.nW = b->n_cp[2];
  o.lD = b->low[0]; o.lH = b->low[1]; o.lW = b->low[2];
  o.AD = b->A[0]; o.AH = b->A[1]; o.AW = b->A[2];
  o.sD = b->up_scale[0]; o.sH = b->up_scale[1]; o.sW = b->up_scale[2];
  o.upsample = b->upsample; o.use_log = b->use_log; o.mag = b->magnitude;
  if (o.nD < 1 || o.nH < 1 || o.nW < 1 || o.lD < 1 || o.lH < 1 || o.lW < 1) return false;
  if (d == 2 && (o.nD != 1 || o.lD != 1)) return false;
  if (!o.AH || !o.AW || (d == 3 && !o.AD)) return false;
  return true;
}

// upsampled (pre-exp) field at a voxel
template <int DIM>
__device__ __forceinline__ float bias_up(const BiasCfg& b, const float* __restrict__ low_n, int z, int y,
                                         int x, const Dims& g, i64 p) {
  if (!b.upsample) return low_n[p];
  UpAxis ux = up_axis(x, b.lW, b.sW), uy = up_axis(y, b.lH, b.sH);
  float acc = 0.f;
#pragma unroll
  for (int dz = 0; dz < (DIM == 3 ? 2 : 1); ++dz) {
    int zi = 0; float lz = 1.f;
    if (DIM == 3) { UpAxis uz = up_axis(z, b.lD, b.sD); zi = dz ? uz.i1 : uz.i0; lz = dz ? uz.l1 : uz.l0; }
    const float* r0 = low_n + ((i64)zi * b.lH + uy.i0) * b.lW;
    const float* r1 = low_n + ((i64)zi * b.lH + uy.i1) * b.lW;
    acc += lz * (uy.l0 * (ux.l0 * __ldg(r0 + ux.i0) + ux.l1 * __ldg(r0 + ux.i1)) +
                 uy.l1 * (ux.l0 * __ldg(r1 + ux.i0) + ux.l1 * __ldg(r1 + ux.i1)));
  }
  return acc;
}

__device__ __forceinline__ float bias_value(const BiasCfg& b, float up, float& braw, bool& pass) {
  braw = b.use_log ? expf(up) : 1.f + up;
  float t = braw - 1.f;
  pass = (t >= -b.mag && t <= b.mag);
  return 1.f + clampf(t, -b.mag, b.mag);
}


// One voxel, one channel of the intensity stage(s).  order: 0 noise, 1 bias, 2 noise->bias,
// 3 bias->noise; dl = delta at the voxel (ignored for order 1), bv = clipped bias value.
__device__ __forceinline__ float intensity_point(int order, float x0, float dl, float ns, float bv,
                                                 int use_ig, float ig) {
  float t = x0;
  if (order == 0 || order == 2) {
    float t0 = t;
    t = t0 + ns * dl;
    if (use_ig && fabsf(t0 - ig) < 1e-8f) t = ig;
  }
  if (order != 0) {
    float t0 = t;
    t = t0 * bv;
    if (use_ig && fabsf(t0 - ig) < 1e-8f) t = ig;
  }
  if (order == 3) {
    float t0 = t;
    t = t0 + ns * dl;
    if (use_ig && fabsf(t0 - ig) < 1e-8f) t = ig;
  }
  return t;
}

// Adjoint of intensity_point: go = dL/d(out). Returns dL/dx0; gd = dL/d(delta); gb += dL/d(bias).
__device__ __forceinline__ float intensity_point_bwd(int order, float go, float x0, float dl, float ns,
                                                     float bv, int use_ig, float ig, float& gd, float& gb) {
  gd = 0.f;
  if (order == 0) {
    if (use_ig && fabsf(x0 - ig) < 1e-8f) go = 0.f;
    gd = ns * go;
  } else if (order == 1) {
    if (use_ig && fabsf(x0 - ig) < 1e-8f) go = 0.f;
    gb += go * x0;
    go *= bv;
  } else if (order == 2) {
    float t1 = x0 + ns * dl;
    bool ig1 = use_ig && fabsf(x0 - ig) < 1e-8f;
    if (ig1) t1 = ig;
    if (use_ig && fabsf(t1 - ig) < 1e-8f) go = 0.f;
    gb += go * t1;
    go *= bv;
    if (ig1) go = 0.f;
    gd = ns * go;
  } else {
    float t1 = x0 * bv;
    bool ig1 = use_ig && fabsf(x0 - ig) < 1e-8f;
    if (ig1) t1 = ig;
    if (use_ig && fabsf(t1 - ig) < 1e-8f) go = 0.f;
    gd = ns * go;
    if (ig1) go = 0.f;
    gb += go * x0;
    go *= bv;
  }
  return go;
}

}  // namespace advk
