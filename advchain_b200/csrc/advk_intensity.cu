// advk_intensity.cu -- the two intensity stages of the chain, AdvNoise (x + eps*delta,
// adv_noise.py:79-90) and AdvBias (x * clip(exp(upsample(bspline(cp)))), adv_bias.py:152-188,
// 279-356), single or fused in either order, forward and backward.
//
// The reference evaluates the B-spline with a conv_transposeNd whose kernel is up to 643x643
// (adv_bias.py:293-301).  Here conv_transpose+crop is folded, per axis, into a small dense matrix
// A_ax (built once on the host from the reference's own kernel construction), so the low-res field
// is  low = (A_D (x) A_H (x) A_W) cp  -- a few hundred FMAs per low-res voxel -- and the full-res
// bias is evaluated on the fly per voxel (linear upsample + exp + clip) without ever being stored.
#include "advk_intensity.cuh"
#include "advk_adjoint.cuh"

namespace advk {

// low[n,z,y,x] = sum_{i,j,k} AD[z,i] AH[y,j] AW[x,k] * (s * cp[n,i,j,k])
template <int DIM>
__global__ void lowfield_fwd_kernel(BiasCfg b, int N, const float* __restrict__ cp, float s,
                                    float* __restrict__ low) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  i64 L = (i64)b.lD * b.lH * b.lW;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)N * L) return;
  int x = (int)(idx % b.lW), y = (int)((idx / b.lW) % b.lH), z = (int)((idx / ((i64)b.lW * b.lH)) % b.lD);
  i64 n = idx / L;
  const float* c = cp + n * ((i64)b.nD * b.nH * b.nW);
  float acc = 0.f;
  for (int i = 0; i < b.nD; ++i) {
    float wz = (DIM == 3) ? b.AD[z * b.nD + i] : 1.f;
    if (wz == 0.f) continue;
    for (int j = 0; j < b.nH; ++j) {
      float wy = wz * b.AH[y * b.nH + j];
      if (wy == 0.f) continue;
      for (int k = 0; k < b.nW; ++k) acc += (wy * b.AW[x * b.nW + k]) * (s * c[((i64)i * b.nH + j) * b.nW + k]);
    }
  }
  low[idx] = acc;
}

// g_cp[n,i,j,k] = s * sum_{z,y,x} AD[z,i] AH[y,j] AW[x,k] g_low[n,z,y,x]   (one block per cp)
template <int DIM>
__global__ void __launch_bounds__(256)
lowfield_bwd_kernel(BiasCfg b, int N, const float* __restrict__ g_low, float s, float* __restrict__ g_cp) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[32];
  i64 ncp = (i64)b.nD * b.nH * b.nW;
  i64 L = (i64)b.lD * b.lH * b.lW;
  i64 e = blockIdx.x;   // n*ncp + cp index
  i64 n = e / ncp;
  int k = (int)(e % b.nW), j = (int)((e / b.nW) % b.nH), i = (int)((e / ((i64)b.nW * b.nH)) % b.nD);
  const float* gl = g_low + n * L;
  float v[1] = {0.f};
  // blockIdx.y splits the low-res volume; partial sums are combined with one atomic per block
  const int Li = (int)L, lW = b.lW, lHW = b.lW * b.lH;
  for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < Li; q += gridDim.y * blockDim.x) {
    int x = q % lW, y = (q / lW) % b.lH, z = q / lHW;
    float w = b.AW[x * b.nW + k] * b.AH[y * b.nH + j];
    if (DIM == 3) w *= b.AD[z * b.nD + i];
    if (w != 0.f) v[0] += w * gl[q];
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(g_cp + e, s * v[0]);
}

// order: 0 noise, 1 bias, 2 noise->bias, 3 bias->noise
template <int DIM>
__global__ void __launch_bounds__(256)
intensity_fwd_kernel(Dims g, int C, int order, const float* __restrict__ x, const float* __restrict__ delta,
                     float ns, const float* __restrict__ low, BiasCfg b, int use_ig, float ig,
                     float* __restrict__ out, float* __restrict__ bias_out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  float bv = 1.f;
  if (order != 0) {
    int xx, yy, zz;
    voxel_xyz(g, (unsigned)p, xx, yy, zz);
    float braw; bool pass;
    bv = bias_value(b, bias_up<DIM>(b, low + (i64)n * b.lD * b.lH * b.lW, zz, yy, xx, g, p), braw, pass);
    if (bias_out) bias_out[(i64)n * g.S + p] = bv;
  }
  for (int c = 0; c < C; ++c) {
    i64 q = ((i64)n * C + c) * g.S + p;
    float t = intensity_point(order, x[q], order != 1 ? delta[q] : 0.f, ns, bv, use_ig, ig);
    out[q] = t;
  }
}

template <int DIM>
__global__ void __launch_bounds__(256)
intensity_bwd_kernel(Dims g, int C, int order, const float* __restrict__ g_out, const float* __restrict__ x,
                     const float* __restrict__ delta, float ns, const float* __restrict__ low, BiasCfg b,
                     int use_ig, float ig, float* __restrict__ g_x, float* __restrict__ g_delta,
                     float* __restrict__ g_up) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  float bv = 1.f, braw = 1.f; bool pass = true;
  if (order != 0) {
    int xx, yy, zz;
    voxel_xyz(g, (unsigned)p, xx, yy, zz);
    bv = bias_value(b, bias_up<DIM>(b, low + (i64)n * b.lD * b.lH * b.lW, zz, yy, xx, g, p), braw, pass);
  }
  float gb = 0.f;
  for (int c = 0; c < C; ++c) {
    i64 q = ((i64)n * C + c) * g.S + p;
    float gd;
    float go = intensity_point_bwd(order, g_out[q], x[q], (order >= 2) ? delta[q] : 0.f, ns, bv, use_ig, ig, gd, gb);
    if (g_delta) g_delta[q] = gd;
    if (g_x) g_x[q] = go;
  }
  if (g_up) g_up[(i64)n * g.S + p] = pass ? gb * (b.use_log ? braw : 1.f) : 0.f;
}

// Adjoint of the linear upsample along one axis (scalar version of the morph one).
__global__ void __launch_bounds__(256)
adjoint_axis_f_kernel(const float* __restrict__ a, float* __restrict__ out, i64 outer, int n_in, int n_out,
                      i64 inner, float scale) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= outer * n_out * inner) return;
  i64 i = idx % inner;
  int j = (int)((idx / inner) % n_out);
  i64 o = idx / (inner * n_out);
  int plo = (int)floorf(((float)j - 0.5f) / scale - 0.5f) - 1;
  int phi = (int)ceilf(((float)j + 1.5f) / scale - 0.5f) + 1;
  if (plo < 0 || j == 0) plo = 0;
  if (phi > n_in - 1) phi = n_in - 1;
  float acc = 0.f;
  for (int p = plo; p <= phi; ++p) {
    UpAxis u = up_axis(p, n_out, scale);
    float w = (u.i0 == j ? u.l0 : 0.f) + (u.i1 == j ? u.l1 : 0.f);
    if (w != 0.f) acc += w * a[(o * n_in + p) * inner + i];
  }
  out[idx] = acc;
}

}  // namespace advk

using namespace advk;

extern "C" int advk_bias_lowfield_fwd(const advk_bias_cfg* cfg, int N, const float* cp, float cp_scale,
                                      float* low, void* stream) {
  BiasCfg b;
  int d = (cfg && cfg->low[0] == 1 && cfg->n_cp[0] == 1 && !cfg->A[0]) ? 2 : 3;
  ADVK_REQUIRE(make_bias(cfg, d, b), "bad bias config");
  ADVK_REQUIRE(cp && low && N >= 1, "null pointer");
  i64 tot = (i64)N * b.lD * b.lH * b.lW;
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 2) ADVK_LAUNCH(K_lowfield_fwd, st, launch_pdl((lowfield_fwd_kernel<2>), blocks_for(tot, 128), 128, 0, st, b, N, cp, cp_scale, low));
  else ADVK_LAUNCH(K_lowfield_fwd, st, launch_pdl((lowfield_fwd_kernel<3>), blocks_for(tot, 128), 128, 0, st, b, N, cp, cp_scale, low));
  return check_launch("bias_lowfield_fwd");
}

extern "C" int advk_bias_lowfield_bwd(const advk_bias_cfg* cfg, int N, const float* g_low, float cp_scale,
                                      float* g_cp, void* stream) {
  BiasCfg b;
  int d = (cfg && cfg->low[0] == 1 && cfg->n_cp[0] == 1 && !cfg->A[0]) ? 2 : 3;
  ADVK_REQUIRE(make_bias(cfg, d, b), "bad bias config");
  ADVK_REQUIRE(g_low && g_cp && N >= 1, "null pointer");
  unsigned blocks = (unsigned)((i64)N * b.nD * b.nH * b.nW);
  i64 L = (i64)b.lD * b.lH * b.lW;
  ADVK_REQUIRE(L < 2147483647LL, "low-res bias field too large");
  unsigned split = (unsigned)((L + 1023) / 1024);
  if (split > 32) split = 32;
  if (split < 1) split = 1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(g_cp, 0, sizeof(float) * blocks, st);
  dim3 grid(blocks, split);
  if (d == 2) ADVK_LAUNCH(K_lowfield_bwd, st, launch_pdl((lowfield_bwd_kernel<2>), grid, 256, 0, st, b, N, g_low, cp_scale, g_cp));
  else ADVK_LAUNCH(K_lowfield_bwd, st, launch_pdl((lowfield_bwd_kernel<3>), grid, 256, 0, st, b, N, g_low, cp_scale, g_cp));
  return check_launch("bias_lowfield_bwd");
}

extern "C" int advk_intensity_fwd(const advk_geom* gg, int C, int order, const float* x, const float* delta,
                                  float noise_scale, const float* low, const advk_bias_cfg* bias,
                                  int use_ignore, float ignore_value, float* out, float* bias_out,
                                  void* stream) {
  Dims g; BiasCfg b = {};
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(order >= 0 && order <= 3 && C >= 1 && x && out, "bad order / null pointer");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  if (order != 1) ADVK_REQUIRE(delta != nullptr, "delta is NULL");
  if (order != 0) {
    ADVK_REQUIRE(make_bias(bias, gg->d, b) && low, "bad bias config / low is NULL");
    if (!b.upsample) ADVK_REQUIRE(b.lD == g.D && b.lH == g.H && b.lW == g.W, "low field size mismatch");
  }
  dim3 grid(blocks_for(g.S, 256), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (gg->d == 2)
    ADVK_LAUNCH(K_intensity_fwd, st, launch_pdl((intensity_fwd_kernel<2>), grid, 256, 0, st, g, C, order, x, delta, noise_scale, low, b, use_ignore, ignore_value, out, bias_out));
  else
    ADVK_LAUNCH(K_intensity_fwd, st, launch_pdl((intensity_fwd_kernel<3>), grid, 256, 0, st, g, C, order, x, delta, noise_scale, low, b, use_ignore, ignore_value, out, bias_out));
  return check_launch("intensity_fwd");
}

extern "C" int advk_intensity_bwd(const advk_geom* gg, int C, int order, const float* g_out, const float* x,
                                  const float* delta, float noise_scale, const float* low,
                                  const advk_bias_cfg* bias, int use_ignore, float ignore_value, float* g_x,
                                  float* g_delta, float* g_up, void* stream) {
  Dims g; BiasCfg b = {};
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(order >= 0 && order <= 3 && C >= 1 && x && g_out, "bad order / null pointer");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  if (order == 2 || order == 3) ADVK_REQUIRE(delta != nullptr, "delta is NULL");
  if (order != 0) {
    ADVK_REQUIRE(make_bias(bias, gg->d, b) && low, "bad bias config / low is NULL");
    if (!b.upsample) ADVK_REQUIRE(b.lD == g.D && b.lH == g.H && b.lW == g.W, "low field size mismatch");
  }
  dim3 grid(blocks_for(g.S, 256), g.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (gg->d == 2)
    ADVK_LAUNCH(K_intensity_bwd, st, launch_pdl((intensity_bwd_kernel<2>), grid, 256, 0, st, g, C, order, g_out, x, delta, noise_scale, low, b, use_ignore, ignore_value, g_x, g_delta, g_up));
  else
    ADVK_LAUNCH(K_intensity_bwd, st, launch_pdl((intensity_bwd_kernel<3>), grid, 256, 0, st, g, C, order, g_out, x, delta, noise_scale, low, b, use_ignore, ignore_value, g_x, g_delta, g_up));
  return check_launch("intensity_bwd");
}

extern "C" size_t advk_bias_scratch_floats(const advk_geom* gg, const advk_bias_cfg* bias) {
  Dims g; BiasCfg b;
  if (!make_dims(gg, g) || !make_bias(bias, gg->d, b)) return 0;
  i64 p1 = (gg->d == 3) ? (i64)g.N * b.lD * g.H * g.W : 0;
  i64 p2 = (i64)g.N * b.lD * b.lH * g.W;
  return (size_t)(p1 + p2);
}

extern "C" int advk_bias_upsample_adjoint(const advk_geom* gg, const advk_bias_cfg* bias, const float* g_up,
                                          float* scratch, float* g_low, void* stream) {
  Dims g; BiasCfg b;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(make_bias(bias, gg->d, b), "bad bias config");
  ADVK_REQUIRE(g_up && g_low, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!b.upsample) {
    cudaMemcpyAsync(g_low, g_up, sizeof(float) * g.N * g.S, cudaMemcpyDeviceToDevice, st);
    return check_launch("bias_upsample_adjoint(copy)");
  }
  ADVK_REQUIRE(scratch != nullptr, "scratch is NULL");
  const float* a = g_up;
  float* s1 = scratch;
  if (gg->d == 3) {
    i64 tot = (i64)g.N * b.lD * g.H * g.W;
    ADVK_LAUNCH(K_adjoint_axis_f, st, launch_pdl((adjoint_axis_f_kernel), blocks_for(tot, 256), 256, 0, st, a, s1, g.N, g.D, b.lD, (i64)g.H * g.W, b.sD));
    a = s1; s1 += tot;
  }
  // the last two axes in one launch (advk_adjoint.cuh); per-axis kernels when a row does not fit shared memory
  if (!launch_adjoint_hw<float>(K_adjoint_axis_f, a, nullptr, 1.f, g_low, (i64)g.N * b.lD, g.H, g.W, b.lH, b.lW, b.sH, b.sW, st)) {
    i64 tot2 = (i64)g.N * b.lD * b.lH * g.W;
    ADVK_LAUNCH(K_adjoint_axis_f, st, launch_pdl((adjoint_axis_f_kernel), blocks_for(tot2, 256), 256, 0, st, a, s1, (i64)g.N * b.lD, g.H, b.lH, g.W, b.sH));
    i64 tot3 = (i64)g.N * b.lD * b.lH * b.lW;
    ADVK_LAUNCH(K_adjoint_axis_f, st, launch_pdl((adjoint_axis_f_kernel), blocks_for(tot3, 256), 256, 0, st, s1, g_low, (i64)g.N * b.lD * b.lH, g.W, b.lW, 1, b.sW));
  }
  return check_launch("bias_upsample_adjoint");
}
