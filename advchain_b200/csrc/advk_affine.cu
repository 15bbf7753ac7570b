// advk_affine.cu -- AdvAffine parameter algebra: Hardtanh -> 2x3 / 3x4 matrix, closed-form
// inverse, and the analytic backward of both (the reference gets these from autograd through
// torch.stack/cos/sin/matmul/.inverse(): adv_affine.py:210-273, 316-324).
// One thread per sample: the work is O(N * 50 flops); it exists so that a PGD step issues two
// tiny launches instead of ~60 ATen ops and a batched LU.
#include "advk_common.cuh"

namespace advk {

constexpr float kPi = 3.14159265358979323846f;

struct AffineCfg {
  float rot[3], scale[3], shift[3];
};

__device__ __forceinline__ float hardtanh(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// d x d inverse (adjugate) + translation: [A | t]^-1 = [A^-1 | -A^-1 t]
template <int DIM>
__device__ __forceinline__ void invert_affine(const float* th, float* inv) {
  constexpr int R = DIM + 1;
  if (DIM == 2) {
    float a = th[0], b = th[1], c = th[R], d = th[R + 1];
    float det = a * d - b * c;
    float id = 1.f / det;
    float i00 = d * id, i01 = -b * id, i10 = -c * id, i11 = a * id;
    inv[0] = i00; inv[1] = i01; inv[2] = -(i00 * th[2] + i01 * th[R + 2]);
    inv[R] = i10; inv[R + 1] = i11; inv[R + 2] = -(i10 * th[2] + i11 * th[R + 2]);
  } else {
    float a00 = th[0], a01 = th[1], a02 = th[2];
    float a10 = th[4], a11 = th[5], a12 = th[6];
    float a20 = th[8], a21 = th[9], a22 = th[10];
    float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    float det = a00 * c00 + a01 * c01 + a02 * c02;
    float id = 1.f / det;
    float m[9];
    m[0] = c00 * id; m[1] = (a02 * a21 - a01 * a22) * id; m[2] = (a01 * a12 - a02 * a11) * id;
    m[3] = c01 * id; m[4] = (a00 * a22 - a02 * a20) * id; m[5] = (a02 * a10 - a00 * a12) * id;
    m[6] = c02 * id; m[7] = (a01 * a20 - a00 * a21) * id; m[8] = (a00 * a11 - a01 * a10) * id;
    float t0 = th[3], t1 = th[7], t2 = th[11];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      inv[i * 4 + 0] = m[i * 3 + 0]; inv[i * 4 + 1] = m[i * 3 + 1]; inv[i * 4 + 2] = m[i * 3 + 2];
      inv[i * 4 + 3] = -(m[i * 3 + 0] * t0 + m[i * 3 + 1] * t1 + m[i * 3 + 2] * t2);
    }
  }
}

template <int DIM>
__device__ __forceinline__ void build_theta(const AffineCfg& c, const float* p, float* th) {
  if (DIM == 2) {
    float ang = p[0] * c.rot[0] * kPi;
    float cs = cosf(ang), sn = sinf(ang);
    float a = 1.f + p[1] * c.scale[0], b = 1.f + p[2] * c.scale[1];
    th[0] = a * cs; th[1] = b * (-sn); th[2] = p[3] * c.shift[0];
    th[3] = a * sn; th[4] = b * cs;   th[5] = p[4] * c.shift[1];
  } else {
    float ph = p[0] * c.rot[0] * kPi, tt = p[1] * c.rot[1] * kPi, ps = p[2] * c.rot[2] * kPi;
    float cph = cosf(ph), sph = sinf(ph), cth = cosf(tt), sth = sinf(tt), cps = cosf(ps), sps = sinf(ps);
    float sx = 1.f + p[3] * c.scale[0], sy = 1.f + p[4] * c.scale[1], sz = 1.f + p[5] * c.scale[2];
    // theta = (T (R S))[:3,:4] = [R diag(s) | t]
    th[0] = (cth * cps) * sx; th[1] = (-cph * sps + sph * sth * cps) * sy; th[2] = (sph * sps + cph * sth * cps) * sz;
    th[4] = (cth * sps) * sx; th[5] = (cph * cps + sph * sth * sps) * sy;  th[6] = (-sph * cps + cph * sth * sps) * sz;
    th[8] = (-sth) * sx;      th[9] = (sph * cth) * sy;                    th[10] = (cph * cth) * sz;
    th[3] = p[6] * c.shift[0]; th[7] = p[7] * c.shift[1]; th[11] = p[8] * c.shift[2];
  }
}

template <int DIM>
__global__ void affine_theta_fwd_kernel(AffineCfg c, const float* __restrict__ param, float pscale,
                                        int N, float* __restrict__ theta, float* __restrict__ theta_inv) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  constexpr int NP = DIM == 2 ? 5 : 9;
  constexpr int NT = DIM * (DIM + 1);
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float p[NP], th[NT], inv[NT];
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = hardtanh(pscale * param[n * NP + i]);
  build_theta<DIM>(c, p, th);
#pragma unroll
  for (int i = 0; i < NT; ++i) theta[n * NT + i] = th[i];
  if (theta_inv) {
    invert_affine<DIM>(th, inv);
#pragma unroll
    for (int i = 0; i < NT; ++i) theta_inv[n * NT + i] = inv[i];
  }
}

template <int DIM>
__global__ void affine_theta_bwd_kernel(AffineCfg c, const float* __restrict__ param, float pscale,
                                        int N, const float* __restrict__ g_theta,
                                        const float* __restrict__ g_theta_inv,
                                        float* __restrict__ g_param) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  constexpr int NP = DIM == 2 ? 5 : 9;
  constexpr int NT = DIM * (DIM + 1);
  constexpr int R = DIM + 1;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float raw[NP], p[NP], th[NT], g[NT];
#pragma unroll
  for (int i = 0; i < NP; ++i) { raw[i] = pscale * param[n * NP + i]; p[i] = hardtanh(raw[i]); }
  build_theta<DIM>(c, p, th);
#pragma unroll
  for (int i = 0; i < NT; ++i) g[i] = g_theta ? g_theta[n * NT + i] : 0.f;
  if (g_theta_inv) {
    // g_M = -Minv^T G Minv^T on the homogeneous matrices; keep the first d rows.
    float inv[NT];
    invert_affine<DIM>(th, inv);
    float Mi[R][R], G[R][R], T1[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        Mi[i][j] = (i < DIM) ? inv[i * R + j] : (j == DIM ? 1.f : 0.f);
        G[i][j] = (i < DIM) ? g_theta_inv[n * NT + i * R + j] : 0.f;
      }
    // T1 = Minv^T G
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) s += Mi[k][i] * G[k][j];
        T1[i][j] = s;
      }
    // g_M = -T1 Minv^T
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k) s += T1[i][k] * Mi[j][k];
        g[i * R + j] -= s;
      }
  }
  float gp[NP];
  if (DIM == 2) {
    float ang = p[0] * c.rot[0] * kPi;
    float cs = cosf(ang), sn = sinf(ang);
    float a = 1.f + p[1] * c.scale[0], b = 1.f + p[2] * c.scale[1];
    gp[0] = c.rot[0] * kPi * (-a * sn * g[0] - b * cs * g[1] + a * cs * g[3] - b * sn * g[4]);
    gp[1] = c.scale[0] * (cs * g[0] + sn * g[3]);
    gp[2] = c.scale[1] * (-sn * g[1] + cs * g[4]);
    gp[3] = c.shift[0] * g[2];
    gp[4] = c.shift[1] * g[5];
  } else {
    float ph = p[0] * c.rot[0] * kPi, tt = p[1] * c.rot[1] * kPi, ps = p[2] * c.rot[2] * kPi;
    float cph = cosf(ph), sph = sinf(ph), cth = cosf(tt), sth = sinf(tt), cps = cosf(ps), sps = sinf(ps);
    float s[3] = {1.f + p[3] * c.scale[0], 1.f + p[4] * c.scale[1], 1.f + p[5] * c.scale[2]};
    float Rm[3][3] = {{cth * cps, -cph * sps + sph * sth * cps, sph * sps + cph * sth * cps},
                      {cth * sps, cph * cps + sph * sth * sps, -sph * cps + cph * sth * sps},
                      {-sth, sph * cth, cph * cth}};
    float dPh[3][3] = {{0.f, sph * sps + cph * sth * cps, cph * sps - sph * sth * cps},
                       {0.f, -sph * cps + cph * sth * sps, -cph * cps - sph * sth * sps},
                       {0.f, cph * cth, -sph * cth}};
    float dTh[3][3] = {{-sth * cps, sph * cth * cps, cph * cth * cps},
                       {-sth * sps, sph * cth * sps, cph * cth * sps},
                       {-cth, -sph * sth, -cph * sth}};
    float dPs[3][3] = {{-cth * sps, -cph * cps - sph * sth * sps, sph * cps - cph * sth * sps},
                       {cth * cps, -cph * sps + sph * sth * cps, sph * sps + cph * sth * cps},
                       {0.f, 0.f, 0.f}};
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, gs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float gij = g[i * 4 + j];
        float gR = gij * s[j];
        a0 += gR * dPh[i][j]; a1 += gR * dTh[i][j]; a2 += gR * dPs[i][j];
        gs[j] += gij * Rm[i][j];
      }
    gp[0] = c.rot[0] * kPi * a0; gp[1] = c.rot[1] * kPi * a1; gp[2] = c.rot[2] * kPi * a2;
    gp[3] = c.scale[0] * gs[0]; gp[4] = c.scale[1] * gs[1]; gp[5] = c.scale[2] * gs[2];
    gp[6] = c.shift[0] * g[3]; gp[7] = c.shift[1] * g[7]; gp[8] = c.shift[2] * g[11];
  }
  // Hardtanh passes gradient on the OPEN interval only (quirk Q7); chain rule through pscale.
#pragma unroll
  for (int i = 0; i < NP; ++i)
    g_param[n * NP + i] = (raw[i] > -1.f && raw[i] < 1.f) ? gp[i] * pscale : 0.f;
}

static bool copy_cfg(const advk_affine_cfg* in, AffineCfg& o) {
  if (!in || (in->d != 2 && in->d != 3)) return false;
  for (int i = 0; i < 3; ++i) { o.rot[i] = in->rot[i]; o.scale[i] = in->scale[i]; o.shift[i] = in->shift[i]; }
  return true;
}

}  // namespace advk

using namespace advk;

extern "C" int advk_affine_theta_fwd(const advk_affine_cfg* cfg, const float* param, float pscale,
                                     int N, float* theta, float* theta_inv, void* stream) {
  AffineCfg c;
  ADVK_REQUIRE(copy_cfg(cfg, c), "bad affine config");
  ADVK_REQUIRE(param && theta && N >= 1, "null pointer / bad N");
  cudaStream_t st = (cudaStream_t)stream;
  int thr = 64, blk = (N + thr - 1) / thr;
  if (cfg->d == 2) ADVK_LAUNCH(K_affine_theta_fwd, st, launch_pdl((affine_theta_fwd_kernel<2>), blk, thr, 0, st, c, param, pscale, N, theta, theta_inv));
  else ADVK_LAUNCH(K_affine_theta_fwd, st, launch_pdl((affine_theta_fwd_kernel<3>), blk, thr, 0, st, c, param, pscale, N, theta, theta_inv));
  return check_launch("affine_theta_fwd");
}

extern "C" int advk_affine_theta_bwd(const advk_affine_cfg* cfg, const float* param, float pscale,
                                     int N, const float* g_theta, const float* g_theta_inv,
                                     float* g_param, void* stream) {
  AffineCfg c;
  ADVK_REQUIRE(copy_cfg(cfg, c), "bad affine config");
  ADVK_REQUIRE(param && g_param && N >= 1 && (g_theta || g_theta_inv), "null pointer / bad N");
  cudaStream_t st = (cudaStream_t)stream;
  int thr = 64, blk = (N + thr - 1) / thr;
  if (cfg->d == 2) ADVK_LAUNCH(K_affine_theta_bwd, st, launch_pdl((affine_theta_bwd_kernel<2>), blk, thr, 0, st, c, param, pscale, N, g_theta, g_theta_inv, g_param));
  else ADVK_LAUNCH(K_affine_theta_bwd, st, launch_pdl((affine_theta_bwd_kernel<3>), blk, thr, 0, st, c, param, pscale, N, g_theta, g_theta_inv, g_param));
  return check_launch("affine_theta_bwd");
}
