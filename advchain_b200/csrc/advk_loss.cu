// advk_loss.cu -- the consistency loss the PGD loop evaluates every step
// (calc_segmentation_consistency, common/loss.py:8-87 with scales=[0], divergence types 'mse'
// and 'contour'; contour_loss, common/loss.py:102-220), forward and backward w.r.t. the
// prediction, as three streaming kernels instead of the reference's ~600-770 ATen calls
// (2 softmaxes per term, a Conv module rebuilt and run per class per Sobel direction, ...).
//
//   pred = softmax(output), tgt = softmax(reference) (or reference itself when is_gt)
//   E_c  = pred_c - tgt_c
//   mse      = sum_c sum_p (m E_c)^2 / (N K S) / (N S)                       (quirk Q9)
//   contour  = 1/(K-1) sum_{c>=1} 1/nk sum_k  sum_p (m (k * E_c))^2 / (N S)
//              k in {sobel_x, sobel_y} (2-D) | {gx, gx, gz} (3-D, quirk Q10: gy := gx)
// The Sobel correlation is linear, so conv(pred) - conv(tgt) = conv(E): one stencil per class.
//
//   K1 loss_softmax : E, pred  <- output, reference ; mse partial sums
//   K2 loss_contour : R_k = mult_k m^2 (k * E_c)    ; contour partial sums      (c >= 1)
//   K3 loss_grad    : g_pred_c = 2 A_mse m^2 E_c + 2 A_cont sum_k k^T * R_k ; softmax backward
#include "advk_common.cuh"

namespace advk {

// 1-D Sobel factors: h = smoothing [1,2,1], hp = derivative [1,0,-1]
__device__ __forceinline__ float sob_h(int o) { return o == 0 ? 2.f : 1.f; }      // o in {-1,0,1}
__device__ __forceinline__ float sob_hp(int o) { return (float)(-o); }           // [1,0,-1]

__global__ void __launch_bounds__(256)
loss_softmax_kernel(i64 S, int N, int K, const float* __restrict__ out, const float* __restrict__ ref,
                    const float* __restrict__ mask, int is_gt, float* __restrict__ E,
                    float* __restrict__ pred, double* __restrict__ acc) {
  __shared__ float red[32];
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  float v[1] = {0.f};
  if (p < S) {
    const float* o = out + (i64)n * K * S + p;
    const float* r = ref + (i64)n * K * S + p;
    float mo = -INFINITY, mr = -INFINITY;
    for (int c = 0; c < K; ++c) { mo = fmaxf(mo, o[c * S]); mr = fmaxf(mr, r[c * S]); }
    float so = 0.f, sr = 0.f;
    for (int c = 0; c < K; ++c) { so += expf(o[c * S] - mo); sr += expf(r[c * S] - mr); }
    const float m = mask ? mask[(i64)n * S + p] : 1.f;
    float s = 0.f;
    for (int c = 0; c < K; ++c) {
      float pc = expf(o[c * S] - mo) / so;
      float tc = is_gt ? r[c * S] : expf(r[c * S] - mr) / sr;
      float e = pc - tc;
      i64 q = ((i64)n * K + c) * S + p;
      E[q] = e;
      pred[q] = pc;
      float me = pc * m - tc * m;      // the reference masks both operands, then subtracts
      s += me * me;
    }
    v[0] = s;
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)v[0]);
}

// One thread per voxel per object class; R has 2 planes per class: [0] = sobel along the "x
// filter" direction (2-D kx | 3-D gx, weight 2), [1] = the other (2-D ky | 3-D gz).
template <int DIM>
__global__ void __launch_bounds__(256)
loss_contour_kernel(Dims g, int K, const float* __restrict__ E, const float* __restrict__ mask,
                    float* __restrict__ R, double* __restrict__ acc) {
  __shared__ float red[32];
  const int n = blockIdx.z / (K - 1);
  const int c = 1 + blockIdx.z % (K - 1);
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  float v[1] = {0.f};
  if (p < g.S) {
    const int x = (int)(p % g.W), y = (int)((p / g.W) % g.H), z = (int)(p / ((i64)g.W * g.H));
    const float* e = E + ((i64)n * K + c) * g.S;
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int a = (DIM == 3 ? -1 : 0); a <= (DIM == 3 ? 1 : 0); ++a) {
      int zz = z + a;
      if (zz < 0 || zz >= g.D) continue;
#pragma unroll
      for (int b = -1; b <= 1; ++b) {
        int yy = y + b;
        if (yy < 0 || yy >= g.H) continue;
#pragma unroll
        for (int cc = -1; cc <= 1; ++cc) {
          int xx = x + cc;
          if (xx < 0 || xx >= g.W) continue;
          float val = __ldg(e + ((i64)zz * g.H + yy) * g.W + xx);
          if (DIM == 2) {
            g0 += sob_h(b) * sob_hp(cc) * val;      // kx[b][cc] = h[b] hp[cc]
            g1 += sob_hp(b) * sob_h(cc) * val;      // ky[b][cc] = hp[b] h[cc]
          } else {
            g0 += sob_h(a) * sob_hp(b) * sob_h(cc) * val;   // gx = h(D) hp(H) h(W)
            g1 += sob_h(a) * sob_h(b) * sob_hp(cc) * val;   // gz = h(D) h(H) hp(W)
          }
        }
      }
    }
    const float m = mask ? mask[(i64)n * g.S + p] : 1.f;
    const float mult0 = (DIM == 3) ? 2.f : 1.f;             // quirk Q10: gx is used for x and y
    float a0 = m * g0, a1 = m * g1;
    v[0] = mult0 * a0 * a0 + a1 * a1;
    i64 rb = (((i64)n * (K - 1) + (c - 1)) * 2) * g.S + p;
    R[rb] = mult0 * m * a0;
    R[rb + g.S] = m * a1;
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v[0]);
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, float a_mse, float a_cont,
                                     float* __restrict__ loss) {
  loss[0] = (float)((double)a_mse * acc[0] + (double)a_cont * acc[1]);
}

// g_out_c = pred_c (g_pred_c - sum_j g_pred_j pred_j) * upstream
template <int DIM>
__global__ void __launch_bounds__(256)
loss_grad_kernel(Dims g, int K, const float* __restrict__ E, const float* __restrict__ pred,
                 const float* __restrict__ R, const float* __restrict__ mask, float a_mse, float a_cont,
                 const float* __restrict__ upstream, float* __restrict__ g_out) {
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.S) return;
  const int x = (int)(p % g.W), y = (int)((p / g.W) % g.H), z = (int)(p / ((i64)g.W * g.H));
  const float m = mask ? mask[(i64)n * g.S + p] : 1.f;
  const float up = upstream ? upstream[0] : 1.f;
  float dot = 0.f;
  // pass 1: g_pred_c, kept in g_out; pass 2: softmax backward
  for (int c = 0; c < K; ++c) {
    i64 q = ((i64)n * K + c) * g.S + p;
    float gp = 2.f * a_mse * m * m * E[q];
    if (c >= 1 && a_cont != 0.f) {
      const float* r0 = R + (((i64)n * (K - 1) + (c - 1)) * 2) * g.S;
      const float* r1 = r0 + g.S;
      float s = 0.f;
      // dL/dE(q) = sum_k sum_o k[o] R_k(q - o): visit source voxel s = q - o
#pragma unroll
      for (int a = (DIM == 3 ? -1 : 0); a <= (DIM == 3 ? 1 : 0); ++a) {
        int zz = z - a;
        if (zz < 0 || zz >= g.D) continue;
#pragma unroll
        for (int b = -1; b <= 1; ++b) {
          int yy = y - b;
          if (yy < 0 || yy >= g.H) continue;
#pragma unroll
          for (int cc = -1; cc <= 1; ++cc) {
            int xx = x - cc;
            if (xx < 0 || xx >= g.W) continue;
            i64 sidx = ((i64)zz * g.H + yy) * g.W + xx;
            if (DIM == 2)
              s += sob_h(b) * sob_hp(cc) * __ldg(r0 + sidx) + sob_hp(b) * sob_h(cc) * __ldg(r1 + sidx);
            else
              s += sob_h(a) * sob_hp(b) * sob_h(cc) * __ldg(r0 + sidx) +
                   sob_h(a) * sob_h(b) * sob_hp(cc) * __ldg(r1 + sidx);
          }
        }
      }
      gp += 2.f * a_cont * s;
    }
    g_out[q] = gp;
    dot += gp * pred[q];
  }
  for (int c = 0; c < K; ++c) {
    i64 q = ((i64)n * K + c) * g.S + p;
    g_out[q] = up * pred[q] * (g_out[q] - dot);
  }
}

}  // namespace advk

using namespace advk;

static void loss_scales(const Dims& g, int K, int d, float w_mse, float w_contour, float& a_mse, float& a_cont) {
  double NS = (double)g.N * (double)g.S;
  a_mse = (float)((double)w_mse / (NS * K) / NS);
  a_cont = (K > 1) ? (float)((double)w_contour / (double)(K - 1) / (double)(d == 2 ? 2 : 3) / NS) : 0.f;
}

extern "C" size_t advk_loss_scratch_floats(const advk_geom* gg, int K) {
  Dims g;
  if (!make_dims(gg, g) || K < 1) return 0;
  // E, pred: N*K*S each; R: N*(K-1)*2*S; 2 doubles (= 4 floats) of accumulators, 16-byte aligned slot
  return (size_t)((i64)g.N * g.S * (2 * K + 2 * (K - 1)) + 8);
}

extern "C" int advk_consistency_loss_fwd(const advk_geom* gg, int K, const float* output,
                                         const float* reference, const float* mask, float w_mse,
                                         float w_contour, int is_gt, float* scratch, float* loss,
                                         void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(K >= 1 && output && reference && scratch && loss, "null pointer / bad K");
  cudaStream_t st = (cudaStream_t)stream;
  const i64 NKS = (i64)g.N * K * g.S;
  double* acc = reinterpret_cast<double*>(scratch);          // scratch must be 8-byte aligned
  float* E = scratch + 8;
  float* pred = E + NKS;
  float* R = pred + NKS;
  float a_mse, a_cont;
  loss_scales(g, K, gg->d, w_mse, w_contour, a_mse, a_cont);
  cudaMemsetAsync(acc, 0, 2 * sizeof(double), st);
  dim3 grid(blocks_for(g.S, 256), g.N);
  ADVK_LAUNCH(K_loss_softmax, st, loss_softmax_kernel<<<grid, 256, 0, st>>>(g.S, g.N, K, output, reference, mask, is_gt, E, pred, acc));
  if (K > 1 && w_contour != 0.f) {
    dim3 grid2(blocks_for(g.S, 256), 1, g.N * (K - 1));
    if (gg->d == 2) ADVK_LAUNCH(K_loss_contour, st, loss_contour_kernel<2><<<grid2, 256, 0, st>>>(g, K, E, mask, R, acc));
    else ADVK_LAUNCH(K_loss_contour, st, loss_contour_kernel<3><<<grid2, 256, 0, st>>>(g, K, E, mask, R, acc));
  }
  ADVK_LAUNCH(K_loss_finalize, st, loss_finalize_kernel<<<1, 1, 0, st>>>(acc, a_mse, a_cont, loss));
  return check_launch("consistency_loss_fwd");
}

extern "C" int advk_consistency_loss_bwd(const advk_geom* gg, int K, const float* mask, float w_mse,
                                         float w_contour, const float* scratch, const float* upstream,
                                         float* g_output, void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(K >= 1 && scratch && g_output, "null pointer / bad K");
  cudaStream_t st = (cudaStream_t)stream;
  const i64 NKS = (i64)g.N * K * g.S;
  const float* E = scratch + 8;
  const float* pred = E + NKS;
  const float* R = pred + NKS;
  float a_mse, a_cont;
  loss_scales(g, K, gg->d, w_mse, w_contour, a_mse, a_cont);
  if (K <= 1) a_cont = 0.f;
  dim3 grid(blocks_for(g.S, 256), g.N);
  if (gg->d == 2) ADVK_LAUNCH(K_loss_grad, st, loss_grad_kernel<2><<<grid, 256, 0, st>>>(g, K, E, pred, R, mask, a_mse, a_cont, upstream, g_output));
  else ADVK_LAUNCH(K_loss_grad, st, loss_grad_kernel<3><<<grid, 256, 0, st>>>(g, K, E, pred, R, mask, a_mse, a_cont, upstream, g_output));
  return check_launch("consistency_loss_bwd");
}
