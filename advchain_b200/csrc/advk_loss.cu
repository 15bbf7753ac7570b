// advk_loss.cu -- the consistency loss the PGD loop evaluates every step
// (calc_segmentation_consistency, common/loss.py:8-87 with scales=[0], divergence types 'mse'
// and 'contour'; contour_loss, common/loss.py:102-220), forward and backward w.r.t. the
// prediction, as three streaming kernels instead of the reference's ~600-770 ATen calls
// (2 softmaxes per term, a Conv module rebuilt and run per class per Sobel direction, ...).
//
//   pred = softmax(output), tgt = softmax(reference) (or reference itself when is_gt)
//   E_c  = pred_c - tgt_c
//   mse      = sum_c sum_p (m E_c)^2 / (N K S) / (N S)                       (quirk Q9)
//   kl       = sum_p m sum_c p_c (log p_c - log_softmax(output)_c) / (N S)   (common/loss.py:223-249;
//              p = softmax(reference), or where(reference == 0, 1e-8, 1 - 1e-8) when is_gt)
//   contour  = 1/(K-1) sum_{c>=1} 1/nk sum_k  sum_p (m (k * E_c))^2 / (N S)
//              k in {sobel_x, sobel_y} (2-D) | {gx, gx, gz} (3-D, quirk Q10: gy := gx)
// The Sobel correlation is linear, so conv(pred) - conv(tgt) = conv(E): one stencil per class.
//
//   K1 loss_softmax : E, pred  <- output, reference ; mse partial sums
//   K2 loss_contour : R_k = mult_k m^2 (k * E_c)    ; contour partial sums      (c >= 1)
//   K3 loss_contour_adj : s_c = sum_k k^T * R_k (c >= 1)
//   K4 loss_grad    : g_pred_c = 2 A_mse m^2 E_c + 2 A_cont s_c ; softmax backward
#include <stdlib.h>
#include "advk_common.cuh"

namespace advk {

// 1-D Sobel factors: h = smoothing [1,2,1], hp = derivative [1,0,-1]
__device__ __forceinline__ float sob_h(int o) { return o == 0 ? 2.f : 1.f; }      // o in {-1,0,1}
__device__ __forceinline__ float sob_hp(int o) { return (float)(-o); }           // [1,0,-1]

// KC > 0: class count known at compile time -> the logits of a voxel are loaded once into registers;
// KC == 0: any K, re-reading them from L1.
template <int KC>
__global__ void __launch_bounds__(256)
loss_softmax_kernel(i64 S, int N, int Krt, const float* __restrict__ out, const float* __restrict__ ref,
                    const float* __restrict__ mask, int is_gt, int want_kl, float* __restrict__ E,
                    float* __restrict__ pred, double* __restrict__ acc) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[64];
  const int K = KC > 0 ? KC : Krt;
  constexpr int KA = KC > 0 ? KC : 1;
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  float v[2] = {0.f, 0.f};
  if (p < S) {
    const float* o = out + (i64)n * K * S + p;
    const float* r = ref + (i64)n * K * S + p;
    float ov[KA], rv[KA];
    if (KC > 0) {
#pragma unroll
      for (int c = 0; c < KA; ++c) { ov[c] = o[c * S]; rv[c] = r[c * S]; }
    }
#define LO(c) (KC > 0 ? ov[(KC > 0) ? (c) : 0] : o[(c) * S])
#define LR(c) (KC > 0 ? rv[(KC > 0) ? (c) : 0] : r[(c) * S])
    float mo = -INFINITY, mr = -INFINITY;
#pragma unroll
    for (int c = 0; c < K; ++c) { mo = fmaxf(mo, LO(c)); mr = fmaxf(mr, LR(c)); }
    float so = 0.f, sr = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) { so += expf(LO(c) - mo); sr += expf(LR(c) - mr); }
    const float m = mask ? mask[(i64)n * S + p] : 1.f;
    const float lso = logf(so), lsr = logf(sr);
    float s = 0.f, plogp = 0.f, plogq = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const float oc = LO(c), rc = LR(c);
      float pc = expf(oc - mo) / so;
      float tc = is_gt ? rc : expf(rc - mr) / sr;
      float e = pc - tc;
      i64 q = ((i64)n * K + c) * S + p;
      E[q] = e;
      pred[q] = pc;
      float me = pc * m - tc * m;      // the reference masks both operands, then subtracts
      s += me * me;
      if (want_kl) {                    // kl_divergence, common/loss.py:223-249
        float pk, lpk;
        if (is_gt) { pk = (rc == 0.f) ? 1e-8f : 1.f - 1e-8f; lpk = logf(pk); }
        else { pk = tc; lpk = (rc - mr) - lsr; }
        plogp += m * (pk * lpk);
        plogq += m * (pk * ((oc - mo) - lso));
      }
    }
#undef LO
#undef LR
    v[0] = s;
    v[1] = plogp - plogq;
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc, (double)v[0]);
    if (want_kl) atomicAdd(acc + 2, (double)v[1]);
  }
}

// Sobel stencils through shared memory.  The 3x3(x3) kernels are products of the 1-D factors
// h = [1,2,1] and hp = [1,0,-1] (common/loss.py:148-203):
//     2-D  kx = h(H) hp(W),  ky = hp(H) h(W)
//     3-D  gx = h(D) hp(H) h(W),  gz = h(D) h(H) hp(W)        (quirk Q10: gy := gx)
// A CTA owns a 32 x 8 (x,y) tile and marches along z: each plane is staged once (tile + 1-voxel
// halo, zero outside the volume like the reference's padding=1 convolutions), every thread
// evaluates the two in-plane factors from 6 shared reads and keeps a 3-plane register window for the
// h(D) factor -- 1.06 global reads per voxel instead of 27 L1 gathers.
constexpr int LT_X = 32, LT_Y = 8, LT_Z = 16;

// in-plane factors at (tx,ty) of a staged (LT_Y+2) x (LT_X+2) tile:
//   a = hp(H) h(W) t,  b = h(H) hp(W) t
__device__ __forceinline__ void sobel_plane(const float (*t)[LT_X + 2], int tx, int ty, float& a, float& b) {
  const float r00 = t[ty][tx], r01 = t[ty][tx + 1], r02 = t[ty][tx + 2];
  const float r10 = t[ty + 1][tx], r12 = t[ty + 1][tx + 2];
  const float r20 = t[ty + 2][tx], r21 = t[ty + 2][tx + 1], r22 = t[ty + 2][tx + 2];
  // hp = [1,0,-1] over offsets (-1,0,+1); h = [1,2,1]
  a = (r00 + 2.f * r01 + r02) - (r20 + 2.f * r21 + r22);
  b = (r00 + 2.f * r10 + r20) - (r02 + 2.f * r12 + r22);
}

__device__ __forceinline__ void stage_plane(float (*t)[LT_X + 2], const float* __restrict__ src, const Dims& g,
                                            int pz, int x0, int y0) {
  const bool zin = pz >= 0 && pz < g.D;
  for (int i = threadIdx.x; i < (LT_Y + 2) * (LT_X + 2); i += LT_X * LT_Y) {
    const int ly = i / (LT_X + 2), lx = i - ly * (LT_X + 2);
    const int gy = y0 + ly - 1, gx = x0 + lx - 1;
    float v = 0.f;
    if (zin && gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) v = __ldg(src + ((i64)pz * g.H + gy) * g.W + gx);
    t[ly][lx] = v;
  }
}

// Register-prefetched variant for the z-marching loops: the (LT_Y+2)(LT_X+2) = 340 cells of a plane are
// two per thread; `fetch` issues the global loads of the NEXT plane while the current one is consumed
// from shared memory, `commit` stores them.
struct PlaneRegs { float a, b; };
__device__ __forceinline__ PlaneRegs fetch_plane(const float* __restrict__ src, const Dims& g, int pz, int x0, int y0) {
  PlaneRegs r; r.a = 0.f; r.b = 0.f;
  if (pz < 0 || pz >= g.D) return r;
  {
    const int i = threadIdx.x;
    const int ly = i / (LT_X + 2), lx = i - ly * (LT_X + 2);
    const int gy = y0 + ly - 1, gx = x0 + lx - 1;
    if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) r.a = __ldg(src + ((i64)pz * g.H + gy) * g.W + gx);
  }
  {
    const int i = threadIdx.x + LT_X * LT_Y;
    if (i < (LT_Y + 2) * (LT_X + 2)) {
      const int ly = i / (LT_X + 2), lx = i - ly * (LT_X + 2);
      const int gy = y0 + ly - 1, gx = x0 + lx - 1;
      if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) r.b = __ldg(src + ((i64)pz * g.H + gy) * g.W + gx);
    }
  }
  return r;
}
__device__ __forceinline__ void commit_plane(float (*t)[LT_X + 2], const PlaneRegs& r) {
  float* flat = &t[0][0];
  flat[threadIdx.x] = r.a;
  const int i = threadIdx.x + LT_X * LT_Y;
  if (i < (LT_Y + 2) * (LT_X + 2)) flat[i] = r.b;
}

// R has 2 planes per object class: [0] = the "x filter" response (2-D kx | 3-D gx, weight 2),
// [1] = the other (2-D ky | 3-D gz); both already multiplied by mult * m^2 for the backward.
template <int DIM>
__global__ void __launch_bounds__(LT_X * LT_Y)
loss_contour_kernel(Dims g, int K, int nzc, const float* __restrict__ E, const float* __restrict__ mask,
                    float* __restrict__ R, double* __restrict__ acc) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float tile[2][LT_Y + 2][LT_X + 2];
  __shared__ float red[32];
  const int tx = threadIdx.x % LT_X, ty = threadIdx.x / LT_X;
  const int zc = blockIdx.z % nzc;
  const int nc = blockIdx.z / nzc;
  const int n = nc / (K - 1), c = 1 + nc % (K - 1);
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const int x = x0 + tx, y = y0 + ty;
  const bool inxy = x < g.W && y < g.H;
  const float* e = E + ((i64)n * K + c) * g.S;
  const float* mk = mask ? mask + (i64)n * g.S : nullptr;
  float* r0 = R + (((i64)n * (K - 1) + (c - 1)) * 2) * g.S;
  float* r1 = r0 + g.S;
  const float mult0 = (DIM == 3) ? 2.f : 1.f;               // quirk Q10: gx is used for x and y
  float v[1] = {0.f};
  if (DIM == 2) {
    stage_plane(tile[0], e, g, 0, x0, y0);
    __syncthreads();
    float a, b;
    sobel_plane(tile[0], tx, ty, a, b);
    if (inxy) {
      const i64 p = (i64)y * g.W + x;
      const float m = mk ? mk[p] : 1.f;
      const float a0 = m * b, a1 = m * a;                   // kx = h(H) hp(W) = b,  ky = hp(H) h(W) = a
      v[0] = a0 * a0 + a1 * a1;
      r0[p] = m * a0;
      r1[p] = m * a1;
    }
  } else {
    const int zb = zc * LT_Z, ze = min(g.D, zb + LT_Z);
    float wa[3] = {0.f, 0.f, 0.f}, wb[3] = {0.f, 0.f, 0.f};
    PlaneRegs nxt = fetch_plane(e, g, zb - 1, x0, y0);
    for (int pz = zb - 1; pz <= ze; ++pz) {
      float (*t)[LT_X + 2] = tile[(pz - zb + 1) & 1];
      commit_plane(t, nxt);
      __syncthreads();
      if (pz < ze) nxt = fetch_plane(e, g, pz + 1, x0, y0);
      wa[0] = wa[1]; wa[1] = wa[2]; wb[0] = wb[1]; wb[1] = wb[2];
      sobel_plane(t, tx, ty, wa[2], wb[2]);
      const int zo = pz - 1;
      if (zo >= zb && inxy) {
        const float g0 = wa[0] + 2.f * wa[1] + wa[2];        // gx = h(D) hp(H) h(W)
        const float g1 = wb[0] + 2.f * wb[1] + wb[2];        // gz = h(D) h(H) hp(W)
        const i64 p = ((i64)zo * g.H + y) * g.W + x;
        const float m = mk ? mk[p] : 1.f;
        const float a0 = m * g0, a1 = m * g1;
        v[0] += mult0 * a0 * a0 + a1 * a1;
        r0[p] = mult0 * m * a0;
        r1[p] = m * a1;
      }
    }
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v[0]);
}

// Adjoint of the Sobel correlations: dL/dE_c(q) = sum_k sum_o k[o] R_k(q - o).  h is symmetric and hp
// antisymmetric, so this is MINUS the same correlation applied to R:  -(gx-stencil(R0) + gz-stencil(R1)).
// Written to `out` (class c's plane of g_output, finished by loss_grad_kernel).
template <int DIM>
__global__ void __launch_bounds__(LT_X * LT_Y)
loss_contour_adj_kernel(Dims g, int K, int nzc, const float* __restrict__ R, float* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float tile[2][2][LT_Y + 2][LT_X + 2];
  const int tx = threadIdx.x % LT_X, ty = threadIdx.x / LT_X;
  const int zc = blockIdx.z % nzc;
  const int nc = blockIdx.z / nzc;
  const int n = nc / (K - 1), c = 1 + nc % (K - 1);
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const int x = x0 + tx, y = y0 + ty;
  const bool inxy = x < g.W && y < g.H;
  const float* r0 = R + (((i64)n * (K - 1) + (c - 1)) * 2) * g.S;
  const float* r1 = r0 + g.S;
  float* o = out + ((i64)n * K + c) * g.S;
  if (DIM == 2) {
    stage_plane(tile[0][0], r0, g, 0, x0, y0);
    stage_plane(tile[0][1], r1, g, 0, x0, y0);
    __syncthreads();
    float a0, b0, a1, b1;
    sobel_plane(tile[0][0], tx, ty, a0, b0);
    sobel_plane(tile[0][1], tx, ty, a1, b1);
    if (inxy) o[(i64)y * g.W + x] = -(b0 + a1);             // kx-stencil(R0) + ky-stencil(R1)
  } else {
    const int zb = zc * LT_Z, ze = min(g.D, zb + LT_Z);
    float wa[3] = {0.f, 0.f, 0.f}, wb[3] = {0.f, 0.f, 0.f};
    PlaneRegs n0 = fetch_plane(r0, g, zb - 1, x0, y0), n1 = fetch_plane(r1, g, zb - 1, x0, y0);
    for (int pz = zb - 1; pz <= ze; ++pz) {
      const int buf = (pz - zb + 1) & 1;
      commit_plane(tile[buf][0], n0);
      commit_plane(tile[buf][1], n1);
      __syncthreads();
      if (pz < ze) { n0 = fetch_plane(r0, g, pz + 1, x0, y0); n1 = fetch_plane(r1, g, pz + 1, x0, y0); }
      wa[0] = wa[1]; wa[1] = wa[2]; wb[0] = wb[1]; wb[1] = wb[2];
      float dummy;
      sobel_plane(tile[buf][0], tx, ty, wa[2], dummy);       // hp(H) h(W) R0
      sobel_plane(tile[buf][1], tx, ty, dummy, wb[2]);       // h(H) hp(W) R1
      const int zo = pz - 1;
      if (zo >= zb && inxy)
        o[((i64)zo * g.H + y) * g.W + x] = -((wa[0] + 2.f * wa[1] + wa[2]) + (wb[0] + 2.f * wb[1] + wb[2]));
    }
  }
}

// 3-D, fused (the default; advk_loss_tune(0) / ADVK_LOSS_FUSED=0 selects the two kernels above): the Sobel
// responses AND their adjoint in ONE z-marching pass.  The adjoint at a voxel needs R on its 3 x 3 x 3
// neighbourhood, i.e. E on 5 x 5 x 5: the CTA stages E planes with a 2-voxel halo (36 x 12 cells), evaluates R
// on the tile + 1-voxel halo (34 x 10 cells, two per thread, 3-plane register windows for the h(D) factor) into
// shared memory, and every thread applies the adjoint stencils to that plane with a second 3-plane window.  R
// (48 MB written and read back per loss at 1 x 4 x 128^3) never exists in global memory; the contour adjoint
// s_c = sum_k k^T * R_k goes to the scratch slot R used to occupy, for loss_grad_kernel.  Same arithmetic per
// cell as loss_contour_kernel / loss_contour_adj_kernel: bit-identical s, R outside the volume is 0.
constexpr int LF_EX = LT_X + 4, LF_EY = LT_Y + 4, LF_EN = LF_EX * LF_EY;     // E plane: halo 2
constexpr int LF_RX = LT_X + 2, LF_RY = LT_Y + 2, LF_RN = LF_RX * LF_RY;     // R plane: halo 1

template <int ROW>
__device__ __forceinline__ void sobel_plane_w(const float* __restrict__ t, int tx, int ty, float& a, float& b) {
  const float* q0 = t + ty * ROW + tx;
  const float* q1 = q0 + ROW;
  const float* q2 = q1 + ROW;
  const float r00 = q0[0], r01 = q0[1], r02 = q0[2];
  const float r10 = q1[0], r12 = q1[2];
  const float r20 = q2[0], r21 = q2[1], r22 = q2[2];
  a = (r00 + 2.f * r01 + r02) - (r20 + 2.f * r21 + r22);
  b = (r00 + 2.f * r10 + r20) - (r02 + 2.f * r12 + r22);
}

__global__ void __launch_bounds__(LT_X * LT_Y)
loss_contour_fused_kernel(Dims g, int K, int nzc, const float* __restrict__ E, const float* __restrict__ mask,
                          float* __restrict__ Sg, double* __restrict__ acc) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  constexpr int NT = LT_X * LT_Y;
  __shared__ float et[2][LF_EN];
  __shared__ float rt[2][2][LF_RN];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int tx = tid % LT_X, ty = tid / LT_X;
  const int zc = blockIdx.z % nzc;
  const int nc = blockIdx.z / nzc;
  const int n = nc / (K - 1), c = 1 + nc % (K - 1);
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const int x = x0 + tx, y = y0 + ty;
  const bool inxy = x < g.W && y < g.H;
  const int HW = g.H * g.W;
  const float* e = E + ((i64)n * K + c) * g.S;
  const float* mk = mask ? mask + (i64)n * g.S : nullptr;
  float* so = Sg + ((i64)n * (K - 1) + (c - 1)) * g.S;
  const int zb = zc * LT_Z, ze = min(g.D, zb + LT_Z);
  // this thread's two E cells (staging) and two R cells (evaluation); in-plane voxel offset, -1 = outside
  int eoff[2], roff[2], rlx[2], rly[2];
  bool rown[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ie = tid + j * NT;
    const int ely = ie / LF_EX, elx = ie - ely * LF_EX;
    const int egy = y0 + ely - 2, egx = x0 + elx - 2;
    eoff[j] = (ie < LF_EN && egy >= 0 && egy < g.H && egx >= 0 && egx < g.W) ? egy * g.W + egx : -1;
    const int ir = tid + j * NT;
    rly[j] = ir / LF_RX; rlx[j] = ir - rly[j] * LF_RX;
    const int rgy = y0 + rly[j] - 1, rgx = x0 + rlx[j] - 1;
    roff[j] = (ir < LF_RN && rgy >= 0 && rgy < g.H && rgx >= 0 && rgx < g.W) ? rgy * g.W + rgx : -1;
    rown[j] = roff[j] >= 0 && rly[j] >= 1 && rly[j] <= LT_Y && rlx[j] >= 1 && rlx[j] <= LT_X;
  }
  const bool has_r1 = tid + NT < LF_RN;
  const float mult0 = 2.f;                                   // quirk Q10: gx is used for x and y
  float wa[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, wb[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float ua[3] = {0.f, 0.f, 0.f}, ub[3] = {0.f, 0.f, 0.f};
  float pe[2], mnext[2] = {1.f, 1.f};
  float v[1] = {0.f};
  {
    const int pz = zb - 2;
    const bool zin = pz >= 0 && pz < g.D;
#pragma unroll
    for (int j = 0; j < 2; ++j) pe[j] = (zin && eoff[j] >= 0) ? __ldg(e + (i64)pz * HW + eoff[j]) : 0.f;
  }
  for (int pz = zb - 2; pz <= ze + 1; ++pz) {
    const int it = pz - (zb - 2);
    float* et_ = et[it & 1];
    et_[tid] = pe[0];
    if (tid + NT < LF_EN) et_[tid + NT] = pe[1];
    __syncthreads();
    const float m0 = mnext[0], m1 = mnext[1];                // mask of plane zr = pz - 1 (loaded one iteration ago)
    if (pz <= ze) {                                          // next E plane, and the mask of this plane for the next round
      const int qz = pz + 1;
      const bool zin = qz >= 0 && qz < g.D;
#pragma unroll
      for (int j = 0; j < 2; ++j) pe[j] = (zin && eoff[j] >= 0) ? __ldg(e + (i64)qz * HW + eoff[j]) : 0.f;
      const bool pin = pz >= 0 && pz < g.D;
#pragma unroll
      for (int j = 0; j < 2; ++j) mnext[j] = (mk && pin && roff[j] >= 0) ? __ldg(mk + (i64)pz * HW + roff[j]) : 1.f;
    }
    wa[0][0] = wa[0][1]; wa[0][1] = wa[0][2]; wb[0][0] = wb[0][1]; wb[0][1] = wb[0][2];
    sobel_plane_w<LF_EX>(et_, rlx[0], rly[0], wa[0][2], wb[0][2]);
    if (has_r1) {
      wa[1][0] = wa[1][1]; wa[1][1] = wa[1][2]; wb[1][0] = wb[1][1]; wb[1][1] = wb[1][2];
      sobel_plane_w<LF_EX>(et_, rlx[1], rly[1], wa[1][2], wb[1][2]);
    }
    if (it < 2) continue;                                    // (block-uniform)
    // E planes pz-2 .. pz are in the windows: R of plane zr = pz - 1 on the tile + halo 1
    const int zr = pz - 1;
    const bool zrin = zr >= 0 && zr < g.D;
    const bool zown = zr >= zb && zr < ze;
    float* r0_ = rt[it & 1][0];
    float* r1_ = rt[it & 1][1];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 1 && !has_r1) break;
      float R0 = 0.f, R1 = 0.f;
      if (zrin && roff[j] >= 0) {
        const float g0 = wa[j][0] + 2.f * wa[j][1] + wa[j][2];        // gx = h(D) hp(H) h(W)
        const float g1 = wb[j][0] + 2.f * wb[j][1] + wb[j][2];        // gz = h(D) h(H) hp(W)
        const float m = j ? m1 : m0;
        const float a0 = m * g0, a1 = m * g1;
        if (zown && rown[j]) v[0] += mult0 * a0 * a0 + a1 * a1;
        R0 = mult0 * m * a0;
        R1 = m * a1;
      }
      r0_[tid + j * NT] = R0;
      r1_[tid + j * NT] = R1;
    }
    __syncthreads();
    ua[0] = ua[1]; ua[1] = ua[2]; ub[0] = ub[1]; ub[1] = ub[2];
    float dummy;
    sobel_plane_w<LF_RX>(r0_, tx, ty, ua[2], dummy);         // hp(H) h(W) R0
    sobel_plane_w<LF_RX>(r1_, tx, ty, dummy, ub[2]);         // h(H) hp(W) R1
    if (it >= 4 && inxy) {                                   // R planes zr-2 .. zr: the adjoint at plane zo = zr - 1
      const int zo = zr - 1;
      so[((i64)zo * g.H + y) * g.W + x] = -((ua[0] + 2.f * ua[1] + ua[2]) + (ub[0] + 2.f * ub[1] + ub[2]));
    }
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v[0]);
}

// 2-D form of the one-pass contour kernel: one plane, so no register windows -- E on tile + 2-voxel halo, R on
// tile + 1-voxel halo in shared memory, the adjoint stencils at the tile's own voxels.  Same arithmetic per cell as
// loss_contour_kernel<2> / loss_contour_adj_kernel<2> (kx = h(H) hp(W), ky = hp(H) h(W)).
__global__ void __launch_bounds__(LT_X * LT_Y)
loss_contour_fused2d_kernel(Dims g, int K, const float* __restrict__ E, const float* __restrict__ mask,
                            float* __restrict__ Sg, double* __restrict__ acc) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  constexpr int NT = LT_X * LT_Y;
  __shared__ float et[LF_EN];
  __shared__ float rt[2][LF_RN];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int tx = tid % LT_X, ty = tid / LT_X;
  const int nc = blockIdx.z;
  const int n = nc / (K - 1), c = 1 + nc % (K - 1);
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const int x = x0 + tx, y = y0 + ty;
  const float* e = E + ((i64)n * K + c) * g.S;
  const float* mk = mask ? mask + (i64)n * g.S : nullptr;
  float* so = Sg + ((i64)n * (K - 1) + (c - 1)) * g.S;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ie = tid + j * NT;
    if (ie < LF_EN) {
      const int ely = ie / LF_EX, elx = ie - ely * LF_EX;
      const int gy = y0 + ely - 2, gx = x0 + elx - 2;
      et[ie] = (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) ? __ldg(e + gy * g.W + gx) : 0.f;
    }
  }
  __syncthreads();
  float v[1] = {0.f};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ir = tid + j * NT;
    if (ir < LF_RN) {
      const int rly = ir / LF_RX, rlx = ir - rly * LF_RX;
      const int gy = y0 + rly - 1, gx = x0 + rlx - 1;
      float R0 = 0.f, R1 = 0.f;
      if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {
        float a, b;
        sobel_plane_w<LF_EX>(et, rlx, rly, a, b);
        const float m = mk ? __ldg(mk + gy * g.W + gx) : 1.f;
        const float a0 = m * b, a1 = m * a;                  // kx = h(H) hp(W) = b,  ky = hp(H) h(W) = a
        if (rly >= 1 && rly <= LT_Y && rlx >= 1 && rlx <= LT_X) v[0] += a0 * a0 + a1 * a1;
        R0 = m * a0;
        R1 = m * a1;
      }
      rt[0][ir] = R0;
      rt[1][ir] = R1;
    }
  }
  __syncthreads();
  float a0, b0, a1, b1;
  sobel_plane_w<LF_RX>(rt[0], tx, ty, a0, b0);
  sobel_plane_w<LF_RX>(rt[1], tx, ty, a1, b1);
  if (x < g.W && y < g.H) so[(i64)y * g.W + x] = -(b0 + a1);   // kx-stencil(R0) + ky-stencil(R1)
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v[0]);
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, float a_mse, float a_cont, float a_kl,
                                     float* __restrict__ loss) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  loss[0] = (float)((double)a_mse * acc[0] + (double)a_cont * acc[1] + (double)a_kl * acc[2]);
}

// g_pred_c = 2 A_mse m^2 E_c + 2 A_cont s_c  (s_c = contour adjoint, already in g_out for c >= 1);
// g_out_c = pred_c (g_pred_c - sum_j g_pred_j pred_j) * upstream
template <int KC>
__global__ void __launch_bounds__(256)
loss_grad_kernel(i64 S, int Krt, const float* __restrict__ E, const float* __restrict__ pred,
                 const float* __restrict__ mask, float a_mse, float a_cont, float a_kl, int is_gt,
                 const float* __restrict__ upstream, const float* sc, int sc_K, int sc_off, float* g_out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  // sc: the contour adjoint s_c, plane (n * sc_K + c - sc_off): the scratch slot of the fused 3-D kernel
  // (sc_K = K-1, sc_off = 1) or g_out itself (sc_K = K, sc_off = 0: loss_contour_adj_kernel wrote it there)
  const int K = KC > 0 ? KC : Krt;
  constexpr int KA = KC > 0 ? KC : 1;
  const int n = blockIdx.y;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S) return;
  const float m = mask ? mask[(i64)n * S + p] : 1.f;
  const float up = upstream ? upstream[0] : 1.f;
  const float cm = 2.f * a_mse * m * m, cc = 2.f * a_cont;
  const i64 q0 = (i64)n * K * S + p;
  if (KC > 0) {
    // everything of the voxel in registers: E, pred and the contour adjoint are read once, g_out written once
    float e[KA], pr[KA], gp[KA];
    float dot = 0.f, psum = 0.f;
#pragma unroll
    for (int c = 0; c < KA; ++c) {
      const i64 q = q0 + (i64)c * S;
      e[c] = E[q]; pr[c] = pred[q];
      gp[c] = cm * e[c];
      if (c >= 1 && a_cont != 0.f) gp[c] += cc * sc[((i64)n * sc_K + (c - sc_off)) * S + p];
      dot += gp[c] * pr[c];
      const float t = pr[c] - e[c];
      psum += is_gt ? ((t == 0.f) ? 1e-8f : 1.f - 1e-8f) : t;
    }
#pragma unroll
    for (int c = 0; c < KA; ++c) {
      float go = pr[c] * (gp[c] - dot);
      if (a_kl != 0.f) {
        const float t = pr[c] - e[c];
        const float pk = is_gt ? ((t == 0.f) ? 1e-8f : 1.f - 1e-8f) : t;
        go += a_kl * m * (pr[c] * psum - pk);
      }
      g_out[q0 + (i64)c * S] = up * go;
    }
    return;
  }
  float dot = 0.f;
  i64 q = q0;
  for (int c = 0; c < K; ++c, q += S) {
    float gp = cm * E[q];
    if (c >= 1 && a_cont != 0.f) gp += cc * sc[((i64)n * sc_K + (c - sc_off)) * S + p];
    g_out[q] = gp;
    dot += gp * pred[q];
  }
  // 'kl' differentiates straight to the logits: d/do_c = A_kl m (softmax(o)_c sum_j p_j - p_c)
  float psum = 0.f;
  if (a_kl != 0.f) {
    q = q0;
    for (int c = 0; c < K; ++c, q += S) {
      const float t = pred[q] - E[q];
      psum += is_gt ? ((t == 0.f) ? 1e-8f : 1.f - 1e-8f) : t;
    }
  }
  q = q0;
  for (int c = 0; c < K; ++c, q += S) {
    float go = pred[q] * (g_out[q] - dot);
    if (a_kl != 0.f) {
      const float t = pred[q] - E[q];
      const float pk = is_gt ? ((t == 0.f) ? 1e-8f : 1.f - 1e-8f) : t;
      go += a_kl * m * (pred[q] * psum - pk);
    }
    g_out[q] = up * go;
  }
}

}  // namespace advk

using namespace advk;

// 1 (default; environment ADVK_LOSS_FUSED): contour term by loss_contour_fused_kernel / loss_contour_fused2d_kernel;
// 0: the two-kernel predecessor.  Must not change between a forward call and the backward call that consumes its scratch.
static int g_loss_fused = -1;
static int loss_fused() {
  if (g_loss_fused < 0) {
    const char* e = getenv("ADVK_LOSS_FUSED");
    g_loss_fused = e ? (atoi(e) != 0) : 1;
  }
  return g_loss_fused;
}
extern "C" int advk_loss_tune(int fused) {
  int prev = loss_fused();
  if (fused >= 0) g_loss_fused = fused != 0;
  return prev;
}

static void loss_scales(const Dims& g, int K, int d, float w_mse, float w_contour, float w_kl, float& a_mse,
                        float& a_cont, float& a_kl) {
  double NS = (double)g.N * (double)g.S;
  a_mse = (float)((double)w_mse / (NS * K) / NS);
  a_cont = (K > 1) ? (float)((double)w_contour / (double)(K - 1) / (double)(d == 2 ? 2 : 3) / NS) : 0.f;
  a_kl = (float)((double)w_kl / NS);
}

extern "C" size_t advk_loss_scratch_floats(const advk_geom* gg, int K) {
  Dims g;
  if (!make_dims(gg, g) || K < 1) return 0;
  // E, pred: N*K*S each; R: N*(K-1)*2*S; 3 doubles (= 6 floats) of accumulators in a 32-byte slot
  return (size_t)((i64)g.N * g.S * (2 * K + 2 * (K - 1)) + 8);
}

extern "C" int advk_consistency_loss_fwd(const advk_geom* gg, int K, const float* output,
                                         const float* reference, const float* mask, float w_mse,
                                         float w_contour, float w_kl, int is_gt, float* scratch, float* loss,
                                         void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(K >= 1 && output && reference && scratch && loss, "null pointer / bad K");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  cudaStream_t st = (cudaStream_t)stream;
  const i64 NKS = (i64)g.N * K * g.S;
  double* acc = reinterpret_cast<double*>(scratch);          // scratch must be 8-byte aligned
  float* E = scratch + 8;
  float* pred = E + NKS;
  float* R = pred + NKS;
  float a_mse, a_cont, a_kl;
  loss_scales(g, K, gg->d, w_mse, w_contour, w_kl, a_mse, a_cont, a_kl);
  cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  dim3 grid(blocks_for(g.S, 256), g.N);
  const int wkl = w_kl != 0.f ? 1 : 0;
#define ADVK_SOFTMAX(KC) ADVK_LAUNCH(K_loss_softmax, st, launch_pdl((loss_softmax_kernel<KC>), grid, 256, 0, st, g.S, g.N, K, output, reference, mask, is_gt, wkl, E, pred, acc))
  switch (K) {
    case 2: ADVK_SOFTMAX(2); break;
    case 3: ADVK_SOFTMAX(3); break;
    case 4: ADVK_SOFTMAX(4); break;
    case 5: ADVK_SOFTMAX(5); break;
    case 8: ADVK_SOFTMAX(8); break;
    default: ADVK_SOFTMAX(0); break;
  }
#undef ADVK_SOFTMAX
  if (K > 1 && w_contour != 0.f) {
    const int nzc = (gg->d == 3) ? (g.D + LT_Z - 1) / LT_Z : 1;
    dim3 grid2((g.W + LT_X - 1) / LT_X, (g.H + LT_Y - 1) / LT_Y, (unsigned)(g.N * (K - 1) * nzc));
    if (gg->d == 2 && loss_fused()) ADVK_LAUNCH(K_loss_contour, st, launch_pdl((loss_contour_fused2d_kernel), grid2, LT_X * LT_Y, 0, st, g, K, E, mask, R, acc));
    else if (gg->d == 2) ADVK_LAUNCH(K_loss_contour, st, launch_pdl((loss_contour_kernel<2>), grid2, LT_X * LT_Y, 0, st, g, K, nzc, E, mask, R, acc));
    else if (loss_fused()) ADVK_LAUNCH(K_loss_contour, st, launch_pdl((loss_contour_fused_kernel), grid2, LT_X * LT_Y, 0, st, g, K, nzc, E, mask, R, acc));
    else ADVK_LAUNCH(K_loss_contour, st, launch_pdl((loss_contour_kernel<3>), grid2, LT_X * LT_Y, 0, st, g, K, nzc, E, mask, R, acc));
  }
  ADVK_LAUNCH(K_loss_finalize, st, launch_pdl((loss_finalize_kernel), 1, 1, 0, st, acc, a_mse, a_cont, a_kl, loss));
  return check_launch("consistency_loss_fwd");
}

extern "C" int advk_consistency_loss_bwd(const advk_geom* gg, int K, const float* mask, float w_mse,
                                         float w_contour, float w_kl, int is_gt, const float* scratch,
                                         const float* upstream, float* g_output, void* stream) {
  Dims g;
  ADVK_REQUIRE(make_dims(gg, g), "bad geometry");
  ADVK_REQUIRE(K >= 1 && scratch && g_output, "null pointer / bad K");
  ADVK_REQUIRE(g.S < 2147483647LL, "more than 2^31 voxels per sample");
  cudaStream_t st = (cudaStream_t)stream;
  const i64 NKS = (i64)g.N * K * g.S;
  const float* E = scratch + 8;
  const float* pred = E + NKS;
  const float* R = pred + NKS;
  float a_mse, a_cont, a_kl;
  loss_scales(g, K, gg->d, w_mse, w_contour, w_kl, a_mse, a_cont, a_kl);
  if (K <= 1) a_cont = 0.f;
  const bool fused = loss_fused() != 0;                     // the forward left s_c where R would have been
  if (a_cont != 0.f && !fused) {
    const int nzc = (gg->d == 3) ? (g.D + LT_Z - 1) / LT_Z : 1;
    dim3 grid2((g.W + LT_X - 1) / LT_X, (g.H + LT_Y - 1) / LT_Y, (unsigned)(g.N * (K - 1) * nzc));
    if (gg->d == 2) ADVK_LAUNCH(K_loss_contour_adj, st, launch_pdl((loss_contour_adj_kernel<2>), grid2, LT_X * LT_Y, 0, st, g, K, nzc, R, g_output));
    else ADVK_LAUNCH(K_loss_contour_adj, st, launch_pdl((loss_contour_adj_kernel<3>), grid2, LT_X * LT_Y, 0, st, g, K, nzc, R, g_output));
  }
  dim3 grid(blocks_for(g.S, 256), g.N);
#define ADVK_LGRAD(KC) ADVK_LAUNCH(K_loss_grad, st, launch_pdl((loss_grad_kernel<KC>), grid, 256, 0, st, g.S, K, E, pred, mask, a_mse, a_cont, a_kl, is_gt, upstream, fused ? R : g_output, fused ? K - 1 : K, fused ? 1 : 0, g_output))
  switch (K) {
    case 2: ADVK_LGRAD(2); break;
    case 3: ADVK_LGRAD(3); break;
    case 4: ADVK_LGRAD(4); break;
    case 5: ADVK_LGRAD(5); break;
    case 8: ADVK_LGRAD(8); break;
    default: ADVK_LGRAD(0); break;
  }
#undef ADVK_LGRAD
  return check_launch("consistency_loss_bwd");
}
