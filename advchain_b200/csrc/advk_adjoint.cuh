// advk_adjoint.cuh -- adjoint of the align_corners=False linear upsample (F.interpolate / nn.Upsample,
// adv_morph.py:462-464, adv_bias.py:313-327) along the LAST TWO axes in one launch.
//
// The adjoint of a separable upsample is applied axis by axis; after the first (largest) axis the data has
// shrunk by the upsampling factor, and the two remaining per-axis launches (grids of 2-32 CTAs) cost 8-12 us
// each in latency at 128^3.  Here one CTA owns one output row (o, j): its 8 warps split the <= 2*H/Hl + 2 input
// rows that carry weight for node j, every lane sums whole columns (coalesced row reads), the 8 partial column
// sums go through shared memory, and each warp then reduces the columns of the output nodes it owns.
// Deterministic (no atomics).  in viewed as [outer][H][W] elements of T, out as [outer][Hl][Wl];
// val = (a - b) (b nullable), the result is scaled by vs.
#pragma once
#include "advk_common.cuh"

namespace advk {

__device__ __forceinline__ void ahw_zero(float& v) { v = 0.f; }
__device__ __forceinline__ void ahw_zero(float2& v) { v = make_float2(0.f, 0.f); }
__device__ __forceinline__ void ahw_zero(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void ahw_fma(float& s, float w, float a) { s += w * a; }
__device__ __forceinline__ void ahw_fma(float2& s, float w, float2 a) { s.x += w * a.x; s.y += w * a.y; }
__device__ __forceinline__ void ahw_fma(float4& s, float w, float4 a) { s.x += w * a.x; s.y += w * a.y; s.z += w * a.z; }
__device__ __forceinline__ float ahw_sub(float a, float b) { return a - b; }
__device__ __forceinline__ float2 ahw_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float4 ahw_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, 0.f); }
__device__ __forceinline__ float ahw_wsum(float v, float s) { return warp_sum(v) * s; }
__device__ __forceinline__ float2 ahw_wsum(float2 v, float s) { return make_float2(warp_sum(v.x) * s, warp_sum(v.y) * s); }
__device__ __forceinline__ float4 ahw_wsum(float4 v, float s) {
  return make_float4(warp_sum(v.x) * s, warp_sum(v.y) * s, warp_sum(v.z) * s, 0.f);
}

// input positions that can carry weight for output node j (widened by one; exact weights come from up_axis)
__device__ __forceinline__ void ahw_support(int j, int n_in, float scale, int& lo, int& hi) {
  lo = (int)floorf(((float)j - 0.5f) / scale - 0.5f) - 1;
  hi = (int)ceilf(((float)j + 1.5f) / scale - 0.5f) + 1;
  if (lo < 0 || j == 0) lo = 0;               // the source coordinate is clamped at 0 from below
  if (hi > n_in - 1) hi = n_in - 1;
}
__device__ __forceinline__ float ahw_weight(int p, int j, int n_out, float scale) {
  const UpAxis u = up_axis(p, n_out, scale);
  return (u.i0 == j ? u.l0 : 0.f) + (u.i1 == j ? u.l1 : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
adjoint_hw_kernel(const T* __restrict__ a, const T* __restrict__ b, float vs, T* __restrict__ out, int H, int W,
                  int Hl, int Wl, float sH, float sW) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  extern __shared__ float4 ahw_smem4[];
  T* colsum = reinterpret_cast<T*>(ahw_smem4);            // [8][W]
  const int j = blockIdx.x;
  const i64 o = blockIdx.y;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  int plo, phi;
  ahw_support(j, H, sH, plo, phi);
  const bool hb = (b != nullptr);
  // (rows outer, columns inner and unrolled: a warp's loads of one row are independent of each other, so they
  // are all in flight together instead of one latency per (row, column chunk))
  for (int x = lane; x < W; x += 32) ahw_zero(colsum[wrp * W + x]);
  for (int p = plo + wrp; p <= phi; p += 8) {
    const float w = ahw_weight(p, j, Hl, sH);
    if (w == 0.f) continue;
    const i64 rowq = (o * H + p) * W;
#pragma unroll 4
    for (int x = lane; x < W; x += 32) {
      T av = __ldg(a + rowq + x);
      if (hb) av = ahw_sub(av, __ldg(b + rowq + x));
      T s = colsum[wrp * W + x];
      ahw_fma(s, w, av);
      colsum[wrp * W + x] = s;
    }
  }
  __syncthreads();
  for (int jl = wrp; jl < Wl; jl += 8) {
    int xlo, xhi;
    ahw_support(jl, W, sW, xlo, xhi);
    T acc;
    ahw_zero(acc);
    for (int x = xlo + lane; x <= xhi; x += 32) {
      const float w = ahw_weight(x, jl, Wl, sW);
      if (w != 0.f) {
#pragma unroll
        for (int r = 0; r < 8; ++r) ahw_fma(acc, w, colsum[r * W + x]);
      }
    }
    acc = ahw_wsum(acc, vs);
    if (lane == 0) out[(o * Hl + j) * Wl + jl] = acc;
  }
}

// false: not launched (row too wide for the shared column sums, or too many outer slices) -- run the per-axis kernels
template <typename T>
static bool launch_adjoint_hw(int kid, const T* a, const T* b, float vs, T* out, i64 outer, int H, int W, int Hl, int Wl,
                              float sH, float sW, cudaStream_t st) {
  const size_t smem = sizeof(T) * 8 * (size_t)W;
  if (smem > 48 * 1024 || outer > 65535 || outer < 1) return false;
  dim3 grid((unsigned)Hl, (unsigned)outer);
  ADVK_LAUNCH(kid, st, (launch_pdl((adjoint_hw_kernel<T>), grid, 256, smem, st, a, b, vs, out, H, W, Hl, Wl, sH, sW)));
  return true;
}

}  // namespace advk
