// advk_smooth_tma.cuh -- 3-D full-resolution Gaussian of the field build (adv_morph.py:377-389, 484-488) as ONE
// launch: xy tiles marching in z, input planes staged by TMA.  Included by advk_morph.cu (uses KT, KR,
// SmoothArgs, smooth_out).
//
// The reference's smoothing is nn.Conv3d(groups=3, kernel 9^3, padding 4): ZERO padding.  A TMA box load
// (cp.async.bulk.tensor.5d, tensor = [N][D][H][W][4 floats]) fills the part of a box that lies outside the
// tensor with zeros, so a haloed plane tile costs one instruction of one thread, no bounds arithmetic, and the
// halo is the convolution's padding exactly.  A CTA owns a 32 x 16 tile of (x, y) and a run of z planes:
//     producer (thread 0): keeps 2-3 plane tiles in flight on an mbarrier ring (arrive.expect_tx + TMA);
//     x pass: 6 warps, one thread per (row, 4 consecutive outputs): 12 LDS.128 of the float4 tile -> 108 FMA
//             (rows are 41 float4 apart = 4 banks mod 32, so a quarter-warp's 8 rows hit 8 different bank groups);
//     y pass: every thread owns one column and 2 rows: 10 LDS.32 per component from the planar x-pass result;
//     z pass: in registers, scatter form -- a new xy-smoothed plane adds w[8-m] * t to the 9 output planes it
//             belongs to, the oldest accumulator is complete and leaves through the output map; the plane loop is
//             unrolled by 9 so that the accumulator ring is addressed statically (no register shifts).
// One __syncthreads per plane (the x-pass result is double-buffered).  The input is ONE field: forward, the
// compose-with-base offsets r, written by the last squaring step; backward, g_field -- the gradient with
// respect to the UNCLAMPED field, which every consumer of the field has already masked by the clamp range
// (the warps clamp on load: advk_warp_field_bwd, the chain adjoints; torch.clamp in user code), so the mask
// the two-launch version applies once more (idempotent) is not repeated here.  The float4 round trip through
// HBM between the xy and z launches and the 90-instruction input map at every halo voxel are gone.
#pragma once
#include <cuda.h>   // CUtensorMap and the encoder's prototype only: the entry point is fetched through the runtime

namespace advk {

constexpr int FT_TX = 32, FT_TY = 16, FT_THREADS = 256;
constexpr int FT_BW = FT_TX + 2 * KR + 1;      // 41 columns: one more than needed, for the bank-friendly row stride
constexpr int FT_BH = FT_TY + 2 * KR;          // 24 rows
constexpr int FT_TILE_BYTES = FT_BW * FT_BH * 16;   // 15744 = 123 * 128
constexpr int FT_TMP_LD = FT_TX + 4;           // 36 floats: row stride of the planar x-pass result (4 banks mod 32)
constexpr int FT_TMP_FLOATS = 3 * FT_BH * FT_TMP_LD;
static_assert(FT_TILE_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
static_assert(KT % 3 == 0, "the plane loop is unrolled by KT: stage = u % 3 must be static");

template <int MODE> struct FtCfg {
  static constexpr int NT = 1;                 // tiles per stage
  static constexpr int NS = 3;                 // stages in flight
  static constexpr int SMEM = NS * NT * FT_TILE_BYTES + NS * FT_TMP_FLOATS * 4 + NS * 8;
};

struct FtArgs {
  Dims g;
  SmoothArgs<3> a;
  int zc, nzc;          // output planes per CTA, CTAs per sample along z
};

__device__ __forceinline__ unsigned ft_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ft_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ft_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FT_DONE_%=;\n"
      "bra FT_WAIT_%=;\n"
      "FT_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ft_tma_load(unsigned dst, const CUtensorMap* map, int x, int y, int z, int n, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"((unsigned long long)map), "r"(0), "r"(x), "r"(y), "r"(z), "r"(n), "r"(bar) : "memory");
}

// PPT = output rows per thread: 2 -> 256 threads (8 warps), 1 -> 512 threads (16 warps; half the accumulator
// registers per thread, twice the warps per SM to hide the shared-memory and FMA latencies)
template <int MODE, int PPT>
__global__ void __launch_bounds__(FT_THREADS * 2 / PPT, 2)
smooth3d_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ FtArgs q) {
  pdl_trigger();               // programmatic dependent launch (advk_common.cuh); pdl_wait() follows the barrier set-up
  constexpr int NT = FtCfg<MODE>::NT, NS = FtCfg<MODE>::NS;
  constexpr int XO = 2 * PPT;                   // consecutive outputs of one x-pass task
  constexpr int XW = 6 * (2 / PPT);             // warps that carry x-pass tasks: 24 rows x (32 / XO) runs / 32
  extern __shared__ __align__(128) unsigned char ft_smem[];
  const float4* s_in = reinterpret_cast<const float4*>(ft_smem);                  // [NS][NT][BH][BW]
  float* s_tmp = reinterpret_cast<float*>(ft_smem + NS * NT * FT_TILE_BYTES);     // [NS][3][BH][TMP_LD]
  const unsigned bar0 = ft_smem_u32(ft_smem + NS * NT * FT_TILE_BYTES + NS * FT_TMP_FLOATS * 4);
  const unsigned in0 = ft_smem_u32(ft_smem);
  const Dims& g = q.g;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * FT_TX, y0 = blockIdx.y * FT_TY;
  const int n = blockIdx.z / q.nzc, zb = (blockIdx.z - n * q.nzc) * q.zc;
  const int zend = min(zb + q.zc, g.D);
  const int np = (zend - zb) + 2 * KR;            // input planes zb-KR .. zend+KR-1 (those outside the volume are zeros)
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) ft_mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();                  // everything above worked on kernel parameters and shared memory only
  // plane i lives in stage i % NS.  Planes outside the volume are requested like the others: the whole box is
  // out of bounds, TMA fills it with zeros without touching memory, and the stage / phase of a plane stays a
  // pure function of i (static after unrolling the plane loop by KT = 3 * NS)
  auto issue = [&](int i, int s) {
    const int zi = zb - KR + i;
    const unsigned bar = bar0 + 8 * s;
    ft_mbar_expect(bar, NT * FT_TILE_BYTES);
    ft_tma_load(in0 + (s * NT) * FT_TILE_BYTES, &mapA, x0 - KR, y0 - KR, zi, n, bar);
    if (NT == 2) ft_tma_load(in0 + (s * NT + 1) * FT_TILE_BYTES, &mapB, x0 - KR, y0 - KR, zi, n, bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i)
      if (i < np) issue(i, i);
  }
  float w[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) w[k] = q.a.w[k];
  // x-pass task of this thread (warps 0 .. XW-1): row = 8 * (warp % 3) + (lane % 8), outputs XO * run .. + XO - 1
  const int wrp = tid >> 5, lane = tid & 31;
  const int xrow = 8 * (wrp % 3) + (lane & 7), xrun = (wrp / 3) * 4 + (lane >> 3);
  // y/z-pass positions of this thread: column tx, rows PPT * yg .. + PPT - 1
  const int tx = lane, yg = wrp;
  float acc[PPT][3][KT];
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int m = 0; m < KT; ++m) acc[j][c][m] = 0.f;
  unsigned kpar = 0;          // parity of base / KT: stage s is on its (base / NS + u / NS)-th use
  const i64 HW = (i64)g.H * g.W;
  const int gx = x0 + tx, gy0 = y0 + PPT * yg;
  // output map, per-thread constants: base coordinates of the column / the rows, border-clip bounds
  const float bx = base_coord_s(gx, g.W, g.stW);
  float by[PPT];
  bool oky[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    by[j] = base_coord_s(gy0 + j, g.H, g.stH);
    oky[j] = gx < g.W && gy0 + j < g.H;
  }
  const float mxW = (float)(g.W - 1), mxH = (float)(g.H - 1), mxD = (float)(g.D - 1);
  const i64 obase = (i64)n * g.S + (i64)gy0 * g.W + gx;
  for (int base = 0; base < np; base += KT) {
#pragma unroll
    for (int u = 0; u < KT; ++u) {
      const int i = base + u;
      if (i >= np) break;
      const int zi = zb - KR + i;
      const bool live = zi >= 0 && zi < g.D;            // block-uniform
      float t[PPT][3];
      // backward: the two fields of the output map are requested a whole plane of work ahead of their use
      float4 pn[PPT], p0[PPT];
      if (MODE == 1 && i >= 2 * KR) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          if (oky[j]) {
            const i64 idx = obase + (i64)(zi - KR) * HW + (i64)j * g.W;
            pn[j] = __ldg(q.a.C + idx); p0[j] = __ldg(q.a.D + idx);
          }
        }
      }
      const int s = u % NS;             // == i % NS (base is a multiple of KT = 3 * NS): static after unrolling
      ft_mbar_wait(bar0 + 8 * s, (kpar ^ (unsigned)(u / NS)) & 1u);
      if (live) {
        float* tmp = s_tmp + s * FT_TMP_FLOATS;
        if (wrp < XW) {
          const float4* rowA = s_in + ((s * NT) * FT_BH + xrow) * FT_BW + XO * xrun;
          float o[3][XO];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < XO; ++j) o[c][j] = 0.f;
#pragma unroll
          for (int e = 0; e < XO + 2 * KR; ++e) {
            const float4 v = rowA[e];
#pragma unroll
            for (int j = 0; j < XO; ++j) {
              const int k = e - j;
              if (k >= 0 && k < KT) { o[0][j] += w[k] * v.x; o[1][j] += w[k] * v.y; o[2][j] += w[k] * v.z; }
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float* d = tmp + (c * FT_BH + xrow) * FT_TMP_LD + XO * xrun;
            if (XO == 4) *reinterpret_cast<float4*>(d) = make_float4(o[c][0], o[c][1], o[c][2], o[c][XO - 1]);
            else *reinterpret_cast<float2*>(d) = make_float2(o[c][0], o[c][1]);
          }
        }
      }
      __syncthreads();                    // x-pass result complete; stage s has been read (or skipped) by everybody
      if (tid == 0 && i + NS < np) issue(i + NS, s);
      if (live) {
        const float* tmp = s_tmp + s * FT_TMP_FLOATS;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float in[PPT + 2 * KR];
#pragma unroll
          for (int e = 0; e < PPT + 2 * KR; ++e) in[e] = tmp[(c * FT_BH + PPT * yg + e) * FT_TMP_LD + tx];
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
            float sacc = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) sacc += w[k] * in[j + k];
            t[j][c] = sacc;
          }
        }
        // z pass: plane zi is tap KT-1-m of output plane zi - KR + m, whose ring slot (u + m) % KT is static after
        // unrolling
#pragma unroll
        for (int m = 0; m < KT; ++m) {
#pragma unroll
          for (int j = 0; j < PPT; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[j][c][(u + m) % KT] += w[KT - 1 - m] * t[j][c];
        }
      }
      if (i >= 2 * KR) {                  // output plane zo = zi - KR is complete: slot u
        const int zo = zi - KR;
        const float bz = base_coord_s(zo, g.D, g.stD);
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          if (!oky[j]) continue;
          const i64 idx = obase + (i64)zo * HW + (i64)j * g.W;
          const float s0 = acc[j][0][u], s1 = acc[j][1][u], s2 = acc[j][2][u];
          if (MODE == 0) {
            q.a.out[idx] = make_float4(s0 + bx, s1 + by[j], s2 + bz, 0.f);          // field = G (*) r + base
          } else {
            // border-clip mask of the compose-with-base grid_sample (smooth_out<3, 1>: gs_index, border padding):
            // the gradient passes where the pixel coordinate of phi_n - phi_0 + base lies strictly inside
            const float px = (((pn[j].x - p0[j].x) + bx + 1.f) / 2.f) * mxW;
            const float py = (((pn[j].y - p0[j].y) + by[j] + 1.f) / 2.f) * mxH;
            const float pz = (((pn[j].z - p0[j].z) + bz + 1.f) / 2.f) * mxD;
            q.a.out[idx] = make_float4((px > 0.f && px < mxW) ? s0 : 0.f, (py > 0.f && py < mxH) ? s1 : 0.f,
                                       (pz > 0.f && pz < mxD) ? s2 : 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < PPT; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[j][c][u] = 0.f;      // slot u now collects output plane zi + KR + 1
    }
    kpar ^= 1u;                           // KT / NS = 3 uses of every stage per block of KT planes: odd
  }
}

// ---- host side -----------------------------------------------------------------------------------------

typedef CUresult (*ft_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static ft_encode_fn ft_encoder() {
  static ft_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (ft_encode_fn)p;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// tensor map of one interleaved 3-D field [N][D][H][W] x float4, box = one haloed plane tile
static bool ft_make_map(const Dims& g, const void* field, CUtensorMap* map) {
  ft_encode_fn enc = ft_encoder();
  if (!enc || ((uintptr_t)field & 15)) return false;
  const cuuint64_t dims[5] = {4, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.D, (cuuint64_t)g.N};
  const cuuint64_t strides[4] = {16, 16ull * g.W, 16ull * g.W * g.H, 16ull * (cuuint64_t)g.S};
  const cuuint32_t box[5] = {4, (cuuint32_t)FT_BW, (cuuint32_t)FT_BH, 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(field), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int ft_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
  }
  return sms;
}

// Launches the fused smoothing; returns false (nothing launched) when TMA cannot be used -- the caller then
// runs the two-launch kernels.  inA: the field the Gaussian is applied to (forward: r; backward: g_field),
// inB: backward only, the field whose clamp range masks inA.
template <int MODE>
static bool launch_smooth_tma(const Dims& g, const MorphCfg& c, const void* inA, const void* inB, const void* C,
                              const void* D, void* out, cudaStream_t st) {
  // rows per thread: 2 (256 threads, 16 warps per SM; default) or 1 (512 threads, 32 warps per SM: 56 % more
  // instructions per voxel, measured slower: 39.9 / 53.8 vs 32.7 / 40.1 us, gpurun_out/r02o); ADVK_FT_PPT selects
  static int ready = -1, ppt = 2;
  if (ready < 0) {
    const char* e = getenv("ADVK_FT_PPT");
    ppt = (e && atoi(e) == 1) ? 1 : 2;
    ready = (cudaFuncSetAttribute(smooth3d_tma_kernel<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  FtCfg<MODE>::SMEM) == cudaSuccess &&
             cudaFuncSetAttribute(smooth3d_tma_kernel<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  FtCfg<MODE>::SMEM) == cudaSuccess) ? 1 : 0;
  }
  if (!ready) { (void)cudaGetLastError(); return false; }
  CUtensorMap mapA, mapB;
  if (!ft_make_map(g, inA, &mapA)) return false;
  mapB = mapA;      // (second operand slot: unused since the backward stages one field)
  FtArgs q;
  q.g = g;
  q.a.A = (const float4*)inA; q.a.B = (const float4*)inB; q.a.C = (const float4*)C; q.a.D = (const float4*)D;
  q.a.out = (float4*)out;
  for (int i = 0; i < KT; ++i) q.a.w[i] = c.w[i];
  const int tx = (g.W + FT_TX - 1) / FT_TX, ty = (g.H + FT_TY - 1) / FT_TY;
  // z runs: as many CTAs as fit the machine at once, never shorter than 8 planes (halo = 8)
  const i64 tiles = (i64)tx * ty * g.N;
  i64 chunks = (2LL * ft_sms()) / tiles;
  if (chunks < 1) chunks = 1;
  int zc = (int)((g.D + chunks - 1) / chunks);
  if (zc < 8) zc = 8;
  if (zc > g.D) zc = g.D;
  q.zc = zc; q.nzc = (g.D + zc - 1) / zc;
  if ((i64)g.N * q.nzc > 65535) return false;
  dim3 grid(tx, ty, (unsigned)(g.N * q.nzc));
  if (ppt == 2)
    ADVK_LAUNCH(MODE ? K_smooth_bwd : K_smooth_fwd, st,
                (launch_pdl((smooth3d_tma_kernel<MODE, 2>), grid, FT_THREADS, FtCfg<MODE>::SMEM, st, mapA, mapB, q)));
  else
    ADVK_LAUNCH(MODE ? K_smooth_bwd : K_smooth_fwd, st,
                (launch_pdl((smooth3d_tma_kernel<MODE, 1>), grid, 2 * FT_THREADS, FtCfg<MODE>::SMEM, st, mapA, mapB, q)));
  return true;
}

}  // namespace advk
