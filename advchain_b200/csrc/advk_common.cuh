// advk_common.cuh -- shared device helpers: geometry, base coordinates, grid_sample index
// math (ATen semantics, align_corners=True), linear-upsample index math (align_corners=False),
// warp/block reductions, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/advk.h"

namespace advk {

typedef long long i64;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---------------------------------------------------------------------------------------
// Kernel registry: every launch goes through ADVK_LAUNCH so that the library can (a) count its
// own launches (bench.py's `gpu_launches`) and (b) bracket chosen kernels with CUDA events on the
// launching stream (bench.py's live roofline measurement; advk_prof_* in include/advk.h).
#define ADVK_KERNELS(X)                                                                         \
  X(affine_theta_fwd) X(affine_theta_bwd) X(lowfield_fwd) X(lowfield_bwd) X(intensity_fwd)       \
  X(intensity_bwd) X(adjoint_axis_f) X(sumsq) X(update) X(clamp) X(clamp_bwd) X(nonzero)         \
  X(smooth_fwd) X(smooth_bwd) X(adjoint_axis) X(lowres_smooth) X(init_phi0) X(ss_step)           \
  X(ss_step_bwd) X(aos_to_planar) X(unorm2) X(warp_fwd) X(warp_bwd) X(loss_softmax)               \
  X(loss_contour) X(loss_contour_adj) X(loss_finalize) X(loss_grad) X(chain_fwd) X(chain_fwd_stage) X(chain_bwd)         \
  X(chain_bwd_stage) X(steps_check) X(chain_img_fwd) X(chain_pk_fwd) X(chain_img_bwd) X(chain_pk_bwd)        \
  X(publish)
enum KernelId {
#define ADVK_X(n) K_##n,
  ADVK_KERNELS(ADVK_X)
#undef ADVK_X
  K_COUNT
};

struct Prof {
  int slot;
  cudaStream_t st;
  Prof(int kid, cudaStream_t s);
  ~Prof();
};
#define ADVK_LAUNCH(kid, st, ...)  \
  do {                             \
    advk::Prof _prof((kid), (st)); \
    __VA_ARGS__;                   \
  } while (0)

#define ADVK_REQUIRE(cond, msg)                 \
  do {                                          \
    if (!(cond)) {                              \
      advk::set_error("%s: %s", __func__, msg); \
      return ADVK_ERR_ARG;                      \
    }                                           \
  } while (0)

// Division by a run-time constant without the (emulated, ~100-instruction) 64-bit divide: the
// multiply-high form used by CUTLASS' FastDivmod; exact for dividends below 2^31.
struct FastDiv {
  unsigned d, mul, shr;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.d = d ? d : 1; f.mul = 0; f.shr = 0;
  if (f.d != 1) {
    unsigned lg = 0;
    while ((1ull << lg) < f.d) ++lg;                    // ceil(log2 d)
    unsigned p = 31 + lg;
    f.mul = (unsigned)(((1ull << p) + f.d - 1) / f.d);
    f.shr = p - 32;
  }
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned fast_div(unsigned n, const FastDiv& f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+).  One PGD iteration is a chain of ~75 dependent launches, most of them
// 5-40 us long: with plain stream order every boundary costs the drain of the last wave plus the launch latency
// of the next grid.  Every kernel of this library starts with
//     pdl_wait();      griddepcontrol.wait: returns once the grids this one depends on have COMPLETED and their
//                      memory is visible -- nothing of the predecessor is read or overwritten before it
//     pdl_trigger();   griddepcontrol.launch_dependents: the next grid may be scheduled as soon as every CTA of
//                      this one has started (its CTAs then sit in the slots the last wave leaves free, blocked in
//                      their own pdl_wait, and run the moment this grid is done)
// and is launched through launch_pdl with cudaLaunchAttributeProgrammaticStreamSerialization.  Both instructions
// are no-ops in a grid launched without the attribute (advk_set_pdl(0) / ADVK_PDL=0, cooperative launches, user
// kernels in between).  A kernel launched by launch_pdl MUST call pdl_wait() before it touches global memory:
// tests/test_host_logic.py checks that over the sources.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();     // advk_misc.cu
void pdl_disable();     // a launch with the attribute was refused: plain launches from here on
template <typename... KA, typename... A>
static inline cudaError_t launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
  if (rc != cudaSuccess && cfg.numAttrs && (rc == cudaErrorNotSupported || rc == cudaErrorInvalidValue)) {
    (void)cudaGetLastError();
    pdl_disable();
    cfg.numAttrs = 0;
    rc = cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
  }
  return rc;
}
#endif

struct Dims {
  int N, D, H, W;
  i64 S;  // D*H*W
  FastDiv fW, fHW;            // voxel index -> (z, y, x), valid while S < 2^31
  float stD, stH, stW;        // linspace step 2/(size-1) per axis (0 when size == 1)
};

inline bool make_dims(const advk_geom* g, Dims& o) {
  if (!g || (g->d != 2 && g->d != 3) || g->N < 1 || g->D < 1 || g->H < 1 || g->W < 1) return false;
  if (g->d == 2 && g->D != 1) return false;
  o.N = g->N; o.D = g->D; o.H = g->H; o.W = g->W;
  o.S = (i64)g->D * g->H * g->W;
  o.fW = make_fastdiv((unsigned)g->W);
  o.fHW = make_fastdiv((unsigned)((i64)g->H * g->W < 0x7fffffffLL ? (i64)g->H * g->W : 1));
  o.stD = g->D > 1 ? 2.0f / (float)(g->D - 1) : 0.f;
  o.stH = g->H > 1 ? 2.0f / (float)(g->H - 1) : 0.f;
  o.stW = g->W > 1 ? 2.0f / (float)(g->W - 1) : 0.f;
  return true;
}

#ifdef __CUDACC__
// p (voxel index within one sample, < 2^31) -> coordinates
__device__ __forceinline__ void voxel_xyz(const Dims& g, unsigned p, int& x, int& y, int& z) {
  const unsigned zz = fast_div(p, g.fHW);
  const unsigned r = p - zz * (unsigned)(g.H * g.W);
  const unsigned yy = fast_div(r, g.fW);
  x = (int)(r - yy * (unsigned)g.W); y = (int)yy; z = (int)zz;
}
#endif

template <int DIM> struct FieldT;
template <> struct FieldT<2> { typedef float2 type; };
template <> struct FieldT<3> { typedef float4 type; };

// ---------------------------------------------------------------------------------------
// Base coordinate = torch.linspace(-1, 1, size)[i]  (RangeFactories: step=(end-start)/(steps-1),
// first half counted up from start, second half counted down from end).
// `single` is the value for size==1: linspace gives -1 (morph base grid, adv_morph.py:27),
// affine_grid's linspace_from_neg_one gives 0.
__device__ __forceinline__ float base_coord(int i, int size, float single = -1.f) {
  if (size <= 1) return single;
  float step = 2.0f / (float)(size - 1);
  return (i < size / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(size - 1 - i));
}
// same value with the step precomputed on the host (Dims::stW/stH/stD: the identical fp32 quotient)
__device__ __forceinline__ float base_coord_s(int i, int size, float step, float single = -1.f) {
  if (size <= 1) return single;
  return (i < size / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(size - 1 - i));
}

// ---------------------------------------------------------------------------------------
// grid_sample source index, align_corners=True (GridSampler.cuh:23-31, 52-80, 85-136, 138-168).
// Returns the (possibly clipped / reflected) pixel coordinate and d(index)/d(coord) in `mult`.
__device__ __forceinline__ float gs_reflect(float in, int twice_low, int twice_high, float& m) {
  if (twice_low == twice_high) { m = 0.f; return 0.f; }
  float mn = (float)twice_low / 2.f;
  float span = (float)(twice_high - twice_low) / 2.f;
  in = in - mn;
  float s = 1.f;
  if (in < 0.f) { s = -1.f; in = -in; }
  float extra = fmodf(in, span);
  int flips = (int)floorf(in / span);
  if (flips % 2 == 0) { m = s; return extra + mn; }
  m = -s;
  return span - extra + mn;
}

__device__ __forceinline__ float gs_index(float coord, int size, int pad, float& mult) {
  float x = ((coord + 1.f) / 2.f) * (float)(size - 1);
  mult = (float)(size - 1) / 2.f;
  if (pad == ADVK_PAD_REFLECTION) {
    float m;
    x = gs_reflect(x, 0, 2 * (size - 1), m);
    mult *= m;
  }
  if (pad != ADVK_PAD_ZEROS) {
    // clip_coordinates: min(size-1, max(x, 0)) (NaN -> 0 like ATen's ::max), gradient 0 on the clipped side
    const float mx = (float)(size - 1);
    const bool inside = (x > 0.f) && (x < mx);
    x = fminf(mx, fmaxf(x, 0.f));
    mult = inside ? mult : 0.f;
  }
  // safe_downgrade_to_int_range: anything non-finite or huge is "far outside"
  if (!(x <= 2147483646.f && x >= -2147483648.f)) x = -100.f;
  return x;
}

// One axis of a linear / nearest sampling stencil.
struct Axis {
  int i0;          // first corner index
  float w0, w1;    // weights of corner i0 and i0+1
  bool v0, v1;     // in-bounds flags
  float mult;      // d(pixel index)/d(normalised coord) incl. padding clip multiplier
};

__device__ __forceinline__ Axis make_axis(float coord, int size, int pad, int interp) {
  Axis a;
  float x = gs_index(coord, size, pad, a.mult);
  if (interp == ADVK_INTERP_NEAREST) {
    float r = nearbyintf(x);
    a.i0 = (int)r;
    a.w0 = 1.f; a.w1 = 0.f;
    a.v0 = (a.i0 >= 0 && a.i0 < size); a.v1 = false;
    a.mult = 0.f;
    return a;
  }
  float f = floorf(x);
  a.i0 = (int)f;
  a.w0 = (f + 1.f) - x;
  a.w1 = x - f;
  a.v0 = (a.i0 >= 0 && a.i0 < size);
  a.v1 = (a.i0 + 1 >= 0 && a.i0 + 1 < size);
  return a;
}

// Border padding + linear interpolation (the squaring steps and the compose-with-base of the field
// build): the clipped pixel coordinate lies in [0, size-1], so the int-range guard is dead code, corner
// i0 is always in bounds and corner i0+1 is in bounds unless the coordinate sits on the last voxel.
__device__ __forceinline__ Axis make_axis_border(float coord, int size) {
  Axis a;
  const float mx = (float)(size - 1);
  float x = ((coord + 1.f) / 2.f) * mx;
  const bool inside = (x > 0.f) && (x < mx);
  x = fminf(mx, fmaxf(x, 0.f));
  a.mult = inside ? mx / 2.f : 0.f;
  const float f = floorf(x);
  a.i0 = (int)f;
  a.w0 = (f + 1.f) - x;
  a.w1 = x - f;
  a.v0 = true;
  a.v1 = a.i0 + 1 < size;
  return a;
}

// ---------------------------------------------------------------------------------------
// F.interpolate(mode=linear, align_corners=False) source index (UpSample.cuh:96-131).
struct UpAxis { int i0, i1; float l0, l1; };
__device__ __forceinline__ UpAxis up_axis(int dst, int in_size, float scale) {
  UpAxis u;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  u.i0 = (int)src;
  if (u.i0 > in_size - 1) u.i0 = in_size - 1;
  u.i1 = u.i0 + ((u.i0 < in_size - 1) ? 1 : 0);
  u.l1 = src - (float)u.i0;
  u.l0 = 1.f - u.l1;
  return u;
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0. `sm` needs NV*32 floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = warp_sum(v[k]);
    if (lane == 0) sm[k * 32 + wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float s = (lane < nw) ? sm[k * 32 + lane] : 0.f;
      v[k] = warp_sum(s);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

inline unsigned blocks_for(i64 n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace advk
