// advk_misc.cu -- error plumbing, device query, the fused PGD parameter update (per-sample L2
// norm + axpy / sign step, adv_transformation_base.py:129-156 and the optimize_parameters
// bodies), and the solver's pointwise glue (intensity clamp, valid-region mask binarisation).
#include <stdarg.h>
#include "advk_common.cuh"

namespace advk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d: %s", what, (int)e, cudaGetErrorString(e));
    return ADVK_ERR_CUDA;
  }
  return ADVK_OK;
}

// ---------------------------------------------------------------------------------------
// Launch accounting + optional per-kernel CUDA-event timing (see ADVK_LAUNCH).
static const char* const g_kernel_names[K_COUNT + 1] = {
#define ADVK_X(n) #n,
    ADVK_KERNELS(ADVK_X)
#undef ADVK_X
    "?"};
static unsigned long long g_launches[K_COUNT] = {0};
static int g_prof_kid = -2;           // -2 off, -1 all kernels, >=0 one kernel id
static int g_prof_cap = 0, g_prof_used = 0;
static cudaEvent_t* g_prof_ev = nullptr;   // 2 events per record
static int* g_prof_ids = nullptr;
static int* g_prof_cnt = nullptr;          // launches covered by a record (> 1 only when runs are grouped)
// advk_prof_group_runs: a run of back-to-back launches of the profiled kernel on one stream (the 8 adjoint
// squaring steps of a field build) is bracketed ONCE -- an event pair around every 40 us launch adds its own
// 2-4 us to each record
static int g_prof_group = 0, g_prof_open = -1;
static cudaStream_t g_prof_open_st = nullptr;

Prof::Prof(int kid, cudaStream_t s) : slot(-1), st(s) {
  if (kid >= 0 && kid < K_COUNT) ++g_launches[kid];
  if (g_prof_kid == -2) return;
  if (g_prof_kid >= 0 && g_prof_kid != kid) { g_prof_open = -1; return; }     // another kernel ends the run
  if (g_prof_group && g_prof_kid >= 0 && g_prof_open >= 0 && g_prof_open_st == s) {
    slot = g_prof_open;                       // same run: keep its start event, move its end event
    ++g_prof_cnt[slot];
    return;
  }
  if (g_prof_used >= g_prof_cap) { g_prof_open = -1; return; }
  slot = g_prof_used++;
  g_prof_ids[slot] = kid;
  g_prof_cnt[slot] = 1;
  cudaEventRecord(g_prof_ev[2 * slot], st);
  if (g_prof_group && g_prof_kid >= 0) { g_prof_open = slot; g_prof_open_st = s; }
}
Prof::~Prof() {
  if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot + 1], st);
}

// sum of squares per sample, accumulated in double across blocks
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, size_t per, double* __restrict__ out) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[32];
  const int n = blockIdx.y;
  const float* src = g + (size_t)n * per;
  float v[1] = {0.f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    float t = src[i];
    v[0] += t * t;
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(out + n, (double)v[0]);
}

__global__ void __launch_bounds__(256)
update_kernel(float* param, const float* grad, float step, int mode, size_t per,
              const double* __restrict__ ss, const float* __restrict__ guard) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  // NaN/Inf guard of the PGD loop (adv_compose_solver.py:345-346) evaluated on the device: a
  // non-finite loss leaves the parameters untouched, without a host round trip
  if (guard) {
    float gv = guard[0];
    if (isnan(gv) || isinf(gv)) return;
  }
  const int n = blockIdx.y;
  float inv = 0.f;
  if (mode == ADVK_UPD_L2_ASCENT || mode == ADVK_UPD_L2_POWER) inv = 1.f / ((float)sqrt(ss[n]) + 1e-20f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    size_t q = (size_t)n * per + i;
    float gval = grad[q];
    float sg = (gval > 0.f) ? 1.f : ((gval < 0.f) ? -1.f : 0.f);
    float r;
    switch (mode) {
      case ADVK_UPD_L2_ASCENT: r = param[q] + step * (gval * inv); break;
      case ADVK_UPD_SIGN_ASCENT: r = param[q] + step * sg; break;
      case ADVK_UPD_L2_POWER: r = gval * inv; break;
      default: r = sg; break;
    }
    param[q] = r;
  }
}

// Small parameters (control points, velocities, affine): norm + update in ONE launch, one CTA per
// sample.  Same arithmetic as sumsq_kernel + update_kernel (fp32 partials, double across warps).
__global__ void __launch_bounds__(256)
small_update_kernel(float* param, const float* grad, float step, int mode, int per,
                    const float* __restrict__ guard) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  __shared__ float red[32];
  __shared__ float s_inv;
  if (guard) {
    float gv = guard[0];
    if (isnan(gv) || isinf(gv)) return;
  }
  const int n = blockIdx.x;
  const float* g = grad + (size_t)n * per;
  float* p = param + (size_t)n * per;
  if (mode == ADVK_UPD_L2_ASCENT || mode == ADVK_UPD_L2_POWER) {
    float v[1] = {0.f};
    for (int i = threadIdx.x; i < per; i += blockDim.x) { float t = g[i]; v[0] += t * t; }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) s_inv = 1.f / ((float)sqrt((double)v[0]) + 1e-20f);
    __syncthreads();
  }
  const float inv = (mode == ADVK_UPD_L2_ASCENT || mode == ADVK_UPD_L2_POWER) ? s_inv : 0.f;
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    float gval = g[i];
    float sg = (gval > 0.f) ? 1.f : ((gval < 0.f) ? -1.f : 0.f);
    float r;
    switch (mode) {
      case ADVK_UPD_L2_ASCENT: r = p[i] + step * (gval * inv); break;
      case ADVK_UPD_SIGN_ASCENT: r = p[i] + step * sg; break;
      case ADVK_UPD_L2_POWER: r = gval * inv; break;
      default: r = sg; break;
    }
    p[i] = r;
  }
}

__global__ void clamp_kernel(const float* __restrict__ x, float lo, float hi, float* __restrict__ out, size_t n) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = fminf(fmaxf(x[i], lo), hi);
}
__global__ void clamp_bwd_kernel(const float* __restrict__ go, const float* __restrict__ x, float lo, float hi,
                                 float* __restrict__ gx, size_t n) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    gx[i] = (v >= lo && v <= hi) ? go[i] : 0.f;
  }
}
__global__ void nonzero_kernel(float* __restrict__ x, size_t n) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = (x[i] != 0.f) ? 1.f : 0.f;
}

static unsigned ew_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  return (unsigned)(b > 148 * 16 ? 148 * 16 : (b ? b : 1));
}

}  // namespace advk

using namespace advk;

extern "C" int advk_abi_version(void) { return ADVK_ABI_VERSION; }
extern "C" const char* advk_last_error(void) { return g_err; }

extern "C" int advk_kernel_count(void) { return K_COUNT; }
extern "C" const char* advk_kernel_name(int kid) {
  return g_kernel_names[(kid >= 0 && kid < K_COUNT) ? kid : K_COUNT];
}
static int g_pdl = -1;
namespace advk {
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ADVK_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
void pdl_disable() { g_pdl = 0; }
}  // namespace advk
extern "C" int advk_set_pdl(int enable) {
  const int prev = advk::pdl_enabled() ? 1 : 0;
  g_pdl = enable ? 1 : 0;
  return prev;
}
extern "C" unsigned long long advk_launch_count(int kid, int reset) {
  unsigned long long t = 0;
  for (int i = 0; i < K_COUNT; ++i)
    if (kid < 0 || kid == i) {
      t += g_launches[i];
      if (reset) g_launches[i] = 0;
    }
  return t;
}
extern "C" int advk_prof_collect(int* kernel_ids, float* ms, int max_records);
extern "C" int advk_prof_configure(int kid, int capacity) {
  ADVK_REQUIRE(kid >= -2 && kid < K_COUNT && capacity >= 0, "bad kernel id / capacity");
  for (int i = 0; i < 2 * g_prof_cap; ++i) cudaEventDestroy(g_prof_ev[i]);
  delete[] g_prof_ev; delete[] g_prof_ids; delete[] g_prof_cnt;
  g_prof_ev = nullptr; g_prof_ids = nullptr; g_prof_cnt = nullptr; g_prof_cap = 0; g_prof_used = 0; g_prof_kid = -2;
  g_prof_open = -1;
  if (kid == -2 || capacity == 0) return ADVK_OK;
  g_prof_ev = new cudaEvent_t[2 * (size_t)capacity];
  g_prof_ids = new int[capacity];
  g_prof_cnt = new int[capacity];
  for (int i = 0; i < 2 * capacity; ++i)
    if (cudaEventCreate(&g_prof_ev[i]) != cudaSuccess) return check_launch("prof_configure");
  g_prof_cap = capacity;
  g_prof_kid = kid;
  return ADVK_OK;
}
extern "C" int advk_prof_group_runs(int enable) {
  const int prev = g_prof_group;
  g_prof_group = enable ? 1 : 0;
  g_prof_open = -1;
  return prev;
}
extern "C" int advk_prof_collect_runs(int* kernel_ids, float* ms, int* launches, int max_records) {
  int n = g_prof_used < max_records ? g_prof_used : max_records;
  if (launches)
    for (int i = 0; i < n; ++i) launches[i] = g_prof_cnt[i];
  return advk_prof_collect(kernel_ids, ms, max_records);
}
extern "C" int advk_prof_collect(int* kernel_ids, float* ms, int max_records) {
  int n = g_prof_used < max_records ? g_prof_used : max_records;
  g_prof_open = -1;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_prof_ev[2 * i + 1]) != cudaSuccess ||
        cudaEventElapsedTime(&ms[i], g_prof_ev[2 * i], g_prof_ev[2 * i + 1]) != cudaSuccess) {
      check_launch("prof_collect");
      return ADVK_ERR_CUDA;
    }
    kernel_ids[i] = g_prof_ids[i];
  }
  g_prof_used = 0;
  return n;
}

extern "C" int advk_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    set_error("advk_device_info: no CUDA device");
    (void)cudaGetLastError();
    return ADVK_ERR_CUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
  return ADVK_OK;
}

// 3-D scaling-and-squaring step rule (adv_morph.py:159-162) checked on the device: counts a
// violation when the smallest n >= 8 with sqrt(norm2)/2^n <= 0.5 differs from `nb_steps`.
__global__ void steps_check_kernel(const float* __restrict__ norm2, int nb_steps, int min_steps,
                                   int* __restrict__ violations) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  float nrm = sqrtf(norm2[0]);
  int n = min_steps;
  while (nrm / exp2f((float)n) > 0.5f && n < 64) ++n;
  if (n != nb_steps) atomicAdd(violations, 1);
}

extern "C" int advk_morph_steps_check(const float* norm2, int nb_steps, int min_steps, int* violations,
                                      void* stream) {
  ADVK_REQUIRE(norm2 && violations, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ADVK_LAUNCH(K_steps_check, st, launch_pdl((steps_check_kernel), 1, 1, 0, st, norm2, nb_steps, min_steps, violations));
  return check_launch("morph_steps_check");
}

// Early verdict of a graph-replayed PGD iteration: one thread bumps the replay's sequence number and stores
// (sequence << 8 | violations & 0xff) as ONE 32-bit word into pinned, device-mapped HOST memory, so that the host
// can poll the word while the rest of the iteration is still running instead of synchronising behind it.
__global__ void publish_kernel(const int* __restrict__ violations, unsigned* __restrict__ seq,
                               volatile unsigned* host_word, const float* __restrict__ norm2,
                               volatile float* host_norm2) {
  pdl_wait(); pdl_trigger();   // programmatic dependent launch: see launch_pdl (advk_common.cuh)
  const unsigned s = *seq + 1u;
  *seq = s;
  if (norm2 && host_norm2) {                // the norm the step rule was checked against, visible BEFORE the word
    *host_norm2 = *norm2;
    __threadfence_system();
  }
  *host_word = (s << 8) | ((unsigned)*violations & 0xffu);
  __threadfence_system();
}

extern "C" int advk_publish_verdict(const int* violations, unsigned* seq, unsigned* host_word, const float* norm2,
                                    float* host_norm2, void* stream) {
  ADVK_REQUIRE(violations && seq && host_word, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ADVK_LAUNCH(K_publish, st, launch_pdl((publish_kernel), 1, 1, 0, st, violations, seq, (volatile unsigned*)host_word,
                                        norm2, (volatile float*)host_norm2));
  return check_launch("publish_verdict");
}

extern "C" int advk_pgd_update(float* param, const float* grad, float step, int mode, int N,
                               size_t per_sample, double* sumsq, void* stream) {
  return advk_pgd_update_guarded(param, grad, step, mode, N, per_sample, sumsq, nullptr, stream);
}

extern "C" int advk_pgd_update_guarded(float* param, const float* grad, float step, int mode, int N,
                                       size_t per_sample, double* sumsq, const float* guard,
                                       void* stream) {
  ADVK_REQUIRE(param && grad && N >= 1 && per_sample >= 1, "null pointer / bad size");
  ADVK_REQUIRE(mode >= 0 && mode <= 3, "bad mode");
  cudaStream_t st = (cudaStream_t)stream;
  if (per_sample <= 8192) {
    ADVK_LAUNCH(K_update, st, launch_pdl((small_update_kernel), N, 256, 0, st, param, grad, step, mode, (int)per_sample, guard));
    return check_launch("pgd_update");
  }
  size_t b = (per_sample + 255) / 256;
  unsigned bx = (unsigned)(b > 1024 ? 1024 : b);
  dim3 grid(bx, N);
  if (mode == ADVK_UPD_L2_ASCENT || mode == ADVK_UPD_L2_POWER) {
    ADVK_REQUIRE(sumsq != nullptr, "sumsq scratch is NULL");
    cudaMemsetAsync(sumsq, 0, sizeof(double) * N, st);
    ADVK_LAUNCH(K_sumsq, st, launch_pdl((sumsq_kernel), grid, 256, 0, st, grad, per_sample, sumsq));
  }
  ADVK_LAUNCH(K_update, st, launch_pdl((update_kernel), grid, 256, 0, st, param, grad, step, mode, per_sample, sumsq, guard));
  return check_launch("pgd_update");
}

extern "C" int advk_clamp(const float* x, float lo, float hi, float* out, size_t n, void* stream) {
  ADVK_REQUIRE(x && out, "null pointer");
  if (n == 0) return ADVK_OK;
  ADVK_LAUNCH(K_clamp, (cudaStream_t)stream, launch_pdl((clamp_kernel), ew_blocks(n), 256, 0, (cudaStream_t)stream, x, lo, hi, out, n));
  return check_launch("clamp");
}
extern "C" int advk_clamp_bwd(const float* g_out, const float* x, float lo, float hi, float* g_x, size_t n,
                              void* stream) {
  ADVK_REQUIRE(x && g_out && g_x, "null pointer");
  if (n == 0) return ADVK_OK;
  ADVK_LAUNCH(K_clamp_bwd, (cudaStream_t)stream, launch_pdl((clamp_bwd_kernel), ew_blocks(n), 256, 0, (cudaStream_t)stream, g_out, x, lo, hi, g_x, n));
  return check_launch("clamp_bwd");
}
extern "C" int advk_nonzero_mask(float* x, size_t n, void* stream) {
  ADVK_REQUIRE(x != nullptr, "null pointer");
  if (n == 0) return ADVK_OK;
  ADVK_LAUNCH(K_nonzero, (cudaStream_t)stream, launch_pdl((nonzero_kernel), ew_blocks(n), 256, 0, (cudaStream_t)stream, x, n));
  return check_launch("nonzero_mask");
}
