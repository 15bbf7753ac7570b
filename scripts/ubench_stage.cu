// ubench_stage.cu -- GPU-box micro-benchmark (not part of the product library): is it worth staging the
// phi tile of a squaring step in shared memory?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/_ubench_stage scripts/ubench_stage.cu
//   scripts/_ubench_stage [amplitude in voxels]
//
// The forward step phi_k = phi_{k-1} o phi_{k-1} (border padding, align_corners=True; adv_morph.py:133-135,
// 166-168 of the reference) runs at one 32-byte sector request per SM per clock, and 1.55 of its 2.55 sectors
// per voxel are gather fills: L1 only captures the reuse inside a CTA of 256 consecutive voxels (3 rows x 2
// planes).  While the displacement is below one voxel (phi_0 ... roughly phi_5 of 8 levels: it doubles per
// level and ends at a few voxels), every corner of a voxel lies within +-1 of it, so a CTA can stage its tile
// + a halo of 1 once (32x8x8 outputs: 34x10x10 float4 = 54 KB, 1.66x the tile = 0.83 sectors per voxel) and
// gather from shared memory.  Whether a CTA may do so only depends on the phi values of its own tile, which
// it has just loaded: a CTA-wide vote picks the staged or the global gathers (exact for any field).
// Kernels: direct (the product's lean forward step) vs staged, same arithmetic, results compared bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Ax { int i0; float w0, w1; int has1; };
__device__ __forceinline__ Ax axis(float c, int size) {
  Ax a;
  const float mx = (float)(size - 1);
  float x = ((c + 1.f) / 2.f) * mx;
  x = fminf(mx, fmaxf(x, 0.f));
  const float f = floorf(x);
  a.i0 = (int)f; a.w0 = (f + 1.f) - x; a.w1 = x - f;
  a.has1 = (a.i0 + 1 < size) ? 1 : 0;           // outside corner: weight 0, redirected to corner 0
  return a;
}

__global__ void __launch_bounds__(256, 8) fwd_direct(int D, int H, int W, const float4* __restrict__ in, float4* __restrict__ out) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= D * H * W) return;
  const float4 f = __ldg(in + p);
  const Ax ax = axis(f.x, W), ay = axis(f.y, H), az = axis(f.z, D);
  const float4* c = in + (az.i0 * H * W + ay.i0 * W + ax.i0);
  const int sx = ax.has1, sy = ay.has1 * W, sz = az.has1 * H * W;
  float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 s = __ldg(c + ((k & 1) * sx + ((k >> 1) & 1) * sy + (k >> 2) * sz));
    const float w = ((k & 1) ? ax.w1 : ax.w0) * (((k >> 1) & 1) ? ay.w1 : ay.w0) * ((k >> 2) ? az.w1 : az.w0);
    ox += s.x * w; oy += s.y * w; oz += s.z * w;
  }
  out[p] = make_float4(ox, oy, oz, 0.f);
}

constexpr int TX = 32, TY = 8, TZ = 8;                 // outputs per CTA; 256 threads, 8 voxels (z) each
constexpr int RX = TX + 2, RY = TY + 2, RZ = TZ + 2;   // staged region (halo 1)

__global__ void __launch_bounds__(256) fwd_staged(int D, int H, int W, const float4* __restrict__ in, float4* __restrict__ out,
                                                  unsigned* fallbacks) {
  extern __shared__ float4 sm[];
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = blockIdx.z * TZ;
  const int HW = H * W;
  // stage the region; coordinates outside the volume are clamped (those cells are never sampled: sampling
  // coordinates are clipped to the volume)
  for (int r = threadIdx.x; r < RX * RY * RZ; r += 256) {
    const int rx = r % RX, ry = (r / RX) % RY, rz = r / (RX * RY);
    const int gx = min(max(x0 - 1 + rx, 0), W - 1), gy = min(max(y0 - 1 + ry, 0), H - 1), gz = min(max(z0 - 1 + rz, 0), D - 1);
    sm[r] = __ldg(in + (gz * HW + gy * W + gx));
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int x = x0 + lx, y = y0 + ly;
  const bool col = x < W && y < H;
  // pass 1: this thread's 8 voxels -- do all their corners lie inside the staged region?
  bool ok = true;
#pragma unroll
  for (int lz = 0; lz < TZ; ++lz) {
    const int z = z0 + lz;
    if (col && z < D) {
      const float4 f = sm[((lz + 1) * RY + (ly + 1)) * RX + (lx + 1)];
      const Ax ax = axis(f.x, W), ay = axis(f.y, H), az = axis(f.z, D);
      const int cx = ax.i0 - (x0 - 1), cy = ay.i0 - (y0 - 1), cz = az.i0 - (z0 - 1);
      ok = ok && cx >= 0 && cx + ax.has1 < RX && cy >= 0 && cy + ay.has1 < RY && cz >= 0 && cz + az.has1 < RZ;
    }
  }
  const bool staged = __syncthreads_and(ok ? 1 : 0) != 0;
  if (!staged && threadIdx.x == 0 && fallbacks) atomicAdd(fallbacks, 1u);
#pragma unroll 2
  for (int lz = 0; lz < TZ; ++lz) {
    const int z = z0 + lz;
    if (!(col && z < D)) continue;
    const float4 f = sm[((lz + 1) * RY + (ly + 1)) * RX + (lx + 1)];
    const Ax ax = axis(f.x, W), ay = axis(f.y, H), az = axis(f.z, D);
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (staged) {
      const float4* c = sm + (((az.i0 - (z0 - 1)) * RY + (ay.i0 - (y0 - 1))) * RX + (ax.i0 - (x0 - 1)));
      const int sx = ax.has1, sy = ay.has1 * RX, sz = az.has1 * RX * RY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 s = c[(k & 1) * sx + ((k >> 1) & 1) * sy + (k >> 2) * sz];
        const float w = ((k & 1) ? ax.w1 : ax.w0) * (((k >> 1) & 1) ? ay.w1 : ay.w0) * ((k >> 2) ? az.w1 : az.w0);
        ox += s.x * w; oy += s.y * w; oz += s.z * w;
      }
    } else {
      const float4* c = in + (az.i0 * HW + ay.i0 * W + ax.i0);
      const int sx = ax.has1, sy = ay.has1 * W, sz = az.has1 * HW;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 s = __ldg(c + ((k & 1) * sx + ((k >> 1) & 1) * sy + (k >> 2) * sz));
        const float w = ((k & 1) ? ax.w1 : ax.w0) * (((k >> 1) & 1) ? ay.w1 : ay.w0) * ((k >> 2) ? az.w1 : az.w0);
        ox += s.x * w; oy += s.y * w; oz += s.z * w;
      }
    }
    out[z * HW + y * W + x] = make_float4(ox, oy, oz, 0.f);
  }
}

template <typename F>
static float time_us(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return 1e3f * ms / reps;
}

int main(int argc, char** argv) {
  const int D = 128, H = 128, W = 128;
  const size_t S = (size_t)D * H * W;
  const size_t smem = sizeof(float4) * RX * RY * RZ;
  CK(cudaFuncSetAttribute(fwd_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float4 *d4, *oa, *ob; unsigned* fb;
  CK(cudaMalloc(&d4, S * 16)); CK(cudaMalloc(&oa, S * 16)); CK(cudaMalloc(&ob, S * 16)); CK(cudaMalloc(&fb, 4));
  std::vector<float4> h4(S), ra(S), rb(S);
  const float amps_default[] = {0.05f, 0.6f, 1.5f, 4.0f};
  std::vector<float> amps(amps_default, amps_default + 4);
  if (argc > 1) { amps.clear(); amps.push_back((float)atof(argv[1])); }
  for (float amp : amps) {
    for (int z = 0; z < D; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
      const size_t p = ((size_t)z * H + y) * W + x;
      const float bx = -1.f + 2.f * x / (W - 1), by = -1.f + 2.f * y / (H - 1), bz = -1.f + 2.f * z / (D - 1);
      h4[p] = make_float4(bx + amp * 2.f / (W - 1) * sinf(3 * bx + 2 * by + bz), by + amp * 2.f / (H - 1) * cosf(2 * bx - by + 3 * bz),
                          bz + amp * 2.f / (D - 1) * sinf(bx + 4 * by - 2 * bz), 0.f);
    }
    CK(cudaMemcpy(d4, h4.data(), S * 16, cudaMemcpyHostToDevice));
    const int grid1 = (int)((S + 255) / 256), reps = 20;
    const dim3 grid2((W + TX - 1) / TX, (H + TY - 1) / TY, (D + TZ - 1) / TZ);
    const float t1 = time_us([&] { fwd_direct<<<grid1, 256>>>(D, H, W, d4, oa); }, reps);
    const float t2 = time_us([&] { fwd_staged<<<grid2, 256, smem>>>(D, H, W, d4, ob, nullptr); }, reps);
    CK(cudaMemset(fb, 0, 4));
    fwd_staged<<<grid2, 256, smem>>>(D, H, W, d4, ob, fb);
    CK(cudaDeviceSynchronize());
    unsigned nfb = 0;
    CK(cudaMemcpy(&nfb, fb, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ra.data(), oa, S * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rb.data(), ob, S * 16, cudaMemcpyDeviceToHost));
    double e = 0;
    for (size_t p = 0; p < S; ++p) e = fmax(e, fmax(fabs(ra[p].x - rb[p].x), fmax(fabs(ra[p].y - rb[p].y), fabs(ra[p].z - rb[p].z))));
    printf("128^3, amplitude %.2f voxels: direct %.1f us   staged 32x8x8 %.1f us   (%u of %u CTAs fell back; max |diff| %.2e)\n",
           amp, t1, t2, nfb, grid2.x * grid2.y * grid2.z, e);
  }
  return 0;
}
