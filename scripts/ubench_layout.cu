// ubench_layout.cu -- GPU-box micro-benchmark (not part of the product library): does a PLANAR field
// layout (x | y | z planes) beat the interleaved float4 (x,y,z,0) layout for the squaring-step kernels?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/_ubench_layout scripts/ubench_layout.cu
//   scripts/_ubench_layout [D H W] [amplitude in voxels]
//
// Both squaring-step kernels run at one 32-byte sector request per SM per clock (DESIGN.md 3.1,
// profiles/r01l_ssb_taken_apart.md).  A quarter of every float4 sector is the padding float; planar
// fields move 25 % fewer sectors for three times the memory instructions.  This program times, on a
// synthetic smooth field, the forward step phi_k = phi_{k-1} o phi_{k-1} (border padding,
// align_corners=True: adv_morph.py:133-135 of the reference) and the scatter half of its adjoint in
// both layouts with CUDA events, and checks that the layouts agree.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Ax { int i0; float w0, w1; int step; };
__device__ __forceinline__ Ax axis(float c, int size, int stride) {
  Ax a;
  const float mx = (float)(size - 1);
  float x = ((c + 1.f) / 2.f) * mx;
  x = fminf(mx, fmaxf(x, 0.f));
  const float f = floorf(x);
  a.i0 = (int)f; a.w0 = (f + 1.f) - x; a.w1 = x - f;
  a.step = (a.i0 + 1 < size) ? stride : 0;      // outside corner: weight 0, redirected
  return a;
}

// ---- forward step ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8) fwd_f4(int D, int H, int W, const float4* __restrict__ in, float4* __restrict__ out) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= D * H * W) return;
  const float4 f = __ldg(in + p);
  const Ax ax = axis(f.x, W, 1), ay = axis(f.y, H, W), az = axis(f.z, D, H * W);
  const float4* c = in + (az.i0 * H * W + ay.i0 * W + ax.i0);
  float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 s = __ldg(c + ((k & 1) * ax.step + ((k >> 1) & 1) * ay.step + (k >> 2) * az.step));
    const float w = ((k & 1) ? ax.w1 : ax.w0) * (((k >> 1) & 1) ? ay.w1 : ay.w0) * ((k >> 2) ? az.w1 : az.w0);
    ox += s.x * w; oy += s.y * w; oz += s.z * w;
  }
  out[p] = make_float4(ox, oy, oz, 0.f);
}

__global__ void __launch_bounds__(256, 8) fwd_planar(int D, int H, int W, const float* __restrict__ in, float* __restrict__ out) {
  const int S = D * H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= S) return;
  const float* px = in; const float* py = in + S; const float* pz = in + 2 * (size_t)S;
  const Ax ax = axis(__ldg(px + p), W, 1), ay = axis(__ldg(py + p), H, W), az = axis(__ldg(pz + p), D, H * W);
  const int c = az.i0 * H * W + ay.i0 * W + ax.i0;
  float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int o = c + (k & 1) * ax.step + ((k >> 1) & 1) * ay.step + (k >> 2) * az.step;
    const float w = ((k & 1) ? ax.w1 : ax.w0) * (((k >> 1) & 1) ? ay.w1 : ay.w0) * ((k >> 2) ? az.w1 : az.w0);
    ox += __ldg(px + o) * w; oy += __ldg(py + o) * w; oz += __ldg(pz + o) * w;
  }
  out[p] = ox; out[p + S] = oy; out[p + 2 * (size_t)S] = oz;
}

// ---- scatter half of the adjoint: out(corner) += g(p) * w, x hand-off between neighbouring lanes ----
__global__ void __launch_bounds__(256) scat_f4(int D, int H, int W, const float4* __restrict__ phi, const float4* __restrict__ g, float4* out) {
  const int S = D * H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = p < S;
  float4 f = make_float4(0, 0, 0, 0), go = f;
  if (live) { f = __ldg(phi + p); go = __ldg(g + p); }
  const Ax ax = axis(f.x, W, 1), ay = axis(f.y, H, W), az = axis(f.z, D, H * W);
  const int a000 = az.i0 * H * W + ay.i0 * W + ax.i0;
  const int next = __shfl_down_sync(0xffffffffu, live ? a000 : -1, 1);
  const bool hand = live && ax.step && lane < 31 && next == a000 + 1;
  const float qx = __shfl_up_sync(0xffffffffu, go.x, 1), qy = __shfl_up_sync(0xffffffffu, go.y, 1), qz = __shfl_up_sync(0xffffffffu, go.z, 1);
  float w0[4], w1[4], ws[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float wr = ((r & 1) ? ay.w1 : ay.w0) * ((r >> 1) ? az.w1 : az.w0);
    w0[r] = ax.w0 * wr; w1[r] = ax.w1 * wr;
    const float got = __shfl_up_sync(0xffffffffu, hand ? w1[r] : 0.f, 1);
    ws[r] = lane ? got : 0.f;
  }
  if (!live) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float4* d = out + (a000 + (r & 1) * ay.step + (r >> 1) * az.step);
    atomicAdd(d, make_float4(go.x * w0[r] + qx * ws[r], go.y * w0[r] + qy * ws[r], go.z * w0[r] + qz * ws[r], 0.f));
    if (ax.step && !hand) atomicAdd(d + 1, make_float4(go.x * w1[r], go.y * w1[r], go.z * w1[r], 0.f));
  }
}

__global__ void __launch_bounds__(256) scat_planar(int D, int H, int W, const float* __restrict__ phi, const float* __restrict__ g, float* out) {
  const int S = D * H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = p < S;
  float fx = 0, fy = 0, fz = 0, gx = 0, gy = 0, gz = 0;
  if (live) {
    fx = __ldg(phi + p); fy = __ldg(phi + S + p); fz = __ldg(phi + 2 * (size_t)S + p);
    gx = __ldg(g + p); gy = __ldg(g + S + p); gz = __ldg(g + 2 * (size_t)S + p);
  }
  const Ax ax = axis(fx, W, 1), ay = axis(fy, H, W), az = axis(fz, D, H * W);
  const int a000 = az.i0 * H * W + ay.i0 * W + ax.i0;
  const int next = __shfl_down_sync(0xffffffffu, live ? a000 : -1, 1);
  const bool hand = live && ax.step && lane < 31 && next == a000 + 1;
  const float qx = __shfl_up_sync(0xffffffffu, gx, 1), qy = __shfl_up_sync(0xffffffffu, gy, 1), qz = __shfl_up_sync(0xffffffffu, gz, 1);
  float w0[4], w1[4], ws[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float wr = ((r & 1) ? ay.w1 : ay.w0) * ((r >> 1) ? az.w1 : az.w0);
    w0[r] = ax.w0 * wr; w1[r] = ax.w1 * wr;
    const float got = __shfl_up_sync(0xffffffffu, hand ? w1[r] : 0.f, 1);
    ws[r] = lane ? got : 0.f;
  }
  if (!live) return;
  float* ox = out; float* oy = out + S; float* oz = out + 2 * (size_t)S;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = a000 + (r & 1) * ay.step + (r >> 1) * az.step;
    atomicAdd(ox + o, gx * w0[r] + qx * ws[r]);
    atomicAdd(oy + o, gy * w0[r] + qy * ws[r]);
    atomicAdd(oz + o, gz * w0[r] + qz * ws[r]);
    if (ax.step && !hand) {
      atomicAdd(ox + o + 1, gx * w1[r]); atomicAdd(oy + o + 1, gy * w1[r]); atomicAdd(oz + o + 1, gz * w1[r]);
    }
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
  const int D = argc > 3 ? atoi(argv[1]) : 128, H = argc > 3 ? atoi(argv[2]) : 128, W = argc > 3 ? atoi(argv[3]) : 128;
  const float amp = argc > 4 ? (float)atof(argv[4]) : (argc == 2 ? (float)atof(argv[1]) : 0.6f);
  const size_t S = (size_t)D * H * W;
  std::vector<float4> h4(S), hg4(S);
  std::vector<float> hp(3 * S), hgp(3 * S);
  for (int z = 0; z < D; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
    const size_t p = ((size_t)z * H + y) * W + x;
    const float bx = -1.f + 2.f * x / (W - 1), by = -1.f + 2.f * y / (H - 1), bz = -1.f + 2.f * z / (D - 1);
    const float fx = bx + amp * 2.f / (W - 1) * sinf(3 * bx + 2 * by + bz);
    const float fy = by + amp * 2.f / (H - 1) * cosf(2 * bx - by + 3 * bz);
    const float fz = bz + amp * 2.f / (D - 1) * sinf(bx + 4 * by - 2 * bz);
    h4[p] = make_float4(fx, fy, fz, 0.f);
    hp[p] = fx; hp[S + p] = fy; hp[2 * S + p] = fz;
    const float g0 = sinf(0.37f * p), g1 = cosf(0.11f * p), g2 = sinf(0.05f * p + 1.f);
    hg4[p] = make_float4(g0, g1, g2, 0.f);
    hgp[p] = g0; hgp[S + p] = g1; hgp[2 * S + p] = g2;
  }
  float4 *d4, *o4, *g4; float *dp, *op, *gp;
  CK(cudaMalloc(&d4, S * 16)); CK(cudaMalloc(&o4, S * 16)); CK(cudaMalloc(&g4, S * 16));
  CK(cudaMalloc(&dp, S * 12)); CK(cudaMalloc(&op, S * 12)); CK(cudaMalloc(&gp, S * 12));
  CK(cudaMemcpy(d4, h4.data(), S * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(g4, hg4.data(), S * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, hp.data(), S * 12, cudaMemcpyHostToDevice)); CK(cudaMemcpy(gp, hgp.data(), S * 12, cudaMemcpyHostToDevice));
  const int grid = (int)((S + 255) / 256), reps = 20;
  const float t1 = time_us([&] { fwd_f4<<<grid, 256>>>(D, H, W, d4, o4); }, reps);
  const float t2 = time_us([&] { fwd_planar<<<grid, 256>>>(D, H, W, dp, op); }, reps);
  std::vector<float4> r4(S); std::vector<float> rp(3 * S);
  CK(cudaMemcpy(r4.data(), o4, S * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rp.data(), op, S * 12, cudaMemcpyDeviceToHost));
  double e1 = 0;
  for (size_t p = 0; p < S; ++p) e1 = fmax(e1, fmax(fabs(r4[p].x - rp[p]), fmax(fabs(r4[p].y - rp[S + p]), fabs(r4[p].z - rp[2 * S + p]))));
  // scatter: the target is NOT cleared between repetitions (timing only needs the traffic); one clean pass for the check
  const float t3 = time_us([&] { scat_f4<<<grid, 256>>>(D, H, W, d4, g4, o4); }, reps);
  const float t4 = time_us([&] { scat_planar<<<grid, 256>>>(D, H, W, dp, gp, op); }, reps);
  CK(cudaMemset(o4, 0, S * 16)); CK(cudaMemset(op, 0, S * 12));
  scat_f4<<<grid, 256>>>(D, H, W, d4, g4, o4); scat_planar<<<grid, 256>>>(D, H, W, dp, gp, op);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(r4.data(), o4, S * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rp.data(), op, S * 12, cudaMemcpyDeviceToHost));
  double e2 = 0, m2 = 0;
  for (size_t p = 0; p < S; ++p) {
    e2 = fmax(e2, fmax(fabs(r4[p].x - rp[p]), fmax(fabs(r4[p].y - rp[S + p]), fabs(r4[p].z - rp[2 * S + p]))));
    m2 = fmax(m2, fabs(r4[p].x));
  }
  printf("%dx%dx%d, displacement amplitude %.2f voxels, %d repetitions\n", D, H, W, amp, reps);
  printf("forward step   float4 %.1f us   planar %.1f us   (max |diff| %.2e)\n", t1, t2, e1);
  printf("adjoint scatter float4 %.1f us   planar %.1f us   (max |diff| %.2e of %.2e)\n", t3, t4, e2, m2);
  return 0;
}
