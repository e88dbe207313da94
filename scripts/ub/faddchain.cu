// Micro-benchmark: how fast does ONE warp walk a chain of dependent FADDs (a) over registers,
// (b) over shared memory with 128-bit loads issued ahead.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_regs(float* out, int n, long long* cyc, float x) {
  float acc = out[threadIdx.x];
  const long long t0 = clock64();
  for (int i = 0; i < n; i += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += x;
    x = -x;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_smem(float* out, int n, long long* cyc) {
  extern __shared__ float4 sm[];   // [units][32 lanes]
  for (int i = threadIdx.x; i < 16 * 32 * 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (i & 1) ? 1.f : -1.f;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  float acc = 0.f;
  const float4* st = sm + threadIdx.x;
  const long long t0 = clock64();
  float4 a[4], b[4];
#define LOAD(d, c0) _Pragma("unroll") for (int j = 0; j < 4; ++j) d[j] = st[((c0) * 4 + j) * 32]
#define ADD(v) _Pragma("unroll") for (int j = 0; j < 4; ++j) { acc += v[j].x; acc += v[j].y; acc += v[j].z; acc += v[j].w; }
  LOAD(a, 0);
  for (int i = 0; i < n; i += 64) {
    LOAD(b, 1); ADD(a); LOAD(a, 2); ADD(b); LOAD(b, 3); ADD(a); ADD(b); LOAD(a, 0);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc + a[0].x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMemset(out, 0, 4096);
  long long h;
  const int n = 7680;
  for (int rep = 0; rep < 2; ++rep) { chain_regs<<<1, 32>>>(out, n, cyc, 1.0f); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("register chain: %lld cycles for %d adds = %.2f cyc/add\n", h, n, (double)h / n);
  for (int rep = 0; rep < 2; ++rep) { chain_smem<<<1, 64, 16 * 512>>>(out, n, cyc); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("smem chain (LDS.128 ahead): %lld cycles for %d adds = %.2f cyc/add\n", h, n, (double)h / n);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
