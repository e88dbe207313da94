// Micro-benchmark: one warp chain-adding rows out of a shared-memory ring that ONE thread keeps
// full with 8 KB cp.async.bulk copies from a contiguous (L2-resident or DRAM) buffer.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../tfplus_b200/csrc/async_copy.cuh"
using namespace kvhbm;
constexpr int STAGES = 10, STAGE_BYTES = 8192;
__global__ void chain_tma(const float* __restrict__ src, int nst, float* out, long long* cyc, int copies_per_stage) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned long long full[STAGES], empty[STAGES];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_async_smem(); }
  __syncthreads();
  const long long t0 = clock64();
  if (wib == 0) {
    float acc = 0.f; long long waited = 0;
    for (int s = 0; s < nst; ++s) {
      const int e = s % STAGES; const unsigned par = (s / STAGES) & 1u;
      const long long w0 = clock64();
      mbar_wait(&full[e], par);
      waited += clock64() - w0;
      const float4* st = reinterpret_cast<const float4*>(ring + (size_t)e * STAGE_BYTES) + lane;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = st[j * 32];
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc += v[j].x; acc += v[j].y; acc += v[j].z; acc += v[j].w; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[e]);
    }
    out[lane] = acc;
    if (lane == 0) { cyc[0] = clock64() - t0; cyc[1] = waited; }
  } else if (wib == 1 && lane == 0) {
    for (int s = 0; s < nst; ++s) {
      const int e = s % STAGES; const int use = s / STAGES;
      if (use > 0) mbar_wait(&empty[e], (use - 1) & 1);
      mbar_arrive_expect_tx(&full[e], STAGE_BYTES);
      const int cb = STAGE_BYTES / copies_per_stage;
      for (int q = 0; q < copies_per_stage; ++q)
        bulk_load(ring + (size_t)e * STAGE_BYTES + q * cb, reinterpret_cast<const unsigned char*>(src) + (size_t)s * STAGE_BYTES + q * cb, cb, &full[e]);
    }
  }
}
// the production consumer: 16-row chunks, loads a chunk ahead, barrier probed under the adds
__global__ void chain_tma_pipe(const float* __restrict__ src, int nst, float* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned long long full[STAGES], empty[STAGES];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_async_smem(); }
  __syncthreads();
  const long long t0 = clock64();
  if (wib == 0) {
    float acc = 0.f;
    unsigned e = 0, par = 0;
    float4 a[4], b[4];
#define KV_LOAD(dst, c0) _Pragma("unroll") for (int j = 0; j < 4; ++j) dst[j] = st[((c0) * 4 + j) * 32]
#define KV_ADD(v) _Pragma("unroll") for (int j = 0; j < 4; ++j) { acc += v[j].x; acc += v[j].y; acc += v[j].z; acc += v[j].w; }
    mbar_wait(&full[e], par);
    const float4* st = reinterpret_cast<const float4*>(ring + (size_t)e * STAGE_BYTES) + lane;
    KV_LOAD(a, 0);
#pragma unroll 1
    for (int s = 0; s < nst; ++s) {
      KV_LOAD(b, 1); KV_ADD(a); KV_LOAD(a, 2); KV_ADD(b); KV_LOAD(b, 3); KV_ADD(a);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[e]);
      if (++e == STAGES) { e = 0; par ^= 1u; }
      if (s + 1 < nst) {
        unsigned ok = mbar_test(&full[e], par);
        KV_ADD(b);
        while (!ok) ok = mbar_test(&full[e], par);
        st = reinterpret_cast<const float4*>(ring + (size_t)e * STAGE_BYTES) + lane;
        KV_LOAD(a, 0);
      } else { KV_ADD(b); }
    }
    out[lane] = acc;
    if (lane == 0) { cyc[0] = clock64() - t0; cyc[1] = 0; }
  } else if (wib == 1 && lane == 0) {
    for (int s = 0; s < nst; ++s) {
      const int e = s % STAGES; const int use = s / STAGES;
      if (use > 0) mbar_wait(&empty[e], (use - 1) & 1);
      mbar_arrive_expect_tx(&full[e], STAGE_BYTES);
      bulk_load(ring + (size_t)e * STAGE_BYTES, reinterpret_cast<const unsigned char*>(src) + (size_t)s * STAGE_BYTES, STAGE_BYTES, &full[e]);
    }
  }
}
int main() {
  const int nst = 120;
  float* src; cudaMalloc(&src, 64 << 20); cudaMemset(src, 0, 64 << 20);
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 16);
  char* flush; cudaMalloc(&flush, 512 << 20);
  cudaFuncSetAttribute(chain_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES);
  long long h[2];
  for (int cps : {1, 2, 8}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 3; ++rep) {
        if (mode == 1) cudaMemset(flush, rep, 512 << 20);   // source in DRAM
        chain_tma<<<1, 64, STAGES * STAGE_BYTES>>>(src + (mode ? (size_t)rep * (8 << 20) / 4 : 0), nst, out, cyc, cps);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("%s, %d copies/stage: %lld cycles for %d rows = %.2f cyc/row (waited %lld)\n", mode ? "DRAM" : "L2  ", cps, h[0], nst * 64,
             (double)h[0] / (nst * 64), h[1]);
    }
  }
  cudaFuncSetAttribute(chain_tma_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES);
  for (int threads : {64, 640}) {
    for (int rep = 0; rep < 3; ++rep) { chain_tma_pipe<<<1, threads, STAGES * STAGE_BYTES>>>(src, nst, out, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("pipelined consumer, %d threads: %lld cycles for %d rows = %.2f cyc/row\n", threads, h[0], nst * 64, (double)h[0] / (nst * 64));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
