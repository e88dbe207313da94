// Micro-benchmark: how fast can ONE block stream randomly placed 128-byte row parts of a
// 16.7 MB matrix (the heavy-id chain's feed), alone and with every other SM busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowgather rowgather.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// each warp: stages of 64 rows; lane -> (row j*4 + lane/8, chunk lane%8); `dep` = fetch the
// positions with a dependent load first (as the chain producers do)
template <int NLOAD>
__global__ void gather_rows(const float* __restrict__ g, const int* __restrict__ list, int c, int dim,
                            float* out, unsigned long long* tns, int nwarps_active) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  if (wib >= nwarps_active) return;
  const int ch = lane & 7, rsub = lane >> 3;
  float4 acc = make_float4(0, 0, 0, 0);
  const unsigned long long t0 = gtime();
  const int* l = list + (size_t)blockIdx.x * c;
  for (int k0 = wib * NLOAD * 4; k0 < c; k0 += nwarps_active * NLOAD * 4) {
    int pz[NLOAD * 4 / 32 > 0 ? NLOAD * 4 / 32 : 1];
#pragma unroll
    for (int q = 0; q < (NLOAD * 4 + 31) / 32; ++q) pz[q] = k0 + q * 32 + lane < c ? __ldg(l + k0 + q * 32 + lane) : -1;
    float4 v[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
      const int rr = j * 4 + rsub;
      const int p = __shfl_sync(0xffffffffu, pz[rr >> 5], rr & 31);
      v[j] = make_float4(0, 0, 0, 0);
      if (p >= 0) v[j] = __ldcs(reinterpret_cast<const float4*>(g + (long long)p * dim + ch * 4));
    }
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  __syncthreads();
  if (threadIdx.x == 0) tns[blockIdx.x] = gtime() - t0;
  (void)nw;
}

int main() {
  const int B = 65536, dim = 64, c = 7680;
  float* g; cudaMalloc(&g, (size_t)B * dim * 4 * 16);   // 16 batches: 268 MB
  cudaMemset(g, 0, (size_t)B * dim * 4 * 16);
  const int NB = 296;
  std::vector<int> h((size_t)NB * c);
  srand(1);
  for (auto& x : h) x = rand() % (B * 16);
  int* l; cudaMalloc(&l, h.size() * 4); cudaMemcpy(l, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, NB * 320 * 4);
  unsigned long long* tns; cudaMalloc(&tns, NB * 8);
  std::vector<unsigned long long> t(NB);
  char* flush; cudaMalloc(&flush, 512 << 20);
  for (int blocks : {1, 2, 74, 296}) {
    for (int warps : {1, 3, 9}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(flush, rep, 512 << 20);
        gather_rows<16><<<blocks, 320, 0>>>(g, l, c, dim, out, tns, warps);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(t.data(), tns, blocks * 8, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (int i = 0; i < blocks; ++i) mx = t[i] > mx ? t[i] : mx;
      printf("NLOAD16 blocks %3d warps %d: block0 %6.2f us, max %6.2f us -> %.2f ns/row (block0)\n", blocks, warps,
             t[0] * 1e-3, mx * 1e-3, (double)t[0] / c);
    }
  }
  for (int warps : {1, 9}) {
    cudaMemset(flush, 3, 512 << 20);
    gather_rows<8><<<1, 320, 0>>>(g, l, c, dim, out, tns, warps);
    cudaDeviceSynchronize();
    cudaMemcpy(t.data(), tns, 8, cudaMemcpyDeviceToHost);
    printf("NLOAD8 blocks 1 warps %d: %6.2f us -> %.2f ns/row\n", warps, t[0] * 1e-3, (double)t[0] / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
