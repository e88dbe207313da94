// Micro-benchmark 2: per-warp throughput of random 128-byte row-part loads, by load flavour,
// loads in flight, and working-set size (DRAM vs L2).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
template <int MODE>
__device__ __forceinline__ float4 ld(const float4* p) {
  float4 v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = __ldcs(p);
  else if (MODE == 2) v = __ldcg(p);
  else asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// rows per lane-group: MODE loads; LAY 0: 8 lanes per 128-B row part (4 rows per instr); LAY 1: 32 lanes x 16 B = one 512-B row (1 row/instr)
template <int NLOAD, int MODE, int LAY>
__global__ void gather_rows(const float* __restrict__ g, const int* __restrict__ list, int c, int dim,
                            float* out, unsigned long long* tns, int nwarps_active) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wib >= nwarps_active) return;
  const int ch = LAY ? lane : (lane & 7), rsub = LAY ? 0 : (lane >> 3);
  constexpr int RPI = LAY ? 1 : 4;   // rows per instruction
  float4 acc = make_float4(0, 0, 0, 0);
  const unsigned long long t0 = gtime();
  const int* l = list + (size_t)blockIdx.x * c;
  for (int k0 = wib * NLOAD * RPI; k0 < c; k0 += nwarps_active * NLOAD * RPI) {
    int pz[(NLOAD * RPI + 31) / 32];
#pragma unroll
    for (int q = 0; q < (NLOAD * RPI + 31) / 32; ++q) pz[q] = k0 + q * 32 + lane < c ? __ldg(l + k0 + q * 32 + lane) : -1;
    float4 v[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
      const int rr = j * RPI + rsub;
      const int p = __shfl_sync(0xffffffffu, pz[rr >> 5], rr & 31);
      v[j] = make_float4(0, 0, 0, 0);
      if (p >= 0) v[j] = ld<MODE>(reinterpret_cast<const float4*>(g + (long long)p * dim + ch * 4));
    }
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  __syncthreads();
  if (threadIdx.x == 0) tns[blockIdx.x] = gtime() - t0;
}

float* g; int* l; float* out; unsigned long long* tns; char* flush;
template <int NLOAD, int MODE, int LAY>
void run(const char* name, int warps, int dim, int c, int blocks = 1, int threads = 320) {
  unsigned long long t[1024];
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(flush, rep, 512 << 20);
    gather_rows<NLOAD, MODE, LAY><<<blocks, threads>>>(g, l, c, dim, out, tns, warps);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(t, tns, 8 * blocks, cudaMemcpyDeviceToHost);
  printf("%-28s nload %2d warps %2d blocks %3d: %7.2f us  %6.2f ns/row  (%.1f GB/s per SM)\n", name, NLOAD, warps, blocks, t[0] * 1e-3,
         (double)t[0] / c, (LAY ? 512.0 : 128.0) * c / t[0]);
}

int main() {
  const int c = 7680;
  const size_t rows = 65536 * 16;
  cudaMalloc(&g, rows * 128 * 4);   // dim up to 128
  cudaMemset(g, 0, rows * 128 * 4);
  std::vector<int> h((size_t)296 * c);
  srand(1);
  for (auto& x : h) x = rand() % rows;
  cudaMalloc(&l, h.size() * 4); cudaMemcpy(l, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 296 * 1024 * 4); cudaMalloc(&tns, 296 * 8); cudaMalloc(&flush, 512 << 20);
  run<4, 0, 0>("ldg", 1, 64, c); run<8, 0, 0>("ldg", 1, 64, c); run<16, 0, 0>("ldg", 1, 64, c); run<32, 0, 0>("ldg", 1, 64, c);
  run<16, 1, 0>("ldcs", 1, 64, c); run<16, 2, 0>("ldcg", 1, 64, c); run<16, 3, 0>("no_allocate", 1, 64, c);
  run<8, 0, 1>("ldg 512B rows (dim128)", 1, 128, c); run<16, 0, 1>("ldg 512B rows (dim128)", 1, 128, c);
  run<16, 0, 0>("ldg", 9, 64, c); run<16, 0, 0>("ldg 32 warps", 32, 64, c, 1, 1024); run<8, 0, 0>("ldg 32 warps", 32, 64, c, 1, 1024);
  run<4, 0, 0>("ldg 32 warps", 32, 64, c, 1, 1024);
  run<16, 0, 0>("ldg 32 warps all SMs", 32, 64, c, 148, 1024);
  // L2-resident working set: indices within the first 8 MB
  for (auto& x : h) x = rand() % 32768;
  cudaMemcpy(l, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  printf("-- L2-resident (8 MB) --\n");
  unsigned long long t[1];
  for (int rep = 0; rep < 3; ++rep) { gather_rows<16, 0, 0><<<1, 320>>>(g, l, c, 64, out, tns, 1); cudaDeviceSynchronize(); }
  cudaMemcpy(t, tns, 8, cudaMemcpyDeviceToHost);
  printf("ldg nload 16 warps 1 (L2): %7.2f us %6.2f ns/row\n", t[0] * 1e-3, (double)t[0] / c);
  for (int rep = 0; rep < 3; ++rep) { gather_rows<16, 0, 0><<<1, 320>>>(g, l, c, 64, out, tns, 9); cudaDeviceSynchronize(); }
  cudaMemcpy(t, tns, 8, cudaMemcpyDeviceToHost);
  printf("ldg nload 16 warps 9 (L2): %7.2f us %6.2f ns/row\n", t[0] * 1e-3, (double)t[0] / c);
  for (int rep = 0; rep < 3; ++rep) { gather_rows<16, 0, 0><<<1, 1024>>>(g, l, c, 64, out, tns, 32); cudaDeviceSynchronize(); }
  cudaMemcpy(t, tns, 8, cudaMemcpyDeviceToHost);
  printf("ldg nload 16 warps 32 (L2): %7.2f us %6.2f ns/row\n", t[0] * 1e-3, (double)t[0] / c);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
