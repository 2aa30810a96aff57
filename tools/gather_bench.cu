// Random-sector gather microbenchmark: how many independent G-byte reads per second does the
// B200 memory system sustain from a table much larger than L2?  Sets the realistic ceiling
// for the probe kernel (one bucket = one G-byte read) next to the streaming-copy peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t rng(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}

// G = bytes per gather (32, 64, 128); U = independent gathers in flight per thread
template <int G, int U>
__global__ void __launch_bounds__(256) k_gather(const char *tab, uint64_t nslots, uint64_t per_thread, unsigned long long *sink) {
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (uint64_t i = 0; i < per_thread; i += U) {
    uint32_t w[U][G / 4 > 8 ? 8 : G / 4];
#pragma unroll
    for (int u = 0; u < U; u++) {
      uint64_t slot = rng(tid * 0x100000001b3ull + i + u) % nslots;
      const char *p = tab + slot * G;
      asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(p));
      if (G >= 64) { uint32_t t[8];
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(p + 32));
        w[u][0] ^= t[0] ^ t[7]; }
      if (G >= 128) { uint32_t t[8];
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(p + 64));
        w[u][1] ^= t[0] ^ t[7];
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(p + 96));
        w[u][2] ^= t[0] ^ t[7]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += w[u][0] ^ w[u][1] ^ w[u][2] ^ w[u][7];
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int G, int U>
static void run(const char *tab, uint64_t bytes, int ctas_per_sm, int sms, unsigned long long *sink) {
  uint64_t nslots = bytes / G;
  int grid = sms * ctas_per_sm;
  uint64_t per_thread = 4096 / U * U;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_gather<G, U><<<grid, 256>>>(tab, nslots, 256, sink);   // warm-up
  cudaEventRecord(a);
  k_gather<G, U><<<grid, 256>>>(tab, nslots, per_thread, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double n = (double)grid * 256 * per_thread;
  printf("G=%3d B  U=%2d  ctas/sm=%d  threads/sm=%4d  %8.2f Ggather/s  %8.1f GB/s useful\n", G, U, ctas_per_sm,
         ctas_per_sm * 256, n / ms / 1e6, n * G / ms / 1e6);
}

int main(int argc, char **argv) {
  uint64_t gb = argc > 1 ? atoll(argv[1]) : 32;
  uint64_t bytes = gb << 30;
  char *tab; unsigned long long *sink;
  if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  cudaMemset(tab, 1, bytes);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, table %llu GiB\n", p.name, sms, (unsigned long long)gb);
  for (int c : {2, 4, 8}) {
    run<32, 1>(tab, bytes, c, sms, sink); run<32, 2>(tab, bytes, c, sms, sink);
    run<32, 4>(tab, bytes, c, sms, sink); run<32, 8>(tab, bytes, c, sms, sink);
  }
  for (int c : {4, 8}) { run<64, 2>(tab, bytes, c, sms, sink); run<64, 4>(tab, bytes, c, sms, sink); }
  for (int c : {4, 8}) { run<128, 2>(tab, bytes, c, sms, sink); run<128, 4>(tab, bytes, c, sms, sink); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
