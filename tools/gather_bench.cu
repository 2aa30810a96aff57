// Random-sector gather microbenchmark: how many independent 32-byte reads per second does the
// B200 memory system sustain, as a function of the table size (L2-resident ... 96 GiB) and of
// the load instruction?  Sets the realistic ceiling for any probe kernel next to the
// streaming-copy peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t rng(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}

// MODE 0: one LDG.256 (nc, L1 no-allocate) per gather     MODE 1: two LDG.128 (plain ld.global)
// MODE 2: 2 lanes x LDG.128 per gather (one 32 B sector per lane pair)
// MODE 3: one LDG.64 per gather (8 B of the sector; request-rate test)
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_gather(const char *tab, uint64_t nslots, uint64_t per_thread, unsigned long long *sink) {
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (uint64_t i = 0; i < per_thread; i += U) {
    uint32_t w[U][8];
#pragma unroll
    for (int u = 0; u < U; u++) {
      uint64_t who = (MODE == 2) ? (tid >> 1) : tid;
      uint64_t slot = rng(who * 0x100000001b3ull + i + u) % nslots;
      const char *p = tab + slot * 32;
      if (MODE == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(p));
      } else if (MODE == 1) {
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]) : "l"(p));
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(p + 16));
      } else if (MODE == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]) : "l"(p + 16 * (tid & 1)));
        w[u][4] = w[u][5] = w[u][6] = w[u][7] = 0;
      } else {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w[u][0]), "=r"(w[u][1]) : "l"(p));
        w[u][2] = w[u][3] = w[u][4] = w[u][5] = w[u][6] = w[u][7] = 0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += w[u][0] ^ w[u][1] ^ w[u][2] ^ w[u][7];
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int MODE, int U>
static void run(const char *tab, uint64_t bytes, int ctas_per_sm, int sms, unsigned long long *sink) {
  uint64_t nslots = bytes / 32;
  int grid = sms * ctas_per_sm;
  uint64_t per_thread = 2048;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_gather<MODE, U><<<grid, 256>>>(tab, nslots, 256, sink);   // warm-up (also warms L2 for small tables)
  cudaEventRecord(a);
  k_gather<MODE, U><<<grid, 256>>>(tab, nslots, per_thread, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double n = (double)grid * 256 * per_thread / (MODE == 2 ? 2 : 1);
  printf("table %8.3f GiB  mode=%d U=%d ctas/sm=%d  %8.2f Gsector/s  %8.1f GB/s\n", bytes / 1073741824.0, MODE, U,
         ctas_per_sm, n / ms / 1e6, n * 32 / ms / 1e6);
}

int main(int argc, char **argv) {
  uint64_t gb = argc > 1 ? atoll(argv[1]) : 96;
  if (argc > 2) {            // L2 fetch granularity hint: 32, 64 or 128 bytes
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2]));
    size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity <- %d: %s, now %zu\n", atoi(argv[2]), cudaGetErrorString(e), v);
  } else { size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity default %zu\n", v); }
  uint64_t bytes = gb << 30;
  char *tab; unsigned long long *sink;
  if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  cudaMemset(tab, 1, bytes);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, L2 %d MB\n", p.name, sms, p.l2CacheSize >> 20);
  // table-size sweep with the kernel's load (LDG.256, 8 in flight per thread, 1024 threads/SM)
  for (uint64_t mb : std::initializer_list<uint64_t>{32, 1024, 4096, 16384, 32768, gb * 1024})
    if (mb <= gb * 1024) run<0, 4>(tab, mb << 20, 4, sms, sink);
  // instruction variants on the big table and on an L2-resident one
  for (uint64_t mb : std::initializer_list<uint64_t>{32, gb * 1024}) {
    run<0, 8>(tab, mb << 20, 8, sms, sink);
    run<1, 4>(tab, mb << 20, 4, sms, sink);
    run<2, 4>(tab, mb << 20, 4, sms, sink);
    run<2, 8>(tab, mb << 20, 8, sms, sink);
    run<3, 4>(tab, mb << 20, 4, sms, sink);
    run<3, 8>(tab, mb << 20, 8, sms, sink);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
