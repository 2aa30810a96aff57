// k_scan<true> instantiation (see scan_kernel.cuh).
#include "scan_kernel.cuh"

int mfn_scan_grid_dump(mfn_ctx *ctx) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan<true>, SCAN_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const char *env = getenv("MFN_SCAN_CTAS_PER_SM");
  if (env && atoi(env) > 0) per_sm = std::min(per_sm, atoi(env));
  return ctx->sm_count * per_sm;
}

void mfn_scan_launch_dump(const ScanParams &P, int grid, cudaStream_t stream) {
  k_scan<true><<<grid, SCAN_THREADS, 0, stream>>>(P);
}
