// k_sweep_probe<false>: pass 2 of the range-partitioned scan for tables with 32-bit slots.
#include "sweep_kernels.cuh"

void mfn_sweep_launch_probe32(const SweepParams &P, int sm_count, cudaStream_t stream) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_probe<false>, SW_THREADS, 0);
  k_sweep_probe<false><<<sm_count * std::max(per_sm, 1), SW_THREADS, 0, stream>>>(P);
}
