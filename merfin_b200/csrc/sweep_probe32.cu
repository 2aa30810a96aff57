// k_sweep_probe<false>: pass 2 of the range-partitioned scan for tables with 32-bit slots.
#include "sweep_kernels.cuh"

// cap > 0: at most that many CTAs per SM (a probe pass that shares the SMs with the next sub-sweep's scatter passes)
void mfn_sweep_launch_probe32(const SweepParams &P, int sm_count, cudaStream_t stream, int cap) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_probe<false>, SW_THREADS, 0);
  if (cap > 0) per_sm = std::min(per_sm, cap);
  k_sweep_probe<false><<<sm_count * std::max(per_sm, 1), SW_THREADS, 0, stream>>>(P);
}
