// k_sweep_over: windows whose bin was full are probed from the side list (see sweep_kernels.cuh).
#include "sweep_kernels.cuh"

void mfn_sweep_launch_over(const SweepParams &P, int grid, cudaStream_t stream) {
  k_sweep_over<0><<<grid, SW_THREADS, 0, stream>>>(P);
}
