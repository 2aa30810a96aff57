// Second half of the streamed restricted table (joint_stream.cu), in its own translation unit: ptxas 12.9
// crashes when two kernel entries of one module share the out-of-line table lookups.
#include "joint_build.cuh"
#include "synth.cuh"

// the assembly multiplicity of every k-mer of the shard; SYNTH: also its synthetic read multiplicity
// (exactly k_synth_from_asm of table.cu: Poisson(cov x copies), 0.5 % collapsed, 0.5 % falsely duplicated)
__global__ void __launch_bounds__(256) k_joint_fill_asm(JointDev J, DevTable keys, DevTable A, int SYNTH, double cov, uint64_t seed) {
  JointCounts cnt = {0, 0, 0, 0};
  mfn_foreach_entry(keys, [&](uint64_t c, uint32_t) {
    const uint32_t va = mfn_value(A, c);
    if (!va) return;
    if (SYNTH) {
      const uint64_t sel = mfn_rng(c + 0x5851f42d4c957f2dull * (seed + 1)) % 1000;
      const double lambda = cov * (double)va * (sel < 5 ? 2.0 : (sel < 10 ? 0.5 : 1.0));
      const uint32_t v = mfn_poisson(lambda, c, seed);
      if (v) { if (J.w64) joint_insert<true>(J, c, 0, v, cnt); else joint_insert<false>(J, c, 0, v, cnt); }
    }
    if (J.w64) joint_insert<true>(J, c, 1, va, cnt); else joint_insert<false>(J, c, 1, va, cnt);
  });
  for (int o = 16; o > 0; o >>= 1) {
    cnt.home += __shfl_down_sync(0xffffffffu, cnt.home, o);
    cnt.spill += __shfl_down_sync(0xffffffffu, cnt.spill, o);
    cnt.stash += __shfl_down_sync(0xffffffffu, cnt.stash, o);
    cnt.big += __shfl_down_sync(0xffffffffu, cnt.big, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt.home) atomicAdd(&J.stats->n_home, (unsigned long long)cnt.home);
    if (cnt.spill) atomicAdd(&J.stats->n_spill, (unsigned long long)cnt.spill);
    if (cnt.stash) atomicAdd(&J.stats->n_stash, (unsigned long long)cnt.stash);
    if (cnt.big) atomicAdd(&J.stats->n_big, (unsigned long long)cnt.big);
  }
}

void mfn_joint_fill_asm_launch(const JointDev &J, const DevTable &keys, const DevTable &A, int synth, double cov, uint64_t seed,
                               int grid, cudaStream_t stream) {
  k_joint_fill_asm<<<grid, 256, 0, stream>>>(J, keys, A, synth, cov, seed);
}
