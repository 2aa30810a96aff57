// The fused (read, assembly) table restricted to the k-mers of one sequence shard (own translation
// unit: ptxas 12.9 crashes when this kernel shares a module with the other joint-table kernels).
#include "joint_build.cuh"

#include <algorithm>
#include <cmath>

// ---- the fused table restricted to the k-mers of a sequence set ----------------------------------
// A -hist scan only ever asks for k-mers that occur in the sequences it scans.  The restricted table
// holds (v_read, v_asm) for exactly those k-mers of `keys` (the counted k-mers of the shard) that
// either database knows: 1.5 G instead of 3.5 G entries at the 3 Gbp benchmark, 1/N of that again
// when N ranks each restrict to their own shard -- which is what the probe pass streams per step.
// A lookup that does not find its k-mer here falls back to the per-database tables (sweep_slow), so
// results stay exact for ANY sequence; only the speed depends on scanning the sequences it was
// built for.
// (one entry point for both slot widths: two entries sharing the out-of-line probe crash ptxas)
__global__ void __launch_bounds__(256) k_joint_fill_keys(JointDev J, DevTable S, DevTable R, DevTable A) {
  JointCounts cnt = {0, 0, 0, 0};
  mfn_foreach_entry(S, [&](uint64_t c, uint32_t) {
    const uint32_t vr = mfn_value(R, c), va = mfn_value(A, c);
    // distinct keys: the two fields of a k-mer are written by the same thread, one after the other
    if (J.w64) {
      if (vr) joint_insert<true>(J, c, 0, vr, cnt);
      if (va) joint_insert<true>(J, c, 1, va, cnt);
    } else {
      if (vr) joint_insert<false>(J, c, 0, vr, cnt);
      if (va) joint_insert<false>(J, c, 1, va, cnt);
    }
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

int mfn_joint_build_restricted(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_table *keys, mfn_joint **out) {
  if (R->dev.k != A->dev.k || R->dev.k != keys->dev.k) { mfn_set_error("mer size mismatch"); return MFN_ERR_K_MISMATCH; }
  if (R->dev.probe_both || A->dev.probe_both) { mfn_set_error("joint table needs canonical databases"); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  mfn_joint *j = nullptr;
  const char *envp = getenv("MFN_JOINT_POW2");
  MFN_TRY(mfn_joint_alloc(ctx, R, A, keys->n_kmers, !(envp && atoi(envp) == 0),
                      [&](uint32_t rmax, uint32_t amax) { return mfn_joint_count_saturated(ctx, R, A, rmax, amax); }, &j));
  const int grid = ctx->sm_count * 8;
  k_joint_fill_keys<<<grid, 256, 0, ctx->stream>>>(j->dev, keys->dev, R->dev, A->dev);
  ctx->launches++;
  int st = mfn_joint_finish(ctx, j, false);
  if (st != MFN_OK) { mfn_joint_free(j); return st; }
  j->restricted = true;
  *out = j;
  return MFN_OK;
}

