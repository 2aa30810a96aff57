// -completeness over two RESIDENT tables: every k-mer the read table holds, joined with the assembly table
// (merfin-completeness.C:48-144 streams both databases from disk; mfn_completeness_shard does that through the
// device decoder).  This entry measures the join itself at any scale without writing a database first: the
// benchmark's tables are synthesised on the device.  (own translation unit: ptxas 12.9 and out-of-line lookups)
#include "scan_kernel.cuh"

__global__ void __launch_bounds__(256) k_join_tables(DevTable R, DevTable A, unsigned long long *acc, PairTable sparse) {
  __shared__ uint32_t s_hist[SH_R * SH_A];
  for (uint32_t i = threadIdx.x; i < SH_R * SH_A; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  // two thirds of a read database are error k-mers the assembly does not hold, with value 1-3: those pairs are
  // counted in registers (one shared-memory address would take them all)
  uint32_t hot[3] = {0, 0, 0};
  mfn_foreach_entry(R, [&](uint64_t c, uint32_t vr) {
    const uint32_t va = mfn_value(A, c);
    if (va == 0 && vr >= 1 && vr <= 3) hot[vr - 1]++;
    else if (vr < SH_R && va < SH_A) atomicAdd(&s_hist[va * SH_R + vr], 1u);
    else if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
    else sparse_add(sparse, acc, vr, va, 1ull);
  });
#pragma unroll
  for (int v = 0; v < 3; v++) if (hot[v]) atomicAdd(&s_hist[v + 1], hot[v]);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < SH_R * SH_A; i += blockDim.x) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
  }
}

extern "C" int mfn_completeness_resident(mfn_ctx *ctx, const mfn_table *reads_raw, const mfn_table *asm_raw, mfn_accum *acc) {
  if (!ctx || !reads_raw || !asm_raw || !acc) return MFN_ERR_INVALID;
  if (reads_raw->streamed) { mfn_set_error("-completeness needs the read database resident or on disk, not streamed"); return MFN_ERR_UNSUPPORTED; }
  if (reads_raw->value_offset != 0 || asm_raw->value_offset != 0) { mfn_set_error("-completeness joins raw values: load both tables with minV = 0"); return MFN_ERR_INVALID; }
  if (reads_raw->dev.k != asm_raw->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", reads_raw->dev.k, asm_raw->dev.k); return MFN_ERR_K_MISMATCH; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  if (reads_raw->wide || asm_raw->wide) return mfn_wide_completeness_resident(ctx, reads_raw, asm_raw, acc);      // k > 32
  k_join_tables<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reads_raw->dev, asm_raw->dev, acc->d_buf, acc->sparse);
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  return MFN_OK;
}
