// Tensor-pipe micro-benchmark (dev tool): cycles per tcgen05.mma for a given N with A from
// shared memory (SS) or TMEM (TS), issued from a warp-uniform loop with hoisted descriptors.
// Operand contents are irrelevant (smem is left uninitialised); only timing is reported.
#include "umma.cuh"
#include "metrpo.h"
#include "common.cuh"

namespace metrpo {

__global__ void __launch_bounds__(128, 1)
mmabench_kernel(int ts_mode, int N, int reps, int two_acc, int a_col, int d_col, int wait_each,
                unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  // zero the operand tiles so that no NaN/denormal paths are exercised
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N);
    const uint64_t ad = smem_desc_sw128(smem_u32(smem));            // A tile [128 x 64]
    const uint64_t bd = smem_desc_sw128(smem_u32(smem + 16384));    // B tile [256 x 64]
    const uint32_t at = tmem + a_col;                                // A in TMEM (32 cols)
    unsigned long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + d_col + ((two_acc && (r & 1)) ? 64u : 0u);
      // optional sync primitives between MMA groups (bitmask), to price them in the issue loop
      bool ok = true;
      if (wait_each & 1) { if ((threadIdx.x & 31) == 0) ok = mbar_try_wait(&bar2, 1); }
      if (wait_each & 2) { if (!__all_sync(0xffffffffu, ok)) break; }
      if (wait_each & 4) tc_fence_after();
      if (wait_each & 16) { if (elect_one()) umma_commit(&bar2); __syncwarp(); }
      if (elect_one()) {
        if (ts_mode) {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_ts(d, at + 8 * j, bd + 2 * j, idesc, 1);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_ss(d, ad + 2 * j, bd + 2 * j, idesc, 1);
        }
      }
      __syncwarp();
    }
    unsigned long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    unsigned long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace metrpo

extern "C" int metrpo_bench_mma(int ts_mode, int N, int reps, int two_acc, int a_col, int d_col,
                                int wait_each, unsigned long long* out_dev, void* stream_) {
  using namespace metrpo;
  if (N < 16 || N > 256 || (N % 16)) return set_error(METRPO_ERR_INVALID, "bench_mma: bad N");
  const size_t smem = 16384 + 32768 + 2048;
  METRPO_CUDA_OK(cudaFuncSetAttribute(mmabench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mmabench_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream_)>>>(ts_mode, N, reps, two_acc, a_col, d_col, wait_each, out_dev);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
