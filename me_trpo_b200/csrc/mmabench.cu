// Tensor-pipe micro-benchmark (dev tool): cycles per tcgen05.mma for a given N with A from
// shared memory (SS) or TMEM (TS), issued from a warp-uniform loop with hoisted descriptors.
// Operand contents are irrelevant (smem is left uninitialised); only timing is reported.
#include "umma.cuh"
#include "metrpo.h"
#include "metrpo_dev.h"
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

// Legacy warp-level tensor path (mma.sync) micro-benchmark: `warps` warps of one CTA each issue
// reps x 8 independent MMAs (8 accumulator chains); kind 0: m16n8k8 tf32, 1: m16n8k16 bf16.
// out_dev[0] = clock64 span of warp 0.  Priced here because the TRPO sample pass could use it.
namespace metrpo {
__global__ void mmasync_bench_kernel(int kind, int reps, unsigned long long* out) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 5, b1 = 9;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kind == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = static_cast<unsigned long long>(s != 12345.f); }
}
}  // namespace metrpo

extern "C" int metrpo_bench_mma_sync(int kind, int warps, int reps, unsigned long long* out_dev, void* stream_) {
  using namespace metrpo;
  if (kind < 0 || kind > 1 || warps < 1 || warps > 32 || reps < 1) return set_error(METRPO_ERR_INVALID, "bench_mma_sync: bad argument");
  mmasync_bench_kernel<<<1, warps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(kind, reps, out_dev);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
