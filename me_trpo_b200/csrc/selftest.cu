// UMMA self-test: one CTA computes C[128,N] = A[128,K] * B[N,K]^T with tcgen05.mma using the
// exact operand layouts / descriptors of the rollout kernel (umma.cuh).  It exists so that the
// descriptor encodings are pinned by a GPU unit test (tests/test_umma_selftest.py) and so that
// per-layout tensor-pipe cycle counts can be measured in isolation (reps > 1).
#include "umma.cuh"
#include "metrpo.h"
#include "metrpo_dev.h"
#include "common.cuh"
#include "fit_gemm.cuh"

namespace metrpo {

// pack row-major bf16 B[N,K] into SW128 K-major sub-tiles: sub-tile kk (64 k) at kk*N*128 bytes
__global__ void selftest_pack_sw128(const __nv_bfloat16* __restrict__ src, uint8_t* __restrict__ dst,
                                    int rows, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  int r = i / K, k = i % K;
  uint32_t off = (k >> 6) * (rows * 128u) + sw128_off(r, k & 63);
  *reinterpret_cast<__nv_bfloat16*>(dst + off) = src[i];
}

// mode 0: A,B SW128 (B via bulk copy of a pre-packed image)   (K % 64 == 0)
// mode 1: A,B no-swizzle core-matrix layout                   (K % 16 == 0)
// mode 2: A in TMEM (tcgen05.st, packed bf16 pairs), B SW128  (K % 64 == 0)
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(int mode, int N, int K, int reps, const __nv_bfloat16* __restrict__ A,
                     const __nv_bfloat16* __restrict__ B, const uint8_t* __restrict__ Bpacked,
                     float* __restrict__ C, unsigned long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  // SW128 tiles need 1024 B alignment in the shared window
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;                                  // 128*K*2 bytes
  uint8_t* sB = smem + ((128 * K * 2 + 1023) & ~1023); // N*K*2 bytes

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t acc = tmem;                 // columns [0, N)
  const uint32_t a_tm = tmem + 256;          // columns [256, 256 + K/2)   (mode 2)

  // ---- stage A (thread == row) ----
  {
    const int r = tid;
    if (mode == 0) {
      for (int k = 0; k < K; k += 8) {
        uint4 v = *reinterpret_cast<const uint4*>(A + r * K + k);
        *reinterpret_cast<uint4*>(sA + (k >> 6) * (128 * 128) + sw128_off(r, k & 63)) = v;
      }
    } else if (mode == 1) {
      for (int k = 0; k < K; k += 8) {
        uint4 v = *reinterpret_cast<const uint4*>(A + r * K + k);
        *reinterpret_cast<uint4*>(sA + noswz_off(r, k, 128)) = v;
      }
    } else {
      // TMEM lane = (warp%4)*32 + lane = tid for a 4-warp CTA
      for (int k = 0; k < K; k += 32) {
        uint32_t v[16];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + r * K + k);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = src[i];
        tmem_st16(a_tm + (static_cast<uint32_t>(warp * 32) << 16) + (k >> 1), v);
      }
      tmem_st_wait();
    }
  }
  // ---- stage B ----
  if (mode == 1) {
    for (int i = tid; i < N * (K / 8); i += 128) {
      int r = i / (K / 8), kc = i % (K / 8);
      uint4 v = *reinterpret_cast<const uint4*>(B + r * K + kc * 8);
      *reinterpret_cast<uint4*>(sB + noswz_off(r, kc * 8, N)) = v;
    }
  } else if (tid == 0) {
    mbar_arrive_expect_tx(&bar_b, N * K * 2);
    bulk_g2s(sB, Bpacked, N * K * 2, &bar_b);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    if (mode != 1) mbar_wait(&bar_b, 0);
    const uint32_t idesc = idesc_bf16_f32(128, N);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    unsigned long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int j = 0; j < K / 16; ++j) {
        uint32_t accum = (j > 0 || rep > 0) ? 1u : 0u;
        if (mode == 0) {
          uint64_t ad = smem_desc_sw128(a0 + (j >> 2) * (128 * 128) + (j & 3) * 32);
          uint64_t bd = smem_desc_sw128(b0 + (j >> 2) * (N * 128) + (j & 3) * 32);
          umma_ss(acc, ad, bd, idesc, accum);
        } else if (mode == 1) {
          uint64_t ad = smem_desc_noswz(a0 + 2 * j * (128 * 16), 128 * 16, 128);
          uint64_t bd = smem_desc_noswz(b0 + 2 * j * (N * 16), N * 16, 128);
          umma_ss(acc, ad, bd, idesc, accum);
        } else {
          uint64_t bd = smem_desc_sw128(b0 + (j >> 2) * (N * 128) + (j & 3) * 32);
          umma_ts(acc, a_tm + j * 8, bd, idesc, accum);
        }
      }
    }
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    unsigned long long t1 = clock64();
    if (cycles) cycles[0] = t1 - t0;
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  // ---- epilogue: thread == row ----
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(acc + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c + i < N) C[tid * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  (void)lane;
}

}  // namespace metrpo

using namespace metrpo;

extern "C" int metrpo_selftest_umma(int mode, int N, int K, int reps, const void* A_bf16,
                                    const void* B_bf16, float* C, unsigned long long* cycles,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (mode < 0 || mode > 2) return set_error(METRPO_ERR_INVALID, "selftest: mode must be 0..2");
  if (N < 16 || N > 256 || (N % 16)) return set_error(METRPO_ERR_INVALID, "selftest: N in [16,256], N%16==0");
  if (K < 16 || K > 256 || (K % 16) || (mode != 1 && (K % 64)))
    return set_error(METRPO_ERR_INVALID, "selftest: K in [16,256]; K%64==0 for SW128 modes");
  if (reps < 1) reps = 1;
  uint8_t* packed = nullptr;
  if (mode != 1) {
    METRPO_CUDA_OK(cudaMallocAsync(&packed, (size_t)N * K * 2, stream));
    int n = N * K;
    selftest_pack_sw128<<<(n + 255) / 256, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(B_bf16), packed, N, K);
  }
  size_t smem = ((128 * K * 2 + 1023) & ~1023) + (size_t)N * K * 2 + 1024;
  METRPO_CUDA_OK(cudaFuncSetAttribute(selftest_umma_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma_kernel<<<1, 128, smem, stream>>>(mode, N, K, reps,
                                                 static_cast<const __nv_bfloat16*>(A_bf16),
                                                 static_cast<const __nv_bfloat16*>(B_bf16), packed,
                                                 C, cycles);
  METRPO_CUDA_OK(cudaGetLastError());
  if (packed) METRPO_CUDA_OK(cudaFreeAsync(packed, stream));
  return METRPO_OK;
}

// Dev hook: the fit's batched TF32 GEMM (fit_gemm.cuh) on caller-provided device arrays.
extern "C" int metrpo_dev_gemm_tf32(int M, int N, int Kd, int models, const float* A, long long lda,
                                    long long strideA, int a_mn, const float* B, long long ldb,
                                    long long strideB, int b_mn, float* C, long long ldc, long long strideC,
                                    int epi, const float* bias, long long strideBias, const float* aux,
                                    long long ldaux, long long strideAux, float* dbg, void* stream_) {
  GemmParams p;
  p.dbg = reinterpret_cast<unsigned long long*>(dbg);
  p.M = M; p.N = N; p.Kd = Kd; p.a_mn = a_mn; p.b_mn = b_mn; p.epi = epi; p.trans_store = 0; p.round_out = 0; p.splits = 1; p.strideSplit = 0;
  p.C = C; p.ldc = ldc; p.strideC = strideC; p.bias = bias; p.strideBias = strideBias; p.colsum = nullptr;
  GemmOperands o;
  o.A = A; o.lda = lda; o.strideA = strideA; o.a_ext = M; o.a_kext = 0;
  o.B = B; o.ldb = ldb; o.strideB = strideB; o.b_ext = N; o.b_kext = 0;
  o.aux = aux; o.ldaux = ldaux; o.strideAux = strideAux;
  const int r = fit_gemm_launch(p, o, models, static_cast<cudaStream_t>(stream_));
  if (r) return set_error(METRPO_ERR_CUDA, "dev_gemm_tf32: launch set-up failed (%d)", r);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
