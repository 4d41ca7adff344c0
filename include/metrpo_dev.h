/* metrpo_dev.h -- DEVELOPMENT tools built into me_trpo_b200/libmetrpo_dev.so, NOT part of the
 * product ABI (include/metrpo.h): the tcgen05 descriptor self-test that pins csrc/umma.cuh's
 * encodings (tests/test_umma_selftest.py) and the tensor-pipe issue-rate micro-benchmarks behind
 * DESIGN.md section 5's cycle numbers (tools/mma_bench.py, tools/mma_sync_bench.py). */
#ifndef METRPO_DEV_H_
#define METRPO_DEV_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* metrpo_last_error(void);

/* Single-CTA tcgen05 GEMM self-test: C[128,N] = A[128,K] * B[N,K]^T (bf16 in, fp32 out) through
 * the same descriptors the rollout kernel uses.  mode 0: SW128 smem operands, B by bulk copy;
 * 1: no-swizzle core-matrix operands; 2: A in TMEM.  cycles (device, may be NULL) receives the
 * clock64 span of `reps` back-to-back K loops. */
int metrpo_selftest_umma(int mode, int N, int K, int reps, const void* A_bf16, const void* B_bf16,
                         float* C, unsigned long long* cycles, void* stream);

/* Dev tool: tensor-pipe micro-benchmark.  Issues reps x 4 tcgen05.mma (M=128, K=16, given N) from
 * a warp-uniform loop; out_dev[0] = clock64 span of the issue loop, out_dev[1] = span until the
 * final commit is observed.  ts_mode 1: A operand from TMEM, 0: from shared memory. */
int metrpo_bench_mma(int ts_mode, int N, int reps, int two_acc, int a_col, int d_col,
                     int wait_each, unsigned long long* out_dev, void* stream);

/* Dev tool: legacy warp-level tensor path (mma.sync) micro-benchmark.  `warps` warps of one CTA
 * issue reps x 8 independent MMAs each; kind 0: m16n8k8 tf32, 1: m16n8k16 bf16.  out_dev[0] =
 * clock64 span of warp 0. */
int metrpo_bench_mma_sync(int kind, int warps, int reps, unsigned long long* out_dev, void* stream);

/* Dev hook: the ensemble fit's batched TF32 tcgen05 GEMM (csrc/fit_gemm.cuh) on caller-provided
 * device arrays.  C[m,n] = epi(sum_k A(m,k) B(n,k)) per model; a_mn / b_mn = 1: operand stored
 * [k][m] / [k][n]; epi 0 plain, 1 relu(acc + bias[n]), 2 aux > 0 ? acc : 0. */
int metrpo_dev_gemm_tf32(int M, int N, int Kd, int models, const float* A, long long lda,
                         long long strideA, int a_mn, const float* B, long long ldb,
                         long long strideB, int b_mn, float* C, long long ldc, long long strideC,
                         int epi, const float* bias, long long strideBias, const float* aux,
                         long long ldaux, long long strideAux, float* dbg_stage, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* METRPO_DEV_H_ */
