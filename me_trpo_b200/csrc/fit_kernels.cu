// Ensemble dynamics fit on the device (include/metrpo.h metrpo_fit_*), SURVEY.md 8(f) N3:
//
//   metrpo_fit_step   one optimizer step of ALL K models of the ensemble on K independent
//                     minibatches (model_based_rl.py:957-970; utils.py:129-131,366-369):
//                     gather + normalise -> MLP forward -> de-normalised MSE loss
//                     (model_based_rl.py:57-71 with training.py:257) -> backward -> Adam
//                     (tf.train.AdamOptimizer, model_based_rl.py:154-163)
//   metrpo_fit_eval   per-model validation loss over a whole data set tiled K times (:934-935,
//                     :973-981) + per-model best-weights snapshot (:998-1007), all on the device
//   metrpo_fit_restore_best   recover_weights (:876-879, :1034)
//
// Default (METRPO_FIT_TF32): every contraction behind layer 0 -- forward, dgrad and wgrad of the
// hidden and output layers, and the layer-0 weight gradient -- runs on the tcgen05 tensor cores
// through ONE hand-written batched TF32 GEMM kernel (fit_gemm.cuh; TMA-fed, K-/MN-major operands
// read in place, bias + ReLU or ReLU mask + bias-gradient column sums fused into the epilogue);
// layer 0 itself (Din <= 88 deep) is fused into the gather kernel on the CUDA cores.  Around them:
// the MSE backward (bias b2, de-normalisation, residual, loss reduction, dO and db2 in one pass)
// and one Adam kernel over the K x P flat parameter block.  METRPO_FIT_FP32 is the fidelity mode:
// true fp32 products through cuBLAS (bit-compatible with the oracle to 1e-7) with separate
// bias / mask kernels.
#include <cublas_v2.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "fit_gemm.cuh"

namespace metrpo {

constexpr uint32_t PHILOX_STREAM_FIT = 0x30000u;   // minibatch row indices
constexpr int FIT_SPLITS = 3;   // split-K of the N <= 128 products when they would leave most SMs idle (40 CTAs at batch 1000)

// per-step scalars of a replayed CUDA graph of the training iteration (device copy, refreshed before
// every launch): kernels that take a `dargs` pointer read these instead of their by-value arguments
struct FitStepArgs { unsigned long long seed, offset; float lr_t; int pad; };

struct FitDims {
  int S, A, SA, drop, Din, H, K;
  int rnd;                               // 1 (TF32 mode): producers round GEMM operands to TF32-nearest
  int Sp, Dp;                            // S and Din rounded up to 32: row pitch of W2 / O and of Z (zero padded)
  int oW0, ob0, oW1, ob1, oW2, ob2, P;   // float offsets inside one model's parameter block (W2 is [H][Sp])
};

// x_data[n][SA] (state, action), y_data[n][S] (next state).  Row r of model k's minibatch is sample
// idx[r*K + k] (np.reshape(x_batch, (batch, -1)) + get_ith_tensor, model_based_rl.py:966-969,
// utils.py:366-369); idx == NULL -> Philox; identity != 0 -> row0 + r for every model (validation:
// np.tile(x, n_models), :934-935).
__global__ void fit_gather_kernel(FitDims d, const float* __restrict__ x_data, const float* __restrict__ y_data,
                                  int n_data, const int* __restrict__ idx, int identity, int row0,
                                  unsigned long long seed, unsigned long long offset, int rows,
                                  const float* __restrict__ norm, float* __restrict__ Z,
                                  float* __restrict__ XS, float* __restrict__ Y, long long strideZ,
                                  long long strideS, const FitStepArgs* __restrict__ dargs) {
  const int r = blockIdx.x * blockDim.y + threadIdx.y, k = blockIdx.y;
  if (r >= rows) return;
  if (dargs) { seed = dargs->seed; offset = dargs->offset; }
  int i;
  if (identity) i = row0 + r;
  else if (idx) i = idx[static_cast<size_t>(r) * d.K + k];
  else i = philox_index(seed, static_cast<uint64_t>(offset), static_cast<uint32_t>(r * d.K + k), PHILOX_STREAM_FIT, n_data);
  i = min(max(i, 0), n_data - 1);
  const float* xr = x_data + static_cast<size_t>(i) * d.SA;
  const float* yr = y_data + static_cast<size_t>(i) * d.S;
  const float* in_mean = norm;
  const float* in_std = norm + d.SA;
  float* z = Z + k * strideZ + static_cast<size_t>(r) * d.Dp;
  float* xs = XS + k * strideS + static_cast<size_t>(r) * d.S;
  float* y = Y + k * strideS + static_cast<size_t>(r) * d.S;
  for (int c = threadIdx.x; c < d.SA; c += blockDim.x) {
    const float v = xr[c];
    if (c >= d.drop) {   // training.py:228,146-154
      const float zv = __fdiv_rn(__fsub_rn(v, in_mean[c]), in_std[c]);
      z[c - d.drop] = d.rnd ? round_tf32(zv) : zv;
    }
    if (c < d.S) { xs[c] = v; y[c] = yr[c]; }
  }
}

// Fused layer 0: gather + normalise + column drop + (z W0 + b0) + ReLU  (training.py:146-154,207-208,228).
// The contraction is only Din (<= 88) deep, so it runs on the CUDA cores inside the pass that has to
// write H0 anyway.  Block = 64 rows x 128 hidden columns of one model, 256 threads, 4 x 8 register
// tile per thread; blockIdx.y == 0 also stores Z / XS / Y for the backward pass and the loss.
constexpr int L0_ROWS = 64, L0_COLS = 128;
__global__ void __launch_bounds__(256) fit_layer0_kernel(FitDims d, const float* __restrict__ x_data,
                                                         const float* __restrict__ y_data, int n_data,
                                                         const int* __restrict__ idx, int identity, int row0,
                                                         unsigned long long seed, unsigned long long offset, int rows,
                                                         const float* __restrict__ norm, const float* __restrict__ theta,
                                                         float* __restrict__ Z, float* __restrict__ XS, float* __restrict__ Y,
                                                         float* __restrict__ H0, long long strideZ, long long strideS,
                                                         long long strideH) {
  extern __shared__ __align__(16) float sm0[];
  const int Dp = d.Din + 1;                    // odd-ish row stride of the z tile
  float* zs = sm0;                             // [64][Dp]
  float* ws = sm0 + ((L0_ROWS * Dp + 3) & ~3); // [Din][128]
  __shared__ int sIdx[L0_ROWS];
  const int tid = threadIdx.x, k = blockIdx.z, r0 = blockIdx.x * L0_ROWS, c0 = blockIdx.y * L0_COLS;
  if (tid < L0_ROWS) {
    const int r = r0 + tid;
    int i = 0;
    if (r < rows) {
      if (identity) i = row0 + r;
      else if (idx) i = idx[static_cast<size_t>(r) * d.K + k];
      else i = philox_index(seed, static_cast<uint64_t>(offset), static_cast<uint32_t>(r * d.K + k), PHILOX_STREAM_FIT, n_data);
      i = min(max(i, 0), n_data - 1);
    }
    sIdx[tid] = i;
  }
  // W0 tile [Din][128 columns of this block], coalesced float4
  const float* W0 = theta + static_cast<size_t>(k) * d.P + d.oW0;
  for (int q = tid; q < d.Din * (L0_COLS / 4); q += 256) {
    const int i = q / (L0_COLS / 4), c4 = q - i * (L0_COLS / 4);
    reinterpret_cast<float4*>(ws + i * L0_COLS)[c4] = *reinterpret_cast<const float4*>(W0 + static_cast<size_t>(i) * d.H + c0 + 4 * c4);
  }
  __syncthreads();
  const float* in_mean = norm;
  const float* in_std = norm + d.SA;
  // the scattered reads of up to 8 elements per thread are issued back to back (a store between two
  // loads would serialise the DRAM latencies), then normalised and stored
  const int total = L0_ROWS * d.SA;
  for (int base = 0; base < total; base += 256 * 8) {
    float xv[8], yv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = base + u * 256 + tid;
      xv[u] = 0.f; yv[u] = 0.f;
      if (q < total) {
        const int rr = q / d.SA, c = q - rr * d.SA;
        if (r0 + rr < rows) {
          xv[u] = __ldg(&x_data[static_cast<size_t>(sIdx[rr]) * d.SA + c]);
          if (blockIdx.y == 0 && c < d.S) yv[u] = __ldg(&y_data[static_cast<size_t>(sIdx[rr]) * d.S + c]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = base + u * 256 + tid;
      if (q < total) {
        const int rr = q / d.SA, c = q - rr * d.SA, r = r0 + rr;
        float zv = 0.f;
        if (r < rows) {
          const float v = xv[u];
          zv = __fdiv_rn(__fsub_rn(v, in_mean[c]), in_std[c]);
          if (blockIdx.y == 0) {
            if (c >= d.drop) Z[k * strideZ + static_cast<size_t>(r) * d.Dp + c - d.drop] = d.rnd ? round_tf32(zv) : zv;
            if (c < d.S) {
              XS[k * strideS + static_cast<size_t>(r) * d.S + c] = v;
              Y[k * strideS + static_cast<size_t>(r) * d.S + c] = yv[u];
            }
          }
        }
        if (c >= d.drop) zs[rr * Dp + c - d.drop] = zv;
      }
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;   // rows 4ty..4ty+3, columns 8tx..8tx+7
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
#pragma unroll 2
  for (int i = 0; i < d.Din; ++i) {
    const float4 w0 = *reinterpret_cast<const float4*>(ws + i * L0_COLS + 8 * tx);
    const float4 w1 = *reinterpret_cast<const float4*>(ws + i * L0_COLS + 8 * tx + 4);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float z = zs[(4 * ty + a) * Dp + i];
      acc[a][0] = fmaf(z, w0.x, acc[a][0]); acc[a][1] = fmaf(z, w0.y, acc[a][1]);
      acc[a][2] = fmaf(z, w0.z, acc[a][2]); acc[a][3] = fmaf(z, w0.w, acc[a][3]);
      acc[a][4] = fmaf(z, w1.x, acc[a][4]); acc[a][5] = fmaf(z, w1.y, acc[a][5]);
      acc[a][6] = fmaf(z, w1.z, acc[a][6]); acc[a][7] = fmaf(z, w1.w, acc[a][7]);
    }
  }
  const float* b0 = theta + static_cast<size_t>(k) * d.P + d.ob0 + c0 + 8 * tx;
  const float4 bb0 = *reinterpret_cast<const float4*>(b0), bb1 = *reinterpret_cast<const float4*>(b0 + 4);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = r0 + 4 * ty + a;
    if (r < rows) {
      float4* o = reinterpret_cast<float4*>(H0 + k * strideH + static_cast<size_t>(r) * d.H + c0 + 8 * tx);
      float4 o0 = make_float4(fmaxf(acc[a][0] + bb0.x, 0.f), fmaxf(acc[a][1] + bb0.y, 0.f), fmaxf(acc[a][2] + bb0.z, 0.f),
                              fmaxf(acc[a][3] + bb0.w, 0.f));
      float4 o1 = make_float4(fmaxf(acc[a][4] + bb1.x, 0.f), fmaxf(acc[a][5] + bb1.y, 0.f), fmaxf(acc[a][6] + bb1.z, 0.f),
                              fmaxf(acc[a][7] + bb1.w, 0.f));
      if (d.rnd) {   // H0 is the A operand of the TF32 layer-1 GEMMs
        o0 = make_float4(round_tf32(o0.x), round_tf32(o0.y), round_tf32(o0.z), round_tf32(o0.w));
        o1 = make_float4(round_tf32(o1.x), round_tf32(o1.y), round_tf32(o1.z), round_tf32(o1.w));
      }
      o[0] = o0; o[1] = o1;
    }
  }
}

// H = relu(H + b)  (training.py:207-208), float4 over [K][rows][Hd]
__global__ void fit_bias_relu_kernel(float* __restrict__ Hbuf, const float* __restrict__ theta, int b_off,
                                     long long P, int rows, int Hd, long long strideH) {
  const int k = blockIdx.y;
  const long long n4 = static_cast<long long>(rows) * Hd / 4;
  float4* h4 = reinterpret_cast<float4*>(Hbuf + k * strideH);
  const float4* b4 = reinterpret_cast<const float4*>(theta + k * P + b_off);
  const int hq = Hd / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = h4[i];
    const float4 b = b4[i % hq];
    v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f);
    v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
    h4[i] = v;
  }
}

// Fused MSE forward/backward of the output layer.  One warp per (model, row):
//   pred = (diff_mean + diff_std * (O + b2)) + x                       (training.py:257)
//   loss_k += sum_s (pred - y)^2 * inv_rows                            (model_based_rl.py:57-71)
//   dO = 2 * inv_rows * diff_std * (pred - y)   (in place, backward only);  db2 partials per block
__global__ void fit_mse_kernel(FitDims d, float* __restrict__ O, const float* __restrict__ XS,
                               const float* __restrict__ Y, const float* __restrict__ theta,
                               const float* __restrict__ norm, int rows, double inv_rows, int backward,
                               long long strideS, long long strideO, int nsplit, long long strideSplit,
                               double* __restrict__ loss_acc, float* __restrict__ part2) {
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float* b2 = theta + static_cast<size_t>(k) * d.P + d.ob2;
  const float* dmean = norm + 2 * d.SA;
  const float* dstd = norm + 2 * d.SA + d.S;
  extern __shared__ float sh_db2[];   // [warps][S] per block
  double lsum = 0.0;
  float db_local[2] = {0.f, 0.f};   // lane owns columns lane, lane + 32 (S <= 64)
  for (int r = blockIdx.x * wpb + wib; r < rows; r += gridDim.x * wpb) {
    const size_t base = k * strideS + static_cast<size_t>(r) * d.S;
    const size_t obase = k * strideO + static_cast<size_t>(r) * d.Sp;   // O rows are Sp floats apart
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int s = lane + 32 * q;
      if (s < d.S) {
        float osum = O[obase + s];                      // split-K partials of H1 W2, added in split order
        for (int sp = 1; sp < nsplit; ++sp) osum = __fadd_rn(osum, O[sp * strideSplit + obase + s]);
        const float o = __fadd_rn(osum, b2[s]);
        const float pred = __fadd_rn(__fadd_rn(dmean[s], __fmul_rn(dstd[s], o)), XS[base + s]);
        const float diff = __fsub_rn(pred, Y[base + s]);
        lsum += static_cast<double>(diff) * diff;
        if (backward) {
          const float g = static_cast<float>(2.0 * inv_rows) * dstd[s] * diff;
          O[obase + s] = d.rnd ? round_tf32(g) : g;
          db_local[q] += g;
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  __shared__ double sh_loss[32];
  if (lane == 0) sh_loss[wib] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {   // one atomic per block (the warps' sums are added in a fixed order first)
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh_loss[w];
    atomicAdd(&loss_acc[k], t * inv_rows);
  }
  if (backward) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int s = lane + 32 * q;
      if (s < d.S) sh_db2[wib * d.S + s] = db_local[q];
    }
    __syncthreads();
    // per-block partial of db2, warps and blocks summed in a fixed order (bit-reproducible)
    for (int s = threadIdx.x; s < d.S; s += blockDim.x) {
      float acc = 0.f;
      for (int w = 0; w < wpb; ++w) acc += sh_db2[w * d.S + s];
      part2[(static_cast<size_t>(k) * gridDim.x + blockIdx.x) * d.S + s] = acc;
    }
  }
}

// dH *= (Hact > 0);  db[j] += sum_rows dH[., j].  Block = (32-column strip, 128-row slab, model):
// 8 x 32 threads, one float4 per thread and row, column sums reduced through shared memory and
// written as one partial row per slab (summed in a fixed order by fit_bias_grad_finish_kernel, so
// the update is bit-reproducible).
constexpr int COLSUM_ROWS = 128;
__global__ void fit_relu_bwd_colsum_kernel(float* __restrict__ dH, const float* __restrict__ Hact, int rows,
                                           int Hd, long long strideH, float* __restrict__ part_out) {
  const int k = blockIdx.z, c4 = blockIdx.x * 8 + threadIdx.x;   // float4 column index
  float4* g = reinterpret_cast<float4*>(dH + k * strideH);
  const float4* a = reinterpret_cast<const float4*>(Hact + k * strideH);
  const int hq = Hd / 4;
  const int r_end = min(rows, (blockIdx.y + 1) * COLSUM_ROWS);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int r = blockIdx.y * COLSUM_ROWS + threadIdx.y; r < r_end; r += 32) {
    const size_t o = static_cast<size_t>(r) * hq + c4;
    const float4 av = a[o];
    float4 v = g[o];
    v.x = av.x > 0.f ? v.x : 0.f; v.y = av.y > 0.f ? v.y : 0.f;
    v.z = av.z > 0.f ? v.z : 0.f; v.w = av.w > 0.f ? v.w : 0.f;
    g[o] = v;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  __shared__ float part[32][33];
  part[threadIdx.y][threadIdx.x * 4 + 0] = acc.x;
  part[threadIdx.y][threadIdx.x * 4 + 1] = acc.y;
  part[threadIdx.y][threadIdx.x * 4 + 2] = acc.z;
  part[threadIdx.y][threadIdx.x * 4 + 3] = acc.w;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int c = threadIdx.x; c < 32; c += 8) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += part[j][c];
      part_out[(static_cast<size_t>(k) * gridDim.y + blockIdx.y) * Hd + blockIdx.x * 32 + c] = sum;
    }
  }
}

// db1, db0, db2 = fixed-order sums of the per-slab / per-block partials.  Block = 32 columns x 8
// slab groups: every thread sums its slabs (stride 8) in order, the 8 group sums are added in order.
__global__ void __launch_bounds__(256) fit_bias_grad_finish_kernel(FitDims d, const float* __restrict__ part1,
                                                                   const float* __restrict__ part0, int nslab,
                                                                   const float* __restrict__ part2, int nblk2,
                                                                   const float* __restrict__ w0part, int nsplit0,
                                                                   long long w0_split_stride, long long w0_model_stride,
                                                                   float* __restrict__ grad) {
  __shared__ float red[8][33];
  const int k = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float* g = grad + static_cast<size_t>(k) * d.P;
  const float* p = nullptr;
  int n = 0, stride = 0, out = -1;
  if (j < d.H) { p = part1 + static_cast<size_t>(k) * nslab * d.H + j; n = nslab; stride = d.H; out = d.ob1 + j; }
  else if (j < 2 * d.H) { p = part0 + static_cast<size_t>(k) * nslab * d.H + (j - d.H); n = nslab; stride = d.H; out = d.ob0 + j - d.H; }
  else if (j < 2 * d.H + d.S) { p = part2 + static_cast<size_t>(k) * nblk2 * d.S + (j - 2 * d.H); n = nblk2; stride = d.S; out = d.ob2 + j - 2 * d.H; }
  else if (w0part != nullptr && j < 2 * d.H + d.S + d.Din * d.H) {   // dW0 = sum of its split-K partials
    const int e0 = j - (2 * d.H + d.S);
    p = w0part + static_cast<size_t>(k) * w0_model_stride + e0; n = nsplit0; stride = 0; out = d.oW0 + e0;
  }
  float s = 0.f;
  if (p) {
#pragma unroll 4
    for (int b = ty; b < n; b += 8) s += (stride ? p[static_cast<size_t>(b) * stride] : p[static_cast<size_t>(b) * w0_split_stride]);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && out >= 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][tx];
    g[out] = t;
  }
}

// tf.train.AdamOptimizer: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// theta -= lr_t * m / (sqrt(v) + eps),  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)  (host-computed)
__global__ void fit_adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                                float* __restrict__ v, long long n4, float lr_t, float beta1, float beta2,
                                float eps, const FitStepArgs* __restrict__ dargs) {
  if (dargs) lr_t = dargs->lr_t;
  float4* t4 = reinterpret_cast<float4*>(theta);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 t = t4[i], mm = m4[i], vv = v4[i];
    const float4 g = g4[i];
#define ADAM1(c)                                                     \
    mm.c = beta1 * mm.c + (1.f - beta1) * g.c;                       \
    vv.c = beta2 * vv.c + (1.f - beta2) * g.c * g.c;                 \
    t.c = t.c - lr_t * mm.c / (sqrtf(vv.c) + eps);
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    t4[i] = t; m4[i] = mm; v4[i] = vv;
  }
}

// per-model snapshot decision (model_based_rl.py:998-1007): mode 1: improved = min > new;
// mode 2: unconditional (initial save, :925-930, :938-946)
__global__ void fit_snapshot_flags_kernel(const double* __restrict__ loss_acc, float* __restrict__ min_losses,
                                          uint8_t* __restrict__ flags, float* __restrict__ losses_out, int K,
                                          int mode) {
  const int k = threadIdx.x;
  if (k >= K) return;
  const float l = static_cast<float>(loss_acc[k]);
  if (losses_out) losses_out[k] = l;
  uint8_t f = 0;
  if (mode == 2 || (mode == 1 && min_losses[k] > l)) { f = 1; min_losses[k] = l; }
  flags[k] = f;
}
__global__ void fit_snapshot_copy_kernel(const float* __restrict__ theta, float* __restrict__ best,
                                         const uint8_t* __restrict__ flags, long long P4) {
  const int k = blockIdx.y;
  if (!flags[k]) return;
  const float4* s = reinterpret_cast<const float4*>(theta) + k * P4;
  float4* t = reinterpret_cast<float4*>(best) + k * P4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < P4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    t[i] = s[i];
}
__global__ void fit_losses_out_kernel(const double* __restrict__ loss_acc, float* __restrict__ out, int K) {
  if (threadIdx.x < K) out[threadIdx.x] = static_cast<float>(loss_acc[threadIdx.x]);
}

}  // namespace metrpo

using namespace metrpo;

struct metrpo_fit {
  metrpo_fit_cfg cfg;
  FitDims d;
  long long P_pad;        // floats between consecutive models' parameter blocks (multiple of 4)
  int R;                  // row capacity of the activation buffers
  cublasHandle_t blas = nullptr;
  cublasComputeType_t compute;
  float *theta = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr, *best = nullptr;
  float *Z = nullptr, *XS = nullptr, *Y = nullptr, *H0 = nullptr, *H1 = nullptr, *O = nullptr, *D1 = nullptr;
  float* norm = nullptr;
  float *part1 = nullptr, *part0 = nullptr, *part2 = nullptr;   // bias-gradient partials
  float* w0part = nullptr;                                       // split-K partials of dW0: [FIT_SPLITS][K][up4(Din*H)]
  double* loss_acc = nullptr;
  float* min_losses = nullptr;
  uint8_t* flags = nullptr;
  long long adam_t = 0;
  bool norm_set = false;
  std::vector<char> w_set;
  int last_launches = 0;
  int nslab_max = 0;      // row slabs of the bias-gradient partial buffers
  int o_splits = 1;       // split-K factor of the last forward's output-layer product (read by the MSE kernel)
  // backward pass: weight-gradient GEMMs run on a side stream next to the data-gradient GEMMs
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_w2 = nullptr, ev_d1 = nullptr, ev_join = nullptr;
  // CUDA graph of the training iteration (TF32 mode, Philox minibatches): captured on `cap` after one
  // eager iteration with the same (x, y, n_data, batch, losses), replayed into the caller's stream
  FitStepArgs* dargs = nullptr;
  cudaStream_t cap = nullptr;
  cudaGraphExec_t gexec = nullptr;
  const float *gx = nullptr, *gy = nullptr;
  int gn = 0, gbatch = 0, gwarm = 0, glaunches = 0;
  int use_graph = 1;
};

static void fit_free(metrpo_fit* h) {
  if (!h) return;
  if (h->blas) cublasDestroy(h->blas);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  if (h->cap) cudaStreamDestroy(h->cap);
  cudaFree(h->dargs);
  if (h->side) cudaStreamDestroy(h->side);
  for (cudaEvent_t e : {h->ev_fork, h->ev_w2, h->ev_d1, h->ev_join}) if (e) cudaEventDestroy(e);
  cudaFree(h->theta); cudaFree(h->grad); cudaFree(h->m); cudaFree(h->v); cudaFree(h->best);
  cudaFree(h->Z); cudaFree(h->XS); cudaFree(h->Y); cudaFree(h->H0); cudaFree(h->H1); cudaFree(h->O);
  cudaFree(h->D1); cudaFree(h->norm); cudaFree(h->part1); cudaFree(h->part0); cudaFree(h->part2); cudaFree(h->w0part); cudaFree(h->loss_acc); cudaFree(h->min_losses); cudaFree(h->flags);
  delete h;
}

#define METRPO_BLAS_OK(expr)                                                                     \
  do {                                                                                           \
    cublasStatus_t _s = (expr);                                                                  \
    if (_s != CUBLAS_STATUS_SUCCESS)                                                             \
      return set_error(METRPO_ERR_CUDA, "%s:%d: %s -> cublas status %d", __FILE__, __LINE__, #expr, (int)_s); \
  } while (0)

extern "C" int metrpo_fit_create(const metrpo_fit_cfg* cfg, metrpo_fit_t** out) {
  if (!cfg || !out) return set_error(METRPO_ERR_INVALID, "fit_create: null argument");
  *out = nullptr;
  const metrpo_fit_cfg& c = *cfg;
  if (c.state_dim < 1 || c.action_dim < 1 || c.n_models < 1 || c.max_rows < 1)
    return set_error(METRPO_ERR_INVALID, "fit_create: S, A, K, max_rows must be >= 1");
  if (c.state_dim > 64 || c.n_models > 64) return set_error(METRPO_ERR_UNSUPPORTED, "fit_create: S <= 64 and K <= 64 in this build");
  if (c.drop_cols < 0 || c.drop_cols > 2 || c.drop_cols >= c.state_dim)
    return set_error(METRPO_ERR_INVALID, "fit_create: drop_cols must be 0, 1 or 2");
  if (c.hidden < 32 || c.hidden % 32)
    return set_error(METRPO_ERR_UNSUPPORTED, "fit_create: hidden width must be a multiple of 32 (got %d)", c.hidden);
  if (c.precision != METRPO_FIT_TF32 && c.precision != METRPO_FIT_FP32)
    return set_error(METRPO_ERR_INVALID, "fit_create: unknown precision %d", c.precision);
  METRPO_CUDA_OK(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  METRPO_CUDA_OK(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10)
    return set_error(METRPO_ERR_UNSUPPORTED, "fit_create: device %d is sm_%d%d; this library is sm_100a only (no fallback)", c.device, prop.major, prop.minor);

  metrpo_fit* h = new metrpo_fit();
  h->cfg = c;
  FitDims& d = h->d;
  d.S = c.state_dim; d.A = c.action_dim; d.SA = d.S + d.A; d.drop = c.drop_cols; d.Din = d.SA - d.drop;
  d.H = c.hidden; d.K = c.n_models;
  // every sub-block starts on a float4 boundary (H % 32 == 0; W0 and W2 are padded up)
  auto up4 = [](int v) { return (v + 3) / 4 * 4; };
  d.Sp = (d.S + 31) / 32 * 32; d.Dp = (d.Din + 31) / 32 * 32;
  d.rnd = c.precision == METRPO_FIT_TF32 ? 1 : 0;
  d.oW0 = 0; d.ob0 = up4(d.Din * d.H); d.oW1 = d.ob0 + d.H; d.ob1 = d.oW1 + d.H * d.H;
  d.oW2 = d.ob1 + d.H; d.ob2 = d.oW2 + d.H * d.Sp; d.P = d.ob2 + up4(d.S);
  h->P_pad = d.P;
  h->R = c.max_rows;
  h->compute = c.precision == METRPO_FIT_FP32 ? CUBLAS_COMPUTE_32F : CUBLAS_COMPUTE_32F_FAST_TF32;
  h->w_set.assign(d.K, 0);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
  };
  const size_t PK = static_cast<size_t>(d.P) * d.K * 4, RK = static_cast<size_t>(h->R) * d.K * 4;
  alloc((void**)&h->theta, PK); alloc((void**)&h->grad, PK); alloc((void**)&h->m, PK);
  alloc((void**)&h->v, PK); alloc((void**)&h->best, PK);
  alloc((void**)&h->Z, RK * d.Dp); alloc((void**)&h->XS, RK * d.S); alloc((void**)&h->Y, RK * d.S);
  alloc((void**)&h->H0, RK * d.H); alloc((void**)&h->H1, RK * d.H); alloc((void**)&h->D1, RK * d.H);
  alloc((void**)&h->O, RK * d.Sp * FIT_SPLITS);                 // [FIT_SPLITS][K][R][Sp]: split-K partials of H1 W2
  alloc((void**)&h->w0part, static_cast<size_t>(FIT_SPLITS) * d.K * ((d.Din * d.H + 3) / 4 * 4) * 4);
  alloc((void**)&h->norm, (2 * d.SA + 2 * d.S) * 4);
  const size_t nslab_max = static_cast<size_t>((h->R + 255) / 256) * 8;   // 32-row slabs of the GEMM epilogue
  h->nslab_max = static_cast<int>(nslab_max);
  alloc((void**)&h->part1, nslab_max * d.K * d.H * 4); alloc((void**)&h->part0, nslab_max * d.K * d.H * 4);
  alloc((void**)&h->part2, static_cast<size_t>(148) * d.K * d.S * 4);
  alloc((void**)&h->loss_acc, d.K * 8); alloc((void**)&h->min_losses, d.K * 4); alloc((void**)&h->flags, d.K);
  if (e != cudaSuccess) {
    fit_free(h);
    return set_error(e == cudaErrorMemoryAllocation ? METRPO_ERR_NOMEM : METRPO_ERR_CUDA, "fit_create: %s", cudaGetErrorString(e));
  }
  {
    const size_t smem0 = (static_cast<size_t>((L0_ROWS * (d.Din + 1) + 3) & ~3) + static_cast<size_t>(d.Din) * L0_COLS) * 4;
    if (smem0 > 48 * 1024)
      cudaFuncSetAttribute(fit_layer0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem0));
  }
  if (cublasCreate(&h->blas) != CUBLAS_STATUS_SUCCESS) {
    fit_free(h);
    return set_error(METRPO_ERR_CUDA, "fit_create: cublasCreate failed");
  }
  e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->dargs, sizeof(FitStepArgs));
  { const char* ev = getenv("METRPO_FIT_GRAPH"); h->use_graph = ev ? atoi(ev) : 1; }
  for (cudaEvent_t* ev : {&h->ev_fork, &h->ev_w2, &h->ev_d1, &h->ev_join})
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    fit_free(h);
    return set_error(METRPO_ERR_CUDA, "fit_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return METRPO_OK;
}

extern "C" int metrpo_fit_destroy(metrpo_fit_t* h) {
  if (!h) return METRPO_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  fit_free(h);
  return METRPO_OK;
}

extern "C" int metrpo_fit_num_params(const metrpo_fit_t* h) {
  if (!h) return 0;
  const FitDims& d = h->d;
  return d.Din * d.H + d.H + d.H * d.H + d.H + d.H * d.S + d.S;
}
extern "C" int metrpo_fit_last_launches(const metrpo_fit_t* h) { return h ? h->last_launches : 0; }

static int fit_copy_weights(metrpo_fit* h, int k, float* const* user, bool to_lib, float* block, cudaStream_t st) {
  const FitDims& d = h->d;
  const int off[6] = {d.oW0, d.ob0, d.oW1, d.ob1, d.oW2, d.ob2};
  const int len[6] = {d.Din * d.H, d.H, d.H * d.H, d.H, d.H * d.S, d.S};
  float* base = block + static_cast<size_t>(k) * d.P;
  for (int i = 0; i < 6; ++i) {
    if (i == 4) {   // W2 lives as [H][Sp] (zero padded columns) inside the block
      if (to_lib) METRPO_CUDA_OK(cudaMemcpy2DAsync(base + off[i], d.Sp * 4, user[i], d.S * 4, d.S * 4, d.H, cudaMemcpyDeviceToDevice, st));
      else METRPO_CUDA_OK(cudaMemcpy2DAsync(user[i], d.S * 4, base + off[i], d.Sp * 4, d.S * 4, d.H, cudaMemcpyDeviceToDevice, st));
      continue;
    }
    if (to_lib) METRPO_CUDA_OK(cudaMemcpyAsync(base + off[i], user[i], len[i] * 4, cudaMemcpyDeviceToDevice, st));
    else METRPO_CUDA_OK(cudaMemcpyAsync(user[i], base + off[i], len[i] * 4, cudaMemcpyDeviceToDevice, st));
  }
  return METRPO_OK;
}

extern "C" int metrpo_fit_set_weights(metrpo_fit_t* h, int k, const float* W0, const float* b0, const float* W1,
                                      const float* b1, const float* W2, const float* b2, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_set_weights: null handle");
  if (k < 0 || k >= h->d.K) return set_error(METRPO_ERR_INVALID, "fit_set_weights: model index %d out of range", k);
  if (!W0 || !b0 || !W1 || !b1 || !W2 || !b2) return set_error(METRPO_ERR_INVALID, "fit_set_weights: null weight pointer");
  float* user[6] = {const_cast<float*>(W0), const_cast<float*>(b0), const_cast<float*>(W1),
                    const_cast<float*>(b1), const_cast<float*>(W2), const_cast<float*>(b2)};
  int rc = fit_copy_weights(h, k, user, true, h->theta, static_cast<cudaStream_t>(stream));
  if (rc == METRPO_OK) h->w_set[k] = 1;
  return rc;
}

extern "C" int metrpo_fit_get_weights(metrpo_fit_t* h, int k, float* W0, float* b0, float* W1, float* b1,
                                      float* W2, float* b2, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_get_weights: null handle");
  if (k < 0 || k >= h->d.K) return set_error(METRPO_ERR_INVALID, "fit_get_weights: model index %d out of range", k);
  if (!W0 || !b0 || !W1 || !b1 || !W2 || !b2) return set_error(METRPO_ERR_INVALID, "fit_get_weights: null weight pointer");
  float* user[6] = {W0, b0, W1, b1, W2, b2};
  return fit_copy_weights(h, k, user, false, h->theta, static_cast<cudaStream_t>(stream));
}

extern "C" int metrpo_fit_set_normalization(metrpo_fit_t* h, const float* in_mean, const float* in_std,
                                            const float* diff_mean, const float* diff_std, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_set_normalization: null handle");
  if (!in_mean || !in_std || !diff_mean || !diff_std) return set_error(METRPO_ERR_INVALID, "fit_set_normalization: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int SA = h->d.SA, S = h->d.S;
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm, in_mean, SA * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + SA, in_std, SA * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + 2 * SA, diff_mean, S * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + 2 * SA + S, diff_std, S * 4, cudaMemcpyDeviceToDevice, st));
  h->norm_set = true;
  return METRPO_OK;
}

extern "C" int metrpo_fit_reset_adam(metrpo_fit_t* h, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_reset_adam: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t PK = static_cast<size_t>(h->d.P) * h->d.K * 4;
  METRPO_CUDA_OK(cudaMemsetAsync(h->m, 0, PK, st));
  METRPO_CUDA_OK(cudaMemsetAsync(h->v, 0, PK, st));
  h->adam_t = 0;
  return METRPO_OK;
}

// row-major C[M,N] = op(A) op(B), batched over the K models (column-major cuBLAS sees the
// transposes: C^T = op(B)^T op(A)^T)
static cublasStatus_t gemm_rm(metrpo_fit* h, bool tA, bool tB, int M, int N, int Kd, const float* A, int lda,
                              long long sA, const float* B, int ldb, long long sB, float* C, int ldc,
                              long long sC) {
  const float one = 1.f, zero = 0.f;
  return cublasGemmStridedBatchedEx(h->blas, tB ? CUBLAS_OP_T : CUBLAS_OP_N, tA ? CUBLAS_OP_T : CUBLAS_OP_N, N, M,
                                    Kd, &one, B, CUDA_R_32F, ldb, sB, A, CUDA_R_32F, lda, sA, &zero, C,
                                    CUDA_R_32F, ldc, sC, h->d.K, h->compute, CUBLAS_GEMM_DEFAULT);
}

static int fit_check_ready(metrpo_fit* h, const char* who) {
  for (int k = 0; k < h->d.K; ++k)
    if (!h->w_set[k]) return set_error(METRPO_ERR_STATE, "%s: weights of model %d were never set", who, k);
  if (!h->norm_set) return set_error(METRPO_ERR_STATE, "%s: normalization constants were never set", who);
  return METRPO_OK;
}

// C = epi(A B^T) on the tcgen05 GEMM (fit_gemm.cuh), batched over the K models
static int own_gemm(metrpo_fit* h, int M, int N, int Kd, const float* A, long long lda, long long sA, int a_mn,
                    int a_ext, const float* B, long long ldb, long long sB, int b_mn, int b_ext, float* C,
                    long long ldc, long long sC, int epi, const float* bias, const float* aux, float* colsum,
                    int trans_store, cudaStream_t st, int a_kext = 0, int b_kext = 0, int splits = 1,
                    long long strideSplit = 0) {
  GemmParams p;
  p.M = M; p.N = N; p.Kd = Kd; p.a_mn = a_mn; p.b_mn = b_mn; p.epi = epi; p.trans_store = trans_store;
  p.round_out = (epi != GEMM_EPI_PLAIN) ? 1 : 0;   // H1 / dH1 / dH0 feed later GEMMs
  p.C = C; p.ldc = ldc; p.strideC = sC; p.bias = bias; p.strideBias = h->d.P; p.colsum = colsum; p.dbg = nullptr;
  p.splits = splits; p.strideSplit = strideSplit;
  GemmOperands o;
  o.A = A; o.lda = lda; o.strideA = sA; o.a_ext = a_ext; o.a_kext = a_kext;
  o.B = B; o.ldb = ldb; o.strideB = sB; o.b_ext = b_ext; o.b_kext = b_kext;
  o.aux = aux; o.ldaux = h->d.H; o.strideAux = static_cast<long long>(h->R) * h->d.H;
  const int r = fit_gemm_launch(p, o, h->d.K, st);
  if (r) return set_error(METRPO_ERR_CUDA, "fit: tcgen05 GEMM set-up failed (code %d; M %d N %d K %d)", r, M, N, Kd);
  return METRPO_OK;
}

// gather + forward of `rows` rows per model; leaves O = H1 W2 (bias b2 is added by fit_mse_kernel)
static int fit_forward(metrpo_fit* h, const float* x, const float* y, int n_data, const int* idx, int identity,
                       int row0, unsigned long long seed, unsigned long long offset, int rows, cudaStream_t st,
                       int& launches, const FitStepArgs* dargs = nullptr) {
  const FitDims& d = h->d;
  const long long R = h->R, sZ = R * d.Dp, sS = R * d.S, sO = R * d.Sp, sH = R * d.H, P = d.P;
  const bool own = h->cfg.precision == METRPO_FIT_TF32;
  const int eb = static_cast<int>(std::min<long long>((static_cast<long long>(rows) * d.H / 4 + 255) / 256, 1184));
  int rc;
  if (!own && d.H % L0_COLS == 0) {   // fp32 fidelity mode: layer 0 on the CUDA cores inside the gather pass
    const size_t smem0 = (static_cast<size_t>((L0_ROWS * (d.Din + 1) + 3) & ~3) + static_cast<size_t>(d.Din) * L0_COLS) * 4;
    fit_layer0_kernel<<<dim3((rows + L0_ROWS - 1) / L0_ROWS, d.H / L0_COLS, d.K), 256, smem0, st>>>(
        d, x, y, n_data, idx, identity, row0, seed, offset, rows, h->norm, h->theta, h->Z, h->XS, h->Y, h->H0, sZ, sS, sH);
    launches += 1;
  } else {   // gather, then layer 0 as a GEMM (one 32-deep K block: Din <= 88 is zero padded to Dp)
    dim3 blk(32, 8), grd((rows + 7) / 8, d.K);
    fit_gather_kernel<<<grd, blk, 0, st>>>(d, x, y, n_data, idx, identity, row0, seed, offset, rows, h->norm,
                                          h->Z, h->XS, h->Y, sZ, sS, dargs);
    if (own) {   // H0 = relu(Z W0 + b0): A = Z [rows][Dp] (zero padded), B = W0 [k = Din][n = H] MN-major
      rc = own_gemm(h, rows, d.H, d.Dp, h->Z, d.Dp, sZ, 0, rows, h->theta + d.oW0, d.H, P, 1, d.H, h->H0, d.H, sH,
                    GEMM_EPI_BIAS_RELU, h->theta + d.ob0, nullptr, nullptr, 0, st, 0, d.Din);
      if (rc != METRPO_OK) return rc;
      launches += 2;
    } else {
      METRPO_BLAS_OK(gemm_rm(h, false, false, rows, d.H, d.Din, h->Z, d.Dp, sZ, h->theta + d.oW0, d.H, P, h->H0, d.H, sH));
      fit_bias_relu_kernel<<<dim3(eb, d.K), 256, 0, st>>>(h->H0, h->theta, d.ob0, P, rows, d.H, sH);
      launches += 3;
    }
  }
  if (own) {
    // H1 = relu(H0 W1 + b1): A = H0 [rows][H] K-major, B = W1 [k][n] MN-major, bias + ReLU in the epilogue
    rc = own_gemm(h, rows, d.H, d.H, h->H0, d.H, sH, 0, rows, h->theta + d.oW1, d.H, P, 1, d.H, h->H1, d.H, sH,
                  GEMM_EPI_BIAS_RELU, h->theta + d.ob1, nullptr, nullptr, 0, st);
    if (rc != METRPO_OK) return rc;
    // O = H1 W2: B = W2 [k = H][n = Sp] MN-major (zero padded columns), 128-column tile
    // (few row tiles: split the reduction over FIT_SPLITS CTAs per tile, the MSE kernel adds the partials)
    h->o_splits = (((rows + 127) / 128) * d.K * FIT_SPLITS <= 160 && d.H / 32 >= 2 * FIT_SPLITS) ? FIT_SPLITS : 1;
    rc = own_gemm(h, rows, d.S, d.H, h->H1, d.H, sH, 0, rows, h->theta + d.oW2, d.Sp, P, 1, d.Sp, h->O, d.Sp, sO,
                  GEMM_EPI_PLAIN, nullptr, nullptr, nullptr, 0, st, 0, 0, h->o_splits, static_cast<long long>(d.K) * sO);
    if (rc != METRPO_OK) return rc;
    launches += 2;
  } else {
    METRPO_BLAS_OK(gemm_rm(h, false, false, rows, d.H, d.H, h->H0, d.H, sH, h->theta + d.oW1, d.H, P, h->H1, d.H, sH));
    fit_bias_relu_kernel<<<dim3(eb, d.K), 256, 0, st>>>(h->H1, h->theta, d.ob1, P, rows, d.H, sH);
    METRPO_BLAS_OK(gemm_rm(h, false, false, rows, d.S, d.H, h->H1, d.H, sH, h->theta + d.oW2, d.Sp, P, h->O, d.Sp, sO));
    h->o_splits = 1;
    launches += 3;
  }
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}

// every launch of one training iteration on `st` (dargs != NULL: the per-step scalars come from device memory)
static int fit_step_issue(metrpo_fit* h, const float* x, const float* y, int n_data, const int32_t* idx, int batch,
                          uint64_t seed, uint64_t offset, float lr_t, float* losses, cudaStream_t st,
                          const FitStepArgs* dargs) {
  int rc;
  const bool own = h->cfg.precision == METRPO_FIT_TF32;
  if (!own) METRPO_BLAS_OK(cublasSetStream(h->blas, st));
  const FitDims& d = h->d;
  const long long R = h->R, sZ = R * d.Dp, sS = R * d.S, sO = R * d.Sp, sH = R * d.H, P = d.P;
  int launches = 0;
  METRPO_CUDA_OK(cudaMemsetAsync(h->loss_acc, 0, d.K * 8, st));
  rc = fit_forward(h, x, y, n_data, idx, 0, 0, seed, offset, batch, st, launches, dargs);
  if (rc != METRPO_OK) return rc;
  const int mb = std::min((batch + 7) / 8, 148);
  fit_mse_kernel<<<dim3(mb, d.K), 256, 8 * d.S * 4, st>>>(d, h->O, h->XS, h->Y, h->theta, h->norm, batch, 1.0 / batch, 1,
                                                     sS, sO, h->o_splits, static_cast<long long>(d.K) * sO, h->loss_acc,
                                                     h->part2);
  int nslab, w0_splits = 1;
  const long long P0 = (static_cast<long long>(d.Din) * d.H + 3) / 4 * 4;   // model stride of the dW0 partials
  if (own) {
    nslab = fit_gemm_colsum_slabs(batch, d.H);
    // The weight-gradient GEMMs (side stream) run next to the data-gradient GEMMs (caller's stream):
    //   side:  dW2 ............ | wait dH1 | dW1 .......................... |
    //   main:  dH1 ... | wait dW2 (it reads H1, whose buffer dH0 reuses) | dH0 ... dW0 | join
    cudaStream_t sd = h->side;
    METRPO_CUDA_OK(cudaEventRecord(h->ev_fork, st));
    METRPO_CUDA_OK(cudaStreamWaitEvent(sd, h->ev_fork, 0));
    // dW2[H,S] = H1^T dO: A = H1 [k = row][m] MN-major, B = dO [k = row][n = Sp] MN-major
    rc = own_gemm(h, d.H, d.S, batch, h->H1, d.H, sH, 1, d.H, h->O, d.Sp, sO, 1, d.Sp, h->grad + d.oW2, d.Sp, P,
                  GEMM_EPI_PLAIN, nullptr, nullptr, nullptr, 0, sd);
    if (rc != METRPO_OK) return rc;
    METRPO_CUDA_OK(cudaEventRecord(h->ev_w2, sd));
    // dH1 = (dO W2^T) * (H1 > 0), db1 = column sums: A = dO [rows][Sp] K-major, B = W2 [n = H][k = Sp] K-major
    rc = own_gemm(h, batch, d.H, d.Sp, h->O, d.Sp, sO, 0, batch, h->theta + d.oW2, d.Sp, P, 0, d.H, h->D1, d.H, sH,
                  GEMM_EPI_MASK, nullptr, h->H1, h->part1, 0, st);
    if (rc != METRPO_OK) return rc;
    METRPO_CUDA_OK(cudaEventRecord(h->ev_d1, st));
    METRPO_CUDA_OK(cudaStreamWaitEvent(sd, h->ev_d1, 0));
    // dW1 = H0^T dH1: both operands MN-major, reduction over the rows
    rc = own_gemm(h, d.H, d.H, batch, h->H0, d.H, sH, 1, d.H, h->D1, d.H, sH, 1, d.H, h->grad + d.oW1, d.H, P,
                  GEMM_EPI_PLAIN, nullptr, nullptr, nullptr, 0, sd);
    if (rc != METRPO_OK) return rc;
    METRPO_CUDA_OK(cudaEventRecord(h->ev_join, sd));
    // dH0 = (dH1 W1^T) * (H0 > 0) -> H1's buffer (H1 is dead once dW2 has read it), db0 = column sums
    METRPO_CUDA_OK(cudaStreamWaitEvent(st, h->ev_w2, 0));
    rc = own_gemm(h, batch, d.H, d.H, h->D1, d.H, sH, 0, batch, h->theta + d.oW1, d.H, P, 0, d.H, h->H1, d.H, sH,
                  GEMM_EPI_MASK, nullptr, h->H0, h->part0, 0, st);
    if (rc != METRPO_OK) return rc;
    // dW0[Din,H] = Z^T dH0, computed as its transpose (M = H fills the 128 tensor-core rows) and
    // stored transposed: A = dH0 [k = row][m = H] MN-major, B = Z [k = row][n = Dp] MN-major
    // (H / 128 row tiles x K models = 40 CTAs at H = 1024: split-K partials, summed with the bias gradients)
    w0_splits = ((d.H + 127) / 128 * d.K * FIT_SPLITS <= 160 && (batch + 31) / 32 >= 2 * FIT_SPLITS) ? FIT_SPLITS : 1;
    if (w0_splits > 1)
      rc = own_gemm(h, d.H, d.Din, batch, h->H1, d.H, sH, 1, d.H, h->Z, d.Dp, sZ, 1, d.Dp, h->w0part, d.H, P0,
                    GEMM_EPI_PLAIN, nullptr, nullptr, nullptr, 1, st, 0, 0, w0_splits, static_cast<long long>(d.K) * P0);
    else
      rc = own_gemm(h, d.H, d.Din, batch, h->H1, d.H, sH, 1, d.H, h->Z, d.Dp, sZ, 1, d.Dp, h->grad + d.oW0, d.H, P,
                    GEMM_EPI_PLAIN, nullptr, nullptr, nullptr, 1, st);
    if (rc != METRPO_OK) return rc;
    METRPO_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
    launches += 5;
  } else {
    nslab = (batch + COLSUM_ROWS - 1) / COLSUM_ROWS;
    // dW2[H,S] = H1^T dO;  dH1 = dO W2^T (masked by H1 > 0, db1 = column sums)
    METRPO_BLAS_OK(gemm_rm(h, true, false, d.H, d.S, batch, h->H1, d.H, sH, h->O, d.Sp, sO, h->grad + d.oW2, d.Sp, P));
    METRPO_BLAS_OK(gemm_rm(h, false, true, batch, d.H, d.S, h->O, d.Sp, sO, h->theta + d.oW2, d.Sp, P, h->D1, d.H, sH));
    fit_relu_bwd_colsum_kernel<<<dim3(d.H / 32, nslab, d.K), dim3(8, 32), 0, st>>>(h->D1, h->H1, batch, d.H, sH, h->part1);
    // dW1 = H0^T dH1;  dH0 = dH1 W1^T -> H1's buffer (H1 is dead now), masked by H0 > 0, db0
    METRPO_BLAS_OK(gemm_rm(h, true, false, d.H, d.H, batch, h->H0, d.H, sH, h->D1, d.H, sH, h->grad + d.oW1, d.H, P));
    METRPO_BLAS_OK(gemm_rm(h, false, true, batch, d.H, d.H, h->D1, d.H, sH, h->theta + d.oW1, d.H, P, h->H1, d.H, sH));
    fit_relu_bwd_colsum_kernel<<<dim3(d.H / 32, nslab, d.K), dim3(8, 32), 0, st>>>(h->H1, h->H0, batch, d.H, sH, h->part0);
    // dW0[Din,H] = Z^T dH0
    METRPO_BLAS_OK(gemm_rm(h, true, false, d.Din, d.H, batch, h->Z, d.Dp, sZ, h->H1, d.H, sH, h->grad + d.oW0, d.H, P));
    launches += 7;
  }
  fit_bias_grad_finish_kernel<<<dim3((2 * d.H + d.S + (w0_splits > 1 ? d.Din * d.H : 0) + 31) / 32, d.K), 256, 0, st>>>(
      d, h->part1, h->part0, nslab, h->part2, mb, w0_splits > 1 ? h->w0part : nullptr, w0_splits,
      static_cast<long long>(d.K) * P0, P0, h->grad);
  // Adam (tf.train.AdamOptimizer defaults beta1 0.9, beta2 0.999, epsilon 1e-8)
  const long long n4 = P * d.K / 4;
  fit_adam_kernel<<<static_cast<int>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, st>>>(
      h->theta, h->grad, h->m, h->v, n4, lr_t, 0.9f, 0.999f, 1e-8f, dargs);
  launches += 3;
  if (losses) { fit_losses_out_kernel<<<1, 64, 0, st>>>(h->loss_acc, losses, d.K); ++launches; }
  METRPO_CUDA_OK(cudaGetLastError());
  h->last_launches = launches;
  return METRPO_OK;
}

extern "C" int metrpo_fit_step(metrpo_fit_t* h, const float* x, const float* y, int n_data, const int32_t* idx,
                               int batch, uint64_t seed, uint64_t offset, double lr, float* losses,
                               void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_step: null handle");
  if (!x || !y || n_data < 1) return set_error(METRPO_ERR_INVALID, "fit_step: x, y and n_data >= 1 are required");
  if (batch < 1 || batch > h->R) return set_error(METRPO_ERR_INVALID, "fit_step: batch must be in [1, max_rows=%d]", h->R);
  int rc = fit_check_ready(h, "fit_step");
  if (rc != METRPO_OK) return rc;
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  h->adam_t += 1;
  const double b1 = 0.9, b2 = 0.999;
  const float lr_t = static_cast<float>(lr * std::sqrt(1.0 - std::pow(b2, (double)h->adam_t)) / (1.0 - std::pow(b1, (double)h->adam_t)));

  // Replay path: the 13 launches + 4 event edges of an iteration as ONE graph launch.  The first call
  // with a given (x, y, n_data, batch) runs eagerly, the second captures, later ones replay; the
  // scalars that change per step (Philox seed / offset, lr_t) travel through h->dargs.
  const bool graphable = h->use_graph && h->cfg.precision == METRPO_FIT_TF32 && idx == nullptr;
  if (graphable) {
    if (h->gx != x || h->gy != y || h->gn != n_data || h->gbatch != batch) {
      if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
      h->gx = x; h->gy = y; h->gn = n_data; h->gbatch = batch; h->gwarm = 0;
    }
    FitStepArgs ha;
    ha.seed = seed; ha.offset = offset; ha.lr_t = lr_t; ha.pad = 0;
    METRPO_CUDA_OK(cudaMemcpyAsync(h->dargs, &ha, sizeof(ha), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    if (h->gwarm >= 1 && !h->gexec) {
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = fit_step_issue(h, x, y, n_data, nullptr, batch, seed, offset, lr_t, nullptr, h->cap, h->dargs);
        const cudaError_t ce = cudaStreamEndCapture(h->cap, &graph);
        if (rc == METRPO_OK && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&h->gexec, graph, 0) == cudaSuccess) {
          h->glaunches = h->last_launches;
        } else {
          h->gexec = nullptr; h->use_graph = 0;     // capture not possible here: stay on the eager path
          cudaGetLastError();
        }
        if (graph) cudaGraphDestroy(graph);
      } else {
        h->use_graph = 0;
        cudaGetLastError();
      }
    }
    if (h->gexec) {
      METRPO_CUDA_OK(cudaGraphLaunch(h->gexec, st));
      h->last_launches = h->glaunches;
      if (losses) {   // the caller's loss buffer changes from call to call: outside the graph
        fit_losses_out_kernel<<<1, 64, 0, st>>>(h->loss_acc, losses, h->d.K);
        METRPO_CUDA_OK(cudaGetLastError());
        ++h->last_launches;
      }
      return METRPO_OK;
    }
    h->gwarm += 1;
    return fit_step_issue(h, x, y, n_data, nullptr, batch, seed, offset, lr_t, losses, st, h->dargs);
  }
  return fit_step_issue(h, x, y, n_data, idx, batch, seed, offset, lr_t, losses, st, nullptr);
}

extern "C" int metrpo_fit_eval(metrpo_fit_t* h, const float* x, const float* y, int n, int snapshot, float* losses,
                               uint8_t* improved, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_eval: null handle");
  if (!x || !y || n < 1) return set_error(METRPO_ERR_INVALID, "fit_eval: x, y and n >= 1 are required");
  if (snapshot < 0 || snapshot > 2) return set_error(METRPO_ERR_INVALID, "fit_eval: snapshot must be 0, 1 or 2");
  if (h->d.K > 64) return set_error(METRPO_ERR_UNSUPPORTED, "fit_eval: K <= 64");
  int rc = fit_check_ready(h, "fit_eval");
  if (rc != METRPO_OK) return rc;
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->cfg.precision != METRPO_FIT_TF32) METRPO_BLAS_OK(cublasSetStream(h->blas, st));
  const FitDims& d = h->d;
  const long long sS = static_cast<long long>(h->R) * d.S, sO = static_cast<long long>(h->R) * d.Sp;
  int launches = 0;
  METRPO_CUDA_OK(cudaMemsetAsync(h->loss_acc, 0, d.K * 8, st));
  for (int row0 = 0; row0 < n; row0 += h->R) {
    const int rows = std::min(h->R, n - row0);
    rc = fit_forward(h, x, y, n, nullptr, 1, row0, 0, 0, rows, st, launches);
    if (rc != METRPO_OK) return rc;
    const int mb = std::min((rows + 7) / 8, 148);
    fit_mse_kernel<<<dim3(mb, d.K), 256, 8 * d.S * 4, st>>>(d, h->O, h->XS, h->Y, h->theta, h->norm, rows, 1.0 / n, 0, sS,
                                                       sO, h->o_splits, static_cast<long long>(d.K) * sO, h->loss_acc,
                                                       h->part2);
    ++launches;
  }
  fit_snapshot_flags_kernel<<<1, 64, 0, st>>>(h->loss_acc, h->min_losses, h->flags, losses, d.K, snapshot);
  ++launches;
  if (snapshot) {
    fit_snapshot_copy_kernel<<<dim3(64, d.K), 256, 0, st>>>(h->theta, h->best, h->flags, d.P / 4);
    ++launches;
  }
  if (improved) METRPO_CUDA_OK(cudaMemcpyAsync(improved, h->flags, d.K, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaGetLastError());
  h->last_launches = launches;
  return METRPO_OK;
}

extern "C" int metrpo_fit_restore_best(metrpo_fit_t* h, void* stream) {
  if (!h) return set_error(METRPO_ERR_INVALID, "fit_restore_best: null handle");
  METRPO_CUDA_OK(cudaMemcpyAsync(h->theta, h->best, static_cast<size_t>(h->d.P) * h->d.K * 4,
                                 cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return METRPO_OK;
}
