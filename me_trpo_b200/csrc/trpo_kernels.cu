// TRPO half of the hot path on the device (include/metrpo.h metrpo_trpo_*):
//
//   metrpo_trpo_process       BaseSampler.process_samples (samplers/base.py:48-105): baseline
//                             prediction, TD residuals, discounted reverse scans, advantage centring,
//                             on the time-major [T,B] buffers of the fused sampler (no path lists)
//   metrpo_trpo_fit_baseline  rllab LinearFeatureBaseline.fit (samplers/base.py:167): normal
//                             equations of the ridge regression + dense solve
//   metrpo_trpo_update        NPO.optimize_policy (algos/npo.py:94-111) -> rllab
//                             ConjugateGradientOptimizer.optimize: surrogate gradient, 10 CG
//                             iterations on Fisher-vector products, step scaling, back-tracking
//                             line search -- every scalar stays on the device
//
// These are HBM-bound scans / reductions plus small fp32 MLP math (P ~ 2k parameters) on CUDA
// cores; nothing here is GEMM-shaped enough for the tensor pipe and it is not forced into one.
// One thread owns one sample; activations of a 128-sample tile live in shared memory
// ([feature][sample], conflict-free), weights are broadcast from shared memory, parameter
// gradients are tile-level outer products accumulated per block and flushed once with fp64 atomics.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace metrpo {

constexpr int TP_MAXL = METRPO_MAX_POLICY_LAYERS;
constexpr int NTPAD = 4;   // row padding of the [feature][sample] tiles (keeps float4 alignment)

struct PolDims {
  int L;                    // weight layers
  int d[TP_MAXL + 1];       // S, hidden.., A
  int w_off[TP_MAXL];       // flat offsets (rllab get_params order: W0,b0,W1,b1,..,log_std)
  int b_off[TP_MAXL];
  int np[TP_MAXL];          // padded row stride of layer l in shared memory (multiple of 4)
  int sw_off[TP_MAXL];      // padded smem offsets of W_l (rows) and b_l
  int sb_off[TP_MAXL];
  int act_row[TP_MAXL + 1]; // first row of layer l's activations in the activation tile
  int logstd_off, P, P_pad, sum_d, max_d, out_tanh;
};

enum { MODE_LOSS = 0, MODE_GRAD = 1, MODE_FVP = 2 };
// accumulator layout (double): [0,P) parameter sums | P: sum lr*adv | P+1: sum kl | P+2: count
constexpr int ACC_EXTRA = 4;

struct PassParams {
  PolDims pd;
  const float* theta;
  const float* vec;
  const float* obs;
  const float* act;
  const float* adv;
  const float* old_mean;
  const float* old_log_std;
  int old_ls_stride;
  const uint8_t* valid;
  long long N;
  double* acc;
  const int* skip_flag;
  float* act_cache;        // tiled pass only, optional: hidden activations per 128-sample tile, [tile][hidden row][128].
                           // The GRAD pass writes it, the Fisher-vector passes of the same update (same theta) read it
                           // instead of recomputing the forward pass
};

// tanh(x) = 1 - 2/(exp(2x)+1) with ex2.approx / rcp.approx: abs error ~1e-7, ~6 instructions (the
// same formula the rollout kernel uses for the policy it samples with, csrc/rollout_kernel.cuh)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[j][n] = act( b[j] + sum_i in[i][n] * W[i][j] )  for this thread's sample n
template <int LD>
__device__ __forceinline__ void layer_forward(const float* __restrict__ W, const float* __restrict__ b,
                                              int nin, int nout, int np, const float* in, float* out,
                                              int n, bool use_tanh) {
  for (int j0 = 0; j0 < nout; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = (j0 + q < nout) ? b[j0 + q] : 0.f;
#pragma unroll 4
    for (int i = 0; i < nin; ++i) {
      const float x = in[i * LD + n];
      const float4* w4 = reinterpret_cast<const float4*>(W + i * np + j0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j0 + 4 * q < np) {
          const float4 w = w4[q];
          acc[4 * q + 0] = fmaf(x, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(x, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(x, w.w, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (j0 + q < nout) out[(j0 + q) * LD + n] = use_tanh ? tanh_fast(acc[q]) : acc[q];
  }
}

// forward-mode tangent: out[j] = (vb[j] + sum_i vW[i][j]*a[i] + W[i][j]*tin[i]) * dact(aout[j])
template <int LD>
__device__ __forceinline__ void layer_tangent(const float* __restrict__ W, const float* __restrict__ vW,
                                              const float* __restrict__ vb, int nin, int nout, int np,
                                              const float* a_in, const float* t_in, const float* a_out,
                                              float* t_out, int n, bool use_tanh) {
  for (int j0 = 0; j0 < nout; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = (j0 + q < nout) ? vb[j0 + q] : 0.f;
#pragma unroll 2
    for (int i = 0; i < nin; ++i) {
      const float x = a_in[i * LD + n];
      const float tx = t_in ? t_in[i * LD + n] : 0.f;
      const float4* v4 = reinterpret_cast<const float4*>(vW + i * np + j0);
      const float4* w4 = reinterpret_cast<const float4*>(W + i * np + j0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j0 + 4 * q < np) {
          const float4 v = v4[q];
          acc[4 * q + 0] = fmaf(x, v.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(x, v.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x, v.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(x, v.w, acc[4 * q + 3]);
          if (t_in) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(tx, w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(tx, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(tx, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(tx, w.w, acc[4 * q + 3]);
          }
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (j0 + q < nout) {
        const float ao = a_out[(j0 + q) * LD + n];
        t_out[(j0 + q) * LD + n] = use_tanh ? acc[q] * (1.f - ao * ao) : acc[q];
      }
  }
}

// d_in[i][n] = (sum_j W[i][j] * d_out[j][n]) * (1 - a_in[i][n]^2)
template <int LD>
__device__ __forceinline__ void layer_backward_data(const float* __restrict__ W, int nin, int nout, int np,
                                                    const float* d_out, const float* a_in, float* d_in,
                                                    int n) {
  for (int j0 = 0; j0 < nout; j0 += 16) {
    float d[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) d[q] = (j0 + q < nout) ? d_out[(j0 + q) * LD + n] : 0.f;
#pragma unroll 4
    for (int i = 0; i < nin; ++i) {
      const float4* w4 = reinterpret_cast<const float4*>(W + i * np + j0);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j0 + 4 * q < np) {
          const float4 w = w4[q];
          s = fmaf(w.x, d[4 * q + 0], s);
          s = fmaf(w.y, d[4 * q + 1], s);
          s = fmaf(w.z, d[4 * q + 2], s);
          s = fmaf(w.w, d[4 * q + 3], s);
        }
      if (j0 > 0) s += d_in[i * LD + n];
      if (j0 + 16 >= nout) {
        const float a = a_in[i * LD + n];
        s *= (1.f - a * a);
      }
      d_in[i * LD + n] = s;
    }
  }
}

// sG[e] += sum_n rowA(i)[n] * rowD(j)[n] for the E = (nin+1)*nout entries of (W_l, b_l); the
// bias entries use the constant-one row.  Entry e is owned by thread e % NT: no atomics.
template <int NT, int LD>
__device__ __forceinline__ void accumulate_outer(float* sG, int nin, int nout, const float* a_in,
                                                 const float* d_out, const float* ones, int tid) {
  const int E = (nin + 1) * nout;
  for (int e = tid; e < E; e += NT) {
    const int i = e / nout, j = e - i * nout;
    const float4* ar = reinterpret_cast<const float4*>(i < nin ? a_in + i * LD : ones);
    const float4* dr = reinterpret_cast<const float4*>(d_out + j * LD);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int q = 0; q < NT / 4; q += 2) {
      const float4 a = ar[q], d = dr[q];
      const float4 a2 = ar[q + 1], d2 = dr[q + 1];
      s0 = fmaf(a.x, d.x, s0); s0 = fmaf(a.y, d.y, s0); s0 = fmaf(a.z, d.z, s0); s0 = fmaf(a.w, d.w, s0);
      s1 = fmaf(a2.x, d2.x, s1); s1 = fmaf(a2.y, d2.y, s1); s1 = fmaf(a2.z, d2.z, s1); s1 = fmaf(a2.w, d2.w, s1);
    }
    sG[e] += s0 + s1;
  }
}

// One pass over all samples.  MODE_LOSS: sum lr*adv, sum kl.  MODE_GRAD: + gradient of
// -sum(lr*adv).  MODE_FVP: sum_n J^T M J vec (Gauss-Newton form of the KL Hessian at old == new,
// equal to the Perlmutter double-backprop the reference uses; log_std block added by the caller).
template <int MODE, int NT>
__global__ void __launch_bounds__(NT) policy_pass_kernel(const __grid_constant__ PassParams p) {
  constexpr int LD = NT + NTPAD;
  extern __shared__ __align__(16) float sm[];
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
  const PolDims& pd = p.pd;
  const int tid = threadIdx.x, L = pd.L, A = pd.d[L], S = pd.d[0];
  float* sW = sm;
  float* sV = sW + pd.P_pad;                                  // FVP only
  float* sG = sV + (MODE == MODE_FVP ? pd.P_pad : 0);        // GRAD / FVP: flat [P] accumulators
  float* sAct = sG + (MODE == MODE_LOSS ? 0 : ((pd.P + 3) & ~3));
  float* sBufA = sAct + pd.sum_d * LD;
  float* sBufB = sBufA + pd.max_d * LD;
  float* sOnes = sBufB + pd.max_d * LD;
  __shared__ double sRed[3][32];
  __shared__ float sLs[32];   // raw log_std parameters

  // stage parameters into the padded layout
  for (int i = tid; i < pd.P_pad; i += NT) { sW[i] = 0.f; if (MODE == MODE_FVP) sV[i] = 0.f; }
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int nin = pd.d[l], nout = pd.d[l + 1];
    for (int e = tid; e < nin * nout; e += NT) {
      const int i = e / nout, j = e - i * nout;
      sW[pd.sw_off[l] + i * pd.np[l] + j] = p.theta[pd.w_off[l] + e];
      if (MODE == MODE_FVP) sV[pd.sw_off[l] + i * pd.np[l] + j] = p.vec[pd.w_off[l] + e];
    }
    for (int j = tid; j < nout; j += NT) {
      sW[pd.sb_off[l] + j] = p.theta[pd.b_off[l] + j];
      if (MODE == MODE_FVP) sV[pd.sb_off[l] + j] = p.vec[pd.b_off[l] + j];
    }
  }
  if (MODE != MODE_LOSS)
    for (int i = tid; i < pd.P; i += NT) sG[i] = 0.f;
  for (int i = tid; i < LD; i += NT) sOnes[i] = 1.f;
  if (tid < 32) sLs[tid] = tid < A ? p.theta[pd.logstd_off + tid] : 0.f;
  __syncthreads();

  double t_surr = 0.0, t_kl = 0.0, t_cnt = 0.0;
  const long long n_tiles = (p.N + NT - 1) / NT;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n0 = tile * NT, ng = n0 + tid;
    const bool inb = ng < p.N;
    const bool ok = inb && (p.valid == nullptr || p.valid[ng] != 0);
    // observations: coalesced read of the tile's [NT,S] block, stored [feature][sample]
    {
      const long long base = n0 * S, lim = p.N * S;
      for (int q = tid; q < NT * S; q += NT) {
        const int smp = q / S, f = q - smp * S;
        sAct[f * LD + smp] = (base + q < lim) ? p.obs[base + q] : 0.f;
      }
    }
    __syncthreads();
    // ---- forward (training.py:99-103) ----
    for (int l = 0; l < L; ++l)
      layer_forward<LD>(sW + pd.sw_off[l], sW + pd.sb_off[l], pd.d[l], pd.d[l + 1], pd.np[l],
                        sAct + pd.act_row[l] * LD, sAct + pd.act_row[l + 1] * LD, tid,
                        (l < L - 1) || pd.out_tanh);
    const float* mu = sAct + pd.act_row[L] * LD;
    float* dOut = sBufA;   // delta at the output pre-activation, [A][LD]
    float* cLs = sBufB;    // GRAD: per-sample d(-lr*adv)/d log_std_a, [A][LD]

    if (MODE == MODE_FVP) {
      // tangent forward through the mean network
      const float* tin = nullptr;
      float* bufs[2] = {sBufA, sBufB};
      for (int l = 0; l < L; ++l) {
        float* tout = bufs[l & 1];
        layer_tangent<LD>(sW + pd.sw_off[l], sV + pd.sw_off[l], sV + pd.sb_off[l], pd.d[l], pd.d[l + 1],
                          pd.np[l], sAct + pd.act_row[l] * LD, tin, sAct + pd.act_row[l + 1] * LD, tout,
                          tid, (l < L - 1) || pd.out_tanh);
        tin = tout;
      }
      dOut = bufs[(L - 1) & 1];
      // delta = M * mu_dot with M = d^2 kl / d mu^2 = 2 / (2 sigma^2 + 1e-8)   (kl_sym, A.3)
      for (int a = 0; a < A; ++a) {
        const float ls = fmaxf(sLs[a], -13.815510557964274f);
        const float sg2 = __expf(2.f * ls);
        float dv = dOut[a * LD + tid] * (2.f / (2.f * sg2 + 1e-8f));
        if (pd.out_tanh) { const float m = mu[a * LD + tid]; dv *= (1.f - m * m); }
        dOut[a * LD + tid] = ok ? dv : 0.f;
      }
      if (ok) t_cnt += 1.0;
    } else {
      // likelihood ratio and KL of this sample (DiagonalGaussian, A.3)
      float ll_new = 0.f, ll_old = 0.f, kl = 0.f;
      float zn[24];
      for (int a = 0; a < A; ++a) {
        const float m = mu[a * LD + tid];
        const float x = inb ? p.act[ng * A + a] : 0.f;
        const float om = inb ? p.old_mean[ng * A + a] : 0.f;
        const float ols = p.old_log_std[((inb && p.old_ls_stride) ? ng * p.old_ls_stride : 0) + a];
        const float ls = fmaxf(sLs[a], -13.815510557964274f);   // min_std 1e-6
        const float sg = expf(ls), osg = expf(ols);
        const float z = (x - m) / sg, zo = (x - om) / osg;
        ll_new += -ls - 0.5f * z * z;
        ll_old += -ols - 0.5f * zo * zo;
        kl += ((om - m) * (om - m) + osg * osg - sg * sg) / (2.f * sg * sg + 1e-8f) + ls - ols;
        if (a < 24) zn[a] = z;
      }
      const float lr = expf(ll_new - ll_old);
      const float ad = inb ? p.adv[ng] : 0.f;
      if (ok) { t_surr += static_cast<double>(lr) * ad; t_kl += kl; t_cnt += 1.0; }
      if (MODE == MODE_GRAD) {
        const float c = ok ? -ad * lr : 0.f;       // d(-lr*adv)/d ll_new
        for (int a = 0; a < A; ++a) {
          const float lsr = sLs[a];
          const float ls = fmaxf(lsr, -13.815510557964274f);
          const float sg = expf(ls);
          const float z = zn[a < 24 ? a : 23];
          float dv = c * z / sg;                   // d ll / d mu = z / sigma
          if (pd.out_tanh) { const float m = mu[a * LD + tid]; dv *= (1.f - m * m); }
          dOut[a * LD + tid] = dv;
          cLs[a * LD + tid] = (lsr > -13.815510557964274f) ? c * (z * z - 1.f) : 0.f;   // d ll / d log_std
        }
      }
    }

    if (MODE != MODE_LOSS) {
      if (MODE == MODE_GRAD) {
        __syncthreads();
        // log_std entries: sum_n cLs[a][n]
        for (int a = tid; a < A; a += NT) {
          const float4* r4 = reinterpret_cast<const float4*>(cLs + a * LD);
          float s = 0.f;
          for (int q = 0; q < NT / 4; ++q) { const float4 v = r4[q]; s += (v.x + v.y) + (v.z + v.w); }
          sG[pd.logstd_off + a] += s;
        }
      }
      // ---- backward: deltas ping-pong between the two buffers ----
      float* dcur = dOut;
      for (int l = L - 1; l >= 0; --l) {
        __syncthreads();   // dcur rows (all samples) complete
        accumulate_outer<NT, LD>(sG + pd.w_off[l], pd.d[l], pd.d[l + 1], sAct + pd.act_row[l] * LD, dcur,
                                 sOnes, tid);
        if (l > 0) {
          float* dnext = (dcur == sBufA) ? sBufB : sBufA;
          __syncthreads();   // everyone finished reading dnext's previous contents (cLs / older delta)
          layer_backward_data<LD>(sW + pd.sw_off[l], pd.d[l], pd.d[l + 1], pd.np[l], dcur,
                                  sAct + pd.act_row[l] * LD, dnext, tid);
          dcur = dnext;
        }
      }
    }
    __syncthreads();   // tile buffers are reused by the next tile
  }

  // ---- flush ----
  if (MODE != MODE_LOSS)
    for (int i = tid; i < pd.P; i += NT)
      if (sG[i] != 0.f) atomicAdd(&p.acc[i], static_cast<double>(sG[i]));
  t_surr = warp_sum(t_surr); t_kl = warp_sum(t_kl); t_cnt = warp_sum(t_cnt);
  if ((tid & 31) == 0) { sRed[0][tid >> 5] = t_surr; sRed[1][tid >> 5] = t_kl; sRed[2][tid >> 5] = t_cnt; }
  __syncthreads();
  if (tid < 3) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += sRed[tid][w];
    if (s != 0.0) atomicAdd(&p.acc[pd.P + tid], s);
  }
}


}  // namespace metrpo
#include "trpo_mma.cuh"
#include "trpo_tiled.cuh"
namespace metrpo {

// ---------------------------------------------------------------------------------------------
// CG / line-search controller: single-CTA kernels over P-vectors in double (rllab keeps the CG
// vectors in float64 and hands float32 parameters / directions to the graph; same here).
// cg buffer: g | x | r | p | prev | descent (6P doubles) then scalars.
// ---------------------------------------------------------------------------------------------
enum { CGS_RR = 0, CGS_LOSS_BEFORE, CGS_LAST_LOSS, CGS_LAST_KL, CGS_NITER, CGS_STEP0, CGS_N, CGS_DHD, CGS_COUNT = 16 };
enum { FLAG_CG_DONE = 0, FLAG_ACCEPTED = 1 };
constexpr int CTRL_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < CTRL_THREADS / 32; ++w) s += sred[w];
  return s;
}

// d^2 kl / d log_std_a^2 at old == new (kl_sym with its 1e-8, A.3); 0 when the min_std clamp is active
__device__ __forceinline__ double logstd_hess(double ls_raw) {
  if (ls_raw <= -13.815510557964274) return 0.0;
  const double s = exp(2.0 * ls_raw), eps = 1e-8;
  return 4.0 * s * (2.0 * s - eps) / ((2.0 * s + eps) * (2.0 * s + eps));
}

__global__ void __launch_bounds__(CTRL_THREADS) k_grad_finish(int P, double* acc, const float* theta,
                                                              double* cg, float* vec_f, int* flags) {
  __shared__ double sred[CTRL_THREADS / 32];
  double* g = cg; double* x = cg + P; double* r = cg + 2 * P; double* pv = cg + 3 * P;
  double* prev = cg + 4 * P; double* sc = cg + 6 * P;
  const double N = acc[P + 2];
  double rr = 0.0;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) {
    const double ge = acc[e] / N;
    g[e] = ge; x[e] = 0.0; r[e] = ge; pv[e] = ge; prev[e] = static_cast<double>(theta[e]);
    vec_f[e] = static_cast<float>(ge);
    rr += ge * ge;
  }
  rr = block_sum(rr, sred);
  if (threadIdx.x == 0) {
    sc[CGS_RR] = rr;
    sc[CGS_LOSS_BEFORE] = -acc[P] / N;
    sc[CGS_LAST_LOSS] = sc[CGS_LOSS_BEFORE];
    sc[CGS_LAST_KL] = acc[P + 1] / N;
    sc[CGS_NITER] = 0.0; sc[CGS_N] = N;
    flags[FLAG_CG_DONE] = 0; flags[FLAG_ACCEPTED] = 0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < P + ACC_EXTRA; e += CTRL_THREADS) acc[e] = 0.0;
}

// krylov.cg body (A.2): z = Hp; v = rr / p.z; x += v p; r -= v z; p = r + (rr'/rr) p
__global__ void __launch_bounds__(CTRL_THREADS) k_cg_step(int P, int logstd_off, int A, double* acc, double* cg,
                                                          float* vec_f, int* flags, double reg, int last) {
  __shared__ double sred[CTRL_THREADS / 32];
  double* x = cg + P; double* r = cg + 2 * P; double* pv = cg + 3 * P; double* prev = cg + 4 * P;
  double* z = cg + 5 * P;   // the descent slot doubles as z during CG
  double* sc = cg + 6 * P;
  if (flags[FLAG_CG_DONE]) return;   // uniform
  const double N = sc[CGS_N];
  double pz = 0.0;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) {
    double ze = acc[e] / N;
    if (e >= logstd_off && e < logstd_off + A) ze = logstd_hess(prev[e]) * pv[e];
    ze += reg * pv[e];
    z[e] = ze;
    pz += pv[e] * ze;
  }
  pz = block_sum(pz, sred);
  const double rr = sc[CGS_RR];
  const double v = rr / pz;
  double nrr = 0.0;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) {
    x[e] += v * pv[e];
    const double re = r[e] - v * z[e];
    r[e] = re;
    nrr += re * re;
  }
  nrr = block_sum(nrr, sred);
  const double mu = nrr / rr;
  const bool done = nrr < 1e-10;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) {
    const double pe = r[e] + mu * pv[e];
    pv[e] = pe;
    vec_f[e] = static_cast<float>((last || done) ? x[e] : pe);
    acc[e] = 0.0;
  }
  if (threadIdx.x < ACC_EXTRA) acc[P + threadIdx.x] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    sc[CGS_RR] = nrr;
    if (done) flags[FLAG_CG_DONE] = 1;
  }
}

// initial_step_size = sqrt(2 max_kl / (d.Hd + 1e-8)); descent = step0 * d   (A.2)
//
// from_residual != 0: d.Hd is taken from the conjugate-gradient recurrence instead of one more
// Fisher-vector pass over all samples: CG maintains r = g - H x, so d.(H d) = d.(g - r) with the same
// regularised operator (rllab evaluates Hx(d) explicitly; the two agree to the rounding of the fp64
// recurrence, ~1e-7 relative, far below the fp32 noise of a Fisher-vector pass).
__global__ void __launch_bounds__(CTRL_THREADS) k_step_finish(int P, int logstd_off, int A, double* acc,
                                                              double* cg, double reg, double max_kl,
                                                              int from_residual) {
  __shared__ double sred[CTRL_THREADS / 32];
  double* x = cg + P; double* prev = cg + 4 * P; double* desc = cg + 5 * P; double* sc = cg + 6 * P;
  const double* g = cg; const double* r = cg + 2 * P;
  const double N = sc[CGS_N];
  double dhd = 0.0;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) {
    double ze;
    if (from_residual) {
      ze = g[e] - r[e];
    } else {
      ze = acc[e] / N;
      if (e >= logstd_off && e < logstd_off + A) ze = logstd_hess(prev[e]) * x[e];
      ze += reg * x[e];
    }
    dhd += x[e] * ze;
  }
  dhd = block_sum(dhd, sred);
  double step0 = sqrt(2.0 * max_kl * (1.0 / (dhd + 1e-8)));
  if (isnan(step0)) step0 = 1.0;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) { desc[e] = step0 * x[e]; acc[e] = 0.0; }
  if (threadIdx.x < ACC_EXTRA) acc[P + threadIdx.x] = 0.0;
  if (threadIdx.x == 0) { sc[CGS_STEP0] = step0; sc[CGS_DHD] = dhd; }
}

__global__ void __launch_bounds__(CTRL_THREADS) k_ls_prepare(int P, const double* cg, float* trial_f,
                                                             const int* flags, double ratio) {
  if (flags[FLAG_ACCEPTED]) return;
  const double* prev = cg + 4 * P; const double* desc = cg + 5 * P;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) trial_f[e] = static_cast<float>(prev[e] - ratio * desc[e]);
}

__global__ void __launch_bounds__(CTRL_THREADS) k_ls_check(int P, double* acc, double* cg, int* flags, int k,
                                                           double max_kl) {
  double* sc = cg + 6 * P;
  if (!flags[FLAG_ACCEPTED] && threadIdx.x == 0) {
    const double N = sc[CGS_N];
    const double loss = -acc[P] / N, kl = acc[P + 1] / N;
    sc[CGS_LAST_LOSS] = loss; sc[CGS_LAST_KL] = kl; sc[CGS_NITER] = k;
    if (loss < sc[CGS_LOSS_BEFORE] && kl <= max_kl) flags[FLAG_ACCEPTED] = 1;
  }
  __syncthreads();
  if (threadIdx.x < ACC_EXTRA) acc[P + threadIdx.x] = 0.0;
}

__global__ void __launch_bounds__(CTRL_THREADS) k_finalize(int P, const double* cg, const float* trial_f,
                                                           float* theta_out, const int* flags, double max_kl,
                                                           double* info) {
  const double* prev = cg + 4 * P; const double* sc = cg + 6 * P;
  const double loss = sc[CGS_LAST_LOSS], kl = sc[CGS_LAST_KL];
  const bool reject = isnan(loss) || isnan(kl) || loss >= sc[CGS_LOSS_BEFORE] || kl >= max_kl;
  for (int e = threadIdx.x; e < P; e += CTRL_THREADS) theta_out[e] = reject ? static_cast<float>(prev[e]) : trial_f[e];
  if (threadIdx.x == 0 && info != nullptr) {
    info[0] = sc[CGS_LOSS_BEFORE]; info[1] = loss; info[2] = kl; info[3] = sc[CGS_NITER];
    info[4] = reject ? 0.0 : 1.0; info[5] = sc[CGS_STEP0]; info[6] = sc[CGS_N]; info[7] = sc[CGS_DHD];
  }
}

// ---------------------------------------------------------------------------------------------
// process_samples on time-major buffers
// ---------------------------------------------------------------------------------------------
// position of every sample inside its path (forward scan per row); rows start a fresh path at t=0
__global__ void pos_scan_kernel(const uint8_t* __restrict__ done, int* __restrict__ pos, int T, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int pcur = 0;
  for (int t = 0; t < T; ++t) {
    const size_t o = static_cast<size_t>(t) * B + b;
    pos[o] = pcur;
    pcur = done[o] ? 0 : pcur + 1;
  }
}

// rllab LinearFeatureBaseline features (A.4): [o, o^2, al, al^2, al^3, 1], o = clip(obs,-10,10), al = pos/100
__device__ __forceinline__ double baseline_value(const float* __restrict__ o, int pos, const double* __restrict__ c, int S) {
  double v = 0.0;
  for (int s = 0; s < S; ++s) {
    const double x = fmin(fmax(static_cast<double>(o[s]), -10.0), 10.0);
    v += c[s] * x + c[S + s] * x * x;
  }
  const double al = pos / 100.0;
  return v + c[2 * S] * al + c[2 * S + 1] * al * al + c[2 * S + 2] * al * al * al + c[2 * S + 3];
}
__global__ void baseline_predict_kernel(const float* __restrict__ obs, const int* __restrict__ pos,
                                        const double* __restrict__ coeffs, double* __restrict__ base,
                                        long long N, int S) {
  extern __shared__ double sc[];
  for (int i = threadIdx.x; i < 2 * S + 4; i += blockDim.x) sc[i] = coeffs[i];
  __syncthreads();
  for (long long n = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; n < N;
       n += static_cast<long long>(gridDim.x) * blockDim.x)
    base[n] = baseline_value(obs + n * S, pos[n], sc, S);
}

// reverse scan per row (samplers/base.py:55-62): delta_t = r_t + g*b_{t+1} - b_t (b = 0 past the path
// end), adv = discount_cumsum(delta, g*lam), ret = discount_cumsum(r, g).  Samples after the last
// done of a row belong to an unfinished path and are marked invalid (obtain_samples returns only
// completed paths).
__global__ void gae_scan_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                const double* __restrict__ base, double gamma, double lam,
                                double* __restrict__ adv_raw, float* __restrict__ ret,
                                uint8_t* __restrict__ valid, int T, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double a = 0.0, rsum = 0.0, next_b = 0.0;
  bool ok = false;
  for (int t = T - 1; t >= 0; --t) {
    const size_t o = static_cast<size_t>(t) * B + b;
    if (done[o]) { a = 0.0; rsum = 0.0; next_b = 0.0; ok = true; }
    const double r = rew[o], bt = base ? base[o] : 0.0;
    const double delta = r + gamma * next_b - bt;
    a = delta + gamma * lam * a;
    rsum = r + gamma * rsum;
    next_b = bt;
    adv_raw[o] = ok ? a : 0.0;
    ret[o] = ok ? static_cast<float>(rsum) : 0.f;
    valid[o] = ok ? 1 : 0;
  }
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) <= v) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
}
// stats: [0] count [1] sum [2] sumsq [3] min [4] mean [5] std
__global__ void moments_kernel(const double* __restrict__ adv_raw, const uint8_t* __restrict__ valid,
                               long long N, double* stats) {
  __shared__ double sred[4][32];
  double c = 0.0, s = 0.0, q = 0.0, mn = 1e300;
  for (long long n = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; n < N;
       n += static_cast<long long>(gridDim.x) * blockDim.x)
    if (valid[n]) { const double v = adv_raw[n]; c += 1.0; s += v; q += v * v; mn = fmin(mn, v); }
  c = warp_sum(c); s = warp_sum(s); q = warp_sum(q);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sred[0][w] = c; sred[1][w] = s; sred[2][w] = q; sred[3][w] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cc = 0, ss = 0, qq = 0, mm = 1e300;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { cc += sred[0][i]; ss += sred[1][i]; qq += sred[2][i]; mm = fmin(mm, sred[3][i]); }
    if (cc > 0) { atomicAdd(&stats[0], cc); atomicAdd(&stats[1], ss); atomicAdd(&stats[2], qq); atomic_min_double(&stats[3], mm); }
  }
}
__global__ void moments_finish_kernel(double* stats) {
  const double n = stats[0];
  const double mean = n > 0 ? stats[1] / n : 0.0;
  const double var = n > 0 ? stats[2] / n - mean * mean : 0.0;
  stats[4] = mean;
  stats[5] = sqrt(fmax(var, 0.0));
}
// center_advantages (A.5): (a - mean) / (std + 1e-8); shift_advantages_to_positive: a - min + 1e-8
__global__ void center_kernel(const double* __restrict__ adv_raw, const uint8_t* __restrict__ valid,
                              const double* __restrict__ stats, int center, int positive,
                              float* __restrict__ adv, long long N) {
  const double mean = stats[4], denom = stats[5] + 1e-8;
  const double mn = center ? (stats[3] - mean) / denom : stats[3];
  for (long long n = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; n < N;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    double v = adv_raw[n];
    if (center) v = (v - mean) / denom;
    if (positive) v = v - mn + 1e-8;
    adv[n] = valid[n] ? static_cast<float>(v) : 0.f;
  }
}

// normal equations of the baseline fit: G[(D+1)][D], rows 0..D-1 = F^T F, row D = F^T returns
template <int NT>
__global__ void __launch_bounds__(NT) gram_kernel(const float* __restrict__ obs, const float* __restrict__ ret,
                                                  const uint8_t* __restrict__ valid, const int* __restrict__ pos,
                                                  long long N, int S, int D, double* G) {
  constexpr int LD = NT + NTPAD;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int E = (D + 1) * D;
  double* sAcc = reinterpret_cast<double*>(smraw);
  float* sF = reinterpret_cast<float*>(smraw + static_cast<size_t>(E) * 8);
  const int tid = threadIdx.x;
  for (int e = tid; e < E; e += NT) sAcc[e] = 0.0;
  const long long n_tiles = (N + NT - 1) / NT;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n = tile * NT + tid;
    const bool ok = n < N && valid[n] != 0;
    __syncthreads();
    {
      const float* o = obs + (ok ? n : 0) * S;
      for (int s = 0; s < S; ++s) {
        const float x = ok ? fminf(fmaxf(o[s], -10.f), 10.f) : 0.f;
        sF[s * LD + tid] = x;
        sF[(S + s) * LD + tid] = x * x;
      }
      const float al = ok ? static_cast<float>(pos[n] / 100.0) : 0.f;
      sF[(2 * S) * LD + tid] = al;
      sF[(2 * S + 1) * LD + tid] = al * al;
      sF[(2 * S + 2) * LD + tid] = al * al * al;
      sF[(2 * S + 3) * LD + tid] = ok ? 1.f : 0.f;
      sF[D * LD + tid] = ok ? ret[n] : 0.f;
    }
    __syncthreads();
    // 4 x 4 register tiles over the (D+1) x D entries (rows strided by tiles_i, columns by tiles_j so
    // that a quarter-warp reads 8 consecutive feature rows: no bank conflicts), reduced over the 128
    // samples with float4 loads along the sample axis: 8 loads per 64 FMAs
    {
      const int tiles_i = (D + 1 + 3) >> 2, tiles_j = (D + 3) >> 2;
      for (int t = tid; t < tiles_i * tiles_j; t += NT) {
        const int ti = t / tiles_j, tj = t - ti * tiles_j;
        const float* ar[4];
        const float* br[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = ti + tiles_i * q, j = tj + tiles_j * q;
          ar[q] = sF + (i <= D ? i : 0) * LD;
          br[q] = sF + (j < D ? j : 0) * LD;
        }
        float2 acc[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) { acc[a][0] = make_float2(0.f, 0.f); acc[a][1] = make_float2(0.f, 0.f); }
#pragma unroll 2
        for (int n4 = 0; n4 < NT / 4; ++n4) {
          float4 av[4], bv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            av[q] = *reinterpret_cast<const float4*>(ar[q] + 4 * n4);
            bv[q] = *reinterpret_cast<const float4*>(br[q] + 4 * n4);
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
              float2 sacc = acc[a][b2];
              sacc = __ffma2_rn(make_float2(av[a].x, av[a].x), make_float2(bv[2 * b2].x, bv[2 * b2 + 1].x), sacc);
              sacc = __ffma2_rn(make_float2(av[a].y, av[a].y), make_float2(bv[2 * b2].y, bv[2 * b2 + 1].y), sacc);
              sacc = __ffma2_rn(make_float2(av[a].z, av[a].z), make_float2(bv[2 * b2].z, bv[2 * b2 + 1].z), sacc);
              sacc = __ffma2_rn(make_float2(av[a].w, av[a].w), make_float2(bv[2 * b2].w, bv[2 * b2 + 1].w), sacc);
              acc[a][b2] = sacc;
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b3 = 0; b3 < 4; ++b3) {
            const int i = ti + tiles_i * a, j = tj + tiles_j * b3;
            if (i <= D && j < D) {
              const float v = (b3 & 1) ? acc[a][b3 >> 1].y : acc[a][b3 >> 1].x;
              sAcc[i * D + j] += static_cast<double>(v);   // entry owned by this thread for the whole kernel
            }
          }
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < E; e += NT)
    if (sAcc[e] != 0.0) atomicAdd(&G[e], sAcc[e]);
}

// coeffs = solve(F^T F + reg I, F^T y), retried with reg *= 10 while NaN (A.4); Gaussian
// elimination with partial pivoting, one CTA, augmented matrix in shared memory
__global__ void __launch_bounds__(128) baseline_solve_kernel(const double* __restrict__ G, int D, double reg,
                                                             double* __restrict__ coeffs) {
  extern __shared__ double sA[];   // [D][D+1]
  __shared__ int s_piv;
  __shared__ int s_bad;
  const int tid = threadIdx.x, W = D + 1;
  for (int attempt = 0; attempt < 5; ++attempt) {
    for (int e = tid; e < D * W; e += blockDim.x) {
      const int i = e / W, j = e - i * W;
      sA[e] = (j < D) ? G[i * D + j] + (i == j ? reg : 0.0) : G[D * D + i];
    }
    __syncthreads();
    for (int k = 0; k < D; ++k) {
      if (tid == 0) {
        int pv = k; double best = fabs(sA[k * W + k]);
        for (int i = k + 1; i < D; ++i) { const double v = fabs(sA[i * W + k]); if (v > best) { best = v; pv = i; } }
        s_piv = pv;
      }
      __syncthreads();
      const int pv = s_piv;
      if (pv != k)
        for (int j = tid; j < W; j += blockDim.x) { const double t = sA[k * W + j]; sA[k * W + j] = sA[pv * W + j]; sA[pv * W + j] = t; }
      __syncthreads();
      const double pivot = sA[k * W + k];
      for (int i = k + 1 + tid; i < D; i += blockDim.x) {
        const double f = sA[i * W + k] / pivot;
        for (int j = k; j < W; ++j) sA[i * W + j] -= f * sA[k * W + j];
      }
      __syncthreads();
    }
    if (tid == 0) {
      int bad = 0;
      for (int i = D - 1; i >= 0; --i) {
        double s = sA[i * W + D];
        for (int j = i + 1; j < D; ++j) s -= sA[i * W + j] * coeffs[j];
        const double c = s / sA[i * W + i];
        coeffs[i] = c;
        if (isnan(c)) bad = 1;
      }
      s_bad = bad;
    }
    __syncthreads();
    if (!s_bad) break;
    reg *= 10.0;
    __syncthreads();
  }
}

}  // namespace metrpo

// =============================================================================================
// C ABI
// =============================================================================================
// One-shot all-reduce over peer memory.  Exchange buffer of a rank: data[2][P2P_NMAX] doubles
// (double-buffered by the parity of the reduction's sequence number) + flags[P2P_MAX_RANKS].
// ---------------------------------------------------------------------------------------------
namespace metrpo {
constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_NMAX = 16384;
constexpr size_t P2P_BYTES = 2 * static_cast<size_t>(P2P_NMAX) * 8 + P2P_MAX_RANKS * 4 + 64;
struct P2PArgs {
  double* data[P2P_MAX_RANKS];
  unsigned* flags[P2P_MAX_RANKS];
  int rank, world;
};
__device__ __forceinline__ unsigned long long p2p_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) k_allreduce_p2p(const P2PArgs a, double* __restrict__ buf, int n, unsigned seq) {
  double* mine = a.data[a.rank] + static_cast<size_t>(seq & 1u) * P2P_NMAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.world) {
    // tell peer `threadIdx.x` that this rank's copy of reduction `seq` is complete ...
    st_release_sys(a.flags[threadIdx.x] + a.rank, seq + 1u);
    // ... and wait until that peer's copy is (bounded: a lost rank traps instead of hanging the GPU)
    const unsigned long long t0 = p2p_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(a.flags[a.rank] + threadIdx.x) < seq + 1u) {
      if ((++spins & 0x3ff) == 0 && p2p_timer_ns() - t0 > 10000000000ull) __trap();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < a.world; ++w)      // rank order: the same sum, bit for bit, on every rank
      s += ld_relaxed_sys_f64(a.data[w] + static_cast<size_t>(seq & 1u) * P2P_NMAX + i);
    buf[i] = s;
  }
}
}  // namespace metrpo

using namespace metrpo;

struct metrpo_trpo {
  metrpo_trpo_cfg cfg;
  PolDims pd;
  int num_sms = 0, max_smem = 0;
  double* acc = nullptr;      // [P + ACC_EXTRA]
  double* cg = nullptr;       // [6P + CGS_COUNT]
  float* vec_f = nullptr;     // [P]
  float* trial_f = nullptr;   // [P]
  int* flags = nullptr;       // [4]
  double* gram = nullptr;     // [(D+1)*D]
  // per-batch workspace (grown on demand)
  long long cap = 0;
  int* pos = nullptr;
  double* base = nullptr;
  double* adv_raw = nullptr;
  metrpo_allreduce_fn ar = nullptr;
  void* ar_user = nullptr;
  // one-shot peer-memory all-reduce (metrpo_trpo_enable_p2p)
  void* p2p_local = nullptr;              // this rank's exchange buffer (cudaMalloc, IPC-exported)
  void* p2p_peer[P2P_MAX_RANKS] = {};     // every rank's buffer as mapped here (own entry = p2p_local)
  int p2p_rank = 0, p2p_world = 0;
  unsigned p2p_seq = 0;
  int last_launches = 0;
  int pass_impl = METRPO_TRPO_PASS_AUTO;
  float* act_cache = nullptr;          // see PassParams::act_cache
  size_t act_cache_floats = 0;
};

static void trpo_free(metrpo_trpo* h) {
  if (!h) return;
  for (int w = 0; w < h->p2p_world; ++w)
    if (w != h->p2p_rank && h->p2p_peer[w]) cudaIpcCloseMemHandle(h->p2p_peer[w]);
  cudaFree(h->p2p_local);
  cudaFree(h->acc); cudaFree(h->cg); cudaFree(h->vec_f); cudaFree(h->trial_f); cudaFree(h->flags);
  cudaFree(h->gram); cudaFree(h->pos); cudaFree(h->base); cudaFree(h->adv_raw);
  delete h;
}

extern "C" int metrpo_trpo_create(const metrpo_trpo_cfg* cfg, metrpo_trpo_t** out) {
  if (!cfg || !out) return set_error(METRPO_ERR_INVALID, "trpo_create: null argument");
  *out = nullptr;
  const metrpo_trpo_cfg& c = *cfg;
  if (c.n_policy_layers < 1 || c.n_policy_layers > METRPO_MAX_POLICY_LAYERS)
    return set_error(METRPO_ERR_INVALID, "trpo_create: n_policy_layers in [1,%d]", METRPO_MAX_POLICY_LAYERS);
  if (c.state_dim < 1 || c.action_dim < 1 || c.policy_dims[0] != c.state_dim ||
      c.policy_dims[c.n_policy_layers] != c.action_dim)
    return set_error(METRPO_ERR_INVALID, "trpo_create: policy_dims must run from S to A");
  if (c.action_dim > 24)
    return set_error(METRPO_ERR_UNSUPPORTED, "trpo_create: action_dim <= 24 in this build (got %d)", c.action_dim);
  for (int l = 0; l <= c.n_policy_layers; ++l)
    if (c.policy_dims[l] < 1 || c.policy_dims[l] > 512)
      return set_error(METRPO_ERR_INVALID, "trpo_create: policy layer width out of range");
  METRPO_CUDA_OK(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  METRPO_CUDA_OK(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10)
    return set_error(METRPO_ERR_UNSUPPORTED, "trpo_create: device %d is sm_%d%d; this library is sm_100a only (no fallback)", c.device, prop.major, prop.minor);

  metrpo_trpo* h = new metrpo_trpo();
  h->cfg = c;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem = static_cast<int>(prop.sharedMemPerBlockOptin);
  PolDims& pd = h->pd;
  std::memset(&pd, 0, sizeof(pd));
  pd.L = c.n_policy_layers;
  pd.out_tanh = c.policy_out_tanh ? 1 : 0;
  int off = 0, soff = 0, row = 0;
  for (int l = 0; l <= pd.L; ++l) {
    pd.d[l] = c.policy_dims[l];
    pd.act_row[l] = row; row += pd.d[l];
    if (pd.d[l] > pd.max_d) pd.max_d = pd.d[l];
  }
  pd.sum_d = row;
  for (int l = 0; l < pd.L; ++l) {
    const int nin = pd.d[l], nout = pd.d[l + 1];
    pd.w_off[l] = off; off += nin * nout;
    pd.b_off[l] = off; off += nout;
    pd.np[l] = (nout + 3) & ~3;
    pd.sw_off[l] = soff; soff += nin * pd.np[l];
    pd.sb_off[l] = soff; soff += pd.np[l];
  }
  pd.logstd_off = off; off += pd.d[pd.L];
  pd.P = off;
  pd.P_pad = (soff + 3) & ~3;

  const int D = 2 * c.state_dim + 4;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
  };
  alloc(reinterpret_cast<void**>(&h->acc), (pd.P + ACC_EXTRA) * 8);
  alloc(reinterpret_cast<void**>(&h->cg), (6 * pd.P + CGS_COUNT) * 8);
  alloc(reinterpret_cast<void**>(&h->vec_f), pd.P * 4);
  alloc(reinterpret_cast<void**>(&h->trial_f), pd.P * 4);
  alloc(reinterpret_cast<void**>(&h->flags), 4 * 4);
  alloc(reinterpret_cast<void**>(&h->gram), static_cast<size_t>(D + 1) * D * 8);
  if (e != cudaSuccess) {
    trpo_free(h);
    return set_error(METRPO_ERR_CUDA, "trpo_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return METRPO_OK;
}

extern "C" int metrpo_trpo_destroy(metrpo_trpo_t* h) {
  if (!h) return METRPO_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  trpo_free(h);
  return METRPO_OK;
}

extern "C" int metrpo_trpo_num_params(const metrpo_trpo_t* h) { return h ? h->pd.P : 0; }
extern "C" int metrpo_trpo_last_launches(const metrpo_trpo_t* h) { return h ? h->last_launches : 0; }

extern "C" int metrpo_trpo_set_pass_impl(metrpo_trpo_t* h, int impl) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_set_pass_impl: null handle");
  if (impl < METRPO_TRPO_PASS_AUTO || impl > METRPO_TRPO_PASS_TF32X3)
    return set_error(METRPO_ERR_INVALID, "trpo_set_pass_impl: unknown implementation %d", impl);
  h->pass_impl = impl;
  return METRPO_OK;
}

extern "C" int metrpo_trpo_set_allreduce(metrpo_trpo_t* h, metrpo_allreduce_fn fn, void* user) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_set_allreduce: null handle");
  h->ar = fn; h->ar_user = user;
  return METRPO_OK;
}

static int ensure_workspace(metrpo_trpo* h, long long N) {
  if (N <= h->cap) return METRPO_OK;
  cudaFree(h->pos); cudaFree(h->base); cudaFree(h->adv_raw);
  h->pos = nullptr; h->base = nullptr; h->adv_raw = nullptr; h->cap = 0;
  METRPO_CUDA_OK(cudaMalloc(&h->pos, N * 4));
  METRPO_CUDA_OK(cudaMalloc(&h->base, N * 8));
  METRPO_CUDA_OK(cudaMalloc(&h->adv_raw, N * 8));
  h->cap = N;
  return METRPO_OK;
}

extern "C" int metrpo_trpo_p2p_handle(metrpo_trpo_t* h, void* handle_out) {
  if (!h || !handle_out) return set_error(METRPO_ERR_INVALID, "trpo_p2p_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == METRPO_IPC_HANDLE_BYTES, "IPC handle size");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  if (!h->p2p_local) {
    METRPO_CUDA_OK(cudaMalloc(&h->p2p_local, P2P_BYTES));
    METRPO_CUDA_OK(cudaMemset(h->p2p_local, 0, P2P_BYTES));
  }
  cudaIpcMemHandle_t hd;
  METRPO_CUDA_OK(cudaIpcGetMemHandle(&hd, h->p2p_local));
  std::memcpy(handle_out, &hd, sizeof(hd));
  return METRPO_OK;
}

extern "C" int metrpo_trpo_enable_p2p(metrpo_trpo_t* h, int rank, int world, const void* handles) {
  if (!h || !handles) return set_error(METRPO_ERR_INVALID, "trpo_enable_p2p: null argument");
  if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world)
    return set_error(METRPO_ERR_INVALID, "trpo_enable_p2p: need 0 <= rank < world <= %d", P2P_MAX_RANKS);
  if (!h->p2p_local) return set_error(METRPO_ERR_STATE, "trpo_enable_p2p: call metrpo_trpo_p2p_handle first");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  for (int w = 0; w < world; ++w) {
    if (w == rank) { h->p2p_peer[w] = h->p2p_local; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const char*>(handles) + static_cast<size_t>(w) * sizeof(hd), sizeof(hd));
    const cudaError_t e = cudaIpcOpenMemHandle(&h->p2p_peer[w], hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int v = 0; v < w; ++v)
        if (v != rank && h->p2p_peer[v]) { cudaIpcCloseMemHandle(h->p2p_peer[v]); h->p2p_peer[v] = nullptr; }
      cudaGetLastError();
      return set_error(METRPO_ERR_CUDA, "trpo_enable_p2p: cannot map rank %d's buffer: %s", w, cudaGetErrorString(e));
    }
  }
  h->p2p_rank = rank; h->p2p_world = world; h->p2p_seq = 0;
  return METRPO_OK;
}

static int allreduce(metrpo_trpo* h, double* buf, int n, cudaStream_t st) {
  if (h->p2p_world > 1) {
    if (n > P2P_NMAX) return set_error(METRPO_ERR_UNSUPPORTED, "all-reduce of %d doubles exceeds the exchange buffer", n);
    P2PArgs a;
    for (int w = 0; w < h->p2p_world; ++w) {
      a.data[w] = static_cast<double*>(h->p2p_peer[w]);
      a.flags[w] = reinterpret_cast<unsigned*>(static_cast<char*>(h->p2p_peer[w]) + 2 * static_cast<size_t>(P2P_NMAX) * 8);
    }
    a.rank = h->p2p_rank; a.world = h->p2p_world;
    k_allreduce_p2p<<<1, 256, 0, st>>>(a, buf, n, h->p2p_seq++);
    METRPO_CUDA_OK(cudaGetLastError());
    ++h->last_launches;
    return METRPO_OK;
  }
  if (!h->ar) return METRPO_OK;
  const int rc = h->ar(h->ar_user, buf, n, st);
  if (rc != 0) return set_error(METRPO_ERR_STATE, "all-reduce callback failed (%d)", rc);
  return METRPO_OK;
}

extern "C" int metrpo_trpo_process(metrpo_trpo_t* h, int T, int B, const float* obs, const float* rew,
                                   const uint8_t* done, const double* baseline_coeffs, double discount,
                                   double gae_lambda, int center_adv, int positive_adv, float* adv,
                                   float* ret, uint8_t* valid, double* stats, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_process: null handle");
  if (T < 1 || B < 1) return set_error(METRPO_ERR_INVALID, "trpo_process: T >= 1 and B >= 1 required");
  if (!obs || !rew || !done || !adv || !ret || !valid || !stats)
    return set_error(METRPO_ERR_INVALID, "trpo_process: null buffer");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const long long N = static_cast<long long>(T) * B;
  int rc = ensure_workspace(h, N);
  if (rc != METRPO_OK) return rc;
  const int S = h->cfg.state_dim;
  const int rb = (B + 127) / 128;
  const int gs = h->num_sms * 8;
  int launches = 0;
  pos_scan_kernel<<<rb, 128, 0, st>>>(done, h->pos, T, B); ++launches;
  if (baseline_coeffs) {
    baseline_predict_kernel<<<gs, 256, (2 * S + 4) * 8, st>>>(obs, h->pos, baseline_coeffs, h->base, N, S);
    ++launches;
  }
  gae_scan_kernel<<<rb, 128, 0, st>>>(rew, done, baseline_coeffs ? h->base : nullptr, discount, gae_lambda,
                                      h->adv_raw, ret, valid, T, B); ++launches;
  const double init[8] = {0, 0, 0, 1e300, 0, 0, 0, 0};
  METRPO_CUDA_OK(cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
  moments_kernel<<<gs, 256, 0, st>>>(h->adv_raw, valid, N, stats); ++launches;
  rc = allreduce(h, stats, 3, st);
  if (rc != METRPO_OK) return rc;
  moments_finish_kernel<<<1, 1, 0, st>>>(stats); ++launches;
  center_kernel<<<gs, 256, 0, st>>>(h->adv_raw, valid, stats, center_adv ? 1 : 0, positive_adv ? 1 : 0, adv, N);
  ++launches;
  METRPO_CUDA_OK(cudaGetLastError());
  h->last_launches = launches;
  return METRPO_OK;
}

extern "C" int metrpo_trpo_fit_baseline(metrpo_trpo_t* h, int T, int B, const float* obs, const float* ret,
                                        const uint8_t* valid, const uint8_t* done, double reg_coeff,
                                        double* coeffs_out, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_fit_baseline: null handle");
  if (T < 1 || B < 1 || !obs || !ret || !valid || !done || !coeffs_out)
    return set_error(METRPO_ERR_INVALID, "trpo_fit_baseline: bad argument");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const long long N = static_cast<long long>(T) * B;
  int rc = ensure_workspace(h, N);
  if (rc != METRPO_OK) return rc;
  const int S = h->cfg.state_dim, D = 2 * S + 4, E = (D + 1) * D;
  pos_scan_kernel<<<(B + 127) / 128, 128, 0, st>>>(done, h->pos, T, B);
  METRPO_CUDA_OK(cudaMemsetAsync(h->gram, 0, static_cast<size_t>(E) * 8, st));
  constexpr int NT = 128;
  const size_t smem = static_cast<size_t>(E) * 8 + static_cast<size_t>(D + 1) * (NT + NTPAD) * 4;
  if (smem > static_cast<size_t>(h->max_smem))
    return set_error(METRPO_ERR_UNSUPPORTED, "trpo_fit_baseline: state_dim %d needs %zu B of shared memory", S, smem);
  METRPO_CUDA_OK(cudaFuncSetAttribute(gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (N + NT - 1) / NT;
  const int grid = static_cast<int>(std::min<long long>(tiles, h->num_sms * 2));
  gram_kernel<NT><<<grid, NT, smem, st>>>(obs, ret, valid, h->pos, N, S, D, h->gram);
  rc = allreduce(h, h->gram, E, st);
  if (rc != METRPO_OK) return rc;
  const size_t ssm = static_cast<size_t>(D) * (D + 1) * 8;
  METRPO_CUDA_OK(cudaFuncSetAttribute(baseline_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
  baseline_solve_kernel<<<1, 128, ssm, st>>>(h->gram, D, reg_coeff, coeffs_out);
  METRPO_CUDA_OK(cudaGetLastError());
  h->last_launches = 3;
  return METRPO_OK;
}

// ---------------------------------------------------------------------------------------------
template <int MODE, int NT>
static int launch_pass_nt(metrpo_trpo* h, const PassParams& p, cudaStream_t st) {
  const PolDims& pd = h->pd;
  const size_t floats = static_cast<size_t>(pd.P_pad) * (MODE == MODE_FVP ? 2 : 1) +
                        (MODE == MODE_LOSS ? 0 : ((pd.P + 3) & ~3)) +
                        static_cast<size_t>(pd.sum_d + 2 * pd.max_d + 1) * (NT + NTPAD);
  const size_t smem = floats * 4;
  if (smem > static_cast<size_t>(h->max_smem)) return 1;   // try a smaller tile
  METRPO_CUDA_OK(cudaFuncSetAttribute(policy_pass_kernel<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  METRPO_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, policy_pass_kernel<MODE, NT>, NT, smem));
  if (per_sm < 1) per_sm = 1;
  const long long tiles = (p.N + NT - 1) / NT;
  const int grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(h->num_sms) * per_sm));
  policy_pass_kernel<MODE, NT><<<grid, NT, smem, st>>>(p);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
// tensor-core pass (trpo_mma.cuh): one persistent CTA of MMA_WARPS warps per SM
template <int MODE, int NS>
static int launch_pass_mma(metrpo_trpo* h, const PassParams& p, cudaStream_t st) {
  const size_t smem = mma_smem_bytes(h->pd, MODE == MODE_FVP, NS == 3);
  if (smem > static_cast<size_t>(h->max_smem)) return 1;
  METRPO_CUDA_OK(cudaFuncSetAttribute(policy_pass_mma_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (p.N + 31) / 32;
  const int grid = static_cast<int>(std::min<long long>((tiles + MMA_WARPS - 1) / MMA_WARPS, h->num_sms));
  policy_pass_mma_kernel<MODE, NS><<<grid, MMA_WARPS * 32, smem, st>>>(p);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
// register-tiled fp32 pass (trpo_tiled.cuh): persistent grid, 2 CTAs of 128 threads per SM
template <int MODE>
static int launch_pass_tiled(metrpo_trpo* h, const PassParams& p, cudaStream_t st) {
  const size_t smem = tiled_smem_bytes(h->pd, MODE);
  if (smem > static_cast<size_t>(h->max_smem)) return 1;
  METRPO_CUDA_OK(cudaFuncSetAttribute(policy_pass_tiled_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  METRPO_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, policy_pass_tiled_kernel<MODE>, TILED_NT, smem));
  if (per_sm < 1) per_sm = 1;
  const long long tiles = (p.N + TILED_NT - 1) / TILED_NT;
  const int grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(h->num_sms) * per_sm));
  policy_pass_tiled_kernel<MODE><<<grid, TILED_NT, smem, st>>>(p);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
template <int MODE>
static int launch_pass(metrpo_trpo* h, const PassParams& p, cudaStream_t st) {
  if (h->pass_impl == METRPO_TRPO_PASS_AUTO && tiled_eligible(h->pd)) {
    const int rc = launch_pass_tiled<MODE>(h, p, st);
    if (rc == METRPO_OK) { ++h->last_launches; return rc; }
    if (rc != 1) return rc;
  }
  if (h->pass_impl == METRPO_TRPO_PASS_TF32 || h->pass_impl == METRPO_TRPO_PASS_TF32X3) {
    if (!mma_eligible(h->pd))
      return set_error(METRPO_ERR_UNSUPPORTED, "trpo: the tensor-core pass covers <= 3 weight layers of width <= 32");
    int rc = h->pass_impl == METRPO_TRPO_PASS_TF32 ? launch_pass_mma<MODE, 1>(h, p, st)
                                                   : launch_pass_mma<MODE, 3>(h, p, st);
    if (rc == METRPO_OK) { ++h->last_launches; return rc; }
    if (rc != 1) return rc;
  }
  int rc = launch_pass_nt<MODE, 128>(h, p, st);
  if (rc == 1) rc = launch_pass_nt<MODE, 64>(h, p, st);
  if (rc == 1) rc = launch_pass_nt<MODE, 32>(h, p, st);
  if (rc == 1) return set_error(METRPO_ERR_UNSUPPORTED, "trpo: policy network too large for the shared-memory tile");
  if (rc == METRPO_OK) ++h->last_launches;
  return rc;
}

static int fill_pass(metrpo_trpo* h, PassParams& p, long long N, const float* obs, const float* act,
                     const float* adv, const float* old_mean, const float* old_log_std,
                     int old_log_std_per_sample, const uint8_t* valid) {
  p.pd = h->pd; p.obs = obs; p.act = act; p.adv = adv; p.old_mean = old_mean; p.old_log_std = old_log_std;
  p.old_ls_stride = old_log_std_per_sample ? h->cfg.action_dim : 0;
  p.valid = valid; p.N = N; p.acc = h->acc; p.skip_flag = nullptr; p.vec = nullptr; p.theta = nullptr;
  p.act_cache = nullptr;
  return METRPO_OK;
}

extern "C" int metrpo_trpo_loss_kl(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                                   const float* adv, const float* old_mean, const float* old_log_std,
                                   int old_log_std_per_sample, const uint8_t* valid, const float* theta,
                                   double* out2, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_loss_kl: null handle");
  if (N < 1 || !obs || !act || !adv || !old_mean || !old_log_std || !theta || !out2)
    return set_error(METRPO_ERR_INVALID, "trpo_loss_kl: bad argument");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  h->last_launches = 0;
  PassParams p;
  fill_pass(h, p, N, obs, act, adv, old_mean, old_log_std, old_log_std_per_sample, valid);
  p.theta = theta;
  METRPO_CUDA_OK(cudaMemsetAsync(h->acc, 0, (h->pd.P + ACC_EXTRA) * 8, st));
  int rc = launch_pass<MODE_LOSS>(h, p, st);
  if (rc != METRPO_OK) return rc;
  rc = allreduce(h, h->acc + h->pd.P, ACC_EXTRA, st);
  if (rc != METRPO_OK) return rc;
  // out2 = (loss, mean_kl): reuse k_ls_check's arithmetic on the host side of the stream
  double hacc[ACC_EXTRA];
  METRPO_CUDA_OK(cudaMemcpyAsync(hacc, h->acc + h->pd.P, sizeof(hacc), cudaMemcpyDeviceToHost, st));
  METRPO_CUDA_OK(cudaStreamSynchronize(st));
  out2[0] = -hacc[0] / hacc[2];
  out2[1] = hacc[1] / hacc[2];
  return METRPO_OK;
}

// debug / test hooks: gradient of the surrogate and one Fisher-vector product, returned as the
// raw device accumulators divided by the sample count (host doubles)
extern "C" int metrpo_trpo_grad(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                                const float* adv, const float* old_mean, const float* old_log_std,
                                int old_log_std_per_sample, const uint8_t* valid, const float* theta,
                                const float* vec, double reg_coeff, double* out_host, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_grad: null handle");
  if (N < 1 || !obs || !act || !adv || !old_mean || !old_log_std || !theta || !out_host)
    return set_error(METRPO_ERR_INVALID, "trpo_grad: bad argument");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int P = h->pd.P;
  h->last_launches = 0;
  PassParams p;
  fill_pass(h, p, N, obs, act, adv, old_mean, old_log_std, old_log_std_per_sample, valid);
  p.theta = theta; p.vec = vec;
  METRPO_CUDA_OK(cudaMemsetAsync(h->acc, 0, (P + ACC_EXTRA) * 8, st));
  int rc = vec ? launch_pass<MODE_FVP>(h, p, st) : launch_pass<MODE_GRAD>(h, p, st);
  if (rc != METRPO_OK) return rc;
  rc = allreduce(h, h->acc, P + ACC_EXTRA, st);
  if (rc != METRPO_OK) return rc;
  std::vector<double> hacc(P + ACC_EXTRA);
  std::vector<float> hth(P), hv(P);
  METRPO_CUDA_OK(cudaMemcpyAsync(hacc.data(), h->acc, (P + ACC_EXTRA) * 8, cudaMemcpyDeviceToHost, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(hth.data(), theta, P * 4, cudaMemcpyDeviceToHost, st));
  if (vec) METRPO_CUDA_OK(cudaMemcpyAsync(hv.data(), vec, P * 4, cudaMemcpyDeviceToHost, st));
  METRPO_CUDA_OK(cudaStreamSynchronize(st));
  const double Nv = hacc[P + 2];
  for (int e = 0; e < P; ++e) {
    double v = hacc[e] / Nv;
    if (vec) {
      if (e >= h->pd.logstd_off) {
        const double ls = hth[e];
        double hh = 0.0;
        if (ls > -13.815510557964274) { const double s = std::exp(2.0 * ls), eps = 1e-8; hh = 4.0 * s * (2.0 * s - eps) / ((2.0 * s + eps) * (2.0 * s + eps)); }
        v = hh * hv[e];
      }
      v += reg_coeff * hv[e];
    }
    out_host[e] = v;
  }
  return METRPO_OK;
}

extern "C" int metrpo_trpo_update(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                                  const float* adv, const float* old_mean, const float* old_log_std,
                                  int old_log_std_per_sample, const uint8_t* valid, float* theta,
                                  double step_size, int cg_iters, double reg_coeff, double backtrack_ratio,
                                  int max_backtracks, double* info, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "trpo_update: null handle");
  if (N < 1 || !obs || !act || !adv || !old_mean || !old_log_std || !theta)
    return set_error(METRPO_ERR_INVALID, "trpo_update: bad argument");
  if (cg_iters < 1 || max_backtracks < 1 || !(step_size > 0))
    return set_error(METRPO_ERR_INVALID, "trpo_update: cg_iters >= 1, max_backtracks >= 1, step_size > 0 required");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const PolDims& pd = h->pd;
  const int P = pd.P, A = pd.d[pd.L];
  h->last_launches = 0;
  int rc;
  PassParams p;
  fill_pass(h, p, N, obs, act, adv, old_mean, old_log_std, old_log_std_per_sample, valid);
  METRPO_CUDA_OK(cudaMemsetAsync(h->acc, 0, (P + ACC_EXTRA) * 8, st));

  // Hidden activations of the policy at theta are the same for the gradient pass and all Fisher-vector
  // passes of this update: the gradient pass stores them (64 floats per sample for 32-32 policies, 1 GB at
  // 4.1 M samples), the Fisher-vector passes load them instead of recomputing the forward pass.
  static const bool use_act_cache = [] { const char* ev = getenv("METRPO_TRPO_ACT_CACHE"); return !ev || atoi(ev) != 0; }();
  if (use_act_cache && h->pass_impl == METRPO_TRPO_PASS_AUTO && tiled_eligible(pd) && !pd.out_tanh) {
    const size_t hid_rows = static_cast<size_t>(pd.sum_d - pd.d[0] - pd.d[pd.L]);
    const size_t need = static_cast<size_t>((N + TILED_NT - 1) / TILED_NT) * hid_rows * TILED_NT;
    if (need > h->act_cache_floats && need * 4 <= (static_cast<size_t>(8) << 30)) {
      if (h->act_cache) { cudaStreamSynchronize(st); cudaFree(h->act_cache); h->act_cache = nullptr; h->act_cache_floats = 0; }
      if (cudaMalloc(&h->act_cache, need * 4) == cudaSuccess) h->act_cache_floats = need;
      else { h->act_cache = nullptr; cudaGetLastError(); }
    }
    if (hid_rows > 0 && need <= h->act_cache_floats) p.act_cache = h->act_cache;
  }

  // loss_before and flat gradient g (one pass yields both)
  p.theta = theta;
  if ((rc = launch_pass<MODE_GRAD>(h, p, st)) != METRPO_OK) return rc;
  if ((rc = allreduce(h, h->acc, P + ACC_EXTRA, st)) != METRPO_OK) return rc;
  k_grad_finish<<<1, CTRL_THREADS, 0, st>>>(P, h->acc, theta, h->cg, h->vec_f, h->flags); ++h->last_launches;

  // d = cg(Hx, g)
  p.vec = h->vec_f;
  for (int it = 0; it < cg_iters; ++it) {
    p.skip_flag = h->flags + FLAG_CG_DONE;
    if ((rc = launch_pass<MODE_FVP>(h, p, st)) != METRPO_OK) return rc;
    if ((rc = allreduce(h, h->acc, P + ACC_EXTRA, st)) != METRPO_OK) return rc;
    k_cg_step<<<1, CTRL_THREADS, 0, st>>>(P, pd.logstd_off, A, h->acc, h->cg, h->vec_f, h->flags, reg_coeff,
                                          it == cg_iters - 1 ? 1 : 0); ++h->last_launches;
  }
  // step0 from d.Hx(d): by default from the CG residual (r = g - Hx is maintained by the recurrence),
  // METRPO_TRPO_EXPLICIT_SHS=1 runs rllab's explicit extra Fisher-vector pass instead
  static const bool explicit_shs = [] { const char* ev = getenv("METRPO_TRPO_EXPLICIT_SHS"); return ev && atoi(ev) != 0; }();
  p.skip_flag = nullptr;
  if (explicit_shs) {
    if ((rc = launch_pass<MODE_FVP>(h, p, st)) != METRPO_OK) return rc;
    if ((rc = allreduce(h, h->acc, P + ACC_EXTRA, st)) != METRPO_OK) return rc;
  }
  k_step_finish<<<1, CTRL_THREADS, 0, st>>>(P, pd.logstd_off, A, h->acc, h->cg, reg_coeff, step_size,
                                            explicit_shs ? 0 : 1); ++h->last_launches;

  // back-tracking line search: ratio in backtrack_ratio ** arange(max_backtracks)
  p.vec = nullptr;
  p.act_cache = nullptr;
  p.theta = h->trial_f;
  p.skip_flag = h->flags + FLAG_ACCEPTED;
  for (int k = 0; k < max_backtracks; ++k) {
    k_ls_prepare<<<1, CTRL_THREADS, 0, st>>>(P, h->cg, h->trial_f, h->flags, std::pow(backtrack_ratio, k)); ++h->last_launches;
    if ((rc = launch_pass<MODE_LOSS>(h, p, st)) != METRPO_OK) return rc;
    if ((rc = allreduce(h, h->acc + P, ACC_EXTRA, st)) != METRPO_OK) return rc;
    k_ls_check<<<1, CTRL_THREADS, 0, st>>>(P, h->acc, h->cg, h->flags, k, step_size); ++h->last_launches;
  }
  k_finalize<<<1, CTRL_THREADS, 0, st>>>(P, h->cg, h->trial_f, theta, h->flags, step_size, info); ++h->last_launches;
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}
