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
#include <cmath>
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
};

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
      if (j0 + q < nout) out[(j0 + q) * LD + n] = use_tanh ? tanhf(acc[q]) : acc[q];
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
        const float ls = fmaxf(sW_logstd(p, a), -13.815510557964274f);
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
        const float ols = p.old_log_std[(p.old_ls_stride ? ng * p.old_ls_stride : 0) * (inb ? 1 : 0) + a];
        const float ls = fmaxf(sW_logstd(p, a), -13.815510557964274f);   // min_std 1e-6
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
          const float lsr = sW_logstd(p, a);
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
