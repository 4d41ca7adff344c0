// fp32-faithful arithmetic mode of the rollout path (cfg.precision = METRPO_PREC_FP32).
//
// The reference evaluates the dynamics MLP with fp32 tf.matmul (training.py:207-208); tcgen05 has
// no fp32 MMA, so the fast path (rollout_kernel.cuh / rollout_duo.cuh) rounds the operands to bf16.
// This mode runs the SAME step semantics (policy, clip, normalise, all-K dynamics, sam_mode
// selection, analytic cost / done, reset, trajectory layout, Philox streams) with every product in
// fp32 FMA arithmetic on CUDA cores, true division by in_std and tanhf, so that the cost of the
// bf16 operands can be MEASURED on the device at full size and long horizons (tests/
// test_bench_parity_gpu.py) and so that a user can trade ~50x of speed for the reference's
// arithmetic (validation costs that drive early stopping, debugging).  It is a fidelity tool, not
// a roofline kernel: one launch per layer and step, a plain 64 x 64 register-tiled SGEMM.
#pragma once
#include "rollout_kernel.cuh"

namespace metrpo {

struct Fp32Params {
  int S, A, SA, drop, Din, H, K, B, T_max, env_id, sam_mode, determ, per_model;
  int n_pol_layers, pol_out_tanh, pol_logstd_off, row_offset;
  PolicyLayer pl[4];
  const float* pol;      // packed policy blob (same layout as the tensor-core path)
  const float* norm;     // in_mean[SA] | in_std[SA] | diff_mean[S] | diff_std[S]
  float gamma;
  // per-step inputs
  int t;                 // step index inside this launch sequence (buffers are indexed by it)
  const float* eps; const int* model_idx; const float* std_noise;
  const float* ext_actions; const float* ext_reset_states;
  const float* reset_pool; int R;
  unsigned long long seed, offset;
  // state
  float* x;              // [sets][B][S]   sets = per_model ? K : 1
  int* ts; int* nreset;  // [B]
  float* a_clip;         // [sets][B][A]
  float* a_raw;          // [sets][B][A] unclipped action of this step (stored in the trajectory)
  float* z;              // [sets][B][Din]
  const float* o;        // [K][B][S] layer-2 output (h1 @ W2 + b2)
  float* pm_acc; float* pm_gpow; float* pm_dmask;   // per_model: [K][B]
  // outputs
  float* obs; float* act; float* mean; float* rew; uint8_t* done;
};

// ---------------------------------------------------------------------------------------------
// policy + Z operand: one thread per (set, row)
// ---------------------------------------------------------------------------------------------
__global__ void fp32_begin_step(const Fp32Params p) {
  const int sets = p.per_model ? p.K : 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sets * p.B) return;
  const int row = i % p.B;
  const float* x = p.x + static_cast<size_t>(i) * p.S;
  float a_mean[24], a_raw[24];
  if (p.ext_actions != nullptr) {
    for (int a = 0; a < p.A; ++a) a_mean[a] = a_raw[a] = p.ext_actions[row * p.A + a];
  } else {
    // mean network (training.py:99-103): tanh hidden layers, identity / tanh output
    float bufA[HPMAX], bufB[HPMAX];
    float* cur = bufA;
    float* nxt = bufB;
    for (int s = 0; s < p.S; ++s) cur[s] = x[s];
    for (int l = 0; l < p.n_pol_layers; ++l) {
      const PolicyLayer L = p.pl[l];
      const bool last = (l == p.n_pol_layers - 1);
      for (int j = 0; j < L.nout; ++j) {
        float acc = p.pol[L.b_off + j];
        for (int q = 0; q < L.nin; ++q) acc = fmaf(cur[q], p.pol[L.w_off + q * L.npad + j], acc);
        if (last) a_mean[j] = p.pol_out_tanh ? tanhf(acc) : acc;
        else nxt[j] = tanhf(acc);
      }
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (p.determ) {
      for (int a = 0; a < p.A; ++a) a_raw[a] = a_mean[a];
    } else {
      for (int a = 0; a < p.A; ++a) {
        float e;
        if (p.eps != nullptr) {
          e = p.eps[(static_cast<size_t>(p.t) * p.B + row) * p.A + a];
        } else {
          float n4[4];
          philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(p.t),
                         static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EPS + (a >> 2), n4);
          e = n4[a & 3];
        }
        const float ls = fmaxf(p.pol[p.pol_logstd_off + a], -13.815510557964274f);   // min_std 1e-6
        a_raw[a] = __fadd_rn(__fmul_rn(e, expf(ls)), a_mean[a]);                       // rllab get_actions
      }
    }
    if (!p.per_model) {
      const size_t o = static_cast<size_t>(p.t) * p.B + row;
      for (int a = 0; a < p.A; ++a) {
        if (p.act) p.act[o * p.A + a] = a_raw[a];
        if (p.mean) p.mean[o * p.A + a] = a_mean[a];
      }
    }
  }
  const float* in_mean = p.norm;
  const float* in_std = p.norm + p.SA;
  float* z = p.z + static_cast<size_t>(i) * p.Din;
  for (int a = 0; a < p.A; ++a) {
    const float u = fminf(fmaxf(a_raw[a], -1.f), 1.f);                                 // env_helpers.py:599
    p.a_clip[static_cast<size_t>(i) * p.A + a] = u;
    p.a_raw[static_cast<size_t>(i) * p.A + a] = a_raw[a];
  }
  // z = ((xgu - mean) / std)[:, drop:]   (training.py:228,146-154) -- a true division here
  for (int f = p.drop; f < p.SA; ++f) {
    const float v = f < p.S ? x[f] : p.a_clip[static_cast<size_t>(i) * p.A + (f - p.S)];
    z[f - p.drop] = __fdiv_rn(__fsub_rn(v, in_mean[f]), in_std[f]);
  }
}

// ---------------------------------------------------------------------------------------------
// C[k] = act(A[k] @ W[k] + b[k]);  A [M,Kd] (batch stride sA, 0 = shared), W [Kd,N], C [M,N]
// 64 x 64 tile, 256 threads, 4 x 4 outputs per thread, fp32 FMA in k order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_gemm_bias_act(const float* __restrict__ A, long long sA,
                                                          const float* __restrict__ W, long long sW,
                                                          const float* __restrict__ bias, long long sB,
                                                          float* __restrict__ C, long long sC, int M, int N,
                                                          int Kd, int relu) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int kb = blockIdx.z;
  A += kb * sA; W += kb * sW; bias += kb * sB; C += kb * sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Kd; k0 += 16) {
    for (int q = threadIdx.x; q < 64 * 16; q += 256) {
      const int r = q >> 4, kk = q & 15;                 // A tile: 64 rows x 16 k
      As[kk][r] = (m0 + r < M && k0 + kk < Kd) ? A[static_cast<size_t>(m0 + r) * Kd + k0 + kk] : 0.f;
      const int kr = q >> 6, cn = q & 63;                // W tile: 16 k x 64 cols
      Ws[kr][cn] = (k0 + kr < Kd && n0 + cn < N) ? W[static_cast<size_t>(k0 + kr) * N + n0 + cn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = __fadd_rn(acc[i][j], bias[n]);
      if (relu) v = fmaxf(v, 0.f);
      C[static_cast<size_t>(m) * N + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// candidate -> select -> reward / done / reset -> trajectory: one thread per row (per (model,row)
// in per-model mode)
// ---------------------------------------------------------------------------------------------
template <int SMAX, int AMAX>
__global__ void fp32_finish_step(const Fp32Params p) {
  const int sets = p.per_model ? p.K : 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sets * p.B) return;
  const int row = i % p.B, S = p.S, A = p.A, K = p.K;
  const float* dmean = p.norm + 2 * p.SA;
  const float* dstd = p.norm + 2 * p.SA + S;
  float* x = p.x + static_cast<size_t>(i) * S;
  float xn[SMAX], u[AMAX];
#pragma unroll
  for (int s = 0; s < SMAX; ++s) xn[s] = 0.f;
#pragma unroll
  for (int a = 0; a < AMAX; ++a) u[a] = a < A ? p.a_clip[static_cast<size_t>(i) * A + a] : 0.f;
  // next_state = (diff_mean + diff_std * nn_output) + x   (training.py:257)
  auto cand = [&](int k, int s) {
    const float o = p.o[(static_cast<size_t>(k) * p.B + row) * S + s];
    return __fadd_rn(__fadd_rn(dmean[s], __fmul_rn(dstd[s], o)), x[s]);
  };
  const int mode = p.sam_mode;
  if (p.per_model) {
    const int k = i / p.B;
    for (int s = 0; s < S; ++s) xn[s] = cand(k, s);
  } else if (K == 1 || mode == METRPO_SAM_ONE_MODEL) {
    for (int s = 0; s < S; ++s) xn[s] = cand(0, s);
  } else if (mode == METRPO_SAM_STEP_RAND || mode == METRPO_SAM_EPS_RAND) {
    int idx;
    if (p.model_idx != nullptr) idx = p.model_idx[static_cast<size_t>(p.t) * p.B + row];
    else if (mode == METRPO_SAM_STEP_RAND)
      idx = philox_index(p.seed, p.offset + static_cast<unsigned long long>(p.t),
                         static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_IDX, K);
    else
      idx = philox_index(p.seed, static_cast<uint64_t>(static_cast<uint32_t>(p.nreset[row])),
                         static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EIDX, K);
    idx = min(max(idx, 0), K - 1);
    for (int s = 0; s < S; ++s) xn[s] = cand(idx, s);
  } else {
    for (int s = 0; s < S; ++s) {
      float m = 0.f;
      for (int k = 0; k < K; ++k) m = __fadd_rn(m, cand(k, s));
      m = __fdiv_rn(m, static_cast<float>(K));
      float outv = m;
      if (mode == METRPO_SAM_MODEL_MEAN_STD) {
        float var = 0.f;
        for (int k = 0; k < K; ++k) { const float d = __fsub_rn(cand(k, s), m); var = __fadd_rn(var, __fmul_rn(d, d)); }
        const float sd = sqrtf(__fdiv_rn(var, static_cast<float>(K)));
        float nz;
        if (p.std_noise != nullptr) {
          nz = p.std_noise[(static_cast<size_t>(p.t) * p.B + row) * S + s];
        } else {
          float n4[4];
          philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(p.t),
                         static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_STD + (s >> 2), n4);
          nz = n4[s & 3];
        }
        outv = __fadd_rn(m, __fmul_rn(nz, sd));
      } else if (mode == METRPO_SAM_MODEL_MED) {
        float lo = 0.f, hi = 0.f;
        const int r_lo = (K - 1) / 2, r_hi = K / 2;
        for (int a = 0; a < K; ++a) {
          const float va = cand(a, s);
          int less = 0, eq = 0;
          for (int b = 0; b < K; ++b) { const float vb = cand(b, s); less += (vb < va); eq += (vb == va); }
          if (less <= r_lo && r_lo < less + eq) lo = va;
          if (less <= r_hi && r_hi < less + eq) hi = va;
        }
        outv = (r_lo == r_hi) ? lo : __fmul_rn(__fadd_rn(lo, hi), 0.5f);
      }
      xn[s] = outv;
    }
  }
  const float reward = -env_cost<SMAX, AMAX>(p.env_id, S, A, xn, u);                 // env_helpers.py:601
  if (p.per_model) {   // model_based_rl.py:133-139
    const float c = __fmul_rn(-reward, 1.f - p.pm_dmask[i]);
    p.pm_acc[i] = __fadd_rn(p.pm_acc[i], __fmul_rn(p.pm_gpow[i], c));
    p.pm_gpow[i] = __fmul_rn(p.pm_gpow[i], p.gamma);
    if (env_is_done<SMAX>(p.env_id, S, xn)) p.pm_dmask[i] = 1.f;
    for (int s = 0; s < S; ++s) x[s] = xn[s];
    return;
  }
  const int ts = p.ts[row] + 1;
  const bool dn = env_is_done<SMAX>(p.env_id, S, xn) || (ts >= p.T_max);             // :603-604
  const size_t o = static_cast<size_t>(p.t) * p.B + row;
  if (p.obs) for (int s = 0; s < S; ++s) p.obs[o * S + s] = x[s];
  if (p.rew) p.rew[o] = reward;
  if (p.done) p.done[o] = dn ? 1 : 0;
  if (dn) {   // :605-606 -> reset(dones)
    const int nr = p.nreset[row];
    const float* src = p.ext_reset_states
                           ? p.ext_reset_states + static_cast<size_t>(row) * S
                           : p.reset_pool + static_cast<size_t>((static_cast<long long>(nr) * p.B + row) % p.R) * S;
    for (int s = 0; s < S; ++s) x[s] = src[s];
    p.nreset[row] = nr + 1;
    p.ts[row] = 0;
  } else {
    for (int s = 0; s < S; ++s) x[s] = xn[s];
    p.ts[row] = ts;
  }
}

__global__ void fp32_fill(float* p, float v, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fp32_tile_states(const float* __restrict__ src, float* __restrict__ dst, int K, size_t per) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < per * K) dst[i] = src[i % per];
}

}  // namespace metrpo
