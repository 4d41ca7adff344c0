// Warp-level tensor-core version (mma.sync m16n8k8 TF32, fp32 accumulate) of the TRPO sample pass
// (loss / gradient / Fisher-vector product) for the reference's policy shapes (<= 3 weight layers,
// every width <= 32: params/*.json use 32-32).  Selectable with metrpo_trpo_set_pass_impl; NOT the
// default.
//
// Why it exists and why it is not the default.  The pass streams N samples once from HBM (obs,
// act, old_mean, adv: 124 B/sample for half-cheetah) and evaluates ~2-8 k MACs of tiny dense
// layers per sample; on CUDA cores that arithmetic and its shared-memory operand traffic are the
// bound, far from the HBM roofline.  The layers are 32-row GEMMs whose outputs feed tanh / masks
// immediately, which suits warp-level MMAs (no cross-warp synchronisation, 32 samples per
// warp-tile, operands in warp-private shared memory).  Measured on B200
// (tools/mma_sync_bench.py, profiles/r1_mma_sync_bench.json): the legacy mma.sync path issues one
// m16n8k8 TF32 MMA per 16.7 cycles per SM sub-partition = 252 MAC/clk/SM, only 2x the FP32 FMA
// rate (bf16 m16n8k16: 504 MAC/clk/SM) and ~1/16 of tcgen05.  With single TF32 products this
// kernel ties the SIMT one (FVP 6.1 vs 6.8 ms on 4.1 M samples), with the fp32-equivalent 3xTF32
// split it is slower (9.6 ms).  Reaching the HBM roofline needs tcgen05 (M = 128 sample tiles,
// operands through TMEM, several tiles in flight to hide the commit/wait round trip of each
// dependent 32-wide GEMM): DESIGN.md section 8.
//
// Precision: NS = 3 evaluates every product as hi*hi + lo*hi + hi*lo of TF32 splits (error
// ~2^-21, fp32-equivalent) -- the reference's graph is fp32; NS = 1 is plain TF32.
#pragma once

namespace metrpo {

constexpr int MMA_WARPS = 6;
constexpr int MMA_LDA = 36;          // activation tile [32 samples][36]: conflict-free A fragments
constexpr int MMA_MAXL = 3;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct MmaLayer {
  int nin, nout, Kp, Np, Mp, LDB;   // Kp = pad8(nin), Np = pad8(nout), Mp = pad16(nin)
  int w_hi, w_lo, v_hi, v_lo, b, vb;   // float offsets into the CTA's weight area
};
struct MmaLayout {
  MmaLayer l[MMA_MAXL];
  int total;   // floats
};
__host__ __device__ inline MmaLayout mma_layout(const PolDims& pd, bool fvp, bool split) {
  MmaLayout m;
  int off = 0;
  for (int i = 0; i < pd.L && i < MMA_MAXL; ++i) {
    MmaLayer& L = m.l[i];
    L.nin = pd.d[i]; L.nout = pd.d[i + 1];
    L.Kp = (L.nin + 7) & ~7; L.Np = (L.nout + 7) & ~7; L.Mp = (L.nin + 15) & ~15;
    L.LDB = (L.Np % 32 == 0) ? L.Np + 8 : L.Np;      // B-fragment reads (k = t, n = g) conflict-free
    const int rows = L.Mp > L.Kp ? L.Mp : L.Kp;      // rows >= nin are zero
    const int sz = rows * L.LDB;
    L.w_hi = off; off += sz;
    L.w_lo = off; off += split ? sz : 0;
    L.v_hi = off; off += fvp ? sz : 0;
    L.v_lo = off; off += (fvp && split) ? sz : 0;
    L.b = off; off += L.Np;
    L.vb = off; off += fvp ? L.Np : 0;
  }
  m.total = (off + 3) & ~3;
  return m;
}
inline bool mma_eligible(const PolDims& pd) {
  if (pd.L > MMA_MAXL) return false;
  for (int i = 0; i <= pd.L; ++i)
    if (pd.d[i] > 32) return false;
  return true;
}
inline size_t mma_smem_bytes(const PolDims& pd, bool fvp, bool split) {
  const MmaLayout m = mma_layout(pd, fvp, split);
  return (static_cast<size_t>(m.total) + static_cast<size_t>(MMA_WARPS) * 6 * 32 * MMA_LDA + pd.P + 64) * 4;
}

// A fragments (both 16-row blocks, all k-steps) of a [32][LDA] tile, split into TF32 hi / lo
template <int NS>
__device__ __forceinline__ void load_a_frags(uint32_t (&ah)[2][4][4], uint32_t (&al)[2][4][4], const float* in,
                                             int Kp, int g, int t) {
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      if (ks * 8 < Kp) {
        const float* r0 = in + (mb * 16 + g) * MMA_LDA + ks * 8 + t;
        const float* r1 = r0 + 8 * MMA_LDA;
        const float v0 = r0[0], v1 = r1[0], v2 = r0[4], v3 = r1[4];
        if (NS == 3) {
          split_tf32(v0, ah[mb][ks][0], al[mb][ks][0]); split_tf32(v1, ah[mb][ks][1], al[mb][ks][1]);
          split_tf32(v2, ah[mb][ks][2], al[mb][ks][2]); split_tf32(v3, ah[mb][ks][3], al[mb][ks][3]);
        } else {
          ah[mb][ks][0] = f2tf32(v0); ah[mb][ks][1] = f2tf32(v1); ah[mb][ks][2] = f2tf32(v2); ah[mb][ks][3] = f2tf32(v3);
        }
      }
}

// C[32 x Np] (+)= A[32 x Kp] * B[Kp x Np]:  A = in (samples x features, smem tile), B = weights
// [k][n] at wh / wl (TRANS: B[k = j][n = i] = W[i][j], the back-propagation of deltas).
// c[mb][nb][4] accumulators are initialised by the caller.  Consecutive MMAs go to different
// accumulators (8 independent chains); every B fragment is loaded once for both row blocks.
template <int NS, bool TRANS>
__device__ __forceinline__ void gemm_rows_t(float (&c)[2][4][4], const float* in, int Kp, int Np, const float* wh,
                                            const float* wl, int LDB, int g, int t) {
  uint32_t ah[2][4][4], al[2][4][4];
  load_a_frags<NS>(ah, al, in, Kp, g, t);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    if (ks * 8 < Kp) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        if (nb * 8 < Np) {
          const int o = TRANS ? (nb * 8 + g) * LDB + ks * 8 + t : (ks * 8 + t) * LDB + nb * 8 + g;
          const int o1 = TRANS ? o + 4 : o + 4 * LDB;
          bh[nb][0] = __float_as_uint(wh[o]); bh[nb][1] = __float_as_uint(wh[o1]);
          if (NS == 3) { bl[nb][0] = __float_as_uint(wl[o]); bl[nb][1] = __float_as_uint(wl[o1]); }
        }
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        if (nb * 8 < Np) { mma_tf32(c[0][nb], ah[0][ks], bh[nb][0], bh[nb][1]); mma_tf32(c[1][nb], ah[1][ks], bh[nb][0], bh[nb][1]); }
      if (NS == 3) {
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          if (nb * 8 < Np) { mma_tf32(c[0][nb], al[0][ks], bh[nb][0], bh[nb][1]); mma_tf32(c[1][nb], al[1][ks], bh[nb][0], bh[nb][1]); }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          if (nb * 8 < Np) { mma_tf32(c[0][nb], ah[0][ks], bl[nb][0], bl[nb][1]); mma_tf32(c[1][nb], ah[1][ks], bl[nb][0], bl[nb][1]); }
      }
    }
}
template <int NS>
__device__ __forceinline__ void gemm_rows(float (&c)[2][4][4], const float* in, int Kp, int Np, const float* wh,
                                          const float* wl, int LDB, int g, int t) {
  gemm_rows_t<NS, false>(c, in, Kp, Np, wh, wl, LDB, g, t);
}
// C[32 x Kout] = D[32 x Np] * W^T
template <int NS>
__device__ __forceinline__ void gemm_rows_wt(float (&c)[2][4][4], const float* din, int Np, int Kout,
                                             const float* wh, const float* wl, int LDB, int g, int t) {
  gemm_rows_t<NS, true>(c, din, Np, Kout, wh, wl, LDB, g, t);
}

// G[Mp x Np] += A^T[Mp x 32] * D[32 x Np]: the parameter-gradient outer products, reduced over the
// 32 samples of the tile.  A[i][smp] = act[smp][i].
template <int NS>
__device__ __forceinline__ void gemm_outer(float (&G)[2][4][4], const float* act, int Mp, const float* dl, int Np,
                                           int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {   // 8 samples per k-step
    uint32_t bh[4][2], bl[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
      if (nb * 8 < Np) {
        const float v0 = dl[(ks * 8 + t) * MMA_LDA + nb * 8 + g], v1 = dl[(ks * 8 + t + 4) * MMA_LDA + nb * 8 + g];
        if (NS == 3) { split_tf32(v0, bh[nb][0], bl[nb][0]); split_tf32(v1, bh[nb][1], bl[nb][1]); }
        else { bh[nb][0] = f2tf32(v0); bh[nb][1] = f2tf32(v1); }
      }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
      if (mb * 16 < Mp) {
        uint32_t ah[4], al[4];
        const float* c0 = act + (ks * 8 + t) * MMA_LDA + mb * 16 + g;
        const float v0 = c0[0], v1 = c0[8], v2 = c0[4 * MMA_LDA], v3 = c0[4 * MMA_LDA + 8];
        if (NS == 3) { split_tf32(v0, ah[0], al[0]); split_tf32(v1, ah[1], al[1]); split_tf32(v2, ah[2], al[2]); split_tf32(v3, ah[3], al[3]); }
        else { ah[0] = f2tf32(v0); ah[1] = f2tf32(v1); ah[2] = f2tf32(v2); ah[3] = f2tf32(v3); }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          if (nb * 8 < Np) {
            mma_tf32(G[mb][nb], ah, bh[nb][0], bh[nb][1]);
            if (NS == 3) {
              mma_tf32(G[mb][nb], al, bh[nb][0], bh[nb][1]);
              mma_tf32(G[mb][nb], ah, bl[nb][0], bl[nb][1]);
            }
          }
      }
  }
}

template <int MODE, int NS>
__global__ void __launch_bounds__(MMA_WARPS * 32, 1) policy_pass_mma_kernel(const __grid_constant__ PassParams p) {
  extern __shared__ __align__(16) float sm[];
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
  const PolDims& pd = p.pd;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int L = pd.L, A = pd.d[L], S = pd.d[0];
  const MmaLayout lay = mma_layout(pd, MODE == MODE_FVP, NS == 3);
  float* sWt = sm;                                                   // weights (all variants)
  float* sTiles = sWt + lay.total;                                   // [warps][6][32][LDA]
  float* sG = sTiles + MMA_WARPS * 6 * 32 * MMA_LDA;                 // [P] CTA-level gradient sums
  float* sMisc = sG + ((pd.P + 3) & ~3);                             // [32] M_a (FVP) / log_std
  __shared__ double sRed[3][MMA_WARPS];

  // ---- stage weights: zero padded, TF32 hi (+ lo) ----
  for (int i = tid; i < lay.total; i += blockDim.x) sWt[i] = 0.f;
  if (MODE != MODE_LOSS)
    for (int i = tid; i < pd.P; i += blockDim.x) sG[i] = 0.f;
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const MmaLayer& ML = lay.l[l];
    for (int e = tid; e < ML.nin * ML.nout; e += blockDim.x) {
      const int i = e / ML.nout, j = e - i * ML.nout;
      const float w = p.theta[pd.w_off[l] + e];
      const float wh = __uint_as_float(f2tf32(w));
      sWt[ML.w_hi + i * ML.LDB + j] = wh;
      if (NS == 3) sWt[ML.w_lo + i * ML.LDB + j] = __uint_as_float(f2tf32(w - wh));
      if (MODE == MODE_FVP) {
        const float v = p.vec[pd.w_off[l] + e];
        const float vh = __uint_as_float(f2tf32(v));
        sWt[ML.v_hi + i * ML.LDB + j] = vh;
        if (NS == 3) sWt[ML.v_lo + i * ML.LDB + j] = __uint_as_float(f2tf32(v - vh));
      }
    }
    for (int j = tid; j < ML.nout; j += blockDim.x) {
      sWt[ML.b + j] = p.theta[pd.b_off[l] + j];
      if (MODE == MODE_FVP) sWt[ML.vb + j] = p.vec[pd.b_off[l] + j];
    }
  }
  if (tid < 32) {
    float v = 0.f;
    if (tid < A) {
      const float lsr = p.theta[pd.logstd_off + tid];
      if (MODE == MODE_FVP) {
        const float ls = fmaxf(lsr, -13.815510557964274f);
        v = 2.f / (2.f * __expf(2.f * ls) + 1e-8f);      // d^2 kl / d mu^2  (kl_sym, A.3)
      } else {
        v = lsr;
      }
    }
    sMisc[tid] = v;
  }
  __syncthreads();

  float* tile = sTiles + warp * 6 * 32 * MMA_LDA;
  for (int i = lane; i < 6 * 32 * MMA_LDA; i += 32) tile[i] = 0.f;   // padding columns stay zero
  __syncwarp();
  float* bAct[MMA_MAXL] = {tile, tile + 32 * MMA_LDA, tile + 2 * 32 * MMA_LDA};   // inputs of layers 0..2
  float* bT0 = tile + 3 * 32 * MMA_LDA;
  float* bT1 = tile + 4 * 32 * MMA_LDA;
  float* bMu = tile + 5 * 32 * MMA_LDA;

  // per-warp gradient accumulators (GRAD / FVP), kept in registers across all tiles
  float G[MMA_MAXL][2][4][4];
  float gb[MMA_MAXL];      // bias gradients: lane j owns column j
  float gls = 0.f;         // log_std gradient: lane a owns entry a
#pragma unroll
  for (int l = 0; l < MMA_MAXL; ++l) {
    gb[l] = 0.f;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) G[l][mb][nb][q] = 0.f;
  }
  double t_surr = 0.0, t_kl = 0.0, t_cnt = 0.0;

  const long long n_tiles = (p.N + 31) / 32;
  for (long long tl = static_cast<long long>(blockIdx.x) * MMA_WARPS + warp; tl < n_tiles;
       tl += static_cast<long long>(gridDim.x) * MMA_WARPS) {
    const long long n0 = tl * 32, ng = n0 + lane;
    const bool inb = ng < p.N;
    const bool ok = inb && (p.valid == nullptr || p.valid[ng] != 0);
    const unsigned okmask = __ballot_sync(0xffffffffu, ok);
    __syncwarp();
    // observations -> bAct[0][smp][f]: the tile's [32,S] block is contiguous in HBM (columns >= S
    // of the tile stay zero from the one-time fill)
    {
      const long long base = n0 * S, lim = p.N * S;
      for (int q = lane; q < 32 * S; q += 32) {
        const int smp = q / S, f = q - smp * S;
        bAct[0][smp * MMA_LDA + f] = (base + q < lim) ? p.obs[base + q] : 0.f;
      }
    }
    __syncwarp();
    // ---- forward (training.py:99-103) ----
#pragma unroll
    for (int l = 0; l < MMA_MAXL; ++l)
      if (l < L) {
        const MmaLayer& ML = lay.l[l];
        float c[2][4][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) {
            const float b0 = (nb * 8 < ML.Np) ? sWt[ML.b + nb * 8 + 2 * t] : 0.f;
            const float b1 = (nb * 8 < ML.Np) ? sWt[ML.b + nb * 8 + 2 * t + 1] : 0.f;
            c[mb][nb][0] = b0; c[mb][nb][1] = b1; c[mb][nb][2] = b0; c[mb][nb][3] = b1;
          }
        gemm_rows<NS>(c, bAct[l], ML.Kp, ML.Np, sWt + ML.w_hi, sWt + ML.w_lo, ML.LDB, g, t);
        const bool use_tanh = (l < L - 1) || pd.out_tanh;
        float* out = (l < L - 1) ? bAct[l + 1] : bMu;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
            if (nb * 8 < ML.Np) {
              float v0 = c[mb][nb][0], v1 = c[mb][nb][1], v2 = c[mb][nb][2], v3 = c[mb][nb][3];
              if (use_tanh) { v0 = tanh_fast(v0); v1 = tanh_fast(v1); v2 = tanh_fast(v2); v3 = tanh_fast(v3); }
              float* o0 = out + (mb * 16 + g) * MMA_LDA + nb * 8 + 2 * t;
              *reinterpret_cast<float2*>(o0) = make_float2(v0, v1);
              *reinterpret_cast<float2*>(o0 + 8 * MMA_LDA) = make_float2(v2, v3);
            }
        // (columns >= Np of the tile, read by the next layer's outer product up to pad16(nout),
        // stay zero from the one-time fill)
        __syncwarp();
      }

    float* dcur;   // delta at the output pre-activation [32][Np_last]
    const int NpL = lay.l[L - 1].Np;
    if (MODE == MODE_FVP) {
      // tangent forward: t_out = (a_in V + vb + t_in W) * act'(a_out)
      float* tin = nullptr;
#pragma unroll
      for (int l = 0; l < MMA_MAXL; ++l)
        if (l < L) {
          const MmaLayer& ML = lay.l[l];
          float c[2][4][4];
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
              const float b0 = (nb * 8 < ML.Np) ? sWt[ML.vb + nb * 8 + 2 * t] : 0.f;
              const float b1 = (nb * 8 < ML.Np) ? sWt[ML.vb + nb * 8 + 2 * t + 1] : 0.f;
              c[mb][nb][0] = b0; c[mb][nb][1] = b1; c[mb][nb][2] = b0; c[mb][nb][3] = b1;
            }
          gemm_rows<NS>(c, bAct[l], ML.Kp, ML.Np, sWt + ML.v_hi, sWt + ML.v_lo, ML.LDB, g, t);
          if (l > 0) gemm_rows<NS>(c, tin, ML.Kp, ML.Np, sWt + ML.w_hi, sWt + ML.w_lo, ML.LDB, g, t);
          const bool last = (l == L - 1);
          const bool use_tanh = !last || pd.out_tanh;
          const float* aout = last ? bMu : bAct[l + 1];
          float* tout = (l & 1) ? bT1 : bT0;
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
              if (nb * 8 < ML.Np) {
                const int r0 = mb * 16 + g, cc = nb * 8 + 2 * t;
                float v0 = c[mb][nb][0], v1 = c[mb][nb][1], v2 = c[mb][nb][2], v3 = c[mb][nb][3];
                if (use_tanh) {
                  const float2 a0 = *reinterpret_cast<const float2*>(aout + r0 * MMA_LDA + cc);
                  const float2 a1 = *reinterpret_cast<const float2*>(aout + (r0 + 8) * MMA_LDA + cc);
                  v0 *= (1.f - a0.x * a0.x); v1 *= (1.f - a0.y * a0.y);
                  v2 *= (1.f - a1.x * a1.x); v3 *= (1.f - a1.y * a1.y);
                }
                if (last) {
                  // delta = M * mu_dot (* act'), zero for masked samples
                  const float m0 = sMisc[cc], m1 = sMisc[cc + 1];
                  const float k0 = ((okmask >> r0) & 1u) ? 1.f : 0.f, k1 = ((okmask >> (r0 + 8)) & 1u) ? 1.f : 0.f;
                  v0 *= m0 * k0; v1 *= m1 * k0; v2 *= m0 * k1; v3 *= m1 * k1;
                  if (pd.out_tanh) {   // back through the output tanh (the tangent above went forward through it)
                    const float2 a0 = *reinterpret_cast<const float2*>(aout + r0 * MMA_LDA + cc);
                    const float2 a1 = *reinterpret_cast<const float2*>(aout + (r0 + 8) * MMA_LDA + cc);
                    v0 *= (1.f - a0.x * a0.x); v1 *= (1.f - a0.y * a0.y);
                    v2 *= (1.f - a1.x * a1.x); v3 *= (1.f - a1.y * a1.y);
                  }
                }
                *reinterpret_cast<float2*>(tout + r0 * MMA_LDA + cc) = make_float2(v0, v1);
                *reinterpret_cast<float2*>(tout + (r0 + 8) * MMA_LDA + cc) = make_float2(v2, v3);
              }
          __syncwarp();
          tin = tout;
        }
      dcur = tin;
      if (ok) t_cnt += 1.0;
    } else {
      // likelihood ratio and KL of sample `lane` (DiagonalGaussian, A.3)
      float ll_new = 0.f, ll_old = 0.f, kl = 0.f;
      const float ad = inb ? p.adv[ng] : 0.f;
      float zs[32];
#pragma unroll
      for (int a = 0; a < 32; ++a)
        if (a < A) {
          const float m = bMu[lane * MMA_LDA + a];
          const float x = inb ? p.act[ng * A + a] : 0.f;
          const float om = inb ? p.old_mean[ng * A + a] : 0.f;
          const float ols = p.old_log_std[((inb && p.old_ls_stride) ? ng * p.old_ls_stride : 0) + a];
          const float ls = fmaxf(sMisc[a], -13.815510557964274f);   // min_std 1e-6
          const float sg = expf(ls), osg = expf(ols);
          const float z = (x - m) / sg, zo = (x - om) / osg;
          ll_new += -ls - 0.5f * z * z;
          ll_old += -ols - 0.5f * zo * zo;
          kl += ((om - m) * (om - m) + osg * osg - sg * sg) / (2.f * sg * sg + 1e-8f) + ls - ols;
          zs[a] = z;
        }
      const float lr = expf(ll_new - ll_old);
      if (ok) { t_surr += static_cast<double>(lr) * ad; t_kl += kl; t_cnt += 1.0; }
      dcur = bT0;
      if (MODE == MODE_GRAD) {
        const float cf = ok ? -ad * lr : 0.f;       // d(-lr*adv)/d ll_new
#pragma unroll
        for (int a = 0; a < 32; ++a)
          if (a < NpL) {
            float dv = 0.f, cl = 0.f;
            if (a < A) {
              const float lsr = sMisc[a];
              const float sg = expf(fmaxf(lsr, -13.815510557964274f));
              dv = cf * zs[a] / sg;                  // d ll / d mu = z / sigma
              if (pd.out_tanh) { const float m = bMu[lane * MMA_LDA + a]; dv *= (1.f - m * m); }
              cl = (lsr > -13.815510557964274f) ? cf * (zs[a] * zs[a] - 1.f) : 0.f;   // d ll / d log_std
            }
            bT0[lane * MMA_LDA + a] = dv;
            if (a < A) {
              cl = static_cast<float>(warp_sum(static_cast<double>(cl)));
              if (lane == a) gls += cl;
            }
          }
        __syncwarp();
      }
    }

    if (MODE != MODE_LOSS) {
      // ---- backward: outer products into the register accumulators, deltas ping-pong ----
#pragma unroll
      for (int l = MMA_MAXL - 1; l >= 0; --l)
        if (l < L) {
          const MmaLayer& ML = lay.l[l];
          gemm_outer<NS>(G[l], bAct[l], ML.Mp, dcur, ML.Np, g, t);
          if (lane < ML.nout) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) s += dcur[r * MMA_LDA + lane];
            gb[l] += s;
          }
          if (l > 0) {
            float c[2][4][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int q = 0; q < 4; ++q) c[mb][nb][q] = 0.f;
            gemm_rows_wt<NS>(c, dcur, ML.Np, ML.Kp, sWt + ML.w_hi, sWt + ML.w_lo, ML.LDB, g, t);
            float* dnext = (dcur == bT0) ? bT1 : bT0;
            const float* ain = bAct[l];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int nb = 0; nb < 4; ++nb)
                if (nb * 8 < ML.Kp) {
                  const int r0 = mb * 16 + g, cc = nb * 8 + 2 * t;
                  const float2 a0 = *reinterpret_cast<const float2*>(ain + r0 * MMA_LDA + cc);
                  const float2 a1 = *reinterpret_cast<const float2*>(ain + (r0 + 8) * MMA_LDA + cc);
                  *reinterpret_cast<float2*>(dnext + r0 * MMA_LDA + cc) =
                      make_float2(c[mb][nb][0] * (1.f - a0.x * a0.x), c[mb][nb][1] * (1.f - a0.y * a0.y));
                  *reinterpret_cast<float2*>(dnext + (r0 + 8) * MMA_LDA + cc) =
                      make_float2(c[mb][nb][2] * (1.f - a1.x * a1.x), c[mb][nb][3] * (1.f - a1.y * a1.y));
                }
            __syncwarp();
            dcur = dnext;
          }
        }
    }
  }

  // ---- flush: warps -> CTA (shared atomics) -> global fp64 atomics ----
  if (MODE != MODE_LOSS) {
#pragma unroll
    for (int l = 0; l < MMA_MAXL; ++l)
      if (l < L) {
        const MmaLayer& ML = lay.l[l];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = mb * 16 + g + ((q & 2) ? 8 : 0), j = nb * 8 + 2 * t + (q & 1);
              if (i < ML.nin && j < ML.nout) atomicAdd(&sG[pd.w_off[l] + i * ML.nout + j], G[l][mb][nb][q]);
            }
        if (lane < ML.nout) atomicAdd(&sG[pd.b_off[l] + lane], gb[l]);
      }
    if (MODE == MODE_GRAD && lane < A) atomicAdd(&sG[pd.logstd_off + lane], gls);
    __syncthreads();
    for (int i = tid; i < pd.P; i += blockDim.x)
      if (sG[i] != 0.f) atomicAdd(&p.acc[i], static_cast<double>(sG[i]));
  }
  t_surr = warp_sum(t_surr); t_kl = warp_sum(t_kl); t_cnt = warp_sum(t_cnt);
  if (lane == 0) { sRed[0][warp] = t_surr; sRed[1][warp] = t_kl; sRed[2][warp] = t_cnt; }
  __syncthreads();
  if (tid < 3) {
    double s = 0.0;
    for (int w = 0; w < MMA_WARPS; ++w) s += sRed[tid][w];
    if (s != 0.0) atomicAdd(&p.acc[pd.P + tid], s);
  }
}

}  // namespace metrpo
