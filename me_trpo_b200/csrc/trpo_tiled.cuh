// Register-tiled version of the TRPO sample pass (loss / gradient / Fisher-vector product) for the
// reference's policy shapes (every layer width <= 32; params/*.json use 32-32).  Default
// implementation for those shapes (METRPO_TRPO_PASS_AUTO); fp32 FMA throughout, same accumulation
// order over the input index as the thread-per-sample kernel it replaces, so forward values are
// bit-identical.
//
// The thread-per-sample kernel (policy_pass_kernel) spends its time on shared-memory operand traffic
// (5 loads per 16 FMAs, one dependent chain per output) and on per-entry dot products for the
// parameter gradients (1 load per 2 FMAs): IPC ~0.2.  Here every dense layer of a 128-sample tile is
// a small GEMM with a classic 2-D register tile: thread (sg, og) owns 8 consecutive samples x 4
// consecutive outputs, so one step of the reduction is 3 shared-memory float4 loads (two of them
// warp-broadcast) for 32 independent FMAs.  The parameter-gradient outer products use a 4 x 4
// (input, output) tile per thread reduced over the 128 samples with float4 loads along the sample
// axis (8 loads per 64 FMAs) and stay in registers for the whole kernel.  Activations keep the
// [feature][sample] layout, so the per-sample likelihood / KL math is unchanged.
#pragma once

namespace metrpo {

constexpr int TILED_NT = 128;
constexpr int TILED_LD = TILED_NT + NTPAD;

inline bool tiled_eligible(const PolDims& pd) {
  for (int i = 0; i <= pd.L; ++i)
    if (pd.d[i] > 32) return false;
  return pd.L >= 1 && pd.L <= TP_MAXL;
}
// transposed weights W^T[l]: [nout][pad4(nin)] (GRAD / FVP back-propagation of deltas)
__host__ __device__ inline int tiled_wt_floats(const PolDims& pd) {
  int n = 0;
  for (int l = 0; l < pd.L; ++l) n += pd.d[l + 1] * ((pd.d[l] + 3) & ~3);
  return (n + 3) & ~3;
}
inline size_t tiled_smem_bytes(const PolDims& pd, int mode) {
  size_t fl = static_cast<size_t>(pd.P_pad) * (mode == MODE_FVP ? 2 : 1) + (mode == MODE_LOSS ? 0 : tiled_wt_floats(pd)) +
              static_cast<size_t>(pd.sum_d + 2 * pd.max_d + 2) * TILED_LD + TILED_NT;
  return fl * 4;
}

// c[p][q] += sum_k A[k][8 sg + p] * B[k][4 og + q]
__device__ __forceinline__ void tile_accum(float (&c)[8][4], const float* __restrict__ A, int K,
                                           const float* __restrict__ B, int ldb) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(A + k * TILED_LD);
    const float4 a1 = *reinterpret_cast<const float4*>(A + k * TILED_LD + 4);
    const float4 w = *reinterpret_cast<const float4*>(B + k * ldb);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) {
      c[pp][0] = fmaf(av[pp], w.x, c[pp][0]); c[pp][1] = fmaf(av[pp], w.y, c[pp][1]);
      c[pp][2] = fmaf(av[pp], w.z, c[pp][2]); c[pp][3] = fmaf(av[pp], w.w, c[pp][3]);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(TILED_NT, 2) policy_pass_tiled_kernel(const __grid_constant__ PassParams p) {
  constexpr int NT = TILED_NT, LD = TILED_LD;
  extern __shared__ __align__(16) float sm[];
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
  const PolDims& pd = p.pd;
  const int tid = threadIdx.x, sg = tid >> 3, og = tid & 7, L = pd.L, A = pd.d[L], S = pd.d[0];
  float* sW = sm;
  float* sV = sW + pd.P_pad;                                     // FVP only
  float* sWT = sV + (MODE == MODE_FVP ? pd.P_pad : 0);           // GRAD / FVP
  float* sAct = sWT + (MODE == MODE_LOSS ? 0 : tiled_wt_floats(pd));
  float* sBufA = sAct + pd.sum_d * LD;
  float* sBufB = sBufA + pd.max_d * LD;
  float* sOnes = sBufB + pd.max_d * LD;
  float* sZero = sOnes + LD;
  float* sOk = sZero + LD;                                       // [NT] 1 / 0 per sample
  __shared__ double sRed[3][NT / 32];
  __shared__ float sLs[32];   // raw log_std parameters

  // ---- stage parameters (zero padded) ----
  for (int i = tid; i < pd.P_pad; i += NT) { sW[i] = 0.f; if (MODE == MODE_FVP) sV[i] = 0.f; }
  if (MODE != MODE_LOSS)
    for (int i = tid; i < tiled_wt_floats(pd); i += NT) sWT[i] = 0.f;
  for (int i = tid; i < LD; i += NT) { sOnes[i] = 1.f; sZero[i] = 0.f; }
  __syncthreads();
  int wt_off[TP_MAXL];
  {
    int o = 0;
    for (int l = 0; l < L; ++l) { wt_off[l] = o; o += pd.d[l + 1] * ((pd.d[l] + 3) & ~3); }
  }
  for (int l = 0; l < L; ++l) {
    const int nin = pd.d[l], nout = pd.d[l + 1], nip = (nin + 3) & ~3;
    for (int e = tid; e < nin * nout; e += NT) {
      const int i = e / nout, j = e - i * nout;
      const float w = p.theta[pd.w_off[l] + e];
      sW[pd.sw_off[l] + i * pd.np[l] + j] = w;
      if (MODE != MODE_LOSS) sWT[wt_off[l] + j * nip + i] = w;
      if (MODE == MODE_FVP) sV[pd.sw_off[l] + i * pd.np[l] + j] = p.vec[pd.w_off[l] + e];
    }
    for (int j = tid; j < nout; j += NT) {
      sW[pd.sb_off[l] + j] = p.theta[pd.b_off[l] + j];
      if (MODE == MODE_FVP) sV[pd.sb_off[l] + j] = p.vec[pd.b_off[l] + j];
    }
  }
  if (tid < 32) sLs[tid] = tid < A ? p.theta[pd.logstd_off + tid] : 0.f;
  __syncthreads();

  // parameter-gradient tiles owned by this thread: layer l, rows 4 ti .. (row nin = bias), cols 4 tj ..
  float g[TP_MAXL][16];
  float gls = 0.f;   // GRAD: log_std gradient entry `tid` (tid < A)
#pragma unroll
  for (int l = 0; l < TP_MAXL; ++l)
#pragma unroll
    for (int q = 0; q < 16; ++q) g[l][q] = 0.f;

  double t_surr = 0.0, t_kl = 0.0, t_cnt = 0.0;
  const long long n_tiles = (p.N + NT - 1) / NT;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n0 = tile * NT, ng = n0 + tid;
    const bool inb = ng < p.N;
    const bool ok = inb && (p.valid == nullptr || p.valid[ng] != 0);
    sOk[tid] = ok ? 1.f : 0.f;
    {  // observations: coalesced read of the tile's [NT,S] block, stored [feature][sample]
      const long long base = n0 * S, lim = p.N * S;
      for (int q = tid; q < NT * S; q += NT) {
        const int smp = q / S, f = q - smp * S;
        sAct[f * LD + smp] = (base + q < lim) ? p.obs[base + q] : 0.f;
      }
    }
    __syncthreads();
    // ---- forward (training.py:99-103) ----
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const int nin = pd.d[l], nout = pd.d[l + 1], np = pd.np[l];
      if (4 * og < np) {
        float c[8][4];
        const float4 b = *reinterpret_cast<const float4*>(sW + pd.sb_off[l] + 4 * og);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) { c[pp][0] = b.x; c[pp][1] = b.y; c[pp][2] = b.z; c[pp][3] = b.w; }
        tile_accum(c, sAct + pd.act_row[l] * LD + 8 * sg, nin, sW + pd.sw_off[l] + 4 * og, np);
        const bool use_tanh = (l < L - 1) || pd.out_tanh;
        float* out = sAct + pd.act_row[l + 1] * LD + 8 * sg;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * og + q < nout) {
            float v[8];
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) v[pp] = use_tanh ? tanh_fast(c[pp][q]) : c[pp][q];
            float4* o4 = reinterpret_cast<float4*>(out + (4 * og + q) * LD);
            o4[0] = make_float4(v[0], v[1], v[2], v[3]);
            o4[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
      }
      __syncthreads();
    }
    const float* mu = sAct + pd.act_row[L] * LD;
    float* dOut = sBufA;   // delta at the output pre-activation, [A][LD]
    float* cLs = sBufB;    // GRAD: per-sample d(-lr*adv)/d log_std_a, [A][LD]

    if (MODE == MODE_FVP) {
      // tangent forward: t_out = (a_in V + vb + t_in W) * act'(a_out)
      const float* tin = nullptr;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const int nin = pd.d[l], nout = pd.d[l + 1], np = pd.np[l];
        float* tout = (l & 1) ? sBufB : sBufA;
        const bool last = (l == L - 1);
        if (4 * og < np) {
          float c[8][4];
          const float4 b = *reinterpret_cast<const float4*>(sV + pd.sb_off[l] + 4 * og);
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) { c[pp][0] = b.x; c[pp][1] = b.y; c[pp][2] = b.z; c[pp][3] = b.w; }
          tile_accum(c, sAct + pd.act_row[l] * LD + 8 * sg, nin, sV + pd.sw_off[l] + 4 * og, np);
          if (l > 0) tile_accum(c, tin + 8 * sg, nin, sW + pd.sw_off[l] + 4 * og, np);
          const bool use_tanh = !last || pd.out_tanh;
          const float* aout = sAct + pd.act_row[l + 1] * LD + 8 * sg;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * og + q < nout) {
              const int j = 4 * og + q;
              float v[8];
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) v[pp] = c[pp][q];
              if (use_tanh) {
                const float4 a0 = *reinterpret_cast<const float4*>(aout + j * LD);
                const float4 a1 = *reinterpret_cast<const float4*>(aout + j * LD + 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) v[pp] *= (1.f - av[pp] * av[pp]);
                if (last) {   // back through the output tanh as well
#pragma unroll
                  for (int pp = 0; pp < 8; ++pp) v[pp] *= (1.f - av[pp] * av[pp]);
                }
              }
              if (last) {
                // delta = M * mu_dot with M = d^2 kl / d mu^2 = 2 / (2 sigma^2 + 1e-8)   (kl_sym, A.3)
                const float ls = fmaxf(sLs[j], -13.815510557964274f);
                const float m = 2.f / (2.f * __expf(2.f * ls) + 1e-8f);
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) v[pp] *= m * sOk[8 * sg + pp];
              }
              float4* o4 = reinterpret_cast<float4*>(tout + j * LD + 8 * sg);
              o4[0] = make_float4(v[0], v[1], v[2], v[3]);
              o4[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        tin = tout;
      }
      dOut = const_cast<float*>(tin);
      if (ok) t_cnt += 1.0;
    } else {
      // likelihood ratio and KL of this thread's sample (DiagonalGaussian, A.3)
      float ll_new = 0.f, ll_old = 0.f, kl = 0.f;
      float zn[32];
#pragma unroll
      for (int a = 0; a < 32; ++a)
        if (a < A) {
          const float m = mu[a * LD + tid];
          const float x = inb ? p.act[ng * A + a] : 0.f;
          const float om = inb ? p.old_mean[ng * A + a] : 0.f;
          const float ols = p.old_log_std[((inb && p.old_ls_stride) ? ng * p.old_ls_stride : 0) + a];
          const float ls = fmaxf(sLs[a], -13.815510557964274f);   // min_std 1e-6
          const float sgm = expf(ls), osg = expf(ols);
          const float z = (x - m) / sgm, zo = (x - om) / osg;
          ll_new += -ls - 0.5f * z * z;
          ll_old += -ols - 0.5f * zo * zo;
          kl += ((om - m) * (om - m) + osg * osg - sgm * sgm) / (2.f * sgm * sgm + 1e-8f) + ls - ols;
          zn[a] = z;
        }
      const float lr = expf(ll_new - ll_old);
      const float ad = inb ? p.adv[ng] : 0.f;
      if (ok) { t_surr += static_cast<double>(lr) * ad; t_kl += kl; t_cnt += 1.0; }
      if (MODE == MODE_GRAD) {
        const float cf = ok ? -ad * lr : 0.f;       // d(-lr*adv)/d ll_new
#pragma unroll
        for (int a = 0; a < 32; ++a)
          if (a < A) {
            const float lsr = sLs[a];
            const float sgm = expf(fmaxf(lsr, -13.815510557964274f));
            float dv = cf * zn[a] / sgm;               // d ll / d mu = z / sigma
            if (pd.out_tanh) { const float m = mu[a * LD + tid]; dv *= (1.f - m * m); }
            dOut[a * LD + tid] = dv;
            cLs[a * LD + tid] = (lsr > -13.815510557964274f) ? cf * (zn[a] * zn[a] - 1.f) : 0.f;   // d ll / d log_std
          }
        __syncthreads();
        if (tid < A) {   // log_std entry tid: sum over the tile's samples
          const float4* r4 = reinterpret_cast<const float4*>(cLs + tid * LD);
          float s = 0.f;
          for (int q = 0; q < NT / 4; ++q) { const float4 v = r4[q]; s += (v.x + v.y) + (v.z + v.w); }
          gls += s;
        }
      }
    }

    if (MODE != MODE_LOSS) {
      // ---- backward: outer products into the register tiles, deltas ping-pong between the buffers ----
      if (MODE == MODE_GRAD) __syncthreads();   // cLs (sBufB) fully consumed before it becomes a delta buffer
      float* dcur = dOut;
#pragma unroll
      for (int l = TP_MAXL - 1; l >= 0; --l)
        if (l < L) {
          const int nin = pd.d[l], nout = pd.d[l + 1];
          const int tiles_j = (nout + 3) >> 2, tiles_i = (nin + 4) >> 2;   // rows 0..nin (row nin = bias)
          if (tid < tiles_i * tiles_j) {
            const int ti = tid / tiles_j, tj = tid - ti * tiles_j;
            const float* ar[4];
            const float* dr[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = 4 * ti + q, j = 4 * tj + q;
              ar[q] = i < nin ? sAct + (pd.act_row[l] + i) * LD : (i == nin ? sOnes : sZero);
              dr[q] = j < nout ? dcur + j * LD : sZero;
            }
#pragma unroll 2
            for (int n4 = 0; n4 < NT / 4; ++n4) {
              float4 av[4], dv[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                av[q] = *reinterpret_cast<const float4*>(ar[q] + 4 * n4);
                dv[q] = *reinterpret_cast<const float4*>(dr[q] + 4 * n4);
              }
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  float s = g[l][4 * a + b];
                  s = fmaf(av[a].x, dv[b].x, s); s = fmaf(av[a].y, dv[b].y, s);
                  s = fmaf(av[a].z, dv[b].z, s); s = fmaf(av[a].w, dv[b].w, s);
                  g[l][4 * a + b] = s;
                }
            }
          }
          if (l > 0) {
            // d_in[i][n] = (sum_j d_out[j][n] W[i][j]) * (1 - a_in[i][n]^2), via W^T rows [j][i]
            const int nip = (nin + 3) & ~3;
            float* dnext = (dcur == sBufA) ? sBufB : sBufA;
            if (4 * og < nip) {
              float c[8][4];
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) { c[pp][0] = 0.f; c[pp][1] = 0.f; c[pp][2] = 0.f; c[pp][3] = 0.f; }
              tile_accum(c, dcur + 8 * sg, nout, sWT + wt_off[l] + 4 * og, nip);
              const float* ain = sAct + pd.act_row[l] * LD + 8 * sg;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (4 * og + q < nin) {
                  const int i = 4 * og + q;
                  const float4 a0 = *reinterpret_cast<const float4*>(ain + i * LD);
                  const float4 a1 = *reinterpret_cast<const float4*>(ain + i * LD + 4);
                  float4* o4 = reinterpret_cast<float4*>(dnext + i * LD + 8 * sg);
                  o4[0] = make_float4(c[0][q] * (1.f - a0.x * a0.x), c[1][q] * (1.f - a0.y * a0.y),
                                      c[2][q] * (1.f - a0.z * a0.z), c[3][q] * (1.f - a0.w * a0.w));
                  o4[1] = make_float4(c[4][q] * (1.f - a1.x * a1.x), c[5][q] * (1.f - a1.y * a1.y),
                                      c[6][q] * (1.f - a1.z * a1.z), c[7][q] * (1.f - a1.w * a1.w));
                }
            }
            __syncthreads();   // dnext complete; dcur's readers are done
            dcur = dnext;
          }
        }
    }
    __syncthreads();   // tile buffers are reused by the next tile
  }

  // ---- flush: register tiles -> global fp64 accumulators ----
  if (MODE != MODE_LOSS) {
#pragma unroll
    for (int l = 0; l < TP_MAXL; ++l)
      if (l < L) {
        const int nin = pd.d[l], nout = pd.d[l + 1];
        const int tiles_j = (nout + 3) >> 2, tiles_i = (nin + 4) >> 2;
        if (tid < tiles_i * tiles_j) {
          const int ti = tid / tiles_j, tj = tid - ti * tiles_j;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int i = 4 * ti + a, j = 4 * tj + b;
              const float v = g[l][4 * a + b];
              if (j < nout && v != 0.f) {
                if (i < nin) atomicAdd(&p.acc[pd.w_off[l] + i * nout + j], static_cast<double>(v));
                else if (i == nin) atomicAdd(&p.acc[pd.b_off[l] + j], static_cast<double>(v));
              }
            }
        }
      }
    if (MODE == MODE_GRAD && tid < A && gls != 0.f) atomicAdd(&p.acc[pd.logstd_off + tid], static_cast<double>(gls));
  }
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
