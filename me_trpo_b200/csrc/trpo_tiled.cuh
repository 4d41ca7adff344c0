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
// warp-broadcast) for 32 independent FMAs; layers with <= 8 outputs (the action layer) run one
// thread per sample instead, so that no thread idles.  All deltas of a tile are produced first
// (output delta in place of the mean rows, hidden deltas in the two scratch buffers); the
// parameter-gradient outer products of ALL layers then run as ONE phase: the 4 x 4 (input, output)
// tiles of every layer are dealt out over the 128 threads (120 tiles for 18-32-32-6), each reduced
// over the 128 samples with float4 loads along the sample axis (8 loads per 64 FMAs) into registers
// that live for the whole kernel; bias gradients are row sums of the delta buffers.  Activations keep
// the [feature][sample] layout, so the per-sample likelihood / KL math is unchanged.
#pragma once

namespace metrpo {

constexpr int TILED_NT = 128;
constexpr int TILED_LD = TILED_NT + NTPAD;

inline bool tiled_eligible(const PolDims& pd) {
  for (int i = 0; i <= pd.L; ++i)
    if (pd.d[i] > 32) return false;
  return pd.L >= 1 && pd.L <= 3;   // deltas of <= 2 hidden layers live in the two scratch buffers
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

// c[p][q] += sum_k A[k][smp(p)] * B[k][4 og + q];  thread sg owns samples 4 sg + (0..3) and
// 64 + 4 sg + (0..3): a quarter-warp then touches 8 consecutive float4 of a row (no bank conflicts
// on loads or stores).  A already points at column 4 sg.
__device__ __forceinline__ void tile_accum(float (&c)[8][4], const float* __restrict__ A, int K,
                                           const float* __restrict__ B, int ldb) {
  // packed FFMA2 (two fp32 FMAs per instruction, each IEEE-rounded like fmaf): on sm_100 the
  // three-register scalar FFMA issues at half rate, the packed form restores the full FMA rate
  float2 c2[8][2];
#pragma unroll
  for (int pp = 0; pp < 8; ++pp) { c2[pp][0] = make_float2(c[pp][0], c[pp][1]); c2[pp][1] = make_float2(c[pp][2], c[pp][3]); }
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(A + k * TILED_LD);
    const float4 a1 = *reinterpret_cast<const float4*>(A + k * TILED_LD + 64);
    const float4 w = *reinterpret_cast<const float4*>(B + k * ldb);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) {
      const float2 aa = make_float2(av[pp], av[pp]);
      c2[pp][0] = __ffma2_rn(aa, w01, c2[pp][0]);
      c2[pp][1] = __ffma2_rn(aa, w23, c2[pp][1]);
    }
  }
#pragma unroll
  for (int pp = 0; pp < 8; ++pp) { c[pp][0] = c2[pp][0].x; c[pp][1] = c2[pp][0].y; c[pp][2] = c2[pp][1].x; c[pp][3] = c2[pp][1].y; }
}

// two products sharing the A operand: cw += A * Bw, cv += A * Bv  (forward + tangent of one layer)
__device__ __forceinline__ void tile_accum2(float (&cw)[8][4], float (&cv)[8][4], const float* __restrict__ A, int K,
                                            const float* __restrict__ Bw, const float* __restrict__ Bv, int ldb) {
  float2 w2[8][2], v2[8][2];
#pragma unroll
  for (int pp = 0; pp < 8; ++pp) {
    w2[pp][0] = make_float2(cw[pp][0], cw[pp][1]); w2[pp][1] = make_float2(cw[pp][2], cw[pp][3]);
    v2[pp][0] = make_float2(cv[pp][0], cv[pp][1]); v2[pp][1] = make_float2(cv[pp][2], cv[pp][3]);
  }
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(A + k * TILED_LD);
    const float4 a1 = *reinterpret_cast<const float4*>(A + k * TILED_LD + 64);
    const float4 w = *reinterpret_cast<const float4*>(Bw + k * ldb);
    const float4 v = *reinterpret_cast<const float4*>(Bv + k * ldb);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
    const float2 v01 = make_float2(v.x, v.y), v23 = make_float2(v.z, v.w);
#pragma unroll
    for (int pp = 0; pp < 8; ++pp) {
      const float2 aa = make_float2(av[pp], av[pp]);
      w2[pp][0] = __ffma2_rn(aa, w01, w2[pp][0]); w2[pp][1] = __ffma2_rn(aa, w23, w2[pp][1]);
      v2[pp][0] = __ffma2_rn(aa, v01, v2[pp][0]); v2[pp][1] = __ffma2_rn(aa, v23, v2[pp][1]);
    }
  }
#pragma unroll
  for (int pp = 0; pp < 8; ++pp) {
    cw[pp][0] = w2[pp][0].x; cw[pp][1] = w2[pp][0].y; cw[pp][2] = w2[pp][1].x; cw[pp][3] = w2[pp][1].y;
    cv[pp][0] = v2[pp][0].x; cv[pp][1] = v2[pp][0].y; cv[pp][2] = v2[pp][1].x; cv[pp][3] = v2[pp][1].y;
  }
}

// per-sample evaluation of a layer with <= 8 outputs: c[j] += sum_k A[k][tid] * B[k][j]
__device__ __forceinline__ void sample_accum(float (&c)[8], const float* __restrict__ A, int K,
                                             const float* __restrict__ B, int ldb) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = A[k * TILED_LD];
    const float2 aa = make_float2(a, a);
    const float4 w0 = *reinterpret_cast<const float4*>(B + k * ldb);
    float2 r0 = __ffma2_rn(aa, make_float2(w0.x, w0.y), make_float2(c[0], c[1]));
    float2 r1 = __ffma2_rn(aa, make_float2(w0.z, w0.w), make_float2(c[2], c[3]));
    c[0] = r0.x; c[1] = r0.y; c[2] = r1.x; c[3] = r1.y;
    if (ldb > 4) {
      const float4 w1 = *reinterpret_cast<const float4*>(B + k * ldb + 4);
      r0 = __ffma2_rn(aa, make_float2(w1.x, w1.y), make_float2(c[4], c[5]));
      r1 = __ffma2_rn(aa, make_float2(w1.z, w1.w), make_float2(c[6], c[7]));
      c[4] = r0.x; c[5] = r0.y; c[6] = r1.x; c[7] = r1.y;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(TILED_NT, 2) policy_pass_tiled_kernel(const __grid_constant__ PassParams p) {
  constexpr int NT = TILED_NT, LD = TILED_LD;
  extern __shared__ __align__(16) float sm[];
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
  const PolDims& pd = p.pd;
  const int tid = threadIdx.x, sg = tid & 15, og = tid >> 4, L = pd.L, A = pd.d[L], S = pd.d[0];
  const int s0 = 4 * sg;   // this thread's samples: s0 .. s0+3 and 64+s0 .. 64+s0+3
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

  // delta at the OUTPUT of layer l: the mean rows for the last layer, then the scratch buffers
  float* mu_rows = sAct + pd.act_row[L] * LD;
  auto dbuf = [&](int l) -> float* { return l == L - 1 ? mu_rows : (l == L - 2 ? sBufA : sBufB); };

  // ---- deal the 4 x 4 gradient tiles of all layers out over the threads (<= 2 per thread) ----
  int it_l[2], it_i[2], it_j[2];
  const float* it_a[2][4];
  const float* it_d[2][4];
  float g[2][16];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    it_l[s2] = -1; it_i[s2] = 0; it_j[s2] = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) g[s2][q] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { it_a[s2][q] = sZero; it_d[s2][q] = sZero; }
    if (MODE != MODE_LOSS) {
      int w = tid + NT * s2;
      for (int l = 0; l < L; ++l) {
        const int nin = pd.d[l], nout = pd.d[l + 1];
        const int tiles_j = (nout + 3) >> 2, tiles_i = (nin + 3) >> 2;
        if (w >= 0 && w < tiles_i * tiles_j) {
          it_l[s2] = l; it_i[s2] = w / tiles_j; it_j[s2] = w - it_i[s2] * tiles_j;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // strided tiles: a quarter-warp (same ti, 8 consecutive tj) reads 8 consecutive delta rows
            const int i = it_i[s2] + tiles_i * q, j = it_j[s2] + tiles_j * q;
            it_a[s2][q] = i < nin ? sAct + (pd.act_row[l] + i) * LD : sZero;
            it_d[s2][q] = j < nout ? dbuf(l) + j * LD : sZero;
          }
          w = -1;
        } else if (w >= 0) {
          w -= tiles_i * tiles_j;
        }
      }
    }
  }
  // bias gradient row owned by this thread: the tid-th output over all layers
  int gb_l = -1, gb_j = 0;
  const float* gb_row = sZero;
  float gb = 0.f, gls = 0.f;   // gls: GRAD log_std gradient entry `tid` (tid < A)
  if (MODE != MODE_LOSS) {
    int w = tid;
    for (int l = 0; l < L; ++l) {
      if (w >= 0 && w < pd.d[l + 1]) { gb_l = l; gb_j = w; gb_row = dbuf(l) + w * LD; w = -1; }
      else if (w >= 0) w -= pd.d[l + 1];
    }
  }

  double t_surr = 0.0, t_kl = 0.0, t_cnt = 0.0;
  const long long n_tiles = (p.N + NT - 1) / NT;
  // hidden-activation cache (PassParams::act_cache): rows act_row[1] .. act_row[L]-1 of sAct, per tile
  const int hid_rows = pd.sum_d - S - A;
  const bool cached = (MODE != MODE_LOSS) && p.act_cache != nullptr && hid_rows > 0;
  const int obs_smp0 = tid / S, obs_f0 = tid - obs_smp0 * S, obs_dsmp = NT / S, obs_df = NT - obs_dsmp * S;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n0 = tile * NT, ng = n0 + tid;
    const bool inb = ng < p.N;
    const bool ok = inb && (p.valid == nullptr || p.valid[ng] != 0);
    sOk[tid] = ok ? 1.f : 0.f;
    {  // observations: coalesced read of the tile's [NT,S] block, stored [feature][sample]; element
       // q = tid + NT*i maps to (sample, feature) = (q / S, q % S), advanced without divisions
      const long long base = n0 * S, lim = p.N * S;
      int smp = obs_smp0, f = obs_f0;
      for (int q = tid; q < NT * S; q += NT) {
        sAct[f * LD + smp] = (base + q < lim) ? p.obs[base + q] : 0.f;
        smp += obs_dsmp; f += obs_df;
        if (f >= S) { f -= S; ++smp; }
      }
    }
    if (MODE == MODE_FVP && cached) {   // hidden activations of this tile, stored by the gradient pass
      const float4* src = reinterpret_cast<const float4*>(p.act_cache + static_cast<size_t>(tile) * hid_rows * NT);
      float* dst = sAct + pd.act_row[1] * LD;
      for (int q = tid; q < hid_rows * (NT / 4); q += NT) {
        const int r = q / (NT / 4), c4 = q - r * (NT / 4);
        *reinterpret_cast<float4*>(dst + r * LD + 4 * c4) = __ldg(src + q);
      }
    }
    __syncthreads();
    // ---- forward (training.py:99-103); in FVP mode it is fused with the tangent pass below ----
#pragma unroll 1
    for (int l = 0; l < (MODE == MODE_FVP ? 0 : L); ++l) {
      const int nin = pd.d[l], nout = pd.d[l + 1], np = pd.np[l];
      const bool use_tanh = (l < L - 1) || pd.out_tanh;
      if (np <= 8) {          // one thread per sample
        float c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = j < np ? sW[pd.sb_off[l] + j] : 0.f;
        sample_accum(c, sAct + pd.act_row[l] * LD + tid, nin, sW + pd.sw_off[l], np);
        float* out = sAct + pd.act_row[l + 1] * LD + tid;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nout) out[j * LD] = use_tanh ? tanh_fast(c[j]) : c[j];
      } else if (4 * og < np) {
        float c[8][4];
        const float4 b = *reinterpret_cast<const float4*>(sW + pd.sb_off[l] + 4 * og);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) { c[pp][0] = b.x; c[pp][1] = b.y; c[pp][2] = b.z; c[pp][3] = b.w; }
        tile_accum(c, sAct + pd.act_row[l] * LD + s0, nin, sW + pd.sw_off[l] + 4 * og, np);
        float* out = sAct + pd.act_row[l + 1] * LD + s0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * og + q < nout) {
            float v[8];
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) v[pp] = use_tanh ? tanh_fast(c[pp][q]) : c[pp][q];
            float4* o4 = reinterpret_cast<float4*>(out + (4 * og + q) * LD);
            o4[0] = make_float4(v[0], v[1], v[2], v[3]);
            o4[16] = make_float4(v[4], v[5], v[6], v[7]);
          }
      }
      __syncthreads();
    }
    if (MODE == MODE_GRAD && cached) {
      float4* dstc = reinterpret_cast<float4*>(p.act_cache + static_cast<size_t>(tile) * hid_rows * NT);
      const float* srcc = sAct + pd.act_row[1] * LD;
      for (int q = tid; q < hid_rows * (NT / 4); q += NT) {
        const int r = q / (NT / 4), c4 = q - r * (NT / 4);
        dstc[q] = *reinterpret_cast<const float4*>(srcc + r * LD + 4 * c4);
      }
    }
    const float* mu = mu_rows;
    float* cLs = sBufB;    // GRAD: per-sample d(-lr*adv)/d log_std_a, [A][LD]

    if (MODE == MODE_FVP) {
      // forward and tangent of a layer in ONE phase (they share the layer input):
      //   a_out = act(a_in W + b),   t_out = (a_in V + vb + t_in W) * act'(a_out);
      // the last layer's tangent, scaled by M = d^2 kl / d mu^2, is the output delta (mean rows)
      const float* tin = nullptr;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const int nin = pd.d[l], nout = pd.d[l + 1], np = pd.np[l];
        const bool last = (l == L - 1);
        float* tout = last ? mu_rows : ((l & 1) ? sBufB : sBufA);
        const bool use_tanh = !last || pd.out_tanh;
        if (np <= 8) {
          float cf[8], c[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { cf[j] = j < np ? sW[pd.sb_off[l] + j] : 0.f; c[j] = j < np ? sV[pd.sb_off[l] + j] : 0.f; }
          // the layer's own activation: hidden layers take it from the cache when there is one, the
          // last layer needs it only behind an output tanh
          const bool need_fwd = last ? use_tanh : !cached;
          if (need_fwd) sample_accum(cf, sAct + pd.act_row[l] * LD + tid, nin, sW + pd.sw_off[l], np);
          sample_accum(c, sAct + pd.act_row[l] * LD + tid, nin, sV + pd.sw_off[l], np);
          if (l > 0) sample_accum(c, tin + tid, nin, sW + pd.sw_off[l], np);
          float* aout = sAct + pd.act_row[l + 1] * LD + tid;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < nout) {
              float v = c[j];
              const float av = need_fwd ? (use_tanh ? tanh_fast(cf[j]) : cf[j]) : (last ? 0.f : aout[j * LD]);
              if (use_tanh) {
                v *= (1.f - av * av);
                if (last) v *= (1.f - av * av);   // back through the output tanh as well
              }
              if (last) {
                const float ls = fmaxf(sLs[j], -13.815510557964274f);
                v *= (2.f / (2.f * __expf(2.f * ls) + 1e-8f)) * (ok ? 1.f : 0.f);   // kl_sym, A.3
              } else {
                aout[j * LD] = av;
              }
              tout[j * LD + tid] = v;
            }
        } else if (4 * og < np) {
          float cf[8][4], c[8][4];
          const float4 bw = *reinterpret_cast<const float4*>(sW + pd.sb_off[l] + 4 * og);
          const float4 b = *reinterpret_cast<const float4*>(sV + pd.sb_off[l] + 4 * og);
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            cf[pp][0] = bw.x; cf[pp][1] = bw.y; cf[pp][2] = bw.z; cf[pp][3] = bw.w;
            c[pp][0] = b.x; c[pp][1] = b.y; c[pp][2] = b.z; c[pp][3] = b.w;
          }
          const bool need_fwd = last ? use_tanh : !cached;
          if (need_fwd) tile_accum2(cf, c, sAct + pd.act_row[l] * LD + s0, nin, sW + pd.sw_off[l] + 4 * og, sV + pd.sw_off[l] + 4 * og, np);
          else tile_accum(c, sAct + pd.act_row[l] * LD + s0, nin, sV + pd.sw_off[l] + 4 * og, np);
          if (l > 0) tile_accum(c, tin + s0, nin, sW + pd.sw_off[l] + 4 * og, np);
          float* aout = sAct + pd.act_row[l + 1] * LD + s0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * og + q < nout) {
              const int j = 4 * og + q;
              float v[8], av[8];
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) { v[pp] = c[pp][q]; av[pp] = need_fwd ? (use_tanh ? tanh_fast(cf[pp][q]) : cf[pp][q]) : 0.f; }
              if (!need_fwd && !last) {   // cached activation of this hidden unit
                const float4 a0 = *reinterpret_cast<const float4*>(aout + j * LD);
                const float4 a1 = *reinterpret_cast<const float4*>(aout + j * LD + 64);
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
              }
              if (use_tanh) {
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) v[pp] *= (1.f - av[pp] * av[pp]);
                if (last) {
#pragma unroll
                  for (int pp = 0; pp < 8; ++pp) v[pp] *= (1.f - av[pp] * av[pp]);
                }
              }
              if (last) {
                const float ls = fmaxf(sLs[j], -13.815510557964274f);
                const float m = 2.f / (2.f * __expf(2.f * ls) + 1e-8f);
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) v[pp] *= m * sOk[(pp < 4 ? s0 : 60 + s0) + pp];
              } else if (need_fwd) {
                float4* a4 = reinterpret_cast<float4*>(aout + j * LD);
                a4[0] = make_float4(av[0], av[1], av[2], av[3]);
                a4[16] = make_float4(av[4], av[5], av[6], av[7]);
              }
              float4* o4 = reinterpret_cast<float4*>(tout + j * LD + s0);
              o4[0] = make_float4(v[0], v[1], v[2], v[3]);
              o4[16] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        tin = tout;
      }
      if (ok) t_cnt += 1.0;
    } else {
      // likelihood ratio and KL of this thread's sample (DiagonalGaussian, A.3)
      float ll_new = 0.f, ll_old = 0.f, kl = 0.f;
      float zn[32];
      // blocks of 8 actions: whole blocks beyond A are skipped by a uniform branch (a fully
      // unrolled, predicated 32-iteration loop would still issue all its instructions)
#pragma unroll
      for (int a0 = 0; a0 < 32; a0 += 8)
       if (a0 < A) {
#pragma unroll
        for (int a = a0; a < a0 + 8; ++a)
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
       }
      const float lr = expf(ll_new - ll_old);
      const float ad = inb ? p.adv[ng] : 0.f;
      if (ok) { t_surr += static_cast<double>(lr) * ad; t_kl += kl; t_cnt += 1.0; }
      if (MODE == MODE_GRAD) {
        const float cf = ok ? -ad * lr : 0.f;       // d(-lr*adv)/d ll_new
#pragma unroll
        for (int a0 = 0; a0 < 32; a0 += 8)
         if (a0 < A) {
#pragma unroll
          for (int a = a0; a < a0 + 8; ++a)
          if (a < A) {
            const float lsr = sLs[a];
            const float sgm = expf(fmaxf(lsr, -13.815510557964274f));
            const float m = mu[a * LD + tid];
            float dv = cf * zn[a] / sgm;               // d ll / d mu = z / sigma
            if (pd.out_tanh) dv *= (1.f - m * m);
            mu_rows[a * LD + tid] = dv;                // output delta, in place of the mean (this thread's element)
            cLs[a * LD + tid] = (lsr > -13.815510557964274f) ? cf * (zn[a] * zn[a] - 1.f) : 0.f;   // d ll / d log_std
          }
         }
        __syncthreads();
        if (tid < A) {   // log_std entry tid: sum over the tile's samples
          const float4* r4 = reinterpret_cast<const float4*>(cLs + tid * LD);
          float s = 0.f;
          for (int q = 0; q < NT / 4; ++q) { const float4 v = r4[q]; s += (v.x + v.y) + (v.z + v.w); }
          gls += s;
        }
        __syncthreads();   // cLs (sBufB) consumed before it may become a delta buffer
      }
    }

    if (MODE != MODE_LOSS) {
      // ---- deltas of the hidden layers: d_in[i][n] = (sum_j d_out[j][n] W[i][j]) * (1 - a_in[i][n]^2) ----
#pragma unroll 1
      for (int l = L - 1; l >= 1; --l) {
        const int nin = pd.d[l], nout = pd.d[l + 1], nip = (nin + 3) & ~3;
        const float* dcur = dbuf(l);
        float* dnext = dbuf(l - 1);
        if (4 * og < nip) {
          float c[8][4];
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) { c[pp][0] = 0.f; c[pp][1] = 0.f; c[pp][2] = 0.f; c[pp][3] = 0.f; }
          tile_accum(c, dcur + s0, nout, sWT + wt_off[l] + 4 * og, nip);
          const float* ain = sAct + pd.act_row[l] * LD + s0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * og + q < nin) {
              const int i = 4 * og + q;
              const float4 a0 = *reinterpret_cast<const float4*>(ain + i * LD);
              const float4 a1 = *reinterpret_cast<const float4*>(ain + i * LD + 64);
              float4* o4 = reinterpret_cast<float4*>(dnext + i * LD + s0);
              o4[0] = make_float4(c[0][q] * (1.f - a0.x * a0.x), c[1][q] * (1.f - a0.y * a0.y),
                                  c[2][q] * (1.f - a0.z * a0.z), c[3][q] * (1.f - a0.w * a0.w));
              o4[16] = make_float4(c[4][q] * (1.f - a1.x * a1.x), c[5][q] * (1.f - a1.y * a1.y),
                                  c[6][q] * (1.f - a1.z * a1.z), c[7][q] * (1.f - a1.w * a1.w));
            }
        }
        __syncthreads();
      }
      // ---- parameter gradients of all layers in one phase ----
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
        if (it_l[s2] >= 0) {
#pragma unroll 2
          for (int n4 = 0; n4 < NT / 4; ++n4) {
            float4 av[4], dv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              av[q] = *reinterpret_cast<const float4*>(it_a[s2][q] + 4 * n4);
              dv[q] = *reinterpret_cast<const float4*>(it_d[s2][q] + 4 * n4);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; b += 2) {
                float2 s = make_float2(g[s2][4 * a + b], g[s2][4 * a + b + 1]);
                s = __ffma2_rn(make_float2(av[a].x, av[a].x), make_float2(dv[b].x, dv[b + 1].x), s);
                s = __ffma2_rn(make_float2(av[a].y, av[a].y), make_float2(dv[b].y, dv[b + 1].y), s);
                s = __ffma2_rn(make_float2(av[a].z, av[a].z), make_float2(dv[b].z, dv[b + 1].z), s);
                s = __ffma2_rn(make_float2(av[a].w, av[a].w), make_float2(dv[b].w, dv[b + 1].w), s);
                g[s2][4 * a + b] = s.x; g[s2][4 * a + b + 1] = s.y;
              }
          }
        }
      if (gb_l >= 0) {   // bias gradient: row sum of the delta buffer
        const float4* r4 = reinterpret_cast<const float4*>(gb_row);
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < NT / 4; ++q) { const float4 v = r4[q]; s += (v.x + v.y) + (v.z + v.w); }
        gb += s;
      }
    }
    __syncthreads();   // tile buffers are reused by the next tile
  }

  // ---- flush: register tiles -> global fp64 accumulators ----
  if (MODE != MODE_LOSS) {
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
      if (it_l[s2] >= 0) {
        const int l = it_l[s2], nin = pd.d[l], nout = pd.d[l + 1];
        const int tiles_j = (nout + 3) >> 2, tiles_i = (nin + 3) >> 2;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = it_i[s2] + tiles_i * a, j = it_j[s2] + tiles_j * b;
            const float v = g[s2][4 * a + b];
            if (i < nin && j < nout && v != 0.f) atomicAdd(&p.acc[pd.w_off[l] + i * nout + j], static_cast<double>(v));
          }
      }
    if (gb_l >= 0 && gb != 0.f) atomicAdd(&p.acc[pd.b_off[gb_l] + gb_j], static_cast<double>(gb));
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
