// Persistent ensemble-rollout kernel (sm_100a).
//
// Replaces, for the whole horizon and with no host round trip, the loop
//   VectorizedSampler.obtain_samples  (samplers/vectorized_sampler.py:60-108)
//     -> policy.get_actions           (rllab GaussianMLPPolicy; training.py:96-117)
//     -> VecSimpleEnv.step            (env_helpers.py:597-607)
//          -> get_next_observation    (env_helpers.py:609-635; all K models, then select)
//          -> cost_np_vec / is_done   (envs/com_*_env.py)
//          -> reset                   (env_helpers.py:585-595)
//
// Work decomposition.  One CTA owns (model k, tile of 128 rollout rows); the K CTAs of a "gang
// slot" own the same row tile and advance in lock step: every step each CTA publishes its
// candidate next states [128,S] to an L2-resident exchange buffer, the gang meets on a counter,
// and every CTA applies the sam_mode selection redundantly (all K candidates are needed by
// model_mean/_med/_mean_std, and the next step of every model needs the selected state).
// Rows never interact across tiles, so there is no grid-wide sync.  When there are more row
// tiles than gang slots the host cuts the (tile x horizon) chains into per-slot segment lists
// (rollout_api.cu: build_schedule) so that all SMs stay busy; a tile may migrate between slots
// once, through the row_state buffers + a release/acquire flag.
//
// Inside a CTA (192 threads):
//   warp 0    producer: streams the model's pre-packed bf16 weight tiles L2 -> smem with
//             cp.async.bulk (TMA engine) through a 3-stage mbarrier ring
//   warp 1    MMA issuer: one thread issues tcgen05.mma (M=128, fp32 accumulators in TMEM)
//   warps 2-5 epilogue/compute: TMEM -> registers -> bias+ReLU -> bf16 -> smem operand tiles,
//             plus the per-step serial section (residual/de-normalise, exchange, select,
//             reward/done/reset, trajectory write, policy MLP, normalise -> Z operand tile)
//
// Per step and model the MLP  z[128,K0] -> H -> H -> S  is evaluated as
//   for nc in H/256 output chunks of layer 1:          (acc1: 256 TMEM columns)
//     for kc in H/64 reduction chunks:
//        L0: acc0[b] = Z * W0[:, 64-chunk kc]           (N=64, K=K0; recomputed per nc pass because
//                                                         128 x H activations do not fit on an SM)
//        epilogue: H0[b] = bf16(relu(acc0[b] + b0))     (SW128 K-major A operand, 16 KB)
//        L1: acc1 += H0[b] * W1[kc chunk, nc chunk]     (N=256, K=64: 4 MMAs)
//     epilogue: 4 x (H1[bb] = bf16(relu(acc1[:,64 cols] + b1)));  L2: acc2 += H1[bb] * W2 chunk
//   next_state = (diff_mean + diff_std * (acc2 + b2)) + x          (training.py:257)
#pragma once
#include "umma.cuh"
#include "philox.cuh"

namespace metrpo {

constexpr int TILE_M = 128;
constexpr int NSTAGE = 3;
constexpr int SMAX = 32;      // max state dim held in registers (v1)
constexpr int AMAX = 8;       // max action dim (v1)
constexpr int HPMAX = 32;     // max policy hidden width (v1)
constexpr int MAX_SEG = 32;   // segments per gang slot
constexpr int W1_TILE_BYTES = 256 * 64 * 2;   // 32 KB: [256 n][64 k] bf16, SW128
constexpr int H_TILE_BYTES = 128 * 64 * 2;    // 16 KB: [128 rows][64 k] bf16, SW128
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;

// TMEM column map (512 allocated)
constexpr uint32_t TM_ACC1 = 0;      // 256 cols
constexpr uint32_t TM_ACC0 = 256;    // 2 x 64 cols
constexpr uint32_t TM_ACC2 = 384;    // S_pad (<= 32) cols

enum {
  B_FULL = 0,        // [NSTAGE] weight stage landed (tx)
  B_EMPTY = 3,       // [NSTAGE] weight stage consumed (commit)
  B_W2FULL = 6,
  B_W2EMPTY = 7,
  B_W0RES = 8,
  B_ZREADY = 9,      // Z operand tile written (128 arrivals)
  B_ACC0FULL = 10,   // [2] L0 chunk accumulated (commit)
  B_H0FULL = 12,     // [2] H0 operand tile written (128 arrivals)
  B_H0FREE = 14,     // [2] H0 operand tile consumed (commit)
  B_ACC1FULL = 16,
  B_ACC1FREE = 17,   // acc1 drained to registers (128 arrivals)
  B_H1FULL = 18,     // [2]
  B_H1FREE = 20,     // [2]
  B_ACC2FULL = 22,
  NUM_BARS = 23
};

struct PolicyLayer {
  int nin, nout, npad;   // npad = 32 for hidden layers, 8 for the output layer
  int w_off, b_off;      // float offsets into the policy blob
};

struct KParams {
  // dims
  int S, A, SA, drop, Din, K0, H, S_pad, K, B, T_max, env_id, sam_mode, determ;
  int NC, KC;
  int n_steps, n_slots, n_tiles;
  int resume;            // 1: state comes from row_state (B1 step / continued run)
  // packed weights (per model: stages | W2 chunks | resident W0 tiles)
  const uint8_t* wstream;
  unsigned long long model_stride;
  uint32_t stage_bytes, w0tile_bytes, w2chunk_bytes, off_w2, off_w0res;
  const float* bias;     // [K][2H + 32]: b0 | b1 | b2 (zero padded)
  const float* norm;     // in_mean[SA] | in_std[SA] | diff_mean[S] | diff_std[S]
  const float* pol;      // policy blob (padded W, b per layer, then log_std[AMAX])
  int pol_floats, n_pol_layers, pol_out_tanh, pol_logstd_off;
  PolicyLayer pl[4];
  // inputs
  const float* init_states;
  const float* reset_pool;
  int R;
  const float* eps;
  const int* model_idx;
  const float* std_noise;
  const float* ext_actions;       // B1 step: sampler-provided actions [B,A]
  const float* ext_reset_states;  // B1 step: [B,S]
  unsigned long long seed, offset;
  // outputs
  float* obs; float* act; float* mean; float* rew; uint8_t* done; float* final_states;
  // workspace
  float* xbuf;            // [n_slots][2][K][S][128]
  unsigned* xctr;         // [n_slots]
  float* row_state;       // [n_tiles*128][S]
  int* row_ts;            // [n_tiles*128]
  int* row_nreset;        // [n_tiles*128]
  unsigned* tile_flag;    // [n_tiles]
  unsigned long long* trace;   // optional event trace of one CTA (dev tool), or nullptr
  int trace_cta, trace_t0, trace_t1;
  unsigned* dbg;          // [DBG_HEADER + grid*6*4] abort flag + wait records
  const int4* segs;       // [n_slots][MAX_SEG] = (tile, t0, t1, wait_flag); tile < 0 -> unused
  // smem carve-up (byte offsets from the 1024-aligned base)
  uint32_t off_stage, off_sw0res, off_z, off_h0, off_h1, off_sw2, off_sbias, off_snorm, off_spol,
      off_bars;
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Bounded waits.  dbg[0] is a device-wide abort flag: the first wait that exceeds the timeout
// sets it and records (tag, progress); every other spinning role sees the flag, records where it
// was and leaves.  The kernel then exits cleanly and the host reports the records
// (metrpo_rollout_status) instead of the GPU hanging.
// ---------------------------------------------------------------------------------------------
#ifndef METRPO_WAIT_TIMEOUT_NS
#define METRPO_WAIT_TIMEOUT_NS 1000000000ull
#endif
constexpr int DBG_WORDS_PER_WARP = 4;
constexpr int DBG_HEADER = 16;
__device__ __forceinline__ void dbg_record(unsigned* dbg, uint32_t tag, uint32_t info, uint32_t why) {
  const int idx = DBG_HEADER + (blockIdx.x * (NUM_THREADS / 32) + (threadIdx.x >> 5)) * DBG_WORDS_PER_WARP;
  dbg[idx + 0] = tag;
  dbg[idx + 1] = info;
  dbg[idx + 2] = why;    // 1 = timed out here, 2 = saw the abort flag here
  dbg[idx + 3] = threadIdx.x;
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, unsigned* dbg, uint32_t tag,
                                         uint32_t info) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xff) == 0) {
      if (*reinterpret_cast<volatile unsigned*>(dbg) != 0u) { dbg_record(dbg, tag, info, 2); return false; }
      if (globaltimer_ns() - t0 > METRPO_WAIT_TIMEOUT_NS) {
        atomicExch(dbg, 1u);
        dbg_record(dbg, tag, info, 1);
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ bool wait_ge(const unsigned* ptr, unsigned target, unsigned* dbg, uint32_t tag,
                                        uint32_t info) {
  if (ld_acquire_gpu(ptr) >= target) return true;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (ld_acquire_gpu(ptr) < target) {
    if ((++spins & 0x3f) == 0) {
      if (*reinterpret_cast<volatile unsigned*>(dbg) != 0u) { dbg_record(dbg, tag, info, 2); return false; }
      if (globaltimer_ns() - t0 > METRPO_WAIT_TIMEOUT_NS) {
        atomicExch(dbg, 1u);
        dbg_record(dbg, tag, info, 1);
        return false;
      }
    }
  }
  return true;
}
#define WAITB(idx, par, info) \
  do { if (!wait_bar(&bars[idx], (par), p.dbg, (idx), (info))) goto bail; } while (0)

// optional event trace (dev tool): role-private logs of (code << 40 | clock) for one CTA and a
// window of steps; enabled by metrpo_rollout_set_trace.  role 0 producer, 1 MMA, 2 epilogue.
constexpr int TRACE_CAP = 4096;
#define TRACE(role, code)                                                                    \
  do {                                                                                       \
    if (tr_on && tr_n < TRACE_CAP)                                                           \
      p.trace[(role) * TRACE_CAP + tr_n++] =                                                 \
          (static_cast<unsigned long long>(code) << 40) | (clock64() & 0xFFFFFFFFFFull);     \
  } while (0)

// per-row analytic cost (reward = -cost); u is the clipped action.  envs/com_*_env.py
__device__ __forceinline__ float env_cost(int env_id, int S, int A, const float (&xn)[SMAX],
                                          const float (&u)[AMAX]) {
  float su2 = 0.f;
#pragma unroll
  for (int i = 0; i < AMAX; ++i)
    if (i < A) su2 = __fadd_rn(su2, __fmul_rn(u[i], u[i]));
  switch (env_id) {
    case METRPO_ENV_SWIMMER:   // -(x'[5] - 0.01*mean(u^2))
      return -(xn[5] - 0.01f * (su2 / static_cast<float>(A)));
    case METRPO_ENV_HALF_CHEETAH: {  // -clip(x'[9] - 0.1*0.5*sum(u^2), -10, 10)
      float v = xn[9] - 0.1f * 0.5f * su2;
      return -fminf(fmaxf(v, -10.f), 10.f);
    }
    case METRPO_ENV_HOPPER: {
      float pen = 0.f;
#pragma unroll
      for (int s = 2; s < SMAX; ++s)
        if (s < S) pen += fmaxf(fabsf(xn[s]) - 100.f, 0.f);
      return -(xn[5] - 0.01f * 0.5f * su2 - 10.f * fmaxf(0.45f - xn[0], 0.f) -
               10.f * fmaxf(fabsf(xn[1]) - 0.2f, 0.f) - pen);
    }
    case METRPO_ENV_ANT:       // -(x'[15] - 1e-2*0.5*sum(u^2) + 0.05)
      return -(xn[15] - 0.01f * 0.5f * su2 + 0.05f);
    case METRPO_ENV_HUMANOID: {  // (x'[-1]-1.5)^2 + 1e-5*sum(u^2)   (S <= SMAX only)
      float hh = 0.f;
#pragma unroll
      for (int s = 0; s < SMAX; ++s)
        if (s == S - 1) hh = xn[s];
      return (hh - 1.5f) * (hh - 1.5f) + 1e-2f * 1e-3f * su2;
    }
    default:                   // snake: -(x'[7] - 0.01*0.5*sum(u^2))
      return -(xn[7] - 0.01f * 0.5f * su2);
  }
}
__device__ __forceinline__ bool env_is_done(int env_id, int S, const float (&xn)[SMAX]) {
  if (env_id != METRPO_ENV_ANT) return false;   // NeuralNetEnv default (env_helpers.py:537)
  bool finite = true;
#pragma unroll
  for (int s = 0; s < SMAX; ++s)
    if (s < S) finite = finite && isfinite(xn[s]);
  return !(xn[2] >= 0.2f && xn[2] <= 1.0f && finite);
}

// one dense layer of the policy on CUDA cores: thread r owns column r of the [n][128] scratch
template <int NP>
__device__ __forceinline__ void dense_layer(const float* in_s, int nin, const float* W,
                                            const float* b, float (&acc)[NP], int r) {
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = b[j];
  for (int i = 0; i < nin; ++i) {
    const float xi = in_s[i * TILE_M + r];
    const float4* w4 = reinterpret_cast<const float4*>(W + i * NP);
#pragma unroll
    for (int j = 0; j < NP / 4; ++j) {
      float4 w = w4[j];
      acc[4 * j + 0] = fmaf(xi, w.x, acc[4 * j + 0]);
      acc[4 * j + 1] = fmaf(xi, w.y, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(xi, w.z, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(xi, w.w, acc[4 * j + 3]);
    }
  }
}

// bias + ReLU + bf16 pack of 64 accumulator columns -> one SW128 row (8 x 16 B chunks)
__device__ __forceinline__ void relu_pack_store(const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                                const float* bias, uint8_t* tile, int row) {
  const uint32_t rbase = (row >> 3) * 1024u + (row & 7u) * 128u;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 8 * c + 2 * i;
      float a0 = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]) + bias[col];
      float a1 = __uint_as_float(col + 1 < 32 ? v0[col + 1] : v1[col + 1 - 32]) + bias[col + 1];
      pk[i] = pack_bf16x2(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
    }
    *reinterpret_cast<uint4*>(tile + rbase + (((c ^ row) & 7) << 4)) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// =============================================================================================
__global__ void __launch_bounds__(NUM_THREADS, 1) rollout_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ int abort_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.x / p.K, k = blockIdx.x % p.K;

  uint8_t* sStage = smem + p.off_stage;
  uint8_t* sW0res = smem + p.off_sw0res;
  uint8_t* sZ = smem + p.off_z;
  uint8_t* sH0 = smem + p.off_h0;
  uint8_t* sH1 = smem + p.off_h1;
  uint8_t* sW2 = smem + p.off_sw2;
  float* sBias = reinterpret_cast<float*>(smem + p.off_sbias);
  float* sNorm = reinterpret_cast<float*>(smem + p.off_snorm);
  float* sPol = reinterpret_cast<float*>(smem + p.off_spol);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);

  const int NCH = p.NC * p.KC;          // chunks per step
  const int4* segs = p.segs + slot * MAX_SEG;
  int total_steps = 0;
  for (int i = 0; i < MAX_SEG; ++i) {
    int4 sg = segs[i];
    if (sg.x >= 0) total_steps += sg.z - sg.y;
  }

  // ---- one-time setup ----
  int st_dbg = 0;   // progress counter reported by the wait diagnostics
  bool tr_on = false;
  int tr_n = 0;
  if (tid == 0) {
    abort_smem = 0;
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
    mbar_init(&bars[B_W2FULL], 1); mbar_init(&bars[B_W2EMPTY], 1); mbar_init(&bars[B_W0RES], 1);
    mbar_init(&bars[B_ZREADY], EPI_THREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_ACC0FULL + i], 1); mbar_init(&bars[B_H0FULL + i], EPI_THREADS);
      mbar_init(&bars[B_H0FREE + i], 1); mbar_init(&bars[B_H1FULL + i], EPI_THREADS);
      mbar_init(&bars[B_H1FREE + i], 1);
    }
    mbar_init(&bars[B_ACC1FULL], 1); mbar_init(&bars[B_ACC1FREE], EPI_THREADS);
    mbar_init(&bars[B_ACC2FULL], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  {  // constants -> smem (generic loads; read-only for the rest of the kernel)
    const float* gb = p.bias + static_cast<size_t>(k) * (2 * p.H + 32);
    for (int i = tid; i < 2 * p.H + 32; i += NUM_THREADS) sBias[i] = gb[i];
    for (int i = tid; i < 2 * p.SA + 2 * p.S; i += NUM_THREADS) sNorm[i] = p.norm[i];
    for (int i = tid; i < p.pol_floats; i += NUM_THREADS) sPol[i] = p.pol[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (total_steps > 0) {
    if (warp == 0) {
      // =========================== producer ===========================
      if (lane == 0) {
        const uint8_t* wm = p.wstream + static_cast<size_t>(k) * p.model_stride;
        const uint64_t pol = l2_policy_evict_last();
        mbar_arrive_expect_tx(&bars[B_W0RES], 2 * p.w0tile_bytes);
        bulk_g2s_hint(sW0res, wm + p.off_w0res, 2 * p.w0tile_bytes, &bars[B_W0RES], pol);
        uint32_t gs = 0, w2n = 0;
        const int w2_at = p.KC > 3 ? 3 : p.KC - 1;
        for (int st = 0; st < total_steps; ++st) {
          for (int nc = 0; nc < p.NC; ++nc) {
            for (int kc = 0; kc < p.KC; ++kc) {
              st_dbg = (st << 8) | (nc * p.KC + kc);
              tr_on = p.trace && blockIdx.x == p.trace_cta && st >= p.trace_t0 && st < p.trace_t1;
              const uint32_t s = gs % NSTAGE, n = gs / NSTAGE;
              TRACE(0, 0x100 | (nc * p.KC + kc));
              WAITB(B_EMPTY + s, (n & 1) ^ 1, (uint32_t)st_dbg);
              TRACE(0, 0x200 | (nc * p.KC + kc));
              mbar_arrive_expect_tx(&bars[B_FULL + s], p.stage_bytes);
              bulk_g2s_hint(sStage + s * p.stage_bytes,
                            wm + static_cast<size_t>(nc * p.KC + kc) * p.stage_bytes, p.stage_bytes,
                            &bars[B_FULL + s], pol);
              ++gs;
              if (kc == w2_at) {
                WAITB(B_W2EMPTY, (w2n & 1) ^ 1, (uint32_t)st_dbg);
                mbar_arrive_expect_tx(&bars[B_W2FULL], p.w2chunk_bytes);
                bulk_g2s_hint(sW2, wm + p.off_w2 + static_cast<size_t>(nc) * p.w2chunk_bytes,
                              p.w2chunk_bytes, &bars[B_W2FULL], pol);
                ++w2n;
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // =========================== MMA issuer ===========================
      if (lane == 0) {
        const uint32_t idesc0 = idesc_bf16_f32(128, 64);
        const uint32_t idesc1 = idesc_bf16_f32(128, 256);
        const uint32_t idesc2 = idesc_bf16_f32(128, p.S_pad);
        const int k0steps = p.K0 / 16;
        const uint64_t zdesc = smem_desc_noswz(smem_u32(sZ), TILE_M * 16, 128);
        const uint32_t z_kstep = (2 * TILE_M * 16) >> 4;     // descriptor-lo units per 16-k step
        const uint32_t w0_kstep = (2 * 64 * 16) >> 4;
        const uint64_t h0desc[2] = {smem_desc_sw128(smem_u32(sH0)),
                                    smem_desc_sw128(smem_u32(sH0 + H_TILE_BYTES))};
        const uint64_t h1desc[2] = {smem_desc_sw128(smem_u32(sH1)),
                                    smem_desc_sw128(smem_u32(sH1 + H_TILE_BYTES))};
        const uint64_t w2desc = smem_desc_sw128(smem_u32(sW2));
        const uint32_t w2_sub = (p.S_pad * 128) >> 4;
        uint64_t stdesc[NSTAGE], stw0desc[NSTAGE];
        for (int s = 0; s < NSTAGE; ++s) {
          stdesc[s] = smem_desc_sw128(smem_u32(sStage + s * p.stage_bytes));
          stw0desc[s] = smem_desc_noswz(smem_u32(sStage + s * p.stage_bytes + W1_TILE_BYTES), 64 * 16, 128);
        }
        const uint64_t w0resdesc[2] = {
            smem_desc_noswz(smem_u32(sW0res), 64 * 16, 128),
            smem_desc_noswz(smem_u32(sW0res + p.w0tile_bytes), 64 * 16, 128)};
        const uint32_t acc1 = tmem + TM_ACC1, acc2 = tmem + TM_ACC2;

        uint32_t gs = 0, gc = 0, hs = 0, zn = 0, a1f = 0, w2n = 0;
        WAITB(B_W0RES, 0, (uint32_t)st_dbg);
        for (int st = 0; st < total_steps; ++st) {
          tr_on = p.trace && blockIdx.x == p.trace_cta && st >= p.trace_t0 && st < p.trace_t1;
          TRACE(1, 0x1000);
          WAITB(B_ZREADY, zn & 1, (uint32_t)st_dbg); ++zn;
          TRACE(1, 0x1001);
          tc_fence_after();
          // prologue: L0 of chunks 0 and 1 from the resident W0 tiles
          for (int c = 0; c < 2 && c < NCH; ++c) {
            const uint32_t b = (gc + c) & 1;
            for (int j = 0; j < k0steps; ++j)
              umma_ss(tmem + TM_ACC0 + b * 64, zdesc + j * z_kstep, w0resdesc[c] + j * w0_kstep,
                      idesc0, j > 0);
            umma_commit(&bars[B_ACC0FULL + b]);
          }
          for (int nc = 0; nc < p.NC; ++nc) {
            for (int kc = 0; kc < p.KC; ++kc) {
              const int g = nc * p.KC + kc;
              st_dbg = (st << 8) | g;
              const uint32_t b = gc & 1, s = gs % NSTAGE;
              TRACE(1, 0x100 | g);
              WAITB(B_FULL + s, (gs / NSTAGE) & 1, (uint32_t)st_dbg);
              TRACE(1, 0x200 | g);
              WAITB(B_H0FULL + b, (gc >> 1) & 1, (uint32_t)st_dbg);
              TRACE(1, 0x300 | g);
              tc_fence_after();
              if (g + 2 < NCH) {   // L0 of chunk g+2 (its W0 tile rides in this stage)
                for (int j = 0; j < k0steps; ++j)
                  umma_ss(tmem + TM_ACC0 + b * 64, zdesc + j * z_kstep, stw0desc[s] + j * w0_kstep,
                          idesc0, j > 0);
                umma_commit(&bars[B_ACC0FULL + b]);
              }
              if (kc == 0) {
                // acc1 is drained once per pass: pass n (global count) waits for drain n-1
                if (a1f > 0) {
                  WAITB(B_ACC1FREE, (a1f - 1) & 1, (uint32_t)st_dbg);
                  tc_fence_after();
                }
                ++a1f;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_ss(acc1, h0desc[b] + 2 * j, stdesc[s] + 2 * j, idesc1, (kc | j) != 0);
              umma_commit(&bars[B_EMPTY + s]);
              umma_commit(&bars[B_H0FREE + b]);
              if (kc == p.KC - 1) umma_commit(&bars[B_ACC1FULL]);
              TRACE(1, 0x400 | g);
              ++gs; ++gc;
            }
            // L2 of pass nc: acc2 += relu(h1 chunk) * W2 chunk
            WAITB(B_W2FULL, w2n & 1, (uint32_t)st_dbg); ++w2n;
            for (int sub = 0; sub < 4; ++sub) {
              const uint32_t bb = hs & 1;
              WAITB(B_H1FULL + bb, (hs >> 1) & 1, (uint32_t)st_dbg);
              tc_fence_after();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_ss(acc2, h1desc[bb] + 2 * j, w2desc + sub * w2_sub + 2 * j, idesc2,
                        (nc | sub | j) != 0);
              umma_commit(&bars[B_H1FREE + bb]);
              ++hs;
            }
            umma_commit(&bars[B_W2EMPTY]);
          }
          umma_commit(&bars[B_ACC2FULL]);
          TRACE(1, 0x1002);
        }
      }
    } else {
      // =========================== epilogue / compute warps ===========================
      const int e = tid - 64;                              // 0..127
      const int r = (warp & 3) * 32 + lane;                // TMEM lane == row inside the tile
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const int S = p.S, A = p.A, K = p.K;
      const float* sB0 = sBias;
      const float* sB1 = sBias + p.H;
      const float* sB2 = sBias + 2 * p.H;
      const float* inMean = sNorm;
      const float* inStd = sNorm + p.SA;
      const float* dMean = sNorm + 2 * p.SA;
      const float* dStd = sNorm + 2 * p.SA + S;
      float* scrA = reinterpret_cast<float*>(sH0);         // [<=64][128] fp32 scratch (32 KB)
      float* scrB = reinterpret_cast<float*>(sH1);         // 2 x [32][128] fp32 scratch
      float* scrC = scrB + 32 * TILE_M;

      uint32_t gc = 0, hs = 0, a1n = 0, a2n = 0, xn_cnt = 0;
      float x[SMAX], a_raw[AMAX], a_mean[AMAX];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) x[s] = 0.f;
      int ts = 0, nreset = 0;

      for (int si = 0; si < MAX_SEG; ++si) {
        const int4 sg = segs[si];
        if (sg.x < 0) continue;
        const int tile = sg.x, t0 = sg.y, t1 = sg.z;
        const int row = tile * TILE_M + r;
        const bool valid = row < p.B;

        // ---- segment start: acquire the tile's state ----
        if (sg.w) {
          if (e == 0 && !wait_ge(&p.tile_flag[tile], 1u, p.dbg, 100u, (uint32_t)tile)) abort_smem = 1;
          named_bar_sync(1, EPI_THREADS);
          if (*reinterpret_cast<volatile int*>(&abort_smem)) goto bail;
        }
        if (t0 == 0 && !p.resume) {
#pragma unroll
          for (int s = 0; s < SMAX; ++s) x[s] = (valid && s < S) ? p.init_states[row * S + s] : 0.f;
          ts = 0; nreset = 0;
        } else {
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            x[s] = (valid && s < S) ? __ldcg(&p.row_state[static_cast<size_t>(row) * S + s]) : 0.f;
          ts = __ldcg(&p.row_ts[tile * TILE_M + r]);
          nreset = __ldcg(&p.row_nreset[tile * TILE_M + r]);
        }

        for (int t = t0; t < t1; ++t) {
          tr_on = p.trace && blockIdx.x == p.trace_cta && e == 0 && (t - t0) >= p.trace_t0 && (t - t0) < p.trace_t1;
          TRACE(2, 0x1000);
          // ================= begin step: action + Z operand =================
          if (p.ext_actions != nullptr) {
#pragma unroll
            for (int i = 0; i < AMAX; ++i) {
              a_raw[i] = (valid && i < A) ? p.ext_actions[row * A + i] : 0.f;
              a_mean[i] = a_raw[i];
            }
          } else {
            // policy mean network (training.py:99-103), fp32 on CUDA cores
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) scrA[s * TILE_M + r] = x[s];
            const float* in_s = scrA;
            float* out_s = scrB;
            for (int l = 0; l < p.n_pol_layers; ++l) {
              const PolicyLayer L = p.pl[l];
              const bool last = (l == p.n_pol_layers - 1);
              if (!last) {
                float acc[HPMAX];
                dense_layer<HPMAX>(in_s, L.nin, sPol + L.w_off, sPol + L.b_off, acc, r);
#pragma unroll
                for (int j = 0; j < HPMAX; ++j) out_s[j * TILE_M + r] = tanhf(acc[j]);
                in_s = out_s;
                out_s = (out_s == scrB) ? scrC : scrB;
              } else {
                float acc[AMAX];
                dense_layer<AMAX>(in_s, L.nin, sPol + L.w_off, sPol + L.b_off, acc, r);
#pragma unroll
                for (int i = 0; i < AMAX; ++i) a_mean[i] = p.pol_out_tanh ? tanhf(acc[i]) : acc[i];
              }
            }
            // a = eps * exp(log_std) + mean   (rllab get_actions)
            if (p.determ) {
#pragma unroll
              for (int i = 0; i < AMAX; ++i) a_raw[i] = a_mean[i];
            } else {
              float ep[AMAX];
              if (p.eps != nullptr) {
#pragma unroll
                for (int i = 0; i < AMAX; ++i)
                  ep[i] = (valid && i < A) ? p.eps[(static_cast<size_t>(t) * p.B + row) * A + i] : 0.f;
              } else {
#pragma unroll
                for (int blk = 0; blk < AMAX / 4; ++blk) {
                  float n4[4];
                  if (blk * 4 < A)
                    philox_normal4(p.seed, static_cast<uint32_t>(p.offset + t),
                                   static_cast<uint32_t>(row), PHILOX_STREAM_EPS + blk, n4);
                  else
                    n4[0] = n4[1] = n4[2] = n4[3] = 0.f;
#pragma unroll
                  for (int q = 0; q < 4; ++q) ep[blk * 4 + q] = n4[q];
                }
              }
#pragma unroll
              for (int i = 0; i < AMAX; ++i) {
                const float ls = fmaxf(sPol[p.pol_logstd_off + i], -13.815510557964274f);
                a_raw[i] = (i < A) ? __fadd_rn(__fmul_rn(ep[i], expf(ls)), a_mean[i]) : 0.f;
              }
            }
          }
          // z = (concat(x, clip(a)) - in_mean) / in_std, drop leading cols  (training.py:228,146-154)
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s < S) scrA[s * TILE_M + r] = __fdiv_rn(__fsub_rn(x[s], inMean[s]), inStd[s]);
#pragma unroll
          for (int i = 0; i < AMAX; ++i)
            if (i < A) {
              const float u = fminf(fmaxf(a_raw[i], -1.f), 1.f);   // env_helpers.py:599
              scrA[(S + i) * TILE_M + r] = __fdiv_rn(__fsub_rn(u, inMean[S + i]), inStd[S + i]);
            }
          for (int c = 0; c < p.K0 / 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int f0 = 8 * c + 2 * i, f1 = f0 + 1;
              const float z0 = f0 < p.Din ? scrA[(f0 + p.drop) * TILE_M + r] : 0.f;
              const float z1 = f1 < p.Din ? scrA[(f1 + p.drop) * TILE_M + r] : 0.f;
              pk[i] = pack_bf16x2(z0, z1);
            }
            *reinterpret_cast<uint4*>(sZ + c * (TILE_M * 16) + r * 16) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&bars[B_ZREADY]);
          TRACE(2, 0x1001);

          // ================= per-chunk epilogues =================
          for (int g = 0; g < NCH; ++g) {
            st_dbg = (t << 8) | g;
            {
              const uint32_t b = gc & 1;
              uint32_t v0[32], v1[32];
              TRACE(2, 0x100 | g);
              WAITB(B_ACC0FULL + b, (gc >> 1) & 1, (uint32_t)st_dbg);
              TRACE(2, 0x200 | g);
              tc_fence_after();
              tmem_ld32(tmem + lane_base + TM_ACC0 + b * 64, v0);
              tmem_ld32(tmem + lane_base + TM_ACC0 + b * 64 + 32, v1);
              tmem_ld_wait();
              TRACE(2, 0x300 | g);
              WAITB(B_H0FREE + b, ((gc >> 1) & 1) ^ 1, (uint32_t)st_dbg);
              TRACE(2, 0x400 | g);
              relu_pack_store(v0, v1, sB0 + (g % p.KC) * 64, sH0 + b * H_TILE_BYTES, r);
              TRACE(2, 0x500 | g);
              fence_proxy_async_smem();
              tc_fence_before();
              mbar_arrive(&bars[B_H0FULL + b]);
              TRACE(2, 0x600 | g);
              ++gc;
            }
            // layer-1 pass epilogue.  Pass nc is drained after the first two chunks of pass nc+1
            // have been staged (keeps the MMA warp fed across the pass boundary); the last pass
            // right after its last chunk.  KC >= 4, so the two triggers never coincide.
            int drain_nc = -1;
            if (g == NCH - 1) drain_nc = p.NC - 1;
            else if (g >= p.KC && (g % p.KC) == 1) drain_nc = g / p.KC - 1;
            if (drain_nc >= 0) {
              TRACE(2, 0x700 | g);
              WAITB(B_ACC1FULL, a1n & 1, (uint32_t)st_dbg); ++a1n;
              TRACE(2, 0x800 | g);
              tc_fence_after();
              for (int sub = 0; sub < 4; ++sub) {
                uint32_t v0[32], v1[32];
                tmem_ld32(tmem + lane_base + TM_ACC1 + sub * 64, v0);
                tmem_ld32(tmem + lane_base + TM_ACC1 + sub * 64 + 32, v1);
                tmem_ld_wait();
                if (sub == 3) { tc_fence_before(); mbar_arrive(&bars[B_ACC1FREE]); }
                const uint32_t bb = hs & 1;
                WAITB(B_H1FREE + bb, ((hs >> 1) & 1) ^ 1, (uint32_t)st_dbg);
                relu_pack_store(v0, v1, sB1 + drain_nc * 256 + sub * 64, sH1 + bb * H_TILE_BYTES, r);
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_H1FULL + bb]);
                ++hs;
              }
              TRACE(2, 0x900 | g);
            }
          }

          // ================= finish step: candidate, exchange, select, reward, reset =========
          float cand[SMAX];
          {
            uint32_t v[32];
            TRACE(2, 0x1002);
            WAITB(B_ACC2FULL, a2n & 1, (uint32_t)st_dbg); ++a2n;
            TRACE(2, 0x1003);
            tc_fence_after();
            tmem_ld32(tmem + lane_base + TM_ACC2, v);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
              const float o = __fadd_rn(__uint_as_float(v[s]), sB2[s]);
              // tf.add(diff_mean + diff_std * nn_output, x)   (training.py:257)
              cand[s] = (s < S) ? __fadd_rn(__fadd_rn(dMean[s], __fmul_rn(dStd[s], o)), x[s]) : 0.f;
            }
          }
          float xnext[SMAX];
          const int mode = p.sam_mode;
          if (K == 1) {
#pragma unroll
            for (int s = 0; s < SMAX; ++s) xnext[s] = cand[s];
          } else {
            float* xb = p.xbuf + static_cast<size_t>((slot * 2 + (xn_cnt & 1)) * K) * S * TILE_M;
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) xb[(k * S + s) * TILE_M + r] = cand[s];
            __threadfence();
            named_bar_sync(1, EPI_THREADS);
            if (e == 0) {
              red_release_gpu_add(&p.xctr[slot], 1u);
              if (!wait_ge(&p.xctr[slot], static_cast<unsigned>(K) * (xn_cnt + 1), p.dbg, 101u,
                           (uint32_t)st_dbg))
                abort_smem = 1;
            }
            named_bar_sync(1, EPI_THREADS);
            if (*reinterpret_cast<volatile int*>(&abort_smem)) goto bail;
            TRACE(2, 0x1004);
            ++xn_cnt;
            if (mode == METRPO_SAM_STEP_RAND || mode == METRPO_SAM_EPS_RAND ||
                mode == METRPO_SAM_ONE_MODEL) {
              int idx = 0;
              if (mode != METRPO_SAM_ONE_MODEL) {
                if (p.model_idx != nullptr)
                  idx = valid ? p.model_idx[static_cast<size_t>(t) * p.B + row] : 0;
                else if (mode == METRPO_SAM_STEP_RAND)
                  idx = philox_index(p.seed, static_cast<uint32_t>(p.offset + t),
                                     static_cast<uint32_t>(row), PHILOX_STREAM_IDX, K);
                else
                  idx = philox_index(p.seed, static_cast<uint32_t>(nreset),
                                     static_cast<uint32_t>(row), PHILOX_STREAM_EIDX, K);
                idx = min(max(idx, 0), K - 1);
              }
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                xnext[s] = (s < S) ? __ldcg(&xb[(idx * S + s) * TILE_M + r]) : 0.f;
            } else {
              // model_mean / model_med / model_mean_std over the K candidates (:624-630)
              for (int s = 0; s < S; ++s) {
                float m = 0.f;
                for (int kk = 0; kk < K; ++kk) m = __fadd_rn(m, __ldcg(&xb[(kk * S + s) * TILE_M + r]));
                m = __fdiv_rn(m, static_cast<float>(K));
                float outv = m;
                if (mode == METRPO_SAM_MODEL_MEAN_STD) {
                  float var = 0.f;
                  for (int kk = 0; kk < K; ++kk) {
                    const float d = __fsub_rn(__ldcg(&xb[(kk * S + s) * TILE_M + r]), m);
                    var = __fadd_rn(var, __fmul_rn(d, d));
                  }
                  const float sd = sqrtf(__fdiv_rn(var, static_cast<float>(K)));
                  float nz;
                  if (p.std_noise != nullptr) {
                    nz = valid ? p.std_noise[(static_cast<size_t>(t) * p.B + row) * S + s] : 0.f;
                  } else {
                    float n4[4];
                    philox_normal4(p.seed, static_cast<uint32_t>(p.offset + t),
                                   static_cast<uint32_t>(row), PHILOX_STREAM_STD + (s >> 2), n4);
                    nz = n4[s & 3];
                  }
                  outv = __fadd_rn(m, __fmul_rn(nz, sd));
                } else if (mode == METRPO_SAM_MODEL_MED) {
                  // rank selection without a local array: count values below / equal
                  float lo = 0.f, hi = 0.f;
                  const int r_lo = (K - 1) / 2, r_hi = K / 2;
                  for (int a = 0; a < K; ++a) {
                    const float va = __ldcg(&xb[(a * S + s) * TILE_M + r]);
                    int less = 0, eq = 0;
                    for (int b2 = 0; b2 < K; ++b2) {
                      const float vb = __ldcg(&xb[(b2 * S + s) * TILE_M + r]);
                      less += (vb < va);
                      eq += (vb == va);
                    }
                    if (less <= r_lo && r_lo < less + eq) lo = va;
                    if (less <= r_hi && r_hi < less + eq) hi = va;
                  }
                  outv = (r_lo == r_hi) ? lo : __fmul_rn(__fadd_rn(lo, hi), 0.5f);
                }
                scrA[s * TILE_M + r] = outv;
              }
#pragma unroll
              for (int s = 0; s < SMAX; ++s) xnext[s] = (s < S) ? scrA[s * TILE_M + r] : 0.f;
            }
          }
          // reward = -cost_np_vec(s, clip(a), s')   (env_helpers.py:601)
          float u[AMAX];
#pragma unroll
          for (int i = 0; i < AMAX; ++i) u[i] = fminf(fmaxf(a_raw[i], -1.f), 1.f);
          const float reward = -env_cost(p.env_id, S, A, xnext, u);
          ts += 1;
          const bool dn = env_is_done(p.env_id, S, xnext) || (ts >= p.T_max);   // :603-604
          if (k == 0 && valid) {
            const size_t o = static_cast<size_t>(t) * p.B + row;
            if (p.obs) {
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) p.obs[o * S + s] = x[s];
            }
#pragma unroll
            for (int i = 0; i < AMAX; ++i)
              if (i < A) {
                if (p.act) p.act[o * A + i] = a_raw[i];
                if (p.mean) p.mean[o * A + i] = a_mean[i];
              }
            if (p.rew) p.rew[o] = reward;
            if (p.done) p.done[o] = dn ? 1 : 0;
          }
          TRACE(2, 0x1005);
#pragma unroll
          for (int s = 0; s < SMAX; ++s) x[s] = xnext[s];
          if (dn) {   // env_helpers.py:605-606 -> reset(dones)
            if (valid) {
              const float* src =
                  p.ext_reset_states
                      ? p.ext_reset_states + static_cast<size_t>(row) * S
                      : p.reset_pool + static_cast<size_t>((static_cast<long long>(nreset) * p.B + row) % p.R) * S;
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) x[s] = src[s];
            }
            nreset += 1;
            ts = 0;
          }
        }  // t

        // ---- segment end: publish the tile's state ----
        if (k == 0) {
          if (valid) {
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) {
                p.row_state[static_cast<size_t>(row) * S + s] = x[s];
                if (t1 == p.n_steps && p.final_states) p.final_states[static_cast<size_t>(row) * S + s] = x[s];
              }
          }
          p.row_ts[tile * TILE_M + r] = ts;
          p.row_nreset[tile * TILE_M + r] = nreset;
          __threadfence();
          named_bar_sync(1, EPI_THREADS);
          if (e == 0) red_release_gpu_add(&p.tile_flag[tile], 1u);
        }
      }  // segments
    }
  }

bail:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace metrpo
