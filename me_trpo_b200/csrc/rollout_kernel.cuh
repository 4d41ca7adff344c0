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
//             cp.async.bulk (TMA engine) through a 4-stage mbarrier ring
//   warp 1    MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma (M=128, fp32
//             accumulators in TMEM); lanes 0/1 probe the next chunk's mbarriers while the
//             current chunk's MMAs are queued
//   warps 2-5 epilogue/compute: TMEM -> registers -> (bias)+ReLU -> bf16 -> back to TMEM as the
//             next layer's A operand, plus the per-step serial section (residual/de-normalise,
//             exchange, select, reward/done/reset, trajectory write, policy MLP, normalise -> Z)
//
// Per step and model the MLP  z[128,K0] -> H -> H -> S  is evaluated as
//   for nc in H/256 output chunks of layer 1:          (acc1: 256 TMEM columns)
//     for kc in H/64 reduction chunks:
//        L0 (every 2nd chunk): acc0 = Z(tmem) * W0[:, 128 columns]   (N=128, K=K0; recomputed per
//                                                         nc pass because 128 x H activations do
//                                                         not fit on an SM; b0 rides in two
//                                                         constant-1 columns of Z.  Every MMA at
//                                                         M=128 costs >= 128 cycles whatever N,
//                                                         hence the widest N that fits TMEM)
//        epilogue: H0[b] = bf16(relu(acc0[64 cols]))    (tcgen05.st back to TMEM: the A operand
//                                                         of L1 never touches shared memory)
//        L1: acc1 += H0[b](tmem) * W1[kc chunk, nc chunk]   (N=256, K=64: 4 MMAs, B from smem)
//     epilogue: 4 x 64 columns of acc1 -> bf16(relu(. + b1)) written IN PLACE into acc1's own
//               columns; L2: acc2 += H1(tmem) * W2 chunk   (N=S_pad, K=256: 16 MMAs)
//   next_state = (diff_mean + diff_std * (acc2 + b2)) + x          (training.py:257)
#pragma once
#include "umma.cuh"
#include "philox.cuh"

namespace metrpo {

constexpr int TILE_M = 128;
constexpr int NSTAGE = 4;     // must stay 4: the EMPTY barriers double as H0-buffer-free signals (stage & 1 == chunk & 1)
// The kernel is compiled for two register budgets <SMAX, AMAX> = <32, 8> (every shipped env but
// humanoid) and <64, 24> (humanoid: S = 55, A = 21, policy 100-50-25); rollout_api.cu picks one.
constexpr int HPB = 32;       // policy hidden layers are evaluated in blocks of 32 output columns
constexpr int HPMAX = 128;    // max policy hidden width
constexpr int MAX_SEG = 128;  // segments per gang slot
// W1 stage: [N1 n][64 k] bf16, SW128, N1 = 256 (32 KB) or 128 (16 KB; configs whose operands do not
// fit TMEM / shared memory next to a 256-column layer-1 accumulator)
constexpr int BIAS_PAD = 64;    // zero-padded b2 slot of the per-model bias blob [b0 | b1 | b2]
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;

// TMEM column map (512 allocated), offsets chosen by the host (KParams::tm_*):
//   acc1  N1 cols fp32 at 0; after the drain: N1/64 x 32 cols of packed bf16 H1
//   acc0  128 cols: layer-0 accumulator of one group (2 chunks)
//   acc2  S_pad cols
//   H0    2 x 32 cols: relu(h0) chunk as packed bf16 pairs (A of L1)
//   Z     K0/2 cols: normalised input as packed bf16 (A of L0)
constexpr uint32_t TM_ACC1 = 0;

enum {
  B_FULL = 0,        // [NSTAGE] W1 stage landed (tx)
  B_EMPTY = 4,       // [NSTAGE] L1 of the chunk in this stage completed (commit): frees the W1 stage
                     //          for the producer AND the H0 buffer (stage & 1) for the epilogue
  B_W2FULL = 8,
  B_W2EMPTY = 9,
  B_W0FULL = 10,     // [2] W0 group tile landed (tx)
  B_ACC0FULL = 12,   // [2] L0 group accumulated (commit); also frees W0 ring slot (group & 1)
  B_ZREADY = 14,     // Z operand written (128 arrivals)
  B_ACC0FREE = 15,   // acc0 loaded to registers (128 arrivals)
  B_H0FULL = 16,     // [2] H0 operand written to TMEM (128 arrivals)
  B_ACC1FULL = 18,   // layer-1 pass accumulated (commit)
  B_H1FULL = 19,     // [4] 64-column slice of acc1 converted in place (128 arrivals)
  B_ACC2FULL = 23,
  NUM_BARS = 24
};
// tcgen05.commit costs ~65 cycles in the (blocking) MMA issue stream, hence one barrier per event
// with several waiters rather than one barrier per waiter.

struct PolicyLayer {
  int nin, nout, npad;   // npad = nout rounded up to 32 for hidden layers, AMAX for the output layer
  int w_off, b_off;      // float offsets into the policy blob
  int in_off, out_off;   // float offsets of the activation rows [n][128] in the scratch area (a layer
                         // of <= 32 outputs runs in place, wider ones ping-pong between two regions)
};

struct KParams {
  // dims
  int S, A, SA, drop, Din, K0, H, S_pad, K, B, T_max, env_id, sam_mode, determ;
  int NC, KC, N1;        // layer-1 passes, 64-wide reduction chunks, pass width (256 or 128)
  uint32_t tm_acc0, tm_acc2, tm_h0, tm_z;   // TMEM column offsets
  int pol_in_smem;       // policy blob staged in shared memory (else read through L1 from global)
  int own_mode;          // 1: row-ownership exchange (step_rand / eps_rand, K > 1): the CTA whose model a
                         //    row selects this step computes reward / done / reset / next action for it
  int rec_stride;        // own_mode: floats per row record in xbuf  [a_raw AMAX | done, pad 3 | x_new S, pad to 4]
  int slot_stride;       // own_mode: floats per compacted policy input row in the scratch area
  unsigned long long xbuf_stride;   // floats per (slot, parity) exchange buffer
  int n_steps, n_slots, n_tiles;
  int resume;            // 1: state comes from row_state (B1 step / continued run)
  int per_model;         // 1: per-model validation-cost rollout (model_based_rl.py:122-142): every
                         //    model rolls its OWN prediction forward with the deterministic policy,
                         //    no exchange / reset / timeout; only the discounted cost is emitted
  float gamma;           // per_model: discount (policy_opt_params.gamma)
  float* pm_cost;        // per_model: [K][B] discounted cost of (model, row)
  int row_offset;        // global index of local row 0 (keys the Philox streams)
  // packed weights (per model: W1 stages | W2 chunks | W0 group tiles)
  const uint8_t* wstream;
  unsigned long long model_stride;
  uint32_t stage_bytes, w0g_bytes, w2chunk_bytes, off_w2, off_w0g;   // w0g: [128 n][K0] tile
  const float* bias;     // [K][2H + BIAS_PAD]: b0 (unused by the kernel) | b1 | b2 (zero padded)
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
  unsigned long long* trace;   // optional event trace of one CTA (METRPO_TRACE builds), or nullptr
  int trace_cta, trace_t0, trace_t1;
  unsigned* dbg;          // [DBG_HEADER + grid*6*4] abort flag + wait records
  const int4* segs;       // [n_slots][MAX_SEG] = (tile, t0, t1, wait_flag); tile < 0 -> unused
  // smem carve-up (byte offsets from the 1024-aligned base)
  uint32_t off_stage, off_sw0g, off_scr, off_sw2, off_sbias, off_snorm, off_spol, off_bars;
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
  const int idx = DBG_HEADER + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * DBG_WORDS_PER_WARP;
  dbg[idx + 0] = tag;
  dbg[idx + 1] = info;
  dbg[idx + 2] = why;    // 1 = timed out here, 2 = saw the abort flag here
  dbg[idx + 3] = threadIdx.x;
}
// sleep_ns > 0: back off between polls (roles whose wake-up latency is not critical must not steal
// issue slots from the epilogue warps that share their SM sub-partition)
__device__ __noinline__ bool wait_bar_slow(uint64_t* bar, uint32_t parity, unsigned* dbg, uint32_t tag,
                                           uint32_t info, uint32_t sleep_ns = 0) {
  // the timer is first read after 256 failed polls: nearly every wait ends long before that, and
  // a %globaltimer read at entry sits on the wake-up path of every wait that misses its first poll
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if ((++spins & 0xff) == 0) {
      if (*reinterpret_cast<volatile unsigned*>(dbg) != 0u) { dbg_record(dbg, tag, info, 2); return false; }
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > METRPO_WAIT_TIMEOUT_NS) {
        atomicExch(dbg, 1u);
        dbg_record(dbg, tag, info, 1);
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, unsigned* dbg, uint32_t tag,
                                         uint32_t info, uint32_t sleep_ns = 0) {
  if (mbar_try_wait(bar, parity)) return true;
  return wait_bar_slow(bar, parity, dbg, tag, info, sleep_ns);
}
__device__ __noinline__ bool wait_ge(const unsigned* ptr, unsigned target, unsigned* dbg, uint32_t tag,
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
#define WAITB(idx, par) \
  do { if (!wait_bar(&bars[idx], (par), p.dbg, (idx), (uint32_t)st_dbg)) goto bail; } while (0)
#ifndef METRPO_PRODUCER_SLEEP_NS
#define METRPO_PRODUCER_SLEEP_NS 0
#endif
#ifndef METRPO_MMA_PROLOGUE_SLEEP_NS
#define METRPO_MMA_PROLOGUE_SLEEP_NS 0
#endif
#define WAITB_P(idx, par) \
  do { if (!wait_bar(&bars[idx], (par), p.dbg, (idx), (uint32_t)st_dbg, METRPO_PRODUCER_SLEEP_NS)) goto bail; } while (0)

// optional event trace (dev tool, -DMETRPO_TRACE): role-private logs of (code << 40 | clock) for
// one CTA and a window of steps.  role 0 producer, 1 MMA, 2 epilogue.
constexpr int TRACE_CAP = 4096;
#ifdef METRPO_TRACE
#define TRACE(role, code)                                                                    \
  do {                                                                                       \
    if (tr_on && tr_n < TRACE_CAP)                                                           \
      p.trace[(role) * TRACE_CAP + tr_n++] =                                                 \
          (static_cast<unsigned long long>(code) << 40) | (clock64() & 0xFFFFFFFFFFull);     \
  } while (0)
#define TRACE_ON(cond) tr_on = p.trace && blockIdx.x == p.trace_cta && (cond)
#else
#define TRACE(role, code) do { } while (0)
#define TRACE_ON(cond) do { } while (0)
#endif

// per-row analytic cost (reward = -cost); u is the clipped action.  envs/com_*_env.py
template <int SMAX, int AMAX>
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
template <int SMAX>
__device__ __forceinline__ bool env_is_done(int env_id, int S, const float (&xn)[SMAX]) {
  if (env_id != METRPO_ENV_ANT) return false;   // NeuralNetEnv default (env_helpers.py:537)
  bool finite = true;
#pragma unroll
  for (int s = 0; s < SMAX; ++s)
    if (s < S) finite = finite && isfinite(xn[s]);
  return !(xn[2] >= 0.2f && xn[2] <= 1.0f && finite);
}

// one dense layer of the policy on CUDA cores: thread r owns column r of the [n][128] scratch
// (thread-private, so layers may run in place).  Packed FFMA2 (two fp32 FMAs per instruction).
template <int NP>
__device__ __forceinline__ void dense_fma_row(float2 (&acc2)[NP / 2], float xi, const float4 (&w)[NP / 4]) {
  const float2 x2 = make_float2(xi, xi);
#pragma unroll
  for (int j = 0; j < NP / 4; ++j) {
    acc2[2 * j] = __ffma2_rn(x2, make_float2(w[j].x, w[j].y), acc2[2 * j]);
    acc2[2 * j + 1] = __ffma2_rn(x2, make_float2(w[j].z, w[j].w), acc2[2 * j + 1]);
  }
}
template <int NP>
__device__ __forceinline__ void dense_load_row(const float* in_s, const float* W, int ld, int i, int r, float& xi,
                                               float4 (&w)[NP / 4]) {
  xi = in_s[i * TILE_M + r];
  const float4* w4 = reinterpret_cast<const float4*>(W + i * ld);
#pragma unroll
  for (int j = 0; j < NP / 4; ++j) w[j] = w4[j];
}
// NP output columns starting at W (row stride ld floats), b
template <int NP>
__device__ __forceinline__ void dense_layer(const float* in_s, int nin, const float* W, int ld,
                                            const float* b, float (&acc)[NP], int r) {
  float2 acc2[NP / 2];
#pragma unroll
  for (int j = 0; j < NP / 2; ++j) acc2[j] = make_float2(b[2 * j], b[2 * j + 1]);
  float4 wa[NP / 4], wb[NP / 4];
  float xa, xb;
  dense_load_row<NP>(in_s, W, ld, 0, r, xa, wa);
  int i = 0;
  for (; i + 2 < nin; i += 2) {          // weight rows i+1 / i+2 are in flight while row i / i+1 is used
    dense_load_row<NP>(in_s, W, ld, i + 1, r, xb, wb);
    dense_fma_row<NP>(acc2, xa, wa);
    dense_load_row<NP>(in_s, W, ld, i + 2, r, xa, wa);
    dense_fma_row<NP>(acc2, xb, wb);
  }
  if (i + 1 < nin) {
    dense_load_row<NP>(in_s, W, ld, i + 1, r, xb, wb);
    dense_fma_row<NP>(acc2, xa, wa);
    dense_fma_row<NP>(acc2, xb, wb);
  } else {
    dense_fma_row<NP>(acc2, xa, wa);
  }
#pragma unroll
  for (int j = 0; j < NP / 2; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
}

// (bias +) ReLU + bf16 pack of 64 accumulator columns -> 32 packed columns (A operand layout);
// relu and the bf16x2 pack are one F2FP.RELU instruction per column pair
template <bool kBias>
__device__ __forceinline__ void relu_pack(const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                          const float* bias, uint32_t (&pk)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {   // 8 columns per step
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = 8 * c + i;
      a[i] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
    }
    if (kBias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * c);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + 8 * c + 4);
      // packed FADD2 (the scalar three-operand form issues at half rate on sm_100)
      const float2 r0 = __fadd2_rn(make_float2(a[0], a[1]), make_float2(b0.x, b0.y));
      const float2 r1 = __fadd2_rn(make_float2(a[2], a[3]), make_float2(b0.z, b0.w));
      const float2 r2 = __fadd2_rn(make_float2(a[4], a[5]), make_float2(b1.x, b1.y));
      const float2 r3 = __fadd2_rn(make_float2(a[6], a[7]), make_float2(b1.z, b1.w));
      a[0] = r0.x; a[1] = r0.y; a[2] = r1.x; a[3] = r1.y; a[4] = r2.x; a[5] = r2.y; a[6] = r3.x; a[7] = r3.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[4 * c + i] = relu_pack_bf16x2(a[2 * i], a[2 * i + 1]);
  }
}

// tanh(x) = 1 - 2/(exp(2x)+1) with ex2.approx / rcp.approx: abs error ~1e-7 (policy hidden layers)
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

// =============================================================================================
template <int SMAX, int AMAX, int N1>
__global__ void __launch_bounds__(NUM_THREADS, 1) rollout_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ int abort_smem;
  // row-ownership exchange: compacted row list, per-row policy noise of the next step, hidden
  // activations of one policy pass.  The narrow instantiation keeps the last two in static shared
  // memory (its scratch area is only 16 KB); the wide one carves them out of its policy scratch area.
  constexpr bool NARROW = (SMAX <= 32);
  __shared__ int sList[TILE_M];
  __shared__ int sCnt[4];
  __shared__ __align__(16) float sEpsStatic[NARROW ? TILE_M * AMAX : 4];
  __shared__ __align__(16) float sHidStatic[NARROW ? 2 * 64 * 33 : 4];   // narrow: 2 layers x 64 rows x (32 + 1)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.x / p.K, k = blockIdx.x % p.K;

  uint8_t* sStage = smem + p.off_stage;
  uint8_t* sW0g = smem + p.off_sw0g;     // 2-slot ring of W0 group tiles
  uint8_t* sW2 = smem + p.off_sw2;
  float* sScr = reinterpret_cast<float*>(smem + p.off_scr);   // [max(S+A,32)][128] fp32 scratch
  float* sBias = reinterpret_cast<float*>(smem + p.off_sbias);
  float* sNorm = reinterpret_cast<float*>(smem + p.off_snorm);
  // the narrow instantiation always stages the policy blob in shared memory (<= 16 KB); the wide one
  // reads it through L1 from global when it does not fit next to the activations
  float* sPolS = reinterpret_cast<float*>(smem + p.off_spol);
  const float* sPol = (SMAX <= 32 || p.pol_in_smem) ? sPolS : p.pol;
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
#ifdef METRPO_TRACE
  bool tr_on = false;
  int tr_n = 0;
#endif
  if (tid == 0) {
    abort_smem = 0;
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
    mbar_init(&bars[B_W2FULL], 1); mbar_init(&bars[B_W2EMPTY], 1);
    mbar_init(&bars[B_ZREADY], EPI_THREADS);
    mbar_init(&bars[B_ACC0FREE], EPI_THREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_W0FULL + i], 1); mbar_init(&bars[B_ACC0FULL + i], 1);
      mbar_init(&bars[B_H0FULL + i], EPI_THREADS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[B_H1FULL + i], EPI_THREADS);
    mbar_init(&bars[B_ACC1FULL], 1);
    mbar_init(&bars[B_ACC2FULL], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  {  // constants -> smem (generic loads; read-only for the rest of the kernel)
    const float* gb = p.bias + static_cast<size_t>(k) * (2 * p.H + BIAS_PAD);
    for (int i = tid; i < 2 * p.H + BIAS_PAD; i += NUM_THREADS) sBias[i] = gb[i];
    for (int i = tid; i < 2 * p.SA + 2 * p.S; i += NUM_THREADS) {
      const float v = p.norm[i];
      sNorm[i] = (i >= p.SA && i < 2 * p.SA) ? __frcp_rn(v) : v;   // in_std slot holds 1 / in_std
    }
    if (SMAX <= 32 || p.pol_in_smem)
      for (int i = tid; i < p.pol_floats; i += NUM_THREADS) sPolS[i] = p.pol[i];
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
        uint32_t s = 0, sphase = 0, w2n = 0, wg = 0;
        const int NG = p.KC / 2;                      // L0 groups (2 chunks) per pass
        const uint32_t total_groups = static_cast<uint32_t>(total_steps) * (NCH / 2);
        const int w2_at = p.KC > 3 ? 3 : p.KC - 1;
        // W0 group tiles are consumed in the cyclic order wg % NG over the whole kernel; they are
        // loaded one group ahead of the W1 stages of the group that precedes them.
#define LOAD_W0()                                                                             \
  do {                                                                                        \
    if (wg < total_groups) {                                                                  \
      const uint32_t ws = wg & 1;                                                             \
      if (wg >= 2) WAITB_P(B_ACC0FULL + ws, ((wg >> 1) - 1) & 1);                             \
      mbar_arrive_expect_tx(&bars[B_W0FULL + ws], p.w0g_bytes);                               \
      bulk_g2s_hint(sW0g + ws * p.w0g_bytes,                                                  \
                    wm + p.off_w0g + static_cast<size_t>(wg % NG) * p.w0g_bytes, p.w0g_bytes, \
                    &bars[B_W0FULL + ws], pol);                                               \
      ++wg;                                                                                   \
    }                                                                                         \
  } while (0)
        LOAD_W0();
        for (int st = 0; st < total_steps; ++st) {
          TRACE_ON(st >= p.trace_t0 && st < p.trace_t1);
          const uint8_t* src = wm;
          for (int nc = 0; nc < p.NC; ++nc) {
            for (int kc = 0; kc < p.KC; ++kc) {
              st_dbg = (st << 8) | (nc * p.KC + kc);
              if ((kc & 1) == 0) LOAD_W0();
              TRACE(0, 0x100 | (nc * p.KC + kc));
              WAITB_P(B_EMPTY + s, sphase ^ 1);
              TRACE(0, 0x200 | (nc * p.KC + kc));
              mbar_arrive_expect_tx(&bars[B_FULL + s], p.stage_bytes);
              bulk_g2s_hint(sStage + s * p.stage_bytes, src, p.stage_bytes, &bars[B_FULL + s], pol);
              src += p.stage_bytes;
              if (++s == NSTAGE) { s = 0; sphase ^= 1; }
              if (kc == w2_at) {
                WAITB_P(B_W2EMPTY, (w2n & 1) ^ 1);
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
      // All 32 lanes run the (warp-uniform) loop so that descriptors live in uniform registers;
      // lanes 0/1 poll the mbarriers, one elected lane issues tcgen05.mma / commit.  The tensor
      // pipe accepts only a couple of queued MMAs before issue blocks, so every cycle this warp
      // spends waiting is pipe idle time: the next chunk's barriers are probed while the current
      // chunk's MMAs are still queued.
      {
        const uint32_t idesc0 = idesc_bf16_f32(128, 128);
        const uint32_t idesc1 = idesc_bf16_f32(128, N1);
        const uint32_t idesc2 = idesc_bf16_f32(128, p.S_pad);
        const int k0steps = p.K0 / 16;
        const uint32_t ztm = tmem + p.tm_z;                    // A of L0: 8 TMEM columns per 16-k step
        const uint32_t w0_kstep = (2 * 128 * 16) >> 4;       // W0 group tile: [128 n][K0] no-swizzle
        const uint64_t w0desc0 = smem_desc_noswz(smem_u32(sW0g), 128 * 16, 128);
        const uint32_t w0slot = p.w0g_bytes >> 4;
        const uint64_t w2desc = smem_desc_sw128(smem_u32(sW2));
        const uint32_t w2_sub = (p.S_pad * 128) >> 4;
        const uint64_t stdesc0 = smem_desc_sw128(smem_u32(sStage));
        const uint32_t ststride = p.stage_bytes >> 4;
        const uint32_t acc0 = tmem + p.tm_acc0, acc1 = tmem + TM_ACC1, acc2 = tmem + p.tm_acc2;
        constexpr int nsl = N1 / 64;   // 64-column slices of a layer-1 pass

#define WAITW1(idx, par)                                                                      \
  do {                                                                                        \
    bool ok_ = true;                                                                          \
    if (lane == 0) ok_ = wait_bar(&bars[idx], (par), p.dbg, (idx), (uint32_t)st_dbg);         \
    if (!__all_sync(0xffffffffu, ok_)) goto bail;                                             \
  } while (0)
#define WAITW2(idxa, para, idxb, parb)                                                        \
  do {                                                                                        \
    bool ok_ = true;                                                                          \
    const int wi_ = lane == 0 ? (idxa) : (idxb);                                              \
    const uint32_t wp_ = lane == 0 ? (para) : (parb);                                         \
    if (lane < 2) ok_ = wait_bar(&bars[wi_], wp_, p.dbg, (uint32_t)wi_, (uint32_t)st_dbg);    \
    if (!__all_sync(0xffffffffu, ok_)) goto bail;                                             \
  } while (0)
// single non-blocking probe of up to four barriers: lanes 0..3 each test their own barrier with ONE
// (data-divergent, not control-divergent) try_wait; idx < 0 -> lane idle.  Result consumed later.
#define PROBE4(var, i0, p0, i1, p1, i2, p2, i3, p3)                                           \
  bool var = true;                                                                            \
  {                                                                                           \
    const int wi_ = lane == 0 ? (i0) : lane == 1 ? (i1) : lane == 2 ? (i2) : lane == 3 ? (i3) : -1; \
    const uint32_t wp_ = lane == 0 ? (p0) : lane == 1 ? (p1) : lane == 2 ? (p2) : (p3);       \
    if (wi_ >= 0) var = mbar_try_wait(&bars[wi_], wp_);                                       \
  }
// blocking wait on up to four barriers in parallel (lanes 0..3); idx < 0 -> lane idle
#define WAITW4(i0, p0, i1, p1, i2, p2, i3, p3)                                                \
  do {                                                                                        \
    bool ok_ = true;                                                                          \
    const int wi_ = lane == 0 ? (i0) : lane == 1 ? (i1) : lane == 2 ? (i2) : lane == 3 ? (i3) : -1; \
    const uint32_t wp_ = lane == 0 ? (p0) : lane == 1 ? (p1) : lane == 2 ? (p2) : (p3);       \
    if (wi_ >= 0) ok_ = wait_bar(&bars[wi_], wp_, p.dbg, (uint32_t)wi_, (uint32_t)st_dbg);    \
    if (!__all_sync(0xffffffffu, ok_)) goto bail;                                             \
  } while (0)

        uint32_t s = 0, sphase = 0;   // W1 stage ring position / phase
        uint32_t gg = 0;              // L0 groups issued so far (acc0 / W0 ring use count)
        uint32_t hg = 0;              // groups consumed by L1 so far (H0 buffer use count)
        uint32_t zn = 0, npass = 0, w2n = 0;
        const int NG = p.KC / 2;      // groups per pass
        const int NGS = NCH / 2;      // groups per step
        for (int st = 0; st < total_steps; ++st) {
          TRACE_ON(lane == 0 && st >= p.trace_t0 && st < p.trace_t1);
          TRACE(1, 0x1000);
          // step prologue: Z ready, W0 tile of the step's first group, acc0 drained (a ~20 k cycle
          // wait for the epilogue's serial section: poll with back-off)
          {
            bool ok_ = true;
            const int wi_ = lane == 0 ? B_ZREADY : lane == 1 ? B_W0FULL + (int)(gg & 1) : (lane == 2 && gg > 0) ? B_ACC0FREE : -1;
            const uint32_t wp_ = lane == 0 ? (zn & 1) : lane == 1 ? ((gg >> 1) & 1) : ((gg - 1) & 1);
            if (wi_ >= 0) ok_ = wait_bar(&bars[wi_], wp_, p.dbg, (uint32_t)wi_, (uint32_t)st_dbg, METRPO_MMA_PROLOGUE_SLEEP_NS);
            if (!__all_sync(0xffffffffu, ok_)) goto bail;
          }
          ++zn;
          TRACE(1, 0x1001);
          tc_fence_after();
          if (elect_one()) {
            for (int j = 0; j < k0steps; ++j)
              umma_ts(acc0, ztm + j * 8, w0desc0 + (gg & 1) * w0slot + j * w0_kstep, idesc0, j > 0);
            umma_commit(&bars[B_ACC0FULL + (gg & 1)]);
          }
          __syncwarp();
          ++gg;
          // operands of group 0 (chunk a) + what the L0 of group 1 needs
          WAITW4(B_FULL + (int)s, sphase, B_H0FULL + 0, hg & 1,
                 NGS > 1 ? B_ACC0FREE : -1, (gg - 1) & 1, NGS > 1 ? B_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
          int gp = 0, nc = 0;   // group position inside the pass / pass index
          for (int G = 0; G < NGS; ++G) {
            st_dbg = (st << 8) | (2 * G);
            const bool next_l0 = (G + 1 < NGS);
            const bool first_of_pass = (gp == 0), last_of_pass = (gp == NG - 1);
            uint32_t s1 = s + 1, sphase1 = sphase;
            if (s1 == NSTAGE) { s1 = 0; sphase1 ^= 1; }
            uint32_t s2 = s1 + 1, sphase2 = sphase1;
            if (s2 == NSTAGE) { s2 = 0; sphase2 ^= 1; }
            TRACE(1, 0x100 | (2 * G));
            // ---- batch 1: L0 of the next group (its epilogue needs a full group of lead time),
            //      then L1 of chunk a
            tc_fence_after();
            if (elect_one()) {
              if (next_l0) {
                for (int j = 0; j < k0steps; ++j)
                  umma_ts(acc0, ztm + j * 8, w0desc0 + (gg & 1) * w0slot + j * w0_kstep, idesc0, j > 0);
                umma_commit(&bars[B_ACC0FULL + (gg & 1)]);
              }
              const uint32_t at = tmem + p.tm_h0;
              const uint64_t bd = stdesc0 + s * ststride;
              umma_ts(acc1, at, bd, idesc1, first_of_pass ? 0u : 1u);
              umma_ts(acc1, at + 8, bd + 2, idesc1, 1);
              umma_ts(acc1, at + 16, bd + 4, idesc1, 1);
              umma_ts(acc1, at + 24, bd + 6, idesc1, 1);
              umma_commit(&bars[B_EMPTY + s]);
            }
            __syncwarp();
            if (next_l0) ++gg;
            // ---- chunk b operands (probe; the MMAs above are still executing)
            {
              PROBE4(rb, B_FULL + (int)s1, sphase1, B_H0FULL + 1, hg & 1, -1, 0, -1, 0);
              if (!__all_sync(0xffffffffu, rb)) WAITW4(B_FULL + (int)s1, sphase1, B_H0FULL + 1, hg & 1, -1, 0, -1, 0);
            }
            TRACE(1, 0x300 | (2 * G));
            tc_fence_after();
            if (elect_one()) {
              const uint32_t at = tmem + p.tm_h0 + 32;
              const uint64_t bd = stdesc0 + s1 * ststride;
              umma_ts(acc1, at, bd, idesc1, 1);
              umma_ts(acc1, at + 8, bd + 2, idesc1, 1);
            }
            __syncwarp();
            // ---- probe everything the next group needs while the MMAs above are queued
            const bool need_l0_2 = (G + 2 < NGS);
            const bool has_next = (G + 1 < NGS);
            PROBE4(rn, has_next ? B_FULL + (int)s2 : -1, sphase2, has_next ? B_H0FULL + 0 : -1, (hg + 1) & 1,
                   need_l0_2 ? B_ACC0FREE : -1, (gg - 1) & 1,
                   need_l0_2 ? B_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
            if (elect_one()) {
              const uint32_t at = tmem + p.tm_h0 + 32;
              const uint64_t bd = stdesc0 + s1 * ststride;
              umma_ts(acc1, at + 16, bd + 4, idesc1, 1);
              umma_ts(acc1, at + 24, bd + 6, idesc1, 1);
              umma_commit(&bars[B_EMPTY + s1]);
              if (last_of_pass) umma_commit(&bars[B_ACC1FULL]);
            }
            __syncwarp();
            TRACE(1, 0x400 | (2 * G));
            if (last_of_pass) {
              // L2 of pass nc: acc2 += H1 (converted in place in acc1's columns) * W2 chunk.
              // The next pass's L1 overwrites acc1 only after these MMAs (same issuing thread,
              // in-order pipe), so no "acc1 free" barrier is needed.
              TRACE(1, 0x2000 | nc);
              // each 64-column slice is consumed as soon as the epilogue has converted it, so the
              // L2 MMAs of slice s overlap the conversion of slice s+1
              WAITW2(B_H1FULL + 0, npass & 1, B_W2FULL, w2n & 1); ++w2n;
              TRACE(1, 0x2200 | (nc << 4));
#pragma unroll 1
              for (int sub = 0; sub < nsl; ++sub) {
                if (sub > 0) WAITW1(B_H1FULL + sub, npass & 1);
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t at = acc1 + sub * 64;
                  const uint64_t bd = w2desc + sub * w2_sub;
#pragma unroll
                  for (int j = 0; j < 4; ++j) umma_ts(acc2, at + 8 * j, bd + 2 * j, idesc2, (nc | sub | j) != 0);
                }
                __syncwarp();
              }
              TRACE(1, 0x2400 | (nc << 4));
              if (elect_one()) umma_commit(&bars[B_W2EMPTY]);
              __syncwarp();
              TRACE(1, 0x2300 | nc);
              ++npass; ++nc; gp = 0;
            } else {
              ++gp;
            }
            if (G + 1 < NGS && !__all_sync(0xffffffffu, rn)) {
              WAITW4(B_FULL + (int)s2, sphase2, B_H0FULL + 0, (hg + 1) & 1,
                     need_l0_2 ? B_ACC0FREE : -1, (gg - 1) & 1,
                     need_l0_2 ? B_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
            }
            TRACE(1, 0x500 | (2 * G));
            ++hg;
            s = s2; sphase = sphase2;
          }
          if (elect_one()) umma_commit(&bars[B_ACC2FULL]);
          __syncwarp();
          TRACE(1, 0x1002);
        }
      }
    } else {
      // =========================== epilogue / compute warps ===========================
      const int e = tid - 64;                              // 0..127
      const int r = (warp & 3) * 32 + lane;                // TMEM lane == row inside the tile
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const int S = p.S, A = p.A, K = p.K;
      const float* sB1 = sBias + p.H;
      const float* sB2 = sBias + 2 * p.H;
      const float* inMean = sNorm;
      const float* inRstd = sNorm + p.SA;   // 1 / in_std
      const float* dMean = sNorm + 2 * p.SA;
      const float* dStd = sNorm + 2 * p.SA + S;
      float* scrA = sScr;
      // own_mode buffers: [128 slots][slot_stride] policy inputs at scrA (shared with the Z staging
      // rows), then -- wide instantiation -- the per-row noise and the hidden activations
      float* sEpsRow = NARROW ? sEpsStatic : scrA + ((TILE_M * p.slot_stride + 3) & ~3);
      float* sHid = NARROW ? sHidStatic : sEpsRow + TILE_M * AMAX;
      const float* polW = NARROW ? static_cast<const float*>(sPolS) : sPol;

      uint32_t gg = 0, a1n = 0, a2n = 0, xn_cnt = 0;   // gg: L0 groups consumed so far
      float x[SMAX], a_raw[AMAX], a_mean[AMAX];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) x[s] = 0.f;
      int ts = 0, nreset = 0;
      int est = 0;   // steps done by this CTA (trace window index, same as the MMA warp's st)
      float ep[AMAX];          // N(0,1) policy noise of the upcoming step
      bool ep_valid = false;
      bool have_action = false;   // own_mode: a_raw of the upcoming step came with the exchange record
      int own_idx = 0;            // own_mode: model this thread's row selects in the current step
      int row = 0;
      bool valid = false;
      float pm_acc = 0.f, pm_gpow = 1.f, pm_dmask = 0.f;   // per_model: cost sum, gamma^t, Ant dones
      auto load_eps = [&](int tt) {
        if (p.eps != nullptr) {
#pragma unroll
          for (int i = 0; i < AMAX; ++i)
            ep[i] = (valid && i < A) ? p.eps[(static_cast<size_t>(tt) * p.B + row) * A + i] : 0.f;
        } else {
#pragma unroll
          for (int blk = 0; blk < AMAX / 4; ++blk) {
            float n4[4];
            if (blk * 4 < A)
              philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(tt), static_cast<uint32_t>(row + p.row_offset),
                             PHILOX_STREAM_EPS + blk, n4);
            else
              n4[0] = n4[1] = n4[2] = n4[3] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) ep[blk * 4 + q] = n4[q];
          }
        }
        ep_valid = true;
      };

      for (int si = 0; si < MAX_SEG; ++si) {
        const int4 sg = segs[si];
        if (sg.x < 0) continue;
        const int tile = sg.x, t0 = sg.y, t1 = sg.z;
        row = tile * TILE_M + r;
        valid = row < p.B;
        ep_valid = false;
        have_action = false;

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
          pm_acc = 0.f; pm_gpow = 1.f; pm_dmask = 0.f;
        } else {
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            x[s] = (valid && s < S) ? __ldcg(&p.row_state[static_cast<size_t>(row) * S + s]) : 0.f;
          ts = __ldcg(&p.row_ts[tile * TILE_M + r]);
          nreset = __ldcg(&p.row_nreset[tile * TILE_M + r]);
        }

        for (int t = t0; t < t1; ++t) {
          TRACE_ON(e == 0 && est >= p.trace_t0 && est < p.trace_t1);
          ++est;
          TRACE(2, 0x1000);
          // ================= begin step: action + Z operand =================
          if (p.ext_actions != nullptr) {
#pragma unroll
            for (int i = 0; i < AMAX; ++i) {
              a_raw[i] = (valid && i < A) ? p.ext_actions[row * A + i] : 0.f;
              a_mean[i] = a_raw[i];
            }
          } else if (have_action) {
            // own_mode: the action of this step was computed by the row's owner at the end of the
            // previous step and arrived with the exchange record (a_raw; act / mean already stored)
          } else {
            // policy mean network (training.py:99-103), fp32 on CUDA cores.  Activations live in
            // the thread's own column of the scratch rows.
            if constexpr (SMAX <= 32) {
              // narrow policies (hidden <= 32): every layer in place, compile-time row stride
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) scrA[s * TILE_M + r] = x[s];
              for (int l = 0; l < p.n_pol_layers; ++l) {
                const PolicyLayer L = p.pl[l];
                const bool last = (l == p.n_pol_layers - 1);
                if (!last) {
                  float acc[HPB];
                  dense_layer<HPB>(scrA, L.nin, sPol + L.w_off, HPB, sPol + L.b_off, acc, r);
#pragma unroll
                  for (int j = 0; j < HPB; ++j) scrA[j * TILE_M + r] = fast_tanh(acc[j]);
                } else {
                  float acc[AMAX];
                  dense_layer<AMAX>(scrA, L.nin, sPol + L.w_off, AMAX, sPol + L.b_off, acc, r);
#pragma unroll
                  for (int i = 0; i < AMAX; ++i) a_mean[i] = p.pol_out_tanh ? tanhf(acc[i]) : acc[i];
                }
              }
            } else {
              // wide policies: hidden layers in blocks of 32 output columns (one block: in place;
              // wider: ping-pong between two scratch regions, see PolicyLayer)
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) scrA[p.pl[0].in_off + s * TILE_M + r] = x[s];
              for (int l = 0; l < p.n_pol_layers; ++l) {
                const PolicyLayer& L = p.pl[l];
                const bool last = (l == p.n_pol_layers - 1);
                if (!last) {
#pragma unroll 1
                  for (int j0 = 0; j0 < L.nout; j0 += HPB) {
                    float acc[HPB];
                    dense_layer<HPB>(scrA + L.in_off, L.nin, sPol + L.w_off + j0, L.npad, sPol + L.b_off + j0, acc, r);
#pragma unroll
                    for (int j = 0; j < HPB; ++j)
                      if (j0 + j < L.nout) scrA[L.out_off + (j0 + j) * TILE_M + r] = fast_tanh(acc[j]);
                  }
                } else {
                  float acc[AMAX];
                  dense_layer<AMAX>(scrA + L.in_off, L.nin, sPol + L.w_off, L.npad, sPol + L.b_off, acc, r);
#pragma unroll
                  for (int i = 0; i < AMAX; ++i) a_mean[i] = p.pol_out_tanh ? tanhf(acc[i]) : acc[i];
                }
              }
            }
            TRACE(2, 0x1010);
            // a = eps * exp(log_std) + mean   (rllab get_actions)
            if (p.determ) {
#pragma unroll
              for (int i = 0; i < AMAX; ++i) a_raw[i] = a_mean[i];
            } else {
              if (!ep_valid) load_eps(t);      // normally prefetched while waiting for the exchange
              ep_valid = false;
#pragma unroll
              for (int i = 0; i < AMAX; ++i) {
                const float ls = fmaxf(sPol[p.pol_logstd_off + i], -13.815510557964274f);
                a_raw[i] = (i < A) ? __fadd_rn(__fmul_rn(ep[i], expf(ls)), a_mean[i]) : 0.f;
              }
            }
            if (p.own_mode && k == 0 && valid) {   // first step of a segment: nobody stored these yet
              const size_t o = static_cast<size_t>(t) * p.B + row;
#pragma unroll
              for (int i = 0; i < AMAX; ++i)
                if (i < A) {
                  if (p.act) p.act[o * A + i] = a_raw[i];
                  if (p.mean) p.mean[o * A + i] = a_mean[i];
                }
            }
          }
          TRACE(2, 0x1011);
          // z = (concat(x, clip(a)) - in_mean) * (1 / in_std), drop leading cols  (training.py:228,146-154)
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s < S) scrA[s * TILE_M + r] = __fmul_rn(__fsub_rn(x[s], inMean[s]), inRstd[s]);
#pragma unroll
          for (int i = 0; i < AMAX; ++i)
            if (i < A) {
              const float u = fminf(fmaxf(a_raw[i], -1.f), 1.f);   // env_helpers.py:599
              scrA[(S + i) * TILE_M + r] = __fmul_rn(__fsub_rn(u, inMean[S + i]), inRstd[S + i]);
            }
          TRACE(2, 0x1012);
          for (int c = 0; c < p.K0 / 16; ++c) {   // 16 k per tcgen05.st: 8 columns of bf16 pairs
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int f0 = 16 * c + 2 * i, f1 = f0 + 1;
              // columns Din, Din+1 are 1.0: they multiply the bf16 hi / lo parts of b0 in W0
              const float z0 = f0 < p.Din ? scrA[(f0 + p.drop) * TILE_M + r] : (f0 < p.Din + 2 ? 1.f : 0.f);
              const float z1 = f1 < p.Din ? scrA[(f1 + p.drop) * TILE_M + r] : (f1 < p.Din + 2 ? 1.f : 0.f);
              pk[i] = pack_bf16x2(z0, z1);
            }
            tmem_st8(tmem + lane_base + p.tm_z + c * 8, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars[B_ZREADY]);
          TRACE(2, 0x1001);
          {
            // own_mode: policy noise of step t+1 for this thread's row, parked in shared memory
            // (same thread reads it back at the end of the step if it owns the row)
            if (p.own_mode) {
              if (!p.determ && t + 1 < t1) {
                load_eps(t + 1);
                ep_valid = false;
#pragma unroll
                for (int i = 0; i < AMAX; i += 4)
                  *reinterpret_cast<float4*>(&sEpsRow[r * AMAX + i]) = make_float4(ep[i], ep[i + 1], ep[i + 2], ep[i + 3]);
              }
              // model index of this step, off the critical path (ts / nreset are final here)
              if (p.model_idx != nullptr)
                own_idx = valid ? p.model_idx[static_cast<size_t>(t) * p.B + row] : 0;
              else if (p.sam_mode == METRPO_SAM_STEP_RAND)
                own_idx = philox_index(p.seed, p.offset + static_cast<unsigned long long>(t),
                                       static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_IDX, K);
              else
                own_idx = philox_index(p.seed, static_cast<uint64_t>(static_cast<uint32_t>(nreset)),
                                       static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EIDX, K);
              own_idx = min(max(own_idx, 0), K - 1);
            }
          }

          // ================= per-group epilogues (2 chunks of 64 hidden columns) =================
          {
            const int NG = p.KC / 2, NGS = NCH / 2;
            int egp = 0, enc = 0;   // group position inside the pass / pass index
            for (int G = 0; G < NGS; ++G) {
              st_dbg = (t << 8) | (2 * G);
              {
                uint32_t v0[32], v1[32], v2[32], v3[32], pk[32], pk2[32];
                TRACE(2, 0x100 | (2 * G));
                WAITB(B_ACC0FULL + (gg & 1), (gg >> 1) & 1);
                TRACE(2, 0x200 | (2 * G));
                tc_fence_after();
                tmem_ld32(tmem + lane_base + p.tm_acc0, v0);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 32, v1);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 64, v2);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 96, v3);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_ACC0FREE]);        // next group's L0 may overwrite acc0
                TRACE(2, 0x300 | (2 * G));
                relu_pack<false>(v0, v1, nullptr, pk);
                relu_pack<false>(v2, v3, nullptr, pk2);
                // H0 buffer 0 is free once L1 of chunk 2*gg-2 completed (its stage's EMPTY barrier)
                if (gg >= 1) WAITB(B_EMPTY + ((2 * gg - 2) & 3), ((2 * gg - 2) >> 2) & 1);
                tmem_st32(tmem + lane_base + p.tm_h0, pk);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_H0FULL + 0]);
                TRACE(2, 0x400 | (2 * G));
                if (gg >= 1) WAITB(B_EMPTY + ((2 * gg - 1) & 3), ((2 * gg - 1) >> 2) & 1);
                tmem_st32(tmem + lane_base + p.tm_h0 + 32, pk2);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[B_H0FULL + 1]);
                TRACE(2, 0x600 | (2 * G));
                ++gg;
              }
              // layer-1 pass epilogue.  Pass nc is drained after the first group of pass nc+1 has
              // been staged (keeps the MMA warp fed across the pass boundary); the last pass right
              // after its last group.
              int drain_nc = -1;
              if (G == NGS - 1) drain_nc = p.NC - 1;
              else if (egp == 0 && enc > 0) drain_nc = enc - 1;
              if (++egp == NG) { egp = 0; ++enc; }
              if (drain_nc >= 0) {
                TRACE(2, 0x700 | (2 * G));
                WAITB(B_ACC1FULL, a1n & 1);
                TRACE(2, 0x800 | (2 * G));
                tc_fence_after();
                {
                  // software pipelined: the TMEM loads of slice s+1 are in flight while slice s is
                  // converted and stored (slices touch disjoint columns)
                  uint32_t va0[32], va1[32], vb0[32], vb1[32], pk[32];
                  const uint32_t a1 = tmem + lane_base + TM_ACC1;
                  const float* bb = sB1 + drain_nc * N1;
                  constexpr int nsl = N1 / 64;
                  tmem_ld32(a1, va0);
                  tmem_ld32(a1 + 32, va1);
                  tmem_ld_wait();
#define DRAIN_SLICE(sub, cur0, cur1, nxt0, nxt1, has_next_)                                   \
                  if ((sub) < nsl) {                                                          \
                  const bool has_next = (has_next_) && (sub) + 1 < nsl;                       \
                  if (has_next) {                                                             \
                    tmem_ld32(a1 + ((sub) + 1) * 64, nxt0);                                   \
                    tmem_ld32(a1 + ((sub) + 1) * 64 + 32, nxt1);                              \
                  }                                                                           \
                  relu_pack<true>(cur0, cur1, bb + (sub) * 64, pk);                           \
                  /* in place: the 64 fp32 columns just read become 32 columns of bf16 pairs */ \
                  tmem_st32(a1 + (sub) * 64, pk);                                             \
                  tmem_st_wait();                                                             \
                  tc_fence_before();                                                          \
                  mbar_arrive(&bars[B_H1FULL + (sub)]);                                       \
                  if (has_next) tmem_ld_wait();                                               \
                  }
                  DRAIN_SLICE(0, va0, va1, vb0, vb1, true)
                  DRAIN_SLICE(1, vb0, vb1, va0, va1, true)
                  DRAIN_SLICE(2, va0, va1, vb0, vb1, true)
                  DRAIN_SLICE(3, vb0, vb1, va0, va1, false)
#undef DRAIN_SLICE
                }
                ++a1n;
                TRACE(2, 0x900 | (2 * G));
              }
            }
          }

          // ================= finish step: candidate, exchange, select, reward, reset =========
          float cand[SMAX];
          {
            TRACE(2, 0x1002);
            WAITB(B_ACC2FULL, a2n & 1); ++a2n;
            TRACE(2, 0x1003);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < SMAX / 32; ++c) {
              uint32_t v[32];
              if (32 * c < p.S_pad) {     // warp-uniform
                tmem_ld32(tmem + lane_base + p.tm_acc2 + 32 * c, v);
                tmem_ld_wait();
              }
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const int s = 32 * c + q;
                if (s < S) {
                  const float o = __fadd_rn(__uint_as_float(v[q]), sB2[s]);
                  // tf.add(diff_mean + diff_std * nn_output, x)   (training.py:257)
                  cand[s] = __fadd_rn(__fadd_rn(dMean[s], __fmul_rn(dStd[s], o)), x[s]);
                } else {
                  cand[s] = 0.f;
                }
              }
            }
          }
          bool handled = false;
          {
            if (p.own_mode) {
              // ============ row-ownership exchange (step_rand / eps_rand) ============
              // Every CTA of the gang knows which model each row selects this step.  The CTA of
              // that model already holds the row's next state (its own candidate), so it alone
              // computes reward / done / reset and the NEXT step's action for the row -- with four
              // threads per owned row (~128/K rows) instead of every CTA redundantly running the
              // policy for all 128 rows -- and publishes one record [x_new | a_raw | done] per row.
              handled = true;
              const int RS = p.rec_stride, SPs = p.slot_stride;
              float* rec = p.xbuf + static_cast<size_t>(slot * 2 + (xn_cnt & 1)) * p.xbuf_stride;
              const bool own = valid && own_idx == k;
              const bool want_pol = (t + 1 < t1);
              TRACE(2, 0x1020);
              float own_reward = 0.f;
              bool own_dn = false;
              if (own) {
                float u[AMAX];
#pragma unroll
                for (int i = 0; i < AMAX; ++i) u[i] = fminf(fmaxf(a_raw[i], -1.f), 1.f);
                const float reward = -env_cost<SMAX, AMAX>(p.env_id, S, A, cand, u);   // env_helpers.py:601
                const bool dn = env_is_done<SMAX>(p.env_id, S, cand) || (ts + 1 >= p.T_max);   // :603-604
                own_reward = reward; own_dn = dn;   // (trajectory stores happen after the release below)
                if (dn) {   // :605-606 -> reset(dones)
                  const float* src = p.reset_pool + static_cast<size_t>((static_cast<long long>(nreset) * p.B + row) % p.R) * S;
#pragma unroll
                  for (int s = 0; s < SMAX; ++s)
                    if (s < S) cand[s] = src[s];
                }
                float4* rq = reinterpret_cast<float4*>(rec + r * RS);
                rq[AMAX / 4] = make_float4(dn ? 1.f : 0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < SMAX / 4; ++q)
                  if (4 * q < S) rq[AMAX / 4 + 1 + q] = make_float4(cand[4 * q], cand[4 * q + 1], cand[4 * q + 2], cand[4 * q + 3]);
              }
              // ---- compact the owned rows of the tile ----
              const unsigned om = __ballot_sync(0xffffffffu, own);
              if (lane == 0) sCnt[warp & 3] = __popc(om);
              named_bar_sync(1, EPI_THREADS);
              int off = 0, n_own = 0;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const int c = sCnt[w];
                if (w < (warp & 3)) off += c;
                n_own += c;
              }
              if (own) {
                const int j = off + __popc(om & ((1u << lane) - 1u));
                sList[j] = r;
                float* in = scrA + j * SPs;
#pragma unroll
                for (int s = 0; s < SMAX; ++s)
                  if (s < S) in[s] = cand[s];
                if (want_pol && !p.determ) {
#pragma unroll
                  for (int i = 0; i < AMAX; ++i)
                    if (i < A) in[S + i] = sEpsRow[r * AMAX + i];
                }
              }
              named_bar_sync(1, EPI_THREADS);
              TRACE(2, 0x1021);
              if constexpr (NARROW) {
              // ---- policy of step t+1 for the owned rows: 2 threads per row, 64 rows per pass (one
              //      pass unless a CTA owns more than half of the tile) ----
              if (want_pol) {
                const int part = e & 1, jl = e >> 1;
                const int nl = p.n_pol_layers;
                for (int base = 0; base < n_own; base += 64) {
                  const int j = base + jl;
                  const bool actv = j < n_own;
                  const float* cur = scrA + (actv ? j : 0) * SPs;
                  for (int l = 0; l < nl - 1; ++l) {
                    const PolicyLayer& L = p.pl[l];
                    float acc[16];
                    {
                      const float4* b4 = reinterpret_cast<const float4*>(sPolS + L.b_off + 16 * part);
#pragma unroll
                      for (int q = 0; q < 4; ++q) {
                        const float4 b = b4[q];
                        acc[4 * q] = b.x; acc[4 * q + 1] = b.y; acc[4 * q + 2] = b.z; acc[4 * q + 3] = b.w;
                      }
                    }
                    const float* W = sPolS + L.w_off + 16 * part;
#pragma unroll 4
                    for (int i = 0; i < L.nin; ++i) {
                      const float xi = cur[i];
                      const float4* w4 = reinterpret_cast<const float4*>(W + i * HPB);
#pragma unroll
                      const float2 xx = make_float2(xi, xi);
#pragma unroll
                      for (int q = 0; q < 4; ++q) {
                        const float4 w = w4[q];
                        const float2 r0 = __ffma2_rn(xx, make_float2(w.x, w.y), make_float2(acc[4 * q], acc[4 * q + 1]));
                        const float2 r1 = __ffma2_rn(xx, make_float2(w.z, w.w), make_float2(acc[4 * q + 2], acc[4 * q + 3]));
                        acc[4 * q] = r0.x; acc[4 * q + 1] = r0.y; acc[4 * q + 2] = r1.x; acc[4 * q + 3] = r1.y;
                      }
                    }
                    float* out = sHid + (l & 1) * (64 * 33) + jl * 33 + 16 * part;
#pragma unroll
                    for (int c = 0; c < 16; ++c) out[c] = fast_tanh(acc[c]);
                    __syncwarp();
                    cur = sHid + (l & 1) * (64 * 33) + jl * 33;
                  }
                  const PolicyLayer& L = p.pl[nl - 1];
                  float mo[AMAX / 2];
#pragma unroll
                  for (int c = 0; c < AMAX / 2; ++c) mo[c] = sPolS[L.b_off + (AMAX / 2) * part + c];
                  {
                    const float* W = sPolS + L.w_off + (AMAX / 2) * part;
#pragma unroll 4
                    for (int i = 0; i < L.nin; ++i) {
                      const float xi = cur[i];
                      const float4 w = *reinterpret_cast<const float4*>(W + i * AMAX);
                      mo[0] = __fmaf_rn(xi, w.x, mo[0]); mo[1] = __fmaf_rn(xi, w.y, mo[1]);
                      mo[2] = __fmaf_rn(xi, w.z, mo[2]); mo[3] = __fmaf_rn(xi, w.w, mo[3]);
                    }
                  }
                  if (p.pol_out_tanh) {
#pragma unroll
                    for (int c = 0; c < AMAX / 2; ++c) mo[c] = tanhf(mo[c]);
                  }
                  if (actv) {
                    const int rowl = sList[j];
                    const float* in = scrA + j * SPs;
                    const size_t o = static_cast<size_t>(t + 1) * p.B + (tile * TILE_M + rowl);
#pragma unroll
                    for (int q = 0; q < AMAX / 2; ++q) {
                      const int c = (AMAX / 2) * part + q;
                      if (c < A) {
                        const float mu = mo[q];
                        float raw = mu;
                        if (!p.determ) {   // a = eps * exp(log_std) + mean   (rllab get_actions)
                          const float ls = fmaxf(sPolS[p.pol_logstd_off + c], -13.815510557964274f);
                          raw = __fadd_rn(__fmul_rn(in[S + c], expf(ls)), mu);
                        }
                        if (p.act) p.act[o * A + c] = raw;
                        if (p.mean) p.mean[o * A + c] = mu;
                        rec[rowl * RS + c] = raw;
                      }
                    }
                  }
                  __syncwarp();
                }
              }
              } else {
              // ---- wide policies: policy of step t+1 for the owned rows, 16 rows per pass.  The fp32
              //      weights (64 KB for 100-50-25) do not fit in shared memory next to the operand
              //      rings, so the CTA streams them from L2 ONCE per pass: thread c owns output column
              //      c of the layer for all 16 rows (coalesced, independent weight loads, 8 in
              //      flight), the 16 inputs of a row of W are read as four broadcast float4 from a
              //      transposed activation tile [input][16 rows]. ----
              if (want_pol) {
                const int nl = p.n_pol_layers;
                float* bufA = sHid;                       // [<=128 inputs][16 rows]
                float* bufB = sHid + HPMAX * 16;
                for (int base = 0; base < n_own; base += 16) {
                  for (int q = e; q < 16 * S; q += EPI_THREADS) {   // transpose the pass's inputs
                    const int rr = q / S, si = q - rr * S;
                    bufA[si * 16 + rr] = (base + rr < n_own) ? scrA[(base + rr) * SPs + si] : 0.f;
                  }
                  named_bar_sync(1, EPI_THREADS);
                  float* bin = bufA;
                  float* bout = bufB;
                  for (int l = 0; l < nl; ++l) {
                    const PolicyLayer& L = p.pl[l];
                    const bool last = (l == nl - 1);
                    if (e < L.npad) {
                      float acc[16];
                      const float bia = polW[L.b_off + e];
#pragma unroll
                      for (int rr = 0; rr < 16; ++rr) acc[rr] = bia;
                      const float* W = polW + L.w_off + e;
#pragma unroll 8
                      for (int i = 0; i < L.nin; ++i) {
                        const float w = W[i * L.npad];
                        const float4* x4 = reinterpret_cast<const float4*>(bin + i * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                          const float4 xv = x4[q];
                          acc[4 * q] = __fmaf_rn(xv.x, w, acc[4 * q]); acc[4 * q + 1] = __fmaf_rn(xv.y, w, acc[4 * q + 1]);
                          acc[4 * q + 2] = __fmaf_rn(xv.z, w, acc[4 * q + 2]); acc[4 * q + 3] = __fmaf_rn(xv.w, w, acc[4 * q + 3]);
                        }
                      }
                      if (!last) {
                        float4* o4 = reinterpret_cast<float4*>(bout + e * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                          o4[q] = make_float4(fast_tanh(acc[4 * q]), fast_tanh(acc[4 * q + 1]), fast_tanh(acc[4 * q + 2]),
                                              fast_tanh(acc[4 * q + 3]));
                      } else if (e < A) {
                        const float sg = p.determ ? 0.f : expf(fmaxf(polW[p.pol_logstd_off + e], -13.815510557964274f));
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr)
                          if (base + rr < n_own) {
                            const int rowl = sList[base + rr];
                            const float mu = p.pol_out_tanh ? tanhf(acc[rr]) : acc[rr];
                            float raw = mu;   // a = eps * exp(log_std) + mean   (rllab get_actions)
                            if (!p.determ) raw = __fadd_rn(__fmul_rn(scrA[(base + rr) * SPs + S + e], sg), mu);
                            const size_t o = static_cast<size_t>(t + 1) * p.B + (tile * TILE_M + rowl);
                            if (p.act) p.act[o * A + e] = raw;
                            if (p.mean) p.mean[o * A + e] = mu;
                            rec[rowl * RS + e] = raw;
                          }
                      }
                    }
                    named_bar_sync(1, EPI_THREADS);
                    float* tmp = bin; bin = bout; bout = tmp;
                  }
                }
              }
              }
              TRACE(2, 0x1022);
              // ---- publish / meet the gang: one release + one acquire poll per warp ----
              __syncwarp();
              int okw = 1;
              if (lane == 0) red_release_gpu_add(&p.xctr[slot], 1u);
              if (own) {   // the owner's part of the trajectory record, off the gang's critical path
                const size_t o = static_cast<size_t>(t) * p.B + row;
                if (p.obs) {
#pragma unroll
                  for (int s = 0; s < SMAX; ++s)
                    if (s < S) p.obs[o * S + s] = x[s];
                }
                if (p.rew) p.rew[o] = own_reward;
                if (p.done) p.done[o] = own_dn ? 1 : 0;
              }
              if (lane == 0)
                okw = wait_ge(&p.xctr[slot], 4u * static_cast<unsigned>(K) * (xn_cnt + 1), p.dbg, 101u,
                              (uint32_t)st_dbg) ? 1 : 0;
              okw = __shfl_sync(0xffffffffu, okw, 0);
              if (!okw) goto bail;
              TRACE(2, 0x1023);
              ++xn_cnt;
              float dnf = 0.f;
              if (valid) {
                const float4* rq = reinterpret_cast<const float4*>(rec + r * RS);
                float4 qa[AMAX / 4], qx[SMAX / 4];
#pragma unroll
                for (int q = 0; q < AMAX / 4; ++q) qa[q] = want_pol ? __ldcg(rq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 qd = __ldcg(rq + AMAX / 4);
#pragma unroll
                for (int q = 0; q < SMAX / 4; ++q)
                  if (4 * q < S) qx[q] = __ldcg(rq + AMAX / 4 + 1 + q);
#pragma unroll
                for (int q = 0; q < AMAX / 4; ++q) {
                  a_raw[4 * q] = qa[q].x; a_raw[4 * q + 1] = qa[q].y; a_raw[4 * q + 2] = qa[q].z; a_raw[4 * q + 3] = qa[q].w;
                }
#pragma unroll
                for (int i = 0; i < AMAX; ++i)
                  if (i >= A) a_raw[i] = 0.f;
#pragma unroll
                for (int q = 0; q < SMAX / 4; ++q)
                  if (4 * q < S) {
                    x[4 * q] = qx[q].x; x[4 * q + 1] = qx[q].y; x[4 * q + 2] = qx[q].z; x[4 * q + 3] = qx[q].w;
                  }
#pragma unroll
                for (int s = 0; s < SMAX; ++s)
                  if (s >= S) x[s] = 0.f;
                dnf = qd.x;
              }
              if (dnf != 0.f) { ts = 0; nreset += 1; } else { ts += 1; }
              have_action = want_pol;
              TRACE(2, 0x1005);
            }
          }
          if (!handled) {
          float xnext[SMAX];
          const int mode = p.sam_mode;
          if (K == 1 || p.per_model) {
#pragma unroll
            for (int s = 0; s < SMAX; ++s) xnext[s] = cand[s];
          } else {
            float* xb = p.xbuf + static_cast<size_t>(slot * 2 + (xn_cnt & 1)) * p.xbuf_stride;
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) xb[(k * S + s) * TILE_M + r] = cand[s];
            TRACE(2, 0x1020);
            // useful work while the candidate stores drain and the peers catch up: the parts of
            // the trajectory record that are already known, and the next step's policy noise
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
            }
            if (!p.determ && p.ext_actions == nullptr && t + 1 < t1) load_eps(t + 1);
            // CTA barrier orders the 128 threads' stores before thread 0's gpu-scope release
            named_bar_sync(1, EPI_THREADS);
            TRACE(2, 0x1021);
            if (e == 0) {
              red_release_gpu_add(&p.xctr[slot], 1u);
              TRACE(2, 0x1022);
              if (!wait_ge(&p.xctr[slot], static_cast<unsigned>(K) * (xn_cnt + 1), p.dbg, 101u,
                           (uint32_t)st_dbg))
                abort_smem = 1;
              TRACE(2, 0x1023);
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
                  idx = philox_index(p.seed, p.offset + static_cast<unsigned long long>(t),
                                     static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_IDX, K);
                else
                  idx = philox_index(p.seed, static_cast<uint64_t>(static_cast<uint32_t>(nreset)),
                                     static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EIDX, K);
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
                    philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(t),
                                   static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_STD + (s >> 2), n4);
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
          TRACE(2, 0x1024);
          // reward = -cost_np_vec(s, clip(a), s')   (env_helpers.py:601)
          float u[AMAX];
#pragma unroll
          for (int i = 0; i < AMAX; ++i) u[i] = fminf(fmaxf(a_raw[i], -1.f), 1.f);
          const float reward = -env_cost<SMAX, AMAX>(p.env_id, S, A, xnext, u);
          if (p.per_model) {
            // _policy_cost += gamma**t * cost_tf(x, u, x_next[, dones]); dones = max(dones,
            // is_done_tf(x, x_next)) AFTER the cost (model_based_rl.py:133-139; Ant masks the cost
            // of rows that already terminated, envs/com_ant_env.py:70-75)
            const float c = __fmul_rn(-reward, 1.f - pm_dmask);
            pm_acc = __fadd_rn(pm_acc, __fmul_rn(pm_gpow, c));
            pm_gpow = __fmul_rn(pm_gpow, p.gamma);
            if (env_is_done<SMAX>(p.env_id, S, xnext)) pm_dmask = 1.f;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) x[s] = xnext[s];
            continue;
          }
          ts += 1;
          const bool dn = env_is_done<SMAX>(p.env_id, S, xnext) || (ts >= p.T_max);   // :603-604
          if (k == 0 && valid) {
            const size_t o = static_cast<size_t>(t) * p.B + row;
            if (K == 1) {   // (with K > 1 these were stored while waiting for the exchange)
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
          }  // !handled
        }  // t

        // ---- segment end: publish the tile's state ----
        if (p.per_model) {
          if (valid) p.pm_cost[static_cast<size_t>(k) * p.B + row] = pm_acc;
        } else if (k == 0) {
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
