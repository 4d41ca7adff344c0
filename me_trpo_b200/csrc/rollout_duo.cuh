// Two-stream ("duo") variant of the persistent ensemble-rollout kernel (sm_100a).
//
// Why.  In rollout_kernel.cuh one CTA runs ONE (model, row tile) chain, and every step ends in a
// serial section (candidate -> reward / done / reset -> policy of the owned rows -> gang exchange
// -> next Z operand, ~19 k of ~85 k cycles at the half-cheetah bench shape) during which the
// tensor pipe idles: there is nothing else on the SM to run.  Here a CTA runs TWO independent
// chains ("streams", two different row tiles of the same model) with one epilogue warpgroup each;
// the MMA warp alternates between them step by step, so the serial section of stream A overlaps
// the layer-1 MMAs of stream B.
//
// Column split (cs = 2).  The bench shape has only 32 tiles x 5 models = 160 chains for 148 SMs,
// i.e. ~1 chain per SM.  With cs = 2 a chain is cut along the hidden dimension: CTA c of a pair
// computes layer-1 output columns [c*H/2, (c+1)*H/2) (its NC/2 passes) for BOTH tiles of the pair,
// and layer 2 yields a partial sum over those columns; the two halves exchange the [128, S] fp32
// partials through L2 every step (release/acquire on a per-(tile, model) counter) and add them in
// a fixed order, so both hold the same candidate.  The owned rows of a tile (rows whose selected
// model is k) are dealt alternately to the two halves, which halves the policy work per CTA.
// Per CTA the MMA count per pair-step equals that of one full-width step of the old kernel.
// With cs = 1 (enough tiles per gang slot: hopper / ant shapes) a CTA simply owns two full-width
// tiles and no partials are exchanged.
//
// Shared between the two streams of a CTA (used strictly one stream at a time, in MMA issue order):
// the W1 / W0 / W2 weight rings, and the TMEM accumulators acc1 / acc0 / acc2 and the H0 operand
// buffers.  Private per stream: the Z operand columns, the epilogue warpgroup with its row state in
// registers, its shared-memory scratch.  An epilogue group touches the shared barriers only after
// its private START barrier fired (committed by the MMA warp behind the first layer-0 group of the
// stream-step: in-order completion implies every MMA of the other stream has retired), which rules
// out parity aliasing on barriers that advanced while the group was busy elsewhere.
//
// Scope: the <32, 8, 256> shapes (every shipped env but humanoid) in row-ownership mode (step_rand /
// eps_rand, K > 1), fused run / continue.  Everything else stays on rollout_kernel.cuh.
#pragma once
#include "rollout_kernel.cuh"

namespace metrpo {

// Three warpgroups so that the register file can be re-partitioned with setmaxnreg (which acts on
// whole, aligned warpgroups): WG0 = producer warp + MMA warp (+ two idle warps), WG1 / WG2 = the two
// epilogue warpgroups.  384 threads launch with 168 registers each; WG0 shrinks to DUO_REGS_LO and
// the epilogue warpgroups grow to DUO_REGS_HI (128 * 56 + 256 * 224 = 64 512 <= 65 536).
constexpr int DUO_THREADS = 3 * EPI_THREADS;
constexpr int DUO_REGS_LO = 56;
constexpr int DUO_REGS_HI = 224;
__device__ __forceinline__ void setmaxnreg_dec_lo() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DUO_REGS_LO));
}
__device__ __forceinline__ void setmaxnreg_inc_hi() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DUO_REGS_HI));
}
constexpr int DUO_POL_ROWS = 32;                    // owned rows per policy pass (2 threads per row); DuoParams::pol_rows
                                                    // shrinks it to 16 / 8 when shared memory is tight (K0 = 48 shapes)

enum {
  D_FULL = 0,        // [NSTAGE] W1 stage landed (tx)
  D_EMPTY = 4,       // [NSTAGE] L1 of the chunk in this stage completed (commit)
  D_W2FULL = 8,
  D_W2EMPTY = 9,
  D_W0FULL = 10,     // [2]
  D_ACC0FULL = 12,   // [2] L0 group accumulated (commit); frees W0 ring slot
  D_ACC0FREE = 14,   // acc0 loaded to registers (128 arrivals)
  D_H0FULL = 15,     // [2] (128 arrivals)
  D_ACC1FULL = 17,
  D_H1FULL = 18,     // [4] (128 arrivals)
  D_ACC2FULL = 22,
  D_ACC2FREE = 23,   // acc2 loaded to registers (128 arrivals): the other stream's L2 may overwrite it
  D_ZREADY = 24,     // [2 streams] Z operand written (128 arrivals)
  D_START = 26,      // [2 streams] first L0 group of the stream-step retired (commit)
  D_ZFREE = 28,      // z_shared: every L0 of a stream-step retired (commit): the single Z slot may be rewritten
  D_NUM_BARS = 29
};

struct DuoParams {
  KParams b;                    // everything the single-stream kernel takes
  int cs;                       // column split: 1 or 2 CTAs per (slot, model)
  int NCp;                      // layer-1 passes per CTA and stream-step = NC / cs
  int n_pairs;                  // tile pairs (segments index pairs)
  float* pbuf;                  // cs = 2: [slot*2 + stream][2 parity][K][2 halves][S][128] acc2 partials
  unsigned* pctr;               // cs = 2: [(slot*2 + stream)*K + k]
  unsigned long long pbuf_stride;   // floats per (slot, stream, parity)
  float* rbuf;                  // [slot*2 + stream][2 parity][128 * rec_stride] row records
  unsigned* rctr;               // [slot*2 + stream]
  unsigned long long rbuf_stride;
  uint32_t off_scr[2], off_hid[2], off_list[2];   // per-stream shared memory areas
  int pol_rows;                 // owned rows per policy pass: DUO_POL_ROWS, or 16 / 8 (sizes the hidden-activation scratch)
  int z_shared;                 // 1: TMEM has room for ONE Z operand only (K0 = 48, ant): the streams take
                                //    turns, a group writes its Z after the other stream's last L0 retired
};

// Fully inlined bounded waits: a register-re-partitioned region (setmaxnreg) must not contain ABI
// calls -- ptxas cannot allocate such a region ("Register allocation failed with register count of
// 224" with the __noinline__ helpers of rollout_kernel.cuh).
__device__ __forceinline__ bool dwait_bar(uint64_t* bar, uint32_t parity, unsigned* dbg, uint32_t tag,
                                          uint32_t info, uint32_t sleep_ns = 0) {
  if (mbar_try_wait(bar, parity)) return true;
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
__device__ __forceinline__ bool dwait_ge(const unsigned* ptr, unsigned target, unsigned* dbg, uint32_t tag,
                                         uint32_t info) {
  if (ld_acquire_gpu(ptr) >= target) return true;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (ld_acquire_gpu(ptr) < target) {
    if ((++spins & 0x3f) == 0) {
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
#define DWAITB(idx, par) \
  do { if (!dwait_bar(&bars[idx], (par), p.dbg, (idx), (uint32_t)st_dbg)) goto bail; } while (0)

template <int SMAX, int AMAX, int N1>
__global__ void __launch_bounds__(DUO_THREADS, 1) rollout_duo_kernel(const __grid_constant__ DuoParams dp) {
  const KParams& p = dp.b;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ int abort_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = dp.cs;
  const int c = blockIdx.x % cs;
  const int k = (blockIdx.x / cs) % p.K;
  const int slot = blockIdx.x / (cs * p.K);
  const int NCp = dp.NCp;

  uint8_t* sStage = smem + p.off_stage;
  uint8_t* sW0g = smem + p.off_sw0g;
  uint8_t* sW2 = smem + p.off_sw2;
  float* sBias = reinterpret_cast<float*>(smem + p.off_sbias);   // [b1 of my passes: NCp*N1 | b2 BIAS_PAD]
  float* sNorm = reinterpret_cast<float*>(smem + p.off_snorm);
  float* sPolS = reinterpret_cast<float*>(smem + p.off_spol);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);

  const int NCHp = NCp * p.KC;          // chunks per stream-step
  const int NGSp = NCHp / 2;            // L0 groups per stream-step (even)
  const int4* segs = p.segs + slot * MAX_SEG;
  int total_psteps = 0;                 // pair-steps of this slot
  for (int i = 0; i < MAX_SEG; ++i) {
    int4 sg = segs[i];
    if (sg.x >= 0) total_psteps += sg.z - sg.y;
  }
  const int total_ssteps = 2 * total_psteps;

  int st_dbg = 0;
#ifdef METRPO_TRACE
  bool tr_on = false;
  int tr_n = 0;
#endif
  if (tid == 0) {
    abort_smem = 0;
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars[D_FULL + i], 1); mbar_init(&bars[D_EMPTY + i], 1); }
    mbar_init(&bars[D_W2FULL], 1); mbar_init(&bars[D_W2EMPTY], 1);
    mbar_init(&bars[D_ACC0FREE], EPI_THREADS);
    mbar_init(&bars[D_ACC2FREE], EPI_THREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[D_W0FULL + i], 1); mbar_init(&bars[D_ACC0FULL + i], 1);
      mbar_init(&bars[D_H0FULL + i], EPI_THREADS);
      mbar_init(&bars[D_ZREADY + i], EPI_THREADS);
      mbar_init(&bars[D_START + i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[D_H1FULL + i], EPI_THREADS);
    mbar_init(&bars[D_ACC1FULL], 1);
    mbar_init(&bars[D_ACC2FULL], 1);
    mbar_init(&bars[D_ZFREE], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  {  // constants -> smem
    const float* gb = p.bias + static_cast<size_t>(k) * (2 * p.H + BIAS_PAD);
    const int nb1 = NCp * N1;
    for (int i = tid; i < nb1; i += DUO_THREADS) sBias[i] = gb[p.H + c * nb1 + i];
    for (int i = tid; i < BIAS_PAD; i += DUO_THREADS) sBias[nb1 + i] = gb[2 * p.H + i];
    for (int i = tid; i < 2 * p.SA + 2 * p.S; i += DUO_THREADS) {
      const float v = p.norm[i];
      sNorm[i] = (i >= p.SA && i < 2 * p.SA) ? __frcp_rn(v) : v;
    }
    for (int i = tid; i < p.pol_floats; i += DUO_THREADS) sPolS[i] = p.pol[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TMEM columns between the two streams' Z operands (0: one shared slot)
  const uint32_t zcols = dp.z_shared ? 0u : static_cast<uint32_t>(p.K0 / 2);

  if (warp < 4) {
    setmaxnreg_dec_lo();
    if (total_ssteps == 0) {
    } else if (warp == 0) {
      // =========================== producer ===========================
      if (lane == 0) {
        const uint8_t* wm = p.wstream + static_cast<size_t>(k) * p.model_stride;
        const uint64_t pol = l2_policy_evict_last();
        uint32_t s = 0, sphase = 0, w2n = 0, wg = 0;
        const int NG = p.KC / 2;
        const uint32_t total_groups = static_cast<uint32_t>(total_ssteps) * NGSp;
        const int w2_at = p.KC > 3 ? 3 : p.KC - 1;
#define DUO_LOAD_W0()                                                                         \
  do {                                                                                        \
    if (wg < total_groups) {                                                                  \
      const uint32_t ws = wg & 1;                                                             \
      if (wg >= 2) {                                                                          \
        if (!dwait_bar(&bars[D_ACC0FULL + ws], ((wg >> 1) - 1) & 1, p.dbg, D_ACC0FULL + ws,    \
                      (uint32_t)st_dbg, METRPO_PRODUCER_SLEEP_NS)) goto bail;                 \
      }                                                                                       \
      mbar_arrive_expect_tx(&bars[D_W0FULL + ws], p.w0g_bytes);                               \
      bulk_g2s_hint(sW0g + ws * p.w0g_bytes,                                                  \
                    wm + p.off_w0g + static_cast<size_t>(wg % NG) * p.w0g_bytes, p.w0g_bytes, \
                    &bars[D_W0FULL + ws], pol);                                               \
      ++wg;                                                                                   \
    }                                                                                         \
  } while (0)
        DUO_LOAD_W0();
        for (int u = 0; u < total_ssteps; ++u) {
          for (int ncl = 0; ncl < NCp; ++ncl) {
            const int nc = c * NCp + ncl;
            const uint8_t* src = wm + static_cast<size_t>(nc) * p.KC * p.stage_bytes;
            for (int kc = 0; kc < p.KC; ++kc) {
              st_dbg = (u << 8) | (ncl * p.KC + kc);
              if ((kc & 1) == 0) DUO_LOAD_W0();
              if (!dwait_bar(&bars[D_EMPTY + s], sphase ^ 1, p.dbg, D_EMPTY + s, (uint32_t)st_dbg,
                            METRPO_PRODUCER_SLEEP_NS)) goto bail;
              mbar_arrive_expect_tx(&bars[D_FULL + s], p.stage_bytes);
              bulk_g2s_hint(sStage + s * p.stage_bytes, src, p.stage_bytes, &bars[D_FULL + s], pol);
              src += p.stage_bytes;
              if (++s == NSTAGE) { s = 0; sphase ^= 1; }
              if (kc == w2_at) {
                if (!dwait_bar(&bars[D_W2EMPTY], (w2n & 1) ^ 1, p.dbg, D_W2EMPTY, (uint32_t)st_dbg,
                              METRPO_PRODUCER_SLEEP_NS)) goto bail;
                mbar_arrive_expect_tx(&bars[D_W2FULL], p.w2chunk_bytes);
                bulk_g2s_hint(sW2, wm + p.off_w2 + static_cast<size_t>(nc) * p.w2chunk_bytes,
                              p.w2chunk_bytes, &bars[D_W2FULL], pol);
                ++w2n;
              }
            }
          }
        }
#undef DUO_LOAD_W0
      }
    } else if (warp == 1) {
      // =========================== MMA issuer ===========================
      const uint32_t idesc0 = idesc_bf16_f32(128, 128);
      const uint32_t idesc1 = idesc_bf16_f32(128, N1);
      const uint32_t idesc2 = idesc_bf16_f32(128, p.S_pad);
      const int k0steps = p.K0 / 16;
      const uint32_t w0_kstep = (2 * 128 * 16) >> 4;
      const uint64_t w0desc0 = smem_desc_noswz(smem_u32(sW0g), 128 * 16, 128);
      const uint32_t w0slot = p.w0g_bytes >> 4;
      const uint64_t w2desc = smem_desc_sw128(smem_u32(sW2));
      const uint32_t w2_sub = (p.S_pad * 128) >> 4;
      const uint64_t stdesc0 = smem_desc_sw128(smem_u32(sStage));
      const uint32_t ststride = p.stage_bytes >> 4;
      const uint32_t acc0 = tmem + p.tm_acc0, acc1 = tmem + TM_ACC1, acc2 = tmem + p.tm_acc2;
      constexpr int nsl = N1 / 64;

#define DWAITW4(i0, p0, i1, p1, i2, p2, i3, p3)                                               \
  do {                                                                                        \
    bool ok_ = true;                                                                          \
    const int wi_ = lane == 0 ? (i0) : lane == 1 ? (i1) : lane == 2 ? (i2) : lane == 3 ? (i3) : -1; \
    const uint32_t wp_ = lane == 0 ? (p0) : lane == 1 ? (p1) : lane == 2 ? (p2) : (p3);       \
    if (wi_ >= 0) ok_ = dwait_bar(&bars[wi_], wp_, p.dbg, (uint32_t)wi_, (uint32_t)st_dbg);    \
    if (!__all_sync(0xffffffffu, ok_)) goto bail;                                             \
  } while (0)
#define DPROBE4(var, i0, p0, i1, p1, i2, p2, i3, p3)                                          \
  bool var = true;                                                                            \
  {                                                                                           \
    const int wi_ = lane == 0 ? (i0) : lane == 1 ? (i1) : lane == 2 ? (i2) : lane == 3 ? (i3) : -1; \
    const uint32_t wp_ = lane == 0 ? (p0) : lane == 1 ? (p1) : lane == 2 ? (p2) : (p3);       \
    if (wi_ >= 0) var = mbar_try_wait(&bars[wi_], wp_);                                       \
  }

      uint32_t s = 0, sphase = 0;   // W1 stage ring position / phase
      uint32_t gg = 0;              // L0 groups issued so far (acc0 / W0 ring use count)
      uint32_t hg = 0;              // groups consumed by L1 so far (H0 buffer use count)
      uint32_t npass = 0, w2n = 0;
      const int NG = p.KC / 2;      // groups per pass
      for (int u = 0; u < total_ssteps; ++u) {
        const int sidx = u & 1;                                   // stream of this stream-step
        const uint32_t ztm = tmem + p.tm_z + sidx * zcols;        // A of L0: this stream's Z columns
        st_dbg = (u << 8) | 0xff;
        TRACE_ON(lane == 0 && (u >> 1) >= p.trace_t0 && (u >> 1) < p.trace_t1);
        TRACE(1, 0x1000 | sidx);
        // stream-step prologue: Z of this stream ready, W0 tile of the first group, acc0 drained
        DWAITW4(D_ZREADY + sidx, (u >> 1) & 1, D_W0FULL + (int)(gg & 1), (gg >> 1) & 1,
                gg > 0 ? D_ACC0FREE : -1, (gg - 1) & 1, -1, 0);
        tc_fence_after();
        if (elect_one()) {
          for (int j = 0; j < k0steps; ++j)
            umma_ts(acc0, ztm + j * 8, w0desc0 + (gg & 1) * w0slot + j * w0_kstep, idesc0, j > 0);
          umma_commit(&bars[D_ACC0FULL + (gg & 1)]);
          umma_commit(&bars[D_START + sidx]);
        }
        __syncwarp();
        ++gg;
        TRACE(1, 0x1001);
        DWAITW4(D_FULL + (int)s, sphase, D_H0FULL + 0, hg & 1,
                NGSp > 1 ? D_ACC0FREE : -1, (gg - 1) & 1, NGSp > 1 ? D_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
        int gp = 0, ncl = 0;
        for (int G = 0; G < NGSp; ++G) {
          st_dbg = (u << 8) | (2 * G);
          TRACE(1, 0x100 | G);
          const bool next_l0 = (G + 1 < NGSp);
          const bool first_of_pass = (gp == 0), last_of_pass = (gp == NG - 1);
          uint32_t s1 = s + 1, sphase1 = sphase;
          if (s1 == NSTAGE) { s1 = 0; sphase1 ^= 1; }
          uint32_t s2 = s1 + 1, sphase2 = sphase1;
          if (s2 == NSTAGE) { s2 = 0; sphase2 ^= 1; }
          tc_fence_after();
          if (elect_one()) {
            if (next_l0) {
              for (int j = 0; j < k0steps; ++j)
                umma_ts(acc0, ztm + j * 8, w0desc0 + (gg & 1) * w0slot + j * w0_kstep, idesc0, j > 0);
              umma_commit(&bars[D_ACC0FULL + (gg & 1)]);
              if (G + 2 == NGSp) umma_commit(&bars[D_ZFREE]);   // that was the stream-step's last L0
            }
            const uint32_t at = tmem + p.tm_h0;
            const uint64_t bd = stdesc0 + s * ststride;
            umma_ts(acc1, at, bd, idesc1, first_of_pass ? 0u : 1u);
            umma_ts(acc1, at + 8, bd + 2, idesc1, 1);
            umma_ts(acc1, at + 16, bd + 4, idesc1, 1);
            umma_ts(acc1, at + 24, bd + 6, idesc1, 1);
            umma_commit(&bars[D_EMPTY + s]);
          }
          __syncwarp();
          if (next_l0) ++gg;
          {
            DPROBE4(rb, D_FULL + (int)s1, sphase1, D_H0FULL + 1, hg & 1, -1, 0, -1, 0);
            if (!__all_sync(0xffffffffu, rb)) DWAITW4(D_FULL + (int)s1, sphase1, D_H0FULL + 1, hg & 1, -1, 0, -1, 0);
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t at = tmem + p.tm_h0 + 32;
            const uint64_t bd = stdesc0 + s1 * ststride;
            umma_ts(acc1, at, bd, idesc1, 1);
            umma_ts(acc1, at + 8, bd + 2, idesc1, 1);
          }
          __syncwarp();
          const bool need_l0_2 = (G + 2 < NGSp);
          const bool has_next = (G + 1 < NGSp);
          DPROBE4(rn, has_next ? D_FULL + (int)s2 : -1, sphase2, has_next ? D_H0FULL + 0 : -1, (hg + 1) & 1,
                  need_l0_2 ? D_ACC0FREE : -1, (gg - 1) & 1,
                  need_l0_2 ? D_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
          if (elect_one()) {
            const uint32_t at = tmem + p.tm_h0 + 32;
            const uint64_t bd = stdesc0 + s1 * ststride;
            umma_ts(acc1, at + 16, bd + 4, idesc1, 1);
            umma_ts(acc1, at + 24, bd + 6, idesc1, 1);
            umma_commit(&bars[D_EMPTY + s1]);
            if (last_of_pass) umma_commit(&bars[D_ACC1FULL]);
          }
          __syncwarp();
          if (last_of_pass) {
            // L2 of this pass: acc2 (+)= H1 (in place in acc1's columns) * W2 chunk.  The first L2
            // of a stream-step overwrites acc2: the OTHER stream's epilogue must have read it.
            TRACE(1, 0x2000 | ncl);
            DWAITW4(D_H1FULL + 0, npass & 1, D_W2FULL, w2n & 1,
                    (ncl == 0 && u > 0) ? D_ACC2FREE : -1, (u - 1) & 1, -1, 0);
            TRACE(1, 0x2100 | ncl);
            ++w2n;
#pragma unroll 1
            for (int sub = 0; sub < nsl; ++sub) {
              if (sub > 0) DWAITW4(D_H1FULL + sub, npass & 1, -1, 0, -1, 0, -1, 0);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t at = acc1 + sub * 64;
                const uint64_t bd = w2desc + sub * w2_sub;
#pragma unroll
                for (int j = 0; j < 4; ++j) umma_ts(acc2, at + 8 * j, bd + 2 * j, idesc2, (ncl | sub | j) != 0);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(&bars[D_W2EMPTY]);
            __syncwarp();
            TRACE(1, 0x2200 | ncl);
            ++npass; ++ncl; gp = 0;
          } else {
            ++gp;
          }
          if (G + 1 < NGSp && !__all_sync(0xffffffffu, rn)) {
            DWAITW4(D_FULL + (int)s2, sphase2, D_H0FULL + 0, (hg + 1) & 1,
                    need_l0_2 ? D_ACC0FREE : -1, (gg - 1) & 1,
                    need_l0_2 ? D_W0FULL + (int)(gg & 1) : -1, (gg >> 1) & 1);
          }
          ++hg;
          s = s2; sphase = sphase2;
        }
        if (elect_one()) umma_commit(&bars[D_ACC2FULL]);
        __syncwarp();
        TRACE(1, 0x1002);
      }
#undef DWAITW4
#undef DPROBE4
    }
  } else {
    setmaxnreg_inc_hi();
    if (total_ssteps > 0) {
      // =========================== epilogue / compute warpgroups ===========================
      const int grp = (warp - 4) >> 2;                       // stream served by this warpgroup
      const int e = tid & (EPI_THREADS - 1);                 // 0..127 inside the group
      const int wq = warp & 3;                               // TMEM lane quarter of this warp
      const int r = wq * 32 + lane;                          // row inside the tile
      const int bar_id = 1 + grp;                            // named barrier of the group
      const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
      const int S = p.S, A = p.A, K = p.K;
      const float* sB1 = sBias;
      const float* sB2 = sBias + NCp * N1;
      const float* inMean = sNorm;
      const float* inRstd = sNorm + p.SA;
      const float* dMean = sNorm + 2 * p.SA;
      const float* dStd = sNorm + 2 * p.SA + S;
      float* scrA = reinterpret_cast<float*>(smem + dp.off_scr[grp]);
      float* sHid = reinterpret_cast<float*>(smem + dp.off_hid[grp]);     // [2][pol_rows][33]
      int* sList = reinterpret_cast<int*>(smem + dp.off_list[grp]);        // [128] + sCnt[4]
      int* sCnt = sList + TILE_M;
      const uint32_t ztm = tmem + lane_base + p.tm_z + grp * zcols;
      const int RS = p.rec_stride, SPs = p.slot_stride;

      float x[SMAX], a_raw[AMAX];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) x[s] = 0.f;
      int ts = 0, nreset = 0;
      int ps = 0;        // pair-steps done by this CTA
      uint32_t xn_cnt = 0;
      bool have_action = false;
      int row = 0;
      bool valid = false;

      for (int si = 0; si < MAX_SEG; ++si) {
        const int4 sg = segs[si];
        if (sg.x < 0) continue;
        const int pair = sg.x, t0 = sg.y, t1 = sg.z;
        const int tile = 2 * pair + grp;
        const bool tile_ok = tile < p.n_tiles;
        row = tile * TILE_M + r;
        valid = tile_ok && row < p.B;
        have_action = false;

        if (sg.w) {   // the pair's state was published by the slot that ran its head
          if (e == 0 && !dwait_ge(&p.tile_flag[pair], 2u, p.dbg, 100u, (uint32_t)pair)) abort_smem = 1;
          named_bar_sync(bar_id, EPI_THREADS);
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
          ts = tile_ok ? __ldcg(&p.row_ts[tile * TILE_M + r]) : 0;
          nreset = tile_ok ? __ldcg(&p.row_nreset[tile * TILE_M + r]) : 0;
        }

        for (int t = t0; t < t1; ++t, ++ps) {
          const uint32_t u = 2u * static_cast<uint32_t>(ps) + grp;   // ordinal of this stream-step
          st_dbg = (t << 8) | 0xf0;
          TRACE_ON(e == 0 && ps >= p.trace_t0 && ps < p.trace_t1);
          TRACE(2 + grp, 0x1000);
          // ================= begin step: action + Z operand =================
          if (!have_action) {
            // first step of a segment: every CTA runs the policy for all rows of the tile
            float a_mean[AMAX];
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) scrA[s * TILE_M + r] = x[s];
            for (int l = 0; l < p.n_pol_layers; ++l) {
              const PolicyLayer L = p.pl[l];
              const bool last = (l == p.n_pol_layers - 1);
              if (!last) {
                float acc[HPB];
                dense_layer<HPB>(scrA, L.nin, sPolS + L.w_off, HPB, sPolS + L.b_off, acc, r);
#pragma unroll
                for (int j = 0; j < HPB; ++j) scrA[j * TILE_M + r] = fast_tanh(acc[j]);
              } else {
                float acc[AMAX];
                dense_layer<AMAX>(scrA, L.nin, sPolS + L.w_off, AMAX, sPolS + L.b_off, acc, r);
#pragma unroll
                for (int i = 0; i < AMAX; ++i) a_mean[i] = p.pol_out_tanh ? tanhf(acc[i]) : acc[i];
              }
            }
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
                  float n4[4] = {0.f, 0.f, 0.f, 0.f};
                  if (blk * 4 < A)
                    philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(t),
                                   static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EPS + blk, n4);
#pragma unroll
                  for (int q = 0; q < 4; ++q) ep[blk * 4 + q] = n4[q];
                }
              }
#pragma unroll
              for (int i = 0; i < AMAX; ++i) {
                const float ls = fmaxf(sPolS[p.pol_logstd_off + i], -13.815510557964274f);
                a_raw[i] = (i < A) ? __fadd_rn(__fmul_rn(ep[i], expf(ls)), a_mean[i]) : 0.f;
              }
            }
            if (k == 0 && c == 0 && valid) {
              const size_t o = static_cast<size_t>(t) * p.B + row;
#pragma unroll
              for (int i = 0; i < AMAX; ++i)
                if (i < A) {
                  if (p.act) p.act[o * A + i] = a_raw[i];
                  if (p.mean) p.mean[o * A + i] = a_mean[i];
                }
            }
          }
          // z = (concat(x, clip(a)) - in_mean) * (1 / in_std), drop leading cols  (training.py:228,146-154)
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s < S) scrA[s * TILE_M + r] = __fmul_rn(__fsub_rn(x[s], inMean[s]), inRstd[s]);
#pragma unroll
          for (int i = 0; i < AMAX; ++i)
            if (i < A) {
              const float uu = fminf(fmaxf(a_raw[i], -1.f), 1.f);   // env_helpers.py:599
              scrA[(S + i) * TILE_M + r] = __fmul_rn(__fsub_rn(uu, inMean[S + i]), inRstd[S + i]);
            }
          if (dp.z_shared && u > 0) DWAITB(D_ZFREE, (u - 1) & 1);   // the other stream's L0s are done with the slot
          for (int cc = 0; cc < p.K0 / 16; ++cc) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int f0 = 16 * cc + 2 * i, f1 = f0 + 1;
              const float z0 = f0 < p.Din ? scrA[(f0 + p.drop) * TILE_M + r] : (f0 < p.Din + 2 ? 1.f : 0.f);
              const float z1 = f1 < p.Din ? scrA[(f1 + p.drop) * TILE_M + r] : (f1 < p.Din + 2 ? 1.f : 0.f);
              pk[i] = pack_bf16x2(z0, z1);
            }
            tmem_st8(ztm + cc * 8, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars[D_ZREADY + grp]);
          TRACE(2 + grp, 0x1001);

          // ================= this stream's MMA phase =================
          DWAITB(D_START + grp, (u >> 1) & 1);
          TRACE(2 + grp, 0x1002);
          {
            const int NG = p.KC / 2;
            int egp = 0, enc = 0;
            for (int G = 0; G < NGSp; ++G) {
              const uint32_t gg = u * static_cast<uint32_t>(NGSp) + G;   // global L0 group index
              st_dbg = (t << 8) | (2 * G);
              {
                uint32_t v0[32], v1[32], pk[32];
                DWAITB(D_ACC0FULL + (gg & 1), (gg >> 1) & 1);
                TRACE(2 + grp, 0x100 | G);
                tc_fence_after();
                tmem_ld32(tmem + lane_base + p.tm_acc0, v0);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 32, v1);
                tmem_ld_wait();
                relu_pack<false>(v0, v1, nullptr, pk);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 64, v0);
                tmem_ld32(tmem + lane_base + p.tm_acc0 + 96, v1);
                // H0 buffer 0 is free once L1 of chunk 2*gg-2 completed (its stage's EMPTY barrier)
                if (gg >= 1) DWAITB(D_EMPTY + ((2 * gg - 2) & 3), ((2 * gg - 2) >> 2) & 1);
                tmem_st32(tmem + lane_base + p.tm_h0, pk);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[D_ACC0FREE]);        // next group's L0 may overwrite acc0
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[D_H0FULL + 0]);
                relu_pack<false>(v0, v1, nullptr, pk);
                if (gg >= 1) DWAITB(D_EMPTY + ((2 * gg - 1) & 3), ((2 * gg - 1) >> 2) & 1);
                tmem_st32(tmem + lane_base + p.tm_h0 + 32, pk);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars[D_H0FULL + 1]);
                TRACE(2 + grp, 0x200 | G);
              }
              int drain_nc = -1;
              if (G == NGSp - 1) drain_nc = NCp - 1;
              else if (egp == 0 && enc > 0) drain_nc = enc - 1;
              if (++egp == NG) { egp = 0; ++enc; }
              if (drain_nc >= 0) {
                const uint32_t a1n = u * static_cast<uint32_t>(NCp) + drain_nc;   // global pass index
                DWAITB(D_ACC1FULL, a1n & 1);
                TRACE(2 + grp, 0x300 | drain_nc);
                tc_fence_after();
                uint32_t va0[32], va1[32], pk[32];
                const uint32_t a1 = tmem + lane_base + TM_ACC1;
                const float* bb = sB1 + drain_nc * N1;
                constexpr int nsl = N1 / 64;
#pragma unroll 1
                for (int sub = 0; sub < nsl; ++sub) {
                  tmem_ld32(a1 + sub * 64, va0);
                  tmem_ld32(a1 + sub * 64 + 32, va1);
                  tmem_ld_wait();
                  relu_pack<true>(va0, va1, bb + sub * 64, pk);
                  tmem_st32(a1 + sub * 64, pk);      // in place: 64 fp32 columns -> 32 columns of bf16 pairs
                  tmem_st_wait();
                  tc_fence_before();
                  mbar_arrive(&bars[D_H1FULL + sub]);
                }
                TRACE(2 + grp, 0x400 | drain_nc);
              }
            }
          }

          // ================= finish step (overlaps the OTHER stream's MMA phase) =================
          st_dbg = (t << 8) | 0xf1;
          float cand[SMAX];
          {
            TRACE(2 + grp, 0x1003);
            DWAITB(D_ACC2FULL, u & 1);
            TRACE(2 + grp, 0x1004);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + p.tm_acc2, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[D_ACC2FREE]);
            if (cs == 2) {
              // exchange the layer-2 partial sums of the two column halves through L2
              float* pb = dp.pbuf + (static_cast<size_t>(slot * 2 + grp) * 2 + (xn_cnt & 1)) * dp.pbuf_stride +
                          static_cast<size_t>(k) * 2 * S * TILE_M;
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) pb[(c * S + s) * TILE_M + r] = __uint_as_float(v[s]);
              __syncwarp();
              unsigned* pc = dp.pctr + (slot * 2 + grp) * K + k;
              int okw = 1;
              if (lane == 0) {
                red_release_gpu_add(pc, 1u);
                okw = dwait_ge(pc, 8u * (xn_cnt + 1), p.dbg, 102u, (uint32_t)st_dbg) ? 1 : 0;
              }
              okw = __shfl_sync(0xffffffffu, okw, 0);
              if (!okw) goto bail;
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) {
                  const float other = __ldcg(&pb[((c ^ 1) * S + s) * TILE_M + r]);
                  const float mine_v = __uint_as_float(v[s]);
                  v[s] = __float_as_uint(c == 0 ? __fadd_rn(mine_v, other) : __fadd_rn(other, mine_v));
                }
            }
            TRACE(2 + grp, 0x1005);
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
              if (s < S) {
                const float o = __fadd_rn(__uint_as_float(v[s]), sB2[s]);
                cand[s] = __fadd_rn(__fadd_rn(dMean[s], __fmul_rn(dStd[s], o)), x[s]);   // training.py:257
              } else {
                cand[s] = 0.f;
              }
            }
          }
          {
            // ============ row-ownership exchange (see rollout_kernel.cuh) ============
            float* rec = dp.rbuf + (static_cast<size_t>(slot * 2 + grp) * 2 + (xn_cnt & 1)) * dp.rbuf_stride;
            int own_idx;
            if (p.model_idx != nullptr)
              own_idx = valid ? p.model_idx[static_cast<size_t>(t) * p.B + row] : 0;
            else if (p.sam_mode == METRPO_SAM_STEP_RAND)
              own_idx = philox_index(p.seed, p.offset + static_cast<unsigned long long>(t),
                                     static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_IDX, K);
            else
              own_idx = philox_index(p.seed, static_cast<uint64_t>(static_cast<uint32_t>(nreset)),
                                     static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EIDX, K);
            own_idx = min(max(own_idx, 0), K - 1);
            const bool own = valid && own_idx == k;
            const bool want_pol = (t + 1 < t1);
            // ---- compact the owned rows of the tile; the halves take alternate entries ----
            const unsigned om = __ballot_sync(0xffffffffu, own);
            if (lane == 0) sCnt[wq] = __popc(om);
            named_bar_sync(bar_id, EPI_THREADS);
            int off = 0, n_own = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int cnt = sCnt[w];
              if (w < wq) off += cnt;
              n_own += cnt;
            }
            const int jglob = off + __popc(om & ((1u << lane) - 1u));
            const bool mine = own && (jglob % cs) == c;
            const int n_mine = (n_own - c + cs - 1) / cs;
            float own_reward = 0.f;
            bool own_dn = false;
            if (mine) {
              float uu[AMAX];
#pragma unroll
              for (int i = 0; i < AMAX; ++i) uu[i] = fminf(fmaxf(a_raw[i], -1.f), 1.f);
              own_reward = -env_cost<SMAX, AMAX>(p.env_id, S, A, cand, uu);                  // env_helpers.py:601
              own_dn = env_is_done<SMAX>(p.env_id, S, cand) || (ts + 1 >= p.T_max);          // :603-604
              if (own_dn) {   // :605-606 -> reset(dones)
                const float* src = p.reset_pool + static_cast<size_t>((static_cast<long long>(nreset) * p.B + row) % p.R) * S;
#pragma unroll
                for (int s = 0; s < SMAX; ++s)
                  if (s < S) cand[s] = src[s];
              }
              float4* rq = reinterpret_cast<float4*>(rec + r * RS);
              rq[AMAX / 4] = make_float4(own_dn ? 1.f : 0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int q = 0; q < SMAX / 4; ++q)
                if (4 * q < S) rq[AMAX / 4 + 1 + q] = make_float4(cand[4 * q], cand[4 * q + 1], cand[4 * q + 2], cand[4 * q + 3]);
              const int j = jglob / cs;     // index in this half's list
              sList[j] = r;
              float* in = scrA + j * SPs;
#pragma unroll
              for (int s = 0; s < SMAX; ++s)
                if (s < S) in[s] = cand[s];
              if (want_pol && !p.determ) {   // policy noise of step t+1 for this row
                if (p.eps != nullptr) {
#pragma unroll
                  for (int i = 0; i < AMAX; ++i)
                    if (i < A) in[S + i] = p.eps[(static_cast<size_t>(t + 1) * p.B + row) * A + i];
                } else {
#pragma unroll
                  for (int blk = 0; blk < AMAX / 4; ++blk) {
                    if (blk * 4 < A) {
                      float n4[4];
                      philox_normal4(p.seed, p.offset + static_cast<unsigned long long>(t + 1),
                                     static_cast<uint32_t>(row + p.row_offset), PHILOX_STREAM_EPS + blk, n4);
#pragma unroll
                      for (int q = 0; q < 4; ++q)
                        if (blk * 4 + q < A) in[S + blk * 4 + q] = n4[q];
                    }
                  }
                }
              }
            }
            named_bar_sync(bar_id, EPI_THREADS);
            TRACE(2 + grp, 0x1006);
            // ---- policy of step t+1 for this half's owned rows: 2 threads per row ----
            if (want_pol) {
              const int part = e & 1, jl = e >> 1;
              const int nl = p.n_pol_layers;
              const int PR = dp.pol_rows;
              // lanes of this warp that take part (a partial warp when pol_rows = 8): the two threads of a
              // row exchange their halves of a hidden layer through sHid under this mask
              const unsigned pmask = __ballot_sync(0xffffffffu, jl < PR);
              for (int base = 0; base < n_mine; base += PR) {
                if (jl < PR) {
                  const int j = base + jl;
                  const bool actv = j < n_mine;
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
                      const float2 xx = make_float2(xi, xi);
#pragma unroll
                      for (int q = 0; q < 4; ++q) {
                        const float4 w = w4[q];
                        const float2 r0 = __ffma2_rn(xx, make_float2(w.x, w.y), make_float2(acc[4 * q], acc[4 * q + 1]));
                        const float2 r1 = __ffma2_rn(xx, make_float2(w.z, w.w), make_float2(acc[4 * q + 2], acc[4 * q + 3]));
                        acc[4 * q] = r0.x; acc[4 * q + 1] = r0.y; acc[4 * q + 2] = r1.x; acc[4 * q + 3] = r1.y;
                      }
                    }
                    float* out = sHid + (l & 1) * (PR * 33) + jl * 33 + 16 * part;
#pragma unroll
                    for (int q = 0; q < 16; ++q) out[q] = fast_tanh(acc[q]);
                    __syncwarp(pmask);
                    cur = sHid + (l & 1) * (PR * 33) + jl * 33;
                  }
                  const PolicyLayer& L = p.pl[nl - 1];
                  float mo[AMAX / 2];
#pragma unroll
                  for (int q = 0; q < AMAX / 2; ++q) mo[q] = sPolS[L.b_off + (AMAX / 2) * part + q];
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
                    for (int q = 0; q < AMAX / 2; ++q) mo[q] = tanhf(mo[q]);
                  }
                  if (actv) {
                    const int rowl = sList[j];
                    const float* in = scrA + j * SPs;
                    const size_t o = static_cast<size_t>(t + 1) * p.B + (tile * TILE_M + rowl);
#pragma unroll
                    for (int q = 0; q < AMAX / 2; ++q) {
                      const int a_i = (AMAX / 2) * part + q;
                      if (a_i < A) {
                        const float mu = mo[q];
                        float raw = mu;
                        if (!p.determ) {   // a = eps * exp(log_std) + mean   (rllab get_actions)
                          const float ls = fmaxf(sPolS[p.pol_logstd_off + a_i], -13.815510557964274f);
                          raw = __fadd_rn(__fmul_rn(in[S + a_i], expf(ls)), mu);
                        }
                        if (p.act) p.act[o * A + a_i] = raw;
                        if (p.mean) p.mean[o * A + a_i] = mu;
                        rec[rowl * RS + a_i] = raw;
                      }
                    }
                  }
                  __syncwarp(pmask);
                }
              }
            }
            // ---- publish / meet the gang: one release + one acquire poll per warp ----
            named_bar_sync(bar_id, EPI_THREADS);   // policy threads of other warps wrote this warp's rows' records
            TRACE(2 + grp, 0x1007);
            int okw = 1;
            unsigned* rc = dp.rctr + (slot * 2 + grp);
            if (lane == 0) red_release_gpu_add(rc, 1u);
            if (mine) {   // the owner's part of the trajectory record, off the gang's critical path
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
              okw = dwait_ge(rc, 4u * static_cast<unsigned>(K * cs) * (xn_cnt + 1), p.dbg, 101u, (uint32_t)st_dbg) ? 1 : 0;
            okw = __shfl_sync(0xffffffffu, okw, 0);
            if (!okw) goto bail;
            TRACE(2 + grp, 0x1008);
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
            TRACE(2 + grp, 0x1009);
            have_action = want_pol;
          }
        }  // t

        // ---- segment end: publish the tile's state (the two groups of the (k = 0, c = 0) CTA) ----
        if (k == 0 && c == 0) {
          if (valid) {
#pragma unroll
            for (int s = 0; s < SMAX; ++s)
              if (s < S) {
                p.row_state[static_cast<size_t>(row) * S + s] = x[s];
                if (t1 == p.n_steps && p.final_states) p.final_states[static_cast<size_t>(row) * S + s] = x[s];
              }
          }
          if (tile_ok) {
            p.row_ts[tile * TILE_M + r] = ts;
            p.row_nreset[tile * TILE_M + r] = nreset;
          }
          named_bar_sync(bar_id, EPI_THREADS);
          if (e == 0) red_release_gpu_add(&p.tile_flag[pair], 1u);
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
