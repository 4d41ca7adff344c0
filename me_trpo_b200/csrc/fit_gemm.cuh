// Batched TF32 GEMM on tcgen05 for the ensemble fit (sm_100a): the three dense products of the
// 1024-wide hidden layer (forward, dgrad, wgrad; model_based_rl.py:154-183 as TF builds them from
// tf.matmul in training.py:207-208 and its gradients), one launch for all K models.
//
//   C[m, n] = epilogue( sum_k A(m, k) * B(n, k) ),   fp32 in / out, TF32 operands, fp32 accumulate
//
// Either operand may be stored K-major (the reduction index is the contiguous one) or MN-major
// (the M / N index is contiguous), so all transposes the backward pass needs are read in place:
//     forward   H1 = relu(H0 W1 + b1)      A = H0 [rows][H]  K-major,  B = W1 [k][n]   MN-major
//     wgrad     dW1 = H0^T dH1             A = H0 [k = row][m] MN-major, B = dH1 [k][n] MN-major
//     dgrad     dH0 = (dH1 W1^T) * (H0>0)  A = dH1 [rows][H] K-major,  B = W1 [n][k]   K-major
//
// Structure (192 threads; big problems: one CTA per 256 x 256 output tile = two M = 128 accumulators of
// 256 TMEM columns each sharing every B tile, 3-stage ring of 64 KB; N <= 128 problems: 128 x 128 tiles,
// 4 stages): warp 0 = TMA producer (tiled tensor maps over the fp32 row-major arrays, 32-deep K
// blocks; out-of-range rows / the reduction tail are zero-filled by the TMA unit, so M, N, K need no
// padding), warp 1 = MMA issuer (one elected lane, tcgen05.mma.kind::tf32, K = 8 per instruction,
// operands straight from the swizzled tiles through shared-memory descriptors), warps 2-5 = epilogue:
// tcgen05.ld -> registers (thread = output row) -> bias + ReLU or the ReLU mask of the backward pass
// (mask tile prefetched by TMA into the idle stage ring) -> swizzled shared-memory tile -> TMA store
// (clipped at M, N by the tensor map); the masked variant also emits per-32-row-slab column sums (the
// bias gradients), summed in a fixed order by the caller.
//
// Store clipping: the TMA unit clips rows at M exactly and columns at 16-byte granularity, i.e. up to the next
// multiple of 4 columns past N may be written (with the epilogue of a zero accumulator: 0, or relu(0) = 0);
// the fit's padded arrays (row pitch a multiple of 32 floats, pads kept at zero) rely on exactly that.
//
// Roofline: fp32 operands make this L2-bandwidth bound at 128-row tiles (48 KB per 32-deep K block
// for 512 tensor-pipe cycles); the 256 x 256 tile halves the bytes per FLOP and leaves 80 CTAs for
// the fit's shapes (K = 5 models x 4 x 4 tiles), one wave.
#pragma once
#include <cuda.h>
#include "umma.cuh"

namespace metrpo {

constexpr int GM_BK = 32;
constexpr int GM_EPI_WARPS = 8;                     // two per TMEM lane quarter (even / odd chunks)
constexpr int GM_THREADS = 64 + 32 * GM_EPI_WARPS;
constexpr int GM_CHUNK_BYTES = 128 * 128;            // one 128-row x 32-column fp32 epilogue chunk
template <int MT, int BN> struct GemmCfg {
  static constexpr int BM = 128 * MT;
  static constexpr int A_BYTES = BM * GM_BK * 4;     // 16 / 32 KB
  static constexpr int B_BYTES = BN * GM_BK * 4;     // 16 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE_BYTES > 49152) ? 3 : 6;   // 192 KB either way
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int NCH = BN / 32;                // epilogue chunks per 128-row sub-tile
  static constexpr int NBUF = RING_BYTES / GM_CHUNK_BYTES;   // 16 KB staging buffers the idle ring provides
  static constexpr int NHEAD = (MT == 1) ? 0 : (NBUF - NCH < NCH ? NBUF - NCH : NCH);   // sub-tile 1 chunks with a buffer of their own
  static constexpr int TMEM_COLS = MT * BN;
  static constexpr int SMEM_BYTES = RING_BYTES + BN * 4 + 256 + 1024;
  static_assert(NCH * GM_CHUNK_BYTES <= RING_BYTES, "epilogue staging must fit the stage ring");
};

enum { GEMM_EPI_PLAIN = 0, GEMM_EPI_BIAS_RELU = 1, GEMM_EPI_MASK = 2 };

struct GemmParams {
  int M, N, Kd;                 // per-model problem size (store guards; operand extents live in the tensor maps)
  int a_mn, b_mn;               // 1: operand stored MN-major ([k][m] / [k][n] row-major)
  int epi;
  int round_out;                // 1: round C to TF32-nearest (it is the operand of a later GEMM)
  int splits;                   // split-K: blockIdx.z = model * splits + split; split s reduces its share of the K
                                //   blocks into C + s * strideSplit (partials, summed by the caller; GEMM_EPI_PLAIN only)
  long long strideSplit;
  int trans_store;              // 1: store C transposed, C[n * ldc + m], from registers (GEMM_EPI_PLAIN only)
  float* C; long long ldc, strideC;                     // ldc % 4 == 0 unless trans_store
  const float* bias; long long strideBias;              // GEMM_EPI_BIAS_RELU: bias[n] per model
  unsigned long long* dbg;      // dev: %globaltimer stamps of CTA (0,0,0): start | set-up done | accumulators full | stores issued | stores done
  float* colsum;                // GEMM_EPI_MASK, optional: per (model, 32-row slab) column sums of C,
                                //   [model][gridDim.y * MT * 4][N] (summed in a fixed order by the caller)
};

// kind::tf32 instruction descriptor with operand major-ness (bit 15: A MN-major, bit 16: B MN-major)
__host__ __device__ constexpr uint32_t idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return idesc_tf32_f32(M, N) | (static_cast<uint32_t>(a_mn & 1) << 15) | (static_cast<uint32_t>(b_mn & 1) << 16);
}
// MN-major TF32 operand tile laid out [atom = 32 mn][k rows of 128 B].  For 32-bit MN-major operands the
// tensor core only accepts the "128 B swizzle with 32 B atomicity" layout (descriptor layout type 1,
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side): 32 B chunk c of k-row r is stored at chunk
// c ^ (r & 3), swizzle groups are 4 k-rows (512 B).  One MMA (K = 8) reads the 8 k-rows starting at
// `saddr`; LBO = bytes between consecutive 32-wide MN atoms, SBO = bytes between 4-row k groups.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;            // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most N of this thread's most recent bulk groups still read their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// Tensor maps (built by fit_make_tmap below), always rank 4:
//   K-major operand  [MN rows][Kd] :  dims {Kd, MN, 1, models},        box {32, tile_mn, 1, 1}
//   MN-major operand [Kd rows][MN] :  dims {32, Kd, MN / 32, models},  box {32, 32, tile_mn / 32, 1}
//   C (store, 32 x 32 boxes) and the mask (load, 32 x 128 boxes): row-major like a K-major operand
template <int MT, int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
fit_gemm_tf32_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux) {
  using Cfg = GemmCfg<MT, BN>;
  extern __shared__ uint8_t gm_smem_raw[];
  uint8_t* smem = gm_smem_raw + ((1024u - (smem_u32(gm_smem_raw) & 1023u)) & 1023u);
  uint8_t* sStage = smem;
  float* sBias = reinterpret_cast<float*>(smem + Cfg::RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + BN * 4);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* accfull = bars + 2 * Cfg::STAGES;
  uint64_t* auxfull = accfull + 1;             // [3] mask tiles landed: sub-tile 0 | head of sub-tile 1 | its tail
  uint64_t* subdone = accfull + 4;             // all epilogue warps are done with sub-tile 0's buffers
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool dbg_on = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg_on && tid == 64) p.dbg[0] = globaltimer_ns();
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * Cfg::BM;
  const int model = blockIdx.z / p.splits, split = blockIdx.z - model * p.splits;
  const int nkb_all = (p.Kd + GM_BK - 1) / GM_BK;
  const int kb_lo = static_cast<int>(static_cast<long long>(nkb_all) * split / p.splits);
  const int nkb = static_cast<int>(static_cast<long long>(nkb_all) * (split + 1) / p.splits) - kb_lo;   // >= 1 (host: splits <= blocks)
  const int nch = min(Cfg::NCH, (p.N - n0 + 31) / 32);   // chunks of this tile that hold real columns

  if (tid == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accfull, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&auxfull[i], 1);
    mbar_init(subdone, GM_EPI_WARPS);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!p.trans_store) tma_prefetch_desc(&tmC);
    if (p.epi == GEMM_EPI_MASK) tma_prefetch_desc(&tmAux);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (dbg_on && tid == 64) p.dbg[1] = globaltimer_ns();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sA = sStage + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        const int k0 = (kb_lo + kb) * GM_BK;
        if (p.a_mn) tma_load_4d(sA, &tmA, 0, k0, m0 / 32, model, &full[s]);
        else        tma_load_4d(sA, &tmA, k0, m0, 0, model, &full[s]);
        if (p.b_mn) tma_load_4d(sB, &tmB, 0, k0, n0 / 32, model, &full[s]);
        else        tma_load_4d(sB, &tmB, k0, n0, 0, model, &full[s]);
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
      if (p.epi == GEMM_EPI_MASK) {
        // mask tiles of the epilogue go into the (now idle) stage ring: sub-tile 0 -> buffers 0 .. NCH-1,
        // the first NHEAD chunks of sub-tile 1 -> buffers NCH .., its remaining chunks reuse sub-tile
        // 0's FIRST buffers (whose stores drain first) once all epilogue warps are done with them
        mbar_wait(accfull, 0);
        mbar_arrive_expect_tx(&auxfull[0], nch * GM_CHUNK_BYTES);
        for (int ch = 0; ch < nch; ++ch)
          tma_load_4d(sStage + ch * GM_CHUNK_BYTES, &tmAux, n0 + ch * 32, m0, 0, model, &auxfull[0]);
        if (MT > 1) {
          const int nhead = min(nch, Cfg::NHEAD);
          if (nhead > 0) {
            mbar_arrive_expect_tx(&auxfull[1], nhead * GM_CHUNK_BYTES);
            for (int ch = 0; ch < nhead; ++ch)
              tma_load_4d(sStage + (Cfg::NCH + ch) * GM_CHUNK_BYTES, &tmAux, n0 + ch * 32, m0 + 128, 0, model, &auxfull[1]);
          }
          if (nch > nhead) {
            mbar_wait(subdone, 0);
            mbar_arrive_expect_tx(&auxfull[2], (nch - nhead) * GM_CHUNK_BYTES);
            for (int ch = nhead; ch < nch; ++ch)
              tma_load_4d(sStage + (ch - nhead) * GM_CHUNK_BYTES, &tmAux, n0 + ch * 32, m0 + 128, 0, model, &auxfull[2]);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32_major(128, BN, p.a_mn, p.b_mn);
    uint32_t s = 0, ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aaddr = smem_u32(sStage + s * Cfg::STAGE_BYTES);
        const uint32_t baddr = aaddr + Cfg::A_BYTES;
        // K-major: the four K = 8 steps of a 32-deep block are 32 B apart inside the 128 B swizzle
        // row; MN-major: 8 k-rows = 1024 B apart, MN atoms GM_BK * 128 B apart.  Either way the
        // second 128-row half of a 256-row A tile starts 16 KB further.
        const uint64_t ad = p.a_mn ? smem_desc_sw128_mn(aaddr, GM_BK * 128) : smem_desc_sw128(aaddr);
        const uint64_t bd = p.b_mn ? smem_desc_sw128_mn(baddr, GM_BK * 128) : smem_desc_sw128(baddr);
        const uint32_t astep = p.a_mn ? (1024 >> 4) : (32 >> 4);
        const uint32_t bstep = p.b_mn ? (1024 >> 4) : (32 >> 4);
#pragma unroll
        for (int ks = 0; ks < GM_BK / 8; ++ks) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            umma_ss_tf32(tmem + mt * BN, ad + ks * astep + mt * (16384 >> 4), bd + ks * bstep, idesc, (kb | ks) != 0);
        }
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(accfull);
      }
      __syncwarp();
      if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes (w % 4) * 32 .. + 31 = rows of that quarter of a sub-tile ----
    const int q = warp & 3;
    const int par = (warp - 2) >> 2;                 // this warp takes the chunks of this parity
    const float* biasm = p.bias ? p.bias + model * p.strideBias : nullptr;
    if (p.epi == GEMM_EPI_BIAS_RELU) {
      for (int i = tid - 64; i < BN; i += 32 * GM_EPI_WARPS) sBias[i] = (n0 + i < p.N) ? biasm[n0 + i] : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * GM_EPI_WARPS) : "memory");
    }
    mbar_wait(accfull, 0);
    tc_fence_after();
    if (dbg_on && tid == 64) p.dbg[2] = globaltimer_ns();
    const int rl = q * 32 + lane;                    // row inside the 128-row sub-tile
    const uint32_t sw = rl & 7;
    const uint32_t tq = tmem + (static_cast<uint32_t>(q * 32) << 16);
    if (p.trans_store) {
      // thread = row m, register j = column: C[n][m] straight from registers, coalesced over lanes
      float* Cm = p.C + model * p.strideC + split * p.strideSplit;
      for (int mt = 0; mt < MT; ++mt) {
        const int m = m0 + mt * 128 + rl;
        for (int ch = par; ch < nch; ch += 2) {
          uint32_t v[32];
          tmem_ld32(tq + mt * BN + ch * 32, v);
          tmem_ld_wait();
          if (m < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + ch * 32 + j < p.N) Cm[static_cast<size_t>(n0 + ch * 32 + j) * p.ldc + m] = __uint_as_float(v[j]);
          }
        }
      }
    } else {
      // chunk groups: (sub-tile 0, all chunks) | (sub-tile 1, chunks with a buffer of their own) |
      // (sub-tile 1, chunks that reuse sub-tile 0's buffers).  Per group: every chunk goes TMEM ->
      // registers -> (mask | bias + ReLU) -> swizzled staging tile; then ONE proxy fence and one TMA
      // store per chunk (this warp's 32-row slab), committed as one bulk group.
      const int nhead = min(nch, Cfg::NHEAD);
      const int ngroups = (MT == 1) ? 1 : 3;
      for (int g = 0; g < ngroups; ++g) {
        const int mt = g == 0 ? 0 : 1;
        const int c_lo = g == 2 ? nhead : 0;
        const int c_hi = g == 1 ? nhead : nch;
        if (c_lo >= c_hi) continue;
        const int boff = g == 1 ? Cfg::NCH : (g == 2 ? -nhead : 0);   // staging buffer of chunk ch = boff + ch
        if (p.epi == GEMM_EPI_MASK) mbar_wait(&auxfull[g], 0);
        else if (g == 2) {
          // the buffers reused here are those of this warp's FIRST stores (one bulk group per chunk): on
          // a full tile NHEAD / 2 of its NCH / 2 + NHEAD / 2 groups have to be done, the rest may be pending
          if (lane == 0) {
            if (nch == Cfg::NCH) bulk_wait_group_read<Cfg::NCH / 2>();
            else bulk_wait_group_read0();
          }
          __syncwarp();
        }

        const int mrow0 = m0 + mt * 128 + q * 32;
        const int c_first = c_lo + ((c_lo ^ par) & 1);   // first chunk of this warp's parity in the group
        uint32_t va[32], vb[32];
        if (c_first < c_hi) tmem_ld32(tq + mt * BN + c_first * 32, va);
        for (int ch = c_first; ch < c_hi; ch += 2) {
          tmem_ld_wait();
          const bool odd = ((ch - c_first) >> 1) & 1;
          // the next chunk's TMEM load is in flight while this one is transformed
          if (ch + 2 < c_hi) {
            if (odd) tmem_ld32(tq + mt * BN + (ch + 2) * 32, va);
            else     tmem_ld32(tq + mt * BN + (ch + 2) * 32, vb);
          }
          uint8_t* rowp = sStage + (boff + ch) * GM_CHUNK_BYTES + rl * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float x0 = __uint_as_float(odd ? vb[4 * c] : va[4 * c]), x1 = __uint_as_float(odd ? vb[4 * c + 1] : va[4 * c + 1]);
            float x2 = __uint_as_float(odd ? vb[4 * c + 2] : va[4 * c + 2]), x3 = __uint_as_float(odd ? vb[4 * c + 3] : va[4 * c + 3]);
            float4* slot = reinterpret_cast<float4*>(rowp + ((c ^ sw) << 4));
            if (p.epi == GEMM_EPI_MASK) {
              const float4 a = *slot;
              x0 = a.x > 0.f ? x0 : 0.f; x1 = a.y > 0.f ? x1 : 0.f; x2 = a.z > 0.f ? x2 : 0.f; x3 = a.w > 0.f ? x3 : 0.f;
            } else if (p.epi == GEMM_EPI_BIAS_RELU) {
              const float4 b = *reinterpret_cast<const float4*>(sBias + ch * 32 + 4 * c);
              x0 = fmaxf(x0 + b.x, 0.f); x1 = fmaxf(x1 + b.y, 0.f); x2 = fmaxf(x2 + b.z, 0.f); x3 = fmaxf(x3 + b.w, 0.f);
            }
            if (p.round_out) { x0 = round_tf32(x0); x1 = round_tf32(x1); x2 = round_tf32(x2); x3 = round_tf32(x3); }
            *slot = make_float4(x0, x1, x2, x3);
          }
        }
        __syncwarp();
        if (dbg_on && tid == 64 && g == 0) p.dbg[6] = globaltimer_ns();
        if (p.colsum) {   // lane j sums column j over the warp's 32 rows (rows past M hold zeros)
          float* csum = p.colsum + (static_cast<size_t>(model) * gridDim.y * MT * 4 + (blockIdx.y * MT + mt) * 4 + q) * p.N;
          for (int ch = c_first; ch < c_hi; ch += 2) {
            const uint8_t* slab = sStage + (boff + ch) * GM_CHUNK_BYTES + q * 4096;
            float cs0 = 0.f, cs1 = 0.f;
#pragma unroll 8
            for (int i = 0; i < 32; i += 2) {
              cs0 += *reinterpret_cast<const float*>(slab + i * 128 + ((((lane >> 2) ^ i) & 7) << 4) + (lane & 3) * 4);
              cs1 += *reinterpret_cast<const float*>(slab + (i + 1) * 128 + ((((lane >> 2) ^ (i + 1)) & 7) << 4) + (lane & 3) * 4);
            }
            const int n = n0 + ch * 32 + lane;
            if (n < p.N) csum[n] = cs0 + cs1;
          }
        }
        if (dbg_on && tid == 64 && g == 0) p.dbg[7] = globaltimer_ns();
        // staged slab -> global by TMA (32-row x 128 B boxes, clipped at M, N by the tensor map), one bulk
        // group per chunk.  (Plain 128-bit stores of full lines were measured too: 6.9 us instead of
        // 5.4 us per 256 x 256 tile -- either way the 80 CTAs write their 20 MB at the same time.)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          for (int ch = c_first; ch < c_hi; ch += 2) {
            tma_store_4d(&tmC, sStage + (boff + ch) * GM_CHUNK_BYTES + q * 4096, n0 + ch * 32, mrow0, split, model);
            bulk_commit_group();
          }
          if (g == 0 && p.epi == GEMM_EPI_MASK && MT > 1 && nch > nhead) {
            // the tail mask tiles of sub-tile 1 overwrite the first NCH - NHEAD buffers: this warp's
            // first (NCH - NHEAD) / 2 stores must have been read
            if (nch == Cfg::NCH) bulk_wait_group_read<Cfg::NCH / 2 - (Cfg::NCH - Cfg::NHEAD) / 2>();
            else bulk_wait_group_read0();
            mbar_arrive(subdone);
          }
        }
        __syncwarp();
      }
      if (dbg_on && tid == 64) p.dbg[3] = globaltimer_ns();
      if (lane == 0) bulk_wait_group0();
      if (dbg_on && tid == 64) p.dbg[4] = globaltimer_ns();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
  if (dbg_on && tid == 64) p.dbg[5] = globaltimer_ns();
}

// ------------------------------- host side -------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled fit_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}
// mn_major 0: array [mn][kd] (ld floats per row); 1: array [kd][mn].  tile_mn: rows (atoms * 32) per box.
// mn / kd: extents of the stored array (reads beyond them are zero-filled, stores are clipped).
// Returns 0 on success.
inline int fit_make_tmap(CUtensorMap* tm, const float* base, int mn_major, int mn, int kd, long long ld,
                         int models, long long stride_model, int tile_mn, int splits = 1, long long stride_split = 0) {
  PFN_tmapEncodeTiled enc = fit_tmap_encoder();
  if (!enc) return -1;
  if ((ld & 3) || (stride_model & 3) || (reinterpret_cast<uintptr_t>(base) & 15)) return -2;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  const cuuint64_t smodel = static_cast<cuuint64_t>(models > 1 ? stride_model : (long long)ld * (mn_major ? kd : mn)) * 4;
  if (!mn_major) {
    dims[0] = kd; dims[1] = mn; dims[2] = splits; dims[3] = models;   // dim 2: split-K partial (stores only)
    strides[0] = static_cast<cuuint64_t>(ld) * 4;
    strides[1] = splits > 1 ? static_cast<cuuint64_t>(stride_split) * 4 : static_cast<cuuint64_t>(ld) * 4 * mn;
    strides[2] = smodel;
    if (splits > 1 && (stride_split & 3)) return -2;
    box[0] = GM_BK; box[1] = tile_mn; box[2] = 1; box[3] = 1;
  } else {
    if (mn & 31) return -3;
    dims[0] = 32; dims[1] = kd; dims[2] = mn / 32; dims[3] = models;
    strides[0] = static_cast<cuuint64_t>(ld) * 4; strides[1] = 128; strides[2] = smodel;
    box[0] = 32; box[1] = GM_BK; box[2] = tile_mn / 32; box[3] = 1;
  }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 100 + static_cast<int>(r);
}

struct GemmOperands {
  const float* A; long long lda, strideA; int a_ext, a_kext;   // a_ext: MN extent of the stored array (>= M when
  const float* B; long long ldb, strideB; int b_ext, b_kext;   //   zero padded); *_kext: K extent (0: Kd)
  const float* aux; long long ldaux, strideAux;                // GEMM_EPI_MASK
};
// (internal linkage: the function-local `attr_set` flag must not be unified between libmetrpo.so and
// libmetrpo_dev.so, each of which carries its own copy of the kernel)
template <int MT, int BN>
static int fit_gemm_launch_t(const GemmParams& p, const GemmOperands& o, int models, cudaStream_t st) {
  using Cfg = GemmCfg<MT, BN>;
  CUtensorMap tmA, tmB, tmC, tmAux;
  int r = fit_make_tmap(&tmA, o.A, p.a_mn, o.a_ext, o.a_kext ? o.a_kext : p.Kd, o.lda, models, o.strideA, Cfg::BM);
  if (r) return 1000 + r;
  r = fit_make_tmap(&tmB, o.B, p.b_mn, o.b_ext, o.b_kext ? o.b_kext : p.Kd, o.ldb, models, o.strideB, BN);
  if (r) return 2000 + r;
  tmC = tmA; tmAux = tmA;
  if (!p.trans_store) {
    r = fit_make_tmap(&tmC, p.C, 0, p.M, p.N, p.ldc, models, p.strideC, 32, p.splits, p.strideSplit);
    if (r) return 4000 + r;
  }
  if (p.splits < 1 || (p.splits > 1 && p.epi != GEMM_EPI_PLAIN) || p.splits > (p.Kd + GM_BK - 1) / GM_BK) return 6000;
  if (p.epi == GEMM_EPI_MASK) {
    r = fit_make_tmap(&tmAux, o.aux, 0, p.M, p.N, o.ldaux, models, o.strideAux, 128);
    if (r) return 5000 + r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fit_gemm_tf32_kernel<MT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess)
      return 3000;
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + Cfg::BM - 1) / Cfg::BM, models * p.splits);
  fit_gemm_tf32_kernel<MT, BN><<<grid, GM_THREADS, Cfg::SMEM_BYTES, st>>>(p, tmA, tmB, tmC, tmAux);
  return 0;
}
// slabs (32 rows) of column-sum partials a launch with these dimensions writes per model
inline int fit_gemm_colsum_slabs(int M, int N) {
  const int bm = N <= 128 ? 128 : 256;
  return (M + bm - 1) / bm * (bm / 32);
}
// A: logical [M][Kd], B: logical [N][Kd] (see the header comment for the storage of each major-ness).
// N <= 128 problems run 128 x 128 tiles, everything else 256 x 256.
inline int fit_gemm_launch(const GemmParams& p, const GemmOperands& o, int models, cudaStream_t st) {
  return p.N <= 128 ? fit_gemm_launch_t<1, 128>(p, o, models, st) : fit_gemm_launch_t<2, 256>(p, o, models, st);
}

}  // namespace metrpo
