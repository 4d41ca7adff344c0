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
// Roofline: fp32 operands make this L2-bandwidth bound at 128-row tiles (48 KB per 32-deep K block
// for 512 tensor-pipe cycles); the 256 x 256 tile halves the bytes per FLOP and leaves 80 CTAs for
// the fit's shapes (K = 5 models x 4 x 4 tiles), one wave.
#pragma once
#include <cuda.h>
#include "umma.cuh"

namespace metrpo {

constexpr int GM_BK = 32;
constexpr int GM_THREADS = 192;
constexpr int GM_CHUNK_BYTES = 128 * 128;            // one 128-row x 32-column fp32 epilogue chunk
template <int MT, int BN> struct GemmCfg {
  static constexpr int BM = 128 * MT;
  static constexpr int A_BYTES = BM * GM_BK * 4;     // 16 / 32 KB
  static constexpr int B_BYTES = BN * GM_BK * 4;     // 16 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE_BYTES > 49152) ? 3 : 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int NCH = BN / 32;                // epilogue chunks per 128-row sub-tile
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
  int trans_store;              // 1: store C transposed, C[n * ldc + m], from registers (GEMM_EPI_PLAIN only)
  float* C; long long ldc, strideC;                     // used by the transposed store only (else tmC)
  const float* bias; long long strideBias;              // GEMM_EPI_BIAS_RELU: bias[n] per model
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
  uint64_t* auxfull = accfull + 1;             // mask tile of the current 128-row sub-tile landed
  uint64_t* subdone = accfull + 2;             // all four epilogue warps finished a sub-tile (4 arrivals)
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * Cfg::BM, model = blockIdx.z;
  const int nkb = (p.Kd + GM_BK - 1) / GM_BK;
  const int nch = min(Cfg::NCH, (p.N - n0 + 31) / 32);   // chunks of this tile that hold real columns

  if (tid == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accfull, 1);
    mbar_init(auxfull, 1);
    mbar_init(subdone, 4);
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

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sA = sStage + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        if (p.a_mn) tma_load_4d(sA, &tmA, 0, kb * GM_BK, m0 / 32, model, &full[s]);
        else        tma_load_4d(sA, &tmA, kb * GM_BK, m0, 0, model, &full[s]);
        if (p.b_mn) tma_load_4d(sB, &tmB, 0, kb * GM_BK, n0 / 32, model, &full[s]);
        else        tma_load_4d(sB, &tmB, kb * GM_BK, n0, 0, model, &full[s]);
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
      if (p.epi == GEMM_EPI_MASK) {
        // mask tiles of the epilogue go into the (now idle) stage ring, one 128-row sub-tile at a time
        mbar_wait(accfull, 0);
        for (int mt = 0; mt < MT; ++mt) {
          if (mt > 0) mbar_wait(subdone, (mt - 1) & 1);
          mbar_arrive_expect_tx(auxfull, nch * GM_CHUNK_BYTES);
          for (int ch = 0; ch < nch; ++ch)
            tma_load_4d(sStage + ch * GM_CHUNK_BYTES, &tmAux, n0 + ch * 32, m0 + mt * 128, 0, model, auxfull);
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
    const float* biasm = p.bias ? p.bias + model * p.strideBias : nullptr;
    if (p.epi == GEMM_EPI_BIAS_RELU) {
      for (int i = tid - 64; i < BN; i += 128) sBias[i] = (n0 + i < p.N) ? biasm[n0 + i] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(accfull, 0);
    tc_fence_after();
    const int rl = q * 32 + lane;                    // row inside the 128-row sub-tile
    const uint32_t sw = rl & 7;
    for (int mt = 0; mt < MT; ++mt) {
      const int mrow0 = m0 + mt * 128 + q * 32;
      if (p.epi == GEMM_EPI_MASK) mbar_wait(auxfull, mt & 1);
      float* csum = p.colsum ? p.colsum + (static_cast<size_t>(model) * gridDim.y * MT * 4 +
                                           (blockIdx.y * MT + mt) * 4 + q) * p.N : nullptr;
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + mt * BN + ch * 32, v);
        tmem_ld_wait();
        if (p.trans_store) {                   // thread = row m, register j = column: C[n][m], coalesced over lanes
          float* Cm = p.C + model * p.strideC;
          if (mrow0 + lane < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + ch * 32 + j < p.N) Cm[static_cast<size_t>(n0 + ch * 32 + j) * p.ldc + mrow0 + lane] = __uint_as_float(v[j]);
          }
          continue;
        }
        uint8_t* buf = sStage + ch * GM_CHUNK_BYTES;
        uint8_t* rowp = buf + rl * 128;
        if (p.epi == GEMM_EPI_MASK) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 a = *reinterpret_cast<const float4*>(rowp + ((c ^ sw) << 4));
            if (!(a.x > 0.f)) v[4 * c] = 0u;
            if (!(a.y > 0.f)) v[4 * c + 1] = 0u;
            if (!(a.z > 0.f)) v[4 * c + 2] = 0u;
            if (!(a.w > 0.f)) v[4 * c + 3] = 0u;
          }
        } else if (p.epi == GEMM_EPI_BIAS_RELU) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 b = *reinterpret_cast<const float4*>(sBias + ch * 32 + 4 * c);
            v[4 * c] = __float_as_uint(fmaxf(__uint_as_float(v[4 * c]) + b.x, 0.f));
            v[4 * c + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * c + 1]) + b.y, 0.f));
            v[4 * c + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * c + 2]) + b.z, 0.f));
            v[4 * c + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * c + 3]) + b.w, 0.f));
          }
        }
        if (p.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(round_tf32(__uint_as_float(v[j])));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        if (csum) {   // lane j sums column j over the warp's 32 rows (rows past M hold zeros)
          float cs = 0.f;
          const uint8_t* slab = buf + q * 4096;
#pragma unroll 8
          for (int i = 0; i < 32; ++i)
            cs += *reinterpret_cast<const float*>(slab + i * 128 + ((((lane >> 2) ^ i) & 7) << 4) + (lane & 3) * 4);
          const int n = n0 + ch * 32 + lane;
          if (n < p.N) csum[n] = cs;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmC, buf + q * 4096, n0 + ch * 32, mrow0, 0, model);
          bulk_commit_group();
        }
      }
      if (MT > 1 && !p.trans_store && mt + 1 < MT) {
        // the next sub-tile reuses the staging chunks: the bulk stores must have read them
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
        if (p.epi == GEMM_EPI_MASK && lane == 0) mbar_arrive(subdone);
      }
    }
    if (lane == 0 && !p.trans_store) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// ------------------------------- host side -------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
inline PFN_tmapEncodeTiled fit_tmap_encoder() {
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
                         int models, long long stride_model, int tile_mn) {
  PFN_tmapEncodeTiled enc = fit_tmap_encoder();
  if (!enc) return -1;
  if ((ld & 3) || (stride_model & 3) || (reinterpret_cast<uintptr_t>(base) & 15)) return -2;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  const cuuint64_t smodel = static_cast<cuuint64_t>(models > 1 ? stride_model : (long long)ld * (mn_major ? kd : mn)) * 4;
  if (!mn_major) {
    dims[0] = kd; dims[1] = mn; dims[2] = 1; dims[3] = models;
    strides[0] = static_cast<cuuint64_t>(ld) * 4; strides[1] = static_cast<cuuint64_t>(ld) * 4 * mn; strides[2] = smodel;
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
template <int MT, int BN>
inline int fit_gemm_launch_t(const GemmParams& p, const GemmOperands& o, int models, cudaStream_t st) {
  using Cfg = GemmCfg<MT, BN>;
  CUtensorMap tmA, tmB, tmC, tmAux;
  int r = fit_make_tmap(&tmA, o.A, p.a_mn, o.a_ext, o.a_kext ? o.a_kext : p.Kd, o.lda, models, o.strideA, Cfg::BM);
  if (r) return 1000 + r;
  r = fit_make_tmap(&tmB, o.B, p.b_mn, o.b_ext, o.b_kext ? o.b_kext : p.Kd, o.ldb, models, o.strideB, BN);
  if (r) return 2000 + r;
  tmC = tmA; tmAux = tmA;
  if (!p.trans_store) {
    r = fit_make_tmap(&tmC, p.C, 0, p.M, p.N, p.ldc, models, p.strideC, 32);
    if (r) return 4000 + r;
  }
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
  dim3 grid((p.N + BN - 1) / BN, (p.M + Cfg::BM - 1) / Cfg::BM, models);
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
