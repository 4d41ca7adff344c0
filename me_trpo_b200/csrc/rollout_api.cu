// C-ABI entry points of the ensemble-rollout path (include/metrpo.h): handle management, weight
// packing into the bf16 tile stream, gang schedule construction and kernel launch.
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <algorithm>

#include "common.cuh"
#include "rollout_kernel.cuh"
#include "rollout_duo.cuh"
#include "rollout_fp32.cuh"

namespace metrpo {

// ---------------------------------------------------------------------------------------------
// packing kernels (one thread per destination element)
// ---------------------------------------------------------------------------------------------
// stage (nc,kc) = W1 tile: n in [N1*nc,+N1) x k in [64kc,+64), SW128 K-major
__global__ void pack_w1_kernel(const float* __restrict__ W1, uint8_t* __restrict__ dst, int H,
                               int KC, int N1, uint32_t stage_bytes) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(H) * H) return;
  const int kglob = static_cast<int>(i / H), n = static_cast<int>(i % H);   // W1[k][n]
  const int nc = n / N1, nl = n % N1, kc = kglob / 64, kl = kglob % 64;
  uint8_t* st = dst + static_cast<size_t>(nc * KC + kc) * stage_bytes;
  *reinterpret_cast<__nv_bfloat16*>(st + sw128_off(nl, kl)) = __float2bfloat16_rn(W1[i]);
}
// W0 group tile j: n in [128j,+128) x k in [0,K0), no-swizzle core-matrix layout, zero padded.
// rows Din / Din+1 hold the bf16 hi / lo parts of b0 (Z carries 1.0 in those two columns)
__global__ void pack_w0_kernel(const float* __restrict__ W0, const float* __restrict__ b0,
                               uint8_t* __restrict__ dst, int H, int Din, int K0,
                               uint32_t w0g_bytes, uint32_t off_w0g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K0 * H) return;
  const int kk = i / H, n = i % H;
  const int j = n / 128, nl = n % 128;
  float val = 0.f;
  if (kk < Din) val = W0[kk * H + n];
  else if (kk == Din) val = __bfloat162float(__float2bfloat16_rn(b0[n]));
  else if (kk == Din + 1) val = b0[n] - __bfloat162float(__float2bfloat16_rn(b0[n]));
  *reinterpret_cast<__nv_bfloat16*>(dst + off_w0g + static_cast<size_t>(j) * w0g_bytes +
                                    noswz_off(nl, kk, 128)) = __float2bfloat16_rn(val);
}
// W2 chunk nc: N1/64 sub-tiles [S_pad rows (s)][64 k] SW128, zero padded s >= S
__global__ void pack_w2_kernel(const float* __restrict__ W2, uint8_t* __restrict__ dst, int H, int S,
                               int S_pad, int N1, uint32_t off_w2, uint32_t w2chunk_bytes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * S_pad) return;
  const int h = i / S_pad, s = i % S_pad;
  const int nc = h / N1, sub = (h % N1) / 64, kl = h % 64;
  const float v = s < S ? W2[h * S + s] : 0.f;
  uint8_t* base = dst + off_w2 + static_cast<size_t>(nc) * w2chunk_bytes + sub * (S_pad * 128);
  *reinterpret_cast<__nv_bfloat16*>(base + sw128_off(s, kl)) = __float2bfloat16_rn(v);
}
__global__ void pack_bias_kernel(const float* __restrict__ b0, const float* __restrict__ b1,
                                 const float* __restrict__ b2, float* __restrict__ dst, int H, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * H + BIAS_PAD) return;
  dst[i] = i < H ? b0[i] : (i < 2 * H ? b1[i - H] : (i - 2 * H < S ? b2[i - 2 * H] : 0.f));
}
__global__ void pack_policy_layer_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                         float* __restrict__ dst_w, float* __restrict__ dst_b, int nin,
                                         int nout, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nin * npad) {
    const int r = i / npad, c = i % npad;
    dst_w[i] = c < nout ? W[r * nout + c] : 0.f;
  }
  if (i < npad) dst_b[i] = i < nout ? b[i] : 0.f;
}
__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n,
                                float pad, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? src[i] : pad;
}
// model_costs[k] = mean over rows of row_costs[k][.]  (cost_tf's tf.reduce_mean over the batch,
// envs/com_*_env.py cost_tf; summed over time per row first, which commutes)
__global__ void mean_rows_kernel(const float* __restrict__ row_costs, float* __restrict__ out, int n) {
  __shared__ double sh[256];
  const float* src = row_costs + static_cast<size_t>(blockIdx.x) * n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(src[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = static_cast<float>(sh[0] / n);
}
__global__ void reset_rows_kernel(const float* __restrict__ states, float* __restrict__ row_state,
                                  int* __restrict__ row_ts, int* __restrict__ row_nreset, int B, int S,
                                  int rows_padded) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * S) row_state[i] = states[i];
  if (i < rows_padded) { row_ts[i] = 0; row_nreset[i] = 0; }
}

__global__ void set_rows_kernel(const int* __restrict__ rows, int n, const float* __restrict__ states,
                                float* __restrict__ row_state, int B, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * S) return;
  const int r = rows[i / S];
  if (r >= 0 && r < B) row_state[static_cast<size_t>(r) * S + i % S] = states[i];
}

}  // namespace metrpo

using namespace metrpo;

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct metrpo_rollout {
  metrpo_rollout_cfg cfg;
  int Din, K0, S_pad, NC, KC, N1, n_tiles, max_slots, num_sms;
  bool big;                       // <SMAX, AMAX> = <64, 24> instantiation (else <32, 8>)
  int smax, amax;
  uint32_t tm_acc0, tm_acc2, tm_h0, tm_z;
  int pol_in_smem;
  int rec_stride = 0;
  size_t xbuf_slot_floats = 0;
  uint32_t stage_bytes, w0g_bytes, w2chunk_bytes, off_w2, off_w0g;
  size_t model_stride;
  // device buffers
  uint8_t* wstream = nullptr;
  float* bias = nullptr;
  float* norm = nullptr;
  float* pol = nullptr;
  float* xbuf = nullptr;
  unsigned* xctr = nullptr;
  float* row_state = nullptr;
  int* row_ts = nullptr;
  int* row_nreset = nullptr;
  unsigned* tile_flag = nullptr;
  unsigned* dbg = nullptr;
  unsigned long long* trace = nullptr;
  int trace_cta = 0, trace_t0 = 0, trace_t1 = 0;
  int dbg_words = 0;
  std::map<long long, int4*> schedules;   // schedule key -> device [max_slots][MAX_SEG]
  std::map<long long, int> schedule_slots;
  float* pm_cost = nullptr;               // [K][n_envs] per-(model,row) validation costs
  // policy blob meta
  PolicyLayer pl[4];
  int pol_floats = 0, pol_logstd_off = 0;
  // smem layout
  uint32_t off_stage, off_sw0g, off_scr, off_sw2, off_sbias, off_snorm, off_spol, off_bars,
      smem_bytes;
  std::vector<char> dyn_set;
  bool norm_set = false, pol_set = false, state_set = false;
  int last_launches = 0;
  int last_warps = NUM_THREADS / 32;   // warps per CTA of the last launch (abort diagnostics)
  bool disable_own = false;   // dev / test switch: force the all-candidates exchange
  // two-stream kernel (rollout_duo.cuh)
  int duo_mode = -1;          // METRPO_DUO: -1 auto, 0 off, 1 force cs = 1, 2 force cs = 2
  bool duo_ok = false;        // shapes / TMEM / shared memory allow it
  int n_pairs = 0;
  float* duo_pbuf = nullptr; unsigned* duo_pctr = nullptr;
  float* duo_rbuf = nullptr; unsigned* duo_rctr = nullptr;
  size_t duo_pbuf_stride = 0, duo_rbuf_stride = 0;
  uint32_t duo_tm_z = 0;
  int duo_pol_rows = 32;
  uint32_t d_off_stage, d_off_sw0g, d_off_sw2, d_off_scr[2], d_off_hid[2], d_off_list[2], d_off_sbias,
      d_off_snorm, d_off_spol, d_off_bars, d_smem_bytes;
  int last_kernel = 0;        // 0: single-stream, 1: duo cs = 1, 2: duo cs = 2, 3: fp32 fidelity path
  // fp32 fidelity mode (rollout_fp32.cuh): raw weights + per-step activations
  bool fp32 = false;
  float *f_W0 = nullptr, *f_b0 = nullptr, *f_W1 = nullptr, *f_b1 = nullptr, *f_W2 = nullptr, *f_b2 = nullptr;
  float *f_x = nullptr, *f_aclip = nullptr, *f_araw = nullptr, *f_z = nullptr, *f_h0 = nullptr, *f_h1 = nullptr,
        *f_o = nullptr, *f_pm = nullptr;
};

static uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

static void free_handle(metrpo_rollout* h) {
  if (!h) return;
  cudaFree(h->wstream); cudaFree(h->bias); cudaFree(h->norm); cudaFree(h->pol); cudaFree(h->xbuf);
  cudaFree(h->xctr); cudaFree(h->row_state); cudaFree(h->row_ts); cudaFree(h->row_nreset);
  cudaFree(h->tile_flag); cudaFree(h->dbg); cudaFree(h->trace); cudaFree(h->pm_cost);
  cudaFree(h->duo_pbuf); cudaFree(h->duo_pctr); cudaFree(h->duo_rbuf); cudaFree(h->duo_rctr);
  cudaFree(h->f_W0); cudaFree(h->f_b0); cudaFree(h->f_W1); cudaFree(h->f_b1); cudaFree(h->f_W2); cudaFree(h->f_b2);
  cudaFree(h->f_x); cudaFree(h->f_aclip); cudaFree(h->f_araw); cudaFree(h->f_z); cudaFree(h->f_h0); cudaFree(h->f_h1);
  cudaFree(h->f_o); cudaFree(h->f_pm);
  for (auto& kv : h->schedules) cudaFree(kv.second);
  delete h;
}

extern "C" int metrpo_rollout_create(const metrpo_rollout_cfg* cfg, metrpo_rollout_t** out) {
  if (!cfg || !out) return set_error(METRPO_ERR_INVALID, "create: null argument");
  *out = nullptr;
  const metrpo_rollout_cfg& c = *cfg;
  if (c.state_dim < 2 || c.action_dim < 1 || c.n_models < 1 || c.n_envs < 1 || c.max_path_length < 1)
    return set_error(METRPO_ERR_INVALID, "create: S>=2, A>=1, K>=1, B>=1, T>=1 required");
  if (c.drop_cols < 0 || c.drop_cols > 2 || c.drop_cols >= c.state_dim)
    return set_error(METRPO_ERR_INVALID, "create: drop_cols must be 0, 1 or 2");
  if (c.env_id < METRPO_ENV_SWIMMER || c.env_id > METRPO_ENV_SNAKE)
    return set_error(METRPO_ERR_INVALID, "create: unknown env_id %d", c.env_id);
  if (c.sam_mode < METRPO_SAM_STEP_RAND || c.sam_mode > METRPO_SAM_ONE_MODEL)
    return set_error(METRPO_ERR_INVALID, "create: unknown sam_mode %d", c.sam_mode);
  if (c.precision != METRPO_PREC_BF16 && c.precision != METRPO_PREC_FP32)
    return set_error(METRPO_ERR_UNSUPPORTED, "create: precision must be METRPO_PREC_BF16 or METRPO_PREC_FP32");
  const bool fp32 = (c.precision == METRPO_PREC_FP32);
  if (fp32 && c.hidden < 1) return set_error(METRPO_ERR_INVALID, "create: hidden >= 1 required");
  if (!fp32) {
    bool big = c.state_dim > 32 || c.action_dim > 8;
    for (int l = 1; l < c.n_policy_layers && l < METRPO_MAX_POLICY_LAYERS; ++l) big = big || c.policy_dims[l] > HPB;
    const int mult = big ? 128 : 256;
    if (c.hidden < mult || c.hidden % mult)
      return set_error(METRPO_ERR_UNSUPPORTED, "create: dynamics hidden width must be a multiple of %d (got %d)", mult, c.hidden);
  }
  if (c.state_dim > 64 || c.action_dim > 24)
    return set_error(METRPO_ERR_UNSUPPORTED, "create: this build covers S <= 64, A <= 24 (got %d, %d)", c.state_dim, c.action_dim);
  if (c.n_policy_layers < 1 || c.n_policy_layers > METRPO_MAX_POLICY_LAYERS)
    return set_error(METRPO_ERR_INVALID, "create: n_policy_layers in [1,%d]", METRPO_MAX_POLICY_LAYERS);
  if (c.policy_dims[0] != c.state_dim || c.policy_dims[c.n_policy_layers] != c.action_dim)
    return set_error(METRPO_ERR_INVALID, "create: policy_dims must run from S to A");
  for (int l = 1; l < c.n_policy_layers; ++l)
    if (c.policy_dims[l] < 1 || c.policy_dims[l] > HPMAX)
      return set_error(METRPO_ERR_UNSUPPORTED, "create: policy hidden width <= %d in this build", HPMAX);
  // env-specific index requirements of the fused cost (envs/com_*_env.py)
  const int need_s[] = {6, 10, 6, 16, 1, 8};
  if (c.state_dim < need_s[c.env_id])
    return set_error(METRPO_ERR_INVALID, "create: env %d needs state_dim >= %d", c.env_id, need_s[c.env_id]);

  METRPO_CUDA_OK(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  METRPO_CUDA_OK(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10)
    return set_error(METRPO_ERR_UNSUPPORTED, "create: device %d is sm_%d%d; this library is sm_100a only (no fallback)", c.device, prop.major, prop.minor);
  if (c.n_models > prop.multiProcessorCount)
    return set_error(METRPO_ERR_UNSUPPORTED, "create: n_models > SM count");

  metrpo_rollout* h = new metrpo_rollout();
  h->cfg = c;
  h->fp32 = fp32;
  h->num_sms = prop.multiProcessorCount;
  h->Din = c.state_dim + c.action_dim - c.drop_cols;
  h->K0 = static_cast<int>(align_up(h->Din + 2, 16));   // +2: ones columns carrying b0 (hi, lo)
  h->S_pad = static_cast<int>(align_up(c.state_dim, 16));
  h->big = c.state_dim > 32 || c.action_dim > 8;
  for (int l = 1; l < c.n_policy_layers; ++l) h->big = h->big || c.policy_dims[l] > HPB;
  h->smax = h->big ? 64 : 32;
  h->amax = h->big ? 24 : 8;
  // TMEM budget (512 columns): acc1 N1 | acc0 128 | acc2 S_pad | H0 64 | Z K0/2.  The wide layer-1
  // pass (N1 = 256) is used whenever it fits; otherwise passes of 128 columns.
  h->N1 = h->big ? 128 : 256;   // the two compiled instantiations
  h->tm_acc0 = h->N1;
  h->tm_acc2 = h->tm_acc0 + 128;
  h->tm_h0 = h->tm_acc2 + std::max(h->S_pad, 32);
  h->tm_z = h->tm_h0 + 64;
  h->NC = std::max(1, c.hidden / h->N1);
  h->KC = std::max(2, c.hidden / 64);
  h->n_tiles = (c.n_envs + TILE_M - 1) / TILE_M;
  h->max_slots = std::min(h->n_tiles, h->num_sms / c.n_models);
  h->w0g_bytes = 128 * h->K0 * 2;
  h->stage_bytes = h->N1 * 64 * 2;
  h->w2chunk_bytes = h->S_pad * h->N1 * 2;
  h->off_w2 = h->NC * h->KC * h->stage_bytes;
  h->off_w0g = h->off_w2 + h->NC * h->w2chunk_bytes;
  h->model_stride = align_up(h->off_w0g + (c.hidden / 128) * h->w0g_bytes, 1024);
  h->dyn_set.assign(c.n_models, 0);
  {
    const char* ev = getenv("METRPO_DISABLE_OWN");   // dev switch: A/B the two exchange protocols
    h->disable_own = ev && ev[0] == '1';
  }

  // policy blob layout
  int off = 0;
  for (int l = 0; l < c.n_policy_layers; ++l) {
    PolicyLayer& L = h->pl[l];
    L.nin = c.policy_dims[l];
    L.nout = c.policy_dims[l + 1];
    L.npad = (l == c.n_policy_layers - 1) ? h->amax : static_cast<int>(align_up(L.nout, HPB));
    L.w_off = off; off += L.nin * L.npad;
    L.b_off = off; off += L.npad;
  }
  h->pol_logstd_off = off; off += h->amax;
  h->pol_floats = off;
  // activation rows of the policy in the scratch area: region A holds the layer-0 input (x) and
  // doubles as the Z staging area [S+A rows]; a hidden layer of <= 32 outputs writes in place,
  // a wider one writes to the other region
  int region_rows[2] = {c.state_dim + c.action_dim + 1, 0};   // +1: odd row stride of the own_mode policy inputs
  {
    int cur = 0;
    for (int l = 0; l < c.n_policy_layers; ++l) {
      PolicyLayer& L = h->pl[l];
      L.in_off = cur;   // region index for now
      region_rows[cur] = std::max(region_rows[cur], L.nin);
      if (l < c.n_policy_layers - 1) {
        if (L.nout > HPB) cur ^= 1;
        region_rows[cur] = std::max(region_rows[cur], L.nout);
      }
      L.out_off = cur;
    }
    for (int l = 0; l < c.n_policy_layers; ++l) {
      PolicyLayer& L = h->pl[l];
      L.in_off = L.in_off ? region_rows[0] * TILE_M : 0;
      L.out_off = L.out_off ? region_rows[0] * TILE_M : 0;
    }
  }
  h->pol_in_smem = !h->big || h->pol_floats * 4 <= 16 * 1024;

  // shared-memory carve-up (A operands live in TMEM; smem holds the weight ring + constants)
  uint32_t o = 0;
  h->off_stage = o; o += NSTAGE * h->stage_bytes;
  h->off_sw0g = o; o += 2 * h->w0g_bytes;
  o = align_up(o, 1024); h->off_sw2 = o; o += h->w2chunk_bytes;
  {
    int scr_floats = (region_rows[0] + region_rows[1]) * TILE_M;
    if (h->big) {   // own_mode buffers of the wide instantiation live in the scratch area
      const int slot_stride = (c.state_dim + c.action_dim) | 1;
      const int need = ((TILE_M * slot_stride + 3) & ~3) + TILE_M * h->amax + 2 * HPMAX * 16;   // inputs | noise | 2 x [<=128 inputs][16 rows]
      scr_floats = std::max(scr_floats, need);
    }
    h->off_scr = o; o += align_up(scr_floats * 4, 16);
  }
  h->off_sbias = o; o += (2 * c.hidden + BIAS_PAD) * 4;
  h->off_snorm = o; o += align_up((2 * (c.state_dim + c.action_dim) + 2 * c.state_dim) * 4, 16);
  h->off_spol = o; o += h->pol_in_smem ? align_up(h->pol_floats * 4, 16) : 0;
  h->off_bars = o; o += NUM_BARS * 8;
  h->smem_bytes = o + 1024;
  if (!fp32 && h->smem_bytes > static_cast<uint32_t>(prop.sharedMemPerBlockOptin)) {
    int need = h->smem_bytes;
    delete h;
    return set_error(METRPO_ERR_UNSUPPORTED, "create: config needs %d B of shared memory per CTA (limit %d)", need, (int)prop.sharedMemPerBlockOptin);
  }
  if (!fp32 && h->tm_z + h->K0 / 2 > 512) {
    delete h;
    return set_error(METRPO_ERR_UNSUPPORTED, "create: TMEM budget exceeded (S_pad %d, padded dynamics input %d)", h->S_pad, h->K0);
  }

  const size_t rows_pad = static_cast<size_t>(h->n_tiles) * TILE_M;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
  };
  if (fp32) {   // raw weights and per-step activations of the fidelity path
    const size_t K = c.n_models, H = c.hidden, B = c.n_envs, S = c.state_dim, A = c.action_dim, Din = h->Din;
    alloc(reinterpret_cast<void**>(&h->f_W0), K * Din * H * 4); alloc(reinterpret_cast<void**>(&h->f_b0), K * H * 4);
    alloc(reinterpret_cast<void**>(&h->f_W1), K * H * H * 4); alloc(reinterpret_cast<void**>(&h->f_b1), K * H * 4);
    alloc(reinterpret_cast<void**>(&h->f_W2), K * H * S * 4); alloc(reinterpret_cast<void**>(&h->f_b2), K * S * 4);
    alloc(reinterpret_cast<void**>(&h->f_x), K * B * S * 4); alloc(reinterpret_cast<void**>(&h->f_aclip), K * B * A * 4);
    alloc(reinterpret_cast<void**>(&h->f_araw), K * B * A * 4);
    alloc(reinterpret_cast<void**>(&h->f_z), K * B * Din * 4);
    alloc(reinterpret_cast<void**>(&h->f_h0), K * B * H * 4); alloc(reinterpret_cast<void**>(&h->f_h1), K * B * H * 4);
    alloc(reinterpret_cast<void**>(&h->f_o), K * B * S * 4); alloc(reinterpret_cast<void**>(&h->f_pm), 3 * K * B * 4);
  }
  alloc(reinterpret_cast<void**>(&h->wstream), fp32 ? 1024 : h->model_stride * c.n_models);
  alloc(reinterpret_cast<void**>(&h->bias), static_cast<size_t>(c.n_models) * (2 * c.hidden + BIAS_PAD) * 4);
  alloc(reinterpret_cast<void**>(&h->norm), (2 * (c.state_dim + c.action_dim) + 2 * c.state_dim) * 4);
  alloc(reinterpret_cast<void**>(&h->pol), h->pol_floats * 4);
  h->rec_stride = h->amax + 4 + static_cast<int>(align_up(c.state_dim, 4));   // [a_raw AMAX | done, pad | x_new]
  h->xbuf_slot_floats = static_cast<size_t>(std::max(c.n_models * c.state_dim, h->rec_stride)) * TILE_M;
  alloc(reinterpret_cast<void**>(&h->xbuf), static_cast<size_t>(h->max_slots) * 2 * h->xbuf_slot_floats * 4);
  alloc(reinterpret_cast<void**>(&h->xctr), h->max_slots * 4);
  alloc(reinterpret_cast<void**>(&h->row_state), rows_pad * c.state_dim * 4);
  alloc(reinterpret_cast<void**>(&h->row_ts), rows_pad * 4);
  alloc(reinterpret_cast<void**>(&h->row_nreset), rows_pad * 4);
  alloc(reinterpret_cast<void**>(&h->tile_flag), h->n_tiles * 4);
  alloc(reinterpret_cast<void**>(&h->pm_cost), static_cast<size_t>(c.n_models) * c.n_envs * 4);
  // ---- two-stream kernel (rollout_duo.cuh): narrow shapes whose second Z operand fits TMEM ----
  {
    const char* ev = getenv("METRPO_DUO");
    h->duo_mode = ev ? atoi(ev) : -1;
    h->n_pairs = (h->n_tiles + 1) / 2;
    h->duo_tm_z = h->tm_z;
    // one Z slot always; two when they fit (else the streams take turns on it: z_shared).  METRPO_DUO_ZSHARED=0
    // keeps the K0 = 48 shapes (ant) on the single-stream kernel.
    const char* evz = getenv("METRPO_DUO_ZSHARED");
    const bool allow_zs = evz ? atoi(evz) != 0 : true;
    const bool tmem_fits = h->tm_z + (allow_zs ? 1 : 2) * (h->K0 / 2) <= 512;
    const bool shape_ok = !fp32 && !h->big && c.n_models > 1 && h->n_tiles >= 2 && (h->KC % 2) == 0 &&
                          2 * c.n_models <= h->num_sms;
    // the hidden-activation scratch of the policy pass is the one elastic item: 32 owned rows per pass
    // when it fits, 16 or 8 when the K0 = 48 shapes (ant) would exceed the shared-memory limit
    for (h->duo_pol_rows = DUO_POL_ROWS; h->duo_pol_rows >= 8; h->duo_pol_rows /= 2) {
      uint32_t q = 0;
      h->d_off_stage = q; q += NSTAGE * h->stage_bytes;
      h->d_off_sw0g = q; q += 2 * h->w0g_bytes;
      q = align_up(q, 1024); h->d_off_sw2 = q; q += h->w2chunk_bytes;
      const uint32_t scr_bytes = align_up(std::max(32, c.state_dim + c.action_dim + 1) * TILE_M * 4, 16);
      for (int g = 0; g < 2; ++g) { h->d_off_scr[g] = q; q += scr_bytes; }
      for (int g = 0; g < 2; ++g) { h->d_off_hid[g] = q; q += 2 * h->duo_pol_rows * 33 * 4; }
      for (int g = 0; g < 2; ++g) { h->d_off_list[g] = q; q += (TILE_M + 4) * 4; }
      h->d_off_sbias = q; q += (c.hidden + BIAS_PAD) * 4;      // b1 of at most all passes | b2
      h->d_off_snorm = q; q += align_up((2 * (c.state_dim + c.action_dim) + 2 * c.state_dim) * 4, 16);
      h->d_off_spol = q; q += align_up(h->pol_floats * 4, 16);
      h->d_off_bars = q; q += D_NUM_BARS * 8;
      h->d_smem_bytes = q + 1024;
      if (h->d_smem_bytes + 64 <= static_cast<uint32_t>(prop.sharedMemPerBlockOptin)) break;
    }
    if (h->duo_pol_rows < 8) h->duo_pol_rows = 8;
    h->duo_ok = shape_ok && tmem_fits && h->d_smem_bytes + 64 <= static_cast<uint32_t>(prop.sharedMemPerBlockOptin) &&
                h->duo_mode != 0;
    if (h->duo_ok) {
      h->duo_pbuf_stride = static_cast<size_t>(c.n_models) * 2 * c.state_dim * TILE_M;
      h->duo_rbuf_stride = static_cast<size_t>(h->rec_stride) * TILE_M;
      const size_t streams = static_cast<size_t>(h->n_pairs) * 2;
      alloc(reinterpret_cast<void**>(&h->duo_pbuf), streams * 2 * h->duo_pbuf_stride * 4);
      alloc(reinterpret_cast<void**>(&h->duo_pctr), streams * c.n_models * 4);
      alloc(reinterpret_cast<void**>(&h->duo_rbuf), streams * 2 * h->duo_rbuf_stride * 4);
      alloc(reinterpret_cast<void**>(&h->duo_rctr), streams * 4);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(rollout_duo_kernel<32, 8, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)h->d_smem_bytes);
    }
  }
  {
    const int max_ctas = std::max(h->max_slots * c.n_models, h->num_sms);
    h->dbg_words = DBG_HEADER + max_ctas * (DUO_THREADS / 32) * DBG_WORDS_PER_WARP;
  }
  alloc(reinterpret_cast<void**>(&h->dbg), h->dbg_words * 4);
  if (e == cudaSuccess && !fp32)
    e = h->big ? cudaFuncSetAttribute(rollout_kernel<64, 24, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes)
               : cudaFuncSetAttribute(rollout_kernel<32, 8, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
  if (e != cudaSuccess) {
    free_handle(h);
    return set_error(e == cudaErrorMemoryAllocation ? METRPO_ERR_NOMEM : METRPO_ERR_CUDA, "create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_destroy(metrpo_rollout_t* h) {
  if (!h) return METRPO_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  free_handle(h);
  return METRPO_OK;
}

extern "C" int metrpo_rollout_set_dynamics(metrpo_rollout_t* h, int k, const float* W0, const float* b0,
                                           const float* W1, const float* b1, const float* W2,
                                           const float* b2, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "set_dynamics: null handle");
  if (k < 0 || k >= h->cfg.n_models) return set_error(METRPO_ERR_INVALID, "set_dynamics: model index %d out of range", k);
  if (!W0 || !b0 || !W1 || !b1 || !W2 || !b2) return set_error(METRPO_ERR_INVALID, "set_dynamics: null weight pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int H = h->cfg.hidden, S = h->cfg.state_dim;
  if (h->fp32) {
    const size_t Din = h->Din;
    auto cp = [&](float* dst, const float* src, size_t n) {
      return cudaMemcpyAsync(dst + static_cast<size_t>(k) * n, src, n * 4, cudaMemcpyDeviceToDevice, st);
    };
    METRPO_CUDA_OK(cp(h->f_W0, W0, Din * H)); METRPO_CUDA_OK(cp(h->f_b0, b0, H));
    METRPO_CUDA_OK(cp(h->f_W1, W1, static_cast<size_t>(H) * H)); METRPO_CUDA_OK(cp(h->f_b1, b1, H));
    METRPO_CUDA_OK(cp(h->f_W2, W2, static_cast<size_t>(H) * S)); METRPO_CUDA_OK(cp(h->f_b2, b2, S));
    h->dyn_set[k] = 1;
    return METRPO_OK;
  }
  uint8_t* dst = h->wstream + static_cast<size_t>(k) * h->model_stride;
  const int T = 256;
  pack_w1_kernel<<<(static_cast<size_t>(H) * H + T - 1) / T, T, 0, st>>>(W1, dst, H, h->KC, h->N1, h->stage_bytes);
  pack_w0_kernel<<<(h->K0 * H + T - 1) / T, T, 0, st>>>(W0, b0, dst, H, h->Din, h->K0, h->w0g_bytes,
                                                        h->off_w0g);
  pack_w2_kernel<<<(H * h->S_pad + T - 1) / T, T, 0, st>>>(W2, dst, H, S, h->S_pad, h->N1, h->off_w2, h->w2chunk_bytes);
  pack_bias_kernel<<<(2 * H + BIAS_PAD + T - 1) / T, T, 0, st>>>(b0, b1, b2, h->bias + static_cast<size_t>(k) * (2 * H + BIAS_PAD), H, S);
  METRPO_CUDA_OK(cudaGetLastError());
  h->dyn_set[k] = 1;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_set_normalization(metrpo_rollout_t* h, const float* in_mean,
                                                const float* in_std, const float* diff_mean,
                                                const float* diff_std, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "set_normalization: null handle");
  if (!in_mean || !in_std || !diff_mean || !diff_std) return set_error(METRPO_ERR_INVALID, "set_normalization: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int SA = h->cfg.state_dim + h->cfg.action_dim, S = h->cfg.state_dim;
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm, in_mean, SA * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + SA, in_std, SA * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + 2 * SA, diff_mean, S * 4, cudaMemcpyDeviceToDevice, st));
  METRPO_CUDA_OK(cudaMemcpyAsync(h->norm + 2 * SA + S, diff_std, S * 4, cudaMemcpyDeviceToDevice, st));
  h->norm_set = true;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_set_policy(metrpo_rollout_t* h, const float* const* W, const float* const* b,
                                         const float* log_std, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "set_policy: null handle");
  if (!W || !b || !log_std) return set_error(METRPO_ERR_INVALID, "set_policy: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  for (int l = 0; l < h->cfg.n_policy_layers; ++l) {
    if (!W[l] || !b[l]) return set_error(METRPO_ERR_INVALID, "set_policy: null layer %d", l);
    const PolicyLayer& L = h->pl[l];
    const int n = std::max(L.nin * L.npad, L.npad);
    pack_policy_layer_kernel<<<(n + 255) / 256, 256, 0, st>>>(W[l], b[l], h->pol + L.w_off, h->pol + L.b_off,
                                                              L.nin, L.nout, L.npad);
  }
  copy_f32_kernel<<<1, 32, 0, st>>>(log_std, h->pol + h->pol_logstd_off, h->cfg.action_dim, 0.f, h->amax);
  METRPO_CUDA_OK(cudaGetLastError());
  h->pol_set = true;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_reset(metrpo_rollout_t* h, const float* states, void* stream_) {
  if (!h || !states) return set_error(METRPO_ERR_INVALID, "reset: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int rows_pad = h->n_tiles * TILE_M;
  const int n = std::max(h->cfg.n_envs * h->cfg.state_dim, rows_pad);
  reset_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(states, h->row_state, h->row_ts, h->row_nreset,
                                                     h->cfg.n_envs, h->cfg.state_dim, rows_pad);
  METRPO_CUDA_OK(cudaGetLastError());
  h->state_set = true;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_set_rows(metrpo_rollout_t* h, const int32_t* rows, int n, const float* states,
                                       void* stream_) {
  if (!h || (n > 0 && (!rows || !states))) return set_error(METRPO_ERR_INVALID, "set_rows: null argument");
  if (!h->state_set) return set_error(METRPO_ERR_STATE, "set_rows: call reset() first");
  if (n <= 0) return METRPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int S = h->cfg.state_dim;
  set_rows_kernel<<<(n * S + 255) / 256, 256, 0, st>>>(rows, n, states, h->row_state, h->cfg.n_envs, S);
  METRPO_CUDA_OK(cudaGetLastError());
  return METRPO_OK;
}

// ---------------------------------------------------------------------------------------------
// gang schedule: cut the (tile x step) chains into per-slot segment lists of equal total length.
// A chain is split at most once; the slot that owns the head (t0 == 0, t1 < T) runs it FIRST,
// the slot that owns the tail runs it LAST and waits on the tile flag (always already set in
// practice: per >= T).  Heads never wait, so the schedule cannot deadlock.
// ---------------------------------------------------------------------------------------------
static int build_schedule_raw(int n_tiles, int n_slots, int T, std::vector<int4>& segs) {
  const long long total = static_cast<long long>(n_tiles) * T;
  const long long per = (total + n_slots - 1) / n_slots;
  segs.assign(static_cast<size_t>(n_slots) * MAX_SEG, make_int4(-1, 0, 0, 0));
  for (int j = 0; j < n_slots; ++j) {
    const long long lo = j * per, hi = std::min(total, lo + per);
    std::vector<int4> head, whole, tail;
    long long pos = lo;
    while (pos < hi) {
      const int tile = static_cast<int>(pos / T), t0 = static_cast<int>(pos % T);
      const int t1 = static_cast<int>(std::min<long long>(T, t0 + (hi - pos)));
      int4 sg = make_int4(tile, t0, t1, t0 > 0 ? 1 : 0);
      if (t0 > 0) tail.push_back(sg);
      else if (t1 < T) head.push_back(sg);
      else whole.push_back(sg);
      pos += t1 - t0;
    }
    std::vector<int4> order;
    order.insert(order.end(), head.begin(), head.end());
    order.insert(order.end(), whole.begin(), whole.end());
    order.insert(order.end(), tail.begin(), tail.end());
    if (order.size() > MAX_SEG) return -1;
    for (size_t i = 0; i < order.size(); ++i) segs[static_cast<size_t>(j) * MAX_SEG + i] = order[i];
  }
  return 0;
}
static int build_schedule(metrpo_rollout* h, int T, std::vector<int4>& segs, int& n_slots) {
  n_slots = h->max_slots;
  return build_schedule_raw(h->n_tiles, n_slots, T, segs);
}

// Host-only helper (no CUDA): the gang schedule for n_tiles row tiles on n_slots gang slots and a
// horizon of T steps.  out receives n_slots * max_seg int32 quadruples (tile, t0, t1, wait);
// tile == -1 marks unused entries.  Returns max_seg, or a negative status.
extern "C" int metrpo_debug_schedule(int n_tiles, int n_slots, int T, int32_t* out, int out_capacity_quads) {
  if (n_tiles < 1 || n_slots < 1 || T < 1 || !out) return set_error(METRPO_ERR_INVALID, "debug_schedule: bad argument");
  if (n_slots > n_tiles) return set_error(METRPO_ERR_INVALID, "debug_schedule: n_slots must be <= n_tiles");
  if (out_capacity_quads < n_slots * MAX_SEG) return set_error(METRPO_ERR_INVALID, "debug_schedule: output too small (need %d quads)", n_slots * MAX_SEG);
  std::vector<int4> segs;
  if (build_schedule_raw(n_tiles, n_slots, T, segs) != 0)
    return set_error(METRPO_ERR_UNSUPPORTED, "debug_schedule: more than %d segments per slot", MAX_SEG);
  std::memcpy(out, segs.data(), segs.size() * sizeof(int4));
  return MAX_SEG;
}

// per-model cost rollouts carry per-(model,row) accumulators in registers, so their tile chains are
// never split: slot j runs tiles j, j + n_slots, .. whole
static int build_schedule_whole(int n_tiles, int n_slots, int T, std::vector<int4>& segs) {
  segs.assign(static_cast<size_t>(n_slots) * MAX_SEG, make_int4(-1, 0, 0, 0));
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int j = tile % n_slots, i = tile / n_slots;
    if (i >= MAX_SEG) return -1;
    segs[static_cast<size_t>(j) * MAX_SEG + i] = make_int4(tile, 0, T, 0);
  }
  return 0;
}

// whole_tiles > 0: per-model schedule over that many row tiles
static int get_schedule(metrpo_rollout* h, int T, int whole_tiles, const int4** dev, int* n_slots, cudaStream_t st) {
  const long long key = (static_cast<long long>(whole_tiles) << 32) | static_cast<unsigned>(T);
  auto it = h->schedules.find(key);
  if (it != h->schedules.end()) {
    *dev = it->second;
    *n_slots = h->schedule_slots[key];
    return METRPO_OK;
  }
  std::vector<int4> segs;
  int ns = 0;
  int brc;
  if (whole_tiles > 0) {
    ns = std::min(whole_tiles, h->max_slots);
    brc = build_schedule_whole(whole_tiles, ns, T, segs);
  } else {
    brc = build_schedule(h, T, segs, ns);
  }
  if (brc != 0)
    return set_error(METRPO_ERR_UNSUPPORTED, "run: schedule needs more than %d segments per slot (n_tiles=%d, slots=%d)", MAX_SEG, h->n_tiles, h->max_slots);
  int4* d = nullptr;
  METRPO_CUDA_OK(cudaMalloc(&d, segs.size() * sizeof(int4)));
  // one-time synchronous upload (first call with this horizon only)
  METRPO_CUDA_OK(cudaMemcpy(d, segs.data(), segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
  (void)st;
  h->schedules[key] = d;
  h->schedule_slots[key] = ns;
  *dev = d;
  *n_slots = ns;
  return METRPO_OK;
}

// fp32 fidelity path (rollout_fp32.cuh): the same step semantics, one launch per layer and step
static int launch_fp32(metrpo_rollout* h, const KParams& kp, cudaStream_t st) {
  const metrpo_rollout_cfg& c = h->cfg;
  const int K = c.n_models, S = c.state_dim, A = c.action_dim, H = c.hidden, Din = h->Din;
  const int B = kp.per_model ? kp.B : c.n_envs;
  const int sets = kp.per_model ? K : 1;
  Fp32Params p;
  std::memset(&p, 0, sizeof(p));
  p.S = S; p.A = A; p.SA = S + A; p.drop = c.drop_cols; p.Din = Din; p.H = H; p.K = K; p.B = B;
  p.T_max = c.max_path_length; p.env_id = c.env_id; p.sam_mode = c.sam_mode; p.determ = kp.determ;
  p.per_model = kp.per_model; p.n_pol_layers = c.n_policy_layers; p.pol_out_tanh = c.policy_out_tanh;
  p.pol_logstd_off = h->pol_logstd_off; p.row_offset = c.row_offset;
  for (int l = 0; l < 4; ++l) p.pl[l] = h->pl[l];
  p.pol = h->pol; p.norm = h->norm; p.gamma = kp.gamma;
  p.eps = kp.eps; p.model_idx = kp.model_idx; p.std_noise = kp.std_noise;
  p.ext_actions = kp.ext_actions; p.ext_reset_states = kp.ext_reset_states;
  p.reset_pool = kp.reset_pool; p.R = kp.R > 0 ? kp.R : 1; p.seed = kp.seed; p.offset = kp.offset;
  p.ts = h->row_ts; p.nreset = h->row_nreset;
  p.a_clip = h->f_aclip; p.a_raw = h->f_araw; p.z = h->f_z; p.o = h->f_o;
  p.pm_acc = h->f_pm; p.pm_gpow = h->f_pm + static_cast<size_t>(K) * c.n_envs; p.pm_dmask = h->f_pm + 2 * static_cast<size_t>(K) * c.n_envs;
  p.obs = kp.obs; p.act = kp.act; p.mean = kp.mean; p.rew = kp.rew; p.done = kp.done;
  const int TPB = 128;
  const size_t nstate = static_cast<size_t>(B) * S;
  if (kp.per_model) {   // every model rolls its own copy of the start states
    p.x = h->f_x;
    fp32_tile_states<<<(nstate * K + 255) / 256, 256, 0, st>>>(kp.init_states, h->f_x, K, nstate);
    const size_t n = static_cast<size_t>(K) * B;
    fp32_fill<<<(n + 255) / 256, 256, 0, st>>>(p.pm_acc, 0.f, n);
    fp32_fill<<<(n + 255) / 256, 256, 0, st>>>(p.pm_gpow, 1.f, n);
    fp32_fill<<<(n + 255) / 256, 256, 0, st>>>(p.pm_dmask, 0.f, n);
  } else {
    p.x = h->row_state;
    if (!kp.resume) {
      const int rows_pad = h->n_tiles * TILE_M;
      const int n = std::max(B * S, rows_pad);
      reset_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(kp.init_states, h->row_state, h->row_ts, h->row_nreset, B, S, rows_pad);
    }
  }
  const int rows = sets * B;
  const dim3 g0((H + 63) / 64, (B + 63) / 64, K), g2((S + 63) / 64, (B + 63) / 64, K);
  int launches = 0;
  for (int t = 0; t < kp.n_steps; ++t) {
    p.t = t;
    fp32_begin_step<<<(rows + TPB - 1) / TPB, TPB, 0, st>>>(p);
    fp32_gemm_bias_act<<<g0, 256, 0, st>>>(h->f_z, kp.per_model ? static_cast<long long>(B) * Din : 0LL, h->f_W0,
                                           static_cast<long long>(Din) * H, h->f_b0, H, h->f_h0,
                                           static_cast<long long>(B) * H, B, H, Din, 1);
    fp32_gemm_bias_act<<<g0, 256, 0, st>>>(h->f_h0, static_cast<long long>(B) * H, h->f_W1, static_cast<long long>(H) * H,
                                           h->f_b1, H, h->f_h1, static_cast<long long>(B) * H, B, H, H, 1);
    fp32_gemm_bias_act<<<g2, 256, 0, st>>>(h->f_h1, static_cast<long long>(B) * H, h->f_W2, static_cast<long long>(H) * S,
                                           h->f_b2, S, h->f_o, static_cast<long long>(B) * S, B, S, H, 0);
    if (h->big) fp32_finish_step<64, 24><<<(rows + TPB - 1) / TPB, TPB, 0, st>>>(p);
    else fp32_finish_step<32, 8><<<(rows + TPB - 1) / TPB, TPB, 0, st>>>(p);
    launches += 5;
  }
  METRPO_CUDA_OK(cudaGetLastError());
  if (kp.per_model) {
    METRPO_CUDA_OK(cudaMemcpyAsync(kp.pm_cost, p.pm_acc, static_cast<size_t>(K) * B * 4, cudaMemcpyDeviceToDevice, st));
  } else if (kp.final_states) {
    METRPO_CUDA_OK(cudaMemcpyAsync(kp.final_states, h->row_state, nstate * 4, cudaMemcpyDeviceToDevice, st));
  }
  METRPO_CUDA_OK(cudaMemsetAsync(h->dbg, 0, 4, st));
  h->last_launches = launches;
  h->last_kernel = 3;
  return METRPO_OK;
}

static int launch(metrpo_rollout* h, KParams& p, cudaStream_t st) {
  for (int k = 0; k < h->cfg.n_models; ++k)
    if (!h->dyn_set[k]) return set_error(METRPO_ERR_STATE, "run: dynamics model %d was never set", k);
  if (!h->norm_set) return set_error(METRPO_ERR_STATE, "run: normalization constants were never set");
  const metrpo_rollout_cfg& c = h->cfg;
  if (h->fp32) return launch_fp32(h, p, st);
  p.S = c.state_dim; p.A = c.action_dim; p.SA = p.S + p.A; p.drop = c.drop_cols; p.Din = h->Din;
  p.K0 = h->K0; p.H = c.hidden; p.S_pad = h->S_pad; p.K = c.n_models;
  if (!p.per_model) p.B = c.n_envs;
  p.T_max = c.max_path_length; p.env_id = c.env_id; p.sam_mode = c.sam_mode;
  p.NC = h->NC; p.KC = h->KC; p.N1 = h->N1; p.row_offset = c.row_offset;
  p.tm_acc0 = h->tm_acc0; p.tm_acc2 = h->tm_acc2; p.tm_h0 = h->tm_h0; p.tm_z = h->tm_z;
  p.pol_in_smem = h->pol_in_smem;
  // row-ownership exchange: selection modes only (every row has exactly one source model per step)
  p.own_mode = (c.n_models > 1 && !p.per_model && p.ext_actions == nullptr &&
                (c.sam_mode == METRPO_SAM_STEP_RAND || c.sam_mode == METRPO_SAM_EPS_RAND) && !h->disable_own) ? 1 : 0;
  p.rec_stride = h->rec_stride;
  p.xbuf_stride = h->xbuf_slot_floats;
  p.slot_stride = (c.state_dim + c.action_dim) | 1;
  p.n_tiles = p.per_model ? (p.B + TILE_M - 1) / TILE_M : h->n_tiles;
  p.wstream = h->wstream; p.model_stride = h->model_stride; p.stage_bytes = h->stage_bytes;
  p.w0g_bytes = h->w0g_bytes; p.w2chunk_bytes = h->w2chunk_bytes; p.off_w2 = h->off_w2;
  p.off_w0g = h->off_w0g; p.bias = h->bias; p.norm = h->norm; p.pol = h->pol;
  p.pol_floats = h->pol_floats; p.n_pol_layers = c.n_policy_layers; p.pol_out_tanh = c.policy_out_tanh;
  p.pol_logstd_off = h->pol_logstd_off;
  for (int l = 0; l < 4; ++l) p.pl[l] = h->pl[l];
  p.xbuf = h->xbuf; p.xctr = h->xctr; p.row_state = h->row_state; p.row_ts = h->row_ts;
  p.row_nreset = h->row_nreset; p.tile_flag = h->tile_flag; p.dbg = h->dbg;
  p.trace = h->trace; p.trace_cta = h->trace_cta; p.trace_t0 = h->trace_t0; p.trace_t1 = h->trace_t1;
  p.off_stage = h->off_stage; p.off_sw0g = h->off_sw0g; p.off_scr = h->off_scr;
  p.off_sw2 = h->off_sw2; p.off_sbias = h->off_sbias; p.off_snorm = h->off_snorm;
  p.off_spol = h->off_spol; p.off_bars = h->off_bars;
  // ---- two-stream kernel: pick the column split that keeps the most SMs busy ----
  if (h->duo_ok && p.own_mode && !p.per_model && p.ext_actions == nullptr) {
    const int K = c.n_models, NCtot = h->NC;
    const int old_ctas = std::min(h->n_tiles, h->num_sms / K) * K;
    int best_cs = 0, best_ctas = 0;
    for (int cs = 1; cs <= 2; ++cs) {
      if (NCtot % cs) continue;
      if (h->duo_mode > 0 && h->duo_mode != cs) continue;
      const int slots = std::min(h->n_pairs, h->num_sms / (K * cs));
      if (slots < 1) continue;
      const int ctas = slots * K * cs;
      if (ctas > best_ctas) { best_ctas = ctas; best_cs = cs; }
    }
    // Measured on B200 (tools/duo_probe.py, profiles/r2_duo_probe.json): without column split a
    // two-stream CTA needs ~12 % fewer cycles per tile-step than a single-stream one; with the
    // column split the extra stream switches eat the gain (half-cheetah K = 5, B = 4096: 80.8 k vs
    // 84 k cycles per tile-step on 140 instead of 145 SMs = a tie).  Automatic selection therefore
    // takes the two-stream kernel only unsplit and only when it keeps as many SMs busy.
    if (best_cs && (h->duo_mode > 0 || (best_cs == 1 && best_ctas >= old_ctas))) {
      DuoParams dpar;
      std::memset(&dpar, 0, sizeof(dpar));
      dpar.cs = best_cs; dpar.NCp = NCtot / best_cs; dpar.n_pairs = h->n_pairs;
      dpar.z_shared = (h->tm_z + 2 * (h->K0 / 2) > 512) ? 1 : 0;
      dpar.pol_rows = h->duo_pol_rows;
      const int slots = best_ctas / (K * best_cs);
      // schedule over tile PAIRS: key space disjoint from the single-stream schedules
      const long long key = (static_cast<long long>(0x40000000 | slots) << 32) | static_cast<unsigned>(p.n_steps);
      const int4* dsegs = nullptr;
      auto it = h->schedules.find(key);
      if (it != h->schedules.end()) {
        dsegs = it->second;
      } else {
        std::vector<int4> segs;
        if (build_schedule_raw(h->n_pairs, slots, p.n_steps, segs) != 0)
          return set_error(METRPO_ERR_UNSUPPORTED, "run: duo schedule needs more than %d segments per slot", MAX_SEG);
        int4* d = nullptr;
        METRPO_CUDA_OK(cudaMalloc(&d, segs.size() * sizeof(int4)));
        METRPO_CUDA_OK(cudaMemcpy(d, segs.data(), segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
        h->schedules[key] = d;
        h->schedule_slots[key] = slots;
        dsegs = d;
      }
      p.segs = dsegs; p.n_slots = slots;
      p.off_stage = h->d_off_stage; p.off_sw0g = h->d_off_sw0g; p.off_sw2 = h->d_off_sw2;
      p.off_sbias = h->d_off_sbias; p.off_snorm = h->d_off_snorm; p.off_spol = h->d_off_spol;
      p.off_bars = h->d_off_bars; p.off_scr = h->d_off_scr[0];
      dpar.b = p;
      dpar.pbuf = h->duo_pbuf; dpar.pctr = h->duo_pctr; dpar.pbuf_stride = h->duo_pbuf_stride;
      dpar.rbuf = h->duo_rbuf; dpar.rctr = h->duo_rctr; dpar.rbuf_stride = h->duo_rbuf_stride;
      for (int g = 0; g < 2; ++g) {
        dpar.off_scr[g] = h->d_off_scr[g]; dpar.off_hid[g] = h->d_off_hid[g]; dpar.off_list[g] = h->d_off_list[g];
      }
      const size_t streams = static_cast<size_t>(h->n_pairs) * 2;
      METRPO_CUDA_OK(cudaMemsetAsync(h->duo_pctr, 0, streams * K * 4, st));
      METRPO_CUDA_OK(cudaMemsetAsync(h->duo_rctr, 0, streams * 4, st));
      METRPO_CUDA_OK(cudaMemsetAsync(h->tile_flag, 0, h->n_tiles * 4, st));
      METRPO_CUDA_OK(cudaMemsetAsync(h->dbg, 0, h->dbg_words * 4, st));
      void* dargs[] = {&dpar};
      METRPO_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rollout_duo_kernel<32, 8, 256>),
                                                 dim3(best_ctas), dim3(DUO_THREADS), dargs, h->d_smem_bytes, st));
      h->last_launches = 1;
      h->last_warps = DUO_THREADS / 32;
      h->last_kernel = best_cs;
      return METRPO_OK;
    }
  }
  int n_slots = 0;
  int rc = get_schedule(h, p.n_steps, p.per_model ? p.n_tiles : 0, &p.segs, &n_slots, st);
  if (rc != METRPO_OK) return rc;
  p.n_slots = n_slots;
  METRPO_CUDA_OK(cudaMemsetAsync(h->xctr, 0, h->max_slots * 4, st));
  METRPO_CUDA_OK(cudaMemsetAsync(h->tile_flag, 0, h->n_tiles * 4, st));
  METRPO_CUDA_OK(cudaMemsetAsync(h->dbg, 0, h->dbg_words * 4, st));
  void* args[] = {&p};
  // cooperative launch: all gang CTAs must be co-resident (they spin on each other)
  METRPO_CUDA_OK(cudaLaunchCooperativeKernel(h->big ? reinterpret_cast<void*>(rollout_kernel<64, 24, 128>)
                                                    : reinterpret_cast<void*>(rollout_kernel<32, 8, 256>),
                                             dim3(n_slots * c.n_models), dim3(NUM_THREADS), args,
                                             h->smem_bytes, st));
  h->last_launches = 1;
  h->last_warps = NUM_THREADS / 32;
  h->last_kernel = 0;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_run(metrpo_rollout_t* h, int n_steps, const float* init_states,
                                  const float* reset_pool, int R, const float* eps,
                                  const int32_t* model_idx, const float* std_noise, uint64_t seed,
                                  uint64_t offset, int determ, float* obs, float* act, float* mean,
                                  float* rew, uint8_t* done, float* final_states, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "run: null handle");
  if (n_steps < 1) return set_error(METRPO_ERR_INVALID, "run: n_steps must be >= 1");
  if (!init_states) return set_error(METRPO_ERR_INVALID, "run: init_states is required");
  if (!reset_pool || R < 1) return set_error(METRPO_ERR_INVALID, "run: a reset pool with R >= 1 states is required");
  if (!h->pol_set) return set_error(METRPO_ERR_STATE, "run: policy was never set");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_steps = n_steps; p.resume = 0; p.determ = determ ? 1 : 0;
  p.init_states = init_states; p.reset_pool = reset_pool; p.R = R; p.eps = eps; p.model_idx = model_idx;
  p.std_noise = std_noise; p.seed = seed; p.offset = offset;
  p.obs = obs; p.act = act; p.mean = mean; p.rew = rew; p.done = done; p.final_states = final_states;
  int rc = launch(h, p, static_cast<cudaStream_t>(stream_));
  if (rc == METRPO_OK) h->state_set = true;
  return rc;
}

extern "C" int metrpo_rollout_continue(metrpo_rollout_t* h, int n_steps, const float* reset_pool, int R,
                                       const float* eps, const int32_t* model_idx, const float* std_noise,
                                       uint64_t seed, uint64_t offset, int determ, float* obs, float* act,
                                       float* mean, float* rew, uint8_t* done, float* final_states,
                                       void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "continue: null handle");
  if (n_steps < 1) return set_error(METRPO_ERR_INVALID, "continue: n_steps must be >= 1");
  if (!reset_pool || R < 1) return set_error(METRPO_ERR_INVALID, "continue: a reset pool with R >= 1 states is required");
  if (!h->pol_set) return set_error(METRPO_ERR_STATE, "continue: policy was never set");
  if (!h->state_set) return set_error(METRPO_ERR_STATE, "continue: no row state yet (call run() or reset() first)");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_steps = n_steps; p.resume = 1; p.determ = determ ? 1 : 0;
  p.reset_pool = reset_pool; p.R = R; p.eps = eps; p.model_idx = model_idx;
  p.std_noise = std_noise; p.seed = seed; p.offset = offset;
  p.obs = obs; p.act = act; p.mean = mean; p.rew = rew; p.done = done; p.final_states = final_states;
  return launch(h, p, static_cast<cudaStream_t>(stream_));
}

extern "C" int metrpo_rollout_model_costs(metrpo_rollout_t* h, int n_steps, int n_rows,
                                          const float* init_states, double gamma, float* row_costs,
                                          float* model_costs, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "model_costs: null handle");
  if (n_steps < 1) return set_error(METRPO_ERR_INVALID, "model_costs: n_steps must be >= 1");
  if (n_rows < 1 || n_rows > h->cfg.n_envs)
    return set_error(METRPO_ERR_INVALID, "model_costs: n_rows must be in [1, n_envs=%d]", h->cfg.n_envs);
  if (!init_states || !model_costs) return set_error(METRPO_ERR_INVALID, "model_costs: init_states and model_costs are required");
  if (!h->pol_set) return set_error(METRPO_ERR_STATE, "model_costs: policy was never set");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_steps = n_steps; p.resume = 0; p.determ = 1; p.per_model = 1; p.B = n_rows;
  p.gamma = static_cast<float>(gamma);
  p.pm_cost = row_costs ? row_costs : h->pm_cost;
  p.init_states = init_states; p.R = 1;
  int rc = launch(h, p, st);
  if (rc != METRPO_OK) return rc;
  mean_rows_kernel<<<h->cfg.n_models, 256, 0, st>>>(p.pm_cost, model_costs, n_rows);
  METRPO_CUDA_OK(cudaGetLastError());
  h->last_launches = 2;
  return METRPO_OK;
}

extern "C" int metrpo_rollout_step(metrpo_rollout_t* h, const float* actions, const int32_t* model_idx,
                                   const float* std_noise, const float* reset_states, uint64_t seed,
                                   uint64_t offset, float* obs_out, float* rew_out, uint8_t* done_out,
                                   void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "step: null handle");
  if (!actions || !reset_states) return set_error(METRPO_ERR_INVALID, "step: actions and reset_states are required");
  if (!h->state_set) return set_error(METRPO_ERR_STATE, "step: call reset() first");
  METRPO_CUDA_OK(cudaSetDevice(h->cfg.device));
  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_steps = 1; p.resume = 1; p.determ = 0;
  p.ext_actions = actions; p.ext_reset_states = reset_states; p.model_idx = model_idx;
  p.std_noise = std_noise; p.seed = seed; p.offset = offset; p.R = 1;
  p.rew = rew_out; p.done = done_out; p.final_states = obs_out;
  return launch(h, p, static_cast<cudaStream_t>(stream_));
}

extern "C" int metrpo_rollout_last_launches(const metrpo_rollout_t* h) { return h ? h->last_launches : 0; }
extern "C" int metrpo_rollout_last_kernel(const metrpo_rollout_t* h) { return h ? h->last_kernel : 0; }

// Synchronise `stream` and report whether the last launch completed: a kernel whose internal
// waits timed out sets an abort flag and leaves; the message lists which role of which CTA was
// waiting on which barrier (diagnostics for the mbarrier protocol).
extern "C" int metrpo_rollout_status(metrpo_rollout_t* h, void* stream_) {
  if (!h) return set_error(METRPO_ERR_INVALID, "status: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  METRPO_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<unsigned> host(h->dbg_words);
  METRPO_CUDA_OK(cudaMemcpy(host.data(), h->dbg, h->dbg_words * 4, cudaMemcpyDeviceToHost));
  if (host[0] == 0) return METRPO_OK;
  char msg[480];
  int n = snprintf(msg, sizeof(msg), "rollout kernel aborted (wait timeout):");
  const int warps = h->last_warps;
  const int n_ctas = (h->dbg_words - DBG_HEADER) / (warps * DBG_WORDS_PER_WARP);
  int shown = 0;
  for (int pass = 1; pass <= 2 && shown < 10; ++pass)   // timed-out waits first, then observers
    for (int b = 0; b < n_ctas && shown < 10; ++b)
      for (int w = 0; w < warps && shown < 10; ++w) {
        const unsigned* r = &host[DBG_HEADER + (b * warps + w) * DBG_WORDS_PER_WARP];
        if (r[2] != (unsigned)pass) continue;
        n += snprintf(msg + n, sizeof(msg) - n, " [cta%d w%d %s bar=%u st=%u g=%u]", b, w,
                      pass == 1 ? "TIMEOUT" : "saw", r[0], r[1] >> 8, r[1] & 0xff);
        ++shown;
      }
  return set_error(METRPO_ERR_STATE, "%s", msg);
}

// Dev tool: record an event trace (role-private logs of code<<40 | clock64) of CTA `cta` for its
// local steps [t0, t1).  out (host, 3*4096 u64) receives the logs after a synchronising
// metrpo_rollout_get_trace.
extern "C" int metrpo_rollout_set_trace(metrpo_rollout_t* h, int cta, int t0, int t1) {
  if (!h) return set_error(METRPO_ERR_INVALID, "set_trace: null handle");
#ifndef METRPO_TRACE
  return set_error(METRPO_ERR_UNSUPPORTED, "set_trace: library was built without -DMETRPO_TRACE");
#endif
  if (!h->trace) {
    METRPO_CUDA_OK(cudaMalloc(&h->trace, 4 * TRACE_CAP * 8));
  }
  METRPO_CUDA_OK(cudaMemset(h->trace, 0, 4 * TRACE_CAP * 8));
  h->trace_cta = cta; h->trace_t0 = t0; h->trace_t1 = t1;
  return METRPO_OK;
}
extern "C" int metrpo_rollout_get_trace(metrpo_rollout_t* h, unsigned long long* out_host) {
  if (!h || !h->trace || !out_host) return set_error(METRPO_ERR_INVALID, "get_trace: no trace");
  METRPO_CUDA_OK(cudaDeviceSynchronize());
  METRPO_CUDA_OK(cudaMemcpy(out_host, h->trace, 4 * TRACE_CAP * 8, cudaMemcpyDeviceToHost));
  return METRPO_OK;
}
