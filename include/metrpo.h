/*
 * metrpo.h -- C ABI of the B200-native ME-TRPO hot path (libmetrpo.so).
 *
 * The reference (thanard/me-trpo) is pure Python + TF 1.4 and has no FFI; its "plugin API" for
 * the imaginary-rollout path is three duck-typed Python sockets.  Each entry
 * point below names the reference interface it replaces (file:line under the reference tree).
 * The Python classes in me_trpo_b200/ that mirror those sockets bind these symbols via ctypes
 * (me_trpo_b200/lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer on the handle's device unless it says "host";
 *     the caller (torch) owns all buffers, the library owns only the handle, its packed weight
 *     images and its workspace;
 *   - every call is asynchronous on the passed cudaStream_t (as void*), no implicit sync;
 *   - returns 0 (METRPO_OK) or a negative metrpo_status_t; message via metrpo_last_error()
 *     (thread-local).  Nothing throws or aborts across the ABI;
 *   - a handle is bound to one device and is not thread-safe; distinct handles are independent;
 *   - weights use the reference's TF layout: W[in, out] row-major, y = x @ W + b
 *     (training.py:187-208);
 *   - there is no CPU fallback: on a device that is not sm_100 create() fails.
 */
#ifndef METRPO_H_
#define METRPO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  METRPO_OK = 0,
  METRPO_ERR_INVALID = -1,     /* bad argument / shape */
  METRPO_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the cudaError) */
  METRPO_ERR_UNSUPPORTED = -3, /* valid reference config this build does not cover yet */
  METRPO_ERR_NOMEM = -4,
  METRPO_ERR_STATE = -5        /* call order (e.g. run before weights were set) */
} metrpo_status_t;

/* envs whose analytic cost / done the kernel fuses (envs/com_*_env.py cost_np_vec / is_done) */
enum {
  METRPO_ENV_SWIMMER = 0,      /* com_swimmer_env.py:112-114 */
  METRPO_ENV_HALF_CHEETAH = 1, /* com_half_cheetah_env.py:72-75 */
  METRPO_ENV_HOPPER = 2,       /* com_hopper_env.py:94-104 */
  METRPO_ENV_ANT = 3,          /* com_ant_env.py:77-101 (cost + is_done) */
  METRPO_ENV_HUMANOID = 4,     /* com_simple_humanoid_env.py:105-109 */
  METRPO_ENV_SNAKE = 5         /* com_snake_env.py:81-84 */
};

/* VecSimpleEnv.get_next_observation sam_mode (env_helpers.py:617-634) */
enum {
  METRPO_SAM_STEP_RAND = 0,     /* per-row random model every step */
  METRPO_SAM_EPS_RAND = 1,      /* per-episode model (cur_model_idx) */
  METRPO_SAM_MODEL_MEAN_STD = 2,
  METRPO_SAM_MODEL_MEAN = 3,
  METRPO_SAM_MODEL_MED = 4,
  METRPO_SAM_ONE_MODEL = 5
};

/* arithmetic of the dynamics MLP:
 *   METRPO_PREC_BF16  tcgen05 tensor cores, bf16 operands, fp32 accumulation and epilogues (fast path)
 *   METRPO_PREC_FP32  fp32 FMA on CUDA cores with the reference's operation order (true division by
 *                     in_std, tanhf): the reference's tf.matmul arithmetic (training.py:207-208), ~50x
 *                     slower; a fidelity mode to measure what the bf16 operands cost (all sam_modes,
 *                     run / continue / step / model_costs). */
enum { METRPO_PREC_BF16 = 0, METRPO_PREC_FP32 = 1 };

#define METRPO_MAX_POLICY_LAYERS 4

typedef struct {
  int32_t state_dim;        /* S */
  int32_t action_dim;       /* A */
  int32_t drop_cols;        /* 0 none, 1 ignore_x_input, 2 ignore_xy_input (training.py:146-154) */
  int32_t hidden;           /* dynamics hidden width; both hidden layers (params/*.json) */
  int32_t n_models;         /* K (params "n_models") */
  int32_t n_envs;           /* B parallel imaginary envs (rows) */
  int32_t max_path_length;  /* T: timeout horizon (env_helpers.py:604) */
  int32_t env_id;           /* METRPO_ENV_* */
  int32_t sam_mode;         /* METRPO_SAM_* */
  int32_t n_policy_layers;  /* number of policy weight matrices (hidden layers + 1) */
  int32_t policy_dims[METRPO_MAX_POLICY_LAYERS + 1]; /* S, h1, .., A */
  int32_t policy_out_tanh;  /* policy output_nonlinearity: 0 tf.identity, 1 tf.tanh */
  int32_t precision;        /* METRPO_PREC_* */
  int32_t device;           /* CUDA device ordinal */
  int32_t row_offset;       /* global index of row 0 (Philox streams are keyed by global row, so a
                               row-sharded multi-GPU run reproduces the single-GPU one) */
} metrpo_rollout_cfg;

typedef struct metrpo_rollout metrpo_rollout_t;

/* library / build info string (static storage) */
const char* metrpo_version(void);
/* message of the last failing call on this thread (static thread-local storage) */
const char* metrpo_last_error(void);

/* Handle for one imaginary vec-env = one NeuralNetEnv + VecSimpleEnv pair
 * (env_helpers.py:532-544, 575-583). */
int metrpo_rollout_create(const metrpo_rollout_cfg* cfg, metrpo_rollout_t** out);
int metrpo_rollout_destroy(metrpo_rollout_t* h);

/* Stage model k of the ensemble: the weights of build_ff_neural_net (training.py:171-214),
 * fp32, TF layout.  W0[Din,H] b0[H] W1[H,H] b1[H] W2[H,S] b2[S], Din = S + A - drop_cols.
 * Packs them into the bf16 tile stream the persistent kernel bulk-copies every step. */
int metrpo_rollout_set_dynamics(metrpo_rollout_t* h, int k, const float* W0, const float* b0,
                                const float* W1, const float* b1, const float* W2,
                                const float* b2, void* stream);

/* RunningMeanStd constants (running_mean_std.py:22-27) used by dynamics_model
 * (training.py:228,257): in_mean/in_std [S+A], diff_mean/diff_std [S]. */
int metrpo_rollout_set_normalization(metrpo_rollout_t* h, const float* in_mean,
                                     const float* in_std, const float* diff_mean,
                                     const float* diff_std, void* stream);

/* Gaussian MLP policy mean network + log_std (training.py:96-117; rllab GaussianMLPPolicy).
 * W[i] is [policy_dims[i], policy_dims[i+1]], b[i] is [policy_dims[i+1]]; W, b are HOST arrays
 * of n_policy_layers DEVICE pointers; log_std is a device vector [A]. */
int metrpo_rollout_set_policy(metrpo_rollout_t* h, const float* const* W, const float* const* b,
                              const float* log_std, void* stream);

/* VecSimpleEnv.reset() with explicit states (env_helpers.py:585-595): states[B,S] become the
 * current observations, ts = 0.  The reference draws them from the real simulator. */
int metrpo_rollout_reset(metrpo_rollout_t* h, const float* states, void* stream);
/* Overwrite the states of `n` rows (device int32 row indices, device states [n,S]) WITHOUT touching
 * their step counters: how the step-granular socket hands the done rows the simulator resets it
 * drew in row order after a step (VecSimpleEnv.reset(dones), env_helpers.py:585-595). */
int metrpo_rollout_set_rows(metrpo_rollout_t* h, const int32_t* rows, int n, const float* states, void* stream);

/* VecSimpleEnv.step (env_helpers.py:597-607), socket B1.  actions[B,A] are the sampler's
 * UNCLIPPED actions; the kernel clips to [-1,1] (:599), evaluates all K models (:612-615),
 * selects per sam_mode, computes reward = -cost_np_vec (:601), done = is_done | ts >= T (:603-604)
 * and replaces done rows by reset_states[row] (:605-606).
 *   model_idx  [B] int32 or NULL: step_rand / eps_rand choice per row (NULL -> Philox(seed,offset))
 *   std_noise  [B,S] or NULL: N(0,1) draws for model_mean_std (:626)
 *   reset_states [B,S]: row i takes reset_states[i] if it finishes this step
 *   obs_out [B,S] post-reset observations, rew_out [B], done_out [B] uint8. */
int metrpo_rollout_step(metrpo_rollout_t* h, const float* actions, const int32_t* model_idx,
                        const float* std_noise, const float* reset_states, uint64_t seed,
                        uint64_t offset, float* obs_out, float* rew_out, uint8_t* done_out,
                        void* stream);

/* VectorizedSampler.obtain_samples (samplers/vectorized_sampler.py:45-116) for n_steps steps of
 * all B rows, socket B2: one persistent kernel runs policy -> K dynamics -> select -> reward /
 * done / reset -> trajectory write for the whole horizon with no host round trip.
 *   init_states [B,S]   observations after the initial vec_env.reset() (:49)
 *   reset_pool  [R,S]   pre-sampled real-env reset states; the n-th reset (n = 0,1,..) of row i
 *                       takes reset_pool[(n*B + i) % R]  (reference order when all rows time out
 *                       together; see DESIGN.md for Ant)
 *   eps         [n_steps,B,A] N(0,1) policy noise or NULL -> Philox(seed, offset) on device
 *   model_idx   [n_steps,B] int32 or NULL -> Philox
 *   std_noise   [n_steps,B,S] or NULL (model_mean_std only; NULL -> Philox)
 *   determ      1 -> actions = mean (obtain_samples(determ=True), :64-65)
 * outputs (any may be NULL to skip the write), time-major:
 *   obs [n_steps,B,S] pre-step observations (:91), act [n_steps,B,A] UNCLIPPED actions (:92),
 *   mean [n_steps,B,A] agent_infos['mean'], rew [n_steps,B], done [n_steps,B] uint8,
 *   final_states [B,S] observations after the last step (post-reset). */
int metrpo_rollout_run(metrpo_rollout_t* h, int n_steps, const float* init_states,
                       const float* reset_pool, int R, const float* eps,
                       const int32_t* model_idx, const float* std_noise, uint64_t seed,
                       uint64_t offset, int determ, float* obs, float* act, float* mean,
                       float* rew, uint8_t* done, float* final_states, void* stream);

/* Continue the rollout of metrpo_rollout_run (or of reset() / step()) for n_steps more steps from
 * the row state the previous launch left behind (observations, time-in-path, resets consumed);
 * identical to one longer run() when `offset` (global step index of the first step: keys the
 * Philox streams) and the eps / model_idx / std_noise / output pointers are advanced by the caller.
 * Lets a host overlap the device->host copy of finished steps with the remaining horizon. */
int metrpo_rollout_continue(metrpo_rollout_t* h, int n_steps, const float* reset_pool, int R,
                            const float* eps, const int32_t* model_idx, const float* std_noise,
                            uint64_t seed, uint64_t offset, int determ, float* obs, float* act,
                            float* mean, float* rew, uint8_t* done, float* final_states, void* stream);

/* Per-model validation cost of the current policy, the `policy_costs` tensors of
 * build_policy_graph (model_based_rl.py:122-142) that optimize_policy evaluates every log_every
 * iterations on the fixed validation initial states (:1237-1248) to feed the stop criterion
 * (utils.py:285-296).  Each of the K models rolls ITS OWN prediction forward for n_steps steps from
 * init_states [n_rows,S] (n_rows <= n_envs) under the deterministic, clipped policy (stochastic = 0,
 * :130); no model selection, no reset, no timeout.  Same persistent kernel as metrpo_rollout_run.
 *   cost_k = sum_t gamma^t * mean_rows cost_tf(x, u, x')   (Ant: rows that already hit is_done_tf
 *            contribute 0, envs/com_ant_env.py:70-75,103-117)
 *   row_costs   [K,n_rows] per-(model,row) discounted sums, or NULL (library workspace)
 *   model_costs [K] floats. */
int metrpo_rollout_model_costs(metrpo_rollout_t* h, int n_steps, int n_rows,
                               const float* init_states, double gamma, float* row_costs,
                               float* model_costs, void* stream);

/* Synchronise the stream and report the outcome of the last launch: METRPO_OK, or METRPO_ERR_STATE
 * with a message naming the stalled barrier if the kernel's bounded waits timed out (the kernel
 * aborts itself instead of hanging the GPU). */
int metrpo_rollout_status(metrpo_rollout_t* h, void* stream);

/* Dev tools: event trace of one CTA (clock64 stamps of its producer / MMA / epilogue roles) for
 * its local steps [t0,t1); out_host receives 3 x 4096 uint64 (code << 40 | clock). */
int metrpo_rollout_set_trace(metrpo_rollout_t* h, int cta, int t0, int t1);
int metrpo_rollout_get_trace(metrpo_rollout_t* h, unsigned long long* out_host);

/* number of kernels the last run()/step() call launched on the stream (bench gpu_launches) */
int metrpo_rollout_last_launches(const metrpo_rollout_t* h);
/* which kernel the last run()/continue() call used: 0 = single-stream rollout_kernel, 1 = two-stream
 * kernel (rollout_duo.cuh) without column split, 2 = two-stream kernel with the hidden dimension
 * split over CTA pairs.  Selection is automatic (env METRPO_DUO=0/1/2 overrides, dev switch). */
int metrpo_rollout_last_kernel(const metrpo_rollout_t* h);

/* Host-only helper (no CUDA call): the gang schedule the library builds for n_tiles row tiles of
 * 128 rows on n_slots gang slots over T steps; out receives n_slots * max_seg quadruples
 * (tile, t0, t1, wait_flag), tile == -1 for unused entries.  Returns max_seg (> 0) or a status. */
int metrpo_debug_schedule(int n_tiles, int n_slots, int T, int32_t* out, int out_capacity_quads);

/* =============================================================================================
 * TRPO half of the inner iteration: sample processing + natural-gradient policy update.
 * All buffers are DEVICE pointers unless marked host; sample index n = t * B + b (the time-major
 * layout metrpo_rollout_run writes), so the trajectory buffers are consumed in place.
 * ============================================================================================= */
typedef struct {
  int32_t state_dim;        /* S */
  int32_t action_dim;       /* A (<= 24) */
  int32_t n_policy_layers;  /* weight matrices of the mean network (training.py:96-103) */
  int32_t policy_dims[METRPO_MAX_POLICY_LAYERS + 1]; /* S, h1, .., A */
  int32_t policy_out_tanh;  /* output_nonlinearity: 0 identity, 1 tanh (training.py:82) */
  int32_t device;
} metrpo_trpo_cfg;

typedef struct metrpo_trpo metrpo_trpo_t;

/* Cross-rank SUM of n doubles at dev_buf, in place, ordered on `stream` (e.g. an NCCL all-reduce
 * issued by the host framework).  Returns 0 on success.  Without a callback the update is
 * single-GPU.  Every reduction is <= P + 4 doubles (latency-bound collectives). */
typedef int (*metrpo_allreduce_fn)(void* user, double* dev_buf, int n, void* stream);

int metrpo_trpo_create(const metrpo_trpo_cfg* cfg, metrpo_trpo_t** out);
int metrpo_trpo_destroy(metrpo_trpo_t* h);
int metrpo_trpo_set_allreduce(metrpo_trpo_t* h, metrpo_allreduce_fn fn, void* user);
/* In-library all-reduce over NVLink / NVSwitch peer memory (one process per GPU on ONE node): every
 * rank exposes a small exchange buffer through CUDA IPC; a reduction is then ONE launch of a one-shot
 * kernel -- publish the local accumulator, flag the peers (system-scope release), wait for their
 * flags, sum all ranks' copies with peer loads in rank order (bitwise identical result on every
 * rank) -- instead of a host callback into NCCL per reduction.
 *   metrpo_trpo_p2p_handle  allocates the exchange buffer and writes its 64-byte IPC handle (host)
 *   metrpo_trpo_enable_p2p  handles: world x 64 bytes in rank order (gathered by the host framework);
 *                           takes precedence over a callback set with metrpo_trpo_set_allreduce
 * Every rank must issue the same sequence of trpo calls (they do: the update's control flow does
 * not depend on data). */
#define METRPO_IPC_HANDLE_BYTES 64
int metrpo_trpo_p2p_handle(metrpo_trpo_t* h, void* handle_out);
int metrpo_trpo_enable_p2p(metrpo_trpo_t* h, int rank, int world, const void* handles);
/* Implementation of the per-sample pass (loss / gradient / Fisher-vector product):
 *   AUTO    the fastest measured one for the shape: the register-tiled fp32 pass (csrc/trpo_tiled.cuh)
 *           for policies whose layers are all <= 32 wide (every shipped params/*.json but
 *           humanoid), else SIMT
 *   SIMT    fp32 FMA on CUDA cores, one thread per sample (any shape)
 *   TF32    warp-level tensor-core MMAs, single TF32 products (10-bit operand mantissa); policies
 *           with <= 3 weight layers of width <= 32 (all shipped params/*.json but humanoid)
 *   TF32X3  same with 3xTF32 split products (fp32-equivalent accuracy) */
enum { METRPO_TRPO_PASS_AUTO = 0, METRPO_TRPO_PASS_SIMT = 1, METRPO_TRPO_PASS_TF32 = 2,
       METRPO_TRPO_PASS_TF32X3 = 3 };
int metrpo_trpo_set_pass_impl(metrpo_trpo_t* h, int impl);
/* length P of the flat parameter vector, rllab get_params(trainable=True) order:
 * W0[in,out], b0, W1, b1, .., log_std[A]   */
int metrpo_trpo_num_params(const metrpo_trpo_t* h);
int metrpo_trpo_last_launches(const metrpo_trpo_t* h);

/* BaseSampler.process_samples (samplers/base.py:48-105) on flat buffers obs[T,B,S], rew[T,B],
 * done[T,B]: baseline prediction (coeffs [2S+4] doubles, NULL = not fitted yet -> zeros,
 * :55), deltas (:57-59), advantages = discount_cumsum(deltas, discount*gae_lambda) (:60-61),
 * returns (:62), center_advantages (:82-83), shift_advantages_to_positive (:85-86).
 * Samples of paths still open at the end of the buffer get valid = 0 and adv = ret = 0
 * (obtain_samples returns completed paths only, samplers/vectorized_sampler.py:80-105).
 * stats (8 doubles): n_valid, sum, sum of squares, min, mean, std of the raw advantages. */
int metrpo_trpo_process(metrpo_trpo_t* h, int T, int B, const float* obs, const float* rew,
                        const uint8_t* done, const double* baseline_coeffs, double discount,
                        double gae_lambda, int center_adv, int positive_adv, float* adv, float* ret,
                        uint8_t* valid, double* stats, void* stream);

/* rllab LinearFeatureBaseline.fit (samplers/base.py:167): ridge regression of the
 * returns on [o, o^2, t/100, (t/100)^2, (t/100)^3, 1], o = clip(obs,-10,10), t = index in path;
 * coeffs_out [2S+4] doubles; reg retried x10 up to 5 times while the solution has NaNs. */
int metrpo_trpo_fit_baseline(metrpo_trpo_t* h, int T, int B, const float* obs, const float* ret,
                             const uint8_t* valid, const uint8_t* done, double reg_coeff,
                             double* coeffs_out, void* stream);

/* NPO.optimize_policy -> ConjugateGradientOptimizer.optimize (algos/npo.py:94-111)
 * over N samples: obs[N,S], act[N,A] (unclipped actions), adv[N], old_mean[N,A], old_log_std
 * ([A], or [N,A] when old_log_std_per_sample), valid[N] or NULL.  theta [P] floats is updated in
 * place: loss_before, flat gradient, cg_iters CG iterations on Hx = Fisher-vector product +
 * reg_coeff x, step = sqrt(2 step_size / (d.Hd + 1e-8)), back-tracking over
 * backtrack_ratio ** arange(max_backtracks) accepting the first trial with loss < loss_before and
 * mean_kl <= step_size, else the previous parameters are restored.  No host synchronisation.
 * info (8 doubles, may be NULL): loss_before, loss_after, mean_kl, backtrack index, accepted,
 * initial step, n_valid, d.Hd. */
int metrpo_trpo_update(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                       const float* adv, const float* old_mean, const float* old_log_std,
                       int old_log_std_per_sample, const uint8_t* valid, float* theta,
                       double step_size, int cg_iters, double reg_coeff, double backtrack_ratio,
                       int max_backtracks, double* info, void* stream);

/* optimizer.loss / optimizer.constraint_val (algos/npo.py:108-114): out2 (HOST, 2 doubles) =
 * (surr_loss, mean_kl) at theta.  Synchronises the stream. */
int metrpo_trpo_loss_kl(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                        const float* adv, const float* old_mean, const float* old_log_std,
                        int old_log_std_per_sample, const uint8_t* valid, const float* theta,
                        double* out2, void* stream);

/* Test hook: vec == NULL -> flat gradient of surr_loss at theta; else Hx(vec) = Fisher-vector
 * product + reg_coeff * vec.  out_host: P doubles (HOST).  Synchronises the stream. */
int metrpo_trpo_grad(metrpo_trpo_t* h, long long N, const float* obs, const float* act,
                     const float* adv, const float* old_mean, const float* old_log_std,
                     int old_log_std_per_sample, const uint8_t* valid, const float* theta,
                     const float* vec, double reg_coeff, double* out_host, void* stream);

/* =============================================================================================
 * Ensemble dynamics fit: optimize_models (model_based_rl.py:881-1051) with the
 * graph of build_dynamics_graph (:23-103) and get_dynamics_optimizer (:154-183).  All K models
 * train on independent minibatches in one stream of kernels; weights, Adam moments and the
 * per-model best snapshots live on the device in fp32.  x rows are [state, action] (S + A),
 * y rows are next states (S) -- the data_collection layout (utils.py:44-131).
 * ============================================================================================= */
enum { METRPO_FIT_TF32 = 0,   /* every contraction on the tcgen05 tensor cores through the library's own batched
                                 TF32 GEMM (csrc/fit_gemm.cuh), fp32 accumulate (default) */
       METRPO_FIT_FP32 = 1 }; /* fidelity mode: true fp32 products (the reference's tf.matmul arithmetic; cuBLAS) */

typedef struct {
  int32_t state_dim;   /* S (<= 64) */
  int32_t action_dim;  /* A */
  int32_t drop_cols;   /* ignore_x_input / ignore_xy_input (training.py:146-154) */
  int32_t hidden;      /* both hidden layers (multiple of 32) */
  int32_t n_models;    /* K (<= 64) */
  int32_t max_rows;    /* row capacity: >= dynamics_opt_params.batch_size; validation sets are
                          processed in chunks of this many rows */
  int32_t precision;   /* METRPO_FIT_* */
  int32_t device;
} metrpo_fit_cfg;

typedef struct metrpo_fit metrpo_fit_t;

int metrpo_fit_create(const metrpo_fit_cfg* cfg, metrpo_fit_t** out);
int metrpo_fit_destroy(metrpo_fit_t* h);
int metrpo_fit_num_params(const metrpo_fit_t* h);   /* trainable floats per model */
int metrpo_fit_last_launches(const metrpo_fit_t* h);

/* weights of model k in the TF layout W[in,out] (training.py:187-208); device fp32 */
int metrpo_fit_set_weights(metrpo_fit_t* h, int k, const float* W0, const float* b0, const float* W1,
                           const float* b1, const float* W2, const float* b2, void* stream);
int metrpo_fit_get_weights(metrpo_fit_t* h, int k, float* W0, float* b0, float* W1, float* b1,
                           float* W2, float* b2, void* stream);
/* RunningMeanStd constants used inside dynamics_model (training.py:228,257) */
int metrpo_fit_set_normalization(metrpo_fit_t* h, const float* in_mean, const float* in_std,
                                 const float* diff_mean, const float* diff_std, void* stream);
/* sess.run(dynamics_adam_init) (model_based_rl.py:906-918): zero moments, step count 0 */
int metrpo_fit_reset_adam(metrpo_fit_t* h, void* stream);

/* One training iteration (model_based_rl.py:957-970): row r of model k's minibatch is sample
 * idx[r*K + k] of (x[n_data,S+A], y[n_data,S]) -- np.reshape(x_batch, (batch, -1)) +
 * get_ith_tensor (utils.py:366-369); idx == NULL draws with replacement from Philox(seed, offset)
 * (data_collection.sample, utils.py:129-131).  Loss of model k = mean_rows sum_s (pred - y)^2 on
 * the de-normalised next state (:57-71); Adam(lr, 0.9, 0.999, 1e-8) in TF's formulation.
 * losses [K] device floats (pre-update training losses) or NULL. */
int metrpo_fit_step(metrpo_fit_t* h, const float* x, const float* y, int n_data, const int32_t* idx,
                    int batch, uint64_t seed, uint64_t offset, double lr, float* losses, void* stream);

/* Per-model loss on (x[n,S+A], y[n,S]) evaluated by ALL K models (np.tile, :934-935) ->
 * losses [K] device floats.  snapshot 0: nothing; 1: models whose loss is below their recorded
 * minimum are saved and the minimum updated (:998-1007); 2: save all and initialise the minima
 * (:925-946).  improved [K] device bytes or NULL.  No host synchronisation. */
int metrpo_fit_eval(metrpo_fit_t* h, const float* x, const float* y, int n, int snapshot, float* losses,
                    uint8_t* improved, void* stream);
/* recover_weights (:876-879,1034): every model back to its best snapshot */
int metrpo_fit_restore_best(metrpo_fit_t* h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* METRPO_H_ */
