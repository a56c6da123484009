/*
 * catb200.h - C ABI of libcatb200.so: the constraints-as-terminations PPO hot path on B200 (sm_100a).
 *
 * The reference (Gepetto/constraints-as-terminations) is pure Python/PyTorch: it has no FFI.  The
 * boundary this library plugs into is therefore its Python manager / trainer API; each entry point
 * below names the reference code it replaces (paths relative to the reference root, with
 * U/ = exts/cat_envs/cat_envs/tasks/utils/).  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add (it is the one `constraints_as_terminations_b200/_lib.py` ships).
 *
 * Conventions
 *  - plain C types only: device pointers, sizes, small POD structs passed by pointer (host memory).
 *  - every function returns CATB200_OK (0) or a negative catb200_status; nothing throws.
 *  - no hidden allocation and no hidden synchronisation: persistent state (running max, episode
 *    statistics, running moments, Adam moments) lives in caller-owned device buffers; scratch comes
 *    from a caller-provided workspace whose size `*_workspace_bytes` reports.  The workspace must be
 *    zero-initialised once (cudaMemset) before its first use; kernels leave it clean for the next call.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are CUDA-graph capturable.
 *  - fp32 arithmetic on the CaT / GAE / moments paths is IEEE round-to-nearest with the reference's
 *    operation order (no FMA contraction where torch rounds twice), so results are bit-comparable.
 */
#ifndef CATB200_H_
#define CATB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATB200_VERSION 100

typedef enum {
  CATB200_OK = 0,
  CATB200_ERR_INVALID_ARGUMENT = -1,
  CATB200_ERR_UNSUPPORTED = -2,
  CATB200_ERR_WORKSPACE_TOO_SMALL = -3,
  CATB200_ERR_CUDA = -4
} catb200_status;

int catb200_version(void);
/* Number of kernels this library has launched in this process (host-side count; launches captured
 * into a CUDA graph are counted once at capture, replays are the caller's to multiply). */
uint64_t catb200_launch_count(void);
/* Static description of a status code; for CATB200_ERR_CUDA also the last CUDA error string. */
const char* catb200_error_string(int status);

/* ------------------------------------------------------------------------------------------------
 * Per-step path: constraint terms -> termination probabilities  (rows a1, a2, a3, a5 of SURVEY.md §8)
 * ---------------------------------------------------------------------------------------------- */
#define CATB200_MAX_SOURCES 16
#define CATB200_MAX_TERMS 32
#define CATB200_MAX_IDS 32
#define CATB200_MAX_COLS 256
#define CATB200_MAX_PEAKS 32

typedef enum { CATB200_F32 = 0, CATB200_U8 = 1 } catb200_dtype;

/* One fused operation per reference term function (U/cat/constraints.py, line of the def). */
typedef enum {
  CATB200_OP_GENERIC = 0,              /* x[id]                       user term already evaluated to [N,J]       */
  CATB200_OP_ABS_MINUS = 1,            /* |x[id]| - p0                joint_position:23 joint_torque:57
                                                                      joint_velocity:68 joint_acceleration:78   */
  CATB200_OP_ABSDIFF_MINUS = 2,        /* |x[id]-y[id]| - p0          joint_range:171                            */
  CATB200_OP_ABSDIFF_MINUS_GATE_Y = 3, /* (|x-y| - p0)*[|cmd.y|<p1]   joint_position_when_moving_forward:34      */
  CATB200_OP_ACTION_RATE = 4,          /* |a-a_prev|/p1 - p0          action_rate:184  (p1 = step_dt, true div)  */
  CATB200_OP_COMPONENT_GT = 5,         /* x[id] > p0  (bool)          upsidedown:88   (id = 2)                   */
  CATB200_OP_CONTACT_ANY = 6,          /* any_b max_h|F[h,b]| > p0    contact:97      (p0 = 1.0; bool, 1 column) */
  CATB200_OP_NORM2_MINUS = 7,          /* |x[0:2]| - p0               base_orientation:113                       */
  CATB200_OP_AIR_TIME = 8,             /* (p0-air[b])*first[b]*[|cmd|>p1]  air_time:122                          */
  CATB200_OP_N_CONTACT = 9,            /* |#b(max_h|F|>1) - p0|*[|cmd|>p1] n_foot_contact:144  (1 column)        */
  CATB200_OP_FORCE_PEAK_MINUS = 10,    /* max_h|F[h,b]| - p0          foot_contact_force:201                     */
  CATB200_OP_LIMIT_MINUS = 11,         /* p0 - x[id]                  min_base_height:214 (id = 2)               */
  CATB200_OP_ABS_MINUS_GATE_STILL = 12 /* (|x[id]|-p0)*[|cmd|<p1]     no_move:223                                */
} catb200_op;

/* A device tensor the terms read: env-major rows, `row_len` elements used per env. */
typedef struct {
  const void* ptr;    /* device pointer to element [0,0]                                  */
  int32_t row_len;    /* elements per env row (e.g. 12 joints, 3*17*3 contact history)   */
  int32_t row_stride; /* elements between consecutive env rows (== row_len if contiguous) */
  int32_t dtype;      /* catb200_dtype                                                    */
  int32_t aux;        /* contact history: number of bodies B (row = [H, B, 3]); else 0    */
  int32_t smem_off;   /* filled by catb200_cat_plan_finalize: byte offset in the tile's smem image */
  uint32_t magic;     /* reserved                                                         */
} catb200_source_t;

typedef struct {
  uint8_t op;        /* catb200_op                                                          */
  uint8_t n_cols;    /* columns this term contributes to the [N,K] constraint matrix        */
  uint8_t n_ids;     /* joint / body ids in `ids` (== n_cols for per-joint ops)             */
  uint8_t src0;      /* main source index                                                   */
  uint8_t src1;      /* second source (y, a_prev, first_contact) or 0xff                    */
  uint8_t src2;      /* command source or 0xff                                              */
  uint8_t stat_slot; /* row of episode_sums / mean_values this term accumulates into        */
  uint8_t reserved;  /* must be 0 on entry to finalize (marks contact terms whose ids were remapped) */
  uint16_t col_offset; /* first column of this term in the [N,K] layout                     */
  uint16_t reserved2;  /* filled by finalize: entry k = index of the term with the k-th largest evaluation cost */
  float p0, p1, p2;
  uint8_t ids[CATB200_MAX_IDS]; /* element / body indices into the source row: 8-bit, i.e. rows of at most 256 entries */
} catb200_term_t;

typedef struct {
  int32_t n_sources, n_terms, n_cols, n_slots;
  int32_t smem_bytes;    /* filled by finalize: dynamic shared memory of one 32-env tile     */
  int32_t n_peaks;       /* filled by finalize: distinct (contact history, body) pairs       */
  int32_t smem_peak_off; /* filled by finalize                                               */
  int32_t smem_ctile_off; /* filled by finalize: [n_cols][32] constraint tile inside the image */
  catb200_source_t sources[CATB200_MAX_SOURCES];
  catb200_term_t terms[CATB200_MAX_TERMS];
  uint8_t col_term[CATB200_MAX_COLS];               /* filled by finalize: term of each column        */
  uint16_t slot_col_begin[CATB200_MAX_TERMS + 2];   /* filled by finalize: [begin,end) cols per slot  */
  uint8_t peak_src[CATB200_MAX_PEAKS];              /* filled by finalize: source of each peak slot    */
  uint8_t peak_body[CATB200_MAX_PEAKS];             /* filled by finalize: body of each peak slot      */
} catb200_plan_t;

/* Scalars of CaT.add (U/cat/constraint_manager.py:39-76), rounded on the host exactly like torch
 * rounds python doubles that meet fp32 tensors. */
typedef struct {
  float tau;           /* fl32(tau)                                          :59 */
  float one_minus_tau; /* fl32(1.0 - tau) evaluated in double first          :59 */
  float min_p;         /* fl32(min_p)                                        :72 */
  float floor_max;     /* fl32(1e-6) clamp of the column max                 :55 */
  float span[CATB200_MAX_TERMS]; /* fl32(max_p - min_p) per term (double subtraction)  :72 */
  /* If not NULL: device array of CATB200_MAX_TERMS floats that replaces span[] -- the curriculum changes max_p at run
   * time (U/cat/curriculums.py:37-41); with the table in device memory a CUDA graph of the step keeps seeing the
   * current values (the host refreshes it with an asynchronous copy), whereas by-value arguments are frozen in it. */
  const float* span_dev;
} catb200_cat_params_t;

/* Validates a plan and fills the derived fields (shared-memory layout, division magics). */
int catb200_cat_plan_finalize(catb200_plan_t* plan);

/* Bytes of zero-initialised device scratch needed by catb200_cat_step for this problem size. */
size_t catb200_cat_workspace_bytes(int32_t num_envs, int32_t n_cols);

/*
 * One env step of ConstraintManager.compute() (U/cat/constraint_manager.py:213-229), i.e. for every
 * term: evaluate the constraint (constraints.py), CaT.add (:39-76: column max over envs, clamp,
 * Polyak running max, violation -> probability), then CaT.get_probs (:78-82: max over all columns)
 * and the per-term episode statistics (:223-227).  Optionally also the reward / dones lines of
 * CaTEnv.step (U/cat/cat_env.py:102-107,118-121) when raw_reward != NULL.
 *
 *   running_max  [n_cols]          fp32, in/out     rm_init [n_cols] int32 in/out (0 = first call assigns)
 *   episode_sums [n_slots,num_envs] fp32, in/out    mean_values [n_slots,num_envs] fp32, in/out
 *   cstr_prob    [num_envs]        fp32, out
 *   raw_reward   [num_envs] fp32 in (nullable)      reset_buf [num_envs] u8/bool in (nullable)
 *   reward_out   [num_envs] fp32 out                dones_out [num_envs] fp32 out
 */
int catb200_cat_step(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                     float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                     float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf,
                     float* reward_out, float* dones_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/*
 * catb200_cat_step followed by ConstraintManager.reset(ids of the envs whose reset_buf is set) in the same two
 * launches -- the order of CaTEnv.step (U/cat/cat_env.py:100 compute, :181 constraint_manager.reset inside
 * _reset_idx; U/cat/constraint_manager.py:190-211): the envs being reset contribute their just-updated statistics
 * divided by episode_length[i] (int64) to the per-term means, are zeroed, and reset_out[2*s], reset_out[2*s+1]
 * receive the same values catb200_cat_reset_stats would write (NaN when no env resets).  reset_buf is required.
 * reset_workspace: catb200_cat_reset_workspace_bytes() of zero-initialised device scratch.
 */
int catb200_cat_step_reset(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                           float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                           float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf,
                           float* reward_out, float* dones_out, void* workspace, size_t workspace_bytes,
                           const int64_t* episode_length, float* reset_out, void* reset_workspace,
                           size_t reset_workspace_bytes, void* stream);

/* Raw constraint values of every term, row-major [num_envs, n_cols] (what CaT keeps as
 * raw_constraints, constraint_manager.py:52; also backs the stand-alone term functions). */
int catb200_cat_eval_terms(const catb200_plan_t* plan, int32_t num_envs, float* out, void* stream);

/* Per-column probabilities [num_envs, n_cols] (what CaT keeps as `probs`, constraint_manager.py:64-74) of the raw
 * constraint values `raw` (row-major [num_envs, n_cols], from catb200_cat_eval_terms) under `running_max`. */
int catb200_cat_probs(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                      const float* running_max, float* probs_out, const float* raw, void* stream);

/*
 * ConstraintManager.reset (U/cat/constraint_manager.py:190-211): for each statistics row s,
 *   out[2*s]   = mean_i(episode_sums[s,i] / episode_length[i]) * 100
 *   out[2*s+1] = mean_i(mean_values[s,i]  / episode_length[i])
 * over the selected envs, then zero both rows at those envs.  Selection: `env_ids` (int64, n_ids
 * entries) or, if env_ids == NULL, `mask` (u8 [num_envs], nonzero = selected) or, if both NULL, all.
 * workspace: catb200_cat_reset_workspace_bytes() of zero-initialised device scratch.
 */
size_t catb200_cat_reset_workspace_bytes(void);
int catb200_cat_reset_stats(const int64_t* env_ids, int32_t n_ids, const uint8_t* mask,
                            const int64_t* episode_length, int32_t num_envs, int32_t n_slots,
                            float* episode_sums, float* mean_values, float* out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Running moments + rollout append  (row a6: U/cleanrl/ppo.py:12-62,203-205,225)
 * ---------------------------------------------------------------------------------------------- */
size_t catb200_rms_workspace_bytes(int32_t dim);

/*
 * RunningMeanStd.forward(x, update=True) for x [rows, dim] (dim = 1 for scalars):
 * batch mean / biased variance, Chan merge into (mean[dim], var[dim], count[1]) and
 * out = (x - mean) / sqrt(var + eps) with the *updated* statistics.  `out` may alias x or be NULL
 * (update only).  update == 0 skips the statistics and only normalises.  If out_op != NULL the normalised
 * rows are also written as the zero-padded tensor-core operand [rows, pad_op] (dim <= pad_op <= 2 * dim) of
 * precision op_prec (catb200_prec: bf16, or fp32 rounded to tf32), the layout the first MLP layer reads, so the
 * rollout needs no separate conversion pass.
 */
int catb200_rms_forward(const float* x, int64_t rows, int32_t dim, float* mean, float* var, float* count,
                        float eps, int32_t update, float* out, void* out_op, int32_t pad_op, int32_t op_prec,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Post-step rollout append (U/cleanrl/ppo.py:203-205,215-216,225 for step t):
 *   rewards[t] = reward; dones[t+1] = done; true_dones[t+1] = float(time_out)
 * where dones / true_dones have T+1 time slots so that slot t+1 doubles as next_done / next_true_done.
 */
int catb200_rollout_append(const float* reward, const float* done, const uint8_t* time_out, int32_t num_envs,
                           float* rewards_t, float* dones_t1, float* true_dones_t1, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAE + value normalisation statistics  (rows a7, a8: U/cleanrl/ppo.py:251-277,287-288)
 * ---------------------------------------------------------------------------------------------- */
size_t catb200_gae_workspace_bytes(void);

/*
 * rewards, values: [T, N]; dones, true_dones: [T+1, N] (slot T = next_done / next_true_done);
 * next_value: [N].  Writes advantages, returns: [T, N].  gamma_lambda = fl32(gamma * lambda) with the
 * product taken in double, as python does before it meets the tensor (ppo.py:271-273).
 * If value_rms != NULL (3 floats: mean, var, count), also performs the two consecutive
 * RunningMeanStd updates of ppo.py:287-288 (first with all values, then with all returns) and writes
 * norm_stats[4] = {mean_after_values, var_after_values, mean_after_returns, var_after_returns}, the
 * statistics b_values resp. b_returns are normalised with.
 */
int catb200_gae(const float* rewards, const float* values, const float* dones, const float* true_dones,
                const float* next_value, int32_t T, int32_t num_envs, float gamma, float gamma_lambda,
                float* advantages, float* returns, float* value_rms, float* norm_stats, void* workspace,
                size_t workspace_bytes, void* stream);

/*
 * GAE for the trainers that keep a single float `dones` buffer.  rewards, values: [T, N]
 * time-major; last_values: [N]; writes advantages, returns: [T, N].
 *   CATB200_GAE_RLGAMES  rl_games' A2CBase.discount_values as CaTA2CAgent.play_steps calls it
 *                        (U/rl_games/cat_common.py:96-103) on the float32 dones of CaTExperienceBuffer
 *                        (U/rl_games/cat_experience.py:27-33): dones [T, N] holds the flag observed BEFORE each
 *                        step, last_dones [N] the one after the last step; coef = fl32(gamma * tau), the product
 *                        taken in double.
 *   CATB200_GAE_SKRL     compute_gae of the CaT skrl agent (U/skrl/ppo.py:397-442, `not_dones = 1 - dones`):
 *                        dones [T, N] is the `terminated` memory tensor, last_dones unused (NULL);
 *                        coef = fl32(lambda).  normalize != 0 also applies :440,
 *                        advantages = (adv - mean) / (std + 1e-8) over all T*N entries (unbiased std).
 * workspace (only read when normalize != 0): catb200_gae_float_dones_workspace_bytes() zero-initialised bytes.
 */
typedef enum { CATB200_GAE_RLGAMES = 0, CATB200_GAE_SKRL = 1 } catb200_gae_variant;
size_t catb200_gae_float_dones_workspace_bytes(void);
int catb200_gae_float_dones(int32_t variant, const float* rewards, const float* values, const float* dones,
                            const float* last_dones, const float* last_values, int32_t T, int32_t num_envs,
                            float gamma, float coef, float* advantages, float* returns, int32_t normalize,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Actor-critic MLP + PPO-clip minibatch update  (rows a9, a10: U/cleanrl/ppo.py:71-123,294-354)
 *
 * Two separate MLPs obs -> h1 -> h2 -> h3 -> {act_dim, 1} with ELU (ppo.py:78-96) and a state-independent
 * log-std (ppo.py:97).  Master parameters, gradients and Adam moments are flat fp32 buffers laid out in
 * the reference's `agent.parameters()` order (critic, actor_mean, actor_logstd) so that the python
 * `Agent` exposes them as views with the reference's state_dict keys.  The hidden-layer GEMMs run on the
 * tensor cores (tcgen05) from compute copies of the weights that the Adam kernel refreshes, in the precision
 * dims->prec names:
 *   CATB200_PREC_TF32  fp32 storage, operands rounded to tf32, fp32 accumulation: the reference's own GPU
 *                      numerics (torch.backends.cuda.matmul.allow_tf32, scripts/clean_rl/train.py:86-87)
 *   CATB200_PREC_BF16  bf16 operands and stored activations, fp32 accumulation (half the bytes, 8-bit mantissa)
 * Heads, loss and optimizer math are fp32 either way.  "operand" below = element of that precision
 * (4 bytes / 2 bytes).
 * ---------------------------------------------------------------------------------------------- */
typedef enum { CATB200_PREC_BF16 = 0, CATB200_PREC_TF32 = 1 } catb200_prec;

typedef struct {
  int32_t obs_dim; /* 45                                  */
  int32_t act_dim; /* 12 (<= 16)                          */
  int32_t h1, h2, h3; /* 512, 256, 128: multiples of 128  */
  int32_t obs_pad; /* obs_dim rounded up to 64            */
  int32_t prec;    /* catb200_prec                        */
} catb200_mlp_dims_t;

/* Offsets (in elements) into the flat fp32 parameter vector and into the operand-precision weight-copy buffer.
 * Index 0 = critic, 1 = actor for every [2] array. */
typedef struct {
  int64_t n_params;                 /* total fp32 parameters (377241 for Solo12)                 */
  int64_t w[2][4], b[2][4];         /* weight / bias offset of layer 0..3 of net 0/1             */
  int64_t logstd;                   /* offset of actor_logstd [act_dim]                          */
  int64_t n_wc;                     /* total operand elements of the compute copies              */
  int64_t wc[2][3];                 /* W_l   [out, in_pad], l = 0..2                             */
  int64_t wtc[2][3];                /* W_l^T [in, out],     l = 1..2 (entry 0 unused = -1)       */
} catb200_mlp_layout_t;

int catb200_mlp_layout(const catb200_mlp_dims_t* dims, catb200_mlp_layout_t* layout);

/* Compute copies (W and W^T of the hidden layers, operand precision) from the fp32 master parameters. */
int catb200_mlp_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* wc, void* stream);

/* fp32 [rows, obs_dim] -> operand [rows, obs_pad] (zero padded), the layout the first GEMM reads. */
int catb200_obs_to_operand(const catb200_mlp_dims_t* dims, const float* obs, int64_t rows, void* obs_op,
                           void* stream);

/* Bytes of activation scratch for a forward (training = 0) or forward + backward (training = 1) pass
 * over `rows` samples.  The scratch must be zero-initialised once; the kernels leave what must stay zero zeroed. */
size_t catb200_mlp_workspace_bytes(const catb200_mlp_dims_t* dims, int32_t rows, int32_t training);

/*
 * Agent.get_action_and_value(x) for the rollout (ppo.py:104-119,208-212): both MLPs forward, then
 *   action = action_in                    if action_in != NULL (evaluate given actions), else
 *   action = mean + exp(logstd) * noise   noise ~ N(0,1): supplied by the caller (`noise`), or drawn on the device
 *                                         from rng_state (Philox4x32-10 + Box-Muller, element (row, j) = counter
 *                                         offset + row * act_dim + j; the offset is advanced by rows * act_dim);
 *                                         both NULL -> action = mean
 *   logprob = sum_j Normal(mean_j, std_j).log_prob(action_j),  value = critic(x)
 * obs_op: operand [rows, obs_pad].  Any of action / logprob / value / mean_out may be NULL.
 */
int catb200_mlp_act(const catb200_mlp_dims_t* dims, const void* obs_op, int32_t rows, const float* params,
                    const void* wc, const float* noise, uint64_t* rng_state, const float* action_in, float* action,
                    float* logprob, float* value, float* mean_out, void* workspace, size_t workspace_bytes,
                    void* stream);

typedef struct {
  float clip_coef, ent_coef, vf_coef;
  int32_t norm_adv, clip_vloss;
} catb200_ppo_hparams_t;

/*
 * Forward + backward of one PPO minibatch (ppo.py:298-352 up to loss.backward()):
 * gathers rows mb_inds[0..mb_rows) of the flattened rollout (obs_op_all operand [B, obs_pad], actions_all
 * [B, act_dim], logprobs_all / advantages_all / returns_all / values_all [B]), normalises advantages per
 * minibatch (unbiased std + 1e-8), normalises returns / old values / new values with norm_stats
 * (see catb200_gae), evaluates the clipped policy loss, clipped value loss and entropy bonus and
 * accumulates d loss / d params into `grads` (flat fp32, must be zero on entry; catb200_adam_step
 * leaves it zeroed).  loss_acc[8] += {pg_loss, v_loss, entropy, approx_kl, clipfrac, old_approx_kl,
 * loss, 1} for this minibatch (running sums the trainer reads once per iteration).
 * The weight gradients of the hidden layers are summed with floating-point atomics (red.global.add): their
 * last bits depend on the order in which CTAs retire.
 */
int catb200_ppo_minibatch_grad(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                               const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                               const float* logprobs_all, const float* advantages_all, const float* returns_all,
                               const float* values_all, const float* norm_stats, const float* params,
                               const void* wc, float* grads, float* loss_acc, void* workspace,
                               size_t workspace_bytes, void* stream);

/*
 * clip_grad_norm_(all params, max_grad_norm) + Adam step (ppo.py:353-354; torch.optim.Adam with
 * eps, betas, no weight decay), then zero `grads` and refresh the operand-precision weight copies.
 * grads are first multiplied by grad_scale (1 / world_size after a sum-allreduce).
 * lr_dev: device float; step_dev: device int32 step counter (incremented here).
 * opt_ws: 64 bytes of zero-initialised device scratch.
 */
int catb200_adam_step(const catb200_mlp_dims_t* dims, float* params, float* grads, float* exp_avg,
                      float* exp_avg_sq, void* wc, const float* lr_dev, int32_t* step_dev, float max_grad_norm,
                      float beta1, float beta2, float eps, float grad_scale, float* grad_norm_out, void* opt_ws,
                      void* stream);

/*
 * catb200_ppo_minibatch_grad + catb200_adam_step of ONE minibatch on one GPU (ppo.py:298-354), with the tail of the two --
 * fold of the weight-gradient accumulators into `grads`, global gradient norm, clip, Adam, operand-copy refresh -- as a
 * single launch (one CTA per SM around a grid barrier) instead of three.  Same arguments and results as the two calls
 * (`grads` zero on entry and on exit; opt_ws: the same 64 bytes of zero-initialised scratch, left clean).
 */
int catb200_ppo_minibatch_update(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                                 const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                                 const float* logprobs_all, const float* advantages_all, const float* returns_all,
                                 const float* values_all, const float* norm_stats, float* params, void* wc, float* grads,
                                 float* loss_acc, void* workspace, size_t workspace_bytes, float* exp_avg,
                                 float* exp_avg_sq, const float* lr_dev, int32_t* step_dev, float max_grad_norm,
                                 float beta1, float beta2, float eps, float grad_scale, float* grad_norm_out,
                                 void* opt_ws, void* stream);

/* The Adam + operand-copy-refresh half of catb200_adam_step alone, for callers that obtained the clip coefficient and the
 * bias corrections in opt_ws from catb200_grad_allreduce_norm (multi-GPU). */
int catb200_adam_apply(const catb200_mlp_dims_t* dims, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                       void* wc, const float* lr_dev, float beta1, float beta2, float eps, float grad_scale,
                       void* opt_ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU gradient exchange over NVLink peer memory  (one exchange step per optimizer step)
 *
 * Each rank owns one peer-visible allocation of catb200_peer_arena_bytes(n_params) bytes:
 *   [flags: 64 x uint32][gradient arena 0: n_pad floats][gradient arena 1: n_pad floats][summed gradient: n_pad floats],
 *   n_pad = n_params rounded up to 64.
 * Minibatch k accumulates its flat gradient into arena k & 1 (`grads` of catb200_ppo_minibatch_grad points there).
 * catb200_grad_allreduce_norm is then ONE kernel per rank: flag handshake with every peer over NVLink, rank-ordered sum
 * of the world arenas of that parity into the private `grad_sum` (bit-identical on every rank), squared norm -> clip
 * coefficient / Adam bias corrections in opt_ws (what catb200_adam_step's first launch does on one GPU), and zeroing of
 * the caller's other arena for the next minibatch.  No host synchronisation, no library collective: graph-capturable.
 * ---------------------------------------------------------------------------------------------- */
size_t catb200_peer_arena_bytes(int64_t n_params);
/* cudaMalloc + zero + cudaIpcGetMemHandle (64 bytes, to be sent to the peers by any host channel). */
int catb200_peer_alloc(size_t bytes, void** ptr, uint8_t* ipc_handle64);
/* Map a peer's allocation (cudaIpcOpenMemHandle, peer access enabled lazily) / unmap it / free an own allocation. */
int catb200_peer_open(const uint8_t* ipc_handle64, void** ptr);
int catb200_peer_close(void* ptr);
int catb200_peer_free(void* ptr);
/*
 * peer_bases[world]: base pointer of every rank's allocation as mapped in THIS process (own one at index `rank`).
 * parity: arena to reduce; must equal (*epoch_dev & 1), epoch_dev being a device-side counter of completed calls that
 * the kernel increments.  err_dev (device int32, caller-zeroed) becomes 1 if a peer did not arrive within ~2 s, 2 on a
 * parity mismatch -- the kernel never spins forever.  grad_sum [n_params] receives the summed gradient; the remaining
 * arguments are those of catb200_adam_step's norm / clip stage.
 */
int catb200_grad_allreduce_norm(void* const* peer_bases, int32_t rank, int32_t world, int64_t n_params, int32_t parity,
                                float* grad_sum, float grad_scale, float max_grad_norm, float beta1, float beta2,
                                int32_t* step_dev, float* grad_norm_out, void* opt_ws, uint32_t* epoch_dev,
                                int32_t* err_dev, void* stream);

/*
 * catb200_ppo_minibatch_update on several GPUs: forward + backward of one minibatch (gradient accumulated straight into
 * this rank's arena `parity`) and then ONE launch for the rest of the optimizer step -- fold of the accumulators into the
 * arena, flag handshake with every peer, rank-ordered sum of the world arenas into the private `grad_sum`, squared norm,
 * clip, Adam (grad_scale = 1 / world) and the operand-copy refresh -- around two local grid barriers.  Replaces
 * catb200_ppo_minibatch_grad + catb200_grad_allreduce_norm + catb200_adam_apply (reference: the optimizer step of
 * U/cleanrl/ppo.py:351-354 after the gradient all-reduce its multi-process front-ends do); peer arguments as above.
 * By default every rank pulls its peers' whole arenas (one handshake).  CATB200_PEER_RS=1 selects reduce-scatter /
 * all-gather instead: every rank sums one slice of the vector over all arenas and writes it into every rank's
 * summed-gradient region (the third n_pad floats of the peer block; `grad_sum` is then unused), the partial squared norms
 * travel with the second handshake -- 1 / world of the traffic per peer, measured slower at 1.5 MB (two handshakes).
 */
int catb200_ppo_minibatch_update_peer(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                                      const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                                      const float* logprobs_all, const float* advantages_all, const float* returns_all,
                                      const float* values_all, const float* norm_stats, float* params, void* wc,
                                      float* loss_acc, void* workspace, size_t workspace_bytes, float* exp_avg,
                                      float* exp_avg_sq, const float* lr_dev, int32_t* step_dev, float max_grad_norm,
                                      float beta1, float beta2, float eps, float* grad_norm_out, void* opt_ws,
                                      void* const* peer_bases, int32_t rank, int32_t world, int32_t parity, float* grad_sum,
                                      uint32_t* epoch_dev, int32_t* err_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side random draws (Philox4x32-10; csrc/philox.cuh, CPU restatement oracle/philox_oracle.py)
 *
 * rng_state: two uint64 in device memory, {seed, offset}.  Every function below reads the state on the device,
 * uses the counters [offset, offset + consumed) of its own Philox stream and advances offset by `consumed`, so
 * calls can be captured in CUDA graphs and replayed with fresh numbers.
 * ---------------------------------------------------------------------------------------------- */

/* Host-side (no GPU) evaluation of the Philox4x32-10 block function the kernels inline: counter4 / key2 / out4 are
 * raw 32-bit words in Random123 order, so its published known-answer vectors apply directly. */
int catb200_philox4x32_10(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4);
/* Host-side evaluation of catb200_random_permutation for a given (seed, offset). */
int catb200_random_permutation_host(int64_t n, uint64_t seed, uint64_t offset, int64_t* out);

/* out[0..n) = a pseudo-random permutation of 0..n-1 (replaces torch.randperm, U/cleanrl/ppo.py:295): keyed 6-round
 * Feistel bijection of [0, 2^ceil(log2 n)) with cycle walking; consumes 2 counters. */
int catb200_random_permutation(int64_t n, uint64_t* rng_state, int64_t* out, void* stream);

/* mask[i] = (u_i < p[i]) for i < n, u_i uniform in (0,1) (counter offset + i; consumes n).  If ids != NULL, also
 * ids[0..*count) = ascending positions of the set entries (= mask.nonzero().flatten()).  The optional stochastic
 * termination mode of the CaT env: p = constraint probabilities (north_star; the reference itself keeps `dones`
 * as a probability, U/cat/cat_env.py:107). */
int catb200_bernoulli_mask(const float* p, int32_t n, uint64_t* rng_state, uint8_t* mask, int64_t* ids,
                           int32_t* count, void* stream);

/* UniformVelocityCommandWithDeadzone._update_command (U/mdp/commands.py:39-93). */
typedef struct {
  float lin_vel_x[2], lin_vel_y[2], ang_vel_z[2], heading[2]; /* cfg.ranges.*                            */
  float velocity_deadzone;                                    /* commands.py:100                         */
  float heading_control_stiffness;
  float rel_heading_envs, rel_standing_envs;
  float p_step;            /* env.physics_dt / env.max_episode_length_s (commands.py:71-73,81-83)       */
  int32_t heading_command; /* cfg.heading_command                                                        */
} catb200_command_cfg_t;

/*
 * vel_command_b [N,3] in place: heading-error yaw rate for heading envs, dead zone, Bernoulli resampling
 * (p = 0.01 for still commands, p_step otherwise) with uniform redraws in the cfg ranges, Bernoulli yaw flip.
 * Env i uses uniforms u[i][0..8): 0 resample, 1-3 lin_x / lin_y / ang_z, 4 heading, 5 is_heading, 6 is_standing,
 * 7 yaw flip -- from u_ext [N,8] if given (parity tests), else Philox (consumes 8 N).  resampled [N] (optional)
 * receives the resample mask (the caller resets time_left / command_counter for those, as CommandTerm._resample does).
 */
int catb200_command_update(const catb200_command_cfg_t* cfg, int32_t num_envs, float* vel_command_b,
                           float* heading_target, const float* heading_w, uint8_t* is_heading_env,
                           uint8_t* is_standing_env, const float* u_ext, uint64_t* rng_state, uint8_t* resampled,
                           void* stream);

/*
 * push_by_setting_velocity_with_random_envs (U/mdp/events.py:59-96): pushed[i] = (u[i][0] < p_push); pushed envs get
 * root_vel_w[i][k] = lo[k] + (hi[k] - lo[k]) * u[i][1 + k], k = 0..5 (x, y, z, roll, pitch, yaw); the others keep
 * their velocity.  u from u_ext [N,7] or Philox (consumes 8 N).
 */
int catb200_push_select(int32_t num_envs, float p_push, const float* range_lo, const float* range_hi,
                        float* root_vel_w, const float* u_ext, uint64_t* rng_state, uint8_t* pushed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Policy observation assembly  (S12/cat_flat_env_cfg.py:137-172)
 *
 * The 45-d policy observation of the Solo12 task is six ObservationTermCfg's of Isaac Lab's ObservationManager: a column
 * selection of a state tensor, additive uniform noise (AdditiveUniformNoiseCfg: data + rand * (n_max - n_min) + n_min),
 * then a per-term or per-column scale, concatenated.  One launch assembles all terms for all envs; the uniforms come from
 * the device Philox stream (element (env, column) = counter offset + env * n_cols + column; consumes N * n_cols) or from
 * u_ext [N, n_cols] (parity tests).
 * ---------------------------------------------------------------------------------------------- */
#define CATB200_OBS_MAX_TERMS 8
#define CATB200_OBS_MAX_COLS 64
typedef struct {
  const float* src;      /* [N, ...] state tensor, row stride in elements below                  */
  int32_t row_stride;
  int32_t n_cols;        /* columns this term contributes                                        */
  float n_min, n_max;    /* uniform noise bounds; n_min == n_max: no noise                       */
  float noise_span;      /* fl32(n_max - n_min), the subtraction done in double like python does */
  uint8_t ids[32];       /* source column of each output column                                  */
  float scale[32];       /* per output column                                                    */
} catb200_obs_term_t;

typedef struct {
  int32_t n_terms, n_cols;
  catb200_obs_term_t terms[CATB200_OBS_MAX_TERMS];
  uint8_t col_term[CATB200_OBS_MAX_COLS]; /* filled by the call: term of each output column      */
  uint8_t col_idx[CATB200_OBS_MAX_COLS];  /* filled by the call: index of the column in its term  */
} catb200_obs_plan_t;

int catb200_obs_assemble(catb200_obs_plan_t* plan, int32_t num_envs, float* obs_out, const float* u_ext,
                         uint64_t* rng_state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CATB200_H_ */
