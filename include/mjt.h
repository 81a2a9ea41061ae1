/* mjt.h -- C ABI of libmjt_b200: the B200-native batched physics step.
 *
 * The reference (vmoens/mujoco-torch) is pure Python and has no FFI layer; its boundary for this path is
 *   mujoco_torch.device_put(MjModel)            reference mujoco_torch/_src/device.py:1029
 *   mujoco_torch.make_data(model)               reference mujoco_torch/_src/io.py:100
 *   mujoco_torch.step(model, data)              reference mujoco_torch/_src/forward.py:463
 *   mujoco_torch.forward(model, data)           reference mujoco_torch/_src/forward.py:373
 * The functions below are what a Python binding of those four calls needs (INTEGRATION.md shows the ctypes
 * stub): plain pointers and sizes, no torch types, no exceptions across the boundary (0 = ok, else see
 * mjt_last_error()), caller-owned buffers, stream-ordered and asynchronous.
 *
 * Memory convention ("env-minor"): every per-environment array of n words for a batch of nenv environments is an
 * (n, nenv) array -- word k of environment e at base[k * nenv + e] -- in the model's compute dtype (float or
 * double, fixed when the blob was packed).
 */
#ifndef MJT_H_
#define MJT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mjtModel mjtModel;

/* Input state of one step / forward call: device pointers, env-minor.  Replaces the Data fields the reference's
 * step reads (reference types.Data :1091-1261): qpos (nq), qvel (nv), act (na), time (1), qacc_warmstart (nv),
 * ctrl (nu), qfrc_applied (nv), xfrc_applied (6*nbody), mocap_pos (3*nmocap), mocap_quat (4*nmocap),
 * sensordata (nsensordata; slots of sensor types without an implementation pass through), qfrc_constraint (nv;
 * only read when the model has no constraint rows), eq_active (neq, int32).  qacc is accepted for symmetry with
 * the reference's _check_state but never influences the result.  qpos, qvel, time (and ctrl when nu > 0) are
 * required; any other pointer may be NULL, which reads as zeros (eq_active: as "every equality active"). */
typedef struct mjtStateIn {
  const void* qpos;
  const void* qvel;
  const void* act;
  const void* time;
  const void* qacc;
  const void* qacc_warmstart;
  const void* ctrl;
  const void* qfrc_applied;
  const void* xfrc_applied;
  const void* mocap_pos;
  const void* mocap_quat;
  const void* sensordata;
  const void* qfrc_constraint;
  const int* eq_active;
} mjtStateIn;

/* Library / ABI version (major * 1000 + minor). */
int mjt_version(void);

/* Last error message of the calling thread ("" if none). */
const char* mjt_last_error(void);

/* device_put(MjModel): copies a packed ModelBlob (mujoco_torch_b200/layout.py) to `device`.
 * Replaces reference device.device_put :1029-1112 for MjModel inputs. */
int mjt_model_create(const void* blob_host, size_t nbytes, int device, mjtModel** out);
void mjt_model_destroy(mjtModel* m);

/* Output layout of the model, fixed by the blob it was created from (mujoco_torch_b200/layout.py):
 *   0 = full:  `out` receives the complete new Data, state and every derived field, like the reference's step;
 *   1 = state: `out` receives only qpos, qvel, act, time, qacc, qacc_warmstart, act_dot, sensordata (+ qfrc_constraint for
 *              a model without constraint rows); everything derived lives in the `work` buffer the caller lends to the call,
 *              in the reduced form the next kernel reads (packed qM, support-compacted efc_J), and what nothing reads is
 *              never written.  Same kernels, same arithmetic: the state fields are bit-identical to layout 0. */
int mjt_model_layout(const mjtModel* m);

/* Sizes a caller needs to allocate buffers: words per env of the `out` buffer, of the `work` buffer (state layout: the
 * workspace; RK4: one more stage buffer + accumulators; 0 if neither applies) and sizeof(compute dtype). */
int mjt_model_data_words(const mjtModel* m);
int mjt_model_scratch_words(const mjtModel* m);
int mjt_model_dtype_size(const mjtModel* m);
/* Word offset and word count of a named Data field inside the `out` buffer; -1 if unknown or not part of this layout. */
int mjt_model_field(const mjtModel* m, const char* name, int* offset, int* words);

/* Call flags.  FIXED_ITERATIONS = the reference's `fixed_iterations=True` argument of step / forward / solve
 * (forward.py:463, solver.py:484-487,536-537): the solver runs exactly opt.iterations bodies and its line search
 * exactly opt.ls_iterations, with no termination test.  UPTO_* (mjt_forward, full layout only) stop the pass after the
 * named stage -- what the reference's stage functions (kinematics ... make_constraint, __init__.py:41-108) need. */
enum {
  MJT_FLAG_FIXED_ITERATIONS = 1,
  MJT_FLAG_UPTO_SMOOTH = 2,      /* kinematics .. rne, passive, actuation, qfrc_smooth (k_smooth) */
  MJT_FLAG_UPTO_COLLISION = 4,   /* + contacts (k_collision) */
  MJT_FLAG_UPTO_CONSTRAINT = 8   /* + efc rows (k_constraint) */
};

/* step(model, data, fixed_iterations): reference forward.step :463-496.  `out` is a (data_words, nenv) device buffer
 * that receives the new Data in the model's layout; `work` is (scratch_words, nenv) or NULL when scratch_words == 0.
 * The input is never modified.  Launches on `stream` (a cudaStream_t) and returns immediately. */
int mjt_step(const mjtModel* m, const mjtStateIn* in, void* out, void* work, int nenv, int flags, void* stream);

/* forward(model, data, fixed_iterations): reference forward.forward :373-401 (no integration, no _check_state). */
int mjt_forward(const mjtModel* m, const mjtStateIn* in, void* out, void* work, int nenv, int flags, void* stream);

/* Number of kernel launches mjt_step issues for this model (used by bench.py's gpu_launches claim). */
int mjt_step_launches(const mjtModel* m);

/* Per-kernel device timing for the roofline report: between begin/end every kernel launched by the calling thread is
 * bracketed by CUDA events on its stream; end() synchronises and writes "name:total_ms:launches;..." into buf. */
void mjt_profile_begin(void);
int mjt_profile_end(char* buf, size_t nbuf);

/* Launch configuration of the solver launch that visits every env (environments per CTA, dynamic shared memory bytes): the
 * live-tile pass for models that have one (many, mostly inactive constraint rows), else the single full-tile launch. */
int mjt_solve_config(const mjtModel* m, int* envs_per_cta, int* smem_bytes);

/* Measured FMA-pipe throughput of the current device in TFLOP/s (FMA = 2 flops), FP64 if is_f64 else FP32: register-resident
 * independent FMA chains, best of 5 launches, timed with CUDA events on `stream`.  Denominator of bench.py's compute roofline
 * (SURVEY.md 8(d)); the reference has no counterpart. */
int mjt_fma_peak(int is_f64, double* tflops, void* stream);

/* ---- RL environment layer (reference mujoco_torch/zoo/base.py `_step` :300-340 and the per-environment classes of zoo/):
 * the per-agent-step epilogue after the FRAME_SKIP physics steps, fused into one launch.  Everything is env-minor. */
typedef struct mjtEnvSpec {
  int nq, nv, nu, na;
  int qpos_skip, qvel_skip;       /* observation = qpos[qpos_skip:] ++ clip(qvel[qvel_skip:]) (zoo/*.py _make_obs) */
  double qvel_clip;               /* <= 0: no clipping */
  double dt;                      /* agent step = opt.timestep * FRAME_SKIP */
  double forward_weight, healthy_reward, ctrl_cost_weight, reward_const;   /* reward = fw*(x - x_prev)/dt + hr*healthy - cw*|a|^2 + const */
  int z_index;                    /* healthy needs z_low <= qpos[z_index] <= z_high; -1: no height rule */
  double z_low, z_high;
  int angle_index;                /* healthy needs |qpos[angle_index]| <= angle_max; -1: no angle rule */
  double angle_max;
  int terminate_when_unhealthy;   /* terminated = !healthy */
  int max_episode_steps;          /* truncation: done = terminated | step_count >= max_episode_steps */
  int auto_reset;                 /* finished envs restart from reset_state (+ noise) inside the same launch */
} mjtEnvSpec;

int mjt_env_obs_dim(const mjtEnvSpec* spec);

/* qpos (nq), qvel (nv), act (na), time (1), qacc_warmstart (nv): the state rows the last physics step wrote (updated in place for
 * envs that are reset); action (nu); xprev (1): qpos[0] at the start of the agent step, in/out; step_count (1, int32), in/out;
 * reset_state: DENSE vector [qpos | qvel | act | time | qacc_warmstart] of the template state (same for every env);
 * noise: (nq + nv, nenv) uniform reset noise or NULL; message: (obs_dim + 3, nenv) = observation | reward | done | terminated. */
int mjt_env_post(const mjtEnvSpec* spec, int is_f64, void* qpos, void* qvel, void* act, void* time, void* qacc_warmstart,
                 const void* action, void* xprev, int* step_count, const void* reset_state, const void* noise, void* message,
                 int nenv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MJT_H_ */
