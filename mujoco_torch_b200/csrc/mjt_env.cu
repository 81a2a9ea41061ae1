// k_env_post: the RL environment layer's per-agent-step epilogue, fused into one launch (reference mujoco_torch/zoo/base.py
// `_step` :300-340 and the per-environment classes zoo/humanoid.py, ant.py, halfcheetah.py, hopper.py, walker2d.py, swimmer.py,
// cartpole.py): observation, reward, termination, truncation, step counting and the auto-reset of finished envs.  One thread per
// env over the env-minor state the physics kernels just wrote; the result is one env-minor message
//     [ observation (obs_dim rows) | reward | done | terminated ]        (obs_dim + 3, nenv)
// which is exactly what sharding.ObsGather sends over NCCL, so a sharded rollout all-gathers it without a pack step.
#include <cuda_runtime.h>

#include "../../include/mjt.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_env_post(mjtEnvSpec s, T* __restrict__ qpos, T* __restrict__ qvel, T* __restrict__ act, T* __restrict__ timev,
                                                  T* __restrict__ warm, const T* __restrict__ action, T* __restrict__ xprev,
                                                  int* __restrict__ step_count, const T* __restrict__ reset_state, const T* __restrict__ noise,
                                                  T* __restrict__ msg, int B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  const size_t Bs = (size_t)B;
  auto Q = [&](int k) -> T& { return qpos[(size_t)k * Bs + e]; };
  auto V = [&](int k) -> T& { return qvel[(size_t)k * Bs + e]; };
  // ---- reward (zoo/*.py _compute_reward): forward velocity of qpos[0] over the agent step, healthy bonus, control cost
  bool healthy = true;
  if (s.z_index >= 0) { const T z = Q(s.z_index); healthy = healthy && (z >= T(s.z_low)) && (z <= T(s.z_high)); }
  if (s.angle_index >= 0) { const T a = Q(s.angle_index); healthy = healthy && ((a < T(0) ? -a : a) <= T(s.angle_max)); }
  T cost = T(0);
  for (int u = 0; u < s.nu; ++u) { const T a = action[(size_t)u * Bs + e]; cost += a * a; }
  const T fwd = (Q(0) - xprev[e]) / T(s.dt);
  const T reward = T(s.forward_weight) * fwd + (healthy ? T(s.healthy_reward) : T(0)) - T(s.ctrl_cost_weight) * cost + T(s.reward_const);
  const bool terminated = s.terminate_when_unhealthy && !healthy;
  const int cnt = step_count[e] + 1;
  const bool done = terminated || (cnt >= s.max_episode_steps);
  // ---- observation of the state that produced the reward (built BEFORE the reset, zoo/base.py:318-320)
  int o = 0;
  for (int k = s.qpos_skip; k < s.nq; ++k) msg[(size_t)(o++) * Bs + e] = Q(k);
  for (int k = s.qvel_skip; k < s.nv; ++k) {
    T v = V(k);
    if (s.qvel_clip > 0) v = v < T(-s.qvel_clip) ? T(-s.qvel_clip) : (v > T(s.qvel_clip) ? T(s.qvel_clip) : v);
    msg[(size_t)(o++) * Bs + e] = v;
  }
  msg[(size_t)o * Bs + e] = reward;
  msg[(size_t)(o + 1) * Bs + e] = done ? T(1) : T(0);
  msg[(size_t)(o + 2) * Bs + e] = terminated ? T(1) : T(0);
  // ---- fused auto-reset (zoo/base.py:322-329, _make_batch :266-272): finished envs restart from the template state plus uniform noise
  if (done && s.auto_reset) {
    const T* r = reset_state;   // [qpos nq | qvel nv | act na | time 1 | qacc_warmstart nv]
    for (int k = 0; k < s.nq; ++k) Q(k) = r[k] + (noise ? noise[(size_t)k * Bs + e] : T(0));
    for (int k = 0; k < s.nv; ++k) V(k) = r[s.nq + k] + (noise ? noise[(size_t)(s.nq + k) * Bs + e] : T(0));
    for (int k = 0; k < s.na; ++k) act[(size_t)k * Bs + e] = r[s.nq + s.nv + k];
    timev[e] = r[s.nq + s.nv + s.na];
    for (int k = 0; k < s.nv; ++k) warm[(size_t)k * Bs + e] = r[s.nq + s.nv + s.na + 1 + k];
    step_count[e] = 0;
  } else {
    step_count[e] = cnt;
  }
  xprev[e] = Q(0);   // reference x of the next agent step (after a reset: the fresh state's)
}

}  // namespace

extern "C" int mjt_env_obs_dim(const mjtEnvSpec* s) { return (s->nq - s->qpos_skip) + (s->nv - s->qvel_skip); }

extern "C" int mjt_env_post(const mjtEnvSpec* spec, int is_f64, void* qpos, void* qvel, void* act, void* time, void* qacc_warmstart,
                            const void* action, void* xprev, int* step_count, const void* reset_state, const void* noise, void* message,
                            int nenv, void* stream) {
  if (!spec || !qpos || !qvel || !time || !qacc_warmstart || !xprev || !step_count || !message || nenv <= 0) return 1;
  if (spec->nu > 0 && !action) return 1;
  if (spec->auto_reset && !reset_state) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256, grid = (nenv + threads - 1) / threads;
  if (is_f64)
    k_env_post<double><<<grid, threads, 0, st>>>(*spec, (double*)qpos, (double*)qvel, (double*)act, (double*)time, (double*)qacc_warmstart,
                                                 (const double*)action, (double*)xprev, step_count, (const double*)reset_state,
                                                 (const double*)noise, (double*)message, nenv);
  else
    k_env_post<float><<<grid, threads, 0, st>>>(*spec, (float*)qpos, (float*)qvel, (float*)act, (float*)time, (float*)qacc_warmstart,
                                                (const float*)action, (float*)xprev, step_count, (const float*)reset_state,
                                                (const float*)noise, (float*)message, nenv);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
