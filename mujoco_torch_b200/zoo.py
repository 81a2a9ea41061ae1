"""Batched RL environments over the B200 physics step -- the caller side of the hot path (reference ``mujoco_torch/zoo``:
``base.py`` ``MujocoTorchEnv._step`` :300-340, ``_reset`` :274-298, ``_make_batch`` :266-272 and the per-environment classes).

Same semantics as the reference's environments (observation, reward, termination, truncation, optional fused auto-reset, the
``FRAME_SKIP`` physics steps per agent step, the "one step from the XML state" reset template), without the TorchRL / tensordict
dependency: ``reset()`` and ``step(action)`` return plain dicts of tensors with the reference's keys and shapes.

What is B200-specific:
  * the ``FRAME_SKIP`` sub-steps run with ``outputs="state"``: nothing but the state (a ~100-word public buffer per env) exists
    between sub-steps, no derived field is ever materialised;
  * the whole epilogue -- observation, reward, healthy / terminated / truncated, step counters, auto-reset of finished envs -- is
    ONE kernel launch (``csrc/mjt_env.cu``) over the env-minor state, writing one env-minor message
    ``[observation | reward | done | terminated]`` that ``sharding.ObsGather.launch_packed`` all-gathers over NCCL as it is.
"""

from __future__ import annotations

import ctypes
import dataclasses
import os

import torch

from . import api, ops
from .mjmodel import MjModelLite

_MODELS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "models")


@dataclasses.dataclass(frozen=True)
class EnvDef:
    """One environment of the reference's zoo, as data (the reference spells each as a class with these constants)."""
    model: str                      # bundled compiled model (assets/models/<model>.npz)
    frame_skip: int = 1             # zoo/base.py:59
    reset_noise: float = 0.01       # zoo/base.py:58
    qpos_skip: int = 0              # observation = qpos[qpos_skip:] ++ qvel[qvel_skip:]
    qvel_skip: int = 0
    qvel_clip: float = 0.0          # > 0: qvel clamped to +-clip in the observation
    forward_weight: float = 1.0
    healthy_reward: float = 0.0
    ctrl_cost_weight: float = 0.0
    reward_const: float = 0.0
    z_index: int = -1
    z_low: float = 0.0
    z_high: float = 0.0
    angle_index: int = -1
    angle_max: float = 0.0
    terminate_when_unhealthy: bool = False


# reference zoo/humanoid.py, ant.py (zoo-patched model: free root + floor, dt 0.01), halfcheetah.py, hopper.py, walker2d.py, cartpole.py
ENVS = {
    "humanoid": EnvDef("humanoid", frame_skip=5, reset_noise=0.01, qpos_skip=2, qvel_clip=10.0, healthy_reward=5.0, ctrl_cost_weight=0.1,
                       z_index=2, z_low=1.0, z_high=2.0, terminate_when_unhealthy=True),
    "ant": EnvDef("ant_zoo", frame_skip=5, reset_noise=0.1, qpos_skip=2, healthy_reward=1.0, ctrl_cost_weight=0.5,
                  z_index=2, z_low=0.2, z_high=1.0, terminate_when_unhealthy=True),
    "halfcheetah": EnvDef("halfcheetah", frame_skip=5, reset_noise=0.1, qpos_skip=1, ctrl_cost_weight=0.1),
    "hopper": EnvDef("hopper", qpos_skip=1, qvel_clip=10.0, healthy_reward=1.0, ctrl_cost_weight=1e-3, z_index=1, z_low=0.7, z_high=float("inf"),
                     angle_index=2, angle_max=0.2, terminate_when_unhealthy=True),
    "walker2d": EnvDef("walker2d", qpos_skip=1, qvel_clip=10.0, healthy_reward=1.0, ctrl_cost_weight=1e-3, z_index=1, z_low=0.8, z_high=2.0,
                       angle_index=2, angle_max=1.0, terminate_when_unhealthy=True),
    "cartpole": EnvDef("cartpole", forward_weight=0.0, reward_const=1.0, angle_index=1, angle_max=0.2, terminate_when_unhealthy=True),
}


class EnvSpecC(ctypes.Structure):
    """include/mjt.h mjtEnvSpec."""
    _fields_ = [("nq", ctypes.c_int), ("nv", ctypes.c_int), ("nu", ctypes.c_int), ("na", ctypes.c_int), ("qpos_skip", ctypes.c_int),
                ("qvel_skip", ctypes.c_int), ("qvel_clip", ctypes.c_double), ("dt", ctypes.c_double), ("forward_weight", ctypes.c_double),
                ("healthy_reward", ctypes.c_double), ("ctrl_cost_weight", ctypes.c_double), ("reward_const", ctypes.c_double),
                ("z_index", ctypes.c_int), ("z_low", ctypes.c_double), ("z_high", ctypes.c_double), ("angle_index", ctypes.c_int),
                ("angle_max", ctypes.c_double), ("terminate_when_unhealthy", ctypes.c_int), ("max_episode_steps", ctypes.c_int),
                ("auto_reset", ctypes.c_int)]


def _bind():
    L = ops.lib()
    if not getattr(L, "_env_bound", False):
        L.mjt_env_obs_dim.argtypes = [ctypes.POINTER(EnvSpecC)]
        L.mjt_env_obs_dim.restype = ctypes.c_int
        L.mjt_env_post.argtypes = [ctypes.POINTER(EnvSpecC), ctypes.c_int] + [ctypes.c_void_p] * 11 + [ctypes.c_int, ctypes.c_void_p]
        L.mjt_env_post.restype = ctypes.c_int
        L._env_bound = True
    return L


class BatchedEnv:
    """``env = BatchedEnv("humanoid", num_envs=4096, device="cuda", auto_reset=True)``; ``out = env.reset()``;
    ``out = env.step(action)`` with ``action`` ``(num_envs, nu)``.  Returned dicts hold ``observation (B, obs_dim)``, ``reward (B, 1)``,
    ``done (B, 1)`` and ``terminated (B, 1)`` (bool) -- the keys of the reference's TensorDict.  ``fused=False`` evaluates the
    epilogue with torch ops written after the reference's formulas instead of the kernel (the parity check of the tests).
    """

    def __init__(self, name: str, num_envs: int = 64, max_episode_steps: int = 1000, device="cuda", dtype=torch.float64,
                 auto_reset: bool = False, frame_skip: int | None = None, seed: int | None = None, fused: bool = True):
        if name not in ENVS:
            raise KeyError(f"unknown environment {name!r}; available: {sorted(ENVS)}")
        self.name, self.defn = name, ENVS[name]
        self.num_envs, self.max_episode_steps, self.auto_reset, self.fused = int(num_envs), int(max_episode_steps), bool(auto_reset), bool(fused)
        self.device, self.dtype = torch.device(device), dtype
        self.FRAME_SKIP = int(frame_skip) if frame_skip is not None else self.defn.frame_skip
        self.RESET_NOISE_SCALE = self.defn.reset_noise
        mj = MjModelLite.load_npz(os.path.join(_MODELS, self.defn.model + ".npz"))
        self._m_mj = mj
        self._dt = float(mj.opt.timestep) * self.FRAME_SKIP
        self.mx = api.device_put(mj, dtype=dtype).to(self.device)
        # reset template: the XML state after ONE physics step (zoo/base.py:124-131 steps once "so all dtypes match")
        self._dx0 = api.step(self.mx, api.make_data(self.mx), outputs="state")
        self._gen = torch.Generator(device=self.device)
        if seed is not None:
            self._gen.manual_seed(int(seed))
        d = self.defn
        self.spec = EnvSpecC(nq=mj.nq, nv=mj.nv, nu=mj.nu, na=mj.na, qpos_skip=d.qpos_skip, qvel_skip=d.qvel_skip, qvel_clip=d.qvel_clip, dt=self._dt,
                             forward_weight=d.forward_weight, healthy_reward=d.healthy_reward, ctrl_cost_weight=d.ctrl_cost_weight,
                             reward_const=d.reward_const, z_index=d.z_index, z_low=d.z_low, z_high=d.z_high, angle_index=d.angle_index,
                             angle_max=d.angle_max, terminate_when_unhealthy=int(d.terminate_when_unhealthy),
                             max_episode_steps=self.max_episode_steps, auto_reset=int(self.auto_reset))
        self.obs_dim = (mj.nq - d.qpos_skip) + (mj.nv - d.qvel_skip)
        x0 = self._dx0
        self._reset_state = torch.cat([x0.qpos.reshape(-1), x0.qvel.reshape(-1), x0.act.reshape(-1), x0.time.reshape(1), x0.qacc_warmstart.reshape(-1)]).contiguous()
        # two message buffers used in turn: step i's message may still be read by a side-stream all-gather while step i+1 runs
        self._messages = [torch.zeros((self.obs_dim + 3, self.num_envs), dtype=dtype, device=self.device) for _ in range(2)]
        self._message = self._messages[0]
        self._xprev = torch.zeros(self.num_envs, dtype=dtype, device=self.device)
        self._step_count = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)

    # ---- reference interface
    def _make_batch(self, n: int):
        """``n`` fresh envs: the template state plus uniform noise on qpos and qvel (zoo/base.py:266-272)."""
        batch = self._dx0.expand(n).clone()
        noise = self.RESET_NOISE_SCALE
        if noise > 0:
            batch.qpos.add_(torch.empty(batch.qpos.shape, dtype=self.dtype, device=self.device).uniform_(-noise, noise, generator=self._gen))
            batch.qvel.add_(torch.empty(batch.qvel.shape, dtype=self.dtype, device=self.device).uniform_(-noise, noise, generator=self._gen))
        return batch

    def _obs_torch(self):
        d = self.defn
        qvel = self._dx.qvel
        if d.qvel_clip > 0:
            qvel = qvel.clamp(-d.qvel_clip, d.qvel_clip)
        return torch.cat([self._dx.qpos[..., d.qpos_skip:], qvel[..., d.qvel_skip:]], dim=-1)

    def reset(self, mask: torch.Tensor | None = None) -> dict:
        """Full reset, or a per-env reset of the envs selected by the boolean ``mask`` (zoo/base.py:274-298)."""
        if mask is None or not hasattr(self, "_dx"):
            self._dx = self._make_batch(self.num_envs)
            self._step_count.zero_()
        elif not self.auto_reset:
            n = int(mask.sum())
            if n > 0:
                fresh = self._make_batch(n)
                for k in ("qpos", "qvel", "act", "time", "qacc_warmstart"):
                    getattr(self._dx, k)[mask] = getattr(fresh, k)
                self._step_count[mask] = 0
        self._xprev.copy_(self._dx.qpos[:, 0])
        z = torch.zeros(self.num_envs, 1, dtype=torch.bool, device=self.device)
        return {"observation": self._obs_torch(), "done": z, "terminated": z.clone()}

    def _physics(self, ctrl):
        d = self._dx.replace(ctrl=ctrl)
        for _ in range(self.FRAME_SKIP):
            d = api.step(self.mx, d, outputs="state")
        return d

    def step(self, action: torch.Tensor) -> dict:
        """One agent step: ``FRAME_SKIP`` physics steps, then observation / reward / done / terminated (+ auto-reset)."""
        action = action.to(self.dtype)
        if not self.fused:
            return self._step_torch(action)
        self._dx = self._physics(action)
        dx, B = self._dx, self.num_envs
        em = lambda t: ops._env_minor(t.reshape(B, -1), True, self.dtype)
        qpos, qvel, act, tm, warm = em(dx.qpos), em(dx.qvel), em(dx.act), em(dx.time.reshape(B, 1)), em(dx.qacc_warmstart)
        for name, e, t in (("qpos", qpos, dx.qpos), ("qvel", qvel, dx.qvel), ("time", tm, dx.time), ("qacc_warmstart", warm, dx.qacc_warmstart)):
            if e.data_ptr() != t.data_ptr():   # the kernel resets finished envs in place: the state must already be env-minor
                raise RuntimeError(f"BatchedEnv: state field {name} is not env-minor (was it replaced by a row-major tensor?)")
        a_em = em(action)
        noise = None
        if self.auto_reset and self.RESET_NOISE_SCALE > 0:
            noise = torch.empty((self.spec.nq + self.spec.nv, B), dtype=self.dtype, device=self.device).uniform_(
                -self.RESET_NOISE_SCALE, self.RESET_NOISE_SCALE, generator=self._gen)
        L = _bind()
        self._message = self._messages[1] if self._message is self._messages[0] else self._messages[0]
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if (t is not None and t.numel()) else None
        with torch.cuda.device(self.device):
            rc = L.mjt_env_post(ctypes.byref(self.spec), 1 if self.dtype == torch.float64 else 0, p(qpos), p(qvel), p(act), p(tm), p(warm), p(a_em),
                                p(self._xprev), p(self._step_count), p(self._reset_state), p(noise), p(self._message), B,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError("mjt_env_post failed")
        return self.unpack(self._message)

    def unpack(self, message: torch.Tensor) -> dict:
        """Views of an env-minor message ``(obs_dim + 3, B)`` (this env's own, or one gathered over the job) with the reference's keys."""
        o = self.obs_dim
        return {"observation": message[:o].t(), "reward": message[o].unsqueeze(-1), "done": message[o + 1].unsqueeze(-1) != 0,
                "terminated": message[o + 2].unsqueeze(-1) != 0}

    @property
    def message(self) -> torch.Tensor:
        """The last agent step's packed env-minor message (what ``ObsGather.launch_packed`` sends)."""
        return self._message

    # ---- the same agent step with torch ops, written after the reference's formulas (zoo/base.py:300-340 + the env classes)
    def _step_torch(self, action):
        d = self.defn
        qpos_before = self._dx.qpos.clone()
        self._dx = self._physics(action)
        self._step_count += 1
        qpos = self._dx.qpos
        healthy = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        if d.z_index >= 0:
            z = qpos[..., d.z_index]
            healthy &= (z >= d.z_low) & (z <= d.z_high)
        if d.angle_index >= 0:
            healthy &= qpos[..., d.angle_index].abs() <= d.angle_max
        forward_vel = (qpos[..., 0] - qpos_before[..., 0]) / self._dt
        ctrl_cost = d.ctrl_cost_weight * (action ** 2).sum(dim=-1)
        reward = d.forward_weight * forward_vel + torch.where(healthy, d.healthy_reward, 0.0) - ctrl_cost + d.reward_const
        terminated = (~healthy if d.terminate_when_unhealthy else torch.zeros_like(healthy)).unsqueeze(-1)
        done = terminated | (self._step_count >= self.max_episode_steps).unsqueeze(-1)
        obs = self._obs_torch().clone()
        if self.auto_reset:
            mask = done.squeeze(-1)
            n = int(mask.sum())
            if n:
                fresh = self._make_batch(n)
                for k in ("qpos", "qvel", "act", "time", "qacc_warmstart"):
                    getattr(self._dx, k)[mask] = getattr(fresh, k)
                self._step_count[mask] = 0
        return {"observation": obs, "reward": reward.unsqueeze(-1).to(self.dtype), "done": done, "terminated": terminated}
