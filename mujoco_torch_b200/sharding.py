"""Multi-GPU plumbing: environments are independent, so a job shards them into contiguous ranges, one per rank.

There is no data-path collective: ``torch.distributed`` is only used for the barrier around a timed
region and for the max-over-ranks time.  Synthetic initial states are generated per *global* chunk of environments so
that the union of the ranks' shards is bit-identical to the batch a single process would generate.
"""

from __future__ import annotations

import numpy as np
import torch

CHUNK = 4096  # environments per RNG stream; a stream is keyed by (seed, global chunk index)
RESET_NOISE = {"humanoid": 0.01, "ant": 0.1, "ant_zoo": 0.1, "halfcheetah": 0.1}
NCTRL = 4  # distinct ctrl draws cycled through by the benchmark


def shard_range(total_envs: int, world: int, rank: int):
    """Contiguous env range [lo, hi) owned by ``rank``; earlier ranks take the remainder."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _chunk_state(mj, name, chunk, seed):
    """Reference benchmark distribution (benchmarks/_helpers.py:25-42, zoo/base.py:266-272): qpos0 + U(-s, s),
    qvel = 0.01 N(0,1), ctrl ~ U(-1, 1); half of the envs of free-root models start lowered so contacts are active."""
    rng = np.random.RandomState((seed * 1000003 + chunk) % (2**32))
    s = RESET_NOISE.get(name, 0.05)
    qpos = np.tile(np.asarray(mj.qpos0), (CHUNK, 1)) + rng.uniform(-s, s, size=(CHUNK, mj.nq))
    qvel = 0.01 * rng.randn(CHUNK, mj.nv)
    if mj.njnt and int(mj.jnt_type[0]) == 0:
        qpos[:, 2] -= rng.uniform(0.0, 0.35, size=CHUNK) * (rng.rand(CHUNK) < 0.5)
    ctrl = rng.uniform(-1, 1, size=(NCTRL, CHUNK, mj.nu))
    return qpos, qvel, ctrl


def initial_state(mj, name: str, lo: int, hi: int, seed: int = 42):
    """States of global envs [lo, hi): (qpos (n, nq), qvel (n, nv), ctrl (NCTRL, n, nu)) as float64 numpy arrays."""
    qs, vs, cs = [], [], []
    for chunk in range(lo // CHUNK, (max(hi, lo + 1) - 1) // CHUNK + 1):
        q, v, c = _chunk_state(mj, name, chunk, seed)
        a, b = max(lo, chunk * CHUNK) - chunk * CHUNK, min(hi, (chunk + 1) * CHUNK) - chunk * CHUNK
        qs.append(q[a:b])
        vs.append(v[a:b])
        cs.append(c[:, a:b])
    return np.concatenate(qs), np.concatenate(vs), np.concatenate(cs, axis=1)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (a device-measured time) over the job; identity when not distributed."""
    import torch.distributed as td

    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def job_rate(envs_per_rank: int, steps: int, ms_per_rank: float, device=None) -> float:
    """Whole-job env-steps/s: every rank's envs x steps over the slowest rank's time."""
    import torch.distributed as td

    world = td.get_world_size() if (td.is_available() and td.is_initialized()) else 1
    return world * envs_per_rank * steps / (max_over_ranks(ms_per_rank, device) * 1e-3)


def bind_to_local_cpus(device_index: int) -> dict:
    """One process per GPU on a multi-socket host: run this rank's Python thread -- and, by first touch, place its pinned
    staging buffers -- on the CPU cores NVML reports as local to the GPU (same socket / PCIe root).  Without it the eight ranks
    of a box land wherever the launcher put them and host<->device copies cross the socket interconnect.  Returns what was done
    (for the bench line); never raises."""
    import os

    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        local = sorted(cpus & allowed)
        if local:
            os.sched_setaffinity(0, local)
            info = {"bound": True, "cpus": len(local), "first_cpu": local[0], "last_cpu": local[-1], "allowed_before": len(allowed)}
        else:
            info = {"bound": False, "reason": "no overlap between the GPU's local cpus and this process's allowed set", "allowed": len(allowed)}
    except Exception as exc:  # noqa: BLE001
        info = {"bound": False, "reason": f"{type(exc).__name__}: {exc}"}
    return info


class ObsGather:
    """Optional per-agent-step all-gather of observation + reward + done over the job: the only collective a
    sharded rollout may want, and it is off the physics path.

    Every rank packs its ``(B, obs_dim)`` observation, ``(B,)`` reward and ``(B,)`` done flag into one **env-minor**
    ``(obs_dim + 2, B)`` message -- the layout the step's outputs already have, so packing ``[d.qpos, d.qvel]`` is a handful of
    contiguous row copies -- and ``all_gather_into_tensor`` collects ``(world, obs_dim + 2, B)`` in rank order (= global env
    order, ranks own contiguous env ranges).  On CUDA the pack and the collective run on a **side stream** that first waits for
    the producer stream, so NCCL overlaps the next physics step; ``wait()`` makes the caller's stream wait for the result.  One
    message per step, sized for launch latency (Humanoid: 65 536 envs x 57 words x 8 ranks = 120 MB f32 per gather).

    ``gather = ObsGather(obs_dim, B, dtype, device)``; per step ``gather.launch(obs, reward, done)`` (``obs`` one ``(B, obs_dim)``
    tensor or a list of ``(B, k_i)`` pieces, e.g. ``[d.qpos, d.qvel]``) ... next ``step`` ...
    ``obs_all, reward_all, done_all = gather.wait()`` with shapes ``(world, B, obs_dim)``, ``(world, B)``, ``(world, B)``
    (strided views of the receive buffer; ``obs_all.reshape(world * B, obs_dim)`` gives the flat global batch).  Without an
    initialised process group it degenerates to a local copy.
    """

    def __init__(self, obs_dim: int, envs_per_rank: int, dtype=torch.float32, device="cpu"):
        import torch.distributed as td

        self.dist = td.is_available() and td.is_initialized() and td.get_world_size() > 1
        self.world = td.get_world_size() if self.dist else 1
        self.obs_dim, self.B = int(obs_dim), int(envs_per_rank)
        self.device = torch.device(device)
        self.send = torch.empty((self.obs_dim + 2, self.B), dtype=dtype, device=self.device)
        self.recv = torch.empty((self.world, self.obs_dim + 2, self.B), dtype=dtype, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._work = None
        self.bytes_per_gather = self.recv.numel() * self.recv.element_size()

    def _pack_and_gather(self, parts, reward, done):
        import torch.distributed as td

        from .hostio import _row_blocks

        # (B, k) logical pieces; env-minor sources (views of a step's output buffer) are contiguous row blocks, and pieces that sit
        # back to back in that buffer (qpos, qvel of a state-layout step) travel as ONE copy
        for blk, r0, rows in _row_blocks([p.reshape(self.B, -1) for p in parts], self.B):
            self.send[r0:r0 + rows].copy_(blk)
        self.send[self.obs_dim].copy_(reward.reshape(self.B))
        self.send[self.obs_dim + 1].copy_(done.reshape(self.B))
        if self.dist:
            self._work = td.all_gather_into_tensor(self.recv.view(self.world * (self.obs_dim + 2), self.B), self.send, async_op=True)   # dim-0 concatenation
        else:
            self.recv[0].copy_(self.send)

    def launch(self, obs, reward: torch.Tensor, done: torch.Tensor):
        """Enqueue pack + all-gather (asynchronous; the inputs may be overwritten only after ``wait``)."""
        parts = list(obs) if isinstance(obs, (list, tuple)) else [obs]
        if any(p.shape[0] != self.B for p in parts) or sum(p[0].numel() for p in parts) != self.obs_dim:
            raise ValueError(f"ObsGather: expected obs ({self.B}, {self.obs_dim}), got {[tuple(p.shape) for p in parts]}")
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self._pack_and_gather(parts, reward, done)
            for t in (*parts, reward, done):   # the caching allocator must not recycle the inputs before the side stream has read them
                try:
                    t.record_stream(self.stream)
                except RuntimeError:           # memory not owned by the caching allocator
                    pass
        else:
            self._pack_and_gather(parts, reward, done)

    def launch_packed(self, message: torch.Tensor):
        """All-gather a message that is ALREADY packed env-minor, ``(rows, B)`` contiguous -- e.g. ``zoo.BatchedEnv.message``,
        written by the fused env epilogue kernel: no pack step at all, the collective reads the kernel's output.  ``rows`` may
        exceed ``obs_dim + 2`` (extra rows such as ``terminated`` travel too); use :meth:`wait_packed` for the result."""
        import torch.distributed as td

        if message.dim() != 2 or message.shape[1] != self.B or not message.is_contiguous():
            raise ValueError(f"ObsGather.launch_packed: expected a contiguous (rows, {self.B}) message, got {tuple(message.shape)}")
        rows = message.shape[0]
        if getattr(self, "_recv_packed", None) is None or self._recv_packed.shape[1] != rows:
            self._recv_packed = torch.empty((self.world, rows, self.B), dtype=message.dtype, device=self.device)

        def run():
            if self.dist:
                self._work = td.all_gather_into_tensor(self._recv_packed.view(self.world * rows, self.B), message, async_op=True)
            else:
                self._recv_packed[0].copy_(message)

        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                run()
            try:
                message.record_stream(self.stream)
            except RuntimeError:
                pass
        else:
            run()

    def wait_packed(self) -> torch.Tensor:
        """``(world, rows, B)`` messages of the whole job in rank order (= global env order), after :meth:`launch_packed`."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self._recv_packed

    def wait(self):
        """``(obs (world, B, obs_dim), reward (world, B), done (world, B) bool)`` of the whole job; rank-major = global env order."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.recv[:, : self.obs_dim].permute(0, 2, 1), self.recv[:, self.obs_dim], self.recv[:, self.obs_dim + 1] != 0
