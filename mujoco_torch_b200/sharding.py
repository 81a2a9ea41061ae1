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
