"""Multi-GPU host logic on CPU: two gloo ranks shard a batch of environments.

The physics has no cross-env term, so the test also proves the sharded job equals the single-process one by stepping
each rank's shard with the oracle (the product itself has no CPU path) and comparing the gathered result.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from mujoco_torch_b200 import sharding
from tests import util

TOTAL = 37  # odd on purpose: ranks get 19 and 18 envs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        mj = util.load_model("halfcheetah")
        lo, hi = sharding.shard_range(TOTAL, world, rank)
        qpos, qvel, ctrl = sharding.initial_state(mj, "halfcheetah", lo, hi, seed=42)
        _, d = util.run_oracle(mj, qpos, qvel, ctrl[0])
        gathered = [None] * world
        td.all_gather_object(gathered, {"lo": lo, "hi": hi, "qpos_in": qpos, "qpos": d.qpos.numpy(), "qvel": d.qvel.numpy()})
        # optional observation / reward / done all-gather (equal shards; rank-major = global env order)
        B, od = 5, mj.nq + mj.nv
        g = sharding.ObsGather(od, B, torch.float64)
        obs = torch.cat([d.qpos[:B], d.qvel[:B]], dim=1)
        g.launch(obs, torch.full((B,), float(rank + 1), dtype=torch.float64), torch.arange(B) % 2 == rank)
        obs_all, rew_all, done_all = g.wait()
        obs_parts = [None] * world
        td.all_gather_object(obs_parts, obs.numpy())
        gather_ok = (tuple(obs_all.shape) == (world, B, od)
                     and np.array_equal(obs_all.reshape(world * B, od).numpy(), np.concatenate(obs_parts))
                     and rew_all.reshape(-1).tolist() == [1.0] * B + [2.0] * B
                     and done_all.reshape(-1).tolist() == [i % 2 == 0 for i in range(B)] + [i % 2 == 1 for i in range(B)])
        # the same collective fed by an already packed env-minor message (what the fused env epilogue kernel writes): no pack step
        msg = torch.cat([obs.t(), torch.full((1, B), float(rank + 1), dtype=torch.float64), torch.zeros(2, B, dtype=torch.float64)]).contiguous()
        g.launch_packed(msg)
        packed = g.wait_packed()
        msgs = [None] * world
        td.all_gather_object(msgs, msg.numpy())
        gather_ok = gather_ok and tuple(packed.shape) == (world, od + 3, B) and np.array_equal(packed.numpy(), np.stack(msgs))
        slow = sharding.max_over_ranks(10.0 + rank)          # rank 1 is the slow one
        rate = sharding.job_rate(100, 10, 10.0 + rank)       # 2 ranks x 100 envs x 10 steps / 11 ms
        if rank == 0:
            torch.save({"gathered": gathered, "slow": slow, "rate": rate, "gather_ok": gather_ok}, out)
    finally:
        td.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    g = sorted(res["gathered"], key=lambda x: x["lo"])
    assert [(x["lo"], x["hi"]) for x in g] == [(0, 19), (19, 37)]
    assert res["slow"] == 11.0
    assert res["gather_ok"], "ObsGather: gathered observations / rewards / done flags are not the rank-ordered union"
    assert abs(res["rate"] - 2 * 100 * 10 / 11e-3) < 1e-6
    mj = util.load_model("halfcheetah")
    qpos, qvel, ctrl = sharding.initial_state(mj, "halfcheetah", 0, TOTAL, seed=42)
    assert np.array_equal(np.concatenate([x["qpos_in"] for x in g]), qpos), "union of shards != single-process batch"
    _, d = util.run_oracle(mj, qpos, qvel, ctrl[0])
    for k in ("qpos", "qvel"):
        got = np.concatenate([x[k] for x in g])
        assert np.abs(got - getattr(d, k).numpy()).max() < 1e-12, k


def test_shard_ranges_partition():
    for total, world in ((65536 * 8, 8), (37, 2), (5, 8), (4096, 3)):
        spans = [sharding.shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_states_cross_chunk_boundaries():
    mj = util.load_model("cartpole")
    full = sharding.initial_state(mj, "cartpole", 0, 3 * sharding.CHUNK, seed=7)
    lo, hi = sharding.CHUNK - 5, 2 * sharding.CHUNK + 9
    part = sharding.initial_state(mj, "cartpole", lo, hi, seed=7)
    assert np.array_equal(part[0], full[0][lo:hi]) and np.array_equal(part[1], full[1][lo:hi])
    assert np.array_equal(part[2], full[2][:, lo:hi])


def test_obs_gather_without_process_group_is_a_local_copy():
    g = sharding.ObsGather(3, 4, torch.float32)
    obs = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    g.launch(obs, torch.ones(4), torch.tensor([True, False, False, True]))
    o, r, dn = g.wait()
    assert o.shape == (1, 4, 3) and torch.equal(o[0], obs) and r[0].tolist() == [1.0] * 4 and dn[0].tolist() == [True, False, False, True]
    assert g.bytes_per_gather == 4 * 5 * 4
    g.launch([obs[:, :1], obs[:, 1:]], torch.ones(4), torch.zeros(4))     # obs given in pieces
    assert torch.equal(g.wait()[0][0], obs)
    with pytest.raises(ValueError):
        g.launch(torch.zeros(3, 3), torch.ones(3), torch.zeros(3))
