"""How many envs of the bench workload go non-finite, per step (f32 / f64, Euler / ImplicitFast, state / full)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, numpy as np, torch
import mujoco_torch_b200 as mt
from mujoco_torch_b200 import sharding
from mujoco_torch_b200.mjmodel import MjModelLite
ap = argparse.ArgumentParser(); ap.add_argument("--dtype", default="f32"); ap.add_argument("--integrator", type=int, default=3)
ap.add_argument("--outputs", default="state"); ap.add_argument("--steps", type=int, default=40); ap.add_argument("--batch", type=int, default=65536)
a = ap.parse_args()
mj = MjModelLite.load_npz("assets/models/humanoid.npz"); mj.opt.integrator = a.integrator
dt = torch.float32 if a.dtype == "f32" else torch.float64
mx = mt.device_put(mj, dtype=dt).to("cuda:0")
qpos, qvel, ctrl = sharding.initial_state(mj, "humanoid", 0, a.batch, seed=42)
T = lambda x: torch.as_tensor(x, device="cuda:0").to(dt)
d = mt.make_data(mx).expand(a.batch).clone().replace(qpos=T(qpos), qvel=T(qvel))
cs = [T(c) for c in ctrl]
for i in range(a.steps):
    d = mt.step(mx, d.replace(ctrl=cs[i % 4]), outputs=a.outputs)
    bad = ~(torch.isfinite(d.qpos).all(1) & torch.isfinite(d.qvel).all(1))
    big = d.qvel.abs().amax(1)
    if i % 4 == 3 or bad.any():
        print(i, "nonfinite envs", int(bad.sum()), "max|qvel| over finite", float(big[~bad].max()), "envs with |qvel|>1e3", int((big[~bad] > 1e3).sum()), flush=True)
    if bad.sum() > 100: break
