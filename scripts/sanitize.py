"""Small steps of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): odd batch sizes, f64 and f32,
full-warp, half-warp, compact and two-dof-per-lane solver instantiations, convex narrowphase, RK4, sensors; both output layouts,
fixed_iterations, the stage-limited forward passes, top-k contacts, height fields, muscles and the env epilogue kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import mujoco_torch_b200 as mt

CASES = [("humanoid", {}, 0.4, 33), ("halfcheetah", {}, 0.6, 37), ("ant", {}, 0.0, 5), ("ant_zoo", {"cone": 1, "solver": 1}, 0.5, 19),
         ("stack", {}, -2.0, 7), ("sink", {}, -3.0, 9), ("sink", {"integrator": 3}, -3.0, 3), ("equality_bodies", {}, 0.0, 4),
         ("convex", {}, -1.0, 3), ("cartpole", {}, 0.0, 1),
         ("capsules_topk", {}, -1.0, 5), ("terrain", {}, -5.0, 6), ("muscle", {}, -4.0, 5), ("muscle", {"integrator": 3}, -4.0, 3),
         ("sensors_extra", {"integrator": 1}, 0.5, 5),
         # live-tile solver with caps so small that most envs overflow into the full-tile pass (5th entry: env overrides)
         ("stack", {"integrator": 0}, -1.0, 7, {"MJT_LIVE_MIN_ROWS": "16", "MJT_LIVE_FRAC": "0.01"}),
         ("ant_zoo", {"cone": 1, "solver": 1}, 0.5, 19, {"MJT_LIVE_MIN_ROWS": "16", "MJT_LIVE_FRAC": "0.01"})]
ONLY = os.environ.get("MJT_SANITIZE_ONLY")   # comma-separated model names (time-boxed runs)
for name, opt, drop, B, *rest in CASES:
    if ONLY and name not in ONLY.split(","):
        continue
    for k in ("MJT_LIVE_MIN_ROWS", "MJT_LIVE_FRAC"):
        os.environ.pop(k, None)
    os.environ.update(rest[0] if rest else {})
    mj = util.load_model(name, **opt)
    act = None
    if drop >= 0:
        qpos, qvel, ctrl = util.seeded_state(mj, B, drop=drop)
    elif drop < -4.5:
        qpos, qvel, ctrl = util.terrain_state(mj, B)
    elif drop < -3.5:
        qpos, qvel, ctrl, act = util.muscle_state(mj, B)
    elif drop < -2.5:
        qpos, qvel, ctrl, act = util.sink_state(mj, B)
    else:
        qpos, qvel, ctrl = util.pile_state(mj, B) if drop > -1.5 else util.stack_state(mj, B)
    for dtype in (torch.float64, torch.float32):
        mx, d = util.run_cuda(mj, qpos, qvel, ctrl, dtype, nsteps=2, act=act)
        assert torch.isfinite(d.qpos).all(), (name, dtype)
        mx, ds = util.run_cuda(mj, qpos, qvel, ctrl, dtype, nsteps=2, act=act, outputs="state")
        assert torch.equal(ds.qpos, d.qpos), (name, dtype, "state layout")
        if not rest:
            util.run_cuda(mj, qpos, qvel, ctrl, dtype, nsteps=1, act=act, outputs="state", fixed_iterations=True)
            mt.collision(mx, d)
            mt.make_constraint(mx, d)
            mt.kinematics(mx, d)
    print("ok", name, opt, B, rest[0] if rest else "", flush=True)

# env layer: the fused epilogue kernel (observation / reward / termination / auto-reset) behind two agent steps
for name in ("humanoid", "hopper", "cartpole"):
    if ONLY and name not in ONLY.split(","):
        continue
    env = mt.zoo.BatchedEnv(name, num_envs=11, max_episode_steps=3, device="cuda:0", dtype=torch.float64, auto_reset=True, seed=1)
    env.reset()
    for _ in range(4):
        out = env.step(torch.rand(11, env._m_mj.nu, device="cuda:0", dtype=torch.float64) * 2 - 1)
    torch.cuda.synchronize()
    print("ok env", name, flush=True)
