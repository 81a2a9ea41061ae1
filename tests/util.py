"""Shared helpers for the parity tests: seeded inputs and oracle / CUDA runners."""

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "assets", "models")


def load_model(name, **opt_overrides):
    from mujoco_torch_b200.mjmodel import MjModelLite

    mj = MjModelLite.load_npz(os.path.join(MODELS, name + ".npz"))
    for k, v in opt_overrides.items():
        if k.endswith("]"):      # one component of a vector option: "gravity[0]", "wind[2]"
            name, idx = k[:-1].split("[")
            vec = np.array(getattr(mj.opt, name), dtype=np.float64)
            vec[int(idx)] = v
            setattr(mj.opt, name, vec)
        else:
            setattr(mj.opt, k, v)
    return mj


def seeded_state(mj, B, seed=42, qpos_noise=0.05, qvel_scale=0.5, drop=0.0, ctrl=True):
    """Synthetic initial states: qpos0 + U(-s, s), qvel ~ N(0, scale^2), ctrl ~ U(-1, 1); root height lowered by
    U(0, drop) for free-root models so contacts are active in a good fraction of envs."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1)) + rng.uniform(-qpos_noise, qpos_noise, size=(B, mj.nq))
    qvel = qvel_scale * rng.randn(B, mj.nv)
    c = rng.uniform(-1, 1, size=(B, mj.nu)) if ctrl else np.zeros((B, mj.nu))
    if drop > 0:
        if mj.njnt and int(mj.jnt_type[0]) == 0:
            qpos[:, 2] -= rng.uniform(0, drop, size=B)
        else:
            qpos[:, min(1, mj.nq - 1)] -= rng.uniform(0, drop, size=B)
    return qpos, qvel, c


def pile_state(mj, B, seed=3):
    """Free bodies dropped into a loose pile above the floor: exercises plane-convex and convex-convex contacts."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1))
    k = 0
    for j in range(mj.njnt):
        if int(mj.jnt_type[j]) != 0:
            continue
        a = int(mj.jnt_qposadr[j])
        qpos[:, a:a + 2] += rng.uniform(-0.15, 0.15, size=(B, 2))
        qpos[:, a + 2] = 0.15 + 0.22 * k + rng.uniform(-0.05, 0.05, size=B)
        q = rng.randn(B, 4) * 0.3 + np.array([1.0, 0, 0, 0])
        qpos[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        k += 1
    qvel = 0.3 * rng.randn(B, mj.nv)
    return qpos, qvel, rng.uniform(-1, 1, size=(B, mj.nu))


def stack_state(mj, B, seed=5):
    """The model's own resting arrangement, squeezed: each free body sunk a few mm into the one below and yawed/tilted
    slightly, so that stacked convex pairs are in multi-point face contact."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1))
    k = 0
    for j in range(mj.njnt):
        if int(mj.jnt_type[j]) != 0:
            continue
        a = int(mj.jnt_qposadr[j])
        k += 1
        qpos[:, a:a + 2] += rng.uniform(-0.01, 0.01, size=(B, 2))
        qpos[:, a + 2] -= 0.004 * k + rng.uniform(0.0, 0.002, size=B)
        q = qpos[:, a + 3:a + 7] + rng.randn(B, 4) * 0.01
        qpos[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qvel = 0.1 * rng.randn(B, mj.nv)
    return qpos, qvel, rng.uniform(-1, 1, size=(B, mj.nu))


def sink_state(mj, B, seed=9):
    """assets/xml/sink.xml: joints pushed past their limits in about half of the envs, free bodies pressed into the floor,
    non-zero actuator activations."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1))
    for j in range(mj.njnt):
        t, a = int(mj.jnt_type[j]), int(mj.jnt_qposadr[j])
        if t == 0:      # free: sink a little, random tilt
            qpos[:, a:a + 2] += rng.uniform(-0.02, 0.02, size=(B, 2))
            qpos[:, a + 2] -= rng.uniform(0.0, 0.03, size=B)
            q = np.array([1.0, 0, 0, 0]) + 0.3 * rng.randn(B, 4)
            qpos[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        elif t == 1:    # ball: rotation angle up to ~1 rad about a random axis (limit 0.6)
            ang = rng.uniform(0.0, 1.0, size=B)
            ax = rng.randn(B, 3)
            ax /= np.linalg.norm(ax, axis=1, keepdims=True)
            qpos[:, a] = np.cos(ang / 2)
            qpos[:, a + 1:a + 4] = ax * np.sin(ang / 2)[:, None]
        else:           # slide / hinge: up to 1.3 x the range
            lo, hi = mj.jnt_range[j]
            qpos[:, a] = rng.uniform(1.3 * lo, 1.3 * hi, size=B)
    qvel = 0.5 * rng.randn(B, mj.nv)
    return qpos, qvel, rng.uniform(-1.5, 1.5, size=(B, mj.nu)), rng.uniform(-0.6, 0.6, size=(B, mj.na))


def muscle_state(mj, B, seed=21):
    """assets/xml/muscle.xml: joint angles over their ranges (so the muscles sweep their length-gain curve), activations in [0, 1],
    controls a little beyond [0, 1] (muscle dynamics clamp them)."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1))
    for j in range(mj.njnt):
        lo, hi = mj.jnt_range[j]
        qpos[:, int(mj.jnt_qposadr[j])] = rng.uniform(0.9 * lo, 0.9 * hi, size=B)
    qvel = 2.0 * rng.randn(B, mj.nv)
    return qpos, qvel, rng.uniform(-0.2, 1.2, size=(B, mj.nu)), rng.uniform(0.0, 1.0, size=(B, mj.na))


def terrain_state(mj, B, seed=31):
    """assets/xml/terrain.xml: free bodies scattered over the height field, at heights around its surface (some resting in it, some
    above it), random orientations."""
    rng = np.random.RandomState(seed)
    qpos = np.tile(np.asarray(mj.qpos0), (B, 1))
    for j in range(mj.njnt):
        if int(mj.jnt_type[j]) != 0:
            continue
        a = int(mj.jnt_qposadr[j])
        qpos[:, a] = rng.uniform(-0.9, 0.9, size=B)
        qpos[:, a + 1] = rng.uniform(-0.7, 0.7, size=B)
        qpos[:, a + 2] = rng.uniform(0.1, 0.45, size=B)
        q = rng.randn(B, 4)
        qpos[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    return qpos, 0.3 * rng.randn(B, mj.nv), np.zeros((B, mj.nu))


# step inputs beyond qpos / qvel / ctrl / act that a fixture may record (tests/golden/make_golden.py: "in_" + name)
EXTRA_INPUTS = ("xfrc_applied", "qfrc_applied", "eq_active", "mocap_pos", "mocap_quat")


def run_oracle(mj, qpos, qvel, ctrl, dtype=torch.float64, nsteps=1, warm=None, act=None, fixed_iterations=False, extra=None):
    from oracle import forward as of
    from oracle import model as om

    m = om.OModel(mj, dtype)
    d = of.make_data(m, qpos.shape[0])
    d.qpos, d.qvel, d.ctrl = (torch.as_tensor(x).to(dtype) for x in (qpos, qvel, ctrl))
    if warm is not None:
        d.qacc_warmstart = torch.as_tensor(warm).to(dtype)
    if act is not None:
        d.act = torch.as_tensor(act).to(dtype)
    for k, v in (extra or {}).items():
        setattr(d, k, torch.as_tensor(v).to(torch.int32 if k == "eq_active" else dtype))
    for _ in range(nsteps):
        d = of.step(m, d, fixed_iterations=fixed_iterations)
    return m, d


def run_cuda(mj, qpos, qvel, ctrl, dtype=torch.float64, nsteps=1, warm=None, device="cuda:0", act=None, fixed_iterations=False,
             outputs="full", extra=None):
    import mujoco_torch_b200 as mt

    mx = mt.device_put(mj, dtype=dtype).to(device)
    d = mt.make_data(mx).expand(qpos.shape[0]).clone()
    kw = dict(qpos=torch.as_tensor(qpos, device=device).to(dtype), qvel=torch.as_tensor(qvel, device=device).to(dtype),
              ctrl=torch.as_tensor(ctrl, device=device).to(dtype))
    if warm is not None:
        kw["qacc_warmstart"] = torch.as_tensor(warm, device=device).to(dtype)
    if act is not None:
        kw["act"] = torch.as_tensor(act, device=device).to(dtype)
    for k, v in (extra or {}).items():
        kw[k] = torch.as_tensor(v, device=device).to(getattr(d, k).dtype if k == "eq_active" else dtype)
    d = d.replace(**kw)
    for _ in range(nsteps):
        d = mt.step(mx, d, fixed_iterations=fixed_iterations, outputs=outputs)
    torch.cuda.synchronize()
    return mx, d


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over the whole field (field-scale relative error)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


ORACLE_FIELD = {"contact_dist": "contact.dist", "contact_pos": "contact.pos", "contact_frame": "contact.frame"}


def get(d, name):
    obj = d
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj
