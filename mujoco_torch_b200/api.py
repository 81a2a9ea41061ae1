"""Public API mirroring ``mujoco_torch`` for the batched physics step.

    mx = device_put(mj_model).to("cuda")          reference device.py:1029
    dx = make_data(mx)                            reference io.py:100
    d_batch = dx.expand(B).clone()                reference benchmarks/_helpers.py:25-42
    d_batch = step(mx, d_batch)                   reference forward.py:463
    d_batch = torch.vmap(lambda d: step(mx, d))(d_batch)

``step`` never mutates its input; unchanged fields are shared with the input Data, everything the step writes is a
view of one freshly allocated env-minor buffer.
"""

from __future__ import annotations

import numpy as np
import torch

from . import layout as L
from . import ops
from .data import Contact, Data
from .mjmodel import ConeType, MjModelLite
from .model import Model, Option, Statistic, device_put_model

_OUT_FIELDS = tuple(n for n, _ in L.D_FIELDS if n not in L.HIDDEN_FIELDS and not n.startswith("contact_"))


def device_put(value, *, dtype: torch.dtype | None = None):
    """Place a MuJoCo struct on the device-side representation.  Mirrors reference ``device_put`` (device.py:1029-1112), duck-typed
    on MuJoCo's attribute names because the ``mujoco`` package may be absent:

      * a compiled model (``mujoco.MjModel`` or ``MjModelLite``: has ``nq`` and ``opt``)  -> :class:`Model` (all static precompute);
      * simulation data (``mujoco.MjData`` or any object with ``qpos, qvel, ctrl ...``)    -> :class:`Data` with batch size ``[]``;
        like the reference (`_data_derived`) contact / efc arrays keep the struct's CURRENT sizes -- ``step`` accepts any incoming
        sizes and always returns the model's static ones; rotation matrices come back as ``(n, 3, 3)``;
      * options (``mujoco.MjOption``: has ``timestep`` and ``gravity``)                   -> :class:`Option`;
      * statistics (``mujoco.MjStatistic``: has ``meaninertia``)                          -> :class:`Statistic`.
    """
    if isinstance(value, (Model, Data, Option, Statistic)):
        return value
    if hasattr(value, "nq") and hasattr(value, "opt"):
        return device_put_model(value, dtype)
    if hasattr(value, "qpos") and hasattr(value, "qvel"):
        return _data_from_struct(value, torch.get_default_dtype() if dtype is None else dtype)
    if hasattr(value, "timestep") and hasattr(value, "gravity"):
        return Option(value)
    if hasattr(value, "meaninertia"):
        return Statistic(value)
    raise NotImplementedError(f"{type(value)} is not supported for device_put.")


_DATA_33 = ("xmat", "ximat", "geom_xmat", "site_xmat", "cam_xmat")   # MuJoCo stores rotation matrices flattened to 9
_DATA_STATE = ("qpos", "qvel", "act", "time", "qacc", "qacc_warmstart", "ctrl", "qfrc_applied", "xfrc_applied", "mocap_pos", "mocap_quat",
               "eq_active", "sensordata", "qfrc_constraint")
_CONTACT_FIELDS = ("dist", "pos", "frame", "includemargin", "friction", "solref", "solreffriction", "solimp", "dim", "geom1", "geom2", "geom",
                   "efc_address")


def _data_from_struct(value, dtype) -> Data:
    """reference device_put(MjData): every Data field the struct carries becomes a tensor (floats in ``dtype``, indices integer)."""
    def t(x, want_int=False):
        a = np.ascontiguousarray(np.asarray(x))
        out = torch.as_tensor(a)
        if want_int or not out.is_floating_point():
            return out.to(torch.int32) if a.dtype.kind in "iub" else out
        return out.to(dtype)

    f = {}
    names = [n for n, _ in L.D_FIELDS if n not in L.HIDDEN_FIELDS and not n.startswith("contact_")]
    names += [n for n in _DATA_STATE if n not in names]
    names += ["userdata", "qLDiagInv", "qfrc_fluid", "qfrc_inverse", "cacc", "cfrc_int", "cfrc_ext", "subtree_linvel", "subtree_angmom",
              "efc_pos", "efc_margin", "efc_type", "solver_niter"]
    for name in names:
        if not hasattr(value, name):
            continue
        v = t(getattr(value, name), want_int=name in ("eq_active", "efc_type", "solver_niter"))
        if name == "time":
            v = v.reshape(())
        if name in _DATA_33 and v.dim() >= 1 and v.shape[-1] == 9:
            v = v.reshape(*v.shape[:-1], 3, 3)
        if name == "efc_J" and v.dim() == 1 and hasattr(value, "qvel"):
            nv = int(np.asarray(value.qvel).size)
            v = v.reshape(-1, nv) if nv and v.numel() % nv == 0 else v
        f[name] = v
    for req in ("qpos", "qvel"):
        if req not in f:
            raise ValueError(f"device_put: the data struct has no {req}")
    nv = f["qvel"].shape[-1]
    f.setdefault("time", torch.zeros((), dtype=dtype))
    f.setdefault("act", torch.zeros(0, dtype=dtype))
    f.setdefault("ctrl", torch.zeros(0, dtype=dtype))
    f.setdefault("qacc_warmstart", torch.zeros(nv, dtype=dtype))
    f.setdefault("qfrc_applied", torch.zeros(nv, dtype=dtype))
    f.setdefault("qfrc_constraint", torch.zeros(nv, dtype=dtype))
    f.setdefault("xfrc_applied", torch.zeros((0, 6), dtype=dtype))
    f.setdefault("mocap_pos", torch.zeros((0, 3), dtype=dtype))
    f.setdefault("mocap_quat", torch.zeros((0, 4), dtype=dtype))
    f.setdefault("sensordata", torch.zeros(0, dtype=dtype))
    f.setdefault("eq_active", torch.zeros(0, dtype=torch.int32))
    # contacts: the struct's current ones (0 after a reset), MuJoCo names (`dim`) mapped to the reference's (`contact_dim`)
    c = getattr(value, "contact", None)
    cf = {}
    for name in _CONTACT_FIELDS:
        if c is not None and hasattr(c, name):
            v = t(getattr(c, name), want_int=name in ("dim", "geom1", "geom2", "geom", "efc_address"))
            if name == "frame" and v.dim() >= 1 and v.shape[-1] == 9:
                v = v.reshape(*v.shape[:-1], 3, 3)
            cf["contact_dim" if name == "dim" else name] = v
    ncon = int(cf["dist"].shape[0]) if "dist" in cf else 0
    z = lambda *sh: torch.zeros(sh, dtype=dtype)
    zi = lambda *sh: torch.zeros(sh, dtype=torch.int64)
    for name, default in (("dist", z(ncon)), ("pos", z(ncon, 3)), ("frame", z(ncon, 3, 3)), ("includemargin", z(ncon)), ("friction", z(ncon, 5)),
                          ("solref", z(ncon, 2)), ("solreffriction", z(ncon, 2)), ("solimp", z(ncon, 5)),
                          ("contact_dim", torch.zeros(ncon, dtype=torch.int32)), ("geom1", zi(ncon)), ("geom2", zi(ncon)), ("geom", zi(ncon, 2)),
                          ("efc_address", zi(ncon))):
        cf.setdefault(name, default)
    f["contact"] = Contact(cf)
    nefc = int(f["efc_D"].shape[0]) if "efc_D" in f else int(getattr(value, "nefc", 0) or 0)
    return Data(f, dict(ne=int(getattr(value, "ne", 0) or 0), nf=int(getattr(value, "nf", 0) or 0), nl=int(getattr(value, "nl", 0) or 0),
                        ncon=ncon, nefc=nefc))


def _contact_static(m: Model, lead, device):
    c = m.con
    if m.topk:   # make_data of a max_contact_points model: K slots (their parameters become per-env data after the first step)
        c = {k: v[: m.ncon] for k, v in c.items()}
    ft = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(m.dtype).to(device)
    it = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a).astype(np.int64)).to(dt).to(device)
    ex = lambda t: t.expand(*lead, *t.shape) if lead else t
    g1, g2 = it(c["g1"], torch.int64), it(c["g2"], torch.int64)
    return dict(
        includemargin=ex(ft(c["include"])), friction=ex(ft(c["friction"])), solref=ex(ft(c["solref"])),
        solreffriction=ex(ft(c["solreffriction"])), solimp=ex(ft(c["solimp"])), contact_dim=ex(it(c["dim"], torch.int32)),
        geom1=ex(g1), geom2=ex(g2), geom=ex(torch.stack([g1, g2], dim=-1)), efc_address=ex(it(m.con_efcadr, torch.int64)),
    )


def _contact_topk(m: Model, src: torch.Tensor, lead, device):
    """max_contact_points models: the contact parameters of every env follow the candidates k_topk selected (``src``: candidate
    index per contact slot, (..., K)); reference collision_driver.collision :822-840 emits the same data-dependent Contact."""
    c = m.con
    idx = src.long()
    ft = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(m.dtype).to(device)[idx]
    it = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a).astype(np.int64)).to(dt).to(device)[idx]
    g1, g2 = it(c["g1"], torch.int64), it(c["g2"], torch.int64)
    adr = torch.as_tensor(np.asarray(m.con_efcadr).astype(np.int64)).to(device)
    return dict(includemargin=ft(c["include"]), friction=ft(c["friction"]), solref=ft(c["solref"]), solreffriction=ft(c["solreffriction"]),
                solimp=ft(c["solimp"]), contact_dim=it(c["dim"], torch.int32), geom1=g1, geom2=g2, geom=torch.stack([g1, g2], dim=-1),
                efc_address=adr.expand(*lead, *adr.shape) if lead else adr)


def make_data(m: Model) -> Data:
    """Zero-initialised unbatched Data with static contact / efc sizes (reference io.make_data :100-221)."""
    if not isinstance(m, Model):
        m = device_put(m)
    dev, dt = m.device, m.dtype
    z = lambda *s: torch.zeros(s, dtype=dt, device=dev)
    f = {}
    for name, _ in L.D_FIELDS:
        if name in L.HIDDEN_FIELDS or name.startswith("contact_"):
            continue
        f[name] = z(*m.field_shape[name])
    f["qpos"] = torch.as_tensor(np.asarray(m.mj.qpos0, dtype=np.float64)).to(dt).to(dev)
    f.update(ctrl=z(m.nu), qfrc_applied=z(m.nv), xfrc_applied=z(m.nbody, 6), mocap_pos=z(m.nmocap, 3), mocap_quat=z(m.nmocap, 4),
             eq_active=torch.as_tensor(np.asarray(m.mj.eq_active0).astype(np.int32)).to(dev), userdata=z(0),
             qLDiagInv=z(m.nv), qfrc_fluid=z(m.nv), qfrc_inverse=z(m.nv), cacc=z(m.nbody, 6), cfrc_int=z(m.nbody, 6),
             cfrc_ext=z(m.nbody, 6), subtree_linvel=z(m.nbody, 3), subtree_angmom=z(m.nbody, 3), efc_pos=z(m.nefc), efc_margin=z(m.nefc),
             efc_type=torch.zeros(m.nefc, dtype=torch.int32, device=dev), solver_niter=torch.zeros((), dtype=torch.int32, device=dev))
    for b in range(m.nbody):
        mid = int(m.mj.body_mocapid[b])
        if mid >= 0:
            f["mocap_pos"][mid] = torch.as_tensor(np.asarray(m.mj.body_pos[b])).to(dt)
            f["mocap_quat"][mid] = torch.as_tensor(np.asarray(m.mj.body_quat[b])).to(dt)
    contact = Contact(dict(dist=z(m.ncon), pos=z(m.ncon, 3), frame=z(m.ncon, 3, 3), **_contact_static(m, (), dev)))
    f["contact"] = contact
    return Data(f, dict(ne=m.ne, nf=m.nf, nl=m.nl, ncon=m.ncon, nefc=m.nefc))


def _run_state(m: Model, d: Data, mode: int, args, lead, flat) -> Data:
    """outputs="state": only the state fields come back (views of one small env-minor buffer); every derived field of the
    returned Data is the input's, untouched."""
    out = torch.ops.mjt.run(m._id, mode | ops.MODE_STATE, *args)
    if flat:
        out = out.reshape(*lead, out.shape[-1])
    sl = m.state_layout
    new = dict(d._fields)
    for name, off in sl["pub_off"].items():
        if off < 0:
            continue
        shp = m.field_shape[name]
        n = int(np.prod(shp)) if shp else 1
        v = out[..., off:off + n]
        new[name] = v.reshape(tuple(lead) + tuple(shp)) if len(shp) != 1 else v
    return Data(new, dict(d._meta))


def _run(m: Model, d: Data, mode: int, outputs: str = "full") -> Data:
    if not isinstance(m, Model):
        raise TypeError("first argument must be a mujoco_torch_b200 Model (see device_put)")
    if outputs not in ("full", "state"):
        raise ValueError(f"outputs must be 'full' or 'state', got {outputs!r}")
    lead = tuple(d.qpos.shape[:-1])
    flat = len(lead) > 1
    args = []
    for name in ops._STATE_ARGS:
        t = getattr(d, name)
        if name == "time":
            t = t.unsqueeze(-1)
        ne = {"xfrc_applied": 2, "mocap_pos": 2, "mocap_quat": 2}.get(name, 1)
        t = t.reshape(*t.shape[: t.dim() - ne], -1)
        if flat:
            t = t.reshape(-1, t.shape[-1])
        args.append(t)
    if outputs == "state":
        return _run_state(m, d, mode, args, lead, flat)
    out = torch.ops.mjt.run(m._id, mode, *args)
    if flat:
        out = out.reshape(*lead, out.shape[-1])
    new = dict(d._fields)
    # one split call makes every field's view (host overhead matters at small batch sizes); fields are laid out back to
    # back in L.D_FIELDS order
    plan = m.__dict__.get("_field_plan")
    if plan is None:
        names = [n for n, _ in L.D_FIELDS]
        sizes = [int(np.prod(m.field_shape[n])) if m.field_shape[n] else 1 for n in names]
        plan = m.__dict__["_field_plan"] = (names, sizes, sum(sizes))
    names, sizes, total = plan
    parts = dict(zip(names, torch.split(out[..., :total], sizes, dim=-1)))
    sl = parts.__getitem__
    for name in _OUT_FIELDS:
        shp = m.field_shape[name]
        v = parts[name]
        new[name] = v if len(shp) == 1 else v.reshape(tuple(lead) + tuple(shp))
    # model-constant contact fields (reference collision_driver._compute_static_contact_dict :691-793); efc_address is
    # re-derived cone-aware every step by the reference (constraint.py:636-646), so it always comes from the model
    if m.topk:
        static = _contact_topk(m, sl("scr_consrc").reshape(*lead, m.ncon), lead, out.device)
    else:
        key = (lead, str(out.device))
        cache = m.__dict__.setdefault("_static_contact_cache", {})
        if key not in cache:
            cache[key] = _contact_static(m, lead, out.device)
        static = cache[key]
    new["contact"] = Contact(dict(dist=sl("contact_dist").reshape(*lead, m.ncon), pos=sl("contact_pos").reshape(*lead, m.ncon, 3),
                                  frame=sl("contact_frame").reshape(*lead, m.ncon, 3, 3), **static))
    meta = dict(d._meta)
    meta.update(ne=m.ne, nf=m.nf, nl=m.nl, ncon=m.ncon, nefc=m.nefc)
    return Data(new, meta)


def step(m: Model, d: Data, fixed_iterations: bool = False, *, outputs: str | None = None) -> Data:
    """Advance the simulation by one timestep (reference forward.step :463-496).

    ``fixed_iterations=True`` runs the constraint solver for exactly ``opt.iterations`` iterations and its line search for
    exactly ``opt.ls_iterations`` (no termination test), like the reference's static-graph mode (solver.py:484-487,
    536-537); the default is per-env early exit, which is what the reference computes in eager mode.

    ``outputs`` (an extension; default ``"full"``, or the model's ``default_outputs``): ``"full"`` materialises every field
    of ``Data`` like the reference; ``"state"`` produces only ``qpos, qvel, act, time, qacc, qacc_warmstart, act_dot,
    sensordata`` -- bit-identical values -- and leaves every derived field of the returned ``Data`` as it was in ``d``.
    A rollout that only feeds the state back in (and reads observations from it) moves 20x less memory per step.
    """
    mode = ops.MODE_STEP | (ops.MODE_FIXED if fixed_iterations else 0)
    return _run(m, d, mode, outputs or getattr(m, "default_outputs", "full"))


def forward(m: Model, d: Data, fixed_iterations: bool = False, *, outputs: str | None = None) -> Data:
    """Forward dynamics without integration (reference forward.forward :373-401)."""
    mode = ops.MODE_FORWARD | (ops.MODE_FIXED if fixed_iterations else 0)
    return _run(m, d, mode, outputs or "full")


# ---- stage-level API (reference __init__.py:41-108): each stage returns `d` with that stage's write-set refreshed
# (SURVEY.md Appendix D).  All stages are evaluated by the fused forward kernels from the state in `d`.
_STAGE_FIELDS = {
    "kinematics": ("qpos", "xanchor", "xaxis", "xpos", "xquat", "xmat", "xipos", "ximat", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat"),
    "com_pos": ("subtree_com", "cinert", "cdof"),
    "crb": ("crb", "qM"),
    "factor_m": ("qLD",),
    "tendon": ("ten_length", "ten_J"),
    "transmission": ("actuator_length", "actuator_moment"),
    "com_vel": ("cvel", "cdof_dot"),
    "rne": ("qfrc_bias",),
    "passive": ("qfrc_passive", "qfrc_gravcomp"),
    "collision": ("contact",),
    "make_constraint": ("efc_J", "efc_D", "efc_aref", "efc_frictionloss", "contact"),
    "fwd_position": ("qpos", "xanchor", "xaxis", "xpos", "xquat", "xmat", "xipos", "ximat", "geom_xpos", "geom_xmat", "site_xpos",
                     "site_xmat", "subtree_com", "cinert", "cdof", "crb", "qM", "qLD", "ten_length", "ten_J", "actuator_length",
                     "actuator_moment", "contact", "efc_J", "efc_D", "efc_aref", "efc_frictionloss"),
    "fwd_velocity": ("actuator_velocity", "ten_velocity", "cvel", "cdof_dot", "qfrc_passive", "qfrc_gravcomp", "qfrc_bias"),
    "fwd_actuation": ("act_dot", "qfrc_actuator", "actuator_force"),
    "fwd_acceleration": ("qfrc_smooth", "qacc_smooth"),
    "solve": ("qacc", "qacc_warmstart", "qfrc_constraint", "efc_force"),
    "sensor_pos": ("sensordata",), "sensor_vel": ("sensordata",), "sensor_acc": ("sensordata",),
}


# how far the fused forward pass has to run for each stage (include/mjt.h MJT_FLAG_UPTO_*): 1 = k_smooth, 2 = + k_collision,
# 3 = + k_constraint, 0 = the whole pass (stages whose outputs come from the solver kernel: factorisation, solve, sensors)
_STAGE_UPTO = {"kinematics": 1, "com_pos": 1, "crb": 1, "tendon": 1, "transmission": 1, "com_vel": 1, "rne": 1, "passive": 1,
               "fwd_velocity": 1, "fwd_actuation": 1, "collision": 2, "make_constraint": 3}


def _make_stage(name):
    upto = _STAGE_UPTO.get(name, 0)

    def stage(m: Model, d: Data) -> Data:
        full = _run(m, d, ops.MODE_FORWARD | (upto << ops.MODE_UPTO_SHIFT), "full")
        return d.replace(**{k: full._fields[k] for k in _STAGE_FIELDS[name]})

    stage.__name__ = name
    stage.__doc__ = (f"Stage ``{name}``: recomputed FROM THE STATE in ``d`` (qpos, qvel, act, ctrl, applied forces, mocap) by the fused "
                     f"kernels{', stopping after stage ' + str(upto) if upto else ''}, then {_STAGE_FIELDS[name]} are refreshed in the "
                     "returned Data.  Unlike the reference's stage functions it does not consume derived fields the caller edited in ``d``.")
    return stage


for _n in _STAGE_FIELDS:
    globals()[_n] = _make_stage(_n)


def constraint_sizes(m: Model):
    """(ne, nf, nl, ncon, nefc) -- reference collision_driver.constraint_sizes :681-688."""
    return m.constraint_sizes_py
