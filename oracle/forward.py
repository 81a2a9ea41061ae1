"""ORACLE (test infrastructure): pipeline driver, passive forces, actuation, integrators.

Follows reference ``mujoco_torch/_src/forward.py`` (_check_state :44-59, _position :73-84, _velocity :87-99,
_actuation :102-219, _acceleration :222-228, _advance :255-310, _euler :313-328, _rungekutta4 :331-370,
forward :373-401, _implicit :404-416, step :463-496), ``passive.py`` :80-200 and ``derivative.py`` :22-68.
All tensors carry a leading env dim.
"""

import copy

import numpy as np
import torch

from . import collision, constraint, sensor, smooth, solver
from . import omath as M
from .model import (BALL, DSBL_ACTUATION, DSBL_CLAMPCTRL, DSBL_DAMPER, DSBL_EULERDAMP, DSBL_GRAVITY, DSBL_SPRING, FREE,
                    HINGE, SLIDE, mjMINVAL)

_MAXVAL = 1e10


class OData:
    """Attribute bag of batched tensors (leading dim = env)."""

    def clone(self):
        return copy.copy(self)


def make_data(m, B=1):
    """reference io.make_data :100-221 (state + output buffers the step reads before writing)."""
    d = OData()
    dt = m.dtype
    z = lambda *s: torch.zeros((B,) + s, dtype=dt)
    d.qpos = m.qpos0.expand(B, m.nq).clone()
    d.qvel, d.qacc, d.qacc_warmstart = z(m.nv), z(m.nv), z(m.nv)
    d.act, d.act_dot, d.ctrl = z(m.na), z(m.na), z(m.nu)
    d.qfrc_applied, d.xfrc_applied = z(m.nv), z(m.nbody, 6)
    d.qfrc_constraint, d.qfrc_gravcomp = z(m.nv), z(m.nv)
    d.time = z()
    d.eq_active = torch.as_tensor(np.asarray(m.mj.eq_active0, dtype=np.int32)).expand(B, m.neq).clone()
    d.sensordata = z(m.nsensordata)
    d.cacc = z(m.nbody, 6)
    # output buffers that some configurations never write (nefc == 0, nu == 0): zeros, as in the reference's make_data
    d.efc_J, d.efc_D, d.efc_aref, d.efc_frictionloss, d.efc_force = z(m.nefc, m.nv), z(m.nefc), z(m.nefc), z(m.nefc), z(m.nefc)
    d.actuator_length, d.actuator_moment, d.actuator_velocity, d.actuator_force = z(m.nu), z(m.nu, m.nv), z(m.nu), z(m.nu)
    d.qfrc_actuator = z(m.nv)
    d.mocap_pos, d.mocap_quat = z(m.nmocap, 3), z(m.nmocap, 4)
    for b in range(m.nbody):
        mid = int(m.mj.body_mocapid[b])
        if mid >= 0:
            d.mocap_pos[:, mid], d.mocap_quat[:, mid] = m.body_pos[b], m.body_quat[b]
    return d


def check_state(m, d):
    bad = lambda x: ~torch.isfinite(x) | (x.abs() > _MAXVAL)
    d.qpos = torch.where(bad(d.qpos), m.qpos0.expand_as(d.qpos), d.qpos)
    d.qvel = torch.where(bad(d.qvel), torch.zeros_like(d.qvel), d.qvel)
    d.qacc = torch.where(bad(d.qacc), torch.zeros_like(d.qacc), d.qacc)
    return d


# Test hook: when set, called as CONTACT_HOOK(m, d) right after the narrowphase.  Parity tests use it to replay the
# constraint/solver stages on a given contact manifold (the reference's manifold point selection breaks exact ties
# by rounding noise, see DESIGN.md "Discontinuities"), so downstream stages can be compared on identical contacts.
CONTACT_HOOK = None


def position(m, d):
    d = smooth.kinematics(m, d)
    d = smooth.com_pos(m, d)
    d = smooth.tendon(m, d)
    d = smooth.crb(m, d)
    d = smooth.tendon_armature(m, d)
    smooth.factor_m(m, d)
    d = collision.collision(m, d)
    if CONTACT_HOOK is not None:
        d = CONTACT_HOOK(m, d)
    d = constraint.make_constraint(m, d)
    d = smooth.transmission(m, d)
    return d


# ---------------------------------------------------------------- passive (passive.py)


def _spring_damper(m, d):
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    qfrc = torch.zeros((B, m.nv), dtype=dt)
    if not (m.disableflags & DSBL_SPRING):
        for j in range(m.njnt):
            jt, qa, da = int(mj.jnt_type[j]), int(mj.jnt_qposadr[j]), int(mj.jnt_dofadr[j])
            k = m.jnt_stiffness[j]
            if jt == FREE:
                qfrc[:, da : da + 3] = -k * (d.qpos[:, qa : qa + 3] - m.qpos_spring[qa : qa + 3])
                qfrc[:, da + 3 : da + 6] = -k * M.quat_sub(d.qpos[:, qa + 3 : qa + 7], m.qpos_spring[qa + 3 : qa + 7].expand(B, 4))
            elif jt == BALL:
                qfrc[:, da : da + 3] = -k * M.quat_sub(d.qpos[:, qa : qa + 4], m.qpos_spring[qa : qa + 4].expand(B, 4))
            else:
                qfrc[:, da] = -k * (d.qpos[:, qa] - m.qpos_spring[qa])
    if not (m.disableflags & DSBL_DAMPER):
        qfrc = qfrc - m.dof_damping * d.qvel
    if m.ntendon:
        z = torch.zeros((B, m.ntendon), dtype=dt)
        frc_spring, frc_damper = z, z
        if not (m.disableflags & DSBL_SPRING):
            below = m.tendon_lengthspring[:, 0] - d.ten_length
            above = m.tendon_lengthspring[:, 1] - d.ten_length
            frc_spring = torch.where(below > 0, m.tendon_stiffness * below, torch.zeros_like(below))
            frc_spring = torch.where(above < 0, m.tendon_stiffness * above, frc_spring)
        if not (m.disableflags & DSBL_DAMPER):
            frc_damper = -m.tendon_damping * d.ten_velocity
        qfrc = qfrc + torch.matmul(d.ten_J.transpose(-1, -2), (frc_spring + frc_damper).unsqueeze(-1)).squeeze(-1)
    return qfrc


def _gravcomp(m, d):
    B = d.qpos.shape[0]
    out = torch.zeros_like(d.qvel)
    for b in range(m.nbody):
        force = (-m.gravity * (m.body_mass[b] * m.body_gravcomp[b])).expand(B, 3)
        jacp, _ = smooth.jac(m, d, d.xipos[:, b], b)
        out = out + (jacp * force[:, None, :]).sum(-1)
    return out


def _fluid(m, d):
    """_inertia_box_fluid_model :30-77 + _fluid :158-173."""
    B, dt = d.qpos.shape[0], d.qpos.dtype
    out = torch.zeros_like(d.qvel)
    root = np.asarray(m.mj.body_rootid)
    for b in range(m.nbody):
        inertia, mass = m.body_inertia[b], m.body_mass[b]
        box = inertia.unsqueeze(0).repeat(3, 1) * (torch.ones((3, 3), dtype=dt) - 2 * torch.eye(3, dtype=dt))
        box = 6.0 * torch.clamp(box.sum(-1), min=1e-12)
        box = torch.sqrt(box / torch.clamp_min(mass, 1e-12)) * (mass > 0.0)
        ximat = d.ximat[:, b]
        offset = d.xipos[:, b] - d.subtree_com[:, int(root[b])]
        lvel = M.transform_motion(d.cvel[:, b], offset, ximat)
        lwind = (ximat.transpose(-1, -2) * m.wind).sum(-1)
        lvel = torch.cat([lvel[:, :3], lvel[:, 3:] - lwind], dim=-1)
        diam = box.mean()
        lfrc_ang = lvel[:, :3] * -torch.pi * diam**3 * m.viscosity
        lfrc_vel = lvel[:, 3:] * -3.0 * torch.pi * diam * m.viscosity
        scale_vel = torch.stack([box[1] * box[2], box[0] * box[2], box[0] * box[1]])
        scale_ang = torch.stack([box[0] * (box[1] ** 4 + box[2] ** 4), box[1] * (box[0] ** 4 + box[2] ** 4),
                                 box[2] * (box[0] ** 4 + box[1] ** 4)])
        lfrc_vel = lfrc_vel - 0.5 * m.density * scale_vel * torch.abs(lvel[:, 3:]) * lvel[:, 3:]
        lfrc_ang = lfrc_ang - (1.0 * m.density * scale_ang * torch.abs(lvel[:, :3]) * lvel[:, :3] / 64.0)
        force, torque = (ximat * lfrc_vel[:, None, :]).sum(-1), (ximat * lfrc_ang[:, None, :]).sum(-1)
        out = out + smooth.apply_ft(m, d, force, torque, d.xipos[:, b], b)
    return out


def passive(m, d):
    """passive.passive :176-200."""
    if m.disableflags & (DSBL_SPRING | DSBL_DAMPER):
        d.qfrc_passive = torch.zeros_like(d.qvel)
        d.qfrc_gravcomp = torch.zeros_like(d.qvel)
        return d
    qfrc_passive = _spring_damper(m, d)
    if m.has_gravcomp and not (m.disableflags & DSBL_GRAVITY):
        d.qfrc_gravcomp = _gravcomp(m, d)
        agc = torch.as_tensor(np.asarray(m.mj.jnt_actgravcomp)[np.asarray(m.mj.dof_jntid)]).to(d.qpos.dtype)
        qfrc_passive = qfrc_passive + d.qfrc_gravcomp * (1 - agc)
    if m.has_fluid:
        qfrc_passive = qfrc_passive + _fluid(m, d)
    d.qfrc_passive = qfrc_passive
    return d


def velocity(m, d):
    mv_ = lambda A, x: torch.matmul(A, x.unsqueeze(-1)).squeeze(-1)
    if m.nu:
        d.actuator_velocity = mv_(d.actuator_moment, d.qvel)
    else:
        d.actuator_velocity = torch.zeros((d.qpos.shape[0], 0), dtype=d.qpos.dtype)
    if m.ntendon:
        d.ten_velocity = mv_(d.ten_J, d.qvel)
    d = smooth.com_vel(m, d)
    d = passive(m, d)
    d = smooth.rne(m, d)
    return d


# ---------------------------------------------------------------- actuation


def _muscle_sigmoid(x):
    sol = x * x * x * (3 * x * (2 * x - 5) + 10)
    sol = torch.where(x <= 0, torch.zeros_like(sol), sol)
    return torch.where(x >= 1, torch.ones_like(sol), sol)


def _muscle_dynamics(ctrl, act, prm):
    """support.muscle_dynamics :205-228."""
    ctrlclamp, actclamp = torch.clamp(ctrl, 0, 1), torch.clamp(act, 0, 1)
    tau_act = prm[0] * (0.5 + 1.5 * actclamp)
    tau_deact = prm[1] / (0.5 + 1.5 * actclamp)
    sw = prm[2]
    dctrl = ctrlclamp - act
    tau_hard = torch.where(dctrl > 0, tau_act, tau_deact)
    tau_smooth = tau_deact + (tau_act - tau_deact) * _muscle_sigmoid(M.safe_div(dctrl, sw) + 0.5)
    tau = torch.where(sw < mjMINVAL, tau_hard, tau_smooth)
    return dctrl / torch.clamp_min(tau, mjMINVAL)


def _muscle_gain_length(length, lmin, lmax):
    a, b = 0.5 * (lmin + 1), 0.5 * (1 + lmax)
    cm = lambda x: torch.clamp_min(x, mjMINVAL)
    out0 = 0.5 * torch.square((length - lmin) / cm(a - lmin))
    out1 = 1 - 0.5 * torch.square((1 - length) / cm(1 - a))
    out2 = 1 - 0.5 * torch.square((length - 1) / cm(b - 1))
    out3 = 0.5 * torch.square((lmax - length) / cm(lmax - b))
    out = torch.where(length <= b, out2, out3)
    out = torch.where(length <= 1, out1, out)
    out = torch.where(length <= a, out0, out)
    return torch.where((lmin <= length) & (length <= lmax), out, torch.zeros_like(out))


def _muscle_gain(length, vel, lengthrange, acc0, prm):
    cm = lambda x: torch.clamp_min(x, mjMINVAL)
    lrange = prm[:2]
    force, scale, lmin, lmax, vmax, fvmax = prm[2], prm[3], prm[4], prm[5], prm[6], prm[8]
    force = torch.where(force < 0, scale / cm(acc0), force)
    L0 = (lengthrange[1] - lengthrange[0]) / cm(lrange[1] - lrange[0])
    L = lrange[0] + (length - lengthrange[0]) / cm(L0)
    V = vel / cm(L0 * vmax)
    FL = _muscle_gain_length(L, lmin, lmax)
    y = fvmax - 1
    FV = torch.where(V <= y, fvmax - torch.square(y - V) / cm(y), fvmax.expand_as(V))
    FV = torch.where(V <= 0, torch.square(V + 1), FV)
    FV = torch.where(V <= -1, torch.zeros_like(FV), FV)
    return -force * FL * FV


def _muscle_bias(length, lengthrange, acc0, prm):
    cm = lambda x: torch.clamp_min(x, mjMINVAL)
    lrange = prm[:2]
    force, scale, lmax, fpmax = prm[2], prm[3], prm[5], prm[7]
    force = torch.where(force < 0, scale / cm(acc0), force)
    L0 = (lengthrange[1] - lengthrange[0]) / cm(lrange[1] - lrange[0])
    L = lrange[0] + (length - lengthrange[0]) / cm(L0)
    b = 0.5 * (1 + lmax)
    out1 = -force * fpmax * 0.5 * torch.square((L - 1) / cm(b - 1))
    out2 = -force * fpmax * (0.5 + (L - b) / cm(b - 1))
    out = torch.where(L <= b, out1, out2)
    return torch.where(L <= 1, torch.zeros_like(out), out)


def actuation(m, d):
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    if not m.nu or m.disableflags & DSBL_ACTUATION:
        d.act_dot = torch.zeros((B, m.na), dtype=dt)
        d.qfrc_actuator = torch.zeros((B, m.nv), dtype=dt)
        return d
    ctrl = d.ctrl
    inf = float("inf")
    if not (m.disableflags & DSBL_CLAMPCTRL):
        lim = torch.as_tensor(np.asarray(mj.actuator_ctrllimited).astype(bool))
        lo = torch.where(lim, m.actuator_ctrlrange[:, 0], torch.full_like(m.actuator_ctrlrange[:, 0], -inf))
        hi = torch.where(lim, m.actuator_ctrlrange[:, 1], torch.full_like(m.actuator_ctrlrange[:, 1], inf))
        ctrl = torch.minimum(torch.maximum(ctrl, lo), hi)
    act_dot = torch.zeros((B, m.na), dtype=dt)
    ctrl_act = ctrl.clone()
    for i in range(m.nu):
        dyn, adr = int(mj.actuator_dyntype[i]), int(mj.actuator_actadr[i])
        if adr < 0:
            continue
        act = d.act[:, adr]
        prm = m.actuator_dynprm[i]
        if dyn == 1:
            ad = ctrl[:, i]
        elif dyn in (2, 3):
            ad = (ctrl[:, i] - act) / torch.clamp_min(prm[0], mjMINVAL)
        elif dyn == 4:
            ad = _muscle_dynamics(ctrl[:, i], act, prm)
        else:
            ad = torch.zeros_like(act)
        act_dot[:, adr] = ad
        ctrl_act[:, i] = d.act[:, adr + int(mj.actuator_actnum[i]) - 1]
    force = torch.zeros((B, m.nu), dtype=dt)
    for i in range(m.nu):
        gt, bt = int(mj.actuator_gaintype[i]), int(mj.actuator_biastype[i])
        gp, bp = m.actuator_gainprm[i], m.actuator_biasprm[i]
        len_, vel = d.actuator_length[:, i], d.actuator_velocity[:, i]
        if gt == 0:
            gain = gp[0]
        elif gt == 1:
            gain = gp[0] + gp[1] * len_ + gp[2] * vel
        elif gt == 2:
            gain = _muscle_gain(len_, vel, m.actuator_lengthrange[i], m.actuator_acc0[i], gp)
        else:
            raise RuntimeError("unrecognized gaintype")
        bias = torch.zeros_like(len_)
        if bt == 1:
            bias = bp[0] + bp[1] * len_ + bp[2] * vel
        elif bt == 2:
            bias = _muscle_bias(len_, m.actuator_lengthrange[i], m.actuator_acc0[i], bp)
        force[:, i] = gain * ctrl_act[:, i] + bias
    flim = torch.as_tensor(np.asarray(mj.actuator_forcelimited).astype(bool))
    flo = torch.where(flim, m.actuator_forcerange[:, 0], torch.full_like(m.actuator_forcerange[:, 0], -inf))
    fhi = torch.where(flim, m.actuator_forcerange[:, 1], torch.full_like(m.actuator_forcerange[:, 1], inf))
    force = torch.minimum(torch.maximum(force, flo), fhi)
    qfrc_actuator = torch.matmul(d.actuator_moment.transpose(-1, -2), force.unsqueeze(-1)).squeeze(-1)
    dof_jnt = torch.as_tensor(np.asarray(mj.dof_jntid), dtype=torch.long)
    if m.has_gravcomp:
        agc = torch.as_tensor(np.asarray(mj.jnt_actgravcomp)).to(dt)[dof_jnt]
        qfrc_actuator = qfrc_actuator + d.qfrc_gravcomp * agc
    alim = torch.as_tensor(np.asarray(mj.jnt_actfrclimited).astype(bool))
    alo = torch.where(alim, m.jnt_actfrcrange[:, 0], torch.full_like(m.jnt_actfrcrange[:, 0], -inf))[dof_jnt]
    ahi = torch.where(alim, m.jnt_actfrcrange[:, 1], torch.full_like(m.jnt_actfrcrange[:, 1], inf))[dof_jnt]
    qfrc_actuator = torch.minimum(torch.maximum(qfrc_actuator, alo), ahi)
    d.act_dot, d.qfrc_actuator, d.actuator_force = act_dot, qfrc_actuator, force
    return d


def acceleration(m, d):
    qfrc_applied = d.qfrc_applied + smooth.xfrc_accumulate(m, d)
    d.qfrc_smooth = d.qfrc_passive - d.qfrc_bias + d.qfrc_actuator + qfrc_applied
    d.qacc_smooth = smooth.solve_m(m, d, d.qfrc_smooth)
    return d


def forward(m, d, fixed_iterations=False):
    d = position(m, d)
    d = sensor.sensor_pos(m, d)
    d = velocity(m, d)
    d = sensor.sensor_vel(m, d)
    d = actuation(m, d)
    d = acceleration(m, d)
    if m.nefc == 0:
        d.qacc = d.qacc_smooth.clone()
        return d
    d = solver.solve(m, d, fixed_iterations=fixed_iterations)
    d = sensor.sensor_acc(m, d)
    return d


# ---------------------------------------------------------------- integrators


def _integrate_pos(m, qpos, qvel, dt_):
    mj = m.mj
    out = qpos.clone()
    for j in range(m.njnt):
        jt, qa, da = int(mj.jnt_type[j]), int(mj.jnt_qposadr[j]), int(mj.jnt_dofadr[j])
        if jt == FREE:
            out[:, qa : qa + 3] = qpos[:, qa : qa + 3] + dt_ * qvel[:, da : da + 3]
            out[:, qa + 3 : qa + 7] = M.quat_integrate(qpos[:, qa + 3 : qa + 7], qvel[:, da + 3 : da + 6], dt_)
        elif jt == BALL:
            out[:, qa : qa + 4] = M.quat_integrate(qpos[:, qa : qa + 4], qvel[:, da : da + 3], dt_)
        else:
            out[:, qa] = qpos[:, qa] + dt_ * qvel[:, da]
    return out


def advance(m, d, act_dot, qacc, qvel=None, dt_=None):
    mj = m.mj
    if dt_ is None:
        dt_ = m.timestep
    act = d.act
    if m.na:
        inf = float("inf")
        act = d.act.clone()
        for i in range(m.nu):
            adr = int(mj.actuator_actadr[i])
            if adr < 0:
                continue
            for a in range(adr, adr + int(mj.actuator_actnum[i])):
                if int(mj.actuator_dyntype[i]) == 3:
                    tau = torch.clamp_min(m.actuator_dynprm[i, 0], mjMINVAL)
                    na_ = d.act[:, a] + act_dot[:, a] * tau * (1 - torch.exp(-dt_ / tau))
                else:
                    na_ = d.act[:, a] + act_dot[:, a] * dt_
                if int(mj.actuator_actlimited[i]):
                    na_ = torch.clamp(na_, m.actuator_actrange[i, 0], m.actuator_actrange[i, 1])
                act[:, a] = na_
    new_qvel = d.qvel + qacc * dt_
    qvel_for_pos = new_qvel if qvel is None else qvel
    d.qpos = _integrate_pos(m, d.qpos, qvel_for_pos, dt_)
    d.qvel, d.act = new_qvel, act
    d.time = d.time + dt_
    return d


def euler(m, d):
    qacc = d.qacc
    if not (m.disableflags & DSBL_EULERDAMP):
        qM = d.qM + torch.diag(m.timestep * m.dof_damping)
        L = smooth.factor_m(m, d, qM)
        qacc = smooth.solve_m(m, d, d.qfrc_smooth + d.qfrc_constraint, L)
    return advance(m, d, d.act_dot, qacc)


def deriv_smooth_vel(m, d):
    """derivative.deriv_smooth_vel :22-68."""
    mj = m.mj
    dt = d.qpos.dtype
    qderiv = None
    if not (m.disableflags & DSBL_ACTUATION):
        bias_vel = m.actuator_biasprm[:, 2] * torch.as_tensor(np.asarray(mj.actuator_biastype) == 1).to(dt)
        gain_vel = m.actuator_gainprm[:, 2] * torch.as_tensor(np.asarray(mj.actuator_gaintype) == 1).to(dt)
        ctrl = d.ctrl.clone()
        for i in range(m.nu):
            if int(mj.actuator_dyntype[i]) != 0:
                # reference: torch.where(dyn_mask, d.act, ctrl) (well defined only when na == nu)
                ctrl[:, i] = d.act[:, int(mj.actuator_actadr[i])]
        vel = bias_vel + gain_vel * ctrl
        mom = d.actuator_moment
        qderiv = torch.matmul(mom.transpose(-1, -2), mom * vel.unsqueeze(-1))
    if not (m.disableflags & DSBL_DAMPER):
        dd = torch.diag(m.dof_damping)
        qderiv = -dd if qderiv is None else qderiv - dd
    if m.ntendon:
        qderiv = qderiv - torch.matmul(d.ten_J.transpose(-1, -2) * m.tendon_damping, d.ten_J)
    if not (m.disableflags & (DSBL_DAMPER | DSBL_SPRING)) and m.has_fluid:
        raise NotImplementedError("fluid drag not supported for implicitfast")
    return qderiv


def implicit(m, d):
    qderiv = deriv_smooth_vel(m, d)
    qacc = d.qacc
    if qderiv is not None:
        qm = d.qM - m.timestep * qderiv
        L = M.small_cholesky(qm)
        qacc = M.small_cholesky_solve(d.qfrc_smooth + d.qfrc_constraint, L)
    return advance(m, d, d.act_dot, qacc)


def rungekutta4(m, d, fixed_iterations=False):
    d_t0 = d.clone()
    A = [0.5, 0.5, 1.0]
    Bw = [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0]
    C = [0.5, 0.5, 1.0]
    h = m.timestep
    kqvel = d.qvel
    qvel, qacc, act_dot = Bw[0] * kqvel, Bw[0] * d.qacc, Bw[0] * d.act_dot
    d = d.clone()
    for i in range(3):
        a, b = A[i], Bw[i + 1]
        dqvel, dqacc, dact_dot = a * kqvel, a * d.qacc, a * d.act_dot
        kqpos = _integrate_pos(m, d_t0.qpos, dqvel, h)
        kact = d_t0.act + dact_dot * h
        kqvel = d_t0.qvel + dqacc * h
        d.qpos, d.qvel, d.act, d.time = kqpos, kqvel, kact, d_t0.time + C[i] * h
        d = forward(m, d, fixed_iterations=fixed_iterations)
        qvel = qvel + b * kqvel
        qacc = qacc + b * d.qacc
        act_dot = act_dot + b * d.act_dot
    return advance(m, d_t0, act_dot, qacc, qvel)


def step(m, d, fixed_iterations=False):
    d = d.clone()
    d = check_state(m, d)
    d = forward(m, d, fixed_iterations=fixed_iterations)
    if m.integrator == 0:
        d = euler(m, d)
    elif m.integrator == 1:
        d = rungekutta4(m, d, fixed_iterations=fixed_iterations)
    elif m.integrator == 3:
        d = implicit(m, d)
    else:
        raise NotImplementedError(f"integrator {m.integrator} not implemented.")
    return d
