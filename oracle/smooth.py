"""ORACLE (test infrastructure): smooth dynamics, batched over a leading env dim.

Follows reference ``mujoco_torch/_src/smooth.py`` (kinematics :34-207, com_pos :210-288, crb :291-308,
factor_m :311-332, solve_m :335-354, com_vel :385-424, rne :427-467, tendon :470-521,
transmission :535-591) and ``support.py`` (make_m :50-80, jac :138-153, xfrc_accumulate :184-194).
The reference's tree scans are replaced by loops in body-index order (parents precede children).
"""

import torch

from . import omath as M
from .model import BALL, FREE, HINGE, SLIDE, DSBL_GRAVITY, mjMINVAL


def kinematics(m, d):
    mj = m.mj
    B = d.qpos.shape[0]
    dt = d.qpos.dtype
    qpos = d.qpos.clone()
    xpos = torch.zeros((B, m.nbody, 3), dtype=dt)
    xquat = torch.zeros((B, m.nbody, 4), dtype=dt)
    xanchor = torch.zeros((B, m.njnt, 3), dtype=dt)
    xaxis = torch.zeros((B, m.njnt, 3), dtype=dt)
    for b in range(m.nbody):
        pos = m.body_pos[b].expand(B, 3)
        quat = m.body_quat[b].expand(B, 4)
        if b > 0:
            p = int(mj.body_parentid[b])
            pos = xpos[:, p] + M.rotate(pos, xquat[:, p])
            quat = M.quat_mul(xquat[:, p], quat)
        for j in range(mj.body_jntadr[b], mj.body_jntadr[b] + mj.body_jntnum[b]):
            jt, qa = int(mj.jnt_type[j]), int(mj.jnt_qposadr[j])
            if jt == FREE:
                anchor = qpos[:, qa : qa + 3]
                axis = torch.tensor([0.0, 0.0, 1.0], dtype=dt).expand(B, 3)
            else:
                anchor = M.rotate(m.jnt_pos[j].expand(B, 3), quat) + pos
                axis = M.rotate(m.jnt_axis[j].expand(B, 3), quat)
            xanchor[:, j], xaxis[:, j] = anchor, axis
            if jt == FREE:
                pos = qpos[:, qa : qa + 3]
                quat = M.normalize(qpos[:, qa + 3 : qa + 7])
                qpos[:, qa + 3 : qa + 7] = quat
            elif jt == BALL:
                qloc = M.normalize(qpos[:, qa : qa + 4])
                qpos[:, qa : qa + 4] = qloc
                quat = M.quat_mul(quat, qloc)
                pos = anchor - M.rotate(m.jnt_pos[j].expand(B, 3), quat)
            elif jt == HINGE:
                angle = qpos[:, qa] - m.qpos0[qa]
                qloc = M.axis_angle_to_quat(m.jnt_axis[j].expand(B, 3), angle)
                quat = M.quat_mul(quat, qloc)
                pos = anchor - M.rotate(m.jnt_pos[j].expand(B, 3), quat)
            elif jt == SLIDE:
                pos = pos + axis * (qpos[:, qa] - m.qpos0[qa]).unsqueeze(-1)
        xpos[:, b], xquat[:, b] = pos, quat
    xmat = M.quat_to_mat(xquat)

    if m.nmocap:
        for b in range(m.nbody):
            mid = int(mj.body_mocapid[b])
            if mid >= 0:
                xpos[:, b] = d.mocap_pos[:, mid]
                mq = M.normalize(d.mocap_quat[:, mid])
                xquat[:, b] = mq
                xmat[:, b] = M.quat_to_mat(mq)

    def l2g(bodyid, lpos, lquat):
        bid = torch.as_tensor(bodyid, dtype=torch.long)
        wp, wq = xpos[:, bid], xquat[:, bid]
        pos = wp + M.rotate(lpos.expand(B, *lpos.shape), wq)
        mat = M.quat_to_mat(M.quat_mul(wq, lquat.expand(B, *lquat.shape)))
        return pos, mat

    d.qpos, d.xanchor, d.xaxis, d.xpos, d.xquat, d.xmat = qpos, xanchor, xaxis, xpos, xquat, xmat
    d.xipos, d.ximat = l2g(range(m.nbody), m.body_ipos, m.body_iquat)
    if m.ngeom:
        d.geom_xpos, d.geom_xmat = l2g(mj.geom_bodyid, m.geom_pos, m.geom_quat)
    if m.nsite:
        d.site_xpos, d.site_xmat = l2g(mj.site_bodyid, m.site_pos, m.site_quat)
    return d


def com_pos(m, d):
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    pos = d.xipos * m.body_mass[None, :, None]
    mass = m.body_mass.clone()
    pos = pos.clone()
    for b in range(m.nbody - 1, 0, -1):
        p = int(mj.body_parentid[b])
        pos[:, p] = pos[:, p] + pos[:, b]
        mass[p] = mass[p] + mass[b]
    cond = (mass < mjMINVAL)[None, :, None]
    subtree_com = torch.where(cond, d.xipos, pos / torch.clamp_min(mass, mjMINVAL)[None, :, None])

    root = torch.as_tensor(mj.body_rootid, dtype=torch.long)
    root_com = subtree_com[:, root]
    off = d.xipos - root_com
    # inert_com (:236-243): h = cross(off, -I) rows; inert = ximat diag(I) ximat^T + h h^T m
    eye = torch.eye(3, dtype=dt)
    h = M.cross(off[:, :, None, :].expand(B, m.nbody, 3, 3), -eye.expand(B, m.nbody, 3, 3))
    inert = torch.matmul(d.ximat * m.body_inertia[None, :, None, :], d.ximat.transpose(-1, -2))
    inert = inert + torch.matmul(h, h.transpose(-1, -2)) * m.body_mass[None, :, None, None]
    ti, tj = [0, 1, 2, 0, 0, 1], [0, 1, 2, 1, 2, 2]
    cinert = torch.cat([inert[:, :, ti, tj], off * m.body_mass[None, :, None], m.body_mass[None, :, None].expand(B, m.nbody, 1)], dim=-1)

    cdof = torch.zeros((B, m.nv, 6), dtype=dt)
    for j in range(m.njnt):
        jt, da, b = int(mj.jnt_type[j]), int(mj.jnt_dofadr[j]), int(mj.jnt_bodyid[j])
        offset = root_com[:, b] - d.xanchor[:, j]
        xmat_t = d.xmat[:, b].transpose(-1, -2)  # rows = body axes
        if jt == FREE:
            cdof[:, da : da + 3, 3:] = eye
            cdof[:, da + 3 : da + 6, :3] = xmat_t
            cdof[:, da + 3 : da + 6, 3:] = M.cross(xmat_t, offset[:, None, :].expand(B, 3, 3))
        elif jt == BALL:
            cdof[:, da : da + 3, :3] = xmat_t
            cdof[:, da : da + 3, 3:] = M.cross(xmat_t, offset[:, None, :].expand(B, 3, 3))
        elif jt == HINGE:
            cdof[:, da, :3] = d.xaxis[:, j]
            cdof[:, da, 3:] = M.cross(d.xaxis[:, j], offset)
        elif jt == SLIDE:
            cdof[:, da, 3:] = d.xaxis[:, j]
    d.subtree_com, d.cinert, d.cdof = subtree_com, cinert, cdof
    return d


def crb(m, d):
    mj = m.mj
    crb_body = d.cinert.clone()
    for b in range(m.nbody - 1, 0, -1):
        p = int(mj.body_parentid[b])
        crb_body[:, p] = crb_body[:, p] + crb_body[:, b]
    crb_body[:, 0] = 0.0
    dof_body = torch.as_tensor(mj.dof_bodyid, dtype=torch.long)
    crb_cdof = M.inert_mul(crb_body[:, dof_body], d.cdof)
    # support.make_m dense (:61-69)
    qm = torch.matmul(crb_cdof, d.cdof.transpose(-1, -2))
    qm = qm + torch.diag(m.dof_armature)
    qm = qm * m.m_mask
    qm = qm + torch.tril(qm, -1).transpose(-1, -2)
    d.crb, d.qM = crb_body, qm
    return d


def tendon(m, d):
    if not m.ntendon:
        return d
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    ten_length = torch.zeros((B, m.ntendon), dtype=dt)
    ten_J = torch.zeros((B, m.ntendon, m.nv), dtype=dt)
    for t in range(m.ntendon):
        for w in range(mj.tendon_adr[t], mj.tendon_adr[t] + mj.tendon_num[t]):
            if int(mj.wrap_type[w]) != 1:
                continue
            j = int(mj.wrap_objid[w])
            coef = float(mj.wrap_prm[w])
            ten_length[:, t] = ten_length[:, t] + coef * d.qpos[:, mj.jnt_qposadr[j]]
            ten_J[:, t, mj.jnt_dofadr[j]] = coef
    d.ten_length, d.ten_J = ten_length, ten_J
    return d


def tendon_armature(m, d):
    if not m.ntendon:
        return d
    JTAJ = torch.matmul(d.ten_J.transpose(-1, -2), d.ten_J * m.tendon_armature[None, :, None])
    d.qM = d.qM + JTAJ
    return d


def factor_m(m, d, qM=None):
    L = M.small_cholesky(d.qM if qM is None else qM)
    if qM is None:
        d.qLD = L
    return L


def solve_m(m, d, x, qLD=None):
    return M.small_cholesky_solve(x, d.qLD if qLD is None else qLD)


def transmission(m, d):
    if not m.nu:
        return d
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    length = torch.zeros((B, m.nu), dtype=dt)
    moment = torch.zeros((B, m.nu, m.nv), dtype=dt)
    for i in range(m.nu):
        trntype, trnid = int(mj.actuator_trntype[i]), int(mj.actuator_trnid[i, 0])
        gear = m.actuator_gear[i]
        if trntype == 3:  # TENDON
            length[:, i] = d.ten_length[:, trnid] * gear[0]
            moment[:, i] = d.ten_J[:, trnid] * gear[0]
            continue
        jt, da, qa = int(mj.jnt_type[trnid]), int(mj.jnt_dofadr[trnid]), int(mj.jnt_qposadr[trnid])
        if jt == FREE:
            if trntype == 1:  # JOINTINPARENT
                quat_neg = M.quat_inv(d.qpos[:, qa + 3 : qa + 7])
                gearaxis = M.rotate(gear[3:].expand(B, 3), quat_neg)
                moment[:, i, da : da + 3] = gear[:3]
                moment[:, i, da + 3 : da + 6] = gearaxis
            else:
                moment[:, i, da : da + 6] = gear[:6]
        elif jt == BALL:
            q = d.qpos[:, qa : qa + 4]
            axis, angle = M.quat_to_axis_angle(q)
            gearaxis = gear[:3].expand(B, 3)
            if trntype == 1:
                gearaxis = M.rotate(gearaxis, M.quat_inv(q))
            length[:, i] = (axis * angle.unsqueeze(-1) * gearaxis).sum(-1)
            moment[:, i, da : da + 3] = gearaxis
        else:
            length[:, i] = d.qpos[:, qa] * gear[0]
            moment[:, i, da] = gear[0]
    d.actuator_length, d.actuator_moment = length, moment
    return d


def com_vel(m, d):
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    cvel = torch.zeros((B, m.nbody, 6), dtype=dt)
    cdof_dot = torch.zeros((B, m.nv, 6), dtype=dt)
    for b in range(1, m.nbody):
        v = cvel[:, int(mj.body_parentid[b])].clone()
        for j in range(mj.body_jntadr[b], mj.body_jntadr[b] + mj.body_jntnum[b]):
            jt, da = int(mj.jnt_type[j]), int(mj.jnt_dofadr[j])
            if jt == FREE:
                v = v + (d.cdof[:, da : da + 3] * d.qvel[:, da : da + 3, None]).sum(1)
                cdof_dot[:, da + 3 : da + 6] = M.motion_cross(v[:, None, :].expand(B, 3, 6), d.cdof[:, da + 3 : da + 6])
                v = v + (d.cdof[:, da + 3 : da + 6] * d.qvel[:, da + 3 : da + 6, None]).sum(1)
            else:
                w = 3 if jt == BALL else 1
                cdof_dot[:, da : da + w] = M.motion_cross(v[:, None, :].expand(B, w, 6), d.cdof[:, da : da + w])
                v = v + (d.cdof[:, da : da + w] * d.qvel[:, da : da + w, None]).sum(1)
        cvel[:, b] = v
    d.cvel, d.cdof_dot = cvel, cdof_dot
    return d


def rne(m, d):
    mj = m.mj
    B, dt = d.qpos.shape[0], d.qpos.dtype
    cacc = torch.zeros((B, m.nbody, 6), dtype=dt)
    if not (m.disableflags & DSBL_GRAVITY):
        cacc[:, 0, 3:] = -m.gravity
    for b in range(1, m.nbody):
        a = cacc[:, int(mj.body_parentid[b])]
        da, dn = int(mj.body_dofadr[b]), int(mj.body_dofnum[b])
        if dn:
            a = a + (d.cdof_dot[:, da : da + dn] * d.qvel[:, da : da + dn, None]).sum(1)
        cacc[:, b] = a
    # reference scan for the world body also adds nothing but carries -gravity: bodies without dofs keep parent's
    frc = M.inert_mul(d.cinert, cacc) + M.motion_cross_force(d.cvel, M.inert_mul(d.cinert, d.cvel))
    cfrc = frc.clone()
    for b in range(m.nbody - 1, 0, -1):
        p = int(mj.body_parentid[b])
        cfrc[:, p] = cfrc[:, p] + cfrc[:, b]
    dof_body = torch.as_tensor(mj.dof_bodyid, dtype=torch.long)
    d.qfrc_bias = (d.cdof * cfrc[:, dof_body]).sum(-1)
    return d


def jac(m, d, point, body_id):
    """support.jac :138-153.  point (B,3); returns jacp, jacr (B, nv, 3)."""
    mask = m.body_dofmask[body_id].to(point.dtype)[None, :, None]
    root = int(m.mj.body_rootid[body_id])
    offset = point - d.subtree_com[:, root]
    jacp = d.cdof[:, :, 3:] + M.cross(d.cdof[:, :, :3], offset[:, None, :].expand_as(d.cdof[:, :, :3]))
    return jacp * mask, d.cdof[:, :, :3] * mask


def apply_ft(m, d, force, torque, point, body_id):
    jacp, jacr = jac(m, d, point, body_id)
    return (jacp * force[:, None, :]).sum(-1) + (jacr * torque[:, None, :]).sum(-1)


def xfrc_accumulate(m, d):
    out = torch.zeros_like(d.qvel)
    for b in range(m.nbody):
        out = out + apply_ft(m, d, d.xfrc_applied[:, b, :3], d.xfrc_applied[:, b, 3:], d.xipos[:, b], b)
    return out
