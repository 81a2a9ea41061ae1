"""ORACLE (test infrastructure): constraint row assembly, batched over envs.

Follows reference ``mujoco_torch/_src/constraint.py``: _kbi :69-113, equality rows :116-212 / :254-299,
friction :215-251, limits :302-405, contacts :408-583, make_constraint :600-768.
Row order: connect, weld, joint-eq, dof friction, tendon friction, ball limits, slide/hinge limits,
tendon limits, condim-1 contacts, then condim 3, 4, 6 contacts.
"""

import numpy as np
import torch

from . import omath as M
from . import smooth
from .model import (BALL, DSBL_EQUALITY, DSBL_FRICTIONLOSS, DSBL_LIMIT, DSBL_REFSAFE, HINGE, SLIDE, mjMINVAL)

mjMINIMP, mjMAXIMP = 1e-4, 0.9999


def kbi(m, solref, solimp, pos):
    """_kbi :69-113.  solref (..., 2), solimp (..., 5), pos (...)."""
    timeconst, dampratio = solref[..., 0], solref[..., 1]
    if not (m.disableflags & DSBL_REFSAFE):
        timeconst = torch.clamp_min(timeconst, 2 * m.timestep) * (timeconst > 0)
    dmin, dmax, width, mid, power = (solimp[..., i] for i in range(5))
    dmin = torch.clamp(dmin, mjMINIMP, mjMAXIMP)
    dmax = torch.clamp(dmax, mjMINIMP, mjMAXIMP)
    width = torch.clamp_min(width, mjMINVAL)
    mid = torch.clamp(mid, mjMINIMP, mjMAXIMP)
    power = torch.clamp_min(power, 1.0)
    k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio)
    b = 2 / (dmax * timeconst)
    k = torch.where(dampratio <= 0, -dampratio / (dmax * dmax), k)
    b = torch.where(timeconst <= 0, -timeconst / dmax, b)
    imp_x = torch.abs(pos) / width
    imp_a = (1.0 / torch.pow(mid, power - 1)) * torch.pow(imp_x, power)
    imp_b = 1 - (1.0 / torch.pow(1 - mid, power - 1)) * torch.pow(1 - imp_x, power)
    imp_y = torch.where(imp_x < mid, imp_a, imp_b)
    imp = dmin + imp_y * (dmax - dmin)
    imp = torch.minimum(torch.maximum(imp, dmin), dmax)  # torch.clamp(imp, dmin, dmax)
    imp = torch.where(imp_x > 1.0, dmax, imp)
    return k, b, imp


class _Rows:
    def __init__(self):
        self.J, self.pos, self.pos_norm, self.invweight, self.solref, self.solimp, self.floss = [], [], [], [], [], [], []

    def add(self, J, pos, pos_norm, invweight, solref, solimp, floss=None):
        """J (B,r,nv); pos,pos_norm (B,r); invweight (r,) or (B,r); solref (r,2); solimp (r,5)."""
        B, r = pos.shape
        self.J.append(J)
        self.pos.append(pos)
        self.pos_norm.append(pos_norm)
        self.invweight.append(invweight.expand(B, r))
        self.solref.append(solref.expand(B, r, 2))
        self.solimp.append(solimp.expand(B, r, 5))
        self.floss.append(torch.zeros_like(pos) if floss is None else floss.expand(B, r))


def make_constraint(m, d):
    mj = m.mj
    B, dt, nv = d.qpos.shape[0], d.qpos.dtype, m.nv
    if m.nefc == 0:
        d.efc_J = torch.zeros((B, 0, nv), dtype=dt)
        d.efc_D = d.efc_aref = d.efc_frictionloss = torch.zeros((B, 0), dtype=dt)
        return d
    rows = _Rows()
    fl = m.disableflags
    L = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.long)

    # ---- equality: connect (:116-156)
    if not (fl & DSBL_EQUALITY) and m.neq:
        et = np.asarray(mj.eq_type)
        for eqt in (0, 1):
            for i in np.nonzero(et == eqt)[0]:
                id1, id2 = int(mj.eq_obj1id[i]), int(mj.eq_obj2id[i])
                data = m.eq_data[i]
                active = d.eq_active[:, i].to(dt)
                anchor1, anchor2 = data[0:3], data[3:6]
                if eqt == 0:
                    pos1 = (d.xmat[:, id1] * anchor1).sum(-1) + d.xpos[:, id1]
                    pos2 = (d.xmat[:, id2] * anchor2).sum(-1) + d.xpos[:, id2]
                    cpos = pos1 - pos2
                    jacp1, _ = smooth.jac(m, d, pos1, id1)
                    jacp2, _ = smooth.jac(m, d, pos2, id2)
                    j = (jacp1 - jacp2).transpose(-1, -2)
                    pn = M.norm(cpos)[:, None].expand(B, 3)
                    a3 = active[:, None]
                    iw = (m.body_invweight0[id1, 0] + m.body_invweight0[id2, 0]).expand(3)
                    rows.add(j * a3[:, :, None], cpos * a3, pn * a3, iw, m.eq_solref[i].expand(3, 2), m.eq_solimp[i].expand(3, 5))
                else:
                    relpose, torquescale = data[6:10], data[10]
                    pos1 = (d.xmat[:, id1] * anchor2).sum(-1) + d.xpos[:, id1]
                    pos2 = (d.xmat[:, id2] * anchor1).sum(-1) + d.xpos[:, id2]
                    cpos = pos1 - pos2
                    jacp1, jacr1 = smooth.jac(m, d, pos1, id1)
                    jacp2, jacr2 = smooth.jac(m, d, pos2, id2)
                    jacdifp = jacp1 - jacp2
                    jacdifr = (jacr1 - jacr2) * torquescale
                    quat = M.quat_mul(d.xquat[:, id1], relpose.expand(B, 4))
                    quat1 = M.quat_inv(d.xquat[:, id2])
                    crot = M.quat_mul(quat1, quat)[:, 1:]
                    q1e = quat1[:, None, :].expand(B, nv, 4)
                    qe = quat[:, None, :].expand(B, nv, 4)
                    jacdifr = 0.5 * M.quat_mul(M.quat_mul_axis(q1e, jacdifr), qe)[..., 1:]
                    j = torch.cat((jacdifp.transpose(-1, -2), jacdifr.transpose(-1, -2)), dim=1)
                    pos = torch.cat((cpos, crot * torquescale), dim=-1)
                    pn = M.norm(pos)[:, None].expand(B, 6)
                    a6 = active[:, None]
                    iw = (m.body_invweight0[id1] + m.body_invweight0[id2]).repeat_interleave(3)
                    rows.add(j * a6[:, :, None], pos * a6, pn * a6, iw, m.eq_solref[i].expand(6, 2), m.eq_solimp[i].expand(6, 5))
        # joint equality (:254-299)
        for i in np.nonzero(et == 2)[0]:
            id1, id2 = int(mj.eq_obj1id[i]), int(mj.eq_obj2id[i])
            data = m.eq_data[i]
            active = d.eq_active[:, i].to(dt)
            qa1, da1 = int(mj.jnt_qposadr[id1]), int(mj.jnt_dofadr[id1])
            # the reference indexes jnt_*adr[id2] even when id2 == -1 (numpy wrap-around) and masks the result
            qa2, da2 = int(mj.jnt_qposadr[id2]), int(mj.jnt_dofadr[id2])
            has2 = float(id2 > -1)
            pos1, pos2 = d.qpos[:, qa1], d.qpos[:, qa2] * has2
            ref1, ref2 = m.qpos0[qa1], m.qpos0[qa2] * has2
            dif = pos2 - ref2
            dif_power = torch.pow(dif[:, None], torch.arange(0, 5, dtype=dt))
            deriv = (data[1:5] * dif_power[:, :4] * torch.arange(1, 5, dtype=dt)).sum(-1)
            j = torch.zeros((B, nv), dtype=dt)
            j[:, da1] = 1.0
            j[:, da2] = -deriv
            pos = pos1 - ref1 - (data[:5] * dif_power).sum(-1)
            iw = (m.dof_invweight0[da1] + m.dof_invweight0[da2] * has2).expand(1)
            pos = (pos * active)[:, None]
            rows.add((j * active[:, None])[:, None, :], pos, pos, iw, m.eq_solref[i].expand(1, 2), m.eq_solimp[i].expand(1, 5))

    # ---- friction loss (:215-251)
    if not (fl & DSBL_FRICTIONLOSS):
        dof_ids = np.nonzero(np.asarray(mj.dof_frictionloss) > 0)[0]
        ten_ids = np.nonzero(np.asarray(mj.tendon_frictionloss) > 0)[0]
        n = dof_ids.size + ten_ids.size
        if n:
            eye = torch.eye(nv, dtype=dt)
            j = eye[L(dof_ids)].expand(B, dof_ids.size, nv)
            if ten_ids.size:
                j = torch.cat([j, d.ten_J[:, L(ten_ids)]], dim=1)
            cat = lambda a, b_: torch.cat([a[L(dof_ids)], b_[L(ten_ids)]]) if ten_ids.size else a[L(dof_ids)]
            z = torch.zeros((B, n), dtype=dt)
            rows.add(j, z, z, cat(m.dof_invweight0, m.tendon_invweight0), cat(m.dof_solref, m.tendon_solref_fri),
                     cat(m.dof_solimp, m.tendon_solimp_fri), cat(m.dof_frictionloss, m.tendon_frictionloss))

    # ---- limits
    if not (fl & DSBL_LIMIT):
        jt, lim = np.asarray(mj.jnt_type), np.asarray(mj.jnt_limited).astype(bool)
        for i in np.nonzero((jt == BALL) & lim)[0]:  # :302-335
            qa, da = int(mj.jnt_qposadr[i]), int(mj.jnt_dofadr[i])
            axis, angle = M.quat_to_axis_angle(d.qpos[:, qa : qa + 4])
            j = torch.zeros((B, nv), dtype=dt)
            j[:, da : da + 3] = -axis
            pos = torch.max(m.jnt_range[i]) - angle - m.jnt_margin[i]
            act = (pos < 0).to(dt)
            pos = (pos * act)[:, None]
            rows.add((j * act[:, None])[:, None, :], pos, pos, m.dof_invweight0[da].expand(1), m.jnt_solref[i].expand(1, 2),
                     m.jnt_solimp[i].expand(1, 5))
        ids = np.nonzero(((jt == SLIDE) | (jt == HINGE)) & lim)[0]  # :338-372
        if ids.size:
            qa, da = L(mj.jnt_qposadr[ids]), L(mj.jnt_dofadr[ids])
            q = d.qpos[:, qa]
            dist_min = q - m.jnt_range[L(ids), 0]
            dist_max = m.jnt_range[L(ids), 1] - q
            val = (dist_min < dist_max).to(dt) * 2 - 1
            pos = torch.minimum(dist_min, dist_max) - m.jnt_margin[L(ids)]
            act = (pos < 0).to(dt)
            j = torch.zeros((B, ids.size, nv), dtype=dt)
            j[:, torch.arange(ids.size), da] = val
            pos = pos * act
            rows.add(j * act[:, :, None], pos, pos, m.dof_invweight0[da], m.jnt_solref[L(ids)], m.jnt_solimp[L(ids)])
        tid = np.nonzero(np.asarray(mj.tendon_limited))[0]  # :375-405
        if tid.size:
            t = L(tid)
            length = d.ten_length[:, t]
            dist_min = length - m.tendon_range[t, 0]
            dist_max = m.tendon_range[t, 1] - length
            pos = torch.minimum(dist_min, dist_max) - m.tendon_margin[t]
            act = (pos < 0).to(dt)
            sign = ((dist_min < dist_max).to(dt) * 2 - 1) * act
            pos = pos * act
            rows.add(d.ten_J[:, t] * sign[:, :, None], pos, pos, m.tendon_invweight0[t], m.tendon_solref_lim[t], m.tendon_solimp_lim[t])

    # ---- contacts
    if m.ncon:
        c = m.contact
        n1, n3, n4, n6 = m.condim_counts
        geom_body = np.asarray(mj.geom_bodyid)
        elliptic = m.cone == 1
        start = 0
        def contact_rows(condim, ci, pi):
            """Rows of contact slot `ci` (dist / pos / frame) with the parameters and bodies of contact-table entry `pi`."""
            dist = d.contact_dist[:, ci] - c["includemargin"][pi]
            body1, body2 = int(geom_body[int(c["geom1"][pi])]), int(geom_body[int(c["geom2"][pi])])
            pos_c, frame = d.contact_pos[:, ci], d.contact_frame[:, ci]
            jacp2, jacr2 = smooth.jac(m, d, pos_c, body2)
            jacp1, jacr1 = smooth.jac(m, d, pos_c, body1)
            diff = torch.matmul(frame, (jacp2 - jacp1).transpose(-1, -2))  # (B,3,nv)
            t = m.body_invweight0[body1, 0] + m.body_invweight0[body2, 0]
            active = (dist < 0).to(dt)
            fri = c["friction"][pi]
            if condim == 1:  # :408-451
                pos = (dist * active)[:, None]
                return (diff[:, 0:1] * active[:, None, None], pos, pos, t.expand(1), c["solref"][pi].expand(1, 2), c["solimp"][pi].expand(1, 5))
            if not elliptic:  # :454-516
                if condim > 3:
                    diff = torch.cat((diff, torch.matmul(frame, (jacr2 - jacr1).transpose(-1, -2))), dim=1)
                ne = (condim - 1) * 2
                f = fri[: condim - 1].repeat_interleave(2) * torch.tensor([1.0, -1.0] * (condim - 1), dtype=dt)
                j = diff[:, 0:1] + diff[:, 1:condim].repeat_interleave(2, dim=1) * f[None, :, None]
                mu = fri[0]
                iw = (t + mu * mu * t) * 2 * mu * mu / m.impratio
                pos = dist[:, None].expand(B, ne) * active[:, None]
                return (j * active[:, None, None], pos, pos, iw.expand(ne), c["solref"][pi].expand(ne, 2), c["solimp"][pi].expand(ne, 5))
            # elliptic :519-583
            j = diff
            if condim > 3:
                j = torch.cat((j, torch.matmul(frame, (jacr2 - jacr1).transpose(-1, -2))[:, : condim - 3]), dim=1)
            srf = c["solreffriction"][pi]
            srf = srf + c["solref"][pi] * float(not bool((srf != 0).any()))
            solref = torch.cat((c["solref"][pi][None], srf[None].expand(condim - 1, 2)), dim=0)
            fr2 = torch.square(fri[0]) / torch.square(fri[1 : condim - 1])
            iw = torch.cat((t[None], (t / m.impratio)[None], (t / m.impratio) * fr2))
            pos_aref = torch.zeros((B, condim), dtype=dt)
            pos_aref[:, 0] = dist
            return (j * active[:, None, None], pos_aref * active[:, None], dist[:, None].expand(B, condim), iw, solref, c["solimp"][pi].expand(condim, 5))

        src = getattr(d, "contact_src", None) if getattr(m, "topk", 0) else None   # max_contact_points: candidate per (env, slot)
        for condim, count in ((1, n1), (3, n3), (4, n4), (6, n6)):
            for ci in range(start, start + count):
                if src is None:
                    rows.add(*contact_rows(condim, ci, ci))
                    continue
                # the slot holds a different candidate in different envs: evaluate each candidate that occurs, pick per env
                merged = None
                for v in torch.unique(src[:, ci]).tolist():
                    r = contact_rows(condim, ci, int(v))
                    nrow = r[1].shape[1]
                    r = (r[0], r[1], r[2], r[3].expand(B, nrow), r[4].expand(B, nrow, 2), r[5].expand(B, nrow, 5))
                    if merged is None:
                        merged = [x.clone() for x in r]
                    else:
                        sel = src[:, ci] == v
                        for x, y in zip(merged, r):
                            x[sel] = y[sel]
                rows.add(*merged)
            start += count

    J = torch.cat(rows.J, dim=1)
    pos, pos_norm = torch.cat(rows.pos, dim=1), torch.cat(rows.pos_norm, dim=1)
    invweight = torch.cat(rows.invweight, dim=1)
    solref, solimp = torch.cat(rows.solref, dim=1), torch.cat(rows.solimp, dim=1)
    k, b, imp = kbi(m, solref, solimp, pos_norm)
    r = torch.clamp_min(invweight * (1 - imp) / imp, mjMINVAL)
    aref = -b * torch.matmul(J, d.qvel[:, :, None]).squeeze(-1) - k * imp * pos
    assert J.shape[1] == m.nefc, (J.shape, m.nefc)
    d.efc_J, d.efc_D, d.efc_aref, d.efc_frictionloss = J, 1 / r, aref, torch.cat(rows.floss, dim=1)
    return d
