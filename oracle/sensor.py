"""ORACLE (test infrastructure): sensors, batched over envs.

Follows reference ``mujoco_torch/_src/sensor.py`` (sensor_pos :56-211, sensor_vel :214-343, sensor_acc
:346-440; mujoco>=3.3.7 accelerometer branch :36-38) and the analytic ray tests of ``ray.py``
(_ray_quad :28, _ray_plane :43, _ray_sphere :60, _ray_capsule :72, _ray_box :132, precompute_ray_data
:292-324, ray_precomputed :327-372).  Sensor types without a branch in the reference (touch, ...)
leave their slots untouched.
"""

import numpy as np
import torch

from . import omath as M
from .model import BOX, CAPSULE, DSBL_SENSOR, PLANE, SPHERE, mjMINVAL

# mjtSensor values (mujoco_torch_b200.mjmodel.SensorType)
S_ACCELEROMETER, S_VELOCIMETER, S_GYRO, S_MAGNETOMETER, S_RANGEFINDER = 1, 2, 3, 6, 7
S_JOINTPOS, S_JOINTVEL, S_ACTUATORPOS, S_ACTUATORVEL, S_ACTUATORFRC, S_JOINTACTFRC = 9, 10, 13, 14, 15, 16
S_TENDONPOS, S_TENDONVEL, S_BALLQUAT, S_BALLANGVEL = 11, 12, 18, 19
S_FRAMEPOS, S_FRAMEQUAT, S_FRAMEXAXIS, S_FRAMEYAXIS, S_FRAMEZAXIS, S_FRAMELINVEL, S_FRAMEANGVEL = 26, 27, 28, 29, 30, 31, 32
S_SUBTREECOM, S_CLOCK = 35, 45
S_FORCE, S_TORQUE, S_TENDONACTFRC, S_SUBTREELINVEL, S_SUBTREEANGMOM = 4, 5, 17, 36, 37
O_UNKNOWN, O_BODY, O_XBODY, O_GEOM, O_SITE, O_CAMERA = 0, 1, 2, 5, 6, 7
STAGE_POS, STAGE_VEL, STAGE_ACC = 1, 2, 3


def _frame(m, d, otype, oid):
    """World position / rotation of a sensor object (sensor.py:63-74 objtype_data)."""
    B = d.qpos.shape[0]
    if otype == O_BODY:
        return d.xipos[:, oid], d.ximat[:, oid]
    if otype == O_XBODY:
        return d.xpos[:, oid], d.xmat[:, oid]
    if otype == O_GEOM:
        return d.geom_xpos[:, oid], d.geom_xmat[:, oid]
    if otype == O_SITE:
        return d.site_xpos[:, oid], d.site_xmat[:, oid]
    if otype == O_UNKNOWN:
        return torch.zeros(B, 3, dtype=d.qpos.dtype), torch.eye(3, dtype=d.qpos.dtype).expand(B, 3, 3)
    raise NotImplementedError(f"sensor object type {otype}")


def _frame_quat(m, d, otype, oid):
    """sensor.py:161-176."""
    mj = m.mj
    B = d.qpos.shape[0]
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=d.qpos.dtype).expand(B, 4)
    if otype == O_XBODY:
        return d.xquat[:, oid]
    if otype == O_BODY:
        return M.quat_mul(d.xquat[:, oid], t(mj.body_iquat[oid]))
    if otype == O_GEOM:
        return M.quat_mul(d.xquat[:, int(mj.geom_bodyid[oid])], t(mj.geom_quat[oid]))
    if otype == O_SITE:
        return M.quat_mul(d.xquat[:, int(mj.site_bodyid[oid])], t(mj.site_quat[oid]))
    if otype == O_UNKNOWN:
        return t([1.0, 0, 0, 0])
    raise NotImplementedError(f"sensor object type {otype}")


def _obj_body(mj, otype, oid):
    if otype in (O_BODY, O_XBODY):
        return oid
    if otype == O_GEOM:
        return int(mj.geom_bodyid[oid])
    if otype == O_SITE:
        return int(mj.site_bodyid[oid])
    return 0
INF = float("inf")


def _cutoff(sensor, cutoff, datatype):
    """_apply_cutoff :41-53; sensor (B, n[,k]), cutoff (n,)."""
    c = torch.as_tensor(np.asarray(cutoff), dtype=sensor.dtype)
    while c.dim() < sensor.dim() - 1:
        c = c.unsqueeze(-1)
    c = c.expand(sensor.shape[1:]) if sensor.dim() > 1 else c
    if datatype == 0:
        return torch.where(c > 0, torch.maximum(torch.minimum(sensor, c), -c), sensor)
    if datatype == 1:
        return torch.where(c > 0, torch.minimum(sensor, c), sensor)
    return sensor


def _ray_quad(a, b, c):
    det = b * b - a * c
    det_2 = torch.sqrt(det)
    x0 = M.safe_div(-b - det_2, a)
    x1 = M.safe_div(-b + det_2, a)
    inv = lambda x: (det < mjMINVAL) | (x < 0)
    inf = torch.full_like(x0, INF)
    return torch.where(inv(x0), inf, x0), torch.where(inv(x1), inf, x1)


def _ray_plane(size, pnt, vec):
    x = -M.safe_div(pnt[..., 2], vec[..., 2])
    valid = vec[..., 2] <= -mjMINVAL
    valid = valid & (x >= 0)
    p = pnt[..., 0:2] + x.unsqueeze(-1) * vec[..., 0:2]
    valid = valid & ((size[0:2] <= 0) | (torch.abs(p) <= size[0:2])).all(-1)
    return torch.where(valid, x, torch.full_like(x, INF))


def _ray_sphere(size, pnt, vec):
    x0, x1 = _ray_quad(M.dot(vec, vec), M.dot(vec, pnt), M.dot(pnt, pnt) - size[0] * size[0])
    return torch.where(torch.isinf(x0), x1, x0)


def _ray_capsule(size, pnt, vec):
    a = M.dot(vec[..., 0:2], vec[..., 0:2])
    b = M.dot(vec[..., 0:2], pnt[..., 0:2])
    c = M.dot(pnt[..., 0:2], pnt[..., 0:2]) - size[0] * size[0]
    x0, x1 = _ray_quad(a, b, c)
    x = torch.where(torch.isinf(x0), x1, x0)
    inf = torch.full_like(x, INF)
    x = torch.where(torch.abs(pnt[..., 2] + x * vec[..., 2]) <= size[1], x, inf)
    zoff = torch.stack([torch.zeros_like(size[1]), torch.zeros_like(size[1]), size[1]])
    dif = pnt - zoff
    x0, x1 = _ray_quad(M.dot(vec, vec), M.dot(vec, dif), M.dot(dif, dif) - size[0] * size[0])
    x = torch.where((pnt[..., 2] + x0 * vec[..., 2] >= size[1]) & (x0 < x), x0, x)
    x = torch.where((pnt[..., 2] + x1 * vec[..., 2] >= size[1]) & (x1 < x), x1, x)
    dif = pnt + zoff
    x0, x1 = _ray_quad(M.dot(vec, vec), M.dot(vec, dif), M.dot(dif, dif) - size[0] * size[0])
    x = torch.where((pnt[..., 2] + x0 * vec[..., 2] <= -size[1]) & (x0 < x), x0, x)
    x = torch.where((pnt[..., 2] + x1 * vec[..., 2] <= -size[1]) & (x1 < x), x1, x)
    return x


def _ray_box(size, pnt, vec):
    iface = torch.tensor([[1, 2], [0, 2], [0, 1], [1, 2], [0, 2], [0, 1]], dtype=torch.long)
    x = torch.cat([M.safe_div(size - pnt, vec), -M.safe_div(size + pnt, vec)], dim=-1)
    p0 = pnt[..., iface[:, 0]] + x * vec[..., iface[:, 0]]
    p1 = pnt[..., iface[:, 1]] + x * vec[..., iface[:, 1]]
    valid = (torch.abs(p0) <= size[iface[:, 0]]) & (torch.abs(p1) <= size[iface[:, 1]]) & (x >= 0)
    return torch.where(valid, x, torch.full_like(x, INF)).min(-1).values


_RAY_FN = {PLANE: _ray_plane, SPHERE: _ray_sphere, CAPSULE: _ray_capsule, BOX: _ray_box}
# iteration order of the reference's _PRIMITIVE_RAY_FUNC dict (ray.py:283-290)
_RAY_ORDER = (PLANE, SPHERE, CAPSULE, 4, 5, BOX)


def _ray(m, d, pnt, vec, bodyexclude):
    """ray_precomputed with flg_static=True and one excluded body.  pnt, vec (B,3) -> dist (B,)."""
    mj = m.mj
    gb = np.asarray(mj.geom_bodyid)
    gfilter = gb != bodyexclude
    matid = np.asarray(mj.geom_matid)
    dyn = (matid != -1) | (np.asarray(mj.geom_rgba)[:, 3] != 0)  # materials are not compiled: matid == -1
    dists = []
    for gt in _RAY_ORDER:
        ids = np.nonzero(gfilter & (np.asarray(mj.geom_type) == gt))[0]
        if ids.size == 0:
            continue
        if gt not in _RAY_FN:
            raise NotImplementedError("ray test for ellipsoid/cylinder")
        for g in ids:
            xmat, xpos = d.geom_xmat[:, g], d.geom_xpos[:, g]
            gp = (xmat.transpose(-1, -2) * (pnt - xpos)[:, None, :]).sum(-1)
            gv = (xmat.transpose(-1, -2) * vec[:, None, :]).sum(-1)
            # the reference keeps geom sizes in float64 here (ray.py:320)
            size = torch.as_tensor(np.asarray(mj.geom_size[g], dtype=np.float64))
            dist = _RAY_FN[gt](size, gp.to(torch.float64), gv.to(torch.float64))
            if not dyn[g]:
                dist = torch.full_like(dist, INF)
            dists.append(dist)
    if not dists:
        return torch.full((pnt.shape[0],), -1.0, dtype=pnt.dtype)
    dists = torch.stack(dists, dim=-1)
    mind = dists.min(-1).values
    return torch.where(torch.isinf(mind), torch.full_like(mind, -1.0), mind)


def _write(d, adr, dim, vals):
    """scatter vals (B, n, dim) at sensor_adr."""
    sd = d.sensordata.clone()
    for k, a in enumerate(adr):
        sd[:, a : a + dim] = vals[:, k].reshape(vals.shape[0], dim).to(sd.dtype)
    d.sensordata = sd


def _stage(m, stage):
    mj = m.mj
    if m.disableflags & DSBL_SENSOR or m.nsensor == 0:
        return []
    st = np.asarray(mj.sensor_type)
    out = []
    for t in sorted(set(st[np.asarray(mj.sensor_needstage) == stage].tolist())):
        idx = np.nonzero(st == t)[0]
        out.append((int(t), idx))
    return out


def sensor_pos(m, d):
    mj = m.mj
    for t, idx in _stage(m, STAGE_POS):
        objid, adr, cutoff = mj.sensor_objid[idx], mj.sensor_adr[idx], mj.sensor_cutoff[idx]
        dtp = int(mj.sensor_datatype[idx][0])
        if t == S_MAGNETOMETER:
            v = torch.stack([(d.site_xmat[:, o].transpose(-1, -2) * m.magnetic).sum(-1) for o in objid], 1)
            _write(d, adr, 3, _cutoff(v, cutoff, dtp))
        elif t == S_RANGEFINDER:
            vals = []
            for o in objid:
                pnt, vec = d.site_xpos[:, o], d.site_xmat[:, o, :, 2]
                vals.append(_ray(m, d, pnt, vec, int(mj.site_bodyid[o])))
            v = torch.stack(vals, 1)
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_JOINTPOS:
            v = d.qpos[:, torch.as_tensor(mj.jnt_qposadr[objid], dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_ACTUATORPOS:
            v = d.actuator_length[:, torch.as_tensor(objid, dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_TENDONPOS:
            v = d.ten_length[:, torch.as_tensor(objid, dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_BALLQUAT:
            vals = []
            for o in objid:
                a = int(mj.jnt_qposadr[o])
                vals.append(M.normalize(d.qpos[:, a : a + 4]))
            _write(d, adr, 4, _cutoff(torch.stack(vals, 1), cutoff, dtp))
        elif t == S_SUBTREECOM:
            v = torch.stack([d.subtree_com[:, o] for o in objid], 1)
            _write(d, adr, 3, _cutoff(v, cutoff, dtp))
        elif t == S_CLOCK:
            v = d.time[:, None].expand(-1, len(idx))
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t in (S_FRAMEPOS, S_FRAMEXAXIS, S_FRAMEYAXIS, S_FRAMEZAXIS, S_FRAMEQUAT):
            vals = []
            for k, o in zip(idx, objid):
                ot, rt, rid = int(mj.sensor_objtype[k]), int(mj.sensor_reftype[k]), int(mj.sensor_refid[k])
                if t == S_FRAMEQUAT:
                    q = _frame_quat(m, d, ot, int(o))
                    if rid != -1:
                        r = _frame_quat(m, d, rt, rid)
                        q = M.quat_mul(M.quat_inv(r), q)
                    vals.append(q)
                    continue
                xpos, xmat = _frame(m, d, ot, int(o))
                rpos, rmat = _frame(m, d, rt, rid) if rid != -1 else (None, None)
                if t == S_FRAMEPOS:
                    v = xpos if rid == -1 else (rmat.transpose(-1, -2) * (xpos - rpos)[:, None, :]).sum(-1)
                else:
                    axis = xmat[:, :, {S_FRAMEXAXIS: 0, S_FRAMEYAXIS: 1, S_FRAMEZAXIS: 2}[t]]
                    v = axis if rid == -1 else (rmat.transpose(-1, -2) * axis[:, None, :]).sum(-1)
                vals.append(v)
            _write(d, adr, 4 if t == S_FRAMEQUAT else 3, _cutoff(torch.stack(vals, 1), cutoff, dtp))
    return d


def sensor_vel(m, d):
    mj = m.mj
    for t, idx in _stage(m, STAGE_VEL):
        objid, adr, cutoff = mj.sensor_objid[idx], mj.sensor_adr[idx], mj.sensor_cutoff[idx]
        dtp = int(mj.sensor_datatype[idx][0])
        if t in (S_VELOCIMETER, S_GYRO):
            vals = []
            for o in objid:
                b = int(mj.site_bodyid[o])
                rot_t = d.site_xmat[:, o].transpose(-1, -2)
                cvel = d.cvel[:, b]
                if t == S_VELOCIMETER:
                    dif = d.site_xpos[:, o] - d.subtree_com[:, int(mj.body_rootid[b])]
                    vec = cvel[:, 3:] - M.cross(dif, cvel[:, :3])
                else:
                    vec = cvel[:, :3]
                vals.append((rot_t * vec[:, None, :]).sum(-1))
            _write(d, adr, 3, _cutoff(torch.stack(vals, 1), cutoff, dtp))
        elif t == S_JOINTVEL:
            v = d.qvel[:, torch.as_tensor(mj.jnt_dofadr[objid], dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_ACTUATORVEL:
            v = d.actuator_velocity[:, torch.as_tensor(objid, dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_TENDONVEL:
            v = d.ten_velocity[:, torch.as_tensor(objid, dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_BALLANGVEL:
            vals = []
            for o in objid:
                a = int(mj.jnt_dofadr[o])
                vals.append(d.qvel[:, a : a + 3])
            _write(d, adr, 3, _cutoff(torch.stack(vals, 1), cutoff, dtp))
        elif t in (S_SUBTREELINVEL, S_SUBTREEANGMOM):
            # sensor.py:260-265 reads d.subtree_linvel / d.subtree_angmom, which nothing in the reference computes (the
            # smooth.subtree_vel call at :223-225 is behind a hasattr that is False): make_data's zeros (io.py:184-185)
            _write(d, adr, 3, _cutoff(torch.zeros(d.qpos.shape[0], len(idx), 3, dtype=d.qpos.dtype), cutoff, dtp))
        elif t in (S_FRAMELINVEL, S_FRAMEANGVEL):   # sensor.py:265-325
            vals = []
            for k, o in zip(idx, objid):
                ot, rt, rid = int(mj.sensor_objtype[k]), int(mj.sensor_reftype[k]), int(mj.sensor_refid[k])
                xpos_o, _ = _frame(m, d, ot, int(o))
                # refid == -1 indexes the *last* element in the reference (python negative index); it is masked out afterwards
                rr = rid if rid != -1 else {O_BODY: mj.nbody, O_XBODY: mj.nbody, O_GEOM: mj.ngeom, O_SITE: mj.nsite, O_UNKNOWN: 1}[rt] - 1
                xpos_r, xmat_r = _frame(m, d, rt, rr)
                ob, rb = _obj_body(mj, ot, int(o)), _obj_body(mj, rt, rr)
                cvel_o, cvel_r = d.cvel[:, ob], d.cvel[:, rb]
                off_o = xpos_o - d.subtree_com[:, int(mj.body_rootid[ob])]
                off_r = xpos_r - d.subtree_com[:, int(mj.body_rootid[rb])]
                ang, angr = cvel_o[:, :3], cvel_r[:, :3]
                rot = lambda v: (xmat_r.transpose(-1, -2) * v[:, None, :]).sum(-1)
                if t == S_FRAMELINVEL:
                    lin = cvel_o[:, 3:] - M.cross(off_o, ang)
                    linr = cvel_r[:, 3:] - M.cross(off_r, angr)
                    rel = lin - linr + M.cross(xpos_o - xpos_r, angr)
                    vals.append(rot(rel) if rid > -1 else lin)
                else:
                    vals.append(rot(ang - angr) if rid > -1 else ang)
            _write(d, adr, 3, _cutoff(torch.stack(vals, 1), cutoff, dtp))
    return d


def sensor_acc(m, d):
    mj = m.mj
    for t, idx in _stage(m, STAGE_ACC):
        objid, adr, cutoff = mj.sensor_objid[idx], mj.sensor_adr[idx], mj.sensor_cutoff[idx]
        dtp = int(mj.sensor_datatype[idx][0])
        if t == S_ACCELEROMETER:
            vals = []
            for o in objid:
                b = int(mj.site_bodyid[o])
                rot_t = d.site_xmat[:, o].transpose(-1, -2)
                cvel, cacc = d.cvel[:, b], d.cacc[:, b]
                dif = d.site_xpos[:, o] - d.subtree_com[:, int(mj.body_rootid[b])]
                rt = lambda v: (rot_t * v[:, None, :]).sum(-1)
                ang = rt(cvel[:, :3])
                lin = rt(cvel[:, 3:] - M.cross(dif, cvel[:, :3]))
                acc = rt(cacc[:, 3:] - M.cross(dif, cacc[:, :3]))
                vals.append(acc + M.cross(ang, lin))
            _write(d, adr, 3, _cutoff(torch.stack(vals, 1), cutoff, dtp))
        elif t == S_ACTUATORFRC:
            v = d.actuator_force[:, torch.as_tensor(objid, dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t == S_JOINTACTFRC:
            v = d.qfrc_actuator[:, torch.as_tensor(mj.jnt_dofadr[objid], dtype=torch.long)]
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        elif t in (S_FORCE, S_TORQUE):
            # sensor.py:394-416 rotates d.cfrc_int, which stays at make_data's zeros (io.py:182): there is no
            # smooth.rne_postconstraint behind the hasattr at :362-363
            _write(d, adr, 3, _cutoff(torch.zeros(d.qpos.shape[0], len(idx), 3, dtype=d.qpos.dtype), cutoff, dtp))
        elif t == S_TENDONACTFRC:   # sensor.py:421-423, mask built in device.py:572-579
            trn, tid = np.asarray(mj.actuator_trntype), np.asarray(mj.actuator_trnid)[:, 0]
            mask = torch.as_tensor(np.stack([(trn == 3) & (tid == o) for o in objid]), dtype=d.actuator_force.dtype)
            v = d.actuator_force @ mask.T
            _write(d, adr, 1, _cutoff(v, cutoff, dtp)[:, :, None])
        # every other type (TOUCH, FRAMELINACC, FRAMEANGACC, ...): `continue` at sensor.py:425-426, slots untouched
    return d
