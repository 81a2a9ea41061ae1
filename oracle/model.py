"""ORACLE (test infrastructure, never the product path): static, per-model precompute.

Restates what the reference derives once at ``device_put`` time:
  * constraint sizes / row layout      reference device.py:212-378
  * collision candidates and groups    reference collision_driver.py:135-174, 299-315, 581-649
  * contact solver-parameter mixing    reference collision_driver.py:177-257
  * static contact ordering            reference collision_driver.py:691-793 (argsort at :764)
Input is any object with mujoco.MjModel attribute names (``MjModelLite`` here).
"""

from __future__ import annotations

import numpy as np
import torch

# enum values (MuJoCo mjmodel.h)
FREE, BALL, SLIDE, HINGE = 0, 1, 2, 3
PLANE, HFIELD, SPHERE, CAPSULE, ELLIPSOID, CYLINDER, BOX, MESH = range(8)
DSBL_CONSTRAINT, DSBL_EQUALITY, DSBL_FRICTIONLOSS, DSBL_LIMIT, DSBL_CONTACT = 1, 2, 4, 8, 16
DSBL_SPRING, DSBL_DAMPER, DSBL_GRAVITY, DSBL_CLAMPCTRL, DSBL_WARMSTART = 32, 64, 128, 256, 512
DSBL_FILTERPARENT, DSBL_ACTUATION, DSBL_REFSAFE, DSBL_SENSOR, DSBL_EULERDAMP = 1024, 2048, 4096, 8192, 1 << 15
mjMINVAL, mjMINMU = 1e-15, 1e-5

# contacts produced per candidate pair (reference collision_primitive.py:224-231, collision_convex.py fn.ncon)
_NCON = {
    (PLANE, SPHERE): 1, (PLANE, CAPSULE): 2, (PLANE, BOX): 4, (PLANE, MESH): 4,
    (SPHERE, SPHERE): 1, (SPHERE, CAPSULE): 1, (SPHERE, BOX): 1, (SPHERE, MESH): 1,
    (CAPSULE, CAPSULE): 1, (CAPSULE, BOX): 2, (CAPSULE, MESH): 2,
    (BOX, BOX): 4, (BOX, MESH): 4, (MESH, MESH): 4,
    (HFIELD, SPHERE): 4, (HFIELD, CAPSULE): 4, (HFIELD, BOX): 4, (HFIELD, MESH): 4,
}


class Candidate:
    __slots__ = ("geom1", "geom2", "ipair", "geomp", "dim")

    def __init__(self, geom1, geom2, ipair, geomp, dim):
        self.geom1, self.geom2, self.ipair, self.geomp, self.dim = geom1, geom2, ipair, geomp, dim


def is_sparse(m) -> bool:
    """reference support.py:43-47."""
    if int(m.opt.jacobian) == 2:
        return m.nv >= 60
    return int(m.opt.jacobian) == 1


def _body_pair_filter(m, b1, b2) -> bool:
    """reference collision_driver.py:299-315."""
    dsbl_filterparent = int(m.opt.disableflags) & DSBL_FILTERPARENT
    weld1, weld2 = m.body_weldid[b1], m.body_weldid[b2]
    parent_weld1 = m.body_weldid[m.body_parentid[weld1]]
    parent_weld2 = m.body_weldid[m.body_parentid[weld2]]
    if weld1 == weld2:
        return True
    if not dsbl_filterparent and weld1 != 0 and weld2 != 0 and (weld1 == parent_weld2 or weld2 == parent_weld1):
        return True
    return False


def collision_candidates(m, convex_shapes=None):
    """reference collision_driver.py:135-174 and :581-615.  Returns an insertion-ordered dict
    key=(t1,t2,meshkey1,meshkey2) -> [Candidate]."""
    result: dict = {}

    def mesh_key(g):
        if convex_shapes is None or convex_shapes[g] is None:
            return ((-1,), (-1,), (-1,))
        return convex_shapes[g]

    def add(g1, g2, ipair=-1):
        g1, g2 = int(g1), int(g2)
        t1, t2 = int(m.geom_type[g1]), int(m.geom_type[g2])
        if t1 > t2:
            t1, t2, g1, g2 = t2, t1, g2, g1
        key = (t1, t2, mesh_key(g1), mesh_key(g2))
        if any((c.geom1, c.geom2) == (g1, g2) for c in result.get(key, [])):
            return
        if ipair > -1:
            cand = Candidate(g1, g2, ipair, -1, int(m.pair_dim[ipair]))
        elif m.geom_priority[g1] != m.geom_priority[g2]:
            gp = g1 if m.geom_priority[g1] > m.geom_priority[g2] else g2
            cand = Candidate(g1, g2, -1, int(gp), int(m.geom_condim[gp]))
        else:
            cand = Candidate(g1, g2, -1, -1, int(max(m.geom_condim[g1], m.geom_condim[g2])))
        result.setdefault(key, []).append(cand)

    for ipair in range(m.npair):
        add(m.pair_geom1[ipair], m.pair_geom2[ipair], ipair)

    exclude = set(int(s) for s in m.exclude_signature)
    body_pairs = []
    for b1 in range(m.nbody):
        for b2 in range(b1, m.nbody):
            if ((b1 << 16) + b2) in exclude:
                continue
            if _body_pair_filter(m, b1, b2):
                continue
            body_pairs.append((b1, b2))
    for b1, b2 in body_pairs:
        for g1 in range(m.body_geomadr[b1], m.body_geomadr[b1] + m.body_geomnum[b1]):
            for g2 in range(m.body_geomadr[b2], m.body_geomadr[b2] + m.body_geomnum[b2]):
                mask = int(m.geom_contype[g1]) & int(m.geom_conaffinity[g2])
                mask |= int(m.geom_contype[g2]) & int(m.geom_conaffinity[g1])
                if mask != 0:
                    add(g1, g2)
    return result


def _max_contact_points(m) -> int:
    """reference collision_driver.py:571-578."""
    for i in range(m.nnumeric):
        name = m.names[m.name_numericadr[i]:].decode("utf-8").split("\x00", 1)[0]
        if name == "max_contact_points":
            return int(m.numeric_data[m.numeric_adr[i]])
    return -1


class OModel:
    """Float tensors in ``dtype`` + integer tables + the static contact/constraint layout."""

    FLOAT_FIELDS = (
        "qpos0 qpos_spring body_pos body_quat body_ipos body_iquat body_mass body_inertia body_gravcomp "
        "body_invweight0 jnt_solref jnt_solimp jnt_pos jnt_axis jnt_stiffness jnt_range jnt_margin jnt_actfrcrange "
        "dof_solref dof_solimp dof_frictionloss dof_armature dof_damping dof_invweight0 geom_solmix geom_solref "
        "geom_solimp geom_size geom_pos geom_quat geom_friction geom_margin geom_gap geom_rbound site_pos site_quat site_size "
        "cam_pos cam_quat cam_pos0 cam_mat0 light_pos light_dir pair_solref pair_solreffriction pair_solimp "
        "pair_margin pair_gap pair_friction eq_solref eq_solimp eq_data tendon_solref_lim tendon_solimp_lim "
        "tendon_solref_fri tendon_solimp_fri tendon_range tendon_margin tendon_stiffness tendon_damping "
        "tendon_armature tendon_frictionloss tendon_lengthspring tendon_invweight0 actuator_dynprm actuator_gainprm "
        "actuator_biasprm actuator_ctrlrange actuator_forcerange actuator_actrange actuator_gear actuator_acc0 "
        "actuator_lengthrange sensor_cutoff"
    ).split()

    def __init__(self, mj, dtype=torch.float64):
        self.mj = mj
        self.dtype = dtype
        for k in ("nq nv nu na nbody njnt ngeom nsite ncam nlight nmesh npair neq ntendon nsensor nmocap nsensordata").split():
            setattr(self, k, int(getattr(mj, k)))
        for k in self.FLOAT_FIELDS:
            v = getattr(mj, k, None)
            if v is not None:
                setattr(self, k, torch.as_tensor(np.asarray(v, dtype=np.float64)).to(dtype))
        o = mj.opt
        self.timestep = float(o.timestep)
        self.gravity = torch.as_tensor(np.asarray(o.gravity, dtype=np.float64)).to(dtype)
        self.wind = torch.as_tensor(np.asarray(o.wind, dtype=np.float64)).to(dtype)
        self.magnetic = torch.as_tensor(np.asarray(o.magnetic, dtype=np.float64)).to(dtype)
        self.density, self.viscosity = float(o.density), float(o.viscosity)
        self.impratio = float(o.impratio)
        self.tolerance, self.ls_tolerance = float(o.tolerance), float(o.ls_tolerance)
        self.iterations, self.ls_iterations = int(o.iterations), int(o.ls_iterations)
        self.solver, self.cone, self.integrator = int(o.solver), int(o.cone), int(o.integrator)
        self.disableflags = int(o.disableflags)
        self.meaninertia = float(mj.stat.meaninertia)
        self.has_fluid = bool(self.density > 0 or self.viscosity > 0 or np.any(np.asarray(o.wind) != 0))
        self.has_gravcomp = bool(np.any(np.asarray(mj.body_gravcomp) != 0))
        self.sparse = is_sparse(mj)
        if self.sparse:
            raise NotImplementedError("sparse-M path (nv >= 60) is a 'next' row (SURVEY.md §8f X4)")

        # dense M mask: (i, j) with j an ancestor dof of i (reference device.py:669-678)
        rows, cols = [], []
        for i in range(self.nv):
            j = i
            while j > -1:
                rows.append(i)
                cols.append(j)
                j = int(mj.dof_parentid[j])
        mask = torch.zeros((self.nv, self.nv), dtype=torch.bool)
        if rows:
            mask[rows, cols] = True
        self.m_mask = mask

        # ancestor-or-self dof mask per body (what support.jac derives with a reverse scan, :138-144)
        anc = np.zeros((self.nbody, self.nv), dtype=bool)
        for b in range(1, self.nbody):
            anc[b] = anc[mj.body_parentid[b]]
            if mj.body_dofnum[b] > 0:
                anc[b, mj.body_dofadr[b]: mj.body_dofadr[b] + mj.body_dofnum[b]] = True
        self.body_dofmask = torch.as_tensor(anc)

        self._convex()
        self._collision()
        self._constraint_sizes()

    # -- convex geoms (filled by oracle.convex when present)
    def _convex(self):
        from . import mesh as omesh

        self.convex = omesh.build_convex(self.mj, self.dtype)

    # -- collision statics
    def _collision(self):
        mj = self.mj
        shapes = [None if c is None else c.shape_key for c in self.convex]
        if self.disableflags & (DSBL_CONSTRAINT | DSBL_CONTACT):
            self.groups, self.ncon, self.contact = [], 0, None
            return
        cands = collision_candidates(mj, shapes)
        groups = []
        for key, lst in cands.items():
            t = (key[0], key[1])
            if t not in _NCON:
                if t[0] == PLANE and t[1] in (PLANE, HFIELD):
                    continue
                raise NotImplementedError(f"geom pair {t} collisions not implemented")
            # reference precompute_collision_indices :440-497 regroups by (pair, priority) type
            typ: dict = {}
            for c in lst:
                typ.setdefault((c.ipair > -1, c.geomp > -1), []).append(c)
            ordered = [c for sub in typ.values() for c in sub]
            groups.append((t, ordered, _NCON[t]))
        self.groups = groups
        if not groups:
            self.ncon, self.contact = 0, None
            return
        dt = self.dtype
        g1l, g2l, diml, im, fr, sr, srf, si = [], [], [], [], [], [], [], []
        for t, lst, n in groups:
            for c in lst:
                p = self._params(c)
                for _ in range(n):
                    g1l.append(c.geom1)
                    g2l.append(c.geom2)
                    diml.append(c.dim)
                    im.append(p["margin"] - p["gap"])
                    fr.append(p["friction"])
                    sr.append(p["solref"])
                    srf.append(p["solreffriction"])
                    si.append(p["solimp"])
        contact_dim = torch.tensor(diml, dtype=torch.int32)
        total = len(diml)
        max_cp = _max_contact_points(mj)
        self.topk = 0
        if max_cp > -1 and total > max_cp:
            # collision_driver.collision :822-840: the K closest candidates survive, per env; the contact table below then keeps
            # one entry per CANDIDATE (group order) and oracle.collision records which candidate sits in which slot
            if len(set(diml)) > 1:
                raise NotImplementedError("max_contact_points with contacts of different condim is not supported")
            self.topk, self.ncand = max_cp, total
            self.topk_perm = torch.argsort(contact_dim[:max_cp])
        # The reference sorts with an UNSTABLE argsort (collision_driver.py:764,842): same call, same torch build.
        sort_idx = torch.argsort(contact_dim) if not self.topk else torch.arange(total)
        self.contact_perm = sort_idx
        f = lambda l: torch.stack([torch.as_tensor(np.asarray(x, dtype=np.float64)) for x in l]).to(dt)[sort_idx]
        self.contact = dict(
            geom1=torch.tensor(g1l)[sort_idx], geom2=torch.tensor(g2l)[sort_idx], dim=contact_dim[sort_idx],
            includemargin=f(im), friction=f(fr), solref=f(sr), solreffriction=f(srf), solimp=f(si),
        )
        self.ncon = self.topk or total

    def _params(self, c):
        """reference collision_driver.py:177-257 for one candidate (float64 host arithmetic)."""
        mj = self.mj
        if c.ipair > -1:
            i = c.ipair
            return dict(friction=np.maximum(mj.pair_friction[i], mjMINMU), solref=mj.pair_solref[i],
                        solreffriction=mj.pair_solreffriction[i], solimp=mj.pair_solimp[i],
                        margin=mj.pair_margin[i], gap=mj.pair_gap[i])
        g1, g2 = c.geom1, c.geom2
        if c.geomp > -1:
            gp = c.geomp
            return dict(friction=mj.geom_friction[gp][[0, 0, 1, 2, 2]], solref=mj.geom_solref[gp], solreffriction=np.zeros(2),
                        solimp=mj.geom_solimp[gp], margin=max(mj.geom_margin[g1], mj.geom_margin[g2]),
                        gap=max(mj.geom_gap[g1], mj.geom_gap[g2]))
        friction = np.maximum(mj.geom_friction[g1], mj.geom_friction[g2])[[0, 0, 1, 2, 2]]
        s1, s2 = mj.geom_solmix[g1], mj.geom_solmix[g2]
        with np.errstate(divide="ignore", invalid="ignore"):
            mix = s1 / (s1 + s2)
        if s1 < mjMINVAL and s2 < mjMINVAL:
            mix = 0.5
        if s1 < mjMINVAL and s2 >= mjMINVAL:
            mix = 0.0
        r1, r2 = mj.geom_solref[g1], mj.geom_solref[g2]
        solref = np.minimum(r1, r2)
        if r1[0] > 0 and r2[0] > 0:
            solref = mix * r1 + (1 - mix) * r2
        solimp = mix * mj.geom_solimp[g1] + (1 - mix) * mj.geom_solimp[g2]
        return dict(friction=friction, solref=solref, solreffriction=np.zeros(2), solimp=solimp,
                    margin=max(mj.geom_margin[g1], mj.geom_margin[g2]), gap=max(mj.geom_gap[g1], mj.geom_gap[g2]))

    # -- constraint sizes (reference device.py:226-264)
    def _constraint_sizes(self):
        mj, fl = self.mj, self.disableflags
        if fl & DSBL_CONSTRAINT:
            self.ne = self.nf = self.nl = self.nefc = 0
            self.ncon = 0
            self.ne_connect = self.ne_weld = self.ne_joint = 0
            self.condim_counts = (0, 0, 0, 0)
            return
        if fl & DSBL_EQUALITY:
            self.ne_connect = self.ne_weld = self.ne_joint = 0
        else:
            et = np.asarray(mj.eq_type)
            self.ne_connect, self.ne_weld, self.ne_joint = int((et == 0).sum()), int((et == 1).sum()), int((et == 2).sum())
        self.ne = self.ne_connect * 3 + self.ne_weld * 6 + self.ne_joint
        self.nf = 0 if fl & DSBL_FRICTIONLOSS else int((np.asarray(mj.dof_frictionloss) > 0).sum()) + int(
            (np.asarray(mj.tendon_frictionloss) > 0).sum())
        self.nl = 0 if fl & DSBL_LIMIT else int(np.asarray(mj.jnt_limited).sum()) + int(np.asarray(mj.tendon_limited).sum())
        nc = 0
        counts = [0, 0, 0, 0]
        if self.ncon:
            dims = self.contact["dim"][: self.ncon]   # max_contact_points: the K smallest condims (make_condim :618-644; all equal here)
            for k, condim in enumerate((1, 3, 4, 6)):
                n = int((dims == condim).sum())
                counts[k] = n
                rows = condim if self.cone == 1 else (condim - 1) * 2
                nc += n * (1 if condim == 1 else rows)
        self.condim_counts = tuple(counts)
        self.nefc = self.ne + self.nf + self.nl + nc
