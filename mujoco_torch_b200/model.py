"""``Model``: the immutable, device-resident compiled model and all static (per-model) precompute.

Mirrors the role of the reference's ``device_put(MjModel)`` (reference device.py:1029-1112 and
``_model_derived`` :632-884): candidate geom pairs and their final contact order, mixed contact solver
parameters, constraint row layout, ancestor-dof masks -- computed once on the host and packed into one
contiguous ModelBlob that every CTA stages into shared memory with a single TMA bulk copy.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from . import convex as _convex
from . import layout as L
from .mjmodel import (ConeType, DisableBit, EqType, GeomType, IntegratorType, JacobianType, JointType, MjModelLite, SensorType,
                      SolverType, TrnType, WrapType, mjMINMU, mjMINVAL)

_G = GeomType
_FN_TABLE = {
    (_G.PLANE, _G.SPHERE): "PLANE_SPHERE", (_G.PLANE, _G.CAPSULE): "PLANE_CAPSULE", (_G.PLANE, _G.BOX): "PLANE_CONVEX",
    (_G.PLANE, _G.MESH): "PLANE_CONVEX", (_G.SPHERE, _G.SPHERE): "SPHERE_SPHERE", (_G.SPHERE, _G.CAPSULE): "SPHERE_CAPSULE",
    (_G.SPHERE, _G.BOX): "SPHERE_CONVEX", (_G.SPHERE, _G.MESH): "SPHERE_CONVEX", (_G.CAPSULE, _G.CAPSULE): "CAPSULE_CAPSULE",
    (_G.CAPSULE, _G.BOX): "CAPSULE_CONVEX", (_G.CAPSULE, _G.MESH): "CAPSULE_CONVEX", (_G.BOX, _G.BOX): "CONVEX_CONVEX",
    (_G.BOX, _G.MESH): "CONVEX_CONVEX", (_G.MESH, _G.MESH): "CONVEX_CONVEX",
    # height fields (reference collision_driver.py:111-114, collision_hfield.py)
    (_G.HFIELD, _G.SPHERE): "HFIELD_SPHERE", (_G.HFIELD, _G.CAPSULE): "HFIELD_CAPSULE", (_G.HFIELD, _G.BOX): "HFIELD_CONVEX",
    (_G.HFIELD, _G.MESH): "HFIELD_CONVEX",
}
HFIELD_MAX_PRISMS = 32   # sub-grid prisms one (pair, env) thread walks; larger objects / finer grids are refused at device_put
# Solver live tile: room for LIVE_FRAC of the constraint rows, at least LIVE_MIN_ROWS.  Measured on the bench states the live
# share is far below that (box stack: <= 112 of 336 rows for 99 % of the envs, Ant: <= 36 of 284); envs that need more take
# the full-tile pass.  Models solved by half warps (nv <= 16) are occupancy-limited by the row vectors rather than by M / L, so
# they get the tighter *_HALFWARP pair (profiles/README.md has the sweep).
LIVE_FRAC, LIVE_MIN_ROWS = 0.33, 64
LIVE_FRAC_HALFWARP, LIVE_MIN_ROWS_HALFWARP = 0.15, 48

# Sensor types the reference has no branch for (TOUCH, FRAMELINACC, FRAMEANGACC, ... ) are not an error there: its stage loops
# `continue` past them and their sensordata slots keep the caller's values (sensor.py:425-426).  k_sensor does the same, so no
# sensor type is refused here.


class Option:
    """Physics options (reference types.Option :503-557)."""

    def __init__(self, o):
        for k in ("timestep", "impratio", "tolerance", "ls_tolerance", "density", "viscosity"):
            setattr(self, k, float(getattr(o, k)))
        for k in ("gravity", "wind", "magnetic"):
            setattr(self, k, np.asarray(getattr(o, k), dtype=np.float64).copy())
        self.iterations, self.ls_iterations = int(o.iterations), int(o.ls_iterations)
        self.integrator = IntegratorType(int(o.integrator))
        self.cone = ConeType(int(o.cone))
        self.jacobian = JacobianType(int(o.jacobian))
        self.solver = SolverType(int(o.solver))
        self.disableflags = DisableBit(int(o.disableflags))
        self.enableflags = int(getattr(o, "enableflags", 0))
        self.has_fluid_params = bool(self.density > 0 or self.viscosity > 0 or np.any(self.wind != 0))


class Statistic:
    def __init__(self, s):
        self.meaninertia = float(s.meaninertia)


def is_sparse(m) -> bool:
    """reference support.is_sparse :43-47."""
    if int(m.opt.jacobian) == JacobianType.AUTO:
        return m.nv >= 60
    return int(m.opt.jacobian) == JacobianType.SPARSE


def _validate(mj):
    """Same refusals as reference device._validate :919-964 for the subset that exists here."""
    if int(mj.opt.integrator) not in (IntegratorType.EULER, IntegratorType.RK4, IntegratorType.IMPLICITFAST):
        raise NotImplementedError(f"{IntegratorType(int(mj.opt.integrator))!r} not implemented.")
    if int(mj.opt.solver) not in (SolverType.CG, SolverType.NEWTON):
        raise NotImplementedError(f"{SolverType(int(mj.opt.solver))!r} not implemented.")
    if any(int(d) not in (1, 3, 4, 6) for d in mj.geom_condim) or any(int(d) not in (1, 3, 4, 6) for d in mj.pair_dim):
        raise NotImplementedError("Only condim=1, 3, 4 and 6 are supported.")
    for t in np.asarray(mj.eq_type):
        if int(t) not in (EqType.CONNECT, EqType.WELD, EqType.JOINT):
            raise NotImplementedError(f"{EqType(int(t))!r} not implemented.")
    # Site-based connect / weld (eq_objtype = mjOBJ_SITE = 6): the reference never reads eq_objtype and indexes bodies with the
    # object ids (device.py:306-307).  Kept for parity; ids that are not body ids would index out of range there.
    for i, t in enumerate(np.asarray(mj.eq_type)):
        if int(t) in (EqType.CONNECT, EqType.WELD) and max(int(mj.eq_obj1id[i]), int(mj.eq_obj2id[i])) >= int(mj.nbody):
            raise NotImplementedError(f"equality {i}: object id beyond nbody (site-based equalities are indexed as bodies, like the reference)")
    for t in np.asarray(mj.actuator_trntype):
        if int(t) not in (TrnType.JOINT, TrnType.JOINTINPARENT, TrnType.TENDON):
            raise NotImplementedError(f"{TrnType(int(t))!r} not implemented.")
    if is_sparse(mj):
        raise NotImplementedError("sparse mass-matrix path (nv >= 60 or jacobian=sparse) is not implemented")


# ------------------------------------------------------------------------------------------ broadphase


def _weld_filtered(mj, b1, b2):
    w1, w2 = int(mj.body_weldid[b1]), int(mj.body_weldid[b2])
    if w1 == w2:
        return True
    if int(mj.opt.disableflags) & DisableBit.FILTERPARENT:
        return False
    pw1 = int(mj.body_weldid[mj.body_parentid[w1]])
    pw2 = int(mj.body_weldid[mj.body_parentid[w2]])
    return w1 != 0 and w2 != 0 and (w1 == pw2 or w2 == pw1)


def candidate_pairs(mj, convex_tables):
    """Static candidate geom pairs in the reference's group order (collision_driver.py:135-174,:581-615,
    regrouped by parameter source as in precompute_collision_indices :440-497).

    Returns a list of dicts {fn, g1, g2, ipair, geomp, dim} and the per-pair contact counts.
    """
    groups: dict = {}

    def shape_key(g):
        c = convex_tables[g]
        return ((-1,),) * 3 if c is None else c.key

    def push(g1, g2, ipair=-1):
        g1, g2 = int(g1), int(g2)
        t1, t2 = int(mj.geom_type[g1]), int(mj.geom_type[g2])
        if t1 > t2:
            g1, g2, t1, t2 = g2, g1, t2, t1
        bucket = groups.setdefault((t1, t2, shape_key(g1), shape_key(g2)), [])
        if any(p["g1"] == g1 and p["g2"] == g2 for p in bucket):
            return
        if ipair >= 0:
            bucket.append(dict(g1=g1, g2=g2, ipair=ipair, geomp=-1, dim=int(mj.pair_dim[ipair])))
            return
        p1, p2 = int(mj.geom_priority[g1]), int(mj.geom_priority[g2])
        if p1 != p2:
            gp = g1 if p1 > p2 else g2
            bucket.append(dict(g1=g1, g2=g2, ipair=-1, geomp=gp, dim=int(mj.geom_condim[gp])))
        else:
            bucket.append(dict(g1=g1, g2=g2, ipair=-1, geomp=-1, dim=int(max(mj.geom_condim[g1], mj.geom_condim[g2]))))

    for ip in range(mj.npair):
        push(mj.pair_geom1[ip], mj.pair_geom2[ip], ip)
    excluded = {int(s) for s in np.asarray(mj.exclude_signature)}
    for b1 in range(mj.nbody):
        for b2 in range(b1, mj.nbody):
            if ((b1 << 16) + b2) in excluded or _weld_filtered(mj, b1, b2):
                continue
            a1, n1 = int(mj.body_geomadr[b1]), int(mj.body_geomnum[b1])
            a2, n2 = int(mj.body_geomadr[b2]), int(mj.body_geomnum[b2])
            for g1 in range(a1, a1 + n1):
                for g2 in range(a2, a2 + n2):
                    if (int(mj.geom_contype[g1]) & int(mj.geom_conaffinity[g2])) | (
                            int(mj.geom_contype[g2]) & int(mj.geom_conaffinity[g1])):
                        push(g1, g2)

    flat = []
    for (t1, t2, _, _), bucket in groups.items():
        tt = (GeomType(t1), GeomType(t2))
        if tt not in _FN_TABLE:
            if t1 == _G.PLANE and t2 in (_G.PLANE, _G.HFIELD):
                continue  # MuJoCo never collides these
            raise NotImplementedError(f"({tt[0]!r}, {tt[1]!r}) collisions not implemented.")
        by_source: dict = {}
        for p in bucket:
            by_source.setdefault((p["ipair"] >= 0, p["geomp"] >= 0), []).append(p)
        for sub in by_source.values():
            for p in sub:
                p["fn"] = _FN_TABLE[tt]
                flat.append(p)
    return flat


def _mix_params(mj, p):
    """Solver parameters of one candidate pair (reference collision_driver.py:177-257)."""
    if p["ipair"] >= 0:
        i = p["ipair"]
        return (np.maximum(mj.pair_friction[i], mjMINMU), mj.pair_solref[i], mj.pair_solreffriction[i], mj.pair_solimp[i],
                float(mj.pair_margin[i]) - float(mj.pair_gap[i]))
    g1, g2 = p["g1"], p["g2"]
    include = max(mj.geom_margin[g1], mj.geom_margin[g2]) - max(mj.geom_gap[g1], mj.geom_gap[g2])
    if p["geomp"] >= 0:
        gp = p["geomp"]
        f = mj.geom_friction[gp]
        return np.array([f[0], f[0], f[1], f[2], f[2]]), mj.geom_solref[gp], np.zeros(2), mj.geom_solimp[gp], include
    f = np.maximum(mj.geom_friction[g1], mj.geom_friction[g2])
    s1, s2 = float(mj.geom_solmix[g1]), float(mj.geom_solmix[g2])
    if s1 < mjMINVAL and s2 < mjMINVAL:
        w = 0.5
    elif s1 < mjMINVAL:
        w = 0.0
    else:
        w = s1 / (s1 + s2)
    r1, r2 = mj.geom_solref[g1], mj.geom_solref[g2]
    solref = w * r1 + (1 - w) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2)
    solimp = w * mj.geom_solimp[g1] + (1 - w) * mj.geom_solimp[g2]
    return np.array([f[0], f[0], f[1], f[2], f[2]]), solref, np.zeros(2), solimp, include


# ------------------------------------------------------------------------------------------ Model


class Model:
    """Compiled model + packed device blob.  Public attributes follow reference types.Model :560-910."""

    def __init__(self, mj, dtype=torch.float64):
        _validate(mj)
        self.mj = mj
        self.dtype = dtype
        self.device = torch.device("cpu")
        for k in ("nq nv nu na nbody njnt ngeom nsite ncam nlight nmesh npair nexclude neq ntendon nwrap nsensor nnumeric "
                  "nmocap nM nsensordata").split():
            setattr(self, k, int(getattr(mj, k, 0)))
        self.opt = Option(mj.opt)
        self.stat = Statistic(mj.stat)
        self._static()
        self._pack()
        self._handle = None
        self._handles = {}   # layout name -> device handle (ops.model_handle)
        self._mid = None

    # ---- tensors mirroring MjModel fields on demand (qpos0, body_mass, ...)
    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        mj = self.__dict__.get("mj")
        if mj is not None and hasattr(mj, name):
            v = getattr(mj, name)
            if isinstance(v, np.ndarray):
                t = torch.as_tensor(v)
                if t.is_floating_point():
                    t = t.to(self.dtype)
                return t.to(self.device)
            return v
        raise AttributeError(name)

    @property
    def _id(self):
        """Registry id used by the torch.library op (assigned lazily, one per Model object)."""
        if self._mid is None:
            from . import ops

            self._mid = ops.register_model(self)
        return self._mid

    @property
    def constraint_sizes_py(self):
        return (self.ne, self.nf, self.nl, self.ncon, self.nefc)

    # ---- static precompute
    def _static(self):
        mj = self.mj
        fl = int(mj.opt.disableflags)
        self.convex_tables = _convex.build(mj)
        contacts_on = not (fl & (DisableBit.CONSTRAINT | DisableBit.CONTACT))
        pairs = candidate_pairs(mj, self.convex_tables) if contacts_on else []
        # group-concat contact list, then the reference's (unstable) argsort by condim
        dims, owner = [], []
        for ip, p in enumerate(pairs):
            n = L.PAIR_NCON[p["fn"]]
            dims += [p["dim"]] * n
            owner += [(ip, k) for k in range(n)]
        ncon = len(dims)
        # max_contact_points (reference collision_driver.py:571-578, 822-840): of the `ncand` candidate contacts only the K with the
        # smallest distance survive, chosen per env at run time.  The candidates then keep their group order (k_collision writes
        # them to scratch slots), k_topk selects, and contact parameters / bodies are looked up through the selected candidate.
        self.topk, self.ncand = 0, 0
        for i in range(getattr(mj, "nnumeric", 0)):
            nm = mj.names[int(mj.name_numericadr[i]):].split(b"\x00", 1)[0].decode()
            if nm == "max_contact_points" and 0 <= int(mj.numeric_data[int(mj.numeric_adr[i])]) < ncon:
                self.topk, self.ncand = int(mj.numeric_data[int(mj.numeric_adr[i])]), ncon
        if self.topk and len(set(dims)) > 1:
            # the reference sizes the efc rows from the K smallest condims of the model (make_condim :618-644) while the contacts that
            # survive are picked by distance: with mixed condims its row table and its contacts disagree -- nothing to be faithful to
            raise NotImplementedError("max_contact_points with contacts of different condim is not supported")
        # The reference orders contacts with torch.argsort(contact_dim) (collision_driver.py:764,842), which is not a
        # stable sort for n > 16; we make the same call once here and bake the permutation.
        order = torch.argsort(torch.tensor(dims, dtype=torch.int32)).tolist() if ncon else []
        if self.topk:
            order = list(range(ncon))   # candidates stay in group order; the sort runs over the K selected (topk_perm)
            self.topk_perm = torch.argsort(torch.tensor(dims[: self.topk], dtype=torch.int32)).tolist()
        slot_of = {owner[src]: dst for dst, src in enumerate(order)}
        self.pairs, self.ncon = pairs, ncon
        self.contact_order = order
        self.pair_slot = np.full((max(len(pairs), 1), 4), -1, dtype=np.int32)
        for (ip, k), dst in slot_of.items():
            self.pair_slot[ip, k] = dst
        con = dict(dim=np.zeros(ncon, np.int32), g1=np.zeros(ncon, np.int32), g2=np.zeros(ncon, np.int32),
                   include=np.zeros(ncon), friction=np.zeros((ncon, 5)), solref=np.zeros((ncon, 2)),
                   solreffriction=np.zeros((ncon, 2)), solimp=np.zeros((ncon, 5)))
        for ip, p in enumerate(pairs):
            fr, sr, srf, si, inc = _mix_params(mj, p)
            for k in range(L.PAIR_NCON[p["fn"]]):
                c = self.pair_slot[ip, k]
                con["dim"][c], con["g1"][c], con["g2"][c] = p["dim"], p["g1"], p["g2"]
                con["include"][c], con["friction"][c], con["solref"][c] = inc, fr, sr
                con["solreffriction"][c], con["solimp"][c] = srf, si
        self.con = con
        if self.topk:
            self.ncon = ncon = self.topk   # contact slots; self.con keeps one entry per CANDIDATE

        # constraint row sources, in the reference's row order (constraint.py:651-674)
        src = []  # (type, id, nrows)
        S = {n: i for i, n in enumerate(L.SRC_TYPES)}
        on = not (fl & DisableBit.CONSTRAINT)
        self.ne = self.nf = self.nl = 0
        if on and not (fl & DisableBit.EQUALITY):
            et = np.asarray(mj.eq_type)
            for tname, tval, rows in (("EQ_CONNECT", EqType.CONNECT, 3), ("EQ_WELD", EqType.WELD, 6), ("EQ_JOINT", EqType.JOINT, 1)):
                for i in np.nonzero(et == tval)[0]:
                    src.append((S[tname], int(i), rows))
                    self.ne += rows
        if on and not (fl & DisableBit.FRICTIONLOSS):
            for i in np.nonzero(np.asarray(mj.dof_frictionloss) > 0)[0]:
                src.append((S["FRIC_DOF"], int(i), 1))
                self.nf += 1
            for i in np.nonzero(np.asarray(mj.tendon_frictionloss) > 0)[0]:
                src.append((S["FRIC_TEN"], int(i), 1))
                self.nf += 1
        if on and not (fl & DisableBit.LIMIT):
            jt, lim = np.asarray(mj.jnt_type), np.asarray(mj.jnt_limited).astype(bool)
            for i in np.nonzero((jt == JointType.BALL) & lim)[0]:
                src.append((S["LIM_BALL"], int(i), 1))
            for i in np.nonzero(((jt == JointType.SLIDE) | (jt == JointType.HINGE)) & lim)[0]:
                src.append((S["LIM_SH"], int(i), 1))
            # the reference sizes nl from jnt_limited.sum() (device.py:247): free joints flagged limited would count too
            for i in np.nonzero(np.asarray(mj.tendon_limited))[0]:
                src.append((S["LIM_TEN"], int(i), 1))
            self.nl = sum(1 for s in src if s[0] in (S["LIM_BALL"], S["LIM_SH"], S["LIM_TEN"]))
        elliptic = int(mj.opt.cone) == ConeType.ELLIPTIC
        if on:
            for c in range(ncon):
                d = int(con["dim"][c])
                rows = 1 if d == 1 else (d if elliptic else 2 * (d - 1))
                src.append((S["CONTACT"], c, rows))
        row, self.src = 0, []
        for t, i, n in src:
            self.src.append((t, i, row))
            row += n
        self.nefc = row
        self.con_efcadr = np.array([r for (t, i, r) in self.src if t == S["CONTACT"]], dtype=np.int32)

        # ancestor-or-self dof masks per body (replaces the reference's reverse tree scan in support.jac :138-144)
        mw = max(1, (self.nv + 31) // 32)
        mask = np.zeros((self.nbody, mw), dtype=np.uint32)
        for b in range(1, self.nbody):
            mask[b] = mask[int(mj.body_parentid[b])]
            for dof in range(int(mj.body_dofadr[b]), int(mj.body_dofadr[b]) + int(mj.body_dofnum[b])):
                mask[b, dof // 32] |= np.uint32(1 << (dof % 32))
        self.body_dofmask = mask
        self.maskwords = mw
        self._row_supports(S)

    def _row_supports(self, S):
        """Static column support of every constraint source (the dofs its Jacobian rows can touch).

        The solver kernel keeps J compacted to these columns in shared memory and assembles the Newton Hessian
        only over them.  Contacts: dofs that are ancestors of exactly one of the two bodies (common ancestors
        cancel exactly, reference support.jac_dif_pair :156-166); limits / friction: the joint's own dofs;
        tendons: the wrapped joints' dofs; equalities: ancestors of either body (or the two joint dofs).
        """
        mj = self.mj
        nv = self.nv

        def anc(b):
            bits = self.body_dofmask[b]
            return {i for i in range(nv) if (int(bits[i // 32]) >> (i % 32)) & 1}

        def tendon_dofs(t):
            out = set()
            for w in range(int(mj.tendon_adr[t]), int(mj.tendon_adr[t]) + int(mj.tendon_num[t])):
                if int(mj.wrap_type[w]) == WrapType.JOINT:
                    out.add(int(mj.jnt_dofadr[int(mj.wrap_objid[w])]))
            return out

        gb = np.asarray(mj.geom_bodyid)
        src_cadr, src_nsup, src_nrows, cols = [], [], [], []
        nrows_of = {}
        for k, (t, i, row) in enumerate(self.src):
            nxt = self.src[k + 1][2] if k + 1 < len(self.src) else self.nefc
            nrows_of[k] = nxt - row
        for k, (t, i, row) in enumerate(self.src):
            name = L.SRC_TYPES[t]
            if name == "CONTACT" and self.topk:
                sup = set()   # any candidate may land in this slot: the union of their supports
                for cnd in range(self.ncand):
                    sup |= anc(int(gb[self.con["g1"][cnd]])) ^ anc(int(gb[self.con["g2"][cnd]]))
            elif name == "CONTACT":
                a, b = anc(int(gb[self.con["g1"][i]])), anc(int(gb[self.con["g2"][i]]))
                sup = a ^ b
            elif name in ("EQ_CONNECT", "EQ_WELD"):
                sup = anc(int(mj.eq_obj1id[i])) | anc(int(mj.eq_obj2id[i]))
            elif name == "EQ_JOINT":
                j2 = int(mj.eq_obj2id[i])
                sup = {int(mj.jnt_dofadr[int(mj.eq_obj1id[i])]), int(mj.jnt_dofadr[j2 if j2 >= 0 else self.njnt + j2])}
            elif name == "FRIC_DOF":
                sup = {i}
            elif name in ("FRIC_TEN", "LIM_TEN"):
                sup = tendon_dofs(i)
            elif name == "LIM_BALL":
                d0 = int(mj.jnt_dofadr[i])
                sup = {d0, d0 + 1, d0 + 2}
            else:  # LIM_SH
                sup = {int(mj.jnt_dofadr[i])}
            sup = sorted(sup)
            src_cadr.append(len(cols))
            src_nsup.append(len(sup))
            src_nrows.append(nrows_of[k])
            cols.extend(sup)
        row_src, row_jadr, jadr = [], [], 0
        for k, (t, i, row) in enumerate(self.src):
            for _ in range(src_nrows[k]):
                row_src.append(k)
                row_jadr.append(jadr)
                jadr += src_nsup[k]
        # transpose of the support table: for every dof the (source, position inside that source's support) pairs, so
        # that J^T f is a per-dof gather (one lane per dof, no scatter)
        per_dof = [[] for _ in range(self.nv)]
        for k in range(len(self.src)):
            for pos in range(src_nsup[k]):
                per_dof[cols[src_cadr[k] + pos]].append(k | (pos << 16))
        dof_sadr = np.concatenate([[0], np.cumsum([len(x) for x in per_dof])]).astype(np.int64)
        dof_sent = [e for x in per_dof for e in x]
        self.sup = dict(src_cadr=src_cadr, src_nsup=src_nsup, src_nrows=src_nrows, cols=cols, row_src=row_src, row_jadr=row_jadr,
                        jc_words=jadr, dof_sadr=dof_sadr, dof_sent=dof_sent)
        # Live tile of the solver (models with many, mostly inactive rows): room for LIVE_FRAC of the rows and of the compact
        # J, never less than 64 rows.  Envs that need more fall back to the full tile (csrc/k_solve.cuh).
        nefc = len(row_src)
        if nefc > 64:
            half = self.nv <= 16
            min_rows = int(os.environ.get("MJT_LIVE_MIN_ROWS", LIVE_MIN_ROWS_HALFWARP if half else LIVE_MIN_ROWS))   # tuning knobs (A/B runs)
            frac = float(os.environ.get("MJT_LIVE_FRAC", LIVE_FRAC_HALFWARP if half else LIVE_FRAC))
            rcap = min(nefc, max(min_rows, -(-int(nefc * frac) // 16) * 16))
            maxsup = max(src_nsup) if src_nsup else 1
            jcap = min(jadr, max(rcap * maxsup * 3 // 4, min_rows))
            self.live_caps = (rcap, jcap) if rcap * 2 <= nefc else (0, 0)   # a tile that does not halve the rows costs more than it saves (HalfCheetah, measured)
        else:
            self.live_caps = (0, 0)

    # ---- blob
    def _pack(self):
        mj = self.mj
        f64 = self.dtype == torch.float64
        I, F = {}, {}
        i32 = lambda x: np.ascontiguousarray(np.asarray(x), dtype=np.int64).astype(np.int32).reshape(-1)
        f = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1)
        for k in ("body_parentid body_rootid body_jntadr body_jntnum body_dofadr body_dofnum body_mocapid jnt_type jnt_qposadr "
                  "jnt_dofadr jnt_bodyid jnt_limited dof_bodyid dof_parentid dof_jntid geom_bodyid geom_type site_bodyid "
                  "jnt_actfrclimited jnt_actgravcomp").split():
            I[k] = i32(getattr(mj, k))
        I["body_dofmask"] = self.body_dofmask.view(np.int32).reshape(-1)
        # geoms / sites grouped by body (k_smooth places them while the body's pose is still in registers)
        for kind, ids in (("geom", np.asarray(mj.geom_bodyid).reshape(-1)), ("site", np.asarray(mj.site_bodyid).reshape(-1))):
            order = np.argsort(ids, kind="stable")
            num = np.bincount(ids.astype(np.int64), minlength=self.nbody) if ids.size else np.zeros(self.nbody, np.int64)
            I[f"body_{kind}adr"] = i32(np.concatenate([[0], np.cumsum(num)[:-1]]))
            I[f"body_{kind}num"] = i32(num)
            I[f"body_{kind}"] = i32(order)
        for k in ("trntype dyntype gaintype biastype actadr actnum ctrllimited forcelimited actlimited").split():
            I["act_" + k] = i32(getattr(mj, "actuator_" + k))
        I["act_trnid"] = i32(np.asarray(mj.actuator_trnid).reshape(-1, 2)[:, 0]) if self.nu else i32([])
        I["ten_adr"], I["ten_num"], I["ten_limited"] = i32(mj.tendon_adr), i32(mj.tendon_num), i32(mj.tendon_limited)
        wj = [int(o) if int(t) == WrapType.JOINT else -1 for t, o in zip(mj.wrap_type, mj.wrap_objid)]
        I["wrap_jnt"] = i32(wj)
        fn_id = {n: i for i, n in enumerate(L.PAIR_FNS)}
        I["pair_fn"] = i32([fn_id[p["fn"]] for p in self.pairs])
        I["pair_g1"], I["pair_g2"] = i32([p["g1"] for p in self.pairs]), i32([p["g2"] for p in self.pairs])
        I["pair_slot"] = i32(self.pair_slot[: len(self.pairs)])
        # height-field pairs (reference collision_driver._hfield_subgrid_size :318-329): the sub-grid of cells around the object,
        # from its bounding radius and the grid pitch -- static; which cells it covers is decided per env in the kernel
        nhf = int(getattr(mj, "nhfield", 0))
        I["hf_nrow"], I["hf_ncol"], I["hf_adr"] = i32(getattr(mj, "hfield_nrow", [])), i32(getattr(mj, "hfield_ncol", [])), i32(getattr(mj, "hfield_adr", []))
        pair_hf, pair_rb = [], []
        for p_ in self.pairs:
            if not p_["fn"].startswith("HFIELD"):
                pair_hf += [-1, 0, 0]
                pair_rb.append(0.0)
                continue
            hid = int(mj.geom_dataid[p_["g1"]])
            hs, nrow, ncol = np.asarray(mj.hfield_size[hid], dtype=np.float64), int(mj.hfield_nrow[hid]), int(mj.hfield_ncol[hid])
            rb = float(mj.geom_rbound[p_["g2"]])
            xtick, ytick = 2.0 * hs[0] / (ncol - 1), 2.0 * hs[1] / (nrow - 1)
            xb, yb = min(int(np.ceil(2 * rb / xtick)) + 1, ncol), min(int(np.ceil(2 * rb / ytick)) + 1, nrow)
            if 2 * xb * yb > HFIELD_MAX_PRISMS:
                raise NotImplementedError(f"hfield pair ({p_['g1']}, {p_['g2']}): sub-grid {xb} x {yb} needs {2 * xb * yb} prisms, more than {HFIELD_MAX_PRISMS}")
            pair_hf += [hid, xb, yb]
            # the radius that places the sub-grid around the object is NOT geom_rbound: hfield_sphere / hfield_convex use
            # max(geom_size) (collision_hfield.py:355,381), hfield_capsule radius + half length (:368)
            gs = np.asarray(mj.geom_size[p_["g2"]], dtype=np.float64)
            pair_rb.append(float(gs[0] + gs[1]) if p_["fn"] == "HFIELD_CAPSULE" else float(gs.max()))
        I["pair_hf"] = i32(pair_hf)
        gb = np.asarray(mj.geom_bodyid)
        I["con_body1"], I["con_body2"] = i32(gb[self.con["g1"]]) if len(self.con["g1"]) else i32([]), i32(gb[self.con["g2"]]) if len(self.con["g2"]) else i32([])
        I["con_dim"], I["con_efcadr"] = i32(self.con["dim"]), i32(self.con_efcadr)
        I["topk_perm"] = i32(getattr(self, "topk_perm", []))
        I["con_geom1"], I["con_geom2"] = i32(self.con["g1"]), i32(self.con["g2"])
        I["src_type"], I["src_id"], I["src_row"] = (i32([s[k] for s in self.src]) for k in range(3))
        sp = self.sup
        I["src_cadr"], I["src_nsup"], I["src_nrows"] = i32(sp["src_cadr"]), i32(sp["src_nsup"]), i32(sp["src_nrows"])
        I["sup_col"], I["row_src"], I["row_jadr"] = i32(sp["cols"]), i32(sp["row_src"]), i32(sp["row_jadr"])
        jc_src = []
        for r, k in enumerate(sp["row_src"]):
            ca, ns = sp["src_cadr"][k], sp["src_nsup"][k]
            jc_src += [r * self.nv + cc for cc in sp["cols"][ca: ca + ns]]
        I["jc_src"] = i32(jc_src)
        I["dof_sadr"], I["dof_sent"] = i32(sp["dof_sadr"]), i32(sp["dof_sent"])
        # rows by decreasing support size (stable): k_solve's J.x passes then see homogeneous trip counts
        I["row_order"] = i32(sorted(range(len(sp["row_src"])), key=lambda r: -sp["src_nsup"][sp["row_src"][r]]))
        # per (dof, source) entry of the J^T f gather, packed: [J offset of the column | ns << 20, first row | nr << 16]
        src_row0 = [s[2] for s in self.src]
        ent = []
        for e in sp["dof_sent"]:
            k, pos = e & 0xffff, e >> 16
            jofs = sp["row_jadr"][src_row0[k]] + pos if sp["src_nrows"][k] else 0
            assert jofs < (1 << 20) and sp["src_nsup"][k] < (1 << 11) and src_row0[k] < (1 << 16) and sp["src_nrows"][k] < (1 << 15)
            ent += [jofs | (sp["src_nsup"][k] << 20), src_row0[k] | (sp["src_nrows"][k] << 16)]
        I["dof_ent"] = i32(ent)
        I["tri_src"] = i32([i * self.nv + j for i in range(self.nv) for j in range(i + 1)])
        I["triu_src"] = i32([j * self.nv + i for i in range(self.nv) for j in range(i)])
        # packed lower-triangle index p -> (row a, col b) as a | b << 16 (Hessian assembly walks support pairs by p)
        I["tri_ab"] = i32([a | (b << 16) for a in range(self.nv) for b in range(a + 1)])
        I["eq_type"], I["eq_obj1"], I["eq_obj2"] = i32(mj.eq_type), i32(mj.eq_obj1id), i32(mj.eq_obj2id)
        I["sens_type"], I["sens_objid"], I["sens_adr"] = i32(mj.sensor_type), i32(mj.sensor_objid), i32(mj.sensor_adr)
        I["sens_dim"], I["sens_datatype"], I["sens_stage"] = i32(mj.sensor_dim), i32(mj.sensor_datatype), i32(mj.sensor_needstage)
        I["sens_objtype"] = i32(mj.sensor_objtype)
        I["sens_reftype"] = i32(getattr(mj, "sensor_reftype", np.zeros(self.nsensor)))
        I["sens_refid"] = i32(getattr(mj, "sensor_refid", np.full(self.nsensor, -1)))
        # rangefinder candidate geoms: reference ray.precompute_ray_data :292-324 iterates geom types in this order
        rf = []
        for gt in (GeomType.PLANE, GeomType.SPHERE, GeomType.CAPSULE, GeomType.ELLIPSOID, GeomType.CYLINDER, GeomType.BOX):
            for g in np.nonzero(np.asarray(mj.geom_type) == gt)[0]:
                visible = (int(mj.geom_matid[g]) != -1) or (float(mj.geom_rgba[g][3]) != 0)
                rf.append(int(g) if visible else -1 - int(g))
        I["rf_geom"] = i32(rf)
        # convex tables
        cv = [(g, c) for g, c in enumerate(self.convex_tables) if c is not None]
        geom_cvx = np.full(self.ngeom, -1, np.int32)
        va = fa = ea = na_ = 0
        cvx = {k: [] for k in "geom nvert nface nfv nedge vertadr faceadr edgeadr normadr".split()}
        vert_all, norm_all, face_all, edge_all = [], [], [], []
        for ci, (g, c) in enumerate(cv):
            geom_cvx[g] = ci
            for k, v in (("geom", g), ("nvert", len(c.vert)), ("nface", c.face.shape[0]), ("nfv", c.face.shape[1]),
                         ("nedge", len(c.edge)), ("vertadr", va), ("faceadr", fa), ("edgeadr", ea), ("normadr", na_)):
                cvx[k].append(v)
            vert_all.append(c.vert.reshape(-1))
            norm_all.append(c.norm.reshape(-1))
            face_all.append(c.face.reshape(-1))
            edge_all.append(c.edge.reshape(-1))
            va += len(c.vert)
            na_ += c.face.shape[0]
            fa += c.face.size
            ea += len(c.edge)
        for k, v in cvx.items():
            I["cvx_" + k] = i32(v)
        I["geom_cvx"] = i32(geom_cvx)
        I["cvx_face"] = i32(np.concatenate(face_all)) if face_all else i32([])
        I["cvx_edge"] = i32(np.concatenate(edge_all)) if edge_all else i32([])

        o = mj.opt
        opt = np.zeros(len(L.OPT_SLOTS))
        vals = dict(timestep=o.timestep, impratio=o.impratio, tolerance=o.tolerance, ls_tolerance=o.ls_tolerance,
                    meaninertia=mj.stat.meaninertia, density=o.density, viscosity=o.viscosity)
        for k3, arr in (("gravity", o.gravity), ("wind", o.wind), ("magnetic", o.magnetic)):
            for a in range(3):
                vals[f"{k3}{a}"] = arr[a]
        for k, v in vals.items():
            opt[L.OPT_SLOTS.index(k)] = float(v)
        F["opt"] = opt
        for k in ("qpos0 qpos_spring body_pos body_quat body_ipos body_iquat body_mass body_subtreemass body_inertia body_gravcomp body_invweight0 "
                  "jnt_pos jnt_axis jnt_stiffness jnt_range jnt_margin jnt_solref jnt_solimp jnt_actfrcrange dof_armature "
                  "dof_damping dof_invweight0 dof_frictionloss dof_solref dof_solimp geom_pos geom_quat geom_size site_pos "
                  "site_quat wrap_prm eq_data eq_solref eq_solimp").split():
            F[k] = f(getattr(mj, k))
        for k in ("gear ctrlrange forcerange actrange gainprm biasprm dynprm lengthrange acc0").split():
            F["act_" + k] = f(getattr(mj, "actuator_" + k))
        for k, src_name in (("ten_stiffness", "tendon_stiffness"), ("ten_damping", "tendon_damping"), ("ten_armature", "tendon_armature"),
                            ("ten_lengthspring", "tendon_lengthspring"), ("ten_range", "tendon_range"), ("ten_margin", "tendon_margin"),
                            ("ten_invweight0", "tendon_invweight0"), ("ten_solref_lim", "tendon_solref_lim"),
                            ("ten_solimp_lim", "tendon_solimp_lim"), ("ten_frictionloss", "tendon_frictionloss"),
                            ("ten_solref_fri", "tendon_solref_fri"), ("ten_solimp_fri", "tendon_solimp_fri")):
            F[k] = f(getattr(mj, src_name))
        F["con_includemargin"], F["con_friction"], F["con_solref"] = f(self.con["include"]), f(self.con["friction"]), f(self.con["solref"])
        F["con_solreffriction"], F["con_solimp"] = f(self.con["solreffriction"]), f(self.con["solimp"])
        F["sens_cutoff"] = f(mj.sensor_cutoff)
        F["cvx_vert"] = np.concatenate(vert_all) if vert_all else f([])
        F["cvx_norm"] = np.concatenate(norm_all) if norm_all else f([])
        F["hf_size"], F["hf_data"], F["pair_rbound"] = f(getattr(mj, "hfield_size", [])), f(getattr(mj, "hfield_data", [])), f(pair_rb)

        # structurally non-zero entries of the packed lower triangle of M: (a, b) with dof b an ancestor-or-self of dof a
        # (support.make_m :50-80); a tendon armature couples arbitrary dofs, so then every entry counts
        nP = self.nv * (self.nv + 1) // 2
        if self.ntendon and np.any(np.asarray(mj.tendon_armature) != 0):
            m_nz = list(range(nP))
        else:
            m_nz = []
            for a in range(self.nv):
                b = a
                while b >= 0:
                    m_nz.append(a * (a + 1) // 2 + b)
                    b = int(mj.dof_parentid[b])
            m_nz.sort()
        I["m_nz"] = i32(m_nz)

        # Data field tables of the two output layouts (layout.py)
        sizes = dict(nq=self.nq, nv=self.nv, nu=self.nu, na=self.na, nbody=self.nbody, njnt=self.njnt, ngeom=self.ngeom,
                     nsite=self.nsite, ntendon=self.ntendon, ncon=self.ncon, nefc=self.nefc, nsensordata=self.nsensordata,
                     nsrc=len(self.src), ncand=self.ncand, ntopk=self.topk)
        self.field_shape, self.field_off, off = {}, {}, 0
        for name, expr in L.D_FIELDS:
            shp = tuple(int(s) for s in eval(expr, {}, sizes))
            self.field_shape[name], self.field_off[name] = shp, off
            off += int(np.prod(shp)) if shp else 1
        self.data_words = off
        nwords = lambda shp: int(np.prod(shp)) if shp else 1
        # -- "state" layout: public buffer (P table) + workspace (D table); see layout.py
        S_ = {n: i for i, n in enumerate(L.SRC_TYPES)}
        has_sens = self.nsensordata > 0
        simple_trn = all(int(t) in (TrnType.JOINT, TrnType.JOINTINPARENT) and int(mj.jnt_type[int(np.asarray(mj.actuator_trnid).reshape(-1, 2)[u, 0])]) in (JointType.HINGE, JointType.SLIDE)
                         for u, t in enumerate(np.asarray(mj.actuator_trntype)))
        need = set("qpos qvel act time act_dot xpos xquat xipos xanchor xaxis subtree_com crb cdof cvel qM qfrc_bias qfrc_passive "
                   "qfrc_gravcomp qfrc_actuator qfrc_smooth scr_cacc scr_cfrc scr_ovf".split())
        if len(self.pairs) or has_sens:
            need |= {"geom_xpos", "geom_xmat"}
        if has_sens:
            need |= {"site_xpos", "site_xmat", "xmat", "ximat", "actuator_length", "actuator_velocity", "actuator_force"}
        if any(t in (S_["EQ_CONNECT"], S_["EQ_WELD"]) for t, _, _ in self.src):
            need.add("xmat")
        if self.opt.has_fluid_params:
            need.add("ximat")
        if self.ntendon:
            need |= {"ten_length", "ten_J", "ten_velocity"}
        if int(mj.opt.integrator) == IntegratorType.IMPLICITFAST and not simple_trn:
            need.add("actuator_moment")
        if self.ncon:
            need |= {"contact_dist", "contact_pos", "contact_frame"}
        if self.topk:
            need |= {"scr_cand_dist", "scr_cand_pos", "scr_cand_frame", "scr_consrc"}
        if self.nefc:
            need |= {"efc_J", "efc_D", "efc_aref", "scr_live"}
            if self.nf:
                need.add("efc_frictionloss")
        ws_shape = dict(self.field_shape)
        ws_shape["qM"], ws_shape["efc_J"] = (nP,), (self.sup["jc_words"],)
        ws_off, off = {}, 0
        for name, _ in L.D_FIELDS:
            if name in need:
                ws_off[name] = off
                off += nwords(ws_shape[name])
            else:
                ws_off[name] = -1
        ws_words = off
        pub_off, off = {}, 0
        for name in L.P_FIELDS:
            if name == "qfrc_constraint" and self.nefc:
                pub_off[name] = -1   # only a model without constraint rows passes it through (forward.py:390-396)
                continue
            pub_off[name] = off
            off += nwords(self.field_shape[name])
        pub_words = off
        # geoms the state layout has to place: members of a collision pair, or everything when sensors may read them
        geom_need = np.zeros(self.ngeom, np.int32)
        if has_sens:
            geom_need[:] = 1
        for pr in self.pairs:
            geom_need[pr["g1"]] = geom_need[pr["g2"]] = 1
        self.state_layout = dict(pub_off=pub_off, pub_words=pub_words, ws_off=ws_off, ws_words=ws_words, ws_shape=ws_shape)

        S = {n: i for i, n in enumerate(L.SRC_TYPES)}
        cnt = lambda t: sum(1 for s in self.src if s[0] == S[t])
        st = np.asarray(mj.sensor_needstage)
        H = dict(magic=L.MAGIC, is_f64=int(f64), nq=self.nq, nv=self.nv, nu=self.nu, na=self.na, nbody=self.nbody, njnt=self.njnt,
                 ngeom=self.ngeom, nsite=self.nsite, nmocap=self.nmocap, ntendon=self.ntendon, nwrap=self.nwrap, neq=self.neq,
                 nsensor=self.nsensor, nsensordata=self.nsensordata, ncon=self.ncon, nefc=self.nefc, ne=self.ne, nf=self.nf,
                 nl=self.nl, npair_cand=len(self.pairs), nsrc=len(self.src), disableflags=int(o.disableflags), solver=int(o.solver),
                 cone=int(o.cone), integrator=int(o.integrator), iterations=int(o.iterations), ls_iterations=int(o.ls_iterations),
                 data_words=self.data_words, has_gravcomp=int(np.any(np.asarray(mj.body_gravcomp) != 0)),
                 has_fluid=int(self.opt.has_fluid_params), nv_pad=self.nv | 1, nlim_sh=cnt("LIM_SH"), nlim_ball=cnt("LIM_BALL"),
                 nfric_dof=cnt("FRIC_DOF"), nfric_ten=cnt("FRIC_TEN"), nlim_ten=cnt("LIM_TEN"), ne_connect=cnt("EQ_CONNECT"),
                 ne_weld=cnt("EQ_WELD"), ne_joint=cnt("EQ_JOINT"), maskwords=self.maskwords, nsens_pos=int((st == 1).sum()),
                 nsens_vel=int((st == 2).sum()), nsens_acc=int((st == 3).sum()), nconvex=len(cv), nrfgeom=len(rf), jc_words=self.sup["jc_words"],
                 live_rcap=self.live_caps[0], live_jcap=self.live_caps[1], n_mnz=len(m_nz), topk=self.topk, ncand=self.ncand,
                 nsrc1=sum(1 for n in self.sup["src_nsup"] if n == 1))
        nH, nI, nF, nD, nPf = len(L.H_SCALARS), len(L.I_ARRAYS), len(L.F_ARRAYS), len(L.D_FIELDS), len(L.P_FIELDS)
        def assemble(keep=None, extra=None, state=False):
            """One blob; with ``keep`` only the named arrays carry data (the others get length 0)."""
            Hh = dict(H)
            Hh.update(extra or {})
            if state:
                Hh.update(layout_state=1, data_words=ws_words, pub_words=pub_words)
                I["geom_need"] = geom_need
            else:
                Hh.update(layout_state=0, pub_words=self.data_words)
                I["geom_need"] = np.ones(self.ngeom, np.int32)
            hdr = np.zeros(nH + nI + nF + 2 * nD + nPf, dtype=np.int32)
            ints, pos = [], hdr.size
            for k, name in enumerate(L.I_ARRAYS):
                a = I[name] if keep is None or name in keep else np.zeros(0, np.int32)
                hdr[nH + k] = pos
                ints.append(a)
                pos += a.size
            pos_pad = (pos + 3) // 4 * 4
            flts, fpos = [], 0
            for k, name in enumerate(L.F_ARRAYS):
                a = F[name] if keep is None or name in keep else np.zeros(0)
                hdr[nH + nI + k] = fpos
                flts.append(a)
                fpos += a.size
                if fpos % 2:  # keep every float array 8-byte aligned in f32 / 16-byte pairs in f64
                    flts.append(np.zeros(1))
                    fpos += 1
            for k, (name, _) in enumerate(L.D_FIELDS):
                hdr[nH + nI + nF + k] = ws_off[name] if state else self.field_off[name]
                hdr[nH + nI + nF + nD + k] = nwords(ws_shape[name] if state else self.field_shape[name])
            for k, name in enumerate(L.P_FIELDS):
                hdr[nH + nI + nF + 2 * nD + k] = pub_off[name] if state else self.field_off[name]
            fl_arr = np.concatenate(flts) if flts else np.zeros(0)
            fl_bytes = fl_arr.astype(np.float64 if f64 else np.float32).view(np.int32)
            Hh["float_base_words"] = pos_pad
            Hh["blob_words"] = (pos_pad + fl_bytes.size + 3) // 4 * 4
            for k, name in enumerate(L.H_SCALARS):
                hdr[k] = np.int32(np.uint32(Hh.get(name, 0)).view(np.int32)) if name == "magic" else Hh.get(name, 0)
            out = np.zeros(Hh["blob_words"], dtype=np.int32)
            out[: hdr.size] = hdr
            ii = np.concatenate(ints) if ints else np.zeros(0, np.int32)
            out[hdr.size: hdr.size + ii.size] = ii
            out[pos_pad: pos_pad + fl_bytes.size] = fl_bytes
            return out

        # the solver kernel stages a reduced blob (only the tables it reads) so that more envs fit in shared memory
        solve_keep = set("jnt_type jnt_qposadr jnt_dofadr act_trntype act_trnid act_gear act_dyntype act_gaintype act_biastype act_actadr act_actnum act_actlimited "
                         "src_row src_nrows src_nsup src_cadr sup_col row_src row_jadr tri_ab dof_sadr dof_sent jc_src m_nz row_order dof_ent "
                         "opt dof_damping act_actrange act_gainprm act_biasprm act_dynprm ten_damping".split())
        def both(state):
            # the state layout stores efc_J already compacted: its solver needs no dense-word table (2 KB of shared memory)
            solve_blob = assemble(solve_keep - ({"jc_src"} if state else set()), None, state)
            full = assemble(None, {"solve_blob_words": int(solve_blob.size)}, state)
            return np.concatenate([full, solve_blob])

        self.blob_np = both(False)
        self.state_layout["blob_np"] = both(True)
        self.blob = torch.from_numpy(self.blob_np.copy())
        self._I, self._F = I, F

    # ---- device movement (reference types._model_to :991-1010)
    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            return self
        new = object.__new__(Model)
        new.__dict__.update(self.__dict__)
        new.device = device
        new.blob = self.blob.to(device)
        new._handle = None
        new._handles = {}
        new._mid = None
        new.__dict__.pop("_static_contact_cache", None)
        return new

    def cuda(self):
        return self.to("cuda")

    def clone(self):
        return self.to(self.device)


def device_put_model(mj, dtype=None) -> Model:
    return Model(mj, torch.get_default_dtype() if dtype is None else dtype)
