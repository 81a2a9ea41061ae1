"""MJCF mini-compiler: XML -> MjModelLite (numpy arrays under mujoco.MjModel attribute names).

The reference obtains its compiled model from the third-party MuJoCo C library
(``mujoco.MjModel.from_xml_*``; call sites reference test_util.py:337-350, zoo/base.py:91),
which is not available in this environment.  This module restates the
part of MuJoCo's published compile semantics the hot-path configs need:
default classes, frames (quat/euler/axisangle/xyaxes/zaxis/fromto), inertia from geoms,
settotalmass, joints/dofs tables, contact pairs/excludes, motors, sites, sensors, convex meshes,
and the ``mj_setConst`` quantities at qpos0 (invweight0, acc0, meaninertia).

PARITY UNPINNED at this boundary: no MuJoCo binary exists here to check against; the oracle and
the CUDA path both consume exactly these compiled numbers, so step parity does not depend on it.
"""

from __future__ import annotations

import copy
import os
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .mjmodel import (
    CamLightType,
    DataType,
    DisableBit,
    EnableBit,
    EqType,
    GeomType,
    JointType,
    MjModelLite,
    MjOptionLite,
    MjStatisticLite,
    ObjType,
    SensorType,
    Stage,
    TrnType,
    WrapType,
    mjMINVAL,
    mjNBIAS,
    mjNDYN,
    mjNEQDATA,
    mjNGAIN,
)

# ----------------------------------------------------------------------------- small math


def _f(s, n=None, default=None):
    if s is None:
        return default
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} values, got {v.size}: {s!r}")
    return v


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    n = np.linalg.norm(v)
    return v / n if n > 0 else v


def quat_mul(a, b):
    return np.array(
        [
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
        ]
    )


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def mat_to_quat(m):
    """Rotation matrix -> unit quaternion (w>=0 branch of Shepperd's method)."""
    m = np.asarray(m, dtype=np.float64)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def rotate(v, q):
    return quat_to_mat(q) @ np.asarray(v, dtype=np.float64)


def axisangle_to_quat(axis, angle):
    axis = _normalize(axis)
    return np.concatenate([[np.cos(angle / 2)], axis * np.sin(angle / 2)])


def z_to_quat(vec):
    """Minimal rotation taking +z to ``vec`` (MuJoCo mjuu_z2quat)."""
    vec = _normalize(vec)
    axis = np.cross([0.0, 0.0, 1.0], vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = np.arctan2(s, vec[2])
    return np.concatenate([[np.cos(ang / 2)], axis * np.sin(ang / 2)])


def euler_to_quat(e, seq="xyz"):
    q = np.array([1.0, 0, 0, 0])
    for ang, c in zip(e, seq):
        ax = {"x": [1.0, 0, 0], "y": [0, 1.0, 0], "z": [0, 0, 1.0]}[c.lower()]
        qa = axisangle_to_quat(ax, ang)
        q = quat_mul(q, qa) if c.islower() else quat_mul(qa, q)
    return q


def _eig_inertia(inertia):
    """Principal moments (descending) and a right-handed eigen-frame quaternion."""
    w, v = np.linalg.eigh(inertia)
    order = np.argsort(-w, kind="stable")
    w, v = w[order], v[:, order]
    # canonical sign: make largest-magnitude component of each axis positive, then fix handedness
    for k in range(3):
        j = int(np.argmax(np.abs(v[:, k])))
        if v[j, k] < 0:
            v[:, k] = -v[:, k]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    return w, mat_to_quat(v)


# ----------------------------------------------------------------------------- defaults

_DEFAULT_ELEMS = (
    "geom",
    "joint",
    "site",
    "camera",
    "light",
    "pair",
    "equality",
    "tendon",
    "general",
    "motor",
    "position",
    "velocity",
    "mesh",
    "material",
)
_ACT_SHORTCUTS = ("general", "motor", "position", "velocity", "intvelocity", "damper", "cylinder", "muscle", "adhesion")


class _Defaults:
    def __init__(self):
        self.classes: dict[str, dict[str, dict[str, str]]] = {"main": {k: {} for k in _DEFAULT_ELEMS}}
        self.parent: dict[str, str | None] = {"main": None}

    def parse(self, node, parent_name=None):
        name = node.get("class", "main" if parent_name is None else None)
        if name is None:
            raise ValueError("nested <default> requires class")
        if parent_name is None and name != "main":
            # top-level default with explicit class name acts as the root class
            self.classes[name] = copy.deepcopy(self.classes["main"])
            self.parent[name] = "main"
        elif parent_name is not None:
            self.classes[name] = copy.deepcopy(self.classes[parent_name])
            self.parent[name] = parent_name
        cls = self.classes[name]
        for child in node:
            if child.tag == "default":
                continue
            tag = child.tag
            cls.setdefault(tag, {})
            cls[tag].update(child.attrib)
            if tag in _ACT_SHORTCUTS and tag != "general":
                # actuator shortcuts share the "general" default slot in MuJoCo
                cls["general"].update(child.attrib)
        for child in node:
            if child.tag == "default":
                self.parse(child, name)

    def resolve(self, tag, elem, childclass):
        cname = elem.get("class", childclass or "main")
        if cname not in self.classes:
            raise ValueError(f"unknown default class {cname!r}")
        src = "general" if tag in _ACT_SHORTCUTS else tag
        attrs = dict(self.classes[cname].get(src, {}))
        if tag in _ACT_SHORTCUTS and tag != "general":
            attrs.update(self.classes[cname].get(tag, {}))
        attrs.update(elem.attrib)
        return attrs


# ----------------------------------------------------------------------------- mesh helpers


def _read_stl(path):
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:5].lower() == b"solid" and b"facet" in data[:512]:
        verts = []
        for mt in re.finditer(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", data):
            verts.append([float(mt.group(1)), float(mt.group(2)), float(mt.group(3))])
        tri = np.array(verts, dtype=np.float64).reshape(-1, 3, 3)
    else:
        n = struct.unpack("<I", data[80:84])[0]
        rec = np.frombuffer(data, dtype=np.uint8, count=n * 50, offset=84).reshape(n, 50)
        tri = rec[:, 12:48].copy().view("<f4").reshape(n, 3, 3).astype(np.float64)
    return tri


def _polyhedron_mass_props(verts, faces):
    """Volume, CoM and inertia-about-CoM (unit density) of a closed triangulated surface."""
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    det = np.einsum("ij,ij->i", v0, np.cross(v1, v2))
    vol = det.sum() / 6.0
    com = ((v0 + v1 + v2) / 4.0 * det[:, None]).sum(0) / 6.0 / vol
    # second moments via canonical tetrahedron integration
    S = np.zeros((3, 3))
    for a, b, c, d in zip(v0 - com, v1 - com, v2 - com, det):
        m = np.stack([a, b, c])
        cov = (m.T @ (np.ones((3, 3)) + np.eye(3)) @ m) / 120.0
        S += d * cov
    inertia = np.trace(S) * np.eye(3) - S
    return vol, com, inertia


def _compile_mesh(path, scale):
    from scipy.spatial import ConvexHull

    tri = _read_stl(path) * scale[None, None, :]
    raw = tri.reshape(-1, 3)
    verts, inv = np.unique(np.round(raw, 12), axis=0, return_inverse=True)
    faces = inv.reshape(-1, 3)
    # orientation: make sure outward (positive volume)
    vol, com, inertia = _polyhedron_mass_props(verts, faces)
    if vol < 0:
        faces = faces[:, ::-1].copy()
        vol, com, inertia = _polyhedron_mass_props(verts, faces)
    w, q = _eig_inertia(inertia)
    R = quat_to_mat(q)
    verts_c = (verts - com) @ R  # express in principal frame centred at CoM
    hull = ConvexHull(verts_c)
    return dict(vert=verts_c, face=faces.astype(np.int32), pos=com, quat=q, volume=vol, inertia=w, hull_vertices=np.sort(hull.vertices))


# ----------------------------------------------------------------------------- compiler


class _Body:
    def __init__(self, name, parent, pos, quat, childclass):
        self.name, self.parent, self.pos, self.quat, self.childclass = name, parent, pos, quat, childclass
        self.joints, self.geoms, self.sites, self.cams, self.lights = [], [], [], [], []
        self.inertial = None
        self.mocap = False
        self.gravcomp = 0.0
        self.id = -1


def _bool(s, default=None):
    if s is None:
        return default
    return {"true": True, "false": False, "auto": None}[s]


class Compiler:
    def __init__(self, root: ET.Element, basedir: str = "."):
        self.root = root
        self.basedir = basedir
        self.defaults = _Defaults()
        c = root.find("compiler")
        ca = c.attrib if c is not None else {}
        self.degree = ca.get("angle", "degree") == "degree"
        self.eulerseq = ca.get("eulerseq", "xyz")
        self.autolimits = _bool(ca.get("autolimits"), True)
        self.inertiafromgeom = ca.get("inertiafromgeom", "auto")
        self.settotalmass = float(ca.get("settotalmass", -1))
        self.boundmass = float(ca.get("boundmass", 0))
        self.boundinertia = float(ca.get("boundinertia", 0))
        self.meshdir = ca.get("meshdir", ca.get("assetdir", ""))
        for d in root.findall("default"):
            self.defaults.parse(d, None)

    # -- frames
    def _ang(self, a):
        return np.deg2rad(a) if self.degree else a

    def _orient(self, a, default=None):
        if "quat" in a:
            return _normalize(_f(a["quat"], 4))
        if "axisangle" in a:
            v = _f(a["axisangle"], 4)
            return axisangle_to_quat(v[:3], self._ang(v[3]))
        if "euler" in a:
            return euler_to_quat(self._ang(_f(a["euler"], 3)), self.eulerseq)
        if "xyaxes" in a:
            v = _f(a["xyaxes"], 6)
            x = _normalize(v[:3])
            y = v[3:] - x * np.dot(x, v[3:])
            y = _normalize(y)
            z = np.cross(x, y)
            return mat_to_quat(np.stack([x, y, z], axis=1))
        if "zaxis" in a:
            return z_to_quat(_f(a["zaxis"], 3))
        return np.array([1.0, 0, 0, 0]) if default is None else default

    # -- tree parse
    def _parse_body(self, node, parent, childclass, bodies):
        a = node.attrib
        cc = a.get("childclass", childclass)
        if parent is None:
            b = _Body("world", None, np.zeros(3), np.array([1.0, 0, 0, 0]), cc)
        else:
            b = _Body(a.get("name"), parent, _f(a.get("pos"), 3, np.zeros(3)), self._orient(a), cc)
            b.mocap = _bool(a.get("mocap"), False)
            b.gravcomp = float(a.get("gravcomp", 0))
        b.id = len(bodies)
        bodies.append(b)
        for ch in node:
            t = ch.tag
            if t in ("joint", "freejoint"):
                if t == "freejoint":
                    attrs = dict(ch.attrib)
                    attrs["type"] = "free"
                else:
                    attrs = self.defaults.resolve("joint", ch, cc)
                b.joints.append(attrs)
            elif t == "geom":
                b.geoms.append(self.defaults.resolve("geom", ch, cc))
            elif t == "site":
                b.sites.append(self.defaults.resolve("site", ch, cc))
            elif t == "camera":
                b.cams.append(self.defaults.resolve("camera", ch, cc))
            elif t == "light":
                b.lights.append(self.defaults.resolve("light", ch, cc))
            elif t == "inertial":
                b.inertial = dict(ch.attrib)
        for ch in node:
            if ch.tag == "body":
                self._parse_body(ch, b, cc, bodies)
        return b

    # -- geoms
    def _geom(self, a, meshes):
        tname = a.get("type", "sphere")
        if "mesh" in a and "type" not in a:
            tname = "mesh"
        gtype = GeomType[tname.upper()]
        size = np.zeros(3)
        s = _f(a.get("size"), None, np.zeros(0))
        size[: s.size] = s
        pos = _f(a.get("pos"), 3, np.zeros(3))
        quat = self._orient(a)
        if "fromto" in a:
            ft = _f(a["fromto"], 6)
            vec = ft[:3] - ft[3:]
            half = np.linalg.norm(vec) / 2
            pos = (ft[:3] + ft[3:]) / 2
            quat = z_to_quat(vec)
            if gtype in (GeomType.CAPSULE, GeomType.CYLINDER):
                size[1] = half
            else:
                size[2] = half
        g = dict(type=gtype, size=size, pos=pos, quat=quat, name=a.get("name"))
        g["contype"] = int(a.get("contype", 1))
        g["conaffinity"] = int(a.get("conaffinity", 1))
        g["condim"] = int(a.get("condim", 3))
        g["priority"] = int(a.get("priority", 0))
        g["group"] = int(a.get("group", 0))
        fr = np.array([1.0, 0.005, 0.0001])
        f = _f(a.get("friction"), None, np.zeros(0))
        fr[: f.size] = f
        g["friction"] = fr
        g["solref"] = _f(a.get("solref"), 2, np.array([0.02, 1.0]))
        si = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        s_ = _f(a.get("solimp"), None, np.zeros(0))
        si[: s_.size] = s_
        g["solimp"] = si
        g["solmix"] = float(a.get("solmix", 1))
        g["margin"] = float(a.get("margin", 0))
        g["gap"] = float(a.get("gap", 0))
        g["density"] = float(a.get("density", 1000))
        g["mass"] = float(a["mass"]) if "mass" in a else None
        rgba = _f(a.get("rgba"), 4, np.array([0.5, 0.5, 0.5, 1.0]))
        g["rgba"] = rgba
        g["matid"] = self._materials["index"].get(a.get("material"), -1) if a.get("material") else -1
        g["dataid"] = -1
        if gtype == GeomType.HFIELD:
            hid = meshes["hfield_index"][a["hfield"]]
            g["dataid"] = hid
            hs = meshes["hfields"][hid]["size"]
            g["size"] = size = np.array([hs[0], hs[1], 0.25 * hs[2] + 0.5 * hs[3]])   # MuJoCo: geom size of a height field (x, y, z half sizes)
        if gtype == GeomType.MESH:
            mid = meshes["index"][a["mesh"]]
            mm = meshes["list"][mid]
            g["dataid"] = mid
            # compose geom frame with mesh CoM / principal frame
            g["pos"] = pos + rotate(mm["pos"], quat)
            g["quat"] = quat_mul(quat, mm["quat"])
        # mass properties in geom frame
        g["volume"], g["inertia_unit"] = self._geom_massprops(g, meshes)
        # bounding radius
        if gtype == GeomType.SPHERE:
            g["rbound"] = size[0]
        elif gtype == GeomType.CAPSULE:
            g["rbound"] = size[0] + size[1]
        elif gtype == GeomType.CYLINDER:
            g["rbound"] = np.sqrt(size[0] ** 2 + size[1] ** 2)
        elif gtype in (GeomType.BOX, GeomType.ELLIPSOID):
            g["rbound"] = np.linalg.norm(size) if gtype == GeomType.BOX else np.max(size)
        elif gtype == GeomType.MESH:
            g["rbound"] = float(np.max(np.linalg.norm(meshes["list"][g["dataid"]]["vert"], axis=1)))
        elif gtype == GeomType.HFIELD:
            hs = meshes["hfields"][g["dataid"]]["size"]
            g["rbound"] = float(np.sqrt(hs[0] ** 2 + hs[1] ** 2 + max(hs[2], hs[3]) ** 2))
        else:
            g["rbound"] = 0.0
        return g

    @staticmethod
    def _geom_massprops(g, meshes):
        """Volume and unit-mass diagonal inertia in the geom frame."""
        t, s = g["type"], g["size"]
        if t == GeomType.SPHERE:
            return 4 / 3 * np.pi * s[0] ** 3, np.full(3, 2 * s[0] ** 2 / 5)
        if t == GeomType.CAPSULE:
            r, h = s[0], 2 * s[1]
            vol = np.pi * (r * r * h + 4 * r**3 / 3)
            sphere_m = 4 * r / (4 * r + 3 * h)
            cyl_m = 1 - sphere_m
            i = np.zeros(3)
            i[0] = i[1] = cyl_m * (3 * r * r + h * h) / 12
            i[2] = cyl_m * r * r / 2
            si = 2 * sphere_m * r * r / 5
            i[0] += si + sphere_m * h * (3 * r + 2 * h) / 8
            i[1] += si + sphere_m * h * (3 * r + 2 * h) / 8
            i[2] += si
            return vol, i
        if t == GeomType.CYLINDER:
            r, h = s[0], 2 * s[1]
            return np.pi * r * r * h, np.array([(3 * r * r + h * h) / 12, (3 * r * r + h * h) / 12, r * r / 2])
        if t == GeomType.ELLIPSOID:
            return 4 / 3 * np.pi * s[0] * s[1] * s[2], np.array(
                [(s[1] ** 2 + s[2] ** 2) / 5, (s[0] ** 2 + s[2] ** 2) / 5, (s[0] ** 2 + s[1] ** 2) / 5]
            )
        if t == GeomType.BOX:
            return 8 * s[0] * s[1] * s[2], np.array(
                [(s[1] ** 2 + s[2] ** 2) / 3, (s[0] ** 2 + s[2] ** 2) / 3, (s[0] ** 2 + s[1] ** 2) / 3]
            )
        if t == GeomType.MESH:
            mm = meshes["list"][g["dataid"]]
            return mm["volume"], mm["inertia"] / mm["volume"]
        return 0.0, np.zeros(3)

    # -- main
    def compile(self) -> MjModelLite:
        root = self.root
        m = MjModelLite()
        # meshes
        meshes = {"index": {}, "list": [], "names": [], "hfield_index": {}, "hfields": []}
        asset = root.find("asset")
        if asset is not None:
            # height fields: <hfield name nrow ncol size="rx ry elevation_z base_z" elevation="nrow*ncol values, row-major"/>; like
            # MuJoCo's compiler the data are shifted / scaled into [0, 1] (file-based height fields are not supported)
            for he in asset.findall("hfield"):
                a = dict(he.attrib)
                if "file" in a:
                    raise NotImplementedError("hfield from file is not supported; give nrow / ncol / elevation")
                nrow, ncol = int(a["nrow"]), int(a["ncol"])
                data = _f(a.get("elevation"), None, np.zeros(nrow * ncol))
                if data.size != nrow * ncol:
                    raise ValueError(f"hfield {a.get('name')}: elevation has {data.size} values, expected nrow * ncol = {nrow * ncol}")
                lo, hi = float(data.min()), float(data.max())
                data = (data - lo) / (hi - lo) if hi > lo else np.zeros_like(data)
                meshes["hfield_index"][a.get("name")] = len(meshes["hfields"])
                meshes["hfields"].append(dict(name=a.get("name"), nrow=nrow, ncol=ncol, size=_f(a["size"], 4), data=data))
            for me in asset.findall("mesh"):
                a = self.defaults.resolve("mesh", me, None)
                path = os.path.join(self.basedir, self.meshdir, a["file"])
                name = a.get("name", os.path.splitext(os.path.basename(a["file"]))[0])
                meshes["index"][name] = len(meshes["list"])
                meshes["names"].append(name)
                meshes["list"].append(_compile_mesh(path, _f(a.get("scale"), 3, np.ones(3))))

        materials = {"index": {}, "rgba": []}
        if asset is not None:
            for ma in asset.findall("material"):
                a = self.defaults.resolve("material", ma, None)
                materials["index"][a.get("name")] = len(materials["rgba"])
                materials["rgba"].append(_f(a.get("rgba"), 4, np.ones(4)))
        self._materials = materials

        bodies: list[_Body] = []
        self._parse_body(root.find("worldbody"), None, None, bodies)
        nbody = len(bodies)

        # ---------------- bodies / joints / dofs
        body_parentid = np.array([0 if b.parent is None else b.parent.id for b in bodies], dtype=np.int32)
        body_pos = np.stack([b.pos for b in bodies])
        body_quat = np.stack([b.quat for b in bodies])
        jnt_type, jnt_bodyid, jnt_pos, jnt_axis, jnt_qposadr, jnt_dofadr = [], [], [], [], [], []
        jnt_stiffness, jnt_range, jnt_limited, jnt_margin, jnt_solref, jnt_solimp = [], [], [], [], [], []
        jnt_actfrcrange, jnt_actfrclimited, jnt_actgravcomp, jnt_group, jnt_names = [], [], [], [], []
        dof_bodyid, dof_jntid, dof_parentid, dof_armature, dof_damping, dof_frictionloss = [], [], [], [], [], []
        dof_solref, dof_solimp = [], []
        qpos0, qpos_spring = [], []
        body_jntadr = np.full(nbody, -1, np.int32)
        body_jntnum = np.zeros(nbody, np.int32)
        body_dofadr = np.full(nbody, -1, np.int32)
        body_dofnum = np.zeros(nbody, np.int32)
        last_dof_of_body = np.full(nbody, -1, np.int32)  # last dof in the body's own chain (incl. ancestors)
        for b in bodies:
            pid = 0 if b.parent is None else b.parent.id
            last = last_dof_of_body[pid] if b.parent is not None else -1
            if b.joints:
                body_jntadr[b.id] = len(jnt_type)
                body_dofadr[b.id] = len(dof_bodyid)
            for a in b.joints:
                jt = JointType[a.get("type", "hinge").upper()]
                jid = len(jnt_type)
                jnt_type.append(int(jt))
                jnt_bodyid.append(b.id)
                jnt_names.append(a.get("name"))
                jnt_qposadr.append(len(qpos0))
                jnt_dofadr.append(len(dof_bodyid))
                jnt_group.append(int(a.get("group", 0)))
                if jt == JointType.FREE:
                    jnt_pos.append(np.zeros(3))
                    jnt_axis.append(np.array([0.0, 0, 1]))
                elif jt == JointType.BALL:
                    jnt_pos.append(_f(a.get("pos"), 3, np.zeros(3)))
                    jnt_axis.append(np.array([0.0, 0, 1]))
                else:
                    jnt_pos.append(_f(a.get("pos"), 3, np.zeros(3)))
                    jnt_axis.append(_normalize(_f(a.get("axis"), 3, np.array([0.0, 0, 1]))))
                jnt_stiffness.append(float(a.get("stiffness", 0)))
                rng = _f(a.get("range"), 2, np.zeros(2))
                if jt in (JointType.HINGE, JointType.BALL):
                    rng = self._ang(rng)
                lim = _bool(a.get("limited"), None)
                has_range = "range" in a and (rng[0] != 0 or rng[1] != 0)
                if lim is None:
                    if not self.autolimits and "range" in a:
                        raise ValueError("limited=auto with autolimits=false")
                    lim = has_range
                jnt_limited.append(1 if lim else 0)
                jnt_range.append(rng)
                jnt_margin.append(float(a.get("margin", 0)))
                jnt_solref.append(_f(a.get("solreflimit"), 2, np.array([0.02, 1.0])))
                si = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                s_ = _f(a.get("solimplimit"), None, np.zeros(0))
                si[: s_.size] = s_
                jnt_solimp.append(si)
                afr = _f(a.get("actuatorfrcrange"), 2, np.zeros(2))
                afl = _bool(a.get("actuatorfrclimited"), None)
                if afl is None:
                    afl = "actuatorfrcrange" in a and (afr[0] != 0 or afr[1] != 0)
                jnt_actfrcrange.append(afr)
                jnt_actfrclimited.append(1 if afl else 0)
                jnt_actgravcomp.append(1 if _bool(a.get("actuatorgravcomp"), False) else 0)
                ref = float(a.get("ref", 0))
                sref = float(a.get("springref", 0))
                if jt in (JointType.HINGE,):
                    ref, sref = self._ang(ref), self._ang(sref)
                if jt == JointType.FREE:
                    q0 = np.concatenate([b.pos, b.quat])
                    qpos0.extend(q0)
                    qpos_spring.extend(q0)
                elif jt == JointType.BALL:
                    qpos0.extend([1.0, 0, 0, 0])
                    qpos_spring.extend([1.0, 0, 0, 0])
                else:
                    qpos0.append(ref)
                    qpos_spring.append(sref)
                sfr = _f(a.get("solreffriction"), 2, np.array([0.02, 1.0]))
                sfi = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                s_ = _f(a.get("solimpfriction"), None, np.zeros(0))
                sfi[: s_.size] = s_
                for _k in range(jt.dof_width()):
                    did = len(dof_bodyid)
                    dof_bodyid.append(b.id)
                    dof_jntid.append(jid)
                    dof_parentid.append(last)
                    last = did
                    dof_armature.append(float(a.get("armature", 0)))
                    dof_damping.append(float(a.get("damping", 0)))
                    dof_frictionloss.append(float(a.get("frictionloss", 0)))
                    dof_solref.append(sfr)
                    dof_solimp.append(sfi)
                body_jntnum[b.id] += 1
                body_dofnum[b.id] += jt.dof_width()
            last_dof_of_body[b.id] = last
        nq, nv, njnt = len(qpos0), len(dof_bodyid), len(jnt_type)

        body_weldid = np.zeros(nbody, np.int32)
        body_rootid = np.zeros(nbody, np.int32)
        body_mocapid = np.full(nbody, -1, np.int32)
        nmocap = 0
        for b in bodies[1:]:
            pid = b.parent.id
            body_weldid[b.id] = b.id if b.joints else body_weldid[pid]
            body_rootid[b.id] = b.id if pid == 0 else body_rootid[pid]
            if b.mocap:
                body_mocapid[b.id] = nmocap
                nmocap += 1

        dof_Madr = np.zeros(nv, np.int32)
        adr = 0
        for i in range(nv):
            dof_Madr[i] = adr
            j = i
            while j >= 0:
                adr += 1
                j = dof_parentid[j]
        nM = adr

        # ---------------- geoms
        geoms = []
        body_geomadr = np.full(nbody, -1, np.int32)
        body_geomnum = np.zeros(nbody, np.int32)
        for b in bodies:
            if b.geoms:
                body_geomadr[b.id] = len(geoms)
            for a in b.geoms:
                g = self._geom(a, meshes)
                g["bodyid"] = b.id
                geoms.append(g)
                body_geomnum[b.id] += 1
        ngeom = len(geoms)

        # ---------------- inertia
        body_mass = np.zeros(nbody)
        body_inertia = np.zeros((nbody, 3))
        body_ipos = np.zeros((nbody, 3))
        body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
        for b in bodies[1:]:
            use_geoms = self.inertiafromgeom == "true" or (self.inertiafromgeom == "auto" and b.inertial is None)
            if not use_geoms and b.inertial is not None:
                ia = b.inertial
                body_mass[b.id] = float(ia["mass"])
                body_ipos[b.id] = _f(ia.get("pos"), 3, np.zeros(3))
                if "fullinertia" in ia:
                    fi = _f(ia["fullinertia"], 6)
                    full = np.array([[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]])
                    w, q = _eig_inertia(full)
                    body_inertia[b.id], body_iquat[b.id] = w, q
                else:
                    body_inertia[b.id] = _f(ia["diaginertia"], 3)
                    body_iquat[b.id] = self._orient(ia)
                continue
            gl = [g for g in geoms if g["bodyid"] == b.id]
            masses = []
            for g in gl:
                gm = g["mass"] if g["mass"] is not None else g["density"] * g["volume"]
                masses.append(gm)
            mtot = float(np.sum(masses)) if masses else 0.0
            if mtot <= 0:
                continue
            com = sum(gm * g["pos"] for gm, g in zip(masses, gl)) / mtot
            full = np.zeros((3, 3))
            for gm, g in zip(masses, gl):
                R = quat_to_mat(g["quat"])
                Ig = R @ np.diag(g["inertia_unit"] * gm) @ R.T
                dpos = g["pos"] - com
                full += Ig + gm * (np.dot(dpos, dpos) * np.eye(3) - np.outer(dpos, dpos))
            body_mass[b.id] = mtot
            body_ipos[b.id] = com
            if len(gl) == 1:
                # single geom: inertial frame coincides with the geom frame (MuJoCo keeps it un-sorted)
                body_inertia[b.id] = gl[0]["inertia_unit"] * masses[0]
                body_iquat[b.id] = gl[0]["quat"]
            else:
                w, q = _eig_inertia(full)
                body_inertia[b.id], body_iquat[b.id] = w, q
        if self.boundmass > 0:
            body_mass[1:] = np.maximum(body_mass[1:], self.boundmass)
        if self.boundinertia > 0:
            body_inertia[1:] = np.maximum(body_inertia[1:], self.boundinertia)
        if self.settotalmass > 0:
            scale = self.settotalmass / body_mass.sum()
            body_mass *= scale
            body_inertia *= scale
        body_subtreemass = body_mass.copy()
        for i in range(nbody - 1, 0, -1):
            body_subtreemass[body_parentid[i]] += body_subtreemass[i]

        # ---------------- sites / cameras / lights
        sites, cams, lights = [], [], []
        for b in bodies:
            for a in b.sites:
                st = a.get("type", "sphere")
                size = np.full(3, 0.005)
                s = _f(a.get("size"), None, np.zeros(0))
                size[: s.size] = s
                pos = _f(a.get("pos"), 3, np.zeros(3))
                quat = self._orient(a)
                if "fromto" in a:
                    ft = _f(a["fromto"], 6)
                    vec = ft[:3] - ft[3:]
                    pos = (ft[:3] + ft[3:]) / 2
                    quat = z_to_quat(vec)
                    size[1] = np.linalg.norm(vec) / 2
                rgba = _f(a.get("rgba"), 4, np.array([0.5, 0.5, 0.5, 1.0]))
                sites.append(
                    dict(name=a.get("name"), bodyid=b.id, type=int(GeomType[st.upper()]), size=size, pos=pos, quat=quat,
                         group=int(a.get("group", 0)), rgba=rgba)
                )
            for a in b.cams:
                cams.append(dict(name=a.get("name"), bodyid=b.id, pos=_f(a.get("pos"), 3, np.zeros(3)), quat=self._orient(a),
                                 mode=a.get("mode", "fixed"), target=a.get("target"), fovy=float(a.get("fovy", 45))))
            for a in b.lights:
                lights.append(dict(name=a.get("name"), bodyid=b.id, pos=_f(a.get("pos"), 3, np.zeros(3)),
                                   dir=_normalize(_f(a.get("dir"), 3, np.array([0.0, 0, -1]))), mode=a.get("mode", "fixed"),
                                   target=a.get("target"), directional=_bool(a.get("directional"), False)))
        nsite, ncam, nlight = len(sites), len(cams), len(lights)

        name2id = {
            "body": {b.name: b.id for b in bodies if b.name},
            "joint": {n: i for i, n in enumerate(jnt_names) if n},
            "geom": {g["name"]: i for i, g in enumerate(geoms) if g["name"]},
            "site": {s["name"]: i for i, s in enumerate(sites) if s["name"]},
            "mesh": dict(meshes["index"]),
        }

        # ---------------- option
        opt = MjOptionLite()
        o = root.find("option")
        oa = o.attrib if o is not None else {}
        opt.timestep = float(oa.get("timestep", 0.002))
        opt.impratio = float(oa.get("impratio", 1))
        opt.tolerance = float(oa.get("tolerance", 1e-8))
        opt.ls_tolerance = float(oa.get("ls_tolerance", 0.01))
        opt.noslip_tolerance = float(oa.get("noslip_tolerance", 1e-6))
        opt.ccd_tolerance = float(oa.get("ccd_tolerance", 1e-6))
        opt.gravity = _f(oa.get("gravity"), 3, np.array([0.0, 0, -9.81]))
        opt.wind = _f(oa.get("wind"), 3, np.zeros(3))
        opt.magnetic = _f(oa.get("magnetic"), 3, np.array([0.0, -0.5, 0]))
        opt.density = float(oa.get("density", 0))
        opt.viscosity = float(oa.get("viscosity", 0))
        opt.o_margin = float(oa.get("o_margin", 0))
        opt.o_solref = _f(oa.get("o_solref"), 2, np.array([0.02, 1.0]))
        opt.o_solimp = _f(oa.get("o_solimp"), 5, np.array([0.9, 0.95, 0.001, 0.5, 2.0]))
        opt.o_friction = _f(oa.get("o_friction"), 5, np.array([1.0, 1.0, 0.005, 0.0001, 0.0001]))
        opt.integrator = {"euler": 0, "rk4": 1, "implicit": 2, "implicitfast": 3}[oa.get("integrator", "Euler").lower()]
        opt.cone = {"pyramidal": 0, "elliptic": 1}[oa.get("cone", "pyramidal").lower()]
        opt.jacobian = {"dense": 0, "sparse": 1, "auto": 2}[oa.get("jacobian", "auto").lower()]
        opt.solver = {"pgs": 0, "cg": 1, "newton": 2}[oa.get("solver", "Newton").lower()]
        opt.iterations = int(oa.get("iterations", 100))
        opt.ls_iterations = int(oa.get("ls_iterations", 50))
        opt.noslip_iterations = int(oa.get("noslip_iterations", 0))
        opt.ccd_iterations = int(oa.get("ccd_iterations", 50))
        opt.sdf_iterations = int(oa.get("sdf_iterations", 10))
        opt.sdf_initpoints = int(oa.get("sdf_initpoints", 40))
        disable, enable = 0, 0
        if o is not None and o.find("flag") is not None:
            fmap = {
                "constraint": DisableBit.CONSTRAINT, "equality": DisableBit.EQUALITY, "frictionloss": DisableBit.FRICTIONLOSS,
                "limit": DisableBit.LIMIT, "contact": DisableBit.CONTACT, "spring": DisableBit.SPRING, "damper": DisableBit.DAMPER,
                "gravity": DisableBit.GRAVITY, "clampctrl": DisableBit.CLAMPCTRL, "warmstart": DisableBit.WARMSTART,
                "filterparent": DisableBit.FILTERPARENT, "actuation": DisableBit.ACTUATION, "refsafe": DisableBit.REFSAFE,
                "sensor": DisableBit.SENSOR, "midphase": DisableBit.MIDPHASE, "eulerdamp": DisableBit.EULERDAMP,
                "autoreset": DisableBit.AUTORESET, "nativeccd": DisableBit.NATIVECCD, "island": DisableBit.ISLAND,
            }
            emap = {"override": EnableBit.OVERRIDE, "energy": EnableBit.ENERGY, "fwdinv": EnableBit.FWDINV,
                    "invdiscrete": EnableBit.INVDISCRETE, "multiccd": EnableBit.MULTICCD, "sleep": EnableBit.SLEEP}
            for k, v in o.find("flag").attrib.items():
                if k == "passive":
                    if v == "disable":
                        disable |= int(DisableBit.SPRING | DisableBit.DAMPER)
                elif k in fmap:
                    if v == "disable":
                        disable |= int(fmap[k])
                elif k in emap:
                    if v == "enable":
                        enable |= int(emap[k])
        opt.disableflags = disable
        opt.enableflags = enable
        opt.disableactuator = 0

        # ---------------- contact pairs / excludes
        pairs, excludes = [], []
        ct = root.find("contact")
        if ct is not None:
            for pe in ct.findall("pair"):
                a = self.defaults.resolve("pair", pe, None)
                g1, g2 = name2id["geom"][a["geom1"]], name2id["geom"][a["geom2"]]
                if geoms[g1]["bodyid"] > geoms[g2]["bodyid"]:
                    g1, g2 = g2, g1
                G1, G2 = geoms[g1], geoms[g2]
                p = dict(geom1=g1, geom2=g2, signature=(G1["bodyid"] << 16) + G2["bodyid"])
                # unspecified attributes are inferred from the two geoms (MuJoCo mjCPair::Compile)
                if G1["priority"] != G2["priority"]:
                    gp = G1 if G1["priority"] > G2["priority"] else G2
                    mixfri = gp["friction"]
                    mixref, miximp = gp["solref"], gp["solimp"]
                    p["dim"] = gp["condim"]
                else:
                    p["dim"] = max(G1["condim"], G2["condim"])
                    mixfri = np.maximum(G1["friction"], G2["friction"])
                    s1, s2 = G1["solmix"], G2["solmix"]
                    if s1 >= mjMINVAL and s2 >= mjMINVAL:
                        mix = s1 / (s1 + s2)
                    elif s1 < mjMINVAL and s2 < mjMINVAL:
                        mix = 0.5
                    elif s1 < mjMINVAL:
                        mix = 0.0
                    else:
                        mix = 1.0
                    if G1["solref"][0] > 0 and G2["solref"][0] > 0:
                        mixref = mix * G1["solref"] + (1 - mix) * G2["solref"]
                    else:
                        mixref = np.minimum(G1["solref"], G2["solref"])
                    miximp = mix * G1["solimp"] + (1 - mix) * G2["solimp"]
                if "condim" in a:
                    p["dim"] = int(a["condim"])
                if "friction" in a:
                    fr = np.array([1.0, 1.0, 0.005, 0.0001, 0.0001])
                    f = _f(a["friction"])
                    fr[: f.size] = f
                    if f.size < 2:
                        fr[1] = fr[0]
                    p["friction"] = fr
                else:
                    p["friction"] = np.array([mixfri[0], mixfri[0], mixfri[1], mixfri[2], mixfri[2]])
                p["solref"] = _f(a.get("solref"), 2, mixref)
                p["solreffriction"] = _f(a.get("solreffriction"), 2, np.zeros(2))
                si = np.array(miximp, dtype=np.float64).copy()
                if "solimp" in a:
                    si = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                    s_ = _f(a["solimp"])
                    si[: s_.size] = s_
                p["solimp"] = si
                p["margin"] = float(a["margin"]) if "margin" in a else max(G1["margin"], G2["margin"])
                p["gap"] = float(a["gap"]) if "gap" in a else max(G1["gap"], G2["gap"])
                p["name"] = a.get("name")
                pairs.append(p)
            for ee in ct.findall("exclude"):
                b1, b2 = name2id["body"][ee.get("body1")], name2id["body"][ee.get("body2")]
                if b1 > b2:
                    b1, b2 = b2, b1
                excludes.append((b1 << 16) + b2)
        pairs.sort(key=lambda p: p["signature"])  # MuJoCo orders pairs by body signature (stable)
        excludes.sort()
        npair = len(pairs)

        # ---------------- equality
        eqs = []
        eq = root.find("equality")
        if eq is not None:
            for ee in eq:
                a = self.defaults.resolve("equality", ee, None)
                data = np.zeros(mjNEQDATA)
                e = dict(active=1 if _bool(a.get("active"), True) else 0,
                         solref=_f(a.get("solref"), 2, np.array([0.02, 1.0])), name=a.get("name"))
                si = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                s_ = _f(a.get("solimp"), None, np.zeros(0))
                si[: s_.size] = s_
                e["solimp"] = si
                if ee.tag in ("connect", "weld") and "site1" in a:
                    # Site semantic (MuJoCo >= 3.1): the objects are sites, eq_objtype = mjOBJ_SITE and the anchors / relative pose
                    # come from the site frames at run time, so eq_data carries only the torque scale.  NOTE: the reference ignores
                    # eq_objtype and indexes *bodies* with these ids (device.py:306-307, constraint.py:116-212); device_put keeps
                    # that behaviour for parity and refuses ids that are not valid body ids.
                    if "site2" not in a or "body1" in a or "body2" in a:
                        raise ValueError(f"equality {ee.tag}: site1 and site2 must be given together, without body1 / body2")
                    e.update(type=int(EqType.CONNECT if ee.tag == "connect" else EqType.WELD), obj1=name2id["site"][a["site1"]],
                             obj2=name2id["site"][a["site2"]], objtype=int(ObjType.SITE))
                    if ee.tag == "weld":
                        data[10] = float(a.get("torquescale", 1))
                elif ee.tag in ("connect", "weld"):
                    b1 = name2id["body"][a["body1"]]
                    b2 = name2id["body"][a["body2"]] if "body2" in a else 0
                    e.update(type=int(EqType.CONNECT if ee.tag == "connect" else EqType.WELD), obj1=b1, obj2=b2,
                             objtype=int(ObjType.BODY))
                    e["_spec"] = (ee.tag, a)
                elif ee.tag == "joint":
                    j1 = name2id["joint"][a["joint1"]]
                    j2 = name2id["joint"][a["joint2"]] if "joint2" in a else -1
                    pc = _f(a.get("polycoef"), None, np.array([0.0, 1, 0, 0, 0]))
                    data[: pc.size] = pc
                    e.update(type=int(EqType.JOINT), obj1=j1, obj2=j2, objtype=int(ObjType.JOINT))
                else:
                    raise NotImplementedError(f"equality {ee.tag}")
                e["data"] = data
                eqs.append(e)
        neq = len(eqs)

        # ---------------- tendons (fixed only)
        tendons, wraps = [], []
        tn = root.find("tendon")
        if tn is not None:
            for te in tn:
                if te.tag != "fixed":
                    raise NotImplementedError("only fixed tendons are supported")
                a = self.defaults.resolve("tendon", te, None)
                t = dict(name=a.get("name"), adr=len(wraps), num=0)
                for je in te.findall("joint"):
                    wraps.append((int(WrapType.JOINT), name2id["joint"][je.get("joint")], float(je.get("coef"))))
                    t["num"] += 1
                rng = _f(a.get("range"), 2, np.zeros(2))
                lim = _bool(a.get("limited"), None)
                if lim is None:
                    lim = "range" in a and (rng[0] != 0 or rng[1] != 0)
                t.update(limited=1 if lim else 0, range=rng, margin=float(a.get("margin", 0)),
                         stiffness=float(a.get("stiffness", 0)), damping=float(a.get("damping", 0)),
                         armature=float(a.get("armature", 0)), frictionloss=float(a.get("frictionloss", 0)),
                         solref_lim=_f(a.get("solreflimit"), 2, np.array([0.02, 1.0])),
                         solref_fri=_f(a.get("solreffriction"), 2, np.array([0.02, 1.0])))
                for key, attr in (("solimp_lim", "solimplimit"), ("solimp_fri", "solimpfriction")):
                    si = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                    s_ = _f(a.get(attr), None, np.zeros(0))
                    si[: s_.size] = s_
                    t[key] = si
                sl = _f(a.get("springlength"), None, np.array([-1.0]))
                t["springlength"] = np.array([sl[0], sl[0] if sl.size == 1 else sl[1]])
                tendons.append(t)
        ntendon, nwrap = len(tendons), len(wraps)
        name2id["tendon"] = {t["name"]: i for i, t in enumerate(tendons) if t["name"]}

        # ---------------- actuators
        acts = []
        an = root.find("actuator")
        if an is not None:
            for ae in an:
                a = self.defaults.resolve(ae.tag, ae, None)
                gear = np.zeros(6)
                g = _f(a.get("gear"), None, np.array([1.0]))
                gear[: g.size] = g
                act = dict(name=a.get("name"), gear=gear, dyntype=0, gaintype=0, biastype=0,
                           dynprm=np.zeros(mjNDYN), gainprm=np.zeros(mjNGAIN), biasprm=np.zeros(mjNBIAS))
                act["dynprm"][0] = 1.0
                act["gainprm"][0] = 1.0
                if "joint" in a:
                    act.update(trntype=int(TrnType.JOINT), trnid=name2id["joint"][a["joint"]])
                elif "jointinparent" in a:
                    act.update(trntype=int(TrnType.JOINTINPARENT), trnid=name2id["joint"][a["jointinparent"]])
                elif "tendon" in a:
                    act.update(trntype=int(TrnType.TENDON), trnid=name2id["tendon"][a["tendon"]])
                else:
                    raise NotImplementedError("actuator transmission")
                if ae.tag == "general":
                    act["dyntype"] = {"none": 0, "integrator": 1, "filter": 2, "filterexact": 3, "muscle": 4}[a.get("dyntype", "none")]
                    act["gaintype"] = {"fixed": 0, "affine": 1, "muscle": 2}[a.get("gaintype", "fixed")]
                    act["biastype"] = {"none": 0, "affine": 1, "muscle": 2}[a.get("biastype", "none")]
                    for key in ("dynprm", "gainprm", "biasprm"):
                        if key in a:
                            v = _f(a[key])
                            act[key][:] = 0
                            act[key][: v.size] = v
                elif ae.tag == "position":
                    kp = float(a.get("kp", 1))
                    kv = float(a.get("kv", 0))
                    act["gainprm"][0] = kp
                    act["biastype"] = 1
                    act["biasprm"][1], act["biasprm"][2] = -kp, -kv
                elif ae.tag == "velocity":
                    kv = float(a.get("kv", 1))
                    act["gainprm"][0] = kv
                    act["biastype"] = 1
                    act["biasprm"][2] = -kv
                elif ae.tag == "muscle":
                    # MuJoCo's <muscle> shortcut: muscle dynamics, gain and bias sharing one parameter vector
                    # [range0 range1 force scale lmin lmax vmax fpmax fvmax]; dynprm = [tau_act tau_deact tausmooth]
                    tc = _f(a.get("timeconst"), 2, np.array([0.01, 0.04]))
                    act["dyntype"], act["gaintype"], act["biastype"] = 4, 2, 2
                    act["dynprm"][:] = 0
                    act["dynprm"][:3] = [tc[0], tc[1], float(a.get("tausmooth", 0))]
                    rg = _f(a.get("range"), 2, np.array([0.75, 1.05]))
                    prm = [rg[0], rg[1], float(a.get("force", -1)), float(a.get("scale", 200)), float(a.get("lmin", 0.5)),
                           float(a.get("lmax", 1.6)), float(a.get("vmax", 1.5)), float(a.get("fpmax", 1.3)), float(a.get("fvmax", 1.2))]
                    for key in ("gainprm", "biasprm"):
                        act[key][:] = 0
                        act[key][:9] = prm
                elif ae.tag != "motor":
                    raise NotImplementedError(f"actuator {ae.tag}")
                # MuJoCo computes missing length ranges of muscles by simulation (mj_setLengthRange); only explicit ones here
                act["lengthrange"] = _f(a.get("lengthrange"), 2, np.zeros(2))
                if (act["gaintype"] == 2 or act["biastype"] == 2) and "lengthrange" not in a:
                    raise NotImplementedError("muscle actuators need an explicit lengthrange (automatic length-range computation is not implemented)")
                for key, rk in (("ctrl", "ctrlrange"), ("force", "forcerange"), ("act", "actrange")):
                    rng = _f(a.get(rk), 2, np.zeros(2))
                    lim = _bool(a.get(key + "limited"), None)
                    if lim is None:
                        lim = rk in a and (rng[0] != 0 or rng[1] != 0)
                    act[key + "range"], act[key + "limited"] = rng, 1 if lim else 0
                acts.append(act)
        nu = len(acts)
        na = 0
        for act in acts:
            if act["dyntype"] != 0:
                act["actadr"], act["actnum"] = na, 1
                na += 1
            else:
                act["actadr"], act["actnum"] = -1, 0

        # ---------------- sensors
        sens = []
        sn = root.find("sensor")
        nsensordata = 0
        if sn is not None:
            spec = {
                "touch": (SensorType.TOUCH, 1, Stage.ACC, DataType.POSITIVE),
                "accelerometer": (SensorType.ACCELEROMETER, 3, Stage.ACC, DataType.REAL),
                "velocimeter": (SensorType.VELOCIMETER, 3, Stage.VEL, DataType.REAL),
                "gyro": (SensorType.GYRO, 3, Stage.VEL, DataType.REAL),
                "force": (SensorType.FORCE, 3, Stage.ACC, DataType.REAL),
                "torque": (SensorType.TORQUE, 3, Stage.ACC, DataType.REAL),
                "magnetometer": (SensorType.MAGNETOMETER, 3, Stage.POS, DataType.REAL),
                "rangefinder": (SensorType.RANGEFINDER, 1, Stage.POS, DataType.REAL),
                "jointpos": (SensorType.JOINTPOS, 1, Stage.POS, DataType.REAL),
                "jointvel": (SensorType.JOINTVEL, 1, Stage.VEL, DataType.REAL),
                "actuatorpos": (SensorType.ACTUATORPOS, 1, Stage.POS, DataType.REAL),
                "actuatorvel": (SensorType.ACTUATORVEL, 1, Stage.VEL, DataType.REAL),
                "actuatorfrc": (SensorType.ACTUATORFRC, 1, Stage.ACC, DataType.REAL),
                "jointactuatorfrc": (SensorType.JOINTACTFRC, 1, Stage.ACC, DataType.REAL),
                "tendonpos": (SensorType.TENDONPOS, 1, Stage.POS, DataType.REAL),
                "tendonvel": (SensorType.TENDONVEL, 1, Stage.VEL, DataType.REAL),
                "ballquat": (SensorType.BALLQUAT, 4, Stage.POS, DataType.QUATERNION),
                "ballangvel": (SensorType.BALLANGVEL, 3, Stage.VEL, DataType.REAL),
                "framepos": (SensorType.FRAMEPOS, 3, Stage.POS, DataType.REAL),
                "framequat": (SensorType.FRAMEQUAT, 4, Stage.POS, DataType.QUATERNION),
                "framexaxis": (SensorType.FRAMEXAXIS, 3, Stage.POS, DataType.AXIS),
                "frameyaxis": (SensorType.FRAMEYAXIS, 3, Stage.POS, DataType.AXIS),
                "framezaxis": (SensorType.FRAMEZAXIS, 3, Stage.POS, DataType.AXIS),
                "framelinvel": (SensorType.FRAMELINVEL, 3, Stage.VEL, DataType.REAL),
                "frameangvel": (SensorType.FRAMEANGVEL, 3, Stage.VEL, DataType.REAL),
                "subtreecom": (SensorType.SUBTREECOM, 3, Stage.POS, DataType.REAL),
                "subtreelinvel": (SensorType.SUBTREELINVEL, 3, Stage.VEL, DataType.REAL),
                "subtreeangmom": (SensorType.SUBTREEANGMOM, 3, Stage.VEL, DataType.REAL),
                "framelinacc": (SensorType.FRAMELINACC, 3, Stage.ACC, DataType.REAL),
                "frameangacc": (SensorType.FRAMEANGACC, 3, Stage.ACC, DataType.REAL),
                "tendonactuatorfrc": (SensorType.TENDONACTFRC, 1, Stage.ACC, DataType.REAL),
                "clock": (SensorType.CLOCK, 1, Stage.POS, DataType.REAL),
            }
            objtypes = {"body": ObjType.BODY, "xbody": ObjType.XBODY, "geom": ObjType.GEOM, "site": ObjType.SITE, "camera": ObjType.CAMERA}
            kind_of = {ObjType.BODY: "body", ObjType.XBODY: "body", ObjType.GEOM: "geom", ObjType.SITE: "site", ObjType.CAMERA: "camera"}
            for se in sn:
                if se.tag not in spec:
                    raise NotImplementedError(f"sensor {se.tag}")
                st, dim, stage, dt = spec[se.tag]
                rt, rid = ObjType.UNKNOWN, -1
                if se.tag.startswith("frame"):
                    ot = objtypes[se.get("objtype")]
                    oid = name2id[kind_of[ot]][se.get("objname")]
                    if "refname" in se.attrib:
                        rt = objtypes[se.get("reftype")]
                        rid = name2id[kind_of[rt]][se.get("refname")]
                elif se.tag == "clock":
                    ot, oid = ObjType.UNKNOWN, -1
                elif "tendon" in se.attrib:
                    ot, oid = ObjType.TENDON, [x["name"] for x in tendons].index(se.get("tendon"))
                elif "body" in se.attrib:
                    ot, oid = ObjType.BODY, name2id["body"][se.get("body")]
                elif "site" in se.attrib:
                    ot, oid = ObjType.SITE, name2id["site"][se.get("site")]
                elif "joint" in se.attrib:
                    ot, oid = ObjType.JOINT, name2id["joint"][se.get("joint")]
                elif "actuator" in se.attrib:
                    ot, oid = ObjType.UNKNOWN, [x["name"] for x in acts].index(se.get("actuator"))
                else:
                    raise NotImplementedError("sensor object")
                sens.append(dict(name=se.get("name"), type=int(st), dim=dim, needstage=int(stage), datatype=int(dt),
                                 objtype=int(ot), objid=oid, reftype=int(rt), refid=rid, adr=nsensordata,
                                 cutoff=float(se.get("cutoff", 0))))
                nsensordata += dim
        nsensor = len(sens)

        # ---------------- numerics (custom)
        numerics = []
        cu = root.find("custom")
        if cu is not None:
            for ne_ in cu.findall("numeric"):
                numerics.append((ne_.get("name"), _f(ne_.get("data"), None, np.zeros(1))))

        # ---------------- assemble
        def arr(lst, shape_tail=(), dtype=np.float64):
            if len(lst) == 0:
                return np.zeros((0,) + shape_tail, dtype=dtype)
            return np.asarray(lst, dtype=dtype).reshape((len(lst),) + shape_tail)

        m.nq, m.nv, m.nu, m.na, m.nbody, m.njnt, m.ngeom = nq, nv, nu, na, nbody, njnt, ngeom
        m.nsite, m.ncam, m.nlight, m.nmesh, m.npair, m.nexclude = nsite, ncam, nlight, len(meshes["list"]), npair, len(excludes)
        m.neq, m.ntendon, m.nwrap, m.nsensor, m.nnumeric, m.nmocap, m.nM = neq, ntendon, nwrap, nsensor, len(numerics), nmocap, nM
        m.nsensordata = nsensordata
        m.nhfield, m.ntex, m.ntexdata, m.nkey = len(meshes["hfields"]), 0, 0, 0
        hfs = meshes["hfields"]
        m.hfield_size = np.array([h["size"] for h in hfs], dtype=np.float64).reshape(len(hfs), 4)
        m.hfield_nrow = np.array([h["nrow"] for h in hfs], dtype=np.int32)
        m.hfield_ncol = np.array([h["ncol"] for h in hfs], dtype=np.int32)
        m.hfield_adr = np.array(np.concatenate([[0], np.cumsum([h["nrow"] * h["ncol"] for h in hfs])[:-1]]) if hfs else [], dtype=np.int32)
        m.hfield_data = np.concatenate([h["data"] for h in hfs]) if hfs else np.zeros(0)
        m.opt = opt
        m.qpos0 = arr(qpos0)
        m.qpos_spring = arr(qpos_spring)
        m.body_parentid, m.body_rootid, m.body_weldid, m.body_mocapid = body_parentid, body_rootid, body_weldid, body_mocapid
        m.body_jntnum, m.body_jntadr, m.body_dofnum, m.body_dofadr = body_jntnum, body_jntadr, body_dofnum, body_dofadr
        m.body_geomnum, m.body_geomadr = body_geomnum, body_geomadr
        m.body_treeid = np.array([-1] + [int(np.unique(body_rootid[1:]).tolist().index(r)) for r in body_rootid[1:]], np.int32)
        m.body_sameframe = np.zeros(nbody, np.uint8)
        m.body_simple = np.zeros(nbody, np.uint8)
        m.body_pos, m.body_quat, m.body_ipos, m.body_iquat = body_pos, body_quat, body_ipos, body_iquat
        m.body_mass, m.body_subtreemass, m.body_inertia = body_mass, body_subtreemass, body_inertia
        m.body_gravcomp = np.array([b.gravcomp for b in bodies])
        m.jnt_type = arr(jnt_type, dtype=np.int32)
        m.jnt_qposadr, m.jnt_dofadr, m.jnt_bodyid = arr(jnt_qposadr, dtype=np.int32), arr(jnt_dofadr, dtype=np.int32), arr(jnt_bodyid, dtype=np.int32)
        m.jnt_group = arr(jnt_group, dtype=np.int32)
        m.jnt_limited = arr(jnt_limited, dtype=np.uint8)
        m.jnt_actfrclimited = arr(jnt_actfrclimited, dtype=np.uint8)
        m.jnt_actgravcomp = arr(jnt_actgravcomp, dtype=np.uint8)
        m.jnt_solref, m.jnt_solimp = arr(jnt_solref, (2,)), arr(jnt_solimp, (5,))
        m.jnt_pos, m.jnt_axis = arr(jnt_pos, (3,)), arr(jnt_axis, (3,))
        m.jnt_stiffness, m.jnt_range, m.jnt_margin = arr(jnt_stiffness), arr(jnt_range, (2,)), arr(jnt_margin)
        m.jnt_actfrcrange = arr(jnt_actfrcrange, (2,))
        m.dof_bodyid, m.dof_jntid, m.dof_parentid = arr(dof_bodyid, dtype=np.int32), arr(dof_jntid, dtype=np.int32), arr(dof_parentid, dtype=np.int32)
        m.dof_treeid = m.body_treeid[m.dof_bodyid] if nv else np.zeros(0, np.int32)
        m.dof_Madr = dof_Madr
        m.dof_simplenum = np.zeros(nv, np.int32)
        m.dof_solref, m.dof_solimp = arr(dof_solref, (2,)), arr(dof_solimp, (5,))
        m.dof_frictionloss, m.dof_armature, m.dof_damping = arr(dof_frictionloss), arr(dof_armature), arr(dof_damping)

        def garr(key, tail=(), dtype=np.float64):
            return arr([g[key] for g in geoms], tail, dtype)

        m.geom_type = garr("type", dtype=np.int32)
        m.geom_contype, m.geom_conaffinity = garr("contype", dtype=np.int32), garr("conaffinity", dtype=np.int32)
        m.geom_condim, m.geom_bodyid = garr("condim", dtype=np.int32), garr("bodyid", dtype=np.int32)
        m.geom_dataid, m.geom_group, m.geom_priority = garr("dataid", dtype=np.int32), garr("group", dtype=np.int32), garr("priority", dtype=np.int32)
        m.geom_sameframe = np.zeros(ngeom, np.uint8)
        m.geom_matid = garr("matid", dtype=np.int32)
        m.nmat = len(self._materials["rgba"])
        m.mat_rgba = arr(self._materials["rgba"], (4,))
        m.geom_solmix, m.geom_solref, m.geom_solimp = garr("solmix"), garr("solref", (2,)), garr("solimp", (5,))
        m.geom_size, m.geom_rbound = garr("size", (3,)), garr("rbound")
        m.geom_aabb = np.zeros((ngeom, 6))
        m.geom_pos, m.geom_quat, m.geom_friction = garr("pos", (3,)), garr("quat", (4,)), garr("friction", (3,))
        m.geom_margin, m.geom_gap, m.geom_rgba = garr("margin"), garr("gap"), garr("rgba", (4,))
        m.geom_fluid = np.zeros((ngeom, 12))
        m.site_type = arr([s["type"] for s in sites], dtype=np.int32)
        m.site_bodyid = arr([s["bodyid"] for s in sites], dtype=np.int32)
        m.site_group = arr([s["group"] for s in sites], dtype=np.int32)
        m.site_sameframe = np.zeros(nsite, np.uint8)
        m.site_size, m.site_pos, m.site_quat = arr([s["size"] for s in sites], (3,)), arr([s["pos"] for s in sites], (3,)), arr([s["quat"] for s in sites], (4,))
        m.site_rgba = arr([s["rgba"] for s in sites], (4,))
        modes = {"fixed": 0, "track": 1, "trackcom": 2, "targetbody": 3, "targetbodycom": 4}
        m.cam_mode = arr([modes[c["mode"]] for c in cams], dtype=np.int32)
        m.cam_bodyid = arr([c["bodyid"] for c in cams], dtype=np.int32)
        m.cam_targetbodyid = arr([name2id["body"].get(c["target"], -1) if c["target"] else -1 for c in cams], dtype=np.int32)
        m.cam_pos, m.cam_quat = arr([c["pos"] for c in cams], (3,)), arr([c["quat"] for c in cams], (4,))
        m.cam_fovy = arr([c["fovy"] for c in cams])
        m.light_mode = arr([modes[l["mode"]] for l in lights], dtype=np.int32)
        m.light_bodyid = arr([l["bodyid"] for l in lights], dtype=np.int32)
        m.light_targetbodyid = arr([name2id["body"].get(l["target"], -1) if l["target"] else -1 for l in lights], dtype=np.int32)
        m.light_pos, m.light_dir = arr([l["pos"] for l in lights], (3,)), arr([l["dir"] for l in lights], (3,))
        m.light_directional = arr([1 if l["directional"] else 0 for l in lights], dtype=np.uint8)
        # meshes
        vadr, fadr, mv, mf = [], [], [], []
        nvert = nface = 0
        for mm in meshes["list"]:
            vadr.append(nvert)
            fadr.append(nface)
            mv.append(mm["vert"])
            mf.append(mm["face"])
            nvert += len(mm["vert"])
            nface += len(mm["face"])
        m.mesh_vertadr, m.mesh_vertnum = arr(vadr, dtype=np.int32), arr([len(x) for x in mv], dtype=np.int32)
        m.mesh_faceadr, m.mesh_facenum = arr(fadr, dtype=np.int32), arr([len(x) for x in mf], dtype=np.int32)
        m.mesh_vert = np.concatenate(mv) if mv else np.zeros((0, 3))
        m.mesh_face = np.concatenate(mf) if mf else np.zeros((0, 3), np.int32)
        m.mesh_pos = arr([mm["pos"] for mm in meshes["list"]], (3,))
        m.mesh_quat = arr([mm["quat"] for mm in meshes["list"]], (4,))
        # pairs
        m.pair_dim = arr([p["dim"] for p in pairs], dtype=np.int32)
        m.pair_geom1, m.pair_geom2 = arr([p["geom1"] for p in pairs], dtype=np.int32), arr([p["geom2"] for p in pairs], dtype=np.int32)
        m.pair_signature = arr([p["signature"] for p in pairs], dtype=np.int32)
        m.pair_solref, m.pair_solreffriction = arr([p["solref"] for p in pairs], (2,)), arr([p["solreffriction"] for p in pairs], (2,))
        m.pair_solimp, m.pair_friction = arr([p["solimp"] for p in pairs], (5,)), arr([p["friction"] for p in pairs], (5,))
        m.pair_margin, m.pair_gap = arr([p["margin"] for p in pairs]), arr([p["gap"] for p in pairs])
        m.exclude_signature = arr(excludes, dtype=np.int32)
        # equality
        m.eq_type = arr([e["type"] for e in eqs], dtype=np.int32)
        m.eq_obj1id, m.eq_obj2id = arr([e["obj1"] for e in eqs], dtype=np.int32), arr([e["obj2"] for e in eqs], dtype=np.int32)
        m.eq_objtype = arr([e["objtype"] for e in eqs], dtype=np.int32)
        m.eq_active0 = arr([e["active"] for e in eqs], dtype=np.uint8)
        m.eq_solref, m.eq_solimp = arr([e["solref"] for e in eqs], (2,)), arr([e["solimp"] for e in eqs], (5,))
        m.eq_data = arr([e["data"] for e in eqs], (mjNEQDATA,))
        # tendons
        m.tendon_adr, m.tendon_num = arr([t["adr"] for t in tendons], dtype=np.int32), arr([t["num"] for t in tendons], dtype=np.int32)
        m.tendon_limited = arr([t["limited"] for t in tendons], dtype=np.uint8)
        m.tendon_actfrclimited = np.zeros(ntendon, np.uint8)
        m.tendon_actfrcrange = np.zeros((ntendon, 2))
        m.tendon_solref_lim, m.tendon_solimp_lim = arr([t["solref_lim"] for t in tendons], (2,)), arr([t["solimp_lim"] for t in tendons], (5,))
        m.tendon_solref_fri, m.tendon_solimp_fri = arr([t["solref_fri"] for t in tendons], (2,)), arr([t["solimp_fri"] for t in tendons], (5,))
        m.tendon_range, m.tendon_margin = arr([t["range"] for t in tendons], (2,)), arr([t["margin"] for t in tendons])
        m.tendon_stiffness, m.tendon_damping = arr([t["stiffness"] for t in tendons]), arr([t["damping"] for t in tendons])
        m.tendon_armature, m.tendon_frictionloss = arr([t["armature"] for t in tendons]), arr([t["frictionloss"] for t in tendons])
        m.tendon_lengthspring = arr([t["springlength"] for t in tendons], (2,))
        m.wrap_type, m.wrap_objid = arr([w[0] for w in wraps], dtype=np.int32), arr([w[1] for w in wraps], dtype=np.int32)
        m.wrap_prm = arr([w[2] for w in wraps])
        # actuators
        m.actuator_trntype = arr([a["trntype"] for a in acts], dtype=np.int32)
        m.actuator_dyntype = arr([a["dyntype"] for a in acts], dtype=np.int32)
        m.actuator_gaintype = arr([a["gaintype"] for a in acts], dtype=np.int32)
        m.actuator_biastype = arr([a["biastype"] for a in acts], dtype=np.int32)
        m.actuator_trnid = arr([[a["trnid"], -1] for a in acts], (2,), np.int32)
        m.actuator_actadr, m.actuator_actnum = arr([a["actadr"] for a in acts], dtype=np.int32), arr([a["actnum"] for a in acts], dtype=np.int32)
        m.actuator_group = np.zeros(nu, np.int32)
        m.actuator_ctrllimited = arr([a["ctrllimited"] for a in acts], dtype=np.uint8)
        m.actuator_forcelimited = arr([a["forcelimited"] for a in acts], dtype=np.uint8)
        m.actuator_actlimited = arr([a["actlimited"] for a in acts], dtype=np.uint8)
        m.actuator_dynprm, m.actuator_gainprm = arr([a["dynprm"] for a in acts], (mjNDYN,)), arr([a["gainprm"] for a in acts], (mjNGAIN,))
        m.actuator_biasprm = arr([a["biasprm"] for a in acts], (mjNBIAS,))
        m.actuator_actearly = np.zeros(nu, np.uint8)
        m.actuator_ctrlrange, m.actuator_forcerange = arr([a["ctrlrange"] for a in acts], (2,)), arr([a["forcerange"] for a in acts], (2,))
        m.actuator_actrange, m.actuator_gear = arr([a["actrange"] for a in acts], (2,)), arr([a["gear"] for a in acts], (6,))
        m.actuator_cranklength = np.zeros(nu)
        m.actuator_lengthrange = arr([a["lengthrange"] for a in acts], (2,))
        # sensors
        for key, dt in (("type", np.int32), ("datatype", np.int32), ("needstage", np.int32), ("objtype", np.int32),
                        ("objid", np.int32), ("dim", np.int32), ("adr", np.int32), ("cutoff", np.float64)):
            setattr(m, "sensor_" + key, arr([s[key] for s in sens], dtype=dt))
        m.sensor_reftype = np.array([x["reftype"] for x in sens], np.int32).reshape(nsensor)
        m.sensor_refid = np.array([x["refid"] for x in sens], np.int32).reshape(nsensor)
        m.sensor_intprm = np.zeros((nsensor, 3), np.int32)
        # numerics + names
        nadr, ndata, cur = [], [], 0
        for _, dat in numerics:
            nadr.append(cur)
            ndata.extend(dat.tolist())
            cur += dat.size
        m.numeric_adr, m.numeric_data = arr(nadr, dtype=np.int32), arr(ndata)
        m.numeric_size = arr([d.size for _, d in numerics], dtype=np.int32)
        names = bytearray()

        def nameadr(lst):
            out = []
            for n in lst:
                out.append(len(names))
                names.extend((n or "").encode() + b"\x00")
            return arr(out, dtype=np.int32)

        m.name_bodyadr = nameadr([b.name for b in bodies])
        m.name_jntadr = nameadr(jnt_names)
        m.name_geomadr = nameadr([g["name"] for g in geoms])
        m.name_siteadr = nameadr([s["name"] for s in sites])
        m.name_camadr = nameadr([c["name"] for c in cams])
        m.name_meshadr = nameadr(meshes["names"])
        m.name_pairadr = nameadr([p["name"] for p in pairs])
        m.name_eqadr = nameadr([e["name"] for e in eqs])
        m.name_tendonadr = nameadr([t["name"] for t in tendons])
        m.name_actuatoradr = nameadr([a["name"] for a in acts])
        m.name_sensoradr = nameadr([s["name"] for s in sens])
        m.name_numericadr = nameadr([n for n, _ in numerics])
        m.names = bytes(names)
        m.name2id = name2id
        m.stat = MjStatisticLite(meaninertia=1.0, meanmass=float(body_mass[1:].mean()) if nbody > 1 else 0.0,
                                 meansize=0.0, extent=1.0, center=np.zeros(3))
        m._eq_specs = [e.get("_spec") for e in eqs]
        set_const(m)
        del m.__dict__["_eq_specs"]
        return m


# ----------------------------------------------------------------------------- mj_setConst


def _kin_at_qpos0(m):
    """Body poses at qpos0 (host float64; mirrors MuJoCo mj_kinematics for the reference config)."""
    nb = m.nbody
    xpos, xquat = np.zeros((nb, 3)), np.zeros((nb, 4))
    xquat[0, 0] = 1
    xanchor, xaxis = np.zeros((m.njnt, 3)), np.zeros((m.njnt, 3))
    for b in range(1, nb):
        p = m.body_parentid[b]
        pos = xpos[p] + rotate(m.body_pos[b], xquat[p])
        quat = quat_mul(xquat[p], m.body_quat[b])
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            jt, qa = m.jnt_type[j], m.jnt_qposadr[j]
            if jt == JointType.FREE:
                xanchor[j], xaxis[j] = m.qpos0[qa : qa + 3], [0, 0, 1]
                pos, quat = m.qpos0[qa : qa + 3].copy(), _normalize(m.qpos0[qa + 3 : qa + 7])
            else:
                xanchor[j] = rotate(m.jnt_pos[j], quat) + pos
                xaxis[j] = rotate(m.jnt_axis[j], quat)
                if jt == JointType.BALL:
                    quat = quat_mul(quat, _normalize(m.qpos0[qa : qa + 4]))
                    pos = xanchor[j] - rotate(m.jnt_pos[j], quat)
                # hinge/slide at qpos0: zero displacement
        xpos[b], xquat[b] = pos, _normalize(quat)
    return xpos, xquat, xanchor, xaxis


def set_const(m: MjModelLite) -> None:
    """Quantities MuJoCo's mj_setConst derives at qpos0."""
    nb, nv = m.nbody, m.nv
    xpos, xquat, xanchor, xaxis = _kin_at_qpos0(m)
    xipos = np.stack([xpos[b] + rotate(m.body_ipos[b], xquat[b]) for b in range(nb)])
    ximat = np.stack([quat_to_mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(nb)])

    # weld/connect equality data that MuJoCo fills in at compile time (relative pose at qpos0)
    for i, spec in enumerate(getattr(m, "_eq_specs", [])):
        if spec is None:
            continue
        tag, a = spec
        b1, b2 = int(m.eq_obj1id[i]), int(m.eq_obj2id[i])
        R1, R2 = quat_to_mat(xquat[b1]), quat_to_mat(xquat[b2])
        data = m.eq_data[i]
        if tag == "connect":
            anchor = _f(a.get("anchor"), 3, np.zeros(3))
            data[0:3] = anchor
            data[3:6] = R2.T @ (xpos[b1] + R1 @ anchor - xpos[b2])
        else:
            anchor = _f(a.get("anchor"), 3, np.zeros(3))
            data[0:3] = anchor  # anchor on body2
            if "relpose" in a and np.any(_f(a["relpose"], 7)[3:] != 0):
                rp = _f(a["relpose"], 7)
                data[3:6] = rp[:3]
                data[6:10] = _normalize(rp[3:])
            else:
                data[3:6] = R1.T @ (xpos[b2] + R2 @ anchor - xpos[b1])
                data[6:10] = quat_mul(quat_conj(xquat[b1]), xquat[b2])
            data[10] = float(a.get("torquescale", 1))

    # per-dof motion axes about world origin, spatial [ang, lin]
    S = np.zeros((nv, 6))
    for j in range(m.njnt):
        jt, da, b = m.jnt_type[j], m.jnt_dofadr[j], m.jnt_bodyid[j]
        R = quat_to_mat(xquat[b])
        if jt == JointType.FREE:
            S[da : da + 3, 3:] = np.eye(3)
            for k in range(3):
                ax = R[:, k]
                S[da + 3 + k, :3] = ax
                S[da + 3 + k, 3:] = np.cross(xpos[b], ax)
        elif jt == JointType.BALL:
            for k in range(3):
                ax = R[:, k]
                S[da + k, :3] = ax
                S[da + k, 3:] = np.cross(xanchor[j], ax)
        elif jt == JointType.HINGE:
            S[da, :3] = xaxis[j]
            S[da, 3:] = np.cross(xanchor[j], xaxis[j])
        else:
            S[da, 3:] = xaxis[j]

    def body_jac(b, point):
        """6 x nv Jacobian [lin; ang] of ``point`` rigidly attached to body b."""
        J = np.zeros((6, nv))
        d = -1
        bb = b
        while bb > 0 and m.body_dofnum[bb] == 0:
            bb = m.body_parentid[bb]
        if bb > 0:
            d = m.body_dofadr[bb] + m.body_dofnum[bb] - 1
        while d >= 0:
            w, v = S[d, :3], S[d, 3:]
            J[3:, d] = w
            J[:3, d] = v + np.cross(w, point)
            d = m.dof_parentid[d]
        return J

    M = np.zeros((nv, nv))
    for b in range(1, nb):
        Jb = body_jac(b, xipos[b])
        Ib = ximat[b] @ np.diag(m.body_inertia[b]) @ ximat[b].T
        M += m.body_mass[b] * Jb[:3].T @ Jb[:3] + Jb[3:].T @ Ib @ Jb[3:]
    M += np.diag(m.dof_armature)
    m.dof_M0 = np.diag(M).copy()
    Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))

    dof_inv = np.diag(Minv).copy() if nv else np.zeros(0)
    for j in range(m.njnt):
        jt, da = m.jnt_type[j], m.jnt_dofadr[j]
        if jt == JointType.FREE:
            dof_inv[da : da + 3] = dof_inv[da : da + 3].mean()
            dof_inv[da + 3 : da + 6] = dof_inv[da + 3 : da + 6].mean()
        elif jt == JointType.BALL:
            dof_inv[da : da + 3] = dof_inv[da : da + 3].mean()
    m.dof_invweight0 = dof_inv
    biw = np.zeros((nb, 2))
    for b in range(1, nb):
        if m.body_weldid[b] == 0:
            continue
        Jb = body_jac(b, xipos[b])
        A = Jb @ Minv @ Jb.T
        biw[b, 0] = max(mjMINVAL, np.trace(A[:3, :3]) / 3)
        biw[b, 1] = max(mjMINVAL, np.trace(A[3:, 3:]) / 3)
    m.body_invweight0 = biw

    # tendons
    nt = m.ntendon
    tJ = np.zeros((nt, nv))
    tl0 = np.zeros(nt)
    for t in range(nt):
        for w in range(m.tendon_adr[t], m.tendon_adr[t] + m.tendon_num[t]):
            j = m.wrap_objid[w]
            tJ[t, m.jnt_dofadr[j]] = m.wrap_prm[w]
            tl0[t] += m.wrap_prm[w] * m.qpos0[m.jnt_qposadr[j]]
    m.tendon_length0 = tl0
    m.tendon_invweight0 = np.array([max(mjMINVAL, tJ[t] @ Minv @ tJ[t]) for t in range(nt)])
    if nt:
        ls = m.tendon_lengthspring
        unset = ls[:, 0] == -1
        ls[unset] = tl0[unset, None]

    # actuators
    acc0 = np.zeros(m.nu)
    for i in range(m.nu):
        mom = np.zeros(nv)
        tt, tid = m.actuator_trntype[i], m.actuator_trnid[i, 0]
        gear = m.actuator_gear[i]
        if tt == TrnType.TENDON:
            mom = tJ[tid] * gear[0]
        else:
            jt, da = m.jnt_type[tid], m.jnt_dofadr[tid]
            if jt == JointType.FREE:
                mom[da : da + 6] = gear
            elif jt == JointType.BALL:
                mom[da : da + 3] = gear[:3]
            else:
                mom[da] = gear[0]
        acc0[i] = np.linalg.norm(Minv @ mom)
    m.actuator_acc0 = acc0

    m.stat.meaninertia = float(np.mean(np.diag(M))) if nv else 1.0
    # camera / light reference poses
    sub_mass = m.body_subtreemass
    sub_com = xipos * m.body_mass[:, None]
    for i in range(nb - 1, 0, -1):
        sub_com[m.body_parentid[i]] += sub_com[i]
    sub_com = np.where(sub_mass[:, None] > 0, sub_com / np.maximum(sub_mass[:, None], mjMINVAL), xipos)
    cam_xpos = np.array([xpos[b] + rotate(p, xquat[b]) for b, p in zip(m.cam_bodyid, m.cam_pos)]).reshape(-1, 3)
    cam_xmat = np.array([quat_to_mat(quat_mul(xquat[b], q)) for b, q in zip(m.cam_bodyid, m.cam_quat)]).reshape(-1, 3, 3)
    m.cam_pos0 = np.array([cx - xpos[b] for cx, b in zip(cam_xpos, m.cam_bodyid)]).reshape(-1, 3)
    m.cam_poscom0 = np.array([cx - sub_com[b] for cx, b in zip(cam_xpos, m.cam_bodyid)]).reshape(-1, 3)
    m.cam_mat0 = cam_xmat.reshape(-1, 9)
    l_xpos = np.array([xpos[b] + rotate(p, xquat[b]) for b, p in zip(m.light_bodyid, m.light_pos)]).reshape(-1, 3)
    m.light_pos0 = np.array([lx - xpos[b] for lx, b in zip(l_xpos, m.light_bodyid)]).reshape(-1, 3)
    m.light_poscom0 = np.array([lx - sub_com[b] for lx, b in zip(l_xpos, m.light_bodyid)]).reshape(-1, 3)
    m.light_dir0 = np.array([rotate(d_, xquat[b]) for b, d_ in zip(m.light_bodyid, m.light_dir)]).reshape(-1, 3)


# ----------------------------------------------------------------------------- entry points


def from_xml_string(xml: str, basedir: str = ".") -> MjModelLite:
    # MuJoCo's parser (tinyxml2) accepts "--" inside comments (ASCII-art tables in the reference's satellite_small.xml); expat
    # does not.  Comments carry no model content, so they are dropped before parsing.
    root = ET.fromstring(re.sub(r"<!--.*?-->", "", xml, flags=re.S))
    if root.tag != "mujoco":
        raise ValueError("root element must be <mujoco>")
    return Compiler(root, basedir).compile()


def from_xml_path(path: str) -> MjModelLite:
    with open(path) as fh:
        xml = fh.read()
    return from_xml_string(xml, os.path.dirname(os.path.abspath(path)))


def load(path_or_xml: str) -> MjModelLite:
    """``mjcf.load(path_or_xml_string)``."""
    if path_or_xml.lstrip().startswith("<"):
        return from_xml_string(path_or_xml)
    return from_xml_path(path_or_xml)


def patch_zoo_ant(xml: str) -> str:
    """Textual patches the reference's RL layer applies to ant.xml before compiling
    (reference zoo/base.py:236-260 and zoo/ant.py:38-53): fixed camera/light, floor plane,
    free root joint, timestep 0.01."""
    xml = re.sub(r"<camera\b[^/]*/>\s*", "", xml)
    xml = re.sub(r"<light\b[^/]*/>\s*", "", xml)
    camera = '<camera name="side" pos="0 -4 3" xyaxes="1 0 0 0 0.45 1" fovy="60"/>'
    light = '<light name="top" pos="0 0 4" dir="0 0 -1" diffuse="0.8 0.8 0.8" ambient="0.3 0.3 0.3" directional="true"/>'
    floor = ""
    if not re.search(r'<geom\b[^>]*type="plane"', xml):
        floor = '\n  <geom name="floor" type="plane" size="10 10 0.1" rgba="0.8 0.85 0.8 1" conaffinity="1" condim="3"/>'
    xml = xml.replace("<worldbody>", f"<worldbody>\n  {camera}\n  {light}{floor}")
    xml = re.sub(r'(<body\s+name="torso"[^>]*>)', r"\1\n      <freejoint name='root'/>", xml, count=1)
    xml = re.sub(r"(<compiler\b[^/]*/>\s*)", r'\1<option timestep="0.01"/>\n  ', xml, count=1)
    return xml
