"""ORACLE (test infrastructure): height-field narrowphase.

Restates reference ``mujoco_torch/_src/collision_hfield.py`` (_sphere_prism :44-91, _capsule_prism :94-168, _convex_prism :171-230,
_build_prisms :233-294, _hfield_collision :297-335, _select_manifold :338-349, hfield_sphere / capsule / convex :352-390),
``mesh.hfield_prism`` (mesh.py:274-344), ``mesh.hfield`` (:347-364) and the driver ``collision_driver._collide_hfield_geoms``
:332-437 with ``_hfield_subgrid_size`` :318-329.

The reference walks the sub-grid in Python per environment (``float(tensor)`` decides which cells an object covers), so it has no
batched form.  The batched semantics used here -- and by the CUDA kernel -- are exactly "the reference, env by env": the sub-grid
SIZE is static (bounding radius and grid pitch), its ORIGIN (cmin, rmin) is per env.  This module therefore loops over envs and
prisms and evaluates every prism with tensors that carry a leading dim of 1, reusing the clip / manifold / SAT helpers of
``oracle.convex``.
"""

import math as _math

import numpy as np
import torch

from . import convex as C
from . import omath as M

_FACE = np.array([[0, 1, 2, 0], [3, 4, 5, 3], [0, 3, 5, 1], [0, 2, 4, 3], [2, 1, 5, 4]])   # mesh.hfield_prism: bottom, top, three sides
_EDGE = np.array([[0, 1], [1, 2], [0, 2], [3, 4], [3, 5], [4, 5], [0, 3], [1, 5], [2, 4]])
_MANIFOLD_TOL = 1e-3


class _Prism:
    """mesh.hfield_prism :274-344: vertices relative to their centroid (= the prism's position), identity orientation."""

    def __init__(self, vert):
        self.pos = vert.mean(0)
        self.vert = vert - self.pos
        self.face = self.vert[torch.as_tensor(_FACE)]                        # (5, 4, 3)
        e0 = self.face[:, 2] - self.face[:, 1]
        e1 = self.face[:, 0] - self.face[:, 1]
        self.facenorm = M.normalize(M.cross(e0, e1))                         # ccw winding, the last (repeated) corner unused
        self.edge = torch.as_tensor(_EDGE)


def subgrid_size(hsize, nrow, ncol, rbound):
    """collision_driver._hfield_subgrid_size :318-329."""
    xtick, ytick = 2.0 * float(hsize[0]) / (ncol - 1), 2.0 * float(hsize[1]) / (nrow - 1)
    return min(int(_math.ceil(2 * rbound / xtick)) + 1, ncol), min(int(_math.ceil(2 * rbound / ytick)) + 1, nrow)


def build_prisms(hsize, nrow, ncol, data, obj_pos, rbound, sub, dt):
    """_build_prisms :233-294.  data[col, row] in [0, 1]; two triangular prisms per grid cell."""
    xhalf, yhalf, zscale, zbase = (float(x) for x in hsize)
    dx, dy = 2.0 * xhalf / (ncol - 1), 2.0 * yhalf / (nrow - 1)
    xfrac = (float(obj_pos[0]) - rbound + xhalf) / (2 * xhalf) * (ncol - 1)
    yfrac = (float(obj_pos[1]) - rbound + yhalf) / (2 * yhalf) * (nrow - 1)
    cmin, rmin = _math.floor(xfrac), _math.floor(yfrac)
    pt = lambda col, row: torch.tensor([dx * col - xhalf, dy * row - yhalf, float(data[col, row]) * zscale], dtype=dt)
    prisms = []
    for r in range(sub[1]):
        for c in range(sub[0]):
            ri, ci = max(0, min(rmin + r, nrow - 2)), max(0, min(cmin + c, ncol - 2))
            p1, p2, p3 = pt(ci, ri), pt(ci + 1, ri + 1), pt(ci, ri + 1)
            for top in (torch.stack([p1, p2, p3]), torch.stack([p1, pt(ci + 1, ri), p2])):
                bottom = top.clone()
                bottom[:, 2] = -zbase
                prisms.append(_Prism(torch.cat([bottom[[0, 2, 1]], top])))
    return prisms


def _sphere_prism(spos, r, p):
    """:44-91 (sphere centre `spos` in the height field's frame)."""
    sp = spos - p.pos
    support = ((sp - p.facenorm * r - p.face[:, 0]) * p.facenorm).sum(-1)
    support = torch.where(support >= 0, torch.full_like(support, -1e12), support)
    best = int(support.argmax())
    face, normal = p.face[best], p.facenorm[best]
    pt = sp - ((sp - face[0]) * normal).sum() * normal
    e0, e1 = torch.roll(face, 1, dims=0), face
    en = M.cross(e1 - e0, normal.expand_as(e0))
    ed = ((pt - e0) * en).sum(-1)
    inside = bool((ed <= 0).all())
    ed = torch.where((en == 0).all(-1) | (ed < 0.0), torch.full_like(ed, 1e12), ed)
    idx = int(ed.argmin())
    if not inside:
        pt = M.closest_segment_point(e0[idx], e1[idx], pt)
    n, d = M.normalize_with_norm(pt - sp)
    spt = sp + n * r
    return (d - r).reshape(1), ((pt + spt) * 0.5 + p.pos).reshape(1, 3), n.reshape(1, 3)


def _capsule_prism(cpos, axis, r, length, p):
    """:94-168."""
    cp = cpos - p.pos
    seg = axis * length
    pts = torch.stack([cp - seg, cp + seg])
    sup = (((pts[None] - (p.facenorm * r)[:, None]) - p.face[:, 0][:, None]) * p.facenorm[:, None]).sum(-1).min(-1).values
    sup = torch.where(sup >= 0, torch.full_like(sup, -1e12), sup)
    best = int(sup.argmax())
    face, normal = p.face[best], p.facenorm[best]
    e0, e1 = torch.roll(face, 1, dims=0), face
    en = M.cross(e1 - e0, normal.expand_as(e0))
    clipped, mask = C._clip_edge_to_planes(pts[0][None], pts[1][None], e0[None], en[None])
    clipped, mask = clipped[0] - normal * r, mask[0]
    face_pts = clipped - ((clipped - face[0]) * normal).sum(-1, keepdim=True) * normal
    pos = (clipped + face_pts) * 0.5
    norm = normal.expand(2, 3).clone()
    pen = torch.where(mask, ((face_pts - clipped) * normal).sum(-1), torch.full_like(pos[:, 0], -1.0))
    ec, cc = M.closest_segment_to_segment_points(e0, e1, pts[0].expand_as(e0), pts[1].expand_as(e0))
    ei = int(((ec - cc) ** 2).sum(-1).argmin())
    edge_axis, edge_dist = M.normalize_with_norm(cc[ei] - ec[ei])
    edge_pos = (ec[ei] + (cc[ei] - edge_axis * r)) * 0.5
    edge_pen = r - edge_dist
    if bool(edge_pen > 0):
        pos, norm, pen = pos.clone(), norm.clone(), pen.clone()
        pos[0], norm[0], pen[0] = edge_pos, edge_axis, edge_pen
    return -pen, pos + p.pos, -norm


def _convex_prism(opos, omat, cv, p):
    """:171-230 (object pose in the height field's frame; `cv` the object's hull tables)."""
    f1, f2 = cv.face, torch.as_tensor(_FACE)
    s1, s2 = f1.shape[-1], f2.shape[-1]
    if s1 < s2:
        f1 = torch.cat([f1, f1[:, -1:].expand(-1, s2 - s1)], dim=1)
    elif s2 < s1:
        f2 = torch.cat([f2, f2[:, -1:].expand(-1, s1 - s2)], dim=1)
    dt = opos.dtype
    eye = torch.eye(3, dtype=dt)
    obj = dict(faces=cv.vert[f1], normals=cv.facenorm, vert=cv.vert, edge=cv.edge, pos=opos, mat=omat)
    pri = dict(faces=p.vert[f2], normals=p.facenorm, vert=p.vert, edge=p.edge, pos=p.pos, mat=eye)
    swapped = cv.vert.shape[0] > p.vert.shape[0]
    a, b = (pri, obj) if swapped else (obj, pri)
    to_pos = b["mat"].T @ (a["pos"] - b["pos"])
    to_mat = b["mat"].T @ a["mat"]
    xf = lambda x: to_pos + x @ to_mat.T
    faces1, normals1, verts1 = xf(a["faces"]), a["normals"] @ to_mat.T, xf(a["vert"])
    edges1, edges2 = verts1[a["edge"]], b["vert"][b["edge"]]
    dist, pos, normal = C._sat_hull_hull(faces1[None], b["faces"][None], verts1[None], b["vert"][None], normals1[None], b["normals"][None],
                                         edges1[None], edges2[None])
    pos = b["pos"] + pos[0] @ b["mat"].T
    normal = normal[0] @ b["mat"].T
    if swapped:
        normal = -normal
    return dist[0], pos, normal


def _select_manifold(dist, pos, n):
    """:338-349."""
    n_mean = n.mean(0)
    mask = dist < torch.minimum(torch.zeros_like(dist), dist.min() + _MANIFOLD_TOL)
    idx = C._manifold_points(pos[None], mask[None], n_mean[None])[0]
    dist, pos, n = dist[idx], pos[idx], n[idx]
    unique = torch.tril(idx == idx[:, None]).sum(dim=1) == 1
    return torch.where(unique, dist, torch.ones_like(dist)), pos, n


def _one_env(kind, hsize, nrow, ncol, data, hpos, hmat, opos, omat, osize, cv, rbound, sub):
    """_hfield_collision :297-335 + hfield_sphere / capsule / convex :352-390 for one env."""
    dt = opos.dtype
    obj_pos = hmat.T @ (opos - hpos)
    obj_mat = hmat.T @ omat
    ds, ps, ns, tops = [], [], [], []
    for p in build_prisms(hsize, nrow, ncol, data, obj_pos, rbound, sub, dt):
        if kind == "sphere":
            d_, p_, n_ = _sphere_prism(obj_pos, osize[0], p)
        elif kind == "capsule":
            d_, p_, n_ = _capsule_prism(obj_pos, obj_mat[:, 2], osize[0], osize[1], p)
        else:
            d_, p_, n_ = _convex_prism(obj_pos, obj_mat, cv, p)
        ds.append(d_)
        ps.append(p_)
        ns.append(n_)
        tops.append(p.facenorm[1].expand(d_.shape[0], 3))
    dist, pos, n, top = torch.cat(ds), torch.cat(ps), -torch.cat(ns), torch.cat(tops)
    n = torch.where((n[:, 2] < 1e-6).unsqueeze(-1), top, n)
    dist, pos, n = _select_manifold(dist, pos, n)
    pos = pos @ hmat.T + hpos
    n = n @ hmat.T
    return dist, pos, M.make_frame(n)


def hfield_pair(m, d, g1, g2):
    """collision_driver._collide_hfield_geoms :332-437 for one (hfield, object) pair, env by env -> dist (B,4), pos (B,4,3), frame (B,4,3,3)."""
    mj = m.mj
    hid = int(mj.geom_dataid[g1])
    nrow, ncol, adr = int(mj.hfield_nrow[hid]), int(mj.hfield_ncol[hid]), int(mj.hfield_adr[hid])
    data = np.asarray(mj.hfield_data[adr: adr + nrow * ncol], dtype=np.float64).reshape((ncol, nrow), order="F")   # mesh.hfield :347-364
    hsize = np.asarray(mj.hfield_size[hid], dtype=np.float64)
    rbound = float(mj.geom_rbound[g2])
    sub = subgrid_size(hsize, nrow, ncol, rbound)
    gt = int(mj.geom_type[g2])
    kind = {2: "sphere", 3: "capsule"}.get(gt, "convex")
    if kind == "sphere":
        rbound_obj = float(np.max(np.asarray(mj.geom_size[g2])))                       # hfield_sphere :355
    elif kind == "capsule":
        rbound_obj = float(mj.geom_size[g2][0] + mj.geom_size[g2][1])                  # hfield_capsule :368
    else:
        rbound_obj = float(np.max(np.asarray(mj.geom_size[g2])))                       # hfield_convex :381
    cv = m.convex[g2] if kind == "convex" else None
    outs = [_one_env(kind, hsize, nrow, ncol, data, d.geom_xpos[b, g1], d.geom_xmat[b, g1], d.geom_xpos[b, g2], d.geom_xmat[b, g2],
                     m.geom_size[g2], cv, rbound_obj, sub) for b in range(d.qpos.shape[0])]
    return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]), torch.stack([o[2] for o in outs])
