"""ORACLE (test infrastructure): convex geom tables (verts, polygon faces, face normals, unique edges).

Restates reference ``mesh.py``: box corner/face tables (:165-218, :406-430), face normals
(_get_face_norm :43-54), unique edges (_get_edge_normals :57-94, _get_unique_edges :359-362),
coplanar-facet merge into CCW polygons (_convex_hull_2d :97-108, _merge_coplanar :111-162,
``_MAX_HULL_FACE_VERTICES = 20`` :32).  The reference builds hulls with ``trimesh`` (absent here);
``scipy.spatial.ConvexHull`` (same Qhull) is used instead, so face/vertex *ordering* is this
module's own deterministic choice -- PARITY UNPINNED for ordering, geometry is identical.
"""

from __future__ import annotations

import itertools

import numpy as np
import torch

BOX, MESH = 6, 7
_MAX_HULL_FACE_VERTICES = 20


class ConvexGeom:
    def __init__(self, vert, face, dtype):
        self.vert_np, self.face_np = vert, face
        fv = vert[face]
        e0 = fv[:, 1] - fv[:, 0]
        e1 = fv[:, -1] - fv[:, 0]
        n = np.cross(e0, e1)
        self.facenorm_np = n / np.linalg.norm(n, axis=1, keepdims=True)
        self.edge_np = _unique_edges(face, self.facenorm_np)
        self.vert = torch.as_tensor(vert).to(dtype)
        self.face = torch.as_tensor(face, dtype=torch.long)
        self.facenorm = torch.as_tensor(self.facenorm_np).to(dtype)
        self.edge = torch.as_tensor(self.edge_np, dtype=torch.long)
        self.shape_key = (tuple(face.shape), tuple(vert.shape), tuple(self.edge_np.shape))


def _unique_edges(face, face_norm):
    r_face = np.roll(face, 1, axis=1)
    fe = np.stack([face, r_face], axis=-1)
    fe.sort(axis=2)
    flat = fe.reshape(-1, 2)
    owner = np.repeat(np.arange(face.shape[0]), face.shape[1])
    emap: dict = {}
    for i in range(flat.shape[0]):
        a, b = int(flat[i, 0]), int(flat[i, 1])
        if a == b:
            continue
        emap.setdefault((a, b), []).append(face_norm[owner[i]])
    edges = [k for k, v in emap.items() if len(v) >= 2]
    return np.array(edges, dtype=np.int64).reshape(-1, 2)


def _hull_2d(points, normal):
    from scipy.spatial import ConvexHull

    best_axis = int(np.abs(normal).argmax())
    axis = np.eye(3)[best_axis]
    keep = [i for i in range(3) if i != best_axis]
    c = ConvexHull(points[:, keep])
    order = 1 if axis.dot(normal) > 0 else -1
    order *= -1 if best_axis == 1 else 1
    return c.vertices[::order]


def _plane_key(eq, scale):
    return tuple(np.round(eq[:3], 4) + 0.0) + (float(np.round(eq[3] / scale, 4)) + 0.0,)


def _mesh_polygons(vert):
    from scipy.spatial import ConvexHull

    hull = ConvexHull(vert)
    scale = float(np.abs(vert).max()) or 1.0
    used = np.sort(np.unique(hull.simplices))
    remap = -np.ones(vert.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    v = vert[used]
    groups: dict = {}
    for simplex, eq in zip(hull.simplices, hull.equations):
        groups.setdefault(_plane_key(eq, scale), []).append(remap[simplex])
    faces = []
    for key in sorted(groups):
        normal = np.array(key[:3])
        idx = np.unique(np.concatenate(groups[key]))
        if idx.size == 3:
            a, b, c = v[idx]
            if np.dot(np.cross(b - a, c - a), normal) < 0:
                idx = idx[[0, 2, 1]]
            faces.append(idx)
            continue
        poly = idx[_hull_2d(v[idx], normal)]
        if poly.size > _MAX_HULL_FACE_VERTICES:
            poly = poly[:: poly.size // _MAX_HULL_FACE_VERTICES + 1]
        faces.append(poly)
    mx = max(f.size for f in faces)
    faces = np.stack([np.pad(f, (0, mx - f.size), "edge") for f in faces])
    return v, faces


_CACHE: dict = {}


def geom_convex(mj, g, dtype):
    t = int(mj.geom_type[g])
    if t == BOX:
        corners = np.array(list(itertools.product((-1, 1), (-1, 1), (-1, 1))), dtype=np.float64)
        vert = corners * np.asarray(mj.geom_size[g], dtype=np.float64)
        face = np.array([[0, 4, 5, 1], [0, 2, 6, 4], [6, 7, 5, 4], [2, 3, 7, 6], [1, 5, 7, 3], [0, 1, 3, 2]])
    elif t == MESH:
        mid = int(mj.geom_dataid[g])
        a = int(mj.mesh_vertadr[mid])
        key = ("mesh", mid)
        if key not in _CACHE:
            _CACHE[key] = _mesh_polygons(np.asarray(mj.mesh_vert[a : a + int(mj.mesh_vertnum[mid])], dtype=np.float64))
        vert, face = _CACHE[key]
    else:
        return None
    return ConvexGeom(vert, face, dtype)
