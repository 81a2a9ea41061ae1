"""Host-side convex tables for box / mesh geoms: vertices, polygon faces (CCW, padded), face normals, edges.

Same role as the reference's ``mesh.get`` (mesh.py:406-447): per-geom hull description consumed by the
convex narrowphase.  Built once at ``device_put``; stored flat in the ModelBlob.
"""

from __future__ import annotations

import numpy as np

from .mjmodel import GeomType

MAX_FACE_VERTS = 20

_BOX_CORNERS = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
_BOX_FACES = np.array([[0, 4, 5, 1], [0, 2, 6, 4], [6, 7, 5, 4], [2, 3, 7, 6], [1, 5, 7, 3], [0, 1, 3, 2]], dtype=np.int64)


class ConvexTable:
    __slots__ = ("vert", "face", "norm", "edge")

    def __init__(self, vert, face):
        self.vert = np.ascontiguousarray(vert, dtype=np.float64)
        self.face = np.ascontiguousarray(face, dtype=np.int64)
        p = self.vert[self.face]
        n = np.cross(p[:, 1] - p[:, 0], p[:, -1] - p[:, 0])
        self.norm = n / np.linalg.norm(n, axis=1, keepdims=True)
        self.edge = self._edges()

    def _edges(self):
        """Undirected edges shared by >= 2 faces, in first-seen order (face-major, then corner)."""
        seen: dict = {}
        nf, k = self.face.shape
        for f in range(nf):
            for c in range(k):
                a, b = int(self.face[f, c]), int(self.face[f, c - 1])
                if a == b:
                    continue
                key = (min(a, b), max(a, b))
                seen[key] = seen.get(key, 0) + 1
        return np.array([e for e, cnt in seen.items() if cnt >= 2], dtype=np.int64).reshape(-1, 2)

    @property
    def key(self):
        return (self.face.shape, self.vert.shape, self.edge.shape)


def _ccw_polygon(pts, idx, normal):
    """Order coplanar vertex ids counter-clockwise about ``normal`` (2D hull in the dominant plane)."""
    from scipy.spatial import ConvexHull

    ax = int(np.argmax(np.abs(normal)))
    keep = [i for i in range(3) if i != ax]
    hull = ConvexHull(pts[:, keep])
    step = 1 if normal[ax] > 0 else -1
    if ax == 1:
        step = -step
    return idx[hull.vertices[::step]]


def _plane_key(eq, scale):
    """Coplanarity bucket of a hull facet: unit normal to 4 decimals, offset to 1e-4 of the mesh size (meshes come
    from float32 STL files, so exactly coplanar triangles differ at the 1e-7 level)."""
    return tuple(np.round(eq[:3], 4) + 0.0) + (float(np.round(eq[3] / scale, 4)) + 0.0,)


def _mesh_table(vert):
    from scipy.spatial import ConvexHull

    hull = ConvexHull(vert)
    scale = float(np.abs(vert).max()) or 1.0
    used = np.unique(hull.simplices)
    local = np.full(len(vert), -1, dtype=np.int64)
    local[used] = np.arange(len(used))
    v = vert[used]
    planes: dict = {}
    for tri, eq in zip(hull.simplices, hull.equations):
        planes.setdefault(_plane_key(eq, scale), []).append(local[tri])
    polys = []
    for eq in sorted(planes):
        n = np.asarray(eq[:3])
        ids = np.unique(np.concatenate(planes[eq]))
        if len(ids) == 3:
            a, b, c = v[ids]
            poly = ids if np.dot(np.cross(b - a, c - a), n) >= 0 else ids[[0, 2, 1]]
        else:
            poly = _ccw_polygon(v[ids], ids, n)
            if len(poly) > MAX_FACE_VERTS:
                poly = poly[:: len(poly) // MAX_FACE_VERTS + 1]
        polys.append(poly)
    width = max(len(p) for p in polys)
    face = np.stack([np.concatenate([p, np.repeat(p[-1], width - len(p))]) for p in polys])
    return ConvexTable(v, face)


def build(mj):
    """Per-geom ConvexTable (or None for non-convex geoms)."""
    cache: dict = {}
    out = []
    for g in range(mj.ngeom):
        t = int(mj.geom_type[g])
        if t == GeomType.BOX:
            out.append(ConvexTable(_BOX_CORNERS * np.asarray(mj.geom_size[g], dtype=np.float64), _BOX_FACES))
        elif t == GeomType.MESH:
            mid = int(mj.geom_dataid[g])
            if mid not in cache:
                a, n = int(mj.mesh_vertadr[mid]), int(mj.mesh_vertnum[mid])
                cache[mid] = _mesh_table(np.asarray(mj.mesh_vert[a : a + n], dtype=np.float64))
            out.append(cache[mid])
        else:
            out.append(None)
    return out
