"""trimesh stand-in (TEST INFRASTRUCTURE): Trimesh + convex hull + coplanar facets on scipy's Qhull."""
import numpy as np
from scipy.spatial import ConvexHull


class Trimesh:
    def __init__(self, vertices=None, faces=None, process=True):
        v = np.asarray(vertices, dtype=np.float64)
        f = np.asarray(faces, dtype=np.int64)
        if f.ndim == 2 and f.shape[1] == 4:  # quads -> two triangles
            f = np.concatenate([f[:, [0, 1, 2]], f[:, [0, 2, 3]]])
        self.vertices, self.faces = v, f
        self._facets = None

    def _compute_facets(self):
        v, f = self.vertices, self.faces
        n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        d = np.einsum("ij,ij->i", n, v[f[:, 0]])
        groups = {}
        for i, key in enumerate(map(tuple, np.round(np.c_[n, d], 6))):
            groups.setdefault(key, []).append(i)
        facets, normals = [], []
        for key, idx in groups.items():
            if len(idx) > 1:
                facets.append(np.array(idx))
                normals.append(np.array(key[:3]))
        self._facets, self._facets_normal = facets, np.array(normals).reshape(-1, 3)

    @property
    def facets(self):
        if self._facets is None:
            self._compute_facets()
        return self._facets

    @property
    def facets_normal(self):
        if self._facets is None:
            self._compute_facets()
        return self._facets_normal


class convex:  # noqa: N801
    @staticmethod
    def convex_hull(mesh):
        pts = np.asarray(mesh.vertices if isinstance(mesh, Trimesh) else mesh, dtype=np.float64)
        hull = ConvexHull(pts)
        used = np.sort(np.unique(hull.simplices))
        remap = -np.ones(len(pts), dtype=np.int64)
        remap[used] = np.arange(len(used))
        v = pts[used]
        faces = remap[hull.simplices]
        c = v.mean(0)
        for i, tri in enumerate(faces):
            a, b, cc = v[tri]
            if np.dot(np.cross(b - a, cc - a), a - c) < 0:
                faces[i] = tri[[0, 2, 1]]
        return Trimesh(vertices=v, faces=faces)
