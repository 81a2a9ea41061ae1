"""ORACLE (test infrastructure): convex geom preprocessing (reference mesh.py:43-110,165-218,392-447)."""

from __future__ import annotations

PLANE, HFIELD, SPHERE, CAPSULE, ELLIPSOID, CYLINDER, BOX, MESH = range(8)


def build_convex(mj, dtype):
    out = []
    for g in range(mj.ngeom):
        t = int(mj.geom_type[g])
        if t in (BOX, MESH):
            from . import convex_prep

            out.append(convex_prep.geom_convex(mj, g, dtype))
        else:
            out.append(None)
    return out
