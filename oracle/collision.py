"""ORACLE (test infrastructure): narrowphase + collision driver, batched over envs.

Follows reference ``collision_primitive.py`` (plane_sphere :41, plane_capsule :48-74, sphere_sphere
:172-192, sphere_capsule :195-201, capsule_capsule :204-221) and ``collision_driver.collision`` :800-875.
Each narrowphase returns (dist (B,n), pos (B,n,3), frame (B,n,3,3)).
"""

import torch

from . import omath as M
from .model import BOX, CAPSULE, HFIELD, MESH, PLANE, SPHERE


def _plane_sphere(n, ppos, spos, r):
    dist = M.dot(spos - ppos, n) - r
    pos = spos - n * (r + 0.5 * dist).unsqueeze(-1)
    return dist, pos


def plane_sphere(m, d, g1, g2):
    n = d.geom_xmat[:, g1, :, 2]
    dist, pos = _plane_sphere(n, d.geom_xpos[:, g1], d.geom_xpos[:, g2], m.geom_size[g2, 0])
    return dist[:, None], pos[:, None], M.make_frame(n)[:, None]


def plane_capsule(m, d, g1, g2):
    n, axis = d.geom_xmat[:, g1, :, 2], d.geom_xmat[:, g2, :, 2]
    dt = n.dtype
    b, b_norm = M.normalize_with_norm(axis - n * M.dot(n, axis).unsqueeze(-1))
    y = torch.tensor([0.0, 1.0, 0.0], dtype=dt).expand_as(n)
    z = torch.tensor([0.0, 0.0, 1.0], dtype=dt).expand_as(n)
    inner = torch.where(((-0.5 < n[:, 1]) & (n[:, 1] < 0.5)).unsqueeze(-1), y, z)
    b = torch.where((b_norm < 0.5).unsqueeze(-1), inner, b)
    frame = torch.stack([n, b, M.cross(n, b)], dim=-2)
    segment = axis * m.geom_size[g2, 1]
    ds, ps = [], []
    for off in (segment, -segment):
        dist, pos = _plane_sphere(n, d.geom_xpos[:, g1], d.geom_xpos[:, g2] + off, m.geom_size[g2, 0])
        ds.append(dist)
        ps.append(pos)
    return torch.stack(ds, 1), torch.stack(ps, 1), torch.stack([frame, frame], 1)


def _sphere_sphere(pos1, r1, pos2, r2):
    n, dist = M.normalize_with_norm(pos2 - pos1)
    x = torch.tensor([1.0, 0.0, 0.0], dtype=n.dtype).expand_as(n)
    n = torch.where((dist == 0.0).unsqueeze(-1), x, n)
    dist = dist - (r1 + r2)
    pos = pos1 + n * (r1 + dist * 0.5).unsqueeze(-1)
    return dist, pos, n


def sphere_sphere(m, d, g1, g2):
    dist, pos, n = _sphere_sphere(d.geom_xpos[:, g1], m.geom_size[g1, 0], d.geom_xpos[:, g2], m.geom_size[g2, 0])
    return dist[:, None], pos[:, None], M.make_frame(n)[:, None]


def sphere_capsule(m, d, g1, g2):
    axis, length = d.geom_xmat[:, g2, :, 2], m.geom_size[g2, 1]
    segment = axis * length
    cpos = d.geom_xpos[:, g2]
    pt = M.closest_segment_point(cpos - segment, cpos + segment, d.geom_xpos[:, g1])
    dist, pos, n = _sphere_sphere(d.geom_xpos[:, g1], m.geom_size[g1, 0], pt, m.geom_size[g2, 0])
    return dist[:, None], pos[:, None], M.make_frame(n)[:, None]


def capsule_capsule(m, d, g1, g2):
    axis1, axis2 = d.geom_xmat[:, g1, :, 2], d.geom_xmat[:, g2, :, 2]
    seg1, seg2 = axis1 * m.geom_size[g1, 1], axis2 * m.geom_size[g2, 1]
    p1, p2 = d.geom_xpos[:, g1], d.geom_xpos[:, g2]
    pt1, pt2 = M.closest_segment_to_segment_points(p1 - seg1, p1 + seg1, p2 - seg2, p2 + seg2)
    dist, pos, n = _sphere_sphere(pt1, m.geom_size[g1, 0], pt2, m.geom_size[g2, 0])
    return dist[:, None], pos[:, None], M.make_frame(n)[:, None]


def _fn(t):
    table = {
        (PLANE, SPHERE): plane_sphere,
        (PLANE, CAPSULE): plane_capsule,
        (SPHERE, SPHERE): sphere_sphere,
        (SPHERE, CAPSULE): sphere_capsule,
        (CAPSULE, CAPSULE): capsule_capsule,
    }
    if t in table:
        return table[t]
    if t[0] == HFIELD:
        from . import hfield

        return hfield.hfield_pair
    from . import convex

    return {
        (PLANE, BOX): convex.plane_convex, (PLANE, MESH): convex.plane_convex,
        (SPHERE, BOX): convex.sphere_convex, (SPHERE, MESH): convex.sphere_convex,
        (CAPSULE, BOX): convex.capsule_convex, (CAPSULE, MESH): convex.capsule_convex,
        (BOX, BOX): convex.convex_convex, (BOX, MESH): convex.convex_convex, (MESH, MESH): convex.convex_convex,
    }[t]


def collision(m, d):
    """collision_driver.collision :800-875: per group narrowphase, concat, permute by the static sort."""
    if m.ncon == 0:
        B, dt = d.qpos.shape[0], d.qpos.dtype
        d.contact_dist = torch.zeros((B, 0), dtype=dt)
        d.contact_pos = torch.zeros((B, 0, 3), dtype=dt)
        d.contact_frame = torch.zeros((B, 0, 3, 3), dtype=dt)
        return d
    dists, poss, frames = [], [], []
    for t, cands, _n in m.groups:
        fn = _fn(t)
        for c in cands:
            dist, pos, frame = fn(m, d, c.geom1, c.geom2)
            dists.append(dist)
            poss.append(pos)
            frames.append(frame)
    perm = m.contact_perm
    d.contact_dist = torch.cat(dists, 1)[:, perm]
    d.contact_pos = torch.cat(poss, 1)[:, perm]
    d.contact_frame = torch.cat(frames, 1)[:, perm]
    if getattr(m, "topk", 0):
        # collision_driver.collision :822-840: `_, idx = torch.topk(-contact.dist, k=max_cp)`, then argsort by condim of the survivors
        _, idx = torch.topk(-d.contact_dist, k=m.topk, dim=1)
        idx = idx[:, m.topk_perm]
        take = lambda x: torch.gather(x, 1, idx.reshape(idx.shape + (1,) * (x.dim() - 2)).expand(idx.shape + x.shape[2:]))
        d.contact_dist, d.contact_pos, d.contact_frame = take(d.contact_dist), take(d.contact_pos), take(d.contact_frame)
        d.contact_src = idx
    return d
