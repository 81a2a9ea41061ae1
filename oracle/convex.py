"""ORACLE (test infrastructure): convex (box / mesh) narrowphase, batched over envs.

Restates reference ``mujoco_torch/_src/collision_convex.py`` (non-``smooth_collisions`` branches):
plane_convex :604-623, sphere_convex :626-699, capsule_convex :702-802, convex_convex :805-856 with
_sat_hull_hull :464-601, _create_contact_manifold :395-449, _clip :330-392, _clip_edge_to_planes :265-327,
_manifold_points :183-235 and the closest-point helpers :39-175.
Every helper takes tensors with arbitrary leading batch dims; ``argmax/argmin`` keep torch's first-index tie rule.
"""

import torch

from . import omath as M


def _take(x, idx):
    """x (..., N, *tail), idx (..., K) -> (..., K, *tail)."""
    tail = x.shape[idx.dim():]
    ix = idx.reshape(idx.shape + (1,) * len(tail)).expand(idx.shape + tail)
    return torch.gather(x, idx.dim() - 1, ix)


def _sel(x, idx):
    """x (..., N, *tail), idx (...) -> (..., *tail)."""
    return _take(x, idx.unsqueeze(-1)).squeeze(idx.dim())


def _geom(m, d, g):
    cv = m.convex[g]
    return d.geom_xpos[:, g], d.geom_xmat[:, g], cv


def _to_world(pos, mat, x):
    """mat @ x + pos for x (..., K, 3)."""
    return pos.unsqueeze(-2) + torch.matmul(x, mat.transpose(-1, -2))


def _seg_plane(a, b, p0, n):
    """_closest_segment_point_plane :39-63 (broadcasting)."""
    dd = (p0 * n).sum(-1)
    denom = (n * (b - a)).sum(-1)
    t = (dd - (n * a).sum(-1)) / (denom + 1e-6 * (denom == 0.0))
    t = torch.clamp(t, 0, 1)
    return a + t.unsqueeze(-1) * (b - a)


def _manifold_points(poly, mask, norm):
    """:183-235.  poly (..., N, 3), mask (..., N), norm (..., 3) -> idx (..., 4)."""
    N = poly.shape[-2]
    dist_mask = torch.where(mask, torch.zeros_like(poly[..., 0]), torch.full_like(poly[..., 0], -1e6))
    a_idx = dist_mask.argmax(-1)
    a = _sel(poly, a_idx).unsqueeze(-2)
    b_idx = (((a - poly) ** 2).sum(-1) + dist_mask).argmax(-1)
    b = _sel(poly, b_idx).unsqueeze(-2)
    nb = norm.unsqueeze(-2)
    ab = M.cross(nb, a - b)
    ap = a - poly
    c_idx = (torch.abs((ap * ab).sum(-1)) + dist_mask).argmax(-1)
    c = _sel(poly, c_idx).unsqueeze(-2)
    ac = M.cross(nb, a - c)
    bc = M.cross(nb, b - c)
    bp = b - poly
    dist_bp = torch.abs((bp * bc).sum(-1)) + dist_mask
    dist_ap = torch.abs((ap * ac).sum(-1)) + dist_mask
    d_idx = torch.cat([dist_bp, dist_ap], dim=-1).argmax(-1) % N
    return torch.stack([a_idx, b_idx, c_idx, d_idx], dim=-1)


def plane_convex(m, d, g1, g2):
    ppos, pmat, _ = _geom(m, d, g1)
    cpos, cmat, cv = _geom(m, d, g2)
    B = ppos.shape[0]
    vert = cv.vert.expand(B, *cv.vert.shape)
    rt = cmat.transpose(-1, -2)
    plane_pos = torch.matmul(rt, (ppos - cpos).unsqueeze(-1)).squeeze(-1)
    n = torch.matmul(rt, pmat[:, :, 2].unsqueeze(-1)).squeeze(-1)
    support = ((plane_pos.unsqueeze(-2) - vert) * n.unsqueeze(-2)).sum(-1)
    idx = _manifold_points(vert, support > 0, n)
    pos = _to_world(cpos, cmat, _take(vert, idx))
    nw = pmat[:, :, 2]
    frame = M.make_frame(nw).unsqueeze(1).expand(B, 4, 3, 3)
    unique = torch.tril(idx.unsqueeze(-2) == idx.unsqueeze(-1)).sum(-1) == 1
    sup = _take(support, idx)
    dist = torch.where(unique, -sup, torch.ones_like(sup))
    return dist, pos, frame


def _best_face(cv, support, B):
    """face with maximal (negative) support, :644-657."""
    masked = torch.where(support >= 0, torch.full_like(support, -1e12), support)
    best = masked.argmax(-1)
    faces = cv.vert[cv.face].expand(B, *cv.face.shape, 3)
    normals = cv.facenorm.expand(B, *cv.facenorm.shape)
    return _sel(faces, best), _sel(normals, best)


def sphere_convex(m, d, g1, g2):
    spos, _, _ = _geom(m, d, g1)
    cpos, cmat, cv = _geom(m, d, g2)
    B = spos.shape[0]
    r = m.geom_size[g1, 0]
    rt = cmat.transpose(-1, -2)
    sp = torch.matmul(rt, (spos - cpos).unsqueeze(-1)).squeeze(-1)
    faces = cv.vert[cv.face]
    normals = cv.facenorm
    support = ((sp.unsqueeze(1) - normals * r - faces[:, 0]) * normals).sum(-1)  # (B, nface)
    face, normal = _best_face(cv, support, B)
    pt = sp - ((sp - face[:, 0]) * normal).sum(-1, keepdim=True) * normal
    e0 = torch.roll(face, 1, dims=1)
    e1 = face
    en = M.cross(e1 - e0, normal.unsqueeze(1).expand_as(e0))
    edge_dist = ((pt.unsqueeze(1) - e0) * en).sum(-1)
    inside = (edge_dist <= 0).all(-1)
    degenerate = (en == 0).all(-1)
    masked = torch.where(degenerate | (edge_dist < 0.0), torch.full_like(edge_dist, 1e12), edge_dist)
    idx = masked.argmin(-1)
    edge_pt = M.closest_segment_point(_sel(e0, idx), _sel(e1, idx), pt)
    pt = torch.where(inside.unsqueeze(-1), pt, edge_pt)
    n, dd = M.normalize_with_norm(pt - sp)
    spt = sp + n * r
    dist = dd - r
    pos = (pt + spt) * 0.5
    n = torch.matmul(cmat, n.unsqueeze(-1)).squeeze(-1)
    pos = torch.matmul(cmat, pos.unsqueeze(-1)).squeeze(-1) + cpos
    return dist[:, None], pos[:, None], M.make_frame(n)[:, None]


def _clip_edge_to_planes(p0, p1, plane_pts, plane_normals):
    """:265-327.  p0, p1 (..., 3); planes (..., P, 3) -> new_ps (..., 2, 3), mask (..., 2)."""
    a, b = p0.unsqueeze(-2), p1.unsqueeze(-2)
    p0_front = ((a - plane_pts) * plane_normals).sum(-1) > 1e-6
    p1_front = ((b - plane_pts) * plane_normals).sum(-1) > 1e-6
    cand = _seg_plane(a, b, plane_pts, plane_normals)

    def clip_point(q0, q1, front):
        pts = torch.where(front.unsqueeze(-1), cand, q0.unsqueeze(-2).expand_as(cand))
        dists = ((pts - q0.unsqueeze(-2)) * (q1 - q0).unsqueeze(-2)).sum(-1)
        return _sel(pts, dists.argmax(-1))

    new_p0 = clip_point(p0, p1, p0_front)
    new_p1 = clip_point(p1, p0, p1_front)
    clipped = torch.stack([new_p0, new_p1], dim=-2)
    mask = ~(p0_front & p1_front).any(-1)
    new_ps = torch.where(mask[..., None, None], clipped, torch.stack([p0, p1], dim=-2))
    crossing = ((p0 - p1) * (new_ps[..., 0, :] - new_ps[..., 1, :])).sum(-1) < 0
    mask = torch.where(crossing, torch.zeros_like(mask), mask)
    return new_ps, torch.stack([mask, mask], dim=-1)


def _project_poly_onto_poly_plane(poly1, norm1, poly2, norm2):
    """:249-257."""
    dd = (poly2[..., 0, :] * norm2).sum(-1)
    denom = (norm1 * norm2).sum(-1)
    t = (dd.unsqueeze(-1) - (poly1 * norm2.unsqueeze(-2)).sum(-1)) / (denom + 1e-6 * (denom == 0.0)).unsqueeze(-1)
    return poly1 + t.unsqueeze(-1) * norm1.unsqueeze(-2)


def _clip(clipping_poly, subject_poly, clipping_normal, subject_normal):
    """:330-392 -> points (..., 2M+2N, 3), mask (..., 2M+2N)."""
    c0 = torch.roll(clipping_poly, 1, dims=-2)
    c1 = clipping_poly
    cn = M.cross(c1 - c0, clipping_normal.unsqueeze(-2).expand_as(c0))
    s0 = torch.roll(subject_poly, 1, dims=-2)
    s1 = subject_poly
    sn = M.cross(s1 - s0, subject_normal.unsqueeze(-2).expand_as(s0))
    e0, m0 = _clip_edge_to_planes(s0, s1, c0.unsqueeze(-3), cn.unsqueeze(-3))
    c0s = _project_poly_onto_poly_plane(c0, clipping_normal, subject_poly, subject_normal)
    c1s = _project_poly_onto_poly_plane(c1, clipping_normal, subject_poly, subject_normal)
    e1, m1 = _clip_edge_to_planes(c0s, c1s, s0.unsqueeze(-3), sn.unsqueeze(-3))
    pts = torch.cat([e0, e1], dim=-3)
    msk = torch.cat([m0, m1], dim=-2)
    return pts.reshape(*pts.shape[:-3], -1, 3), msk.reshape(*msk.shape[:-2], -1)


def _create_contact_manifold(clipping_poly, subject_poly, clipping_norm, subject_norm, sep_axis):
    """:395-449."""
    poly_incident, mask = _clip(clipping_poly, subject_poly, clipping_norm, subject_norm)
    nn = M.normalize(clipping_norm).unsqueeze(-2)
    p0 = clipping_poly[..., 0:1, :]
    poly_ref = poly_incident - ((poly_incident - p0) * nn).sum(-1, keepdim=True) * nn
    behind = ((poly_incident - p0) * (-clipping_norm).unsqueeze(-2)).sum(-1) > 1e-6
    mask = mask & behind
    best = _manifold_points(poly_ref, mask, clipping_norm)
    contact_pts = _take(poly_ref, best)
    mask_pts = _take(mask, best)
    pen_dir = _take(poly_incident, best) - contact_pts
    penetration = (pen_dir * (-clipping_norm).unsqueeze(-2)).sum(-1)
    dist = torch.where(mask_pts, -penetration, torch.ones_like(penetration))
    normal = (-sep_axis).unsqueeze(-2).expand_as(contact_pts)
    return dist, contact_pts, normal


def _sat_hull_hull(faces_a, faces_b, verts_a, verts_b, normals_a, normals_b, edges_a, edges_b):
    """:464-601.  All inputs carry a leading env dim."""
    da = edges_a[..., 0, :] - edges_a[..., 1, :]
    db = edges_b[..., 0, :] - edges_b[..., 1, :]
    na_, nb_ = da.shape[-2], db.shape[-2]
    da_r = da.repeat(*([1] * (da.dim() - 2)), nb_, 1)          # tile
    db_r = db.repeat_interleave(na_, dim=-2)
    ee = M.normalize(M.cross(da_r, db_r))
    axes = torch.cat([normals_a, normals_b, ee], dim=-2)
    sa = torch.matmul(axes, verts_a.transpose(-1, -2))
    sb = torch.matmul(axes, verts_b.transpose(-1, -2))
    dist1 = sa.max(-1).values - sb.min(-1).values
    dist2 = sb.max(-1).values - sa.min(-1).values
    sign = torch.where(dist1 > dist2, -1, 1)
    support = torch.minimum(dist1, dist2)
    support = torch.where(~(axes == 0.0).all(-1), support, torch.full_like(support, 1e6))
    best_idx = support.argmin(-1)
    best_sign = _sel(sign, best_idx)
    best_axis = _sel(axes, best_idx)
    is_edge = best_idx >= (normals_a.shape[-2] + normals_b.shape[-2])
    dist_a = (best_axis.unsqueeze(-2) * normals_a).sum(-1)
    dist_b = (best_axis.unsqueeze(-2) * normals_b).sum(-1)
    a_max, b_max, a_min, b_min = dist_a.argmax(-1), dist_b.argmax(-1), dist_a.argmin(-1), dist_b.argmin(-1)
    pos_sign = best_sign > 0
    w3 = pos_sign[..., None, None]
    w1 = pos_sign[..., None]
    ref_face = torch.where(w3, _sel(faces_a, a_max), _sel(faces_b, b_max))
    ref_norm = torch.where(w1, _sel(normals_a, a_max), _sel(normals_b, b_max))
    inc_face = torch.where(w3, _sel(faces_b, b_min), _sel(faces_a, a_min))
    inc_norm = torch.where(w1, _sel(normals_b, b_min), _sel(normals_a, a_min))
    dist, pos, normal = _create_contact_manifold(ref_face, inc_face, ref_norm, inc_norm, -best_sign.unsqueeze(-1).to(best_axis.dtype) * best_axis)
    idx = dist.argmin(-1)
    d_at, p_at = _sel(dist, idx), _sel(pos, idx)
    one = torch.ones_like(d_at)
    dist = torch.where(is_edge.unsqueeze(-1), torch.stack([d_at, one, one, one], dim=-1), dist)
    pos = torch.where(is_edge[..., None, None], p_at.unsqueeze(-2).expand_as(pos), pos)
    return dist, pos, normal


def capsule_convex(m, d, g1, g2):
    cap_pos_w, cap_mat, _ = _geom(m, d, g1)
    cpos, cmat, cv = _geom(m, d, g2)
    B = cpos.shape[0]
    r, length = m.geom_size[g1, 0], m.geom_size[g1, 1]
    rt = cmat.transpose(-1, -2)
    cap_pos = torch.matmul(rt, (cap_pos_w - cpos).unsqueeze(-1)).squeeze(-1)
    axis = torch.matmul(rt, cap_mat[:, :, 2].unsqueeze(-1)).squeeze(-1)
    seg = axis * length
    cap_pts = torch.stack([cap_pos - seg, cap_pos + seg], dim=1)  # (B,2,3)
    faces = cv.vert[cv.face]
    normals = cv.facenorm
    pts = cap_pts.unsqueeze(1) - (normals * r).unsqueeze(1)  # (B,nface,2,3)
    sup = ((pts - faces[:, 0].unsqueeze(1)) * normals.unsqueeze(1)).sum(-1).min(-1).values  # (B,nface)
    has_support = (sup < 0).all(-1)
    face, normal = _best_face(cv, sup, B)
    e0 = torch.roll(face, 1, dims=1)
    e1 = face
    en = M.cross(e1 - e0, normal.unsqueeze(1).expand_as(e0))
    clipped, mask = _clip_edge_to_planes(cap_pts[:, 0], cap_pts[:, 1], e0, en)
    clipped = clipped - (normal * r).unsqueeze(1)
    face_pts = clipped - ((clipped - face[:, 0:1]) * normal.unsqueeze(1)).sum(-1, keepdim=True) * normal.unsqueeze(1)
    pos = (clipped + face_pts) * 0.5
    norm = normal.unsqueeze(1).expand(B, 2, 3)
    penetration = torch.where(mask & has_support.unsqueeze(-1), ((face_pts - clipped) * normal.unsqueeze(1)).sum(-1),
                              torch.full_like(pos[..., 0], -1.0))
    nE = e0.shape[1]
    edge_closest, cap_closest = M.closest_segment_to_segment_points(e0, e1, cap_pts[:, 0:1].expand(B, nE, 3), cap_pts[:, 1:2].expand(B, nE, 3))
    e_idx = ((edge_closest - cap_closest) ** 2).sum(-1).argmin(-1)
    cap_c, edge_c = _sel(cap_closest, e_idx), _sel(edge_closest, e_idx)
    edge_axis, edge_dist = M.normalize_with_norm(cap_c - edge_c)
    edge_pos = (edge_c + (cap_c - edge_axis * r)) * 0.5
    edge_pen = r - edge_dist
    has_edge = (edge_pen > 0)
    pos_u = torch.cat([edge_pos.unsqueeze(1), pos[:, 1:]], dim=1)
    pos = torch.where(has_edge[:, None, None], pos_u, pos)
    norm_u = torch.cat([edge_axis.unsqueeze(1), norm[:, 1:]], dim=1)
    n = -torch.where(has_edge[:, None, None], norm_u, norm)
    pos = _to_world(cpos, cmat, pos)
    n = torch.matmul(n, cmat.transpose(-1, -2))
    pen_u = torch.cat([edge_pen.unsqueeze(1), penetration[:, 1:]], dim=1)
    dist = -torch.where(has_edge[:, None], pen_u, penetration)
    return dist, pos, M.make_frame(n)


def convex_convex(m, d, g1, g2):
    p1, r1, c1 = _geom(m, d, g1)
    p2, r2, c2 = _geom(m, d, g2)
    B = p1.shape[0]
    f1, f2 = c1.face, c2.face
    s1, s2 = f1.shape[-1], f2.shape[-1]
    if s1 < s2:
        f1 = torch.cat([f1, f1[:, -1:].expand(-1, s2 - s1)], dim=1)
    elif s2 < s1:
        f2 = torch.cat([f2, f2[:, -1:].expand(-1, s1 - s2)], dim=1)
    swapped = c1.vert.shape[0] > c2.vert.shape[0]
    if swapped:
        p1, r1, c1, f1, p2, r2, c2, f2 = p2, r2, c2, f2, p1, r1, c1, f1
    to_pos = torch.matmul(r2.transpose(-1, -2), (p1 - p2).unsqueeze(-1)).squeeze(-1)
    to_mat = torch.matmul(r2.transpose(-1, -2), r1)
    xf = lambda x: to_pos[:, None, :] + torch.matmul(x, to_mat.transpose(-1, -2))
    faces1 = c1.vert[f1]  # (nf, nfv, 3)
    nf1 = faces1.shape[0]
    faces1 = xf(faces1.reshape(1, -1, 3).expand(B, -1, 3)).reshape(B, nf1, -1, 3)
    normals1 = torch.matmul(c1.facenorm.expand(B, -1, 3), to_mat.transpose(-1, -2))
    verts1 = xf(c1.vert.expand(B, -1, 3))
    faces2 = c2.vert[f2].expand(B, *f2.shape, 3)
    normals2 = c2.facenorm.expand(B, -1, 3)
    verts2 = c2.vert.expand(B, -1, 3)
    edges1 = _take(verts1, c1.edge.reshape(1, -1).expand(B, -1)).reshape(B, -1, 2, 3)
    edges2 = verts2[:, c2.edge]
    dist, pos, normal = _sat_hull_hull(faces1, faces2, verts1, verts2, normals1, normals2, edges1, edges2)
    pos = _to_world(p2, r2, pos)
    normal = torch.matmul(normal, r2.transpose(-1, -2))
    if swapped:
        normal = -normal
    return dist, pos, M.make_frame(normal)
