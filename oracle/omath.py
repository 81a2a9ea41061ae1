"""ORACLE (test infrastructure, never the product path): batched restatement of the reference's
math helpers.  Every function works on tensors with arbitrary leading batch dims ``(..., k)`` and
cites the single-env reference function it follows (paths relative to the reference tree,
``mujoco_torch/_src/math.py``).
"""

import torch

mjMINVAL = 1e-15


def cross(a, b):
    """math.py:63-76."""
    return torch.stack(
        [
            a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
            a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
            a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0],
        ],
        dim=-1,
    )


def dot(a, b):
    return (a * b).sum(-1)


def safe_div(num, den):
    """math.py:79-81."""
    return num / (den + mjMINVAL * (den == 0))


def norm(x):
    """math.py:196-213 (norm over the last axis; all-zero vectors give exactly 0)."""
    is_zero = (x == 0).all(dim=-1, keepdim=True)
    xs = torch.where(is_zero, torch.ones_like(x), x)
    n = torch.linalg.norm(xs, dim=-1)
    return torch.where(is_zero.squeeze(-1), torch.zeros_like(n), n)


def normalize_with_norm(x):
    """math.py:216-230."""
    n = norm(x)
    return x / (n + 1e-6 * (n == 0.0)).unsqueeze(-1), n


def normalize(x):
    return normalize_with_norm(x)[0]


def rotate(vec, quat):
    """math.py:246-261."""
    s, u = quat[..., 0:1], quat[..., 1:]
    r = 2 * (dot(u, vec).unsqueeze(-1) * u) + (s * s - dot(u, u).unsqueeze(-1)) * vec
    return r + 2 * s * cross(u, vec)


def quat_inv(q):
    """math.py:264-273."""
    return q * torch.tensor([1, -1, -1, -1], dtype=q.dtype)


def quat_mul(u, v):
    """math.py:283-300."""
    return torch.stack(
        [
            u[..., 0] * v[..., 0] - u[..., 1] * v[..., 1] - u[..., 2] * v[..., 2] - u[..., 3] * v[..., 3],
            u[..., 0] * v[..., 1] + u[..., 1] * v[..., 0] + u[..., 2] * v[..., 3] - u[..., 3] * v[..., 2],
            u[..., 0] * v[..., 2] - u[..., 1] * v[..., 3] + u[..., 2] * v[..., 0] + u[..., 3] * v[..., 1],
            u[..., 0] * v[..., 3] + u[..., 1] * v[..., 2] - u[..., 2] * v[..., 1] + u[..., 3] * v[..., 0],
        ],
        dim=-1,
    )


def quat_mul_axis(q, axis):
    """math.py:303-320."""
    return torch.stack(
        [
            -q[..., 1] * axis[..., 0] - q[..., 2] * axis[..., 1] - q[..., 3] * axis[..., 2],
            q[..., 0] * axis[..., 0] + q[..., 2] * axis[..., 2] - q[..., 3] * axis[..., 1],
            q[..., 0] * axis[..., 1] + q[..., 3] * axis[..., 0] - q[..., 1] * axis[..., 2],
            q[..., 0] * axis[..., 2] + q[..., 1] * axis[..., 1] - q[..., 2] * axis[..., 0],
        ],
        dim=-1,
    )


def quat_to_mat(q):
    """math.py:323-351 -> (..., 3, 3)."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r0 = torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], dim=-1)
    r1 = torch.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], dim=-1)
    r2 = torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], dim=-1)
    return torch.stack([r0, r1, r2], dim=-2)


def quat_to_axis_angle(q):
    """math.py:354-360."""
    axis, sin_a_2 = normalize_with_norm(q[..., 1:])
    angle = 2 * torch.arctan2(sin_a_2, q[..., 0])
    angle = torch.where(angle > torch.pi, angle - 2 * torch.pi, angle)
    return axis, angle


def axis_angle_to_quat(axis, angle):
    """math.py:363-374."""
    s, c = torch.sin(angle * 0.5), torch.cos(angle * 0.5)
    return torch.cat([c.unsqueeze(-1), axis * s.unsqueeze(-1)], dim=-1)


def quat_integrate(q, v, dt):
    """math.py:377-383."""
    v, norm_ = normalize_with_norm(v)
    angle = dt * norm_
    q_res = axis_angle_to_quat(v, angle)
    q_res = quat_mul(q, q_res)
    return normalize(q_res)


def quat_sub(u, v):
    """math.py:276-280."""
    q = quat_mul(quat_inv(v), u)
    axis, angle = quat_to_axis_angle(q)
    return axis * angle.unsqueeze(-1)


_TRI_ID = [[0, 3, 4], [3, 1, 5], [4, 5, 2]]


def inert_mul(i, v):
    """math.py:415-429.  i: (..., 10), v: (..., 6)."""
    inr = i[..., _TRI_ID]  # (..., 3, 3)
    pos, mass = i[..., 6:9], i[..., 9:10]
    ang = (inr * v[..., None, :3]).sum(-1) + cross(pos, v[..., 3:])
    vel = mass * v[..., 3:] - cross(pos, v[..., :3])
    return torch.cat((ang, vel), dim=-1)


def motion_cross(u, v):
    """math.py:455-467."""
    ang = cross(u[..., :3], v[..., :3])
    vel = cross(u[..., 3:], v[..., :3]) + cross(u[..., :3], v[..., 3:])
    return torch.cat((ang, vel), dim=-1)


def motion_cross_force(v, f):
    """math.py:470-482."""
    ang = cross(v[..., :3], f[..., :3]) + cross(v[..., 3:], f[..., 3:])
    vel = cross(v[..., :3], f[..., 3:])
    return torch.cat((ang, vel), dim=-1)


def transform_motion(vel, offset, rotmat):
    """math.py:437-452."""
    ang, lin = vel[..., :3], vel[..., 3:]
    rt = rotmat.transpose(-1, -2)
    lin = (rt * (lin - cross(offset, ang))[..., None, :]).sum(-1)
    ang = (rt * ang[..., None, :]).sum(-1)
    return torch.cat([ang, lin], dim=-1)


def orthogonals(a):
    """math.py:485-493."""
    y = torch.tensor([0.0, 1.0, 0.0], dtype=a.dtype).expand_as(a)
    z = torch.tensor([0.0, 0.0, 1.0], dtype=a.dtype).expand_as(a)
    cond = ((-0.5 < a[..., 1]) & (a[..., 1] < 0.5)).unsqueeze(-1)
    b = torch.where(cond, y, z)
    b = b - a * dot(a, b).unsqueeze(-1)
    b = normalize(b) * (a != 0).any(dim=-1, keepdim=True)
    return b, cross(a, b)


def make_frame(a):
    """math.py:496-500 -> (..., 3, 3) rows [a, b, c]."""
    a = normalize(a)
    b, c = orthogonals(a)
    return torch.stack([a, b, c], dim=-2)


def closest_segment_point(a, b, pt):
    """math.py:506-510."""
    ab = b - a
    t = dot(pt - a, ab) / (dot(ab, ab) + 1e-6)
    return a + torch.clamp(t, 0.0, 1.0).unsqueeze(-1) * ab


def closest_segment_point_and_dist(a, b, pt):
    """math.py:513-520."""
    closest = closest_segment_point(a, b, pt)
    diff = pt - closest
    return closest, dot(diff, diff)


def closest_segment_to_segment_points(a0, a1, b0, b1):
    """math.py:523-569."""
    dir_a, len_a = normalize_with_norm(a1 - a0)
    dir_b, len_b = normalize_with_norm(b1 - b0)
    half_len_a = len_a * 0.5
    half_len_b = len_b * 0.5
    a_mid = a0 + dir_a * half_len_a.unsqueeze(-1)
    b_mid = b0 + dir_b * half_len_b.unsqueeze(-1)
    trans = a_mid - b_mid
    dira_dot_dirb = dot(dir_a, dir_b)
    dira_dot_trans = dot(dir_a, trans)
    dirb_dot_trans = dot(dir_b, trans)
    denom = 1 - dira_dot_dirb * dira_dot_dirb
    orig_t_a = (-dira_dot_trans + dira_dot_dirb * dirb_dot_trans) / (denom + 1e-6)
    orig_t_b = dirb_dot_trans + orig_t_a * dira_dot_dirb
    t_a = torch.maximum(torch.minimum(orig_t_a, half_len_a), -half_len_a)
    t_b = torch.maximum(torch.minimum(orig_t_b, half_len_b), -half_len_b)
    best_a = a_mid + dir_a * t_a.unsqueeze(-1)
    best_b = b_mid + dir_b * t_b.unsqueeze(-1)
    new_a, d1 = closest_segment_point_and_dist(a0, a1, best_b)
    new_b, d2 = closest_segment_point_and_dist(b0, b1, best_a)
    c = (d1 < d2).unsqueeze(-1)
    best_a = torch.where(c, new_a, best_a)
    best_b = torch.where(c, best_b, new_b)
    return best_a, best_b


_INLINE_CHOLESKY_MAX_SIZE = 16


def small_cholesky(A):
    """math.py:87-129.  A: (..., n, n)."""
    n = A.shape[-1]
    if n > _INLINE_CHOLESKY_MAX_SIZE:
        A = A + 1e-10 * torch.eye(n, dtype=A.dtype)
        return torch.linalg.cholesky(A)
    L = [[torch.zeros_like(A[..., 0, 0]) for _ in range(n)] for _ in range(n)]
    for j in range(n):
        s = A[..., j, j]
        for k in range(j):
            s = s - L[j][k] * L[j][k]
        L[j][j] = torch.sqrt(torch.clamp_min(s, 1e-12))
        for i in range(j + 1, n):
            s = A[..., i, j]
            for k in range(j):
                s = s - L[i][k] * L[j][k]
            L[i][j] = s / L[j][j]
    return torch.stack([torch.stack(row, dim=-1) for row in L], dim=-2)


def small_cholesky_solve(x, L):
    """math.py:132-168.  x: (..., n), L: (..., n, n)."""
    n = L.shape[-1]
    if n > _INLINE_CHOLESKY_MAX_SIZE:
        return torch.cholesky_solve(x.unsqueeze(-1), L).squeeze(-1)
    y = []
    for i in range(n):
        s = x[..., i]
        for k in range(i):
            s = s - L[..., i, k] * y[k]
        y.append(s / L[..., i, i])
    result = [None] * n
    for i in range(n - 1, -1, -1):
        s = y[i]
        for k in range(i + 1, n):
            s = s - L[..., k, i] * result[k]
        result[i] = s / L[..., i, i]
    return torch.stack(result, dim=-1)
