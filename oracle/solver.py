"""ORACLE (test infrastructure): primal CG / Newton constraint solver, batched over envs.

Follows reference ``mujoco_torch/_src/solver.py:244-553``.  The reference's per-env ``while`` loops
(``while_loop`` :29-61; masked under vmap per patches/_pr175526_while_loop_vmap.py:83-127) are
reproduced with a per-env "still running" mask: the body is evaluated for every env and the old
carry kept wherever that env's condition is false -- net effect per env == the eager Python loop.
"""

import torch

from . import omath as M
from .model import DSBL_FRICTIONLOSS, DSBL_WARMSTART, mjMINVAL

CG, NEWTON = 1, 2


class Ctx:
    FIELDS = ("qacc", "qfrc_constraint", "Jaref", "efc_force", "Ma", "grad", "Mgrad", "search", "active", "gauss", "cost",
              "prev_cost", "niter")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def copy(self):
        return Ctx(**self.__dict__)

    def where(self, mask, other):
        """Keep self where mask (B,) is True, else other."""
        out = {}
        for k in self.FIELDS:
            a, b = getattr(self, k), getattr(other, k)
            mk = mask.reshape((-1,) + (1,) * (a.dim() - 1))
            out[k] = torch.where(mk, a, b)
        return Ctx(**out)


def solve(m, d, fixed_iterations=False):
    nv, nefc = m.nv, m.nefc
    dt = d.qpos.dtype
    B = d.qpos.shape[0]
    J, D, aref, floss = d.efc_J, d.efc_D, d.efc_aref, d.efc_frictionloss
    JT = J.transpose(-1, -2)
    qfrc_smooth, qacc_smooth = d.qfrc_smooth, d.qacc_smooth
    dense_M = d.qM
    ne_nf = m.ne + m.nf
    use_floss = m.nf > 0 and not (m.disableflags & DSBL_FRICTIONLOSS)
    inertia_scale = m.meaninertia * max(1, nv)
    tolerance, ls_tolerance = m.tolerance, m.ls_tolerance
    eq_fric_mask = torch.arange(nefc) < ne_nf
    mv_ = lambda A, x: torch.matmul(A, x.unsqueeze(-1)).squeeze(-1)

    def update_constraint(ctx):
        """:320-357."""
        active = (ctx.Jaref < 0) | eq_fric_mask
        floss_force = torch.zeros((B, nefc), dtype=dt)
        floss_cost = torch.zeros((B,), dtype=dt)
        if use_floss:
            has = floss > 0
            r = 1.0 / (D + (D == 0.0) * mjMINVAL)
            lin_neg = (ctx.Jaref <= -r * floss) & has
            lin_pos = (ctx.Jaref >= r * floss) & has
            active = active & ~lin_neg & ~lin_pos
            floss_force = torch.where(lin_neg, floss, torch.where(lin_pos, -floss, torch.zeros_like(floss)))
            floss_cost = (lin_neg * (-0.5 * r * floss * floss - floss * ctx.Jaref)).sum(-1) + (
                lin_pos * (-0.5 * r * floss * floss + floss * ctx.Jaref)).sum(-1)
        efc_force = D * -ctx.Jaref * active + floss_force
        qfrc_con = mv_(JT, efc_force)
        gauss = 0.5 * ((ctx.Ma - qfrc_smooth) * (ctx.qacc - qacc_smooth)).sum(-1)
        cost = 0.5 * torch.sum(D * ctx.Jaref * ctx.Jaref * active, dim=-1) + gauss + floss_cost
        out = ctx.copy()
        out.qfrc_constraint, out.gauss, out.cost, out.prev_cost = qfrc_con, gauss, cost, ctx.cost
        out.efc_force, out.active = efc_force, active
        return out

    def update_gradient(ctx):
        """:359-376."""
        grad = ctx.Ma - qfrc_smooth - ctx.qfrc_constraint
        if m.solver == CG:
            mgrad = M.small_cholesky_solve(grad, d.qLD)
        elif m.solver == NEWTON:
            wJT = JT * D.unsqueeze(1) * ctx.active.unsqueeze(1)
            if nv <= 16:
                h = dense_M + (wJT.unsqueeze(-1) * J.unsqueeze(1)).sum(-2)
            else:
                h = dense_M + torch.matmul(wJT, J)
            L = M.small_cholesky(h)
            mgrad = M.small_cholesky_solve(grad, L)
        else:
            raise NotImplementedError(f"unsupported solver type: {m.solver}")
        out = ctx.copy()
        out.grad, out.Mgrad = grad, mgrad
        return out

    def create_context(qacc, qfrc_con, grad_flag=True):
        """:293-318."""
        jaref = mv_(J, qacc) - aref
        ma = (dense_M * qacc.unsqueeze(1)).sum(-1)
        z = lambda *s: torch.zeros((B,) + s, dtype=dt)
        ctx = Ctx(qacc=qacc, qfrc_constraint=qfrc_con, Jaref=jaref, efc_force=z(nefc), Ma=ma, grad=z(nv), Mgrad=z(nv),
                  search=z(nv), active=torch.zeros((B, nefc), dtype=torch.bool), gauss=z(),
                  cost=torch.full((B,), float("inf"), dtype=dt), prev_cost=z(), niter=torch.zeros((B,), dtype=torch.long))
        ctx = update_constraint(ctx)
        if grad_flag:
            ctx = update_gradient(ctx)
            ctx.search = -ctx.Mgrad
        return ctx

    def linesearch(ctx):
        """:378-497."""
        smag = M.norm(ctx.search) * inertia_scale
        gtol = tolerance * ls_tolerance * smag
        mv = (dense_M * ctx.search.unsqueeze(1)).sum(-1)
        jv = mv_(J, ctx.search)
        quad_gauss = torch.stack((ctx.gauss, (ctx.search * ctx.Ma).sum(-1) - (ctx.search * qfrc_smooth).sum(-1),
                                  0.5 * (ctx.search * mv).sum(-1)), dim=-1)  # (B,3)
        quad = torch.stack((0.5 * ctx.Jaref * ctx.Jaref, jv * ctx.Jaref, 0.5 * jv * jv), dim=-1) * D.unsqueeze(-1)  # (B,nefc,3)

        def point_fn(alpha):
            x = ctx.Jaref + alpha.unsqueeze(-1) * jv
            active = (x < 0) | eq_fric_mask
            floss_adjust = torch.zeros((B, 3), dtype=dt)
            if use_floss:
                has = floss > 0
                r = 1.0 / (D + (D == 0.0) * mjMINVAL)
                rf = r * floss
                fl_ln = (x <= -rf) & has
                fl_lp = (x >= rf) & has
                qf0 = (fl_ln * floss * (-0.5 * rf - ctx.Jaref)).sum(-1) + (fl_lp * floss * (-0.5 * rf + ctx.Jaref)).sum(-1)
                qf1 = (fl_ln * (-floss * jv)).sum(-1) + (fl_lp * (floss * jv)).sum(-1)
                floss_adjust = torch.stack([qf0, qf1, torch.zeros_like(qf0)], dim=-1)
                active = active & ~fl_ln & ~fl_lp
            quad_total = quad_gauss + torch.sum(quad * active.unsqueeze(-1), dim=1) + floss_adjust
            cost = alpha * alpha * quad_total[:, 2] + alpha * quad_total[:, 1] + quad_total[:, 0]
            deriv_0 = 2 * alpha * quad_total[:, 2] + quad_total[:, 1]
            deriv_1 = 2 * quad_total[:, 2] + (quad_total[:, 2] == 0) * mjMINVAL
            return (alpha, cost, deriv_0, deriv_1)

        sel = lambda c, x, y: tuple(torch.where(c, a, b) for a, b in zip(x, y))

        def ls_cond(lo, hi, swap, it):
            done = it >= m.ls_iterations
            done = done | (~swap)
            done = done | ((lo[2] < 0) & (lo[2] > -gtol))
            done = done | ((hi[2] > 0) & (hi[2] < gtol))
            return ~done

        def ls_body(lo, hi):
            lo_next = point_fn(lo[0] - lo[2] / lo[3])
            hi_next = point_fn(hi[0] - hi[2] / hi[3])
            mid = point_fn(0.5 * (lo[0] + hi[0]))
            in_bracket = lambda x, y: ((x < y) & (y < 0)) | ((x > y) & (y > 0))
            not_bracketed = (lo[2] < 0) == (hi[2] < 0)
            swp = lambda cur, cand: in_bracket(cur, cand) | (not_bracketed & (torch.abs(cand) < torch.abs(cur)))
            s1 = swp(lo[2], lo_next[2])
            lo = sel(s1, lo_next, lo)
            s2 = swp(lo[2], mid[2])
            lo = sel(s2, mid, lo)
            s3 = swp(lo[2], hi_next[2])
            lo = sel(s3, hi_next, lo)
            s4 = swp(hi[2], hi_next[2])
            hi = sel(s4, hi_next, hi)
            s5 = swp(hi[2], mid[2])
            hi = sel(s5, mid, hi)
            s6 = swp(hi[2], lo_next[2])
            hi = sel(s6, lo_next, hi)
            return lo, hi, s1 | s2 | s3 | s4 | s5 | s6

        p0 = point_fn(torch.zeros((B,), dtype=dt))
        p1 = point_fn(p0[0] - p0[2] / p0[3])
        early = torch.abs(p1[2]) < gtol
        lesser = p1[2] < p0[2]
        hi = sel(lesser, p0, p1)
        lo = sel(lesser, p1, p0)
        swap = ~early
        it = torch.zeros((B,), dtype=torch.long)
        if fixed_iterations:
            for _ in range(m.ls_iterations):
                lo, hi, swap = ls_body(lo, hi)
        else:
            run = ls_cond(lo, hi, swap, it)
            while bool(run.any()):
                nlo, nhi, nswap = ls_body(lo, hi)
                lo, hi = sel(run, nlo, lo), sel(run, nhi, hi)
                swap = torch.where(run, nswap, swap)
                it = torch.where(run, it + 1, it)
                run = run & ls_cond(lo, hi, swap, it)
        improved = ((lo[1] < p0[1]) | (hi[1] < p0[1])).to(dt)
        alpha = torch.where(lo[1] < hi[1], lo[0], hi[0])
        out = ctx.copy()
        out.qacc = ctx.qacc + improved.unsqueeze(-1) * ctx.search * alpha.unsqueeze(-1)
        out.Ma = ctx.Ma + improved.unsqueeze(-1) * mv * alpha.unsqueeze(-1)
        out.Jaref = ctx.Jaref + improved.unsqueeze(-1) * jv * alpha.unsqueeze(-1)
        return out

    def cond(ctx):
        improvement = (ctx.prev_cost - ctx.cost) / inertia_scale
        gradient = torch.linalg.norm(ctx.grad, dim=-1) / inertia_scale
        done = ctx.niter >= m.iterations
        done = done | (improvement < tolerance)
        done = done | (gradient < tolerance)
        return ~done

    def body(ctx):
        ctx = linesearch(ctx)
        prev_grad, prev_Mgrad = ctx.grad, ctx.Mgrad
        ctx = update_constraint(ctx)
        ctx = update_gradient(ctx)
        if m.solver == NEWTON:
            search = -ctx.Mgrad
        else:
            beta = (ctx.grad * (ctx.Mgrad - prev_Mgrad)).sum(-1)
            beta = beta / torch.clamp_min((prev_grad * prev_Mgrad).sum(-1), mjMINVAL)
            beta = torch.clamp_min(beta, 0)
            search = -ctx.Mgrad + beta.unsqueeze(-1) * ctx.search
        ctx.search = search
        ctx.niter = ctx.niter + 1
        return ctx

    qacc = qacc_smooth
    qfrc_constraint = d.qfrc_constraint
    if not (m.disableflags & DSBL_WARMSTART):
        warm = create_context(d.qacc_warmstart, qfrc_constraint, grad_flag=False)
        smth = create_context(qacc_smooth, qfrc_constraint, grad_flag=False)
        qacc = torch.where((warm.cost < smth.cost).unsqueeze(-1), d.qacc_warmstart, qacc_smooth)

    ctx = create_context(qacc, qfrc_constraint)
    if m.iterations == 1:
        ctx = body(ctx)
    elif fixed_iterations:
        for _ in range(m.iterations):
            ctx = body(ctx)
    else:
        run = cond(ctx)
        while bool(run.any()):
            new = body(ctx)
            ctx = new.where(run, ctx)
            run = run & cond(ctx)
    d.qacc_warmstart = ctx.qacc.clone()
    d.qacc, d.qfrc_constraint, d.efc_force = ctx.qacc, ctx.qfrc_constraint, ctx.efc_force
    d.solver_niter_oracle = ctx.niter
    return d
