"""Parity of the CUDA step against the CPU oracle on seeded inputs (one step, f64 1e-9 / f32 1e-4 relative)."""

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

FIELDS_SMOOTH = ["xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "subtree_com", "cinert",
                 "crb", "cdof", "cdof_dot", "cvel", "qM", "qLD", "actuator_length", "actuator_moment", "actuator_velocity",
                 "actuator_force", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth", "qacc_smooth"]
FIELDS_EFC = ["efc_J", "efc_D", "efc_aref", "efc_frictionloss", "efc_force", "qfrc_constraint"]
FIELDS_STATE = ["qpos", "qvel", "qacc", "qacc_warmstart", "act", "time"]

CASES = [
    ("cartpole", {}, 0.0),
    ("humanoid", {}, 0.4),
    ("humanoid", {"integrator": 3}, 0.4),
    ("halfcheetah", {}, 0.6),
    ("ant", {}, 0.0),
    ("ant_zoo", {"cone": 1, "solver": 1}, 0.5),
    ("hopper", {}, 0.6),
    ("walker2d", {}, 0.8),
]


def _compare(mj, B, dtype, tol, drop, nsteps=1):
    qpos, qvel, ctrl = util.seeded_state(mj, B, drop=drop)
    mo, do = util.run_oracle(mj, qpos, qvel, ctrl, dtype, nsteps)
    mx, dc = util.run_cuda(mj, qpos, qvel, ctrl, dtype, nsteps)
    assert (mx.ne, mx.nf, mx.nl, mx.ncon, mx.nefc) == (mo.ne, mo.nf, mo.nl, mo.ncon, mo.nefc)
    bad = {}
    if mo.ncon:
        # active contact set must match exactly
        act_o = (do.contact_dist - mo.contact["includemargin"]) < 0
        act_c = (dc.contact.dist.cpu() - dc.contact.includemargin.cpu()) < 0
        assert torch.equal(act_o, act_c), "active contact set differs"
        assert torch.equal(mo.contact["geom1"], dc.contact.geom1[0].cpu()) and torch.equal(mo.contact["geom2"], dc.contact.geom2[0].cpu())
        for k, ko in (("contact.dist", "contact_dist"), ("contact.pos", "contact_pos"), ("contact.frame", "contact_frame")):
            e = util.rel_err(util.get(dc, k), getattr(do, ko))
            if e > tol:
                bad[k] = e
    for k in FIELDS_SMOOTH + (FIELDS_EFC if mo.nefc else []) + FIELDS_STATE:
        if not hasattr(do, k):
            continue
        e = util.rel_err(getattr(dc, k), getattr(do, k))
        if e > tol:
            bad[k] = e
    assert not bad, f"fields beyond tol {tol}: {bad}"


@pytest.mark.parametrize("name,opt,drop", CASES)
def test_step_f64(name, opt, drop):
    _compare(util.load_model(name, **opt), 64, torch.float64, 1e-9, drop)


@pytest.mark.parametrize("name,opt,drop", CASES)
def test_step_f32(name, opt, drop):
    _compare(util.load_model(name, **opt), 64, torch.float32, 1e-4, drop)
