"""Golden vectors recorded from the REFERENCE'S OWN code (tests/golden/make_golden.py: the unmodified
`/root/reference/mujoco_torch/_src` step run under import shims, single-env eager float64).

  * CPU tier:  the oracle restatement must reproduce them (this is what pins the oracle).
  * GPU tier:  the CUDA path, through the C ABI, must reproduce them -- every one of them, including the recordings of
               `step(..., fixed_iterations=True)`.
Tolerance 1e-9 relative per field (field-scale), float64.  Contact frames of coincident-point contacts are arbitrary
(see test_gpu_step) and skipped; everything else, including contact order, geom ids, efc layout, is compared.
"""

import glob
import os

import numpy as np
import pytest
import torch

from tests import util

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
GOLDEN_CPU_ONLY = []   # every recording is held against both the oracle (CPU tier) and the CUDA path (GPU tier)
FIELDS = ["qpos", "qvel", "qacc", "qacc_warmstart", "act", "time", "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis",
          "geom_xpos", "geom_xmat", "subtree_com", "cinert", "crb", "cdof", "cdof_dot", "cvel", "qM", "qLD", "actuator_length",
          "actuator_moment", "actuator_velocity", "actuator_force", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth",
          "qacc_smooth", "qfrc_constraint", "efc_J", "efc_D", "efc_aref", "efc_frictionloss", "efc_force", "sensordata"]
TOL = 1e-9


def _load(path):
    z = np.load(path)
    opt = {str(k): (int(v) if float(v).is_integer() else float(v)) for k, v in zip(z["opt_keys"], z["opt_vals"])}
    opt.pop("fixed_iterations", None)   # an argument of step(), recorded next to the model options (see _fixed)
    mj = util.load_model(str(z["model"]), **opt)
    return z, mj


def _extra(z):
    """Recorded step inputs beyond qpos / qvel / ctrl / act (applied forces, equality switches, mocap poses)."""
    return {k: z["in_" + k] for k in util.EXTRA_INPUTS if "in_" + k in z.files}


def _fixed(z):
    return bool(dict(zip((str(k) for k in z["opt_keys"]), z["opt_vals"])).get("fixed_iterations", 0))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))


def _check(z, get, get_contact, label):
    bad = {}
    nefc = z["out_efc_J"].shape[1]
    degenerate = np.zeros(z["in_qpos"].shape[0], dtype=bool)
    if z["contact_dist"].shape[1]:
        for k in ("contact_dim", "geom1", "geom2", "efc_address"):
            got = np.asarray(get_contact(k))
            assert np.array_equal(got.astype(np.int64), z["contact_" + k].astype(np.int64)), f"{label}: contact.{k} differs"
        for k in ("includemargin", "friction", "solref", "solreffriction", "solimp", "dist"):
            e = _rel(get_contact(k), z["contact_" + k])
            if e > TOL:
                bad["contact." + k] = e
        # positions: an inactive slot of a convex (box / mesh) pair holds an arbitrary tie-broken manifold point -- skip those
        gtype = np.asarray(util.load_model(str(z["model"])).geom_type)
        manifold = (gtype[z["contact_geom2"].astype(np.int64)] >= 6) | (gtype[z["contact_geom1"].astype(np.int64)] == 1)   # convex / hfield pairs
        skip = manifold & (z["contact_dist"] - z["contact_includemargin"] >= 0)
        pe = np.abs(np.asarray(get_contact("pos"), dtype=np.float64) - z["contact_pos"]).max(-1) / max(np.abs(z["contact_pos"]).max(), 1e-6)
        if np.where(skip, 0.0, pe).max() > TOL:
            bad["contact.pos"] = float(np.where(skip, 0.0, pe).max())
        # frames: skip contacts whose two closest points coincide (normal undefined)
        fr_ref, fr = z["contact_frame"], np.asarray(get_contact("frame"), dtype=np.float64)
        diff = np.abs(fr - fr_ref).max(axis=(-1, -2))
        nondeg = diff < 1e-6
        degenerate = (~nondeg).any(axis=1)
        # a differing frame is only acceptable for a coincident-point contact: both frames must still be orthonormal
        for b, c in zip(*np.nonzero(~nondeg)):
            assert abs(np.linalg.det(fr[b, c]) - 1) < 1e-6 or np.allclose(fr[b, c], 0), f"{label}: bad frame env {b} contact {c}"
        if nondeg.any() and diff[nondeg].max() > 10 * TOL:
            bad["contact.frame"] = float(diff[nondeg].max())
    for k in FIELDS:
        ref = z["out_" + k]
        got = np.asarray(get(k), dtype=np.float64)
        if k.startswith("efc_") or k in ("qfrc_constraint",):
            if nefc == 0:
                continue
            sel = ~degenerate
            e = _rel(got[sel], ref[sel]) if sel.any() else 0.0
            if k == "qfrc_constraint":
                e = float(np.abs(got[sel] - ref[sel]).max() / max(np.abs(z["out_qfrc_smooth"]).max(), 1e-6)) if sel.any() else 0.0
        else:
            e = _rel(got, ref)
        if e > (10 * TOL if k in ("efc_force", "qfrc_constraint", "qacc", "qacc_warmstart", "qvel") else TOL):
            bad[k] = e
    assert not bad, f"{label}: fields beyond {TOL}: {bad}"


_UNSTABLE_CACHE = {}


def _oracle_unstable(path, z, mj):
    """Envs on which the reference ALGORITHM is discontinuous at this input: the oracle's own qacc moves by more than 100x the
    tolerance when qpos / qvel are perturbed by a few ulp (a line-search bracket decision or an `improved` test sitting on a
    rounding boundary -- DESIGN.md 4; seen with solvers stopped before convergence).  Only such an env may differ from the
    recording (see _check_excluding_discontinuities)."""
    if path in _UNSTABLE_CACHE:
        return _UNSTABLE_CACHE[path]
    B = z["in_qpos"].shape[0]
    unstable = np.zeros(B, dtype=bool)
    if z["out_efc_J"].shape[1]:
        kw = dict(nsteps=int(z["nsteps"]), act=z["in_act"] if "in_act" in z.files else None, fixed_iterations=_fixed(z), extra=_extra(z))
        base = util.run_oracle(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, **kw)[1].qacc.numpy()
        rng = np.random.RandomState(17)
        scale = max(np.abs(base).max(), 1e-6)
        for _ in range(4):
            pq = z["in_qpos"] * (1 + 4e-16 * rng.randn(*z["in_qpos"].shape))
            pv = z["in_qvel"] * (1 + 4e-16 * rng.randn(*z["in_qvel"].shape))
            q = util.run_oracle(mj, pq, pv, z["in_ctrl"], torch.float64, **kw)[1].qacc.numpy()
            unstable |= np.abs(q - base).max(1) / scale > 100 * TOL
    _UNSTABLE_CACHE[path] = unstable
    return unstable


def _check_excluding_discontinuities(path, z, mj, zs, sel, get, get_contact, label):
    """Every env must reproduce the recording, except envs that (a) differ in the integrated state AND (b) sit on a
    discontinuity of the reference algorithm (the oracle's own answer jumps under a 4e-16 input perturbation) -- at most 20 %
    of the fixture.  The excluded count goes to the warnings summary so that the round's test log shows it."""
    try:
        _check(zs, get, get_contact, label)
        return
    except AssertionError as exc:
        first = exc
    n = zs["in_qpos"].shape[0]
    differs = np.zeros(n, dtype=bool)
    for k in ("qpos", "qvel", "qacc"):
        ref = zs["out_" + k]
        differs |= np.abs(np.asarray(get(k), dtype=np.float64) - ref).reshape(n, -1).max(1) / max(np.abs(ref).max(), 1e-6) > 10 * TOL
    unstable = _oracle_unstable(path, z, mj)[sel]
    if not differs.any() or (differs & ~unstable).any() or differs.sum() > max(1, n // 5):
        raise first
    import warnings

    warnings.warn(f"{label} vs {os.path.basename(path)}: {int(differs.sum())}/{n} envs excluded -- they differ from the recording and the "
                  f"oracle itself is unstable there under 4e-16 input perturbation: envs {np.nonzero(differs)[0].tolist()}")
    keep = ~differs
    _check(_restrict(zs, keep), lambda k: np.asarray(get(k))[keep], lambda k: np.asarray(get_contact(k))[keep], label)


def _restrict(z, sel):
    if sel.all():
        return z
    batched = lambda k: k not in ("opt_keys", "opt_vals") and z[k].ndim and z[k].shape[0] == sel.size
    keys = z.files if hasattr(z, "files") else z.keys()
    return {k: (z[k][sel] if batched(k) else z[k]) for k in keys}


def _recorded_branch(z, mj, get_contact, get):
    """Convex manifolds: the reference breaks exact argmax ties (rectangular faces) by rounding noise, so an env can
    legitimately pick another corner than the recording did (tests/test_gpu_step.py::_check_manifolds; even this
    oracle, batched, does not always reproduce the per-env recording).  Returns the envs that took the recorded branch
    -- they must be the majority -- and the fixture restricted to them."""
    sel = np.ones(z["in_qpos"].shape[0], dtype=bool)
    if not ((np.asarray(mj.geom_type) >= 6).any() and z["contact_dist"].shape[1]):
        return sel, z
    sel = np.abs(np.asarray(get_contact("dist")) - z["contact_dist"]).max(1) < 1e-7
    for k in ("qvel", "qacc"):  # RK4: a tie flipped at an earlier stage only shows in the integrated state
        sel &= np.abs(np.asarray(get(k)) - z["out_" + k]).max(1) < 1e-7 * max(np.abs(z["out_" + k]).max(), 1.0)
    # RK4 runs the narrowphase four times per step: four times the ties to get through
    need = 0.25 if int(mj.opt.integrator) == 1 else 0.5
    assert sel.mean() >= need, f"only {int(sel.sum())}/{sel.size} envs took the recorded manifold branch"
    if sel.all():
        return sel, z
    batched = lambda k: k not in ("opt_keys", "opt_vals") and z[k].ndim and z[k].shape[0] == sel.size
    return sel, {k: (z[k][sel] if batched(k) else z[k]) for k in z.files}


@pytest.mark.parametrize("path", GOLDEN + GOLDEN_CPU_ONLY, ids=[os.path.basename(p)[:-4] for p in GOLDEN + GOLDEN_CPU_ONLY])
def test_oracle_matches_reference(path):
    z, mj = _load(path)
    mo, do = util.run_oracle(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, nsteps=int(z["nsteps"]),
                             act=z["in_act"] if "in_act" in z.files else None, fixed_iterations=_fixed(z), extra=_extra(z))
    ocon = {"dist": "contact_dist", "pos": "contact_pos", "frame": "contact_frame"}

    def get_contact(k):
        if k in ocon:
            return getattr(do, ocon[k]).numpy()
        B = z["in_qpos"].shape[0]
        if k == "efc_address":
            dims = mo.contact["dim"].numpy()[: mo.ncon]
            rows = np.where(dims == 1, 1, dims if mo.cone == 1 else 2 * (dims - 1))
            adr = mo.ne + mo.nf + mo.nl + np.concatenate([[0], np.cumsum(rows)[:-1]])
            return np.tile(adr, (B, 1))
        key = {"contact_dim": "dim"}.get(k, k)
        if getattr(mo, "topk", 0):   # max_contact_points: the slot's parameters are those of the candidate selected in that env
            return mo.contact[key][do.contact_src].numpy()
        v = mo.contact[key].numpy()
        return np.tile(v, (B,) + (1,) * v.ndim)

    sel, zs = _recorded_branch(z, mj, get_contact, lambda k: getattr(do, k).numpy())
    _check_excluding_discontinuities(path, z, mj, zs, sel, lambda k: getattr(do, k).numpy()[sel], lambda k: np.asarray(get_contact(k))[sel], "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference(path):
    z, mj = _load(path)
    mx, dc = util.run_cuda(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, nsteps=int(z["nsteps"]),
                           act=z["in_act"] if "in_act" in z.files else None, fixed_iterations=_fixed(z), extra=_extra(z))
    sel, zs = _recorded_branch(z, mj, lambda k: getattr(dc.contact, k).cpu().numpy(), lambda k: getattr(dc, k).cpu().numpy())
    _check_excluding_discontinuities(path, z, mj, zs, sel, lambda k: getattr(dc, k).cpu().numpy()[sel],
                                     lambda k: getattr(dc.contact, k).cpu().numpy()[sel], "cuda")


def test_fixed_iterations_recording_differs_from_early_exit():
    """The CG `fixed_iterations=True` recording must tell the two solver modes apart: the oracle's early-exit path stops at
    `tolerance` and lands ~4e-6 away from it, far beyond the 1e-9 bar the fixed path is held to above."""
    path = [p for p in GOLDEN if p.endswith("ant_zoo_cg_fixed_iterations.npz")][0]
    z, mj = _load(path)
    assert _fixed(z)
    _, early = util.run_oracle(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, fixed_iterations=False)
    _, fixed = util.run_oracle(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, fixed_iterations=True)
    scale = np.abs(z["out_qfrc_smooth"]).max()
    assert np.abs(early.qfrc_constraint.numpy() - z["out_qfrc_constraint"]).max() / scale > 1e-7
    assert np.abs(fixed.qfrc_constraint.numpy() - z["out_qfrc_constraint"]).max() / scale < 1e-9
