"""CPU tier: the oracle against every literal known answer the reference's own tests hold for this path
(SURVEY.md §8c, g1-g10).  These run without a GPU and without the reference tree."""

import numpy as np
import pytest
import torch

from mujoco_torch_b200 import mjcf
from oracle import collision as ocol
from oracle import constraint as ocon
from oracle import forward as of
from oracle import model as om
from oracle import omath as M
from oracle import smooth as osm
from oracle import solver as osol


def _model(xml, dtype=torch.float64, **opt):
    mj = mjcf.from_xml_string(xml)
    for k, v in opt.items():
        setattr(mj.opt, k, v)
    return mj, om.OModel(mj, dtype)


FREE_BODY = '<mujoco><option timestep="0.01"{flag}/><worldbody><body><joint type="free"/><geom size="0.1"/></body></worldbody></mujoco>'


def test_g1_free_fall():
    """reference test/smooth_test.py:181-208: one Euler step of a free body under gravity."""
    _, m = _model(FREE_BODY.format(flag=""))
    d = of.step(m, of.make_data(m, 1))
    np.testing.assert_array_almost_equal(d.qpos[0].numpy(), [0, 0, -9.81e-4, 1, 0, 0, 0], decimal=7)


def test_g1_free_fall_gravity_disabled():
    xml = '<mujoco><option timestep="0.01"><flag gravity="disable"/></option><worldbody><body><joint type="free"/><geom size="0.1"/></body></worldbody></mujoco>'
    _, m = _model(xml)
    d = of.step(m, of.make_data(m, 1))
    assert d.qpos[0].tolist() == [0, 0, 0, 1, 0, 0, 0]


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_g2_frictionloss_saturation(sign):
    """reference test/solver_test.py:77-108: 1-dof, J=[[1]], D=[1], aref=[0], frictionloss=[2], M=[[1]]."""
    xml = '<mujoco><worldbody><body><joint type="slide" frictionloss="2"/><geom size="0.1" mass="1"/></body></worldbody></mujoco>'
    _, m = _model(xml)
    assert (m.ne, m.nf, m.nl, m.nefc) == (0, 1, 0, 1)
    d = of.make_data(m, 1)
    one = torch.ones((1, 1), dtype=torch.float64)
    d.efc_J, d.efc_D, d.efc_aref, d.efc_frictionloss = one[:, None, :].clone(), one.clone(), 0 * one, 2 * one
    d.qM, d.qLD = one[:, None, :].clone(), one[:, None, :].clone()
    d.qfrc_smooth = d.qacc_smooth = d.qacc_warmstart = sign * 100 * one
    d = osol.solve(m, d)
    np.testing.assert_allclose(d.efc_force[0].numpy(), [-sign * 2], atol=1e-9)
    np.testing.assert_allclose(d.qfrc_constraint[0].numpy(), [-sign * 2], atol=1e-9)
    np.testing.assert_allclose(d.qacc[0].numpy(), [sign * 100 - sign * 2], atol=1e-9)


def test_g3_parallel_capsules():
    """reference test/collision_driver_test.py:179-204 (same fixture, same tolerances)."""
    xml = """<mujoco><worldbody>
      <body><joint type="free"/><geom fromto="-0.5 0.1 0.25 0.5 0.1 0.25" size="0.1" type="capsule"/></body>
      <body><joint type="free"/><geom fromto="-0.5 0.1 0.1 0.5 0.1 0.1" size="0.1" type="capsule"/></body>
    </worldbody></mujoco>"""
    _, m = _model(xml)
    d = of.make_data(m, 1)
    d = osm.kinematics(m, d)
    d = ocol.collision(m, d)
    assert m.ncon == 1
    np.testing.assert_allclose(d.contact_dist[0].numpy(), [-0.05], rtol=1e-6)
    np.testing.assert_allclose(d.contact_pos[0, 0].numpy(), [0.0, 0.1, (0.15 + 0.2) / 2.0], atol=1e-5)
    np.testing.assert_allclose(d.contact_frame[0, 0, 0].numpy(), [0, 0.0, -1.0], atol=1e-5)


def test_g5_ncon_ant(models_dir):
    """reference test/collision_driver_test.py:451-468: ant.xml has 60 contacts, 0 with contacts disabled."""
    from mujoco_torch_b200.mjmodel import MjModelLite

    mj = MjModelLite.load_npz(f"{models_dir}/ant.npz")
    assert om.OModel(mj).ncon == 60
    mj.opt.disableflags = int(mj.opt.disableflags) | 16
    assert om.OModel(mj).ncon == 0


def test_g5_parent_child_filter():
    """reference test/collision_driver_test.py:388-445: parent/child geoms never collide unless FILTERPARENT is disabled."""
    xml = """<mujoco><worldbody><body><joint type="free"/><geom size="0.1"/>
      <body pos="0.05 0 0"><joint type="hinge"/><geom size="0.1"/></body></body></worldbody></mujoco>"""
    _, m = _model(xml)
    assert m.ncon == 0
    _, m2 = _model(xml, disableflags=1 << 10)
    assert m2.ncon == 1


def _collide(xml):
    _, m = _model(xml)
    d = of.make_data(m, 1)
    d = osm.kinematics(m, d)
    return m, ocol.collision(m, d)


CAPSULES_BODY = """<worldbody>
  <body pos="0 0 0.54"><joint axis="1 0 0" type="free"/><geom fromto="-0.4 0 0 0.4 0 0" size="0.05" type="capsule"/></body>
  <body pos="0 0 0.54"><joint axis="1 0 0" type="free"/><geom fromto="-0.4 0 0 0.4 0 0" size="0.05" type="capsule"/></body>
  <body pos="0 0 0.54"><joint axis="1 0 0" type="free"/><geom fromto="-0.4 0 0 0.4 0 0" size="0.05" type="capsule"/></body>
</worldbody>"""


def test_g5_top_k_contacts():
    """reference test/collision_driver_test.py:470-524 (TopKContactTest): three coincident capsules give 3 contacts, 2 with
    `max_contact_points = 2`; the product's device_put sizes its contact slots the same way."""
    import mujoco_torch_b200 as mt

    m_all, d_all = _collide(f"<mujoco>{CAPSULES_BODY}</mujoco>")
    xml_k = f'<mujoco><custom><numeric data="2" name="max_contact_points"/></custom>{CAPSULES_BODY}</mujoco>'
    m_k, d_k = _collide(xml_k)
    assert m_all.ncon == 3 and d_all.contact_dist.shape == (1, 3)
    assert m_k.ncon == 2 and d_k.contact_dist.shape == (1, 2) and d_k.contact_src.shape == (1, 2)
    mx = mt.device_put(mjcf.from_xml_string(xml_k))
    assert (mx.topk, mx.ncand, mx.ncon) == (2, 3, 2) and mx.field_shape["scr_cand_dist"] == (3,) and mx.field_shape["contact_dist"] == (2,)
    assert mt.device_put(mjcf.from_xml_string(f"<mujoco>{CAPSULES_BODY}</mujoco>")).topk == 0
    # the K survivors are the closest candidates, in ascending distance (torch.topk(-dist))
    mixed = """<mujoco><custom><numeric data="2" name="max_contact_points"/></custom><worldbody>
      <geom type="plane" size="5 5 0.1"/>
      <body pos="0 0 0.03"><joint type="free"/><geom size="0.05"/></body>
      <body pos="1 0 0.045"><joint type="free"/><geom size="0.05"/></body>
      <body pos="2 0 0.2"><joint type="free"/><geom size="0.05"/></body>
      <body pos="3 0 0.01"><joint type="free"/><geom size="0.05"/></body></worldbody></mujoco>"""
    m_m, d_m = _collide(mixed)
    plane_sphere = [i for i in range(m_m.ncand) if int(m_m.contact["geom1"][i]) == 0]
    chosen = [int(m_m.contact["geom2"][int(i)]) for i in d_m.contact_src[0]]
    assert chosen == [4, 1] and m_m.ncand > m_m.ncon == 2 and len(plane_sphere) == 4     # geoms 4 (z 0.01) and 1 (z 0.03) are deepest
    np.testing.assert_allclose(d_m.contact_dist[0].numpy(), [0.01 - 0.05, 0.03 - 0.05], atol=1e-12)


def test_g4_box_plane_at_45_degrees():
    """reference test/collision_driver_test.py:261-282: a unit box rotated 45 degrees about x above a plane touches it along an
    edge: the first two manifold points penetrate, the last two do not."""
    m, d = _collide("""<mujoco><worldbody><geom size="40 40 40" type="plane"/>
      <body pos="0 0 0.7" euler="45 0 0"><joint axis="1 0 0" type="free"/><geom size="0.5 0.5 0.5" type="box"/></body></worldbody></mujoco>""")
    dist = d.contact_dist[0].numpy()
    assert dist.shape == (4,) and (dist[:2] < 0).all() and (dist[2:] > 0).all()
    np.testing.assert_allclose(dist[:2], 0.7 - 0.5 * np.sqrt(2.0), atol=1e-9)
    np.testing.assert_allclose(d.contact_frame[0, :2, 0].numpy(), [[0, 0, 1.0]] * 2, atol=1e-12)


def test_g4_box_box_face():
    """reference test/collision_driver_test.py:284-308: a small tilted box resting on a larger one: 4 contacts, all penetrating,
    at z ~ 0.39 (2 decimals), normal +z."""
    m, d = _collide("""<mujoco><worldbody>
      <body pos="0.0 1.0 0.2"><joint axis="1 0 0" type="free"/><geom size="0.2 0.2 0.2" type="box"/></body>
      <body pos="0.1 1.0 0.495" euler="0.1 -0.1 0"><joint axis="1 0 0" type="free"/><geom size="0.1 0.1 0.1" type="box"/></body>
    </worldbody></mujoco>""")
    assert d.contact_pos.shape == (1, 4, 3) and (d.contact_dist[0] < 0).all()
    np.testing.assert_array_almost_equal(d.contact_pos[0, :, 2].numpy(), [0.39] * 4, 2)
    np.testing.assert_array_almost_equal(d.contact_frame[0, :, 0, :].numpy(), [[0.0, 0.0, 1.0]] * 4)


def test_g4_box_box_edge():
    """reference test/collision_driver_test.py:310-336: an edge-edge box contact: exactly one penetrating point."""
    m, d = _collide("""<mujoco><worldbody>
      <body pos="-1.0 -1.0 0.2"><joint axis="1 0 0" type="free"/><geom size="0.2 0.2 0.2" type="box"/></body>
      <body pos="-1.0 -1.2 0.55" euler="0 45 30"><joint axis="1 0 0" type="free"/><geom size="0.1 0.1 0.1" type="box"/></body>
    </worldbody></mujoco>""")
    dist = d.contact_dist[0].numpy()
    assert (dist[:1] < 0).all() and (dist[1:] > 0).all()


def test_g4_capsule_box_edge():
    """reference test/collision_driver_test.py:230-254: a tilted capsule touching a box edge: two slots, the second one separated."""
    m, d = _collide("""<mujoco><worldbody>
      <body pos="0.5 0 0.55" euler="0 30 0"><joint axis="1 0 0" type="free"/><geom fromto="-0.6 0 0 0.6 0 0" size="0.05" type="capsule"/></body>
      <body><joint axis="1 0 0" type="free"/><geom size="0.5 0.5 0.5" type="box"/></body></worldbody></mujoco>""")
    assert d.contact_pos.shape == (1, 2, 3) and float(d.contact_dist[0, 1]) > 0 and float(d.contact_dist[0, 0]) < 0


def test_g6_segment_geometry():
    """reference test/math_test.py:131-188 (closest points between segments; literals)."""
    t = lambda *v: torch.tensor([v], dtype=torch.float64)
    a, b = M.closest_segment_to_segment_points(t(0, 0, -1), t(0, 0, 1), t(-1, 0, 0.5), t(1, 0, 0.5))
    np.testing.assert_allclose(a[0].numpy(), [0, 0, 0.5], atol=1e-5)
    np.testing.assert_allclose(b[0].numpy(), [0, 0, 0.5], atol=1e-5)
    # parallel, overlapping
    a, b = M.closest_segment_to_segment_points(t(0, 0, -1), t(0, 0, 1), t(1, 0, -1), t(1, 0, 1))
    assert abs(float((a - b).norm()) - 1.0) < 1e-5
    b1, c1 = M.orthogonals(M.normalize(t(1.0, 2.0, 3.0)))
    n = M.normalize(t(1.0, 2.0, 3.0))
    assert abs(float(M.dot(b1, c1))) < 1e-12 and abs(float(M.dot(b1, n))) < 1e-12 and abs(float(M.dot(c1, n))) < 1e-12


@pytest.mark.parametrize("refsafe_disabled", [True, False])
def test_g7_kbi_refsafe(refsafe_disabled):
    """reference test/constraint_test.py:128-146: k = 1/(dmax^2 tc^2) with refsafe disabled, 1/(dmax^2 (2dt)^2) enabled."""
    xml = '<mujoco><option timestep="0.01"/><worldbody><body><joint type="slide"/><geom size="0.1"/></body></worldbody></mujoco>'
    _, m = _model(xml, disableflags=(1 << 12) if refsafe_disabled else 0)
    tc, dmax = 0.005, 0.99
    solref = torch.tensor([tc, 1.0], dtype=torch.float64)
    solimp = torch.tensor([0.9, dmax, 0.001, 0.5, 2.0], dtype=torch.float64)
    k, b, imp = ocon.kbi(m, solref, solimp, torch.tensor(1.0, dtype=torch.float64))
    tce = tc if refsafe_disabled else 2 * 0.01
    np.testing.assert_allclose(float(k), 1 / (dmax * dmax * tce * tce), rtol=1e-12)


def test_g8_euler_without_damping():
    """reference test/forward_test.py:128-140: EULERDAMP disabled, qvel=qacc=1 -> qvel = 1 + dt."""
    xml = """<mujoco><option timestep="0.002"><flag eulerdamp="disable" gravity="disable"/></option>
      <worldbody><body><joint type="slide" damping="5"/><geom size="0.1"/></body></worldbody></mujoco>"""
    _, m = _model(xml)
    d = of.make_data(m, 1)
    d.qvel[:] = 1.0
    d.qacc, d.act_dot = torch.ones((1, 1), dtype=torch.float64), torch.zeros((1, 0), dtype=torch.float64)
    d.qfrc_smooth = d.qfrc_constraint = torch.zeros((1, 1), dtype=torch.float64)
    d = of.euler(m, d)
    np.testing.assert_allclose(d.qvel[0].numpy(), [1 + 0.002], rtol=1e-12)


def test_g9_slide_limit_sign():
    """reference test/constraint_test.py:108-125: qpos=2.0 beyond range 1.8 -> J[0,0] = -1; an un-ranged joint has no row."""
    xml = """<mujoco><worldbody><body><joint type="slide" range="-1.8 1.8"/><geom size="0.1"/>
      <body><joint type="slide" axis="0 1 0"/><geom size="0.1"/></body></body></worldbody></mujoco>"""
    _, m = _model(xml)
    assert m.nl == 1 and m.nefc == 1
    d = of.make_data(m, 1)
    d.qpos[0, 0] = 2.0
    d = of.position(m, d)
    assert d.efc_J[0, 0, 0].item() == -1.0


@pytest.mark.parametrize("cone,condim,rows", [(0, 3, 4), (1, 3, 3), (0, 4, 6), (1, 4, 4), (0, 1, 1), (1, 6, 6), (0, 6, 10)])
def test_g10_rows_per_contact(cone, condim, rows):
    """reference test/constraint_test.py:536-549, 578-591: elliptic rows = condim, pyramidal = 2(condim-1)."""
    xml = f"""<mujoco><worldbody><geom type="plane" size="1 1 1" condim="{condim}"/>
      <body pos="0 0 0.05"><joint type="free"/><geom size="0.1" condim="{condim}"/></body></worldbody></mujoco>"""
    _, m = _model(xml, cone=cone)
    assert m.ncon == 1 and m.nefc == rows
    d = of.step(m, of.make_data(m, 2))
    assert d.efc_J.shape == (2, rows, 6) and torch.isfinite(d.qpos).all()


def test_every_recording_has_a_recipe_and_every_recipe_a_recording():
    """tests/golden/*.npz and the CASES list of tests/golden/make_golden.py (the committed generating script) name the same
    fixtures, and every recording says which model / options / step count it was made with."""
    import ast
    import glob
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "golden", "make_golden.py")).read()
    tree = ast.parse(src)
    cases = next(n.value for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "CASES")
    names = sorted(ast.literal_eval(e.elts[0]) for e in cases.elts)
    files = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(here, "golden", "*.npz")))
    assert names == files, (set(names) ^ set(files))
    for f in files:
        z = np.load(os.path.join(here, "golden", f + ".npz"))
        assert {"in_qpos", "in_qvel", "in_ctrl", "nsteps", "opt_keys", "opt_vals", "model", "out_qpos", "out_qacc"} <= set(z.files), f
        assert z["in_qpos"].shape[0] >= 32 and z["out_qpos"].shape == z["in_qpos"].shape, f
