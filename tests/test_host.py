"""CPU tier: C-ABI library surface, layout agreement, host-side static precompute vs the oracle's, Data container API."""

import ctypes
import os
import re
import types as pytypes

import numpy as np
import pytest
import torch

import mujoco_torch_b200 as mt
from mujoco_torch_b200 import layout as L
from mujoco_torch_b200 import ops
from oracle import model as om
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = ["cartpole", "humanoid", "halfcheetah", "ant", "ant_zoo", "hopper", "walker2d", "swimmer", "pendula", "frictionloss_dof", "convex",
          "equality", "satellite_small", "satellite_large", "equality_bodies", "sink", "shapes", "stack"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mjt.h")).read()
    declared = sorted(set(re.findall(r"\b(mjt_[a-z_]+)\s*\(", hdr)))
    assert declared, "no declarations found in include/mjt.h"
    lib = ops.lib()
    for sym in declared:
        assert hasattr(lib, sym), f"libmjt_b200.so does not export {sym}"
    assert sorted(ops.EXPORTS) == declared
    assert lib.mjt_version() >= 1000
    assert isinstance(lib.mjt_last_error(), bytes)


def test_layout_header_in_sync():
    path = os.path.join(ROOT, "mujoco_torch_b200", "csrc", "mjt_layout.h")
    assert open(path).read() == L.header_text(), "run `python -m mujoco_torch_b200.layout`"


def test_no_cpu_fallback():
    mx = mt.device_put(util.load_model("cartpole"), dtype=torch.float64)
    d = mt.make_data(mx)
    with pytest.raises((NotImplementedError, RuntimeError)):
        mt.step(mx, d)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mujoco_torch_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.parametrize("name", MODELS)
def test_static_precompute_matches_oracle(name):
    """candidate pairs / final contact order / mixed solver params / efc sizes: product packer vs oracle restatement."""
    mj = util.load_model(name)
    mx = mt.device_put(mj, dtype=torch.float64)
    mo = om.OModel(mj, torch.float64)
    assert (mx.ne, mx.nf, mx.nl, mx.ncon, mx.nefc) == (mo.ne, mo.nf, mo.nl, mo.ncon, mo.nefc)
    if mo.ncon:
        c = mo.contact
        assert mx.con["g1"].tolist() == c["geom1"].tolist() and mx.con["g2"].tolist() == c["geom2"].tolist()
        assert mx.con["dim"].tolist() == c["dim"].tolist()
        for k_prod, k_or in (("include", "includemargin"), ("friction", "friction"), ("solref", "solref"),
                             ("solreffriction", "solreffriction"), ("solimp", "solimp")):
            np.testing.assert_allclose(mx.con[k_prod], c[k_or].numpy(), rtol=1e-14, atol=0)
    # blob header round trip
    w = mx.blob_np
    assert w[L.h_index("nv")] == mj.nv and w[L.h_index("nefc")] == mo.nefc and w[L.h_index("data_words")] == mx.data_words
    assert w[L.h_index("blob_words")] + w[L.h_index("solve_blob_words")] == w.size


def test_sizes_of_benchmark_configs():
    """SURVEY.md §8(d) config table."""
    expect = {"cartpole": (2, 2, 1, 0, 0), "halfcheetah": (9, 9, 6, 16, 70), "humanoid": (28, 27, 21, 8, 53), "ant": (8, 8, 8, 60, 248)}
    for name, (nq, nv, nu, ncon, nefc) in expect.items():
        mj = util.load_model(name)
        mx = mt.device_put(mj, dtype=torch.float64)
        assert (mj.nq, mj.nv, mj.nu, mx.ncon, mx.nefc) == (nq, nv, nu, ncon, nefc), name
    mj = util.load_model("ant_zoo", cone=1, solver=1)
    mx = mt.device_put(mj, dtype=torch.float32)
    assert (mj.nq, mj.nv, mj.nu, mx.ncon, mx.nefc) == (15, 14, 8, 69, 215)


def test_data_container_api():
    mx = mt.device_put(util.load_model("humanoid"), dtype=torch.float64)
    d = mt.make_data(mx)
    assert tuple(d.batch_size) == () and d.qpos.shape == (28,) and d.contact.dist.shape == (8,)
    assert d.efc_J.shape == (53, 27) and (d.ne, d.nf, d.nl, d.ncon, d.nefc) == (0, 0, 21, 8, 53)
    b = d.expand(4).clone()
    assert tuple(b.batch_size) == (4,) and b.qpos.shape == (4, 28) and b.contact.frame.shape == (4, 8, 3, 3)
    b2 = b.replace(qvel=torch.ones(4, 27, dtype=torch.float64))
    assert b2.qvel.sum() == 4 * 27 and b.qvel.sum() == 0 and b2.qpos is b.qpos
    b2.update_(ctrl=torch.full((4, 21), 0.5, dtype=torch.float64))
    assert float(b2.ctrl[0, 0]) == 0.5
    b3 = b.tree_replace({"contact.dist": torch.ones(4, 8, dtype=torch.float64)})
    assert b3.contact.dist.sum() == 32 and b.contact.dist.sum() == 0
    assert b[1].qpos.shape == (28,) and tuple(b[1:3].batch_size) == (2,)
    mask = torch.tensor([True, False, True, False])
    fresh = mt.make_data(mx).expand(2).clone().replace(qvel=torch.full((2, 27), 7.0, dtype=torch.float64))
    b[mask] = fresh
    assert float(b.qvel[0, 0]) == 7.0 and float(b.qvel[1, 0]) == 0.0
    s = torch.stack([d, d, d])
    assert tuple(s.batch_size) == (3,) and s.contact.pos.shape == (3, 8, 3)
    f32 = s.to(torch.float32)
    assert f32.qpos.dtype == torch.float32 and f32.contact.geom1.dtype == torch.int64


def test_graphed_step_refuses_what_it_cannot_capture():
    """GraphedStep is CUDA-only like step(): wrong model type, unbatched or CPU Data and nsteps < 1 fail at construction."""
    mx = mt.device_put(util.load_model("cartpole"), dtype=torch.float64)
    d = mt.make_data(mx).expand(4).clone()
    with pytest.raises(TypeError):
        mt.GraphedStep(object(), d)
    with pytest.raises(ValueError):
        mt.GraphedStep(mx, mt.make_data(mx))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        mt.GraphedStep(mx, d)
    with pytest.raises(ValueError):
        mt.GraphedStep(mx, d, 0)


def test_masked_assignment_materialises_expanded_fields():
    """The RL reset pattern `d[mask] = fresh` on a Data whose fields are stride-0 expansions (an un-cloned expand(); the
    model-constant Contact fields of every step output): the write lands in the selected envs only and never in the shared source."""
    mx = mt.device_put(util.load_model("humanoid"), dtype=torch.float64)
    one = mt.make_data(mx)
    b = one.expand(4)
    assert b.qvel.stride()[0] == 0
    fresh = mt.make_data(mx).expand(2).clone().replace(qvel=torch.full((2, 27), 3.0, dtype=torch.float64))
    fresh.contact.friction.fill_(0.25)
    b[torch.tensor([False, True, False, True])] = fresh
    assert b.qvel[:, 0].tolist() == [0.0, 3.0, 0.0, 3.0] and float(one.qvel.abs().max()) == 0.0
    assert float(b.contact.friction[1].min()) == 0.25 and float(b.contact.friction[0].max()) != 0.25
    assert float(one.contact.friction.max()) != 0.25
    b[0:2] = fresh
    assert b.qvel[:, 0].tolist() == [3.0, 3.0, 0.0, 3.0]


def test_data_is_a_pytree_for_vmap():
    mx = mt.device_put(util.load_model("cartpole"), dtype=torch.float64)
    d = mt.make_data(mx).expand(5).clone()
    out = torch.vmap(lambda x: x.replace(qvel=x.qvel + x.qpos.sum()))(d)
    assert isinstance(out, mt.Data) and out.qvel.shape == (5, 2)


def test_unsupported_features_raise():
    mj = util.load_model("cartpole")
    mj.opt.integrator = 2  # IMPLICIT (full) is not supported by the reference either
    with pytest.raises(NotImplementedError):
        mt.device_put(mj)
    mj = util.load_model("cartpole")
    mj.opt.solver = 0  # PGS
    with pytest.raises(NotImplementedError):
        mt.device_put(mj)


def test_out_of_scope_models_raise_at_device_put():
    """sparse-M models (nv >= 60), collisions the reference has no function for fail
    loudly at device_put -- never at step time, never silently."""
    from mujoco_torch_b200 import mjcf

    links = "".join(f'<body pos="0 0 -0.1"><joint type="hinge" axis="0 1 0"/><geom type="capsule" size="0.02" fromto="0 0 0 0 0 -0.1"/>'
                    for _ in range(60)) + "</body>" * 60
    big = mjcf.from_xml_string(f"<mujoco><worldbody>{links}</worldbody></mujoco>")
    assert big.nv == 60
    with pytest.raises(NotImplementedError, match="sparse"):
        mt.device_put(big)
    cyl = mjcf.from_xml_string('<mujoco><worldbody><geom type="plane" size="1 1 .1"/><body pos="0 0 1"><freejoint/>'
                               '<geom type="cylinder" size="0.1 0.2"/></body></worldbody></mujoco>')
    with pytest.raises(NotImplementedError, match="collisions not implemented"):
        mt.device_put(cyl)
    frc = mjcf.from_xml_string('<mujoco><worldbody><body pos="0 0 1"><joint type="hinge"/><geom type="sphere" size="0.1"/>'
                               '<site name="s"/></body></worldbody><sensor><force site="s"/></sensor></mujoco>')
    # no sensor type is refused: the reference has none it rejects either (types without a branch keep their slots, sensor.py:425)
    assert mt.device_put(frc).nsensordata == 3


def test_live_tile_caps_and_blob_header():
    """Solver live tile: only models whose tile would at least halve the rows get one; the caps travel in the
    blob header, and the two hidden flag fields are the last Data fields of the layout."""
    caps = {n: mt.device_put(util.load_model(n)).live_caps for n in ("humanoid", "halfcheetah", "hopper", "ant", "ant_zoo", "stack")}
    assert caps["humanoid"] == (0, 0) and caps["halfcheetah"] == (0, 0) and caps["hopper"] == (0, 0)
    for n in ("ant", "ant_zoo", "stack"):
        mx = mt.device_put(util.load_model(n))
        rcap, jcap = mx.live_caps
        assert 0 < 2 * rcap <= mx.nefc and rcap % 16 == 0 and 0 < jcap <= mx.sup["jc_words"]
        hdr = np.asarray(mx.blob[: len(L.H_SCALARS)])
        assert int(hdr[L.H_SCALARS.index("live_rcap")]) == rcap and int(hdr[L.H_SCALARS.index("live_jcap")]) == jcap
        assert mx.field_shape["scr_live"] == (len(mx.src),) and mx.field_shape["scr_ovf"] == ()
    names = [n for n, _ in L.D_FIELDS]
    assert names.index("scr_ovf") == names.index("scr_live") + 1 and {"scr_live", "scr_ovf"} <= L.HIDDEN_FIELDS


def test_bench_algorithmic_counts_match_survey():
    """bench.py's roofline numerators: SURVEY.md §8(d) gives 1 712 B (f64) of minimal state I/O and ~0.17 MFLOP per Humanoid
    env-step; HalfCheetah 568 B."""
    import bench

    mj = util.load_model("humanoid")
    assert bench.algorithmic_bytes(mj, 8) == 1712 and bench.algorithmic_bytes(mj, 4) == 856
    assert bench.algorithmic_bytes(util.load_model("halfcheetah"), 8) == 568
    fl = bench.algorithmic_flops(mj, mt.device_put(mj))
    assert 1.6e5 < fl["total"] < 1.9e5 and fl["total"] == pytest.approx(sum(fl["by_kernel"].values()))
    assert 1.3e5 < fl["by_kernel"]["k_solve"] < 1.6e5      # solver + factor_m / qacc_smooth: the dominant kernel's numerator
    mj.opt.integrator = 1   # RK4: four forwards
    assert bench.algorithmic_flops(mj, mt.device_put(mj))["total"] == pytest.approx(4 * fl["total"])
    # measured DRAM traffic of the committed ncu capture is keyed by kernel, model, dtype, batch and output mode
    t = bench.measured_traffic("k_solve", "humanoid", "f64", 65536, "state")
    assert t is not None and t > 0 and bench.measured_traffic("k_solve", "humanoid", "f64", 4096, "state") is None


def test_device_get_into_follows_the_reference_rules():
    """reference device.device_get_into :1119-1205: one struct keeps the batch dim, a list gets one env each; rotation matrices
    travel flattened, efc_J / actuator_moment / qM / qLD are skipped, contact_dim lands in `dim`, shape mismatches are skipped."""
    import types as pytypes

    mx = mt.device_put(util.load_model("humanoid"))
    B = 3
    d = mt.make_data(mx).expand(B).clone()
    d = d.replace(qpos=d.qpos + torch.arange(B, dtype=d.qpos.dtype)[:, None], xmat=torch.randn(B, mx.nbody, 3, 3, dtype=d.qpos.dtype))
    d.contact.frame.copy_(torch.randn_like(d.contact.frame))

    def blank(batch):
        r = pytypes.SimpleNamespace(contact=pytypes.SimpleNamespace())
        r.qvel = np.zeros(batch + (mx.nv + 1,))      # wrong shape on purpose: must be left alone
        return r

    one = blank((B,))
    mt.device_get_into(one, d)
    assert one.qpos.shape == (B, mx.nq) and np.array_equal(one.qpos, d.qpos.numpy())
    assert one.xmat.shape == (B, mx.nbody, 9) and np.array_equal(one.xmat, d.xmat.reshape(B, mx.nbody, 9).numpy())
    assert one.contact.frame.shape == (B, mx.ncon, 9) and one.contact.dim.shape == (B, mx.ncon)
    assert not hasattr(one.contact, "contact_dim")
    for skipped in ("efc_J", "actuator_moment", "qM", "qLD"):
        assert not hasattr(one, skipped)
    assert one.qvel.shape == (B, mx.nv + 1) and not one.qvel.any()
    assert (one.ncon, one.nefc) == (mx.ncon, mx.nefc)

    many = [blank(()) for _ in range(B)]
    mt.device_get_into(many, d)
    for i, r in enumerate(many):
        assert r.qpos.shape == (mx.nq,) and np.array_equal(r.qpos, d.qpos[i].numpy())
        assert r.contact.frame.shape == (mx.ncon, 9)
    with pytest.raises(ValueError, match="doesn't match value batch size"):
        mt.device_get_into(many[:2], d)
    with pytest.raises(ValueError, match="unrecognizable batch dimension"):
        mt.device_get_into([blank(())], mt.make_data(mx))


def test_device_put_accepts_data_option_and_statistic_structs():
    """reference device.device_put :1029-1112 maps MjModel / MjData / MjOption / MjStatistic; ours is duck-typed on MuJoCo's
    attribute names.  An MjData-like struct (flattened rotation matrices, `contact.dim`, current ncon = 0 after a reset, like
    benchmarks/_helpers.py:28-33) becomes a Data whose state feeds `step`; options and statistics become Option / Statistic."""
    mj = util.load_model("humanoid")
    mx = mt.device_put(mj)
    rng = np.random.RandomState(1)
    md = pytypes.SimpleNamespace(qpos=np.asarray(mj.qpos0) + 0.01 * rng.randn(mj.nq), qvel=0.1 * rng.randn(mj.nv), act=np.zeros(0), time=0.25,
                                 ctrl=rng.uniform(-1, 1, mj.nu), qacc_warmstart=np.zeros(mj.nv), qfrc_applied=np.zeros(mj.nv),
                                 xfrc_applied=np.zeros((mj.nbody, 6)), mocap_pos=np.zeros((0, 3)), mocap_quat=np.zeros((0, 4)),
                                 eq_active=np.zeros(0, np.uint8), sensordata=np.zeros(0), xmat=np.tile(np.eye(3).reshape(9), (mj.nbody, 1)),
                                 xpos=np.zeros((mj.nbody, 3)), ncon=0, nefc=0,
                                 contact=pytypes.SimpleNamespace(dist=np.zeros(0), pos=np.zeros((0, 3)), frame=np.zeros((0, 9)), dim=np.zeros(0, np.int32),
                                                                 geom1=np.zeros(0, np.int32), geom2=np.zeros(0, np.int32)))
    for dtype in (None, torch.float32):
        d = mt.device_put(md, dtype=dtype)
        want = torch.get_default_dtype() if dtype is None else dtype
        assert isinstance(d, mt.Data) and d.batch_size == torch.Size([]) and d.qpos.dtype == want
        assert d.xmat.shape == (mj.nbody, 3, 3) and d.contact.dist.shape == (0,) and d.contact.frame.shape == (0, 3, 3)
        assert d.contact.contact_dim.dtype == torch.int32 and float(d.time) == 0.25 and d.ncon == 0
        assert torch.allclose(d.ctrl, torch.as_tensor(md.ctrl).to(want))
    # every input of the step op is there, with the shapes make_data gives
    ref = mt.make_data(mx)
    from mujoco_torch_b200 import ops
    for name in ops._STATE_ARGS:
        assert getattr(d, name).shape == getattr(ref, name).shape, name
    batch = torch.stack([d, d])
    assert batch.qpos.shape == (2, mj.nq) and batch.contact.dist.shape == (2, 0)
    opt = mt.device_put(mj.opt)
    assert isinstance(opt, mt.Option) and opt.timestep == float(mj.opt.timestep) and opt.iterations == int(mj.opt.iterations)
    assert isinstance(mt.device_put(mj.stat), mt.Statistic) and mt.device_put(mj.stat).meaninertia == float(mj.stat.meaninertia)
    with pytest.raises(NotImplementedError):
        mt.device_put(object())


def test_env_minor_inputs_need_no_copy():
    """The op reads env-minor (n, B) arrays.  A batched Data cloned from `make_data(...).expand(B)` and every output of a step
    already have that memory layout, so `ops._env_minor` hands their storage to the kernels as it is; only a row-major tensor
    the caller puts into the Data is transposed (once per step -- there is no cache that could go stale)."""
    import mujoco_torch_b200 as mt

    mx = mt.device_put(util.load_model("humanoid"), dtype=torch.float32)
    d = mt.make_data(mx).expand(8).clone()
    for name in ("qpos", "qvel", "qfrc_applied", "xfrc_applied", "ctrl"):
        t = getattr(d, name)
        flat = t.reshape(t.shape[0], -1)
        e = ops._env_minor(flat, True, torch.float32)
        assert e.is_contiguous() and e.shape == (flat.shape[1], 8)
        assert e.data_ptr() == t.data_ptr(), f"{name}: a cloned batch should already be env-minor"
    assert torch.equal(d.qpos, mt.make_data(mx).qpos.expand(8, -1))            # same values, same logical shape
    d.xfrc_applied[3, 2, 1] = 5.0                                                 # in-place writes are seen (no copy in between)
    assert float(ops._env_minor(d.xfrc_applied.reshape(8, -1), True, torch.float32)[2 * 6 + 1, 3]) == 5.0
    x = torch.zeros(8, 6)                                                         # row-major: transposed copy
    a = ops._env_minor(x, True, torch.float32)
    assert a.shape == (6, 8) and a.is_contiguous() and a.data_ptr() != x.data_ptr()
    assert not hasattr(ops, "_env_minor_cached")


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys, the
    metric / unit / workload of our own arm, zero transfer bytes; non-zero ranks of a torchrun launch print nothing."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--model", "halfcheetah", "--cpu-batch", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=util.ROOT).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == "env-steps/sec" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["config"]["workload"].startswith("halfcheetah")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    silent = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=util.ROOT, env=env)
    assert silent.returncode == 0 and silent.stdout.strip() == ""


def test_mjcf_site_equalities_and_comment_tables():
    """MJCF surface of the reference's own test files: site-based weld / connect (`equality.xml`: eq_objtype = mjOBJ_SITE, ids
    are site ids, only the torque scale in eq_data) and comments containing "--" (`satellite_small.xml`), which expat rejects."""
    xml = """<mujoco><!-- table  | a -- b |
      ------+------ -->
      <worldbody>
        <body name="a" pos="0 0 1"><freejoint/><geom size="0.1"/><site name="sa" pos="0.1 0 0"/></body>
        <body name="b" pos="0 0 2"><freejoint/><geom size="0.1"/><site name="sb" pos="0 0.1 0"/></body>
      </worldbody>
      <equality>
        <weld name="w" site1="sa" site2="sb" torquescale="2" active="false"/>
        <connect name="c" site1="sb" site2="sa"/>
        <connect name="cb" body1="a" body2="b" anchor="0 0 0.5"/>
      </equality></mujoco>"""
    mj = mt.mjcf.load(xml)
    assert mj.eq_type.tolist() == [1, 0, 0] and mj.eq_objtype.tolist() == [6, 6, 1]
    assert mj.eq_obj1id.tolist() == [0, 1, 1] and mj.eq_obj2id.tolist() == [1, 0, 2] and mj.eq_active0.tolist() == [0, 1, 1]
    assert mj.eq_data[0, 10] == 2.0 and not mj.eq_data[0, :10].any() and not mj.eq_data[1].any()
    assert np.allclose(mj.eq_data[2, :3], [0, 0, 0.5]) and np.allclose(mj.eq_data[2, 3:6], [0, 0, -0.5])
    mt.device_put(mj)   # ids are valid body ids: accepted (indexed as bodies, like the reference)
    with pytest.raises(ValueError, match="site1 and site2"):
        mt.mjcf.load(xml.replace('site1="sb" site2="sa"', 'site1="sb" body2="a"'))
