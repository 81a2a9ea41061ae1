"""Generate golden input/output vectors by running the REFERENCE'S OWN step code on our compiled models.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz (run in the build container only)

The reference (`/root/reference/mujoco_torch/_src/*.py`, unmodified, imported from where it lies) needs `mujoco`,
`tensordict`, `pyvers` and `trimesh`; none exist in this image, so the minimal stand-ins of `oracle/refshim/` are put
on sys.path first (see its README).  `mujoco_torch/__init__.py` is bypassed (it monkey-patches torch internals for
torch.compile and imports the renderer); the physics modules under `_src` are imported directly.  Each environment is
stepped single-env, eager, float64 -- the reference's own eager semantics.
"""

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("MJT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, ROOT)
pkg = types.ModuleType("mujoco_torch")
pkg.__path__ = [os.path.join(REF, "mujoco_torch")]
sys.modules["mujoco_torch"] = pkg

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

import mujoco  # noqa: E402  (the shim)
from mujoco_torch._src import device as ref_device  # noqa: E402
from mujoco_torch._src import forward as ref_forward  # noqa: E402
from mujoco_torch._src import io as ref_io  # noqa: E402

from mujoco_torch._src import mesh as ref_mesh  # noqa: E402
from mujoco_torch._src import types as ref_types  # noqa: E402

from mujoco_torch_b200 import convex as prod_convex  # noqa: E402
from tests import util  # noqa: E402


# The reference builds its convex-geom tables with trimesh (absent here).  Like the MJCF->MjModel boundary, this
# preprocessing is therefore supplied by our own host code: the reference's collision code then runs unmodified on it.
def _mesh_get(m):
    tabs = prod_convex.build(m._lite)
    pick = lambda f: [None if t is None else f(t) for t in tabs]
    return {"geom_convex_face": pick(lambda t: t.face), "geom_convex_vert": pick(lambda t: t.vert),
            "geom_convex_edge": pick(lambda t: t.edge), "geom_convex_facenormal": pick(lambda t: t.norm)}


def _mesh_convex(m, data_id):
    tabs = prod_convex.build(m._lite)
    g = next(i for i in range(m.ngeom) if int(m.geom_dataid[i]) == data_id)
    t = tabs[g]
    return ref_types.ConvexMesh(torch.tensor(t.vert), torch.tensor(t.vert[t.face]), torch.tensor(t.norm), torch.tensor(t.edge),
                                torch.zeros((len(t.edge), 2, 3)), batch_size=[])


ref_mesh.get = _mesh_get
ref_mesh.convex = _mesh_convex
ref_device.mesh.get = _mesh_get
ref_device.mesh.convex = _mesh_convex

OUT_FIELDS = ["qpos", "qvel", "qacc", "qacc_warmstart", "act", "time", "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis",
              "geom_xpos", "geom_xmat", "subtree_com", "cinert", "crb", "cdof", "cdof_dot", "cvel", "qM", "qLD", "actuator_length",
              "actuator_moment", "actuator_velocity", "actuator_force", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth",
              "qacc_smooth", "qfrc_constraint", "efc_J", "efc_D", "efc_aref", "efc_frictionloss", "efc_force", "sensordata"]
CONTACT_FIELDS = ["dist", "pos", "frame", "includemargin", "friction", "solref", "solreffriction", "solimp", "contact_dim", "geom1",
                  "geom2", "efc_address"]

# (fixture name, model, option overrides, envs, drop height, steps)
CASES = [
    ("cartpole", "cartpole", {}, 4, 0.0, 1),
    ("humanoid", "humanoid", {}, 6, 0.4, 1),
    # NOTE: ImplicitFast cannot be recorded for actuated models: reference derivative.py:48 does
    # torch.where(dyn_mask (nu,), d.act (na,), ctrl (nu,)) which raises for na != nu (Humanoid: nu=21, na=0).
    ("frictionloss_dof_implicitfast", "frictionloss_dof", {"integrator": 3}, 3, 0.0, 1),
    ("humanoid_converged", "humanoid", {"iterations": 100, "ls_iterations": 50}, 4, 0.4, 1),
    ("halfcheetah", "halfcheetah", {}, 6, 0.6, 1),
    ("halfcheetah_3steps", "halfcheetah", {}, 3, 0.6, 3),
    ("ant", "ant", {}, 3, 0.0, 1),
    ("ant_zoo_cg_elliptic", "ant_zoo", {"cone": 1, "solver": 1}, 4, 0.5, 1),
    ("hopper", "hopper", {}, 4, 0.6, 1),
    ("walker2d_rk4", "walker2d", {"integrator": 1}, 3, 0.8, 1),
    ("swimmer", "swimmer", {}, 3, 0.0, 1),
    ("pendula", "pendula", {}, 3, 0.0, 1),
    ("frictionloss_dof", "frictionloss_dof", {}, 3, 0.0, 1),
    ("equality_bodies", "equality_bodies", {}, 4, 0.0, 1),
    ("sink", "sink", {}, 6, -3.0, 1),
    ("sink_elliptic", "sink", {"cone": 1}, 4, -3.0, 1),
    ("convex", "convex", {}, 4, -1.0, 1),
    ("shapes", "shapes", {}, 4, -1.0, 1),
    ("shapes_elliptic", "shapes", {"cone": 1}, 6, -1.0, 1),
    ("stack_rk4", "stack", {}, 8, -2.0, 1),
    # the reference's own equality.xml (site-based weld / connect, inactive; joint equality; box, sphere and plane contacts).

    ("equality_sites", "equality", {}, 4, 0.3, 1),
    # the reference's control-moment-gyro satellites (free-floating, no contacts: smooth dynamics + actuation only)
    ("satellite_small", "satellite_small", {}, 3, 0.0, 2),
    ("satellite_large", "satellite_large", {}, 3, 0.0, 2),
    # step(..., fixed_iterations=True): exactly `iterations` solver and `ls_iterations` line-search iterations, no termination
    # test (solver.py:484-487, 536-537).  "fixed_iterations" is not a model option: it is passed to step() and stripped here.
    ("halfcheetah_fixed_iterations", "halfcheetah", {"iterations": 6, "ls_iterations": 5, "fixed_iterations": 1}, 4, 0.6, 1),
    # CG run past convergence: the early-exit path stops at `tolerance` (2e-7 relative away in qacc), the fixed path keeps going.
    # (A short unconverged CG run is not usable as a pin: the oracle's own answer moves by 3e-3 under 2e-16 input perturbations.)
    ("ant_zoo_cg_fixed_iterations", "ant_zoo", {"cone": 1, "solver": 1, "iterations": 60, "ls_iterations": 20, "fixed_iterations": 1}, 4, 0.2, 1),
    # muscle actuators: dynamics, length / velocity gain curves and passive bias (support.py:196-296), on a joint, on a tendon,
    # and as a gain-only general actuator (drop -4 = tests/util.muscle_state), Euler and ImplicitFast
    ("muscle", "muscle", {}, 32, -4.0, 3),
    ("muscle_implicitfast", "muscle", {"integrator": 3}, 32, -4.0, 2),
    # max_contact_points = 6 of 19 candidate contacts: data-dependent contact selection, parameters and geom ids per env
    # (collision_driver.py:822-840); loose pile of capsules and a ball on the floor, two steps
    ("capsules_topk", "capsules_topk", {}, 32, -1.0, 2),
    # height-field narrowphase: sphere, capsule and box scattered over a bumpy 9 x 11 terrain (drop -5 = tests/util.terrain_state);
    # the reference walks the sub-grid in Python per env (collision_hfield.py), which is exactly how it is recorded here
    ("terrain", "terrain", {}, 32, -5.0, 1),
    # sensors the reference accepts without feeding them: force / torque / subtreelinvel / subtreeangmom read make_data's zeros,
    # touch / framelinacc / frameangacc have no branch and keep their slots, tendonactuatorfrc is computed (sensor.py:356-426)
    ("sensors_extra", "sensors_extra", {}, 32, 0.5, 2),
    # mjtDisableBit handling (opt.disableflags): every bit the reference tests somewhere on the path, on the scene that has the
    # feature -- contacts / limits + friction loss / gravity + warm start + ctrl clamping / springs + dampers + actuation /
    # the whole constraint stage / sensors + refsafe (sink), equalities (equality_bodies), parent-child filtering (raw ant: more
    # candidate pairs), and Humanoid with its XML's eulerdamp="disable" cleared (the implicit damping re-solve of _euler)
    ("sink_nocontact", "sink", {"disableflags": 16}, 32, -3.0, 2),
    ("sink_nolimit_nofrictionloss", "sink", {"disableflags": 12}, 32, -3.0, 2),
    ("sink_nogravity_nowarmstart_noclamp", "sink", {"disableflags": 896}, 32, -3.0, 2),
    ("sink_nospring_nodamper_noactuation", "sink", {"disableflags": 2144}, 32, -3.0, 2),
    ("sink_noconstraint", "sink", {"disableflags": 1}, 32, -3.0, 2),
    ("sink_nosensor_norefsafe", "sink", {"disableflags": 12288}, 32, -3.0, 2),
    ("equality_bodies_noequality", "equality_bodies", {"disableflags": 2}, 32, 0.0, 2),
    ("ant_nofilterparent", "ant", {"disableflags": 1024}, 32, 0.0, 1),
    ("humanoid_eulerdamp", "humanoid", {"disableflags": 0}, 32, 0.4, 1),
    # step inputs beyond qpos / qvel / ctrl (EXTRA below): Cartesian and joint-space applied forces (support.xfrc_accumulate
    # :184-194, forward.py:222-228), equalities switched per env through eq_active (constraint.py:195-200), mocap poses driving
    # kinematics (smooth.py:105-113)
    ("humanoid_applied", "humanoid", {"iterations": 100, "ls_iterations": 50}, 32, 0.4, 2),   # converged: the 1-iteration XML solver is discontinuous at most of these states
    ("equality_bodies_toggled", "equality_bodies", {}, 32, 0.0, 2),
    ("equality_sites_toggled", "equality", {}, 32, 0.3, 2),
    ("pendula_mocap", "pendula", {}, 32, 0.0, 2),
    # solver / option combinations not recorded above: CG with the pyramidal cone, CG run to convergence on Humanoid, impratio != 1
    # in both cones (constraint.py:487, 549), the zoo Ant with its default Newton / pyramidal options, Walker2d with Euler
    ("halfcheetah_cg", "halfcheetah", {"solver": 1}, 32, 0.6, 2),
    ("humanoid_cg_converged", "humanoid", {"solver": 1, "iterations": 100, "ls_iterations": 50}, 32, 0.4, 1),
    ("sink_impratio", "sink", {"impratio": 5.0}, 32, -3.0, 2),
    ("sink_elliptic_impratio", "sink", {"cone": 1, "impratio": 0.5}, 32, -3.0, 2),
    ("ant_zoo_default", "ant_zoo", {}, 32, 0.5, 2),
    ("walker2d", "walker2d", {}, 32, 0.8, 2),
    # ctrl three times its range: clamping active in most envs (forward.py:117-126), and the same with DSBL_CLAMPCTRL;
    # sink at three times ctrl also drives its position actuator into its forcerange
    ("halfcheetah_bigctrl", "halfcheetah", {}, 32, 0.6, 2),
    ("halfcheetah_bigctrl_noclamp", "halfcheetah", {"disableflags": 256}, 32, 0.6, 2),
    ("sink_bigctrl", "sink", {}, 32, -3.0, 2),
    # vector options: tilted gravity (converged solver), wind + viscosity on top of the swimmer's density (inertia-box fluid model)
    ("humanoid_gravity_tilted", "humanoid", {"iterations": 100, "ls_iterations": 50, "gravity[0]": 3.0, "gravity[2]": -7.0}, 32, 0.4, 2),
    ("swimmer_wind_viscosity", "swimmer", {"wind[0]": 0.7, "wind[1]": -0.4, "viscosity": 0.05}, 32, 0.0, 2),
    # integrator x feature and cone x solver x model combinations that only appeared separately above
    ("muscle_rk4", "muscle", {"integrator": 1}, 32, -4.0, 2),
    ("sink_rk4", "sink", {"integrator": 1}, 32, -3.0, 1),
    ("satellite_small_rk4", "satellite_small", {"integrator": 1}, 32, 0.0, 2),
    # (a CG run that stops at `tolerance` is not usable as a pin: hopper + CG + elliptic lands 2e-8 from the recording in the
    # oracle itself -- one iteration more or less at the threshold; the CG recordings above use fixed_iterations or stay far from it)
    ("walker2d_elliptic", "walker2d", {"cone": 1, "iterations": 100, "ls_iterations": 50}, 32, 0.8, 2),
    ("humanoid_elliptic_converged", "humanoid", {"cone": 1, "iterations": 100, "ls_iterations": 50}, 32, 0.4, 1),
]


def extra_inputs(name, lite, B):
    """Seeded extra step inputs of the fixtures named *_applied / *_toggled / *_mocap (recorded as in_<field>)."""
    rng = np.random.RandomState(sum(map(ord, name)))
    out = {}
    if name.endswith("_applied") or name.endswith("_mocap"):
        x = 20.0 * rng.randn(B, lite.nbody, 6)
        x[:, 0] = 0.0
        x *= (rng.rand(B, lite.nbody, 1) < 0.35)      # most bodies carry exactly zero wrench (the skip path)
        out["xfrc_applied"] = x
        out["qfrc_applied"] = 2.0 * rng.randn(B, lite.nv)
    if name.endswith("_toggled"):
        out["eq_active"] = (rng.rand(B, lite.neq) < 0.6).astype(np.uint8)
    if name.endswith("_mocap"):
        ids = [b for b in range(lite.nbody) if int(lite.body_mocapid[b]) >= 0]
        pos = np.stack([np.asarray(lite.body_pos[b]) for b in ids])[None] + 0.2 * rng.randn(B, len(ids), 3)
        quat = rng.randn(B, len(ids), 4)
        out["mocap_pos"], out["mocap_quat"] = pos, quat / np.linalg.norm(quat, axis=-1, keepdims=True) * rng.uniform(0.5, 2.0, (B, len(ids), 1))
    return out


# every fixture records at least this many envs (the per-fixture counts in CASES are the historical minimum)
MIN_ENVS = int(os.environ.get("MJT_GOLDEN_ENVS", "32"))
OUT_DIR = os.environ.get("MJT_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))


def run_case(name, model, opt, B, drop, nsteps):
    B = max(B, MIN_ENVS)
    fixed = bool(opt.get("fixed_iterations", 0))
    lite = util.load_model(model, **{k: v for k, v in opt.items() if k != "fixed_iterations"})
    qpos, qvel, ctrl = util.seeded_state(lite, B, drop=drop)
    act = None
    if drop < -4.5:
        qpos, qvel, ctrl = util.terrain_state(lite, B)
    elif drop < -3.5:
        qpos, qvel, ctrl, act = util.muscle_state(lite, B)
    elif drop < -2.5:
        qpos, qvel, ctrl, act = util.sink_state(lite, B)
    elif drop < 0:  # pile-up state: every free body lowered and tilted so that convex pairs interpenetrate
        qpos, qvel, ctrl = util.pile_state(lite, B) if drop > -1.5 else util.stack_state(lite, B)
    if "_bigctrl" in name:
        ctrl = 3.0 * ctrl
    extra = extra_inputs(name, lite, B)
    mx = ref_device.device_put(mujoco.MjModel(lite))
    dx0 = ref_io.make_data(mx)
    outs = {k: [] for k in OUT_FIELDS}
    cons = {k: [] for k in CONTACT_FIELDS}
    for b in range(B):
        d = dx0.replace(qpos=torch.as_tensor(qpos[b]), qvel=torch.as_tensor(qvel[b]), ctrl=torch.as_tensor(ctrl[b]))
        if act is not None:
            d = d.replace(act=torch.as_tensor(act[b]))
        for k, v in extra.items():
            d = d.replace(**{k: torch.as_tensor(v[b]).to(getattr(dx0, k).dtype)})
        for _ in range(nsteps):
            d = ref_forward.step(mx, d, fixed_iterations=fixed)
        for k in OUT_FIELDS:
            outs[k].append(np.asarray(getattr(d, k)))
        for k in CONTACT_FIELDS:
            cons[k].append(np.asarray(getattr(d.contact, k)))
    save = {"in_qpos": qpos, "in_qvel": qvel, "in_ctrl": ctrl, **({"in_act": act} if act is not None else {}),
            **{"in_" + k: v for k, v in extra.items()}, "nsteps": np.int64(nsteps),
            "opt_keys": np.array(list(opt.keys()), dtype="U32"), "opt_vals": np.array(list(opt.values()), dtype=np.float64),
            "model": np.array(model), "torch_version": np.array(torch.__version__)}
    for k, v in outs.items():
        save["out_" + k] = np.stack(v)
    for k, v in cons.items():
        save["contact_" + k] = np.stack(v)
    path = os.path.join(OUT_DIR, name + ".npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **save)
    print(f"{name}: B={B} steps={nsteps} qpos[0,:3]={save['out_qpos'][0, :3]} -> {os.path.relpath(path, ROOT)}", flush=True)


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case[0] not in only:
            continue
        run_case(*case)
