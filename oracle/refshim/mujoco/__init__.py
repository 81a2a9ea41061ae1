"""mujoco stand-in (TEST INFRASTRUCTURE): the constants / enums the reference reads at import time (values from
MuJoCo's public headers, SURVEY.md Appendix A -- kept identical to mujoco_torch_b200.mjmodel) and an ``MjModel``
facade over ``MjModelLite`` so that the reference's ``device_put`` can ingest our compiled models."""
import enum
import types as _types

import numpy as np

__version__ = "3.5.0"

mjMINVAL, mjMAXVAL, mjMINMU, mjMINIMP, mjMAXIMP = 1e-15, 1e10, 1e-5, 1e-4, 0.9999
mjNREF, mjNIMP, mjNGROUP, mjNEQDATA, mjNDYN, mjNGAIN, mjNBIAS, mjPI = 2, 5, 6, 11, 10, 10, 10, np.pi


def _enum(name, prefix, members, start=0):
    if isinstance(members, str):
        members = {m: i + start for i, m in enumerate(members.split())}
    return enum.IntEnum(name, {prefix + k: v for k, v in members.items()})


mjtJoint = _enum("mjtJoint", "mjJNT_", "FREE BALL SLIDE HINGE")
mjtGeom = _enum("mjtGeom", "mjGEOM_", "PLANE HFIELD SPHERE CAPSULE ELLIPSOID CYLINDER BOX MESH SDF")
mjtIntegrator = _enum("mjtIntegrator", "mjINT_", "EULER RK4 IMPLICIT IMPLICITFAST")
mjtCone = _enum("mjtCone", "mjCONE_", "PYRAMIDAL ELLIPTIC")
mjtJacobian = _enum("mjtJacobian", "mjJAC_", "DENSE SPARSE AUTO")
mjtSolver = _enum("mjtSolver", "mjSOL_", "PGS CG NEWTON")
mjtEq = _enum("mjtEq", "mjEQ_", "CONNECT WELD JOINT TENDON FLEX DISTANCE")
mjtWrap = _enum("mjtWrap", "mjWRAP_", "NONE JOINT PULLEY SITE SPHERE CYLINDER")
mjtTrn = _enum("mjtTrn", "mjTRN_", "JOINT JOINTINPARENT SLIDERCRANK TENDON SITE BODY")
mjtDyn = _enum("mjtDyn", "mjDYN_", "NONE INTEGRATOR FILTER FILTEREXACT MUSCLE USER")
mjtGain = _enum("mjtGain", "mjGAIN_", "FIXED AFFINE MUSCLE USER")
mjtBias = _enum("mjtBias", "mjBIAS_", "NONE AFFINE MUSCLE USER")
mjtConstraint = _enum("mjtConstraint", "mjCNSTR_", "EQUALITY FRICTION_DOF FRICTION_TENDON LIMIT_JOINT LIMIT_TENDON CONTACT_FRICTIONLESS "
                      "CONTACT_PYRAMIDAL CONTACT_ELLIPTIC")
mjtCamLight = _enum("mjtCamLight", "mjCAMLIGHT_", "FIXED TRACK TRACKCOM TARGETBODY TARGETBODYCOM")
mjtObj = _enum("mjtObj", "mjOBJ_", "UNKNOWN BODY XBODY JOINT DOF GEOM SITE CAMERA LIGHT")
mjtDataType = _enum("mjtDataType", "mjDATATYPE_", "REAL POSITIVE AXIS QUATERNION")
mjtStage = _enum("mjtStage", "mjSTAGE_", "NONE POS VEL ACC")
mjtSensor = _enum("mjtSensor", "mjSENS_", "TOUCH ACCELEROMETER VELOCIMETER GYRO FORCE TORQUE MAGNETOMETER RANGEFINDER CAMPROJECTION "
                  "JOINTPOS JOINTVEL TENDONPOS TENDONVEL ACTUATORPOS ACTUATORVEL ACTUATORFRC JOINTACTFRC TENDONACTFRC BALLQUAT "
                  "BALLANGVEL JOINTLIMITPOS JOINTLIMITVEL JOINTLIMITFRC TENDONLIMITPOS TENDONLIMITVEL TENDONLIMITFRC FRAMEPOS "
                  "FRAMEQUAT FRAMEXAXIS FRAMEYAXIS FRAMEZAXIS FRAMELINVEL FRAMEANGVEL FRAMELINACC FRAMEANGACC SUBTREECOM "
                  "SUBTREELINVEL SUBTREEANGMOM INSIDESITE GEOMDIST GEOMNORMAL GEOMFROMTO CONTACT E_POTENTIAL E_KINETIC CLOCK "
                  "TACTILE PLUGIN USER")
_dsbl = "CONSTRAINT EQUALITY FRICTIONLOSS LIMIT CONTACT SPRING DAMPER GRAVITY CLAMPCTRL WARMSTART FILTERPARENT ACTUATION REFSAFE " \
        "SENSOR MIDPHASE EULERDAMP AUTORESET NATIVECCD ISLAND"
_d = {"mjDSBL_" + n: 1 << i for i, n in enumerate(_dsbl.split())}
_d["mjNDISABLE"] = len(_dsbl.split())
mjtDisableBit = enum.IntEnum("mjtDisableBit", _d)
_e = {"mjENBL_" + n: 1 << i for i, n in enumerate("OVERRIDE ENERGY FWDINV INVDISCRETE MULTICCD SLEEP".split())}
_e["mjNENABLE"] = 6
mjtEnableBit = enum.IntEnum("mjtEnableBit", _e)

_structs = _types.SimpleNamespace(_MjContactList=type("_MjContactList", (), {}))


class MjOption:
    def __init__(self, lite_opt):
        object.__setattr__(self, "_o", lite_opt)

    cone = property(lambda self: self._o.cone)
    integrator = property(lambda self: self._o.integrator)
    solver = property(lambda self: self._o.solver)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_o"), name)

    def __setattr__(self, name, value):
        setattr(self._o, name, value)


class MjStatistic:
    def __init__(self, lite_stat):
        self._s = lite_stat

    def __getattr__(self, name):
        return getattr(self._s, name)


def _zeros(shape, dtype=np.float64):
    return lambda m: np.zeros(shape(m) if callable(shape) else shape, dtype=dtype)


# MjModel fields the reference copies but our mini-compiler does not produce (cosmetic / unused on the step path)
_DEFAULTS = {
    "light_castshadow": lambda m: np.zeros(m.nlight, np.uint8), "light_cutoff": lambda m: np.full(m.nlight, 45.0),
    "light_diffuse": lambda m: np.full((m.nlight, 3), 0.7), "light_ambient": lambda m: np.zeros((m.nlight, 3)),
    "light_specular": lambda m: np.full((m.nlight, 3), 0.3), "light_attenuation": lambda m: np.tile([1.0, 0, 0], (m.nlight, 1)),
    "light_active": lambda m: np.ones(m.nlight, np.uint8), "light_type": lambda m: np.asarray(m.light_directional, np.int32),
    "mesh_normaladr": lambda m: np.zeros(m.nmesh, np.int32), "mesh_normalnum": lambda m: np.zeros(m.nmesh, np.int32),
    "mesh_graphadr": lambda m: np.full(m.nmesh, -1, np.int32), "mesh_normal": lambda m: np.zeros((0, 3)),
    "mesh_graph": lambda m: np.zeros(0, np.int32), "mesh_texcoordadr": lambda m: np.full(m.nmesh, -1, np.int32),
    "mesh_texcoordnum": lambda m: np.zeros(m.nmesh, np.int32), "mesh_texcoord": lambda m: np.zeros((0, 2)),
    "hfield_size": lambda m: np.zeros((0, 4)), "hfield_nrow": lambda m: np.zeros(0, np.int32), "hfield_ncol": lambda m: np.zeros(0, np.int32),
    "hfield_adr": lambda m: np.zeros(0, np.int32), "hfield_data": lambda m: np.zeros(0),
    "mat_rgba": lambda m: np.zeros((0, 4)), "mat_texid": lambda m: np.zeros((0, 10), np.int32), "mat_texrepeat": lambda m: np.zeros((0, 2)),
    "mat_texuniform": lambda m: np.zeros(0, np.uint8), "tex_type": lambda m: np.zeros(0, np.int32), "tex_height": lambda m: np.zeros(0, np.int32),
    "tex_width": lambda m: np.zeros(0, np.int32), "tex_nchannel": lambda m: np.zeros(0, np.int32), "tex_adr": lambda m: np.zeros(0, np.int32),
    "tex_data": lambda m: np.zeros(0, np.uint8), "cam_resolution": lambda m: np.ones((m.ncam, 2), np.int32),
    "cam_sensorsize": lambda m: np.zeros((m.ncam, 2)), "cam_intrinsic": lambda m: np.zeros((m.ncam, 4)),
    "key_time": lambda m: np.zeros(0), "key_qpos": lambda m: np.zeros((0, m.nq)), "key_qvel": lambda m: np.zeros((0, m.nv)),
    "key_act": lambda m: np.zeros((0, m.na)), "key_mpos": lambda m: np.zeros((0, 3 * m.nmocap)), "key_mquat": lambda m: np.zeros((0, 4 * m.nmocap)),
    "key_ctrl": lambda m: np.zeros((0, m.nu)),
}


class MjModel:
    """Facade exposing an ``MjModelLite`` under the class the reference type-dispatches on."""

    def __init__(self, lite):
        object.__setattr__(self, "_lite", lite)
        object.__setattr__(self, "_opt", MjOption(lite.opt))
        object.__setattr__(self, "_stat", MjStatistic(lite.stat))

    opt = property(lambda self: self._opt)
    stat = property(lambda self: self._stat)
    actuator_biastype = property(lambda self: self._lite.actuator_biastype)
    actuator_dyntype = property(lambda self: self._lite.actuator_dyntype)
    actuator_gaintype = property(lambda self: self._lite.actuator_gaintype)
    actuator_trntype = property(lambda self: self._lite.actuator_trntype)
    eq_type = property(lambda self: self._lite.eq_type)

    def __getattr__(self, name):
        lite = object.__getattribute__(self, "_lite")
        if hasattr(lite, name):
            return getattr(lite, name)
        if name in _DEFAULTS:
            return _DEFAULTS[name](lite)
        raise AttributeError(name)


class MjData:
    """Only what reference device._probe_kinematics_static touches."""

    def __init__(self, model):
        self.model = model
        self.qpos = np.array(model.qpos0, dtype=np.float64)
        self.xaxis = np.zeros((model.njnt, 3))
        self.xanchor = np.zeros((model.njnt, 3))


def mj_normalizeQuat(model, qpos):
    pass


def mj_kinematics(model, data):
    # make every probed field look qpos-dependent: nothing is reported as "static" (compile-mode-only optimisation)
    data.xaxis = np.full((model.njnt, 3), float(np.sum(data.qpos)))
    data.xanchor = np.full((model.njnt, 3), float(np.sum(data.qpos)) + 1.0)


def mj_forward(model, data):
    raise NotImplementedError("MuJoCo C is not available")
