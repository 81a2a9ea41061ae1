"""Host-side compiled-model container and the MuJoCo constants/enums the step path needs.

The reference reads these from the third-party ``mujoco`` package (absent here): see
SURVEY.md Appendix A.  Values follow MuJoCo's public headers (mjmodel.h, pinned 3.5.0 in
the reference's uv.lock).  ``MjModelLite`` is a duck-typed stand-in for ``mujoco.MjModel``:
attribute names and array shapes match MjModel, so ``device_put`` accepts either.
"""

from __future__ import annotations

import enum
import io
import json

import numpy as np

# --- numeric constants (mjmodel.h) -------------------------------------------------------
mjMINVAL = 1e-15
mjMAXVAL = 1e10
mjMINMU = 1e-5
mjMINIMP = 1e-4
mjMAXIMP = 0.9999
mjNREF = 2
mjNIMP = 5
mjNGAIN = 10
mjNBIAS = 10
mjNDYN = 10
mjNEQDATA = 11
mjPI = np.pi


class JointType(enum.IntEnum):
    FREE = 0
    BALL = 1
    SLIDE = 2
    HINGE = 3

    def dof_width(self) -> int:
        return (6, 3, 1, 1)[self.value]

    def qpos_width(self) -> int:
        return (7, 4, 1, 1)[self.value]


class GeomType(enum.IntEnum):
    PLANE = 0
    HFIELD = 1
    SPHERE = 2
    CAPSULE = 3
    ELLIPSOID = 4
    CYLINDER = 5
    BOX = 6
    MESH = 7


class IntegratorType(enum.IntEnum):
    EULER = 0
    RK4 = 1
    IMPLICIT = 2
    IMPLICITFAST = 3


class ConeType(enum.IntEnum):
    PYRAMIDAL = 0
    ELLIPTIC = 1


class JacobianType(enum.IntEnum):
    DENSE = 0
    SPARSE = 1
    AUTO = 2


class SolverType(enum.IntEnum):
    PGS = 0
    CG = 1
    NEWTON = 2


class EqType(enum.IntEnum):
    CONNECT = 0
    WELD = 1
    JOINT = 2
    TENDON = 3
    FLEX = 4
    DISTANCE = 5


class TrnType(enum.IntEnum):
    JOINT = 0
    JOINTINPARENT = 1
    SLIDERCRANK = 2
    TENDON = 3
    SITE = 4
    BODY = 5


class DynType(enum.IntEnum):
    NONE = 0
    INTEGRATOR = 1
    FILTER = 2
    FILTEREXACT = 3
    MUSCLE = 4
    USER = 5


class GainType(enum.IntEnum):
    FIXED = 0
    AFFINE = 1
    MUSCLE = 2
    USER = 3


class BiasType(enum.IntEnum):
    NONE = 0
    AFFINE = 1
    MUSCLE = 2
    USER = 3


class WrapType(enum.IntEnum):
    NONE = 0
    JOINT = 1
    PULLEY = 2
    SITE = 3
    SPHERE = 4
    CYLINDER = 5


class CamLightType(enum.IntEnum):
    FIXED = 0
    TRACK = 1
    TRACKCOM = 2
    TARGETBODY = 3
    TARGETBODYCOM = 4


class ObjType(enum.IntEnum):
    UNKNOWN = 0
    BODY = 1
    XBODY = 2
    JOINT = 3
    DOF = 4
    GEOM = 5
    SITE = 6
    CAMERA = 7
    LIGHT = 8
    TENDON = 18
    ACTUATOR = 19


class DisableBit(enum.IntFlag):
    """mjtDisableBit, MuJoCo >= 3.3.6 layout (separate SPRING / DAMPER bits)."""

    CONSTRAINT = 1 << 0
    EQUALITY = 1 << 1
    FRICTIONLOSS = 1 << 2
    LIMIT = 1 << 3
    CONTACT = 1 << 4
    SPRING = 1 << 5
    DAMPER = 1 << 6
    GRAVITY = 1 << 7
    CLAMPCTRL = 1 << 8
    WARMSTART = 1 << 9
    FILTERPARENT = 1 << 10
    ACTUATION = 1 << 11
    REFSAFE = 1 << 12
    SENSOR = 1 << 13
    MIDPHASE = 1 << 14
    EULERDAMP = 1 << 15
    AUTORESET = 1 << 16
    NATIVECCD = 1 << 17
    ISLAND = 1 << 18


class EnableBit(enum.IntFlag):
    OVERRIDE = 1 << 0
    ENERGY = 1 << 1
    FWDINV = 1 << 2
    INVDISCRETE = 1 << 3
    MULTICCD = 1 << 4
    SLEEP = 1 << 5


class SensorType(enum.IntEnum):
    """mjtSensor (MuJoCo 3.5 numbering)."""

    TOUCH = 0
    ACCELEROMETER = 1
    VELOCIMETER = 2
    GYRO = 3
    FORCE = 4
    TORQUE = 5
    MAGNETOMETER = 6
    RANGEFINDER = 7
    CAMPROJECTION = 8
    JOINTPOS = 9
    JOINTVEL = 10
    TENDONPOS = 11
    TENDONVEL = 12
    ACTUATORPOS = 13
    ACTUATORVEL = 14
    ACTUATORFRC = 15
    JOINTACTFRC = 16
    TENDONACTFRC = 17
    BALLQUAT = 18
    BALLANGVEL = 19
    JOINTLIMITPOS = 20
    JOINTLIMITVEL = 21
    JOINTLIMITFRC = 22
    TENDONLIMITPOS = 23
    TENDONLIMITVEL = 24
    TENDONLIMITFRC = 25
    FRAMEPOS = 26
    FRAMEQUAT = 27
    FRAMEXAXIS = 28
    FRAMEYAXIS = 29
    FRAMEZAXIS = 30
    FRAMELINVEL = 31
    FRAMEANGVEL = 32
    FRAMELINACC = 33
    FRAMEANGACC = 34
    SUBTREECOM = 35
    SUBTREELINVEL = 36
    SUBTREEANGMOM = 37
    INSIDESITE = 38
    GEOMDIST = 39
    GEOMNORMAL = 40
    GEOMFROMTO = 41
    CONTACT = 42
    E_POTENTIAL = 43
    E_KINETIC = 44
    CLOCK = 45
    TACTILE = 46
    PLUGIN = 47
    USER = 48


class DataType(enum.IntEnum):
    REAL = 0
    POSITIVE = 1
    AXIS = 2
    QUATERNION = 3


class Stage(enum.IntEnum):
    NONE = 0
    POS = 1
    VEL = 2
    ACC = 3


class _Struct:
    """Plain attribute bag with npz-friendly (de)serialisation."""

    _scalar_int: tuple[str, ...] = ()
    _scalar_float: tuple[str, ...] = ()

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def as_dict(self) -> dict:
        return dict(self.__dict__)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(sorted(self.__dict__))})"


class MjOptionLite(_Struct):
    """Stand-in for ``mujoco.MjOption`` (fields read by reference types.Option :503-557)."""


class MjStatisticLite(_Struct):
    """Stand-in for ``mujoco.MjStatistic``."""


class MjModelLite(_Struct):
    """Stand-in for ``mujoco.MjModel``: numpy arrays under MjModel attribute names."""

    def save_npz(self, path) -> None:
        flat = {}
        meta = {"scalars": {}, "opt_scalars": {}, "stat_scalars": {}, "names_hex": None, "name2id": {}}
        for k, v in self.__dict__.items():
            if k == "opt":
                for ok, ov in v.__dict__.items():
                    if isinstance(ov, np.ndarray):
                        flat["opt." + ok] = ov
                    else:
                        meta["opt_scalars"][ok] = ov
            elif k == "stat":
                for ok, ov in v.__dict__.items():
                    if isinstance(ov, np.ndarray):
                        flat["stat." + ok] = ov
                    else:
                        meta["stat_scalars"][ok] = ov
            elif k == "names":
                meta["names_hex"] = v.hex()
            elif k == "name2id":
                meta["name2id"] = {kk: vv for kk, vv in v.items()}
            elif isinstance(v, np.ndarray):
                flat[k] = v
            elif isinstance(v, (int, float, str, bool)):
                meta["scalars"][k] = v
            else:
                raise TypeError(f"cannot serialise field {k}: {type(v)}")
        flat["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(path, **flat)

    @classmethod
    def load_npz(cls, path) -> "MjModelLite":
        with np.load(path, allow_pickle=False) as z:
            meta = json.loads(bytes(z["__meta__"]).decode())
            m = cls()
            opt, stat = MjOptionLite(), MjStatisticLite()
            for k in z.files:
                if k == "__meta__":
                    continue
                if k.startswith("opt."):
                    setattr(opt, k[4:], z[k])
                elif k.startswith("stat."):
                    setattr(stat, k[5:], z[k])
                else:
                    setattr(m, k, z[k])
            opt.__dict__.update(meta["opt_scalars"])
            stat.__dict__.update(meta["stat_scalars"])
            m.__dict__.update(meta["scalars"])
            m.opt, m.stat = opt, stat
            if meta.get("names_hex") is not None:
                m.names = bytes.fromhex(meta["names_hex"])
            m.name2id = meta.get("name2id", {})
        return m

    def to_bytes(self) -> bytes:
        buf = io.BytesIO()
        self.save_npz(buf)
        return buf.getvalue()
